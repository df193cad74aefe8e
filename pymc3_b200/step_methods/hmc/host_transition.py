"""One NUTS / HamiltonianMC transition driven from the host, for potentials that are USER Python objects.

The reference lets users pass their own `QuadPotential` subclass as `potential=` and calls its `velocity`,
`energy`, `random` and `update` from the integrator (pymc3/step_methods/hmc/quadpotential.py:91-132,
integration.py:39-109; contract test pymc3/tests/test_quadpotential.py:138-155).  Python callbacks cannot run
inside the device state machine, so for such potentials -- and only for them; the built-in diagonal potentials
never come here -- the tree logic of one chain runs on the host and calls the user's object, while the density
and its gradient are still evaluated on the device (`b2_logp_dlogp` through `ValueGradFunction`; without the CUDA
library this path fails like every other).  The tree builder is the same iterative, binary-counter formulation
as the device state machine (csrc/b2_core.cuh: after leaf n, ctz(~n) merges), semantics of nuts.py:254-406.
Randomness comes from NumPy's global stream, as in the reference (nuts.py:30-33, 177).
"""
import collections

import numpy as np

Edge = collections.namedtuple("Edge", "q p v grad energy logp")
Sub = collections.namedtuple("Sub", "first last p_sum q grad energy logp log_size log_accept n_leaf")


def _logaddexp(a, b):
    return float(np.logaddexp(a, b))


class HostIntegrator:
    """integration.py:39-109 with a user potential; `f(q) -> (logp, grad)` runs on the device."""

    def __init__(self, potential, f):
        self.pot, self.f = potential, f

    def start(self, q, p):
        logp, grad = self.f(q)
        v = self.pot.velocity(p)
        return Edge(q, p, v, grad, float(self.pot.energy(p, velocity=v)) - float(logp), float(logp))

    def step(self, eps, s):
        p_mid = s.p + 0.5 * eps * s.grad
        q_new = s.q + eps * self.pot.velocity(p_mid)
        logp, grad = self.f(q_new)
        p_new = p_mid + 0.5 * eps * grad
        v_new = self.pot.velocity(p_new)
        return Edge(q_new, p_new, v_new, grad, float(self.pot.energy(p_new, velocity=v_new)) - float(logp), float(logp))


def _turning(p_sum, a, b):
    return bool(np.dot(p_sum, a.v) <= 0 or np.dot(p_sum, b.v) <= 0)


def _log_uniform():
    return np.log(np.random.uniform())


def nuts_transition(integ, start, step_size, max_depth, emax):
    """-> (q, grad, stats dict).  Main-tree state: left / right edge, p_sum, proposal, weights."""
    e0 = start.energy
    left = right = start
    p_sum = start.p.copy()
    prop = (start.q, start.grad, start.energy, start.logp)
    log_size, log_accept, n_prop, depth, max_de = 0.0, -np.inf, 0, 0, 0.0
    diverged = turned = False
    while depth < max_depth and not (diverged or turned):
        forward = _log_uniform() < np.log(0.5)                       # nuts.py:177
        eps = step_size if forward else -step_size
        edge = right if forward else left
        stack = []                                                   # completed sub-trees, smallest on top
        ok = True
        n_leaves = 1 << depth
        leaves_done = 0
        for n in range(n_leaves):
            edge = integ.step(eps, edge)
            leaves_done = n + 1
            de = edge.energy - e0
            if np.isnan(de):
                de = np.inf
            if abs(de) > abs(max_de):
                max_de = de
            if not abs(de) < emax:                                   # nuts.py:338-345
                diverged, ok = True, False
                break
            cur = Sub(edge, edge, edge.p.copy(), edge.q, edge.grad, edge.energy, edge.logp, -de, -de + min(0.0, -de), 1)
            m = n
            while m & 1:                                             # one merge per trailing one-bit of n
                t1 = stack.pop()
                t2 = cur
                ps = t1.p_sum + t2.p_sum
                turn = _turning(ps, t1.first, t2.last)
                if t1.n_leaf > 1 and not turn:                       # nuts.py:364-370
                    turn = _turning(t1.p_sum + t2.first.p, t1.first, t2.first) or \
                        _turning(t1.last.p + t2.p_sum, t1.last, t2.last)
                if turn:
                    turned, ok = True, False
                    break
                ls = _logaddexp(t1.log_size, t2.log_size)
                take2 = _log_uniform() < t2.log_size - ls            # nuts.py:375-378
                src = t2 if take2 else t1
                cur = Sub(t1.first, t2.last, ps, src.q, src.grad, src.energy, src.logp, ls,
                          _logaddexp(t1.log_accept, t2.log_accept), t1.n_leaf + t2.n_leaf)
                m >>= 1
            if not ok:
                break
            stack.append(cur)
        depth += 1
        n_prop += leaves_done
        if not ok:
            break
        sub = stack.pop()                                            # the whole new sub-tree: top-level merge, :283-309
        if _log_uniform() < sub.log_size - log_size:
            prop = (sub.q, sub.grad, sub.energy, sub.logp)
        log_size = _logaddexp(log_size, sub.log_size)
        log_accept = _logaddexp(log_accept, sub.log_accept)
        # the two halves in time order: t1 | t2 (a backward sub-tree is traversed from its far end)
        if forward:
            first1, last1, sum1, first2, last2, sum2 = left, right, p_sum, sub.first, sub.last, sub.p_sum
            right = sub.last
        else:
            first1, last1, sum1, first2, last2, sum2 = sub.last, sub.first, sub.p_sum, left, right, p_sum
            left = sub.last
        p_sum = sum1 + sum2
        turned = _turning(p_sum, first1, last2) or _turning(sum1 + first2.p, first1, first2) or \
            _turning(last1.p + sum2, last1, last2)                    # :298-307
    accept = 0.0
    if log_size > 0:                                                 # nuts.py:391-397: exp(log_accept) / expm1(log_size), in log space
        log_den = log_size + np.log1p(-np.exp(-log_size)) if log_size > 30 else np.log(np.expm1(log_size))
        accept = float(np.exp(log_accept - log_den))
    q, grad, energy, logp = prop
    stats = {"depth": depth, "mean_tree_accept": accept, "energy_error": energy - e0, "energy": energy,
             "tree_size": float(n_prop), "max_energy_error": max_de, "model_logp": logp, "diverging": bool(diverged)}
    return q, grad, stats


def hmc_transition(integ, start, step_size, path_length, max_steps, emax):
    """hmc.py:110-152."""
    n_steps = max(1, int(path_length / step_size))
    n_steps = min(max_steps, n_steps)
    state = start
    for _ in range(n_steps):
        state = integ.step(step_size, state)
    div = not np.isfinite(state.energy)
    de = start.energy - state.energy
    if np.isnan(de):
        de = -np.inf
    if abs(de) > emax:
        div = True
    accept = min(1.0, float(np.exp(de)))
    accepted = (not div) and not (np.random.rand() >= accept)
    end = state if accepted else start
    stats = {"path_length": path_length, "n_steps": n_steps, "accept": accept, "accepted": accepted,
             "energy_error": de, "energy": state.energy, "model_logp": state.logp, "diverging": bool(div)}
    return end.q, end.grad, stats
