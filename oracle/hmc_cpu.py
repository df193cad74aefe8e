"""NUTS / HamiltonianMC on the CPU, restating the reference.  TEST INFRASTRUCTURE.

Follows, operation by operation:
  pymc3/step_methods/hmc/integration.py:39-47, 81-109   (leapfrog, energy)
  pymc3/step_methods/step_sizes.py:21-58                (dual averaging)
  pymc3/step_methods/hmc/nuts.py:168-188, 220-406       (tree doubling, recursive)
  pymc3/step_methods/hmc/hmc.py:110-152                 (fixed-length HMC + MH)
  pymc3/step_methods/hmc/base_hmc.py:133-199            (one transition + adaptation)
  pymc3/sampling.py:914-936                             (draw loop, tune switch)

The tree builder is deliberately *recursive* like the reference; the CUDA engine is
iterative.  Agreement between the two is therefore a test of the
iterative restructuring as well as of the arithmetic.
"""
from collections import namedtuple

import numpy as np

PhasePoint = namedtuple("PhasePoint", "q p v grad energy logp")
Candidate = namedtuple("Candidate", "q grad energy log_w_accept logp")
Branch = namedtuple("Branch", "first last p_sum pick log_size log_accept n_leaf")


class BadInitialEnergy(RuntimeError):
    """base_hmc.py:138-158 raises SamplingError('Bad initial energy')."""


class StepSizeAdapter:
    """step_sizes.py:21-58."""

    def __init__(self, step0, target, gamma=0.05, k=0.75, t0=10):
        self.log_step = np.log(step0)
        self.log_bar = self.log_step
        self.target = target
        self.hbar = 0.0
        self.k, self.t0, self.gamma = k, t0, gamma
        self.count = 1
        self.mu = np.log(10 * step0)

    def current(self, tune):
        return np.exp(self.log_step) if tune else np.exp(self.log_bar)   # :34-38

    def update(self, accept, tune):
        if not tune:                                                      # :41-43
            return
        w = 1.0 / (self.count + self.t0)
        self.hbar = (1 - w) * self.hbar + w * (self.target - accept)
        self.log_step = self.mu - self.hbar * np.sqrt(self.count) / self.gamma
        mk = self.count ** -self.k
        self.log_bar = mk * self.log_step + (1 - mk) * self.log_bar
        self.count += 1

    def stats(self):
        return {"step_size": np.exp(self.log_step), "step_size_bar": np.exp(self.log_bar)}


class Leapfrog:
    """integration.py:28-109."""

    def __init__(self, potential, logp_dlogp):
        self.pot = potential
        self.f = logp_dlogp
        self.n_grad = 0

    def start(self, q, p):
        logp, grad = self.f(q)                                   # :43
        self.n_grad += 1
        v = self.pot.velocity(p)
        energy = self.pot.kinetic(p, v) - logp                   # :45-46
        return PhasePoint(q, p, v, grad, energy, logp)

    def step(self, eps, s):
        half = 0.5 * eps                                         # :90
        p_mid = s.p + half * s.grad                              # :94
        v_mid = self.pot.velocity(p_mid)                         # :96
        q_new = s.q + eps * v_mid                                # :99
        logp, grad = self.f(q_new)                               # :101
        self.n_grad += 1
        p_new = p_mid + half * grad                              # :104
        v_new = self.pot.velocity(p_new)                         # :106
        energy = self.pot.kinetic(p_new, v_new) - logp           # :106-107
        return PhasePoint(q_new, p_new, v_new, grad, energy, logp)


def _log_lt(u, log_p):
    """nuts.py:30-33 (logbern): log(U) < log_p; NaN is an error there."""
    if np.isnan(log_p):
        raise FloatingPointError("log_p can't be nan.")
    with np.errstate(divide="ignore"):
        return np.log(u) < log_p


def _turning(p_sum, a, b):
    return (p_sum.dot(a.v) <= 0) or (p_sum.dot(b.v) <= 0)


class _Trajectory:
    """nuts.py:220-406 (_Tree)."""

    def __init__(self, integ, start, eps, emax, rng, t):
        self.integ, self.start, self.eps, self.emax = integ, start, eps, emax
        self.rng, self.t = rng, t
        self.e0 = start.energy
        self.left = self.right = start
        self.pick = Candidate(start.q, start.grad, start.energy, 1.0, start.logp)   # :244-245
        self.depth = 0
        self.log_size = 0.0
        self.log_accept = -np.inf
        self.n_leaf = 0
        self.p_sum = start.p.copy()
        self.max_de = 0.0
        self._leaf_in_branch = 0

    # -- one doubling: nuts.py:254-309
    def grow(self, direction):
        self._leaf_in_branch = 0
        if direction > 0:
            br, div, turn = self._branch(self.right, self.depth, self.eps)
            lm_begin, lm_end = self.left, self.right
            rm_begin, rm_end = br.first, br.last
            lm_psum, rm_psum = self.p_sum, br.p_sum
            self.right = br.last
        else:
            br, div, turn = self._branch(self.left, self.depth, -self.eps)
            lm_begin, lm_end = br.last, br.first
            rm_begin, rm_end = self.left, self.right
            lm_psum, rm_psum = br.p_sum, self.p_sum
            self.left = br.last
        old_depth = self.depth
        self.depth += 1
        self.n_leaf += br.n_leaf
        if div or turn:                                          # :286-287
            return div, turn
        if _log_lt(self.rng.top_u(self.t, old_depth), br.log_size - self.log_size):   # :289-291
            self.pick = br.pick
        self.log_size = np.logaddexp(self.log_size, br.log_size)
        self.log_accept = np.logaddexp(self.log_accept, br.log_accept)
        self.p_sum = self.p_sum + br.p_sum
        # :298-307 (self.depth > 0 always holds here)
        turn = _turning(self.p_sum, self.left, self.right)
        ps1 = lm_psum + rm_begin.p
        turn1 = (ps1.dot(lm_begin.v) <= 0) or (ps1.dot(rm_begin.v) <= 0)
        ps2 = lm_end.p + rm_psum
        turn2 = (ps2.dot(lm_end.v) <= 0) or (ps2.dot(rm_end.v) <= 0)
        return div, bool(turn | turn1 | turn2)

    # -- a single leaf: nuts.py:311-345
    def _leaf(self, frm, eps):
        nxt = self.integ.step(eps, frm)
        leaf_idx = self._leaf_in_branch
        self._leaf_in_branch += 1
        de = nxt.energy - self.e0
        if np.isnan(de):
            de = np.inf
        if abs(de) > abs(self.max_de):
            self.max_de = de
        if abs(de) < self.emax:
            log_w_accept = -de + min(0.0, -de)                   # :331 (sic)
            cand = Candidate(nxt.q, nxt.grad, nxt.energy, log_w_accept, nxt.logp)
            return Branch(nxt, nxt, nxt.p, cand, -de, log_w_accept, 1), False, False
        return Branch(None, None, None, None, -np.inf, -np.inf, 1), True, False

    # -- recursive subtree: nuts.py:347-389
    def _branch(self, frm, depth, eps):
        if depth == 0:
            return self._leaf(frm, eps)
        b1, div, turn = self._branch(frm, depth - 1, eps)
        if div or turn:
            return b1, div, turn
        b2, div, turn = self._branch(b1.last, depth - 1, eps)
        first, last = b1.first, b2.last
        if not (div or turn):
            p_sum = b1.p_sum + b2.p_sum
            turn = _turning(p_sum, first, last)
            if depth - 1 > 0:                                    # :365-370
                ps1 = b1.p_sum + b2.first.p
                t1 = (ps1.dot(b1.first.v) <= 0) or (ps1.dot(b2.first.v) <= 0)
                ps2 = b1.last.p + b2.p_sum
                t2 = (ps2.dot(b1.last.v) <= 0) or (ps2.dot(b2.last.v) <= 0)
                turn = bool(turn | t1 | t2)
            log_size = np.logaddexp(b1.log_size, b2.log_size)
            log_accept = np.logaddexp(b1.log_accept, b2.log_accept)
            u = self.rng.merge_u(self.t, self.depth, depth, self._leaf_in_branch - 1)
            pick = b2.pick if _log_lt(u, b2.log_size - log_size) else b1.pick     # :375-378
        else:
            p_sum, log_size, log_accept, pick = b1.p_sum, b1.log_size, b1.log_accept, b1.pick
        return Branch(first, last, p_sum, pick, log_size, log_accept,
                      b1.n_leaf + b2.n_leaf), div, turn

    def summary(self):                                           # :391-406
        mean_accept = 0.0
        if self.log_size > 0:
            with np.errstate(over="ignore", invalid="ignore"):
                mean_accept = np.exp(self.log_accept) / np.expm1(self.log_size)
            if not np.isfinite(mean_accept):
                # NOT in the reference: there exp()/expm1() overflows to inf/inf = NaN once the energy
                # drops by more than ~709 (< Emax) and dual averaging is NaN from then on.  The engine
                # evaluates the same ratio in log space in exactly that case (deviations).
                mean_accept = np.exp(self.log_accept - (self.log_size + np.log1p(-np.exp(-self.log_size))))
        return {
            "depth": self.depth,
            "mean_tree_accept": mean_accept,
            "energy_error": self.pick.energy - self.start.energy,
            "energy": self.pick.energy,
            "tree_size": self.n_leaf,
            "max_energy_error": self.max_de,
            "model_logp": self.pick.logp,
        }


class CpuHMCBase:
    """base_hmc.py:36-199 for one chain."""

    default_target = 0.8

    def __init__(self, logp_dlogp, ndim, potential, rng, step_scale=0.25, emax=1000.0,
                 target_accept=None, gamma=0.05, k=0.75, t0=10, adapt_step_size=True):
        self.f, self.ndim, self.pot, self.rng = logp_dlogp, ndim, potential, rng
        self.emax = emax
        self.adapt_step_size = adapt_step_size
        self.step_size = step_scale / ndim ** 0.25                # base_hmc.py:93
        target = self.default_target if target_accept is None else target_accept
        self.adapter = StepSizeAdapter(self.step_size, target, gamma, k, t0)
        self.integ = Leapfrog(potential, logp_dlogp)
        self.tune = True
        self.iter_count = 0
        self.n_diverging_after_tune = 0

    def _jitter(self, step, t):
        return step

    def transition(self, q0):
        t = self.iter_count
        p0 = self.pot.draw(self.rng.momentum(t, self.ndim))      # base_hmc.py:135
        start = self.integ.start(q0, p0)
        if not np.isfinite(start.energy):
            self.pot.check()
            raise BadInitialEnergy("Bad initial energy")
        adapt = self.tune and self.adapt_step_size
        step = self.adapter.current(adapt)                        # :160-162
        self.step_size = step
        step = self._jitter(step, t)                              # :164-165
        end_q, end_grad, accept, diverged, stats = self._trajectory(start, step, t)
        self.adapter.update(accept, adapt)                        # :169
        self.pot.update(end_q, self.tune)                         # :170
        if diverged and not self.tune:
            self.n_diverging_after_tune += 1
        self.iter_count += 1
        out = {"tune": self.tune, "diverging": bool(diverged)}
        out.update(stats)
        out.update(self.adapter.stats())                          # :194-197
        return end_q, out


class CpuNUTS(CpuHMCBase):
    """nuts.py:36-208."""

    default_target = 0.8

    def __init__(self, *a, max_treedepth=10, early_max_treedepth=8, **kw):
        super().__init__(*a, **kw)
        self.max_treedepth = max_treedepth
        self.early_max_treedepth = early_max_treedepth
        self.n_max_depth_after_tune = 0

    def _trajectory(self, start, step, t):
        limit = self.early_max_treedepth if (self.tune and self.iter_count < 200) \
            else self.max_treedepth                               # nuts.py:169-172
        traj = _Trajectory(self.integ, start, step, self.emax, self.rng, t)
        div = turn = False
        for d in range(limit):
            go_right = _log_lt(self.rng.direction_u(t, d), np.log(0.5))   # :177
            div, turn = traj.grow(1 if go_right else -1)
            if div or turn:
                break
        else:
            if not self.tune:
                self.n_max_depth_after_tune += 1
        st = traj.summary()
        return traj.pick.q, traj.pick.grad, st["mean_tree_accept"], div, st


class CpuHMC(CpuHMCBase):
    """hmc.py:30-152."""

    default_target = 0.65

    def __init__(self, *a, path_length=2.0, max_steps=1024, jitter=True, **kw):
        super().__init__(*a, **kw)
        self.path_length, self.max_steps, self.jitter = path_length, max_steps, jitter

    def _jitter(self, step, t):
        return self.rng.hmc_jitter(t) * step if self.jitter else step     # hmc.py:26-27

    def _trajectory(self, start, step, t):
        n_steps = min(self.max_steps, max(1, int(self.path_length / step)))   # :111-112
        state = start
        for _ in range(n_steps):
            state = self.integ.step(step, state)
        diverged = not np.isfinite(state.energy)                  # :123-125
        de = start.energy - state.energy
        if np.isnan(de):
            de = -np.inf
        if abs(de) > self.emax:                                   # :129
            diverged = True
        with np.errstate(over="ignore"):
            accept = min(1.0, np.exp(de))
        # hmc.py:136: `div_info is not None or rand() >= accept` short-circuits: the
        # uniform is NOT consumed for a divergent trajectory (matters in legacy mode).
        if diverged:
            accepted = False
        else:
            accepted = not (self.rng.hmc_accept_u(t) >= accept)
        end = state if accepted else start
        stats = {"path_length": self.path_length, "n_steps": n_steps, "accept": accept,
                 "energy_error": de, "energy": state.energy, "accepted": accepted,
                 "model_logp": state.logp}
        return end.q, end.grad, accept, diverged, stats


def run_chain(sampler, q0, draws, tune, seed=None):
    """sampling.py:883-884, 914-936: seed once, `draws` includes tuning; stop tuning at i == tune.

    Returns (positions [draws, D], dict of stat arrays).
    """
    if seed is not None:
        sampler.rng.seed(seed)
    sampler.tune = bool(tune)
    sampler.iter_count = 0
    q = np.array(q0, dtype="d")
    qs = np.empty((draws, len(q)))
    stats = []
    for i in range(draws):
        if i == tune:
            sampler.tune = False
        q, st = sampler.transition(q)
        qs[i] = q
        stats.append(st)
    keys = stats[0].keys() if stats else []
    return qs, {k: np.array([s[k] for s in stats]) for k in keys}
