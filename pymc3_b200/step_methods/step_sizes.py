"""Host-side view of the dual-averaging state that lives on the device.

The recursion of pymc3/step_methods/step_sizes.py:21-58 runs inside the CUDA state machine
(csrc/b2_core.cuh, b2_end_transition).  This object keeps the reference's attribute surface
(`current`, `stats`, `warnings`) for one chain, fed from the device trace / chain report.
"""
import numpy as np
from scipy import special

from ..backends.report import SamplerWarning, WarningType


def acceptance_in_interval(mean_accept, n_draws, target):
    """step_sizes.py:66-70 for many chains at once: is `target` inside the central 95 % interval of
    Beta(n_good + 1, n_bad + 1), the acceptance rate counted over (at most) 100 draws?  NaN / no draws -> True."""
    mean_accept = np.atleast_1d(np.asarray(mean_accept, dtype="f8"))
    ok = np.ones(mean_accept.shape, dtype=bool)
    if not n_draws:
        return ok
    n_bound = min(100, int(n_draws))
    fin = np.isfinite(mean_accept)
    a, b = mean_accept[fin] * n_bound + 1, (1 - mean_accept[fin]) * n_bound + 1
    lower, upper = special.betaincinv(a, b, 0.025), special.betaincinv(a, b, 0.975)
    ok[fin] = (target >= lower) & (target <= upper)
    return ok


def acceptance_warnings(mean_accept, n_draws, target, ok=None):
    """step_sizes.py:60-79: the warning when the target is outside that interval"""
    if not n_draws or not np.isfinite(mean_accept):
        return []
    if ok is None:
        ok = bool(acceptance_in_interval(mean_accept, n_draws, target)[0])
    if not ok:
        msg = ("The acceptance probability does not match the target. It is %s, but should be close "
               "to %s. Try to increase the number of tuning steps." % (mean_accept, target))
        info = {"target": target, "actual": mean_accept}
        return [SamplerWarning(WarningType.BAD_ACCEPTANCE, msg, "warn", None, None, info)]
    return []


class DualAverageAdaptation:
    def __init__(self, initial_step, target, gamma, k, t0):
        self._initial_step = initial_step
        self._target = target
        self._gamma, self._k, self._t0 = gamma, k, t0
        self._log_step = np.log(initial_step)
        self._log_bar = self._log_step
        self._hbar, self._count, self._mu = 0.0, 1, np.log(10 * initial_step)
        self._tuned_stats = []

    def current(self, tune):
        return np.exp(self._log_step) if tune else np.exp(self._log_bar)

    def update(self, accept_stat, tune):
        """step_sizes.py:40-52 on the host -- used only by host-driven transitions (user potentials); the batched
        path runs the same recursion on the device (b2_core.cuh, b2_end_transition)."""
        if not tune:
            self._tuned_stats.append(float(accept_stat))
            return
        count, k, t0, gamma = self._count, self._k, self._t0, self._gamma
        w = 1.0 / (count + t0)
        self._hbar = (1 - w) * self._hbar + w * (self._target - accept_stat)
        self._log_step = self._mu - self._hbar * np.sqrt(count) / gamma
        mk = count ** -k
        self._log_bar = mk * self._log_step + (1 - mk) * self._log_bar
        self._count += 1

    def sync(self, step_size, step_size_bar, post_tune_accept):
        """Adopt the device state after a run."""
        self._log_step, self._log_bar = np.log(step_size), np.log(step_size_bar)
        self._tuned_stats.extend(np.asarray(post_tune_accept, dtype="f8").tolist())

    def stats(self):
        return {"step_size": np.exp(self._log_step), "step_size_bar": np.exp(self._log_bar)}

    def warnings(self):
        """step_sizes.py:60-79: acceptance rate vs target over (at most) the last 100 draws."""
        accept = np.array(self._tuned_stats)
        if accept.size == 0:
            return []
        return acceptance_warnings(float(np.mean(accept)), len(accept), self._target)
