"""Host-side view of the dual-averaging state that lives on the device.

The recursion of pymc3/step_methods/step_sizes.py:21-58 runs inside the CUDA state machine
(csrc/b2_core.cuh, b2_end_transition).  This object keeps the reference's attribute surface
(`current`, `stats`, `warnings`) for one chain, fed from the device trace / chain report.
"""
import numpy as np
from scipy import stats

from ..backends.report import SamplerWarning, WarningType


def acceptance_warnings(mean_accept, n_draws, target):
    """step_sizes.py:60-79: is the target inside the 95 % beta interval of the mean acceptance rate, counted over
    (at most) 100 draws?"""
    if not n_draws or not np.isfinite(mean_accept):
        return []
    n_bound = min(100, int(n_draws))
    n_good, n_bad = mean_accept * n_bound, (1 - mean_accept) * n_bound
    lower, upper = stats.beta(n_good + 1, n_bad + 1).interval(0.95)
    if target < lower or target > upper:
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
        self._tuned_stats = []

    def current(self, tune):
        return np.exp(self._log_step) if tune else np.exp(self._log_bar)

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
