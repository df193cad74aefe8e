"""NUTS step method (pymc3/step_methods/hmc/nuts.py:36-208) over the device engine."""
import numpy as np

from ... import _capi
from ...backends.report import SamplerWarning, WarningType
from ..arraystep import Competence
from .base_hmc import BaseHMC

__all__ = ["NUTS"]


class NUTS(BaseHMC):
    """No-U-Turn sampler; the tree doubling of nuts.py:220-406 runs iteratively on the device
    (csrc/b2_core.cuh).  Sampler statistics as in nuts.py:41-103."""

    name = "nuts"
    _kind = _capi.B2_NUTS
    default_blocked = True
    generates_stats = True
    stats_dtypes = [{
        "depth": np.int64, "step_size": np.float64, "tune": np.bool_, "mean_tree_accept": np.float64,
        "step_size_bar": np.float64, "tree_size": np.float64, "diverging": np.bool_,
        "energy_error": np.float64, "energy": np.float64, "max_energy_error": np.float64,
        "model_logp": np.float64,
    }]

    def __init__(self, vars=None, max_treedepth=10, early_max_treedepth=8, **kwargs):
        super().__init__(vars, **kwargs)
        if not (1 <= max_treedepth <= 12 and 1 <= early_max_treedepth <= 12):
            raise ValueError("tree depths must be in [1, 12] on the device")
        self.max_treedepth = max_treedepth
        self.early_max_treedepth = early_max_treedepth
        self._reached_max_treedepth = 0

    def _account(self, stats, it):
        super()._account(stats, it)
        limit = self.early_max_treedepth if (self.tune and it < 200) else self.max_treedepth
        if not self.tune and int(stats["depth"]) >= limit and not bool(stats["diverging"]):
            self._reached_max_treedepth += 1             # nuts.py:182-184 (upper bound: turning at the last level)

    @staticmethod
    def competence(var, has_grad):
        """nuts.py:190-195."""
        dtype = getattr(var, "dtype", np.dtype("float64"))
        if np.issubdtype(np.dtype(dtype), np.floating) and has_grad:
            return Competence.IDEAL
        return Competence.INCOMPATIBLE

    def _treedepth_warning(self, n_samples, n_treedepth):
        if n_samples > 0 and n_treedepth / float(n_samples) > 0.05:       # nuts.py:197-208
            msg = ("The chain reached the maximum tree depth. Increase max_treedepth, increase "
                   "target_accept or reparameterize.")
            return [SamplerWarning(WarningType.TREEDEPTH, msg, "warn", None, None, None)]
        return []

    def warnings(self):
        return super().warnings() + self._treedepth_warning(self._samples_after_tune, self._reached_max_treedepth)

    def _chain_warnings(self, report, mean_accept_post, n_post, diverging_rows, tune_flags, accept_ok=None):
        return (super()._chain_warnings(report, mean_accept_post, n_post, diverging_rows, tune_flags, accept_ok)
                + self._treedepth_warning(report.n_post, report.n_maxdepth_post))
