"""HamiltonianMC step method (pymc3/step_methods/hmc/hmc.py:30-159) over the device engine."""
import numpy as np

from ... import _capi
from ..arraystep import Competence
from .base_hmc import BaseHMC

__all__ = ["HamiltonianMC"]


def unif(step_size, elow=.85, ehigh=1.15):
    """hmc.py:26-27 -- on the device the jitter uses the chain's Philox stream."""
    return np.random.uniform(elow, ehigh) * step_size


unif._b2_unif = True


class HamiltonianMC(BaseHMC):
    """Fixed-length trajectories with a Metropolis accept (hmc.py:110-152)."""

    name = "hmc"
    _kind = _capi.B2_HMC
    default_blocked = True
    generates_stats = True
    stats_dtypes = [{
        "step_size": np.float64, "n_steps": np.int64, "tune": np.bool_, "step_size_bar": np.float64,
        "accept": np.float64, "diverging": np.bool_, "energy_error": np.float64, "energy": np.float64,
        "path_length": np.float64, "accepted": np.bool_, "model_logp": np.float64,
    }]

    def __init__(self, vars=None, path_length=2., max_steps=1024, **kwargs):
        kwargs.setdefault("step_rand", unif)             # hmc.py:104-105
        kwargs.setdefault("target_accept", 0.65)
        super().__init__(vars, **kwargs)
        self.path_length = path_length
        self.max_steps = max_steps

    @staticmethod
    def competence(var, has_grad):
        """hmc.py:154-159."""
        dtype = getattr(var, "dtype", np.dtype("float64"))
        if not np.issubdtype(np.dtype(dtype), np.floating) or not has_grad:
            return Competence.INCOMPATIBLE
        return Competence.COMPATIBLE
