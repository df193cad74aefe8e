"""pymc3_b200 -- a B200-native NUTS / HamiltonianMC engine behind PyMC3's sampler API.

    import pymc3_b200 as pm
    with pm.EightSchoolsNCP() as model:
        trace = pm.sample(1000, tune=500, chains=4, step=pm.NUTS())
    trace["mu"], trace.get_sampler_stats("depth")

Host code is Python; the hot path is hand-written CUDA for sm_100a reached through the C ABI in
include/b200nuts.h (pymc3_b200/_capi.py).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import glm, stats  # noqa: F401
# stats_device (torch tensor ops on the device trace) is imported on demand: `from pymc3_b200 import stats_device`
from .backends import MultiTrace, NDArray, load_trace, merge_traces, save_trace  # noqa: F401
from .exceptions import ParallelSamplingError, SamplingError  # noqa: F401
from .model import (EightSchoolsNCP, HierLinearNCP, LogisticGLM, Model, StdNormal, StochVol,  # noqa: F401
                    ValueGradFunction, modelcontext)
from .sampling import init_nuts, iter_sample, sample, trace_cov  # noqa: F401
from .stats import ess, rhat  # noqa: F401
from .step_methods.hmc import NUTS, HamiltonianMC  # noqa: F401
from .step_methods.hmc.quadpotential import (QuadPotentialDiag, QuadPotentialDiagAdapt,  # noqa: F401
                                             quad_potential)
