"""ctypes binding of include/b200nuts.h (libb200nuts.so).

This is the "thin C-ABI layer" of BASELINE.json's north star: host code stays Python, device
buffers are PyTorch tensors, and every hot-path call goes through the functions below.
There is no CPU fallback: if the library (or a CUDA device) is missing, loading fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200nuts.so")

ABI_VERSION = 4          # include/b200nuts.h B2_ABI_VERSION this binding's struct layouts were written for
B2_F32, B2_F64 = 0, 1
B2_STD_NORMAL, B2_EIGHT_SCHOOLS_NCP, B2_GLM_LOGIT, B2_HIER_LINEAR_NCP, B2_STOCH_VOL = 0, 2, 3, 4, 5
B2_NUTS, B2_HMC = 0, 1
B2_EXEC_AUTO, B2_EXEC_PERSISTENT, B2_EXEC_LOCKSTEP = 0, 1, 2
B2_GLM_AUTO, B2_GLM_GROUP, B2_GLM_SIMT, B2_GLM_TCGEN05 = 0, 1, 2, 3
PHASE_DONE, PHASE_FAILED = 3, 4
FAIL_BAD_INITIAL_ENERGY = 1


class ModelDesc(C.Structure):
    _fields_ = [("family", C.c_int32), ("D", C.c_int32), ("N", C.c_int32), ("G", C.c_int32),
                ("d_aux0", C.c_void_p), ("d_aux1", C.c_void_p), ("d_X", C.c_void_p), ("d_y", C.c_void_p),
                ("d_floor", C.c_void_p), ("d_grp_off", C.c_void_p), ("hp", C.c_double * 4)]


class SamplerOpts(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_iters", C.c_int32), ("tune_until", C.c_int32),
                ("max_treedepth", C.c_int32), ("early_max_treedepth", C.c_int32),
                ("Emax", C.c_double), ("target_accept", C.c_double),
                ("gamma", C.c_double), ("k", C.c_double), ("t0", C.c_double),
                ("adapt_step_size", C.c_int32), ("adapt_mass", C.c_int32),
                ("path_length", C.c_double), ("max_steps", C.c_int32), ("hmc_jitter", C.c_int32),
                ("exec_mode", C.c_int32), ("glm_path", C.c_int32), ("run_ahead", C.c_int32)]


TRACE_FIELDS = ["d_q", "d_energy", "d_energy_error", "d_max_energy_error", "d_mean_tree_accept",
                "d_step_size", "d_step_size_bar", "d_model_logp", "d_accept", "d_depth", "d_tree_size",
                "d_n_steps", "d_diverging", "d_tune", "d_accepted"]


class TraceOut(C.Structure):
    _fields_ = [(name, C.c_void_p) for name in TRACE_FIELDS]


class ChainReport(C.Structure):
    _fields_ = [("phase", C.c_int32), ("fail_code", C.c_int32), ("iter", C.c_int32),
                ("n_div_post", C.c_int32), ("n_maxdepth_post", C.c_int32), ("n_post", C.c_int32),
                ("n_grad", C.c_int64), ("step_size", C.c_double), ("step_size_bar", C.c_double)]


EXPORTS = ["b2_abi_version", "b2_last_error", "b2_engine_create", "b2_engine_destroy", "b2_logp_dlogp",
           "b2_set_state", "b2_set_position", "b2_sample_run", "b2_get_chain_reports", "b2_get_mass_var", "b2_get_position",
           "b2_kernel_launches", "b2_set_profiling", "b2_get_profile", "b2_get_profile_advance", "b2_step_begin", "b2_step_likelihood",
           "b2_step_advance", "b2_step_active", "b2_step_end", "b2_leapfrog", "b2_set_dense_mass"]

_lib = None


class B2Error(RuntimeError):
    pass


def load_library(path=None):
    """dlopen libb200nuts.so and declare prototypes.  Raises if the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise B2Error("%s not found: build it with `python -m pymc3_b200.build` "
                      "(there is no CPU fallback for the sampler hot path)" % path)
    lib = C.CDLL(path)
    lib.b2_abi_version.restype = C.c_int
    lib.b2_last_error.restype = C.c_char_p
    lib.b2_engine_create.argtypes = [C.POINTER(ModelDesc), C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    lib.b2_engine_destroy.argtypes = [C.c_void_p]
    lib.b2_logp_dlogp.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.b2_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                 C.c_double, C.c_int32, C.c_void_p]
    lib.b2_set_position.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.b2_sample_run.argtypes = [C.c_void_p, C.POINTER(SamplerOpts), C.POINTER(TraceOut), C.c_void_p]
    lib.b2_get_chain_reports.argtypes = [C.c_void_p, C.POINTER(ChainReport)]
    lib.b2_get_mass_var.argtypes = [C.c_void_p, C.c_void_p]
    lib.b2_get_position.argtypes = [C.c_void_p, C.c_void_p]
    lib.b2_kernel_launches.argtypes = [C.c_void_p]
    lib.b2_kernel_launches.restype = C.c_int64
    lib.b2_step_begin.argtypes = [C.c_void_p, C.POINTER(SamplerOpts), C.POINTER(TraceOut), C.c_void_p]
    lib.b2_step_likelihood.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.b2_step_advance.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.b2_step_active.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]
    lib.b2_step_end.argtypes = [C.c_void_p]
    lib.b2_set_dense_mass.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.b2_leapfrog.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.b2_get_profile_advance.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.b2_set_profiling.argtypes = [C.c_void_p, C.c_int32]
    lib.b2_get_profile.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    for name in EXPORTS:
        getattr(lib, name)
    if lib.b2_abi_version() != ABI_VERSION:
        raise B2Error("%s has ABI version %d, this binding expects %d: rebuild it with "
                      "`python -m pymc3_b200.build --force`" % (path, lib.b2_abi_version(), ABI_VERSION))
    if path == LIB_PATH:
        _lib = lib
    return lib


def check(rc, lib=None):
    if rc != 0:
        lib = lib or load_library()
        raise B2Error("libb200nuts error %d: %s" % (rc, lib.b2_last_error().decode()))
