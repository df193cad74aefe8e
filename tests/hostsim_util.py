"""ctypes driver of the CPU build of the engine's state machine (tests/hostsim).  TEST HARNESS."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymc3_b200 import _capi  # noqa: E402  (struct definitions only; no library is loaded)
from tests.hostsim import build as _build  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.hostsim_normal.restype = C.c_double
        _lib.hostsim_normal.argtypes = [C.c_uint32] * 4
    return _lib


class HostArrays:
    """Builds a b2_model_desc whose 'device' pointers are host NumPy buffers."""

    def __init__(self, model):
        self.keep = []
        self.desc = model._describe(self._upload)

    def _upload(self, arr, dtype):
        a = np.ascontiguousarray(arr, dtype=dtype)
        self.keep.append(a)
        return a.ctypes.data


NUTS_DEFAULTS = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05,
                     k=0.75, t0=10.0, adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024,
                     hmc_jitter=1, exec_mode=0, glm_path=0)


def run(model, q0, seeds, n_iters, tune, kind="nuts", dtype="float64", step_size0=None, mass_mean=None,
        mass_var=None, mass_weight=10.0, window=101, **opts):
    ha = HostArrays(model)
    D = ha.desc.D
    q0 = np.ascontiguousarray(q0, dtype="f8").reshape(-1, D)
    Cn = q0.shape[0]
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    o = dict(NUTS_DEFAULTS)
    if kind == "hmc":
        o["target_accept"] = 0.65
    o.update(opts)
    so = _capi.SamplerOpts(kind=_capi.B2_NUTS if kind == "nuts" else _capi.B2_HMC, n_iters=n_iters,
                           tune_until=tune, **o)
    npd = np.dtype(dtype)
    out = {"q": np.zeros((n_iters, Cn, D), dtype=npd)}
    tr = _capi.TraceOut()
    tr.d_q = out["q"].ctypes.data
    codes = {"energy": "f8", "energy_error": "f8", "max_energy_error": "f8", "mean_tree_accept": "f8",
             "step_size": "f8", "step_size_bar": "f8", "model_logp": "f8", "accept": "f8", "depth": "i4",
             "tree_size": "i4", "n_steps": "i4", "diverging": "u1", "tune": "u1", "accepted": "u1"}
    for name, code in codes.items():
        out[name] = np.zeros((n_iters, Cn), dtype=code)
        setattr(tr, "d_" + name, out[name].ctypes.data)
    rep = (_capi.ChainReport * Cn)()
    if step_size0 is None:
        step_size0 = 0.25 / D ** 0.25
    mm = np.zeros(D) if mass_mean is None else np.ascontiguousarray(mass_mean, dtype="f8")
    mv = np.ones(D) if mass_var is None else np.ascontiguousarray(mass_var, dtype="f8")
    rc = lib().hostsim_run(C.byref(ha.desc), Cn, 1 if npd == np.float64 else 0, q0.ctypes.data_as(C.c_void_p),
                           seeds.ctypes.data_as(C.c_void_p), C.c_double(step_size0), mm.ctypes.data_as(C.c_void_p),
                           mv.ctypes.data_as(C.c_void_p), C.c_double(mass_weight), window, C.byref(so), C.byref(tr), rep)
    assert rc == 0
    out["reports"] = list(rep)
    return out


def logp_dlogp(model, q, dtype="float64"):
    ha = HostArrays(model)
    D = ha.desc.D
    q = np.ascontiguousarray(q, dtype="f8").reshape(-1, D)
    lp = np.zeros(len(q))
    g = np.zeros_like(q)
    lib().hostsim_logp(C.byref(ha.desc), 1 if np.dtype(dtype) == np.float64 else 0, q.ctypes.data_as(C.c_void_p),
                       len(q), lp.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p))
    return lp, g
