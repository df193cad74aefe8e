"""Small run of every kernel family at tiny and ragged sizes, both dtypes, both exec modes (written for compute-sanitizer
memcheck, which this pool no longer allows; kept as a crash / launch-error smoke run)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import pymc3_b200 as pm
from pymc3_b200 import _capi
from tests import models_util

OPTS = dict(max_treedepth=6, early_max_treedepth=5, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75, t0=10.0,
            adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=64, hmc_jitter=1, exec_mode=0, glm_path=0)
pairs = models_util.pairs(glm_n=700, glm_k=9, hier_n=900, hier_g=7, sv_t=40)
for name, (model, oracle) in pairs.items():
    for dtype in ("float32", "float64"):
        for mode in (_capi.B2_EXEC_PERSISTENT, _capi.B2_EXEC_LOCKSTEP):
            eng = model.engine(5, dtype=dtype)
            D = eng.D
            eng.set_state(np.random.default_rng(0).uniform(-0.5, 0.5, (5, D)), np.arange(5) + 1, 0.05, np.zeros(D), np.ones(D), 10.0)
            o = dict(OPTS); o["exec_mode"] = mode
            eng.run(_capi.B2_NUTS, 12, 8, o)
            eng.run(_capi.B2_HMC, 6, 0, o)
            eng.close()
# chain-batched kernels at ragged sizes
X, y = models_util.glm_data(1000 + 37, 13, seed=7)
m = pm.LogisticGLM(X, y)
for dtype, paths in (("float32", (1, 2, 3)), ("float64", (1, 2))):
    eng = m.engine(70, dtype=dtype)
    q = np.random.default_rng(1).normal(size=(70, 14)) * 0.3
    for p in paths:
        eng.logp_dlogp(q, glm_path=p)
    eng.set_state(q * 0.1, np.arange(70), 0.05, np.zeros(14), np.ones(14), 10.0)
    o = dict(OPTS); o["exec_mode"] = _capi.B2_EXEC_LOCKSTEP
    eng.run(_capi.B2_NUTS, 10, 6, o)
    eng.close()
idx, fl, yy = models_util.hier_data(40000 + 123, 85, seed=2)
h = pm.HierLinearNCP(idx, fl, yy, 85)
eng = h.engine(130, dtype="float32")
eng.logp_dlogp(np.random.default_rng(2).normal(size=(130, 175)) * 0.2)
eng.close()
# block-per-chain persistent kernel (D > 1024): default layout and the two-blocks-per-SM layout (7 slots + data on chip)
import os
from pymc3_b200.model import sp500_log_returns
sv = pm.StochVol(sp500_log_returns()[:1100])
for dtype in ("float32", "float64"):
    for ctas in ("1", "2"):
        os.environ["B2_PBLOCK_CTAS"] = ctas
        eng = sv.engine(3, dtype=dtype)
        D = eng.D
        tp = sv.dict_to_array(sv.test_point)
        eng.set_state(tp + np.random.default_rng(3).uniform(-0.3, 0.3, (3, D)), np.arange(3) + 1, 0.02, np.zeros(D), np.ones(D), 10.0)
        o = dict(OPTS); o["exec_mode"] = _capi.B2_EXEC_PERSISTENT
        eng.run(_capi.B2_NUTS, 6, 4, o)
        eng.run(_capi.B2_NUTS, 3, 0, o)                  # continuation launch
        eng.close()
os.environ.pop("B2_PBLOCK_CTAS")
# dense metric: reparameterised lock-step run, both dtypes, NUTS and HMC
for dtype in ("float32", "float64"):
    eng = m.engine(70, dtype=dtype)
    A = np.random.default_rng(5).normal(size=(14, 14)) * 0.1
    eng.set_dense_mass(np.linalg.cholesky(A @ A.T + 0.05 * np.eye(14)))
    eng.set_state(np.random.default_rng(6).normal(size=(70, 14)) * 0.05, np.arange(70), 0.05, np.zeros(14), np.ones(14), 0.0)
    o = dict(OPTS); o["adapt_mass"] = 0
    eng.run(_capi.B2_NUTS, 8, 5, o)
    eng.run(_capi.B2_HMC, 4, 0, o)
    eng.position()
    eng.close()
print("sanitize smoke done")
