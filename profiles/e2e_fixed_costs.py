"""Host-side fixed costs of one C2 engine: create (upload + allocations), first run (lazy tcgen05 setup, X pre-tiling,
module loading, graph capture), a second run, destroy."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
sys.argv = ["x"]
import bench
from pymc3_b200 import _capi
import pymc3_b200 as pm

X, y = bench.glm_synthetic(100000, 100)
model = pm.LogisticGLM(X, y)
C, D = 1024, 101
opts = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75, t0=10.0,
            adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=0, exec_mode=_capi.B2_EXEC_AUTO, glm_path=0)
torch.zeros(1, device="cuda")
torch.cuda.synchronize()
for rep in range(3):
    t = [time.perf_counter()]
    eng = model.engine(C, dtype="float32")
    torch.cuda.synchronize(); t.append(time.perf_counter())
    eng.set_state(bench.start_points(D, C, 0), bench.chain_seeds(C, 0), 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
    trace = eng.alloc_trace(_capi.B2_NUTS, 2000)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    eng.run(_capi.B2_NUTS, 1, 1000, opts, out=trace, row0=0)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    eng.run(_capi.B2_NUTS, 1, 1000, opts, out=trace, row0=1)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    del trace
    eng.close()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    print("rep %d: create %.1f ms | set_state + trace alloc %.1f ms | first transition %.1f ms | second %.1f ms | destroy %.1f ms" % (
        (rep,) + tuple((b - a) * 1e3 for a, b in zip(t[:-1], t[1:]))))
