"""Target for `ncu -k regex:k_glm_tc_post`: a short C2 lock-step job (1024 chains, 100 000 x 100) so that the state-machine
kernel is captured in its working regime (tuned step size, ~974 live chains, a few transition ends per launch).
B2_GRAPH=0 keeps every launch a plain kernel launch.  Usage: see profiles/calls/call_r2ao.sh."""
import os
import sys
import numpy as np
import torch
os.environ["B2_GRAPH"] = "0"
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
eng = model.engine(C, dtype="float32")
eng.set_state(bench.start_points(D, C, 0), bench.chain_seeds(C, 0), 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
trace = eng.alloc_trace(_capi.B2_NUTS, 150)
eng.run(_capi.B2_NUTS, 150, 150, opts, out=trace, row0=0)
torch.cuda.synchronize()
print("launches:", eng.kernel_launches(), "leapfrogs:", sum(r.n_grad for r in eng.reports()))
eng.close()
