"""Target for `ncu -k regex:k_persistent_block`: config 4 (stochastic volatility, D = 2907) on 296 chains = one full
wave of two blocks per SM.  Launch 1 tunes for 120 transitions, launch 2 (the one to capture: `-s 1 -c 1`) runs 4 more."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
N_CHAINS = int(sys.argv[1]) if len(sys.argv) > 1 else 296
sys.argv = ["x"]
import bench
from pymc3_b200 import _capi
import pymc3_b200 as pm

C = N_CHAINS
model = pm.StochVol()
D = model.ndim
opts = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75, t0=10.0,
            adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=0, exec_mode=_capi.B2_EXEC_AUTO, glm_path=0)
eng = model.engine(C, dtype="float32")
tp = model.dict_to_array(model.test_point)
q0 = np.stack([tp + np.random.default_rng([7, c]).uniform(-1, 1, size=D) for c in range(C)])
eng.set_state(q0, bench.chain_seeds(C, 0), 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
trace = eng.alloc_trace(_capi.B2_NUTS, 124)
eng.run(_capi.B2_NUTS, 120, 1000, opts, out=trace, row0=0)
torch.cuda.synchronize()
g0 = sum(r.n_grad for r in eng.reports())
t0 = time.perf_counter()
eng.run(_capi.B2_NUTS, 4, 1000, opts, out=trace, row0=120)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
g = sum(r.n_grad for r in eng.reports()) - g0
print("second launch: %d chains x 4 transitions, %d leapfrogs in %.1f ms = %.2f M grad-evals/s" % (C, g, dt * 1e3, g / dt / 1e6))
eng.close()
