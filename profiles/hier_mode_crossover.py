"""Persistent (warp per chain evaluates its own likelihood) vs lock-step (chain-batched k_hier_slab + advance kernel) on the
hierarchical model at mid sizes: where should B2_EXEC_AUTO switch?"""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
sys.argv = ["x"]
import bench
from pymc3_b200 import _capi
import pymc3_b200 as pm

for n, C in ((20000, 1024), (20000, 128), (50000, 1024), (100000, 1024), (100000, 4096)):
    idx, floor, y, g = bench.hier_synthetic(n)
    model = pm.HierLinearNCP(idx, floor, y, g)
    D = model.ndim
    tp = model.dict_to_array(model.test_point)
    q0 = np.stack([tp + np.random.default_rng([7, c]).uniform(-1, 1, size=D) for c in range(C)])
    for mode, name in ((_capi.B2_EXEC_LOCKSTEP, "lock-step"), (_capi.B2_EXEC_PERSISTENT, "persistent")):
        opts = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75, t0=10.0,
                    adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=0, exec_mode=mode, glm_path=0)
        eng = model.engine(C, dtype="float32")
        eng.set_state(q0, bench.chain_seeds(C, 0), 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
        trace = eng.alloc_trace(_capi.B2_NUTS, 40)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            eng.run(_capi.B2_NUTS, 40, 40, opts, out=trace, row0=0)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            gr = sum(r.n_grad for r in eng.reports())
            print("N=%6d C=%4d %-10s %7.3f s  %9.3f M grad-evals/s  (mean tree %.0f)" % (n, C, name, dt, gr / dt / 1e6, gr / C / 40))
        except Exception as err:
            print("N=%6d C=%4d %-10s failed: %s" % (n, C, name, err))
        eng.close()
