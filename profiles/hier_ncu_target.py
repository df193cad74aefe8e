"""Target for `ncu -k regex:k_hier_slab`: config 3's likelihood launch (1 000 000 observations, 85 groups, 4096 chains, all
live) through b2_logp_dlogp, five times; prints the CUDA-event time per call."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
sys.argv = ["x"]
import bench
import pymc3_b200 as pm

idx, floor, y, g = bench.hier_synthetic(1000000)
eng = pm.HierLinearNCP(idx, floor, y, g).engine(4096, dtype="float32")
truth = np.concatenate([[1.5, np.log(0.3), -0.7, np.log(0.3)], np.zeros(2 * g), [np.log(0.7)]])
q = torch.as_tensor(truth + np.random.default_rng(1).normal(size=(4096, 2 * g + 5)) * 0.05, device="cuda", dtype=torch.float32)
for _ in range(2):
    eng.logp_dlogp(q)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.logp_dlogp(q)
e1.record()
torch.cuda.synchronize()
print("k_hier_compact + k_hier_slab + k_hier_finalize: %.1f us per call at 4096 live chains x 1 M observations" % (e0.elapsed_time(e1) * 1e3 / 3))
eng.close()
