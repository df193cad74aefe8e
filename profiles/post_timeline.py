"""clock64 phases of chain 0 inside the fused lock-step companion kernel k_glm_tc_post (B2_TC_TIMELINE=1)."""
import ctypes as C
import os
import sys
import numpy as np
os.environ["B2_TC_TIMELINE"] = "1"
sys.path.insert(0, ".")
sys.argv = ["x"]
import bench
from pymc3_b200 import _capi
import pymc3_b200 as pm
X, y = bench.glm_synthetic(100000, 100)
model = pm.LogisticGLM(X, y)
C_ = 1024
eng = model.engine(C_, dtype="float32")
eng.set_state(bench.start_points(101, C_, 0), bench.chain_seeds(C_, 0), 0.25 / 101 ** 0.25, np.zeros(101), np.ones(101), 10.0)
opts = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75, t0=10.0,
            adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=0, exec_mode=0, glm_path=0)
eng.run(_capi.B2_NUTS, 400, 300, opts)
out = np.zeros((4096, 16), dtype=np.int64)
rc = _capi.load_library().b2_debug_post_timeline(eng.handle, out.ctypes.data_as(C.c_void_p))
assert rc == 0
rows = out[(out[:, 0] > 0) & (out[:, 8] > 0)]
names = ["state load", "finalize", "finish leaf/merges", "top merge", "end transition", "begin transition", "begin doubling+prepare", "state store"]
d = np.diff(rows[:, :9], axis=1)
kinds = {"plain leaf": (d[:, 3] < 50) & (d[:, 4] < 50), "transition end": d[:, 4] > 50}
for k, m in kinds.items():
    if m.sum() == 0:
        continue
    print("%-15s n=%4d total %7.0f cycles:" % (k, m.sum(), rows[m, 8].mean() - rows[m, 0].mean()),
          ", ".join("%s %.0f" % (n, v) for n, v in zip(names, d[m].mean(axis=0))))
