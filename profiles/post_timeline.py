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
rows = out[(out[:, 0] > 0) & (out[:, 8] > out[:, 0]) & (out[:, 7] > 0)]
names = ["state load", "finalize", "finish leaf/merges", "top merge", "end transition", "begin transition", "begin doubling+prepare", "state store"]
d = np.diff(rows[:, :9], axis=1)
tot = rows[:, 8] - rows[:, 0]
ended = rows[:, 12] != rows[:, 11]
nm = rows[:, 10]
print("%-28s %5s %8s %8s   %s" % ("kind", "n", "mean", "p95", " | ".join(names)))
for label, m in [("leaf, 0 merges", (nm == 0) & ~ended), ("leaf, 1-2 merges", (nm >= 1) & (nm <= 2) & ~ended),
                 ("leaf, 3-5 merges", (nm >= 3) & (nm <= 5) & ~ended), ("leaf, 6+ merges", (nm >= 6) & ~ended),
                 ("transition end (tuning)", ended & (rows[:, 11] < 300)), ("transition end (sampling)", ended & (rows[:, 11] >= 300))]:
    if m.sum() == 0:
        continue
    print("%-28s %5d %8.0f %8.0f   %s" % (label, m.sum(), tot[m].mean(), np.percentile(tot[m], 95),
                                         " | ".join("%6.0f" % v for v in d[m].mean(axis=0))))
print("all rows: mean %.0f p50 %.0f p95 %.0f p99 %.0f max %.0f cycles" % (tot.mean(), np.percentile(tot, 50), np.percentile(tot, 95), np.percentile(tot, 99), tot.max()))
