"""Dumps the clock64 timeline of CTA (0,0) of k_glm_tc_main (B2_TC_TIMELINE=1): per tile, when the MMA
thread issued GEMM1/GEMM2 and when the first epilogue warp of each pair waited / computed / stored."""
import ctypes as C
import os
import sys
import numpy as np
import torch
os.environ["B2_TC_TIMELINE"] = "1"
sys.path.insert(0, ".")
sys.argv = ["x"]
import bench
from pymc3_b200 import _capi
import pymc3_b200 as pm
X, y = bench.glm_synthetic(100000, 100)
model = pm.LogisticGLM(X, y)
eng = model.engine(1024, dtype="float32")
q = torch.randn(1024, 101, device="cuda") * 0.1
for _ in range(3):
    eng.logp_dlogp(q, glm_path=_capi.B2_GLM_TCGEN05)
torch.cuda.synchronize()
out = np.zeros((48, 256), dtype=np.int64)
lib = _capi.load_library()
rc = lib.b2_debug_tc_timeline(eng.handle, out.ctypes.data_as(C.c_void_p))
assert rc == 0, rc
names = ["G1_issue_begin", "G1_issue_end", "G2_issue_begin", "G2_issue_end", "epi_wait_begin", "epi_S_ready",
         "epi_S_loaded", "epi_math_done", "epi_P_stored"]
t0 = out[out > 0].min()
T = 87
print("tile " + " ".join("%15s" % n for n in names))
for t in list(range(0, 14)) + list(range(40, 48)) + list(range(80, 87)):
    print("%4d " % t + " ".join("%15d" % (out[e, t] - t0 if out[e, t] else -1) for e in range(9)))
d = np.diff(out[8, 20:80:2])
print("pair A period per own tile (cycles):", d.mean(), " -> per tile", d.mean() / 2)
print("mean epi wait for S:", (out[5, 20:80] - out[4, 20:80]).mean(), " S load:", (out[6, 20:80] - out[5, 20:80]).mean(),
      " math:", (out[7, 20:80] - out[6, 20:80]).mean(), " P store:", (out[8, 20:80] - out[7, 20:80]).mean())
print("mean G1 issue time:", (out[1, 20:80] - out[0, 20:80]).mean(), " G2 issue time:", (out[3, 20:80] - out[2, 20:80]).mean())
ev = []
for e in range(9):
    for t in range(38, 52):
        if out[e, t]:
            ev.append((out[e, t] - t0, names[e], t))
print("--- merged event log, tiles 38..51")
for c, n, t in sorted(ev):
    who = "MMA " if n.startswith("G") else ("epiA" if t % 2 == 0 else "epiB")
    print("%8d %s %-16s tile %d" % (c, who, n, t))

print("--- per-warp skew (cycles after the pair's first warp) for S_loaded / P_stored, tiles 40..47")
for t in range(40, 48):
    ws_ = [w for w in range(16) if ((w >> 2) >> 1) == (t & 1)]
    sl = np.array([out[9 + 2 * w, t] for w in ws_]); ps = np.array([out[10 + 2 * w, t] for w in ws_])
    print("tile %d warps %s  S_loaded +%s   P_stored +%s" % (t, [w + 4 for w in ws_], (sl - sl.min()).tolist(), (ps - ps.min()).tolist()))
