"""Times the chain-batched GLM likelihood (b2_logp_dlogp, tcgen05 path) alone for several chain
counts: per-tile period of k_glm_tc_main = time * clock / (tiles per CTA).  Used to separate
per-CTA compute limits from shared L2 bandwidth (profiles/README.md)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
sys.argv = ["x"]
import bench
from pymc3_b200 import _capi
import pymc3_b200 as pm

X, y = bench.glm_synthetic(100000, 100)
model = pm.LogisticGLM(X, y)
for chains in (128, 256, 512, 1024, 2048):
    eng = model.engine(chains, dtype="float32")
    q = torch.randn(chains, 101, device="cuda") * 0.1
    for _ in range(5):
        eng.logp_dlogp(q, glm_path=_capi.B2_GLM_TCGEN05)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        eng.logp_dlogp(q, glm_path=_capi.B2_GLM_TCGEN05)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    ctiles = (chains + 127) // 128
    splits = max(1, 148 // ctiles)
    tiles = -(-1563 // splits)
    print("chains %5d  %8.1f us per call (incl. finalize + host alloc)  chain_tiles %2d splits %3d tiles/CTA %3d  -> %.2f us per tile"
          % (chains, us, ctiles, splits, tiles, us / tiles))
    eng.close()
