"""Wide tcgen05 GLM likelihood (b2_glm_tcw.cu) timed alone through the parity hook b2_logp_dlogp: all chains live."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pymc3_b200 import model as pm, _capi

dev = torch.device("cuda", 0)
for rows, k, chains in [(3000000, 256, 256), (3000000, 256, 512), (3000000, 256, 1024), (1000000, 160, 1024)]:
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    X = torch.randn((rows, k), generator=gen, device=dev, dtype=torch.float32)
    y = (torch.rand(rows, generator=gen, device=dev) < 0.5).float()
    model = pm.LogisticGLM(X, y)
    eng = model.engine(chains, dtype="float32")
    q = torch.randn((chains, k + 1), generator=gen, device=dev, dtype=torch.float32) * 0.05
    for path, name in [(_capi.B2_GLM_TCGEN05, "tcgen05-wide"), (_capi.B2_GLM_SIMT, "simt")]:
        if name == "simt" and chains > 256:
            continue
        eng.logp_dlogp(q, glm_path=path)
        torch.cuda.synchronize()
        n = 5 if name == "simt" else 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            eng.logp_dlogp(q, glm_path=path)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 4.0 * rows * (k + 1) * chains
        print("%-13s rows %d K %d chains %4d: %.3f ms per all-chains likelihood (3 kernels), %.0f TFLOP/s algorithmic, X tiles %.0f GB/s"
              % (name, rows, k, chains, ms, fl / ms / 1e9, rows * 256 * 4 * (chains // 128) / ms / 1e6), flush=True)
    eng.close()
    del X, y, model
