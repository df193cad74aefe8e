"""Small target for `ncu -k regex:k_glm_tc_main`: the C2 likelihood launch (100 000 x 100, 1024 chains, all live) through
b2_logp_dlogp, six times.  Usage (one GPU, plain run first):
    python profiles/tc_ncu_target.py && ncu --set full --clock-control none --import-source on \
        -k regex:k_glm_tc_main -s 3 -c 2 -o gpurun_out/prof_main_r2 python profiles/tc_ncu_target.py"""
import sys
import torch
sys.path.insert(0, ".")
sys.argv = ["x"]
import bench
from pymc3_b200 import _capi
import pymc3_b200 as pm

X, y = bench.glm_synthetic(100000, 100)
eng = pm.LogisticGLM(X, y).engine(1024, dtype="float32")
q = torch.randn(1024, 101, device="cuda") * 0.1
for _ in range(6):
    eng.logp_dlogp(q, glm_path=_capi.B2_GLM_TCGEN05)
torch.cuda.synchronize()
eng.close()
print("ok")
