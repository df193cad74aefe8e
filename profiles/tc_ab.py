"""A/B timing of k_glm_tc_main builds through b2_logp_dlogp at C2 size (1024 chains, all live): microseconds per call
(compact + reference refresh + main + finalize + ~15 us of host overhead, identical across variants)."""
import os
import sys
import torch
sys.path.insert(0, ".")
sys.argv = ["x"]
import bench
from pymc3_b200 import _capi
import pymc3_b200 as pm

X, y = bench.glm_synthetic(100000, 100)
model = pm.LogisticGLM(X, y)
KEYS = ("B2_TC_EPI", "B2_TC_FLUSH", "B2_TC_NOREF")


def run(chains, n=200, **env):
    for k in KEYS:
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    eng = model.engine(chains, dtype="float32")
    q = torch.randn(chains, 101, device="cuda") * 0.1
    for _ in range(10):
        eng.logp_dlogp(q, glm_path=_capi.B2_GLM_TCGEN05)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.logp_dlogp(q, glm_path=_capi.B2_GLM_TCGEN05)
    e1.record()
    torch.cuda.synchronize()
    eng.close()
    return e0.elapsed_time(e1) * 1e3 / n


for rep in range(2):
    for env in (dict(B2_TC_EPI=0, B2_TC_NOREF=1), dict(B2_TC_EPI=1, B2_TC_NOREF=1), dict(B2_TC_EPI=0), dict(B2_TC_EPI=1),
                dict(B2_TC_EPI=1, B2_TC_FLUSH=32), dict(B2_TC_EPI=1, B2_TC_FLUSH=44)):
        print("rep %d  %-44s 1024 chains: %7.1f us   512 chains: %7.1f us" % (rep, env, run(1024, **env), run(512, **env)))
