"""A/B timing of the wide tensor-core GLM likelihood (k_glm_tcw_main) through b2_logp_dlogp: 3 M rows x 256 features (bf16-
representable X, like config C5's generator), 256 chains live.  us per call incl. compact + reference refresh + finalize."""
import os
import sys
import torch
sys.path.insert(0, ".")
from pymc3_b200 import _capi
import pymc3_b200 as pm

rows, k, chains = 3 * 2 ** 20, 256, 256
gen = torch.Generator(device="cuda")
gen.manual_seed(1)
X = torch.randn((rows, k), generator=gen, device="cuda").bfloat16().float()
y = (torch.rand(rows, generator=gen, device="cuda") < 0.5).float()
q = torch.randn(chains, k + 1, device="cuda") * 0.02


def run(n=20, **env):
    for key in ("B2_TCW_PASSES", "B2_TCW_FLUSH", "B2_TC_NOREF", "B2_TCW_EPI"):
        os.environ.pop(key, None)
    for key, v in env.items():
        os.environ[key] = str(v)
    eng = pm.LogisticGLM(X, y).engine(chains, dtype="float32")
    for _ in range(3):
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


for env in (dict(B2_TC_NOREF=1, B2_TCW_EPI=0), dict(B2_TC_NOREF=1, B2_TCW_EPI=1), dict(B2_TC_NOREF=1, B2_TCW_EPI=1, B2_TCW_FLUSH=1000000),
            dict(B2_TC_NOREF=1, B2_TCW_PASSES=3, B2_TCW_EPI=0), dict(B2_TC_NOREF=1, B2_TCW_PASSES=3, B2_TCW_EPI=1), dict()):
    print("%-70s %8.1f us per call" % (env, run(**env)))
