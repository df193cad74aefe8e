"""Target for `ncu -k regex:k_glm_tcw_main`: the wide tensor-core GLM likelihood on a C5-like shard (3 M rows x 256
features, bf16-representable X -> 2 split passes, 256 chains live) through b2_logp_dlogp, six times."""
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
eng = pm.LogisticGLM(X, y).engine(chains, dtype="float32")
for _ in range(3):
    eng.logp_dlogp(q, glm_path=_capi.B2_GLM_TCGEN05)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.logp_dlogp(q, glm_path=_capi.B2_GLM_TCGEN05)
e1.record()
torch.cuda.synchronize()
print("%.1f us per call (compact + reference refresh + k_glm_tcw_main + finalize)" % (e0.elapsed_time(e1) * 1e3 / 3))
eng.close()
