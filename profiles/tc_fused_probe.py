"""Where does a fused launch of k_glm_tc_main spend its time?  (round 2)

(1) likelihood alone through b2_logp_dlogp for 512 / 1024 chains: two-kernel build (6-stage ring, 96 registers)
    vs the fused build with no state-machine work (B2_TC_HOOK_FUSED=1; 80 registers, 5- or 4-stage ring), both
    epilogues;
(2) a short lock-step NUTS job at C2 size with B2_TC_ROLE_CLOCKS=1: cycles of the likelihood CTAs and of the
    state-machine warps over all fused launches, for several shared-memory budgets (a budget <= 195 KB
    leaves 32 KB of L1 for the state machine's local-memory traffic).
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.argv = ["x"]
import bench
from pymc3_b200 import _capi
import pymc3_b200 as pm

X, y = bench.glm_synthetic(100000, 100)
model = pm.LogisticGLM(X, y)


def set_env(**kw):
    for k in ("B2_TC_FUSED", "B2_TC_EPI", "B2_TC_STAGES", "B2_TC_POST_LEVELS", "B2_TC_SMEM_CAP", "B2_TC_HOOK_FUSED",
              "B2_TC_ROLE_CLOCKS"):
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ[k] = str(v)


def time_hook(chains, n=40):
    eng = model.engine(chains, dtype="float32")
    q = torch.randn(chains, 101, device="cuda") * 0.1
    for _ in range(5):
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


print("== (1) likelihood alone (b2_logp_dlogp = compact + main + finalize), us per call")
for chains in (512, 1024):
    for epi in (0, 1):
        set_env(B2_TC_EPI=epi)
        a = time_hook(chains)
        set_env(B2_TC_EPI=epi, B2_TC_HOOK_FUSED=1, B2_TC_STAGES=5)
        b = time_hook(chains)
        set_env(B2_TC_EPI=epi, B2_TC_HOOK_FUSED=1, B2_TC_STAGES=4)
        c = time_hook(chains)
        print("chains %4d epi %d: two-kernel build %7.1f | fused build, 5 stages %7.1f | 4 stages %7.1f" % (chains, epi, a, b, c))

print("== (2) lock-step job, 1024 chains, 40 iterations from a jittered start")
opts = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75, t0=10.0,
            adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=0, exec_mode=_capi.B2_EXEC_LOCKSTEP,
            glm_path=_capi.B2_GLM_TCGEN05)
lib = _capi.load_library()
lib.b2_debug_tc_role_clocks.argtypes = [C.c_void_p, C.c_void_p]
for label, env in [("two-kernel", dict(B2_TC_FUSED=0)),
                   ("fused 5 stages, smem cap 227 KB", dict(B2_TC_STAGES=5)),
                   ("fused 4 stages, smem cap 227 KB", dict(B2_TC_STAGES=4)),
                   ("fused 4 stages, smem cap 194 KB", dict(B2_TC_STAGES=4, B2_TC_SMEM_CAP=194 * 1024)),
                   ("fused 4 stages, smem cap 162 KB", dict(B2_TC_STAGES=4, B2_TC_SMEM_CAP=162 * 1024))]:
    set_env(B2_TC_ROLE_CLOCKS=1, **env)
    C_, D = 1024, 101
    eng = model.engine(C_, dtype="float32")
    eng.set_state(bench.start_points(D, C_, 0), bench.chain_seeds(C_, 0), 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
    trace = eng.alloc_trace(_capi.B2_NUTS, 40)
    eng.run(_capi.B2_NUTS, 10, 40, opts, out=trace, row0=0)
    torch.cuda.synchronize()
    g0 = sum(r.n_grad for r in eng.reports())
    l0 = eng.kernel_launches()
    t0 = time.perf_counter()
    eng.run(_capi.B2_NUTS, 30, 40, opts, out=trace, row0=10)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    g1 = sum(r.n_grad for r in eng.reports())
    l1 = eng.kernel_launches()
    line = "%-34s %.3f M grad-evals/s, %5.1f us per launch, %4.0f chain-grads per launch" % (
        label, (g1 - g0) / dt / 1e6, dt * 1e6 / (l1 - l0), (g1 - g0) / (l1 - l0))
    buf = np.zeros(6, dtype=np.int64)
    rc = lib.b2_debug_tc_role_clocks(eng.handle, buf.ctypes.data_as(C.c_void_p))
    if rc == 0 and buf[1] > 0 and buf[4] > 0:
        line += " | cycles mean / max: likelihood CTAs %6.0f / %6.0f, state-machine warps %6.0f / %6.0f" % (
            buf[3] / buf[4], buf[5], buf[0] / buf[1], buf[2])
    print(line)
    eng.close()
