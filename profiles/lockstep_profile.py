"""Per-launch durations of the C2 lock-step step (likelihood kernel / state-machine kernel) over one fixed job,
measured with the engine's CUDA events.  Runs unchanged against round 1's tree (A/B on the same box)."""
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
C, D = 1024, 101
opts = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75, t0=10.0,
            adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=0, exec_mode=_capi.B2_EXEC_AUTO, glm_path=0)
eng = model.engine(C, dtype="float32")
eng.set_state(bench.start_points(D, C, 0), bench.chain_seeds(C, 0), 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
trace = eng.alloc_trace(_capi.B2_NUTS, 1000)
for s in range(3):
    eng.run(_capi.B2_NUTS, 100, 500, opts, out=trace, row0=s * 100)
torch.cuda.synchronize()
for prof in (False, True):
    eng.set_profiling(prof)
    g0 = sum(r.n_grad for r in eng.reports())
    l0 = eng.kernel_launches()
    t0 = time.perf_counter()
    base = 300 + (350 if prof else 0)
    for s in range(3):
        eng.run(_capi.B2_NUTS, 100, 500, opts, out=trace, row0=base + s * 100)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    g = sum(r.n_grad for r in eng.reports()) - g0
    line = "profiling=%d: %.3f M grad-evals/s over 300 transitions (%d launches)" % (prof, g / dt / 1e6, eng.kernel_launches() - l0)
    if prof:
        ms, n = eng.profile()
        adv = eng.profile_advance()
        line += " | likelihood %.1f us, state machine %.1f us per step, %.0f chain-grads per step, step %.1f us" % (
            ms * 1e3 / n, adv * 1e3 / n, g / n, dt * 1e6 / n)
    print(line)
eng.close()
