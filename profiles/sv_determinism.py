"""Is the block-per-chain persistent kernel reproducible?  Same chains, same seeds, twice -> the traces must be equal bit for bit."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
sys.argv = ["x"]
import bench
from pymc3_b200 import _capi
import pymc3_b200 as pm

C, N = 160, 60
model = pm.StochVol()
D = model.ndim
opts = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75, t0=10.0,
            adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=0, exec_mode=_capi.B2_EXEC_AUTO, glm_path=0)
tp = model.dict_to_array(model.test_point)
q0 = np.stack([tp + np.random.default_rng([7, c]).uniform(-1, 1, size=D) for c in range(C)])


def once():
    eng = model.engine(C, dtype="float32")
    eng.set_state(q0, bench.chain_seeds(C, 0), 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
    trace = eng.alloc_trace(_capi.B2_NUTS, N)
    eng.run(_capi.B2_NUTS, N, 1000, opts, out=trace, row0=0)
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy().copy() for k, v in trace.items()}
    eng.close()
    return out


for cfg in (dict(B2_PBLOCK_CTAS="1"), dict(B2_PBLOCK_CTAS="2"), dict(B2_PBLOCK_CTAS="1", B2_PBLOCK_NT="256")):
    for k in ("B2_PBLOCK_CTAS", "B2_PBLOCK_NT"):
        os.environ.pop(k, None)
    os.environ.update(cfg)
    a, b = once(), once()
    bad_rows = np.nonzero((a["q"] != b["q"]).any(axis=2))
    first = int(bad_rows[0].min()) if len(bad_rows[0]) else -1
    chains = sorted(set(bad_rows[1].tolist()))
    print(cfg, "identical" if first < 0 else "DIFFER from row %d on, %d chains: %s; tree_size equal: %s" % (
        first, len(chains), chains[:8], bool((a["tree_size"] == b["tree_size"]).all())))


def chunked():
    eng = model.engine(C, dtype="float32")
    eng.set_state(q0, bench.chain_seeds(C, 0), 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
    trace = eng.alloc_trace(_capi.B2_NUTS, N)
    for r in range(0, N, 20):
        eng.run(_capi.B2_NUTS, 20, 1000, opts, out=trace, row0=r)
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy().copy() for k, v in trace.items()}
    eng.close()
    return out


for k in ("B2_PBLOCK_CTAS", "B2_PBLOCK_NT"):
    os.environ.pop(k, None)
a, b, c = once(), chunked(), chunked()
print("chunked vs chunked:", "identical" if (b["q"] == c["q"]).all() else "DIFFER")
print("chunked vs one launch:", "identical" if (a["q"] == b["q"]).all() else "DIFFER from row %d" % int(np.nonzero((a["q"] != b["q"]).any(axis=2))[0].min()))

import pymc3_b200 as pm2
starts = bench.start_dicts(model, 64, 0)
seeds = [int(x) for x in bench.chain_seeds(64, 0)]
res = []
for rep in range(2):
    with model:
        step = pm2.NUTS(target_accept=0.8)
        tr = pm2.sample(40, tune=40, chains=64, step=step, start=starts, random_seed=seeds, progressbar=False,
                        compute_convergence_checks=False, discard_tuned_samples=False)
    res.append(np.stack([tr.get_values(v, combine=False) for v in ["step_size_log__" if "step_size_log__" in tr.varnames else tr.varnames[0]]]))
print("pm.sample twice:", "identical" if (res[0] == res[1]).all() else "DIFFER")
