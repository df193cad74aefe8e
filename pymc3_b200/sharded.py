"""Observation-sharded sampling (BASELINE.json config 5): the one configuration with a
data-path collective.

Every rank holds a contiguous row shard of (X, y) and runs EVERY chain's NUTS state machine
redundantly; per leapfrog the ranks exchange one packed buffer `[C, D+1]` (partial logp, partial dlogp)
with an all-reduce(sum).  The all-reduce returns bit-identical values on every rank, so the replicated
tree decisions stay in lock-step without any further synchronisation.  The reference has no such path
(it never shards a likelihood, SURVEY section 5 "long-context" row); the seam is the same
`ValueGradFunction.__call__` (model.py:645-666), whose value becomes a sum over ranks.

The collective is injected (`allreduce(tensor)`, in place): `torch.distributed.all_reduce` over NCCL on
the GPUs; the CPU test tier drives the same loop over `gloo` with a stand-in engine.
"""
import ctypes as C

import numpy as np

from . import _capi


def row_shard(n_rows, world_size, rank):
    """Contiguous, balanced [lo, hi) row range of `rank`."""
    bounds = np.linspace(0, n_rows, world_size + 1).astype(np.int64)
    return int(bounds[rank]), int(bounds[rank + 1])


def run_lockstep_sharded(engine, kind, n_iters, tune_until, opts, allreduce, world_size, check_every=16,
                         out=None, row0=0, profile=False):
    """Drives `n_iters` transitions of all chains of `engine` (this rank's shard) in lock-step with the
    other ranks.  Returns the device trace dict (identical on every rank).  `profile`: CUDA events around the
    three parts of the first 256 leapfrogs; their mean durations land in `engine.last_lockstep_profile`
    ({"likelihood_us", "allreduce_us", "advance_us", "allreduce_share"})."""
    torch = engine.torch
    lib = engine.lib
    if out is None:
        out = engine.alloc_trace(kind, n_iters)
        view = out
    else:
        view = {k: v[row0:row0 + n_iters] for k, v in out.items()}
    tr = _capi.TraceOut()
    for name, t in view.items():
        setattr(tr, "d_" + name, t.data_ptr())
    o = _capi.SamplerOpts(kind=kind, n_iters=int(n_iters), tune_until=int(tune_until), **opts)
    packed = torch.zeros((engine.n_chains, engine.D + 1), dtype=torch.float64, device=engine.dev)
    stream = engine._stream()
    _capi.check(lib.b2_step_begin(engine.handle, C.byref(o), C.byref(tr), stream), lib)
    active = C.c_int32(1)
    steps = 0
    events = []
    while True:
        for _ in range(check_every):
            timed = profile and len(events) < 256
            if timed:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                ev[0].record()
            _capi.check(lib.b2_step_likelihood(engine.handle, packed.data_ptr(), stream), lib)
            if timed:
                ev[1].record()
            if world_size > 1:
                allreduce(packed)
            if timed:
                ev[2].record()
            _capi.check(lib.b2_step_advance(engine.handle, packed.data_ptr(), int(world_size), stream), lib)
            if timed:
                ev[3].record()
                events.append(ev)
            steps += 1
        _capi.check(lib.b2_step_active(engine.handle, C.byref(active), stream), lib)
        if active.value == 0:
            break
    _capi.check(lib.b2_step_end(engine.handle), lib)
    engine.iter_done += int(n_iters)
    engine.last_lockstep_steps = steps
    engine.last_lockstep_profile = None
    if events:
        torch.cuda.synchronize(engine.dev)
        parts = np.array([[e[i].elapsed_time(e[i + 1]) * 1e3 for i in range(3)] for e in events]).mean(axis=0)
        engine.last_lockstep_profile = {"likelihood_us": float(parts[0]), "allreduce_us": float(parts[1]),
                                        "advance_us": float(parts[2]), "allreduce_share": float(parts[1] / parts.sum()),
                                        "leapfrogs_timed": len(events)}
    return view


def sample_glm_sharded(X_shard, y_shard, n_chains, draws, tune, seeds, start, step_size0=None, dtype="float32",
                       device=0, opts=None, prior_tau=1e-6):
    """Convenience wrapper used by bench.py / the multi-GPU test: NUTS for the Bernoulli-logit GLM with the
    rows sharded over the ranks of the default torch.distributed process group."""
    import torch.distributed as dist
    from .model import LogisticGLM
    world = dist.get_world_size() if dist.is_initialized() else 1
    model = LogisticGLM(X_shard, y_shard, prior_tau=prior_tau)
    eng = model.engine(n_chains, dtype=dtype, device=device)
    D = eng.D
    eng.set_state(start, seeds, step_size0 or 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
    o = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75, t0=10.0,
             adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=0,
             exec_mode=_capi.B2_EXEC_LOCKSTEP, glm_path=_capi.B2_GLM_AUTO)
    o.update(opts or {})
    trace = run_lockstep_sharded(eng, _capi.B2_NUTS, draws + tune, tune, o,
                                 (lambda t: dist.all_reduce(t)) if world > 1 else None, world)
    return eng, trace
