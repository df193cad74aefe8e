"""Multi-process CPU sampling with the oracle, driven like the reference's `_mp_sample`.

TEST / BASELINE INFRASTRUCTURE.  One OS process per chain, exactly like
pymc3/parallel_sampling.py:353-445 (ParallelSampler) -- each child owns its step object
(per-chain adaptation) and returns one draw per message; OMP/BLAS threads are pinned to 1 per
process.  Used by bench.py's `cpu_baseline` leg and `--impl reference`.
"""
import multiprocessing as mp
import os
import time

import numpy as np


def _worker(conn, factory, q0, seed, tune, sampler_kind, sampler_kwargs):
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    from oracle.hmc_cpu import CpuHMC, CpuNUTS
    from oracle.potentials import DiagAdaptPotential
    from oracle.rng import PhiloxRNG
    model = factory()
    D = model.ndim
    pot = DiagAdaptPotential(D, np.zeros(D), np.ones(D), 10)
    cls = CpuNUTS if sampler_kind == "nuts" else CpuHMC
    sampler = cls(model, D, pot, PhiloxRNG(seed), **sampler_kwargs)
    sampler.tune = tune > 0
    q = np.array(q0, dtype="d")
    it = 0
    conn.send(("ready", None))
    while True:
        msg = conn.recv()
        if msg[0] == "stop":
            break
        n = msg[1]
        rows, stats = [], []
        g0 = sampler.integ.n_grad
        for _ in range(n):
            if it == tune:
                sampler.tune = False
            q, st = sampler.transition(q)
            rows.append(q.copy())
            stats.append(st)
            it += 1
        conn.send(("draws", (np.array(rows), stats, sampler.integ.n_grad - g0)))
    conn.close()


class CpuChains:
    """`n_chains` oracle chains in `n_chains` processes (cores = n_chains)."""

    def __init__(self, factory, q0, seeds, tune, kind="nuts", **sampler_kwargs):
        ctx = mp.get_context("fork")
        self.conns, self.procs = [], []
        for c in range(len(q0)):
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(child, factory, q0[c], int(seeds[c]), tune, kind, sampler_kwargs),
                            daemon=True)
            p.start()
            self.conns.append(parent)
            self.procs.append(p)
        for conn in self.conns:
            assert conn.recv()[0] == "ready"

    def advance(self, n_iters):
        """Every chain does n_iters transitions; returns (q [C, n, D], stats, grad evals, seconds)."""
        t0 = time.perf_counter()
        for conn in self.conns:
            conn.send(("go", n_iters))
        outs = [conn.recv()[1] for conn in self.conns]
        dt = time.perf_counter() - t0
        q = np.stack([o[0] for o in outs])
        stats = [o[1] for o in outs]
        n_grad = int(sum(o[2] for o in outs))
        return q, stats, n_grad, dt

    def close(self):
        for conn in self.conns:
            try:
                conn.send(("stop",))
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()
