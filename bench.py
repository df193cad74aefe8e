#!/usr/bin/env python
"""Benchmark of the NUTS hot path (contract: see the task statement; layout: DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2]

Metric (BASELINE.json): leapfrog gradient evaluations / second, whole job (all chains, all GPUs),
with min bulk ESS / second reported beside it.  Headline workload (config 2): Bernoulli-logit GLM
N=100 000, D=100 (+intercept), 1024 chains per GPU, synthetic data of SURVEY 8(d).  With the default
`--workload all` the line also carries `configs`: shortened but complete jobs (tune + draws, ESS,
roofline, e2e) of config 1 (eight schools, with the CPU arm's ESS/s beside it), config 3 (radon NCP,
N = 1 M, 4096 chains split over the GPUs), config 4 (stochastic volatility, 512 chains split over the
GPUs), config 5 (logistic regression, rows sharded over the GPUs, all-reduce per leapfrog) and `c2s`
(a reduced config 2 that the CPU arm can finish: the ESS/s ratio on the same job).

A "step" is `--iters-per-step` (100) NUTS transitions of every chain, continuing one sampling job:
the first half of the (W+K)*100 iterations tunes (dual averaging + diagonal mass adaptation), the
second half draws.  With the defaults W=3, K=17 this is exactly tune=1000 / draws=1000.

  value  grad-evals/s over the K timed steps, data and state resident in HBM (CUDA events)
  e2e    same metric for the same job run through the public API, pymc3_b200.sample(): model upload,
         sampling, trace download (chunk by chunk, overlapped), MultiTrace construction
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md section 6

`--impl reference` times the CPU implementation of the same path (the oracle port of the
reference's step methods, one process per chain like pymc3/parallel_sampling.py) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------- workloads
def glm_synthetic(n=100000, k=100, seed=20200420):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, k), dtype=np.float32)
    beta = rng.normal(0.0, 0.5, size=k)
    eta = 0.3 + X.astype(np.float64) @ beta
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float32)
    return X, y


def hier_synthetic(n=1000000, seed=3):
    from pymc3_b200.model import radon_county_counts
    info = radon_county_counts()
    w = np.asarray(info["counts"], dtype="f8")
    rng = np.random.default_rng(seed)
    g = len(w)
    idx = rng.choice(g, size=n, p=w / w.sum())
    floor = (rng.random(n) < 0.17).astype(np.uint8)
    a = 1.5 + 0.3 * rng.standard_normal(g)
    b = -0.7 + 0.3 * rng.standard_normal(g)
    y = (a[idx] + b[idx] * floor + 0.7 * rng.standard_normal(n)).astype(np.float32)
    return idx, floor, y, g


def make_workload(name, args):
    """-> dict(model, oracle_factory, chains, label, flops_per_chain_grad, bytes_per_chain_grad, bound)"""
    import pymc3_b200 as pm
    if name == "c2":
        n, k = args.n_obs or 100000, args.n_features or 100
        X, y = glm_synthetic(n, k)

        def oracle_factory():
            from oracle.densities import LogisticGLM
            return LogisticGLM(X, y)
        return dict(model=pm.LogisticGLM(X, y), oracle_factory=oracle_factory, chains=args.chains or 1024,
                    label="Bernoulli-logit GLM N=%d D=%d (+intercept)" % (n, k),
                    flops_per_chain_grad=4.0 * n * (k + 1), bytes_per_chain_grad=None, bound="tensor")
    if name == "c1":
        def oracle_factory():
            from oracle.densities import EightSchoolsNCP
            return EightSchoolsNCP()
        return dict(model=pm.EightSchoolsNCP(), oracle_factory=oracle_factory, chains=args.chains or 4,
                    label="eight-schools NCP", flops_per_chain_grad=None, bytes_per_chain_grad=None, bound=None)
    if name == "c3":
        n = args.n_obs or 1000000
        idx, floor, y, g = hier_synthetic(n)

        def oracle_factory():
            from oracle.densities import HierLinearNCP
            return HierLinearNCP(idx, floor, y, g)
        return dict(model=pm.HierLinearNCP(idx, floor, y, g), oracle_factory=oracle_factory,
                    chains=args.chains or 4096, label="radon-style hierarchical NCP, 85 groups N=%d" % n,
                    flops_per_chain_grad=12.0 * n, bytes_per_chain_grad=6.0 * n / 128, bound="hbm")
    if name == "c4":
        def oracle_factory():
            from oracle.densities import StochVol
            from pymc3_b200.model import sp500_log_returns
            return StochVol(sp500_log_returns())
        return dict(model=pm.StochVol(), oracle_factory=oracle_factory, chains=args.chains or 512,
                    label="stochastic volatility T=2905", flops_per_chain_grad=40.0 * 2905,
                    bytes_per_chain_grad=6.0 * 2905 * 4, bound="hbm")
    raise SystemExit("unknown workload %r" % name)


# timing rule: say how the timed iterations avoid measuring a warm-L2 replay of identical work
L2_NOTES = {
    "c1": "one persistent launch per step; state lives in shared memory, nothing to replay",
    "c2": "pre-tiled X (51 MB bf16 hi/lo) is re-read from HBM every leapfrog (ncu: 52 MB DRAM traffic per launch); "
          "positions, gradients and residuals change every launch",
    "c3": "observations (6 MB) stay L2-resident by design; the chains' positions change every launch",
    "c4": "one persistent launch per step; hot state in shared memory, tree stack (0.5 GB for 512 chains) > L2",
}


def start_points(ndim, chains, first_chain):
    """test point + U(-1, 1), keyed by global chain id (sampling.py:1920-1926)."""
    out = np.empty((chains, ndim))
    for c in range(chains):
        out[c] = np.random.default_rng([7, first_chain + c]).uniform(-1, 1, size=ndim)
    return out


def chain_seeds(chains, first_chain):
    return np.array([0x5EED0000 + first_chain + c for c in range(chains)], dtype=np.uint64)


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.device = device

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json; bf16 sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------- CPU arms
def cpu_reference_run(wl, steps, warmup, budget_s, cores, iters_per_step=1):
    """The reference's CPU path (oracle port), one process per chain on `cores` host cores.
    A step = `iters_per_step` NUTS transitions of each of the `cores` chains (tuning phase)."""
    from oracle.cpu_sampler import CpuChains
    ndim = wl["model"].ndim
    q0 = start_points(ndim, cores, 0)
    chains = CpuChains(wl["oracle_factory"], q0, chain_seeds(cores, 0), tune=10 ** 9)
    leap_t, dt_t, done = 0, 0.0, 0
    t_begin = time.perf_counter()
    try:
        for s in range(warmup + steps):
            _, stats, _, dt = chains.advance(iters_per_step)
            if s >= warmup:
                leap_t += int(sum(st["tree_size"] for per_chain in stats for st in per_chain))
                dt_t += dt
                done += 1
            if time.perf_counter() - t_begin > budget_s and done >= 1:
                break
    finally:
        chains.close()
    return leap_t / dt_t, done, dt_t


# ------------------------------------------------------------------------------- GPU arm
class Ctx:
    """rank / device / collectives of this process (one process per GPU)"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        self.dev = torch.device("cuda", self.local_rank)
        torch.cuda.set_device(self.dev)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def reduce(self, seconds, counts):
        """max over ranks of `seconds`, sum over ranks of `counts` (pymc3_b200/distributed.py)"""
        from pymc3_b200 import distributed as b2d
        return b2d.reduce_job_metrics(seconds, counts, device=self.dev)

    def min_ess(self, ess_vec):
        from pymc3_b200 import distributed as b2d
        return b2d.combine_ess(ess_vec, device=self.dev)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def nuts_opts(args, exec_mode=0):
    return dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75,
                t0=10.0, adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=0,
                exec_mode=exec_mode, glm_path={"auto": 0, "group": 1, "simt": 2, "tcgen05": 3}[args.glm_path])


def device_ess(q_draws, model=None, max_cols=512):
    """min-bulk-ESS input: rank-normalised split bulk ESS of every free scalar, computed where the trace lives
    (pymc3_b200/stats_device.py, pinned to the NumPy estimator).  Bulk ESS is rank based, so the monotone
    back-transforms (exp of the log-scale variables) have the same ESS as the free variables.
    q_draws: [draws, C, D] device tensor -> (ess [K] numpy, note)"""
    import torch
    from pymc3_b200 import stats_device
    Dq = q_draws.shape[2]
    note = "every scalar of every free variable (= of every back-transformed one: bulk ESS is rank based)"
    cols = None
    if Dq > max_cols:
        cols = np.unique(np.concatenate([np.arange(8), np.arange(Dq - 8, Dq), np.linspace(8, Dq - 9, 112).astype(int)]))
        q_draws = q_draws[:, :, torch.as_tensor(cols, device=q_draws.device)]
        note = "%d of %d free scalars (first/last 8 + 112 evenly spaced)" % (len(cols), Dq)
    out = []
    for lo in range(0, q_draws.shape[2], 32):                       # 32 columns at a time: bounded workspace
        x = stats_device.from_trace(q_draws[:, :, lo:lo + 32])
        out.append(stats_device.ess_bulk(x).cpu().numpy())
    return np.concatenate(out), note


def start_dicts(model, chains, first_chain):
    """jittered start points as pm.sample takes them: test point + U(-1, 1) (sampling.py:1920-1926), keyed by
    global chain id"""
    tp = model.dict_to_array(model.test_point)
    out = []
    for c in range(chains):
        q = tp + np.random.default_rng([7, first_chain + c]).uniform(-1, 1, size=len(tp))
        out.append(model.array_to_dict(q))
    return out


def sample_config(ctx, args, wl, chains_total, tune, draws, split="strong", profile_chunks=0, chunk=100, run_ahead=True):
    """One complete NUTS job through the PUBLIC API (pymc3_b200.sample) on this rank's share of the chains.
    Returns the per-config result object (rank 0) -- value from the CUDA-event time of the sampling launches inside
    the call, e2e from the wall clock around the whole call (model upload, sampling, trace download, MultiTrace)."""
    import pymc3_b200 as pm
    from pymc3_b200 import distributed as b2d
    torch = ctx.torch
    model = wl["model"]
    if split == "strong":
        lo, hi = b2d.shard_chains(chains_total, ctx.world, ctx.rank)
    else:                                                     # weak: every rank runs chains_total chains of its own
        lo, hi = ctx.rank * chains_total, (ctx.rank + 1) * chains_total
    chains = hi - lo
    starts = start_dicts(model, chains, lo)
    seeds = [int(x) for x in chain_seeds(chains, lo)]
    ctx.barrier()
    t0 = time.perf_counter()
    with model:
        step = pm.NUTS(device=ctx.local_rank, dtype=args.dtype, glm_path=args.glm_path)
        step._profile_last_chunks = profile_chunks
        step._log_chunk_grads = True          # leapfrog counters after every chunk (for e2e over the timed steps)
        trace = pm.sample(draws, tune=tune, chains=chains, step=step, start=starts, random_seed=seeds, chunk=chunk,
                          run_ahead=run_ahead, discard_tuned_samples=False, compute_convergence_checks=False, progressbar=False)
    torch.cuda.synchronize(ctx.dev)
    wall = time.perf_counter() - t0
    ctx.barrier()
    n_grad = step._last_n_grad
    reports = step._last_reports
    failed = sum(1 for r in reports if r.phase != 3)
    # min bulk ESS over post-tune draws, from the host trace pushed back to the device in slabs (cheap next to the job)
    ess_vec, ess_note = None, None
    if draws >= 100:
        names = model.free_RVs
        cols = [np.stack(trace.get_values(n, burn=tune, combine=False)).reshape(chains, draws, -1) for n in names]
        q = np.concatenate(cols, axis=2)                      # [C, draws, D]
        qd = torch.as_tensor(np.ascontiguousarray(np.swapaxes(q, 0, 1)), device=ctx.dev)
        ess_vec, ess_note = device_ess(qd)
        del qd, q, cols
    stats_tree = trace.get_sampler_stats("tree_size")
    depth = trace.get_sampler_stats("depth")
    secs, counts = ctx.reduce([step._last_device_seconds, wall], [n_grad, step._last_kernel_launches, failed, chains])
    res = {"value": counts[0] / secs[0], "unit": "grad-evals/s", "chains_total": int(counts[3]), "chains_this_rank": chains,
           "tune": tune, "draws": draws, "job_seconds_device": secs[0], "job_seconds_wall": secs[1],
           "e2e": {"value": counts[0] / secs[1], "unit": "grad-evals/s",
                   "through": "pymc3_b200.sample(): model upload, sampling, trace download into pinned staging, MultiTrace"},
           "gpu_launches": int(counts[1]), "failed_chains": int(counts[2]),
           "mean_tree_size": float(stats_tree.mean()), "mean_depth_post_tune": float(depth.reshape(chains, -1)[:, tune:].mean()),
           "scaling": split}
    if ess_vec is not None:
        mn, _ = ctx.min_ess(ess_vec)
        res["ess"] = {"min_bulk_ess": mn, "n_scalars": int(len(ess_vec)), "over": ess_note}
        res["min_bulk_ess_per_sec"] = mn / secs[1]           # whole job incl. tuning, wall clock (benchmarks.py:163-169)
    res["chunk_log"] = getattr(step, "_last_chunk_log", None)       # (rows done, seconds since the call began, device seconds)
    res["host_phases_s"] = {k: round(v, 4) for k, v in getattr(step, "_last_timing", {}).items()}
    res["host_phases_s"]["step_and_sample_call"] = round(wall, 4)
    res["host_phases_s"]["finish_detail"] = {k: round(v, 4) for k, v in (getattr(step, "_last_finish_phases", None) or {}).items()}
    res["_profile"] = step._last_profile
    del trace
    return res


def strip_logs(res):
    """drop the per-chunk bookkeeping that only the headline's e2e arithmetic needs"""
    res.pop("chunk_log", None)
    return res


def lockstep_roofline(prof, wl, peaks, bound):
    """roofline object of a lock-step workload from the launches timed inside the job's last chunks"""
    if not prof or not prof.get("like_n"):
        return None
    per_launch_s = prof["like_ms"] / 1e3 / prof["like_n"]
    units = prof["n_grad"] / prof["like_n"]                  # chain-gradients one likelihood launch processed
    out = {"kernel": "chain-batched likelihood (logp+dlogp, all live chains)", "launches_timed": int(prof["like_n"]),
           "chain_grads_per_launch": units, "avg_launch_us": per_launch_s * 1e6,
           "kernel_share_of_slice": prof["like_ms"] / 1e3 / prof["seconds"],
           "advance_kernel_avg_us": prof["adv_ms"] * 1e3 / prof["like_n"],
           "advance_share_of_slice": prof["adv_ms"] / 1e3 / prof["seconds"],
           "timed_in": "the last chunks of the job, CUDA events around every launch (adds ~7 % to those chunks)",
           "peak_source": peaks["source"], "traffic": None}
    if bound == "tensor":
        ach = wl["flops_per_chain_grad"] * units / per_launch_s / 1e12
        out.update({"bound": "tensor", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": ach / peaks["tflops"]})
    else:
        ach = wl["bytes_per_chain_grad"] * units / per_launch_s / 1e9
        out.update({"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]})
        if wl.get("flops_per_chain_grad"):
            # once a tile of observations is shared by 128 chains the bound is FP32 issue, not HBM
            fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12          # nominal: SMs x FP32 lanes x 2 x max SM clock
            fl = wl["flops_per_chain_grad"] * units / per_launch_s / 1e12
            out.update({"fp32_achieved_tflops": fl, "fp32_peak_tflops_nominal": fp32_peak, "fp32_frac": fl / fp32_peak})
    return out


def run_c2_headline(ctx, args):
    """Headline (BASELINE.json config 2): value from an engine-level job with everything resident in HBM, e2e from
    the same job through pymc3_b200.sample, roofline from a short profiled slice that continues the first job."""
    import torch
    from pymc3_b200 import _capi
    wl = make_workload("c2", args)
    model, chains = wl["model"], wl["chains"]          # chains per GPU (weak scaling)
    first_chain = ctx.rank * chains
    ndim = model.ndim
    ips, K, W = args.iters_per_step, args.steps, args.warmup
    total = (K + W) * ips
    tune = total // 2
    opts = nuts_opts(args)
    q0 = start_points(ndim, chains, first_chain)
    seeds = chain_seeds(chains, first_chain)
    dev = ctx.dev

    eng = model.engine(chains, dtype=args.dtype, device=ctx.local_rank)
    eng.set_state(q0, seeds, 0.25 / ndim ** 0.25, np.zeros(ndim), np.ones(ndim), 10.0)
    trace = eng.alloc_trace(_capi.B2_NUTS, total)          # the whole job's device trace, allocated up front
    ctx.barrier()
    tw0 = time.perf_counter()
    ahead = args.run_ahead

    def grads(e):
        return sum(r.n_grad for r in e.reports())

    for s in range(W):
        eng.run(_capi.B2_NUTS, ips, tune, opts, out=trace, row0=s * ips, run_ahead=ahead)
    torch.cuda.synchronize(dev)
    wall_warm = time.perf_counter() - tw0
    clocks = ClockSampler(ctx.local_rank)
    ctx.barrier()
    if ctx.rank == 0 and not args.no_clocks:
        clocks.start()
    eng.set_profiling(False)
    launches0 = eng.kernel_launches()
    grads0 = grads(eng)                      # leapfrogs are counted when they are executed (engine counters)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for s in range(K):
        eng.run(_capi.B2_NUTS, ips, tune, opts, out=trace, row0=(W + s) * ips, run_ahead=ahead)
    ev1.record()
    ctx.barrier()
    wall_timed = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clock_info = clocks.stop() if ctx.rank == 0 else None
    launches = eng.kernel_launches() - launches0
    reports = eng.reports()
    leap_timed = sum(r.n_grad for r in reports) - grads0
    failed = sum(1 for r in reports if r.phase != _capi.PHASE_DONE)

    # ---- profiled slice: the same chains go on for a few more (post-tuning) steps with CUDA events around every
    #      likelihood / advance launch; the roofline's kernel duration and the chain-gradients per launch come from here
    prof = None
    if not args.no_profile and ctx.rank == 0:
        P = 2
        scratch = eng.alloc_trace(_capi.B2_NUTS, P * ips)
        eng.set_profiling(True)
        pg0 = grads(eng)
        tp0 = time.perf_counter()
        for s in range(P):
            eng.run(_capi.B2_NUTS, ips, tune, opts, out=scratch, row0=s * ips, run_ahead=ahead)
        torch.cuda.synchronize(dev)
        like_ms, like_n = eng.profile()
        prof = {"seconds": time.perf_counter() - tp0, "n_grad": grads(eng) - pg0, "like_ms": like_ms, "like_n": like_n,
                "adv_ms": eng.profile_advance()}
        del scratch
    eng.close()

    secs, cnt = ctx.reduce([dev_ms / 1e3, wall_timed, wall_warm], [leap_timed, launches, failed])
    t_timed = secs[0]
    value = cnt[0] / t_timed
    ess_info = None
    if not args.skip_ess and total - tune >= 100:
        ess_vec, ess_note = device_ess(trace["q"][tune:])
        mn, _ = ctx.min_ess(ess_vec)
        ess_info = {"min_bulk_ess": mn, "n_scalars": int(len(ess_vec)), "over": ess_note}
    del trace                                # its blocks stay in torch's caching allocator: the e2e job below asks for the same
                                             # sizes and reuses them (a fresh 1 GB cudaMalloc cost 0.02-0.34 s depending on the box)

    # ---- e2e: the same job through the call a user makes
    e2e = None
    if not args.skip_e2e:
        r = sample_config(ctx, args, wl, chains, tune, total - tune, split="weak", chunk=ips, run_ahead=ahead)
        x_bytes = sum(t.numel() * t.element_size() for t in [torch.as_tensor(model.X), torch.as_tensor(model.y)])
        d2h = (chains * ndim * 4 + chains * (7 * 8 + 2 * 4 + 2)) * ips          # one step's rows of q + 11 stats
        # Like `value`, the headline e2e number is over the K timed steps (the W warm-up steps, early tuning with its
        # deep trees, are excluded from both): leapfrogs of those steps / (their wall time inside sample() + the
        # call's fixed host work -- start points, engine build and upload, final copies, MultiTrace -- pro rata).
        log = r.pop("chunk_log")
        t_w, g_w = [(t, g) for rows_, t, _, g in log if rows_ == W * ips][0]
        chunks_begin = r["host_phases_s"]["start_array"] + r["host_phases_s"]["engines"]
        fixed = r["job_seconds_wall"] - (log[-1][1] - chunks_begin)          # everything outside the chunk loop
        wall_timed_e2e = (log[-1][1] - t_w) + fixed * K / (K + W)
        lf_timed = float(log[-1][3] - g_w)        # leapfrogs executed after the warm-up chunks (engine counters, as for `value`)
        sv, cv = ctx.reduce([wall_timed_e2e], [lf_timed])
        e2e = {"value": cv[0] / sv[0], "whole_call_value": r["e2e"]["value"], "unit": "grad-evals/s",
               "h2d_bytes_per_step": int(x_bytes // (K + W)),
               "d2h_bytes_per_step": int(d2h), "job_seconds_wall": r["job_seconds_wall"],
               "job_seconds_device": r["job_seconds_device"], "through": r["e2e"]["through"],
               "host_phases_s": r.get("host_phases_s"),
               "chunk_log": [[int(a), round(b, 4), round(c, 4), int(d)] for a, b, c, d in log],
               "note": "value = the K timed steps of one pymc3_b200.sample() call for the whole job (whole_call_value includes the "
                       "warm-up steps): X, y are uploaded once per call "
                       "(h2d per step = that upload / steps), every chunk of %d transitions is copied to the host while "
                       "the next one samples; convergence checks off (the bench computes ESS itself)" % ips}
    if ctx.rank != 0:
        return None

    peaks = measured_peaks()
    roofline = lockstep_roofline(prof, wl, peaks, "tensor")
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if roofline is not None and os.path.exists(tpath):         # one ncu --set full capture, per launch (constant, not live)
        with open(tpath) as f:
            roofline["traffic"] = json.load(f).get("k_glm_tc_main", {}).get("dram_bytes_per_launch")
    cpu = None
    if ctx.world == 1 and not args.skip_cpu:
        cores = args.cpu_cores or os.cpu_count()
        v, done, dt = cpu_reference_run(wl, steps=10 ** 6, warmup=1, budget_s=args.cpu_budget, cores=cores)
        cpu = {"value": v, "unit": "grad-evals/s", "cores": cores, "kind": "port",
               "sample": "%d tuning transitions on each of %d chains (1 process/chain, full data), %.1f s" % (done, cores, dt)}
    job_s = t_timed + secs[2]
    return {
        "metric": "leapfrog_grad_evals_per_sec", "value": value, "unit": "grad-evals/s", "n_gpus": ctx.world,
        "steps": K, "warmup": W, "ms_per_step": t_timed * 1e3 / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "chains_per_gpu": chains, "chains_total": chains * ctx.world,
                   "iters_per_step": ips, "tune": tune, "draws": total - tune, "sampler": "NUTS target_accept=0.8",
                   "l2": L2_NOTES["c2"], "grad_evals_counted": "engine leapfrog counters read before and after the timed steps",
                   "step_boundaries": ("lock-step chains that finish a step early run ahead into the next step's rows"
                                       if ahead else "all chains stop at every step boundary")},
        "min_bulk_ess_per_sec": ess_info["min_bulk_ess"] / job_s if ess_info else None,   # whole job incl. tuning
        "job_seconds": job_s, "ess": ess_info, "e2e": e2e, "gpu_launches": int(cnt[1]), "failed_chains": int(cnt[2]),
        "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu,
    }


# shortened-but-complete jobs of the other BASELINE.json configs (tune + draws, ESS, roofline, e2e); under --gpus N
# config 3 and 4 split their chains over the ranks (strong scaling), config 5 shards its rows (weak in rows)
# (config 3: the first mass-matrix window is estimated from the transient of the jittered start and the second from a
#  chain that mixes poorly under that estimate, so -- like the reference -- the sampler needs ~1000 warm-up transitions
#  at tree depth 8-10 before its trees shrink: 4096 x 2000 x ~500 leapfrogs is a 15-minute job on one GPU.  The side
#  config is therefore a 200-transition job: throughput and roofline are representative, its ESS is that of warm-up.)
CONFIG_JOBS = {"c1": dict(chains=4, tune=500, draws=1000), "c3": dict(chains=4096, tune=160, draws=40),
               "c4": dict(chains=512, tune=150, draws=150)}


def run_config(ctx, args, name):
    peaks = measured_peaks()
    job = dict(CONFIG_JOBS[name])
    if name == "c1" and ctx.world > 1:
        return {"skipped": "4 chains, latency-bound by construction: run at N = 1 only"}
    wl = make_workload(name, args)
    res = sample_config(ctx, args, wl, job["chains"], job["tune"], job["draws"], split="strong",
                        profile_chunks=1 if name == "c3" else 0)
    prof = res.pop("_profile", None)
    strip_logs(res)
    res["config"] = {"workload": wl["label"], "sampler": "NUTS target_accept=0.8", "l2": L2_NOTES.get(name)}
    if name == "c3":
        res["roofline"] = lockstep_roofline(prof, wl, peaks, "hbm")
    if name == "c4":
        ach = wl["bytes_per_chain_grad"] * res["value"] / max(ctx.world, 1) / 1e9
        res["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                           "traffic": None, "kernel": "k_persistent_block (whole NUTS transition loop, one launch per chunk)",
                           "regime": "SURVEY 8d's 70 KB per chain-grad if the state streamed through HBM, set against the whole-job "
                                     "rate; the hot state is resident in shared memory, so the kernel is instruction-issue bound, not HBM bound",
                           "peak_source": peaks["source"]}
    if name == "c1" and ctx.world == 1 and not args.skip_cpu:
        res["cpu"] = cpu_ess_run(wl, chains=4, tune=job["tune"], draws=job["draws"])
        if res.get("min_bulk_ess_per_sec") and res["cpu"].get("min_bulk_ess_per_sec"):
            res["ess_per_sec_vs_cpu"] = res["min_bulk_ess_per_sec"] / res["cpu"]["min_bulk_ess_per_sec"]
    return res


def cpu_ess_run(wl, chains, tune, draws):
    """the CPU arm's complete job (oracle port, one process per chain): grad-evals/s and min bulk ESS/s incl. tuning"""
    from oracle.cpu_sampler import CpuChains
    from pymc3_b200 import stats as b2stats
    ndim = wl["model"].ndim
    tp = wl["model"].dict_to_array(wl["model"].test_point)
    q0 = np.stack([tp + np.random.default_rng([7, c]).uniform(-1, 1, size=ndim) for c in range(chains)])
    cc = CpuChains(wl["oracle_factory"], q0, chain_seeds(chains, 0), tune=tune)
    t0 = time.perf_counter()
    try:
        _, _, g1, _ = cc.advance(tune)
        q, _, g2, _ = cc.advance(draws)
    finally:
        cc.close()
    wall = time.perf_counter() - t0
    ess = float(np.min(b2stats.ess(q)))
    return {"kind": "port", "cores": chains, "chains": chains, "tune": tune, "draws": draws, "job_seconds": wall,
            "value": (g1 + g2) / wall, "unit": "grad-evals/s", "min_bulk_ess": ess, "min_bulk_ess_per_sec": ess / wall}


SMALL_ESS_JOBS = {"c2s": ("c2", dict(n_obs=20000, n_features=100, chains=1024)),       # chains: the configs' own counts
                  "c3s": ("c3", dict(n_obs=20000, chains=4096))}


def run_small_ess(ctx, args, name):
    """min-bulk-ESS/s of the GPU engine against the CPU arm ON THE SAME JOB: a reduced C2 (N = 20 000, D = 100) or a
    reduced C3 (N = 20 000 observations, 85 groups), 200 tune + 200 draws -- sizes the CPU port finishes in ~30 s on
    8 cores; the GPU runs the same job with the chain count of the full config (1024 / 4096)."""
    import argparse as _ap
    base, shape = SMALL_ESS_JOBS[name]
    a2 = _ap.Namespace(**vars(args))
    a2.n_obs, a2.n_features, a2.chains = shape.get("n_obs", 0), shape.get("n_features", 0), 0
    wl = make_workload(base, a2)
    gpu = sample_config(ctx, a2, wl, shape["chains"], 200, 200, split="strong")
    gpu.pop("_profile", None)
    strip_logs(gpu)
    out = {"workload": wl["label"], "gpu": {k: gpu[k] for k in ("value", "chains_total", "tune", "draws", "job_seconds_wall",
                                                            "min_bulk_ess_per_sec", "ess", "e2e")}}
    if ctx.world == 1 and not args.skip_cpu:
        out["cpu"] = cpu_ess_run(wl, chains=min(8, os.cpu_count() or 8), tune=200, draws=200)
        out["ess_per_sec_vs_cpu"] = gpu["min_bulk_ess_per_sec"] / out["cpu"]["min_bulk_ess_per_sec"]
        out["grad_evals_per_sec_vs_cpu"] = gpu["e2e"]["value"] / out["cpu"]["value"]
    return out


def run_b200(args):
    ctx = Ctx()
    line = None
    if args.workload in ("all", "c2"):
        line = run_c2_headline(ctx, args)
    elif args.workload == "c5":
        line = run_c5(ctx, args, rows=args.n_obs or 6250000, tune_draws=None)
    else:
        res = run_config(ctx, args, args.workload)
        if ctx.rank == 0:
            line = {"metric": "leapfrog_grad_evals_per_sec", "n_gpus": ctx.world, "steps": args.steps, "warmup": args.warmup,
                    "higher_is_better": True, "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64",
                    "data": "synthetic"}
            line.update(res)
    if args.workload == "all":
        configs = {}
        for name in [c for c in args.configs.split(",") if c]:
            t0 = time.perf_counter()
            try:
                if name == "c5":
                    # rows are weak-scaled, so the posterior narrows with the world size and the first transitions (far
                    # start, step size not yet adapted) get long: at >= 4 ranks the side job is half as many transitions
                    # (30 took 48 s at 8 ranks x 6.25 M rows); the per-leapfrog breakdown does not depend on the job length
                    n5 = args.c5_iters if ctx.world <= 2 else max(12, args.c5_iters // 2)
                    res = run_c5(ctx, args, rows=args.c5_rows, tune_draws=(n5 // 2, n5 - n5 // 2))
                elif name in SMALL_ESS_JOBS:
                    if ctx.world > 1:              # a CPU-ratio job: one GPU against the host cores, nothing to scale
                        res = {"skipped": "GPU-vs-CPU ESS/s ratio on a reduced job: run at N = 1 only"}
                    else:
                        res = run_small_ess(ctx, args, name)
                else:
                    res = run_config(ctx, args, name)
            except Exception as err:                         # a failing side config must not take the headline down
                if ctx.world > 1:
                    raise
                res = {"error": "%s: %s" % (type(err).__name__, str(err)[:300])}
            if isinstance(res, dict):
                res["bench_seconds"] = time.perf_counter() - t0
            configs[name] = res
            ctx.torch.cuda.empty_cache()
        if ctx.rank == 0:
            line["configs"] = configs
    if ctx.rank == 0 and line is not None:
        print(json.dumps(line))
    ctx.close()


def run_c5(ctx, args, rows, tune_draws):
    """Config 5: logistic regression with the OBSERVATIONS sharded over the GPUs (weak scaling in rows:
    6.25 M rows x 256 features per GPU = 50 M rows on 8), every rank runs all 256 chains redundantly and
    the ranks all-reduce the packed [C, D+1] (logp, dlogp) partials once per leapfrog over NCCL."""
    import torch
    import torch.distributed as dist
    from pymc3_b200 import _capi
    from pymc3_b200.model import LogisticGLM
    from pymc3_b200.sharded import run_lockstep_sharded
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    k = args.n_features or 256
    chains = args.chains or 256
    t_setup = time.perf_counter()
    gen = torch.Generator(device=dev)
    gen.manual_seed(5000 + rank)                                 # shard-keyed stream
    X = torch.randn((rows, k), generator=gen, device=dev, dtype=torch.float32).bfloat16().float()
    beta = torch.as_tensor(np.random.default_rng(5).normal(0, 0.5 / np.sqrt(k / 100.0), size=k), dtype=torch.float32, device=dev)
    y = (torch.rand(rows, generator=gen, device=dev) < torch.sigmoid(0.3 + X @ beta)).float()
    model = LogisticGLM(X, y)
    passes = 2 if bool(torch.equal(X.bfloat16().float(), X)) else 3     # the wide kernel drops the Q.Xlo / R.Xlo pass when Xlo == 0
    ndim = k + 1
    if tune_draws is None:
        ips, K, W = args.iters_per_step, args.steps, args.warmup
        total = (K + W) * ips
        tune = total // 2
        warm_iters = W * ips
    else:
        tune, draws = tune_draws
        total = tune + draws
        warm_iters = 0
    eng = model.engine(chains, dtype=args.dtype, device=ctx.local_rank)
    eng.set_state(start_points(ndim, chains, 0) * 0.1, chain_seeds(chains, 0), 0.25 / ndim ** 0.25, np.zeros(ndim),
                  np.ones(ndim), 10.0)
    opts = nuts_opts(args, exec_mode=_capi.B2_EXEC_LOCKSTEP)
    allreduce = (lambda t: dist.all_reduce(t)) if world > 1 else None
    trace = eng.alloc_trace(_capi.B2_NUTS, total)
    del X, y
    torch.cuda.synchronize(dev)
    setup_s = time.perf_counter() - t_setup

    chunk = total if tune_draws is not None else args.iters_per_step      # side config: one chunk, no idle tails
    done = 0
    while done < warm_iters:
        run_lockstep_sharded(eng, _capi.B2_NUTS, chunk, tune, opts, allreduce, world, out=trace, row0=done)
        done += chunk
    ctx.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.kernel_launches()
    t0 = time.perf_counter()
    ev0.record()
    steps = 0
    prof = None
    while done < total:
        n = min(chunk, total - done)
        last = done + n >= total
        out = run_lockstep_sharded(eng, _capi.B2_NUTS, n, tune, opts, allreduce, world, out=trace, row0=done,
                                   profile=last)
        steps += eng.last_lockstep_steps
        if last:
            prof = eng.last_lockstep_profile
        done += n
    ev1.record()
    ctx.barrier()
    wall = time.perf_counter() - t0
    t_local = ev0.elapsed_time(ev1) / 1e3
    secs, _ = ctx.reduce([t_local, wall], [0])
    t_timed = secs[0]
    leap = int(trace["tree_size"][warm_iters:].sum().item())          # replicated chains: count them once
    launches = eng.kernel_launches() - launches0
    ess_info = None
    if total - tune >= 20:
        ess_vec, note = device_ess(trace["q"][tune:])
        ess_info = {"min_bulk_ess": float(np.nanmin(ess_vec)), "n_scalars": int(len(ess_vec)), "over": note,
                    "caveat": "only %d draws per chain" % (total - tune)}
    if rank != 0:
        return None
    peaks = measured_peaks()
    per_step = t_timed / max(steps, 1)
    x_bytes = rows * k * 2.0 * 2                              # bf16 hi | lo tiles are streamed together (lo is all zeros here)
    units = leap / max(steps, 1)                              # chain-gradients a lock-step step really processed
    flops = 4.0 * rows * (k + 1) * units
    res = {"metric": "leapfrog_grad_evals_per_sec", "value": leap / t_timed, "unit": "grad-evals/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_timed * 1e3 / max(args.steps, 1), "higher_is_better": True,
           "scaling": "weak (rows per GPU fixed)", "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64",
           "data": "synthetic",
           "config": {"workload": "logistic regression, observations sharded: %d rows x %d features per GPU, %d rows total"
                      % (rows, k, rows * world), "chains": chains, "tune": tune, "draws": total - tune,
                      "collective": "all-reduce(sum) of [C, D+1] fp64 = %d bytes per leapfrog" % (chains * (k + 1) * 8),
                      "split_passes": passes,
                      "setup_seconds": setup_s},
           "lockstep_steps_timed": steps, "chain_grads_per_step": units, "ms_per_leapfrog_all_chains": per_step * 1e3,
           "gpu_launches": launches, "job_seconds_device": t_timed, "job_seconds_wall": secs[1], "ess": ess_info,
           "e2e": {"value": leap / secs[1], "unit": "grad-evals/s",
                   "through": "pymc3_b200.sharded.run_lockstep_sharded (data generated on the device: nothing to upload; the trace "
                              "stays on the device)"},
           "roofline": {"bound": "tensor", "achieved": flops / per_step / 1e12, "peak": peaks["tflops"], "unit": "TFLOP/s",
                        "frac": flops / per_step / 1e12 / peaks["tflops"], "traffic": None,
                        "hbm_achieved_gbs": x_bytes / per_step / 1e9, "hbm_frac": x_bytes / per_step / 1e9 / peaks["hbm_gbs"],
                        "note": "per GPU, whole lock-step step (likelihood + all-reduce + advance); flops = 4 N (D+1) x live chains; "
                                "bytes = the bf16 X tiles read once per step", "peak_source": peaks["source"]}}
    if prof:
        res["step_breakdown_us"] = prof
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    wl = make_workload("c2" if args.workload == "all" else args.workload, args)
    cores = args.cpu_cores or os.cpu_count()
    v, done, dt = cpu_reference_run(wl, steps=args.steps, warmup=args.warmup, budget_s=args.ref_budget, cores=cores)
    sample = "%d timed steps; step = 1 NUTS tuning transition on each of %d chains (1 process/chain, full data)" % (done, cores)
    line = {
        "impl": "reference", "metric": "leapfrog_grad_evals_per_sec", "value": v, "unit": "grad-evals/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": done, "warmup": args.warmup,
        "ms_per_step": dt * 1e3 / max(done, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "chains": cores, "sampler": "NUTS target_accept=0.8 (oracle port of "
                   "pymc3/step_methods/hmc, NumPy/BLAS logp_dlogp; Theano is not installable here)"},
        "cpu_baseline": {"value": v, "unit": "grad-evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "grad-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=17)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "c1", "c2", "c3", "c4", "c5"],
                    help="all = config 2 as the headline line + the other configs as shortened complete jobs under 'configs'")
    ap.add_argument("--configs", default="c1,c3,c4,c5,c2s,c3s",
                    help="side configs of --workload all (c2s / c3s = reduced C2 / C3 for the CPU ESS/s ratio)")
    ap.add_argument("--c5-rows", type=int, default=6250000, help="rows per GPU of the side config c5")
    ap.add_argument("--c5-iters", type=int, default=30, help="transitions (tune + draws) of the side config c5")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--iters-per-step", type=int, default=100)
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the workload's)")
    ap.add_argument("--n-obs", type=int, default=0)
    ap.add_argument("--n-features", type=int, default=0)
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--glm-path", default="auto", choices=["auto", "group", "simt", "tcgen05"])
    ap.add_argument("--run-ahead", action="store_true",
                    help="lock-step chains that finish a step early go on into the next step's rows (measured on C2: same job "
                         "time -- the slowest chains are slow throughout -- and the idle tail moves into the last timed steps)")
    ap.add_argument("--skip-ess", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="do not time individual likelihood launches")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi during the timed region")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--cpu-cores", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU baseline sampling")
    ap.add_argument("--ref-budget", type=float, default=150.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3                      # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
