#!/usr/bin/env python
"""Benchmark of the NUTS hot path (contract: see the task statement; layout: DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2]

Metric (BASELINE.json): leapfrog gradient evaluations / second, whole job (all chains, all GPUs),
with min bulk ESS / second reported beside it.  Workload (default, `c2`): Bernoulli-logit GLM
N=100 000, D=100 (+intercept), 1024 chains per GPU, synthetic data of SURVEY 8(d).

A "step" is `--iters-per-step` (100) NUTS transitions of every chain, continuing one sampling job:
the first half of the (W+K)*100 iterations tunes (dual averaging + diagonal mass adaptation), the
second half draws.  With the defaults W=3, K=17 this is exactly tune=1000 / draws=1000.

  value  grad-evals/s over the K timed steps, data and state resident in HBM (CUDA events)
  e2e    same metric for a second, identical job in which every step also copies the step's inputs
         (X, y) host->device from pinned memory and reads the step's trace + stats back to the host
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
def run_b200(args):
    import torch
    import torch.distributed as dist
    from pymc3_b200 import _capi, stats as b2stats

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    wl = make_workload(args.workload, args)
    model, chains = wl["model"], wl["chains"]          # chains per GPU (weak scaling)
    first_chain = rank * chains
    ndim = model.ndim
    ips = args.iters_per_step
    K, W = args.steps, args.warmup
    total = (K + W) * ips
    tune = total // 2
    opts = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75,
                t0=10.0, adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=0,
                exec_mode=_capi.B2_EXEC_AUTO,
                glm_path={"auto": 0, "group": 1, "simt": 2, "tcgen05": 3}[args.glm_path])
    q0 = start_points(ndim, chains, first_chain)
    seeds = chain_seeds(chains, first_chain)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def new_engine():
        eng = model.engine(chains, dtype=args.dtype, device=local_rank)
        eng.set_state(q0, seeds, 0.25 / ndim ** 0.25, np.zeros(ndim), np.ones(ndim), 10.0)
        return eng

    # ---- pass A: data and chain state resident in HBM
    eng = new_engine()
    trace = eng.alloc_trace(_capi.B2_NUTS, total)          # the whole job's device trace, allocated up front
    barrier()
    tw0 = time.perf_counter()
    ahead = args.run_ahead                 # lock-step runs: fast chains go on into the next step's rows (no-op otherwise)

    def grads(e):
        return sum(r.n_grad for r in e.reports())

    for s in range(W):
        eng.run(_capi.B2_NUTS, ips, tune, opts, out=trace, row0=s * ips, run_ahead=ahead)
    torch.cuda.synchronize(dev)
    wall_warm = time.perf_counter() - tw0
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0 and not args.no_clocks:
        clocks.start()
    eng.set_profiling(False)                 # per-launch events cost ~7 % of a lock-step step: separate pass P below
    launches0 = eng.kernel_launches()
    grads0 = grads(eng)                      # leapfrogs are counted when they are executed (engine counters)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for s in range(K):
        eng.run(_capi.B2_NUTS, ips, tune, opts, out=trace, row0=(W + s) * ips, run_ahead=ahead)
    ev1.record()
    barrier()
    wall_timed = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clock_info = clocks.stop() if rank == 0 else None
    launches = eng.kernel_launches() - launches0
    reports = eng.reports()
    leap_timed = sum(r.n_grad for r in reports) - grads0
    failed = sum(1 for r in reports if r.phase != _capi.PHASE_DONE)
    eng.close()

    tree = trace["tree_size"]                                           # [total, C]
    leap_all = int(tree.sum().item())
    t_vec = torch.tensor([dev_ms / 1e3, wall_timed, wall_warm], dtype=torch.float64, device=dev)
    cnt = torch.tensor([leap_timed, leap_all, launches, failed], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    t_timed = float(t_vec[0].item())
    value = float(cnt[0].item()) / t_timed

    # ---- min bulk ESS over every scalar of every free and back-transformed variable (post-tune draws)
    ess_info = None
    if not args.skip_ess and total - tune >= 100:
        q = trace["q"][tune:]                                              # [draws, C, D]
        ess_note = "every scalar of every free and back-transformed variable"
        if q.shape[2] > 512:
            # rank-normalised ESS of ~3000 scalars x 512k draws takes minutes on the host: the hyper-parameters
            # (first and last 8 columns) plus 112 evenly spaced latent states; bulk ESS is rank based, so the
            # elementwise monotone back-transforms do not change it
            Dq = q.shape[2]
            cols = np.unique(np.concatenate([np.arange(8), np.arange(Dq - 8, Dq),
                                             np.linspace(8, Dq - 9, 112).astype(int)]))
            qs = q[:, :, torch.as_tensor(cols, device=q.device)]
            qs = qs.permute(1, 0, 2).contiguous().cpu().numpy().astype("f8")
            ess_vec = np.ravel(b2stats.ess(qs))
            ess_note = "%d of %d free scalars (first/last 8 + 112 evenly spaced)" % (len(cols), Dq)
        else:
            q = q.permute(1, 0, 2).contiguous().cpu().numpy().astype("f8")  # [C, draws, D]
            vals = model.expand(q)
            ess_vec = np.concatenate([np.ravel(b2stats.ess(v)) for v in vals.values()])
        ess_t = torch.tensor(ess_vec, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ess_t, op=dist.ReduceOp.SUM)                    # independent chain sets add
        ess_info = {"min_bulk_ess": float(ess_t.min().item()), "n_scalars": int(ess_t.numel()), "over": ess_note}
    del trace

    # ---- pass P: the same job again with CUDA events around every likelihood / advance launch of the K timed
    #      steps (kernel durations for the roofline; its own elapsed time is the denominator of the shares)
    like_ms, like_n, adv_ms, prof_ms, leap_prof = 0.0, 0, 0.0, 0.0, 0
    if not args.no_profile and rank == 0 and args.workload in ("c2", "c3"):     # the lock-step workloads
        eng = new_engine()
        ptrace = eng.alloc_trace(_capi.B2_NUTS, total)
        for s in range(W):
            eng.run(_capi.B2_NUTS, ips, tune, opts, out=ptrace, row0=s * ips, run_ahead=ahead)
        torch.cuda.synchronize(dev)
        eng.set_profiling(True)
        pg0 = grads(eng)
        pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pv0.record()
        for s in range(K):
            eng.run(_capi.B2_NUTS, ips, tune, opts, out=ptrace, row0=(W + s) * ips, run_ahead=ahead)
        pv1.record()
        torch.cuda.synchronize(dev)
        prof_ms = pv0.elapsed_time(pv1)
        like_ms, like_n = eng.profile()
        adv_ms = eng.profile_advance()
        leap_prof = grads(eng) - pg0
        eng.close()
        del ptrace

    # ---- pass B: end to end -- every step uploads its inputs from pinned host memory and reads
    #      its trace + stats back into pinned host memory (same seeds => same job)
    eng = new_engine()
    host_in = [t.cpu().pin_memory() for t in eng._keep]
    h2d = sum(t.numel() * t.element_size() for t in host_in)
    pinned, d2h = {}, 0
    etrace = eng.alloc_trace(_capi.B2_NUTS, total)
    for s in range(W):
        eng.run(_capi.B2_NUTS, ips, tune, opts, out=etrace, row0=s * ips, run_ahead=ahead)
    barrier()
    eg0 = grads(eng)
    e0 = time.perf_counter()
    for s in range(K):
        for src, dst in zip(host_in, eng._keep):
            dst.copy_(src, non_blocking=True)
        out = eng.run(_capi.B2_NUTS, ips, tune, opts, out=etrace, row0=(W + s) * ips, run_ahead=ahead)
        for name, t in out.items():                    # this step's rows are complete for every chain
            if name not in pinned:
                pinned[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            pinned[name].copy_(t, non_blocking=True)
        torch.cuda.synchronize(dev)
    barrier()
    e2e_s = time.perf_counter() - e0
    leap_e2e = grads(eng) - eg0
    del etrace
    d2h = sum(t.numel() * t.element_size() for t in pinned.values())
    eng.close()
    e_vec = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    l_vec = torch.tensor([leap_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(l_vec, op=dist.ReduceOp.SUM)
    e2e_value = float(l_vec.item()) / float(e_vec.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    roofline = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if args.workload == "c2" and os.path.exists(tpath):        # one ncu --set full capture, per launch
        with open(tpath) as f:
            traffic = json.load(f).get("k_glm_tc_main", {}).get("dram_bytes_per_launch")
    if args.workload == "c4":
        # persistent kernel: one launch per step, so the per-unit figure of SURVEY 8d (70 KB per chain-grad if the
        # state streamed through HBM) is set against the whole-step rate
        ach = wl["bytes_per_chain_grad"] * value / max(world, 1) / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                    "traffic": None, "kernel": "k_persistent_block (whole NUTS transition loop, one launch per step)",
                    "regime": "hot state (q, p, grad, p_sum, proposal, mass) is resident in shared memory, so the kernel is "
                              "instruction-issue bound (ncu: issue slots 44 % busy, DRAM 27 GB/s), not HBM bound",
                    "peak_source": peaks["source"]}
    if like_n > 0 and wl["bound"]:
        per_launch_s = like_ms / 1e3 / like_n
        # units one launch processes = chains still inside a trajectory (the launch skips finished chains and,
        # on the tensor-core path, compacts the live ones into dense tiles): counted, not assumed
        units = leap_prof / like_n
        if wl["bound"] == "tensor":
            ach = wl["flops_per_chain_grad"] * units / per_launch_s / 1e12
            roofline = {"bound": "tensor", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                        "frac": ach / peaks["tflops"], "traffic": traffic}
        else:
            ach = wl["bytes_per_chain_grad"] * units / per_launch_s / 1e9
            roofline = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": ach / peaks["hbm_gbs"], "traffic": None}
            if wl.get("flops_per_chain_grad"):
                # once a tile of observations is shared by 128 chains the bound is FP32 issue, not HBM
                fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12          # nominal: SMs x FP32 lanes x 2 x max SM clock
                fl = wl["flops_per_chain_grad"] * units / per_launch_s / 1e12
                roofline.update({"fp32_achieved_tflops": fl, "fp32_peak_tflops_nominal": fp32_peak, "fp32_frac": fl / fp32_peak})
        roofline.update({"kernel": "chain-batched likelihood (logp+dlogp, all chains)", "launches_timed": int(like_n),
                         "chain_grads_per_launch": units,
                         "avg_launch_us": per_launch_s * 1e6, "kernel_share_of_step": like_ms / prof_ms,
                         "advance_kernel_avg_us": adv_ms * 1e3 / like_n, "advance_share_of_step": adv_ms / prof_ms,
                         "timed_in": "a separate identical pass with CUDA events around every launch (adds ~7 % to a step); value is measured without them",
                         "peak_source": peaks["source"]})

    cpu = None
    if world == 1 and not args.skip_cpu:
        cores = args.cpu_cores or os.cpu_count()
        v, done, dt = cpu_reference_run(wl, steps=10 ** 6, warmup=1, budget_s=args.cpu_budget, cores=cores)
        cpu = {"value": v, "unit": "grad-evals/s", "cores": cores, "kind": "port",
               "sample": "%d tuning transitions on each of %d chains (1 process/chain, full data), %.1f s"
                         % (done, cores, dt)}

    line = {
        "metric": "leapfrog_grad_evals_per_sec", "value": value, "unit": "grad-evals/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": t_timed * 1e3 / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "chains_per_gpu": chains, "chains_total": chains * world,
                   "iters_per_step": ips, "tune": tune, "draws": total - tune, "sampler": "NUTS target_accept=0.8",
                   "l2": L2_NOTES.get(args.workload, "inputs change every launch"),
                   "grad_evals_counted": "engine leapfrog counters read before and after the timed steps",
                   "step_boundaries": ("lock-step chains that finish a step early run ahead into the next step's rows; a step "
                                       "ends when every chain has done its transitions" if ahead else "all chains stop at every step boundary")},
        # whole sampling job incl. tuning (mirrors benchmarks/benchmarks/benchmarks.py:163-169)
        "min_bulk_ess_per_sec": (ess_info["min_bulk_ess"] / (t_timed + float(t_vec[2].item()))
                                 if ess_info else None),
        "job_seconds": t_timed + float(t_vec[2].item()),
        "ess": ess_info,
        "e2e": {"value": e2e_value, "unit": "grad-evals/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(cnt[2].item()),
        "failed_chains": int(cnt[3].item()),
        "clocks": clock_info,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_c5(args):
    """Config 5: logistic regression with the OBSERVATIONS sharded over the GPUs (weak scaling in rows:
    6.25 M rows x 256 features per GPU = 50 M rows on 8), every rank runs all 256 chains redundantly and
    the ranks all-reduce the packed [C, D+1] (logp, dlogp) partials once per leapfrog over NCCL."""
    import torch
    import torch.distributed as dist
    from pymc3_b200 import _capi
    from pymc3_b200.model import LogisticGLM
    from pymc3_b200.sharded import run_lockstep_sharded
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    rows = args.n_obs or 6250000
    k = args.n_features or 256
    chains = args.chains or 256
    gen = torch.Generator(device=dev)
    gen.manual_seed(5000 + rank)                                 # shard-keyed stream
    X = torch.randn((rows, k), generator=gen, device=dev, dtype=torch.float32).bfloat16().float()
    beta = torch.as_tensor(np.random.default_rng(5).normal(0, 0.5 / np.sqrt(k / 100.0), size=k), dtype=torch.float32, device=dev)
    y = (torch.rand(rows, generator=gen, device=dev) < torch.sigmoid(0.3 + X @ beta)).float()
    model = LogisticGLM(X, y)
    ndim = k + 1
    ips, K, W = args.iters_per_step, args.steps, args.warmup
    total = (K + W) * ips
    tune = total // 2
    eng = model.engine(chains, dtype=args.dtype, device=local_rank)
    eng.set_state(start_points(ndim, chains, 0) * 0.1, chain_seeds(chains, 0), 0.25 / ndim ** 0.25, np.zeros(ndim),
                  np.ones(ndim), 10.0)
    opts = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75, t0=10.0,
                adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=0,
                exec_mode=_capi.B2_EXEC_LOCKSTEP, glm_path={"auto": 0, "group": 1, "simt": 2, "tcgen05": 3}[args.glm_path])
    allreduce = (lambda t: dist.all_reduce(t)) if world > 1 else None
    trace = eng.alloc_trace(_capi.B2_NUTS, total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for s in range(W):
        run_lockstep_sharded(eng, _capi.B2_NUTS, ips, tune, opts, allreduce, world, out=trace, row0=s * ips)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.kernel_launches()
    ev0.record()
    steps = 0
    for s in range(K):
        run_lockstep_sharded(eng, _capi.B2_NUTS, ips, tune, opts, allreduce, world, out=trace, row0=(W + s) * ips)
        steps += eng.last_lockstep_steps
    ev1.record()
    barrier()
    t_local = ev0.elapsed_time(ev1) / 1e3
    t_all = torch.tensor([t_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    t_timed = float(t_all.item())
    leap = int(trace["tree_size"][W * ips:].sum().item())          # replicated chains: count them once
    launches = eng.kernel_launches() - launches0
    if rank == 0:
        peaks = measured_peaks()
        per_step = t_timed / max(steps, 1)
        x_bytes = rows * k * 4.0
        units = leap / max(steps, 1)                              # chain-gradients a lock-step step really processed
        flops = 4.0 * rows * (k + 1) * units
        line = {"metric": "leapfrog_grad_evals_per_sec", "value": leap / t_timed, "unit": "grad-evals/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": t_timed * 1e3 / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
                "config": {"workload": "logistic regression, observations sharded: %d rows x %d features per GPU, %d rows total"
                           % (rows, k, rows * world), "chains": chains, "iters_per_step": ips, "tune": tune, "draws": total - tune,
                           "collective": "all-reduce(sum) of [C, D+1] fp64 = %d bytes per leapfrog" % (chains * (k + 1) * 8)},
                "lockstep_steps_timed": steps, "chain_grads_per_step": units, "ms_per_leapfrog_all_chains": per_step * 1e3, "gpu_launches": launches,
                "e2e": None,
                "roofline": {"bound": "hbm", "achieved": x_bytes / per_step / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": x_bytes / per_step / 1e9 / peaks["hbm_gbs"], "traffic": None,
                             "tensor_achieved_tflops": flops / per_step / 1e12, "tensor_frac": flops / per_step / 1e12 / peaks["tflops"],
                             "note": "per GPU, whole lock-step step (likelihood + all-reduce + advance); bytes = X shard read once"}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    wl = make_workload(args.workload, args)
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
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
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
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
