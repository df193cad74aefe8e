"""`sample` / `init_nuts` / `iter_sample` with the reference's signatures and semantics
(pymc3/sampling.py:230-579, 786-845, 1837-2014), dispatching to the chain-batched device engine.

The reference chooses between `_mp_sample` (one OS process per chain, parallel_sampling.py:353)
and `_sample_many` (sequential); this module adds the third branch SURVEY 8b describes: when the
step method advertises `_batched`, ALL chains run inside one engine per GPU and the trace comes
back in bulk.  Everything around that (argument normalisation, per-chain seeds, jittered starts,
`discard_tuned_samples`, MultiTrace, report, convergence checks, errors) keeps the reference's
behaviour.  Chains shard over `devices` with no communication.
"""
import collections
import logging
import threading
import time
from collections.abc import Iterable

import numpy as np

from . import _capi
from .backends.base import MultiTrace
from .backends.ndarray import NDArray
from .exceptions import SamplingError
from .model import modelcontext
from .step_methods import step_sizes
from .step_methods.hmc import NUTS
from .step_methods.hmc.quadpotential import QuadPotentialDiagAdapt, QuadPotentialFull

_log = logging.getLogger("pymc3")

__all__ = ["sample", "iter_sample", "init_nuts"]


def _cpu_count():
    """pymc3/parallel_sampling.py:448-459."""
    import multiprocessing
    try:
        return max(1, multiprocessing.cpu_count() // 2)
    except NotImplementedError:
        return 1


def sample(draws=500, step=None, init="auto", n_init=200000, start=None, trace=None, chain_idx=0,
           chains=None, cores=None, tune=500, progressbar=True, model=None, random_seed=None,
           discard_tuned_samples=True, compute_convergence_checks=True, callback=None, devices=None,
           chunk=None, run_ahead=True, **kwargs):
    """Draw samples from the posterior using the given step method (sampling.py:230).

    `cores` is accepted for signature compatibility; chain parallelism is the GPU's.
    `devices`: CUDA device indices to shard chains over (default: the step's device).
    `callback(trace, draw)` is called for every chain and draw (sampling.py:1396-1398) after each chunk of `chunk`
    transitions (default: 1 with a callback, else 100); raising KeyboardInterrupt in it, or pressing Ctrl-C,
    returns the draws completed so far (sampling.py:1407-1409).  `run_ahead=False` makes every chain stop at every
    chunk boundary instead of running ahead into the next chunk's rows (per-chunk accounting in bench.py).
    Extra keyword arguments configure the auto-assigned NUTS sampler (sampling.py:439-451).
    """
    model = modelcontext(model)
    if cores is None:
        cores = min(4, _cpu_count())                      # :389-390
    if chains is None:
        chains = max(2, cores)                            # :402-403
    if isinstance(start, dict):
        start = [start] * chains
    if random_seed == -1:
        random_seed = None
    if chains == 1 and isinstance(random_seed, int):
        random_seed = [random_seed]
    if random_seed is None or isinstance(random_seed, int):
        if random_seed is not None:
            np.random.seed(random_seed)
        random_seed = [np.random.randint(2 ** 30) for _ in range(chains)]   # :410-413
    if not isinstance(random_seed, Iterable):
        raise TypeError("Invalid value for `random_seed`. Must be tuple, list or int")
    random_seed = list(random_seed)
    if len(random_seed) != chains:
        raise ValueError("Need one random seed per chain (%d), got %d" % (chains, len(random_seed)))
    if start is not None:
        for start_vals in start:
            _check_start_shape(model, start_vals)

    draws = int(draws)
    if draws < 0 or tune < 0:
        raise ValueError("draws and tune must be >= 0")
    draws += tune                                         # :434
    if draws < 1:
        raise ValueError("Argument `draws` must be greater than 0.")

    if isinstance(step, (list, tuple)):                   # sampling.py:142-165 builds a CompoundStep from a list
        if len(step) != 1:
            raise NotImplementedError("CompoundStep (several step methods on disjoint variables, step_methods/compound.py) "
                                      "is outside the device path: the engine samples all continuous variables jointly")
        step = step[0]
    if step is None:
        _log.info("Auto-assigning NUTS sampler...")
        start_, step = init_nuts(init=init, chains=chains, n_init=n_init, model=model,
                                 random_seed=random_seed, progressbar=progressbar, **kwargs)
        if start is None:
            start = start_
    elif kwargs:
        raise ValueError("Unknown arguments to `sample`: %s" % sorted(kwargs))   # :132-134
    if not hasattr(step, "_host_driven"):
        raise NotImplementedError("pymc3_b200.sample runs NUTS / HamiltonianMC step methods only")
    if start is None:
        start = [model.test_point] * chains
    if trace is not None:
        # resume: new draws are appended to an in-memory MultiTrace and every chain starts from its last
        # point (sampling.py:893-894, ndarray.py:221-231); like the reference, sampler state is not restored
        if not isinstance(trace, MultiTrace):
            raise NotImplementedError("only in-memory MultiTrace continuation is supported")
        if trace.nchains != chains:
            raise ValueError("trace has %d chains, but chains=%d" % (trace.nchains, chains))
        start = [{k: v for k, v in trace.point(-1, chain=c).items() if k in model.free_RVs} for c in trace.chains]

    t_start = time.time()
    if getattr(step, "_batched", False):
        mtrace = _sample_batched(step, model, draws, tune, chains, start, random_seed, devices, chain_idx,
                                 callback=callback, chunk=chunk, run_ahead=run_ahead, progressbar=progressbar)
    else:
        # a user potential / step_rand: chains one after the other, draws one at a time through step.step(),
        # like the reference's _sample_many -> _iter_sample (sampling.py:786-936)
        mtrace = _sample_sequential(step, model, draws, tune, chains, start, random_seed, chain_idx, callback)
    t_sampling = time.time() - t_start

    discard = tune if discard_tuned_samples else 0
    mtrace = mtrace[discard:]                             # :556-557
    if trace is not None:
        mtrace = _append_traces(trace, mtrace)
    mtrace.report._n_tune = int(tune)
    mtrace.report._n_draws = int(max(0, len(mtrace) - (0 if discard_tuned_samples else tune)))
    mtrace.report._t_sampling = t_sampling
    if compute_convergence_checks:
        if len(mtrace) - (0 if discard_tuned_samples else tune) < 100:
            _log.warning("The number of samples is too small to check convergence reliably.")
        else:
            mtrace.report._run_convergence_checks(mtrace, model)
    mtrace.report._log_summary()
    return mtrace


def _append_traces(old, new):
    """Per chain: concatenate the draws and sampler stats of `new` behind those of `old`."""
    straces = []
    for c_old, c_new in zip(old.chains, new.chains):
        a, b = old._straces[c_old], new._straces[c_new]
        samples = {k: np.concatenate([a.samples[k], b.samples[k]]) for k in b.samples}
        stats = None
        if a._stats is not None and b._stats is not None:
            stats = {k: np.concatenate([a._stats[0][k], b._stats[0][k]]) for k in b._stats[0]}
        st = NDArray.from_arrays(new._straces[c_new].model, c_old, samples, stats)
        st._add_warnings(getattr(a, "_warnings", []) + getattr(b, "_warnings", []))
        straces.append(st)
    return MultiTrace(straces)


def _check_start_shape(model, start):
    """sampling.py:582-604."""
    if not isinstance(start, dict):
        raise TypeError("start argument must be a dict or an array-like of dicts")
    e = ""
    for name, shape in model.free:
        if name in start:
            got = np.shape(start[name])
            if tuple(got) != tuple(shape):
                e += "\nExpected shape {} for var '{}', got: {}".format(tuple(shape), name, got)
    if e != "":
        raise ValueError("Bad shape for start argument:{}".format(e))


def _start_array(model, start, chains):
    """[chains, ndim] start positions; variables a start dict lacks come from the test point (update_start_vals,
    sampling.py:1900-1926).  One pass per variable over all chains, not a dict_to_array per chain."""
    out = np.empty((chains, model.ndim))
    test_point = model.test_point
    if all(start[c] is start[0] for c in range(chains)):          # start=None or a single start point
        point = dict(test_point)
        point.update({k: v for k, v in start[0].items() if k in point})
        out[:] = model.dict_to_array(point)
        return out
    for vm in model.ordering().vmap:
        default = np.ravel(test_point[vm.var])
        out[:, vm.slc] = np.stack([np.ravel(s[vm.var]) if vm.var in s else default for s in start[:chains]])
    return out


Draw = collections.namedtuple("Draw", ["chain", "is_last", "draw_idx", "tuning", "stats", "point", "warnings"])
"""What `callback(trace=, draw=)` receives -- the fields of pymc3/parallel_sampling.py:348-351."""


def _choose_chains(traces, tune):
    """sampling.py:1417-1443: after an interrupt keep the set of chains that maximises the number of draws."""
    if tune is None:
        tune = 0
    if not traces:
        return []
    lengths = [max(0, len(trace) - tune) for trace in traces]
    if not sum(lengths):
        raise ValueError("Not enough samples to build a trace.")
    idxs = np.argsort(lengths)[::-1]
    l_sort = np.array(lengths)[idxs]
    final_length = l_sort[0]
    last_total = 0
    for i, length in enumerate(l_sort):
        total = (i + 1) * length
        if total < last_total:
            use_until = i
            break
        last_total = total
        final_length = length
    else:
        use_until = len(lengths)
    return [traces[idx] for idx in idxs[:use_until]], final_length + tune


class _ShardRun:
    """One device's share of the chains: engine, device trace for the whole job, and a host copy that a
    background thread fills chunk by chunk (device -> pinned staging pieces -> host array) while the next chunk
    samples -- the sampling call releases the GIL, so copies and MCMC overlap."""

    PIECE_BYTES = 48 << 20                    # staging buffers: 2 x 48 MB of pinned memory per shard

    def __init__(self, step, dev, q0, seeds, draws):
        import queue
        import torch
        self.torch, self.step, self.draws = torch, step, draws
        t0 = time.perf_counter()
        self.eng = step._make_engine(len(q0), device=dev)
        t1 = time.perf_counter()
        step._init_engine_state(self.eng, q0, seeds)
        t2 = time.perf_counter()
        self.trace = self.eng.alloc_trace(step._kind, draws)
        self.host = {k: np.empty(tuple(v.shape), dtype=_np_dtype(v)) for k, v in self.trace.items()}
        self.copy_stream = torch.cuda.Stream(device=self.eng.dev)
        self.init_phases = {"engine": t1 - t0, "state": t2 - t1, "trace_buffers": time.perf_counter() - t2}
        self.rows_done = 0
        self.device_seconds = 0.0
        self.chunk_log = []                   # (rows done, wall clock, device seconds, leapfrogs so far) after every chunk
        self.log_grads = bool(getattr(step, "_log_chunk_grads", False))
        self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self._jobs = queue.Queue()
        self._error = None
        self._thread = threading.Thread(target=self._copier, daemon=True)
        self._thread.start()

    def _copier(self):
        torch = self.torch
        try:
            torch.cuda.set_device(self.eng.dev)
            row_bytes = sum(int(np.prod(t.shape[1:])) * t.element_size() for t in self.trace.values())
            piece = max(1, self.PIECE_BYTES // max(row_bytes, 1))
            staging = [{name: torch.empty((piece,) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=True)
                        for name, t in self.trace.items()} for _ in range(2)]
            events = [None, None]
            spans = [None, None]
            for arr in self.host.values():        # first touch of the (lazily mapped) host arrays happens here, not
                arr.fill(0)                       # on the critical path of the last copies

            def drain(k):
                if events[k] is not None:
                    events[k].synchronize()
                    lo, hi = spans[k]
                    for name, buf in staging[k].items():
                        self.host[name][lo:hi] = buf[: hi - lo].numpy()
                    events[k] = None
            k = 0
            while True:
                job = self._jobs.get()
                if job is None:
                    break
                (lo, hi), ready = job
                for p_lo in range(lo, hi, piece):
                    p_hi = min(hi, p_lo + piece)
                    drain(k)
                    with torch.cuda.stream(self.copy_stream):
                        self.copy_stream.wait_event(ready)
                        for name, t in self.trace.items():
                            staging[k][name][: p_hi - p_lo].copy_(t[p_lo:p_hi], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(self.copy_stream)
                    events[k], spans[k] = ev, (p_lo, p_hi)
                    k ^= 1
            drain(0)
            drain(1)
        except BaseException as err:          # surfaced by finish()
            self._error = err

    def run_chunk(self, n, tune, run_ahead):
        """`n` more transitions of every chain of this shard; their rows are handed to the copier thread."""
        torch = self.torch
        lo, hi = self.rows_done, self.rows_done + n
        with torch.cuda.device(self.eng.dev):
            self._ev[0].record()
            self.eng.run(self.step._kind, n, tune, self.step._opts(), out=self.trace, row0=lo, run_ahead=run_ahead)
            self._ev[1].record()
            ready = torch.cuda.Event()
            ready.record()
            self._jobs.put(((lo, hi), ready))
            self._ev[1].synchronize()
            self.device_seconds += self._ev[0].elapsed_time(self._ev[1]) / 1e3
        self.rows_done = hi
        n_grad = sum(rep.n_grad for rep in self.eng.reports()) if self.log_grads else 0
        self.chunk_log.append((hi, time.perf_counter(), self.device_seconds, n_grad))
        return lo, hi

    def flush(self):
        """block until every row handed over so far is in the host arrays"""
        self._jobs.put(None)
        self._thread.join()
        if self._error is not None:
            raise self._error
        self._thread = threading.Thread(target=self._copier, daemon=True)
        self._thread.start()

    def finish(self):
        t0 = time.perf_counter()
        self._jobs.put(None)
        self._thread.join()
        if self._error is not None:
            raise self._error
        t1 = time.perf_counter()
        out = (self.eng.reports(), self.eng.mass_var(), self.eng.kernel_launches())
        t2 = time.perf_counter()
        self.eng.close()
        t3 = time.perf_counter()
        self.trace = None
        self.finish_phases = {"copies": t1 - t0, "reports": t2 - t1, "engine_close": t3 - t2, "trace_free": time.perf_counter() - t3}
        return out


def _np_dtype(t):
    import torch
    return {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32, torch.uint8: np.uint8,
            torch.int64: np.int64}[t.dtype]


def _progress(total, chains, enabled):
    """the reference's progress bar (sampling.py:1370-1382, fastprogress there), advanced once per chunk; only on a
    terminal, so logs and pipes stay clean"""
    import sys
    if not enabled or not sys.stderr.isatty():
        return None
    try:
        from tqdm import tqdm
    except ImportError:
        return None
    return tqdm(total=total, unit="draw", desc="Sampling %d chains" % chains, leave=False, mininterval=0.5)


def _sample_batched(step, model, draws, tune, chains, start, seeds, devices, chain_idx=0, callback=None,
                    chunk=None, run_ahead=True, progressbar=False):
    """All chains in one engine per device; chains are split contiguously over devices.

    The job runs in chunks of `chunk` transitions (default 100; 1 when a `callback` is given, so that it sees
    every draw like the reference's per-draw loop, sampling.py:1384-1398): chains that finish a chunk early run
    ahead into the following rows of the preallocated device trace, each finished chunk is copied to the host
    while the next one samples, `callback(trace=, draw=)` is called per chain and draw of the chunk, and a
    KeyboardInterrupt (from the callback or the user) returns what is complete so far, chosen like
    sampling.py:1407-1443."""
    devices = list(devices) if devices is not None else [step.device]
    tm = {"t0": time.perf_counter()}
    q0 = _start_array(model, start, chains)
    seeds = np.asarray(seeds, dtype=np.uint64)
    bounds = np.linspace(0, chains, len(devices) + 1).astype(int)
    shards = [(dev, bounds[i], bounds[i + 1]) for i, dev in enumerate(devices) if bounds[i + 1] > bounds[i]]
    if chunk is None:
        chunk = 1 if callback is not None else 100
    chunk = max(1, min(int(chunk), draws))
    tm["start_array"] = time.perf_counter()
    runs = [_ShardRun(step, dev, q0[lo:hi], seeds[lo:hi], draws) for dev, lo, hi in shards]
    tm["engines"] = time.perf_counter()
    stat_dtypes = step.stats_dtypes[0]
    interrupted = False
    done = 0

    def run_all(n):
        errors = []
        if len(runs) == 1:
            runs[0].run_chunk(n, tune, run_ahead)
            return

        def work(r):
            try:
                r.run_chunk(n, tune, run_ahead)
            except BaseException as err:      # surfaced on the caller's thread
                errors.append(err)
        threads = [threading.Thread(target=work, args=(r,)) for r in runs]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]

    # bench.py hook: time the individual likelihood / advance launches of the last chunks of the job
    n_chunks = (draws + chunk - 1) // chunk
    prof_from = n_chunks - int(getattr(step, "_profile_last_chunks", 0) or 0)
    prof = None
    bar = _progress(draws, chains, progressbar)
    try:
        while done < draws:
            n = min(chunk, draws - done)
            if done // chunk == prof_from and prof_from >= 0 and prof is None:
                for r in runs:
                    r.eng.set_profiling(True)
                prof = {"n_grad0": sum(rep.n_grad for r in runs for rep in r.eng.reports()), "t0": time.perf_counter()}
            run_all(n)
            done += n
            if bar is not None:
                bar.update(n)
            if callback is not None:
                for r in runs:
                    r.flush()
                _chunk_callbacks(callback, step, model, runs, shards, done - n, done, draws, tune, chain_idx, stat_dtypes)
    except KeyboardInterrupt:
        interrupted = True
    finally:
        if bar is not None:
            bar.close()

    if prof is not None:
        like = [r.eng.profile() for r in runs]
        prof.update(seconds=time.perf_counter() - prof["t0"],
                    n_grad=sum(rep.n_grad for r in runs for rep in r.eng.reports()) - prof["n_grad0"],
                    like_ms=sum(x[0] for x in like), like_n=sum(x[1] for x in like),
                    adv_ms=sum(r.eng.profile_advance() for r in runs), shards=len(runs))
    step._last_profile = prof
    tm["chunks"] = time.perf_counter()
    results = [r.finish() for r in runs]
    tm["finish"] = time.perf_counter()
    rows = min(r.rows_done for r in runs) if not interrupted else done
    if len(runs) == 1:
        host = {name: arr[:rows] for name, arr in runs[0].host.items()}
    else:
        host = {name: np.concatenate([r.host[name][:rows] for r in runs], axis=1) for name in runs[0].host}
    reports = [rep for r in results for rep in r[0]]
    mass_var = np.concatenate([r[1] for r in results])
    step._last_kernel_launches = sum(r[2] for r in results)
    step._last_reports = reports
    step._last_device_seconds = max(r.device_seconds for r in runs)
    step._last_chunk_log = [(rows_, t_ - tm["t0"], d_, g_) for rows_, t_, d_, g_ in runs[0].chunk_log]
    step._last_n_grad = int(sum(rep.n_grad for rep in reports))
    step._last_finish_phases = dict(runs[0].finish_phases, **{"init_" + k: v for k, v in runs[0].init_phases.items()})

    for c, rep in enumerate(reports):
        if rep.phase == _capi.PHASE_FAILED:
            try:
                step._raise_bad_energy(chain_idx + c, mass_var[c])
            except SamplingError as err:
                raise SamplingError("Bad initial energy (chain %d)" % (chain_idx + c)) from err

    straces = _bulk_straces(step, model, host, reports, rows, tune, chains, chain_idx, stat_dtypes)
    tm["straces"] = time.perf_counter()
    keys = ["start_array", "engines", "chunks", "finish", "straces"]
    step._last_timing = {k: tm[k] - tm[p] for k, p in zip(keys, ["t0"] + keys[:-1])}      # seconds per host phase
    if interrupted:
        straces, length = _choose_chains(straces, tune)
        return MultiTrace(straces)[:length]

    # leave the step object in the state the reference's would be in after sampling
    accept_key = "mean_tree_accept" if "mean_tree_accept" in host else "accept"
    step.iter_count = draws
    step.tune = not (tune < draws)
    step.step_size = float(reports[0].step_size)
    step.step_adapt.sync(reports[0].step_size, reports[0].step_size_bar, host[accept_key][tune:, 0])
    step._samples_after_tune = reports[0].n_post
    step._num_divs_sample = reports[0].n_div_post
    if hasattr(step, "_reached_max_treedepth"):
        step._reached_max_treedepth = reports[0].n_maxdepth_post
    if hasattr(step.potential, "sync"):
        step.potential.sync(mass_var[0])
    return MultiTrace(straces)


def _bulk_straces(step, model, host, reports, rows, tune, chains, chain_idx, stat_dtypes):
    """[rows, C, D] host trace -> one NDArray per chain holding VIEWS of the bulk arrays (vectorised replacement of
    the per-draw record() of ndarray.py:258-277; no per-draw Python work, no dtype round trip)."""
    q = np.swapaxes(host["q"], 0, 1)            # [C, rows, D] VIEW of the bulk trace, engine dtype (no copy)
    values = model.expand(q)
    stats = {}
    for key, dt in stat_dtypes.items():
        if key == "path_length":
            stats[key] = np.full((chains, rows), float(step.path_length))
        else:
            stats[key] = host[key].T.astype(dt, copy=False)
    accept_key = "mean_tree_accept" if "mean_tree_accept" in stats else "accept"
    # sampler warnings for all chains from bulk reductions; per-draw divergence records only where there are any
    div_rows = [np.nonzero(stats["diverging"][c])[0] for c in range(chains)] if stats["diverging"].any() else None
    n_post = max(0, rows - tune)
    mean_acc = stats[accept_key][:, tune:].mean(axis=1) if n_post else np.full(chains, np.nan)
    accept_ok = step_sizes.acceptance_in_interval(mean_acc, n_post, step.target_accept)     # all chains at once
    names, snames = list(values.keys()), list(stats.keys())
    vals, svals = [values[n] for n in names], [stats[k] for k in snames]
    straces = []
    for c in range(chains):
        st = NDArray.from_arrays(model, chain_idx + c, dict(zip(names, [v[c] for v in vals])),
                                 dict(zip(snames, [v[c] for v in svals])))
        st._add_warnings(step._chain_warnings(reports[c], mean_acc[c], n_post,
                                              div_rows[c] if div_rows is not None else (), stats["tune"][c],
                                              accept_ok=bool(accept_ok[c])))
        straces.append(st)
    return straces


def _chunk_callbacks(callback, step, model, runs, shards, lo, hi, draws, tune, chain_idx, stat_dtypes):
    """callback(trace=, draw=) for every chain and every draw of rows [lo, hi), in the reference's shape: `trace`
    is the chain's trace up to and including the draw (len(trace) works), `draw` a Draw tuple."""
    for r, (dev, c_lo, c_hi) in zip(runs, shards):
        q = r.host["q"]
        for c in range(c_hi - c_lo):
            for i in range(lo, hi):
                qc = np.ascontiguousarray(q[: i + 1, c])
                values = model.expand(qc)
                stats = {k: (np.full(i + 1, float(step.path_length)) if k == "path_length"
                             else r.host[k][: i + 1, c].astype(dt, copy=False)) for k, dt in stat_dtypes.items()}
                st = NDArray.from_arrays(model, chain_idx + c_lo + c, values, stats)
                point = {k: v[i] for k, v in values.items()}
                row = {k: v[i] for k, v in stats.items()}
                callback(trace=st, draw=Draw(chain_idx + c_lo + c, i == draws - 1, i, i < tune, [row], point, None))


def _sample_sequential(step, model, draws, tune, chains, start, seeds, chain_idx, callback):
    """sampling.py:641-690, 847-936 for step methods that are driven from the host (user potentials)."""
    straces = []
    try:
        for c in range(chains):
            np.random.seed(int(seeds[c]) % (2 ** 32))
            point = dict(model.test_point)
            point.update({k: v for k, v in start[c].items() if k in point})
            strace = NDArray(model=model)
            strace.setup(draws, chain_idx + c, step.stats_dtypes)
            straces.append(strace)
            step.tune = bool(tune)
            step.iter_count = 0
            for i in range(draws):
                if i == tune:
                    step.stop_tuning()
                point, stats = step.step(point)
                strace.record(point, stats)
                if callback is not None:
                    warns = step.warnings() if i == draws - 1 else None
                    callback(trace=strace, draw=Draw(chain_idx + c, i == draws - 1, i, i < tune, stats, point, warns))
            strace.close()
            strace._add_warnings(step.warnings())
    except KeyboardInterrupt:
        for st in straces:
            st.close()
        straces, length = _choose_chains(straces, tune)
        return MultiTrace(straces)[:length]
    return MultiTrace(straces)


def iter_sample(draws, step, start=None, trace=None, chain=0, tune=None, model=None, random_seed=None,
                callback=None):
    """Generator over single draws of ONE chain (sampling.py:786-845): yields the trace so far.
    Drives `step.step(point)`, i.e. the same device code one transition at a time."""
    model = modelcontext(model)
    draws = int(draws)
    if draws < 1:
        raise ValueError("Argument `draws` must be greater than 0.")
    if random_seed is not None:
        np.random.seed(random_seed)
    point = dict(model.test_point)
    if start is not None:
        point.update(start)
    strace = NDArray(model=model)
    strace.setup(draws, chain, step.stats_dtypes)
    step.tune = bool(tune)
    step.iter_count = 0
    for i in range(draws):
        if i == tune:
            step.stop_tuning()
        point, stats = step.step(point)
        strace.record(point, stats)
        yield MultiTrace([strace])[: i + 1]
    strace.close()


def trace_cov(trace, vars=None, model=None):
    """Covariance matrix of the flattened free variables over a trace (pymc3/tuning/scaling.py:113-141)."""
    model = modelcontext(model)
    names = list(model.free_RVs) if model is not None else list(vars if vars is not None else trace.varnames)
    cols = []
    for name in names:
        x = np.asarray(trace[str(name)])
        cols.append(x.reshape(x.shape[0], -1))
    return np.cov(np.concatenate(cols, axis=1).T)


def init_nuts(init="auto", chains=1, n_init=500000, model=None, random_seed=None, progressbar=True, **kwargs):
    """Starting points + an adaptive diagonal potential for NUTS (sampling.py:1837-2014).

    Implements 'adapt_diag' and 'jitter+adapt_diag' (= 'auto'), sampling.py:1915-1929; the
    ADVI / MAP based initialisations are outside the hot path.
    """
    model = modelcontext(model)
    if not isinstance(init, str):
        raise TypeError("init must be a string.")
    init = init.lower()
    if init == "auto":
        init = "jitter+adapt_diag"
    _log.info("Initializing NUTS using {}...".format(init))
    if random_seed is not None:
        random_seed = int(np.atleast_1d(random_seed)[0])
        np.random.seed(random_seed)
    if init == "adapt_diag":
        start = [model.test_point] * chains
    elif init == "jitter+adapt_diag":
        start = []
        for _ in range(chains):
            mean = {var: np.array(val, dtype="f8", copy=True) for var, val in model.test_point.items()}
            for val in mean.values():
                val[...] += 2 * np.random.rand(*val.shape) - 1
            start.append(mean)
    elif init == "nuts":
        # sampling.py:2002-2007: a pilot NUTS run; its covariance becomes a dense static metric (QuadPotentialFull, which
        # the device runs in z = L^-1 q) and random pilot draws become the start points
        with model:
            pilot = sample(draws=n_init, step=NUTS(**kwargs), tune=n_init // 2, random_seed=random_seed,
                           progressbar=progressbar, compute_convergence_checks=False)
        cov = np.atleast_1d(trace_cov(pilot, model=model))
        start = [pilot.point(int(i)) for i in np.random.randint(len(pilot), size=chains)]
        return start, NUTS(potential=QuadPotentialFull(cov), model=model, **kwargs)
    elif init in ("advi+adapt_diag_grad", "advi+adapt_diag", "advi", "advi_map", "map"):
        raise NotImplementedError("init=%r needs the variational / MAP subsystems, which are outside the "
                                  "sampler hot path; use 'adapt_diag', 'jitter+adapt_diag' or 'nuts'" % init)
    else:
        raise ValueError("Unknown initializer: {}.".format(init))
    mean = np.mean([model.dict_to_array(vals) for vals in start], axis=0)
    var = np.ones_like(mean)
    potential = QuadPotentialDiagAdapt(model.ndim, mean, var, 10)
    step = NUTS(potential=potential, model=model, **kwargs)
    return start, step
