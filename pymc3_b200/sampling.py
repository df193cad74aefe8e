"""`sample` / `init_nuts` / `iter_sample` with the reference's signatures and semantics
(pymc3/sampling.py:230-579, 786-845, 1837-2014), dispatching to the chain-batched device engine.

The reference chooses between `_mp_sample` (one OS process per chain, parallel_sampling.py:353)
and `_sample_many` (sequential); this module adds the third branch SURVEY 8b describes: when the
step method advertises `_batched`, ALL chains run inside one engine per GPU and the trace comes
back in bulk.  Everything around that (argument normalisation, per-chain seeds, jittered starts,
`discard_tuned_samples`, MultiTrace, report, convergence checks, errors) keeps the reference's
behaviour.  Chains shard over `devices` with no communication.
"""
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
from .step_methods.hmc import NUTS
from .step_methods.hmc.quadpotential import QuadPotentialDiagAdapt

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
           **kwargs):
    """Draw samples from the posterior using the given step method (sampling.py:230).

    `cores` is accepted for signature compatibility; chain parallelism is the GPU's.
    `devices`: CUDA device indices to shard chains over (default: the step's device).
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

    if step is None:
        _log.info("Auto-assigning NUTS sampler...")
        start_, step = init_nuts(init=init, chains=chains, n_init=n_init, model=model,
                                 random_seed=random_seed, progressbar=progressbar, **kwargs)
        if start is None:
            start = start_
    elif kwargs:
        raise ValueError("Unknown arguments to `sample`: %s" % sorted(kwargs))   # :132-134
    if not getattr(step, "_batched", False):
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
    mtrace = _sample_batched(step, model, draws, tune, chains, start, random_seed, devices, chain_idx)
    t_sampling = time.time() - t_start

    discard = tune if discard_tuned_samples else 0
    mtrace = mtrace[discard:]                             # :556-557
    if trace is not None:
        mtrace = _append_traces(trace, mtrace)
    mtrace.report._n_tune = int(tune)
    mtrace.report._n_draws = int(draws - tune)
    mtrace.report._t_sampling = t_sampling
    if compute_convergence_checks:
        if draws - tune < 100:
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
    out = np.empty((chains, model.ndim))
    for c in range(chains):
        point = dict(model.test_point)
        point.update({k: v for k, v in start[c].items() if k in point})       # update_start_vals
        out[c] = model.dict_to_array(point)
    return out


def _sample_batched(step, model, draws, tune, chains, start, seeds, devices, chain_idx=0):
    """All chains in one engine per device; chains are split contiguously over devices."""
    devices = list(devices) if devices is not None else [step.device]
    q0 = _start_array(model, start, chains)
    seeds = np.asarray(seeds, dtype=np.uint64)
    bounds = np.linspace(0, chains, len(devices) + 1).astype(int)
    shards = [(dev, bounds[i], bounds[i + 1]) for i, dev in enumerate(devices) if bounds[i + 1] > bounds[i]]
    results = [None] * len(shards)
    errors = []

    def work(k, dev, lo, hi):
        try:
            eng = step._make_engine(hi - lo, device=dev)
            step._init_engine_state(eng, q0[lo:hi], seeds[lo:hi])
            out = eng.run(step._kind, draws, tune, step._opts())
            host = {name: t.cpu().numpy() for name, t in out.items()}
            results[k] = (host, eng.reports(), eng.mass_var(), eng.kernel_launches())
            eng.close()
        except Exception as err:  # surfaced on the caller's thread
            errors.append(err)

    if len(shards) == 1:
        work(0, *shards[0])
    else:
        threads = [threading.Thread(target=work, args=(k,) + sh) for k, sh in enumerate(shards)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]

    host = {name: np.concatenate([r[0][name] for r in results], axis=1) for name in results[0][0]}
    reports = [rep for r in results for rep in r[1]]
    mass_var = np.concatenate([r[2] for r in results])
    step._last_kernel_launches = sum(r[3] for r in results)
    step._last_reports = reports

    for c, rep in enumerate(reports):
        if rep.phase == _capi.PHASE_FAILED:
            try:
                step._raise_bad_energy(chain_idx + c, mass_var[c])
            except SamplingError as err:
                raise SamplingError("Bad initial energy (chain %d)" % (chain_idx + c)) from err

    # [draws, C, D] -> per-variable [C, draws, *shape] (vectorised replacement of ndarray.py:266)
    q = np.ascontiguousarray(np.swapaxes(host.pop("q"), 0, 1)).astype("f8")
    values = model.expand(q)
    stat_dtypes = step.stats_dtypes[0]
    stats = {}
    for key, dt in stat_dtypes.items():
        if key == "path_length":
            stats[key] = np.full((chains, draws), float(step.path_length))
        else:
            stats[key] = np.ascontiguousarray(host[key].T).astype(dt)
    accept_key = "mean_tree_accept" if "mean_tree_accept" in stats else "accept"
    straces = []
    for c in range(chains):
        st = NDArray.from_arrays(model, chain_idx + c, {n: v[c] for n, v in values.items()},
                                 {k: v[c] for k, v in stats.items()})
        st._add_warnings(step._chain_warnings(reports[c], stats[accept_key][c, tune:], stats["diverging"][c],
                                              stats["tune"][c]))
        straces.append(st)

    # leave the step object in the state the reference's would be in after sampling
    step.iter_count = draws
    step.tune = not (tune < draws)
    step.step_size = float(reports[0].step_size)
    step.step_adapt.sync(reports[0].step_size, reports[0].step_size_bar, stats[accept_key][0, tune:])
    step._samples_after_tune = reports[0].n_post
    step._num_divs_sample = reports[0].n_div_post
    if hasattr(step, "_reached_max_treedepth"):
        step._reached_max_treedepth = reports[0].n_maxdepth_post
    if hasattr(step.potential, "sync"):
        step.potential.sync(mass_var[0])
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
    elif init in ("advi+adapt_diag_grad", "advi+adapt_diag", "advi", "advi_map", "map", "nuts"):
        raise NotImplementedError("init=%r needs the variational / MAP subsystems, which are outside the "
                                  "sampler hot path; use 'adapt_diag' or 'jitter+adapt_diag'" % init)
    else:
        raise ValueError("Unknown initializer: {}.".format(init))
    mean = np.mean([model.dict_to_array(vals) for vals in start], axis=0)
    var = np.ones_like(mean)
    potential = QuadPotentialDiagAdapt(model.ndim, mean, var, 10)
    step = NUTS(potential=potential, model=model, **kwargs)
    return start, step
