"""GPU parity tests proper: the CUDA engine, called through the C ABI, against the CPU oracle.

Tolerances (BASELINE.json north_star): logp/dlogp within 1e-6 relative in the fp64 check build,
1e-4 in the fp32 production build; tree-building decisions reproducible for a fixed RNG stream.
"""
import numpy as np
import pytest

from oracle.hmc_cpu import CpuHMC, CpuNUTS, run_chain
from oracle.potentials import DiagAdaptPotential
from oracle.rng import PhiloxRNG
from pymc3_b200 import _capi
from tests import models_util

pytestmark = pytest.mark.gpu

NUTS_OPTS = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75,
                 t0=10.0, adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=1,
                 exec_mode=0, glm_path=0)


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name", ["std_normal", "eight_schools", "glm", "hier", "stoch_vol"])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-6), ("float32", 1e-4)])
def test_logp_dlogp_matches_oracle(name, dtype, tol):
    model, oracle = models_util.pairs()[name]
    rng = np.random.default_rng(5)
    q = rng.normal(size=(16, oracle.ndim)) * 0.4
    eng = model.engine(16, dtype=dtype)
    logp, grad = eng.logp_dlogp(q)
    logp, grad = logp.cpu().numpy(), grad.cpu().numpy().astype("f8")
    for i in range(len(q)):
        l0, g0 = oracle(q[i].astype(dtype).astype("f8"))
        assert abs(logp[i] - l0) <= tol * max(1.0, abs(l0)), (name, i, logp[i], l0)
        assert _rel(grad[i], g0) <= tol, (name, i)
    eng.close()


def test_glm_tcgen05_path_matches_oracle():
    """the tensor-core kernel: 3 row splits x 2 chain tiles, ragged N (zero-padded rows), odd K"""
    from oracle import densities as od
    from pymc3_b200 import model as pm
    for n, k, chains in [(1000 + 37, 13, 70), (5000, 100, 300), (64, 127, 5)]:
        X, y = models_util.glm_data(n, k, seed=7)
        model, oracle = pm.LogisticGLM(X, y), od.LogisticGLM(X, y)
        rng = np.random.default_rng(6)
        q = (rng.normal(size=(chains, oracle.ndim)) * 0.5).astype("f4")
        eng = model.engine(chains, dtype="float32")
        logp, grad = eng.logp_dlogp(q, glm_path=_capi.B2_GLM_TCGEN05)
        logp, grad = logp.cpu().numpy(), grad.cpu().numpy().astype("f8")
        for i in range(0, chains, max(1, chains // 16)):
            l0, g0 = oracle(q[i].astype("f8"))
            assert abs(logp[i] - l0) <= 1e-4 * abs(l0), (n, k, i, logp[i], l0)
            assert _rel(grad[i], g0) <= 1e-4, (n, k, i, _rel(grad[i], g0))
        eng.close()


def test_glm_wide_tcgen05_path_matches_oracle():
    """128 <= K <= 256 features (config C5's D = 256): CTA pairs own the two feature halves, the intercept is
    added in the epilogue; ragged N (masked padding rows), partial second half, partial chain tile"""
    from oracle import densities as od
    from pymc3_b200 import model as pm
    for n, k, chains in [(3000 + 17, 256, 140), (2048, 200, 40), (500, 128, 3), (40000, 256, 256)]:
        X, y = models_util.glm_data(n, k, seed=8)
        model, oracle = pm.LogisticGLM(X, y), od.LogisticGLM(X, y)
        rng = np.random.default_rng(6)
        q = (rng.normal(size=(chains, oracle.ndim)) * 0.3).astype("f4")
        eng = model.engine(chains, dtype="float32")
        logp, grad = eng.logp_dlogp(q, glm_path=_capi.B2_GLM_TCGEN05)
        logp, grad = logp.cpu().numpy(), grad.cpu().numpy().astype("f8")
        for i in range(0, chains, max(1, chains // 12)):
            l0, g0 = oracle(q[i].astype("f8"))
            assert abs(logp[i] - l0) <= 1e-4 * abs(l0), (n, k, i, logp[i], l0)
            assert _rel(grad[i], g0) <= 1e-4, (n, k, i, _rel(grad[i], g0))
        eng.close()


def test_glm_wide_tcgen05_lockstep_agrees_with_simt_lockstep():
    """K = 160: the auto path is the wide tensor-core likelihood + plain advance kernel; same chains on the SIMT
    likelihood agree for the first transitions (fp32 round-off diverges chaotically afterwards)"""
    from pymc3_b200 import model as pm
    X, y = models_util.glm_data(6000, 160, seed=11)
    model = pm.LogisticGLM(X, y)
    C, D = 20, 161
    q0 = np.random.default_rng(12).uniform(-1, 1, size=(C, D)) * 0.1
    seeds = np.arange(C) + 500
    a = _run_engine(model, q0, seeds, 12, 12, _capi.B2_NUTS, "float32", _capi.B2_EXEC_LOCKSTEP, glm_path=_capi.B2_GLM_TCGEN05)
    b = _run_engine(model, q0, seeds, 12, 12, _capi.B2_NUTS, "float32", _capi.B2_EXEC_LOCKSTEP, glm_path=_capi.B2_GLM_SIMT)
    assert (a["depth"][:6] == b["depth"][:6]).mean() > 0.95
    assert np.abs(a["q"][:4] - b["q"][:4]).max() < 5e-3
    assert all(r.phase == _capi.PHASE_DONE for r in a["reports"])
    assert np.isfinite(a["energy"]).all()


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-6), ("float32", 1e-4)])
@pytest.mark.parametrize("path", [_capi.B2_GLM_GROUP, _capi.B2_GLM_SIMT])
def test_glm_chain_batched_paths(dtype, tol, path):
    """ragged sizes: N not a multiple of the 64-row tile, chains not a multiple of 64, K odd."""
    from oracle import densities as od
    from pymc3_b200 import model as pm
    X, y = models_util.glm_data(1000 + 37, 13, seed=7)
    model, oracle = pm.LogisticGLM(X, y), od.LogisticGLM(X, y)
    rng = np.random.default_rng(6)
    q = rng.normal(size=(70, oracle.ndim)) * 0.5
    eng = model.engine(70, dtype=dtype)
    logp, grad = eng.logp_dlogp(q, glm_path=path)
    logp, grad = logp.cpu().numpy(), grad.cpu().numpy().astype("f8")
    for i in range(len(q)):
        l0, g0 = oracle(q[i].astype(dtype).astype("f8"))
        assert abs(logp[i] - l0) <= tol * abs(l0)
        assert _rel(grad[i], g0) <= tol
    eng.close()


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-6), ("float32", 1e-4)])
def test_hier_chain_batched_kernel(dtype, tol):
    from oracle import densities as od
    from pymc3_b200 import model as pm
    idx, floor, y = models_util.hier_data(40000 + 123, 85, seed=2)
    model, oracle = pm.HierLinearNCP(idx, floor, y, 85), od.HierLinearNCP(idx, floor, y, 85)
    rng = np.random.default_rng(8)
    q = rng.normal(size=(130, oracle.ndim)) * 0.3
    eng = model.engine(130, dtype=dtype)       # 130 * 40123 >= 2^22 -> chain-batched slab kernel
    logp, grad = eng.logp_dlogp(q)
    logp, grad = logp.cpu().numpy(), grad.cpu().numpy().astype("f8")
    for i in range(0, len(q), 7):
        l0, g0 = oracle(q[i].astype(dtype).astype("f8"))
        assert abs(logp[i] - l0) <= tol * abs(l0)
        assert _rel(grad[i], g0) <= tol
    eng.close()


def _run_engine(model, q0, seeds, n, tune, kind, dtype, exec_mode, step0=None, **over):
    eng = model.engine(len(q0), dtype=dtype)
    D = eng.D
    eng.set_state(q0, seeds, step0 or 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
    opts = dict(NUTS_OPTS)
    if kind == _capi.B2_HMC:
        opts["target_accept"] = 0.65
    opts.update(over)
    opts["exec_mode"] = exec_mode
    out = eng.run(kind, n, tune, opts)
    out = {k: v.cpu().numpy() for k, v in out.items()}
    out["reports"] = eng.reports()
    eng.close()
    return out


@pytest.mark.parametrize("name", ["std_normal", "eight_schools", "glm", "hier", "stoch_vol"])
@pytest.mark.parametrize("exec_mode", [_capi.B2_EXEC_PERSISTENT, _capi.B2_EXEC_LOCKSTEP])
def test_nuts_decisions_match_recursive_oracle(name, exec_mode):
    """fp64, fixed step size: 150 transitions (100 tuning with mass adaptation) draw for draw."""
    model, oracle = models_util.pairs()[name]
    D = oracle.ndim
    rng = np.random.default_rng(1)
    C = 3
    q0 = rng.uniform(-1, 1, size=(C, D))
    seeds = [11, 2 ** 40 + 5, 123456789]
    # stochastic volatility trajectories are chaotic enough to amplify summation-order
    # round-off past 1e-7 after ~100 draws (same on the CPU, see tests/test_hostsim.py): shorter run
    n, tune = (60, 40) if name == "stoch_vol" else (150, 110)
    step0 = 0.02 if name == "hier" else None        # the default start step diverges on this posterior
    out = _run_engine(model, q0, seeds, n, tune, _capi.B2_NUTS, "float64", exec_mode, step0=step0,
                      adapt_step_size=0)
    for c in range(C):
        s = CpuNUTS(oracle, D, DiagAdaptPotential(D, np.zeros(D), np.ones(D), 10), PhiloxRNG(seeds[c]),
                    adapt_step_size=False)
        if step0:
            s.adapter.log_step = s.adapter.log_bar = np.log(step0)
        qs, st = run_chain(s, q0[c], n, tune)
        assert (st["depth"] == out["depth"][:, c]).all()
        assert (st["tree_size"] == out["tree_size"][:, c]).all()
        assert (st["diverging"] == out["diverging"][:, c].astype(bool)).all()
        assert np.abs(qs - out["q"][:, c]).max() < 1e-7
        assert np.abs(st["energy"] - out["energy"][:, c]).max() < 1e-6
        assert np.abs(st["mean_tree_accept"] - out["mean_tree_accept"][:, c]).max() < 1e-7
        assert out["reports"][c].phase == _capi.PHASE_DONE
        assert out["reports"][c].n_grad == st["tree_size"].sum() + 1


def test_nuts_with_dual_averaging_matches_oracle_early():
    """with step-size adaptation the tuning dynamics amplify round-off, so compare the first draws"""
    model, oracle = models_util.pairs()["eight_schools"]
    D = oracle.ndim
    q0 = np.random.default_rng(2).uniform(-1, 1, size=(2, D))
    seeds = [7, 8]
    out = _run_engine(model, q0, seeds, 30, 30, _capi.B2_NUTS, "float64", _capi.B2_EXEC_AUTO)
    for c in range(2):
        s = CpuNUTS(oracle, D, DiagAdaptPotential(D, np.zeros(D), np.ones(D), 10), PhiloxRNG(seeds[c]))
        qs, st = run_chain(s, q0[c], 30, 30)
        assert (st["depth"][:20] == out["depth"][:20, c]).all()
        assert np.abs(qs[:15] - out["q"][:15, c]).max() < 1e-6
        assert np.abs(st["step_size"][:15] - out["step_size"][:15, c]).max() < 1e-6
        assert np.abs(st["step_size_bar"][:15] - out["step_size_bar"][:15, c]).max() < 1e-6


@pytest.mark.parametrize("exec_mode", [_capi.B2_EXEC_PERSISTENT, _capi.B2_EXEC_LOCKSTEP])
def test_hmc_matches_oracle(exec_mode):
    model, oracle = models_util.pairs()["std_normal"]
    D = oracle.ndim
    q0 = np.random.default_rng(3).uniform(-1, 1, size=(2, D))
    seeds = [21, 22]
    out = _run_engine(model, q0, seeds, 200, 150, _capi.B2_HMC, "float64", exec_mode, adapt_step_size=0)
    for c in range(2):
        s = CpuHMC(oracle, D, DiagAdaptPotential(D, np.zeros(D), np.ones(D), 10), PhiloxRNG(seeds[c]),
                   adapt_step_size=False)
        qs, st = run_chain(s, q0[c], 200, 150)
        assert (st["n_steps"] == out["n_steps"][:, c]).all()
        assert (st["accepted"] == out["accepted"][:, c].astype(bool)).all()
        assert np.abs(qs - out["q"][:, c]).max() < 1e-6
        assert np.abs(st["accept"] - out["accept"][:, c]).max() < 1e-6


def test_fp32_production_build_runs_and_is_sane():
    """priors of the reference's published eight-schools table: mu~N(0,5),
    tau~HalfCauchy(5) -> posterior mu 4.46 (sd 3.31), tau 3.59 (sd 3.23)."""
    from pymc3_b200 import model as pm
    model = pm.EightSchoolsNCP(mu_sd=5.0, tau_beta=5.0)
    D = 10
    C = 64
    rng = np.random.default_rng(4)
    q0 = rng.uniform(-1, 1, size=(C, D))
    out = _run_engine(model, q0, np.arange(C) + 100, 600, 300, _capi.B2_NUTS, "float32", _capi.B2_EXEC_AUTO)
    acc = out["mean_tree_accept"][300:].mean()
    assert 0.7 < acc < 0.95
    mu = out["q"][300:, :, 8]
    tau = np.exp(out["q"][300:, :, 9])
    assert abs(mu.mean() - 4.46) < 0.35 and abs(mu.std() - 3.31) < 0.35
    assert abs(tau.mean() - 3.59) < 0.4
    assert out["tune"][:300].all() and not out["tune"][300:].any()


def test_bad_initial_energy_is_reported():
    from pymc3_b200 import model as pm
    model = pm.StdNormal(3)
    eng = model.engine(2, dtype="float64")
    q0 = np.array([[0.0, 0.0, 0.0], [np.inf, 0.0, 0.0]])
    eng.set_state(q0, [1, 2], 0.1, np.zeros(3), np.ones(3), 10.0)
    eng.run(_capi.B2_NUTS, 5, 5, dict(NUTS_OPTS))
    rep = eng.reports()
    assert rep[0].phase == _capi.PHASE_DONE
    assert rep[1].phase == _capi.PHASE_FAILED and rep[1].fail_code == _capi.FAIL_BAD_INITIAL_ENERGY
    eng.close()


@pytest.mark.parametrize("exec_mode", [_capi.B2_EXEC_PERSISTENT, _capi.B2_EXEC_LOCKSTEP])
def test_block_per_chain_group_on_full_stochastic_volatility(exec_mode):
    """D = 2907 > 1024 -> a whole block owns a chain (config C4's mapping)."""
    from oracle import densities as od
    from pymc3_b200 import model as pm
    model, oracle = pm.StochVol(), od.StochVol(pm.sp500_log_returns())
    D = oracle.ndim
    assert D == 2907
    rng = np.random.default_rng(9)
    q0 = rng.uniform(-1, 1, size=(2, D)) * 0.1
    q0[:, 0] = -3.0
    q0[:, -1] = 2.0
    eng = model.engine(2, dtype="float64")
    logp, grad = eng.logp_dlogp(q0)
    for c in range(2):
        l0, g0 = oracle(q0[c])
        assert abs(logp.cpu().numpy()[c] - l0) <= 1e-9 * abs(l0)
        assert _rel(grad.cpu().numpy()[c], g0) <= 1e-9
    eng.close()
    seeds = [31, 32]
    n = 12
    out = _run_engine(model, q0, seeds, n, n, _capi.B2_NUTS, "float64", exec_mode, adapt_step_size=0,
                      early_max_treedepth=5)
    for c in range(2):
        s = CpuNUTS(oracle, D, DiagAdaptPotential(D, np.zeros(D), np.ones(D), 10), PhiloxRNG(seeds[c]),
                    adapt_step_size=False, early_max_treedepth=5)
        qs, st = run_chain(s, q0[c], n, n)
        assert (st["depth"] == out["depth"][:, c]).all()
        assert (st["tree_size"] == out["tree_size"][:, c]).all()
        assert np.abs(qs - out["q"][:, c]).max() < 1e-7
        assert np.abs(st["mean_tree_accept"] - out["mean_tree_accept"][:, c]).max() < 1e-7


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_block_per_chain_accept_statistic_stays_a_probability(dtype):
    """Regression: the warps of a block-per-chain group once raced on the per-level log-size /
    log-accept-sum scalars (nuts.py:369-380 executed redundantly by every lane), which folded
    sub-trees in twice, pushed mean_tree_accept past 1 and blew up dual averaging.  Deep trees
    (early_max_treedepth 8) with adaptation on: the statistic is a probability and no chain stalls."""
    from pymc3_b200 import model as pm
    model = pm.StochVol()
    C, D, n = 16, 2907, 160
    q0 = np.random.default_rng(10).uniform(-1, 1, size=(C, D))
    out = _run_engine(model, q0, np.arange(C) + 900, n, n, _capi.B2_NUTS, dtype, _capi.B2_EXEC_PERSISTENT)
    acc = out["mean_tree_accept"]
    assert np.isfinite(acc).all() and acc.max() <= 1.0 + 1e-5 and acc.min() >= 0.0
    assert all(r.phase == _capi.PHASE_DONE for r in out["reports"])
    assert out["depth"].max() >= 7
    assert 0.6 < acc[-60:].mean() < 0.95
    assert out["step_size"][-1].max() < 1.0


@pytest.mark.parametrize("exec_mode", [_capi.B2_EXEC_PERSISTENT, _capi.B2_EXEC_LOCKSTEP])
def test_chunked_runs_continue_the_same_chains(exec_mode):
    """b2_sample_run called 3 x 20 iterations == one call of 60 (device state persists)."""
    model, oracle = models_util.pairs()["eight_schools"]
    D = oracle.ndim
    q0 = np.random.default_rng(12).uniform(-1, 1, size=(5, D))
    seeds = np.arange(5) + 40
    one = _run_engine(model, q0, seeds, 60, 40, _capi.B2_NUTS, "float64", exec_mode)
    eng = model.engine(5, dtype="float64")
    eng.set_state(q0, seeds, 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
    opts = dict(NUTS_OPTS)
    opts["exec_mode"] = exec_mode
    parts = [eng.run(_capi.B2_NUTS, 20, 40, opts) for _ in range(3)]
    q = np.concatenate([p["q"].cpu().numpy() for p in parts])
    depth = np.concatenate([p["depth"].cpu().numpy() for p in parts])
    tune = np.concatenate([p["tune"].cpu().numpy() for p in parts])
    assert np.array_equal(q, one["q"]) and np.array_equal(depth, one["depth"])
    assert tune[:40].all() and not tune[40:].any()
    assert all(r.iter == 60 for r in eng.reports())
    eng.close()


@pytest.mark.parametrize("which", ["eight_schools_f64", "glm_simt_f32", "glm_tcgen05_f32"])
def test_run_ahead_chunks_equal_one_run(which):
    """lock-step chunks with run_ahead: chains that finish a chunk early go on into the following rows of
    the preallocated trace; the per-chain results are those of one uninterrupted run, bit for bit."""
    from pymc3_b200 import model as pm
    if which == "eight_schools_f64":
        model, dtype, C = models_util.pairs()["eight_schools"][0], "float64", 37
        D = 10
        extra = {}
    else:
        X, y = models_util.glm_data(4000, 20, seed=21)
        model, dtype, C, D = pm.LogisticGLM(X, y), "float32", 150, 21
        extra = {"glm_path": _capi.B2_GLM_TCGEN05 if which == "glm_tcgen05_f32" else _capi.B2_GLM_SIMT}
    q0 = np.random.default_rng(13).uniform(-1, 1, size=(C, D)) * 0.5
    seeds = np.arange(C) + 70
    one = _run_engine(model, q0, seeds, 45, 30, _capi.B2_NUTS, dtype, _capi.B2_EXEC_LOCKSTEP, **extra)
    eng = model.engine(C, dtype=dtype)
    eng.set_state(q0, seeds, 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
    opts = dict(NUTS_OPTS)
    opts.update(extra)
    opts["exec_mode"] = _capi.B2_EXEC_LOCKSTEP
    trace = eng.alloc_trace(_capi.B2_NUTS, 45)
    ahead = []
    for i in range(3):
        eng.run(_capi.B2_NUTS, 15, 30, opts, out=trace, row0=15 * i, run_ahead=True)
        ahead.append(max(r.iter for r in eng.reports()) - 15 * (i + 1))
    got = {k: v.cpu().numpy() for k, v in trace.items()}
    if which == "glm_tcgen05_f32":
        # the tensor-core launch re-divides its grid over the live chain tiles, so the slab summation order (and
        # the last bits) depend on which chains are live together: equal decisions early on, not bit-equal
        assert (got["depth"][:8] == one["depth"][:8]).mean() > 0.97
        assert np.abs(got["q"][:5] - one["q"][:5]).max() < 1e-3
        assert np.isfinite(got["energy"]).all() and np.array_equal(got["tune"], one["tune"])
    else:
        for k in ("q", "depth", "tree_size", "energy", "step_size", "tune", "diverging"):
            assert np.array_equal(got[k], one[k]), k
    assert ahead[0] > 0 and ahead[2] == 0            # somebody ran ahead; nobody beyond the last row
    assert all(r.iter == 45 and r.phase == _capi.PHASE_DONE for r in eng.reports())
    eng.close()


def test_fused_tensor_core_lockstep_agrees_with_simt_lockstep():
    """fp32 NUTS on a GLM big enough for the chain-batched kernels: the fused tcgen05 step
    ({likelihood, finalize+advance+repack}) against the SIMT path -- same decisions early on,
    same posterior summaries later (round-off separates individual trajectories)."""
    from pymc3_b200 import model as pm
    X, y = models_util.glm_data(6000, 20, seed=11)
    model = pm.LogisticGLM(X, y)
    D, C = 21, 256
    rng = np.random.default_rng(3)
    q0 = rng.uniform(-0.2, 0.2, size=(C, D))
    seeds = np.arange(C) + 1000
    runs = {}
    for name, path in (("tc", _capi.B2_GLM_TCGEN05), ("simt", _capi.B2_GLM_SIMT)):
        runs[name] = _run_engine(model, q0, seeds, 300, 150, _capi.B2_NUTS, "float32", _capi.B2_EXEC_LOCKSTEP,
                                 glm_path=path)
    a, b = runs["tc"], runs["simt"]
    assert (a["depth"][:5] == b["depth"][:5]).mean() > 0.97
    assert np.abs(a["q"][:3] - b["q"][:3]).max() < 5e-3
    ma, mb = a["q"][150:].mean(axis=(0, 1)), b["q"][150:].mean(axis=(0, 1))
    sa, sb = a["q"][150:].std(axis=(0, 1)), b["q"][150:].std(axis=(0, 1))
    assert np.abs(ma - mb).max() < 0.1 * sa.max() and np.allclose(sa, sb, rtol=0.1)
    acc_a, acc_b = a["mean_tree_accept"][150:].mean(), b["mean_tree_accept"][150:].mean()
    assert abs(acc_a - acc_b) < 0.02 and 0.7 < acc_a < 0.95
    assert all(r.phase == _capi.PHASE_DONE for r in a["reports"])


def test_stepwise_and_observation_sharded_runs_match_the_plain_run():
    """(a) the stepwise API with one rank reproduces b2_sample_run bit for bit;
    (b) two row shards on one GPU, their packed (logp, dlogp) summed as the all-reduce would, give the
    same chains as the unsharded engine (fp64, to summation-order round-off)."""
    import torch
    from pymc3_b200 import model as pm
    from pymc3_b200.sharded import row_shard, run_lockstep_sharded
    import ctypes as C
    X, y = models_util.glm_data(3000, 12, seed=21)
    D, Cn, n, tune = 13, 6, 60, 40
    rng = np.random.default_rng(5)
    q0 = rng.uniform(-0.3, 0.3, size=(Cn, D))
    seeds = np.arange(Cn) + 77
    opts = dict(NUTS_OPTS)
    opts["exec_mode"] = _capi.B2_EXEC_LOCKSTEP
    opts["glm_path"] = _capi.B2_GLM_SIMT
    opts["adapt_step_size"] = 0
    ref = _run_engine(pm.LogisticGLM(X, y), q0, seeds, n, tune, _capi.B2_NUTS, "float64", _capi.B2_EXEC_LOCKSTEP,
                      glm_path=_capi.B2_GLM_SIMT, adapt_step_size=0)
    # (a) one rank, stepwise
    eng = pm.LogisticGLM(X, y).engine(Cn, dtype="float64")
    eng.set_state(q0, seeds, 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
    one = run_lockstep_sharded(eng, _capi.B2_NUTS, n, tune, opts, None, 1)
    assert np.array_equal(one["q"].cpu().numpy(), ref["q"]) and np.array_equal(one["depth"].cpu().numpy(), ref["depth"])
    eng.close()
    # (b) two shards, all-reduce emulated by adding the two packed buffers
    engs = []
    for r in range(2):
        lo, hi = row_shard(len(y), 2, r)
        e = pm.LogisticGLM(X[lo:hi], y[lo:hi]).engine(Cn, dtype="float64")
        e.set_state(q0, seeds, 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
        engs.append(e)
    lib = engs[0].lib
    traces = [e.alloc_trace(_capi.B2_NUTS, n) for e in engs]
    packed = [torch.zeros((Cn, D + 1), dtype=torch.float64, device="cuda") for _ in engs]
    o = _capi.SamplerOpts(kind=_capi.B2_NUTS, n_iters=n, tune_until=tune, **opts)
    for e, tr_ in zip(engs, traces):
        tr = _capi.TraceOut()
        for name, t in tr_.items():
            setattr(tr, "d_" + name, t.data_ptr())
        _capi.check(lib.b2_step_begin(e.handle, C.byref(o), C.byref(tr), e._stream()), lib)
    active = C.c_int32(1)
    while active.value:
        for _ in range(8):
            for e, p in zip(engs, packed):
                _capi.check(lib.b2_step_likelihood(e.handle, p.data_ptr(), e._stream()), lib)
            total = packed[0] + packed[1]
            for e, p in zip(engs, packed):
                p.copy_(total)
                _capi.check(lib.b2_step_advance(e.handle, p.data_ptr(), 2, e._stream()), lib)
        _capi.check(lib.b2_step_active(engs[0].handle, C.byref(active), engs[0]._stream()), lib)
    a, b = traces[0]["q"].cpu().numpy(), traces[1]["q"].cpu().numpy()
    assert np.array_equal(a, b)                                       # replicated state machines stay identical
    assert (traces[0]["depth"].cpu().numpy() == ref["depth"]).all()
    assert np.abs(a - ref["q"]).max() < 1e-8
    assert np.abs(traces[0]["model_logp"].cpu().numpy() - ref["model_logp"]).max() < 1e-6
    for e in engs:
        lib.b2_step_end(e.handle)
        e.close()


def test_tensor_core_lockstep_is_reproducible_despite_chain_compaction():
    """Active chains are packed into dense tiles through atomics (slot order varies run to run); a chain's
    results must not depend on which tile row it lands in: two identical runs give bit-identical traces."""
    from pymc3_b200 import model as pm
    X, y = models_util.glm_data(5000, 30, seed=13)
    model = pm.LogisticGLM(X, y)
    D, C = 31, 300                       # 3 chain tiles, chains finish their chunks at different steps
    q0 = np.random.default_rng(8).uniform(-0.5, 0.5, size=(C, D))
    seeds = np.arange(C) + 5000
    runs = [_run_engine(model, q0, seeds, 80, 50, _capi.B2_NUTS, "float32", _capi.B2_EXEC_LOCKSTEP,
                        glm_path=_capi.B2_GLM_TCGEN05) for _ in range(2)]
    for key in ("q", "depth", "tree_size", "energy", "step_size"):
        assert np.array_equal(runs[0][key], runs[1][key]), key
    assert all(r.phase == _capi.PHASE_DONE for r in runs[0]["reports"])
