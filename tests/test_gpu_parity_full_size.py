"""Parity of the chain-batched kernels AT THE SIZES bench.py runs them (BASELINE.json configs 2, 3, 5), through
the C ABI, against the CPU oracle evaluated on a subset of the chains; plus the reference's leapfrog
reversibility test and the MCSE z-test of the fp32 tensor-core lock-step path against CPU NUTS.

Tolerances (BASELINE.json north_star): fp32 production build 1e-4 relative for logp and dlogp (dlogp relative to
its largest component, as numpy.testing.assert_allclose(rtol) on a vector norm would).  The relative bound on
logp admits several nats at |logp| ~ 7e4, which says little about the energies NUTS decisions hinge on, so the
GLM tests also assert an ABSOLUTE bound on logp.
"""
import json
import os

import numpy as np
import pytest

from pymc3_b200 import _capi
from tests import models_util

pytestmark = pytest.mark.gpu

ABS_LOGP = 0.02            # nats; energy errors of this size move an acceptance probability by < 2 %

NUTS_OPTS = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75,
                 t0=10.0, adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=1,
                 exec_mode=0, glm_path=0)


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _c2():
    import bench
    return bench.glm_synthetic(100000, 100)


# ------------------------------------------------------------------------------- config C2
def _glm_mode(X, y, iters=8):
    """posterior mode of the logistic regression by Newton's method (fp64; the N(0, 1e6) prior is negligible)"""
    Xa = np.concatenate([np.ones((len(y), 1)), X.astype("f8")], axis=1)
    b = np.zeros(Xa.shape[1])
    for _ in range(iters):
        p = 1.0 / (1.0 + np.exp(-(Xa @ b)))
        b += np.linalg.solve((Xa * (p * (1 - p))[:, None]).T @ Xa, Xa.T @ (y - p))
    return b


def test_tcgen05_glm_at_c2_size_all_chains_live():
    """k_glm_tc_main exactly as the headline benchmark launches it through b2_logp_dlogp: 100 000 x 100, 1024
    chains = 8 chain tiles x 18 row slabs of 1563 64-row tiles; oracle on 16 of the chains.

    Two families of positions: posterior-scale ones (mode + 3 posterior sds of jitter: where the sampler lives after
    warm-up) with the absolute bound, and far-off ones (|eta| up to ~10 and ~40, early warm-up / saturated
    sigmoids) where the tensor core's truncating fp32 accumulation (a relative bias of ~2e-7 on eta, towards zero)
    times sum_i (y_i - sigmoid_i) eta_i ~ 3e4 .. 2e5 shows as 0.01 .. 0.08 nats: bounded there at 3x the absolute
    bound or 2e-6 |logp| (fp32-level: 1e5 terms of size ~2 summed in fp32 tiles), whichever is larger -- energy
    differences at that distance from the mode are hundreds of nats."""
    from oracle import densities as od
    from pymc3_b200 import model as pm
    X, y = _c2()
    model, oracle = pm.LogisticGLM(X, y), od.LogisticGLM(X, y)
    C = 1024
    rng = np.random.default_rng(60)
    mode = _glm_mode(X, y)
    cases = [("posterior", mode + rng.normal(size=(C, oracle.ndim)) * 0.025, ABS_LOGP),
             ("far 0.3", rng.normal(size=(C, oracle.ndim)) * 0.3 / np.sqrt(10.0), 3 * ABS_LOGP),
             ("far 1.5", rng.normal(size=(C, oracle.ndim)) * 1.5 / np.sqrt(10.0), 3 * ABS_LOGP)]
    eng = model.engine(C, dtype="float32")
    for name, q, bound in cases:
        q = q.astype("f4")
        logp, grad = eng.logp_dlogp(q, glm_path=_capi.B2_GLM_TCGEN05)
        logp, grad = logp.cpu().numpy(), grad.cpu().numpy().astype("f8")
        worst, worst_g = 0.0, 0.0
        for i in range(5, C, 64):
            l0, g0 = oracle(q[i].astype("f8"))
            assert abs(logp[i] - l0) <= 1e-4 * abs(l0), (name, i, logp[i], l0)
            assert abs(logp[i] - l0) < max(bound, 2e-6 * abs(l0)), (name, i, logp[i] - l0)     # far cases: fp32-level relative
            assert _rel(grad[i], g0) <= 1e-4, (name, i, _rel(grad[i], g0))
            worst, worst_g = max(worst, abs(logp[i] - l0)), max(worst_g, _rel(grad[i], g0))
        print("C2 logp_dlogp %-9s: max |dlogp| %.2e nats, max rel dlogp error %.2e" % (name, worst, worst_g))
    eng.close()


@pytest.mark.parametrize("fused", ["1", "0"])
def test_tcgen05_lockstep_energies_at_c2_size(fused, monkeypatch):
    """The lock-step step at C2 size (likelihood + state machine, chains compacted into dense tiles as they
    finish their trees at different leapfrogs: 60-95 % live): every traced model_logp was produced by the
    tensor-core kernel for the traced position -- hold it against the oracle at that position, absolutely."""
    from oracle import densities as od
    from pymc3_b200 import model as pm
    monkeypatch.setenv("B2_TC_FUSED", fused)
    X, y = _c2()
    model, oracle = pm.LogisticGLM(X, y), od.LogisticGLM(X, y)
    C, D, n = 1024, 101, 14
    q0 = np.random.default_rng(61).uniform(-1, 1, size=(C, D)) * 0.05
    eng = model.engine(C, dtype="float32")
    eng.set_state(q0, np.arange(C) + 3000, 0.02, np.zeros(D), np.full(D, 1e-4), 10.0)
    opts = dict(NUTS_OPTS)
    opts.update(exec_mode=_capi.B2_EXEC_LOCKSTEP, glm_path=_capi.B2_GLM_TCGEN05, early_max_treedepth=6)
    out = eng.run(_capi.B2_NUTS, n, n, opts)
    q = out["q"].cpu().numpy().astype("f8")
    lp = out["model_logp"].cpu().numpy()
    sizes = out["tree_size"].cpu().numpy()
    assert all(r.phase == _capi.PHASE_DONE for r in eng.reports())
    assert len(np.unique(sizes[3:])) > 3                         # trees of different lengths: compaction was exercised
    worst = 0.0
    for c in range(7, C, 128):
        for t in (0, 1, n // 2, n - 1):
            l0, _ = oracle(q[t, c])
            assert abs(lp[t, c] - l0) < ABS_LOGP, (fused, t, c, lp[t, c], l0)
            worst = max(worst, abs(lp[t, c] - l0))
    print("C2 lock-step (fused=%s): max |model_logp - oracle| %.2e nats" % (fused, worst))
    eng.close()


# ------------------------------------------------------------------------------- config C5 (per-GPU shard)
def test_wide_tcgen05_glm_at_c5_shard_size():
    """k_glm_tcw_main on >= 3 M rows x 256 features x 256 chains (C5 runs 6.25 M rows per GPU through the same
    launch geometry: the row slabs only get longer).  X is generated on the device like bench.py does
    (bf16-representable fp32); the oracle sees the same rows in 8 chunks (logp and dlogp add over rows)."""
    import torch
    from oracle import densities as od
    from pymc3_b200 import model as pm
    rows, k, C = 3 * 2 ** 20 + 77, 256, 256
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(5005)
    X = torch.randn((rows, k), generator=gen, device=dev, dtype=torch.float32).bfloat16().float()
    beta = torch.as_tensor(np.random.default_rng(5).normal(0, 0.3, size=k), dtype=torch.float32, device=dev)
    y = (torch.rand(rows, generator=gen, device=dev) < torch.sigmoid(0.3 + X @ beta)).float()
    model = pm.LogisticGLM(X, y)
    eng = model.engine(C, dtype="float32")
    rng = np.random.default_rng(62)
    # chains 0..127: positions a few posterior sds (~1.3e-3) from the generating coefficients; chains 128..255: the
    # same positions moved by half a posterior sd -- the pairs give energy DIFFERENCES at the scale of a trajectory
    q = (rng.normal(size=(C, k + 1)) * 0.004 + np.concatenate([[0.3], beta.cpu().numpy()])).astype("f4")
    q[C // 2:] = q[:C // 2] + (rng.normal(size=(C // 2, k + 1)) * 6e-4).astype("f4")
    logp, grad = eng.logp_dlogp(q, glm_path=_capi.B2_GLM_TCGEN05)
    logp, grad = logp.cpu().numpy(), grad.cpu().numpy().astype("f8")
    eng.close()
    Xh, yh = X.cpu().numpy(), y.cpu().numpy()
    del X, y
    subset = list(range(3, C // 2, 32)) + [i + C // 2 for i in range(3, C // 2, 32)]
    l0 = np.zeros(len(subset))
    g0 = np.zeros((len(subset), k + 1))
    bounds = np.linspace(0, rows, 9).astype(int)
    tau = 1e-6
    for a, b in zip(bounds[:-1], bounds[1:]):
        chunk = od.LogisticGLM(Xh[a:b], yh[a:b], prior_tau=tau)
        for j, i in enumerate(subset):
            l, g = chunk(q[i].astype("f8"))
            l0[j] += l
            g0[j] += g
    for j, i in enumerate(subset):                                # the prior was counted once per chunk
        b_ = q[i, 1:].astype("f8")
        l0[j] -= 7 * np.sum(0.5 * (-tau * b_ * b_ + np.log(tau) - np.log(2 * np.pi)))
        g0[j, 1:] -= 7 * (-tau * b_)
        assert abs(logp[i] - l0[j]) <= 1e-4 * abs(l0[j]), (i, logp[i], l0[j])
        assert abs(logp[i] - l0[j]) <= 5e-7 * abs(l0[j]), (i, logp[i] - l0[j])  # fp32-level: |logp| ~ 2e6 here
        assert _rel(grad[i], g0[j]) <= 1e-4, (i, _rel(grad[i], g0[j]))
    # energy differences between neighbouring positions: what a NUTS decision sees
    half = len(subset) // 2
    d_gpu = logp[subset[half:]] - logp[subset[:half]]
    d_ref = l0[half:] - l0[:half]
    assert np.abs(d_gpu - d_ref).max() < 5 * ABS_LOGP, (d_gpu - d_ref)
    print("C5 shard: max |dlogp| %.2e nats (|logp| %.1e), max error of energy differences %.2e nats, max rel grad %.2e"
          % (np.abs(logp[subset] - l0).max(), np.abs(l0).max(), np.abs(d_gpu - d_ref).max(),
             max(_rel(grad[i], g0[j]) for j, i in enumerate(subset))))


# ------------------------------------------------------------------------------- config C3
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-6), ("float32", 1e-4)])
def test_hier_slab_kernel_at_c3_size(dtype, tol):
    """k_hier_slab on 1 000 000 observations in 85 unbalanced groups (bench.hier_synthetic), 512 chains; the fp32
    build folds its packed partial sums into doubles every staged tile, so 1e-4 holds with a wide margin."""
    import bench
    from oracle import densities as od
    from pymc3_b200 import model as pm
    idx, floor, y, g = bench.hier_synthetic(1000000)
    model, oracle = pm.HierLinearNCP(idx, floor, y, g), od.HierLinearNCP(idx, floor, y, g)
    C = 512
    rng = np.random.default_rng(63)
    truth = np.concatenate([[1.5, np.log(0.3), -0.7, np.log(0.3)], np.zeros(2 * g), [np.log(0.7)]])
    q = truth + rng.normal(size=(C, oracle.ndim)) * 0.05
    eng = model.engine(C, dtype=dtype)
    logp, grad = eng.logp_dlogp(q)
    logp, grad = logp.cpu().numpy(), grad.cpu().numpy().astype("f8")
    worst_l, worst_g = 0.0, 0.0
    for i in range(1, C, 37):
        l0, g0 = oracle(q[i].astype(dtype).astype("f8"))
        worst_l = max(worst_l, abs(logp[i] - l0) / abs(l0))
        worst_g = max(worst_g, _rel(grad[i], g0))
        assert abs(logp[i] - l0) <= tol * abs(l0), (i, logp[i], l0)
        assert _rel(grad[i], g0) <= tol, (i, _rel(grad[i], g0))
    print("C3 %s: max rel logp %.2e, max rel grad %.2e" % (dtype, worst_l, worst_g))
    eng.close()


# ------------------------------------------------------------------------------- leapfrog (R3)
@pytest.mark.parametrize("name", ["eight_schools", "glm", "hier", "stoch_vol"])
@pytest.mark.parametrize("dtype,rtol", [("float64", 1e-5), ("float32", 1e-4)])
def test_leapfrog_reversible(name, dtype, rtol):
    """pymc3/tests/test_hmc.py:27-46 against the device integrator (b2_leapfrog = compute_state + n x step):
    n steps forward then n steps with -epsilon return to the start within rtol 1e-5 (the reference's, met by the
    fp64 build; the fp32 production build is held to its 1e-4: 40 steps of fp32 round-off reach 1e-5 on single
    components of these 8..300-dimensional models); random diagonal scaling."""
    model, oracle = models_util.pairs()[name]
    rng = np.random.default_rng(42)
    D = oracle.ndim
    C = 3
    var = rng.random(D) + 0.05
    q0 = rng.normal(size=(C, D)) * 0.3
    p0 = rng.normal(size=(C, D)) / np.sqrt(var)
    eng = model.engine(C, dtype=dtype)
    for eps in (0.01, 0.1):
        scale = 0.1 if name in ("hier", "stoch_vol") else 1.0      # keep the 20-step paths inside the stable region
        for n_steps in (1, 2, 3, 4, 20):
            q1, p1, _ = eng.leapfrog(q0, p0, var, eps * scale, n_steps)
            q2, p2, _ = eng.leapfrog(q1, p1, var, -eps * scale, n_steps)
            # components near zero are compared on the scale of the vector they belong to
            for got, ref in ((q2, q0), (p2, p0)):
                got, ref = got.cpu().numpy().astype("f8"), ref.astype(dtype).astype("f8")
                assert np.abs(got - ref).max(axis=1).max() <= rtol * np.abs(ref).max(), (name, eps, n_steps)
    eng.close()


@pytest.mark.parametrize("name", ["eight_schools", "glm", "hier", "stoch_vol"])
def test_leapfrog_matches_oracle_integrator(name):
    """integration.py:81-109 step by step: q', p' and the energy after n steps equal the oracle's Leapfrog (fp64)."""
    from oracle.hmc_cpu import Leapfrog
    from oracle.potentials import DiagPotential
    model, oracle = models_util.pairs()[name]
    rng = np.random.default_rng(43)
    D, C = oracle.ndim, 2
    var = rng.random(D) + 0.05
    q0 = rng.normal(size=(C, D)) * 0.3
    p0 = rng.normal(size=(C, D)) / np.sqrt(var)
    eps = 0.003 if name in ("hier", "stoch_vol") else 0.03
    eng = model.engine(C, dtype="float64")
    q1, p1, en = eng.leapfrog(q0, p0, var, eps, 7)
    q1, p1, en = q1.cpu().numpy(), p1.cpu().numpy(), en.cpu().numpy()
    for c in range(C):
        integ = Leapfrog(DiagPotential(var), oracle)
        s = integ.start(q0[c], p0[c])
        for _ in range(7):
            s = integ.step(eps, s)
        np.testing.assert_allclose(q1[c], s.q, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(p1[c], s.p, rtol=1e-9, atol=1e-10)
        assert abs(en[c] - s.energy) <= 1e-9 * max(1.0, abs(s.energy))
    eng.close()


# ------------------------------------------------------------------------------- HMC (R7) beyond the toy
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_hmc_on_glm_matches_oracle(dtype):
    """HamiltonianMC (hmc.py:110-152) on a non-trivial model: fp64 draw for draw against the oracle; fp32 the
    same decisions and positions for the first transitions (round-off separates trajectories later)."""
    from oracle.hmc_cpu import CpuHMC, run_chain
    from oracle.potentials import DiagAdaptPotential
    from oracle.rng import PhiloxRNG
    model, oracle = models_util.pairs()["glm"]
    D, C, n, tune = oracle.ndim, 3, 60, 40
    q0 = np.random.default_rng(44).uniform(-1, 1, size=(C, D)) * 0.3
    seeds = [61, 62, 63]
    eng = model.engine(C, dtype=dtype)
    eng.set_state(q0, seeds, 0.05, np.zeros(D), np.ones(D), 10.0)
    opts = dict(NUTS_OPTS)
    opts.update(target_accept=0.65, adapt_step_size=0, path_length=0.6, exec_mode=_capi.B2_EXEC_LOCKSTEP)
    out = {k: v.cpu().numpy() for k, v in eng.run(_capi.B2_HMC, n, tune, opts).items()}
    eng.close()
    for c in range(C):
        s = CpuHMC(oracle, D, DiagAdaptPotential(D, np.zeros(D), np.ones(D), 10), PhiloxRNG(seeds[c]),
                   adapt_step_size=False, path_length=0.6)
        s.adapter.log_step = s.adapter.log_bar = np.log(0.05)
        qs, st = run_chain(s, q0[c], n, tune)
        if dtype == "float64":
            assert (st["n_steps"] == out["n_steps"][:, c]).all()
            assert (st["accepted"] == out["accepted"][:, c].astype(bool)).all()
            assert np.abs(qs - out["q"][:, c]).max() < 1e-6
            assert np.abs(st["accept"] - out["accept"][:, c]).max() < 1e-6
        else:
            assert (st["n_steps"][:10] == out["n_steps"][:10, c]).all()
            assert (st["accepted"][:10] == out["accepted"][:10, c].astype(bool)).all()
            assert np.abs(qs[:10] - out["q"][:10, c]).max() < 2e-3
            assert np.abs(st["accept"][:10] - out["accept"][:10, c]).max() < 1e-3


# ------------------------------------------------------------------------------- posterior, tensor-core path
def test_tcgen05_lockstep_posterior_agrees_with_cpu_nuts_by_mcse_z_test():
    """BASELINE.json north star: posterior means and sds of the fp32 tcgen05 lock-step path (the path the headline
    number comes from) agree with CPU NUTS within an MCSE-based |z| < 4, on a C2-shaped model (N = 20 000,
    D = 100).  CPU arm: 16 oracle chains x 2500 draws, tests/golden/glm_c2shape_posterior.json (made by
    tests/golden/make_glm_posterior.py with the same seeded data)."""
    import pymc3_b200 as pm
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glm_c2shape_posterior.json")))
    X, y = models_util.glm_data(gold["n"], gold["k"], seed=gold["seed"])
    model = pm.LogisticGLM(X, y)
    C, D, tune, draws = 256, gold["k"] + 1, 500, 500
    q0 = np.random.default_rng(45).uniform(-1, 1, size=(C, D)) * 0.1
    eng = model.engine(C, dtype="float32")
    eng.set_state(q0, np.arange(C) + 8000, 0.25 / D ** 0.25, np.zeros(D), np.ones(D), 10.0)
    opts = dict(NUTS_OPTS)
    opts.update(exec_mode=_capi.B2_EXEC_LOCKSTEP, glm_path=_capi.B2_GLM_TCGEN05, hmc_jitter=0)
    out = eng.run(_capi.B2_NUTS, tune + draws, tune, opts)
    assert all(r.phase == _capi.PHASE_DONE for r in eng.reports())
    q = out["q"][tune:].permute(1, 0, 2).cpu().numpy().astype("f8")        # [C, draws, D]
    acc = out["mean_tree_accept"][tune:].mean().item()
    eng.close()
    assert 0.7 < acc < 0.92
    mean_c, sd_c, ess_c = np.array(gold["mean"]), np.array(gold["sd"]), np.array(gold["ess_bulk"])
    # conservative standard errors: NUTS draws are antithetic (bulk ESS above the number of draws); no credit for that
    ess_c = np.minimum(ess_c, gold["chains"] * gold["draws"])
    ess_g = np.minimum(np.ravel(pm.stats.ess(q)), C * draws)
    mean_g, sd_g = q.mean(axis=(0, 1)), q.reshape(-1, D).std(axis=0)
    assert float(np.ravel(pm.stats.rhat(q)).max()) < 1.02
    se_mean = np.sqrt(sd_g ** 2 / ess_g + sd_c ** 2 / ess_c)
    z = (mean_g - mean_c) / se_mean
    print("z(mean): mean square %.2f (1 = calibrated), max |z| %.2f at %d" % (np.mean(z ** 2), np.abs(z).max(), np.abs(z).argmax()))
    assert np.abs(z).max() < 4, (int(np.abs(z).argmax()), float(np.abs(z).max()), float(np.mean(z ** 2)))
    # sd: se(sd) ~ sd / sqrt(2 ESS) for a near-Gaussian marginal (tail ESS is lower than bulk: use half of it)
    se_sd = np.sqrt(sd_g ** 2 / ess_g + sd_c ** 2 / ess_c)       # >= the asymptotic se of an sd (sd / sqrt(2 ESS))
    z_sd = (sd_g - sd_c) / se_sd
    assert np.abs(z_sd).max() < 4, (int(np.abs(z_sd).argmax()), float(np.abs(z_sd).max()))
    print("z-test tcgen05 lock-step: max |z| mean %.2f, sd %.2f; min ESS gpu %.0f cpu %.0f" %
          (np.abs(z).max(), np.abs(z_sd).max(), ess_g.min(), ess_c.min()))
