"""The engine's state machine (pymc3_b200/csrc/b2_core.cuh, b2_models.cuh), compiled for the CPU
with a one-lane thread group (tests/hostsim), against the recursive oracle.  This is how the
iterative tree builder, the Philox stream and the adaptation logic are verified without a GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle.hmc_cpu import CpuHMC, CpuNUTS, run_chain
from oracle.potentials import DiagAdaptPotential, DiagPotential
from oracle.rng import PhiloxRNG, philox4x32_10
from tests import hostsim_util as hs
from tests import models_util


def test_philox_matches_numpy_twin():
    rng = np.random.default_rng(0)
    out = (C.c_uint32 * 4)()
    for _ in range(50):
        c = [int(v) for v in rng.integers(0, 2 ** 32, size=4)]
        k = [int(v) for v in rng.integers(0, 2 ** 32, size=2)]
        hs.lib().hostsim_philox(*c, *k, out)
        assert tuple(out) == philox4x32_10(c, k)
    # known-answer vector of Random123 (Salmon et al.): philox4x32-10, counter=key=0
    hs.lib().hostsim_philox(0, 0, 0, 0, 0, 0, out)
    assert tuple(out) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    r = PhiloxRNG(2 ** 40 + 17)
    z = r.momentum(5, 9)
    z2 = [hs.lib().hostsim_normal(r.key[0], r.key[1], 5, i) for i in range(9)]
    assert np.abs(z - np.array(z2)).max() < 1e-14


@pytest.mark.parametrize("name", ["std_normal", "eight_schools", "glm", "hier", "stoch_vol"])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 2e-5)])
def test_logp_dlogp_evaluators(name, dtype, tol):
    model, oracle = models_util.pairs()[name]
    q = np.random.default_rng(5).normal(size=(4, oracle.ndim)) * 0.4
    lp, g = hs.logp_dlogp(model, q, dtype=dtype)
    for i in range(len(q)):
        l0, g0 = oracle(q[i].astype(dtype).astype("f8"))
        assert abs(lp[i] - l0) <= tol * max(1.0, abs(l0))
        assert np.abs(g[i] - g0).max() <= tol * max(1.0, np.abs(g0).max())


@pytest.mark.parametrize("name,step0,n", [("std_normal", None, 400), ("eight_schools", None, 400),
                                          ("glm", None, 300), ("hier", 0.02, 150), ("stoch_vol", None, 60)])
def test_iterative_nuts_equals_recursive_reference_logic(name, step0, n):
    """fixed step size, adaptive mass matrix (window swap at draw 102), tuning stops at 0.7 n."""
    model, oracle = models_util.pairs()[name]
    D = oracle.ndim
    tune = int(0.7 * n)
    q0 = np.random.default_rng(1).uniform(-1, 1, size=(2, D))
    seeds = [11, 2 ** 40 + 5]
    res = hs.run(model, q0, seeds, n, tune, adapt_step_size=0, step_size0=step0)
    for c in range(2):
        s = CpuNUTS(oracle, D, DiagAdaptPotential(D, np.zeros(D), np.ones(D), 10), PhiloxRNG(seeds[c]),
                    adapt_step_size=False)
        if step0:
            s.adapter.log_step = s.adapter.log_bar = np.log(step0)
        qs, st = run_chain(s, q0[c], n, tune)
        assert (st["depth"] == res["depth"][:, c]).all()
        assert (st["tree_size"] == res["tree_size"][:, c]).all()
        assert (st["diverging"] == res["diverging"][:, c].astype(bool)).all()
        assert (st["tune"] == res["tune"][:, c].astype(bool)).all()
        assert np.abs(qs - res["q"][:, c]).max() < 1e-8
        for key in ("energy", "energy_error", "max_energy_error", "mean_tree_accept", "model_logp", "step_size"):
            assert np.allclose(st[key], res[key][:, c], rtol=1e-7, atol=1e-7), key
        assert res["reports"][c].n_grad == st["tree_size"].sum() + 1
        assert res["reports"][c].n_maxdepth_post == s.n_max_depth_after_tune
        assert res["reports"][c].n_div_post == s.n_diverging_after_tune


def test_dual_averaging_and_early_treedepth():
    """with step-size adaptation round-off is amplified by the tuning dynamics: compare the start"""
    model, oracle = models_util.pairs()["eight_schools"]
    q0 = np.random.default_rng(2).uniform(-1, 1, size=(1, 10))
    res = hs.run(model, q0, [7], 40, 40)
    s = CpuNUTS(oracle, 10, DiagAdaptPotential(10, np.zeros(10), np.ones(10), 10), PhiloxRNG(7))
    qs, st = run_chain(s, q0[0], 40, 40)
    assert (st["depth"][:25] == res["depth"][:25, 0]).all()
    assert np.abs(qs[:15] - res["q"][:15, 0]).max() < 1e-7
    assert np.allclose(st["step_size"][:15], res["step_size"][:15, 0], rtol=1e-7)
    assert np.allclose(st["step_size_bar"][:15], res["step_size_bar"][:15, 0], rtol=1e-7)
    assert res["depth"].max() <= 8           # early_max_treedepth while tuning and iter < 200


def test_static_diag_potential_and_max_treedepth():
    model, oracle = models_util.pairs()["std_normal"]
    var = np.array([0.5, 2.0, 1.0, 4.0, 9.0])
    q0 = np.zeros((1, 5))
    res = hs.run(model, q0, [3], 60, 0, adapt_step_size=0, adapt_mass=0, mass_var=var, step_size0=0.01,
                 max_treedepth=4)
    s = CpuNUTS(oracle, 5, DiagPotential(var), PhiloxRNG(3), adapt_step_size=False, max_treedepth=4)
    s.adapter.log_step = s.adapter.log_bar = np.log(0.01)
    qs, st = run_chain(s, q0[0], 60, 0)
    assert np.abs(qs - res["q"][:, 0]).max() < 1e-10
    assert (res["depth"] == 4).all() and res["reports"][0].n_maxdepth_post == 60 == s.n_max_depth_after_tune
    assert not res["tune"].any()


def test_hmc_equals_reference_logic():
    model, oracle = models_util.pairs()["std_normal"]
    D = oracle.ndim
    q0 = np.random.default_rng(3).uniform(-1, 1, size=(2, D))
    seeds = [21, 22]
    res = hs.run(model, q0, seeds, 300, 200, kind="hmc", adapt_step_size=0)
    for c in range(2):
        s = CpuHMC(oracle, D, DiagAdaptPotential(D, np.zeros(D), np.ones(D), 10), PhiloxRNG(seeds[c]),
                   adapt_step_size=False)
        qs, st = run_chain(s, q0[c], 300, 200)
        assert (st["n_steps"] == res["n_steps"][:, c]).all()
        assert (st["accepted"] == res["accepted"][:, c].astype(bool)).all()
        assert np.abs(qs - res["q"][:, c]).max() < 1e-8
        assert np.allclose(st["accept"], res["accept"][:, c], atol=1e-9)
        assert np.allclose(st["energy"], res["energy"][:, c], rtol=1e-9)


def test_bad_initial_energy_and_fp32_sanity():
    from pymc3_b200 import model as pm
    res = hs.run(pm.StdNormal(3), np.array([[np.inf, 0.0, 0.0]]), [1], 5, 5)
    assert res["reports"][0].phase == 4 and res["reports"][0].fail_code == 1
    model = pm.EightSchoolsNCP(mu_sd=5.0, tau_beta=5.0)
    q0 = np.random.default_rng(4).uniform(-1, 1, size=(32, 10))
    res = hs.run(model, q0, np.arange(32) + 100, 500, 250, dtype="float32")
    mu, tau = res["q"][250:, :, 8], np.exp(res["q"][250:, :, 9])
    assert abs(mu.mean() - 4.46) < 0.4 and abs(tau.mean() - 3.59) < 0.45      # published table, SURVEY section 6
    assert 0.7 < res["mean_tree_accept"][250:].mean() < 0.92
