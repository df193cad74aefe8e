"""Pins the oracle to the reference's own golden vectors."""
import json
import os

import numpy as np
import pytest

from oracle import densities as od
from oracle.hmc_cpu import CpuHMC, CpuNUTS, run_chain
from oracle.potentials import DiagAdaptPotential, DiagPotential, clip_precision
from oracle.rng import LegacyRNG


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "step_kat.json")) as f:
        return json.load(f)


def test_nuts_known_answer_trace(kat):
    """pymc3/tests/test_step.py:371-474, setup :505-527: NUTS(scaling=model.test_point) ->
    guess_scaling: diag Hessian 2 -> QuadPotentialDiag(1/2); 100 tuning draws, random_seed=1."""
    model = od.NormalPair()
    pot = DiagPotential(1.0 / clip_precision(np.array([2.0])))
    sampler = CpuNUTS(model, 1, pot, LegacyRNG())
    qs, stats = run_chain(sampler, [0.0], 100, 100, seed=1)
    assert np.abs(qs[:, 0] - np.array(kat["NUTS"])).max() < 1e-8
    assert set(stats) == {"depth", "diverging", "energy", "energy_error", "model_logp", "max_energy_error",
                          "mean_tree_accept", "step_size", "step_size_bar", "tree_size", "tune"}   # test_step.py:980-1008
    assert np.allclose(stats["model_logp"], [model.logp(q) for q in qs])


def test_hmc_known_answer_trace(kat):
    """pymc3/tests/test_step.py:163-266: HamiltonianMC() default potential QuadPotentialDiagAdapt(1,0,1,10)."""
    model = od.NormalPair()
    pot = DiagAdaptPotential(1, np.zeros(1), np.ones(1), 10)
    sampler = CpuHMC(model, 1, pot, LegacyRNG())
    qs, _ = run_chain(sampler, [0.0], 100, 100, seed=1)
    assert np.abs(qs[:, 0] - np.array(kat["HamiltonianMC"])).max() < 1e-8


def test_developer_guide_logp_dlogp(golden_dir):
    """docs/source/developer_guide.rst:151-155, 572-575, 715-737."""
    with open(os.path.join(golden_dir, "devguide_logp.json")) as f:
        d = json.load(f)
    logp, grad = od.DevGuideModel()(np.array(d["z"] + d["x"]))
    assert abs(logp - d["logp"]) < 5e-8
    assert np.abs(grad - np.array(d["dlogp"])).max() < 5e-8
    z = d["scalar_model"]["z"]
    x_logp = od.normal_logp(5.0, z, 1.0)
    assert abs(x_logp - d["scalar_model"]["x_logp"]) < 5e-7      # the guide prints 8 significant digits
    assert abs(x_logp + od.normal_logp(z, 0.0, 5.0) - d["scalar_model"]["model_logp"]) < 5e-7


def test_value_grad_function_vector():
    """pymc3/tests/test_model.py:299-304: cost = extra1*val1.sum() + val2.sum() at ones -> 21, [5,5,5,1,...]."""
    extra1, val1, val2 = 5.0, np.ones(3), np.ones((2, 3))
    assert extra1 * val1.sum() + val2.sum() == 21
    # :320-337 (edge case #2948): Lognormal(0, tau=1)[3] + HalfCauchy(10) at the test point: gradient 0
    u = np.zeros(3)           # sigma_log__ test value = log(median=1) = 0
    g_lognormal = -u          # d/du [ -u^2/2 - u (pdf) + u (jacobian) ]
    nu = 10.0                 # HalfCauchy test value = beta
    w = (nu / 10.0) ** 2
    g_hc = 1.0 - 2.0 * w / (1.0 + w)
    assert np.allclose(np.concatenate([g_lognormal, [g_hc]]), 0.0, atol=1e-5)
