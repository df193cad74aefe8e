"""CPU-tier checks of the boundary and the host logic (no compute calls without a GPU)."""
import ast
import ctypes
import os
import re

import numpy as np
import pytest

import pymc3_b200 as pm
from pymc3_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "b200nuts.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z_0-9]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from pymc3_b200 import build
    build.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = _header_functions()
    assert "b2_sample_run" in names and "b2_logp_dlogp" in names and len(names) >= 12
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_capi.EXPORTS)
    assert _capi.load_library().b2_abi_version() == _capi.ABI_VERSION == 4


def test_ctypes_structs_match_header_layout():
    """sizes the C compiler computes for the ABI structs (LP64): guards against field drift."""
    assert ctypes.sizeof(_capi.ModelDesc) == 16 + 6 * 8 + 32
    assert ctypes.sizeof(_capi.SamplerOpts) == 5 * 4 + 4 + 5 * 8 + 2 * 4 + 8 + 5 * 4 + 4
    assert ctypes.sizeof(_capi.TraceOut) == 15 * 8
    assert ctypes.sizeof(_capi.ChainReport) == 6 * 4 + 8 + 16


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pm.StdNormal(2):
        with pytest.raises(_capi.B2Error) as err:
            pm.sample(5, tune=5, chains=2)
    assert "no CPU fallback" in str(err.value)


def test_product_never_imports_the_oracle_or_the_host_simulator():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pymc3_b200")):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    mods = [node.module or ""]
                for m in mods:
                    assert not m.startswith("oracle") and "hostsim" not in m and not m.startswith("tests"), (f, m)
    for f in os.listdir(os.path.join(ROOT, "pymc3_b200", "csrc")):
        if f.endswith(".cu"):
            assert "B2HostGroup" not in open(os.path.join(ROOT, "pymc3_b200", "csrc", f)).read()


def test_model_layout_and_names():
    m = pm.EightSchoolsNCP()
    assert m.free_RVs == ["eta", "mu", "tau_log__"] and m.unobserved_RVs == ["eta", "mu", "tau_log__", "tau"]
    od = m.ordering()
    assert od.size == 10 and od["mu"].slc == slice(8, 9) and od["eta"].shp == (8,)
    arr = np.arange(10.0)
    assert np.array_equal(m.dict_to_array(m.array_to_dict(arr)), arr)
    full = m.expand(np.zeros((3, 4, 10)))
    assert full["eta"].shape == (3, 4, 8) and full["tau"].shape == (3, 4) and np.all(full["tau"] == 1.0)
    g = pm.LogisticGLM(np.zeros((5, 3)), np.zeros(5))
    assert g.free_RVs == ["Intercept", "x0", "x1", "x2"]            # glm/linear.py:72-99
    with pytest.raises(TypeError):
        pm.LogisticGLM(np.zeros((5, 3)), np.zeros((5, 1)))
    h = pm.HierLinearNCP([2, 0, 1, 2, 0], [0, 1, 0, 1, 1], [1.0, 2.0, 3.0, 4.0, 5.0], 3)
    assert h.ndim == 11 and list(h.grp_off) == [0, 2, 3, 5]
    assert list(h.y) == [2.0, 5.0, 3.0, 1.0, 4.0] and list(h.floor) == [1, 1, 0, 0, 1]
    sv = pm.StochVol()
    assert sv.ndim == 2907 and sv.free_RVs == ["step_size_log__", "volatility", "nu_log__"]
    with pytest.raises(TypeError):
        pm.modelcontext(None)


def test_step_method_surface_without_gpu():
    with pm.StdNormal(4):
        nuts = pm.NUTS()
        hmc = pm.HamiltonianMC()
    assert nuts.name == "nuts" and hmc.name == "hmc" and nuts.generates_stats and nuts.default_blocked
    assert abs(nuts.step_size - 0.25 / 4 ** 0.25) < 1e-15           # base_hmc.py:93
    assert nuts.target_accept == 0.8 and hmc.target_accept == 0.65 and hmc.path_length == 2.0
    assert isinstance(nuts.potential, pm.QuadPotentialDiagAdapt)
    assert set(nuts.stats_dtypes[0]) == {"depth", "step_size", "tune", "mean_tree_accept", "step_size_bar",
                                         "tree_size", "diverging", "energy_error", "energy", "max_energy_error",
                                         "model_logp"}
    assert nuts.stats_dtypes[0]["tree_size"] is np.float64 and nuts.stats_dtypes[0]["depth"] is np.int64
    assert nuts.tune is True
    nuts.stop_tuning()
    assert nuts.tune is False
    from pymc3_b200.step_methods import Competence

    class V:
        dtype = np.dtype("float64")
    assert pm.NUTS.competence(V, True) == Competence.IDEAL and pm.NUTS.competence(V, False) == Competence.INCOMPATIBLE
    assert pm.HamiltonianMC.competence(V, True) == Competence.COMPATIBLE
    o = nuts._opts()
    assert o["max_treedepth"] == 10 and o["early_max_treedepth"] == 8 and o["Emax"] == 1000.0 and o["adapt_mass"] == 1
    assert hmc._opts()["hmc_jitter"] == 1 and nuts._opts()["hmc_jitter"] == 0
    with pm.StdNormal(4):
        with pytest.raises(ValueError):
            pm.NUTS(bogus=1)
        with pytest.raises(ValueError):
            pm.NUTS(max_treedepth=20)


def test_quadpotential_protocol():
    """tests/test_quadpotential.py:25-135 for the diagonal family."""
    with pytest.raises(ValueError):
        pm.quad_potential(np.array([1.0, -1.0]), True)
    pot = pm.quad_potential(np.array([0.25, 4.0]), False)              # precision -> variance
    assert np.allclose(pot.v, [4.0, 0.25])
    x = np.array([1.0, 2.0])
    v = pot.velocity(x)
    assert np.allclose(v, [4.0, 0.5]) and np.isclose(pot.energy(x), 0.5 * (4.0 + 1.0))
    out = np.empty(2)
    assert np.isclose(pot.velocity_energy(x, out), pot.energy(x, velocity=v)) and np.allclose(out, v)
    np.random.seed(0)
    draws = np.array([pot.random() for _ in range(4000)])
    assert np.allclose(draws.std(axis=0), 1 / np.sqrt(pot.v), rtol=0.06)
    adapt = pm.QuadPotentialDiagAdapt(2, np.zeros(2), np.array([1.0, 2.0]), 10)
    init = adapt.device_init()
    assert init["weight"] == 10 and init["window"] == 101 and init["adapt"] == 1
    with pytest.raises(ValueError):
        pm.QuadPotentialDiagAdapt(2, np.zeros(3))
    adapt.sync(np.array([1.0, 0.0]))
    from pymc3_b200.model import VarMap
    with pytest.raises(ValueError) as err:
        adapt.raise_ok([VarMap("a", slice(0, 1), (), "f8"), VarMap("b", slice(1, 2), (), "f8")])
    assert "The derivative of RV `b`.ravel()[0] is zero." in str(err.value)


def test_sample_argument_normalisation_needs_no_gpu():
    with pm.StdNormal(2):
        with pytest.raises(TypeError):
            pm.sample(5, tune=5, chains=2, random_seed=0.5)
        with pytest.raises(ValueError):
            pm.sample(5, tune=5, chains=2, random_seed=[1, 2, 3])
        with pytest.raises(ValueError):
            pm.sample(5, tune=5, chains=2, start={"x": np.zeros(3)})
        with pytest.raises(ValueError):
            pm.sample(0, tune=0, chains=2)
        with pytest.raises(NotImplementedError):
            pm.init_nuts(init="advi")
        with pytest.raises(ValueError):
            pm.init_nuts(init="nonsense")
        np.random.seed(3)
        start, step = pm.init_nuts(init="jitter+adapt_diag", chains=3)
        assert len(start) == 3 and all(np.all(np.abs(s["x"]) <= 1) for s in start)
        assert np.allclose(step.potential._initial_mean, np.mean([s["x"] for s in start], axis=0))
        assert step.potential._initial_weight == 10                     # sampling.py:1929


def test_ndarray_backend_and_multitrace_contract(tmp_path):
    """tests/backend_fixtures.py selection / slicing / stats contract on host data."""
    model = pm.EightSchoolsNCP()
    rng = np.random.default_rng(0)
    straces = []
    for c in range(3):
        q = rng.normal(size=(20, 10))
        st = pm.NDArray.from_arrays(model, c, model.expand(q), {"depth": np.arange(20), "tune": np.arange(20) < 5})
        straces.append(st)
    mt = pm.MultiTrace(straces)
    assert mt.nchains == 3 and len(mt) == 20 and mt.chains == [0, 1, 2]
    assert mt["eta"].shape == (60, 8) and mt.get_values("eta", combine=False)[1].shape == (20, 8)
    assert mt.get_values("mu", chains=1).shape == (20,) and mt.get_values("mu", burn=5, thin=5).shape == (9,)
    assert mt.get_sampler_stats("depth").shape == (60,) and mt.stat_names == {"depth", "tune"}
    assert mt.get_sampler_stats("tree_depth", chains=[0]).shape == (20,)
    assert len(mt[5:]) == 15 and mt[5:].get_sampler_stats("depth", chains=0)[0] == 5
    assert mt.mu.shape == (60,) and mt.depth.shape == (60,)
    with pytest.raises(KeyError):
        mt["nope"]
    with pytest.raises(ValueError):
        pm.MultiTrace([straces[0], straces[0]])
    assert len(list(mt.points([0]))) == 20
    d = pm.save_trace(mt, str(tmp_path / "tr"))
    with pytest.raises(OSError):
        pm.save_trace(mt, d)
    back = pm.load_trace(d, model=model)
    assert np.array_equal(back["tau"], mt["tau"]) and np.array_equal(back.depth, mt.depth)
    other = pm.MultiTrace([pm.NDArray.from_arrays(model, 7, model.expand(rng.normal(size=(20, 10))),
                                                  {"depth": np.arange(20), "tune": np.arange(20) < 5})])
    merged = pm.merge_traces([mt, other])
    assert merged.nchains == 4 and merged.chains[-1] == 7
    # draw-at-a-time recording used by the step() loop
    nd = pm.NDArray(model=model)
    nd.setup(3, 0, [{"depth": np.int64}])
    for i in range(2):
        nd.record(model.test_point, [{"depth": i}])
    nd.close()
    assert len(nd) == 2 and nd.get_values("tau").shape == (2,) and nd.get_sampler_stats("depth")[1] == 1


def test_diagnostics():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(4, 1000, 2))
    e = pm.ess(x)
    assert e.shape == (2,) and np.all(e > 3000) and np.all(e < 5200)
    assert np.all(np.abs(pm.rhat(x) - 1) < 0.01)
    y = np.zeros((4, 2000))
    for t in range(1, 2000):
        y[:, t] = 0.9 * y[:, t - 1] + rng.normal(size=4) * np.sqrt(0.19)
    assert 250 < pm.ess(y) < 700                                    # theory: 8000 * 0.1/1.9 = 421
    assert pm.rhat(y + np.arange(4)[:, None] * 3) > 1.5            # chains stuck at different places
    assert abs(pm.stats.mcse_mean(x)[0] - 1 / np.sqrt(4000)) < 0.004
    assert pm.stats.bfmi(rng.normal(size=(2, 500))).shape == (2,)


def test_host_driven_transitions_sample_the_target():
    """pymc3_b200/step_methods/hmc/host_transition.py (the path of user potentials) on a NumPy density: stationary
    standard deviations of a diagonal normal, NUTS and HMC, dense and diagonal potentials."""
    from pymc3_b200.step_methods.hmc import host_transition as ht
    from pymc3_b200.step_methods.hmc.quadpotential import QuadPotentialDiag, QuadPotentialFull
    sig = np.array([1.0, 2.0, 0.5])

    def f(q):
        return -0.5 * np.sum((q / sig) ** 2), -q / sig ** 2

    np.random.seed(4)
    for pot in (QuadPotentialDiag(np.ones(3)), QuadPotentialFull(np.diag(sig ** 2) + 0.05)):
        integ = ht.HostIntegrator(pot, f)
        q, draws, acc = np.zeros(3), [], []
        for _ in range(3000):
            q, _, st = ht.nuts_transition(integ, integ.start(q, pot.random()), 0.35, 10, 1000.0)
            draws.append(q)
            acc.append(st["mean_tree_accept"])
        assert np.allclose(np.std(draws, axis=0), sig, rtol=0.1) and 0.6 < np.mean(acc) <= 1.0
    integ = ht.HostIntegrator(QuadPotentialDiag(np.ones(3)), f)
    q, draws = np.zeros(3), []
    for _ in range(3000):
        q, _, st = ht.hmc_transition(integ, integ.start(q, integ.pot.random()), 0.3, 2.0, 1024, 1000.0)
        draws.append(q)
    assert np.allclose(np.std(draws, axis=0), sig, rtol=0.1)


def test_sample_with_an_adaptive_dense_potential_runs_host_driven_chains():
    """sample() with a step method that is driven from the host (QuadPotentialFullAdapt, quadpotential.py:482-572;
    tests/test_quadpotential.py:274-290): sequential chains, per-draw record, sampler stats; the density here is a
    NumPy stand-in for the device ValueGradFunction."""
    from pymc3_b200.step_methods.hmc.quadpotential import QuadPotentialFullAdapt
    sig = np.array([1.0, 2.0, 0.5])

    class FakeVG:
        size, dtype = 3, np.dtype("f8")

        def __call__(self, q, grad_out=None):
            return np.array(-0.5 * np.sum((q / sig) ** 2)), -q / sig ** 2

    class Model(pm.StdNormal):
        def logp_dlogp_function(self, *a, **k):
            return FakeVG()

    seen = []
    with Model(3, sigma=sig):
        with pytest.warns(UserWarning):
            pot = QuadPotentialFullAdapt(3, np.zeros(3), np.diag(sig ** 2) + 0.05, 5)
        step = pm.NUTS(potential=pot, dtype="float64")
        assert not step._batched
        dense = pm.NUTS(scaling=np.diag(sig ** 2) + 0.05, is_cov=True, dtype="float64")
        assert type(dense.potential).__name__ == "QuadPotentialFull" and dense._batched      # static dense: device path
        trace = pm.sample(400, tune=200, chains=2, step=step, random_seed=11, compute_convergence_checks=False,
                          callback=lambda trace, draw: seen.append(draw.draw_idx))
    x = trace["x"]
    assert x.shape == (800, 3) and np.allclose(x.std(axis=0), sig, rtol=0.2) and np.abs(x.mean(axis=0)).max() < 0.4
    assert len(seen) == 2 * 600 and 0.6 < trace.get_sampler_stats("mean_tree_accept").mean() < 0.99
    assert np.isfinite(trace.get_sampler_stats("step_size")).all()


def test_running_covariance_and_adaptive_dense_potential():
    """tests/test_quadpotential.py:158-271 restated for the host classes: the Welford co-moment accumulator equals
    numpy's estimates (also when seeded with pseudo-observations), momenta drawn from the potential have covariance
    M, the update window / adaptation window bookkeeping, the not-invertible error and the experimental warning."""
    from pymc3_b200.step_methods.hmc import quadpotential as qp
    rng = np.random.default_rng(5432)
    L = np.tril(rng.normal(size=(10, 10)))
    L[np.diag_indices(10)] = np.exp(np.diag(L))
    samples = rng.multivariate_normal(rng.normal(size=10), L @ L.T, size=100)
    est = qp._WeightedCovariance(10)
    for x in samples:
        est.add_sample(x, 1)
    assert np.allclose(est.current_mean(), samples.mean(axis=0)) and np.allclose(est.current_covariance(), np.cov(samples, rowvar=0))
    est2 = qp._WeightedCovariance(10, samples[:10].mean(axis=0), np.cov(samples[:10], rowvar=0, bias=True), 10)
    for x in samples[10:]:
        est2.add_sample(x, 1)
    assert np.allclose(est2.current_mean(), samples.mean(axis=0)) and np.allclose(est2.current_covariance(), np.cov(samples, rowvar=0))

    np.random.seed(4566)
    m = np.array([[3.0, -2.0], [-2.0, 4.0]])
    var = np.array([[2 * m[0, 0], m[1, 0] ** 2 + m[1, 1] * m[0, 0]], [m[0, 1] ** 2 + m[1, 1] * m[0, 0], 2 * m[1, 1]]])
    with pytest.warns(UserWarning):
        pot = qp.QuadPotentialFullAdapt(2, np.zeros(2), np.linalg.inv(m), 1)
    draws = np.array([pot.random() for _ in range(1000)])
    assert np.all(np.abs(m - np.cov(draws, rowvar=0)) < 5 * np.sqrt(var / 1000))       # Wishart spread

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        init = np.array([[1.0, 0.02], [0.02, 0.8]])
        pot = qp.QuadPotentialFullAdapt(2, np.zeros(2), init, 1, update_window=50)
        for _ in range(49):
            pot.update(np.random.randn(2), None, True)
        assert np.allclose(pot._cov, init)
        pot.update(np.random.randn(2), None, True)
        assert not np.allclose(pot._cov, init)
        pot = qp.QuadPotentialFullAdapt(2, np.zeros(2), np.eye(2), 1, adaptation_window=10)
        for _ in range(11):
            pot.update(np.random.randn(2), None, True)
        assert pot._previous_update == 10 and pot.adaptation_window == 10 * pot.adaptation_window_multiplier
        pot = qp.QuadPotentialFullAdapt(2, np.zeros(2), np.eye(2), 0, adaptation_window=10)
        for _ in range(11):
            pot.update(np.ones(2), None, True)
        with pytest.raises(ValueError):
            pot.raise_ok(None)
        # gradient-based diagonal adaptation (quadpotential.py:272-310): after 150 draws var = (n / sum |grad|)^2
        g = qp.QuadPotentialDiagAdaptGrad(2, np.zeros(2), np.ones(2), 10)
        for _ in range(300):
            g.update(np.random.randn(2), np.array([2.0, -0.5]), True)
        assert np.allclose(g._var, [0.25, 4.0], rtol=0.05)       # (the windows restart from one pseudo-gradient of 1)


def test_step_lists_are_unwrapped_or_refused():
    """sampling.py:142-165: a list of step methods becomes a CompoundStep; one HMC-family step is just that step"""
    with pm.StdNormal(2):
        with pytest.raises(NotImplementedError, match="CompoundStep"):
            pm.sample(5, tune=0, chains=1, step=[object(), object()])
        with pytest.raises(NotImplementedError, match="NUTS / HamiltonianMC"):
            pm.sample(5, tune=0, chains=1, step=[object()])


def test_trace_cov_matches_numpy():
    """pymc3/tuning/scaling.py:113-141"""
    rng = np.random.default_rng(0)
    tr = {"a": rng.normal(size=(200, 2)), "b": rng.normal(size=200)}
    tr["b"] = tr["b"] + tr["a"][:, 0]

    class M:
        free_RVs = ["a", "b"]
    flat = np.column_stack([tr["a"], tr["b"]])
    assert np.allclose(pm.trace_cov(tr, model=M()), np.cov(flat.T))

