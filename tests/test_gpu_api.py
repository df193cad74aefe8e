"""The reference-facing Python surface on a real GPU.  These read like the reference's own tests:
pymc3/tests/test_step.py:943-1008, test_hmc.py:49-66, test_sampling.py:41-236,
test_ndarray_backend.py, sampler_fixtures.py:24-56,150-176."""
import numpy as np
import pytest

import pymc3_b200 as pm
from pymc3_b200.exceptions import SamplingError

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def schools_trace():
    with pm.EightSchoolsNCP(mu_sd=5.0, tau_beta=5.0) as model:
        trace = pm.sample(500, tune=500, chains=8, random_seed=3, step=pm.NUTS(dtype="float64"))
    return model, trace


def test_sampler_stats_names_shapes_and_model_logp(schools_trace):
    """test_step.py:980-1008."""
    model, trace = schools_trace
    expected = {"depth", "diverging", "energy", "energy_error", "model_logp", "max_energy_error",
                "mean_tree_accept", "step_size", "step_size_bar", "tree_size", "tune"}
    assert trace.stat_names == expected
    for name in expected:
        assert trace.get_sampler_stats(name, chains=0).shape == (500,)
    assert trace.get_sampler_stats("tree_depth").shape == (8 * 500,)
    f = model.logp_dlogp_function(dtype="float64")
    model_logp = trace.get_sampler_stats("model_logp", chains=2)
    for i in (0, 17, 499):
        q = model.dict_to_array({k: v for k, v in trace.point(i, chain=2).items() if k in model.free_RVs})
        assert abs(float(f(q)[0]) - model_logp[i]) < 1e-8 * abs(model_logp[i])


def test_trace_selection_api(schools_trace):
    model, trace = schools_trace
    assert trace.nchains == 8 and len(trace) == 500
    assert set(trace.varnames) == {"eta", "mu", "tau_log__", "tau"}
    assert trace["eta"].shape == (4000, 8)
    assert trace.get_values("mu", combine=False)[3].shape == (500,)
    assert trace.get_values("mu", burn=100, thin=2, chains=[0, 1]).shape == (400,)
    assert np.allclose(trace["tau"], np.exp(trace["tau_log__"]))
    assert len(trace[100:]) == 400 and len(trace[::5]) == 100
    assert trace["mu", 100::2].shape == (8 * 200,)
    assert set(trace.point(3).keys()) == set(trace.varnames)
    assert trace.report.n_tune == 500 and trace.report.n_draws == 500 and trace.report.t_sampling > 0
    assert not trace.get_sampler_stats("tune").any()


def test_posterior_matches_published_table(schools_trace):
    """SURVEY section 6 (Diagnosing_biased_Inference_with_Divergences.ipynb:1566-1575), MCSE-based z < 4."""
    # tau's posterior is heavy-tailed (half-Cauchy prior): its sample sd over 8 x 500 draws moves by +-1 from
    # seed to seed, so the table is checked on 64 chains (32 000 draws) and the 8-chain fixture only below
    with pm.EightSchoolsNCP(mu_sd=5.0, tau_beta=5.0):
        big = pm.sample(500, tune=500, chains=64, random_seed=5, step=pm.NUTS(dtype="float64"),
                        compute_convergence_checks=False)
    model, trace = schools_trace
    pub = {"mu": (4.46, 3.31), "tau": (3.59, 3.23)}
    for name, (mean, sd) in pub.items():
        draws = np.stack(big.get_values(name, combine=False))
        mcse = pm.stats.mcse_mean(draws)
        assert abs(draws.mean() - mean) < 4 * mcse + 0.15          # 0.15: the table's own MC error
        assert abs(draws.std() - sd) < 0.5
    acc = trace.get_sampler_stats("mean_tree_accept")
    assert abs(acc.mean() - 0.8) < 0.08                             # sampler_fixtures.py:174-176
    assert float(np.max(pm.rhat(np.stack(trace.get_values("mu", combine=False))))) < 1.05


def test_discard_tuned_samples_and_lengths():
    """test_sampling.py:138-147."""
    with pm.StdNormal(3):
        t1 = pm.sample(50, tune=30, chains=2, random_seed=1, discard_tuned_samples=True,
                       compute_convergence_checks=False)
        t2 = pm.sample(50, tune=30, chains=2, random_seed=1, discard_tuned_samples=False,
                       compute_convergence_checks=False)
    assert len(t1) == 50 and len(t2) == 80
    assert t2.get_sampler_stats("tune", chains=0)[:30].all() and not t2.get_sampler_stats("tune", chains=0)[30:].any()
    assert np.array_equal(t1["x"], t2[30:]["x"])                    # same seeds -> same chains


def test_seeds_reproduce_and_differ():
    """test_sampling.py:46-73."""
    with pm.StdNormal(2):
        a = pm.sample(20, tune=20, chains=2, random_seed=7, compute_convergence_checks=False)
        b = pm.sample(20, tune=20, chains=2, random_seed=7, compute_convergence_checks=False)
        c = pm.sample(20, tune=20, chains=2, random_seed=8, compute_convergence_checks=False)
        d = pm.sample(20, tune=20, chains=2, random_seed=[11, 12], compute_convergence_checks=False)
    assert np.array_equal(a["x"], b["x"])
    assert not np.array_equal(a["x"], c["x"])
    assert d.nchains == 2


def test_bad_arguments():
    """test_sampling.py:99-111, 175-192."""
    with pm.StdNormal(2) as model:
        with pytest.raises(ValueError):
            pm.sample(10, tune=5, chains=1, step=pm.NUTS(), foo=1)
        with pytest.raises(ValueError):
            pm.NUTS(foo=1)
        with pytest.raises(ValueError):
            pm.sample(10, tune=5, chains=2, start={"x": np.zeros(5)})
        with pytest.raises(TypeError):
            pm.sample(10, tune=5, chains=2, random_seed=1.5)
        with pytest.raises(ValueError):
            pm.NUTS(scaling=np.ones(2), potential=pm.QuadPotentialDiag(np.ones(2)))


def test_bad_initial_energy_raises_sampling_error():
    """test_hmc.py:59-66, test_step.py:943-956."""
    with pm.StdNormal(2):
        with pytest.raises(SamplingError) as err:
            pm.sample(10, tune=5, chains=2, start={"x": np.array([np.inf, 0.0])}, step=pm.NUTS(),
                      compute_convergence_checks=False)
        assert "Bad initial energy" in str(err.value)


def test_tuning_reset_step_size_constant_after_tune():
    """test_hmc.py:49-57."""
    with pm.StdNormal(4, sigma=[1.0, 2.0, 0.5, 3.0]):
        step = pm.NUTS()
        trace = pm.sample(100, tune=200, chains=2, step=step, discard_tuned_samples=False, random_seed=2,
                          compute_convergence_checks=False)
    assert step.tune is False
    ss = trace.get_sampler_stats("step_size", chains=0)
    assert np.all(ss[200:] == ss[200]) and len(np.unique(ss[:200])) > 50
    assert np.allclose(step.potential._var, [1.0, 4.0, 0.25, 9.0], rtol=0.6)


def test_divergences_are_reported():
    """test_step.py:958-978: a funnel-like posterior sampled with a too-large fixed step."""
    with pm.EightSchoolsNCP(tau_beta=25.0):
        step = pm.NUTS(adapt_step_size=False, step_scale=6.0)
        trace = pm.sample(200, tune=0, chains=2, step=step, random_seed=5, compute_convergence_checks=False)
    assert trace.get_sampler_stats("diverging").any()
    msgs = [w.message for w in trace.report._warnings]
    assert any("diverg" in m.lower() for m in msgs)       # "N divergences" or "only diverging samples"
    assert not trace.report.ok
    with pytest.raises(ValueError):
        trace.report.raise_ok()


def test_step_interface_single_chain_loop():
    """step.step(point) one draw at a time (sampling.py:914-936) + iter_sample (test_sampling.py:113)."""
    with pm.StdNormal(3) as model:
        step = pm.NUTS(dtype="float64")
        np.random.seed(4)
        point = model.test_point
        xs = []
        for i in range(60):
            if i == 40:
                step.stop_tuning()
            point, stats = step.step(point)
            assert set(stats[0]) == set(step.stats_dtypes[0])
            assert stats[0]["tune"] == (i < 40)
            xs.append(point["x"])
        assert np.std(np.array(xs)) > 0.3
        traces = list(pm.iter_sample(10, pm.NUTS(), tune=5, random_seed=3))
        assert len(traces) == 10 and len(traces[-1]) == 10 and len(traces[3]) == 4


def test_hamiltonian_mc_runs_with_reference_stats():
    with pm.StdNormal(5, sigma=np.arange(1.0, 6.0)):
        trace = pm.sample(400, tune=400, chains=4, step=pm.HamiltonianMC(), random_seed=9,
                          compute_convergence_checks=False)
    assert trace.stat_names == {"step_size", "n_steps", "tune", "step_size_bar", "accept", "diverging",
                                "energy_error", "energy", "path_length", "accepted", "model_logp"}
    assert np.allclose(trace["x"].std(axis=0), np.arange(1.0, 6.0), rtol=0.25)
    assert abs(trace.get_sampler_stats("accept").mean() - 0.65) < 0.15


def test_scaling_argument_and_guess_scaling():
    """test_step.py:505-527 builds NUTS(scaling=model.test_point): diag Hessian -> QuadPotentialDiag."""
    with pm.StdNormal(3, sigma=[1.0, 2.0, 0.5]) as model:
        step = pm.NUTS(scaling=model.test_point)
        assert isinstance(step.potential, pm.QuadPotentialDiag)
        assert np.allclose(step.potential.v, [1.0, 4.0, 0.25], rtol=1e-3)
        trace = pm.sample(300, tune=300, chains=2, step=step, random_seed=1, compute_convergence_checks=False)
    assert np.allclose(trace["x"].std(axis=0), [1.0, 2.0, 0.5], rtol=0.25)


def test_value_grad_function_interface():
    """test_model.py:244-318 (interface part) + developer_guide dict<->array round trip."""
    model = pm.EightSchoolsNCP()
    f = model.logp_dlogp_function()
    assert f.size == 10 and f.dtype == np.float64
    with pytest.raises(TypeError):
        model.logp_dlogp_function(dtype="int32")
    arr = np.arange(10, dtype="f8") / 10
    point = f.array_to_dict(arr)
    assert np.allclose(f.dict_to_array(point), arr)
    logp, grad = f(arr)
    out = np.empty(10)
    logp2 = f(arr, grad_out=out)
    assert logp == logp2 and np.array_equal(grad, out)
    with pytest.raises(ValueError):
        f(np.zeros(3))


def test_multi_device_chain_sharding_is_device_count_invariant():
    import torch
    n = torch.cuda.device_count()
    devices = list(range(min(n, 2))) * (2 if n < 2 else 1)         # 1 GPU: two engines on the same device
    with pm.StdNormal(3):
        a = pm.sample(30, tune=30, chains=6, random_seed=5, compute_convergence_checks=False)
        b = pm.sample(30, tune=30, chains=6, random_seed=5, devices=devices, compute_convergence_checks=False)
    assert np.array_equal(a["x"], b["x"])


def test_save_load_trace_roundtrip(tmp_path, schools_trace):
    """test_ndarray_backend.py:203-280."""
    model, trace = schools_trace
    d = pm.save_trace(trace[:50], str(tmp_path / "t"), overwrite=True)
    back = pm.load_trace(d, model=model)
    assert back.nchains == 8 and len(back) == 50
    assert np.array_equal(back["mu"], trace[:50]["mu"])
    assert np.array_equal(back.get_sampler_stats("depth"), trace[:50].get_sampler_stats("depth"))


def test_posterior_agrees_with_cpu_nuts_by_mcse_z_test():
    """BASELINE.json north star: posterior means agree with the reference's CPU NUTS within an
    MCSE-based |z| < 4 (CPU arm = the oracle port of the reference's step methods, 4 chains)."""
    from oracle import densities as od
    from oracle.hmc_cpu import CpuNUTS, run_chain
    from oracle.potentials import DiagAdaptPotential
    from oracle.rng import PhiloxRNG
    from tests import models_util
    X, y = models_util.glm_data(400, 4, seed=31)
    oracle = od.LogisticGLM(X, y)
    D = oracle.ndim
    cpu = []
    for c in range(4):
        s = CpuNUTS(oracle, D, DiagAdaptPotential(D, np.zeros(D), np.ones(D), 10), PhiloxRNG(900 + c))
        qs, _ = run_chain(s, np.zeros(D), 900, 400)
        cpu.append(qs[400:])
    cpu = np.stack(cpu)                                            # [4, 500, D]
    with pm.LogisticGLM(X, y):
        trace = pm.sample(500, tune=400, chains=64, random_seed=4, step=pm.NUTS(), compute_convergence_checks=False)
    names = ["Intercept"] + ["x%d" % i for i in range(4)]
    for j, name in enumerate(names):
        gpu = np.stack(trace.get_values(name, combine=False))      # [64, 500]
        se = np.hypot(float(pm.stats.mcse_mean(gpu)), float(pm.stats.mcse_mean(cpu[:, :, j])))
        z = (gpu.mean() - cpu[:, :, j].mean()) / se
        assert abs(z) < 4, (name, z)
        assert abs(gpu.std() / cpu[:, :, j].std() - 1) < 0.12, name


def test_resume_from_existing_trace():
    """sampling.py:893-894: passing trace= continues every chain from its last point and appends."""
    with pm.StdNormal(3):
        first = pm.sample(40, tune=40, chains=3, random_seed=1, compute_convergence_checks=False)
        more = pm.sample(25, tune=0, chains=3, random_seed=2, trace=first, step=pm.NUTS(adapt_step_size=False),
                         compute_convergence_checks=False)
    assert len(more) == 65 and more.nchains == 3
    assert np.array_equal(more.get_values("x", chains=1)[:40], first.get_values("x", chains=1))
    assert more.get_sampler_stats("depth", chains=0).shape == (65,)


def test_device_diagnostics_match_host_diagnostics_on_a_real_trace():
    """bulk ESS / R-hat computed on the device trace (no host copy) == the NumPy estimators"""
    import torch
    from pymc3_b200 import _capi, stats_device
    model = pm.EightSchoolsNCP()
    C, D = 32, 10
    eng = model.engine(C, dtype="float32")
    eng.set_state(np.random.default_rng(3).uniform(-1, 1, size=(C, D)), np.arange(C) + 7, 0.25 / D ** 0.25,
                  np.zeros(D), np.ones(D), 10.0)
    opts = dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75, t0=10.0,
                adapt_step_size=1, adapt_mass=1, path_length=2.0, max_steps=1024, hmc_jitter=0, exec_mode=0, glm_path=0)
    out = eng.run(_capi.B2_NUTS, 500, 250, opts)
    q = out["q"][250:]                                                # [draws, chains, D] on the GPU
    assert q.is_cuda
    x = stats_device.from_trace(q)
    ess_d, rhat_d = stats_device.ess_bulk(x), stats_device.rhat(x)
    assert ess_d.is_cuda and rhat_d.is_cuda
    host = x.cpu().numpy()
    np.testing.assert_allclose(ess_d.cpu().numpy(), pm.stats.ess(host), rtol=1e-7)
    np.testing.assert_allclose(rhat_d.cpu().numpy(), pm.stats.rhat(host), rtol=1e-9)
    assert float(rhat_d.max()) < 1.05 and float(ess_d.min()) > 1000
    eng.close()


def test_user_potential_is_called():
    """pymc3/tests/test_quadpotential.py:138-155: a user subclass of QuadPotentialDiag passed as potential= must have
    its methods called while sampling (host-driven transitions, logp / dlogp on the device)."""
    from pymc3_b200.step_methods.hmc import quadpotential
    called = []

    class Potential(quadpotential.QuadPotentialDiag):
        def energy(self, x, velocity=None):
            called.append(1)
            return super().energy(x, velocity)

    with pm.StdNormal(1):
        step = pm.NUTS(potential=Potential(np.ones(1)), dtype="float64")
        trace = pm.sample(10, init=None, step=step, chains=1, tune=10, compute_convergence_checks=False)
    assert called and len(trace) == 10
    assert set(trace.stat_names) >= {"depth", "tree_size", "mean_tree_accept", "energy", "diverging", "step_size"}


def test_dense_mass_matrix_on_the_device():
    """test_step.py:534-565 in spirit: NUTS with a dense `scaling=C` (QuadPotentialFull / FullInv, quadpotential.py:
    400-479) runs all chains on the device in z = L^-1 q (b2_set_dense_mass) and reproduces the moments of the target;
    the adaptive dense potential (FullAdapt) takes the host-driven path with device gradients."""
    sig = np.array([1.0, 2.0, 0.5])
    C = np.diag(sig ** 2) + 0.05
    for kw in (dict(scaling=C, is_cov=True), dict(scaling=np.linalg.inv(C), is_cov=False)):
        with pm.StdNormal(3, sigma=sig):
            step = pm.NUTS(dtype="float64", **kw)
            assert type(step.potential).__name__ in ("QuadPotentialFull", "QuadPotentialFullInv") and step._batched
            trace = pm.sample(500, tune=300, chains=64, step=step, random_seed=11, compute_convergence_checks=False)
        x = trace["x"]
        assert x.shape == (64 * 500, 3)
        assert np.abs(x.mean(axis=0) / sig).max() < 0.05 and np.allclose(x.std(axis=0), sig, rtol=0.03)
        assert 0.7 < trace.get_sampler_stats("mean_tree_accept").mean() < 0.95
    with pm.StdNormal(3, sigma=sig):                      # HamiltonianMC through the same reparameterised seam
        step = pm.HamiltonianMC(scaling=C, is_cov=True, dtype="float64")
        assert step._batched and step._dense
        trace = pm.sample(500, tune=300, chains=64, step=step, random_seed=12, compute_convergence_checks=False)
    x = trace["x"]
    assert np.abs(x.mean(axis=0) / sig).max() < 0.06 and np.allclose(x.std(axis=0), sig, rtol=0.04)
    q1, stats = step.step({"x": np.zeros(3)})                # one chain, one draw: positions in and out are q, not z
    assert np.isfinite(q1["x"]).all() and np.abs(q1["x"]).max() < 20 and "accept" in stats[0]
    from pymc3_b200.step_methods.hmc.quadpotential import QuadPotentialFullAdapt
    with pm.StdNormal(3, sigma=sig):
        with pytest.warns(UserWarning):
            pot = QuadPotentialFullAdapt(3, np.zeros(3), np.eye(3), 1)
        step = pm.NUTS(potential=pot, dtype="float64")
        assert not step._batched
        trace = pm.sample(300, tune=300, chains=2, step=step, random_seed=11, compute_convergence_checks=False)
    x = trace["x"]
    assert np.abs(x.mean(axis=0)).max() < 0.4 and np.allclose(x.std(axis=0), sig, rtol=0.2)
    assert np.allclose(np.sqrt(np.diag(pot._cov)), sig, rtol=0.3)            # the potential learnt the scales


def test_dense_metric_with_a_diagonal_factor_equals_the_diagonal_potential():
    """The reparameterised run must be the same Markov chain as the metric it stands for: with L = diag(s) the dense
    path (z = q / s, unit mass) and the diagonal potential var = s^2 see the same momenta (p_z = n, p_q = n / s), so
    the fp64 traces agree to rounding until chaos takes over, and all tree statistics agree on the first draws."""
    from pymc3_b200 import _capi
    rng = np.random.default_rng(3)
    X = rng.normal(size=(400, 4)).astype("f4")
    y = (rng.uniform(size=400) < 0.5).astype("f4")
    model = pm.LogisticGLM(X, y)
    D, Cn = model.ndim, 8
    s = np.array([0.3, 0.5, 0.2, 0.4, 0.25])
    opts = dict(max_treedepth=6, early_max_treedepth=6, Emax=1000.0, target_accept=0.8, gamma=0.05, k=0.75, t0=10.0,
                adapt_step_size=1, adapt_mass=0, path_length=2.0, max_steps=64, hmc_jitter=0,
                exec_mode=_capi.B2_EXEC_LOCKSTEP, glm_path=_capi.B2_GLM_SIMT)
    q0 = rng.normal(size=(Cn, D)) * 0.1
    seeds = np.arange(Cn, dtype=np.uint64) + 77
    out = []
    for dense in (False, True):
        eng = model.engine(Cn, dtype="float64")
        if dense:
            eng.set_dense_mass(np.diag(s))
            eng.set_state(q0, seeds, 0.2, np.zeros(D), np.ones(D), 0.0)
        else:
            eng.set_state(q0, seeds, 0.2, np.zeros(D), s ** 2, 0.0)
        tr = eng.run(_capi.B2_NUTS, 12, 12, opts)
        out.append({k: v.cpu().numpy() for k, v in tr.items()})
        pos = eng.position()
        assert np.allclose(pos, out[-1]["q"][-1], rtol=1e-12)              # position() speaks q as well
        eng.close()
    a, b = out
    assert (a["tree_size"][:6] == b["tree_size"][:6]).all() and (a["depth"][:6] == b["depth"][:6]).all()
    assert np.allclose(a["q"][:6], b["q"][:6], rtol=1e-7, atol=1e-9)
    assert np.allclose(a["energy"][:6], b["energy"][:6], rtol=1e-9) and np.allclose(a["step_size"][:6], b["step_size"][:6], rtol=1e-9)
    # a dense metric is static: asking the engine to adapt the diagonal next to it is refused
    eng = model.engine(Cn, dtype="float64")
    eng.set_dense_mass(np.diag(s))
    eng.set_state(q0, seeds, 0.2, np.zeros(D), np.ones(D), 0.0)
    with pytest.raises(_capi.B2Error, match="static"):
        eng.run(_capi.B2_NUTS, 2, 2, dict(opts, adapt_mass=1))
    eng.close()


def test_dense_metric_on_a_correlated_posterior_through_the_tensor_core_likelihood():
    """A GLM with correlated regressors in fp32 (tcgen05 likelihood under the reparameterisation): the dense metric built
    from a pilot run's covariance gives the same posterior means as the adaptive diagonal run (|z| < 4 with the
    chains' spread as scale) and needs shorter trees."""
    rng = np.random.default_rng(8)
    base = rng.normal(size=(20000, 1))
    X = (0.9 * base + 0.45 * rng.normal(size=(20000, 6))).astype("f4")          # pairwise correlation ~0.8
    beta = np.array([0.5, -0.3, 0.2, 0.1, -0.4, 0.3])
    y = (rng.uniform(size=20000) < 1 / (1 + np.exp(-(0.2 + X @ beta)))).astype("f4")
    with pm.LogisticGLM(X, y) as model:
        pilot = pm.sample(300, tune=300, chains=128, step=pm.NUTS(), random_seed=5, compute_convergence_checks=False)
    names = model.free_RVs
    flat = np.concatenate([pilot[n].reshape(len(pilot[n]), -1) for n in names], axis=1)
    cov = np.cov(flat, rowvar=0)
    with model:
        step = pm.NUTS(scaling=cov, is_cov=True)
        assert step._batched and step._dense
        dense = pm.sample(300, tune=300, chains=128, step=step, random_seed=6, compute_convergence_checks=False)
    flat_d = np.concatenate([dense[n].reshape(len(dense[n]), -1) for n in names], axis=1)
    sd = flat.std(axis=0)
    z = (flat_d.mean(axis=0) - flat.mean(axis=0)) / (sd / np.sqrt(128 * 300 / 20.0))      # ESS >= draws / 20, conservatively
    assert np.abs(z).max() < 4.0, z
    assert np.allclose(flat_d.std(axis=0), sd, rtol=0.1)
    t_diag = pilot.get_sampler_stats("tree_size")[-128 * 100:].mean()
    t_dense = dense.get_sampler_stats("tree_size")[-128 * 100:].mean()
    assert t_dense < 0.7 * t_diag, (t_dense, t_diag)


def test_sample_callback_and_cancel():
    """pymc3/tests/test_sampling.py:194-219: the callback sees every draw; KeyboardInterrupt from it returns the
    draws made so far."""
    seen = []
    with pm.StdNormal(2):
        pm.sample(10, tune=0, chains=2, step=pm.NUTS(), random_seed=3, compute_convergence_checks=False,
                  callback=lambda trace, draw: seen.append((draw.chain, draw.draw_idx, len(trace))))
    assert len(seen) == 20 and all(n == i + 1 for _, i, n in seen)

    def cancel(trace, draw):
        if len(trace) >= 5:
            raise KeyboardInterrupt()

    with pm.StdNormal(2):
        trace = pm.sample(10, tune=0, chains=1, step=pm.NUTS(), random_seed=3, callback=cancel,
                          compute_convergence_checks=False)
    assert len(trace) == 5


def test_init_nuts_pilot_run_gives_a_dense_metric():
    """sampling.py:2002-2007: init='nuts' runs a pilot NUTS job, takes `trace_cov` of it as a static dense metric
    (QuadPotentialFull, on the device) and pilot draws as start points."""
    rng = np.random.default_rng(4)
    base = rng.normal(size=(4000, 1))
    X = (0.9 * base + 0.45 * rng.normal(size=(4000, 3))).astype("f4")
    y = (rng.uniform(size=4000) < 1 / (1 + np.exp(-(X @ np.array([0.4, -0.2, 0.3]))))).astype("f4")
    with pm.LogisticGLM(X, y) as model:
        start, step = pm.init_nuts(init="nuts", chains=16, n_init=150, random_seed=3, progressbar=False)
        assert type(step.potential).__name__ == "QuadPotentialFull" and step._batched and len(start) == 16
        cov = step.potential._cov
        assert cov.shape == (model.ndim, model.ndim) and np.all(np.linalg.eigvalsh(cov) > 0)
        off = cov / np.sqrt(np.outer(np.diag(cov), np.diag(cov)))
        assert np.abs(off - np.eye(model.ndim)).max() > 0.3            # the regressors are correlated: so is the posterior
        trace = pm.sample(150, tune=100, chains=16, init="nuts", n_init=150, random_seed=5, progressbar=False,
                          compute_convergence_checks=False)
    assert np.allclose(pm.trace_cov(trace, model=model), cov, rtol=0.6, atol=0.3 * np.diag(cov).max())
    assert 0.6 < trace.get_sampler_stats("mean_tree_accept").mean() < 0.97

