"""Density parity exactly as the reference checks it: against scipy.stats on its own grids
(pymc3/tests/test_distributions.py:139-158 domains, :456-465 check_logp, 6 decimals)."""
import numpy as np
import pytest
from scipy import stats as sp

from oracle import densities as od
from tests import models_util

R = [-np.inf, -2.1, -1, -0.01, 0.0, 0.01, 1, 2.1, np.inf][1:-1]
Rplus = [0, 0.01, 0.1, 0.9, 0.99, 1, 1.5, 2, 100][1:-1]
Rplusbig = [0, 0.5, 0.9, 0.99, 1, 1.5, 2, 20][1:-1]
Unit = [0, 0.001, 0.1, 0.5, 0.75, 0.99, 1][1:-1]
Bool = [0, 0, 1, 1]


def test_normal():          # test_distributions.py:546
    for x in R:
        for mu in R:
            for s in Rplus:
                assert abs(od.normal_logp(x, mu, s) - sp.norm.logpdf(x, mu, s)) < 1e-6 * max(1, abs(sp.norm.logpdf(x, mu, s)))


def test_half_normal():     # :566
    for x in Rplus:
        for s in Rplus:
            assert abs(od.half_normal_logp(x, s) - sp.halfnorm.logpdf(x, scale=s)) < 1e-6


def test_exponential():     # :626
    for x in Rplus:
        for lam in Rplus:
            assert abs(od.exponential_logp(x, lam) - sp.expon.logpdf(x, 0, 1 / lam)) < 1e-6


def test_half_cauchy():     # :667
    for x in Rplus:
        for b in Rplusbig:
            assert abs(od.half_cauchy_logp(x, b) - sp.halfcauchy.logpdf(x, scale=b)) < 1e-6


def test_student_t():       # :651
    for x in R:
        for nu in Rplus:
            for mu in R:
                for lam in Rplus:
                    ref = sp.t.logpdf(x, nu, mu, lam ** -0.5)
                    assert abs(od.student_t_logp(x, nu, mu, lam) - ref) < 1e-6 * max(1, abs(ref))


def test_bernoulli_and_binomial():   # :725, :734
    for y in Bool:
        for p in Unit:
            ref = sp.bernoulli.logpmf(y, p)
            assert abs(od.binomial_n1_logp(np.float64(y), p) - ref) < 1e-6
            assert abs(od.bernoulli_logit_logp(y, np.log(p / (1 - p))) - ref) < 1e-6


def test_gaussian_random_walk_is_sum_of_normals():
    """timeseries.py:237-256 (parity otherwise unpinned in the reference)."""
    rng = np.random.default_rng(0)
    x = rng.normal(size=20).cumsum()
    s = 0.3
    ref = sp.norm.logpdf(x[1:], x[:-1], s).sum()
    m = od.StochVol(rng.normal(size=20) * 0.01)
    a, c = np.log(s), np.log(5.0)
    q = np.concatenate([[a], x, [c]])
    lam = np.exp(-2 * x)
    expect = (sp.expon.logpdf(s, 0, 1 / 10.0) + a + ref + sp.expon.logpdf(5.0, 0, 10.0) + c
              + sp.t.logpdf(m.r, 5.0, 0, lam ** -0.5).sum())
    assert abs(m.logp(q) - expect) < 1e-8 * abs(expect)


@pytest.mark.parametrize("name", ["std_normal", "eight_schools", "glm", "hier", "stoch_vol"])
def test_analytic_gradients_match_finite_differences(name):
    _, oracle = models_util.pairs()[name]
    rng = np.random.default_rng(1)
    for _ in range(3):
        q = rng.normal(size=oracle.ndim) * 0.4
        _, g = oracle(q)
        fd = od.finite_difference_grad(oracle, q)
        assert np.abs(g - fd).max() <= 2e-6 * max(1.0, np.abs(g).max())


def test_models_compose_elementary_densities():
    m = od.EightSchoolsNCP()
    q = np.random.default_rng(2).normal(size=10) * 0.5
    eta, mu, u = q[:8], q[8], q[9]
    tau = np.exp(u)
    expect = (sp.norm.logpdf(eta).sum() + sp.norm.logpdf(mu, 0, 1e6) + sp.halfcauchy.logpdf(tau, scale=25) + u
              + sp.norm.logpdf(m.y, mu + tau * eta, m.sigma).sum())
    assert abs(m.logp(q) - expect) < 1e-9 * abs(expect)
