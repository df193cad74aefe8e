"""fp64 log-densities and analytic gradients of the model families.  TEST INFRASTRUCTURE.

Elementary densities restate the reference's expressions:
  Normal        pymc3/distributions/continuous.py:518-537 (tau form, get_tau_sigma :105-144)
  HalfNormal    :888-905        HalfCauchy :2432-2447      Exponential :1549-1563
  StudentT      :2021-2040      Flat :300-314
  Bernoulli(logit_p) pymc3/distributions/discrete.py:350   Binomial :104 (+dist_math.py:78-91)
  log transform pymc3/distributions/transforms.py:164-181, 203-216
  GaussianRandomWalk pymc3/distributions/timeseries.py:237-256
and the five benchmark models follow SURVEY appendix C (model sources cited per class).
Gradients are derived by hand (Theano's tt.grad, model.py:625, is unavailable) and are
checked against central finite differences in tests/test_oracle_densities.py.

Flat-vector layout: free (transformed) variables concatenated in *creation order*
(blocking.py:33-59 applied to model.free_RVs), which is what
``model.logp_dlogp_function()`` uses (model.py:885-887).
"""
import numpy as np
from scipy.special import gammaln, digamma, expit

LOG_2PI = np.log(2.0 * np.pi)


# ----------------------------------------------------------------- elementary densities
def normal_logp(x, mu, sigma=None, tau=None):
    if tau is None:
        tau = sigma ** -2.0                      # continuous.py:105-144
    return (-tau * (x - mu) ** 2 + np.log(tau / np.pi / 2.0)) / 2.0     # :535-536


def half_normal_logp(x, sigma):
    tau = sigma ** -2.0
    return np.where(x >= 0, -0.5 * tau * x ** 2 + 0.5 * np.log(tau * 2.0 / np.pi), -np.inf)


def half_cauchy_logp(x, beta):
    return np.where(x >= 0, np.log(2.0) - np.log(np.pi) - np.log(beta) - np.log1p((x / beta) ** 2),
                    -np.inf)                     # :2445-2447


def exponential_logp(x, lam):
    return np.where(x >= 0, np.log(lam) - lam * x, -np.inf)             # :1562-1563


def student_t_logp(x, nu, mu, lam):
    return (gammaln((nu + 1.0) / 2.0) + 0.5 * np.log(lam / (nu * np.pi)) - gammaln(nu / 2.0)
            - (nu + 1.0) / 2.0 * np.log1p(lam * (x - mu) ** 2 / nu))   # :2036-2040


def bernoulli_logit_logp(y, eta):
    # discrete.py:350: switch(value, -log1pexp(-logit_p)?, ...) == y*eta - softplus(eta)
    return y * eta - np.logaddexp(0.0, eta)


def binomial_n1_logp(y, p):
    # glm/families.py:115-119 -> Binomial(n=1, p=sigmoid(eta)); discrete.py:104;
    # binomln(1, y) == 0 for y in {0, 1}; logpow(0, 0) == 0 (dist_math.py:78-91)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.where(y == 0, 0.0, y * np.log(p))
        b = np.where(1 - y == 0, 0.0, (1 - y) * np.log1p(-p))
    return a + b


# ----------------------------------------------------------------------- model scaffolding
class OracleModel:
    """Holds the ordering of free variables (name, slice, shape) and deterministics."""

    free = ()            # list of (name, shape)
    log_transformed = () # names of positive variables sampled on the log scale

    def _finish(self):
        self.slices, off = {}, 0
        for name, shape in self.free:
            n = int(np.prod(shape)) if shape else 1
            self.slices[name] = (slice(off, off + n), shape)
            off += n
        self.ndim = off

    def pack(self, point):
        out = np.empty(self.ndim)
        for name, (slc, shape) in self.slices.items():
            out[slc] = np.ravel(point[name])
        return out

    def unpack(self, q):
        out = {}
        for name, (slc, shape) in self.slices.items():
            out[name] = q[slc].reshape(shape) if shape else q[slc][0]
        return out

    def test_point(self):
        return np.zeros(self.ndim)

    def __call__(self, q):
        return self.logp_dlogp(np.asarray(q, dtype="d"))

    def logp(self, q):
        return self.logp_dlogp(q)[0]


class NormalPair(OracleModel):
    """tests/test_step.py:505-527: x ~ N(0,1); y ~ N(x,1) observed 1."""

    free = (("x", ()),)

    def __init__(self, y=1.0):
        self.y = y
        self._finish()

    def logp_dlogp(self, q):
        x = q[0]
        lp = normal_logp(x, 0.0, 1.0) + normal_logp(self.y, x, 1.0)
        return lp, np.array([-x + (self.y - x)])


class DevGuideModel(OracleModel):
    """docs/source/developer_guide.rst:560-575: z~N(0,10)[10]; x~N(z,1)[10]; y~N(sum x,1) obs 2.5."""

    free = (("z", (10,)), ("x", (10,)))

    def __init__(self):
        self._finish()

    def logp_dlogp(self, q):
        z, x = q[:10], q[10:]
        s = x.sum()
        lp = normal_logp(z, 0.0, 10.0).sum() + normal_logp(x, z, 1.0).sum() + normal_logp(2.5, s, 1.0)
        gz = -z / 100.0 + (x - z)
        gx = -(x - z) + (2.5 - s)
        return lp, np.concatenate([gz, gx])


class StdNormal(OracleModel):
    """x ~ N(0, 1)[n] -- benchmarks.py:75-91 overhead model / sampler_fixtures Normal."""

    def __init__(self, n=1, sigma=1.0):
        self.free = (("x", (n,)),)
        self.sigma = np.broadcast_to(np.asarray(sigma, dtype="d"), (n,)).copy()
        self._finish()

    def logp_dlogp(self, q):
        return normal_logp(q, 0.0, self.sigma).sum(), -q / self.sigma ** 2


class EightSchoolsNCP(OracleModel):
    """pymc3/examples/gelman_schools.py:26-40 (C1).  Free: eta[J], mu, tau_log__."""

    log_transformed = ("tau",)

    def __init__(self, y=None, sigma=None, mu_sd=1e6, tau_beta=25.0):
        self.y = np.array([28, 8, -3, 7, -1, 1, 18, 12], dtype="d") if y is None else np.asarray(y, "d")
        self.sigma = (np.array([15, 10, 16, 11, 9, 11, 10, 18], dtype="d") if sigma is None
                      else np.asarray(sigma, "d"))
        self.J = len(self.y)
        self.mu_sd, self.tau_beta = float(mu_sd), float(tau_beta)
        self.free = (("eta", (self.J,)), ("mu", ()), ("tau_log__", ()))
        self._finish()

    def logp_dlogp(self, q):
        J = self.J
        eta, mu, u = q[:J], q[J], q[J + 1]
        tau = np.exp(u)
        resid = self.y - mu - tau * eta
        r = resid / self.sigma ** 2
        lp = (normal_logp(eta, 0.0, 1.0).sum() + normal_logp(mu, 0.0, self.mu_sd)
              + float(half_cauchy_logp(tau, self.tau_beta)) + u
              + normal_logp(self.y, mu + tau * eta, self.sigma).sum())
        w = (tau / self.tau_beta) ** 2
        g = np.empty(J + 2)
        g[:J] = -eta + tau * r
        g[J] = -mu / self.mu_sd ** 2 + r.sum()
        g[J + 1] = 1.0 - 2.0 * w / (1.0 + w) + tau * np.dot(r, eta)
        return lp, g

    def deterministics(self, qs):
        return {"tau": np.exp(qs[..., self.J + 1])}


class LogisticGLM(OracleModel):
    """glm/linear.py:49-101 + glm/families.py:115-119 (C2, C5).

    Free (creation order): Intercept ~ Flat, x0..x{D-1} ~ Normal(0, tau=1e-6).
    The logit form is evaluated (identical to Binomial(n=1, p=sigmoid) where finite).
    """

    def __init__(self, X, y, prior_tau=1e-6, intercept=True):
        self.X = np.asarray(X, dtype="d")
        self.y = np.asarray(y, dtype="d")
        self.N, self.D = self.X.shape
        self.prior_tau = prior_tau
        self.intercept = intercept
        names = (["Intercept"] if intercept else []) + ["x%d" % i for i in range(self.D)]
        self.free = tuple((n, ()) for n in names)
        self._finish()

    def logp_dlogp(self, q):
        if self.intercept:
            b0, beta = q[0], q[1:]
        else:
            b0, beta = 0.0, q
        eta = b0 + self.X @ beta
        lp = normal_logp(beta, 0.0, tau=self.prior_tau).sum() + bernoulli_logit_logp(self.y, eta).sum()
        resid = self.y - expit(eta)
        gb = self.X.T @ resid - self.prior_tau * beta
        g = np.concatenate([[resid.sum()], gb]) if self.intercept else gb
        return lp, g


class HierLinearNCP(OracleModel):
    """benchmarks/benchmarks/benchmarks.py:25-45 (C3, radon NCP).

    Free: mu_a, sigma_a_log__, mu_b, sigma_b_log__, a[G], b[G], eps_log__   (D = 2G+5).
    NB the benchmark literally writes sd=100**2 for mu_a/mu_b.
    """

    log_transformed = ("sigma_a", "sigma_b", "eps")

    def __init__(self, group_idx, floor, y, n_groups, mu_sd=100.0 ** 2, hc_beta=5.0):
        self.idx = np.asarray(group_idx, dtype=np.int64)
        self.floor = np.asarray(floor, dtype="d")
        self.y = np.asarray(y, dtype="d")
        self.G = int(n_groups)
        self.mu_sd, self.hc_beta = float(mu_sd), float(hc_beta)
        G = self.G
        self.free = (("mu_a", ()), ("sigma_a_log__", ()), ("mu_b", ()), ("sigma_b_log__", ()),
                     ("a", (G,)), ("b", (G,)), ("eps_log__", ()))
        self._finish()

    def logp_dlogp(self, q):
        G = self.G
        mu_a, ua, mu_b, ub = q[0], q[1], q[2], q[3]
        a, b, ue = q[4:4 + G], q[4 + G:4 + 2 * G], q[4 + 2 * G]
        sa, sb, eps = np.exp(ua), np.exp(ub), np.exp(ue)
        A = mu_a + sa * a
        B = mu_b + sb * b
        yhat = A[self.idx] + B[self.idx] * self.floor
        resid = self.y - yhat
        lp = (normal_logp(mu_a, 0.0, self.mu_sd) + normal_logp(mu_b, 0.0, self.mu_sd)
              + float(half_cauchy_logp(sa, self.hc_beta)) + ua
              + float(half_cauchy_logp(sb, self.hc_beta)) + ub
              + float(half_cauchy_logp(eps, self.hc_beta)) + ue
              + normal_logp(a, 0.0, 1.0).sum() + normal_logp(b, 0.0, 1.0).sum()
              + normal_logp(self.y, yhat, eps).sum())
        r = resid / eps ** 2
        S0 = np.bincount(self.idx, weights=r, minlength=G)
        S1 = np.bincount(self.idx, weights=r * self.floor, minlength=G)
        ss = np.dot(resid, resid)

        def hc_du(s):
            w = (s / self.hc_beta) ** 2
            return 1.0 - 2.0 * w / (1.0 + w)

        g = np.empty(self.ndim)
        g[0] = -mu_a / self.mu_sd ** 2 + S0.sum()
        g[1] = hc_du(sa) + sa * np.dot(S0, a)
        g[2] = -mu_b / self.mu_sd ** 2 + S1.sum()
        g[3] = hc_du(sb) + sb * np.dot(S1, b)
        g[4:4 + G] = -a + sa * S0
        g[4 + G:4 + 2 * G] = -b + sb * S1
        g[4 + 2 * G] = hc_du(eps) - len(self.y) + ss / eps ** 2
        return lp, g

    def deterministics(self, qs):
        G = self.G
        return {"sigma_a": np.exp(qs[..., 1]), "sigma_b": np.exp(qs[..., 3]),
                "eps": np.exp(qs[..., 4 + 2 * G])}


class StochVol(OracleModel):
    """docs/source/notebooks/stochastic_volatility.ipynb cell 10 (C4).

    step_size ~ Exponential(10); volatility ~ GaussianRandomWalk(sigma=step_size, shape=T)
    (init = Flat, timeseries.py:188-256); nu ~ Exponential(0.1);
    returns ~ StudentT(nu, lam=exp(-2 volatility)).
    Free: step_size_log__, volatility[T], nu_log__.
    """

    log_transformed = ("step_size", "nu")

    def __init__(self, returns, step_lam=10.0, nu_lam=0.1):
        self.r = np.asarray(returns, dtype="d")
        self.T = len(self.r)
        self.step_lam, self.nu_lam = step_lam, nu_lam
        self.free = (("step_size_log__", ()), ("volatility", (self.T,)), ("nu_log__", ()))
        self._finish()

    def logp_dlogp(self, q):
        T = self.T
        a, vol, c = q[0], q[1:1 + T], q[1 + T]
        s, nu = np.exp(a), np.exp(c)
        d = vol[1:] - vol[:-1]
        lam = np.exp(-2.0 * vol)
        z = lam * self.r ** 2 / nu
        lp = (float(exponential_logp(s, self.step_lam)) + a
              + normal_logp(vol[1:], vol[:-1], s).sum()
              + float(exponential_logp(nu, self.nu_lam)) + c
              + student_t_logp(self.r, nu, 0.0, lam).sum())
        g = np.empty(self.ndim)
        # d/da: Exp prior (-lam*s) + jacobian 1 + GRW: sum(d^2/s^2 - 1)
        g[0] = -self.step_lam * s + 1.0 + (d ** 2).sum() / s ** 2 - (T - 1)
        gv = np.zeros(T)
        gv[1:] += -d / s ** 2
        gv[:-1] += d / s ** 2
        # StudentT wrt vol_i: 0.5*(-2) + (nu+1)/2 * 2 z/(1+z)
        gv += -1.0 + (nu + 1.0) * z / (1.0 + z)
        g[1:1 + T] = gv
        # d/dnu of sum_i [...]
        dnu = (0.5 * digamma((nu + 1.0) / 2.0) - 0.5 * digamma(nu / 2.0) - 0.5 / nu
               - 0.5 * np.log1p(z) + (nu + 1.0) / 2.0 * z / (nu * (1.0 + z)))
        g[1 + T] = -self.nu_lam * nu + 1.0 + nu * dnu.sum()
        return lp, g

    def deterministics(self, qs):
        return {"step_size": np.exp(qs[..., 0]), "nu": np.exp(qs[..., 1 + self.T])}


def finite_difference_grad(model, q, h=1e-6):
    q = np.asarray(q, dtype="d")
    g = np.empty_like(q)
    for i in range(len(q)):
        e = np.zeros_like(q)
        e[i] = h * max(1.0, abs(q[i]))
        g[i] = (model.logp(q + e) - model.logp(q - e)) / (2 * e[i])
    return g
