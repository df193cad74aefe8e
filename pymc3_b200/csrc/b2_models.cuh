// Fused logp + dlogp of the model families, evaluated by ONE thread group per chain.
//
// These replace the Theano-compiled `ValueGradFunction.__call__` (pymc3/model.py:645-666)
// for the families named in BASELINE.json.  Densities follow the reference's expressions:
//   Normal        pymc3/distributions/continuous.py:518-537     (tau form)
//   HalfCauchy    :2432-2447  + log transform pymc3/distributions/transforms.py:164-181,203-216
//   Exponential   :1549-1563      StudentT :2021-2040       Flat :300-314
//   Bernoulli/Binomial(n=1) logit form  pymc3/distributions/discrete.py:104,350; glm/families.py:115-119
//   GaussianRandomWalk  pymc3/distributions/timeseries.py:237-256
// and the models are SURVEY appendix C (gelman_schools.py:26-40, glm/linear.py:49-101,
// benchmarks/benchmarks/benchmarks.py:25-45, stochastic_volatility.ipynb cell 10).
//
// The group evaluators here serve (a) the persistent NUTS kernel, where a chain's whole
// likelihood fits one warp/block, (b) the parity hook b2_logp_dlogp, and (c) as the slow but
// simple cross-check of the chain-batched kernels for the large-N families (b2_glm_*.cu,
// b2_hier.cu).  Vectors are of type T (float production / double check build); every sum is
// accumulated in double.
#pragma once
#include "b2_core.cuh"

enum {
    B2_FAMILY_STD_NORMAL = 0,        // x_i ~ N(mu_i, sigma_i)             aux0=mu[D] aux1=sigma[D]
    B2_FAMILY_EIGHT_SCHOOLS_NCP = 2, // eta[J], mu, tau_log__               aux0=y[J]  aux1=sigma[J]  hp={mu_sd, tau_beta}
    B2_FAMILY_GLM_LOGIT = 3,         // Intercept, x0..x{K-1}               X[N,K] y[N]               hp={prior_tau}
    B2_FAMILY_HIER_LINEAR_NCP = 4,   // mu_a,sa_log,mu_b,sb_log,a[G],b[G],eps_log   (sorted by group) hp={mu_sd, hc_beta}
    B2_FAMILY_STOCH_VOL = 5,         // step_log, vol[T], nu_log            aux0=returns[T]           hp={step_lam, nu_lam}
};

struct B2ModelData {
    int family;
    int D;                 // number of free (unconstrained) parameters
    int N;                 // observations (J / N / T)
    int G;                 // groups (hier) or regressors K (glm)
    const double* aux0;    // fp64 data vectors (small models)
    const double* aux1;
    const float* X;        // [N, K] row-major fp32 (glm)
    const float* yf;       // [N] fp32 responses (glm: 0/1; hier: y)
    const unsigned char* floor_u8;   // [N] hier
    const int* grp_off;    // [G+1] hier: observations are sorted by group
    void* scratch;         // [C, N] T, glm group evaluator only
    const void* aux0_chip; // optional on-chip copy of aux0 in the vector dtype (persistent block kernel, stoch. vol.)
    double hp[4];
};

#define B2_LOG_2PI 1.8378770664093454835606594728112
#define B2_LOG_PI 1.1447298858494001741434273513531
#define B2_LOG_2 0.693147180559945309417232121458

// element-wise transcendentals in the vector dtype (fp32 production build: float versions)
B2_HD float b2_exp_t(float x) { return expf(x); }
B2_HD double b2_exp_t(double x) { return exp(x); }
B2_HD float b2_log1p_t(float x) { return log1pf(x); }
B2_HD double b2_log1p_t(double x) { return log1p(x); }

B2_HD double b2_digamma(double x) {
    double r = 0.0;
    while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
    const double f = 1.0 / (x * x);
    const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 +
                     f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
    return r + log(x) - 0.5 / x + t;
}

// Normal.logp with sigma given:  (-tau (x-mu)^2 + log(tau/(2 pi)))/2
B2_HD double b2_normal_logp(double x, double mu, double sigma) {
    const double tau = 1.0 / (sigma * sigma);
    const double d = x - mu;
    return 0.5 * (-tau * d * d + log(tau) - B2_LOG_2PI);
}

// d/du [ HalfCauchy(e^u | beta) + u ] and its value
B2_HD double b2_halfcauchy_log_logp(double u, double beta, double* du) {
    const double x = exp(u);
    const double w = (x / beta) * (x / beta);
    *du = 1.0 - 2.0 * w / (1.0 + w);
    return B2_LOG_2 - B2_LOG_PI - log(beta) - log1p(w) + u;
}

// ------------------------------------------------------------------------------ std normal
template <typename T, typename G>
B2_HD double b2_eval_std_normal(const G& g, const B2ModelData& m, const T* q, T* grad) {
    double acc[1] = {0.0};
    for (int i = g.lane(); i < m.D; i += G::NT) {
        const double mu = m.aux0 ? m.aux0[i] : 0.0, sg = m.aux1 ? m.aux1[i] : 1.0;
        const double x = (double)q[i];
        acc[0] += b2_normal_logp(x, mu, sg);
        grad[i] = (T)(-(x - mu) / (sg * sg));
    }
    g.allsum(acc);
    return acc[0];
}

// --------------------------------------------------------------------------- eight schools
template <typename T, typename G>
B2_HD double b2_eval_eight_schools(const G& g, const B2ModelData& m, const T* q, T* grad) {
    const int J = m.N;
    const double mu = (double)q[J], u = (double)q[J + 1];
    const double tau = exp(u);
    double acc[3] = {0.0, 0.0, 0.0};                 // logp, sum r, sum r*eta
    for (int j = g.lane(); j < J; j += G::NT) {
        const double eta = (double)q[j], sg = m.aux1[j];
        const double r = (m.aux0[j] - mu - tau * eta) / (sg * sg);
        acc[0] += b2_normal_logp(eta, 0.0, 1.0) + b2_normal_logp(m.aux0[j], mu + tau * eta, sg);
        acc[1] += r;
        acc[2] += r * eta;
        grad[j] = (T)(-eta + tau * r);
    }
    g.allsum(acc);
    double du;
    const double lp = acc[0] + b2_normal_logp(mu, 0.0, m.hp[0]) + b2_halfcauchy_log_logp(u, m.hp[1], &du);
    if (g.lane() == 0) {
        grad[J] = (T)(-mu / (m.hp[0] * m.hp[0]) + acc[1]);
        grad[J + 1] = (T)(du + tau * acc[2]);
    }
    return lp;
}

// ------------------------------------------------------------------- Bernoulli-logit GLM
// group evaluator (small N / cross-check): pass 1 over observations, pass 2 over regressors
template <typename T, typename G>
B2_HD double b2_eval_glm_logit(const G& g, const B2ModelData& m, const T* q, T* grad, int chain) {
    const int K = m.G, N = m.N;
    T* resid = (T*)m.scratch + (size_t)chain * N;
    const T b0 = q[0];
    double acc[2] = {0.0, 0.0};                      // logp, sum resid
    for (int i = g.lane(); i < N; i += G::NT) {
        const float* xr = m.X + (size_t)i * K;
        T eta = b0;
        for (int k = 0; k < K; ++k) eta += (T)xr[k] * q[1 + k];
        const T y = (T)m.yf[i];
        const T e = (T)exp(-fabs((double)eta));       // softplus(eta) = max(eta,0) + log1p(e^-|eta|)
        const T sp = (eta > 0 ? eta : (T)0) + (T)log1p((double)e);
        const T sig = eta >= 0 ? (T)1 / ((T)1 + e) : e / ((T)1 + e);
        const T r = y - sig;
        resid[i] = r;
        acc[0] += (double)(y * eta - sp);
        acc[1] += (double)r;
    }
    g.sync();
    const double ptau = m.hp[0];
    double pr[1] = {0.0};
    for (int k = g.lane(); k < K; k += G::NT) {
        double s = 0.0;
        for (int i = 0; i < N; ++i) s += (double)((T)m.X[(size_t)i * K + k] * resid[i]);
        const double b = (double)q[1 + k];
        grad[1 + k] = (T)(s - ptau * b);
        pr[0] += 0.5 * (-ptau * b * b + log(ptau) - B2_LOG_2PI);
    }
    g.allsum(acc);
    g.allsum(pr);
    if (g.lane() == 0) grad[0] = (T)acc[1];
    return acc[0] + pr[0];
}

// ------------------------------------------------------------ hierarchical linear (radon NCP)
template <typename T, typename G>
B2_HD double b2_eval_hier(const G& g, const B2ModelData& m, const T* q, T* grad) {
    const int NG = m.G;
    const double mu_a = (double)q[0], ua = (double)q[1], mu_b = (double)q[2], ub = (double)q[3];
    const double ue = (double)q[4 + 2 * NG];
    const double sa = exp(ua), sb = exp(ub), eps = exp(ue);
    const double inv_e2 = 1.0 / (eps * eps);
    // logp_lik, ss, sumS0, sumS1, S0.a, S1.b, prior(a,b)
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int gi = g.lane(); gi < NG; gi += G::NT) {
        const double a = (double)q[4 + gi], b = (double)q[4 + NG + gi];
        const T A = (T)(mu_a + sa * a), B = (T)(mu_b + sb * b);
        double s0 = 0.0, s1 = 0.0, ss = 0.0;
        for (int i = m.grp_off[gi]; i < m.grp_off[gi + 1]; ++i) {
            const T fl = (T)m.floor_u8[i];
            const T r = (T)m.yf[i] - (A + B * fl);
            s0 += (double)r;
            s1 += (double)(r * fl);
            ss += (double)(r * r);
        }
        s0 *= inv_e2; s1 *= inv_e2;
        acc[1] += ss; acc[2] += s0; acc[3] += s1; acc[4] += s0 * a; acc[5] += s1 * b;
        acc[6] += -0.5 * (a * a + b * b) - B2_LOG_2PI;
        grad[4 + gi] = (T)(-a + sa * s0);
        grad[4 + NG + gi] = (T)(-b + sb * s1);
    }
    g.allsum(acc);
    double dua, dub, due;
    const double N = (double)m.N;
    double lp = acc[6] + b2_normal_logp(mu_a, 0.0, m.hp[0]) + b2_normal_logp(mu_b, 0.0, m.hp[0]);
    lp += b2_halfcauchy_log_logp(ua, m.hp[1], &dua) + b2_halfcauchy_log_logp(ub, m.hp[1], &dub) +
          b2_halfcauchy_log_logp(ue, m.hp[1], &due);
    lp += -0.5 * inv_e2 * acc[1] + N * (-ue - 0.5 * B2_LOG_2PI);
    if (g.lane() == 0) {
        const double pm = 1.0 / (m.hp[0] * m.hp[0]);
        grad[0] = (T)(-mu_a * pm + acc[2]);
        grad[1] = (T)(dua + sa * acc[4]);
        grad[2] = (T)(-mu_b * pm + acc[3]);
        grad[3] = (T)(dub + sb * acc[5]);
        grad[4 + 2 * NG] = (T)(due - N + acc[1] * inv_e2);
    }
    return lp;
}

// ------------------------------------------------------------------- stochastic volatility
template <typename T, typename G>
B2_HD double b2_eval_stoch_vol(const G& g, const B2ModelData& m, const T* q, T* grad) {
    const int Tn = m.N;
    const double a = (double)q[0], c = (double)q[1 + Tn];
    const double s = exp(a), nu = exp(c);
    const double inv_s2 = 1.0 / (s * s);
    const double hnu1 = 0.5 * (nu + 1.0);
    const T* vol = q + 1;
    // logp terms, sum d^2, sum dnu-part: a lane's few elements are summed in the vector dtype, the
    // cross-lane reduction is always fp64
    T part[3] = {(T)0, (T)0, (T)0};
    const T inv_s2_t = (T)inv_s2, inv_nu = (T)(1.0 / nu), hnu1_t = (T)hnu1, nu1_t = (T)(nu + 1.0);
    const T hnu1_over_nu = (T)(hnu1 / nu);
    for (int i = g.lane(); i < Tn; i += G::NT) {
        const T vi = vol[i];
        T gv = (T)0;
        if (i > 0) {
            const T d = vi - vol[i - 1];
            part[1] += d * d;
            gv -= d * inv_s2_t;
        }
        if (i + 1 < Tn) gv += (vol[i + 1] - vi) * inv_s2_t;
        const T r = m.aux0_chip ? static_cast<const T*>(m.aux0_chip)[i] : (T)m.aux0[i];
        const T z = b2_exp_t((T)-2 * vi) * (r * r * inv_nu);           // lam r^2 / nu
        const T l1 = b2_log1p_t(z);
        const T zr = z / ((T)1 + z);
        part[0] += -vi - hnu1_t * l1;                                  // 0.5 log(lam) = -vol
        part[2] += -(T)0.5 * l1 + hnu1_over_nu * zr;
        grad[1 + i] = gv + (nu1_t * zr - (T)1);
    }
    // The terms that depend on (a, c) only -- two lgamma, two digamma, the logs of the hyper-parameters: ~1000 fp64
    // instructions -- were evaluated by every thread of the group on the critical path of every leapfrog.  Each is now
    // one job done by one lane of a different warp (a one-warp or one-lane group does them all), and the group sum
    // below delivers the totals to everybody.
    double acc[5] = {(double)part[0], (double)part[1], (double)part[2], 0.0, 0.0};
    {
        constexpr int NW = (G::NT + 31) / 32;
        const int wid = g.lane() >> 5;
        if ((g.lane() & 31) == 0) {
            if (0 % NW == wid) acc[3] += Tn * lgamma(hnu1);
            if (1 % NW == wid) acc[3] -= Tn * lgamma(0.5 * nu);
            if (2 % NW == wid) acc[4] += Tn * 0.5 * b2_digamma(hnu1);
            if (3 % NW == wid) acc[4] -= Tn * 0.5 * b2_digamma(0.5 * nu);
            if (4 % NW == wid) acc[3] += log(m.hp[0]) + log(m.hp[1]);
        }
    }
    g.allsum(acc);
    const double n1 = (double)(Tn - 1);
    double lp = (-m.hp[0] * s + a)                                                // Exp(s|10) + jacobian   (log lam in acc[3])
              + (-0.5 * inv_s2 * acc[1] + n1 * (-a - 0.5 * B2_LOG_2PI))            // GRW innovations (init Flat)
              + (-m.hp[1] * nu + c)                                                // Exp(nu|0.1) + jacobian
              + acc[0] + acc[3] + Tn * (-0.5 * c - 0.5 * B2_LOG_PI);               // log(nu) = c
    if (g.lane() == 0) {
        grad[0] = (T)(-m.hp[0] * s + 1.0 + acc[1] * inv_s2 - n1);
        const double dnu = acc[4] - Tn * 0.5 / nu + acc[2];
        grad[1 + Tn] = (T)(-m.hp[1] * nu + 1.0 + nu * dnu);
    }
    return lp;
}

template <typename T, typename G>
B2_HD double b2_eval_model(const G& g, const B2ModelData& m, const T* q, T* grad, int chain) {
    switch (m.family) {
    case B2_FAMILY_STD_NORMAL: return b2_eval_std_normal<T, G>(g, m, q, grad);
    case B2_FAMILY_EIGHT_SCHOOLS_NCP: return b2_eval_eight_schools<T, G>(g, m, q, grad);
    case B2_FAMILY_GLM_LOGIT: return b2_eval_glm_logit<T, G>(g, m, q, grad, chain);
    case B2_FAMILY_HIER_LINEAR_NCP: return b2_eval_hier<T, G>(g, m, q, grad);
    case B2_FAMILY_STOCH_VOL: return b2_eval_stoch_vol<T, G>(g, m, q, grad);
    default: return NAN;
    }
}
