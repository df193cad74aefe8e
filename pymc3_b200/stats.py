"""Convergence diagnostics: rank-normalised split R-hat and bulk / tail ESS, MCSE.

The reference re-exports these from arviz (pymc3/stats/__init__.py:42-55; used by
backends/report.py:110-123 and benchmarks/benchmarks/benchmarks.py:168), which is not available
here, so the published algorithms are implemented directly:
Vehtari, Gelman, Simpson, Carpenter, Buerkner (2021) "Rank-normalization, folding, and
localization: an improved R-hat", and Geyer's initial monotone sequence estimator as in Stan.
Input layout everywhere: [chains, draws, *shape]; the result has shape `shape`.
"""
import numpy as np
from scipy import special

__all__ = ["ess", "rhat", "mcse_mean", "bfmi", "summary_min_ess"]


def _split(x):
    """[chains, draws, K] -> [2*chains, draws//2, K]."""
    n = x.shape[1] // 2
    return np.concatenate([x[:, :n], x[:, -n:]], axis=0)


def _rank_normalise(x):
    """Pooled fractional ranks -> normal scores, per column K.  x: [chains, draws, K]."""
    c, n, k = x.shape
    flat = x.reshape(c * n, k)
    order = np.argsort(flat, axis=0, kind="stable")
    ranks = np.empty_like(order, dtype="f8")
    idx = np.arange(1, c * n + 1, dtype="f8")
    for j in range(k):                       # average ranks for ties
        col = flat[order[:, j], j]
        r = idx.copy()
        if np.any(col[1:] == col[:-1]):
            _, inv, cnt = np.unique(col, return_inverse=True, return_counts=True)
            ends = np.cumsum(cnt)
            r = ((ends - cnt + 1 + ends) / 2.0)[inv]
        ranks[order[:, j], j] = r
    z = special.ndtri((ranks - 0.375) / (c * n + 0.25))
    return z.reshape(c, n, k)


def _autocov(x):
    """Biased autocovariance along draws via FFT.  x: [chains, draws, K] -> same shape."""
    n = x.shape[1]
    m = 1
    while m < 2 * n:
        m *= 2
    xc = x - x.mean(axis=1, keepdims=True)
    f = np.fft.rfft(xc, n=m, axis=1)
    ac = np.fft.irfft(f * np.conj(f), n=m, axis=1)[:, :n]
    return ac / n


def _ess_core(x):
    """ESS of [chains, draws, K] by Geyer's initial monotone sequence on the chain-averaged
    autocorrelation with the multi-chain variance estimate (Stan / Vehtari et al. 2021 eq. 10-13)."""
    c, n, k = x.shape
    if n < 4:
        return np.full(k, np.nan)
    acov = _autocov(x)
    chain_mean = x.mean(axis=1)
    mean_var = acov[:, 0].mean(axis=0) * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if c > 1:
        var_plus = var_plus + chain_mean.var(axis=0, ddof=1)
    rho = 1.0 - (mean_var[None, :] - acov.mean(axis=0)) / np.where(var_plus > 0, var_plus, np.nan)[None, :]
    rho[0] = 1.0
    out = np.empty(k)
    total = c * n
    for j in range(k):
        r = rho[:, j]
        if not np.isfinite(r[1]):
            out[j] = np.nan
            continue
        # pair sums P_t = rho_{2t} + rho_{2t+1}: initial positive, then monotone sequence
        npair = (n - 1) // 2
        pairs = r[0:2 * npair:2] + r[1:2 * npair:2]
        neg = np.nonzero(pairs < 0)[0]
        last = neg[0] if len(neg) else npair           # pairs[:last] are positive
        p = np.minimum.accumulate(pairs[:last]) if last > 0 else pairs[:0]
        tau = -1.0 + 2.0 * p.sum()
        # Stan's "improved estimate": add the first even-lag term of the truncated pair if positive
        if last < npair and r[2 * last] > 0:
            tau += r[2 * last]
        tau = max(tau, 1.0 / np.log10(total))
        out[j] = total / tau
    return out


def _as3d(draws):
    a = np.asarray(draws, dtype="f8")
    if a.ndim == 1:
        a = a[None, :]
    shape = a.shape[2:]
    return a.reshape(a.shape[0], a.shape[1], -1), shape


def ess(draws, method="bulk"):
    """Effective sample size.  method: 'bulk' (rank-normalised split chains), 'mean' (split, raw),
    'tail' (min of the 5% / 95% quantile indicators)."""
    x, shape = _as3d(draws)
    if method == "bulk":
        val = _ess_core(_split(_rank_normalise(x)))
    elif method == "mean":
        val = _ess_core(_split(x))
    elif method == "tail":
        lo = np.quantile(x, 0.05, axis=(0, 1))
        hi = np.quantile(x, 0.95, axis=(0, 1))
        val = np.minimum(_ess_core(_split((x <= lo).astype("f8"))), _ess_core(_split((x <= hi).astype("f8"))))
    else:
        raise ValueError("unknown ESS method %r" % method)
    return val.reshape(shape)


def _rhat_core(x):
    c, n, k = x.shape
    chain_mean = x.mean(axis=1)
    chain_var = x.var(axis=1, ddof=1)
    between = n * chain_mean.var(axis=0, ddof=1)
    within = chain_var.mean(axis=0)
    return np.sqrt(((n - 1.0) / n * within + between / n) / within)


def rhat(draws):
    """Rank-normalised split R-hat: max of the bulk and the folded (tail) version."""
    x, shape = _as3d(draws)
    if x.shape[0] < 2 and x.shape[1] < 4:
        return np.full(shape, np.nan)
    bulk = _rhat_core(_split(_rank_normalise(x)))
    folded = np.abs(x - np.median(x, axis=(0, 1)))
    tail = _rhat_core(_split(_rank_normalise(folded)))
    return np.maximum(bulk, tail).reshape(shape)


def mcse_mean(draws):
    """Monte-Carlo standard error of the posterior mean: sd / sqrt(ESS_mean)."""
    x, shape = _as3d(draws)
    sd = x.reshape(-1, x.shape[2]).std(axis=0, ddof=1)
    return (sd / np.sqrt(_ess_core(_split(x)))).reshape(shape)


def bfmi(energy):
    """Energy Bayesian fraction of missing information per chain.  energy: [chains, draws]."""
    e = np.asarray(energy, dtype="f8")
    if e.ndim == 1:
        e = e[None, :]
    return np.square(np.diff(e, axis=1)).mean(axis=1) / e.var(axis=1)


def summary_min_ess(trace, varnames=None):
    """min over every scalar of every (free and back-transformed) variable of the bulk ESS -- the
    numerator of BASELINE.json's `min bulk ESS/sec` (mirrors benchmarks.py:163-169)."""
    names = varnames if varnames is not None else trace.varnames
    worst = np.inf
    for n in names:
        v = np.stack(trace.get_values(n, combine=False, squeeze=False))
        worst = min(worst, float(np.nanmin(ess(v))))
    return worst
