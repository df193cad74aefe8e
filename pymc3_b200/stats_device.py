"""Rank-normalised split bulk ESS and R-hat computed where the trace lives.

Same estimators as `pymc3_b200.stats` (Vehtari et al. 2021; Geyer's initial monotone sequence as in Stan;
reference call sites: pymc3/stats/__init__.py:42-55, backends/report.py:101-168), written with tensor ops
(sort, scatter, FFT, cumulative min) so that a `[draws, chains, D]` device trace of thousands of chains
never has to be copied to the host for its convergence checks.  Device plumbing only -- there is no custom
kernel here; `tests/test_stats_device.py` pins it to the NumPy implementation.
Input layout: `[chains, draws, K]` tensors (any device); results are `[K]` float64 tensors on that device.
"""
import math

import torch

__all__ = ["ess_bulk", "rhat", "from_trace"]


def from_trace(q):
    """engine trace layout [draws, chains, D] -> [chains, draws, D] float64 (a view + cast, no host copy)"""
    return q.permute(1, 0, 2).to(torch.float64)


def _split(x):
    n = x.shape[1] // 2
    return torch.cat([x[:, :n], x[:, -n:]], dim=0)


def _rank_normalise(x):
    """pooled fractional ranks (ties get their average rank) -> normal scores, per column"""
    c, n, k = x.shape
    total = c * n
    flat = x.reshape(total, k)
    order = torch.argsort(flat, dim=0, stable=True)
    sorted_vals = torch.gather(flat, 0, order)
    pos = torch.arange(total, device=x.device, dtype=torch.int64)[:, None].expand(total, k)
    first = torch.ones((total, k), dtype=torch.bool, device=x.device)
    first[1:] = sorted_vals[1:] != sorted_vals[:-1]
    last = torch.ones((total, k), dtype=torch.bool, device=x.device)
    last[:-1] = first[1:]
    start = torch.where(first, pos, torch.zeros_like(pos)).cummax(dim=0).values
    end = torch.where(last, pos, torch.full_like(pos, total - 1)).flip(0).cummin(dim=0).values.flip(0)
    avg_rank = (start + end).to(torch.float64) * 0.5 + 1.0
    ranks = torch.empty((total, k), dtype=torch.float64, device=x.device)
    ranks.scatter_(0, order, avg_rank)
    z = torch.special.ndtri((ranks - 0.375) / (total + 0.25))
    return z.reshape(c, n, k)


def _autocov(x):
    n = x.shape[1]
    m = 1
    while m < 2 * n:
        m *= 2
    xc = x - x.mean(dim=1, keepdim=True)
    f = torch.fft.rfft(xc, n=m, dim=1)
    ac = torch.fft.irfft(f * torch.conj(f), n=m, dim=1)[:, :n]
    return ac / n


def _ess_core(x):
    c, n, k = x.shape
    if n < 4:
        return torch.full((k,), float("nan"), dtype=torch.float64, device=x.device)
    acov = _autocov(x)
    chain_mean = x.mean(dim=1)
    mean_var = acov[:, 0].mean(dim=0) * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if c > 1:
        var_plus = var_plus + chain_mean.var(dim=0, unbiased=True)
    denom = torch.where(var_plus > 0, var_plus, torch.full_like(var_plus, float("nan")))
    rho = 1.0 - (mean_var[None, :] - acov.mean(dim=0)) / denom[None, :]          # [n, k]
    rho[0] = 1.0
    total = c * n
    npair = (n - 1) // 2
    pairs = rho[0:2 * npair:2] + rho[1:2 * npair:2]                                # [npair, k]
    positive_prefix = (pairs < 0).to(torch.int64).cumsum(dim=0) == 0               # pairs before the first negative one
    last = positive_prefix.sum(dim=0)                                              # index of the first negative pair
    mono = torch.cummin(torch.where(positive_prefix, pairs, torch.full_like(pairs, float("inf"))), dim=0).values
    tau = -1.0 + 2.0 * torch.where(positive_prefix, mono, torch.zeros_like(mono)).sum(dim=0)
    # Stan's "improved estimate": the first even-lag term of the truncated pair, if positive
    idx = torch.clamp(2 * last, max=n - 1)
    extra = torch.gather(rho, 0, idx[None, :])[0]
    tau = tau + torch.where((last < npair) & (extra > 0), extra, torch.zeros_like(extra))
    tau = torch.clamp(tau, min=1.0 / math.log10(total))
    out = total / tau
    return torch.where(torch.isfinite(rho[1]), out, torch.full_like(out, float("nan")))


def ess_bulk(x):
    """bulk effective sample size of `[chains, draws, K]` (rank-normalised, split chains)"""
    x = x.to(torch.float64)
    return _ess_core(_split(_rank_normalise(x)))


def _rhat_core(x):
    c, n, k = x.shape
    chain_mean = x.mean(dim=1)
    chain_var = x.var(dim=1, unbiased=True)
    between = n * chain_mean.var(dim=0, unbiased=True)
    within = chain_var.mean(dim=0)
    return torch.sqrt(((n - 1.0) / n * within + between / n) / within)


def rhat(x):
    """rank-normalised split R-hat: max of the bulk and the folded (tail) version"""
    x = x.to(torch.float64)
    bulk = _rhat_core(_split(_rank_normalise(x)))
    srt = torch.sort(x.reshape(-1, x.shape[2]), dim=0).values            # numpy's median: mean of the two middle values
    m = srt.shape[0]
    med = 0.5 * (srt[(m - 1) // 2] + srt[m // 2])
    folded = (x - med[None, None, :]).abs()
    tail = _rhat_core(_split(_rank_normalise(folded)))
    return torch.maximum(bulk, tail)
