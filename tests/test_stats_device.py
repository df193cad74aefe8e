"""Device (tensor-op) diagnostics against the NumPy implementation they mirror -- run on CPU tensors here, on
the GPU trace in test_gpu_api.py."""
import numpy as np
import pytest
import torch

from pymc3_b200 import stats, stats_device


def _ar1(c, n, k, phi, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros((c, n, k))
    e = rng.normal(size=(c, n, k))
    for t in range(1, n):
        x[:, t] = phi * x[:, t - 1] + e[:, t]
    return x + rng.normal(size=(c, 1, k)) * 0.05


@pytest.mark.parametrize("c,n,k,phi", [(4, 200, 3, 0.0), (8, 301, 5, 0.7), (2, 64, 2, 0.95), (16, 100, 4, -0.4)])
def test_bulk_ess_and_rhat_match_numpy(c, n, k, phi):
    x = _ar1(c, n, k, phi, seed=c + n)
    t = torch.as_tensor(x)
    np.testing.assert_allclose(stats_device.ess_bulk(t).numpy(), stats.ess(x), rtol=1e-9)
    np.testing.assert_allclose(stats_device.rhat(t).numpy(), stats.rhat(x), rtol=1e-9)


def test_ties_get_average_ranks_like_numpy():
    rng = np.random.default_rng(3)
    x = rng.integers(0, 7, size=(4, 150, 2)).astype("f8")          # heavy ties
    t = torch.as_tensor(x)
    np.testing.assert_allclose(stats_device.ess_bulk(t).numpy(), stats.ess(x), rtol=1e-9)
    np.testing.assert_allclose(stats_device.rhat(t).numpy(), stats.rhat(x), rtol=1e-9)


def test_degenerate_inputs():
    const = torch.ones((3, 50, 1), dtype=torch.float64)
    assert torch.isnan(stats_device.ess_bulk(const)).all() == np.isnan(stats.ess(const.numpy())).all()
    short = torch.as_tensor(np.random.default_rng(0).normal(size=(2, 3, 1)))
    assert torch.isnan(stats_device.ess_bulk(short)).all()          # fewer than 4 draws per split chain


def test_engine_trace_layout_helper():
    q = torch.arange(2 * 3 * 4, dtype=torch.float32).reshape(2, 3, 4)        # [draws, chains, D]
    x = stats_device.from_trace(q)
    assert x.shape == (3, 2, 4) and x.dtype == torch.float64 and x[1, 0, 2] == q[0, 1, 2]
