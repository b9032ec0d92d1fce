"""pytest configuration: `gpu` marker, repo root on sys.path, shared synthetic-data helpers."""

import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `-m gpu`")
    config.addinivalue_line("markers", "slow: about a minute of CPU oracle work (deselect with -m 'gpu and not slow')")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def make_class_data(n, d, c, seed=0, mean_scale=0.2, offset=0.0, skew=False, dtype=torch.float32):
    """Synthetic labeled data of SURVEY.md section 8(d): low-rank class-scaled signal + noise +
    class means, optional constant offset (cancellation stress) and skewed class sizes."""
    g = torch.Generator().manual_seed(seed)
    r = min(32, d)
    if skew:
        w = 1.0 / (torch.arange(c, dtype=torch.float64) + 1.0)
        y = torch.multinomial(w / w.sum(), n, replacement=True, generator=g)
    else:
        y = torch.randint(0, c, (n,), generator=g)
    basis = torch.randn(r, d, generator=g, dtype=torch.float64) / r**0.5
    scales = 0.5 + torch.rand(c, generator=g, dtype=torch.float64)
    means = mean_scale * torch.randn(c, d, generator=g, dtype=torch.float64)
    z = torch.randn(n, r, generator=g, dtype=torch.float64) * scales[y][:, None]
    x = z @ basis + 0.5 * torch.randn(n, d, generator=g, dtype=torch.float64) + means[y] + offset
    return x.to(dtype), y


def rel_err(a, b):
    """Relative Frobenius error of a against the (higher precision) truth b."""
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))
