"""CPU, build container only: the oracle against the LIVE reference on fresh random inputs.
Skipped wherever /root/reference is absent (e.g. the GPU box)."""

import pytest
import torch

from oracle import ref_loader
from oracle import sqfa_oracle as O

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")


def spd(n, m, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(n, m, m + 3, generator=g, dtype=torch.float64)
    return a @ a.transpose(1, 2) / (m + 3) + 0.05 * torch.eye(m, dtype=torch.float64)


@pytest.mark.parametrize("seed", [0, 1])
@pytest.mark.parametrize("estimator", ["empirical", "oas"])
def test_class_statistics_live(seed, estimator):
    R = ref_loader.load()
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(500, 9, generator=g) + 0.7
    y = torch.randint(0, 5, (500,), generator=g)
    a, b = R.statistics.class_statistics(X, y, estimator), O.class_statistics(X, y, estimator)
    for k in a:
        assert torch.allclose(a[k], b[k], rtol=1e-5, atol=1e-6), k


@pytest.mark.parametrize("na,nb,m", [(1, 1, 2), (4, 1, 4), (3, 5, 6), (6, 6, 9)])
def test_distances_live(na, nb, m):
    R = ref_loader.load()
    A, B = spd(na, m, 1), spd(nb, m, 2)
    if na == 1:
        A = A[0]
    for name in ("affine_invariant_sq", "affine_invariant", "log_euclidean_sq", "log_euclidean"):
        a, b = getattr(R.distances, name)(A, B), getattr(O, name)(A, B)
        assert a.shape == b.shape, name
        assert torch.allclose(a, b, rtol=1e-8, atol=1e-10), name
    a, b = R.linalg.generalized_eigenvalues(A, B), O.generalized_eigenvalues(A, B)
    assert a.shape == b.shape and torch.allclose(a, b, rtol=1e-8)
    F = torch.randn(2, m, dtype=torch.float64)
    assert torch.allclose(R.linalg.conjugate_matrix(A, F), O.conjugate_matrix(A, F), rtol=1e-10)


@pytest.mark.parametrize("kind", ["second_moments", "full"])
def test_closure_live(kind):
    R = ref_loader.load()
    g = torch.Generator().manual_seed(5)
    X = torch.randn(800, 10, generator=g, dtype=torch.float64) * (0.5 + torch.rand(10, generator=g, dtype=torch.float64))
    y = torch.randint(0, 6, (800,), generator=g)
    X = X + 0.2 * y[:, None]
    stats = R.statistics.class_statistics(X, y)
    F0 = torch.randn(3, 10, generator=g)
    cls = R.model.SecondMomentsSQFA if kind == "second_moments" else R.model.SQFA
    m = cls(n_dim=10, feature_noise=0.01, n_filters=3, filters=F0).double()
    d = m.get_class_distances(stats, regularized=True)
    tri = torch.tril_indices(6, 6, -1)
    loss = -d[tri[0], tri[1]].mean()
    loss.backward()
    l2, g2, d2 = O.loss_and_grad(kind, stats, F0.double(), noise=0.01)
    assert torch.allclose(loss.detach(), l2, rtol=1e-10)
    assert torch.allclose(m.parametrizations.filters.original.grad, g2, rtol=1e-7, atol=1e-10)
