"""The oracle against the known answers and properties the reference's OWN test-suite pins:

* generalized eigenvalues / matrix square root / matrix logarithm against scipy
  (/root/reference/tests/test_linalg.py:182-245, 334-389), whitening identity (:312-331),
  conjugation against the plain triple product (:264-309);
* affine-invariant / log-Euclidean: symmetric self-distance matrix, zero diagonal, invariance to
  inversion, AI(A, I) == LE(A, I) (/root/reference/tests/test_distances.py:40-97); Fisher-Rao lower
  bound and the mean-covariance plug-ins: symmetric, zero diagonal (:100-170);
* class statistics of constant / hand-computable data (/root/reference/tests/test_statistics.py:24-50).

These run on the CPU (no GPU, no /root/reference needed): they pin the CHECKER. The CUDA path is
compared with the same oracle in the `-m gpu` tests and repeats these properties there
(tests/test_hp2_gpu.py::test_distance_properties, ::test_generalized_eigenvalues_and_spd_functions,
tests/test_hp1_gpu.py::test_class_statistics_constant_data).
"""

import os
import sys

import pytest
import scipy.linalg
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import sqfa_oracle as O  # noqa: E402

F64 = torch.float64


def random_spd(n, m, seed):
    """SPD matrices with a spectrum in [0.01, 2.01] (the range of the reference's fixture, make_examples.py:15-20)."""
    g = torch.Generator().manual_seed(seed)
    lam = 2 * torch.rand(n, m, generator=g, dtype=F64) ** 2 + 0.01
    Q, _ = torch.linalg.qr(torch.randn(n, m, m, generator=g, dtype=F64))
    return (Q * lam[:, None, :]) @ Q.mT


def gaussians(n, m, seed):
    g = torch.Generator().manual_seed(seed + 100)
    return {"means": torch.randn(n, m, generator=g, dtype=F64), "covariances": random_spd(n, m, seed)}


@pytest.mark.parametrize("n_a,n_b", [(1, 1), (4, 1), (4, 8), (8, 8)])
@pytest.mark.parametrize("m", [2, 4, 6])
def test_generalized_eigenvalues_against_scipy(n_a, n_b, m):
    A, B = random_spd(n_a, m, 1), random_spd(n_b, m, 2)
    got = O.generalized_eigenvalues(A, B).reshape(n_a, n_b, m)
    for i in range(n_a):
        for j in range(n_b):
            want = scipy.linalg.eigh(A[i].numpy(), B[j].numpy(), eigvals_only=True)[::-1].copy()
            assert torch.allclose(got[i, j], torch.from_numpy(want), rtol=1e-9, atol=1e-11)
    # descending order, all positive (linalg.py:69-70)
    assert (got[..., :-1] >= got[..., 1:]).all() and (got > 0).all()


@pytest.mark.parametrize("m", [2, 4, 17])
def test_spd_log_and_inv_sqrt_against_scipy(m):
    A = random_spd(4, m, 3)
    logs = O.spd_log(A)
    W = O.spd_inv_sqrt(A)
    eye = torch.eye(m, dtype=F64)
    for i in range(4):
        want = torch.from_numpy(scipy.linalg.logm(A[i].numpy()).real)
        assert torch.allclose(logs[i], want, atol=1e-8)
        # whitening, not the symmetric root: W A W^T = I (test_linalg.py:312-331)
        assert torch.allclose(W[i] @ A[i] @ W[i].T, eye, atol=1e-8)
        # W^T W is the inverse
        assert torch.allclose(W[i].T @ W[i], torch.linalg.inv(A[i]), rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize("m", [2, 4, 6])
@pytest.mark.parametrize("n", [1, 4, 8])
@pytest.mark.parametrize("k", [1, 4, 7])
def test_conjugate_matrix_is_the_triple_product(m, n, k):
    A = random_spd(n, m, 4)
    F = torch.randn(k, m, generator=torch.Generator().manual_seed(5), dtype=F64)
    got = O.conjugate_matrix(A, F)
    want = torch.stack([F @ A[i] @ F.T for i in range(n)])
    if n == 1:  # size-1 batch dimensions are squeezed (linalg.py:44-45)
        want = want[0]
    if k == 1:
        want = want.reshape(got.shape)
    assert got.shape == want.shape
    assert torch.allclose(got, want, atol=1e-12)


@pytest.mark.parametrize("n", [1, 4, 8])
@pytest.mark.parametrize("m", [2, 4, 6])
def test_spd_distance_properties(n, m):
    A = random_spd(n, m, 6)
    Ainv = torch.linalg.inv(A)
    eye = torch.eye(m, dtype=F64)
    for dist in (O.affine_invariant_sq, O.log_euclidean_sq):
        D = dist(A, A)
        assert D.shape == ((n, n) if n > 1 else ())  # squeezed for single matrices
        D = D.reshape(n, n)
        assert torch.allclose(D, D.T, atol=1e-9)
        assert torch.allclose(D.diagonal(), torch.zeros(n, dtype=F64), atol=1e-9)
        assert torch.allclose(D, dist(Ainv, Ainv).reshape(n, n), atol=1e-7)
    assert torch.allclose(O.affine_invariant_sq(A, eye), O.log_euclidean_sq(A, eye), atol=1e-9)
    # the non-squared forms: sqrt(d^2 + 1e-6) (distances.py:29, 89, 138)
    assert torch.allclose(O.affine_invariant(A, A) ** 2, O.affine_invariant_sq(A, A) + O.EPSILON, atol=1e-12)
    assert torch.allclose(O.log_euclidean(A, A) ** 2, O.log_euclidean_sq(A, A) + O.EPSILON, atol=1e-12)


@pytest.mark.parametrize("n", [1, 4, 8])
@pytest.mark.parametrize("m", [2, 4, 6])
def test_gaussian_distance_properties(n, m):
    st = gaussians(n, m, 7)
    zero = torch.zeros(n, dtype=F64)
    for dist in (O.fisher_rao_lower_bound_sq, O.bhattacharyya, O.mahalanobis_sq, O.hellinger, O.fisher_rao_same_cov):
        D = dist(st, st).reshape(n, n)
        assert torch.isfinite(D).all()
        assert torch.allclose(D, D.T, atol=1e-9), dist.__name__
        # hellinger is sqrt(1 - exp(-bhattacharyya) + 1e-6): its diagonal is the constant sqrt(EPSILON) (distances.py:392)
        diag = zero + O.EPSILON**0.5 if dist is O.hellinger else zero
        assert torch.allclose(D.diagonal(), diag, atol=1e-7), dist.__name__
        off = D[~torch.eye(n, dtype=torch.bool)]
        assert (off > 0).all()
    # equal covariances: the lower bound's embedding distance reduces to a function of the Mahalanobis
    # distance only, and Bhattacharyya to Mahalanobis^2 / 8
    same = {"means": st["means"], "covariances": st["covariances"][:1].expand(n, m, m).clone()}
    if n > 1:
        assert torch.allclose(O.bhattacharyya(same, same), O.mahalanobis_sq(same, same) / 8, atol=1e-10)


def test_class_statistics_known_answers():
    # constant data per class: mean = the constant, zero covariance, second moment = mu mu^T
    # (test_statistics.py:24-50)
    n_per, D, C = 7, 5, 4
    labels = torch.arange(C).repeat_interleave(n_per)
    X = (labels.to(F64) + 1)[:, None] * torch.ones(D, dtype=F64)
    st = O.class_statistics(X, labels)
    for c in range(C):
        mu = torch.full((D,), c + 1.0, dtype=F64)
        assert torch.allclose(st["means"][c], mu, atol=1e-14)
        assert torch.allclose(st["covariances"][c], torch.zeros(D, D, dtype=F64), atol=1e-14)
        assert torch.allclose(st["second_moments"][c], torch.outer(mu, mu), atol=1e-13)
    # a hand-computable two-class case, shuffled rows, a label gap (class 1 is empty -> NaN like the reference)
    X = torch.tensor([[1.0, 0.0], [0.0, 4.0], [3.0, 0.0], [0.0, 0.0]], dtype=F64)
    y = torch.tensor([0, 2, 0, 2])
    st = O.class_statistics(X, y)
    assert st["means"].shape == (3, 2)  # C = max(label) + 1 (statistics.py:29)
    assert torch.equal(st["means"][0], torch.tensor([2.0, 0.0], dtype=F64))
    assert torch.equal(st["covariances"][0], torch.tensor([[2.0, 0.0], [0.0, 0.0]], dtype=F64))
    assert torch.equal(st["means"][2], torch.tensor([0.0, 2.0], dtype=F64))
    assert torch.equal(st["covariances"][2], torch.tensor([[0.0, 0.0], [0.0, 8.0]], dtype=F64))
    assert torch.isnan(st["means"][1]).all()
    # the bucket permutation is the stable sort of the labels (statistics.py:37, ascending row ids per class)
    perm, offsets = O.bucket_permutation(y, 3)
    assert perm.tolist() == [0, 2, 1, 3] and offsets.tolist() == [0, 2, 2, 4]
    assert torch.equal(perm, torch.sort(y, stable=True).indices)


def test_pca_known_answer():
    # axis-aligned variances: components are the coordinate axes in descending variance (statistics.py:127-160)
    g = torch.Generator().manual_seed(0)
    X = torch.randn(4000, 4, generator=g, dtype=F64) * torch.tensor([1.0, 3.0, 0.5, 2.0], dtype=F64)
    comps = O.pca(X, 3)
    assert comps.shape == (3, 4)
    assert comps.abs().argmax(dim=1).tolist() == [1, 3, 0]
    assert torch.allclose(comps @ comps.T, torch.eye(3, dtype=F64), atol=1e-10)
    with pytest.raises(ValueError):
        O.pca(X, 5)
