"""CPU: the oracle restatement against golden vectors produced by the REAL reference
(oracle/make_golden.py). This is what pins the oracle; the GPU tests then compare against it."""

import os

import numpy as np
import pytest
import torch

from oracle import sqfa_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with np.load(os.path.join(GOLD, name)) as z:
        return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def close(a, b, tol=1e-10):
    return float((a.double() - b.double()).abs().max()) <= tol * max(1.0, float(b.double().abs().max()))


def test_class_statistics_golden():
    g = load("class_statistics.npz")
    X, y = g["X"], g["y"]
    emp = O.class_statistics(X, y, "empirical")
    assert close(emp["means"], g["means"]) and close(emp["covariances"], g["covariances"])
    assert close(emp["second_moments"], g["second_moments"])
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)  # the goldens were generated like the reference's tests
    try:
        oas = O.class_statistics(X, y, "oas")
    finally:
        torch.set_default_dtype(prev)
    assert close(oas["covariances"], g["oas_covariances"]) and close(oas["second_moments"], g["oas_second_moments"])
    emp32 = O.class_statistics(X.float(), y, "empirical")
    assert close(emp32["means"], g["means_f32"], 1e-6) and close(emp32["covariances"], g["covariances_f32"], 1e-5)
    assert close(O.sample_covariance(X), g["sample_cov"])
    assert close(O.sample_covariance(X, True), g["sample_cov_centered"])
    perm, offsets = O.bucket_permutation(y, 4)
    assert torch.equal(perm, g["perm"])
    assert O.subspace_angle(O.pca(X, 3), g["pca3"]) < 1e-6
    assert O.subspace_angle(O.pca_from_scatter(emp["second_moments"], 3), g["pca_scatter3"]) < 1e-6


def test_distances_golden():
    g = load("distances.npz")
    A, B = g["A"], g["B"]
    sa = {"means": g["mu_a"], "covariances": A}
    sb = {"means": g["mu_b"], "covariances": B}
    assert close(O.conjugate_matrix(A, g["F"]), g["conj"])
    assert close(O.generalized_eigenvalues(A, B), g["geneig_ab"], 1e-9)
    assert close(O.generalized_eigenvalues(A, A), g["geneig_aa"], 1e-9)
    assert close(O.spd_log(A), g["spd_log"], 1e-9)
    for name in ("affine_invariant", "log_euclidean"):
        short = {"affine_invariant": "ai", "log_euclidean": "le"}[name]
        assert close(getattr(O, name + "_sq")(A, B), g[short + "_sq_ab"], 1e-9)
        assert close(getattr(O, name)(A, B), g[short + "_ab"], 1e-9)
        assert close(getattr(O, name + "_sq")(A, A), g[short + "_sq_aa"], 1e-8)
        assert close(getattr(O, name)(A, A), g[short + "_aa"], 1e-6)  # sqrt near 0 on the diagonal
    assert close(O.fisher_rao_lower_bound_sq(sa, sb), g["fr_sq_ab"], 1e-9)
    assert close(O.fisher_rao_lower_bound(sa, sb), g["fr_ab"], 1e-9)
    assert close(O.fisher_rao_lower_bound_sq(sa, sa), g["fr_sq_aa"], 1e-8)


@pytest.mark.parametrize("tag,kind,dist", [("sm", "second_moments", None), ("sm_le", "second_moments", "log_euclidean"),
                                           ("full", "full", None)])
def test_closure_golden(tag, kind, dist):
    g = load("closure.npz")
    stats = {k: g[k] for k in ("means", "covariances", "second_moments")}
    dfun = getattr(O, dist) if dist else None
    # the reference model was built from float32 filters / float32 noise_mat and then .double()d
    noise32 = float(torch.tensor(float(g["noise"]), dtype=torch.float32))
    loss, grad, dmat = O.loss_and_grad(kind, stats, g["F0"].float().double(), noise=noise32, distance=dfun)
    i, j = torch.tril_indices(4, 4, -1)
    assert close(dmat[i, j], g[tag + "_dist"][i, j], 1e-9)
    assert close(loss, g[tag + "_loss"], 1e-10)
    assert close(grad, g[tag + "_grad"], 1e-8)


@pytest.mark.parametrize("tag,kind", [("sm", "second_moments"), ("full", "full")])
def test_fit_golden(tag, kind):
    g = load("closure.npz")
    stats = {k: g[k] for k in ("means", "covariances", "second_moments")}
    noise32 = float(torch.tensor(float(g["noise"]), dtype=torch.float32))
    F, losses, _ = O.fit_lbfgs(kind, stats, g["F0"].float().double(), noise=noise32, max_epochs=200)
    ref_losses = g[tag + "_fit_losses"]
    assert abs(float(losses[-1]) - float(ref_losses[-1])) < 1e-5
    assert abs(float(losses[0]) - float(ref_losses[0])) < 1e-6  # loss lists are stored in float32
    assert O.subspace_angle(F, g[tag + "_fit_filters"]) < 2e-3


def test_plugin_distances_golden():
    """The reference's other distance_fun plug-ins (distances.py:240-432): oracle vs the real reference."""
    g = load("distances.npz")
    sa = {"means": g["mu_a"], "covariances": g["A"]}
    sb = {"means": g["mu_b"], "covariances": g["B"]}
    for name, key in (("bhattacharyya", "bhatt"), ("hellinger", "hell"), ("fisher_rao_same_cov", "frsc")):
        assert close(getattr(O, name)(sa, sb), g[key + "_ab"], 1e-10), name
        assert close(getattr(O, name)(sa, sa), g[key + "_aa"], 1e-9), name
    assert close(O.mahalanobis_sq(sa, sb), g["maha_sq_ab"], 1e-10)
    assert close(O.mahalanobis_sq(sa, sa), g["maha_sq_aa"], 1e-10)
    assert close(O.mahalanobis(sa, sb), g["maha_ab"], 1e-10)
