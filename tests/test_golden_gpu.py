"""GPU: the CUDA path against golden vectors produced by the REAL reference."""

import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import sqfa_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with np.load(os.path.join(GOLD, name)) as z:
        return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def test_class_statistics_golden():
    from sqfa_b200 import statistics as S

    g = load("class_statistics.npz")
    X, y = g["X"].float().cuda(), g["y"].cuda()
    emp = S.class_statistics(X, y)
    for key in ("means", "covariances", "second_moments"):
        assert rel_err(emp[key], g[key]) < 1e-5, key
    oas = S.class_statistics(X, y, estimator="oas")
    assert rel_err(oas["covariances"], g["oas_covariances"]) < 1e-5
    assert rel_err(oas["second_moments"], g["oas_second_moments"]) < 1e-5
    perm, offsets, counts = S.bucket_labels(y)
    assert torch.equal(perm.cpu().long(), g["perm"])  # bit-exact bucketing
    assert rel_err(S.sample_covariance(X), g["sample_cov"]) < 1e-5
    assert rel_err(S.sample_covariance(X, assume_centered=True), g["sample_cov_centered"]) < 1e-5
    assert O.subspace_angle(S.pca(X, 3).cpu(), g["pca3"]) < 1e-3


def test_distances_golden():
    from sqfa_b200 import distances as Dn
    from sqfa_b200 import linalg as Ln

    g = load("distances.npz")
    A, B = g["A"].float().cuda(), g["B"].float().cuda()
    sa = {"means": g["mu_a"].float().cuda(), "covariances": A}
    sb = {"means": g["mu_b"].float().cuda(), "covariances": B}
    assert rel_err(Ln.conjugate_matrix(A, g["F"].float().cuda()), g["conj"]) < 1e-5
    assert rel_err(Ln.generalized_eigenvalues(A, B), g["geneig_ab"]) < 1e-4
    assert rel_err(Ln.spd_log(A), g["spd_log"]) < 1e-4
    i, j = torch.tril_indices(5, 5, -1)
    for fn, key in ((Dn.affine_invariant_sq, "ai_sq"), (Dn.affine_invariant, "ai"), (Dn.log_euclidean_sq, "le_sq"),
                    (Dn.log_euclidean, "le")):
        assert rel_err(fn(A, B), g[key + "_ab"]) < 1e-4, key
        assert rel_err(fn(A, A)[i, j], g[key + "_aa"][i, j]) < 1e-4, key
    assert rel_err(Dn.fisher_rao_lower_bound_sq(sa, sb), g["fr_sq_ab"]) < 1e-4
    assert rel_err(Dn.fisher_rao_lower_bound(sa, sb), g["fr_ab"]) < 1e-4
    assert rel_err(Dn.fisher_rao_lower_bound(sa, sa)[i, j], g["fr_aa"][i, j]) < 1e-4


@pytest.mark.parametrize("tag,kind,dist", [("sm", "sm", None), ("sm_le", "sm", "log_euclidean"), ("full", "full", None)])
def test_closure_and_fit_golden(tag, kind, dist):
    from sqfa_b200 import distances as Dn
    from sqfa_b200.model import SQFA, SecondMomentsSQFA

    g = load("closure.npz")
    stats = {k: g[k].float().cuda() for k in ("means", "covariances", "second_moments")}
    cls = SecondMomentsSQFA if kind == "sm" else SQFA
    dfun = getattr(Dn, dist) if dist else None
    m = cls(n_dim=12, feature_noise=0.01, n_filters=3, filters=g["F0"].float(), distance_fun=dfun).cuda()
    out = m._fused_loss_plan(stats)()
    out[0].backward()
    assert abs(float(out[0]) - float(g[tag + "_loss"])) < 1e-4 * abs(float(g[tag + "_loss"]))
    assert rel_err(m.parametrizations.filters.original.grad, g[tag + "_grad"]) < 2e-3
    if dist is None:
        # Converged fit. The default stopping rule (3 epochs with |delta loss| < 1e-6) fires while float32
        # L-BFGS still creeps along the flat optimum at ~1e-6 per epoch; WHERE it fires depends on the
        # rounding inside the optimiser's dot products (measured on this problem: angle to the reference's
        # float64 filters 1.5e-3 or 7e-5 with the stock atol, for two summation orders of the same
        # algorithm), so the fit is run to its float32 fixed point (atol = 1e-9: 1.0e-4 / 5e-5). The
        # reference's own float32 fit ends 3e-5 .. 7e-5 from its float64 fit.
        m = cls(n_dim=12, feature_noise=0.01, n_filters=3, filters=g["F0"].float())
        losses, _ = m.fit(data_statistics=stats, max_epochs=200, show_progress=False, return_loss=True, atol=1e-9)
        ref = g[kind + "_fit_losses"]
        assert abs(float(losses[-1]) - float(ref[-1])) < 1e-4 * abs(float(ref[-1]))
        # north_star: learned filters within 1e-3 in subspace angle
        angle = O.subspace_angle(m.filters.detach().cpu(), g[kind + "_fit_filters"])
        print(f"{kind}: converged fit, subspace angle to the reference's float64 filters {angle:.2e}")
        assert angle < 1e-3
        # with the reference's default atol the loss still agrees to 1e-4 (the filters to a few 1e-3)
        m = cls(n_dim=12, feature_noise=0.01, n_filters=3, filters=g["F0"].float())
        losses, _ = m.fit(data_statistics=stats, max_epochs=200, show_progress=False, return_loss=True)
        assert abs(float(losses[-1]) - float(ref[-1])) < 1e-4 * abs(float(ref[-1]))
        assert O.subspace_angle(m.filters.detach().cpu(), g[kind + "_fit_filters"]) < 5e-3
