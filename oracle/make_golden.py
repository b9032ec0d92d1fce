"""Generate tests/golden/*.npz from the REAL reference (dherrera1911/sqfa at /root/reference).

Run in the build container (needs /root/reference): `python oracle/make_golden.py`.
The fixtures pin the oracle (tests/test_oracle_golden.py) and the CUDA path
(tests/test_golden_gpu.py) to outputs of the reference's own code; they are small, seeded and
committed together with this script. Everything is computed by the reference in float64 (its own
test-suite runs in float64, tests/test_distances.py:15) except the `*_f32` entries.
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def spd(n, m, g):
    ev = 2 * torch.rand(n, m, generator=g, dtype=torch.float64) ** 2 + 0.01
    low = torch.tril(torch.randn(n, m, m, generator=g, dtype=torch.float64), diagonal=-1)
    Q = torch.matrix_exp(low - low.transpose(1, 2))
    return torch.einsum("ijk,ik,ilk->ijl", Q, ev, Q)


def npz(name, **arrays):
    np.savez_compressed(os.path.join(OUT, name), **{k: np.asarray(v) for k, v in arrays.items()})


def main():
    R = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    torch.set_default_dtype(torch.float64)  # as the reference's tests do (also for torch.eye in OAS)
    g = torch.Generator().manual_seed(20240607)

    # ---- HP1: class statistics on non-trivial data (uneven classes, non-zero means)
    n, d, c = 700, 12, 4
    y = torch.multinomial(torch.tensor([0.4, 0.3, 0.2, 0.1]), n, replacement=True, generator=g)
    X = torch.randn(n, d, generator=g) * (0.5 + torch.rand(d, generator=g)) + 0.3 * y[:, None] + 1.0
    emp = R.statistics.class_statistics(X, y, estimator="empirical")
    oas = R.statistics.class_statistics(X, y, estimator="oas")
    emp32 = R.statistics.class_statistics(X.float(), y, estimator="empirical")
    npz(
        "class_statistics.npz", X=X, y=y,
        means=emp["means"], covariances=emp["covariances"], second_moments=emp["second_moments"],
        oas_covariances=oas["covariances"], oas_second_moments=oas["second_moments"],
        means_f32=emp32["means"], covariances_f32=emp32["covariances"],
        pca3=R.statistics.pca(X, 3), pca_scatter3=R.statistics.pca_from_scatter(emp["second_moments"], 3),
        sample_cov=R.statistics.sample_covariance(X), sample_cov_centered=R.statistics.sample_covariance(X, True),
        perm=torch.sort(y, stable=True).indices,
    )

    # ---- HP2: linalg + distances
    A, B = spd(5, 6, g), spd(3, 6, g)
    mu_a, mu_b = torch.randn(5, 6, generator=g), torch.randn(3, 6, generator=g)
    sa, sb = {"means": mu_a, "covariances": A}, {"means": mu_b, "covariances": B}
    F = torch.randn(3, 6, generator=g)
    D = R.distances
    npz(
        "distances.npz", A=A, B=B, mu_a=mu_a, mu_b=mu_b, F=F,
        conj=R.linalg.conjugate_matrix(A, F),
        geneig_ab=R.linalg.generalized_eigenvalues(A, B), geneig_aa=R.linalg.generalized_eigenvalues(A, A),
        spd_log=R.linalg.spd_log(A),
        ai_sq_ab=D.affine_invariant_sq(A, B), ai_ab=D.affine_invariant(A, B),
        ai_sq_aa=D.affine_invariant_sq(A, A), ai_aa=D.affine_invariant(A, A),
        le_sq_ab=D.log_euclidean_sq(A, B), le_ab=D.log_euclidean(A, B),
        le_sq_aa=D.log_euclidean_sq(A, A), le_aa=D.log_euclidean(A, A),
        fr_sq_ab=D.fisher_rao_lower_bound_sq(sa, sb), fr_ab=D.fisher_rao_lower_bound(sa, sb),
        fr_sq_aa=D.fisher_rao_lower_bound_sq(sa, sa), fr_aa=D.fisher_rao_lower_bound(sa, sa),
        # plug-in distances between Gaussians (distances.py:240-432)
        bhatt_ab=D.bhattacharyya(sa, sb), bhatt_aa=D.bhattacharyya(sa, sa),
        maha_sq_ab=D.mahalanobis_sq(sa, sb), maha_ab=D.mahalanobis(sa, sb), maha_sq_aa=D.mahalanobis_sq(sa, sa),
        hell_ab=D.hellinger(sa, sb), hell_aa=D.hellinger(sa, sa),
        frsc_ab=D.fisher_rao_same_cov(sa, sb), frsc_aa=D.fisher_rao_same_cov(sa, sa),
    )

    # ---- HP2: the closure (loss + gradient) and a converged fit, both model kinds
    out = {}
    F0 = torch.randn(3, d, generator=g)
    stats = {k: v.clone() for k, v in emp.items()}
    tri = torch.tril_indices(c, c, -1)
    for kind, cls in (("sm", R.model.SecondMomentsSQFA), ("full", R.model.SQFA)):
        for dname in (None, "log_euclidean") if kind == "sm" else (None,):
            tag = kind + ("_le" if dname else "")
            dfun = getattr(R.distances, dname) if dname else None
            m = cls(n_dim=d, feature_noise=0.01, n_filters=3, filters=F0.float(), distance_fun=dfun)
            m = m.double()
            dist = m.get_class_distances(stats, regularized=True)
            loss = -dist[tri[0], tri[1]].mean()
            loss.backward()
            out[f"{tag}_dist"] = dist.detach()
            out[f"{tag}_loss"] = loss.detach()
            out[f"{tag}_grad"] = m.parametrizations.filters.original.grad.detach()
        m = cls(n_dim=d, feature_noise=0.01, n_filters=3, filters=F0.float()).double()
        losses, _ = m.fit(data_statistics=stats, max_epochs=200, show_progress=False, return_loss=True)
        out[f"{kind}_fit_losses"] = losses
        out[f"{kind}_fit_filters"] = m.filters.detach()
    npz("closure.npz", F0=F0, noise=0.01, means=stats["means"], covariances=stats["covariances"],
        second_moments=stats["second_moments"], **out)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
