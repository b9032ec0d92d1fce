"""CPU oracle for the SQFA hot paths -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, in plain torch-CPU tensor algebra, of what the reference computes on the two hot
paths (class statistics; projection -> pairwise SPD distance -> loss -> gradient). Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may
import this module; nothing under `sqfa_b200/` does.

Every function cites the reference lines it follows (paths relative to
/root/reference/src/sqfa/). The floating-point building blocks the reference delegates to torch
(`einsum`/`bmm`, `torch.linalg.eigh`/`eigvalsh` = LAPACK syevd, autograd, `torch.optim.LBFGS`;
torch is unpinned upstream -- `torch>=1.8`, pyproject.toml:16 -- and is 2.11.0 here) are used as
the same library calls, so timing this module on the host is a fair stand-in for the reference's
CPU path.

Parity pin: `tests/golden/*.npz` were produced by the REAL reference (imported from
/root/reference by oracle/make_golden.py in the build container); tests/test_oracle_golden.py
checks this module against them, and tests/test_oracle_vs_reference.py compares it live against
the reference whenever /root/reference is present.
"""

import math

import torch

EPSILON = 1e-6  # distances.py:29


# ------------------------------------------------------------------------------------------------
# HP1: class statistics
# ------------------------------------------------------------------------------------------------
def sample_covariance(points, assume_centered=False):
    """statistics.py:97-124"""
    n = points.shape[0]
    if assume_centered:
        return points.T @ points / n  # :116
    centred = points - points.mean(dim=0)  # :118-119
    return centred.T @ centred / (n - 1)  # :120-122


def oas_covariance(points, assume_centered=False):
    """statistics.py:57-94 (OAS shrinkage, Chen et al. 2010)"""
    n, d = points.shape
    s = sample_covariance(points, assume_centered=assume_centered)  # :81
    tr = torch.trace(s)  # :84
    tr_sq = (s * s).sum()  # :85
    rho = ((1 - 2 / d) * tr_sq + tr * tr) / ((n + 1 - 2 / d) * (tr_sq - tr * tr / d))  # :86-88
    rho = rho if bool(rho < 1.0) else 1.0  # python min(1.0, rho), :89 (NaN -> 1.0)
    target = torch.eye(d) * tr / d  # :92 (default-dtype eye, as upstream)
    return (1 - rho) * s + rho * target  # :93


def bucket_permutation(labels, n_classes):
    """Row ids of every class in ascending order, concatenated -- what statistics.py:37 builds one
    class at a time. Returns (perm, offsets[n_classes+1])."""
    chunks, offsets = [], [0]
    for c in range(n_classes):
        idx = (labels == c).nonzero().squeeze(1)  # :37
        chunks.append(idx)
        offsets.append(offsets[-1] + idx.numel())
    perm = torch.cat(chunks) if chunks else torch.zeros(0, dtype=torch.int64)
    return perm, torch.tensor(offsets, dtype=torch.int64)


def class_statistics(points, labels, estimator="empirical"):
    """statistics.py:8-54"""
    n_classes = int(labels.max() + 1)  # :29
    d = points.shape[-1]
    means = torch.zeros(n_classes, d, dtype=points.dtype)  # :32-34
    covs = torch.zeros(n_classes, d, d, dtype=points.dtype)
    sms = torch.zeros(n_classes, d, d, dtype=points.dtype)
    for c in range(n_classes):  # :36
        rows = points[(labels == c).nonzero().squeeze(1)]  # :37-38
        means[c] = rows.mean(dim=0)  # :40
        if estimator == "empirical":
            cov = sample_covariance(rows)  # :43
        elif estimator == "oas":
            cov = oas_covariance(rows)  # :45
        else:
            raise ValueError(estimator)
        covs[c] = torch.as_tensor(cov, dtype=points.dtype)  # :46
        sms[c] = covs[c] + torch.outer(means[c], means[c])  # :47
    return {"means": means, "covariances": covs, "second_moments": sms}


def pca(points, n_components=None):
    """statistics.py:127-160"""
    n, d = points.shape
    if n_components is None:
        n_components = min(n, d)
    if n_components > d:
        raise ValueError("n_components must be less than or equal to n_dim.")
    _, vecs = torch.linalg.eigh(sample_covariance(points))  # :152-155
    return torch.flip(vecs[:, -n_components:], dims=[1]).T  # :157-158


def pca_from_scatter(scatters, n_components=None):
    """statistics.py:163-192 -- NB upstream feeds the mean scatter to pca() as a point cloud."""
    d = scatters.shape[-1]
    if n_components is None:
        n_components = d
    if n_components > d:
        raise ValueError("n_components must be less than or equal to n_dim.")
    return pca(scatters.mean(dim=0), n_components=n_components)  # :189-190


# ------------------------------------------------------------------------------------------------
# HP2: linalg + distances
# ------------------------------------------------------------------------------------------------
def conjugate_matrix(A, B):
    """linalg.py:19-45: out[n, b] = B_b A_n B_b^T, size-1 batch dims squeezed."""
    if A.dim() == 2:
        A = A.unsqueeze(0)
    if B.dim() < 2:
        raise ValueError("B must have at least 2 dimensions.")
    if B.dim() == 2:
        out = torch.matmul(torch.matmul(B, A), B.T)  # (n, o, o)
        return torch.squeeze(out, dim=0)
    out = torch.matmul(torch.matmul(B[None], A[:, None]), B.transpose(-2, -1)[None])  # (n, b, o, o)
    return torch.squeeze(out, dim=(0, 1))


def spd_inv_sqrt(M):
    """linalg.py:144-162: whitening diag(lambda^-1/2) V^T (not the symmetric root)."""
    lam, V = torch.linalg.eigh(M)
    return (V * torch.sqrt(1.0 / lam).unsqueeze(-2)).transpose(-2, -1)


def spd_log(M):
    """linalg.py:165-183"""
    lam, V = torch.linalg.eigh(M)
    return (V * torch.log(lam).unsqueeze(-2)) @ V.transpose(-2, -1)


def generalized_eigenvalues(A, B):
    """linalg.py:48-70: eigvalsh(W_b A_n W_b^T), descending."""
    W = spd_inv_sqrt(B)  # :67
    return torch.linalg.eigvalsh(conjugate_matrix(A, W)).flip(-1)  # :68-70


def affine_invariant_sq(A, B):
    """distances.py:46-67"""
    return (torch.log(generalized_eigenvalues(A, B)) ** 2).sum(dim=-1)


def affine_invariant(A, B):
    """distances.py:70-89"""
    return torch.sqrt(affine_invariant_sq(A, B) + EPSILON)


def log_euclidean_sq(A, B):
    """distances.py:92-116"""
    if A.dim() == 2:
        A = A.unsqueeze(0)
    diff = spd_log(A)[:, None] - spd_log(B)[None]  # :111-114
    return torch.squeeze((diff * diff).sum(dim=(-2, -1)))  # :115-116


def log_euclidean(A, B):
    """distances.py:119-138"""
    return torch.sqrt(log_euclidean_sq(A, B) + EPSILON)


def embed_gaussian(stats):
    """distances.py:141-174: [[Sigma + mu mu^T, mu], [mu^T, 1]]"""
    mu, cov = stats["means"], stats["covariances"]
    if mu.dim() == 1:
        mu = mu.unsqueeze(0)
    if cov.dim() == 2:
        cov = cov.unsqueeze(0)
    c, k = mu.shape
    E = torch.zeros(c, k + 1, k + 1, dtype=mu.dtype)
    E[:, :k, :k] = cov + mu[:, :, None] * mu[:, None, :]  # :167-168
    E[:, :k, k] = mu  # :170-173
    E[:, k, :k] = mu
    E[:, k, k] = 1.0
    return E


def fisher_rao_lower_bound_sq(stats_a, stats_b):
    """distances.py:177-207"""
    return affine_invariant_sq(embed_gaussian(stats_a), embed_gaussian(stats_b)) / 2


def fisher_rao_lower_bound(stats_a, stats_b):
    """distances.py:210-237"""
    return torch.sqrt(fisher_rao_lower_bound_sq(stats_a, stats_b) + EPSILON)


def _pair_mean_cov(stats_a, stats_b):
    mu_a, mu_b = stats_a["means"], stats_b["means"]
    cov_a, cov_b = stats_a["covariances"], stats_b["covariances"]
    if mu_a.dim() == 1:
        mu_a = mu_a.unsqueeze(0)
    if mu_b.dim() == 1:
        mu_b = mu_b.unsqueeze(0)
    if cov_a.dim() == 2:
        cov_a = cov_a.unsqueeze(0)
    if cov_b.dim() == 2:
        cov_b = cov_b.unsqueeze(0)
    return mu_a[:, None] - mu_b[None, :], (cov_a[:, None] + cov_b[None, :]) / 2, cov_a, cov_b


def mahalanobis_sq(stats_a, stats_b):
    """distances.py:283-330: d^T ((Sigma_a + Sigma_b) / 2)^-1 d for every pair (not squeezed)."""
    diff, mean_cov, _, _ = _pair_mean_cov(stats_a, stats_b)
    return torch.einsum("ijk,ijkl,ijl->ij", diff, torch.linalg.inv(mean_cov), diff)  # :323-329


def mahalanobis(stats_a, stats_b):
    """distances.py:333-361"""
    return torch.sqrt(mahalanobis_sq(stats_a, stats_b) + EPSILON)


def bhattacharyya(stats_a, stats_b):
    """distances.py:240-280: maha / 8 + (logdet mean_cov - (logdet A + logdet B) / 2) / 2, squeezed."""
    diff, mean_cov, cov_a, cov_b = _pair_mean_cov(stats_a, stats_b)
    term1 = torch.einsum("ijk,ijkl,ijl->ij", diff, torch.linalg.inv(mean_cov), diff)  # :268-270
    term2 = torch.logdet(mean_cov) - (torch.logdet(cov_a)[:, None] + torch.logdet(cov_b)[None, :]) * 0.5  # :272-276
    return torch.squeeze(term1 / 8 + term2 * 0.5)  # :278-280


def hellinger(stats_a, stats_b):
    """distances.py:364-393"""
    return torch.sqrt(1 - torch.exp(-bhattacharyya(stats_a, stats_b)) + EPSILON)


def fisher_rao_same_cov(stats_a, stats_b):
    """distances.py:396-432"""
    return 2.0**0.5 * torch.acosh(1 + mahalanobis_sq(stats_a, stats_b) / 4)


# ------------------------------------------------------------------------------------------------
# HP2: model forward and the closure loss
# ------------------------------------------------------------------------------------------------
def apply_constraint(raw, constraint):
    """constraints.py:37 (sphere: row-normalise) / :75 (identity)."""
    if constraint == "sphere":
        return raw / raw.norm(dim=-1, keepdim=True)
    if constraint == "none":
        return raw
    raise ValueError(f"oracle supports constraints 'sphere' and 'none', got {constraint!r}")


def stats_to_scatter(stats):
    """model.py:24-53"""
    if isinstance(stats, dict):
        mu = stats["means"]
        return stats["covariances"] + mu[:, :, None] * mu[:, None, :]  # :46-49
    return stats


def class_distances_second_moments(stats, filters, noise=0.0, distance=affine_invariant):
    """SecondMomentsSQFA.get_class_distances(regularized=True), model.py:190-220."""
    psi = conjugate_matrix(stats_to_scatter(stats), filters)  # :212-214
    psi = psi + noise * torch.eye(filters.shape[0], dtype=psi.dtype)[None]  # :216-217
    return distance(psi, psi)  # :219


def class_distances_full(stats, filters, noise=0.0, distance=fisher_rao_lower_bound):
    """SQFA.get_class_distances(regularized=True), model.py:508-546."""
    mu = stats["means"] @ filters.T  # :534 -> :236
    cov = conjugate_matrix(stats["covariances"], filters)  # :535
    cov = cov + noise * torch.eye(filters.shape[0], dtype=cov.dtype)[None]  # :537-538
    fs = {"means": mu, "covariances": cov}
    return distance(fs, fs)  # :545


def closure_loss(distances):
    """_optim.py:88-94: minus the mean over the strict lower triangle; NaN/inf guard :16-30."""
    c = distances.shape[0]
    i, j = torch.tril_indices(c, c, offset=-1)
    tril = distances[i, j]
    if torch.isnan(tril).any():
        raise ValueError("Some distances between classes are NaN.")
    if torch.isinf(tril).any():
        raise ValueError("Some distances between classes are inf.")
    return -tril.mean()


def loss_and_grad(kind, stats, raw_filters, noise=0.0, constraint="sphere", distance=None):
    """One closure evaluation (_optim.py:90-96): loss and d loss / d raw_filters via autograd.
    kind: "second_moments" (SecondMomentsSQFA) or "full" (SQFA)."""
    raw = raw_filters.detach().clone().requires_grad_(True)
    F = apply_constraint(raw, constraint)
    if kind == "second_moments":
        d = class_distances_second_moments(stats, F, noise, distance or affine_invariant)
    elif kind == "full":
        d = class_distances_full(stats, F, noise, distance or fisher_rao_lower_bound)
    else:
        raise ValueError(kind)
    loss = closure_loss(d)
    loss.backward()
    return loss.detach(), raw.grad.detach(), d.detach()


def fit_lbfgs(kind, stats, raw_filters, noise=0.0, constraint="sphere", distance=None, max_epochs=300, lr=0.1,
              atol=1e-6, **lbfgs_kwargs):
    """fitting_loop, _optim.py:78-134: torch LBFGS on the raw filters, stop after 3 consecutive
    epochs with |delta loss| < atol. Returns (constrained filters, loss per epoch, closure evals)."""
    raw = raw_filters.detach().clone().requires_grad_(True)
    opt = torch.optim.LBFGS([raw], lr=lr, **lbfgs_kwargs)  # :78-82
    n_eval = [0]

    def closure():  # :90-96
        opt.zero_grad()
        F = apply_constraint(raw, constraint)
        if kind == "second_moments":
            d = class_distances_second_moments(stats, F, noise, distance or affine_invariant)
        else:
            d = class_distances_full(stats, F, noise, distance or fisher_rao_lower_bound)
        loss = closure_loss(d)
        loss.backward()
        n_eval[0] += 1
        return loss

    losses, prev, hits = [], 0.0, 0
    for _ in range(max_epochs):  # :105
        cur = opt.step(closure).item()  # :108
        hits = hits + 1 if abs(prev - cur) < atol else 0  # :111-123
        prev = cur
        losses.append(cur)
        if hits >= 3:  # :130
            break
    return apply_constraint(raw.detach(), constraint), torch.tensor(losses), n_eval[0]


def subspace_angle(F1, F2):
    """Largest principal angle (radians) between the row spaces of two (k, D) filter matrices."""
    q1, _ = torch.linalg.qr(F1.double().T)
    q2, _ = torch.linalg.qr(F2.double().T)
    s = torch.linalg.svdvals(q1.T @ q2).clamp(max=1.0)
    return math.acos(float(s.min()))
