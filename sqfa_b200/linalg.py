"""Utility functions for matrix algebra -- B200-native drop-in for `sqfa.linalg`.

Public names and shape conventions follow /root/reference/src/sqfa/linalg.py. The functions on
the SQFA hot path (`conjugate_matrix` with a filter matrix, `generalized_eigenvalues`,
`spd_log`, `spd_inv_sqrt`) run in the sm_100a kernels; `spd_sqrt` and `generalized_eigenvectors`
are not used by any model path and are composed from the native
eigendecomposition with a few device-side torch ops.
"""

import torch

from . import _lib, _ops

__all__ = [
    "conjugate_matrix",
    "generalized_eigenvalues",
    "generalized_eigenvectors",
    "spd_sqrt",
    "spd_log",
    "spd_inv_sqrt",
]


def __dir__():
    return __all__


def _to_device(*tensors):
    dev = _lib.compute_device(*tensors)
    return dev, [None if t is None else t.to(dev) for t in tensors]


def conjugate_matrix(A, B):
    """
    Conjugate matrix A by B, i.e. compute B A B^T (reference linalg.py:19-45).

    A: (n_batch_A, n_dim, n_dim) or (n_dim, n_dim); B: (n_out, n_dim) or (n_batch_B, n_out, n_dim).
    Returns (n_batch_A, n_batch_B, n_out, n_out) with size-1 batch dimensions squeezed.
    With a 2-D B (a filter matrix, n_out <= 32) and symmetric float32 A this is the native
    projection kernel (one streaming pass over A); other cases are batched matmuls on the device.
    """
    if A.dim() == 2:
        A = A.unsqueeze(0)
    if B.dim() < 2:
        raise ValueError("B must have at least 2 dimensions.")
    out_dev = A.device
    dev, (Ad, Bd) = _to_device(A, B)
    native = (
        B.dim() == 2 and B.shape[0] <= _ops.MAX_FILTERS and A.dtype == torch.float32 and A.dim() == 3
        and A.shape[-1] == A.shape[-2] == B.shape[-1]
    )
    with torch.cuda.device(dev):
        if native:
            C, _ = _ops.Project.apply(Bd, Ad.contiguous(), None)
        else:
            C = torch.einsum("...ij,njk,...kl->n...il", Bd, Ad, Bd.transpose(-2, -1))
    squeeze_dim = (0) if B.dim() == 2 else (0, 1)
    return torch.squeeze(C, dim=squeeze_dim).to(out_dev)


def _eig_native(M):
    """Eigen-decomposition of SPD matrices (..., m, m) with the per-class Jacobi kernel.
    Returns (V, lam, logM) on the compute device; eigenvalues are NOT sorted."""
    dev = _lib.compute_device(M)
    Md = _ops.f32c(M, dev).reshape(-1, M.shape[-2], M.shape[-1])
    m = Md.shape[-1]
    with torch.cuda.device(dev):
        W, _ = _ops.class_factor_raw(Md, _ops.DIST_LE)
    V = W[:, : m * m].reshape(-1, m, m)
    lam = W[:, m * m : m * m + m]
    logM = W[:, m * m + 2 * m :].reshape(-1, m, m)
    return V, lam, logM


def generalized_eigenvalues(A, B):
    """
    Generalized eigenvalues of the SPD pairs (A_a, B_b), descending (reference linalg.py:48-70).
    Shape (n_batch_A, n_batch_B, n_dim), size-1 batch dimensions squeezed. Forward only.
    """
    a3 = A.unsqueeze(0) if A.dim() == 2 else A
    b3 = B.unsqueeze(0) if B.dim() == 2 else B
    dev = _lib.compute_device(A, B)
    n_a, m, _ = a3.shape
    n_b = b3.shape[0]
    with torch.cuda.device(dev):
        Wa, _ = _ops.class_factor_raw(_ops.f32c(a3, dev), _ops.DIST_AI)
        Wb, _ = _ops.class_factor_raw(_ops.f32c(b3, dev), _ops.DIST_AI)
        lam = torch.empty(n_a, n_b, m, dtype=torch.float32, device=dev)
        _ops.pair_raw(Wa, Wb, n_a, n_b, m, _ops.DIST_AI, False, eig_out=lam)
    # squeeze rules of conjugate_matrix (linalg.py:44-45): a 2-D B drops its batch dim; any
    # remaining leading batch dim of size 1 is squeezed
    if B.dim() == 2:
        out = torch.squeeze(lam.squeeze(1), dim=0)
    else:
        out = torch.squeeze(lam, dim=(0, 1))
    return out.to(device=A.device, dtype=A.dtype)


def generalized_eigenvectors(A, B):
    """
    Generalized eigenvectors / eigenvalues of (A, B), descending (reference linalg.py:73-118).
    Not on a model path; composed from device-side torch ops.
    """
    dev = _lib.compute_device(A, B)
    a3 = (A.unsqueeze(0) if A.dim() == 2 else A).to(dev)
    b3 = (B.unsqueeze(0) if B.dim() == 2 else B).to(dev)
    W = spd_inv_sqrt(b3).to(dev)
    conj = W[None] @ a3[:, None] @ W.transpose(-2, -1)[None]
    vals, vecs = torch.linalg.eigh(conj)
    vals, vecs = vals.flip(-1), vecs.flip(-1)
    vecs = torch.einsum("bij,abjk->abik", W.transpose(-2, -1), vecs)
    vecs = vecs / torch.linalg.norm(vecs, dim=-2, keepdim=True)
    return torch.squeeze(vecs, dim=(0, 1)).to(A.device), torch.squeeze(vals, dim=(0, 1)).to(A.device)


def spd_sqrt(M):
    """Symmetric square root of SPD matrices (reference linalg.py:121-141)."""
    V, lam, _ = _eig_native(M)
    out = (V * torch.sqrt(lam).unsqueeze(-2)) @ V.transpose(-2, -1)
    return out.reshape(M.shape).to(device=M.device, dtype=M.dtype)


def spd_inv_sqrt(M):
    """
    Whitening matrices diag(lambda^-1/2) V^T of SPD matrices (reference linalg.py:144-162): for
    W = spd_inv_sqrt(M), W M W^T = I. (Rows are ordered by the Jacobi solver, not by eigenvalue.)
    """
    V, lam, _ = _eig_native(M)
    out = (V * torch.rsqrt(lam).unsqueeze(-2)).transpose(-2, -1)
    return out.reshape(M.shape).to(device=M.device, dtype=M.dtype)


def spd_log(M):
    """Matrix logarithm of SPD matrices (reference linalg.py:165-183), native Jacobi kernel."""
    _, _, logM = _eig_native(M)
    return logM.reshape(M.shape).to(device=M.device, dtype=M.dtype)
