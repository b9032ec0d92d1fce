"""Utility functions for matrix algebra -- B200-native drop-in for `sqfa.linalg`.

Public names and shape conventions follow /root/reference/src/sqfa/linalg.py. The functions on
the SQFA hot path (`conjugate_matrix` with a filter matrix, `generalized_eigenvalues`,
`spd_log`, `spd_inv_sqrt`) run in the sm_100a kernels; `spd_sqrt` and `generalized_eigenvectors`
are not used by any model path and are composed from the native
eigendecomposition with a few device-side torch ops. Every function is differentiable like its
reference counterpart (user `distance_fun`s are built from them, docs/source/tutorials/distances.md
of the reference), follows the dtype of its input and accepts matrices of any size.
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


_SYMMETRY_CACHE = {}


def _is_symmetric(A):
    """A (n, d, d) equals its transpose (exactly). One device pass, remembered per tensor version so a
    fitting loop that conjugates the same statistics every evaluation checks them once."""
    key = (A.data_ptr(), tuple(A.shape), A._version, A.device)
    hit = _SYMMETRY_CACHE.get(key)
    if hit is None:
        if len(_SYMMETRY_CACHE) > 64:
            _SYMMETRY_CACHE.clear()
        hit = bool(torch.equal(A, A.transpose(-2, -1)))
        _SYMMETRY_CACHE[key] = hit
    return hit


def conjugate_matrix(A, B):
    """
    Conjugate matrix A by B, i.e. compute B A B^T (reference linalg.py:19-45).

    A: (n_batch_A, n_dim, n_dim) or (n_dim, n_dim); B: (n_out, n_dim) or (n_batch_B, n_out, n_dim).
    Returns (n_batch_A, n_batch_B, n_out, n_out) with size-1 batch dimensions squeezed.
    With a 2-D B (a filter matrix, n_out <= 32) and float32 A this is the native projection kernel
    (one streaming pass over A). Its backward w.r.t. B uses the saved product B A, which is the
    gradient only for symmetric A: when B requires a gradient, A is checked (once per tensor) and a
    non-symmetric A takes the batched-matmul path, like every other case.
    """
    if A.dim() == 2:
        A = A.unsqueeze(0)
    if B.dim() < 2:
        raise ValueError("B must have at least 2 dimensions.")
    out_dev = A.device
    dev, (Ad, Bd) = _to_device(A, B)
    native = (
        B.dim() == 2 and B.shape[0] <= _ops.MAX_FILTERS and A.dtype == torch.float32 and A.dim() == 3
        and B.dtype == torch.float32 and A.shape[-1] == A.shape[-2] == B.shape[-1]
    )
    with torch.cuda.device(dev):
        if native and (B.requires_grad or A.requires_grad) and torch.is_grad_enabled():
            native = _is_symmetric(Ad)
        if native:
            C, _ = _ops.Project.apply(Bd, Ad.contiguous(), None)
        else:
            C = torch.einsum("...ij,njk,...kl->n...il", Bd, Ad, Bd.transpose(-2, -1))
    squeeze_dim = (0) if B.dim() == 2 else (0, 1)
    return torch.squeeze(C, dim=squeeze_dim).to(out_dev)


class _SpdEigh(torch.autograd.Function):
    """Eigendecomposition M = V diag(lam) V^T of SPD matrices (n, m, m), float32, m <= 64, with the
    per-matrix one-sided Jacobi kernel (`sqfa_class_factor`). Eigenvalues come in the solver's order
    (not sorted). The backward is the adjoint of a symmetric eigendecomposition -- what autograd
    applies to the `torch.linalg.eigh` calls of the reference (linalg.py:137, 159, 179):
        gM = V [ diag(g_lam) + skew(V^T g_V) / (lam_j - lam_i) ] V^T      (symmetrised)."""

    @staticmethod
    def forward(ctx, M):
        W, _ = _ops.class_factor_raw(M.contiguous(), _ops.DIST_LE)
        m = M.shape[-1]
        V = W[:, : m * m].reshape(-1, m, m)
        lam = W[:, m * m : m * m + m]
        ctx.save_for_backward(lam, V)
        return lam.clone(), V.clone()

    @staticmethod
    def backward(ctx, g_lam, g_V):
        lam, V = ctx.saved_tensors
        inner = torch.zeros_like(V)
        if g_V is not None:
            K = V.transpose(-2, -1) @ g_V
            K = 0.5 * (K - K.transpose(-2, -1))
            gap = lam.unsqueeze(-2) - lam.unsqueeze(-1)  # gap[i, j] = lam_j - lam_i
            eye = torch.eye(lam.shape[-1], dtype=torch.bool, device=lam.device)
            inner = torch.where(eye, torch.zeros_like(K), K / gap.masked_fill(eye, 1.0))
        if g_lam is not None:
            inner = inner + torch.diag_embed(g_lam)
        gM = V @ inner @ V.transpose(-2, -1)
        return 0.5 * (gM + gM.transpose(-2, -1))


def _native_ok(M):
    return M.dtype == torch.float32 and M.shape[-1] <= _ops.MAX_M and M.shape[-1] == M.shape[-2]


def _spd_eigh(M):
    """(lam, V, dev) of SPD matrices (..., m, m), differentiable, computed on the CUDA device in the
    dtype of M. float32 matrices up to 64 x 64 use the native Jacobi kernel; larger matrices (e.g.
    whitening a 784 x 784 data covariance) and float64 inputs -- the reference follows the input dtype
    and handles any size -- use `torch.linalg.eigh` on the device."""
    dev = _lib.compute_device(M)
    Md = M.to(dev).reshape(-1, M.shape[-2], M.shape[-1])
    with torch.cuda.device(dev):
        if _native_ok(Md):
            lam, V = _SpdEigh.apply(Md)
        else:
            lam, V = torch.linalg.eigh(Md)
    return lam, V


def generalized_eigenvalues(A, B):
    """
    Generalized eigenvalues of the SPD pairs (A_a, B_b), descending (reference linalg.py:48-70).
    Shape (n_batch_A, n_batch_B, n_dim), size-1 batch dimensions squeezed. float32 inputs up to
    64 x 64 that need no gradient run in the pair kernel; otherwise the reference's composition
    (whiten, conjugate, eigenvalues) runs on the device in the input dtype, differentiable.
    """
    a3 = A.unsqueeze(0) if A.dim() == 2 else A
    b3 = B.unsqueeze(0) if B.dim() == 2 else B
    dev = _lib.compute_device(A, B)
    n_a, m, _ = a3.shape
    n_b = b3.shape[0]
    needs_grad = torch.is_grad_enabled() and (A.requires_grad or B.requires_grad)
    with torch.cuda.device(dev):
        if _native_ok(a3) and _native_ok(b3) and not needs_grad:
            Wa, _ = _ops.class_factor_raw(_ops.f32c(a3, dev), _ops.DIST_AI)
            Wb, _ = _ops.class_factor_raw(_ops.f32c(b3, dev), _ops.DIST_AI)
            lam = torch.empty(n_a, n_b, m, dtype=torch.float32, device=dev)
            _ops.pair_raw(Wa, Wb, n_a, n_b, m, _ops.DIST_AI, False, eig_out=lam)
        else:
            W = spd_inv_sqrt(b3.to(dev))
            conj = W[None] @ a3.to(dev)[:, None] @ W.transpose(-2, -1)[None]
            vals, _ = _spd_eigh(conj)
            lam = torch.sort(vals.reshape(n_a, n_b, m), dim=-1, descending=True).values
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
    Not on a model path; composed from the eigendecomposition above with device-side torch ops.
    """
    dev = _lib.compute_device(A, B)
    a3 = (A.unsqueeze(0) if A.dim() == 2 else A).to(dev)
    b3 = (B.unsqueeze(0) if B.dim() == 2 else B).to(dev)
    with torch.cuda.device(dev):
        W = spd_inv_sqrt(b3)
        conj = W[None] @ a3[:, None] @ W.transpose(-2, -1)[None]
        n_a, n_b, m, _ = conj.shape
        vals, vecs = _spd_eigh(conj)
        order = torch.argsort(vals, dim=-1, descending=True)
        vals = torch.gather(vals, -1, order).reshape(n_a, n_b, m)
        vecs = torch.gather(vecs, -1, order.unsqueeze(-2).expand_as(vecs)).reshape(n_a, n_b, m, m)
        vecs = torch.einsum("bij,abjk->abik", W.transpose(-2, -1), vecs)
        vecs = vecs / torch.linalg.norm(vecs, dim=-2, keepdim=True)
    return torch.squeeze(vecs, dim=(0, 1)).to(A.device), torch.squeeze(vals, dim=(0, 1)).to(A.device)


def spd_sqrt(M):
    """Symmetric square root of SPD matrices (reference linalg.py:121-141); differentiable."""
    lam, V = _spd_eigh(M)
    out = (V * torch.sqrt(lam).unsqueeze(-2)) @ V.transpose(-2, -1)
    return out.reshape(M.shape).to(device=M.device, dtype=M.dtype)


def spd_inv_sqrt(M):
    """
    Whitening matrices diag(lambda^-1/2) V^T of SPD matrices (reference linalg.py:144-162): for
    W = spd_inv_sqrt(M), W M W^T = I; differentiable. (With the native solver the rows are in the
    Jacobi solver's order, not sorted by eigenvalue.)
    """
    lam, V = _spd_eigh(M)
    out = (V * torch.rsqrt(lam).unsqueeze(-2)).transpose(-2, -1)
    return out.reshape(M.shape).to(device=M.device, dtype=M.dtype)


def spd_log(M):
    """Matrix logarithm of SPD matrices (reference linalg.py:165-183); differentiable."""
    if _native_ok(M) and not (torch.is_grad_enabled() and M.requires_grad):
        dev = _lib.compute_device(M)
        with torch.cuda.device(dev):  # the Jacobi kernel forms V log(lam) V^T itself
            W, _ = _ops.class_factor_raw(_ops.f32c(M, dev).reshape(-1, M.shape[-2], M.shape[-1]), _ops.DIST_LE)
        m = M.shape[-1]
        return W[:, m * m + 2 * m :].reshape(M.shape).to(device=M.device, dtype=M.dtype)
    lam, V = _spd_eigh(M)
    out = (V * torch.log(lam).unsqueeze(-2)) @ V.transpose(-2, -1)
    return out.reshape(M.shape).to(device=M.device, dtype=M.dtype)
