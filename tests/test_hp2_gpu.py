"""GPU parity of hot path 2 (projection, pairwise SPD distances, fused loss + analytic backward, fit)
against the CPU oracle, through the Python host API that sits on the C ABI."""

import os

import pytest
import torch

from conftest import make_class_data, rel_err
from oracle import sqfa_oracle as O

pytestmark = pytest.mark.gpu

DIST_TOL = 1e-4   # north_star: pairwise distances and loss within 1e-4 relative
GRAD_TOL = 2e-3   # gradient of the loss w.r.t. the raw filters, relative Frobenius vs fp64 oracle
ANGLE_TOL = 1e-3  # north_star: learned filters within 1e-3 in subspace angle


def sample_spd(n, m, seed=0, dtype=torch.float64):
    """Random SPD matrices like the reference fixture tests/make_examples.py:15-20."""
    g = torch.Generator().manual_seed(seed)
    ev = 2 * torch.rand(n, m, generator=g, dtype=dtype) ** 2 + 0.01
    low = torch.tril(torch.randn(n, m, m, generator=g, dtype=dtype), diagonal=-1)
    Q = torch.matrix_exp(low - low.transpose(1, 2))
    return torch.einsum("ijk,ik,ilk->ijl", Q, ev, Q)


def stats_for(n, d, c, seed=0):
    X, y = make_class_data(n, d, c, seed=seed)
    X = X / (X.std() * d**0.5)
    return O.class_statistics(X.double(), y)


def to_f32_cuda(stats):
    return {k: v.float().cuda() for k, v in stats.items()}


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("c,d,k", [(1, 16, 2), (4, 33, 3), (10, 784, 4), (19, 104, 8), (3, 1027, 16), (2, 300, 32)])
def test_conjugate_matrix_matches_oracle(c, d, k):
    from sqfa_b200.linalg import conjugate_matrix

    S = sample_spd(c, d, seed=d) if d <= 128 else stats_for(40 * c + 200, d, c, seed=d)["second_moments"]
    F = torch.randn(k, d, generator=torch.Generator().manual_seed(k), dtype=torch.float64)
    ref = O.conjugate_matrix(S, F)
    got = conjugate_matrix(S.float().cuda(), F.float().cuda())
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 1e-5
    with pytest.raises(ValueError):
        conjugate_matrix(S.float().cuda(), F[0].float().cuda())


@pytest.mark.parametrize("n,m", [(1, 2), (4, 4), (8, 6), (12, 9), (20, 17), (6, 32), (5, 33), (3, 64)])
def test_distances_match_oracle(n, m):
    from sqfa_b200 import distances as Dn

    A = sample_spd(n, m, seed=m)
    B = sample_spd(max(n - 1, 1), m, seed=m + 100)
    Ac, Bc = A.float().cuda(), B.float().cuda()
    for name in ("affine_invariant_sq", "affine_invariant", "log_euclidean_sq", "log_euclidean"):
        ref_self = getattr(O, name)(A, A)
        got_self = getattr(Dn, name)(Ac, Ac)
        assert got_self.shape == ref_self.shape, name
        ref_cross = getattr(O, name)(A, B)
        got_cross = getattr(Dn, name)(Ac, Bc)
        assert got_cross.shape == ref_cross.shape, name
        if n > 1:
            i, j = torch.tril_indices(n, n, -1)
            assert rel_err(got_self[i, j], ref_self[i, j]) < DIST_TOL, name
            assert torch.equal(got_self, got_self.T)
        assert rel_err(got_cross, ref_cross) < DIST_TOL, name
    # Fisher-Rao lower bound on Gaussians
    mu = torch.randn(n, m, generator=torch.Generator().manual_seed(7), dtype=torch.float64)
    if m + 1 <= 64:
        sd = {"means": mu, "covariances": A}
        sc = {"means": mu.float().cuda(), "covariances": Ac}
        for name in ("fisher_rao_lower_bound_sq", "fisher_rao_lower_bound"):
            ref = getattr(O, name)(sd, sd)
            got = getattr(Dn, name)(sc, sc)
            assert got.shape == ref.shape
            if n > 1:
                i, j = torch.tril_indices(n, n, -1)
                assert rel_err(got[i, j], ref[i, j]) < DIST_TOL, name


@pytest.mark.parametrize("n", [1, 4, 8])
@pytest.mark.parametrize("m", [2, 4, 6])
def test_distance_properties(n, m):
    """Property tests of the reference: tests/test_distances.py:40-124."""
    from sqfa_b200 import distances as Dn

    A = sample_spd(n, m, seed=10 * n + m).float().cuda()
    ai = Dn.affine_invariant_sq(A, A)
    le = Dn.log_euclidean_sq(A, A)
    assert ai.shape == ((n, n) if n != 1 else ())
    assert le.shape == ((n, n) if n != 1 else ())
    if n > 1:
        assert torch.allclose(ai, ai.T, atol=1e-5) and torch.allclose(le, le.T, atol=1e-5)
        assert torch.allclose(torch.diagonal(ai), torch.zeros(n, device="cuda"), atol=1e-5)
    Ainv = torch.linalg.inv(A.double()).float()
    assert torch.allclose(ai, Dn.affine_invariant_sq(Ainv, Ainv), rtol=2e-4, atol=1e-4)
    assert torch.allclose(le, Dn.log_euclidean_sq(Ainv, Ainv), rtol=2e-4, atol=1e-4)
    eye = torch.eye(m, device="cuda")
    assert torch.allclose(Dn.affine_invariant_sq(A, eye), Dn.log_euclidean_sq(A, eye), rtol=2e-4, atol=1e-4)


@pytest.mark.parametrize("na,nb,m", [(1, 1, 2), (4, 1, 4), (4, 8, 6), (8, 4, 17)])
def test_generalized_eigenvalues_and_spd_functions(na, nb, m):
    """vs the oracle (itself pinned to scipy by the reference's tests/test_linalg.py:182-245)."""
    from sqfa_b200 import linalg as Ln

    A = sample_spd(na, m, seed=1)
    B = sample_spd(nb, m, seed=2)
    ref = O.generalized_eigenvalues(A, B)
    got = Ln.generalized_eigenvalues(A.float().cuda(), B.float().cuda())
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 1e-4
    assert rel_err(Ln.spd_log(A.float().cuda()), O.spd_log(A)) < 1e-4
    W = Ln.spd_inv_sqrt(B.float().cuda()).double().cpu()
    eye = torch.eye(m, dtype=torch.float64).expand(nb, m, m)
    assert torch.allclose(W @ B @ W.transpose(-2, -1), eye, atol=1e-4)
    R = Ln.spd_sqrt(A.float().cuda()).double().cpu()
    assert rel_err(R @ R, A) < 1e-4


# ------------------------------------------------------------------------------------------------
CASES = [
    # kind, n, d, c, k, distance name
    ("second_moments", 3000, 64, 10, 4, None),
    ("full", 3000, 64, 10, 4, None),
    ("full", 4000, 104, 19, 8, None),
    ("second_moments", 3000, 48, 6, 8, "log_euclidean"),
    ("full", 3000, 40, 5, 32, None),      # m = 33: one problem per warp, 17 lanes
    ("second_moments", 2000, 784, 10, 4, None),
    ("full", 10000, 32, 200, 4, None),    # 19 900 pairs, m = 5: 2 x 2 pair tiles, two of the ten problems a warp could hold
    ("second_moments", 10000, 24, 190, 6, "log_euclidean"),
]


def _models(kind, d, k, F0, distance=None, noise=0.01):
    from sqfa_b200 import distances as Dn
    from sqfa_b200.model import SQFA, SecondMomentsSQFA

    cls = SecondMomentsSQFA if kind == "second_moments" else SQFA
    dfun = getattr(Dn, distance) if distance else None
    return cls(n_dim=d, feature_noise=noise, n_filters=k, filters=F0.clone(), distance_fun=dfun)


@pytest.mark.parametrize("kind,n,d,c,k,distance", CASES)
def test_fused_loss_and_gradient_match_oracle(kind, n, d, c, k, distance):
    stats = stats_for(n, d, c, seed=k)
    F0 = torch.randn(k, d, generator=torch.Generator().manual_seed(3))
    odist = getattr(O, distance) if distance else None
    loss64, grad64, d64 = O.loss_and_grad(kind, stats, F0.double(), noise=0.01, distance=odist)

    model = _models(kind, d, k, F0, distance).cuda()
    sc = to_f32_cuda(stats)
    # fused native closure
    plan = model._fused_loss_plan(sc)
    assert plan is not None
    out = plan()
    out[0].backward()
    g = model.parametrizations.filters.original.grad
    assert float(out[1]) == 0.0
    assert abs(float(out[0]) - float(loss64)) <= DIST_TOL * abs(float(loss64))
    assert rel_err(g, grad64) < GRAD_TOL
    # generic path: get_class_distances + autograd through the native ops
    model.zero_grad()
    dmat = model.get_class_distances(sc, regularized=True)
    i, j = torch.tril_indices(c, c, -1)
    assert rel_err(dmat[i, j], d64[i, j]) < DIST_TOL
    (-dmat[i, j].mean()).backward()
    g2 = model.parametrizations.filters.original.grad
    assert rel_err(g2, grad64) < GRAD_TOL


@pytest.mark.parametrize("kind,d,c,k,world", [("full", 96, 37, 6, 4), ("second_moments", 200, 12, 4, 3),
                                              ("full", 64, 5, 16, 8)])
def test_class_and_pair_sharded_closure_phases_on_one_device(kind, d, c, k, world):
    """sqfa_fused_loss_sharded (multi-GPU fits: classes of the projection AND pairs sharded over ranks) with
    `world` emulated ranks on one device: every rank runs the three phases on its own workspace, the two
    exchange spans and dF are summed over the ranks the way the all-reduces would. Loss and dLoss/dF must
    equal the unsharded sqfa_fused_loss and the oracle (also with more ranks than classes)."""
    import ctypes

    from sqfa_b200 import _lib, _ops
    from sqfa_b200._stats_driver import class_share

    lib = _lib.load()
    stats = stats_for(60 * c, d, c, seed=c)
    sc = to_f32_cuda(stats)
    F = torch.randn(k, d, generator=torch.Generator().manual_seed(5)).cuda()
    F = (F / F.norm(dim=1, keepdim=True)).contiguous()
    model = _models(kind, d, k, F.cpu()).cuda()
    S, M, dist = model._fused_inputs(sc)
    noise = model._noise_scalar()
    ref = _ops.fused_loss_raw(F, S, M, noise, dist)
    P = c * (c - 1) // 2
    ranks = []
    for r in range(world):
        p0, p1 = _ops.shard_pairs(P, c, r, world)
        c0, c1 = class_share(c, r, world)
        ws = torch.empty(lib.sqfa_fused_loss_workspace_bytes(c, d, k, dist, p0, p1), dtype=torch.uint8, device="cuda")
        spans = []
        for which in (0, 1):
            off, nb = ctypes.c_size_t(0), ctypes.c_size_t(0)
            _lib.check(lib.sqfa_fused_loss_exchange_span(c, d, k, dist, p0, p1, which, ctypes.byref(off),
                                                         ctypes.byref(nb)), "span")
            spans.append(ws[off.value:off.value + nb.value].view(torch.float32))
        ranks.append({"p": (p0, p1), "c": (c0, c1), "ws": ws, "spans": spans,
                      "dF": torch.empty(k, d, device="cuda")})

    def phase(i):
        for q in ranks:
            _lib.check(lib.sqfa_fused_loss_sharded(i, _lib.ptr(S), _lib.ptr(M), _lib.ptr(F), c, d, k, float(noise), dist,
                                                   q["c"][0], q["c"][1], q["p"][0], q["p"][1], _lib.ptr(q["dF"]),
                                                   _lib.ptr(q["ws"]), q["ws"].numel(), _lib.stream_ptr()), "sharded")

    def all_reduce(which):
        total = torch.stack([q["spans"][which] for q in ranks]).sum(0)
        for q in ranks:
            q["spans"][which].copy_(total)

    phase(0)
    all_reduce(0)
    phase(1)
    all_reduce(1)
    phase(2)
    dF = torch.stack([q["dF"] for q in ranks]).sum(0)
    loss, bad = ranks[0]["spans"][1][-64:-62].tolist()
    assert bad == 0.0
    assert abs(loss - float(ref[0])) <= 1e-6 * abs(float(ref[0]))
    assert rel_err(dF, ref[4:].view(k, d)) < 1e-5
    loss64, _, _ = O.loss_and_grad(kind, stats, F.double().cpu(), noise=0.01)
    assert abs(loss - float(loss64)) <= DIST_TOL * abs(float(loss64))


def test_fit_matches_oracle_trajectory():
    """Filters after ONE epoch (well conditioned) and the converged loss."""
    n, d, c, k = 4000, 32, 6, 4
    stats = stats_for(n, d, c, seed=11)
    stats32 = {kk: v.float() for kk, v in stats.items()}
    F0 = O.pca_from_scatter(stats32["second_moments"], k)
    stats64 = {kk: v.double() for kk, v in stats.items()}
    for kind in ("second_moments", "full"):
        # fixed iteration count: one LBFGS epoch of 6 inner iterations. (A full 20-iteration epoch
        # amplifies fp32 rounding differences beyond 1e-3 in the REFERENCE itself: its own
        # fp32 and fp64 runs differ by more)
        Fo, losses_o, _ = O.fit_lbfgs(kind, stats64, F0.double(), noise=0.01, max_epochs=1, max_iter=6)
        model = _models(kind, d, k, F0)
        loss, _ = model.fit(data_statistics=stats32, max_epochs=1, show_progress=False, return_loss=True, max_iter=6)
        assert model.filters.device.type == "cpu"  # the model returns to where it lived
        assert O.subspace_angle(model.filters.detach(), Fo) < ANGLE_TOL
        assert abs(float(loss[0]) - float(losses_o[0])) < DIST_TOL * abs(float(losses_o[0]))
        # converged
        Fo, losses_o, _ = O.fit_lbfgs(kind, stats64, F0.double(), noise=0.01, max_epochs=100)
        model = _models(kind, d, k, F0)
        loss, _ = model.fit(data_statistics=stats32, max_epochs=100, show_progress=False, return_loss=True)
        assert abs(float(loss[-1]) - float(losses_o[-1])) < 1e-4 * abs(float(losses_o[-1]))
        print(kind, "converged angle", O.subspace_angle(model.filters.detach(), Fo))


def test_fit_from_points_pairwise_and_transform():
    from sqfa_b200.model import SQFA

    X, y = make_class_data(3000, 24, 5, seed=2)
    X = X / (X.std() * 24**0.5)
    model = SQFA(n_dim=24, feature_noise=0.01, n_filters=4)
    model.fit_pca(X)
    assert O.subspace_angle(model.filters.detach(), O.pca(X.double(), 4).float()) < 1e-3
    loss, t = model.fit(X, y, max_epochs=5, show_progress=False, return_loss=True, pairwise=True)
    assert torch.isfinite(loss).all() and loss.numel() > 0
    Z = model.transform(X)
    assert Z.shape == (3000, 4)
    assert rel_err(Z, X.double() @ model.filters.detach().double().T) < 1e-5
    cov = model.transform_scatters(O.class_statistics(X, y)["covariances"])
    assert cov.shape == (5, 4, 4)


def test_error_types_match_reference():
    """Error matrix of the reference: tests/test_training.py:90-101, 201-259."""
    from sqfa_b200.model import SQFA, SecondMomentsSQFA

    stats = {k: v.float() for k, v in stats_for(500, 8, 3).items()}
    with pytest.raises(ValueError):
        SQFA(n_dim=4, n_filters=6)
    with pytest.raises(ValueError):
        SQFA(n_dim=8, n_filters=2).fit()
    with pytest.raises(ValueError):
        SQFA(n_dim=8, n_filters=2).fit_pca()
    with pytest.raises(TypeError):
        SQFA(n_dim=8, n_filters=2).fit(data_statistics=stats["second_moments"], show_progress=False)
    with pytest.raises(TypeError):
        SecondMomentsSQFA(n_dim=8, n_filters=2).fit(data_statistics=[1, 2, 3], show_progress=False)
    with pytest.raises(ValueError):
        SecondMomentsSQFA(n_dim=8, n_filters=2).fit(data_statistics={"means": stats["means"]}, show_progress=False)
    with pytest.raises(ValueError):
        SecondMomentsSQFA(n_dim=8, n_filters=3).fit(
            data_statistics=stats, pairwise=True, show_progress=False, max_epochs=2
        )
    with pytest.raises(TypeError):
        SQFA(n_dim=8, n_filters=2).get_class_distances(stats["second_moments"])


@pytest.mark.parametrize("tc", ["default", "0", "1", "tma"])
@pytest.mark.parametrize("C,D,k", [(3, 1027, 5), (7, 40, 3), (2, 3072, 8), (130, 96, 16), (5, 520, 32), (3, 1024, 24),
                                   (40, 512, 16), (2, 132, 9), (6, 300, 31)])
def test_project_forward_shapes(C, D, k, tc, monkeypatch):
    """T = F S, Psi = T F^T, mu' = F m for row splits, partial column strips, D % 4 != 0, every KT -- through
    the SIMT pass (SQFA_PROJECT_TC=0), the tcgen05 pass forced for every 8 < k <= 32 (=1: ragged row tiles and
    column stages, k not a multiple of 16), its TMA-staged variant, and the default choice (tensor cores, register
    staging, for k > 16)."""
    from sqfa_b200 import _ops

    monkeypatch.delenv("SQFA_PROJECT_TC_BULK", raising=False)
    if tc == "default":
        monkeypatch.delenv("SQFA_PROJECT_TC", raising=False)
    elif tc == "tma":  # the tcgen05 pass with the raw tile staged by TMA (cp.async.bulk.tensor), for every 8 < k <= 32
        monkeypatch.setenv("SQFA_PROJECT_TC", "1")
        monkeypatch.setenv("SQFA_PROJECT_TC_BULK", "1")
    else:
        monkeypatch.setenv("SQFA_PROJECT_TC", tc)

    g = torch.Generator().manual_seed(C * D + k)
    A = torch.randn(C, D, D, generator=g)
    S = (A + A.transpose(1, 2)) / 2
    M = torch.randn(C, D, generator=g)
    F = torch.randn(k, D, generator=g) / D**0.5
    T, Psi, Mu = _ops.project_fwd_raw(S.cuda(), M.cuda(), F.cuda())
    T64 = torch.einsum("fi,cij->cfj", F.double(), S.double())
    assert rel_err(T, T64) < 1e-5
    assert rel_err(Psi, torch.einsum("cfj,gj->cfg", T64, F.double())) < 1e-5
    assert rel_err(Mu, M.double() @ F.double().T) < 1e-5


# ------------------------------------------------------------------------------------------------
# autograd through the public linalg functions (user distance_funs are built from them)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["spd_sqrt", "spd_log", "spd_inv_sqrt"])
@pytest.mark.parametrize("n,m", [(3, 4), (5, 9)])
def test_linalg_functions_are_differentiable(name, n, m):
    """d/dM of a scalar function of spd_sqrt / spd_log / spd_inv_sqrt against fp64 autograd through
    torch.linalg.eigh (what the reference differentiates through, linalg.py:121-183)."""
    from sqfa_b200 import linalg as Ln

    M = sample_spd(n, m, seed=7 * m + n)
    Wt = torch.randn(n, m, m, generator=torch.Generator().manual_seed(1), dtype=torch.float64)

    def scalar(fun_out):
        if name == "spd_inv_sqrt":  # rows of the whitening matrix come in the solver's order: use W^T W = M^-1
            fun_out = fun_out.transpose(-2, -1) @ fun_out
        return (fun_out * Wt.to(fun_out)).sum()

    M64 = M.clone().requires_grad_(True)
    lam, V = torch.linalg.eigh(M64)
    ref_out = {"spd_sqrt": (V * lam.sqrt().unsqueeze(-2)) @ V.transpose(-2, -1),
               "spd_log": (V * lam.log().unsqueeze(-2)) @ V.transpose(-2, -1),
               "spd_inv_sqrt": (V * lam.rsqrt().unsqueeze(-2)).transpose(-2, -1)}[name]
    scalar(ref_out).backward()
    g_ref = 0.5 * (M64.grad + M64.grad.transpose(-2, -1))

    M32 = M.float().cuda().requires_grad_(True)
    out = getattr(Ln, name)(M32)
    assert out.requires_grad and out.grad_fn is not None
    scalar(out).backward()
    g_got = 0.5 * (M32.grad + M32.grad.transpose(-2, -1))
    assert rel_err(g_got, g_ref) < 1e-3
    # float64 input: computed in float64 (the reference follows the input dtype)
    M64d = M.cuda().requires_grad_(True)
    out64 = getattr(Ln, name)(M64d)
    assert out64.dtype == torch.float64
    scalar(out64).backward()
    assert rel_err(0.5 * (M64d.grad + M64d.grad.transpose(-2, -1)), g_ref) < 1e-9


def _bw_distance_sq(A, B, linalg):
    """The Bures-Wasserstein distance_fun of the reference's tutorial (docs/source/tutorials/distances.md)."""
    tr_A = torch.einsum("ijj->i", A)
    tr_B = torch.einsum("ijj->i", B)
    A_sqrt = linalg.spd_sqrt(A)
    C = linalg.conjugate_matrix(B, A_sqrt)
    tr_C = torch.sum(torch.sqrt(torch.linalg.eigvalsh(C)), dim=-1)
    return tr_A[None, :] + tr_B[:, None] - 2 * tr_C


def test_user_distance_fun_trains_with_correct_gradients():
    """A fit with the tutorial's Bures-Wasserstein distance: the gradient of the first closure
    evaluation (through spd_sqrt) and the losses of a short fit against the oracle in fp64."""
    from sqfa_b200 import linalg as Ln
    from sqfa_b200.model import SecondMomentsSQFA

    class OracleLinalg:  # the reference's formulas for the two functions the tutorial uses
        conjugate_matrix = staticmethod(O.conjugate_matrix)

        @staticmethod
        def spd_sqrt(M):
            lam, V = torch.linalg.eigh(M)
            return torch.einsum("...ij,...j,...kj->...ik", V, torch.sqrt(lam), V)

    def bw(A, B):
        return torch.sqrt(torch.abs(_bw_distance_sq(A, B, Ln)) + 1e-6)

    def bw_oracle(A, B):
        return torch.sqrt(torch.abs(_bw_distance_sq(A, B, OracleLinalg)) + 1e-6)

    n, d, c, k = 3000, 20, 5, 3
    stats = stats_for(n, d, c, seed=21)
    F0 = torch.randn(k, d, generator=torch.Generator().manual_seed(2))
    loss64, grad64, _ = O.loss_and_grad("second_moments", stats, F0.double(), noise=0.01, distance=bw_oracle)
    model = SecondMomentsSQFA(n_dim=d, feature_noise=0.01, n_filters=k, filters=F0.clone(), distance_fun=bw).cuda()
    sc = to_f32_cuda(stats)
    dmat = model.get_class_distances(sc, regularized=True)
    i, j = torch.tril_indices(c, c, -1)
    loss = -dmat[i, j].mean()
    loss.backward()
    assert abs(float(loss) - float(loss64)) < DIST_TOL * abs(float(loss64))
    assert rel_err(model.parametrizations.filters.original.grad, grad64) < GRAD_TOL
    _, losses_o, _ = O.fit_lbfgs("second_moments", stats, F0.double(), noise=0.01, distance=bw_oracle, max_epochs=2,
                                 max_iter=5)
    model = SecondMomentsSQFA(n_dim=d, feature_noise=0.01, n_filters=k, filters=F0.clone(), distance_fun=bw)
    losses, _ = model.fit(data_statistics={kk: v.float() for kk, v in stats.items()}, max_epochs=2,
                          show_progress=False, return_loss=True, max_iter=5)
    assert torch.allclose(losses.double(), losses_o.double(), rtol=1e-3)


def test_large_and_double_inputs_take_the_composed_path():
    """Matrices above 64 x 64 and float64 inputs: same semantics as the reference (any size, output dtype
    follows the input), computed with device-side library calls instead of the pair kernels."""
    from sqfa_b200 import distances as Dn
    from sqfa_b200 import linalg as Ln

    A = sample_spd(3, 80, seed=1)
    B = sample_spd(2, 80, seed=2)
    got = Dn.affine_invariant(A.float().cuda(), B.float().cuda())
    assert got.dtype == torch.float32 and rel_err(got, O.affine_invariant(A, B)) < 1e-3
    W = Ln.spd_inv_sqrt(A.float().cuda()).double().cpu()
    assert torch.allclose(W @ A @ W.transpose(-2, -1), torch.eye(80, dtype=torch.float64).expand(3, 80, 80), atol=1e-3)
    A6, B6 = sample_spd(4, 6, seed=3), sample_spd(3, 6, seed=4)
    for name in ("affine_invariant_sq", "affine_invariant", "log_euclidean_sq", "log_euclidean"):
        got = getattr(Dn, name)(A6.cuda(), B6.cuda())
        assert got.dtype == torch.float64
        assert rel_err(got, getattr(O, name)(A6, B6)) < 1e-10, name
    mu = torch.randn(4, 6, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    sd = {"means": mu, "covariances": A6}
    got = Dn.fisher_rao_lower_bound({k: v.cuda() for k, v in sd.items()}, {k: v.cuda() for k, v in sd.items()})
    assert got.dtype == torch.float64 and rel_err(got, O.fisher_rao_lower_bound(sd, sd)) < 1e-8
    lam = Ln.generalized_eigenvalues(A6.cuda(), B6.cuda())
    assert lam.dtype == torch.float64 and rel_err(lam, O.generalized_eigenvalues(A6, B6)) < 1e-10


# ------------------------------------------------------------------------------------------------
# the reference's other distance_fun plug-ins (distances.py:240-432): native warp-per-pair kernels
# ------------------------------------------------------------------------------------------------
PLUGINS = ("bhattacharyya", "mahalanobis_sq", "mahalanobis", "hellinger", "fisher_rao_same_cov")


def test_plugin_distances_golden():
    import numpy as np

    from sqfa_b200 import distances as Dn

    with np.load(os.path.join(os.path.dirname(__file__), "golden", "distances.npz")) as z:
        g = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
    sa = {"means": g["mu_a"].float().cuda(), "covariances": g["A"].float().cuda()}
    sb = {"means": g["mu_b"].float().cuda(), "covariances": g["B"].float().cuda()}
    for name, key in (("bhattacharyya", "bhatt"), ("hellinger", "hell"), ("fisher_rao_same_cov", "frsc"),
                      ("mahalanobis_sq", "maha_sq")):
        got = getattr(Dn, name)(sa, sb)
        assert got.shape == g[key + "_ab"].shape, name
        assert rel_err(got, g[key + "_ab"]) < 1e-4, name
    assert rel_err(Dn.mahalanobis(sa, sb), g["maha_ab"]) < 1e-4
    i, j = torch.tril_indices(5, 5, -1)
    assert rel_err(Dn.bhattacharyya(sa, sa)[i, j], g["bhatt_aa"][i, j]) < 1e-4
    assert rel_err(Dn.mahalanobis_sq(sa, sa)[i, j], g["maha_sq_aa"][i, j]) < 1e-4
    # float64 inputs: float64 results through the composed path
    sa64 = {"means": g["mu_a"].cuda(), "covariances": g["A"].cuda()}
    sb64 = {"means": g["mu_b"].cuda(), "covariances": g["B"].cuda()}
    got = Dn.bhattacharyya(sa64, sb64)
    assert got.dtype == torch.float64 and rel_err(got, g["bhatt_ab"]) < 1e-10


@pytest.mark.parametrize("name", PLUGINS)
@pytest.mark.parametrize("na,nb,k", [(1, 1, 2), (5, 3, 4), (7, 7, 9), (40, 3, 17), (3, 2, 40)])
def test_plugin_distances_match_oracle_with_gradients(name, na, nb, k):
    """values and d/d(means, covariances) against fp64 autograd through the reference's formulas
    (torch.linalg.inv / logdet), for A != B and for A is B."""
    from sqfa_b200 import distances as Dn

    g = torch.Generator().manual_seed(na * 100 + nb * 10 + k)
    A, B = sample_spd(na, k, seed=k), sample_spd(nb, k, seed=k + 1)
    mu_a = torch.randn(na, k, generator=g, dtype=torch.float64)
    mu_b = torch.randn(nb, k, generator=g, dtype=torch.float64)
    Wt = torch.randn(na, nb, generator=g, dtype=torch.float64)

    def run(fun, tensors, weights):
        leaves = [t.clone().requires_grad_(True) for t in tensors]
        sa = {"means": leaves[0], "covariances": leaves[1]}
        sb = {"means": leaves[2], "covariances": leaves[3]}
        D = fun(sa, sb)
        (D.reshape(na, nb) * weights.to(D)).sum().backward()
        return D.detach(), [t.grad for t in leaves]

    D_ref, g_ref = run(getattr(O, name), (mu_a, A, mu_b, B), Wt)
    D_got, g_got = run(getattr(Dn, name), tuple(t.float().cuda() for t in (mu_a, A, mu_b, B)), Wt)
    assert D_got.shape == D_ref.shape
    assert rel_err(D_got, D_ref) < 1e-4
    for a, b in zip(g_got, g_ref):
        b = 0.5 * (b + b.transpose(-2, -1)) if b.dim() == 3 else b
        a = 0.5 * (a + a.transpose(-2, -1)) if a.dim() == 3 else a
        assert rel_err(a, b) < 1e-3
    # (fisher_rao_same_cov: d acosh(1 + x / 4) / dx is infinite at the diagonal's x = 0, in the reference too)
    if na == nb and name != "fisher_rao_same_cov":  # the same statistics on both sides
        leaves = [mu_a.float().cuda().requires_grad_(True), A.float().cuda().requires_grad_(True)]
        s = {"means": leaves[0], "covariances": leaves[1]}
        D = getattr(Dn, name)(s, s)
        l64 = [mu_a.clone().requires_grad_(True), A.clone().requires_grad_(True)]
        s64 = {"means": l64[0], "covariances": l64[1]}
        D64 = getattr(O, name)(s64, s64)
        off = ~torch.eye(na, dtype=torch.bool)
        if na > 1:
            (D.reshape(na, na)[off.cuda()] * Wt[off].float().cuda()).sum().backward()
            (D64.reshape(na, na)[off] * Wt[off]).sum().backward()
            assert rel_err(leaves[0].grad, l64[0].grad) < 1e-3
            assert rel_err(0.5 * (leaves[1].grad + leaves[1].grad.transpose(1, 2)),
                           0.5 * (l64[1].grad + l64[1].grad.transpose(1, 2))) < 1e-3


def test_fit_with_plugin_distance():
    """SQFA.fit with distance_fun=bhattacharyya: losses of a short fit against the oracle in fp64."""
    from sqfa_b200 import distances as Dn
    from sqfa_b200.model import SQFA

    n, d, c, k = 3000, 20, 5, 3
    stats = stats_for(n, d, c, seed=31)
    F0 = torch.randn(k, d, generator=torch.Generator().manual_seed(4))
    loss64, grad64, _ = O.loss_and_grad("full", stats, F0.double(), noise=0.01, distance=O.bhattacharyya)
    model = SQFA(n_dim=d, feature_noise=0.01, n_filters=k, filters=F0.clone(), distance_fun=Dn.bhattacharyya).cuda()
    dmat = model.get_class_distances(to_f32_cuda(stats), regularized=True)
    i, j = torch.tril_indices(c, c, -1)
    loss = -dmat[i, j].mean()
    loss.backward()
    assert abs(float(loss) - float(loss64)) < DIST_TOL * abs(float(loss64))
    assert rel_err(model.parametrizations.filters.original.grad, grad64) < GRAD_TOL
    _, losses_o, _ = O.fit_lbfgs("full", stats, F0.double(), noise=0.01, distance=O.bhattacharyya, max_epochs=2, max_iter=5)
    model = SQFA(n_dim=d, feature_noise=0.01, n_filters=k, filters=F0.clone(), distance_fun=Dn.bhattacharyya)
    losses, _ = model.fit(data_statistics={kk: v.float() for kk, v in stats.items()}, max_epochs=2,
                          show_progress=False, return_loss=True, max_iter=5)
    assert torch.allclose(losses.double(), losses_o.double(), rtol=1e-3)
