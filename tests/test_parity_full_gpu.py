"""GPU: parity against the CPU oracle ON BASELINE.json's own configurations, at full size.

* c1-c4 `class_statistics`: every class's mean, covariance and second moment against the fp64 oracle
  (north_star: 1e-5 relative).
* the closure (loss and gradient) at full c1, c2, c3 and at the c5 shape (C=100, k=32 -> m=33).
* c4 (C=1000, m=17, 499 500 pairs): >= 5 000 pairs' distances and their part of dLoss/dE against the
  oracle, addressed through `pair_begin` / `pair_end`; plus one full oracle closure (`slow`, ~1 min).
* `transform` at the shapes its tuned kernel is claimed on (50 000 x 3072, k = 8 / 16 / 32, a ragged
  row count, strided and unaligned rows).
* `pca_from_scatter` / `fit_pca(data_statistics=)` against the reference's golden vector.
* the NaN / inf guard of the fitting loop (reference _optim.py:16-30) with the reference's messages.
"""

import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import sqfa_oracle as O

pytestmark = pytest.mark.gpu

CONFIGS = {
    # name: (N, D, C, k, model kind)
    "c1": (60000, 784, 10, 4, "second_moments"),
    "c2": (50000, 3072, 10, 8, "full"),
    "c3": (200000, 104, 19, 8, "full"),
    "c4": (1280000, 512, 1000, 16, "full"),
}
STAT_TOL = 1e-5   # north_star: class means and second moments within 1e-5 relative (fp32)
DIST_TOL = 1e-4   # north_star: pairwise distances and loss within 1e-4 relative
GRAD_TOL = 1e-3   # dLoss/dF against fp64 autograd through the reference's formulas (measured ~1e-6)


def synth(n, d, c, seed=0):
    """SURVEY.md 8(d) data, generated on the device (the CPU copy feeds the oracle)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    y = torch.randint(0, c, (n,), generator=g, device="cuda")
    basis = torch.randn(32, d, generator=g, device="cuda") / 32**0.5
    scales = 0.5 + torch.rand(c, generator=g, device="cuda")
    means = 0.2 * torch.randn(c, d, generator=g, device="cuda")
    x = (torch.randn(n, 32, generator=g, device="cuda") * scales[y][:, None]) @ basis
    x += 0.5 * torch.randn(n, d, generator=g, device="cuda")
    x += means[y]
    x /= x.std() * d**0.5
    return x.contiguous(), y


def per_class_rel_err(got, ref):
    """max over classes of |got_c - ref_c|_F / |ref_c|_F"""
    got, ref = got.double().cpu(), ref.double()
    c = ref.shape[0]
    num = (got - ref).reshape(c, -1).norm(dim=1)
    den = ref.reshape(c, -1).norm(dim=1).clamp_min(1e-300)
    return float((num / den).max())


_STATS_CACHE = {}


def full_size_statistics(cfg):
    """(device statistics dict, fp64 oracle statistics dict) of a BASELINE config, computed once."""
    if cfg not in _STATS_CACHE:
        from sqfa_b200.statistics import class_statistics

        n, d, c, _, _ = CONFIGS[cfg]
        X, y = synth(n, d, c, seed=len(cfg) + n % 7)
        got = class_statistics(X, y)
        ref = O.class_statistics(X.cpu().double(), y.cpu())
        del X, y
        torch.cuda.empty_cache()
        _STATS_CACHE.clear()  # one configuration resident at a time (c2 / c4 statistics are GBs)
        _STATS_CACHE[cfg] = (got, ref)
    return _STATS_CACHE[cfg]


# ------------------------------------------------------------------------------------------------
# HP1 at full size, every class
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4"])
def test_class_statistics_full_size_every_class(cfg):
    got, ref = full_size_statistics(cfg)
    for key in ("means", "covariances", "second_moments"):
        assert got[key].shape == ref[key].shape
        err = per_class_rel_err(got[key], ref[key])
        assert err < STAT_TOL, (cfg, key, err)


# ------------------------------------------------------------------------------------------------
# HP2 closure at full size
# ------------------------------------------------------------------------------------------------
def _closure_vs_oracle(kind, stats_dev, stats64, k, seed):
    from sqfa_b200.model import SQFA, SecondMomentsSQFA

    d = stats64["means"].shape[1]
    F0 = torch.randn(k, d, generator=torch.Generator().manual_seed(seed))
    loss64, grad64, _ = O.loss_and_grad(kind, stats64, F0.double(), noise=0.01)
    cls = SecondMomentsSQFA if kind == "second_moments" else SQFA
    model = cls(n_dim=d, feature_noise=0.01, n_filters=k, filters=F0.clone()).cuda()
    # the graph-free closure the fitting loop uses ...
    direct = model._fused_direct_plan(stats_dev)
    assert direct is not None
    loss_d, bad, _ = direct().tolist()
    g_direct = model.parametrizations.filters.original.grad.clone()
    # ... and the autograd one
    model.zero_grad()
    out = model._fused_loss_plan(stats_dev)()
    out[0].backward()
    g_auto = model.parametrizations.filters.original.grad
    assert bad == 0 and float(out[1]) == 0
    for val in (loss_d, float(out[0])):
        assert abs(val - float(loss64)) <= DIST_TOL * abs(float(loss64)), (val, float(loss64))
    e_direct, e_auto = rel_err(g_direct, grad64), rel_err(g_auto, grad64)
    print(f"closure {kind} C={stats64['means'].shape[0]} D={d} k={k}: loss {loss_d:.7f} (oracle {float(loss64):.7f}), "
          f"grad rel err direct {e_direct:.2e} autograd {e_auto:.2e}")
    assert e_direct < GRAD_TOL and e_auto < GRAD_TOL


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3"])
def test_closure_full_size_matches_oracle(cfg):
    _, _, _, k, kind = CONFIGS[cfg]
    got, ref = full_size_statistics(cfg)
    stats_dev = {kk: v for kk, v in got.items()}
    _closure_vs_oracle(kind, stats_dev, ref, k, seed=3)


def test_closure_c5_shape_matches_oracle():
    """BASELINE config 5's closure: C = 100, D = 1024, k = 32 -> Fisher-Rao on 33 x 33 embeddings
    (two Jacobi columns per lane), 4 950 pairs. Statistics of 300 000 synthetic rows."""
    from sqfa_b200.statistics import class_statistics

    X, y = synth(300000, 1024, 100, seed=55)
    got = class_statistics(X, y)
    ref = {kk: v.double().cpu() for kk, v in got.items()}  # the closure is what is under test here
    del X, y
    _closure_vs_oracle("full", got, ref, 32, seed=5)


def _c4_embedded(k=16, noise=0.01):
    """E (C, k+1, k+1) of the c4 statistics under random unit filters: float32 on the device through
    the native projection + embedding, and the fp64 oracle restatement (distances.py:141-174)."""
    from sqfa_b200 import _ops

    got, ref = full_size_statistics("c4")
    d = ref["means"].shape[1]
    F = torch.randn(k, d, generator=torch.Generator().manual_seed(16))
    F = F / F.norm(dim=1, keepdim=True)
    T, Psi, Mu = _ops.project_fwd_raw(got["covariances"], got["means"], F.cuda())
    E = _ops.embed_fwd_raw(Psi, Mu, noise, _ops.DIST_FR)
    F64 = F.double()
    cov64 = O.conjugate_matrix(ref["covariances"], F64) + noise * torch.eye(k, dtype=torch.float64)
    E64 = O.embed_gaussian({"means": ref["means"] @ F64.T, "covariances": cov64})
    assert rel_err(E, E64) < 1e-5
    return E, E64


def test_c4_pair_subsets_match_oracle():
    """>= 5 000 of c4's 499 500 pairs, as four slices of the linearised pair list: distances and the
    slices' contribution to dLoss/dE against the oracle (autograd through eigvalsh, fp64)."""
    from sqfa_b200 import _ops

    E, E64 = _c4_embedded()
    C, m, _ = E64.shape
    P = C * (C - 1) // 2
    W, flag = _ops.class_factor_raw(E, _ops.DIST_FR)
    assert int(flag) == 0
    starts = [0, 123_457, 311_111, P - 1500]
    weight = -1.0 / P
    dist_out = torch.zeros(C, C, device="cuda")
    gE = torch.zeros(C, m, m, device="cuda")
    loss = torch.zeros(2, device="cuda")
    for p0 in starts:
        _ops.pair_raw(W, W, C, C, m, _ops.DIST_FR, True, weight=weight, dist_out=dist_out, loss=loss, gEa=gE, gEb=gE,
                      pair_range=(p0, p0 + 1500))
    # oracle: the same pairs, row by row
    E64 = E64.clone().requires_grad_(True)
    total, n_pairs, worst = 0.0, 0, 0.0
    for p0 in starts:
        ps = torch.arange(p0, p0 + 1500)
        ii = torch.floor((1 + torch.sqrt(1 + 8 * ps.double())) / 2).long()
        ii = torch.where(ii * (ii - 1) // 2 > ps, ii - 1, ii)
        ii = torch.where((ii + 1) * ii // 2 <= ps, ii + 1, ii)
        jj = ps - ii * (ii - 1) // 2
        assert bool(((jj >= 0) & (jj < ii)).all())
        for i in ii.unique().tolist():
            cols = jj[ii == i]
            d_ref = torch.sqrt(O.affine_invariant_sq(E64[i : i + 1], E64[cols]).reshape(-1) / 2 + O.EPSILON)
            d_got = dist_out[i, cols.cuda()].double().cpu()
            worst = max(worst, float(((d_got - d_ref.detach()).abs() / d_ref.detach()).max()))
            assert torch.equal(dist_out[i, cols.cuda()], dist_out[cols.cuda(), i])  # mirrored
            total = total + d_ref.sum()
            n_pairs += cols.numel()
    assert n_pairs == 6000
    assert worst < DIST_TOL, worst
    (weight * total).backward()
    g_ref = 0.5 * (E64.grad + E64.grad.transpose(1, 2))
    g_got = gE.double().cpu()
    g_got = 0.5 * (g_got + g_got.transpose(1, 2))
    touched = g_ref.reshape(C, -1).norm(dim=1) > 0
    assert int(touched.sum()) >= 900  # the slices' columns sweep almost every class
    err = float((g_got - g_ref).norm() / g_ref.norm())
    print(f"c4 pair slices: worst distance rel err {worst:.2e}, dLoss/dE rel err {err:.2e}")
    assert err < GRAD_TOL
    if bool((~touched).any()):
        assert float(g_got[~touched].abs().max()) == 0.0
    assert abs(float(loss[0]) - float(total)) < DIST_TOL * abs(float(total)) and float(loss[1]) == 0


@pytest.mark.slow
def test_c4_full_closure_matches_oracle():
    """ONE full oracle evaluation of the c4 closure (1000 classes, 10^6 LAPACK eigenproblems and the
    eigh backward: about a minute of CPU) against the fused native closure."""
    got, ref = full_size_statistics("c4")
    _closure_vs_oracle("full", got, ref, 16, seed=4)


# ------------------------------------------------------------------------------------------------
# transform at the shapes of its tuned kernel
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize(
    "n,d,k,ld",
    [
        (50000, 3072, 8, None),    # c2, the shape DESIGN.md quotes
        (50000, 3072, 16, None),
        (50000, 3072, 32, None),
        (50001, 3072, 8, None),    # last row tile has one row
        (20000, 784, 4, None),     # c1
        (30000, 1024, 32, 1100),   # strided rows (a view into a wider matrix), still 16-byte aligned
        (30000, 1024, 16, 1027),   # row stride not a multiple of 4 floats: scalar fallback
        (4097, 104, 8, None),      # c3
    ],
)
def test_transform_tuned_shapes(n, d, k, ld):
    from sqfa_b200.model import SQFA

    g = torch.Generator(device="cuda").manual_seed(n + d + k)
    if ld is None:
        X = torch.randn(n, d, generator=g, device="cuda")
    else:
        X = torch.randn(n, ld, generator=g, device="cuda")[:, :d]
        assert X.stride(0) == ld and not X.is_contiguous()
    F = torch.randn(k, d, generator=g, device="cuda")
    model = SQFA(n_dim=d, n_filters=k, filters=F.cpu(), constraint="none").cuda()
    Z = model.transform(X)
    assert Z.shape == (n, k)
    ref = X.double() @ model.filters.detach().double().T
    assert rel_err(Z, ref) < 1e-5
    # per-row check: a wrong row (tile edge, stride) would hide in a Frobenius norm
    row_err = (Z.double() - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-30)
    assert float(row_err.max()) < 1e-4


# ------------------------------------------------------------------------------------------------
# pca_from_scatter and fit_pca(data_statistics=) against the reference's golden vector
# ------------------------------------------------------------------------------------------------
def _golden(name):
    with np.load(os.path.join(os.path.dirname(__file__), "golden", name)) as z:
        return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def test_pca_from_scatter_and_fit_pca_statistics_golden():
    """reference tests/test_training.py:184-188: fit_pca(data_statistics=) == pca_from_scatter of the
    second moments (the reference hands the mean scatter matrix to pca() as a point cloud)."""
    from sqfa_b200.model import SQFA, SecondMomentsSQFA
    from sqfa_b200.statistics import pca_from_scatter

    g = _golden("class_statistics.npz")
    sm = g["second_moments"].float().cuda()
    comps = pca_from_scatter(sm, 3)
    assert comps.shape == (3, 12)
    assert O.subspace_angle(comps.cpu(), g["pca_scatter3"]) < 1e-3
    # rows are unit eigenvectors, in descending order of variance: compare row by row up to sign
    for r in range(3):
        cos = abs(float(comps[r].double().cpu() @ g["pca_scatter3"][r]))
        assert cos > 1 - 1e-5, (r, cos)
    stats = {k: g[k].float().cuda() for k in ("means", "covariances")}
    for cls in (SecondMomentsSQFA, SQFA):
        model = cls(n_dim=12, n_filters=3)
        model.fit_pca(data_statistics=stats)
        assert model.filters.shape == (3, 12)
        assert O.subspace_angle(model.filters.detach().cpu(), g["pca_scatter3"]) < 1e-3
        model = cls(n_dim=12, n_filters=3)
        model.fit_pca(X=g["X"].float())
        assert O.subspace_angle(model.filters.detach().cpu(), g["pca3"]) < 1e-3
    with pytest.raises(ValueError):
        pca_from_scatter(sm, 13)


# ------------------------------------------------------------------------------------------------
# the NaN / inf guard
# ------------------------------------------------------------------------------------------------
NAN_MSG = "Some distances between classes are NaN. Try using float64 or a different regularization parameter."
INF_MSG = "Some distances between classes are inf. Try using float64 or a different regularization parameter."


def _small_stats(c=5, d=12, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(c, d, d + 4, generator=g)
    return {"means": 0.1 * torch.randn(c, d, generator=g), "covariances": A @ A.transpose(1, 2) / d}


@pytest.mark.parametrize("path", ["direct", "autograd", "generic"])
@pytest.mark.parametrize("poison", ["indefinite", "nan"])
def test_fit_guard_raises_reference_error(path, poison, monkeypatch):
    """A class matrix that is not SPD (negative definite) or holds NaN makes some pair distances NaN;
    the reference raises ValueError with its message from check_distances_valid (_optim.py:16-30)."""
    from sqfa_b200 import distances as Dn
    from sqfa_b200.model import SecondMomentsSQFA

    stats = _small_stats()
    scatters = (stats["covariances"] + torch.einsum("ci,cj->cij", stats["means"], stats["means"])).clone()
    if poison == "indefinite":
        scatters[2] = -scatters[2]
    else:
        scatters[3, 1, 1] = float("nan")
    dfun = None
    if path == "generic":  # a user-supplied distance_fun: native projection + torch ops + check_distances_valid
        def dfun(A, B):
            return Dn.affine_invariant(A, B)
    model = SecondMomentsSQFA(n_dim=12, feature_noise=0.0, n_filters=3, distance_fun=dfun)
    if path == "autograd":
        monkeypatch.setattr(model, "_fused_direct_plan", lambda data_statistics: None)
    with pytest.raises(ValueError) as info:
        model.fit(data_statistics=scatters, max_epochs=2, show_progress=False)
    assert str(info.value) == NAN_MSG


def test_fit_guard_inf_message_and_flag():
    """inf distances: the reference's second message. The native closure reports the number of
    non-finite pair distances in out[1]; check_distances_valid distinguishes NaN from inf."""
    from sqfa_b200._optim import check_distances_valid
    from sqfa_b200.model import SecondMomentsSQFA

    d = torch.full((4, 4), 1.0, device="cuda")
    check_distances_valid(d)
    d[2, 1] = float("inf")
    with pytest.raises(ValueError) as info:
        check_distances_valid(d)
    assert str(info.value) == INF_MSG
    d[2, 1] = 1.0
    d[1, 2] = float("nan")  # upper triangle is not inspected (reference reads the strict lower triangle)
    check_distances_valid(d)
    d[3, 0] = float("nan")
    with pytest.raises(ValueError) as info:
        check_distances_valid(d)
    assert str(info.value) == NAN_MSG
    # the fused kernel's counter: two poisoned classes of five -> 4 + 3 = 7 non-finite pairs
    stats = _small_stats()
    scatters = stats["covariances"].clone()
    scatters[0] = -scatters[0]
    scatters[4, 0, 0] = float("nan")
    model = SecondMomentsSQFA(n_dim=12, feature_noise=0.0, n_filters=3).cuda()
    out = model._fused_loss_plan(scatters.cuda())()
    assert float(out[1]) == 7.0
    assert float(out[0]) != float(out[0])  # NaN loss
