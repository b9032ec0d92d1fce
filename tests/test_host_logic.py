"""CPU: host-side logic of the drop-in API that needs no GPU -- constructor / validation error
types of the reference (tests/test_training.py:205-259), module state names, pair sharding."""

import pytest
import torch

from sqfa_b200 import _lib
from sqfa_b200._stats_driver import pair_range
from sqfa_b200.model import SQFA, SecondMomentsSQFA, _check_statistics


def test_namespace_mirrors_reference():
    import sqfa_b200

    for name in ("statistics", "linalg", "distances", "constraints", "_optim", "model"):
        assert hasattr(sqfa_b200, name)
    assert sqfa_b200.statistics.__all__ == ["class_statistics", "oas_covariance", "pca", "pca_from_scatter"]
    assert set(sqfa_b200.linalg.__all__) >= {"conjugate_matrix", "generalized_eigenvalues", "spd_log", "spd_inv_sqrt"}
    assert set(sqfa_b200.distances.__all__) >= {"affine_invariant", "fisher_rao_lower_bound", "log_euclidean"}


def test_module_state_names_and_defaults():
    m = SQFA(n_dim=8, feature_noise=0.01, n_filters=3)
    assert [n for n, _ in m.named_parameters()] == ["parametrizations.filters.original"]
    assert [n for n, _ in m.named_buffers()] == ["noise_mat"]
    assert m.noise_mat.shape == (3, 3) and m.noise_mat.dtype == torch.float32
    assert torch.allclose(m.filters.norm(dim=1), torch.ones(3), atol=1e-6)  # sphere constraint
    assert m.distance_fun.__name__ == "fisher_rao_lower_bound"
    assert SecondMomentsSQFA(n_dim=8).distance_fun.__name__ == "affine_invariant"
    assert SecondMomentsSQFA(n_dim=8, constraint="none").constraint == "none"
    o = SecondMomentsSQFA(n_dim=8, n_filters=2, constraint="orthogonal")
    assert torch.allclose(o.filters @ o.filters.T, torch.eye(2), atol=1e-5)
    sd = m.state_dict()
    m2 = SQFA(n_dim=8, feature_noise=0.01, n_filters=3)
    m2.load_state_dict(sd)
    assert torch.equal(m2.filters, m.filters)


def test_constructor_and_validation_errors():
    with pytest.raises(ValueError):
        SQFA(n_dim=4, n_filters=6)
    with pytest.raises(ValueError):
        SecondMomentsSQFA(n_dim=8).fit()
    with pytest.raises(ValueError):
        SecondMomentsSQFA(n_dim=8).fit(X=torch.randn(4, 8))
    with pytest.raises(ValueError):
        SQFA(n_dim=8).fit_pca()
    with pytest.raises(TypeError):
        _check_statistics([1, 2, 3])
    with pytest.raises(TypeError):
        _check_statistics(torch.zeros(2, 3, 3), needs_dict=True)
    with pytest.raises(ValueError):
        _check_statistics({"means": torch.zeros(2, 3)})
    with pytest.raises(TypeError):
        SQFA(n_dim=8).fit(data_statistics=torch.zeros(2, 8, 8), show_progress=False)
    with pytest.raises(TypeError):
        SQFA(n_dim=8).get_class_distances(torch.zeros(2, 8, 8))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_compute_fails_loudly_without_gpu():
    from sqfa_b200.statistics import class_statistics

    with pytest.raises(_lib.SqfaNativeError):
        class_statistics(torch.randn(16, 4), torch.zeros(16, dtype=torch.long))
    with pytest.raises(_lib.SqfaNativeError):
        SQFA(n_dim=4).transform_scatters(torch.eye(4)[None])


def test_estimator_and_dtype_validation():
    from sqfa_b200.statistics import class_statistics

    with pytest.raises(ValueError):
        class_statistics(torch.randn(4, 2), torch.zeros(4), estimator="bogus")


@pytest.mark.parametrize("n_pairs,world", [(0, 2), (1, 4), (45, 2), (499500, 8), (4950, 3)])
def test_pair_ranges_tile_the_pair_list(n_pairs, world):
    ranges = [pair_range(n_pairs, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n_pairs
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1


def test_lbfgs_on_cpu_parameters_is_torch_lbfgs():
    """sqfa_b200._lbfgs.LBFGS only takes the native path for one float32 CUDA parameter without line
    search; anything else must behave exactly like torch.optim.LBFGS (same iterates, same state)."""
    from sqfa_b200._lbfgs import LBFGS

    def run(cls, **kw):
        torch.manual_seed(0)
        x = torch.nn.Parameter(torch.randn(50))
        A = torch.diag(torch.linspace(0.5, 5.0, 50))
        opt = cls([x], lr=0.5, history_size=6, **kw)

        def closure():
            opt.zero_grad()
            loss = 0.5 * x @ A @ x - x.sum()
            loss.backward()
            return loss

        losses = [opt.step(closure).item() for _ in range(3)]
        return x.detach().clone(), losses, opt.state[opt._params[0]]

    for kw in ({}, {"line_search_fn": "strong_wolfe"}):
        x_ref, l_ref, _ = run(torch.optim.LBFGS, **kw)
        x_got, l_got, state = run(LBFGS, **kw)
        assert torch.equal(x_ref, x_got) and l_ref == l_got
        assert "sqfa_native" not in state


def test_streaming_statistics_need_the_gpu_and_validate_arguments():
    from sqfa_b200.statistics import StreamingClassStatistics

    with pytest.raises(ValueError):
        StreamingClassStatistics(0, 3)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.SqfaNativeError):
            StreamingClassStatistics(8, 3)


def test_direct_closure_plan_only_for_closed_form_constraints():
    """The graph-free closure needs a CUDA float32 parameter and constraints with a closed-form adjoint;
    on the CPU (this suite) and for torch's orthogonal parametrisation it must decline, so fitting_loop
    falls back to the autograd path instead of silently computing something else."""
    stats = {"means": torch.zeros(3, 6), "covariances": torch.eye(6).repeat(3, 1, 1)}
    for constraint in ("sphere", "none", "orthogonal"):
        model = SQFA(n_dim=6, n_filters=2, feature_noise=0.01, constraint=constraint)
        assert model._fused_direct_plan(stats) is None  # CPU parameter
    model = SecondMomentsSQFA(n_dim=6, n_filters=2, feature_noise=0.01)
    assert model._fused_direct_plan(stats["covariances"]) is None


def test_odd_even_transposition_meets_every_column_pair_once():
    """The pairing of the column-pair Jacobi kernel (pairs.cu, pair_cp_kernel): n alternating steps of
    adjacent pairs, every rotation followed by a swap of the two columns. One sweep must bring every pair
    of columns together exactly once and leave the columns in reversed order (the kernel locates the zero
    padding column of an odd m from the parity of the sweep count)."""
    for n in (2, 4, 6, 10, 18, 34):
        pos = list(range(n))
        met = set()
        for step in range(n):
            for p in range(step % 2, n - 1, 2):
                a, b = pos[p], pos[p + 1]
                key = (min(a, b), max(a, b))
                assert key not in met
                met.add(key)
                pos[p], pos[p + 1] = b, a
        assert len(met) == n * (n - 1) // 2
        assert pos == list(range(n))[::-1]


def test_column_pair_jacobi_model_matches_lapack():
    """float32 numpy model of the pair kernel's Jacobi (pairs.cu, pair_cp_kernel): one-sided rotations in the
    odd-even transposition ordering, every rotation followed by the swap, columns kept as scale * vector
    ("fast" rotations: y + tau1 x, x - tau2 y, scales multiplied by c and folded back once per sweep), carried
    norms, the same convergence rules. The squared column norms must be the generalized eigenvalues that
    LAPACK finds, and the zero padding column of an odd m must sit where the sweep parity says."""
    import numpy as np

    rng = np.random.default_rng(0)
    tol, last = 1e-6, 3e-4
    for m in (4, 9, 17, 33):
        n = m + (m & 1)
        a = rng.standard_normal((m, m + 4))
        b = rng.standard_normal((m, m + 4))
        Ei, Ej = a @ a.T / (m + 4) + 0.05 * np.eye(m), b @ b.T / (m + 4) + 0.05 * np.eye(m)
        Li, Lj = np.linalg.cholesky(Ei), np.linalg.cholesky(Ej)
        A = np.zeros((n, n), np.float32)
        A[:m, :m] = (np.linalg.solve(Lj, Li)).T.astype(np.float32)  # columns of (L_j^-1 L_i)^T
        scale = np.ones(n, np.float32)
        sweeps = 0
        for _ in range(24):
            A *= scale
            scale[:] = 1
            norms = (A * A).sum(0)
            rotated = False
            for step in range(n):
                for p in range(step % 2, n - 1, 2):
                    x, y = A[:, p].copy(), A[:, p + 1].copy()
                    al, be = norms[p], norms[p + 1]
                    ab = np.float32(scale[p] * scale[p + 1] * np.dot(x, y))
                    t, c, n0, n1 = np.float32(0), np.float32(1), be, al
                    if ab * ab > tol * tol * al * be:
                        d, h = be - al, ab + ab
                        hs = h if not np.signbit(d) else -h
                        t = np.float32(hs / (abs(d) + np.sqrt(d * d + h * h)))
                        c = np.float32(1 / np.sqrt(t * t + 1))
                        n0, n1 = be + t * ab, al - t * ab
                        rotated = rotated or ab * ab > last * last * al * be
                    tq = t / (scale[p] * scale[p + 1])
                    A[:, p] = y + (tq * scale[p] * scale[p]) * x          # position p <- c (y + t x)
                    A[:, p + 1] = x - (tq * scale[p + 1] * scale[p + 1]) * y  # position p + 1 <- c (x - t y)
                    scale[p], scale[p + 1] = c * scale[p + 1], c * scale[p]
                    norms[p], norms[p + 1] = n0, n1
            sweeps += 1
            if not rotated:
                break
        A *= scale
        lam = (A.astype(np.float64) ** 2).sum(0)
        if m & 1:  # the padding column: last position after an even number of sweeps, first after an odd number
            pad = 0 if sweeps & 1 else n - 1
            assert lam[pad] == 0.0
            lam = np.delete(lam, pad)
        ref = np.linalg.eigvalsh(np.linalg.solve(Lj, Ei) @ np.linalg.inv(Lj).T)
        assert sweeps <= 10
        assert np.allclose(np.sort(lam), ref, rtol=2e-5)
        d2 = (np.log(lam) ** 2).sum()
        assert abs(d2 - (np.log(ref) ** 2).sum()) <= 1e-4 * (np.log(ref) ** 2).sum()


class _ToyModel(torch.nn.Module):
    """A model with the one method `fitting_loop` needs, in plain CPU torch: distances between the rows of
    `stats` seen through a learnable diagonal scaling (bounded, so L-BFGS converges to a plateau)."""

    def __init__(self, n_dim, poison=None):
        super().__init__()
        self.w = torch.nn.Parameter(torch.linspace(-0.5, 0.5, n_dim))
        self.poison = poison

    def get_class_distances(self, data_statistics, regularized=False):
        z = data_statistics * torch.tanh(self.w)
        d = (z[:, None, :] - z[None, :, :]).pow(2).sum(-1)
        if self.poison is not None:
            d = d + torch.tril(torch.full_like(d, self.poison), -1)
        return d


@pytest.mark.parametrize("max_epochs,atol", [(30, 1e-6), (2, 1e-6), (0, 1e-6), (6, 1e3)])
def test_fitting_loop_host_semantics_match_reference(max_epochs, atol, capsys):
    """Epoch loop, stopping rule, history and messages of `fitting_loop` against the REAL reference's
    (_optim.py:78-145) on a toy CPU model: same losses, same epoch count, same final message."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference sources not present")
    from sqfa_b200._optim import fitting_loop

    R = ref_loader.load()
    stats = torch.randn(5, 7, generator=torch.Generator().manual_seed(3))
    out = {}
    for name, loop in (("ref", R._optim.fitting_loop), ("ours", fitting_loop)):
        model = _ToyModel(7)
        loss, seconds = loop(model, stats, max_epochs=max_epochs, lr=0.1, atol=atol, show_progress=False,
                             return_loss=True)
        captured = capsys.readouterr()
        out[name] = (loss, seconds, model.w.detach().clone(), captured.out + captured.err)
    (l_ref, t_ref, w_ref, msg_ref), (l_got, t_got, w_got, msg_got) = out["ref"], out["ours"]
    assert l_got.shape == l_ref.shape == t_got.shape and torch.equal(l_got, l_ref)
    assert torch.equal(w_got, w_ref)
    assert msg_got.strip() == msg_ref.strip()
    assert fitting_loop(_ToyModel(7), stats, max_epochs=1, show_progress=False) is None


@pytest.mark.parametrize("poison,word", [(float("nan"), "NaN"), (float("inf"), "inf")])
def test_fitting_loop_guard_messages_match_reference(poison, word):
    """NaN / inf distances raise the reference's ValueError texts (_optim.py:16-30) on the generic path."""
    from oracle import ref_loader

    from sqfa_b200._optim import fitting_loop

    stats = torch.randn(4, 3, generator=torch.Generator().manual_seed(1))
    with pytest.raises(ValueError) as got:
        fitting_loop(_ToyModel(3, poison), stats, max_epochs=2, show_progress=False)
    assert word in str(got.value)
    if ref_loader.available():
        with pytest.raises(ValueError) as ref:
            ref_loader.load()._optim.fitting_loop(_ToyModel(3, poison), stats, max_epochs=2, show_progress=False)
        assert str(got.value) == str(ref.value)


@pytest.mark.parametrize("kind", ["SQFA", "SecondMomentsSQFA"])
@pytest.mark.parametrize("constraint", ["sphere", "none", "orthogonal"])
def test_state_dict_interchange_with_reference(kind, constraint):
    """A model saved by the reference loads here and the other way round: same state-dict keys, shapes and
    dtypes, same constrained filters from the same raw parameter (reference model.py:148-170, constraints.py)."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference sources not present")
    import sqfa_b200

    R = ref_loader.load()
    F0 = torch.randn(3, 9, generator=torch.Generator().manual_seed(7))
    kw = dict(n_dim=9, feature_noise=0.02, n_filters=3, filters=F0.clone(), constraint=constraint)
    ref, ours = getattr(R.model, kind)(**kw), getattr(sqfa_b200.model, kind)(**kw)
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert list(sd_ref) == list(sd_ours)
    for key in sd_ref:
        assert sd_ref[key].shape == sd_ours[key].shape and sd_ref[key].dtype == sd_ours[key].dtype, key
    assert torch.allclose(ref.filters, ours.filters, atol=1e-6)
    assert torch.equal(ref.noise_mat, ours.noise_mat)

    # fresh models with other filters, then cross-load
    kw["filters"] = torch.randn(3, 9, generator=torch.Generator().manual_seed(8))
    ref2, ours2 = getattr(R.model, kind)(**kw), getattr(sqfa_b200.model, kind)(**kw)
    ours2.load_state_dict(sd_ref)
    ref2.load_state_dict(sd_ours)
    assert torch.allclose(ours2.filters, ref.filters, atol=1e-6)
    assert torch.allclose(ref2.filters, ours.filters, atol=1e-6)
    assert type(ours.distance_fun).__name__ == type(ref.distance_fun).__name__
    assert ours.distance_fun.__name__ == ref.distance_fun.__name__ and ours.constraint == ref.constraint


def test_tf32_split_model_keeps_the_gram_at_float32_accuracy():
    """numpy model of the operand split of the tcgen05 Gram and projection kernels (ptx.cuh `to_tf32`,
    gram.cu / project_tc.cu producers): hi = (bits + 0x1000) & 0xFFFFE000 (round to nearest on the 10-bit
    TF32 mantissa, two integer instructions), lo = x - hi in float32. The split must be exact (hi + lo == x),
    hi must be a TF32 number, |lo| <= 2^-11 |x|, and the three products the kernels issue
    (hi hi + hi lo + lo hi, the tensor core reading only the TF32 part of lo) must reproduce the float64 Gram
    to ~2^-21 -- well inside the 1e-5 the statistics are held to."""
    import numpy as np

    def to_tf32(x):
        return ((x.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)

    def tf32_part(x):  # what kind::tf32 reads of an fp32 operand word: the low 13 mantissa bits are ignored
        return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)

    rng = np.random.default_rng(0)
    x = (rng.standard_normal((4096, 48)) * np.exp(rng.uniform(-3, 3, (1, 48)))).astype(np.float32)
    x[0, :4] = [0.0, -0.0, 1.0, -1.5]  # exactly representable values split into (x, 0)
    hi = to_tf32(x)
    lo = x - hi
    assert hi.dtype == np.float32 and lo.dtype == np.float32
    assert np.array_equal(hi + lo, x)  # exact split
    assert not (hi.view(np.uint32) & np.uint32(0x1FFF)).any()
    assert (np.abs(lo) <= np.abs(x) * 2.0**-11).all()
    assert np.array_equal(lo[0, :4], np.zeros(4, np.float32))
    # ties round away from zero (add half an ulp to the magnitude, then truncate)
    tie = np.array([1.0 + 2.0**-11, -(1.0 + 2.0**-11)], dtype=np.float32)
    assert np.array_equal(to_tf32(tie), np.array([1.0 + 2.0**-10, -(1.0 + 2.0**-10)], dtype=np.float32))
    # non-finite inputs stay non-finite (the guard downstream reports them; nothing is silently zeroed)
    bad = to_tf32(np.array([np.inf, -np.inf, np.nan], dtype=np.float32))
    assert np.isinf(bad[:2]).all() and np.isnan(bad[2])

    h, l = hi.astype(np.float64), tf32_part(lo).astype(np.float64)
    model = h.T @ h + h.T @ l + l.T @ h
    truth = x.astype(np.float64).T @ x.astype(np.float64)
    scale = np.sqrt(np.outer(np.diag(truth), np.diag(truth)))  # entry-wise, relative to the diagonal scale
    assert (np.abs(model - truth) / scale).max() < 2.0**-20
    # one TF32 product alone would not do: 2^-11 operand rounding
    assert (np.abs(h.T @ h - truth) / scale).max() > 2.0**-16


def test_design_document_cites_existing_tests_and_files():
    """DESIGN.md / README.md name tests (`file.py::test_name`) and source files as evidence: every one must exist."""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "DESIGN.md")).read() + open(os.path.join(root, "README.md")).read()
    sources = {f: open(os.path.join(root, "tests", f)).read() for f in os.listdir(os.path.join(root, "tests"))
               if f.endswith(".py")}
    last_file = None
    for m in re.finditer(r"(?:(test_\w+\.py))?::(test_\w+)", text):
        last_file = m.group(1) or last_file
        name = m.group(2).rstrip("_")
        assert last_file in sources, last_file
        hay = sources[last_file] if m.group(1) else "".join(sources.values())
        assert re.search(r"def " + re.escape(name), hay), (last_file, name)
    for f in set(re.findall(r"`(?:tests/)?(test_\w+\.py)", text)):
        assert f in sources, f
    for f in set(re.findall(r"`((?:sqfa_b200|tools|oracle|profiles|include)/[\w/\.]+\.(?:py|cu|cuh|h|sh|json|txt|csv|md))`", text)):
        assert os.path.exists(os.path.join(root, f)), f


def test_partitions_hold_for_arbitrary_sizes():
    """Property test (hypothesis) of the three partitions the multi-GPU paths rest on, for arbitrary class
    counts and world sizes: `shard_pairs` (pair list: contiguous, complete, whole rows at multiples of 4),
    `class_share` (classes: contiguous, complete, equal to the Gram kernel's completion-counter groups
    c * W / C) and `peer_push_schedule` (every foreign non-empty group once, distinct destinations per step)."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from sqfa_b200._ops import shard_pairs
    from sqfa_b200._stats_driver import class_share, peer_push_schedule

    @settings(max_examples=300, deadline=None)
    @given(st.integers(min_value=1, max_value=5000), st.integers(min_value=1, max_value=16))
    def check(C, W):
        P = C * (C - 1) // 2
        cuts = [shard_pairs(P, C, r, W) for r in range(W)]
        assert cuts[0][0] == 0 and cuts[-1][1] == P
        for (a0, a1), (b0, b1) in zip(cuts[:-1], cuts[1:]):
            assert a0 <= a1 == b0 <= b1
        for _, e in cuts[:-1]:  # a cut is the first pair of a row i with i % 4 == 0 (or an end of the list)
            i = (1 + int(round((1 + 8 * e) ** 0.5))) // 2
            assert i * (i - 1) // 2 == e and (i % 4 == 0 or e in (0, P))
        shares = [class_share(C, r, W) for r in range(W)]
        assert shares[0][0] == 0 and shares[-1][1] == C
        assert all(a[1] == b[0] and a[0] <= a[1] for a, b in zip(shares, shares[1:] + [(C, C)]))
        owner = [c * W // C for c in range(C)]
        for r, (lo, hi) in enumerate(shares):
            assert owner[lo:hi] == [r] * (hi - lo)
        for r in range(W):
            sched = peer_push_schedule(r, W, shares)
            groups = [g for g, _, _ in sched]
            assert r not in groups and len(set(groups)) == len(groups)
            assert set(groups) == {g for g in range(W) if g != r and shares[g][1] > shares[g][0]}
            assert all((lo, hi) == shares[g] for g, lo, hi in sched)

    check()
