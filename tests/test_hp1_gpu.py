"""GPU parity of hot path 1 (class_statistics) against the CPU oracle, through the C ABI."""

import os

import pytest
import torch

from conftest import make_class_data, rel_err
from oracle import sqfa_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north_star: class means and second moments within 1e-5 relative (fp32)


def _probe(K, N):
    """D[128 x N] = A^T B through the test-only probe library (tests/native/umma_probe.cu)."""
    import ctypes

    from sqfa_b200 import _lib, build

    path = build.PROBE_LIB if os.path.exists(build.PROBE_LIB) else build.build_probe()
    lib = ctypes.CDLL(path)
    u32, i32, ptr = ctypes.c_uint32, ctypes.c_int32, ctypes.c_void_p
    lib.sqfa_debug_umma_probe.restype = ctypes.c_int
    lib.sqfa_debug_umma_probe.argtypes = [ptr, ptr, ptr, i32, i32, i32, u32, u32, u32, u32, u32, u32, ptr]
    g = torch.Generator().manual_seed(1)
    A = torch.randint(-4, 5, (K, 128), generator=g).float().cuda()
    B = torch.randint(-4, 5, (K, N), generator=g).float().cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = lib.sqfa_debug_umma_probe(
        _lib.ptr(A), _lib.ptr(B), _lib.ptr(D), K, N, 1, 128 * 16, 128, 0, 0, 0, 2 * 128 * 16, _lib.stream_ptr()
    )
    torch.cuda.synchronize()
    assert rc == 0
    return (D.double() - A.double().T @ B.double()).abs().max().item()


@pytest.mark.parametrize("K,N", [(8, 32), (16, 256), (32, 128)])
def test_umma_operand_layout(K, N):
    """Small-integer operands: products are exact in tf32, any error is a layout bug."""
    assert _probe(K, N) == 0.0


@pytest.mark.parametrize("n,c", [(1000, 3), (5000, 10), (8193, 7), (50000, 10), (200000, 19), (300000, 255), (4099, 300),
                                 (70000, 1000)])
def test_bucket_labels_bit_exact(n, c):
    from sqfa_b200.statistics import bucket_labels

    g = torch.Generator().manual_seed(n + c)
    y = torch.randint(0, c, (n,), generator=g)
    perm, offsets, counts = bucket_labels(y.cuda())
    ref = torch.sort(y, stable=True).indices
    assert torch.equal(perm.cpu().long(), ref)
    ref_counts = torch.bincount(y, minlength=int(y.max()) + 1)
    assert torch.equal(counts.cpu()[:-1], ref_counts)
    assert int(counts[-1]) == 0
    assert torch.equal(offsets.cpu()[1:-1], torch.cumsum(ref_counts, 0))


def test_bucket_labels_out_of_range_rows_dropped():
    from sqfa_b200.statistics import bucket_labels

    y = torch.tensor([2, -1, 0, 7, 1, 0, -5, 2, 1, 9])
    perm, offsets, counts = bucket_labels(y.cuda(), n_classes=3)
    assert perm.cpu().tolist() == [2, 5, 4, 8, 0, 7, 1, 3, 6, 9]
    assert counts.cpu().tolist() == [2, 2, 2, 4]
    assert offsets.cpu().tolist() == [0, 2, 4, 6, 10]


@pytest.mark.parametrize(
    "n,d,c,kw",
    [
        (1000, 4, 3, {}),                       # reference's own test shape (tests/test_statistics.py:24)
        (3000, 104, 19, {}),                    # stereo-disparity D
        (6000, 784, 10, {}),                    # MNIST D (ragged tiles: 784 = 3*256 + 16)
        (4000, 512, 40, {"skew": True}),        # ragged class sizes
        (2500, 300, 5, {"offset": 3.0}),        # uncentred data: cancellation stress
        (1500, 1027, 4, {}),                    # D not a multiple of 4 -> scalar load path
        (5000, 3072, 2, {}),                    # CIFAR D, full 128x256 tiles
    ],
)
def test_class_statistics_matches_oracle(n, d, c, kw):
    from sqfa_b200.statistics import class_statistics

    X, y = make_class_data(n, d, c, seed=d, **kw)
    got = class_statistics(X.cuda(), y.cuda())
    ref32 = O.class_statistics(X, y)
    ref64 = O.class_statistics(X.double(), y)
    for key in ("means", "covariances", "second_moments"):
        assert got[key].shape == ref32[key].shape
        assert rel_err(got[key], ref64[key]) < TOL, key
        assert rel_err(got[key], ref32[key]) < TOL, key
    # symmetric output, both triangles written
    cov = got["covariances"]
    assert torch.equal(cov, cov.transpose(1, 2))


def test_class_statistics_constant_data():
    """Known-answer test of the reference: tests/test_statistics.py:24-50."""
    from sqfa_b200.statistics import class_statistics

    X = torch.ones(1000, 4)
    y = torch.randint(0, 3, (1000,))
    s = class_statistics(X, y)  # CPU in -> CPU out
    assert s["means"].device.type == "cpu"
    assert s["means"].shape == (3, 4) and s["covariances"].shape == (3, 4, 4)
    assert torch.allclose(s["means"], torch.ones(3, 4), atol=1e-6)
    assert torch.allclose(s["covariances"], torch.zeros(3, 4, 4), atol=1e-5)
    assert torch.allclose(s["second_moments"], torch.ones(3, 4, 4), atol=1e-5)


def test_class_statistics_oas_and_edge_classes():
    from sqfa_b200.statistics import class_statistics

    X, y = make_class_data(2000, 64, 6, seed=3)
    y[y == 4] = 5          # class 4: a single row
    y[0] = 4
    y[y == 2] = 1          # class 2: empty
    got = class_statistics(X.cuda(), y.cuda(), estimator="oas")
    ref = O.class_statistics(X, y, estimator="oas")
    for key in ("means", "covariances", "second_moments"):
        g, r = got[key].cpu(), ref[key]
        assert torch.equal(torch.isnan(g), torch.isnan(r)), key
        ok = ~torch.isnan(r)
        assert rel_err(g[ok], r[ok]) < TOL, key


def test_sample_covariance_and_pca():
    from sqfa_b200 import statistics as S

    X, _ = make_class_data(3000, 96, 1, seed=5)
    assert rel_err(S.sample_covariance(X.cuda()), O.sample_covariance(X.double())) < TOL
    assert rel_err(S.sample_covariance(X.cuda(), assume_centered=True),
                   O.sample_covariance(X.double(), assume_centered=True)) < TOL
    assert rel_err(S.oas_covariance(X.cuda()), O.oas_covariance(X.double())) < TOL
    comp = S.pca(X.cuda(), 4).cpu()
    assert comp.shape == (4, 96)
    assert O.subspace_angle(comp, O.pca(X.double(), 4).float()) < 1e-3
    with pytest.raises(ValueError):
        S.pca(X.cuda(), 97)


@pytest.mark.parametrize("n,d,c,kw", [(20000, 256, 3, {}), (9000, 520, 7, {"skew": True}), (700, 2048, 2, {})])
def test_gram_long_chains_and_k_split(n, d, c, kw):
    """Classes much longer than one accumulation chain (512 samples) and few tiles (K split with
    red.add): the per-chain drains keep the tensor-core truncation bias out of the result."""
    from sqfa_b200.statistics import class_statistics

    X, y = make_class_data(n, d, c, seed=n, **kw)
    ref64 = O.class_statistics(X.double(), y)
    got = class_statistics(X.cuda(), y.cuda())
    for key in ("means", "covariances", "second_moments"):
        assert rel_err(got[key], ref64[key]) < TOL / 3, key


@pytest.mark.parametrize("n,d,c", [(6000, 784, 4), (3000, 1027, 3), (9000, 520, 7), (4000, 200, 5)])
def test_packed_gram_exchange_layout(n, d, c):
    """The packed upper-tile Gram (the buffer ranks all-reduce) gives bit-identical statistics to the
    full (C, D, D) layout, also for D that is not a multiple of the tile or of 4 and with a K split."""
    from sqfa_b200 import _lib
    from sqfa_b200._stats_driver import CudaStatsOps

    X, y = make_class_data(n, d, c, seed=d)
    X, y = X.cuda(), y.cuda()
    ops = CudaStatsOps()
    perm, offsets, counts = ops.bucket(y, c)
    means = ops.class_means(ops.class_sums(X, perm, offsets, c), counts[:c].clone())
    full = ops.class_gram(X, perm, offsets, means, c)
    packed = ops.class_gram(X, perm, offsets, means, c, packed=True)
    assert packed.numel() == _lib.load().sqfa_gram_packed_floats(d, c)
    cov_f, sm_f = ops.finalize(full, means, counts[:c].clone(), 0, 1, True)
    cov_p, sm_p = ops.finalize(packed, means, counts[:c].clone(), 0, 1, True, packed=True)
    assert torch.equal(cov_f, cov_p) and torch.equal(sm_f, sm_p)
    cov_o, sm_o = ops.finalize(packed, means, counts[:c].clone(), 1, 1, True, packed=True)  # OAS from packed
    ref = O.class_statistics(X.double().cpu(), y.cpu(), estimator="oas")
    assert rel_err(cov_o, ref["covariances"]) < TOL and rel_err(sm_o, ref["second_moments"]) < TOL


@pytest.mark.parametrize("n,d,c,world", [(6000, 512, 10, 8), (5000, 784, 3, 4), (9000, 520, 7, 2)])
def test_reduce_scatter_building_blocks_on_one_device(n, d, c, world):
    """The pieces of the fused reduce-scatter by class (multi-GPU class_statistics with sharded output),
    exercised on ONE device with `world` emulated ranks that each hold a slice of the rows:
    * sqfa_class_gram with completion counters: classes run owner group by owner group, rotated so that a
      rank's own classes come last; the tiles equal those of the default job order and every group's counter
      ends at sqfa_class_gram_group_signals;
    * sqfa_peer_push copies a finished group into the owner's slot [source rank];
    * sqfa_stats_epilogue_reduce sums the ranks' partials in rank order: statistics of the union."""
    import ctypes

    from sqfa_b200 import _lib
    from sqfa_b200._stats_driver import CudaStatsOps, class_share, peer_push_schedule

    lib = _lib.load()
    X, y = make_class_data(n, d, c, seed=3 * d)
    X, y = X.cuda(), y.cuda()
    ops = CudaStatsOps()
    perm_all, off_all, cnt_all = ops.bucket(y, c)
    counts = cnt_all[:c].clone()
    means = ops.class_means(ops.class_sums(X, perm_all, off_all, c), counts)
    shares = [class_share(c, r, world) for r in range(world)]
    per_class = lib.sqfa_gram_packed_floats(d, c) // c
    stride = max(hi - lo for lo, hi in shares) * per_class
    cuts = [(n * r) // world for r in range(world + 1)]
    # every "rank": packed partial Gram of its rows, centred by the global means
    slots = torch.full((world, world, stride), float("nan"), device="cuda")  # [owner][source][share]
    partials = []
    for r in range(world):
        Xr, yr = X[cuts[r]:cuts[r + 1]], y[cuts[r]:cuts[r + 1]]
        perm, offsets, _ = ops.bucket(yr, c)
        plain = ops.class_gram(Xr, perm, offsets, means, c, packed=True)
        done = torch.zeros(world, dtype=torch.int32, device="cuda")
        own_hi = shares[r][1]
        rotated = ops.class_gram(Xr, perm, offsets, means, c, packed=True, done=done, n_groups=world,
                                 first_class=own_hi % c)
        # same tiles in another job order (entries of edge tiles beyond D are never written: compare through
        # the epilogue; K parts of small problems are summed with red.add, hence a tolerance)
        ca = ops.finalize(plain, means, counts, 0, 1, False, packed=True)[0]
        cb = ops.finalize(rotated, means, counts, 0, 1, False, packed=True)[0]
        both = torch.isfinite(ca) & torch.isfinite(cb)  # classes absent from this slice divide 0 by n - 1 < 0
        assert float((ca[both] - cb[both]).norm()) <= 1e-6 * float(ca[both].norm())
        expected = [lib.sqfa_class_gram_group_signals(Xr.shape[0], d, c, world, g) for g in range(world)]
        assert done.tolist() == expected
        partials.append(rotated)
        for g, lo, hi in peer_push_schedule(r, world, shares):
            dst = slots[g, r].data_ptr()
            src = rotated.data_ptr() + 4 * lo * per_class
            _lib.check(lib.sqfa_peer_push(ctypes.c_void_p(dst), ctypes.c_void_p(src), 4 * (hi - lo) * per_class,
                                          _lib.stream_ptr()), "sqfa_peer_push")
    ref = O.class_statistics(X.double().cpu(), y.cpu())
    for r in range(world):
        lo, hi = shares[r]
        nc = hi - lo
        if nc == 0:
            continue
        cov = torch.empty(nc, d, d, device="cuda")
        sm = torch.empty(nc, d, d, device="cuda")
        ws = torch.empty(lib.sqfa_stats_epilogue_workspace_bytes(nc), dtype=torch.uint8, device="cuda")
        own = partials[r].data_ptr() + 4 * lo * per_class
        _lib.check(lib.sqfa_stats_epilogue_reduce(
            ctypes.c_void_p(own), _lib.ptr(slots[r]), stride, world, r, _lib.ptr(means[lo:hi]), None,
            _lib.ptr(counts[lo:hi]), d, nc, 0, 1, _lib.ptr(cov), _lib.ptr(sm), _lib.ptr(ws), ws.numel(),
            _lib.stream_ptr()), "sqfa_stats_epilogue_reduce")
        assert rel_err(cov, ref["covariances"][lo:hi]) < TOL
        assert rel_err(sm, ref["second_moments"][lo:hi]) < TOL
        # the same sum formed on the host side in rank order, through the ordinary epilogue: identical bits
        total = torch.zeros(nc * per_class, device="cuda")
        for q in range(world):
            total += partials[q][lo * per_class:hi * per_class]
        total = torch.nan_to_num(total)  # (never-written padding of edge tiles)
        cov2, _ = ops.finalize(total, means[lo:hi], counts[lo:hi], 0, 1, True, packed=True)
        assert torch.equal(cov, cov2)


@pytest.mark.parametrize("n,d,c,est", [(5000, 512, 6, "empirical"), (3000, 203, 4, "oas"), (2000, 1027, 3, "empirical")])
def test_single_call_entry_matches_stepwise(n, d, c, est):
    """sqfa_class_statistics (one host call) enqueues exactly the step-by-step sequence: identical bits."""
    from sqfa_b200._stats_driver import CudaStatsOps, run_class_statistics

    X, y = make_class_data(n, d, c, seed=7 * d)
    X, y = X.cuda(), y.cuda()
    ops = CudaStatsOps()
    est_id = 1 if est == "oas" else 0
    m1, cov1, sm1, (perm1, off1, cnt1) = run_class_statistics(ops, X, y, est_id)  # fused path
    ops.gram_events = []  # instrumented -> step by step
    m2, cov2, sm2, (perm2, off2, cnt2) = run_class_statistics(ops, X, y, est_id)
    ops.gram_events = None
    assert torch.equal(m1, m2) and torch.equal(cov1, cov2) and torch.equal(sm1, sm2)
    assert torch.equal(perm1, perm2) and torch.equal(off1, off2) and torch.equal(cnt1, cnt2)
    ref = O.class_statistics(X.double().cpu(), y.cpu(), estimator=est)
    assert rel_err(cov1, ref["covariances"]) < TOL


def test_class_count_change_between_calls():
    """Once the class count has repeated, a call is enqueued for that count before the label maximum has reached
    the host; a call whose labels have another maximum must notice and run again with the right count."""
    from sqfa_b200.statistics import class_statistics

    d = 96
    Xa, ya = make_class_data(3000, d, 5, seed=1)
    Xb, yb = make_class_data(3000, d, 8, seed=2)
    Xc, yc = make_class_data(3000, d, 3, seed=3)
    for _ in range(3):
        got = class_statistics(Xa.cuda(), ya.cuda())
    assert got["means"].shape[0] == 5
    for X, y, c in ((Xb, yb, 8), (Xc, yc, 3), (Xa, ya, 5), (Xa, ya, 5), (Xa, ya, 5), (Xb, yb, 8)):
        got = class_statistics(X.cuda(), y.cuda())
        ref = O.class_statistics(X.double(), y)
        assert got["covariances"].shape == (c, d, d)
        for key in ref:
            assert rel_err(got[key], ref[key]) < TOL


def test_label_max_published_to_mapped_host_memory():
    """The label maximum lands in pinned host memory without a device-to-host copy, also when calls
    from several streams are in flight and for empty / all-negative labels."""
    from sqfa_b200._stats_driver import CudaStatsOps

    ops = CudaStatsOps()
    g = torch.Generator().manual_seed(3)
    streams = [torch.cuda.Stream() for _ in range(4)]
    for rep in range(40):  # more calls than scratch slots
        hi = int(torch.randint(1, 5000, (1,), generator=g))
        y = torch.randint(0, hi, (int(torch.randint(1, 200000, (1,), generator=g)),), generator=g)
        with torch.cuda.stream(streams[rep % 4]):
            yd = y.cuda(non_blocking=False)
            assert ops.label_max_host(yd) == int(y.max())
    assert ops.label_max_host(torch.empty(0, dtype=torch.int64, device="cuda")) == -1
    assert ops.label_max_host(torch.full((1000,), -7, dtype=torch.int64, device="cuda")) == -1


@pytest.mark.parametrize("est", ["empirical", "oas"])
def test_streaming_statistics_match_one_shot(est):
    """Chunked accumulation (ragged chunks, a class absent from the first chunk, rows with labels out
    of range) equals class_statistics on the concatenation and the oracle."""
    from sqfa_b200.statistics import StreamingClassStatistics, class_statistics

    n, d, c = 20000, 300, 7
    X, y = make_class_data(n, d, c, seed=11)
    X = X + 3.0  # means far from zero: the shift matters
    order = torch.argsort((y == 6).float(), stable=True)  # class 6 only shows up late
    X, y = X[order].contiguous(), y[order].contiguous()
    ref = O.class_statistics(X.double(), y, estimator=est)
    one = class_statistics(X.cuda(), y.cuda(), estimator=est)
    acc = StreamingClassStatistics(d, c)
    cuts = [0, 1, 5000, 5000, 12001, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        acc.update(X[a:b], y[a:b])  # CPU chunks: uploaded by update
    junk = torch.randn(50, d)
    acc.update(junk, torch.full((50,), c + 3))  # out-of-range labels are dropped
    got = acc.finalize(estimator=est)
    assert acc.n_rows == n + 50
    for key in ("means", "covariances", "second_moments"):
        assert rel_err(got[key], ref[key]) < TOL, key
        assert rel_err(got[key], one[key].double().cpu()) < TOL, key
