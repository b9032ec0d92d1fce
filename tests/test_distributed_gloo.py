"""CPU, world_size 2 / 3, gloo: (i) the sample-sharded class_statistics driver (three all-reduces) with
the local kernels replaced by oracle-backed CPU ops -- sharded == unsharded; (ii) the pair-sharded closure
(shard_pairs slices, one all-reduce of [loss, flag, dF], constraint adjoint after it) with the oracle's
arithmetic in place of the kernels -- equals the replicated closure, identical on every rank."""

import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleStatsOps:
    """CPU stand-ins for CudaStatsOps' local steps, written with the oracle's arithmetic."""

    def label_max(self, y):
        return y.max().reshape(1).clone() if y.numel() else torch.tensor([-1])

    def bucket(self, y, C):
        key = torch.where((y < 0) | (y >= C), torch.full_like(y, C), y)
        perm = torch.sort(key, stable=True).indices.to(torch.int32)
        counts = torch.bincount(key, minlength=C + 1)
        offsets = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(counts, 0)])
        return perm, offsets, counts

    def class_sums(self, X, perm, offsets, C, out=None):
        sums = torch.stack([X[perm[offsets[c]:offsets[c + 1]].long()].sum(0) for c in range(C)])
        if out is None:
            return sums
        out.copy_(sums)
        return out

    def class_means(self, sums, counts):
        return sums / counts[:, None].to(sums.dtype)

    def class_gram(self, X, perm, offsets, centre, C):
        out = []
        for c in range(C):
            Z = X[perm[offsets[c]:offsets[c + 1]].long()] - centre[c]
            out.append(Z.T @ Z)
        return torch.stack(out)

    def finalize(self, gram, means, counts, estimator_id, ddof, want_sm, class_range=None):
        c0, c1 = (0, means.shape[0]) if class_range is None else class_range
        cov = gram[c0:c1] / (counts[c0:c1, None, None].to(gram.dtype) - ddof)
        return cov, cov + means[c0:c1, :, None] * means[c0:c1, None, :]


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sqfa_b200._stats_driver import run_class_statistics

    g = torch.Generator().manual_seed(1)
    X = torch.randn(1000, 7, generator=g, dtype=torch.float64) + 0.5
    y = torch.randint(0, 5, (1000,), generator=g)
    y[y == 4] = 3 if rank == 0 else 4  # class 4 only exists on rank 1's shard
    lo, hi = (0, 400) if rank == 0 else (400, 1000)  # uneven shards
    y_full = torch.randint(0, 5, (1000,), generator=torch.Generator().manual_seed(1))
    means, cov, sm, _ = run_class_statistics(OracleStatsOps(), X[lo:hi], y[lo:hi], 0, group=dist.group.WORLD)
    if rank == 0:
        ret["means"], ret["cov"], ret["sm"] = means, cov, sm
    # sharded output: every rank finalises its share of the classes
    means_s, cov_s, sm_s, _ = run_class_statistics(OracleStatsOps(), X[lo:hi], y[lo:hi], 0, group=dist.group.WORLD,
                                                   shard_output=True)
    ret[f"share{rank}"] = (means_s, cov_s, sm_s)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_statistics_equal_unsharded():
    from oracle import sqfa_oracle as O

    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        means, cov, sm = ret["means"], ret["cov"], ret["sm"]
        shares = [ret["share0"], ret["share1"]]
    g = torch.Generator().manual_seed(1)
    X = torch.randn(1000, 7, generator=g, dtype=torch.float64) + 0.5
    y = torch.randint(0, 5, (1000,), generator=g)
    y0, y1 = y.clone(), y.clone()
    y0[y0 == 4] = 3
    y_union = torch.cat([y0[:400], y1[400:]])
    ref = O.class_statistics(X, y_union)
    assert torch.allclose(means, ref["means"], rtol=1e-10, atol=1e-12)
    assert torch.allclose(cov, ref["covariances"], rtol=1e-9, atol=1e-12)
    assert torch.allclose(sm, ref["second_moments"], rtol=1e-9, atol=1e-12)
    # the two ranks' shares, concatenated in rank order, are the full result (classes 0-2 | 3-4)
    assert shares[0][1].shape[0] == 3 and shares[1][1].shape[0] == 2
    assert torch.equal(shares[0][0], means) and torch.equal(shares[1][0], means)
    assert torch.equal(torch.cat([shares[0][1], shares[1][1]]), cov)
    assert torch.equal(torch.cat([shares[0][2], shares[1][2]]), sm)


def test_pair_shards_cover_the_pair_list():
    """shard_pairs: contiguous, disjoint, complete; cuts fall on rows that are multiples of 4."""
    from sqfa_b200._ops import shard_pairs
    from sqfa_b200._stats_driver import class_share

    for C in (2, 5, 10, 19, 100, 1000):
        P = C * (C - 1) // 2
        for world in (1, 2, 3, 4, 8):
            cuts = [shard_pairs(P, C, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == P
            for (a0, a1), (b0, b1) in zip(cuts[:-1], cuts[1:]):
                assert a1 == b0 and a0 <= a1
            for b, e in cuts[:-1]:
                i = int((1 + (1 + 8 * e) ** 0.5) / 2)  # e = i (i - 1) / 2 for a row i that is a multiple of 4
                assert i * (i - 1) // 2 == e and (i % 4 == 0 or e == P or e == 0)
            if C == 1000 and world == 8:  # balanced to a few percent
                sizes = [e - b for b, e in cuts]
                assert max(sizes) / min(sizes) < 1.1
            shares = [class_share(C, r, world) for r in range(world)]
            assert shares[0][0] == 0 and shares[-1][1] == C
            assert all(a[1] == b[0] for a, b in zip(shares[:-1], shares[1:]))
            # a rank's share is the class group of the Gram kernel's completion counter `rank`
            for r, (lo, hi) in enumerate(shares):
                assert all((c * world) // C == r for c in range(lo, hi))


def test_peer_push_schedule_is_staggered_and_complete():
    """Fused reduce-scatter: every rank pushes every non-empty foreign group exactly once, never its own, and
    at each step the destinations of the ranks are pairwise different (no peer is hit by two pushes)."""
    from sqfa_b200._stats_driver import class_share, peer_push_schedule

    for C in (3, 10, 19, 1000):
        for world in (2, 4, 8):
            shares = [class_share(C, r, world) for r in range(world)]
            for r in range(world):
                sched = peer_push_schedule(r, world, shares)
                assert [g for g, _, _ in sched] == [g for g in [(r + s) % world for s in range(1, world)]
                                                    if shares[g][1] > shares[g][0]]
                assert all((lo, hi) == shares[g] for g, lo, hi in sched) and r not in [g for g, _, _ in sched]
            for step in range(1, world):
                dests = [(r + step) % world for r in range(world)]
                assert len(set(dests)) == world


def _closure_worker(rank, world, port, ret):
    """One rank of the pair-sharded closure (model._fused_direct_plan.run_sharded) with the kernels replaced
    by the oracle's float64 arithmetic: loss and dLoss/dF over THIS rank's slice of the linearised pair list
    (p = i (i - 1) / 2 + j, weight 1 / P), ONE all-reduce of [loss, flag, -, -, dF], then the sphere adjoint."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import sqfa_oracle as O
    from sqfa_b200._ops import shard_pairs

    stats, W, noise = _closure_problem()
    C, (k, D) = stats["means"].shape[0], W.shape
    P = C * (C - 1) // 2
    p0, p1 = shard_pairs(P, C, rank, world)
    nrm = W.norm(dim=-1, keepdim=True)
    F = (W / nrm).detach().requires_grad_(True)
    d = O.class_distances_full(stats, F, noise)
    i, j = torch.tril_indices(C, C, offset=-1)
    p = i * (i - 1) // 2 + j
    mine = (p >= p0) & (p < p1)
    loss = -(d[i, j] * mine).sum() / P
    loss.backward()
    packed = torch.cat([loss.detach().view(1), torch.zeros(3, dtype=W.dtype), F.grad.reshape(-1)])
    dist.all_reduce(packed)
    dF = packed[4:].view(k, D)
    Fd = F.detach()
    dW = (dF - (dF * Fd).sum(dim=-1, keepdim=True) * Fd) / nrm
    ret[rank] = (packed[0].clone(), dW.clone(), int(mine.sum()))
    dist.barrier()
    dist.destroy_process_group()


def _closure_problem():
    from oracle import sqfa_oracle as O

    g = torch.Generator().manual_seed(3)
    C, D, k = 13, 10, 3
    X = torch.randn(C * 40, D, generator=g, dtype=torch.float64) * (0.5 + torch.rand(D, generator=g, dtype=torch.float64))
    y = torch.arange(C).repeat_interleave(40)
    X = X + 0.3 * torch.randn(C, D, generator=g, dtype=torch.float64)[y]
    stats = O.class_statistics(X, y)
    W = torch.randn(k, D, generator=g, dtype=torch.float64)
    return stats, W, 0.01


@pytest.mark.parametrize("world", [2, 3])
def test_pair_sharded_closure_sums_to_the_replicated_one(world):
    from oracle import sqfa_oracle as O

    port = 31500 + (os.getpid() % 2000) + world
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_closure_worker, args=(world, port, ret), nprocs=world, join=True)
        got = [ret[r] for r in range(world)]
    stats, W, noise = _closure_problem()
    loss, grad, _ = O.loss_and_grad("full", stats, W, noise=noise, constraint="sphere")
    C = stats["means"].shape[0]
    assert sum(n for _, _, n in got) == C * (C - 1) // 2  # every pair evaluated by exactly one rank
    for l, dW, _ in got:  # identical on every rank (the optimiser state stays replicated)
        assert torch.equal(l, got[0][0]) and torch.equal(dW, got[0][1])
    assert torch.allclose(got[0][0], loss, rtol=1e-12)
    assert torch.allclose(got[0][1], grad, rtol=1e-9, atol=1e-13)


def _three_phase_worker(rank, world, port, ret):
    """One rank of the class- AND pair-sharded closure (sqfa_fused_loss_sharded, _ops.fused_loss_sharded_raw) in
    the oracle's float64 arithmetic: (0) project the rank's classes, all-reduce (Psi, mu') [zeros elsewhere];
    (1) embedding + distances of the rank's pairs, all-reduce (gPsi, gmu, loss); (2) projection adjoint over
    the rank's classes, all-reduce dF."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import sqfa_oracle as O
    from sqfa_b200._ops import shard_pairs
    from sqfa_b200._stats_driver import class_share

    stats, W, noise = _closure_problem()
    S, M = stats["covariances"], stats["means"]
    C, (k, D) = M.shape[0], W.shape
    P = C * (C - 1) // 2
    p0, p1 = shard_pairs(P, C, rank, world)
    c0, c1 = class_share(C, rank, world)
    F = W / W.norm(dim=-1, keepdim=True)
    # phase 0
    psi = torch.zeros(C, k, k, dtype=W.dtype)
    mu = torch.zeros(C, k, dtype=W.dtype)
    T = F @ S[c0:c1]  # saved for the adjoint: (c1 - c0, k, D)
    psi[c0:c1] = T @ F.T
    mu[c0:c1] = M[c0:c1] @ F.T
    dist.all_reduce(psi)
    dist.all_reduce(mu)
    # phase 1
    psi.requires_grad_(True)
    mu.requires_grad_(True)
    cov = psi + noise * torch.eye(k, dtype=W.dtype)
    fs = {"means": mu, "covariances": cov}
    d = O.fisher_rao_lower_bound(fs, fs)
    i, j = torch.tril_indices(C, C, offset=-1)
    p = i * (i - 1) // 2 + j
    mine = (p >= p0) & (p < p1)
    loss = -(d[i, j] * mine).sum() / P
    loss.backward()
    g_psi, g_mu, loss = psi.grad.clone(), mu.grad.clone(), loss.detach().clone()
    for t in (g_psi, g_mu, loss):
        dist.all_reduce(t)
    # phase 2: dF = sum over own classes of (gPsi + gPsi^T) T_c + gmu_c m_c^T
    sym = g_psi[c0:c1] + g_psi[c0:c1].mT
    dF = (sym @ T).sum(0) + g_mu[c0:c1].T @ M[c0:c1]
    dist.all_reduce(dF)
    ret[rank] = (loss, dF.clone())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_class_and_pair_sharded_closure_phases_sum_to_the_replicated_one(world):
    from oracle import sqfa_oracle as O

    port = 33500 + (os.getpid() % 2000) + world
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_three_phase_worker, args=(world, port, ret), nprocs=world, join=True)
        got = [ret[r] for r in range(world)]
    stats, W, noise = _closure_problem()
    F = W / W.norm(dim=-1, keepdim=True)
    loss, dF, _ = O.loss_and_grad("full", stats, F, noise=noise, constraint="none")  # gradient at the constrained filters
    for l, g in got:
        assert torch.equal(l, got[0][0]) and torch.equal(g, got[0][1])
    assert torch.allclose(got[0][0], loss, rtol=1e-12)
    assert torch.allclose(got[0][1], dF, rtol=1e-9, atol=1e-13)
