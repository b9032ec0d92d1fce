"""CPU, world_size 2, gloo: the sample-sharded class_statistics driver (three all-reduces) with
the local kernels replaced by oracle-backed CPU ops. Checks that sharded == unsharded."""

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
