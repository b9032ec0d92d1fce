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

    def class_sums(self, X, perm, offsets, C):
        return torch.stack([X[perm[offsets[c]:offsets[c + 1]].long()].sum(0) for c in range(C)])

    def class_means(self, sums, counts):
        return sums / counts[:, None].to(sums.dtype)

    def class_gram(self, X, perm, offsets, centre, C):
        out = []
        for c in range(C):
            Z = X[perm[offsets[c]:offsets[c + 1]].long()] - centre[c]
            out.append(Z.T @ Z)
        return torch.stack(out)

    def finalize(self, gram, means, counts, estimator_id, ddof, want_sm):
        cov = gram / (counts[:, None, None].to(gram.dtype) - ddof)
        return cov, cov + means[:, :, None] * means[:, None, :]


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
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_statistics_equal_unsharded():
    from oracle import sqfa_oracle as O

    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        means, cov, sm = ret["means"], ret["cov"], ret["sm"]
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
