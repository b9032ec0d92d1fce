"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the two places where the hot paths shard.

* class_statistics with the samples sharded over the ranks == the single-GPU result
* the closure with the class-pair list (and the projection's classes) sharded over the ranks == the
  replicated closure: loss and gradient to 1e-6, and a short fit produces the same losses
Each rank is a spawned process (`torch.multiprocessing`), rendezvous on 127.0.0.1.
"""

import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch.distributed as dist

    from conftest import make_class_data
    from sqfa_b200 import statistics as S
    from sqfa_b200.model import SQFA

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        res = {}
        # ---- HP1: samples sharded (ragged split), every rank gets the statistics of the union
        n, d, c = 9001, 300, 12
        X, y = make_class_data(n, d, c, seed=3)
        X = X / (X.std() * d**0.5)
        cut = [0, 4000, n]
        full = S.class_statistics(X.cuda(), y.cuda())
        part = S.class_statistics(X[cut[rank]:cut[rank + 1]].cuda(), y[cut[rank]:cut[rank + 1]].cuda(),
                                  group=dist.group.WORLD)
        for key in full:
            res["stats_" + key] = float((part[key] - full[key]).norm() / full[key].norm())
        # a class absent from one rank's shard and OAS
        y2 = y.clone()
        y2[: cut[1]][y2[: cut[1]] == 5] = 4
        full2 = S.class_statistics(X.cuda(), y2.cuda(), estimator="oas")
        part2 = S.class_statistics(X[cut[rank]:cut[rank + 1]].cuda(), y2[cut[rank]:cut[rank + 1]].cuda(),
                                   estimator="oas", group=dist.group.WORLD)
        res["stats_oas_absent_class"] = float((part2["covariances"] - full2["covariances"]).norm()
                                              / full2["covariances"].norm())
        # ---- HP1, sharded output: reduce-scatter by class fused into the Gram kernel (copy-engine pushes into
        # peer-mapped slots) and the epilogue; against the single-GPU result, against the NCCL all-reduce
        # path, bit-reproducible, over several calls (slot sets alternate) and from two streams
        n, d, c = 9001, 512, 12  # D = 512: the packed tile list is smaller than (C, D, D), the layout both paths move
        X, y = make_class_data(n, d, c, seed=4)
        X = X / (X.std() * d**0.5)
        full = S.class_statistics(X.cuda(), y.cuda())
        Xs, ys = X[cut[rank]:cut[rank + 1]].cuda(), y[cut[rank]:cut[rank + 1]].cuda()
        sh = S.class_statistics(Xs, ys, group=dist.group.WORLD, shard_output=True)
        lo, hi = sh["class_range"]
        ops = S._cuda_ops()
        states = getattr(ops, "_peer_states", {})
        res["peer_path_used"] = float(any(st.get("ok") for st in states.values()))
        res["peer_why"] = ";".join(st.get("why", "") for st in states.values())
        res["peer_means"] = float((sh["means"] - full["means"]).norm() / full["means"].norm())
        for key in ("covariances", "second_moments"):
            res["peer_" + key] = float((sh[key] - full[key][lo:hi]).norm() / full[key][lo:hi].norm())
        again = [S.class_statistics(Xs, ys, group=dist.group.WORLD, shard_output=True) for _ in range(3)]
        res["peer_reproducible"] = float(all(torch.equal(a["covariances"], sh["covariances"]) for a in again))
        other = torch.cuda.Stream()
        other.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(other):
            on_other = S.class_statistics(Xs, ys, group=dist.group.WORLD, shard_output=True)
        on_main = S.class_statistics(Xs, ys, estimator="oas", group=dist.group.WORLD, shard_output=True)
        torch.cuda.synchronize()
        res["peer_two_streams"] = float(torch.equal(on_other["covariances"], sh["covariances"]))
        full_oas = S.class_statistics(X.cuda(), y.cuda(), estimator="oas")
        res["peer_oas"] = float((on_main["covariances"] - full_oas["covariances"][lo:hi]).norm()
                                / full_oas["covariances"][lo:hi].norm())
        os.environ["SQFA_PEER_REDUCE"] = "0"
        nccl = S.class_statistics(Xs, ys, group=dist.group.WORLD, shard_output=True)
        os.environ["SQFA_PEER_REDUCE"] = "1"
        res["peer_vs_nccl"] = float((nccl["covariances"] - sh["covariances"]).norm() / sh["covariances"].norm())
        # fewer classes than ranks would own (C = 1 on 2 ranks: rank 1 owns nothing)
        y1 = torch.zeros_like(ys)
        one = S.class_statistics(Xs, y1, group=dist.group.WORLD, shard_output=True)
        full1 = S.class_statistics(X.cuda(), torch.zeros_like(y).cuda())
        lo1, hi1 = one["class_range"]
        res["peer_one_class_rows"] = float(one["covariances"].shape[0] == hi1 - lo1)
        if hi1 > lo1:
            res["peer_one_class"] = float((one["covariances"] - full1["covariances"][lo1:hi1]).norm()
                                          / full1["covariances"][lo1:hi1].norm())
        # ---- HP2: pair list sharded. C = 40 -> 780 pairs
        n, d, c, k = 12000, 64, 40, 6
        X, y = make_class_data(n, d, c, seed=9)
        X = X / (X.std() * d**0.5)
        stats = S.class_statistics(X.cuda(), y.cuda())
        F0 = torch.randn(k, d, generator=torch.Generator().manual_seed(0))
        model = SQFA(n_dim=d, feature_noise=0.01, n_filters=k, filters=F0.clone()).cuda()
        rep = model._fused_direct_plan(stats)().clone()
        g_rep = model.parametrizations.filters.original.grad.clone()
        model._process_group = dist.group.WORLD
        sh = model._fused_direct_plan(stats)().clone()
        g_sh = model.parametrizations.filters.original.grad.clone()
        # the same with only the pairs sharded (every rank projects all classes)
        os.environ["SQFA_SHARD_CLASSES"] = "0"
        sh_p = model._fused_direct_plan(stats)().clone()
        g_sh_p = model.parametrizations.filters.original.grad.clone()
        os.environ["SQFA_SHARD_CLASSES"] = "1"
        model._process_group = None
        res["closure_pairs_only_loss_rel"] = abs(float(sh_p[0]) - float(rep[0])) / abs(float(rep[0]))
        res["closure_pairs_only_grad_rel"] = float((g_sh_p - g_rep).norm() / g_rep.norm())
        res["closure_loss_rel"] = abs(float(sh[0]) - float(rep[0])) / abs(float(rep[0]))
        res["closure_grad_rel"] = float((g_sh - g_rep).norm() / g_rep.norm())
        res["closure_bad"] = float(sh[1])
        m1 = SQFA(n_dim=d, feature_noise=0.01, n_filters=k, filters=F0.clone())
        l1, _ = m1.fit(data_statistics=stats, max_epochs=2, show_progress=False, return_loss=True, max_iter=6)
        m2 = SQFA(n_dim=d, feature_noise=0.01, n_filters=k, filters=F0.clone())
        l2, _ = m2.fit(data_statistics=stats, max_epochs=2, show_progress=False, return_loss=True, max_iter=6,
                       process_group=dist.group.WORLD)
        res["fit_loss_rel"] = float(((l1 - l2).abs() / l1.abs()).max())
        res["fit_filter_rel"] = float((m1.filters.detach() - m2.filters.detach()).norm() / m1.filters.detach().norm())
        # every rank must hold the same filters (identical all-reduced gradients)
        Fl = m2.filters.detach().cuda()
        Fo = Fl.clone()
        dist.broadcast(Fo, src=0)
        res["fit_rank_divergence"] = float((Fl - Fo).abs().max())
        torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < WORLD, reason="needs 2 GPUs")
def test_sharded_statistics_and_pair_sharded_closure(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(WORLD, port, str(tmp_path)), nprocs=WORLD, join=True)
    for rank in range(WORLD):
        res = torch.load(os.path.join(str(tmp_path), f"rank{rank}.pt"))
        print(rank, res)
        for key in ("stats_means", "stats_covariances", "stats_second_moments", "stats_oas_absent_class"):
            assert res[key] < 1e-5, (rank, key, res[key])
        assert res["peer_path_used"] == 1.0, res["peer_why"]
        for key in ("peer_means", "peer_covariances", "peer_second_moments", "peer_oas", "peer_vs_nccl"):
            assert res[key] < 1e-5, (rank, key, res[key])
        assert res["peer_reproducible"] == 1.0 and res["peer_two_streams"] == 1.0
        assert res["peer_one_class_rows"] == 1.0 and res.get("peer_one_class", 0.0) < 1e-5
        assert res["closure_bad"] == 0
        assert res["closure_loss_rel"] < 1e-6 and res["closure_grad_rel"] < 1e-5, res
        assert res["closure_pairs_only_loss_rel"] < 1e-6 and res["closure_pairs_only_grad_rel"] < 1e-5, res
        assert res["fit_loss_rel"] < 1e-4 and res["fit_filter_rel"] < 1e-3, res
        assert res["fit_rank_divergence"] == 0.0
