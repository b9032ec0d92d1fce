"""Multi-GPU class_statistics step at c2 (50 000 x 3072 per rank): replicated output (NCCL all-reduce of the
packed Gram) against sharded output through NCCL and through the fused peer reduce-scatter.
Run under torchrun with >= 2 ranks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from sqfa_b200 import statistics as S

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, D, C = 50000, 3072, 10
g = torch.Generator(device=dev).manual_seed(rank)
X = torch.randn(n, D, device=dev, generator=g)
y = torch.randint(0, C, (n,), device=dev, generator=g)


def timeit(fn, reps=10):
    for _ in range(4):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


W = dist.get_world_size()
t_single = timeit(lambda: S.class_statistics(X, y))
t_rep = timeit(lambda: S.class_statistics(X, y, group=dist.group.WORLD))
os.environ["SQFA_PEER_REDUCE"] = "0"
t_sh_nccl = timeit(lambda: S.class_statistics(X, y, group=dist.group.WORLD, shard_output=True))
os.environ["SQFA_PEER_REDUCE"] = "1"
t_sh_peer = timeit(lambda: S.class_statistics(X, y, group=dist.group.WORLD, shard_output=True))
ops = S._cuda_ops()
used = [st.get("ok") for st in getattr(ops, "_peer_states", {}).values()]
why = [st.get("why") for st in getattr(ops, "_peer_states", {}).values()]
if rank == 0:
    print(f"world {W}: single-GPU call {t_single:.3f} ms | replicated (NCCL all-reduce) {t_rep:.3f} | sharded NCCL "
          f"{t_sh_nccl:.3f} | sharded peer push {t_sh_peer:.3f} (peer path ok: {used} {why}) | "
          f"efficiency vs single: {t_single / t_rep:.3f} / {t_single / t_sh_nccl:.3f} / {t_single / t_sh_peer:.3f}")
dist.destroy_process_group()
