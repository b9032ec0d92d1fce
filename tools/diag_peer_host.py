"""How much of the multi-GPU class_statistics step is the host (the label-maximum read and the Python between
it and the Gram launch)? Sharded-output step with and without a known n_classes. Run under torchrun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from sqfa_b200 import statistics as S
from sqfa_b200._stats_driver import run_class_statistics

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, D, C = 50000, 3072, 10
g = torch.Generator(device=dev).manual_seed(rank)
X = torch.randn(n, D, device=dev, generator=g)
y = torch.randint(0, C, (n,), device=dev, generator=g)
ops = S._cuda_ops()
G = dist.group.WORLD


def timeit(fn, reps=10):
    for _ in range(4):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t), host * 1e3


a = timeit(lambda: run_class_statistics(ops, X, y, 0, group=G, shard_output=True))
b = timeit(lambda: run_class_statistics(ops, X, y, 0, group=G, shard_output=True, n_classes=C))
c = timeit(lambda: run_class_statistics(ops, X, y, 0))
if rank == 0:
    print(f"world {dist.get_world_size()}: sharded step {a[0]:.3f} ms (host enqueue {a[1]:.3f}) | with n_classes given "
          f"{b[0]:.3f} ms (host {b[1]:.3f}) | single-GPU fused {c[0]:.3f} ms (host {c[1]:.3f})")
dist.destroy_process_group()
