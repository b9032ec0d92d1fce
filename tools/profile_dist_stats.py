"""Where the multi-GPU class_statistics step spends its time (run under torchrun, 2+ ranks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from sqfa_b200 import statistics as S, _stats_driver as drv

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, D, C = 50000, 3072, 10
g = torch.Generator(device=dev).manual_seed(rank)
X = torch.randn(n, D, device=dev, generator=g)
y = torch.randint(0, C, (n,), device=dev, generator=g)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_full = timeit(lambda: S.class_statistics(X, y, group=dist.group.WORLD))
t_single = timeit(lambda: S.class_statistics(X, y))
ops = S._cuda_ops()
buf = torch.empty(ops.lib.sqfa_gram_packed_floats(D, C), device=dev)
t_ar = timeit(lambda: dist.all_reduce(buf))
small = torch.empty(C * D, device=dev)
t_small = timeit(lambda: dist.all_reduce(small))
orig = drv._all_reduce
drv._all_reduce = lambda t, group, op=None: None
t_nocomm = timeit(lambda: S.class_statistics(X, y, group=dist.group.WORLD))
drv._all_reduce = orig
if rank == 0:
    print(f"world {dist.get_world_size()}: full {t_full:.3f} ms | single-GPU path {t_single:.3f} | stepwise without collectives "
          f"{t_nocomm:.3f} | all_reduce packed gram ({buf.numel() * 4 / 1e6:.0f} MB) {t_ar:.3f} | small all_reduce {t_small:.3f}")
dist.destroy_process_group()
