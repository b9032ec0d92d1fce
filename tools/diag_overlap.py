"""Diagnostic (torchrun, >= 2 ranks): timeline of the overlapped Gram all-reduce at the c2 shape.
Knobs: SQFA_GRAM_GROUPS, SQFA_GRAM_RESERVE_SMS, SQFA_NCCL_CTAS, SQFA_GRAM_OVERLAP."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from sqfa_b200 import statistics as S

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, D, C = (int(v) for v in os.environ.get("SQFA_DIAG_SHAPE", "50000,3072,10").split(","))
g = torch.Generator(device=dev).manual_seed(rank)
X = torch.randn(n, D, device=dev, generator=g)
y = torch.randint(0, C, (n,), device=dev, generator=g)
ops = S._cuda_ops()
for _ in range(4):
    S.class_statistics(X, y, group=dist.group.WORLD)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    S.class_statistics(X, y, group=dist.group.WORLD)
e1.record(); torch.cuda.synchronize()
step_ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
ops.overlap_trace = []
for _ in range(3):
    S.class_statistics(X, y, group=dist.group.WORLD)
torch.cuda.synchronize()
if rank == 0:
    cfg = {k: os.environ.get(k) for k in ("SQFA_GRAM_GROUPS", "SQFA_GRAM_RESERVE_SMS", "SQFA_NCCL_CTAS", "SQFA_GRAM_OVERLAP")}
    print("world", world, cfg, "step %.3f ms" % float(step_ms))
    for t0, t1, marks in ops.overlap_trace[-1:]:
        print("   gram kernel end %.3f ms; all-reduce group ends" % t0.elapsed_time(t1),
              ["%.3f" % t0.elapsed_time(m) for m in marks])
dist.barrier(); dist.destroy_process_group()
