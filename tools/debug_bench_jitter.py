"""Per-step timings of the bench loop under torchrun with toggles (debug aid)."""
import gc, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
from sqfa_b200 import statistics as S

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
w = bench.WORKLOAD
X, y = bench.synth(w["N"], w["D"], w["C"], dev, 1234 + rank)
group = dist.group.WORLD


def run(tag, hold, sampler_on, gc_off, synthetic_labels=False):
    yy = torch.randint(0, w["C"], (w["N"],), device=dev) if synthetic_labels else y
    for _ in range(3):
        st = S.class_statistics(X, yy, group=group)
        if not hold:
            del st
    if gc_off:
        gc.collect(); gc.disable()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    smp = bench.ClockSampler(local) if (sampler_on and rank == 0) else None
    if smp:
        smp.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    evs[0].record()
    for i in range(10):
        st = S.class_statistics(X, yy, group=group)
        if not hold:
            del st
        evs[i + 1].record()
    torch.cuda.synchronize()
    if smp:
        smp.stop()
    gc.enable()
    t = [evs[i].elapsed_time(evs[i + 1]) for i in range(10)]
    print(f"rank {rank} {tag}: total {sum(t):.2f} ms  steps " + " ".join(f"{x:.2f}" for x in t), flush=True)
    dist.barrier()


run("plain", False, False, False)
run("hold", True, False, False)
run("hold+gcoff", True, False, True)
run("hold+gcoff+sampler", True, True, True)
run("plain again", False, False, False)
run("randlabels", False, False, False, True)
dist.destroy_process_group()
