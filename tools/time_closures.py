"""Closure evaluation time (graph-free path, CUDA-graph replay allowed) at the BASELINE shapes.
Usage: time_closures.py [c1 c2 ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sqfa_b200.model import SQFA, SecondMomentsSQFA

SHAPES = {"c1": (10, 784, 4), "c2": (10, 3072, 8), "c3": (19, 104, 8), "c4": (1000, 512, 16), "c5": (100, 1024, 32)}
dev = torch.device("cuda")
for cfg in (sys.argv[1:] or list(SHAPES)):
    C, D, k = SHAPES[cfg]
    g = torch.Generator(device=dev).manual_seed(0)
    A = torch.randn(C, D, D + 8, device=dev, generator=g)
    cov = ((A @ A.transpose(1, 2)) / (D + 8) / D).contiguous()
    del A
    means = 0.05 * torch.randn(C, D, device=dev, generator=g) / D**0.5
    stats = {"means": means, "covariances": cov}
    model = SQFA(n_dim=D, feature_noise=0.01, n_filters=k).to(dev)
    plan = model._fused_direct_plan(stats)
    for _ in range(5):
        plan()
    torch.cuda.synchronize()
    reps = 200 if C < 500 else 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = plan()
    e1.record()
    torch.cuda.synchronize()
    print(f"{cfg}: C={C} D={D} k={k}  closure {e0.elapsed_time(e1) / reps * 1e3:.1f} us  loss {float(out[0]):.6f}")
    del cov, stats, model, plan
    torch.cuda.empty_cache()
