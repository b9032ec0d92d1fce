"""A few closure evaluations (the fitting loop's graph-free path) at a BASELINE config, for ncu launch
lists / captures. Usage: run_closure_once.py c2 [reps] [graph]   (graph: allow CUDA-graph replay)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if not (len(sys.argv) > 3 and sys.argv[3] == "graph"):
    os.environ["SQFA_GRAPH_CLOSURE"] = "0"  # individual kernel launches are what a profiler should see
import torch
from sqfa_b200.model import SQFA

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
C, D, k = {"c1": (10, 784, 4), "c2": (10, 3072, 8), "c3": (19, 104, 8), "c4": (1000, 512, 16), "c5": (100, 1024, 32)}[cfg]
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(C, D, D + 8, device=dev, generator=g)
cov = ((A @ A.transpose(1, 2)) / (D + 8) / D).contiguous()
del A
means = 0.05 * torch.randn(C, D, device=dev, generator=g) / D**0.5
stats = {"means": means, "covariances": cov}
model = SQFA(n_dim=D, feature_noise=0.01, n_filters=k).to(dev)
plan = model._fused_direct_plan(stats)
for _ in range(reps):
    out = plan()
torch.cuda.synchronize()
print("ok", out.tolist())
