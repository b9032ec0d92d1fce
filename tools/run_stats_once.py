"""One class_statistics call at a BASELINE config (for ncu captures). Usage: run_stats_once.py [c1|c2|c4] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sqfa_b200 import statistics as S

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n, D, C = {"c1": (60000, 784, 10), "c2": (50000, 3072, 10), "c3": (200000, 104, 19), "c4": (1280000, 512, 1000)}[cfg]
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(n, D, device="cuda", generator=g)
y = torch.randint(0, C, (n,), device="cuda", generator=g)
for _ in range(reps):
    s = S.class_statistics(X, y)
torch.cuda.synchronize()
print("ok", float(s["covariances"][0, 0, 0]))
