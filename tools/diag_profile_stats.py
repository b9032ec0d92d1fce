"""Diagnostic: cProfile of the host side of class_statistics at a BASELINE config."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sqfa_b200 import statistics as S
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
n, D, C = {"c1": (60000, 784, 10), "c2": (50000, 3072, 10), "c3": (200000, 104, 19), "c4": (1280000, 512, 1000)}[cfg]
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(n, D, device="cuda", generator=g)
y = torch.randint(0, C, (n,), device="cuda", generator=g)
for _ in range(5):
    S.class_statistics(X, y)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    S.class_statistics(X, y)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
