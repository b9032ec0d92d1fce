"""Pair kernel timing: forward only vs forward+backward, for a few (C, m)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sqfa_b200 import _ops

dev = torch.device("cuda")
def spd(n, m, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    a = torch.randn(n, m, m + 4, generator=g, device=dev)
    return (a @ a.transpose(1, 2) / (m + 4) + 0.05 * torch.eye(m, device=dev)).contiguous()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
for C, m, dist in [(1000, 17, _ops.DIST_FR), (1000, 16, _ops.DIST_AI), (1000, 9, _ops.DIST_FR), (300, 33, _ops.DIST_FR), (1000, 17, _ops.DIST_LE)]:
    E = spd(C, m, m)
    W, _ = _ops.class_factor_raw(E, dist)
    P = C * (C - 1) // 2
    out = torch.zeros(2, device=dev); gE = torch.zeros(C, m, m, device=dev)
    tf = t(lambda: _ops.pair_raw(W, W, C, C, m, dist, True, weight=-1.0 / P, loss=out))
    tb = t(lambda: _ops.pair_raw(W, W, C, C, m, dist, True, weight=-1.0 / P, loss=out, gEa=gE, gEb=gE))
    print(f"C={C} m={m} dist={dist}: P={P} fwd {tf:.3f} ms ({P/tf/1e3:.1f} Mpairs/s)  fwd+bwd {tb:.3f} ms ({P/tb/1e3:.1f} Mpairs/s)")
