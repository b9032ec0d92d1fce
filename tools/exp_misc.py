"""Timings that are not bench lines: transform (Z = X F^T) bandwidth, c5-shard class_statistics."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sqfa_b200 import _ops, statistics as S


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n, D, k in [(50000, 3072, 8), (60000, 784, 4), (200000, 104, 8), (1280000, 512, 16), (1000000, 1024, 32)]:
    X = torch.randn(n, D, device="cuda")
    F = torch.randn(k, D, device="cuda")
    ms = timeit(lambda: _ops.transform_raw(X, F))
    print(f"transform n={n} D={D} k={k}: {ms*1e3:.1f} us -> {n*D*4/ms/1e6:.0f} GB/s")
    del X
if "c5" in sys.argv:
    n, d, c = 12_500_000, 1024, 100
    y = torch.randint(0, c, (n,), device="cuda")
    X = torch.empty(n, d, device="cuda")
    for lo in range(0, n, 500_000):
        X[lo:lo + 500_000].normal_()
    ms = timeit(lambda: S.class_statistics(X, y), reps=3)
    print(f"class_statistics c5 shard (12.5M x 1024, C=100): {ms:.1f} ms -> {n/ms/1e3:.1f} M samples/s, "
          f"{2*n*d*d/ms/1e9:.0f} algorithmic TFLOP/s")
