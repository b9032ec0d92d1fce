"""Diagnostic: where does a class_statistics call spend host / device time? Usage: diag_stats_host.py c3"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sqfa_b200 import statistics as S

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
n, D, C = {"c1": (60000, 784, 10), "c2": (50000, 3072, 10), "c3": (200000, 104, 19), "c4": (1280000, 512, 1000)}[cfg]
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(n, D, device="cuda", generator=g)
y = torch.randint(0, C, (n,), device="cuda", generator=g)
ops = S._cuda_ops()
for _ in range(3):
    S.class_statistics(X, y)
torch.cuda.synchronize()
def tm(f, reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = f()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / reps * 1e3, (t2 - t0) / reps * 1e3
print(cfg, "class_statistics     host-enqueue %.3f ms, with final sync %.3f ms per call" % tm(lambda: S.class_statistics(X, y)))
print(cfg, "label_max_host       %.3f / %.3f" % tm(lambda: ops.label_max_host(y)))
print(cfg, "fused (C known)      %.3f / %.3f" % tm(lambda: ops.fused(X, y, C, 0, 1, True)))
print(cfg, "bucket               %.3f / %.3f" % tm(lambda: ops.bucket(y, C)))
perm, offsets, counts = ops.bucket(y, C)
print(cfg, "class_sums           %.3f / %.3f" % tm(lambda: ops.class_sums(X, perm, offsets, C)))
sums = ops.class_sums(X, perm, offsets, C)
means = ops.class_means(sums, counts[:C].clone())
print(cfg, "class_gram           %.3f / %.3f" % tm(lambda: ops.class_gram(X, perm, offsets, means, C)))
gram = ops.class_gram(X, perm, offsets, means, C)
print(cfg, "finalize             %.3f / %.3f" % tm(lambda: ops.finalize(gram.clone(), means, counts[:C].clone(), 0, 1, True)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): S.class_statistics(X, y)
e1.record(); torch.cuda.synchronize()
print(cfg, "event-timed per call %.3f ms" % (e0.elapsed_time(e1) / 20))
