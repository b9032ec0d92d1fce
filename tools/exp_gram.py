"""Experiment: Gram error vs accumulation chain length (ksplit) and first timings. Run on GPU."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sqfa_b200 import _lib, statistics as S
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_class_data

lib = _lib.load()
dev = torch.device("cuda")

def gram_with(X, perm, offsets, means, C, ks):
    n, D = X.shape
    g = torch.empty(C, D, D, device=dev)
    wsb = lib.sqfa_class_gram_workspace_bytes(n, D, C)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr()
    _lib.check(lib.sqfa_class_gram(_lib.ptr(X), X.stride(0), _lib.ptr(perm), _lib.ptr(offsets), _lib.ptr(means),
                                   n, D, C, _lib.ptr(g), 0, ks, None, 0, 0, 0, _lib.ptr(ws), wsb, st), "gram")
    return g

def time_it(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts)//2]

out = {}
# --- error vs chain length
n, D, C = 5000, 3072, 2
X, y = make_class_data(n, D, C, seed=D)
Xd, yd = X.cuda(), y.cuda()
perm, offsets, counts = S.bucket_labels(yd)
means = torch.stack([X[y == c].double().mean(0) for c in range(C)]).float().cuda()
ref = torch.stack([(X[y == c].double() - means[c].double().cpu()).T @ (X[y == c].double() - means[c].double().cpu()) for c in range(C)])
iu = torch.triu_indices(D, D)
for ks in (4096, 2048, 1024, 512, 256, 128):
    g = gram_with(Xd, perm, offsets, means, C, ks).double().cpu()
    e = ((g - ref)[:, iu[0], iu[1]].norm() / ref[:, iu[0], iu[1]].norm()).item()
    dg = torch.diagonal(g, dim1=1, dim2=2); dr = torch.diagonal(ref, dim1=1, dim2=2)
    bias = ((dg - dr) / dr).mean().item()
    out[f"err_ks{ks}"] = (e, bias)
    print("ks", ks, "relerr", e, "mean rel diag bias", bias, flush=True)

# --- timing at the c2 bench shape
n, D, C = 50000, 3072, 10
g = torch.Generator(device="cuda").manual_seed(0)
Xb = torch.randn(n, D, device=dev, generator=g)
yb = torch.randint(0, C, (n,), device=dev, generator=g)
perm, offsets, counts = S.bucket_labels(yb)
means = torch.zeros(C, D, device=dev)
for ks in (2048, 1024, 512, 256):
    tmin, tmed = time_it(lambda: gram_with(Xb, perm, offsets, means, C, ks))
    tz = 0.0
    flops = 3 * 2 * n * D * D * (156 / 288)  # executed: 3 passes, 156 of 288 128x256 tiles
    print(f"gram ks={ks}: {tmin:.3f} ms (chain_rows=ks; excl {tz:.3f} ms) -> {flops / ((tmin - tz) * 1e-3) / 1e12:.1f} TF/s executed tf32", flush=True)
    out[f"gram_ms_ks{ks}"] = (tmin, tz)
tmin, tmed = time_it(lambda: S.class_statistics(Xb, yb))
print(f"class_statistics c2: {tmin:.3f} ms min {tmed:.3f} med -> {n / (tmin * 1e-3):.3e} samples/s", flush=True)
out["class_statistics_c2_ms"] = (tmin, tmed)
tb, _ = time_it(lambda: S.bucket_labels(yb))
print(f"bucket: {tb:.3f} ms")
n, D, C = 60000, 784, 10
Xb = torch.randn(n, D, device=dev, generator=g); yb = torch.randint(0, C, (n,), device=dev, generator=g)
tmin, tmed = time_it(lambda: S.class_statistics(Xb, yb))
print(f"class_statistics c1: {tmin:.3f} ms -> {n / (tmin * 1e-3):.3e} samples/s", flush=True)
n, D, C = 1280000, 512, 1000
Xb = torch.randn(n, D, device=dev, generator=g); yb = torch.randint(0, C, (n,), device=dev, generator=g)
tmin, tmed = time_it(lambda: S.class_statistics(Xb, yb))
print(f"class_statistics c4: {tmin:.3f} ms -> {n / (tmin * 1e-3):.3e} samples/s", flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "exp_gram.json"), "w"))
