"""e2e pipelining experiment: where does the time go when uploads/downloads of consecutive steps overlap?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sqfa_b200 import statistics as S, _stats_driver as drv

dev = torch.device("cuda")
n, D, C = 50000, 3072, 10
X = torch.randn(n, D, device=dev)
y = torch.randint(0, C, (n,), device=dev)
Xh, yh = X.cpu().pin_memory(), y.cpu().pin_memory()
st0 = S.class_statistics(X, y)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
out_h = [{k: torch.empty(v.shape).pin_memory() for k, v in st0.items()} for _ in range(2)]
Xd = [torch.empty_like(X) for _ in range(2)]
yd = [torch.empty_like(y) for _ in range(2)]
ops = S._cuda_ops()


def wall(fn, reps=8):
    fn(0); fn(1)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(reps):
        fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


keep = [None, None]


def step(i, two_streams, known_c):
    b = i % 2
    s = streams[b] if two_streams else torch.cuda.current_stream()
    with torch.cuda.stream(s):
        yd[b].copy_(yh, non_blocking=True)
        Xd[b].copy_(Xh, non_blocking=True)
        if known_c:
            m, cov, sm, _ = drv.run_class_statistics(ops, Xd[b], yd[b], 0, n_classes=C)
            st = {"means": m, "covariances": cov, "second_moments": sm}
        else:
            st = S.class_statistics(Xd[b], yd[b])
        for k, v in st.items():
            out_h[b][k].copy_(v, non_blocking=True)
    keep[b] = st


def copies_only(i, both):
    b = i % 2
    with torch.cuda.stream(streams[0]):
        Xd[b].copy_(Xh, non_blocking=True)
    if both:
        with torch.cuda.stream(streams[1]):
            for k, v in st0.items():
                out_h[b][k].copy_(v, non_blocking=True)


print(f"H2D only            {wall(lambda i: copies_only(i, False)):.2f} ms/step")
print(f"H2D || D2H          {wall(lambda i: copies_only(i, True)):.2f} ms/step")
print(f"sequential          {wall(lambda i: step(i, False, False)):.2f} ms/step")
print(f"sequential, known C {wall(lambda i: step(i, False, True)):.2f} ms/step")
print(f"2 streams           {wall(lambda i: step(i, True, False)):.2f} ms/step")
print(f"2 streams, known C  {wall(lambda i: step(i, True, True)):.2f} ms/step")
