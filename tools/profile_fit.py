"""Where does the time of SQFA.fit go? Per-kernel-stage timings of one closure + per-epoch wall times."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sqfa_b200 import _ops
from sqfa_b200.model import SQFA, SecondMomentsSQFA
import sqfa_b200._optim as optim_mod

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
C, D, k = {"c1": (10, 784, 4), "c2": (10, 3072, 8), "c3": (19, 104, 8), "c4": (1000, 512, 16), "c5": (100, 1024, 32)}[cfg]
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(C, D, D + 8, device=dev, generator=g)
cov = (A @ A.transpose(1, 2)) / (D + 8) / D
del A
means = 0.05 * torch.randn(C, D, device=dev, generator=g) / D**0.5
stats = {"means": means, "covariances": cov.contiguous()}
model = SQFA(n_dim=D, feature_noise=0.01, n_filters=k).to(dev)

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

F = model.filters.detach().contiguous()
for rep in range(2):
    t = [ev()]
    T, Psi, Mu = _ops.project_fwd_raw(stats["covariances"], stats["means"], F); t.append(ev())
    E = _ops.embed_fwd_raw(Psi, Mu, 0.01, _ops.DIST_FR); t.append(ev())
    W, flag = _ops.class_factor_raw(E, _ops.DIST_FR); t.append(ev())
    m = E.shape[-1]; P = C * (C - 1) // 2
    out = torch.zeros(2, device=dev); gE = torch.zeros(C, m, m, device=dev); t.append(ev())
    _ops.pair_raw(W, W, C, C, m, _ops.DIST_FR, True, weight=-1.0 / P, loss=out, gEa=gE, gEb=gE); t.append(ev())
    gPsi, gMu = _ops.embed_bwd_raw(gE, Mu, k, _ops.DIST_FR); t.append(ev())
    dF = _ops.project_bwd_raw(gPsi, gMu, T, stats["means"]); t.append(ev())
    torch.cuda.synchronize()
names = ["project_fwd", "embed_fwd", "class_factor", "zeros", "pair_fwd_bwd", "embed_bwd", "project_bwd"]
print(cfg, "C,D,k,m,P =", C, D, k, m, P)
for nm, a, b in zip(names, t, t[1:]):
    print(f"  {nm:14s} {a.elapsed_time(b)*1e3:9.1f} us")
print(f"  total          {t[0].elapsed_time(t[-1])*1e3:9.1f} us   (S read at HBM peak would take {4*C*D*D/6.5e12*1e6:.1f} us)")

plan = model._fused_loss_plan(stats)
for _ in range(3):
    model.zero_grad(); plan()[0].backward()
torch.cuda.synchronize()
a = ev()
for _ in range(20):
    model.zero_grad(); plan()[0].backward()
b = ev(); torch.cuda.synchronize()
print(f"closure (no host sync) {a.elapsed_time(b)/20*1e3:.1f} us")
t0 = time.perf_counter()
for _ in range(20):
    model.zero_grad(); o = plan(); v = o.detach().tolist(); o[0].backward()
torch.cuda.synchronize()
print(f"closure (with host read) {(time.perf_counter()-t0)/20*1e6:.1f} us")

# per-epoch wall time and closure count through the real fit
count = [0]
orig = _ops.FusedLoss.apply
def counting(*a, **kw):
    count[0] += 1
    return orig(*a, **kw)
_ops.FusedLoss.apply = counting
epochs = 4 if C < 500 else 2
import cProfile, pstats, io
import copy
warm = copy.deepcopy(model)
warm.fit(data_statistics=stats, max_epochs=1, atol=0.0, show_progress=False)  # imports, allocator, autotune
torch.cuda.synchronize()
count[0] = 0
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
loss, tt = model.fit(data_statistics=stats, max_epochs=epochs, atol=0.0, show_progress=False, return_loss=True)
pr.disable()
torch.cuda.synchronize()
sio = io.StringIO(); pstats.Stats(pr, stream=sio).sort_stats("tottime").print_stats(22); print(sio.getvalue()[-3600:])
print(f"fit: {epochs} epochs, {count[0]} closure evals, {time.perf_counter()-t0:.3f} s; epoch end times {tt.tolist()}")
print("losses", loss.tolist())
