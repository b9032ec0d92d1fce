"""Diagnostic: per-epoch loss / subspace angle of the golden converged fit (tests/golden/closure.npz)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import sqfa_oracle as O
from sqfa_b200.model import SQFA, SecondMomentsSQFA
import sqfa_b200._optim as optim_mod

z = np.load(os.path.join(ROOT, "tests", "golden", "closure.npz"))
g = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
stats = {k: g[k].float().cuda() for k in ("means", "covariances", "second_moments")}
for kind, cls in (("sm", SecondMomentsSQFA), ("full", SQFA)):
    for opt_name in ("native", "torch"):
        optim_mod.LBFGS = optim_mod.LBFGS if opt_name == "native" else torch.optim.LBFGS
        for atol in (1e-6, 1e-9):
            m = cls(n_dim=12, feature_noise=0.01, n_filters=3, filters=g["F0"].float())
            losses, _ = m.fit(data_statistics=stats, max_epochs=200, show_progress=False, return_loss=True, atol=atol)
            ang = O.subspace_angle(m.filters.detach().cpu(), g[kind + "_fit_filters"])
            print(kind, opt_name, "atol", atol, "epochs", len(losses), "evals", m._last_fit_evaluations,
                  "final loss %.9f ref %.9f" % (float(losses[-1]), float(g[kind + "_fit_losses"][-1])), "angle %.2e" % ang)
            print("   losses", [round(float(x), 7) for x in losses[-6:]])
