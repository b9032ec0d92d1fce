"""GPU: float64 inputs. The reference follows the dtype of its inputs
(statistics.py:28, model `.double()`), recommends float64 when distances turn NaN / inf
(_optim.py:28-30) and runs its own test-suite in float64; the goldens are float64, so parity is checked
at float64 tolerances: class statistics through the native DFMA kernels, the closure and a fit through
float64 device-side library calls."""

import os

import numpy as np
import pytest
import torch

from conftest import make_class_data, rel_err
from oracle import sqfa_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with np.load(os.path.join(GOLD, name)) as z:
        return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def test_class_statistics_float64_golden():
    from sqfa_b200 import statistics as S

    g = load("class_statistics.npz")
    X, y = g["X"].cuda(), g["y"].cuda()
    assert X.dtype == torch.float64
    emp = S.class_statistics(X, y)
    for key in ("means", "covariances", "second_moments"):
        assert emp[key].dtype == torch.float64 and emp[key].is_cuda
        assert rel_err(emp[key], g[key]) < 1e-12, key
    oas = S.class_statistics(X, y, estimator="oas")
    assert rel_err(oas["covariances"], g["oas_covariances"]) < 1e-12
    assert rel_err(oas["second_moments"], g["oas_second_moments"]) < 1e-12
    assert rel_err(S.sample_covariance(X), g["sample_cov"]) < 1e-12
    assert rel_err(S.sample_covariance(X, assume_centered=True), g["sample_cov_centered"]) < 1e-12
    assert O.subspace_angle(S.pca(X, 3).cpu(), g["pca3"]) < 1e-7
    cpu = S.class_statistics(g["X"], g["y"])  # CPU tensors in, CPU float64 tensors out
    assert cpu["covariances"].dtype == torch.float64 and not cpu["covariances"].is_cuda
    assert torch.equal(cpu["covariances"], emp["covariances"].cpu())  # deterministic


@pytest.mark.parametrize("n,d,c,kw", [(3000, 104, 19, {}), (2500, 300, 5, {"offset": 3.0}), (4000, 130, 7, {"skew": True}),
                                      (600, 65, 3, {})])
def test_class_statistics_float64_matches_oracle(n, d, c, kw):
    from sqfa_b200.statistics import class_statistics

    X, y = make_class_data(n, d, c, seed=d, dtype=torch.float64, **kw)
    y[y == c - 1] = 0 if c == 3 else y[y == c - 1]  # c == 3: last class empty -> NaN statistics like the reference
    got = class_statistics(X.cuda(), y.cuda())
    ref = O.class_statistics(X, y)
    for key in ("means", "covariances", "second_moments"):
        a, b = got[key].cpu(), ref[key]
        assert torch.equal(torch.isnan(a), torch.isnan(b)), key
        ok = ~torch.isnan(b)
        assert float((a[ok] - b[ok]).norm() / b[ok].norm()) < 1e-11, key


@pytest.mark.parametrize("tag,kind,dist", [("sm", "sm", None), ("sm_le", "sm", "log_euclidean"), ("full", "full", None)])
def test_closure_and_fit_float64_golden(tag, kind, dist):
    """model.double() + float64 statistics: loss / gradient of the reference to 1e-9, its converged fit
    to 1e-5 in subspace angle."""
    from sqfa_b200 import distances as Dn
    from sqfa_b200.model import SQFA, SecondMomentsSQFA

    g = load("closure.npz")
    stats = {k: g[k].cuda() for k in ("means", "covariances", "second_moments")}
    cls = SecondMomentsSQFA if kind == "sm" else SQFA
    dfun = getattr(Dn, dist) if dist else None
    m = cls(n_dim=12, feature_noise=0.01, n_filters=3, filters=g["F0"].float(), distance_fun=dfun).double().cuda()
    assert m._fused_loss_plan(stats) is None  # float64 takes the composed path
    dmat = m.get_class_distances(stats, regularized=True)
    assert dmat.dtype == torch.float64
    i, j = torch.tril_indices(4, 4, -1)
    assert rel_err(dmat[i, j], g[tag + "_dist"][i, j]) < 1e-9
    loss = -dmat[i, j].mean()
    loss.backward()
    assert abs(float(loss) - float(g[tag + "_loss"])) < 1e-9 * abs(float(g[tag + "_loss"]))
    assert rel_err(m.parametrizations.filters.original.grad, g[tag + "_grad"]) < 1e-7
    if dist is None:
        m = cls(n_dim=12, feature_noise=0.01, n_filters=3, filters=g["F0"].float()).double()
        losses, _ = m.fit(data_statistics={k: v.cpu() for k, v in stats.items()}, max_epochs=200,
                          show_progress=False, return_loss=True)
        ref = g[kind + "_fit_losses"]
        assert m.filters.dtype == torch.float64
        assert abs(float(losses[-1]) - float(ref[-1])) < 1e-6 * abs(float(ref[-1]))
        assert O.subspace_angle(m.filters.detach().cpu(), g[kind + "_fit_filters"]) < 1e-3
