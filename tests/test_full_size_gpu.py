"""GPU: BASELINE.json configurations at FULL size, checked through size-independent properties
(the oracle would need minutes to hours at these sizes):

* bucketing: bit-exact against torch.sort(stable=True) on the device
* total second moment: sum_c [(n_c - 1) cov_c + n_c mu_c mu_c^T] == X^T X   (fp64 cuBLAS check)
* class means: n_c mu_c summed over classes == column sums of X
* symmetry, second_moments - covariances == mu mu^T
* the closure: loss finite and equal between the fused kernel and the generic (matrix) path,
  pairwise distances symmetric with the sqrt(1e-6) diagonal, a fit decreases the loss
"""

import pytest
import torch

pytestmark = pytest.mark.gpu

CONFIGS = {
    "c1": (60000, 784, 10, 4),
    "c2": (50000, 3072, 10, 8),
    "c3": (200000, 104, 19, 8),
    "c4": (1280000, 512, 1000, 16),
}


def synth(n, d, c, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    y = torch.randint(0, c, (n,), generator=g, device="cuda")
    basis = torch.randn(32, d, generator=g, device="cuda") / 32**0.5
    scales = 0.5 + torch.rand(c, generator=g, device="cuda")
    means = 0.2 * torch.randn(c, d, generator=g, device="cuda")
    x = (torch.randn(n, 32, generator=g, device="cuda") * scales[y][:, None]) @ basis
    x += 0.5 * torch.randn(n, d, generator=g, device="cuda")
    x += means[y]
    x /= x.std() * d**0.5
    return x.contiguous(), y


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4"])
def test_class_statistics_full_size_properties(cfg):
    from sqfa_b200.statistics import bucket_labels, class_statistics

    n, d, c, _ = CONFIGS[cfg]
    X, y = synth(n, d, c)
    perm, offsets, counts = bucket_labels(y)
    assert torch.equal(perm.long(), torch.sort(y, stable=True).indices)
    assert torch.equal(counts[:-1], torch.bincount(y, minlength=c))
    s = class_statistics(X, y)
    mu, cov, sm = s["means"], s["covariances"], s["second_moments"]
    assert mu.shape == (c, d) and cov.shape == (c, d, d) and sm.shape == (c, d, d)
    assert torch.equal(cov, cov.transpose(1, 2))
    nc = counts[:-1].double()
    # column sums
    col_sum = X.double().sum(0)
    got_sum = (nc[:, None] * mu.double()).sum(0)
    assert float((got_sum - col_sum).norm() / col_sum.norm().clamp_min(1e-30)) < 1e-5 or float(
        (got_sum - col_sum).abs().max()) < 1e-4 * float(X.abs().max()) * n**0.5
    # total second moment (fp64 reference GEMM on the device, class by class to bound memory)
    total = torch.zeros(d, d, dtype=torch.float64, device="cuda")
    for lo in range(0, n, 100000):
        xb = X[lo:lo + 100000].double()
        total += xb.T @ xb
    recon = torch.zeros(d, d, dtype=torch.float64, device="cuda")
    for k in range(c):
        recon += (nc[k] - 1) * cov[k].double() + nc[k] * torch.outer(mu[k].double(), mu[k].double())
    assert float((recon - total).norm() / total.norm()) < 1e-5
    # second moments are covariance + outer(mean, mean) (statistics.py:47)
    k = c // 2
    assert torch.allclose(sm[k], cov[k] + torch.outer(mu[k], mu[k]), rtol=1e-6, atol=1e-9)
    del s, total, recon
    torch.cuda.empty_cache()


@pytest.mark.parametrize("cfg", ["c1", "c3", "c4"])
def test_closure_and_fit_full_size_properties(cfg):
    from sqfa_b200.model import SQFA, SecondMomentsSQFA

    n, d, c, k = CONFIGS[cfg]
    g = torch.Generator(device="cuda").manual_seed(1)
    # class statistics of the right shape without materialising N x D data: random SPD + means
    A = torch.randn(c, d, d + 8, generator=g, device="cuda")
    cov = (A @ A.transpose(1, 2) / (d + 8) / d).contiguous()
    del A
    mu = 0.05 * torch.randn(c, d, generator=g, device="cuda") / d**0.5
    stats = {"means": mu, "covariances": cov}
    cls = SecondMomentsSQFA if cfg == "c1" else SQFA
    model = cls(n_dim=d, feature_noise=0.01, n_filters=k).cuda()
    out = model._fused_loss_plan(stats)()
    loss_fused, bad = out.detach().tolist()
    assert bad == 0 and loss_fused == loss_fused and abs(loss_fused) < 1e6
    if c <= 100:  # the generic path materialises the full C x C matrix
        dmat = model.get_class_distances(stats, regularized=True)
        assert torch.allclose(dmat, dmat.T)
        assert torch.allclose(torch.diagonal(dmat), torch.full((c,), 1e-6**0.5, device="cuda"), rtol=1e-5)
        i, j = torch.tril_indices(c, c, -1, device="cuda")
        assert abs(float(-dmat[i, j].mean()) - loss_fused) < 1e-5 * abs(loss_fused)
    epochs = 2 if c >= 1000 else 4
    loss, _ = model.fit(data_statistics=stats, max_epochs=epochs, show_progress=False, return_loss=True)
    assert torch.isfinite(loss).all()
    assert float(loss[-1]) <= float(loss[0]) + 1e-7  # the objective (minus mean distance) goes down


def test_c5_shard_full_size_and_streaming():
    """BASELINE config 5, the shard one of 8 GPUs holds: 12.5 M rows x 1024 dims (51 GB), 100 classes.
    One-shot class_statistics on the resident shard against (a) the streaming accumulator fed in four
    chunks (different code path: fixed shift + accumulate) and (b) an fp64 computation of two classes."""
    from sqfa_b200.statistics import StreamingClassStatistics, class_statistics

    n, d, c = 12_500_000, 1024, 100
    free, _ = torch.cuda.mem_get_info()
    if free < 80e9:
        pytest.skip("needs 80 GB of free device memory")
    g = torch.Generator(device="cuda").manual_seed(5)
    y = torch.randint(0, c, (n,), generator=g, device="cuda")
    means = 0.3 * torch.randn(c, d, generator=g, device="cuda")
    scale = 0.5 + torch.rand(c, 1, generator=g, device="cuda")
    X = torch.empty(n, d, device="cuda")
    step = 500_000
    for lo in range(0, n, step):  # chunked generation: no second 51 GB temporary
        yy = y[lo:lo + step]
        blk = torch.randn(yy.numel(), d, generator=g, device="cuda")
        X[lo:lo + step] = blk * scale[yy] + means[yy]
    del blk
    s = class_statistics(X, y)
    acc = StreamingClassStatistics(d, c)
    cuts = [0, 1_000_003, 5_000_000, 9_999_999, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        acc.update(X[a:b], y[a:b])
    t = acc.finalize()
    for key in ("means", "covariances", "second_moments"):
        err = float((s[key].double() - t[key].double()).norm() / t[key].double().norm())
        assert err < 1e-5, (key, err)
    for cls in (0, 57):
        rows = X[y == cls].double()
        mu = rows.mean(0)
        xc = rows - mu
        cov = xc.T @ xc / (rows.shape[0] - 1)
        assert float((s["means"][cls].double() - mu).norm() / mu.norm()) < 1e-5
        assert float((s["covariances"][cls].double() - cov).norm() / cov.norm()) < 1e-5
        sm = cov + torch.outer(mu, mu)
        assert float((s["second_moments"][cls].double() - sm).norm() / sm.norm()) < 1e-5
