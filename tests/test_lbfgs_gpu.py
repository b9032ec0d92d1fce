"""GPU: the device-side L-BFGS direction update against torch.optim.LBFGS (the reference's optimiser,
_optim.py:78-79) -- same iterates up to the summation order inside the dot products."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def _quadratic(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    diag = torch.exp(torch.empty(n, device="cuda").uniform_(-4.0, 4.0, generator=g))  # condition number ~3000
    U = torch.randn(n, 8, generator=g, device="cuda") / n**0.5
    b = torch.randn(n, generator=g, device="cuda")

    def f(x):  # 0.5 x^T (diag + U U^T) x - b^T x
        return 0.5 * (x * diag * x).sum() + 0.5 * ((x @ U) ** 2).sum() - (b * x).sum()

    return f


def _run(opt_cls, f, n, steps, **kw):
    x = torch.nn.Parameter(torch.zeros(n, device="cuda"))
    opt = opt_cls([x], **kw)
    losses = []

    def closure():
        opt.zero_grad()
        loss = f(x)
        loss.backward()
        return loss

    for _ in range(steps):
        losses.append(float(opt.step(closure)))
    return x.detach().clone(), losses, opt


@pytest.mark.parametrize("n,hist", [(3000, 5), (20000, 100), (100000, 7), (8192 * 16, 3)])
def test_direction_update_matches_torch_lbfgs(n, hist):
    from sqfa_b200._lbfgs import LBFGS

    f = _quadratic(n, n)
    kw = dict(lr=0.3, history_size=hist, max_iter=12)
    x_ref, l_ref, opt_ref = _run(torch.optim.LBFGS, f, n, 3, **kw)
    x_got, l_got, opt = _run(LBFGS, f, n, 3, **kw)
    st, st_ref = opt.state[opt._params[0]], opt_ref.state[opt_ref._params[0]]
    assert "sqfa_native" in st, "the native path was not taken"
    assert st["n_iter"] == st_ref["n_iter"] == 36 and st["func_evals"] == st_ref["func_evals"]
    assert float((x_got - x_ref).norm() / x_ref.norm()) < 1e-4
    for a, b in zip(l_got, l_ref):
        assert abs(a - b) <= 1e-5 * max(1.0, abs(b))
    assert int(st["sqfa_native"]["meta"][1]) == len(st_ref["old_dirs"]) == min(hist, 35)  # the ring wrapped


def test_unsupported_configurations_fall_back_to_torch():
    from sqfa_b200._lbfgs import LBFGS

    f = _quadratic(500, 1)
    x_ref, l_ref, _ = _run(torch.optim.LBFGS, f, 500, 2, lr=1.0, line_search_fn="strong_wolfe")
    x_got, l_got, opt = _run(LBFGS, f, 500, 2, lr=1.0, line_search_fn="strong_wolfe")
    assert "sqfa_native" not in opt.state[opt._params[0]]
    assert torch.equal(x_ref, x_got) and l_ref == l_got


def test_fit_with_device_lbfgs_matches_torch_lbfgs(monkeypatch):
    """SQFA.fit loss per epoch: device-side direction update vs plain torch.optim.LBFGS."""
    import sqfa_b200._optim as optim_mod
    from conftest import make_class_data
    from sqfa_b200.model import SQFA
    from sqfa_b200.statistics import class_statistics

    X, y = make_class_data(4000, 60, 5, seed=2)
    stats = class_statistics(X.cuda(), y.cuda())
    F0 = torch.randn(3, 60, generator=torch.Generator().manual_seed(0))
    out = {}
    for name, cls in (("native", optim_mod.LBFGS), ("torch", torch.optim.LBFGS)):
        monkeypatch.setattr(optim_mod, "LBFGS", cls)
        model = SQFA(n_dim=60, feature_noise=0.01, n_filters=3, filters=F0.clone())
        loss, _ = model.fit(data_statistics=stats, max_epochs=3, atol=0.0, show_progress=False, return_loss=True,
                            max_iter=8)
        out[name] = (loss, model.filters.detach().cpu())
    # loss and gradient are deterministic (fixed-order reductions); what differs between the two
    # optimisers is the summation order inside the dot products of the two-loop recursion
    assert torch.allclose(out["native"][0], out["torch"][0], rtol=2e-4, atol=1e-6)
    Fa, Fb = out["native"][1], out["torch"][1]
    assert float((Fa - Fb).norm() / Fb.norm()) < 5e-3


@pytest.mark.parametrize("constraint", ["sphere", "none"])
def test_fit_constraints_direct_closure_vs_autograd(monkeypatch, constraint):
    """The graph-free closure (closed-form constraint adjoint) and the autograd closure give the same
    fit. (`orthogonal`, torch's parametrisation, always takes the autograd closure -- checked on the
    CPU in test_host_logic.py; its matrix-exponential dynamics are too chaotic for a trajectory test.)"""
    from conftest import make_class_data
    from sqfa_b200.model import SQFA
    from sqfa_b200.statistics import class_statistics

    X, y = make_class_data(3000, 40, 4, seed=5)
    stats = class_statistics(X.cuda(), y.cuda())
    F0 = torch.linalg.qr(torch.randn(40, 3, generator=torch.Generator().manual_seed(1)))[0].T.contiguous()
    runs = {}
    for mode in ("direct", "autograd"):
        model = SQFA(n_dim=40, feature_noise=0.01, n_filters=3, filters=F0.clone(), constraint=constraint)
        if mode == "autograd":
            monkeypatch.setattr(model, "_fused_direct_plan", lambda data_statistics: None)
        else:
            assert model.cuda()._fused_direct_plan(stats) is not None
        loss, _ = model.fit(data_statistics=stats, max_epochs=2, atol=0.0, show_progress=False, return_loss=True,
                            max_iter=5)
        runs[mode] = (loss, model.filters.detach().cpu())
    assert torch.allclose(runs["direct"][0], runs["autograd"][0], rtol=2e-4, atol=1e-6)
    assert float((runs["direct"][1] - runs["autograd"][1]).norm() / runs["autograd"][1].norm()) < 5e-3
    assert runs["direct"][0][-1] < runs["direct"][0][0]  # the loss went down
    F = runs["direct"][1]
    if constraint == "sphere":
        assert torch.allclose(F.norm(dim=1), torch.ones(3), atol=1e-5)


@pytest.mark.parametrize("c,k,kind", [(12, 4, "full"), (200, 8, "full"), (40, 6, "second_moments")])
def test_closure_and_fit_are_bit_reproducible(c, k, kind):
    """No floating-point atomics anywhere in the closure: the same inputs give the same bits, through
    the directly launched closure, its CUDA-graph replay, and a whole fit (C = 200: 19 900 pairs, 2 x 2 pair tiles)."""
    from conftest import make_class_data
    from sqfa_b200.model import SQFA, SecondMomentsSQFA
    from sqfa_b200.statistics import class_statistics

    d = 48
    X, y = make_class_data(60 * c, d, c, seed=c)
    stats = class_statistics(X.cuda(), y.cuda())
    cls = SQFA if kind == "full" else SecondMomentsSQFA
    F0 = torch.randn(k, d, generator=torch.Generator().manual_seed(k))
    outs, grads = [], []
    for _ in range(2):
        model = cls(n_dim=d, feature_noise=0.01, n_filters=k, filters=F0.clone()).cuda()
        plan = model._fused_direct_plan(stats)
        for _ in range(4):  # call 1: direct launch, call 2: capture, calls 3-4: graph replays
            outs.append(plan().clone())
            grads.append(model.parametrizations.filters.original.grad.clone())
    for o, g in zip(outs[1:], grads[1:]):
        assert torch.equal(o, outs[0]) and torch.equal(g, grads[0])
    fits = []
    for _ in range(2):
        model = cls(n_dim=d, feature_noise=0.01, n_filters=k, filters=F0.clone())
        loss, _ = model.fit(data_statistics=stats, max_epochs=3, atol=0.0, show_progress=False, return_loss=True,
                            max_iter=7)
        fits.append((loss, model.filters.detach().clone()))
    assert torch.equal(fits[0][0], fits[1][0]) and torch.equal(fits[0][1], fits[1][1])


def test_pipelined_iterations_equal_waited_ones(monkeypatch):
    """The fitting loop enqueues the evaluation at the new iterate behind the direction kernel (which
    applies the step itself) and waits once per iteration; with SQFA_LBFGS_PIPELINE=0 it waits after each
    kernel. Same kernels on the same data in the same order: identical losses, filters and evaluation counts."""
    from conftest import make_class_data
    from sqfa_b200.model import SQFA
    from sqfa_b200.statistics import class_statistics

    d, c, k = 64, 15, 5
    X, y = make_class_data(50 * c, d, c, seed=21)
    stats = class_statistics(X.cuda(), y.cuda())
    F0 = torch.randn(k, d, generator=torch.Generator().manual_seed(2))
    runs = []
    for pipe in ("1", "0"):
        monkeypatch.setenv("SQFA_LBFGS_PIPELINE", pipe)
        model = SQFA(n_dim=d, feature_noise=0.01, n_filters=k, filters=F0.clone())
        loss, _ = model.fit(data_statistics=stats, max_epochs=4, atol=0.0, show_progress=False, return_loss=True,
                            max_iter=9)
        runs.append((loss, model.filters.detach().clone(), model._last_fit_evaluations))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    assert runs[0][2] == runs[1][2]
    # a fit that converges inside an epoch (the optimiser's own stopping tests fire): same again
    for pipe in ("1", "0"):
        monkeypatch.setenv("SQFA_LBFGS_PIPELINE", pipe)
        model = SQFA(n_dim=d, feature_noise=0.01, n_filters=k, filters=F0.clone())
        loss, _ = model.fit(data_statistics=stats, max_epochs=40, atol=1e-7, show_progress=False, return_loss=True,
                            tolerance_change=1e-7, tolerance_grad=1e-5)
        runs.append((loss, model.filters.detach().clone(), model._last_fit_evaluations))
    assert torch.equal(runs[2][0], runs[3][0]) and torch.equal(runs[2][1], runs[3][1])
    assert runs[2][2] == runs[3][2]
