"""Class statistics of labeled data points -- B200-native drop-in for `sqfa.statistics`.

Same public names and semantics as /root/reference/src/sqfa/statistics.py (`class_statistics`,
`oas_covariance`, `sample_covariance`, `pca`, `pca_from_scatter`); the arithmetic runs in the
sm_100a kernels of libsqfa_b200.so (label bucketing, segmented column sums, tcgen05 3xTF32 Gram,
statistics epilogue). Inputs may live on the CPU or on a CUDA device; outputs are returned on the
device of `points`. There is no CPU compute path.
"""

import torch

from . import _lib

__all__ = ["class_statistics", "oas_covariance", "pca", "pca_from_scatter"]

_ESTIMATORS = {"empirical": 0, "oas": 1}


def _as_device_points(points, dev):
    if not isinstance(points, torch.Tensor):
        raise TypeError("points must be a torch.Tensor")
    if points.dim() != 2:
        raise ValueError("points must have shape (n_points, n_dim)")
    if points.dtype != torch.float32:
        raise TypeError(
            f"sqfa_b200 kernels compute in float32; got points of dtype {points.dtype}"
        )
    X = points.detach().to(dev, non_blocking=True)
    if X.stride(1) != 1 or X.stride(0) < X.shape[1]:
        X = X.contiguous()
    return X


def _as_device_labels(labels, dev):
    """int64 labels on `dev`. Rows whose label equals no integer class (negative or non-integral
    values) get -1: the reference's `labels == i` (statistics.py:37) never selects them."""
    y = torch.as_tensor(labels).detach().reshape(-1)
    if y.dtype.is_floating_point:
        y = y.to(dev, non_blocking=True)
        yi = y.to(torch.int64)
        yi = torch.where(yi.to(y.dtype) == y, yi, torch.full_like(yi, -1))
        return yi.contiguous()
    if y.dtype == torch.bool:
        y = y.to(torch.int64)
    return y.to(dev, dtype=torch.int64, non_blocking=True).contiguous()


def bucket_labels(labels, n_classes=None):
    """Stable bucketing of row ids by label on the device (kernel K1).

    Returns (perm int32 [n], offsets int64 [C+2], counts int64 [C+1]) as CUDA tensors; entry C of
    counts / offsets is the bucket of rows whose label is outside [0, C). `perm[:offsets[C]]`
    equals `torch.sort(labels, stable=True).indices` for in-range labels (bit-exact).
    """
    lib = _lib.load()
    dev = _lib.compute_device(labels)
    y = _as_device_labels(labels, dev)
    n = y.numel()
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        if n_classes is None:
            mx = torch.empty(1, dtype=torch.int64, device=dev)
            _lib.check(lib.sqfa_label_max(_lib.ptr(y), n, _lib.ptr(mx), st), "sqfa_label_max")
            n_classes = int(mx.item()) + 1
        C = int(n_classes)
        counts = torch.empty(C + 1, dtype=torch.int64, device=dev)
        offsets = torch.empty(C + 2, dtype=torch.int64, device=dev)
        perm = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        ws_bytes = lib.sqfa_bucket_workspace_bytes(n, C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(
            lib.sqfa_bucket_labels(
                _lib.ptr(y), n, C, _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(perm), _lib.ptr(ws), ws_bytes, st
            ),
            "sqfa_bucket_labels",
        )
    return perm[:n], offsets, counts


def _device_statistics(X, perm, offsets, counts, C, estimator_id, ddof=1, shift=None, want_sm=True):
    """means / covariances / second moments of the bucketed rows of X (all on X.device).

    Two passes over X, exactly like the reference (mean, then Gram of the centred rows,
    statistics.py:118-120): pass 1 = per-class column sums -> means, pass 2 = tensor-core Gram of
    (x - mean_c). `shift` overrides the centring vector (used with ddof=0 for assume_centered).
    """
    lib = _lib.load()
    dev = X.device
    n, D = X.shape
    st = _lib.stream_ptr(dev)
    ldx = X.stride(0)

    sums = torch.empty(C, D, dtype=torch.float32, device=dev)
    means = torch.empty(C, D, dtype=torch.float32, device=dev)
    ws_bytes = lib.sqfa_class_sums_workspace_bytes(n, D, C)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    _lib.check(
        lib.sqfa_class_sums(
            _lib.ptr(X), ldx, _lib.ptr(perm), _lib.ptr(offsets), None, n, D, C, _lib.ptr(sums), 0,
            _lib.ptr(ws), ws_bytes, st,
        ),
        "sqfa_class_sums",
    )
    _lib.check(
        lib.sqfa_class_means(_lib.ptr(sums), _lib.ptr(counts), None, D, C, _lib.ptr(means), st),
        "sqfa_class_means",
    )

    centre = means if shift is None else shift
    # the Gram lands directly in the covariance buffer; the epilogue rescales / mirrors it in place
    cov = torch.empty(C, D, D, dtype=torch.float32, device=dev)
    sm = torch.empty(C, D, D, dtype=torch.float32, device=dev) if want_sm else None
    gws_bytes = lib.sqfa_class_gram_workspace_bytes(C)
    gws = torch.empty(gws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(
        lib.sqfa_class_gram(
            _lib.ptr(X), ldx, _lib.ptr(perm), _lib.ptr(offsets), _lib.ptr(centre), D, C, _lib.ptr(cov), 0, 0,
            _lib.ptr(gws), gws_bytes, st,
        ),
        "sqfa_class_gram",
    )
    ews_bytes = lib.sqfa_stats_epilogue_workspace_bytes(C)
    ews = torch.empty(ews_bytes, dtype=torch.uint8, device=dev)
    _lib.check(
        lib.sqfa_stats_epilogue(
            _lib.ptr(cov), _lib.ptr(means), None, _lib.ptr(counts), D, C, estimator_id, ddof, _lib.ptr(cov),
            _lib.ptr(sm), _lib.ptr(ews), ews_bytes, st,
        ),
        "sqfa_stats_epilogue",
    )
    return means, cov, sm


def class_statistics(points, labels, estimator="empirical", keep_on_device=False):
    """
    Compute the mean, covariance and second moment matrix of each class.

    Drop-in for `sqfa.statistics.class_statistics` (reference statistics.py:8-54).

    Parameters
    ----------
    points : torch.Tensor
        Data points with shape (n_points, n_dim), float32, on the CPU or a CUDA device.
    labels : torch.Tensor
        Class labels of each point with shape (n_points).
    estimator:
        Covariance estimator to use. Options are "empirical" and "oas". Default is "empirical".

    Returns
    -------
    statistics_dict : dict
        "means" (n_classes, n_dim), "covariances" and "second_moments" (n_classes, n_dim, n_dim),
        on the device of `points`.
    """
    if estimator not in _ESTIMATORS:
        raise ValueError(f"estimator must be 'empirical' or 'oas', got {estimator!r}")
    dev = _lib.compute_device(points, labels)
    out_dev = points.device if isinstance(points, torch.Tensor) else dev
    X = _as_device_points(points, dev)
    y = _as_device_labels(labels, dev)
    if y.numel() != X.shape[0]:
        raise ValueError("labels must have one entry per row of points")
    if y.numel() == 0:
        raise RuntimeError("class_statistics: empty input (max() of an empty labels tensor)")
    with torch.cuda.device(dev):
        perm, offsets, counts = bucket_labels(y)
        C = counts.numel() - 1
        means, cov, sm = _device_statistics(X, perm, offsets, counts, C, _ESTIMATORS[estimator])
    stats = {"means": means, "covariances": cov, "second_moments": sm}
    if out_dev != dev and not keep_on_device:
        stats = {k: v.to(out_dev) for k, v in stats.items()}
    return stats


def _single_class(points, estimator_id, assume_centered):
    dev = _lib.compute_device(points)
    X = _as_device_points(points, dev)
    n, D = X.shape
    with torch.cuda.device(dev):
        perm = torch.arange(n, dtype=torch.int32, device=dev)
        offsets = torch.tensor([0, n, n], dtype=torch.int64, device=dev)
        counts = torch.tensor([n, 0], dtype=torch.int64, device=dev)
        shift = torch.zeros(1, D, dtype=torch.float32, device=dev) if assume_centered else None
        _, cov, _ = _device_statistics(
            X, perm, offsets, counts, 1, estimator_id, ddof=0 if assume_centered else 1, shift=shift, want_sm=False
        )
    cov = cov[0]
    return cov if points.device == dev else cov.to(points.device)


def sample_covariance(points, assume_centered=False):
    """Sample covariance matrix of the given points (reference statistics.py:97-124)."""
    return _single_class(points, 0, assume_centered)


def oas_covariance(points, assume_centered=False):
    """OAS shrinkage covariance of the given points (reference statistics.py:57-94)."""
    return _single_class(points, 1, assume_centered)


def pca(points, n_components=None):
    """
    Principal components of the given points, shape (n_components, n_dim), descending variance
    (reference statistics.py:127-160). The D x D `eigh` is a one-off library call on the device.
    """
    n_points, n_dim = points.shape
    if n_components is None:
        n_components = min(n_points, n_dim)
    if n_components > n_dim:
        raise ValueError("n_components must be less than or equal to n_dim.")
    covariance = sample_covariance(points)
    dev = _lib.compute_device(covariance)
    _, components = torch.linalg.eigh(covariance.to(dev))
    components = components[:, -n_components:]
    components = torch.flip(components, dims=[1]).T
    return components.to(points.device)


def pca_from_scatter(scatters, n_components=None):
    """
    Principal components from class scatter matrices (reference statistics.py:163-192). As in the
    reference, the mean scatter matrix is handed to `pca` as if it were a point cloud.
    """
    n_classes, n_dim, _ = scatters.shape
    if n_components is None:
        n_components = n_dim
    if n_components > n_dim:
        raise ValueError("n_components must be less than or equal to n_dim.")
    average_scatter = torch.mean(scatters, dim=0)
    return pca(average_scatter, n_components=n_components)
