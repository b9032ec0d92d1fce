"""Class statistics of labeled data points -- B200-native drop-in for `sqfa.statistics`.

Same public names and semantics as /root/reference/src/sqfa/statistics.py (`class_statistics`,
`oas_covariance`, `sample_covariance`, `pca`, `pca_from_scatter`); the arithmetic runs in the
sm_100a kernels of libsqfa_b200.so (label bucketing, segmented column sums, tcgen05 3xTF32 Gram,
statistics epilogue). Inputs may live on the CPU or on a CUDA device; outputs are returned on the
device of `points`. There is no CPU compute path.
"""

import torch

from . import _lib
from ._stats_driver import CudaStatsOps, run_class_statistics

__all__ = ["class_statistics", "oas_covariance", "pca", "pca_from_scatter"]

_ESTIMATORS = {"empirical": 0, "oas": 1}
_ops_singleton = None


def _cuda_ops():
    global _ops_singleton
    if _ops_singleton is None:
        _ops_singleton = CudaStatsOps()
    return _ops_singleton


def _as_device_points(points, dev):
    if not isinstance(points, torch.Tensor):
        raise TypeError("points must be a torch.Tensor")
    if points.dim() != 2:
        raise ValueError("points must have shape (n_points, n_dim)")
    if points.dtype != torch.float32:
        raise TypeError(f"sqfa_b200 kernels compute in float32; got points of dtype {points.dtype}")
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
    dev = _lib.compute_device(labels)
    y = _as_device_labels(labels, dev)
    ops = _cuda_ops()
    with torch.cuda.device(dev):
        if n_classes is None:
            n_classes = ops.label_max_host(y) + 1
        return ops.bucket(y, int(n_classes))


def class_statistics(points, labels, estimator="empirical", keep_on_device=False, group=None):
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
    keep_on_device : bool
        (extension) leave the result on the CUDA device even if `points` lives on the CPU.
    group : torch.distributed process group
        (extension) `points` / `labels` are this rank's shard of the samples; the statistics of the
        union over all ranks are returned on every rank (three all-reduces, see _stats_driver).

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
    if y.numel() == 0 and group is None:
        raise RuntimeError("class_statistics: empty input (max() of an empty labels tensor)")
    with torch.cuda.device(dev):
        means, cov, sm, _ = run_class_statistics(_cuda_ops(), X, y, _ESTIMATORS[estimator], group=group)
    stats = {"means": means, "covariances": cov, "second_moments": sm}
    if out_dev != dev and not keep_on_device:
        stats = {k: v.to(out_dev) for k, v in stats.items()}
    return stats


def _single_class(points, estimator_id, assume_centered):
    dev = _lib.compute_device(points)
    X = _as_device_points(points, dev)
    n, D = X.shape
    with torch.cuda.device(dev):
        y = torch.zeros(n, dtype=torch.int64, device=dev)
        centre = torch.zeros(1, D, dtype=torch.float32, device=dev) if assume_centered else None
        _, cov, _, _ = run_class_statistics(
            _cuda_ops(), X, y, estimator_id, n_classes=1, ddof=0 if assume_centered else 1, centre=centre,
            want_sm=False,
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
