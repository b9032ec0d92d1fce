"""Driver of hot path 1 (class_statistics): the sequence of local compute steps and -- when the
samples are sharded over the ranks of a process group -- the collectives between them.

Per-class (count, sum x, sum (x - mu)(x - mu)^T) are additive over samples, so every rank
processes its own rows and three all-reduces make the result global:
max label -> n_classes, [class sums | class counts] -> global means, Gram partials -> global
covariances. The local steps are supplied by an `ops` object: `CudaStatsOps` (the sm_100a kernels
through the C ABI) in the product; the CPU test-suite injects oracle-backed ops to exercise this
host logic under the gloo backend.
"""

import torch

from . import _lib


class CudaStatsOps:
    """Local steps of class_statistics on the current CUDA device (kernels K1, K2a, K2, K3)."""

    def __init__(self):
        self.lib = _lib.load()
        self.gram_events = None  # set to a list to record (start, end) CUDA events around K2

    def label_max(self, y):
        mx = torch.empty(1, dtype=torch.int64, device=y.device)
        _lib.check(
            self.lib.sqfa_label_max(_lib.ptr(y), y.numel(), _lib.ptr(mx), _lib.stream_ptr(y.device)),
            "sqfa_label_max",
        )
        return mx

    _HOST_SLOTS = 32

    def label_max_host(self, y):
        """max(labels) as a Python int: the kernel stores it in mapped pinned host memory and the
        host waits on an event. No device-to-host copy, so the read does not queue behind bulk
        transfers other streams have on the copy engine (`.item()` would)."""
        if getattr(self, "_host_mx", None) is None:
            self._host_mx = torch.empty(self._HOST_SLOTS, dtype=torch.int64).pin_memory()
            self._host_next = 0
        k = self._host_next
        self._host_next = (k + 1) % self._HOST_SLOTS
        slot = self._host_mx[k : k + 1]
        _lib.check(
            self.lib.sqfa_label_max(_lib.ptr(y), y.numel(), _lib.ptr(slot), _lib.stream_ptr(y.device)),
            "sqfa_label_max",
        )
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(y.device))
        ev.synchronize()
        return int(slot[0])

    def bucket(self, y, C):
        lib, dev, n = self.lib, y.device, y.numel()
        counts = torch.empty(C + 1, dtype=torch.int64, device=dev)
        offsets = torch.empty(C + 2, dtype=torch.int64, device=dev)
        perm = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        ws_bytes = lib.sqfa_bucket_workspace_bytes(n, C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(
            lib.sqfa_bucket_labels(
                _lib.ptr(y), n, C, _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(perm), _lib.ptr(ws), ws_bytes,
                _lib.stream_ptr(dev),
            ),
            "sqfa_bucket_labels",
        )
        return perm[:n], offsets, counts

    def class_sums(self, X, perm, offsets, C):
        lib, dev = self.lib, X.device
        n, D = X.shape
        sums = torch.empty(C, D, dtype=torch.float32, device=dev)
        ws_bytes = lib.sqfa_class_sums_workspace_bytes(n, D, C)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        _lib.check(
            lib.sqfa_class_sums(
                _lib.ptr(X), X.stride(0), _lib.ptr(perm), _lib.ptr(offsets), None, n, D, C, _lib.ptr(sums), 0,
                _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev),
            ),
            "sqfa_class_sums",
        )
        return sums

    def class_sums_shifted(self, X, perm, offsets, shift, C):
        """sum over each class of (x - shift_c) (streaming accumulator)."""
        lib, dev = self.lib, X.device
        n, D = X.shape
        sums = torch.empty(C, D, dtype=torch.float32, device=dev)
        ws_bytes = lib.sqfa_class_sums_workspace_bytes(n, D, C)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        _lib.check(
            lib.sqfa_class_sums(
                _lib.ptr(X), X.stride(0), _lib.ptr(perm), _lib.ptr(offsets), _lib.ptr(shift), n, D, C, _lib.ptr(sums),
                0, _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev),
            ),
            "sqfa_class_sums",
        )
        return sums

    def class_means(self, sums, counts):
        C, D = sums.shape
        means = torch.empty_like(sums)
        _lib.check(
            self.lib.sqfa_class_means(
                _lib.ptr(sums), _lib.ptr(counts), None, D, C, _lib.ptr(means), _lib.stream_ptr(sums.device)
            ),
            "sqfa_class_means",
        )
        return means

    supports_packed = True  # class_gram / finalize understand the packed upper-tile layout

    def packed_is_smaller(self, D, C):
        return self.lib.sqfa_gram_packed_floats(D, C) < C * D * D  # padding to 256 can outweigh the triangle

    def class_gram(self, X, perm, offsets, centre, C, packed=False):
        """Upper triangle of sum_{i in c} (x_i - centre_c)(x_i - centre_c)^T per class (tcgen05).

        packed=True returns the flat list of 256 x 256 upper tiles (what ranks all-reduce: about
        half the bytes of (C, D, D)); `finalize(..., packed=True)` consumes it."""
        lib, dev = self.lib, X.device
        n, D = X.shape
        if packed:
            gram = torch.empty(lib.sqfa_gram_packed_floats(D, C), dtype=torch.float32, device=dev)
        else:
            gram = torch.empty(C, D, D, dtype=torch.float32, device=dev)
        ws_bytes = lib.sqfa_class_gram_workspace_bytes(n, D, C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        if self.gram_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _lib.check(
            lib.sqfa_class_gram(
                _lib.ptr(X), X.stride(0), _lib.ptr(perm), _lib.ptr(offsets), _lib.ptr(centre), n, D, C,
                _lib.ptr(gram), 2 if packed else 0, 0, _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev),
            ),
            "sqfa_class_gram",
        )
        if self.gram_events is not None:
            ev[1].record()
            self.gram_events.append(ev)
        return gram

    def multicast_buffer(self, group, C, D, dev):
        """(buffer, symmetric-memory handle) for the fused Gram + all-reduce: a packed-Gram buffer that
        sits at the same offset on every rank, with an NVSwitch multicast alias; None unless enabled
        (SQFA_GRAM_MULTICAST=1) and supported by the platform. Collective: every rank of the group calls it
        with the same arguments. The buffers are cached per (group, size)."""
        import os

        import torch.distributed as dist

        # Opt-in (SQFA_GRAM_MULTICAST=1). Measured on 2 B200s at c2: 4.39 ms per step against 2.85 ms
        # with the packed NCCL all-reduce -- the 16-byte multimem.red instructions issued by the four
        # epilogue warps per CTA are latency-limited and hold up the accumulator hand-over; a bulk
        # (TMA-sized) push is what this path needs before it can be the default.
        if os.environ.get("SQFA_GRAM_MULTICAST") != "1" or dist.get_world_size(group) < 2:
            return None
        numel = self.lib.sqfa_gram_packed_floats(D, C)
        cache = self.__dict__.setdefault("_mc_cache", {})
        key = (id(group), numel)
        if key not in cache:
            entry = None
            try:
                import torch.distributed._symmetric_memory as symm_mem

                buf = symm_mem.empty(numel, dtype=torch.float32, device=dev)
                hdl = symm_mem.rendezvous(buf, group.group_name)
                ok = int(hdl.multicast_ptr) != 0
                entry = (buf, hdl)
            except Exception:  # no symmetric memory on this platform / backend
                ok = False
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)  # all ranks take the same path
            cache[key] = entry if int(flag.item()) == 1 else None
        return cache[key]

    def class_gram_multicast(self, X, perm, offsets, centre, C, buf, multicast_ptr):
        """This rank's Gram partials, added tile by tile into every rank's `buf` through NVSwitch."""
        import ctypes

        lib, dev = self.lib, X.device
        n, D = X.shape
        ws_bytes = lib.sqfa_class_gram_workspace_bytes(n, D, C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(
            lib.sqfa_class_gram_multicast(
                _lib.ptr(X), X.stride(0), _lib.ptr(perm), _lib.ptr(offsets), _lib.ptr(centre), n, D, C, _lib.ptr(buf),
                ctypes.c_void_p(int(multicast_ptr)), 0, _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev),
            ),
            "sqfa_class_gram_multicast",
        )

    def finalize(self, gram, means, counts, estimator_id, ddof, want_sm, packed=False):
        """cov (in place over gram unless it is packed; mirrored), optional OAS shrinkage, second moments."""
        lib, dev = self.lib, gram.device
        C, D = means.shape
        cov = torch.empty(C, D, D, dtype=torch.float32, device=dev) if packed else gram
        sm = torch.empty(C, D, D, dtype=torch.float32, device=dev) if want_sm else None
        ws_bytes = lib.sqfa_stats_epilogue_workspace_bytes(C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(
            lib.sqfa_stats_epilogue(
                _lib.ptr(gram), _lib.ptr(means), None, _lib.ptr(counts), D, C, estimator_id | (16 if packed else 0),
                ddof, _lib.ptr(cov), _lib.ptr(sm), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev),
            ),
            "sqfa_stats_epilogue",
        )
        return cov, sm


    def fused(self, X, y, C, estimator_id, ddof, want_sm):
        """All local steps from ONE C call (sqfa_class_statistics): no host work between kernels."""
        lib, dev = self.lib, X.device
        n, D = X.shape
        means = torch.empty(C, D, dtype=torch.float32, device=dev)
        cov = torch.empty(C, D, D, dtype=torch.float32, device=dev)
        sm = torch.empty(C, D, D, dtype=torch.float32, device=dev) if want_sm else None
        meta = torch.empty(2 * C + 4, dtype=torch.int64, device=dev)  # counts [C+1] | offsets [C+2]
        counts, offsets = meta[: C + 1], meta[C + 1 : 2 * C + 3]
        perm = torch.empty(n, dtype=torch.int32, device=dev)
        ws_bytes = lib.sqfa_class_statistics_workspace_bytes(n, D, C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(
            lib.sqfa_class_statistics(
                _lib.ptr(X), X.stride(0), _lib.ptr(y), n, D, C, estimator_id, ddof, _lib.ptr(means), _lib.ptr(cov),
                _lib.ptr(sm), _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(perm), _lib.ptr(ws), ws_bytes,
                _lib.stream_ptr(dev),
            ),
            "sqfa_class_statistics",
        )
        return means, cov, sm, (perm, offsets, counts)


def _all_reduce(t, group, op=None):
    import torch.distributed as dist

    dist.all_reduce(t, op=op if op is not None else dist.ReduceOp.SUM, group=group)


def run_class_statistics(ops, X, y, estimator_id, group=None, n_classes=None, ddof=1, centre=None, want_sm=True):
    """means, covariances, second moments of the labeled rows; with `group`, of the union of the
    rows held by all ranks (every rank returns the full global result).

    Two passes over the local rows, like the reference (mean first, then the Gram of the centred
    rows, statistics.py:118-120). `centre` (C, D) overrides the centring vectors.
    """
    if n_classes is None:
        # the one host read the reference also does (statistics.py:29)
        if group is None and getattr(ops, "label_max_host", None) is not None:
            n_classes = ops.label_max_host(y) + 1
        else:
            mx = ops.label_max(y)
            if group is not None:
                import torch.distributed as dist

                _all_reduce(mx, group, dist.ReduceOp.MAX)
            if getattr(ops, "label_max_host", None) is not None:
                n_classes = ops.label_max_host(mx) + 1  # max of one element: publishes it to the host
            else:
                n_classes = int(mx.item()) + 1
    C = n_classes
    if (
        group is None and centre is None and C > 0 and y.numel() > 0
        and getattr(ops, "fused", None) is not None and getattr(ops, "gram_events", None) is None
    ):
        return ops.fused(X, y, C, estimator_id, ddof, want_sm)
    perm, offsets, counts = ops.bucket(y, C)
    mc = None
    if group is not None and centre is None and getattr(ops, "multicast_buffer", None) is not None:
        mc = ops.multicast_buffer(group, C, X.shape[1], X.device)
        if mc is not None:
            # zero-filled BEFORE the all-reduce below: that collective cannot complete anywhere until
            # every rank has contributed, i.e. has passed this point in its stream
            mc[0].zero_()
    sums = ops.class_sums(X, perm, offsets, C)
    class_counts = counts[:C].clone()
    if group is not None:
        _all_reduce(sums, group)
        _all_reduce(class_counts, group)
    means = ops.class_means(sums, class_counts)
    shift = means if centre is None else centre
    if mc is not None:
        # Gram fused with its all-reduce: every finished tile is multimem.red-added into the copy of
        # every rank while the tensor cores work on the next one; one barrier, no collective
        buf, hdl = mc
        ops.class_gram_multicast(X, perm, offsets, shift, C, buf, hdl.multicast_ptr)
        hdl.barrier()
        cov, sm = ops.finalize(buf, means, class_counts, estimator_id, ddof, want_sm, packed=True)
    elif group is not None and getattr(ops, "supports_packed", False) and ops.packed_is_smaller(X.shape[1], C):
        # the one large collective: partial Gram sums over NVLink, upper 256 x 256 tiles only
        gram = ops.class_gram(X, perm, offsets, shift, C, packed=True)
        _all_reduce(gram, group)
        cov, sm = ops.finalize(gram, means, class_counts, estimator_id, ddof, want_sm, packed=True)
    else:
        gram = ops.class_gram(X, perm, offsets, shift, C)
        if group is not None:
            _all_reduce(gram, group)
        cov, sm = ops.finalize(gram, means, class_counts, estimator_id, ddof, want_sm)
    return means, cov, sm, (perm, offsets, counts)


def pair_range(n_pairs, rank, world):
    """Contiguous slice [begin, end) of the linearised class-pair list owned by `rank`."""
    return (n_pairs * rank) // world, (n_pairs * (rank + 1)) // world
