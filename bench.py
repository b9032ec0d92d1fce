#!/usr/bin/env python
"""Benchmark of the SQFA hot paths on B200 -- one JSON line on stdout (rank 0).

  python bench.py --gpus 1 --steps 10 --warmup 3            # our arm, single GPU
  torchrun --nproc-per-node N ... bench.py --gpus N ...      # our arm, N GPUs (weak scaling)
  python bench.py --impl reference ...                       # the reference's CPU path (oracle port)

Workload = BASELINE.json configs[1] (CIFAR-10-shaped: N=50000, D=3072, 10 classes, n_filters=8).
A step is one `class_statistics` pass over the (per-rank) batch; with N GPUs every rank holds its
own 50000-sample shard (weak scaling) and the statistics of the union are all-reduced.
  value    : samples/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e      : the same through the public API with HOST buffers (pinned H2D of X, y and D2H of the
             three result tensors inside the timed region)
  roofline : the tcgen05 Gram kernel against the TF32 tensor peak (half the measured bf16 peak)
  fit      : closure evaluations/s and LBFGS epochs/s of SQFA.fit on the same statistics
Inputs (614 MB) exceed the 126 MB L2, so no explicit L2 flush is needed between iterations.
"""

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {"name": "configs[1] CIFAR-10-shaped synthetic", "N": 50000, "D": 3072, "C": 10, "k": 8}
METRIC = "class_statistics samples/sec"
# ours, per step at this size (n <= 65536 rows: single-block bucketing): label_max, bucket_small, class sums +
# finalize, means, gram_plan, gram_tf32x3, stats_epilogue
KERNELS_PER_STEP = 8


def synth(n, d, c, device, seed):
    """Seeded synthetic class data: low-rank class-scaled signal + noise + means."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    y = torch.randint(0, c, (n,), generator=g, device=device)
    r = 32
    basis = torch.randn(r, d, generator=g, device=device) / r**0.5
    scales = 0.5 + torch.rand(c, generator=g, device=device)
    means = 0.2 * torch.randn(c, d, generator=g, device=device)
    x = (torch.randn(n, r, generator=g, device=device) * scales[y][:, None]) @ basis
    x += 0.5 * torch.randn(n, d, generator=g, device=device)
    x += means[y]
    x /= x.std() * d**0.5
    x -= x.mean(0)
    return x.contiguous(), y


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region (NVML, every ~10 ms; falls back
    to polling nvidia-smi if the NVML bindings are missing)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()
        # NVML queries take a driver lock that kernel launches also need: sample sparsely
        self.period = float(os.environ.get("SQFA_BENCH_CLOCK_PERIOD", "0.01"))
        self.max_mhz, self.nvml, self.handle = None, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            bits = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flags = {
            "hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        self.samples.append((mhz, [k for k, v in flags.items() if bits & v]))

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                    self._stop_evt.wait(self.period)
                    continue
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    f = [x.strip() for x in out.split(",")]
                    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                    self.max_mhz = int(float(f[1]))
                    self.samples.append((int(float(f[0])), [n for n, v in zip(names, f[3:7]) if v.lower() == "active"]))
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(s[0] for s in self.samples)
        reasons = sorted({r for s in self.samples for r in s[1]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def cpu_class_statistics_rate(n, d, c, reps):
    """The oracle (port of the reference's torch-CPU path) on all host cores."""
    import torch

    from oracle import sqfa_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    x, y = synth(n, d, c, "cpu", 1234)
    O.class_statistics(x[: max(n // 10, 100)], y[: max(n // 10, 100)])  # warm the thread pool
    t0 = time.perf_counter()
    for _ in range(reps):
        O.class_statistics(x, y)
    dt = (time.perf_counter() - t0) / reps
    return n / dt, dt


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port; the reference
    is pure Python/torch and cannot travel to the GPU box). Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch

    from oracle import sqfa_oracle as O

    w = WORKLOAD
    torch.set_num_threads(os.cpu_count() or 1)
    x, y = synth(w["N"], w["D"], w["C"], "cpu", 1234)
    for _ in range(max(args.warmup, 1)):
        O.class_statistics(x[:5000], y[:5000])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.class_statistics(x, y)
    dt = time.perf_counter() - t0
    value = w["N"] * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "N_per_gpu": w["N"], "D": w["D"], "classes": w["C"], "n_filters": w["k"],
                   "l2": "host arm: not applicable",
                   "parallelism": f"reference CPU path, {torch.get_num_threads()} host threads (rank 0 only)"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"full workload, {args.steps} repetitions of N={w['N']} (warm-up on 5000 rows)"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from sqfa_b200 import statistics as S
    from sqfa_b200.model import SQFA

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    w = WORKLOAD
    n, d, c, k = w["N"], w["D"], w["C"], w["k"]
    X, y = synth(n, d, c, dev, 1234 + rank)
    ops = S._cuda_ops()

    def step():
        return S.class_statistics(X, y, group=group)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        stats = step()  # held like in the timed loop: the caching allocator reaches its steady state here
    # ---- timed region: HBM-resident inputs (no Python garbage collection pauses inside it)
    import gc

    gc.collect()
    gc.disable()
    # NVML is initialised and the sampler thread started BEFORE the barrier: done after it, rank 0
    # enters the timed region tens of ms late and every other rank's clock counts the wait
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        stats = step()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    # the same K steps once more with CUDA events around the Gram launch (on the launching stream);
    # the instrumented repetition takes the step-by-step entry points, the timed one above the
    # single-call sqfa_class_statistics when there is one GPU
    ops.gram_events = []
    for _ in range(args.steps):
        step()
    sync_all()
    clocks = sampler.stop() if sampler else None
    gram_ms = sum(a.elapsed_time(b) for a, b in ops.gram_events) / max(len(ops.gram_events), 1)
    ops.gram_events = None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    value = n * world * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers in, host results out, through the public API.
    # Every step copies its inputs from pinned host memory, calls class_statistics and copies the
    # three result tensors back to pinned host memory. Consecutive steps rotate over a few CUDA
    # streams (and as many sets of buffers), so step i+1's upload overlaps step i's compute and download
    # (PCIe is full duplex, the B200 has separate copy engines per direction); the timed region ends
    # when every step's results are on the host.
    Xh, yh = X.cpu().pin_memory(), y.cpu().pin_memory()
    nbuf = int(os.environ.get("SQFA_BENCH_E2E_BUFS", "3"))  # buffer sets / streams in flight
    streams = [torch.cuda.Stream(device=dev) for _ in range(nbuf)]
    out_h = [{key: torch.empty(v.shape, dtype=v.dtype).pin_memory() for key, v in stats.items()} for _ in range(nbuf)]
    Xd, yd = [torch.empty_like(X) for _ in range(nbuf)], [torch.empty_like(y) for _ in range(nbuf)]
    del stats

    def upload(i):  # step i's inputs, pinned host -> device, on the stream of its buffer set
        b = i % nbuf
        with torch.cuda.stream(streams[b]):
            yd[b].copy_(yh, non_blocking=True)
            Xd[b].copy_(Xh, non_blocking=True)

    def compute_and_download(i):
        b = i % nbuf
        with torch.cuda.stream(streams[b]):
            st = S.class_statistics(Xd[b], yd[b], group=group)
            for key, v in st.items():
                out_h[b][key].copy_(v, non_blocking=True)
        return st

    def e2e_run(k):
        """k steps; the upload of step i+1 is enqueued before step i is computed (input prefetch),
        every step's upload and download happen inside the run"""
        last = [None] * nbuf
        upload(0)
        for i in range(k):
            if i + 1 < k:
                upload(i + 1)
            last[i % nbuf] = compute_and_download(i)
        torch.cuda.synchronize()  # all steps' results are in host memory
        return last

    # PCIe and host memory are shared with whatever else runs on the host: the K-step run is
    # repeated and the median repetition reported (all repetitions are listed in the JSON)
    e2e_steps = max(4, args.steps)
    e2e_run(2 * nbuf)  # every buffer set twice: the per-stream allocator pools reach their steady state
    e2e_ms = []
    for _ in range(3):
        sync_all()
        e0.record()
        last = e2e_run(e2e_steps)
        e1.record()
        sync_all()
        e2e_ms.append(e0.elapsed_time(e1))
    e2e_med = sorted(e2e_ms)[1]
    stats = {key: v.to(dev) for key, v in out_h[(e2e_steps - 1) % nbuf].items()}
    ems = torch.tensor([e2e_med], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = n * world * e2e_steps / (float(ems) * 1e-3)
    h2d = Xh.numel() * 4 + yh.numel() * 8
    d2h = sum(v.numel() * 4 for v in out_h[0].values())
    del last, Xd, yd

    gc.enable()
    # ---- second hot path: SQFA fit on the statistics just computed (rank-replicated)
    fit = None
    if rank == 0:
        model = SQFA(n_dim=d, feature_noise=0.01, n_filters=k)
        model.fit_pca(data_statistics=stats)
        model = model.to(dev)
        # one closure evaluation = loss forward + gradient w.r.t. the raw filter parameter, the way
        # fitting_loop evaluates it (native, no autograd graph for the sphere constraint)
        plan = model._fused_direct_plan({kk: v for kk, v in stats.items()})
        for _ in range(5):
            plan()
        torch.cuda.synchronize()
        n_eval = 50
        e0.record()
        for _ in range(n_eval):
            plan()
        e1.record()
        torch.cuda.synchronize()
        closure_ms = e0.elapsed_time(e1) / n_eval
        # one throw-away epoch on a copy: the first torch.optim.LBFGS in a process imports half of
        # torch (~2.5 s), which is not part of an epoch
        warm = SQFA(n_dim=d, feature_noise=0.01, n_filters=k, filters=model.filters.detach().clone())
        with contextlib.redirect_stdout(sys.stderr):  # fitting_loop prints like the reference does
            warm.fit(data_statistics=stats, max_epochs=1, atol=0.0, show_progress=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        epochs = 5
        with contextlib.redirect_stdout(sys.stderr):
            model.fit(data_statistics=stats, max_epochs=epochs, atol=0.0, show_progress=False)
        torch.cuda.synchronize()
        fit_s = time.perf_counter() - t0
        fit = {"model": "SQFA fisher_rao_lower_bound, n_filters=8, pca init, feature_noise=0.01",
               "closure_evals_per_s": 1e3 / closure_ms, "closure_ms": closure_ms,
               "epochs_per_s": epochs / fit_s, "algorithmic_bytes_per_closure": 4 * c * d * d,
               "closure_hbm_gbs": 4 * c * d * d / (closure_ms * 1e-3) / 1e9}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    tiles = sum((d + 255) // 256 - (tm >> 1) for tm in range((d + 127) // 128))
    executed = 3 * 2.0 * n * (tiles * 128 * 256)  # 3 TF32 passes over the computed 128x256 tiles
    tf32_peak = peaks["bf16_tflops"] / 2.0
    achieved = executed / (gram_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "gram_tf32x3_kernel", "achieved": achieved, "peak": tf32_peak,
                "unit": "TFLOP/s", "frac": achieved / tf32_peak, "traffic": None,
                "kernel_ms": gram_ms, "executed_flop_per_launch": executed,
                "algorithmic_flop_per_launch": 2.0 * n * d * d,
                "peak_source": f"{peaks['source']}: dense TF32 = bf16_tflops / 2"}
    prof = os.path.join(ROOT, "profiles", "gram_traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            roofline["traffic"] = json.load(f).get("dram_bytes_per_launch")

    cpu = None
    if world == 1:
        rate, dt = cpu_class_statistics_rate(n, d, c, reps=3)
        import torch as _t

        cpu = {"value": rate, "unit": "samples/s", "cores": _t.get_num_threads(), "kind": "port",
               "sample": f"full workload N={n}, 3 repetitions, {dt:.2f} s each"}
        if fit is not None:
            from oracle import sqfa_oracle as O

            cstats = {kk: v.cpu() for kk, v in stats.items()}
            F0 = model.parametrizations.filters.original.detach().cpu()
            O.loss_and_grad("full", cstats, F0, noise=0.01)
            t0 = time.perf_counter()
            for _ in range(3):
                O.loss_and_grad("full", cstats, F0, noise=0.01)
            fit["cpu_closure_evals_per_s"] = 3 / (time.perf_counter() - t0)

    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (3xTF32 tensor-core split, fp32 accumulate)",
        "data": "synthetic",
        "config": {"workload": w["name"], "N_per_gpu": n, "D": d, "classes": c, "n_filters": k,
                   "l2": "inputs (614 MB per GPU) exceed the 126 MB L2; no flush needed",
                   "parallelism": f"samples sharded over {world} GPU(s), 3 all-reduces" if world > 1 else "single GPU"},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "repetitions_ms": [round(x, 2) for x in e2e_ms], "reported": "median repetition",
                "overlap": f"{nbuf} CUDA streams / buffer sets, inputs of step i+1 uploaded while step i computes and downloads"},
        "gpu_launches": KERNELS_PER_STEP * args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "fit": fit,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
