#!/usr/bin/env python
"""Benchmark of the SQFA hot paths on B200 -- one JSON line on stdout (rank 0).

  python bench.py --gpus 1 --steps 10 --warmup 3            # our arm, single GPU
  torchrun --nproc-per-node N ... bench.py --gpus N ...      # our arm, N GPUs (weak scaling)
  python bench.py --impl reference ...                       # the reference's CPU path (oracle port)

Headline workload = BASELINE.json configs[1] (CIFAR-10-shaped: N=50000, D=3072, 10 classes,
n_filters=8). A step is one `class_statistics` pass over the (per-rank) batch; with N GPUs every
rank holds its own 50000-sample shard (weak scaling) and the statistics of the union are combined.
  value     : samples/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       : the same through the public API with HOST buffers (pinned H2D of X, y and D2H of the
              result tensors inside the timed region)
  roofline  : the tcgen05 Gram kernel against the dense TF32 tensor peak MEASURED IN THIS RUN
              (cuBLAS TF32 8192^3, best of 10)
  hp1_table : class_statistics at the other BASELINE configs (c1, c3, c4, the c5 shard of one GPU)
              with their fraction of the co-bound (HBM, tensor) roofline
  fit       : closure evaluations/s and LBFGS epochs/s of SQFA.fit on the c2 statistics
  fit_c4    : the second hot path on configs[3] (C=1000, D=512, k=16, 499500 pairs per evaluation):
              closure evaluations/s, epochs/s, pairs/s, per-kernel rooflines, the CPU closure timed on
              this box (2 evaluations, extrapolated by the evaluation count of the GPU fit) and the
              resulting fit speed-up; with N GPUs the pair list is sharded over the ranks
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
# ours, per step at this size (one 8-bit bucketing pass): label_max, radix_hist, scan_offsets, radix_scatter,
# class sums + finalize, means, gram_plan, gram_tf32x3, stats_epilogue
KERNELS_PER_STEP = 10
# N > 1 (step-by-step entry points + the fused reduce-scatter): one more label_max (publishes the all-reduced
# maximum to the host) and the epilogue in its multi-source form -- torch's element-wise kernels that pack the
# counts, NCCL's kernels and the copy-engine pushes are not counted
KERNELS_PER_STEP_MULTI = 11
PARALLELISM_NOTE = ("samples sharded over {world} GPUs; statistics of the union, every rank finalising its share of the "
                    "classes: reduce-scatter by class fused into the Gram kernel (copy-engine pushes over NVLink into "
                    "peer-mapped slots) and the epilogue; 3 small NCCL all-reduces (max label, sums + counts, barrier)")


def workload_config(world):
    """The `config` object of the JSON line: the same for both arms at the same N (the driver compares them)."""
    w = WORKLOAD
    return {"workload": w["name"], "N_per_gpu": w["N"], "D": w["D"], "classes": w["C"], "n_filters": w["k"],
            "l2": "inputs (614 MB per GPU) exceed the 126 MB L2; no flush needed",
            "parallelism": PARALLELISM_NOTE.format(world=world) if world > 1 else "single GPU"}


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


CONFIGS = {
    # BASELINE.json configs other than the headline one: (N, D, C, k)
    "c1": (60000, 784, 10, 4),
    "c3": (200000, 104, 19, 8),
    "c4": (1280000, 512, 1000, 16),
    "c5_shard": (12500000, 1024, 100, 32),  # the rows one of 8 GPUs holds of N = 1e8
}


def measure_tf32_peak(dev):
    """Dense TF32 tensor peak of this GPU, measured in this run: cuBLAS fp32 matmul with TF32 allowed,
    8192^3, best of 10 (burst) -- SURVEY.md 8(d) / BASELINE.md section 3. TFLOP/s."""
    import torch

    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b, c
        return 2.0 * n**3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def gram_executed_flop(n, d):
    """TF32 flop the Gram kernel executes for n samples of dimension d: 3 passes (hi*hi, hi*lo, lo*hi)
    over the 128 x 256 half tiles that intersect the upper triangle of the 256-padded D x D matrix."""
    from sqfa_b200 import _lib

    return 3 * 2.0 * n * _lib.load().sqfa_gram_executed_tile_area(d)


def synth_big(n, d, c, dev, seed):
    """Like synth() but generated in chunks (no second copy of a 51 GB matrix)."""
    import torch

    g = torch.Generator(device=dev).manual_seed(seed)
    y = torch.randint(0, c, (n,), generator=g, device=dev)
    means = 0.2 * torch.randn(c, d, generator=g, device=dev) / d**0.5
    scale = (0.5 + torch.rand(c, 1, generator=g, device=dev)) / d**0.5
    X = torch.empty(n, d, device=dev)
    step = 500_000
    for lo in range(0, n, step):
        yy = y[lo:lo + step]
        X[lo:lo + step] = torch.randn(yy.numel(), d, generator=g, device=dev) * scale[yy] + means[yy]
    return X, y


def hp1_table(dev, peaks, tf32_peak, reps=5):
    """class_statistics at the other BASELINE configs, single GPU: ms per call, samples/s, the Gram
    kernel's time, and the fraction of the co-bound roofline max(t_hbm, t_tensor) / t_measured with
    t_hbm = 4 N D / HBM peak (one read of X) and t_tensor = executed TF32 flop / TF32 peak."""
    import torch

    from sqfa_b200 import statistics as S

    ops = S._cuda_ops()
    rows = {}
    for name, (n, d, c, _k) in CONFIGS.items():
        need = n * d * 4 * 1.15 + 3 * c * d * d * 4 + (2 << 30)
        free, _ = torch.cuda.mem_get_info()
        if free < need:
            rows[name] = {"skipped": f"needs {need / 1e9:.0f} GB of free device memory, {free / 1e9:.0f} GB free"}
            continue
        X, y = (synth_big if n * d * 4 > (8 << 30) else synth)(n, d, c, dev, 77)
        # warm-up long enough for the clocks to be up (the legs before this one leave the GPU idle for tens
        # of seconds of CPU work), then a timed region of at least ~50 ms
        torch.cuda.synchronize()
        t_w = time.perf_counter()
        n_w = 0
        while n_w < 3 or (time.perf_counter() - t_w < 0.1 and n_w < 500):
            S.class_statistics(X, y)
            n_w += 1
        torch.cuda.synchronize()
        per_call = (time.perf_counter() - t_w) / n_w
        reps_t = int(min(200, max(reps, 0.05 / max(per_call, 1e-5))))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps_t):
            st = S.class_statistics(X, y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps_t
        ops.gram_events = []  # instrumented repetition: events around the Gram launch on its stream
        for _ in range(reps):
            S.class_statistics(X, y)
        torch.cuda.synchronize()
        gram_ms = sum(a.elapsed_time(b) for a, b in ops.gram_events) / max(len(ops.gram_events), 1)
        ops.gram_events = None
        executed = gram_executed_flop(n, d)
        t_hbm = 4.0 * n * d / (peaks["hbm_gbs"] * 1e9) * 1e3
        t_tensor = executed / (tf32_peak * 1e12) * 1e3
        bound = "hbm" if t_hbm >= t_tensor else "tensor"
        rows[name] = {
            "N": n, "D": d, "classes": c, "ms_per_call": ms, "samples_per_s": n / (ms * 1e-3),
            "gram_kernel_ms": gram_ms, "bound": bound, "roofline_ms": max(t_hbm, t_tensor),
            "frac_call": max(t_hbm, t_tensor) / ms, "frac_gram_kernel": max(t_hbm, t_tensor) / gram_ms,
            "hbm_gbs_call": 4.0 * n * d / (ms * 1e-3) / 1e9,
            "executed_tflops_gram_kernel": executed / (gram_ms * 1e-3) / 1e12,
            "algorithmic_tflops_call": 2.0 * n * d * d / (ms * 1e-3) / 1e12,
        }
        del X, y, st
        torch.cuda.empty_cache()
    return rows


def time_closure(plan, n_eval):
    import torch

    for _ in range(3):
        plan()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_eval):
        plan()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n_eval


def closure_kernel_times(stats, model, dist_id, reps=7):
    """The stages of one closure evaluation timed one by one through the step-by-step entry points (the
    fused call runs the same kernels back to back): projection (stream + finish), per-class
    factorisation, the pair kernel, the projection adjoint. ms each."""
    import torch

    from sqfa_b200 import _ops

    S_, M_ = stats["covariances"], stats["means"]
    F = model.filters.detach().contiguous()
    C, k = S_.shape[0], F.shape[0]
    P = C * (C - 1) // 2

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # the wrappers allocate their outputs and workspaces between the events (host time, sometimes a cudaMalloc
    # of hundreds of MB): two warm-up rounds, then the MEDIAN over the repetitions, not the mean
    keys = ("project_fwd", "class_factor", "pair", "project_bwd")
    samples = {key: [] for key in keys}
    for it in range(reps + 2):
        marks = [ev() for _ in range(5)]
        marks[0].record()
        T, Psi, Mu = _ops.project_fwd_raw(S_, M_, F)
        marks[1].record()
        E = _ops.embed_fwd_raw(Psi, Mu, 0.01, dist_id)
        W, _ = _ops.class_factor_raw(E, dist_id)
        marks[2].record()
        gE = torch.zeros_like(E)
        loss = torch.zeros(2, device=E.device)
        _ops.pair_raw(W, W, C, C, E.shape[-1], dist_id, True, weight=-1.0 / P, loss=loss, gEa=gE, gEb=gE)
        marks[3].record()
        gPsi, gMu = _ops.embed_bwd_raw(gE, Mu, k, dist_id)
        _ops.project_bwd_raw(gPsi, gMu, T, M_)
        marks[4].record()
        torch.cuda.synchronize()
        if it < 2:
            continue  # warm-up
        for key, a, b in zip(keys, marks[:-1], marks[1:]):
            samples[key].append(a.elapsed_time(b))
    return {key: sorted(v)[len(v) // 2] for key, v in samples.items()}


def fit_leg(stats, d, c, k, dev, group, world, peaks, n_eval, epochs, label, cpu_evals=0, kernels=False):
    """Closure evaluations/s and LBFGS epochs/s of SQFA.fit (Fisher-Rao lower bound, PCA init,
    feature_noise 0.01, lr 0.1, default L-BFGS, atol 0) on the given class statistics; with a process
    group the pair list of every evaluation AND the classes of the projection are sharded over the ranks."""
    import torch

    from sqfa_b200 import _ops
    from sqfa_b200.model import SQFA

    model = SQFA(n_dim=d, feature_noise=0.01, n_filters=k)
    model.fit_pca(data_statistics=stats)
    model = model.to(dev)
    model._process_group = group if world > 1 else None
    # one closure evaluation = loss forward + gradient w.r.t. the raw filter parameter, the way
    # fitting_loop evaluates it (native, no autograd graph for the sphere constraint)
    plan = model._fused_direct_plan(stats)
    closure_ms = time_closure(plan, n_eval)
    model._process_group = None
    P = c * (c - 1) // 2
    # one throw-away epoch on a copy: the first torch.optim.LBFGS in a process imports half of
    # torch (~2.5 s), which is not part of an epoch
    warm = SQFA(n_dim=d, feature_noise=0.01, n_filters=k, filters=model.filters.detach().clone())
    pg = group if world > 1 else None
    with contextlib.redirect_stdout(sys.stderr):  # fitting_loop prints like the reference does
        warm.fit(data_statistics=stats, max_epochs=1, atol=0.0, show_progress=False, process_group=pg)
    torch.cuda.synchronize()
    F_start = model.filters.detach().clone()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sys.stderr):
        losses, _ = model.fit(data_statistics=stats, max_epochs=epochs, atol=0.0, show_progress=False,
                              return_loss=True, process_group=pg)
    torch.cuda.synchronize()
    fit_s = time.perf_counter() - t0
    evals = int(getattr(model, "_last_fit_evaluations", 0)) or None
    out = {
        "model": f"SQFA fisher_rao_lower_bound, C={c}, D={d}, n_filters={k}, pca init, feature_noise=0.01, {label}",
        "pairs_per_closure": P, "closure_ms": closure_ms, "closure_evals_per_s": 1e3 / closure_ms,
        "pairs_per_s": P / (closure_ms * 1e-3), "epochs": epochs, "fit_s": fit_s, "epochs_per_s": epochs / fit_s,
        "closure_evals_in_fit": evals, "loss_first_last": [float(losses[0]), float(losses[-1])],
        "algorithmic_bytes_per_closure": 4 * c * d * d,
        "closure_hbm_gbs": 4 * c * d * d / (closure_ms * 1e-3) / 1e9,
        "pair_list": (f"pairs AND the projection's classes sharded over {world} ranks; three small all-reduces per "
                      "evaluation ([Psi | mu'] partials, [gPsi | gMu | loss, flag], dF)") if world > 1 else "single GPU",
    }
    if kernels:
        kt = closure_kernel_times(stats, model, _ops.DIST_FR)
        proj_bytes = 4.0 * c * d * d
        out["kernels_ms"] = kt
        out["roofline_project"] = {
            "bound": "hbm", "kernel": "project_stream_kernel + project_finish_kernel",
            "achieved": proj_bytes / (kt["project_fwd"] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": proj_bytes / (kt["project_fwd"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "algorithmic_bytes_per_launch": proj_bytes, "kernel_ms": kt["project_fwd"]}
        pair = {"bound": "sm instruction issue", "kernel": "pair kernel (one-sided Jacobi, registers)",
                "pairs_per_s": P / (kt["pair"] * 1e-3), "kernel_ms": kt["pair"], "issue_active_pct": None}
        prof = os.path.join(ROOT, "profiles", "pair_issue.json")
        if os.path.exists(prof):  # ncu --set full summary of the current kernel, refreshed with the kernel
            with open(prof) as f:
                pj = json.load(f)
            pair["issue_active_pct"] = pj.get("issue_active_pct")
            pair["ncu_capture"] = pj.get("capture")
        out["roofline_pair"] = pair
    if cpu_evals > 0:
        from oracle import sqfa_oracle as O

        torch.set_num_threads(os.cpu_count() or 1)
        cstats = {kk: v.cpu() for kk, v in stats.items() if kk in ("means", "covariances")}
        F0 = F_start.cpu()
        t0 = time.perf_counter()
        for _ in range(cpu_evals):
            cl, _, _ = O.loss_and_grad("full", cstats, F0, noise=0.01)
        cpu_s = (time.perf_counter() - t0) / cpu_evals
        out["cpu_closure_s"] = cpu_s
        out["cpu_closure_evals_per_s"] = 1.0 / cpu_s
        out["cpu_cores"] = torch.get_num_threads()
        out["cpu_loss_at_start"] = float(cl)
        out["closure_speedup_vs_cpu"] = cpu_s / (closure_ms * 1e-3)
        if evals:
            est = evals * cpu_s
            out["cpu_fit_s_extrapolated"] = est
            out["cpu_fit_extrapolation"] = (f"{cpu_evals} CPU closure evaluations timed ({cpu_s:.1f} s each, "
                                            f"{torch.get_num_threads()} threads) x the {evals} evaluations of the GPU fit "
                                            "(BASELINE.md section 3)")
            out["fit_speedup_vs_cpu"] = est / fit_s
            out["target_fit_speedup"] = 100.0
    return out


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
        "config": workload_config(max(int(args.gpus), 1)),  # this repo's arm's config at the same N
        "host_arm": f"reference CPU path, {torch.get_num_threads()} host threads (rank 0 only); no device, l2 not applicable",
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
        # N > 1: the result of the job is sharded by class over the ranks ("ReduceScatter by
        # class"); `value_replicated_output` below is the same job with the full result on every rank
        return S.class_statistics(X, y, group=group, shard_output=world > 1)

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
    ms_rep = None
    if world > 1:  # the same K steps with replicated output (NCCL all-reduce of the packed Gram partials)
        for _ in range(3):
            S.class_statistics(X, y, group=group)
        sync_all()
        e0.record()
        for _ in range(args.steps):
            rep_stats = S.class_statistics(X, y, group=group)
        e1.record()
        sync_all()
        ms_rep = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms_rep, op=dist.ReduceOp.MAX)
        del rep_stats
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
    # full statistics on the device: input of the fit leg below (rank 0's own shard when the job's are sharded)
    stats_full = stats if world == 1 else (S.class_statistics(X, y) if rank == 0 else None)
    # with several ranks every rank finalises and downloads ITS share of the classes (shard_output): the
    # ranks' host buffers together hold the job's result, nothing is computed or copied twice
    sharded = world > 1
    probe = S.class_statistics(X, y, group=group, shard_output=sharded)
    keys = ("means", "covariances", "second_moments")
    out_h = [{key: torch.empty(probe[key].shape, dtype=probe[key].dtype).pin_memory() for key in keys}
             for _ in range(nbuf)]
    del probe
    Xd, yd = [torch.empty_like(X) for _ in range(nbuf)], [torch.empty_like(y) for _ in range(nbuf)]

    def upload(i):  # step i's inputs, pinned host -> device, on the stream of its buffer set
        b = i % nbuf
        with torch.cuda.stream(streams[b]):
            yd[b].copy_(yh, non_blocking=True)
            Xd[b].copy_(Xh, non_blocking=True)

    def compute_and_download(i):
        b = i % nbuf
        with torch.cuda.stream(streams[b]):
            st = S.class_statistics(Xd[b], yd[b], group=group, shard_output=sharded)
            for key in keys:
                out_h[b][key].copy_(st[key], non_blocking=True)
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
    ems = torch.tensor([e2e_med], device=dev)
    nbytes = torch.tensor([Xh.numel() * 4 + yh.numel() * 8, sum(v.numel() * 4 for v in out_h[0].values())],
                          dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        dist.all_reduce(nbytes)  # whole-job bytes per step: all ranks' uploads and downloads
    e2e_value = n * world * e2e_steps / (float(ems) * 1e-3)
    h2d, d2h = int(nbytes[0]), int(nbytes[1])
    del last, Xd, yd, out_h
    stats = stats_full

    gc.enable()
    peaks = load_peaks()
    # Dense TF32 peak: measured in this run with cuBLAS (every rank measures, which keeps the ranks in
    # step). The roofline denominator is the LARGER of that and half the measured bf16 peak of
    # MEASURED_PEAKS.json (the hardware ratio): cuBLAS TF32 does not reach the pipe's peak on this part
    # (the Gram kernel runs above it), and a fraction against the smaller figure would flatter the kernel.
    tf32_cublas = measure_tf32_peak(dev)
    tf32_peak = max(tf32_cublas, peaks["bf16_tflops"] / 2.0)
    # ---- second hot path on the headline config: SQFA fit on the statistics just computed
    fit = fit_leg(stats, d, c, k, dev, None, 1, peaks, n_eval=50, epochs=5, label="configs[1]") if rank == 0 else None
    del stats, stats_full
    torch.cuda.empty_cache()
    # ---- second hot path on configs[3] (C = 1000): all ranks, pair list sharded when world > 1
    n4, d4, c4, k4 = CONFIGS["c4"]
    X4, y4 = synth(n4, d4, c4, dev, 4321)  # same seed on every rank: the fit needs replicated statistics
    stats4 = S.class_statistics(X4, y4)
    del X4, y4
    torch.cuda.empty_cache()
    fit_c4 = fit_leg({kk: stats4[kk] for kk in ("means", "covariances")}, d4, c4, k4, dev, group, world, peaks,
                     n_eval=10, epochs=2, label="configs[3]", cpu_evals=2 if (world == 1 and not args.no_cpu) else 0,
                     kernels=(rank == 0))
    del stats4
    torch.cuda.empty_cache()
    table = hp1_table(dev, peaks, tf32_peak) if (world == 1 and not args.no_table) else None

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    executed = gram_executed_flop(n, d)
    achieved = executed / (gram_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "gram_tf32x3_kernel", "achieved": achieved, "peak": tf32_peak,
                "unit": "TFLOP/s", "frac": achieved / tf32_peak, "traffic": None,
                "kernel_ms": gram_ms, "executed_flop_per_launch": executed,
                "algorithmic_flop_per_launch": 2.0 * n * d * d,
                "peak_source": "max(cuBLAS TF32 matmul 8192^3 measured in this run (best of 10), "
                               "MEASURED_PEAKS.json bf16_tflops / 2)",
                "tf32_cublas_measured_in_run": tf32_cublas, "bf16_over_2": peaks["bf16_tflops"] / 2.0,
                "frac_of_cublas_tf32": achieved / tf32_cublas}
    prof = os.path.join(ROOT, "profiles", "gram_traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            roofline["traffic"] = json.load(f).get("dram_bytes_per_launch")

    cpu = None
    if world == 1 and not args.no_cpu:
        rate, dt = cpu_class_statistics_rate(n, d, c, reps=3)
        import torch as _t

        cpu = {"value": rate, "unit": "samples/s", "cores": _t.get_num_threads(), "kind": "port",
               "sample": f"full workload N={n}, 3 repetitions, {dt:.2f} s each"}

    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (3xTF32 tensor-core split, fp32 accumulate)",
        "data": "synthetic",
        "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "repetitions_ms": [round(x, 2) for x in e2e_ms], "reported": "median repetition",
                "overlap": f"{nbuf} CUDA streams / buffer sets, inputs of step i+1 uploaded while step i computes and downloads",
                "bytes": "whole job (all ranks); with N > 1 every rank downloads its share of the classes"},
        "gpu_launches": (KERNELS_PER_STEP if world == 1 else KERNELS_PER_STEP_MULTI) * args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "fit": fit, "fit_c4": fit_c4,
        "hp1_table": table, "tf32_peak_measured_tflops": tf32_cublas,
    }
    if ms_rep is not None:
        line["value_replicated_output"] = {
            "value": n * world * args.steps / (float(ms_rep) * 1e-3), "unit": "samples/s",
            "ms_per_step": float(ms_rep) / args.steps,
            "note": "full statistics on every rank: NCCL all-reduce of the packed Gram partials behind the kernel"}
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baselines (profiling runs)")
    ap.add_argument("--no-table", action="store_true", help="skip the per-config class_statistics table")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
