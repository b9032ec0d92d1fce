"""Pin the tcgen05 operand-layout assumptions on hardware (run on a B200 through gpurun).

Every hypothesis (smem writer mode, LBO, SBO, swizzle type, major-ness bits, K-step stride) runs
in its own subprocess, so a trap in one does not poison the CUDA context of the others. Results go
to gpurun_out/umma_probe.json: max |D - A^T B| for each hypothesis (exact products: the inputs
are small integers, so any mismatch is a layout error, not rounding).
"""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_one(cfg):
    import torch

    import ctypes

    from sqfa_b200 import _lib, build

    lib = ctypes.CDLL(build.PROBE_LIB if os.path.exists(build.PROBE_LIB) else build.build_probe())
    u32, i32, ptr = ctypes.c_uint32, ctypes.c_int32, ctypes.c_void_p
    lib.sqfa_debug_umma_probe.restype = ctypes.c_int
    lib.sqfa_debug_umma_probe.argtypes = [ptr, ptr, ptr, i32, i32, i32, u32, u32, u32, u32, u32, u32, ptr]
    K, N = cfg["K"], cfg["N"]
    g = torch.Generator().manual_seed(1)
    A = torch.randint(-4, 5, (K, 128), generator=g).float().cuda()
    B = torch.randint(-4, 5, (K, N), generator=g).float().cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = lib.sqfa_debug_umma_probe(
        _lib.ptr(A), _lib.ptr(B), _lib.ptr(D), K, N, cfg["mode"], cfg["lbo"], cfg["sbo"], cfg["layout"],
        cfg["a_major"], cfg["b_major"], cfg["kstep"], _lib.stream_ptr(),
    )
    torch.cuda.synchronize()
    if cfg["mode"] >= 2:
        tab = D[:, :8] if cfg["mode"] == 2 else D[:8, :].T  # [mn][k] -> word index fetched
        print(json.dumps({"rc": rc, "table": tab.long().cpu().tolist()}))
        return
    ref = A.double().T @ B.double()
    err = (D.double() - ref).abs().max().item()
    print(json.dumps({"rc": rc, "max_err": err}))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_one(json.loads(sys.argv[2]))
        return
    hyps = []
    # what gram.cu uses: K-major, no swizzle, LBO = next 4 samples, SBO = next 8 columns
    for K in (8, 16, 32):
        for N in (32, 128, 256):
            hyps.append(dict(name="k_major_noswz", K=K, N=N, mode=1, lbo=128 * 16, sbo=128, layout=0, a_major=0,
                             b_major=0, kstep=2 * 128 * 16))
    # decode probes: which smem word does the hardware fetch for A(k, m) / B(k, n)?
    for layout in (0, 2, 4, 6):
        for major in (1, 0):
            for lbo, sbo in ((2048, 1024), (1024, 2048), (4096, 512), (256, 4096)):
                hyps.append(dict(name=f"decodeA_l{layout}_mj{major}_lbo{lbo}_sbo{sbo}", K=8, N=16, mode=2, lbo=lbo,
                                 sbo=sbo, layout=layout, a_major=major, b_major=0, kstep=0))
    hyps.append(dict(name="decodeB_l2_mj1_lbo2048_sbo1024", K=8, N=256, mode=3, lbo=2048, sbo=1024, layout=2,
                     a_major=0, b_major=1, kstep=0))
    out = []
    for h in hyps:
        try:
            res = subprocess.run([sys.executable, __file__, "--one", json.dumps(h)], capture_output=True, text=True,
                                 timeout=120)
            line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
            r = json.loads(line) if line.startswith("{") else {"rc": res.returncode, "stderr": res.stderr[-400:]}
        except subprocess.TimeoutExpired:
            r = {"timeout": True}
        r.update(h)
        out.append(r)
        print({k: v for k, v in r.items() if k != "table"}, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "umma_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
