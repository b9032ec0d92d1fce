# Summarise the second set of round-2 ncu captures (gpurun_out/r2b_*) into profiles/ (tracked).
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import json, subprocess, os
caps = [("r2b_gram_c2", "gram_tf32x3_kernel", "c2 class_statistics (N=50000, D=3072, C=10): tcgen05 cta_group::2 3xTF32 Gram"),
        ("r2b_pair_c4", "pair_cp_kernel", "c4 closure (C=1000, m=17, 499500 pairs): column-pair Jacobi, 3 problems per warp, 3x3 pair tiles"),
        ("r2b_pair_c5", "pair_cp_kernel", "c5 closure (C=100, m=33, 4950 pairs): column-pair Jacobi, one problem per warp"),
        ("r2b_ptc_c5", "project_tc", "c5 closure (C=100, D=1024, k=32): tcgen05 projection T_c^T = S_c F^T, 3xTF32")]
out = []
for rep, regex, what in caps:
    path = f"gpurun_out/{rep}.ncu-rep"
    if not os.path.exists(path):
        continue
    r = subprocess.run(["python", "tools/ncu_summary.py", path, regex, what], capture_output=True, text=True)
    if r.returncode == 0:
        out.append(json.loads(r.stdout))
json.dump(out, open("profiles/r02b_kernels_ncu.json", "w"), indent=1)
by = {d["capture"]: d for d in out}
if "r2b_gram_c2" in by:
    g = by["r2b_gram_c2"]
    rd, wr = g["dram_read_MB"] * 1e6, g["dram_write_MB"] * 1e6
    json.dump({"kernel": "gram_tf32x3_kernel", "workload": "c2 (N=50000, D=3072, C=10)",
               "capture": "r2b_gram_c2 (ncu --set full --clock-control none), profiles/r02b_kernels_ncu.json",
               "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
               "algorithmic_bytes_per_launch": 818872320}, open("profiles/gram_traffic.json", "w"), indent=1)
if "r2b_pair_c4" in by:
    p = by["r2b_pair_c4"]
    json.dump({"kernel": p["kernel"], "workload": "c4 closure (C=1000, m=17, 499500 pairs)",
               "capture": "r2b_pair_c4 (ncu --set full --clock-control none), profiles/r02b_kernels_ncu.json",
               "issue_active_pct": p["issue_active_pct"], "warp_instructions": p["warp_instructions"],
               "duration_ms": p["duration_ms"]}, open("profiles/pair_issue.json", "w"), indent=1)
PY
for f in bench_launches cl_c1 cl_c2 cl_c3 cl_c4 cl_c5 st_c1 st_c3 st_c4; do
  [ -f gpurun_out/r2b_$f.csv ] && cp gpurun_out/r2b_$f.csv profiles/r02b_${f}_ncu.csv
done
{
  echo "# cuobjdump -sass sqfa_b200/libsqfa_b200.so | grep -c <mnemonic>   (sm_100a SASS of the shipped library)"
  for m in UTCHMMA.2CTA UTCHMMA LDTM UTCBAR SETMAXREG FFMA2 FMUL2 DFMA SHFL REDG RED.E ATOMG UTMALDG UTMASTG SYNCS; do
    printf "%-14s %s\n" "$m" "$(cuobjdump -sass sqfa_b200/libsqfa_b200.so | grep -c "$m")"
  done
  echo
  echo "# per kernel: tcgen05.mma (UTCHMMA) instructions"
  cuobjdump -sass sqfa_b200/libsqfa_b200.so | awk '/Function :/ {fn=$3} /UTCHMMA/ {c[fn]++} END {for (f in c) print c[f], f}' | sort -rn | c++filt | cut -c1-140
} > profiles/r02_sass_mnemonics.txt
ls profiles
