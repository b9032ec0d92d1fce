"""Summarise an `ncu --set full` report of one kernel as a small JSON dict (what profiles/*.json hold).

Usage: python tools/ncu_summary.py REPORT.ncu-rep KERNEL_REGEX ["free-text description"]
Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv` (stall samples per SASS
instruction); prints JSON to stdout.
"""
import csv
import io
import json
import subprocess
import sys

RAW = {
    "duration_ms": ("gpu__time_duration.sum", 1.0),
    "tensor_pipe_active_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0),
    "issue_active_pct": ("sm__inst_issued.avg.pct_of_peak_sustained_active", 1.0),
    "dram_read_MB": ("dram__bytes_read.sum", 1.0),
    "dram_write_MB": ("dram__bytes_write.sum", 1.0),
    "l2_hit_pct": ("lts__t_sector_hit_rate.pct", 1.0),
    "regs": ("launch__registers_per_thread", 1.0),
    "dyn_smem_KB": ("launch__shared_mem_per_block_dynamic", 1.0),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "dram_throughput_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "l1tex_throughput_pct": ("l1tex__throughput.avg.pct_of_peak_sustained_active", 1.0),
    "warp_instructions": ("smsp__inst_executed.sum", 1.0),
}


def run(args):
    return subprocess.run(args, check=True, capture_output=True, text=True).stdout


def main():
    rep, regex = sys.argv[1], sys.argv[2]
    what = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + regex]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {"capture": rep.split("/")[-1].replace(".ncu-rep", ""), "what": what, "kernel": vals[hdr.index("Kernel Name")][:80]}
    for key, (metric, scale) in RAW.items():
        if metric in hdr:
            i = hdr.index(metric)
            try:
                out[key] = round(float(vals[i].replace(",", "")) * scale, 3)
                out[key + "_unit"] = units[i]
            except ValueError:
                pass
    # durations are reported by ncu in a unit of its choosing: normalise to ms
    u = out.get("duration_ms_unit", "ms")
    if "duration_ms" in out:
        out["duration_ms"] = round(out["duration_ms"] * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0), 5)
    out = {k: v for k, v in out.items() if not (k.endswith("_unit") and v in ("", "%", "ms", "us", "ns", "Mbyte", "register/thread"))}
    src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex]))))
    h = src[1]
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    tot = {c: 0 for c in stalls}
    nsamp = 0
    for r in src[2:]:
        if len(r) != len(h):
            continue
        nsamp += int(r[h.index("# Samples")] or 0)
        for c in stalls:
            tot[c] += int(r[h.index(c)] or 0)
    out["stall_samples_total"] = nsamp
    out["stall_samples_top"] = dict(sorted(tot.items(), key=lambda kv: -kv[1])[:6])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
