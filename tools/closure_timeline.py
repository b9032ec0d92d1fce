"""Per-kernel timeline of one fused closure evaluation from an ncu launch list (warm caches).
Usage (GPU box): ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
                 --log-file gpurun_out/cl.csv python tools/run_closure_once.py c2 4
       then:     python tools/closure_timeline.py gpurun_out/cl.csv
"""
import csv, sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mn, mv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
seq = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) > mv and r[mn] == "gpu__time_duration.sum"]
starts = [i for i, (k, v) in enumerate(seq) if "project_stream" in k or "project_partial" in k or "project_tc" in k]
i0 = starts[-1]
tot = 0.0
for k, v in seq[i0:]:
    name = k.split("(")[0].split("::")[-1][:60]
    print(f"{v / 1000:9.1f} us  {name}")
    tot += v
print(f"sum {tot / 1000:.1f} us over {len(seq) - i0} launches")
