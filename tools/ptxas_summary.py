"""Summarise the `-Xptxas -v` logs of the last build (sqfa_b200/_build/*.ptxas.log) into one table:
kernel, registers, spill bytes, static shared memory -- the static evidence the profiling recipe asks
to look at before spending GPU time. Usage: python tools/ptxas_summary.py > profiles/r02_ptxas_resources.txt"""

import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for n in out:
        n = re.sub(r"\(anonymous namespace\)::", "", n)
        n = re.sub(r"^void ", "", n)
        n = re.sub(r"\(.*$", "", n)  # drop the parameter list
        short.append(n.replace("sqfa::", ""))
    return short


def main():
    rows = []
    for log in sorted(glob.glob(os.path.join(ROOT, "sqfa_b200", "_build", "*.ptxas.log"))):
        text = open(log).read()
        for m in re.finditer(
            r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
            r"(\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?",
            text, flags=re.S,
        ):
            name, stack, st, ld, regs, smem = m.groups()
            rows.append((os.path.basename(log)[:-10] + ".cu", name, int(regs), int(stack), int(st), int(ld), int(smem or 0)))
    names = demangle([r[1] for r in rows])
    print("# ptxas -v (nvcc 12.9, -O3, sm_100a) per kernel; dynamic shared memory is set at launch and not listed")
    print(f"{'file':15s} {'kernel':58s} {'regs':>5s} {'stack':>6s} {'spill st':>9s} {'spill ld':>9s} {'static smem':>12s}")
    for (f, _, regs, stack, st, ld, smem), n in zip(rows, names):
        print(f"{f:15s} {n[:58]:58s} {regs:5d} {stack:6d} {st:9d} {ld:9d} {smem:12d}")


if __name__ == "__main__":
    main()
