"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per kernel name launches, total / mean
device time and share. Usage: python tools/summarize_launches.py gpurun_out/launches.csv [out.md]"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(CUtensorMap_st.*", "", name)
    name = name.replace("void ", "").replace("vy::", "").replace("__nv_bfloat16", "bf16").replace("(int)", "").replace("(bool)", "")
    return name[:110]


def main():
    path = sys.argv[1]
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rd:
        v = float(r[ival].replace(",", ""))
        if r[iunit] == "us":
            v *= 1e3
        elif r[iunit] == "ms":
            v *= 1e6
        a = agg[short(r[iname])]
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    out = ["| kernel | launches | total us | mean us | share |", "|---|---:|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {t / 1e3:.1f} | {t / 1e3 / n:.1f} | {100 * t / total:.1f}% |")
    out.append(f"| **total** | {sum(a[0] for a in agg.values())} | {total / 1e3:.1f} | | 100% |")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
