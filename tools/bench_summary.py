"""Prints the headline numbers and the per-family kernel table of a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "gemm frac",
      round(d["roofline"]["frac"], 3), "sum kernels", d["kernel_breakdown_sum_ms"], "launches", d.get("gpu_launches"))
for k, v in d["kernel_breakdown_ms"].items():
    print(f"  {k:24s} {v['calls']:6.0f} {v['ms']:.3f}")
if "decode" in d:
    print("decode", json.dumps(d["decode"])[:400])
