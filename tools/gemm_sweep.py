"""Sweep kernel flavour x tile width x K split for every GEMM shape of the captioner step (development aid).

    python tools/gemm_sweep.py [--iters 10] [--json gpurun_out/gemm_sweep.json]

Uses the vy_gemm_tune_override hook to pin (single CTA | CTA pair, BN, splits) in-process, times each with CUDA events
over L2-exceeding buffer rotations, and prints per shape the automatic choice next to the best measured one — the data
the time model in gemm.cu (choose_tiling) is calibrated against.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gemm_bench  # noqa: E402
from vyomai_b200 import _lib, gemm_tune  # noqa: E402

gemm_tune.ENABLED = False  # the sweep pins every candidate itself


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    lib = _lib.lib()
    out = []
    for c in gemm_bench.bench_cases():
        name, M, N, K, a_mn, b_mn, epi = c
        if args.only and args.only not in name:
            continue
        lib.vy_gemm_tune_override(-1, 0, 0)
        auto = gemm_bench.case(*c, iters=args.iters, cublas_too=False)["us"]
        rows = []
        splits = [0] if epi not in ("accum", "splitk_ok") else [1, 2, 3, 4, 6]
        for pair in (0, 1):
            for bn in (128, 192, 256):
                if pair and b_mn and bn == 192:
                    continue
                for sp in splits:
                    lib.vy_gemm_tune_override(pair, bn, sp)
                    try:
                        t = gemm_bench.case(*c, iters=args.iters, cublas_too=False)["us"]
                    except Exception as e:  # a forced split that the shape does not admit
                        continue
                    rows.append(dict(pair=pair, bn=bn, splits=sp, us=round(t, 1)))
        lib.vy_gemm_tune_override(-1, 0, 0)
        rows.sort(key=lambda r: r["us"])
        best = rows[0]
        bs = [r for r in rows if r["pair"] == 0][0]
        bp = [r for r in rows if r["pair"] == 1][0]
        print(f"{name:20s} M={M:6d} N={N:6d} K={K:6d} auto {auto:7.1f} us | best {best['us']:7.1f} (pair={best['pair']} bn={best['bn']} sp={best['splits']})"
              f" | single {bs['us']:7.1f} (bn={bs['bn']} sp={bs['splits']}) | pair {bp['us']:7.1f} (bn={bp['bn']} sp={bp['splits']})", flush=True)
        out.append(dict(name=name, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, epi=epi, auto_us=round(auto, 1), rows=rows))
    print("poisoned", lib.vy_gemm_poisoned())
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
