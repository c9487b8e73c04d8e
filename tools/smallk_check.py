"""Repeatability / correctness of vy_gemm on weight-gradient shapes with very few tokens (K < one k-block), every
tiling candidate pinned in turn (development aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vyomai_b200 import _lib, gemm_tune, ops  # noqa: E402

gemm_tune.ENABLED = False


def main():
    lib = _lib.lib()
    torch.manual_seed(0)
    bad = 0
    for (M, N, K) in [(256, 1024, 52), (1024, 256, 52), (256, 256, 52), (768, 256, 56), (256, 1024, 8), (256, 1024, 120), (5003, 256, 52)]:
        at = torch.randn(K, M, device="cuda", dtype=torch.bfloat16)
        bt = torch.randn(K, N, device="cuda", dtype=torch.bfloat16)
        ref = at.float().t() @ bt.float()
        for pair in (0, 1):
            for bn in (128, 192, 256):
                lib.vy_gemm_tune_override(pair, bn, 0)
                outs = []
                for rep in range(6):
                    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
                    junk = torch.randn(64 << 20, device="cuda")  # disturb L2 / timing between repeats
                    ops.gemm(at.t(), bt.t(), out=out, allow_split_k=True)
                    del junk
                    outs.append(out)
                torch.cuda.synchronize()
                p = lib.vy_gemm_poisoned()
                err = max(float((o.float() - ref).abs().max()) for o in outs)
                same = all(torch.equal(outs[0], o) for o in outs[1:])
                if p or not same or not err < 0.25:
                    bad += 1
                    print(f"BAD {M}x{N}x{K} pair={pair} bn={bn}: poisoned={p} repeatable={same} maxerr={err:.3f}", flush=True)
        print(f"{M}x{N}x{K} done", flush=True)
    lib.vy_gemm_tune_override(-1, 0, 0)
    print("bad:", bad)


if __name__ == "__main__":
    main()
