"""Runs every tiling candidate of the self-test GEMM shapes one by one and reports the ones whose kernel raised the
barrier-protocol flag (vy_gemm_poisoned) or produced a wrong result (development aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vyomai_b200 import _lib, gemm_tune, ops  # noqa: E402

gemm_tune.ENABLED = False


def main():
    lib = _lib.lib()
    torch.manual_seed(0)
    dt = torch.bfloat16
    bad = 0
    shapes = [(256, 256, 128), (384, 768, 768), (1024, 3072, 768), (1000, 520, 264), (8192, 768, 3072), (300, 50265, 768),
              (512, 768, 768), (784, 768, 768), (768, 768, 8192), (1280, 768, 8192), (2304, 768, 12608), (768, 3072, 4096),
              (520, 264, 5000)]
    for (M, N, K) in shapes:
        K8 = (K + 7) // 8 * 8
        a = torch.randn(M, K8, device="cuda", dtype=dt)
        b = torch.randn(N, K8, device="cuda", dtype=dt) / K8 ** 0.5
        ref = a.float() @ b.float().t()
        for amn in (0, 1):
            for bmn in (0, 1):
                if (amn or bmn) and (M % 8 or N % 8):
                    continue
                aa = a.t().contiguous().t() if amn else a
                bb = b.t().contiguous().t() if bmn else b
                for pair in (0, 1):
                    for bn in (128, 192, 256):
                        for sp in (0, 1, 2, 3, 4, 6, 8):
                            lib.vy_gemm_tune_override(pair, bn, sp)
                            ldo = (N + 7) // 8 * 8
                            out = torch.empty(M, ldo, device="cuda", dtype=dt)[:, :N]
                            try:
                                ops.gemm(aa, bb, out=out, allow_split_k=sp > 0)
                            except _lib.VyomError:
                                continue
                            torch.cuda.synchronize()
                            p = lib.vy_gemm_poisoned()
                            err = (out.float() - ref).abs().max().item()
                            if p != 0 or not err < 0.15:
                                bad += 1
                                print(f"BAD {M}x{N}x{K8} a_mn={amn} b_mn={bmn} pair={pair} bn={bn} splits={sp}: poisoned={p} maxerr={err:.3f}", flush=True)
        print(f"shape {M}x{N}x{K8} done", flush=True)
    lib.vy_gemm_tune_override(-1, 0, 0)
    print("bad configs:", bad)


if __name__ == "__main__":
    main()
