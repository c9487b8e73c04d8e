"""First-contact check of the CTA-pair (cta_group::2) GEMM kernels: a few small products against torch, with the
barrier-protocol flag (vy_gemm_poisoned) read after each one. Run with VY_GEMM_PAIR=1. Exits non-zero at the first
wrong or poisoned result so that a driver script can stop before running anything bigger."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vyomai_b200 import _lib, ops  # noqa: E402


def main():
    torch.manual_seed(0)
    lib = _lib.lib()
    bad = 0
    for (M, N, K, amn, bmn) in [(256, 256, 128, 0, 0), (256, 128, 64, 0, 0), (384, 192, 256, 0, 0), (512, 768, 768, 0, 0),
                                (512, 768, 768, 0, 1), (512, 768, 768, 1, 0), (768, 512, 2048, 1, 1), (8192, 3072, 768, 0, 0)]:
        a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) / K ** 0.5
        ref = a.float() @ b.float().t()
        aa = a.t().contiguous().t() if amn else a
        bb = b.t().contiguous().t() if bmn else b
        out = ops.gemm(aa, bb)
        torch.cuda.synchronize()
        err = (out.float() - ref).abs().max().item()
        rows = (out.float() - ref).abs().amax(dim=1)
        cols = (out.float() - ref).abs().amax(dim=0)
        p = lib.vy_gemm_poisoned()
        ok = err < 0.1 and p == 0
        print(f"{M}x{N}x{K} a_mn={amn} b_mn={bmn}: maxerr={err:.4f} poisoned={p} {'ok' if ok else 'BAD'}", flush=True)
        if not ok:
            br = (rows > 0.1).nonzero().flatten().tolist()
            bc = (cols > 0.1).nonzero().flatten().tolist()
            print(f"   bad rows: {len(br)} first {br[:4]} last {br[-4:]}; bad cols: {len(bc)} first {bc[:4]} last {bc[-4:]}", flush=True)
            bad += 1
            if p != 0 or bad >= 3:
                break
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
