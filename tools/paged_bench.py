"""vy_attn_decode: paged pools + per-row context lengths vs the contiguous cache, same logical contents (config-3
decode shapes: B 32, 12 q heads, 4 | 12 kv heads, context 640, bf16). Prints one JSON line per case with the time per
call, the algorithmic bytes (2 * B * Hkv * ctx * 64 * 2) and the fraction of the measured HBM bandwidth."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vyomai_b200 import ops  # noqa: E402

HBM_GBS = 6535.0
try:
    HBM_GBS = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
                    .get("hbm_gbs", HBM_GBS))
except Exception:
    pass


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    dev = "cuda"
    B, Hq, ctx, bs, layers = 32, 12, 640, 256, 8  # `layers` distinct caches are rotated so that the reads miss L2
    for Hkv in (4, 12):
        nblk = (ctx + 1 + bs - 1) // bs
        pools_k = [torch.randn(B * nblk, bs, Hkv, 64, device=dev, dtype=torch.bfloat16) for _ in range(layers)]
        pools_v = [torch.randn_like(p) for p in pools_k]
        table = torch.randperm(B * nblk, device=dev).to(torch.int32).view(B, nblk)
        seq = torch.full((B,), ctx, dtype=torch.int32, device=dev)
        ragged = (torch.arange(B, device=dev) * 19 % ctx + 1).to(torch.int32)  # per-row lengths 1..ctx (mean ~ ctx / 2)
        cont_k = [torch.randn(B, Hkv, nblk * bs, 64, device=dev, dtype=torch.bfloat16) for _ in range(layers)]
        cont_v = [torch.randn_like(c) for c in cont_k]
        qkv = torch.randn(B, (Hq + 2 * Hkv) * 64, device=dev, dtype=torch.bfloat16)
        cos = torch.rand(nblk * bs, 32, device=dev)
        sin = torch.rand(nblk * bs, 32, device=dev)
        out = torch.empty(B, Hq * 64, device=dev, dtype=torch.bfloat16)
        state = {"i": 0}

        def contiguous():
            i = state["i"] = (state["i"] + 1) % layers
            ops.attn_decode(qkv, cont_k[i], cont_v[i], ctx, Hq, Hkv, cos, sin, out=out)

        def paged():
            i = state["i"] = (state["i"] + 1) % layers
            ops.attn_decode(qkv, pools_k[i], pools_v[i], ctx, Hq, Hkv, cos, sin, out=out, seqlens=seq, block_table=table)

        def paged_ragged():
            i = state["i"] = (state["i"] + 1) % layers
            ops.attn_decode(qkv, pools_k[i], pools_v[i], ctx, Hq, Hkv, cos, sin, out=out, seqlens=ragged, block_table=table)

        full = 2.0 * B * Hkv * ctx * 64 * 2
        rag = 2.0 * float(ragged.sum()) * Hkv * 64 * 2
        for name, fn, nbytes in (("contiguous", contiguous, full), ("paged", paged, full), ("paged_ragged", paged_ragged, rag)):
            us = timeit(fn)
            print(json.dumps({"kernel": "vy_attn_decode", "mode": name, "B": B, "n_kv_heads": Hkv, "ctx": ctx, "block_size": bs,
                              "us_per_call": round(us, 2), "algorithmic_MB": round(nbytes / 1e6, 2),
                              "GBs": round(nbytes / us / 1e3, 1), "hbm_frac_of_measured": round(nbytes / us / 1e3 / HBM_GBS, 3)}), flush=True)


if __name__ == "__main__":
    main()
