"""Timing of vy_attn_fwd / vy_attn_bwd on the attention shapes of the bench workload (development aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vyomai_b200 import ops  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
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
    for name, B, Hq, Hkv, S, causal, pad in [("decoder", 64, 12, 4, 128, True, True), ("vit", 64, 12, 12, 197, False, False)]:
        d = 64
        q = torch.randn(B, Hq, S, d, device=dev).bfloat16()
        k = torch.randn(B, Hkv, S, d, device=dev).bfloat16()
        v = torch.randn(B, Hkv, S, d, device=dev).bfloat16()
        kpm = None
        if pad:
            lens = torch.randint(8, S + 1, (B,), device=dev)
            kpm = (torch.arange(S, device=dev)[None] < lens[:, None]).to(torch.uint8).contiguous()
        o, lse = ops.attn_fwd(q, k, v, causal=causal, q_pos0=0, key_padding_mask=kpm, need_lse=True)
        dout = torch.randn(B, S, Hq * d, device=dev).bfloat16()
        N = (Hq + 2 * Hkv) * d
        dqkv = torch.empty(B * S, N, device=dev, dtype=torch.bfloat16)
        cos = torch.rand(S, 32, device=dev)
        sin = torch.rand(S, 32, device=dev)
        tf = timeit(lambda: ops.attn_fwd(q, k, v, causal=causal, q_pos0=0, key_padding_mask=kpm, need_lse=True))
        tb = timeit(lambda: ops.attn_bwd(q, k, v, o, dout, lse, causal=causal, q_pos0=0, key_padding_mask=kpm, rope_cos=cos,
                                         rope_sin=sin, dq=dqkv[:, : Hq * d], dk=dqkv[:, Hq * d:(Hq + Hkv) * d],
                                         dv=dqkv[:, (Hq + Hkv) * d:]))
        fl = 4.0 * B * Hq * S * S * d
        print(f"{name:8s} B={B} Hq={Hq} Hkv={Hkv} S={S}: fwd {tf:7.1f} us ({fl / tf / 1e6:6.1f} TF)   bwd {tb:7.1f} us ({2.5 * fl / tb / 1e6:6.1f} TF useful)")


if __name__ == "__main__":
    main()
