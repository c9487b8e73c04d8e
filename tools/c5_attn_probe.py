"""The two attention shapes of the PaliGemma-scale prefill on the mma.sync kernel: SigLIP (B 8, 16 heads of 72, 256 tokens) and
Gemma (B 8, 8 q heads / 1 kv head of 256, 264 tokens); timing + a launch sequence for ncu (4 warm-up launches, then 2 more)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vyomai_b200 import ops

def t(fn, n=20):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3

B = 8
qs, ks, vs = (torch.randn(B, 16, 256, 72, device="cuda").bfloat16() for _ in range(3))
qg = torch.randn(B, 8, 264, 256, device="cuda").bfloat16()
kg, vg = (torch.randn(B, 1, 264, 256, device="cuda").bfloat16() for _ in range(2))
for _ in range(2):
    ops.attn_fwd(qs, ks, vs); ops.attn_fwd(qg, kg, vg)
ops.attn_fwd(qs, ks, vs); ops.attn_fwd(qg, kg, vg)
us_s, us_g = t(lambda: ops.attn_fwd(qs, ks, vs)), t(lambda: ops.attn_fwd(qg, kg, vg))
fl_s, fl_g = 4.0 * B * 16 * 256 * 256 * 72, 4.0 * B * 8 * 264 * 264 * 256
print(f"siglip attention (B{B} h16 S256 d72): {us_s:.1f} us, {fl_s / us_s / 1e6:.1f} TFLOP/s")
print(f"gemma prefill attention (B{B} h8/kv1 S264 d256): {us_g:.1f} us, {fl_g / us_g / 1e6:.1f} TFLOP/s")
