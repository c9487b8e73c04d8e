"""Times vy_softmax_xent at the captioner's shape (rows x V bf16 logits, padded rows): the two-kernel form (loss + gradient,
then vy_colsum of the gradient) against the fused form (column sums taken in the same pass). CUDA events, L2 flushed by the
0.8 GB working set itself. Usage: python tools/xent_bench.py [rows] [V] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from vyomai_b200 import ops

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 129
V = int(sys.argv[2]) if len(sys.argv) > 2 else 50265
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ld = (V + 7) // 8 * 8
g = torch.Generator(device="cuda").manual_seed(0)
src = (torch.randn(rows, ld, device="cuda", generator=g) * 2.5).to(torch.bfloat16)
labels = torch.randint(0, V, (rows,), device="cuda", generator=g)
labels[::129] = -100
inv = torch.tensor([1.0 / rows], device="cuda")
buf = torch.empty_like(src)
lg = buf[:, :V]
part = ops.xent_colsum_part(lg)
out = torch.empty(V, device="cuda", dtype=torch.float32)


def timed(fn):
    ts = []
    for i in range(iters + 2):
        buf.copy_(src)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def two_kernel():
    ops.softmax_xent(lg, labels, grad_scale_ptr=inv, write_grad=True)
    ops.colsum(lg, out=out)


def fused():
    ops.softmax_xent(lg, labels, grad_scale_ptr=inv, write_grad=True, colsum_part=part)
    ops.colsum_finish(part, out=out)


t2, tf = timed(two_kernel), timed(fused)
gb = rows * V * 2 / 1e9
print(f"rows {rows} V {V}: two-kernel {t2:.1f} us, fused {tf:.1f} us; gradient buffer {gb:.2f} GB -> fused moves "
      f"{2 * gb / tf * 1e3:.2f} TB/s algorithmic (1 read + 1 write)")
