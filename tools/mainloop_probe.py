import torch, sys, os
sys.path.insert(0, "/root/repo")
from vyomai_b200 import ops
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
M, N, K = 3072, 768, 8192
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16); b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
at = torch.randn(K, M, device="cuda", dtype=torch.bfloat16); bt = torch.randn(K, N, device="cuda", dtype=torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
fl = 2.0 * M * N * K
us = t(lambda: ops.gemm(a, b, out=out)); print(f"BN={os.environ.get('VY_GEMM_FORCE_BN')} dbg={os.environ.get('VY_GEMM_DEBUG')} KK  {us:7.1f} us {fl/us/1e6:7.1f} TF")
us = t(lambda: ops.gemm(at.t(), bt.t(), out=out)); print(f"BN={os.environ.get('VY_GEMM_FORCE_BN')} dbg={os.environ.get('VY_GEMM_DEBUG')} TT  {us:7.1f} us {fl/us/1e6:7.1f} TF")
