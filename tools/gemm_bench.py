"""Per-shape timing of vy_gemm on the GEMMs of one captioner training step (development aid).

    python tools/gemm_bench.py [--iters 20] [--json gpurun_out/gemm_bench.json]

For every (role, M, N, K, operand majors, epilogue) of the bench workload it prints the CUDA-event
time of vy_gemm, the TFLOP/s, and the same contraction through torch.matmul (cuBLAS, no epilogue)
as a yardstick. Inputs are rotated over enough buffers to exceed the 126 MB L2.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vyomai_b200 import ops  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16


def timeit(fn, iters, nbuf):
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def case(name, M, N, K, a_mn=False, b_mn=False, epi="none", iters=20, out_dtype=BF, cublas_too=True, quiet=False):
    """A logical (M,K), B logical (N,K). a_mn/b_mn: stored transposed."""
    bytes_per = (M * K + N * K + M * N) * 2
    nbuf = max(1, min(8, int(160e6 // bytes_per) + 1))
    As = [torch.randn((K, M) if a_mn else (M, K), device=DEV, dtype=BF) for _ in range(nbuf)]
    Bs = [torch.randn((K, N) if b_mn else (N, K), device=DEV, dtype=BF) * (K ** -0.5) for _ in range(nbuf)]
    ldo = (N + 7) // 8 * 8
    outs = [torch.empty((M, ldo), device=DEV, dtype=out_dtype)[:, :N] for _ in range(nbuf)]
    bias = torch.randn(N, device=DEV, dtype=BF)
    kw = {}
    if epi in ("bias",):
        kw = dict(bias=bias)
    elif epi == "gelu_aux":
        auxs = [torch.empty((M, ldo), device=DEV, dtype=BF)[:, :N] for _ in range(nbuf)]
        kw = dict(bias=bias, act="gelu")
    elif epi == "dgelu":
        auxs = [torch.randn((M, ldo), device=DEV, dtype=BF)[:, :N] for _ in range(nbuf)]
        kw = dict(act="dgelu")
    elif epi == "addend":
        adds = [torch.randn((M, ldo), device=DEV, dtype=BF)[:, :N] for _ in range(nbuf)]
        kw = dict(bias=bias)
    elif epi in ("accum", "splitk_ok"):
        pass
    elif epi in ("swiglu", "swiglu_aux"):  # B rows are interleaved (gate, up): N = 2 * intermediate, N / 2 output columns
        outs = [torch.empty((M, N // 2), device=DEV, dtype=out_dtype) for _ in range(nbuf)]
        auxs = [torch.empty((M, N), device=DEV, dtype=BF) for _ in range(nbuf)]
        kw = dict(act="swiglu")

    def ours(i):
        a = As[i].t() if a_mn else As[i]
        b = Bs[i].t() if b_mn else Bs[i]
        k2 = dict(kw)
        if epi in ("gelu_aux", "dgelu", "swiglu_aux"):
            k2["aux"] = auxs[i]
        if epi == "addend":
            k2["addend"] = adds[i]
        if epi == "accum":
            k2["addend"] = outs[i]
            k2["allow_split_k"] = True
        if epi == "splitk_ok":
            k2["allow_split_k"] = True
        ops.gemm(a, b, out=outs[i], **k2)

    def qkv(i):
        B_, S = M // 128, 128
        nq, nkv = (12, 4) if N == 1280 else (12, 12)
        q, k, v = qkv_bufs[i]
        ops.qkv_rope_gemm(As[i], Bs[i], bias, tokens_per_seq=S, start_pos=0, n_q_heads=nq, n_kv_heads=nkv, head_dim=64,
                          rope_cos=cos if N == 1280 else None, rope_sin=sin if N == 1280 else None, q_out=q, k_out=k, v_out=v)

    if epi == "qkv":
        B_, S = M // 128, 128
        nq, nkv = (12, 4) if N == 1280 else (12, 12)
        qkv_bufs = [(torch.empty((B_, nq, S, 64), device=DEV, dtype=BF), torch.empty((B_, nkv, S, 64), device=DEV, dtype=BF),
                     torch.empty((B_, nkv, S, 64), device=DEV, dtype=BF)) for _ in range(nbuf)]
        cos = torch.rand((S, 32), device=DEV)
        sin = torch.rand((S, 32), device=DEV)
        fn = qkv
    else:
        fn = ours

    def cublas(i):
        a = As[i].t() if a_mn else As[i]
        b = Bs[i].t() if b_mn else Bs[i]
        torch.matmul(a, b.t())

    t = timeit(fn, iters, nbuf)
    if not cublas_too:
        return dict(name=name, us=t)
    tc = timeit(cublas, iters, nbuf)
    fl = 2.0 * M * N * K
    r = dict(name=name, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, epi=epi, us=round(t, 1), tflops=round(fl / t / 1e6, 1),
             cublas_us=round(tc, 1), cublas_tflops=round(fl / tc / 1e6, 1), ratio=round(tc / t, 2))
    print(f"{name:28s} M={M:6d} N={N:6d} K={K:6d} {'T' if a_mn else 'N'}{'T' if b_mn else 'N'} {epi:9s} "
          f"{t:9.1f} us {r['tflops']:7.1f} TF | cuBLAS {tc:9.1f} us {r['cublas_tflops']:7.1f} TF | x{r['ratio']:.2f}", flush=True)
    del As, Bs, outs
    return r


def bench_cases():
    """(role, M, N, K, A stored transposed, B stored transposed, epilogue) of every GEMM shape of one captioner step."""
    T, Tv = 64 * 128, 64 * 197  # decoder / ViT token rows of the bench batch
    return [
        # forward
        ("dec.qkv_rope", T, 1280, 768, False, False, "qkv"),
        ("vit.qkv", Tv, 2304, 768, False, False, "bias"),
        ("dec.attn_out+res", T, 768, 768, False, False, "addend"),
        ("vit.attn_out+res", Tv, 768, 768, False, False, "addend"),
        ("dec.ffn1 gelu", T, 3072, 768, False, False, "gelu_aux"),
        ("vit.ffn1 gelu", Tv, 3072, 768, False, False, "gelu_aux"),
        ("dec.ffn2+res", T, 768, 3072, False, False, "addend"),
        ("vit.ffn2+res", Tv, 768, 3072, False, False, "addend"),
        ("lm_head", T, 50265, 768, False, False, "bias"),
        # dgrad: dX = dY W, W [out,in] is the MN-major B operand
        ("dec.d attn_out", T, 768, 768, False, True, "none"),
        ("dec.d qkv (+res)", T, 768, 1280, False, True, "addend"),
        ("vit.d qkv (+res)", Tv, 768, 2304, False, True, "addend"),
        ("dec.d ffn2 dgelu", T, 3072, 768, False, True, "dgelu"),
        ("vit.d ffn2 dgelu", Tv, 3072, 768, False, True, "dgelu"),
        ("dec.d ffn1", T, 768, 3072, False, True, "none"),
        ("lm_head dgrad", T, 768, 50272, False, True, "splitk_ok"),
        # wgrad: dW = dY^T X, both MN-major, accumulated into the gradient buffer
        ("dec.w attn_out", 768, 768, T, True, True, "accum"),
        ("dec.w qkv", 1280, 768, T, True, True, "accum"),
        ("vit.w qkv", 2304, 768, Tv, True, True, "accum"),
        ("dec.w ffn1", 3072, 768, T, True, True, "accum"),
        ("dec.w ffn2", 768, 3072, T, True, True, "accum"),
        ("vit.w ffn1", 3072, 768, Tv, True, True, "accum"),
        ("vit.w ffn2", 768, 3072, Tv, True, True, "accum"),
        ("lm_head wgrad", 50272, 768, T, True, True, "accum"),
        # gated MLP of the RMSNorm / SwiGLU family (hidden 1024, intermediate 3072: the notebooks' Qwen3-0.6B block)
        ("gated gate|up swiglu", T, 6144, 1024, False, False, "swiglu"),
        ("gated gate|up +save z", T, 6144, 1024, False, False, "swiglu_aux"),
    ]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    cases = bench_cases()
    res = []
    for c in cases:
        if args.only and args.only not in c[0]:
            continue
        res.append(case(*c, iters=args.iters))
    from vyomai_b200 import gemm_tune
    for key, best, us, model_us in gemm_tune.LOG:
        print(f"tuned {key[:4]} mn={key[5:7]}: {best or 'model choice kept'} {us:.1f} us (model's choice {model_us:.1f} us)")
    if args.json:
        json.dump(res, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
