"""-m gpu kernel-level parity: every C-ABI kernel on the path against the CPU ORACLE (oracle/vyom_oracle.py) at the
shapes the benchmarks actually run — multi-tile online softmax, causal tile skipping, CTA-pair (cta_group::2) and
split-K GEMMs, ragged S = 197, left padding and fully padded rows — not only the toy shapes of the golden fixtures.

Inputs are made on the CPU from seeded generators, rounded to the storage dtype the kernel receives, handed to the
oracle in fp32 and to the kernel through ops.* (ctypes -> libvyom_b200.so). Stated tolerances (rel-L2 vs the oracle):
  bf16 operands, fp32 accumulate : attention / GEMM outputs 1e-2, attention gradients 2e-2
  tf32 GEMMs (fp32 tensors)      : 2e-3
  LayerNorm / AdamW / RoPE fp32  : 2e-5 or tighter (given per test)
Integer results (cache slots touched, argmax ids, dropout keep masks) are bit-exact.
"""
import math

import pytest
import torch

from oracle import vyom_oracle as O
from tests.conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _randn(shape, seed, dtype=torch.bfloat16, scale=1.0):
    """CPU fp32 values already rounded to `dtype` (so oracle and kernel see the same numbers)."""
    return (torch.randn(shape, generator=_gen(seed)) * scale).to(dtype).float()


def _kpm(B, Skv, pad):
    if pad is None:
        return None
    m = torch.ones(B, Skv, dtype=torch.long)
    for b in range(B):
        n = max(1, Skv - 3 - 5 * b)
        if pad == "right":
            m[b, n:] = 0
        elif pad == "left":
            m[b, : Skv - n] = 0
        else:
            m[b, :] = 0
    return m


def _oracle_mask(B, Sq, Skv, causal, q_pos0, kpm):
    if causal:
        return O.decoder_mask(B, Sq, kpm if kpm is not None else None, q_pos0, torch.float32)
    if kpm is None:
        return None
    return O.encoder_mask(kpm, torch.float32)


ATTN_CASES = [
    # B, Hq, Hkv, Sq, Skv, causal, q_pos0, pad
    (2, 2, 2, 128, 128, False, 0, None),
    (3, 12, 4, 17, 17, False, 0, "right"),
    (8, 12, 4, 128, 128, False, 0, "right"),       # config 1 (encoder 8 x 128, 12 / 4 heads)
    (64, 12, 12, 197, 197, False, 0, None),        # config 2 (ViT batch 64, ragged 197-key tile)
    (32, 12, 4, 512, 512, True, 0, None),          # config 3 prefill (4 query tiles x 4 key tiles, causal skipping)
    (64, 12, 4, 128, 128, True, 0, "right"),       # config 4 decoder (the bench's attention shape)
    (2, 12, 4, 248, 248, True, 0, "right"),        # notebook-II sequence length
    (2, 4, 2, 5, 12, True, 7, None),               # cached continuation: q_pos0 > 0
    (2, 4, 4, 300, 700, False, 0, "right"),        # cross-attention shape: Sq != Skv
    (2, 4, 2, 130, 130, True, 0, "left"),          # left padding: fully masked leading rows (uniform mean, quirk Q4)
    (1, 2, 1, 40, 40, False, 0, "all"),            # every key padded
]


@pytest.mark.parametrize("case", ATTN_CASES, ids=lambda c: "B%d_h%d_%d_S%d_%d_c%d_p%d_%s" % c)
def test_attn_fwd_vs_oracle(case):
    from vyomai_b200 import ops
    B, Hq, Hkv, Sq, Skv, causal, qp, pad = case
    q = _randn((B, Hq, Sq, 64), 1)
    k = _randn((B, Hkv, Skv, 64), 2)
    v = _randn((B, Hkv, Skv, 64), 3)
    kpm = _kpm(B, Skv, pad)
    ref = O.merge_heads(O.sdpa(q, O.repeat_kv(k, Hq // Hkv), O.repeat_kv(v, Hq // Hkv), _oracle_mask(B, Sq, Skv, causal, qp, kpm)))
    kbuf = torch.zeros(B, Hkv, Skv + 9, 64, device=DEV, dtype=torch.bfloat16)  # strided like a kv-cache view
    vbuf = torch.zeros_like(kbuf)
    kbuf[:, :, :Skv] = k.to(DEV)
    vbuf[:, :, :Skv] = v.to(DEV)
    kp = None if kpm is None else kpm.to(torch.uint8).to(DEV)
    out, lse = ops.attn_fwd(q.bfloat16().to(DEV), kbuf[:, :, :Skv], vbuf[:, :, :Skv], causal=causal, q_pos0=qp,
                            key_padding_mask=kp, need_lse=True, out_dtype=torch.float32)
    assert rel_l2(out.cpu(), ref) <= 6e-3
    out16, _ = ops.attn_fwd(q.bfloat16().to(DEV), kbuf[:, :, :Skv], vbuf[:, :, :Skv], causal=causal, q_pos0=qp, key_padding_mask=kp)
    assert rel_l2(out16.float().cpu(), ref) <= 1e-2
    # lse = log2-domain logsumexp of the scaled, masked scores wherever a row has a visible key
    sc = (q @ O.repeat_kv(k, Hq // Hkv).transpose(-1, -2)) / 8.0
    mk = _oracle_mask(B, Sq, Skv, causal, qp, kpm)
    vis = torch.ones(B, 1, Sq, Skv, dtype=torch.bool) if mk is None else (mk == 0).expand(B, 1, Sq, Skv)
    has = vis.any(-1).expand(B, Hq, Sq)
    ref_lse = torch.logsumexp(sc.masked_fill(~vis.expand(B, Hq, Sq, Skv), float("-inf")), -1) / math.log(2.0)
    if bool(has.any()):
        assert float((lse.cpu()[has] - ref_lse[has]).abs().max()) <= 2e-2


BWD_CASES = [
    # B, Hq, Hkv, S, causal, pad, rope
    (2, 2, 2, 128, False, None, False),
    (3, 12, 4, 17, False, "right", True),
    (8, 12, 4, 128, False, "right", True),     # config 1
    (16, 12, 12, 197, False, None, False),     # config 2 (ViT), fused kernel MHA S <= 256
    (64, 12, 4, 128, True, "right", True),     # config 4 decoder (bench shape, fused kernel)
    (2, 12, 12, 248, True, "right", True),
    (4, 12, 4, 512, True, None, True),         # config 3 length: general dK/dV + dQ kernel pair
    (2, 4, 2, 300, False, "right", False),
    (2, 4, 4, 130, True, "left", True),
    (3, 6, 2, 100, True, "right", True),
]


@pytest.mark.parametrize("case", BWD_CASES, ids=lambda c: "B%d_h%d_%d_S%d_c%d_%s_r%d" % c)
def test_attn_bwd_vs_oracle(case):
    """dq / dk / dv of attention INCLUDING the inverse rotation: the leaves are the pre-RoPE projections, the oracle
    differentiates apply_rope -> repeat_kv -> sdpa -> merge_heads with autograd on the CPU."""
    from vyomai_b200 import ops
    B, Hq, Hkv, S, causal, pad, use_rope = case
    d = 64
    qp = _randn((B, Hq, S, d), 4).requires_grad_(True)
    kp = _randn((B, Hkv, S, d), 5).requires_grad_(True)
    vp = _randn((B, Hkv, S, d), 6).requires_grad_(True)
    freqs = O.rope_freqs(S, d)
    qr, kr = O.apply_rope(qp, kp, freqs) if use_rope else (qp, kp)
    kpm = _kpm(B, S, pad)
    o_ref = O.merge_heads(O.sdpa(qr, O.repeat_kv(kr, Hq // Hkv), O.repeat_kv(vp, Hq // Hkv), _oracle_mask(B, S, S, causal, 0, kpm)))
    dout = _randn((B, S, Hq * d), 7)
    o_ref.backward(dout)

    q16, k16, v16 = (t.detach().bfloat16().to(DEV) for t in (qr, kr, vp))
    kpd = None if kpm is None else kpm.to(torch.uint8).to(DEV)
    o, lse = ops.attn_fwd(q16, k16, v16, causal=causal, q_pos0=0, key_padding_mask=kpd, need_lse=True)
    N = (Hq + 2 * Hkv) * d
    dqkv = torch.zeros(B * S, N, device=DEV, dtype=torch.float32)
    cos = freqs[0].cos().contiguous().to(DEV) if use_rope else None
    sin = freqs[0].sin().contiguous().to(DEV) if use_rope else None
    ops.attn_bwd(q16, k16, v16, o, dout.bfloat16().to(DEV), lse, causal=causal, q_pos0=0, key_padding_mask=kpd, rope_cos=cos,
                 rope_sin=sin, dq=dqkv[:, : Hq * d], dk=dqkv[:, Hq * d:(Hq + Hkv) * d], dv=dqkv[:, (Hq + Hkv) * d:])
    g = dqkv.view(B, S, Hq + 2 * Hkv, d).permute(0, 2, 1, 3).cpu()
    assert rel_l2(g[:, :Hq], qp.grad) <= 2e-2
    assert rel_l2(g[:, Hq:Hq + Hkv], kp.grad) <= 2e-2
    assert rel_l2(g[:, Hq + Hkv:], vp.grad) <= 2e-2


@pytest.mark.parametrize("cdt,tol", [(torch.bfloat16, 8e-3), (torch.float32, 2e-5)])
@pytest.mark.parametrize("case", [(32, 12, 12, 640, 768, 0), (32, 12, 4, 513, 768, 0), (3, 12, 4, 0, 16, 0),
                                  (2, 12, 12, 5, 16, 1), (1, 8, 1, 300, 384, 0), (4, 12, 4, 100, 128, 3),
                                  # long contexts: the bf16 copy-engine kernel refills its 3-stage ring (16 / 6 chunks per CTA)
                                  (2, 12, 4, 2000, 2048, 1), (2, 12, 12, 1500, 1536, 2), (2, 12, 2, 700, 768, 0),
                                  (3, 8, 4, 129, 256, 0), (2, 16, 4, 128, 256, 1), (1, 12, 4, 1, 8, 0)],
                         ids=lambda c: "B%d_h%d_%d_pos%d_len%d_split%d" % c)
def test_attn_decode_vs_oracle(case, cdt, tol):
    """Single-token decode: RoPE of q / k at `start`, append to the cache, unmasked attention over [0, start] (quirk Q3).
    The cache must change in slot `start` only (bit-exact elsewhere, and in the spare batch row)."""
    from vyomai_b200 import ops
    B, Hq, Hkv, start, clen, splits = case
    d = 64
    qkv = _randn((B, (Hq + 2 * Hkv) * d), 8, cdt)
    kc0 = torch.zeros(B + 1, Hkv, clen, d)
    vc0 = torch.zeros(B + 1, Hkv, clen, d)
    kc0[:, :, :start] = _randn((B + 1, Hkv, start, d), 9, cdt)
    vc0[:, :, :start] = _randn((B + 1, Hkv, start, d), 10, cdt)
    freqs = O.rope_freqs(clen, d)
    y = O.split_heads(qkv[:, None, :], d)                       # (B, Hq + 2 Hkv, 1, d)
    qh, knew = O.apply_rope(y[:, :Hq], y[:, Hq:Hq + Hkv], freqs[:, start:start + 1])
    vnew = y[:, Hq + Hkv:]
    cache = O.StaticCacheOneOracle(1, B, Hkv, clen, d)
    cache.key_cache[0].copy_(kc0[:B])
    cache.value_cache[0].copy_(vc0[:B])
    kk, vv = cache.update(0, knew, vnew, start)  # attention sees the unrounded new row (the kernel takes it from smem)
    ref = O.merge_heads(O.sdpa(qh, O.repeat_kv(kk, Hq // Hkv), O.repeat_kv(vv, Hq // Hkv), None)).reshape(B, Hq * d)

    kc, vc = kc0.to(cdt).to(DEV), vc0.to(cdt).to(DEV)
    out = ops.attn_decode(qkv.to(cdt).to(DEV), kc, vc, start, Hq, Hkv, freqs[0].cos().contiguous().to(DEV),
                          freqs[0].sin().contiguous().to(DEV), splits=splits, out_dtype=torch.float32)
    assert rel_l2(out.cpu(), ref) <= tol
    kexp, vexp = kc0.clone(), vc0.clone()
    kexp[:B, :, start] = knew[:, :, 0]
    vexp[:B, :, start] = vnew[:, :, 0]
    touched = torch.zeros(B + 1, Hkv, clen, dtype=torch.bool)
    touched[:B, :, start] = True
    for got, exp, old in ((kc, kexp, kc0), (vc, vexp, vc0)):
        got = got.float().cpu()
        assert torch.equal(got[~touched], old.to(cdt).float()[~touched])  # indexing: nothing else moved
        assert rel_l2(got[touched], exp[touched]) <= (4e-3 if cdt == torch.bfloat16 else 1e-6)


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 8e-3), (torch.float32, 2e-3)])
@pytest.mark.parametrize("case", [(8, 128, 12, 4, 0, 128), (3, 17, 12, 12, 0, 32), (2, 5, 12, 4, 7, 40), (64, 128, 12, 4, 0, 128)],
                         ids=lambda c: "B%d_S%d_h%d_%d_pos%d_cap%d" % c)
def test_qkv_rope_gemm_vs_oracle(case, dtype, tol):
    """Packed q|k|v projection with the bias + RoPE + head-split + cache-append epilogue against
    linear -> split_heads -> apply_rope -> StaticCacheOne.update of the oracle."""
    from vyomai_b200 import ops
    B, S, hq, hkv, start, cap = case
    d, H = 64, 768
    N = (hq + 2 * hkv) * d
    x = _randn((B * S, H), 11, dtype)
    w = _randn((N, H), 12, dtype, H ** -0.5)
    bias = _randn((N,), 13, dtype)
    freqs = O.rope_freqs(cap, d)
    y = O.split_heads(O.linear(x, w, bias).view(B, S, N), d)
    f = freqs[:, start:start + S]
    # the reference rounds cos / sin to the model dtype before the multiply (quirk Q6); the kernel's tables carry the same rounding
    emb = torch.cat((f, f), -1)
    cos, sin = emb.cos().to(dtype).float()[:, None], emb.sin().to(dtype).float()[:, None]
    rot = lambda t: t * cos + O.rotate_half(t) * sin  # noqa: E731
    q_ref, k_ref, v_ref = rot(y[:, :hq]), rot(y[:, hq:hq + hkv]), y[:, hq + hkv:]

    cosd = freqs[0].cos().to(dtype).float().contiguous().to(DEV)
    sind = freqs[0].sin().to(dtype).float().contiguous().to(DEV)
    q = torch.zeros(B, hq, S, d, device=DEV, dtype=dtype)
    kc = torch.zeros(B, hkv, cap, d, device=DEV, dtype=dtype)
    vc = torch.zeros(B, hkv, cap, d, device=DEV, dtype=dtype)
    ops.qkv_rope_gemm(x.to(dtype).to(DEV), w.to(dtype).to(DEV), bias.to(dtype).to(DEV), tokens_per_seq=S, start_pos=start,
                      n_q_heads=hq, n_kv_heads=hkv, head_dim=d, rope_cos=cosd, rope_sin=sind, q_out=q, k_out=kc, v_out=vc)
    assert rel_l2(q.float().cpu(), q_ref) <= tol
    assert rel_l2(kc[:, :, start:start + S].float().cpu(), k_ref) <= tol
    assert rel_l2(vc[:, :, start:start + S].float().cpu(), v_ref) <= tol
    assert bool((kc[:, :, :start] == 0).all()) and bool((kc[:, :, start + S:] == 0).all())  # slots outside stay untouched
    assert bool((vc[:, :, :start] == 0).all()) and bool((vc[:, :, start + S:] == 0).all())


def test_qkv_rope_gemm_refuses_to_write_past_the_cache():
    """kv_cap / rope_rows: an append that does not fit returns an error instead of writing out of bounds."""
    from vyomai_b200 import _lib, ops
    x = torch.zeros(4, 768, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(20 * 64, 768, device=DEV, dtype=torch.bfloat16)
    q = torch.zeros(1, 12, 4, 64, device=DEV, dtype=torch.bfloat16)
    kc = torch.zeros(1, 4, 8, 64, device=DEV, dtype=torch.bfloat16)
    cos = torch.zeros(16, 32, device=DEV)
    with pytest.raises(_lib.VyomError, match="do not fit"):
        ops.qkv_rope_gemm(x, w, None, tokens_per_seq=4, start_pos=6, n_q_heads=12, n_kv_heads=4, head_dim=64, rope_cos=cos,
                          rope_sin=cos, q_out=q, k_out=kc, v_out=kc.clone())
    with pytest.raises(_lib.VyomError, match="RoPE tables"):
        ops.qkv_rope_gemm(x, w, None, tokens_per_seq=4, start_pos=14, kv_dst_pos0=0, n_q_heads=12, n_kv_heads=4, head_dim=64,
                          rope_cos=cos, rope_sin=cos, q_out=q, k_out=kc, v_out=kc.clone())
    with pytest.raises(_lib.VyomError, match="RoPE tables"):
        ops.attn_decode(torch.zeros(1, 20 * 64, device=DEV, dtype=torch.bfloat16), kc, kc.clone(), 7, 12, 4, cos[:4], cos[:4])


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 4e-3)])
@pytest.mark.parametrize("rows,H", [(8192, 768), (1024, 768), (51, 768), (333, 1152), (64, 2048), (7, 64)])
def test_add_layernorm_fwd_bwd_vs_oracle(rows, H, dtype, tol):
    from vyomai_b200 import ops
    x, r = _randn((rows, H), 14, dtype), _randn((rows, H), 15, dtype)
    g, b = _randn((H,), 16, dtype), _randn((H,), 17, dtype)
    y, s, mean, rstd = ops.add_layernorm(x.to(dtype).to(DEV), r.to(dtype).to(DEV), g.to(dtype).to(DEV), b.to(dtype).to(DEV),
                                         1e-5, save_stats=True, save_sum=True)
    assert rel_l2(s.float().cpu(), x + r) <= tol
    sref = s.float().cpu().clone().requires_grad_(True)  # gradients at the (rounded) saved sum
    gref, bref = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yref = O.layer_norm(sref, gref, bref, 1e-5)
    assert rel_l2(y.float().cpu(), yref) <= tol
    dy = _randn((rows, H), 18, dtype)
    yref.backward(dy)
    dx, dg, db, dbias = ops.add_layernorm_bwd(dy.to(dtype).to(DEV), s, g.to(dtype).to(DEV), mean, rstd, want_dbias=True)
    assert rel_l2(dx.float().cpu(), sref.grad) <= (tol if dtype == torch.bfloat16 else 2e-5)
    assert rel_l2(dg.cpu(), gref.grad) <= (2e-5 if dtype == torch.float32 else 2e-3)
    assert rel_l2(db.cpu(), bref.grad) <= 2e-5
    # bias gradient of the producing Linear = column sums of dx (summed in fp32 BEFORE dx is rounded to the storage dtype)
    assert rel_l2(dbias.cpu(), sref.grad.sum(0)) <= (2e-5 if dtype == torch.float32 else 2e-3)


GEMM_SHAPES = [(128, 128, 64), (1000, 520, 264), (1024, 3072, 768), (8192, 768, 3072), (8192, 3072, 768), (8192, 1280, 768),
               (1024, 50265, 768)]


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 6e-3), (torch.float32, 2e-3)])
@pytest.mark.parametrize("shape", GEMM_SHAPES, ids=lambda s: "%dx%dx%d" % s)
def test_gemm_layouts_vs_oracle(shape, dtype, tol):
    """x W^T + b for K-major / MN-major operands (what forward, dgrad and wgrad feed the kernel), library-chosen tiling."""
    from vyomai_b200 import ops
    M, N, K = shape
    if dtype == torch.float32 and N > 4096:
        N = 1000
    a, b, bias = _randn((M, K), 19, dtype), _randn((N, K), 20, dtype, K ** -0.5), _randn((N,), 21, dtype)
    ref = O.linear(a, b, bias)
    ad, bd, biasd = a.to(dtype).to(DEV), b.to(dtype).to(DEV), bias.to(dtype).to(DEV)
    ldo = (N + 7) // 8 * 8
    outbuf = torch.empty(M, ldo, device=DEV, dtype=dtype)
    assert rel_l2(ops.gemm(ad, bd, bias=biasd, out=outbuf[:, :N]).float().cpu(), ref) <= tol
    if M % 8 == 0 and N % 8 == 0:
        at, bt = ad.t().contiguous().t(), bd.t().contiguous().t()
        assert rel_l2(ops.gemm(at, bd, bias=biasd).float().cpu(), ref) <= tol
        assert rel_l2(ops.gemm(ad, bt, bias=biasd).float().cpu(), ref) <= tol
        assert rel_l2(ops.gemm(at, bt, bias=biasd).float().cpu(), ref) <= tol


@pytest.mark.parametrize("pair,bn,splits", [(0, 128, 1), (0, 256, 1), (1, 128, 1), (1, 192, 1), (1, 256, 1), (1, 128, 4), (0, 128, 3),
                                            (1, 256, 2)])
def test_gemm_pinned_flavours_vs_oracle(pair, bn, splits):
    """The kernels the bench runs — bf16 CTA-pair (cta_group::2) tiles, single-CTA tiles and split-K — each pinned through
    the tuning hook and checked on a forward shape and on a weight-gradient shape (both operands MN-major, K = tokens)."""
    from vyomai_b200 import _lib, ops
    L = _lib.lib()
    try:
        for (M, N, K, mn) in ((2048, 3072, 768, False), (768, 3072, 8192, True)):
            if splits > 1 and not mn:
                continue
            a, b = _randn((M, K), 22), _randn((N, K), 23, scale=K ** -0.5)
            ref = a @ b.t()
            ad, bd = a.bfloat16().to(DEV), b.bfloat16().to(DEV)
            if mn:
                ad, bd = ad.t().contiguous().t(), bd.t().contiguous().t()
            L.vy_gemm_tune_override(pair, bn, splits if mn else 0)
            out = ops.gemm(ad, bd, allow_split_k=mn)
            assert rel_l2(out.float().cpu(), ref) <= 6e-3, (M, N, K, pair, bn, splits)
    finally:
        L.vy_gemm_tune_override(-1, 0, 0)
    assert L.vy_gemm_poisoned() == 0


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 6e-3), (torch.float32, 2e-3)])
def test_gemm_epilogues_vs_oracle(dtype, tol):
    """bias + exact-erf GELU (+ saved pre-activation), tanh GELU, residual addend, dGELU (the dgrad epilogue), swap-AB
    decode tiles and the ViT stem's row remap, against the oracle's linear / gelu_erf / patch arithmetic."""
    from vyomai_b200 import ops
    for (M, N, K) in ((512, 768, 768), (8192, 3072, 768)):
        a, b, bias, res = _randn((M, K), 24, dtype), _randn((N, K), 25, dtype, K ** -0.5), _randn((N,), 26, dtype), _randn((M, N), 27, dtype)
        z = O.linear(a, b, bias)
        ad, bd, biasd, resd = (t.to(dtype).to(DEV) for t in (a, b, bias, res))
        aux = torch.empty(M, N, device=DEV, dtype=dtype)
        assert rel_l2(ops.gemm(ad, bd, bias=biasd, act="gelu", aux=aux).float().cpu(), O.gelu_erf(z)) <= tol
        assert rel_l2(aux.float().cpu(), z) <= tol
        assert rel_l2(ops.gemm(ad, bd, bias=biasd, act="gelu_tanh").float().cpu(), O.gelu_tanh(z)) <= tol
        assert rel_l2(ops.gemm(ad, bd, bias=biasd, addend=resd).float().cpu(), z + res) <= tol
        zz = aux.float().cpu().requires_grad_(True)
        O.gelu_erf(zz).sum().backward()
        assert rel_l2(ops.gemm(ad, bd, act="dgelu", aux=aux).float().cpu(), (a @ b.t()) * zz.grad) <= 2 * tol
    for (M, N, K) in ((32, 768, 768), (32, 3072, 768), (32, 768, 3072), (3, 2304, 768), (64, 50265, 768), (1, 768, 768)):
        a, b, bias, res = _randn((M, K), 28, dtype), _randn((N, K), 29, dtype, K ** -0.5), _randn((N,), 30, dtype), _randn((M, N), 31, dtype)
        ref = O.gelu_erf(O.linear(a, b, bias)) + res
        out = ops.gemm(a.to(dtype).to(DEV), b.to(dtype).to(DEV), bias=bias.to(dtype).to(DEV), act="gelu", addend=res.to(dtype).to(DEV),
                       swap_ab=True)
        assert rel_l2(out.float().cpu(), ref) <= tol, (M, N, K)
    Bn, P, Hd, Kp = 4, 196, 768, 768  # ViT stem: 2 * (patch W^T + b + pos[1 + p]) at row b * 197 + 1 + p
    a, w, bias, pos = _randn((Bn * P, Kp), 32, dtype), _randn((Hd, Kp), 33, dtype, Kp ** -0.5), _randn((Hd,), 34, dtype), _randn((P + 1, Hd), 35, dtype)
    outb = torch.zeros(Bn * (P + 1), Hd, device=DEV, dtype=dtype)
    ops.gemm(a.to(dtype).to(DEV), w.to(dtype).to(DEV), bias=bias.to(dtype).to(DEV), addend=pos.to(dtype).to(DEV), addend_row_mod=P,
             addend_row_off=1, out=outb, out_scale=2.0, out_row_group=P, out_row_group_stride=P + 1, out_row_off=1)
    ref = 2 * (O.linear(a, w, bias).view(Bn, P, Hd) + pos[1:])
    assert rel_l2(outb.view(Bn, P + 1, Hd)[:, 1:].float().cpu(), ref) <= tol
    assert bool((outb.view(Bn, P + 1, Hd)[:, 0] == 0).all())


@pytest.mark.parametrize("case", [(32, 768, 768), (32, 1280, 768), (32, 3072, 768), (32, 768, 3072), (32, 50265, 768), (1, 768, 768),
                                  (7, 1000, 256), (17, 40, 128), (24, 2048, 2048), (9, 776, 1536),
                                  # the PaliGemma-scale projections: long K (6 double-buffered chunks per warp, 4-CTA clusters), 1 and 32 rows
                                  (1, 2048, 16384), (32, 2048, 16384), (1, 2560, 2048)],
                         ids=lambda c: "n%d_F%d_K%d" % c)
def test_small_batch_gemm_vs_oracle(case):
    """The weight-streaming kernel behind swap-AB vy_gemm calls with <= 32 activation rows (csrc/gemm_skinny.cu): bias, both
    GELUs, residual addend, bf16 / fp32 outputs, out_scale, strided output, feature counts that are no multiple of 16 — against
    the oracle's linear / gelu, and bit-for-bit run to run (cluster reduction in fixed rank order)."""
    import ctypes
    from vyomai_b200 import _lib, ops
    n, F, K = case
    dt = torch.bfloat16
    x, w, bias, res = _randn((n, K), 40, dt), _randn((F, K), 41, dt, K ** -0.5), _randn((F,), 42, dt), _randn((n, F), 43, dt)
    xd, wd, bd, rd = (t.to(dt).to(DEV) for t in (x, w, bias, res))
    z = O.linear(x, w, bias)
    with torch.no_grad():
        o1 = ops.gemm(xd, wd, bias=bd, swap_ab=True, out_dtype=torch.float32)
        assert rel_l2(o1.cpu(), z) <= 2e-3
        o2 = ops.gemm(xd, wd, bias=bd, act="gelu", addend=rd, swap_ab=True)
        assert rel_l2(o2.float().cpu(), O.gelu_erf(z) + res) <= 6e-3
        o3 = ops.gemm(xd, wd, act="gelu_tanh", swap_ab=True, out_dtype=torch.float32, out_scale=0.5)
        assert rel_l2(o3.cpu(), 0.5 * O.gelu_tanh(O.linear(x, w, None))) <= 2e-3
        wide = torch.full((n, F + 24), 7.0, device=DEV, dtype=torch.float32)  # strided output: columns beyond F untouched
        ops.gemm(xd, wd, bias=bd, addend=res.to(DEV), swap_ab=True, out=wide[:, :F])
        assert rel_l2(wide[:, :F].cpu(), z + res) <= 2e-3 and bool((wide[:, F:] == 7.0).all())
        assert torch.equal(o1, ops.gemm(xd, wd, bias=bd, swap_ab=True, out_dtype=torch.float32))
    # this call is served by the small-batch kernel (not the tensor-memory tiles)
    st = _lib.STRUCTS["VyGemm"]()
    st.M, st.N, st.K, st.in_dtype, st.transposed_out = F, n, K, _lib.CONSTS["VY_BF16"], 1
    assert _lib.lib().vy_gemm_is_small_batch(ctypes.byref(st)) == (1 if F <= 9472 else 0)  # vocabulary-sized F keeps the tensor-memory tiles


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_apply_rotary_pos_emb_public_helper(dtype):
    """The stand-alone apply_rotary_pos_emb(q, k, freqs) of layers/positional_embeddings.py (vy_rope_apply) against the
    oracle's apply_rope, which follows the reference's rounding of cos / sin to q.dtype (quirk Q6)."""
    from vyomai_b200.layers.positional_embeddings import RotaryEmbedding, apply_rotary_pos_emb
    from dataclasses import make_dataclass
    C = make_dataclass("C", [("hidden_size", int, 768), ("num_attention_heads", int, 12), ("max_position_embeddings", int, 64)])
    freqs = RotaryEmbedding(C())(64)
    assert rel_l2(freqs, O.rope_freqs(64, 64)) <= 1e-7
    q, k = _randn((3, 12, 40, 64), 36, dtype), _randn((3, 4, 40, 64), 37, dtype)
    qr, kr = O.apply_rope(q.to(dtype), k.to(dtype), freqs[:, 5:45])
    q2, k2 = apply_rotary_pos_emb(q.to(dtype).to(DEV), k.to(dtype).to(DEV), freqs[:, 5:45])
    tol = 1e-6 if dtype == torch.float32 else 8e-3
    assert rel_l2(q2.float().cpu(), qr.float()) <= tol and rel_l2(k2.float().cpu(), kr.float()) <= tol


def test_argmax_rows_first_index_rule():
    """Greedy selection = torch.topk(k=1): the FIRST index of the row maximum (models/decoder.py:489-496)."""
    from vyomai_b200 import ops
    x = torch.randn(32, 50265, generator=_gen(38))
    x[3, 40000] = x[3, 17] = 9.0   # tie: the lower index wins
    x[7, 50264] = 11.0             # maximum in the last column
    x[9, 0] = 12.0
    for dt in (torch.float32, torch.bfloat16):
        xx = x.to(dt)
        want = torch.topk(xx.float(), 1, dim=-1)[1].reshape(-1)
        ld = (50265 + 7) // 8 * 8
        buf = torch.zeros(32, ld, device=DEV, dtype=dt)
        buf[:, :50265] = xx.to(DEV)
        got = ops.argmax_rows(buf[:, :50265]).cpu()
        vals = xx.float()
        assert torch.equal(vals[torch.arange(32), got], vals[torch.arange(32), want])  # same maximum
        tie_free = torch.tensor([int((vals[r] == vals[r].max()).sum()) == 1 for r in range(32)])
        assert torch.equal(got[tie_free], want[tie_free])
        assert int(got[3]) == int((vals[3] == vals[3].max()).nonzero()[0])


def test_embedding_backward_skips_padding_rows():
    """nn.Embedding(padding_idx=pad_token_id): neither the word-table row of the pad token nor the learned-position row
    `pad_token_id` (the reference builds AbsoluteEncoding that way too) receives a gradient; every other row matches
    the oracle's index_add."""
    from vyomai_b200 import ops
    V, H, B, S, pad = 50, 64, 3, 10, 1
    ids = torch.randint(0, V, (B, S), generator=_gen(39))
    ids[:, -3:] = pad
    dout = _randn((B * S, H), 40, torch.float32)
    ref_t = torch.zeros(V, H).index_add_(0, ids.reshape(-1), dout)
    ref_t[pad] = 0
    ref_p = dout.view(B, S, H).sum(0)
    ref_p[pad] = 0
    dt, dp = torch.zeros(V, H, device=DEV), torch.zeros(S, H, device=DEV)
    ops.embed_bwd(ids.reshape(-1).to(DEV), dout.to(DEV), rows=B * S, H=H, tokens_per_seq=S, dtable=dt, dpos=dp, padding_idx=pad,
                  pos_padding_idx=pad)
    assert rel_l2(dt.cpu(), ref_t) <= 1e-6 and bool((dt[pad] == 0).all())
    assert rel_l2(dp.cpu(), ref_p) <= 1e-6 and bool((dp[pad] == 0).all())


def test_decode_through_the_bare_layer_api_with_a_freqs_slice():
    """DecoderAttention called like the reference's tests call it — `freqs` is the (1, S, d/2) angle slice of
    the CURRENT positions, start_pos > 0 — must rotate the new token by position start_pos, not by table row 0."""
    from dataclasses import make_dataclass
    from vyomai_b200.layers.attention import DecoderAttentionGqa
    from vyomai_b200.layers.kv_cache import StaticCache
    C = make_dataclass("C", [("hidden_size", int, 128), ("num_attention_heads", int, 2), ("num_key_value_heads", int, 1),
                             ("max_position_embeddings", int, 32), ("hidden_dropout_prob", float, 0.0), ("layer_norm_eps", float, 1e-5)])
    cfg = C()
    torch.manual_seed(0)
    att = DecoderAttentionGqa(cfg, layer_idx=0).to(DEV).eval()
    att.cache = StaticCache(cfg, is_gqa=True)
    sd = {k: v.detach().float().cpu() for k, v in att.state_dict().items()}
    x = _randn((1, 6, 128), 41, torch.float32)
    freqs = O.rope_freqs(32, 64)
    ocfg = O.Cfg(128, 2, 1, 32, 1, 0, 1e-5, "gelu")
    ocache = O.PerLayerCacheAdapter(O.StaticCacheOneOracle(1, 1, 1, 32, 64))
    with torch.no_grad():
        mask = O.decoder_mask(1, 5, None, 0, torch.float32)
        ref0 = O.self_attention(sd, "", x[:, :5], mask, freqs[:, :5], ocfg, "gqa", cache=ocache, layer_idx=0, start_pos=0)
        ref1 = O.self_attention(sd, "", x[:, 5:], None, freqs[:, 5:6], ocfg, "gqa", cache=ocache, layer_idx=0, start_pos=5)
        got0 = att(x[:, :5].to(DEV), mask.to(DEV), freqs=freqs[:, :5], use_cache=True, start_pos=0)
        got1 = att(x[:, 5:].to(DEV), None, freqs=freqs[:, 5:6], use_cache=True, start_pos=5)
    assert rel_l2(got0.cpu(), ref0) <= 6e-3
    assert rel_l2(got1.cpu(), ref1) <= 6e-3


# ---------------------------------------------------------------------------------------------
# softmax cross-entropy with the vocabulary bias gradient (column sums of the written gradient) taken in the same pass
# ---------------------------------------------------------------------------------------------
XENT_COLSUM_CASES = [
    # rows, V, row stride, ignored rows
    (300, 50272, 50272, "some"),   # the captioner's vocabulary: 7 vectors per thread, the last one partly populated
    (37, 1000, 1024, "some"),      # fewer rows than CTAs, padded rows, one vector per thread for 125 threads
    (160, 8192, 8192, "none"),
    (129, 50265, 50272, "some"),   # bench.py's vocabulary: the last vector holds one column and seven padding lanes
    (40, 1003, 1008, "none"),
    (150, 57344, 57344, "all"),    # the widest vocabulary the fused path takes; every row ignored -> zeros
]


@pytest.mark.parametrize("rows,V,ld,ignored", XENT_COLSUM_CASES)
def test_softmax_xent_with_fused_column_sums(rows, V, ld, ignored):
    """vy_softmax_xent(colsum_part) + vy_colsum_finish against the oracle's loss (cross_entropy_shifted's formula:
    logsumexp - picked logit, ignore_index rows dropped) and autograd's gradients of it: row losses, the gradient written
    over the logits, and its column sums (the LM head's bias gradient) — also against the two-kernel path (vy_softmax_xent
    without colsum_part, then vy_colsum). Tolerances: losses 2e-5 (fp32 math on the same bf16 logits), gradient 4e-3 rel-L2
    (bf16 storage), column sums 2e-3 rel-L2 of the oracle's fp32 sums (the fused sums add unrounded fp32 terms)."""
    from vyomai_b200 import ops
    g = _gen(1234 + rows)
    x = (torch.randn(rows, V, generator=g) * 2.5).to(torch.bfloat16)
    labels = torch.randint(0, V, (rows,), generator=g)
    if ignored == "some":
        labels[::5] = -100
        labels[rows - 1] = -100
    elif ignored == "all":
        labels[:] = -100
    gs = 1.0 / max(1, int((labels != -100).sum()))
    xr = x.float().requires_grad_(True)
    keep = labels != -100
    lse = torch.logsumexp(xr, dim=-1)
    picked = xr.gather(1, labels.clamp(min=0)[:, None])[:, 0]
    ref_rows = (lse - picked) * keep
    (ref_rows.sum() * gs).backward()
    ref_grad, ref_bias = xr.grad, xr.grad.sum(0)

    def device_logits():
        buf = torch.full((rows, ld), float("nan"), dtype=torch.bfloat16, device=DEV)  # padding columns hold garbage
        buf[:, :V] = x.to(DEV)
        return buf[:, :V]

    lg = device_logits()
    part = ops.xent_colsum_part(lg)
    assert part is not None and part.shape[1] == V
    inv = torch.tensor([gs], dtype=torch.float32, device=DEV)
    loss_rows = ops.softmax_xent(lg, labels.to(DEV), ignore_index=-100, grad_scale_ptr=inv, write_grad=True, colsum_part=part)
    two = torch.tensor([2.0], dtype=torch.float32, device=DEV)
    bias = ops.colsum_finish(part, out_dtype=torch.float32)
    bias2 = ops.colsum_finish(part, out=bias.clone(), accumulate=True, scale_ptr=two)  # b + 2 b
    torch.cuda.synchronize()
    assert torch.allclose(loss_rows.cpu(), ref_rows.detach(), rtol=2e-5, atol=2e-5)
    if ignored == "all":
        assert not lg.any() and not bias.any()
        return
    assert rel_l2(lg.float().cpu(), ref_grad) <= 4e-3
    assert not lg[~keep.to(DEV)].any()
    pad = torch.as_strided(lg, (rows, (V + 7) // 8 * 8), (ld, 1), lg.storage_offset())[:, V:]
    assert not pad.any()  # the lanes of a last partial vector beyond V are written with zeros
    assert rel_l2(bias.cpu(), ref_bias) <= 2e-3
    assert torch.allclose(bias2, 3.0 * bias, rtol=1e-6, atol=0.0)
    # the two-kernel path on the same inputs
    lg2 = device_logits()
    loss_rows2 = ops.softmax_xent(lg2, labels.to(DEV), ignore_index=-100, grad_scale_ptr=inv, write_grad=True)
    bias_sep = ops.colsum(lg2, out_dtype=torch.float32)
    assert torch.allclose(loss_rows2, loss_rows, rtol=2e-5, atol=2e-5)
    assert rel_l2(lg.float(), lg2.float()) <= 4e-3
    assert rel_l2(bias, bias_sep) <= 4e-3


def test_fused_column_sums_are_declined_outside_their_shapes():
    from vyomai_b200 import ops
    assert ops.xent_colsum_part(torch.zeros(4, 1000, device=DEV)) is None                          # fp32 logits
    assert ops.xent_colsum_part(torch.zeros(4, 1004, device=DEV, dtype=torch.bfloat16)) is None    # rows do not cover V rounded up to 8
    assert ops.xent_colsum_part(torch.zeros(4, 57352, device=DEV, dtype=torch.bfloat16)) is None   # wider than the accumulators
