"""Fused transformer-block building blocks on top of the C ABI (ops.py).

Each function mirrors one stretch of the reference's layer code and cites it; tensors stay on the
GPU and every arithmetic step is one of the sm_100a kernels — there is no torch-op fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops

HEAD_DIM = 64  # the tcgen05 attention / QKV-RoPE epilogues are specialised for head_dim 64

# decode-shaped GEMMs (few tokens, many output features) run swap-AB
SWAP_AB_MAX_ROWS = 64


@dataclass
class MaskSpec:
    """Factored form of the reference's additive float masks.

    encoder / ViT / cross-attention: (1 - m[:, None, None, :]) * finfo.min  (models/encoder.py:161-164)
        -> key_padding = m (uint8 [B, Skv]), causal = False
    decoder, seqlen > 1: (1 - causal * m) * finfo.min  (models/decoder.py:355-362, 376-419)
        -> key_padding = m, causal = True, q_pos0 = start_pos
    decoder, seqlen == 1: mask is None (quirk Q3) -> MaskSpec(None, False, 0)
    """

    key_padding: Optional[torch.Tensor] = None
    causal: bool = False
    q_pos0: int = 0

    @staticmethod
    def from_attention_mask(attention_mask: Optional[torch.Tensor], causal: bool, q_pos0: int = 0) -> "MaskSpec":
        kpm = None
        if attention_mask is not None:
            kpm = (attention_mask != 0).to(torch.uint8).contiguous()
        return MaskSpec(kpm, causal, q_pos0)

    @staticmethod
    def from_dense(mask: Optional[torch.Tensor], seqlen_q: int) -> "MaskSpec":
        """Accepts the dense additive mask the reference's layer API takes ((B,1,1,Skv) or
        (B,1,Sq,Skv), 0 = visible) and recovers the factored form; raises if the mask is not one
        of the two families the reference builds."""
        if mask is None:
            return MaskSpec()
        if isinstance(mask, MaskSpec):
            return mask
        if mask.dim() != 4 or mask.shape[1] != 1:
            raise _lib.VyomError(f"unsupported attention mask shape {tuple(mask.shape)}")
        vis = mask[:, 0] == 0  # (B, Sq|1, Skv)
        if vis.shape[1] == 1:
            return MaskSpec(vis[:, 0].to(torch.uint8).contiguous(), False, 0)
        B, Sq, Skv = vis.shape
        q_pos0 = Skv - Sq
        kpm = vis[:, -1]  # the last query row sees every non-padded key
        kk = torch.arange(Skv, device=mask.device)[None, :]
        ll = torch.arange(Sq, device=mask.device)[:, None]
        rebuilt = (kk <= q_pos0 + ll)[None] & kpm[:, None, :]
        if not bool((rebuilt == vis).all()):
            raise _lib.VyomError("dense attention masks other than key-padding and causal*key-padding are not supported")
        return MaskSpec(kpm.to(torch.uint8).contiguous(), True, q_pos0)


# ----------------------------------------------------------------------------------------------
# flat packing of q/k/v projection weights
# ----------------------------------------------------------------------------------------------
def _adjacent(ts) -> bool:
    """True if the tensors are back-to-back slices of ONE storage (so a single strided view can span them).
    Address adjacency alone is not enough: the caching allocator happily places separate small allocations
    (e.g. three bias vectors) next to each other."""
    p = ts[0].data_ptr()
    base = ts[0].untyped_storage().data_ptr()
    for t in ts:
        if not t.is_contiguous() or t.data_ptr() != p or t.untyped_storage().data_ptr() != base:
            return False
        p += t.numel() * t.element_size()
    return True


def pack_linears(linears) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Makes the weights (and biases) of several nn.Linear layers adjacent rows of one buffer, so a
    single GEMM computes all projections (q/k/v at layers/attention.py:87-95,161-173). The
    Parameters keep their identity (`.data` is re-pointed into the packed buffer), so state_dict
    keys, optimizers and `.grad` are unaffected. Re-packs lazily after `.to()` moved them."""
    ws = [l.weight for l in linears]
    if not _adjacent([w.data for w in ws]):
        flat = torch.empty((sum(w.shape[0] for w in ws), ws[0].shape[1]), device=ws[0].device, dtype=ws[0].dtype)
        r = 0
        for w in ws:
            flat[r:r + w.shape[0]].copy_(w.data)
            w.data = flat[r:r + w.shape[0]]
            r += w.shape[0]
    w0 = ws[0].data
    W = torch.as_strided(w0, (sum(w.shape[0] for w in ws), w0.shape[1]), (w0.stride(0), 1), w0.storage_offset())
    bs = [l.bias for l in linears]
    Bv = None
    if bs[0] is not None:
        if not _adjacent([b.data for b in bs]):
            flatb = torch.empty(sum(b.shape[0] for b in bs), device=bs[0].device, dtype=bs[0].dtype)
            r = 0
            for b in bs:
                flatb[r:r + b.shape[0]].copy_(b.data)
                b.data = flatb[r:r + b.shape[0]]
                r += b.shape[0]
        b0 = bs[0].data
        Bv = torch.as_strided(b0, (sum(b.shape[0] for b in bs),), (1,), b0.storage_offset())
    return W, Bv


# ----------------------------------------------------------------------------------------------
# RoPE tables
# ----------------------------------------------------------------------------------------------
class RopeTables:
    """cos/sin of the reference's angle table `freqs` (layers/positional_embeddings.py:127-137),
    rounded to the model dtype before use exactly like apply_rotary_pos_emb does (:173-175, quirk
    Q6), kept as contiguous fp32 [max_pos, d/2] device tensors for the fused epilogues."""

    def __init__(self, freqs: torch.Tensor):
        self.freqs = freqs  # (1, max_pos, d/2) fp32, wherever the model keeps it
        self._cache = {}

    def get(self, device: torch.device, dtype: torch.dtype):
        key = (device, dtype)
        if key not in self._cache:
            f = self.freqs[0].to(torch.float32).cpu()
            cos = f.cos().to(dtype).to(torch.float32).contiguous().to(device)
            sin = f.sin().to(dtype).to(torch.float32).contiguous().to(device)
            self._cache[key] = (cos, sin)
        return self._cache[key]


def _lin(x2d: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], **kw) -> torch.Tensor:
    """nn.Linear through vy_gemm; swap-AB for decode-shaped calls."""
    return ops.gemm(x2d, w, bias=b, swap_ab=x2d.shape[0] <= SWAP_AB_MAX_ROWS, **kw)


# ----------------------------------------------------------------------------------------------
# attention block
# ----------------------------------------------------------------------------------------------
@dataclass
class KVTarget:
    """Where the new keys/values go and what attention reads. k_buf/v_buf are [B', Hkv, cap, 64]
    (B' >= B); rows [start_pos, start_pos + S) are written, [0, start_pos + S) are attended."""

    k_buf: torch.Tensor
    v_buf: torch.Tensor
    start_pos: int


def attention_core(
    x2d: torch.Tensor,
    B: int,
    S: int,
    w_qkv: torch.Tensor,
    b_qkv: Optional[torch.Tensor],
    n_q_heads: int,
    n_kv_heads: int,
    mask: MaskSpec,
    rope: Optional[Tuple[torch.Tensor, torch.Tensor, Optional[int]]],
    kv: Optional[KVTarget],
    decode_no_mask: bool,
    need_lse: bool = False,
    pos0: int = 0,
):
    """q/k/v projection -> head split -> RoPE -> kv-cache append -> (GQA) attention -> merged heads.
    Mirrors layers/attention.py:114-132 / :190-213 / :264-287 / :350-377 / :607-623 and
    models/decoder.py:87-111 / :171-199. Returns (attn_out [B*S, Hq*64], saved-for-backward tuple)."""
    dev, T = x2d.device, x2d.dtype
    d = HEAD_DIM
    cos = sin = None
    rope_pos = pos0  # row of the rope tables that belongs to the first new token
    if rope is not None:
        cos, sin, base = rope
        rope_pos = 0 if base is None else pos0 - base
    start = kv.start_pos if kv is not None else 0  # cache slot of the first new token

    if kv is not None and S == 1 and decode_no_mask:
        # single-token decode: plain projection, then the fused RoPE + append + attention kernel
        qkv = _lin(x2d, w_qkv, b_qkv)
        # vy_attn_decode addresses the tables by cache slot + offset: a caller holding only the angle rows of the
        # current positions (the bare layer API) has rope_pos = 0 while the slot is start_pos
        out = ops.attn_decode(qkv, kv.k_buf, kv.v_buf, start, n_q_heads, n_kv_heads, cos, sin, out_dtype=T,
                              rope_pos_off=rope_pos - start)
        return out, None

    q = torch.empty((B, n_q_heads, S, d), device=dev, dtype=torch.bfloat16)
    if kv is None:
        k_att = torch.empty((B, n_kv_heads, S, d), device=dev, dtype=torch.bfloat16)
        v_att = torch.empty((B, n_kv_heads, S, d), device=dev, dtype=torch.bfloat16)
        k_dst, v_dst = k_att, v_att
        dst_off = 0
    else:
        k_dst, v_dst = kv.k_buf, kv.v_buf
        dst_off = start
    ops.qkv_rope_gemm(
        x2d, w_qkv, b_qkv, tokens_per_seq=S, start_pos=rope_pos, kv_dst_pos0=dst_off, n_q_heads=n_q_heads,
        n_kv_heads=n_kv_heads, head_dim=d, rope_cos=cos, rope_sin=sin, q_out=q, k_out=k_dst[:B], v_out=v_dst[:B],
    )
    if kv is not None:
        skv = start + S
        if kv.k_buf.dtype == torch.bfloat16:
            k_att, v_att = kv.k_buf[:B, :, :skv], kv.v_buf[:B, :, :skv]
        else:  # fp32 cache (the reference's StaticCacheOne default): bf16 operand copies
            k_att = ops.cast4d(kv.k_buf[:B, :, :skv], torch.bfloat16)
            v_att = ops.cast4d(kv.v_buf[:B, :, :skv], torch.bfloat16)
    out, lse = ops.attn_fwd(q, k_att, v_att, causal=mask.causal, q_pos0=mask.q_pos0,
                            key_padding_mask=mask.key_padding, out_dtype=T, need_lse=need_lse)
    return out.view(B * S, n_q_heads * d), (q, k_att, v_att, lse)


def project_kv(enc2d: torch.Tensor, B: int, Skv: int, key: nn.Linear, value: nn.Linear, n_kv_heads: int):
    """k / v projections of the encoder states as [B, h_kv, Skv, 64] bf16 (one GEMM over the packed key | value weights)."""
    w_kv, b_kv = pack_linears([key, value])
    k = torch.empty((B, n_kv_heads, Skv, HEAD_DIM), device=enc2d.device, dtype=torch.bfloat16)
    v = torch.empty((B, n_kv_heads, Skv, HEAD_DIM), device=enc2d.device, dtype=torch.bfloat16)
    ops.qkv_rope_gemm(enc2d, w_kv, b_kv, tokens_per_seq=Skv, start_pos=0, n_q_heads=0, n_kv_heads=n_kv_heads, head_dim=HEAD_DIM,
                      rope_cos=None, rope_sin=None, q_out=None, k_out=k, v_out=v)
    return k, v


def cross_attention_core(x2d: torch.Tensor, B: int, Sq: int, enc2d: Optional[torch.Tensor], Skv: int, query: nn.Linear,
                         key: nn.Linear, value: nn.Linear, n_q_heads: int, n_kv_heads: int, mask: MaskSpec,
                         cached_kv: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, need_lse: bool = False):
    """Cross-attention core (layers/attention.py:431-468, 535-571): q from the decoder stream, k / v from the encoder
    states (or from the cache that holds them after the first generation step), no RoPE (the reference leaves it commented
    out), key-padding mask of the ENCODER sequence, GQA by head index. Returns (attn [B*Sq, Hq*64], (q, k, v, lse))."""
    dev = x2d.device
    d = HEAD_DIM
    q = torch.empty((B, n_q_heads, Sq, d), device=dev, dtype=torch.bfloat16)
    ops.qkv_rope_gemm(x2d, query.weight, query.bias, tokens_per_seq=Sq, start_pos=0, n_q_heads=n_q_heads, n_kv_heads=0,
                      head_dim=d, rope_cos=None, rope_sin=None, q_out=q, k_out=None, v_out=None)
    if cached_kv is not None:
        k, v = cached_kv
        if k.dtype != torch.bfloat16:
            k, v = ops.cast4d(k, torch.bfloat16), ops.cast4d(v, torch.bfloat16)
    else:
        k, v = project_kv(enc2d, B, Skv, key, value, n_kv_heads)
    out, lse = ops.attn_fwd(q, k, v, causal=False, q_pos0=0, key_padding_mask=mask.key_padding, out_dtype=x2d.dtype, need_lse=need_lse)
    return out.view(B * Sq, n_q_heads * d), (q, k, v, lse)


def dropout_state(module: nn.Module, p: float) -> Optional[ops.DropoutState]:
    """The reference's nn.Dropout(hidden_dropout_prob) is live in .train() (attention.py:55,70; ffn.py:24,38): a fresh
    mask identity per call then, None in eval / p = 0."""
    if module.training and p > 0.0:
        return ops.DropoutState(p)
    return None


def self_output(attn2d: torch.Tensor, residual2d: torch.Tensor, dense: nn.Linear, ln: nn.LayerNorm, save: bool = False,
                dropout: Optional[ops.DropoutState] = None):
    """AttentionSelfOutput.forward (layers/attention.py:57-72): LN(dropout(dense(attn)) + residual). Without dropout the
    residual add rides in the GEMM epilogue; with it the GEMM writes dense(attn) and the norm kernel drops, adds the
    residual and normalises in one pass (the mask is regenerated in backward, never stored)."""
    if dropout is None:
        s = _lin(attn2d, dense.weight, dense.bias, addend=residual2d)
        y, _, mean, rstd = ops.add_layernorm(s, None, ln.weight, ln.bias, ln.eps, save_stats=save)
        return y, (s, mean, rstd)
    x = _lin(attn2d, dense.weight, dense.bias)
    y, s, mean, rstd = ops.add_layernorm(x, residual2d.contiguous(), ln.weight, ln.bias, ln.eps, save_stats=save, save_sum=save,
                                         dropout=dropout)
    return y, (s, mean, rstd)


def feed_forward(h2d: torch.Tensor, input2d: torch.Tensor, inter: nn.Linear, out: nn.Linear, ln: nn.LayerNorm,
                 act: str = "gelu", save: bool = False, dropout: Optional[ops.DropoutState] = None):
    """FeedForward.forward (layers/ffn.py:32-40): LN(dropout(out(act(intermediate(h)))) + input_tensor) with
    bias+GELU fused into the first GEMM's epilogue and bias+residual into the second's (dropout: see self_output)."""
    z = torch.empty((h2d.shape[0], inter.weight.shape[0]), device=h2d.device, dtype=h2d.dtype) if save else None
    a = _lin(h2d, inter.weight, inter.bias, act=act, aux=z)
    if dropout is None:
        s = _lin(a, out.weight, out.bias, addend=input2d)
        y, _, mean, rstd = ops.add_layernorm(s, None, ln.weight, ln.bias, ln.eps, save_stats=save)
    else:
        x = _lin(a, out.weight, out.bias)
        y, s, mean, rstd = ops.add_layernorm(x, input2d.contiguous(), ln.weight, ln.bias, ln.eps, save_stats=save, save_sum=save,
                                             dropout=dropout)
    return y, (z, a, s, mean, rstd)


def lm_head(h2d: torch.Tensor, dense: nn.Linear, ln: nn.LayerNorm, decoder_w: torch.Tensor, decoder_b: Optional[torch.Tensor],
            save: bool = False):
    """LMHead.forward (models/decoder.py:267-275): decoder(LN(gelu(dense(h)))). The logits buffer is
    allocated with a row stride rounded up to 8 elements so vocab sizes like 50265 keep 16-byte
    aligned rows; the returned tensor is the [:, :V] view."""
    z = torch.empty((h2d.shape[0], dense.weight.shape[0]), device=h2d.device, dtype=h2d.dtype) if save else None
    a = _lin(h2d, dense.weight, dense.bias, act="gelu", aux=z)
    n, _, mean, rstd = ops.add_layernorm(a, None, ln.weight, ln.bias, ln.eps, save_stats=save)
    V = decoder_w.shape[0]
    ld = (V + 7) // 8 * 8
    buf = torch.empty((h2d.shape[0], ld), device=h2d.device, dtype=h2d.dtype)
    logits = _lin(n, decoder_w, decoder_b, out=buf[:, :V])
    return logits, (z, a, n, mean, rstd)
