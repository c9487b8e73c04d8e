"""Autograd glue: the fused blocks as torch.autograd.Functions.

Forward and backward are both sequences of C-ABI kernel calls (ops.py); torch only records the
graph edges. When no gradient is required (inference, generation) the functions call the forward
kernels directly and save nothing.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops
from . import functional as F
from .functional import KVTarget, MaskSpec


def _needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


def _qkv_params(mod):
    lin = mod._packed()
    ws = [l.weight for l in lin]
    bs = [l.bias for l in lin if l.bias is not None]
    return lin, ws, bs


# ----------------------------------------------------------------------------------------------
# attention block: qkv projection + RoPE + attention + output projection + residual + LayerNorm
# ----------------------------------------------------------------------------------------------
def attention_block_fn(mod, x2d: torch.Tensor, B: int, S: int, mask: MaskSpec, rope, kv: Optional[KVTarget],
                       start_pos: int, decode_no_mask: bool) -> torch.Tensor:
    lin, ws, bs = _qkv_params(mod)
    dense, ln = mod.out.dense, mod.out.layernorm
    if _needs_grad(x2d, *ws, dense.weight, ln.weight):
        if kv is not None:
            raise _lib.VyomError("gradients through a kv-cache forward are not supported (use_cache=True is an inference path)")
        from .autograd_train import AttentionBlockFn
        flat = [x2d] + ws + bs + [dense.weight] + ([dense.bias] if dense.bias is not None else []) + [ln.weight, ln.bias]
        return AttentionBlockFn.apply(mod, B, S, mask, rope, start_pos, *flat)
    w_qkv, b_qkv = F.pack_linears(lin)
    attn, _ = F.attention_core(x2d, B, S, w_qkv, b_qkv, mod.num_attention_heads, mod._kv_heads, mask, rope, kv,
                               decode_no_mask, pos0=start_pos)
    y, _ = F.self_output(attn, x2d, dense, ln, dropout=F.dropout_state(mod.out, mod.out.dropout.p))
    return y


def cross_attention_block_fn(mod, x2d: torch.Tensor, B: int, Sq: int, enc2d: Optional[torch.Tensor], Skv: int, mask: MaskSpec,
                             cached_kv=None) -> torch.Tensor:
    """EncoderDecoderAttention[Gqa].forward: LN(dense(cross_attention(x, enc)) + x)."""
    dense, ln = mod.out.dense, mod.out.layernorm
    drop = F.dropout_state(mod.out, mod.out.dropout.p)
    if cached_kv is None and _needs_grad(x2d, enc2d, mod.query.weight, mod.key.weight, dense.weight, ln.weight):
        from .autograd_train import CrossAttentionBlockFn
        return CrossAttentionBlockFn.apply(mod, B, Sq, Skv, mask, drop, x2d, enc2d, mod.query.weight, mod.query.bias, mod.key.weight,
                                           mod.key.bias, mod.value.weight, mod.value.bias, dense.weight, dense.bias, ln.weight, ln.bias)
    attn, _ = F.cross_attention_core(x2d, B, Sq, enc2d, Skv, mod.query, mod.key, mod.value, mod.num_attention_heads, mod._kv_heads,
                                     mask, cached_kv)
    y, _ = F.self_output(attn, x2d, dense, ln, dropout=drop)
    return y


def linear_fn(x2d: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """y = x W^T + b on vy_gemm, differentiable (plain nn.Linear semantics for callers outside the fused blocks)."""
    if _needs_grad(x2d, w, b):
        from .autograd_train import LinearFn
        return LinearFn.apply(x2d, w, b)
    return F._lin(x2d, w, b)


def self_output_fn(attn2d: torch.Tensor, residual2d: torch.Tensor, dense: nn.Linear, ln: nn.LayerNorm,
                   dropout=None) -> torch.Tensor:
    if _needs_grad(attn2d, residual2d, dense.weight, ln.weight):
        from .autograd_train import SelfOutputFn
        return SelfOutputFn.apply(attn2d, residual2d, dense.weight, dense.bias, ln.weight, ln.bias, ln.eps, dropout)
    y, _ = F.self_output(attn2d, residual2d, dense, ln, dropout=dropout)
    return y


def feed_forward_fn(mod, h2d: torch.Tensor, input2d: torch.Tensor) -> torch.Tensor:
    inter, out, ln = mod.intermediate, mod.out, mod.layernorm
    if _needs_grad(h2d, input2d, inter.weight, out.weight, ln.weight):
        from .autograd_train import FeedForwardFn
        return FeedForwardFn.apply(mod._act_name, ln.eps, F.dropout_state(mod, mod.dropout.p), h2d, input2d, inter.weight,
                                   inter.bias, out.weight, out.bias, ln.weight, ln.bias)
    y, _ = F.feed_forward(h2d, input2d, inter, out, ln, act=mod._act_name, dropout=F.dropout_state(mod, mod.dropout.p))
    return y


def lm_head_fn(mod, h2d: torch.Tensor) -> torch.Tensor:
    dense, ln, dec = mod.dense, mod.layer_norm, mod.decoder
    if _needs_grad(h2d, dense.weight, dec.weight, ln.weight):
        from .autograd_train import LMHeadFn
        return LMHeadFn.apply(ln.eps, h2d, dense.weight, dense.bias, ln.weight, ln.bias, dec.weight, mod.bias)
    logits, _ = F.lm_head(h2d, dense, ln, dec.weight, mod.bias)
    return logits


def lm_head_loss_fn(mod, h2d: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """Mean token cross-entropy of the LM head's logits against `labels` (one per row of h2d) as one autograd node."""
    from .autograd_train import LMHeadLossFn
    dense, ln, dec = mod.dense, mod.layer_norm, mod.decoder
    return LMHeadLossFn.apply(ln.eps, ignore_index, h2d, dense.weight, dense.bias, ln.weight, ln.bias, dec.weight, mod.bias, labels)


def embed_fn(ids: torch.Tensor, table: torch.Tensor, pos_table: Optional[torch.Tensor], pos_row_off: int,
             tokens_per_seq: int, extra: Optional[torch.Tensor] = None, padding_idx: Optional[int] = None,
             pos_padding_idx: Optional[int] = None) -> torch.Tensor:
    """hidden rows [B * (tokens_per_seq + (extra is not None)), H] = table[ids] (+ position rows), with an
    optional leading row per sequence copied from `extra` [B, H] (the captioner's image vector). `padding_idx` /
    `pos_padding_idx`: rows of the word / learned-position tables that nn.Embedding(padding_idx=...) keeps gradient-free."""
    from .autograd_train import EmbedFn
    if _needs_grad(table, pos_table, extra):
        return EmbedFn.apply(ids, table, pos_table, pos_row_off, tokens_per_seq, extra, padding_idx, pos_padding_idx)
    return EmbedFn.forward(_NoCtx(), ids, table, pos_table, pos_row_off, tokens_per_seq, extra, padding_idx, pos_padding_idx)


class SlotMergeFn(torch.autograd.Function):
    """rows = where(slot >= 0, image_rows[slot], word_rows): the masked_scatter of image features into the <image> token
    positions (Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1, VisionLanguageModel.forward). Gradients: the word
    rows get dout with the image positions zeroed (masked_scatter overwrote them), the image rows get their dout rows."""

    @staticmethod
    def forward(ctx, word_rows, image_rows, slot):
        ctx.save_for_backward(slot)
        ctx.n_img = image_rows.shape[0]
        return ops.slot_merge(word_rows.contiguous(), image_rows.contiguous(), slot)

    @staticmethod
    def backward(ctx, dout):
        (slot,) = ctx.saved_tensors
        da, db = ops.slot_merge_bwd(dout.contiguous(), slot, ctx.n_img, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return da, db, None


def slot_merge_fn(word_rows: torch.Tensor, image_rows: torch.Tensor, slot: torch.Tensor) -> torch.Tensor:
    if _needs_grad(word_rows, image_rows):
        return SlotMergeFn.apply(word_rows, image_rows, slot)
    return ops.slot_merge(word_rows.contiguous(), image_rows.contiguous(), slot)


class _NoCtx:
    """Stand-in ctx so a Function's forward can be reused on the no-grad path without autograd."""

    needs_input_grad = ()

    def save_for_backward(self, *a):
        pass


def vit_stem(mod, pixels: torch.Tensor):
    """ViT stem forward (no grad bookkeeping). Returns (hidden [B*(nP+1), H], patches [B*nP, K])."""
    w4 = mod.pixel_seq.weight
    T = w4.dtype
    B = pixels.shape[0]
    nP = mod.num_patches
    H = w4.shape[0]
    w2 = w4.view(H, -1)
    if mod.cls_token.shape[-1] != H:
        raise _lib.VyomError("Vit: cls_token width (C*p*p) must equal hidden_size, as in the reference's torch.cat")
    patches = ops.patchify(pixels.contiguous(), tuple(mod.patch_size), T)
    pos = mod.position_embeddings.pos_embeddings.view(nP + 1, -1)
    hidden = torch.empty((B * (nP + 1), H), device=pixels.device, dtype=T)
    # patch rows: 2 * (conv(patch) + bias + pos[1 + p]) at row b*(nP+1) + 1 + p
    ops.gemm(patches, w2, bias=mod.pixel_seq.bias, addend=pos, addend_row_mod=nP, addend_row_off=1, out=hidden,
             out_scale=2.0, out_row_group=nP, out_row_group_stride=nP + 1, out_row_off=1)
    # cls rows: 2 * (cls + pos[0]) at row b*(nP+1)
    ops.embed(None, mod.cls_token.view(1, -1), out=hidden, rows=B, tokens_per_seq=1, out_group_stride=nP + 1,
              out_row_off=0, pos=pos, pos_row_off=0, out_scale=2.0, broadcast_src=True)
    return hidden, patches


def vit_stem_fn(mod, pixels: torch.Tensor) -> torch.Tensor:
    if _needs_grad(mod.pixel_seq.weight, mod.cls_token, mod.position_embeddings.pos_embeddings):
        from .autograd_train import VitStemFn
        return VitStemFn.apply(mod, pixels, mod.pixel_seq.weight, mod.pixel_seq.bias, mod.cls_token,
                               mod.position_embeddings.pos_embeddings)
    hidden, _ = vit_stem(mod, pixels)
    return hidden
