"""One-launch decode step (vy_decode_step, csrc/decode_step.cu): the body of DecoderModel.generate's greedy loop
(reference: models/decoder.py:470-513) as ONE persistent kernel — embedding, every layer's projections / cache append /
attention / LayerNorms / MLP, LM head, argmax, token write-back, position increment. `FusedDecodeStep` binds a model, a
StaticCacheOne and the loop's static buffers into the C-ABI parameter struct once; `launch()` enqueues one step.

There is no fallback inside: `supported()` says whether the kernel's constraints hold (bf16 weights and caches, batch <= 32,
head_dim 64, hidden size a multiple of 256 up to 1024, RoPE or no position table); the caller picks the per-op path
otherwise."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from . import _lib
from . import functional as F

# Opt-in (VY_DECODE_FUSED=1): measured on B200 (profiles/r02_decode_*.md) the one-launch step is at parity with the graph
# of per-op kernels on config 3 (GQA 415 vs 382 us / step, MHA 362 vs 432 us) — both are bound by chains of dependent
# L2 / DRAM round trips (~0.5 / ~2.5 us each on this part), not by bytes — so it does not replace the default path yet.
ENABLED = os.environ.get("VY_DECODE_FUSED", "0") != "0"
MAX_BATCH = 32


def supported(model, kv_cache, batch: int) -> bool:
    if not ENABLED:
        return False
    w = model.word_embeddings.weight
    if w.dtype != torch.bfloat16 or not w.is_cuda or batch > MAX_BATCH:
        return False
    if model.position_embeddings is not None:  # positions enter through RoPE (or not at all) on this path
        return False
    layers = model.all_layer
    if len(layers) > _lib.CONSTS["VY_DECODE_MAX_LAYERS"]:
        return False
    att = layers[0].attention
    H = w.shape[1]
    ffn = layers[0].feed_forward.intermediate.weight.shape[0]
    if att._head != 64 or H != att.num_attention_heads * 64 or H % 256 or H > 1024 or ffn % 256 or ffn > 4096:
        return False
    if (H // 256) not in (1, 2, 3, 4) or (ffn // 256) not in (1, 2, 3, 4, 8, 12, 16):
        return False
    if att.num_attention_heads // att._kv_heads not in (1, 2, 3, 4, 6, 8):
        return False
    if layers[0].feed_forward._act_name != "gelu":
        return False
    kc = kv_cache.key_cache[0]
    return kc.dtype == torch.bfloat16 and kc.is_cuda and kc.shape[1] == att._kv_heads and kc.shape[3] == 64 and kc.shape[0] >= batch


class FusedDecodeStep:
    def __init__(self, model, kv_cache, batch: int, tok: torch.Tensor, pos: torch.Tensor, tokens: Optional[torch.Tensor],
                 pos_bound: int, logits: Optional[torch.Tensor] = None, trace: Optional[torch.Tensor] = None):
        """tok int64 [B], pos int32 [1], tokens int64 [B, max_len] (or None): the loop's static device buffers."""
        L = _lib.lib()
        self.model, self.cache = model, kv_cache
        dev = tok.device
        att0 = model.all_layer[0].attention
        H = model.word_embeddings.weight.shape[1]
        ffn = model.all_layer[0].feed_forward.intermediate.weight.shape[0]
        nbytes = L.vy_decode_step_workspace_bytes(batch, H, att0.num_attention_heads, att0._kv_heads, ffn)
        self.workspace = torch.zeros(nbytes + 256, dtype=torch.uint8, device=dev)  # zeroed once: holds the barrier counters
        ws_ptr = (self.workspace.data_ptr() + 255) // 256 * 256
        n_layers = len(model.all_layer)
        self._keep = []  # tensors whose addresses are baked into the struct
        arr = (_lib.STRUCTS["VyDecodeLayer"] * n_layers)()
        ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        for li, layer in enumerate(model.all_layer):
            att, ff = layer.attention, layer.feed_forward
            w_qkv, b_qkv = F.pack_linears(att._packed())
            if not w_qkv.is_contiguous():
                raise _lib.VyomError("decode step: packed q|k|v weights must be contiguous")
            kc, vc = kv_cache.key_cache[li], kv_cache.value_cache[li]
            if kc.stride() != vc.stride() or kc.stride() != kv_cache.key_cache[0].stride():
                raise _lib.VyomError("decode step: every layer's caches must share one layout")
            self._keep += [w_qkv, b_qkv, kc, vc]
            e = arr[li]
            e.w_qkv, e.b_qkv = ptr(w_qkv), ptr(b_qkv)
            e.w_o, e.b_o = ptr(att.out.dense.weight), ptr(att.out.dense.bias)
            e.ln1_g, e.ln1_b = ptr(att.out.layernorm.weight), ptr(att.out.layernorm.bias)
            e.w_1, e.b_1 = ptr(ff.intermediate.weight), ptr(ff.intermediate.bias)
            e.w_2, e.b_2 = ptr(ff.out.weight), ptr(ff.out.bias)
            e.ln2_g, e.ln2_b = ptr(ff.layernorm.weight), ptr(ff.layernorm.bias)
            e.k_cache, e.v_cache = ptr(kc), ptr(vc)
        self._layers = arr
        cos = sin = None
        if model._rope is not None:
            cos, sin = model._rope.get(dev, torch.bfloat16)
        self._keep += [cos, sin, tok, pos, tokens, logits, trace]
        head = model.lm_head
        kc0 = kv_cache.key_cache[0]
        st = _lib.STRUCTS["VyDecodeStep"]()
        fields = dict(
            B=batch, H=H, n_q_heads=att0.num_attention_heads, n_kv_heads=att0._kv_heads, head_dim=64, ffn=ffn,
            vocab=head.decoder.weight.shape[0], n_layers=n_layers, layers=ctypes.addressof(arr),
            emb=model.word_embeddings.weight.data_ptr(), pos_table=None, rope_cos=ptr(cos), rope_sin=ptr(sin),
            rope_rows=0 if cos is None else cos.shape[0], w_d=head.dense.weight.data_ptr(), b_d=ptr(head.dense.bias),
            ln_head_g=head.layer_norm.weight.data_ptr(), ln_head_b=head.layer_norm.bias.data_ptr(),
            w_v=head.decoder.weight.data_ptr(), b_v=ptr(head.bias), eps_layer=float(att0.out.layernorm.eps),
            eps_head=float(head.layer_norm.eps), cache_len=kc0.shape[2], cache_sb=kc0.stride(0), cache_sh=kc0.stride(1),
            cache_sl=kc0.stride(2), pos_bound=pos_bound, pos=pos.data_ptr(), tok=tok.data_ptr(),
            tokens_out=ptr(tokens), ld_tokens=0 if tokens is None else tokens.stride(0), logits=ptr(logits),
            ld_logits=0 if logits is None else logits.stride(0), workspace=ws_ptr, workspace_bytes=nbytes, trace=ptr(trace),
        )
        for k, v in fields.items():
            if v is not None:
                setattr(st, k, v)
        self._st = st
        self._ws_ptr = ws_ptr

    def launch(self) -> None:
        """Enqueue one decode step on the current stream."""
        self._st.stream = torch.cuda.current_stream().cuda_stream
        fn = _lib.lib().vy_decode_step
        rc = _lib._timed("vy_decode_step", {}, lambda: fn(ctypes.byref(self._st)))
        _lib.check(rc, "vy_decode_step")

    def check(self) -> None:
        """Synchronising health check: raises if a grid barrier of an earlier step timed out."""
        if _lib.lib().vy_decode_step_status(self._ws_ptr) != 0:
            raise _lib.VyomError("vy_decode_step: a grid-wide barrier timed out; the generated tokens are invalid")
