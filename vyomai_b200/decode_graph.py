"""Greedy decoding with ONE CUDA-graph replay per token.

The reference's generate loop (models/decoder.py:470-513) runs one Python forward per token: ~30 kernel launches and
a host sync (`all(eos_reached)`) each, which on a B200 costs ~1 ms of host time for ~60 us of GPU work. Here the whole
single-token step — embedding, every layer's swap-AB projections, vy_attn_decode (RoPE + cache append + attention with
the position read from DEVICE memory), LayerNorms, LM head, argmax, token write-back, position increment — is captured
once and replayed; the host only enqueues replays and looks at the eos flags every few dozen tokens.
"""
from __future__ import annotations

import os

import torch

from . import _lib, decode_step
from . import functional as F
from . import ops


class GreedyDecodeGraph:
    def __init__(self, model, kv_cache, batch: int, max_len: int):
        if model._rope is None:
            raise ValueError("GreedyDecodeGraph supports RoPE models (positions enter only through the attention kernel)")
        self.model, self.cache, self.B, self.max_len = model, kv_cache, batch, max_len
        dev = model.word_embeddings.weight.device
        self.tok = torch.zeros(batch, dtype=torch.long, device=dev)          # token fed to the next step
        self.pos = torch.zeros(1, dtype=torch.int32, device=dev)             # cache slot / position of that token
        self.tokens = torch.zeros((batch, max_len), dtype=torch.long, device=dev)
        self.graph = None
        # the whole step as ONE persistent kernel (csrc/decode_step.cu) when the model fits its constraints
        self.fused = None
        if decode_step.supported(model, kv_cache, batch):
            self.fused = decode_step.FusedDecodeStep(model, kv_cache, batch, self.tok, self.pos, self.tokens,
                                                     pos_bound=kv_cache.key_cache[0].shape[2] - 1)

    # one decode step on the static buffers (eager or under capture)
    def _step(self) -> None:
        if self.fused is not None:
            self.fused.launch()
            return
        m = self.model
        B = self.B
        T = m.word_embeddings.weight.dtype
        x = m._embed(self.tok.view(B, 1), 0)  # [B, H]; RoPE models add no position rows
        cos, sin = m._rope.get(x.device, T)
        limit = self.cache.key_cache[0].shape[2] - 1
        for li, layer in enumerate(m.all_layer):
            att = layer.attention
            w_qkv, b_qkv = F.pack_linears(att._packed())
            qkv = F._lin(x, w_qkv, b_qkv)
            a = ops.attn_decode(qkv, self.cache.key_cache[li], self.cache.value_cache[li], limit, att.num_attention_heads,
                                att._kv_heads, cos, sin, out_dtype=T, start_pos_dev=self.pos)
            y, _ = F.self_output(a, x, att.out.dense, att.out.layernorm)
            ff = layer.feed_forward
            x, _ = F.feed_forward(y, x, ff.intermediate, ff.out, ff.layernorm, act=ff._act_name)
        head = m.lm_head
        logits, _ = F.lm_head(x, head.dense, head.layer_norm, head.decoder.weight, head.bias)
        ops.argmax_advance(logits, self.tok, self.pos, self.tokens)  # tok = argmax; pos += 1; tokens[:, pos] = tok

    def capture(self) -> None:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        pos0, tok0 = self.pos.clone(), self.tok.clone()
        with torch.cuda.stream(side), torch.no_grad():
            self._step()  # warm-up: allocator pools, tensor-map cache, lazy inits
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.pos.copy_(pos0)
        self.tok.copy_(tok0)
        self.graph = torch.cuda.CUDAGraph()
        # the step is a chain of ~30 short kernels: programmatic dependent launch lets each one's prologue overlap its
        # predecessor's tail (-21 us per step, profiles/r02_decode_attn_ab.txt). VY_DECODE_PDL=0 keeps plain stream order.
        pdl = os.environ.get("VY_DECODE_PDL", "1") != "0"
        prev = _lib.lib().vy_set_pdl(1) if pdl else None
        try:
            with torch.cuda.graph(self.graph), torch.no_grad():
                self._step()
        finally:
            if pdl:
                _lib.lib().vy_set_pdl(prev)
        self.pos.copy_(pos0)
        self.tok.copy_(tok0)

    def run(self, first_token: torch.Tensor, start_pos: int, steps: int) -> None:
        """Feeds `first_token` ([B], the token at position start_pos, not yet in the cache) and generates `steps` more;
        token i lands in self.tokens[:, start_pos + 1 + i]."""
        self.tok.copy_(first_token)
        self.pos.fill_(start_pos)
        if self.graph is None:
            self.capture()
        for _ in range(steps):
            self.graph.replay()
        if self.fused is not None:
            self.fused.check()
