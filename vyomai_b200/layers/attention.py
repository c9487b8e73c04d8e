"""Attention layers — host-side mirror of VyomAI/layers/attention.py (and of the second copies of
DecoderAttention[Gqa] in VyomAI/models/decoder.py:44-201 that take a shared kv_cache).

Class names, constructor arguments, sub-module / parameter names (`query`, `key`, `value`, `qkv`,
`out.dense`, `out.layernorm`), `ValueError`s and forward signatures are the reference's. The
forward bodies run the fused sm_100a path:

    q/k/v nn.Linear x3 + rearrange + apply_rotary_pos_emb + cache.update + repeat_kv + SDPA + rearrange
        -> ONE vy_gemm (packed q|k|v weights, bias + RoPE + head-split + cache-append epilogue)
         + ONE vy_attn_fwd (tcgen05 flash attention; GQA by head index)        [seqlen > 1]
        or ONE swap-AB vy_gemm + ONE vy_attn_decode                             [seqlen == 1, cached]
    AttentionSelfOutput (dense + dropout + residual + LayerNorm)
        -> ONE vy_gemm (bias + residual epilogue) + ONE vy_add_layernorm_fwd            [eval / p = 0]
        or ONE vy_gemm (bias) + ONE vy_add_layernorm_fwd (Philox dropout + residual)    [.train(), p > 0]
"""
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib
from .. import functional as F
from ..functional import KVTarget, MaskSpec


def repeat_kv(hidden_states: torch.Tensor, n_rep: int) -> torch.Tensor:
    """(batch, kv_heads, seqlen, head_dim) -> (batch, kv_heads * n_rep, seqlen, head_dim)
    (reference: attention.py:8-19). Kept for API parity; the kernels never materialise it."""
    batch, num_key_value_heads, slen, head_dim = hidden_states.shape
    if n_rep == 1:
        return hidden_states
    hidden_states = hidden_states[:, :, None, :, :].expand(batch, num_key_value_heads, n_rep, slen, head_dim)
    return hidden_states.reshape(batch, num_key_value_heads * n_rep, slen, head_dim)


repeat_kv_einops = repeat_kv  # reference: attention.py:22-39 (same result, einops spelling)


def _check_heads(config) -> None:
    if config.hidden_size % config.num_attention_heads != 0:
        raise ValueError(
            f"The hidden size ({config.hidden_size}) is not a multiple of the number of attention "
            f"heads ({config.num_attention_heads})"
        )


class AttentionSelfOutput(nn.Module):
    """dense -> dropout -> LayerNorm(. + input) (reference: attention.py:42-72)."""

    def __init__(self, config, bias: Optional[bool] = True, out_features: Optional[int] = None):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size if out_features is None else out_features, bias=bias)
        self.layernorm = nn.LayerNorm(config.hidden_size, eps=getattr(config, "layer_norm_eps", 1e-6))
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states: torch.Tensor, input_tensor: torch.Tensor) -> torch.Tensor:
        shape = input_tensor.shape
        H = shape[-1]
        from ..autograd import self_output_fn
        y = self_output_fn(hidden_states.reshape(-1, hidden_states.shape[-1]), input_tensor.reshape(-1, H), self.dense,
                           self.layernorm, dropout=F.dropout_state(self, self.dropout.p))
        return y.view(shape)


class _SelfAttentionBase(nn.Module):
    """Shared machinery of the five self-attention variants."""

    is_gqa = False
    causal_hint = False

    def _setup(self, config, layer_idx: int, gqa: bool, fused_qkv: bool = False) -> None:
        _check_heads(config)
        self.layer_idx = layer_idx
        self.attention_bias = getattr(config, "attention_bias", True)
        self.num_attention_heads = config.num_attention_heads
        head = int(config.hidden_size // config.num_attention_heads)
        if gqa:
            self.head_dim = head
            self.num_key_value_heads = getattr(config, "num_key_value_heads", 4)
            self.num_key_value_groups = self.num_attention_heads // max(self.num_key_value_heads, 1)
            if self.num_attention_heads % self.num_key_value_heads != 0 or self.num_attention_heads < self.num_key_value_heads:
                raise ValueError(
                    f"num_key_value_heads {self.num_key_value_heads }  should be less than equal num_attention_heads {config.num_attention_heads} and  multiple of num_attention_heads {config.num_attention_heads} "
                )
            kv_out = self.num_key_value_heads * head
        else:
            self.head_size = head
            kv_out = config.hidden_size
        self._head = head
        self._kv_heads = self.num_key_value_heads if gqa else self.num_attention_heads
        self.flash = True
        if fused_qkv:
            self.qkv = nn.Linear(config.hidden_size, 3 * config.hidden_size)
        else:
            self.query = nn.Linear(config.hidden_size, config.hidden_size, bias=self.attention_bias)
            self.key = nn.Linear(config.hidden_size, kv_out, bias=self.attention_bias)
            self.value = nn.Linear(config.hidden_size, kv_out, bias=self.attention_bias)
        self.out = AttentionSelfOutput(config=config, bias=self.attention_bias)
        self._rope = None  # RopeTables, attached by the owning model

    def _packed(self):
        if hasattr(self, "qkv"):
            return [self.qkv]
        return [self.query, self.key, self.value]

    def _run(self, hidden_state: torch.Tensor, attention_mask, freqs, kv: Optional[KVTarget], start_pos: int) -> torch.Tensor:
        if self._head != F.HEAD_DIM:
            raise _lib.VyomError(f"the sm_100a attention path is specialised for head_dim 64 (got {self._head})")
        B, S, H = hidden_state.shape
        mask = attention_mask if isinstance(attention_mask, MaskSpec) else MaskSpec.from_dense(attention_mask, S)
        rope = None
        if freqs is not None:
            rope = self._rope_tables(freqs, hidden_state)
        from ..autograd import attention_block_fn
        y = attention_block_fn(self, hidden_state.reshape(B * S, H), B, S, mask, rope, kv, start_pos,
                               decode_no_mask=attention_mask is None)
        return y.view(B, S, H)

    def _rope_tables(self, freqs, x):
        # inside a model `freqs` is the model's RopeTables handle; through the bare layer API it is
        # the reference's (1, S, d/2) angle slice starting at the current position
        # returns (cos, sin, base): fp32 [rows, d/2] tables whose row 0 is absolute position `base`
        if isinstance(freqs, F.RopeTables):
            cos, sin = freqs.get(x.device, x.dtype)
            return cos, sin, 0
        key = (freqs.data_ptr(), tuple(freqs.shape), x.dtype)
        cached = getattr(self, "_freq_cache", None)
        if cached is None or cached[0] != key:
            f = freqs[0].to(torch.float32).cpu()
            cos = f.cos().to(x.dtype).to(torch.float32).contiguous().to(x.device)
            sin = f.sin().to(x.dtype).to(torch.float32).contiguous().to(x.device)
            self._freq_cache = (key, (cos, sin))
        cos, sin = self._freq_cache[1]
        return cos, sin, None  # None: the slice starts at the current position


class EncoderAttention(_SelfAttentionBase):
    """reference: attention.py:75-133"""

    def __init__(self, config, layer_idx: int) -> None:
        super().__init__()
        self._setup(config, layer_idx, gqa=False)

    def forward(self, hidden_state: torch.Tensor, attention_mask: torch.Tensor, freqs: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._run(hidden_state, attention_mask, freqs, None, 0)


class EncoderAttentionGqa(_SelfAttentionBase):
    """reference: attention.py:136-215"""

    is_gqa = True

    def __init__(self, config, layer_idx: int) -> None:
        super().__init__()
        self._setup(config, layer_idx, gqa=True)

    def forward(self, hidden_state: torch.Tensor, attention_mask: torch.Tensor, freqs: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._run(hidden_state, attention_mask, freqs, None, 0)


class VisionAttention(_SelfAttentionBase):
    """Fused-qkv attention of the ViT (reference: attention.py:576-624)."""

    def __init__(self, config, layer_idx: int) -> None:
        super().__init__()
        self._setup(config, layer_idx, gqa=False, fused_qkv=True)

    def forward(self, hidden_state: torch.Tensor, attention_mask: torch.Tensor, freqs: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._run(hidden_state, attention_mask, freqs, None, 0)


class _DecoderAttentionBase(_SelfAttentionBase):
    """Per-layer-cache flavour used by the captioner (reference: attention.py:218-379): the cache
    object is attached as `self.cache` by model._setup_cache()."""

    def forward(self, hidden_state: torch.Tensor, attention_mask: torch.Tensor, freqs: Optional[torch.Tensor] = None,
                use_cache: Optional[bool] = False, start_pos: Optional[int] = 0) -> torch.Tensor:
        kv = None
        if use_cache:
            cache = getattr(self, "cache", None)
            if cache is None:
                raise ValueError("you need to setup cache for every attention layer with model._setup_cache()")
            B, S, _ = hidden_state.shape
            k, v, pos = cache.slot(B, self._kv_heads, S, self._head, start_pos, hidden_state.device, self._cache_dtype(hidden_state))
            kv = KVTarget(k, v, pos)
        return self._run(hidden_state, attention_mask, freqs, kv, start_pos)

    @staticmethod
    def _cache_dtype(x: torch.Tensor) -> torch.dtype:
        return x.dtype


class DecoderAttention(_DecoderAttentionBase):
    """reference: attention.py:218-289"""

    def __init__(self, config, layer_idx: int) -> None:
        super().__init__()
        self._setup(config, layer_idx, gqa=False)


class DecoderAttentionGqa(_DecoderAttentionBase):
    """reference: attention.py:292-379"""

    is_gqa = True

    def __init__(self, config, layer_idx: int) -> None:
        super().__init__()
        self._setup(config, layer_idx, gqa=True)


class _SharedCacheDecoderAttentionBase(_SelfAttentionBase):
    """Shared-cache flavour used by DecoderModel (reference: models/decoder.py:44-201): forward takes
    the whole-model kv_cache and returns (out, kv_cache)."""

    def forward(self, hidden_state: torch.Tensor, attention_mask: torch.Tensor, freqs: Optional[torch.Tensor] = None,
                use_cache: Optional[bool] = False, kv_cache=None, start_pos: Optional[int] = 0):
        kv = None
        if use_cache:
            if kv_cache is None:
                raise ValueError("you need to pass kv_cache")
            B, S, _ = hidden_state.shape
            k, v, pos = kv_cache.slot(self.layer_idx, B, self._kv_heads, S, self._head, start_pos, hidden_state.device,
                                      hidden_state.dtype)
            kv = KVTarget(k, v, pos)
        return self._run(hidden_state, attention_mask, freqs, kv, start_pos), kv_cache


class SharedCacheDecoderAttention(_SharedCacheDecoderAttentionBase):
    def __init__(self, config, layer_idx: int) -> None:
        super().__init__()
        self._setup(config, layer_idx, gqa=False)


class SharedCacheDecoderAttentionGqa(_SharedCacheDecoderAttentionBase):
    is_gqa = True

    def __init__(self, config, layer_idx: int) -> None:
        super().__init__()
        self._setup(config, layer_idx, gqa=True)


class _CrossAttentionBase(nn.Module):
    """Cross-attention of the seq2seq decoder (reference: attention.py:382-573): queries from the decoder stream, keys /
    values from the encoder states, computed once per generation and then read back from the per-layer cache
    (`len(cache) == 0` decides, :445-462); no RoPE; the mask is the ENCODER's key-padding mask."""

    def _setup(self, config, layer_idx: int, gqa: bool) -> None:
        _check_heads(config)
        self.layer_idx = layer_idx
        self.attention_bias = getattr(config, "attention_bias", True)
        self.num_attention_heads = config.num_attention_heads
        head = int(config.hidden_size // config.num_attention_heads)
        if gqa:
            self.head_dim = head
            self.is_casual = True
            self.num_key_value_heads = getattr(config, "num_key_value_heads", 4)
            self.num_key_value_groups = self.num_attention_heads // max(self.num_key_value_heads, 1)
            if self.num_attention_heads % self.num_key_value_heads != 0 or self.num_attention_heads < self.num_key_value_heads:
                raise ValueError(
                    f"num_key_value_heads {self.num_key_value_heads }  should be less than equal num_attention_heads {config.num_attention_heads} and  multiple of num_attention_heads {config.num_attention_heads} "
                )
            kv_out = self.num_key_value_heads * head
        else:
            self.head_size = head
            kv_out = config.hidden_size
        self._head = head
        self._kv_heads = self.num_key_value_heads if gqa else self.num_attention_heads
        self.flash = True
        self.query = nn.Linear(config.hidden_size, config.hidden_size, bias=self.attention_bias)
        self.key = nn.Linear(config.hidden_size, kv_out, bias=self.attention_bias)
        self.value = nn.Linear(config.hidden_size, kv_out, bias=self.attention_bias)
        self.out = AttentionSelfOutput(config=config, bias=self.attention_bias)

    def forward(self, hidden_state: torch.Tensor, encoder_hidden_state: torch.Tensor, encoder_attention_mask: torch.Tensor,
                freqs: Optional[torch.Tensor] = None, use_cache: Optional[bool] = False) -> torch.Tensor:
        if self._head != F.HEAD_DIM:
            raise _lib.VyomError(f"the sm_100a attention path is specialised for head_dim 64 (got {self._head})")
        B, Sq, H = hidden_state.shape
        mask = encoder_attention_mask if isinstance(encoder_attention_mask, MaskSpec) else MaskSpec.from_dense(encoder_attention_mask, 1)
        if mask.causal:
            raise _lib.VyomError("cross-attention takes the encoder's key-padding mask, not a causal one")
        from ..autograd import cross_attention_block_fn
        x2d = hidden_state.reshape(B * Sq, H)
        cached = None
        enc2d, Skv = None, 0
        if use_cache:
            cache = getattr(self, "cache", None)
            if cache is None:
                raise ValueError("use_cache is True please enable model._setup_cache() to use kv-cache")
            if len(cache) != 0:
                cached = cache.get()
                Skv = cached[0].shape[2]
        if cached is None:
            Skv = encoder_hidden_state.shape[1]
            enc2d = encoder_hidden_state.reshape(B * Skv, H).to(hidden_state.dtype)
        if use_cache and cached is None:
            # first cached step: project k / v once and keep them (they do not change during generation)
            with torch.no_grad():
                k, v = F.project_kv(enc2d, B, Skv, self.key, self.value, self._kv_heads)
            cached = self.cache.update(k, v)
        y = cross_attention_block_fn(self, x2d, B, Sq, enc2d, Skv, mask, cached)
        return y.view(B, Sq, H)


class EncoderDecoderAttention(_CrossAttentionBase):
    """reference: attention.py:382-470"""

    def __init__(self, config, layer_idx: int) -> None:
        super().__init__()
        self._setup(config, layer_idx, gqa=False)


class EncoderDecoderAttentionGqa(_CrossAttentionBase):
    """reference: attention.py:473-573"""

    def __init__(self, config, layer_idx: int) -> None:
        super().__init__()
        self._setup(config, layer_idx, gqa=True)
