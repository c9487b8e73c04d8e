"""KV caches — host-side mirror of VyomAI/layers/kv_cache.py.

Same four classes, constructor signatures, `update` / `get` / `__len__` semantics and the
[batch, kv_heads, seq, head_dim] layout. What changes is who writes them: inside the models the
QKV-projection epilogue (prefill) and vy_attn_decode (single token) append rotated k / raw v
straight into the cache storage through `slot()`, so nothing is concatenated or copied per
token. The public `update()` keeps working for callers that hold materialised k/v tensors; it
copies them in with the vy_cast4d kernel.

The dynamic caches grow geometrically instead of `torch.cat`-ing every step
(kv_cache.py:56-57,229-234 do an O(context) copy per token); `key_cache` / `value_cache` expose the
valid prefix as views, so observable contents and shapes match the reference.
"""
from typing import List, Optional, Tuple

import torch

from .. import ops


def _heads(config, is_gqa_attr_only: bool, is_gqa: bool) -> int:
    if is_gqa_attr_only:
        # StaticCacheOne sizes its heads from config.num_key_value_heads whenever the attribute
        # exists, ignoring is_gqa (kv_cache.py:275-282)
        h = getattr(config, "num_key_value_heads", None)
        return config.num_attention_heads if h is None else h
    if is_gqa:
        h = getattr(config, "num_key_value_heads", None)
        if h is None:
            raise ValueError("you are using is_gqa=True and config.num_key_value_heads is not available")
        return h
    return config.num_attention_heads


def _copy_in(dst: torch.Tensor, src: torch.Tensor) -> None:
    ops.cast4d(src if src.stride(3) == 1 else src.contiguous(), dst.dtype, out=dst)


class _Growable:
    """One layer's k/v storage with amortised growth. Valid region: [:, :, :length]."""

    def __init__(self) -> None:
        self.k: Optional[torch.Tensor] = None
        self.v: Optional[torch.Tensor] = None
        self.length = 0

    def reserve(self, batch: int, heads: int, need: int, head_dim: int, device, dtype) -> None:
        if self.k is not None and self.k.shape[2] >= need and self.k.shape[0] >= batch:
            return
        cap = max(need, 64)
        if self.k is not None:
            cap = max(cap, 2 * self.k.shape[2])
        nk = torch.zeros((batch, heads, cap, head_dim), device=device, dtype=dtype)
        nv = torch.zeros((batch, heads, cap, head_dim), device=device, dtype=dtype)
        if self.k is not None and self.length > 0:
            _copy_in(nk[: self.k.shape[0], :, : self.length], self.k[:, :, : self.length])
            _copy_in(nv[: self.v.shape[0], :, : self.length], self.v[:, :, : self.length])
        self.k, self.v = nk, nv


class DynamicCache:
    """Per-layer growing cache (reference: kv_cache.py:11-78)."""

    def __init__(self, config, is_gqa: Optional[bool] = False) -> None:
        self._store = _Growable()
        self._seen_tokens = False

    @property
    def key_cache(self):
        return None if self._store.k is None else self._store.k[:, :, : self._store.length]

    @property
    def value_cache(self):
        return None if self._store.v is None else self._store.v[:, :, : self._store.length]

    def __len__(self) -> int:
        return self._store.length

    def slot(self, batch: int, heads: int, seqlen: int, head_dim: int, start_pos: int, device, dtype):
        """Storage the fused kernels append into: rows [len, len+seqlen) of the returned buffers.
        Like the reference's torch.cat the dynamic cache appends at its current length."""
        pos = self._store.length
        self._store.reserve(batch, heads, pos + seqlen, head_dim, device, dtype)
        self._store.length = pos + seqlen
        self._seen_tokens = True
        return self._store.k, self._store.v, pos

    def update(self, key_states: torch.Tensor, value_states: torch.Tensor, start_pos: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        b, h, s, d = key_states.shape
        k, v, pos = self.slot(b, h, s, d, start_pos, key_states.device, key_states.dtype)
        _copy_in(k[:b, :, pos:pos + s], key_states)
        _copy_in(v[:b, :, pos:pos + s], value_states)
        return self.key_cache, self.value_cache

    def get(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._seen_tokens:
            return self.key_cache, self.value_cache
        raise ValueError("there is no token available in kv-cache")

    def get_seq_length(self, layer_idx: Optional[int] = 0) -> int:
        return self._store.length

    def get_max_length(self) -> Optional[int]:
        return None


class StaticCache:
    """Per-layer fixed-size cache, batch 1 only (reference: kv_cache.py:81-168): zeros
    (1, heads, max_position_embeddings, head_dim) that follow the first k's device/dtype."""

    def __init__(self, config, is_gqa: Optional[bool] = False) -> None:
        self.head_size = int(config.hidden_size // config.num_attention_heads)
        self.heads = _heads(config, False, bool(is_gqa))
        self.max_len = config.max_position_embeddings
        self.key_cache: torch.Tensor = torch.zeros(1, self.heads, self.max_len, self.head_size)
        self.value_cache: torch.Tensor = torch.zeros(1, self.heads, self.max_len, self.head_size)
        self._seen_tokens = False
        self.first_update_len = 0

    def slot(self, batch: int, heads: int, seqlen: int, head_dim: int, start_pos: int, device, dtype):
        if seqlen > self.key_cache.size()[2] or start_pos + seqlen > self.key_cache.size()[2]:
            # the reference fails here too (on the slice assignment at kv_cache.py:142-145 once start_pos + seqlen runs
            # past the buffer); the fused append must never be handed a slot range outside the cache
            raise ValueError(f"{(batch, heads, seqlen, head_dim)} at position {start_pos} is more than init k_cache size {self.key_cache.shape}")
        assert batch == 1, "Only support batch size 1"
        if self.key_cache.device != device or self.key_cache.dtype != dtype:
            self.key_cache = self.key_cache.to(device=device, dtype=dtype)
            self.value_cache = self.value_cache.to(device=device, dtype=dtype)
        self._seen_tokens = True
        self.first_update_len = seqlen
        return self.key_cache, self.value_cache, start_pos

    def update(self, k: torch.Tensor, v: torch.Tensor, start_pos: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        bsz, head, seqlen, d = k.shape
        kc, vc, pos = self.slot(bsz, head, seqlen, d, start_pos, k.device, k.dtype)
        _copy_in(kc[:bsz, :, pos:pos + seqlen], k)
        _copy_in(vc[:bsz, :, pos:pos + seqlen], v)
        return kc[:bsz, :, : pos + seqlen], vc[:bsz, :, : pos + seqlen]

    def get(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._seen_tokens:
            return self.key_cache[:, :, : self.first_update_len], self.value_cache[:, :, : self.first_update_len]
        raise ValueError("there is no token available in kv-cache")

    def __len__(self) -> int:
        if self._seen_tokens is False:
            return 0
        return self.key_cache.shape[2]


class DynamicCacheOne:
    """Whole-model growing cache, one entry per layer (reference: kv_cache.py:171-252)."""

    def __init__(self, config, is_gqa: bool = False) -> None:
        self.layers = config.num_hidden_layers
        self._stores = [_Growable() for _ in range(self.layers)]
        self._seen_tokens = False

    @property
    def key_cache(self) -> List:
        return [[] if s.k is None else s.k[:, :, : s.length] for s in self._stores]

    @property
    def value_cache(self) -> List:
        return [[] if s.v is None else s.v[:, :, : s.length] for s in self._stores]

    def __len__(self) -> int:
        return self._stores[0].length

    def slot(self, index: int, batch: int, heads: int, seqlen: int, head_dim: int, start_pos: int, device, dtype):
        st = self._stores[index]
        pos = st.length
        st.reserve(batch, heads, pos + seqlen, head_dim, device, dtype)
        st.length = pos + seqlen
        self._seen_tokens = True
        return st.k, st.v, pos

    def update(self, index: int, key_states: torch.Tensor, value_states: torch.Tensor, start_pos: int = 0):
        b, h, s, d = key_states.shape
        k, v, pos = self.slot(index, b, h, s, d, start_pos, key_states.device, key_states.dtype)
        _copy_in(k[:b, :, pos:pos + s], key_states)
        _copy_in(v[:b, :, pos:pos + s], value_states)
        st = self._stores[index]
        return st.k[:, :, : st.length], st.v[:, :, : st.length]

    def get(self, index: int):
        if self._seen_tokens:
            st = self._stores[index]
            return st.k[:, :, : st.length], st.v[:, :, : st.length]
        raise ValueError("there is no token available in kv-cache")

    def get_seq_length(self, layer_idx: Optional[int] = 0) -> int:
        return self._stores[layer_idx].length

    def get_max_length(self) -> Optional[int]:
        return None


class StaticCacheOne:
    """Whole-model preallocated cache (reference: kv_cache.py:255-377): zeros
    (batch_size, kv_heads, max_cache_len, head_dim) per layer on the CUDA device, `update` writes
    [start_pos, start_pos + S) and returns the views [:B, :, :start_pos + S]."""

    def __init__(self, config, max_cache_len: int = None, dtype: torch.dtype = torch.float32, batch_size: int = 1,
                 is_gqa: bool = False) -> None:
        self.head_size = int(config.hidden_size // config.num_attention_heads)
        self.batch_size = batch_size
        self.heads = _heads(config, True, is_gqa)
        self.max_cache_len = config.max_position_embeddings if max_cache_len is None else max_cache_len
        self.dtype = dtype
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.cache_shape = (self.batch_size, self.heads, self.max_cache_len, self.head_size)
        self._seen_tokens = False
        self.layers = config.num_hidden_layers
        self.key_cache: List[torch.Tensor] = []
        self.value_cache: List[torch.Tensor] = []
        for _ in range(self.layers):
            self.key_cache.append(torch.zeros(self.cache_shape, dtype=self.dtype, device=self.device))
            self.value_cache.append(torch.zeros(self.cache_shape, dtype=self.dtype, device=self.device))

    def slot(self, index: int, batch: int, heads: int, seqlen: int, head_dim: int, start_pos: int, device, dtype):
        kc = self.key_cache[index]
        if seqlen > kc.size()[2] or start_pos + seqlen > kc.size()[2]:
            raise ValueError(f"{(batch, heads, seqlen, head_dim)} at {start_pos} is more than init k_cache size {tuple(kc.shape)}")
        if heads != kc.shape[1] or batch > kc.shape[0]:
            raise ValueError(f"k of shape {(batch, heads, seqlen, head_dim)} does not fit the cache {tuple(kc.shape)}")
        self._seen_tokens = True
        return kc, self.value_cache[index], start_pos

    def update(self, index: int, key_states: torch.Tensor, value_states: torch.Tensor, start_pos: int = 0):
        bsz, head, seqlen, d = key_states.shape
        kc, vc, pos = self.slot(index, bsz, head, seqlen, d, start_pos, key_states.device, key_states.dtype)
        _copy_in(kc[:bsz, :, pos:pos + seqlen], key_states)
        _copy_in(vc[:bsz, :, pos:pos + seqlen], value_states)
        return kc[:bsz, :, : pos + seqlen], vc[:bsz, :, : pos + seqlen]

    def get(self, index: int):
        if self._seen_tokens:
            return self.key_cache[index], self.value_cache[index]
        raise ValueError("there is no token available in kv-cache")

    def get_seq_length(self, layer_idx: Optional[int] = 0) -> int:
        return self.key_cache[layer_idx].shape[-2]

    def get_max_length(self) -> Optional[int]:
        return None
