"""Paged kv-cache and continuous batching for the decoder family.

Host-side mirror of `Examples/simple_vllm.ipynb` cell 2 of the reference — `SequenceState`, `PagedKVManager`
(`can_allocate` / `allocate` / `free`, pools of shape (max_blocks, block_size, kv_heads, head_dim) per layer) and
`ContinuousBatchEngine` (`add_sequence`, `step`, waiting room, greedy next token, eos / length retirement) — on top of
`vy_attn_decode`'s paged mode: one kernel launch per layer does the new token's RoPE, writes its k/v into the
sequence's current block (the notebook's `k_cache[slots // block_size, slots % block_size] = k`) and attends over the
blocks named by the block table with a per-row context length (the notebook's
`flash_attn_with_kvcache(q, k_cache, v_cache, cache_seqlens=..., block_table=..., causal=True)`).

Differences from the notebook, on purpose:
  * the model is this package's `DecoderModel` (head_dim 64, LayerNorm / GELU blocks), not the notebook's Qwen3 clone;
  * a step that holds both prefilling and decoding sequences runs the prefills (one forward per new sequence, k/v
    scattered into its blocks) AND a proper paged decode for the others — the notebook sends such a mixed batch through
    `flash_attn_varlen_func` over the new tokens only, so a decoding sequence attends to nothing but its own token
    during that step;
  * eos ids are a constructor argument instead of the hard-wired Qwen ids.
"""
from __future__ import annotations

import itertools
from collections import deque
from typing import Dict, Iterable, List

import torch

from . import functional as F
from . import ops
from .layers.kv_cache import StaticCacheOne


class SequenceState:
    """Runtime state of one request (notebook: SequenceState)."""

    def __init__(self, sid: int, prompt_ids: List[int], max_gen_len: int, block_size: int, device):
        self.id, self.device, self.block_size = sid, device, block_size
        p_len = len(prompt_ids)
        self.max_total_len = p_len + max_gen_len
        self.tokens = torch.zeros(self.max_total_len, dtype=torch.long, device=device)
        self.tokens[:p_len] = torch.tensor(prompt_ids, dtype=torch.long, device=device)
        self.num_tokens, self.is_prefill = p_len, True
        max_blocks = (self.max_total_len + block_size - 1) // block_size
        self.block_table = torch.zeros(max_blocks, dtype=torch.int32, device=device)  # which pool blocks hold this sequence
        self.block_count = 0

    def slots(self, start: int, end: int) -> torch.Tensor:
        """Pool slot (block * block_size + offset) of token positions [start, end)."""
        idx = torch.arange(start, end, device=self.device)
        return self.block_table[idx // self.block_size].long() * self.block_size + idx % self.block_size


class PagedKVManager:
    """Block pools (max_blocks, block_size, kv_heads, 64) per layer and a free list (notebook: PagedKVManager)."""

    def __init__(self, model, max_blocks: int, block_size: int, dtype=None, device=None):
        att = model.all_layer[0].attention
        self.block_size = block_size
        self.kv_heads, self.head_dim = att._kv_heads, F.HEAD_DIM
        dev = device or model.word_embeddings.weight.device
        dtype = dtype or model.word_embeddings.weight.dtype
        self.free_blocks = deque(range(max_blocks))
        shape = (max_blocks, block_size, self.kv_heads, self.head_dim)
        self.k_cache = [torch.zeros(shape, device=dev, dtype=dtype) for _ in model.all_layer]
        self.v_cache = [torch.zeros(shape, device=dev, dtype=dtype) for _ in model.all_layer]

    def can_allocate(self, num_tokens: int) -> bool:
        """Enough free blocks for a prompt of `num_tokens`?"""
        return len(self.free_blocks) >= (num_tokens + self.block_size - 1) // self.block_size

    def allocate(self, state: SequenceState) -> None:
        needed = (state.num_tokens + self.block_size - 1) // self.block_size
        while state.block_count < needed:
            if not self.free_blocks:
                raise RuntimeError("KV Cache full! (Engine should have prevented this)")
            state.block_table[state.block_count] = self.free_blocks.popleft()
            state.block_count += 1

    def free(self, state: SequenceState) -> None:
        for i in range(state.block_count):
            self.free_blocks.append(int(state.block_table[i]))
        state.block_count = 0


class ContinuousBatchEngine:
    """Greedy continuous batching over a paged cache (notebook: ContinuousBatchEngine)."""

    def __init__(self, model, kv_mgr: PagedKVManager, max_batch_size: int = 16, eos_token_ids: Iterable[int] = ()):
        if model._rope is None:
            raise ValueError("ContinuousBatchEngine supports RoPE models (positions enter only through the attention kernel)")
        self.model = model.eval()
        self.kv_mgr = kv_mgr
        self.active: Dict[int, SequenceState] = {}
        self.id_gen = itertools.count()
        self.device = model.word_embeddings.weight.device
        self.max_batch = max_batch_size
        self.eos = set(int(t) for t in eos_token_ids)
        self.waiting_room: deque = deque()

    def add_sequence(self, prompt_ids: List[int], max_gen_len: int = 128) -> int:
        """Queues a request; it becomes active when the pool has blocks for its prompt."""
        # every position the sequence can reach must have a row in the model's position / RoPE tables: the decode kernel
        # indexes them without passing through DecoderModel.forward's own check
        self.model._check_positions(len(prompt_ids) + max_gen_len)
        sid = next(self.id_gen)
        self.waiting_room.append({"sid": sid, "prompt_ids": list(prompt_ids), "max_gen_len": max_gen_len})
        return sid

    def _try_schedule_waiting(self) -> None:
        while self.waiting_room and len(self.active) < self.max_batch:
            req = self.waiting_room[0]
            if not self.kv_mgr.can_allocate(len(req["prompt_ids"])):
                break  # pool full: retry after some sequence retires
            self.waiting_room.popleft()
            self.active[req["sid"]] = SequenceState(req["sid"], req["prompt_ids"], req["max_gen_len"], self.kv_mgr.block_size,
                                                    self.device)

    # ---- one prompt: ordinary prefill, then its k/v rows are scattered into the sequence's blocks ----
    @torch.no_grad()
    def _prefill(self, s: SequenceState) -> int:
        m, mgr = self.model, self.kv_mgr
        P = s.num_tokens
        cache = StaticCacheOne(m.config, max_cache_len=P, batch_size=1, dtype=mgr.k_cache[0].dtype)
        if cache.heads != mgr.kv_heads:  # the cache sizes its heads from the config alone (quirk Q12): follow the model
            shape = (1, mgr.kv_heads, P, mgr.head_dim)
            cache.key_cache = [torch.zeros(shape, dtype=cache.dtype, device=cache.device) for _ in m.all_layer]
            cache.value_cache = [torch.zeros(shape, dtype=cache.dtype, device=cache.device) for _ in m.all_layer]
        ids = s.tokens[:P].view(1, P)
        out = m(ids, attention_mask=torch.ones_like(ids), use_cache=True, kv_cache=cache, start_pos=0)
        slots = s.slots(0, P)
        for li in range(len(m.all_layer)):
            mgr.k_cache[li].view(-1, mgr.kv_heads, mgr.head_dim)[slots] = cache.key_cache[li][0, :, :P].transpose(0, 1)
            mgr.v_cache[li].view(-1, mgr.kv_heads, mgr.head_dim)[slots] = cache.value_cache[li][0, :, :P].transpose(0, 1)
        return int(ops.argmax_rows(out.logits[0, -1:].contiguous())[0])

    # ---- all decoding sequences: one batched single-token step over the paged pools ----
    @torch.no_grad()
    def _decode(self, states: List[SequenceState]) -> List[int]:
        m, mgr = self.model, self.kv_mgr
        B = len(states)
        T = m.word_embeddings.weight.dtype
        tok = torch.stack([s.tokens[s.num_tokens - 1] for s in states])
        seqlens = torch.tensor([s.num_tokens - 1 for s in states], dtype=torch.int32, device=self.device)
        width = max(s.block_table.numel() for s in states)
        table = torch.zeros((B, width), dtype=torch.int32, device=self.device)
        for i, s in enumerate(states):
            table[i, : s.block_table.numel()] = s.block_table
        longest = max(s.num_tokens - 1 for s in states)
        x = m._embed(tok.view(B, 1), 0)  # [B, H]
        cos, sin = m._rope.get(x.device, T)
        for li, layer in enumerate(m.all_layer):
            att = layer.attention
            w_qkv, b_qkv = F.pack_linears(att._packed())
            qkv = F._lin(x, w_qkv, b_qkv)
            a = ops.attn_decode(qkv, mgr.k_cache[li], mgr.v_cache[li], longest, att.num_attention_heads, att._kv_heads, cos, sin,
                                out_dtype=T, seqlens=seqlens, block_table=table)
            y, _ = F.self_output(a, x, att.out.dense, att.out.layernorm)
            ff = layer.feed_forward
            x, _ = F.feed_forward(y, x, ff.intermediate, ff.out, ff.layernorm, act=ff._act_name)
        head = m.lm_head
        logits, _ = F.lm_head(x, head.dense, head.layer_norm, head.decoder.weight, head.bias)
        return ops.argmax_rows(logits).tolist()

    def step(self) -> Dict[int, List[int]]:
        """One engine iteration: schedule, grow block tables, prefill the new sequences, decode the others, retire the
        finished ones. Returns {sequence id: all its tokens} for the sequences that finished in this step."""
        self._try_schedule_waiting()
        if not self.active:
            return {}
        states = list(self.active.values())
        for s in states:
            self.kv_mgr.allocate(s)
        nxt: Dict[int, int] = {}
        decoding = [s for s in states if not s.is_prefill]
        for s in states:
            if s.is_prefill:
                nxt[s.id] = self._prefill(s)
        if decoding:
            for s, t in zip(decoding, self._decode(decoding)):
                nxt[s.id] = t
        finished: Dict[int, List[int]] = {}
        for s in states:
            token_id = nxt[s.id]
            s.is_prefill = False
            s.tokens[s.num_tokens] = token_id
            s.num_tokens += 1
            if token_id in self.eos or s.num_tokens >= s.max_total_len:
                finished[s.id] = s.tokens[: s.num_tokens].tolist()
                self.kv_mgr.free(s)
                del self.active[s.id]
        return finished
