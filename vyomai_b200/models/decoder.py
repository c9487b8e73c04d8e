"""GPT-style decoder — host-side mirror of VyomAI/models/decoder.py: DecoderModel with the shared
whole-model kv-cache (`StaticCacheOne` / `DynamicCacheOne`), `create_mask_for_decoder`, greedy /
sampled `generate`. Same names, signatures, outputs and state_dict keys as the reference."""
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn

from .. import ops
from ..functional import MaskSpec
from ..layers.attention import SharedCacheDecoderAttention as DecoderAttention
from ..layers.attention import SharedCacheDecoderAttentionGqa as DecoderAttentionGqa
from ..layers.ffn import FeedForward
from ..layers.kv_cache import DynamicCacheOne, StaticCacheOne
from ._common import LMHead, TextStem, back_to, ensure_cuda


@dataclass
class DecoderOutput(object):
    logits: torch.Tensor


@dataclass
class CLMOutput(object):
    hidden_state: torch.Tensor
    logits: torch.Tensor
    kv_cache: List[torch.FloatTensor] = None


class DecoderLayer(nn.Module):
    "decoder layer for decoder model (reference: models/decoder.py:204-250)"

    def __init__(self, config, layer_idx: int, attention_type: Optional[str] = None) -> None:
        super().__init__()
        self.attention = (
            DecoderAttentionGqa(config, layer_idx=layer_idx) if attention_type == "gqa"
            else DecoderAttention(config, layer_idx=layer_idx)
        )
        if attention_type == "gqa" and layer_idx == 0:
            print("Decoder Using GQA Attention")
        self.feed_forward = FeedForward(config)
        self.layer_idx = layer_idx

    def forward(self, hidden_state, attention_mask, freqs=None, use_cache: Optional[bool] = False, kv_cache=None,
                start_pos: Optional[int] = 0):
        out, kv_cache = self.attention(hidden_state=hidden_state, attention_mask=attention_mask, freqs=freqs,
                                       use_cache=use_cache, kv_cache=kv_cache, start_pos=start_pos)
        out = self.feed_forward(out, hidden_state)
        return out, kv_cache


class DecoderModel(nn.Module, TextStem):
    "Decoder model for language modeling (reference: models/decoder.py:278-514)"

    def __init__(self, config, pos_embedding_type: Optional[str] = "absolute", attention_type: Optional[str] = None) -> None:
        super().__init__()
        self.is_gqa = True if attention_type == "gqa" else False
        self._build_stem(config, pos_embedding_type, "Decoder")
        self.all_layer = nn.ModuleList(
            [DecoderLayer(config, layer_idx, attention_type) for layer_idx in range(config.num_hidden_layers)]
        )
        self.lm_head = LMHead(config=config)
        self.config = config

    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                use_cache: Optional[bool] = False, kv_cache=None, start_pos: Optional[int] = 0,
                _logits_last_only: bool = False) -> CLMOutput:
        dev, origin, (input_ids, attention_mask) = ensure_cuda(self, input_ids, attention_mask)
        _bsz, seqlen = input_ids.shape
        self._check_positions(start_pos + seqlen)
        hidden = self._embed(input_ids, start_pos)
        mask = None
        if seqlen > 1:  # decoder.py:355-362 — single-token steps carry NO mask (quirk Q3)
            mask = self.create_mask_for_decoder(input_ids=input_ids, attention_mask=attention_mask, start_pos=start_pos,
                                                _factored=True)
        hidden = hidden.view(_bsz, seqlen, -1)
        for layer in self.all_layer:
            hidden, kv_cache = layer(hidden, mask, freqs=self._rope, use_cache=use_cache, kv_cache=kv_cache,
                                     start_pos=start_pos)
        head_in = hidden[:, -1:, :].contiguous() if _logits_last_only else hidden
        logits = self.lm_head(head_in)
        return CLMOutput(hidden_state=back_to(origin, hidden), logits=back_to(origin, logits), kv_cache=kv_cache)

    def create_mask_for_decoder(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                                start_pos: Optional[int] = 0, _factored: bool = False):
        """causal (with a start_pos prefix of ones) x key padding (reference: decoder.py:376-419).
        Returns the dense (B,1,S,start_pos+S) 0/1 tensor like the reference; the model itself asks for
        the factored MaskSpec the kernels consume."""
        batch_size, seq_length = input_ids.shape
        if _factored:
            if attention_mask is not None and attention_mask.shape[1] != start_pos + seq_length:
                raise ValueError(
                    f"attention_mask has {attention_mask.shape[1]} columns, expected start_pos + seqlen = {start_pos + seq_length}"
                )
            return MaskSpec.from_attention_mask(attention_mask, causal=True, q_pos0=start_pos)
        device = input_ids.device
        if attention_mask is None:
            attention_mask = torch.ones(seq_length + start_pos).repeat(batch_size, 1).to(device)
        seq_ids = torch.arange(seq_length).to(device)
        causal_mask = (seq_ids[None, None, :].repeat(batch_size, seq_length, 1) <= seq_ids[None, :, None]).to(attention_mask.dtype)
        if start_pos > 0:
            causal_mask = torch.cat(
                [torch.ones((batch_size, seq_length, start_pos), device=device, dtype=causal_mask.dtype), causal_mask], axis=-1)
        return causal_mask[:, None, :, :] * attention_mask[:, None, None, :]

    @classmethod
    def from_config(cls, config, pos_embedding_type: Optional[str] = "absolute", attention_type: Optional[str] = None) -> nn.Module:
        return cls(config, pos_embedding_type, attention_type)

    use_decode_graph = True  # class-level switch: False forces the per-token Python loop of the reference

    def _generate_graph(self, tokens, attention_mask, kv_cache, prompt_len: int, max_len: int, pad_id: int) -> torch.Tensor:
        """Greedy continuation of full-length prompts with a static cache: eager prefill, then one CUDA-graph replay
        per token (decode_graph.GreedyDecodeGraph). Token ids, cache contents and the early stop are those of the
        reference loop (decoder.py:470-513): the reference stops after the first step at which every sequence has
        produced an eos token and leaves the remaining positions at pad_id, which is applied here after the fact."""
        from ..decode_graph import GreedyDecodeGraph
        bsz = tokens.shape[0]
        dev = tokens.device
        out = self.forward(input_ids=tokens[:, :prompt_len].contiguous(), attention_mask=attention_mask, use_cache=True,
                           kv_cache=kv_cache, start_pos=0, _logits_last_only=True)
        first = ops.argmax_rows(out.logits[:, -1])
        tokens[:, prompt_len] = first
        steps = max_len - prompt_len - 1
        if steps > 0:
            g = getattr(self, "_decode_graph", None)
            if g is None or g.cache is not kv_cache:
                g = self._decode_graph = GreedyDecodeGraph(self, kv_cache, bsz, max_len)
            g.run(first, prompt_len, steps)
            tokens[:, prompt_len + 1:] = g.tokens[:, prompt_len + 1:]
        stop_tokens = torch.as_tensor(getattr(self.config, "eos_token_id", 2), device=dev)
        gen = tokens[:, prompt_len:]
        hit = torch.isin(gen, stop_tokens).to(torch.int32).cummax(dim=1).values.bool()  # eos seen at or before step t
        all_hit = hit.all(dim=0)                                                          # the reference breaks after step t
        after = torch.cat([torch.zeros(1, dtype=torch.bool, device=dev), all_hit[:-1]]).to(torch.int32).cummax(dim=0).values.bool()
        gen[:, after] = pad_id
        return tokens

    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, max_len: int = 5, temperature: float = 1.0,
                 use_cache: bool = True, do_sample: bool = False, use_static_cache: bool = False) -> torch.Tensor:
        """Greedy / sampled continuation (reference: decoder.py:430-514). The loop structure, prompt
        handling and eos bookkeeping are the reference's; each step is one fused forward that computes
        the LM head for the last position only, and greedy selection is the vy_argmax_rows kernel
        (first index on ties, like torch.topk(k=1))."""
        dev, origin, (input_ids, attention_mask) = ensure_cuda(self, input_ids, attention_mask)
        all_prompt_size = [t.size()[0] for t in input_ids]
        min_prompt_len, max_prompt_len = min(all_prompt_size), max(all_prompt_size)
        max_len = max_len + max_prompt_len
        pad_id = getattr(self.config, "pad_token_id", 1)
        bsz, _ = input_ids.size()
        tokens = torch.full((bsz, max_len), pad_id, dtype=torch.long, device=dev)
        kv_cache = None
        live_dropout = self.training and getattr(self.config, "hidden_dropout_prob", 0.0) > 0.0  # random masks: eager loop
        graph_path = (self.use_decode_graph and use_cache and use_static_cache and not do_sample and self._rope is not None
                      and not live_dropout and min_prompt_len == input_ids.shape[1] and bool((input_ids != pad_id).all()))
        if graph_path:
            # the replayed steps index the RoPE table with positions up to max_len - 2 without passing through forward():
            # apply its position check up front, like the eager loop (and the reference) would hit it on the way
            self._check_positions(max_len - 1)
        if use_cache:
            if use_static_cache and graph_path:
                # the captured step is bound to its cache buffers: keep both and reuse them for calls of the same shape
                # (slots beyond the current position are never read, so a stale cache needs no clearing)
                # the captured step bakes in the weights' addresses: re-capture when parameters were re-homed
                # (trainer.FlatParams, pack_linears after .to(), load_state_dict(assign=True))
                key = (bsz, max_len, self.word_embeddings.weight.dtype, dev, tuple(p.data_ptr() for p in self.parameters()))
                if getattr(self, "_decode_graph_key", None) != key:
                    self._decode_graph_key = key
                    self._decode_graph_cache = StaticCacheOne(self.config, max_cache_len=max_len, batch_size=bsz,
                                                              dtype=self.word_embeddings.weight.dtype)
                    self._decode_graph = None
                kv_cache = self._decode_graph_cache
            elif use_static_cache:
                kv_cache = StaticCacheOne(self.config, max_cache_len=max_len, batch_size=bsz,
                                          dtype=self.word_embeddings.weight.dtype)
            else:
                kv_cache = DynamicCacheOne(self.config)
        tokens[:, : input_ids.shape[1]] = input_ids
        if graph_path:
            return back_to(origin, self._generate_graph(tokens, attention_mask, kv_cache, min_prompt_len, max_len, pad_id))
        prev_pos = 0
        eos_reached = torch.zeros(bsz, dtype=torch.bool, device=dev)
        input_text_mask = tokens != pad_id
        stop_tokens = torch.as_tensor(getattr(self.config, "eos_token_id", 2), device=dev)
        for cur_pos in range(min_prompt_len, max_len):
            outputs = self.forward(input_ids=tokens[:, prev_pos:cur_pos].contiguous(), attention_mask=attention_mask,
                                   use_cache=use_cache, kv_cache=kv_cache, start_pos=prev_pos, _logits_last_only=True)
            kv_cache = outputs.kv_cache
            last = outputs.logits[:, -1]
            if do_sample:
                next_token = torch.multinomial(last.float() / temperature, num_samples=1).reshape(-1)
            else:
                next_token = ops.argmax_rows(last)  # argmax is invariant to the positive temperature scale
            next_token = torch.where(input_text_mask[:, cur_pos], tokens[:, cur_pos], next_token)
            tokens[:, cur_pos] = next_token
            eos_reached |= (~input_text_mask[:, cur_pos]) & torch.isin(next_token, stop_tokens)
            if use_cache:
                prev_pos = cur_pos
            attention_mask = torch.cat([attention_mask, torch.ones((bsz, 1), device=dev, dtype=attention_mask.dtype)], dim=-1)
            if bool(eos_reached.all()):
                break
        return back_to(origin, tokens)
