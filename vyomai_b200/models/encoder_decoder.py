"""BART-style seq2seq — host-side mirror of VyomAI/models/encoder_decoder.py, kept working on the same kernels as the rest
of the path (not a benchmarked model, but its cross-attention is §8(a) and the reference's
tests/test_encoder_decoder.py must keep passing): `Seq2SeqDecoderLayer` = causal self-attention (per-layer cache) ->
cross-attention over the encoder states (k / v cached once) -> FeedForward whose residual is the layer input (quirk Q2),
`Seq2SeqDecoderModel`, `EncoderDecoderModel` with its own `LMHead` (state_dict keys `lm_head.{dense,layer_norm,vocab}.*`,
`lm_head.bias`), `_setup_cache` / `_clean_cache` attaching caches to BOTH attention modules of every layer."""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from .. import functional as F
from ..functional import MaskSpec
from ..layers.attention import (DecoderAttention, DecoderAttentionGqa, EncoderDecoderAttention, EncoderDecoderAttentionGqa)
from ..layers.ffn import FeedForward
from ..layers.kv_cache import DynamicCache, StaticCache  # noqa: F401
from ._common import TextStem, back_to, ensure_cuda
from .encoder import EncoderModel


@dataclass
class Seq2SeqOutput(object):
    logits: torch.Tensor
    key_value_states: torch.Tensor


class Seq2SeqDecoderLayer(nn.Module):
    "decoder layer for Seq2Seq model (reference: models/encoder_decoder.py:33-87)"

    def __init__(self, config, layer_idx: Optional[int] = 0, attention_type: Optional[str] = None) -> None:
        super().__init__()
        gqa = attention_type == "gqa"
        self.attention = DecoderAttentionGqa(config, layer_idx=layer_idx) if gqa else DecoderAttention(config, layer_idx=layer_idx)
        if gqa and layer_idx == 0:
            print("Decoder Using GQA Attention")
        self.cross_attention = (EncoderDecoderAttentionGqa(config, layer_idx=layer_idx) if gqa
                                else EncoderDecoderAttention(config, layer_idx=layer_idx))
        if gqa and layer_idx == 0:
            print("Using GQA in Cross Attention")
        self.feed_forward = FeedForward(config)
        self.layer_idx = layer_idx

    def forward(self, hidden_state, attention_mask, encoder_hidden_state=None, encoder_attention_mask=None, freqs=None,
                use_cache: Optional[bool] = False, start_pos: Optional[int] = 0) -> torch.Tensor:
        out = self.attention(hidden_state=hidden_state, attention_mask=attention_mask, freqs=freqs, use_cache=use_cache,
                             start_pos=start_pos)
        out = self.cross_attention(hidden_state=out, encoder_hidden_state=encoder_hidden_state,
                                   encoder_attention_mask=encoder_attention_mask, freqs=freqs, use_cache=use_cache)
        return self.feed_forward(out, hidden_state)


class LMHead(nn.Module):
    """dense -> GELU -> LayerNorm -> vocab (reference: models/encoder_decoder.py:90-113; the output projection is called
    `vocab` here, and `bias` / `vocab.bias` are one Parameter)."""

    def __init__(self, config) -> None:
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.layer_norm = nn.LayerNorm(config.hidden_size, eps=getattr(config, "layer_norm_eps", 1e-6))
        self.vocab = nn.Linear(config.hidden_size, config.vocab_size)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))
        self.vocab.bias = self.bias

    @property
    def decoder(self) -> nn.Linear:  # the name the shared LM-head kernels path looks the projection up under
        return self.vocab

    def forward(self, hidden_state: torch.Tensor) -> torch.Tensor:
        from ..autograd import lm_head_fn
        shape = hidden_state.shape
        logits = lm_head_fn(self, hidden_state.reshape(-1, shape[-1]))
        return logits.view(*shape[:-1], logits.shape[-1])

    def loss(self, hidden_state: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
        """Token cross-entropy of self(hidden) against `labels` (one per hidden row) with the head and the loss fused."""
        from ..autograd import lm_head_loss_fn
        return lm_head_loss_fn(self, hidden_state.reshape(-1, hidden_state.shape[-1]), labels, ignore_index)


class Seq2SeqDecoderModel(nn.Module, TextStem):
    """Seq2Seq decoder model (reference: models/encoder_decoder.py:116-278)"""

    def __init__(self, config, pos_embedding_type: Optional[str] = "absolute", attention_type: Optional[str] = None) -> None:
        super().__init__()
        self._build_stem(config, pos_embedding_type, "Decoder")
        self.all_layer = nn.ModuleList(
            [Seq2SeqDecoderLayer(config, layer_idx, attention_type=attention_type) for layer_idx in range(config.num_hidden_layers)]
        )

    def forward(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, encoder_hidden_state: Optional[torch.Tensor] = None,
                encoder_attention_mask=None, use_cache: Optional[bool] = False, start_pos: Optional[int] = 0) -> torch.Tensor:
        dev, _origin, (input_ids, attention_mask, encoder_hidden_state) = ensure_cuda(self, input_ids, attention_mask,
                                                                                     encoder_hidden_state)
        _bsz, seqlen = input_ids.shape
        self._check_positions(start_pos + seqlen)
        hidden = self._embed(input_ids, start_pos).view(_bsz, seqlen, -1)
        mask = None
        if seqlen > 1:  # encoder_decoder.py:197-204 — single-token steps carry no self-attention mask (quirk Q3)
            if attention_mask is not None and attention_mask.shape[1] != start_pos + seqlen:
                raise ValueError(f"attention_mask has {attention_mask.shape[1]} columns, expected start_pos + seqlen = {start_pos + seqlen}")
            mask = MaskSpec.from_attention_mask(attention_mask, causal=True, q_pos0=start_pos)
        if encoder_attention_mask is not None and not isinstance(encoder_attention_mask, MaskSpec):
            encoder_attention_mask = MaskSpec.from_dense(encoder_attention_mask.to(dev), 1)
        for layer in self.all_layer:
            hidden = layer(hidden, mask, encoder_hidden_state=encoder_hidden_state, encoder_attention_mask=encoder_attention_mask,
                           freqs=self._rope, use_cache=use_cache, start_pos=start_pos)
        return hidden

    def create_mask_for_decoder(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                                start_pos: Optional[int] = 0) -> torch.Tensor:
        """causal (with a start_pos prefix of ones) x key padding, dense (B,1,S,start_pos+S) like the reference (:214-262)."""
        device = input_ids.device
        batch_size, seq_length = input_ids.shape
        if attention_mask is None:
            attention_mask = torch.ones(seq_length + start_pos).repeat(batch_size, 1).to(device)
        seq_ids = torch.arange(seq_length).to(device)
        causal_mask = (seq_ids[None, None, :].repeat(batch_size, seq_length, 1) <= seq_ids[None, :, None]).to(attention_mask.dtype)
        if start_pos > 0:
            causal_mask = torch.cat(
                [torch.ones((batch_size, seq_length, start_pos), device=device, dtype=causal_mask.dtype), causal_mask], axis=-1)
        return causal_mask[:, None, :, :] * attention_mask[:, None, None, :]

    @classmethod
    def from_config(cls, config) -> nn.Module:
        return cls(config)


class EncoderDecoderModel(nn.Module):
    """reference: models/encoder_decoder.py:281-391"""

    def __init__(self, encoder_config, decoder_config, encoder: Optional[nn.Module] = None,
                 encoder_pos_embedding_type: Optional[str] = "absolute", encoder_attention_type: Optional[str] = None,
                 decoder_pos_embedding_type: Optional[str] = "absolute", decoder_attention_type: Optional[str] = None) -> None:
        super().__init__()
        self.is_gqa = True if decoder_attention_type == "gqa" else False
        self.encoder = (encoder if encoder is not None
                        else EncoderModel(config=encoder_config, pos_embedding_type=encoder_pos_embedding_type,
                                          attention_type=encoder_attention_type))
        self.decoder = Seq2SeqDecoderModel(config=decoder_config, pos_embedding_type=decoder_pos_embedding_type,
                                           attention_type=decoder_attention_type)
        self.lm_head = LMHead(config=decoder_config)

    def forward(self, input_ids=None, attention_mask=None, decoder_input_ids=None, decoder_attention_mask=None,
                encoder_output=None, use_cache: Optional[bool] = False, start_pos: Optional[int] = 0) -> Seq2SeqOutput:
        dev, origin, (input_ids, attention_mask, decoder_input_ids, decoder_attention_mask, encoder_output) = ensure_cuda(
            self, input_ids, attention_mask, decoder_input_ids, decoder_attention_mask, encoder_output)
        if encoder_output is None:
            encoder_output = self.encoder(input_ids=input_ids, attention_mask=attention_mask).logits
        enc_mask = MaskSpec.from_attention_mask(attention_mask, causal=False)  # None = every encoder position visible
        hidden = self.decoder(input_ids=decoder_input_ids, attention_mask=decoder_attention_mask,
                              encoder_hidden_state=encoder_output, encoder_attention_mask=enc_mask, use_cache=use_cache,
                              start_pos=start_pos)
        logits = self.lm_head(hidden)
        return Seq2SeqOutput(key_value_states=back_to(origin, encoder_output), logits=back_to(origin, logits))

    def get_encoder(self) -> nn.Module:
        return self.encoder

    def get_encoder_output(self, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> object:
        return self.encoder(input_ids=input_ids, attention_mask=attention_mask)

    def get_decoder(self) -> Seq2SeqDecoderModel:
        return self.decoder

    def _setup_cache(self, config, cls: Optional[object] = StaticCache) -> None:
        """setup kv-cache hooks for every self-attention and cross-attention layer"""
        for layer in self.decoder.all_layer:
            layer.attention.cache = cls(config, is_gqa=self.is_gqa)
            layer.cross_attention.cache = cls(config, is_gqa=self.is_gqa)

    def _clean_cache(self) -> None:
        for layer in self.decoder.all_layer:
            layer.attention.cache = None
            layer.cross_attention.cache = None

    @classmethod
    def from_config(cls, encoder_config, decoder_config, encoder: Optional[nn.Module] = None,
                    encoder_pos_embedding_type: Optional[str] = "absolute", encoder_attention_type: Optional[str] = None,
                    decoder_pos_embedding_type: Optional[str] = "absolute", decoder_attention_type: Optional[str] = None) -> nn.Module:
        return cls(encoder_config, decoder_config, encoder, encoder_pos_embedding_type, encoder_attention_type,
                   decoder_pos_embedding_type, decoder_attention_type)
