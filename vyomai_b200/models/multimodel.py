"""Image-text fusion captioner — host-side mirror of VyomAI/models/multimodel.py: a ViT whose CLS
vector is prepended as token 0 of a GPT-style decoder with per-layer kv-caches. Same classes,
signatures, state_dict keys (`encoder.*`, `decoder.*`) and cache hooks as the reference."""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from ..functional import MaskSpec
from ..layers.attention import DecoderAttention, DecoderAttentionGqa
from ..layers.ffn import FeedForward
from ..layers.kv_cache import DynamicCache, StaticCache
from ._common import LMHead, TextStem, back_to, ensure_cuda


@dataclass
class DecoderOutput(object):
    logits: torch.Tensor


class DecoderLayer(nn.Module):
    "decoder layer with a per-layer cache hook (reference: models/multimodel.py:27-69)"

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

    def forward(self, hidden_state, attention_mask, freqs=None, use_cache: Optional[bool] = False,
                start_pos: Optional[int] = 0) -> torch.Tensor:
        out = self.attention(hidden_state=hidden_state, attention_mask=attention_mask, freqs=freqs, use_cache=use_cache,
                             start_pos=start_pos)
        out = self.feed_forward(out, hidden_state)
        return out


class VisionLanguageDecoderModel(nn.Module, TextStem):
    "VisionLanguageModel decoder (reference: models/multimodel.py:97-255)"

    def __init__(self, config, pos_embedding_type: Optional[str] = "absolute", attention_type: Optional[str] = None) -> None:
        super().__init__()
        self.is_gqa = True if attention_type == "gqa" else False
        self._build_stem(config, pos_embedding_type, "Decoder")
        self.all_layer = nn.ModuleList(
            [DecoderLayer(config, layer_idx, attention_type) for layer_idx in range(config.num_hidden_layers)]
        )
        self.lm_head = LMHead(config=config)

    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                encoder_hidden_state: Optional[torch.Tensor] = None, use_cache: Optional[bool] = False,
                start_pos: Optional[int] = 0) -> DecoderOutput:
        origin, hidden = self._hidden(input_ids, attention_mask, encoder_hidden_state, use_cache, start_pos)
        logits = self.lm_head(hidden)
        return DecoderOutput(logits=back_to(origin, logits))

    def forward_loss(self, input_ids, attention_mask, encoder_hidden_state, labels_full, ignore_index: int = -100):
        """Training entry point: the token cross-entropy of forward(...).logits against `labels_full` ([B, S+1],
        aligned with the logits rows) — same math as `loss_fn` of the reference's captioner notebook, with the LM
        head and the loss fused into one autograd node."""
        _origin, hidden = self._hidden(input_ids, attention_mask, encoder_hidden_state, False, 0)
        return self.lm_head.loss(hidden, labels_full, ignore_index)

    def _hidden(self, input_ids, attention_mask, encoder_hidden_state, use_cache, start_pos):
        dev, origin, (input_ids, attention_mask, encoder_hidden_state) = ensure_cuda(
            self, input_ids, attention_mask, encoder_hidden_state)
        _bsz, ntok = input_ids.shape
        table = self.word_embeddings.weight
        extra = 1 if start_pos == 0 else 0  # the image vector is token 0 of the first call (multimodel.py:163-169)
        seqlen = ntok + extra
        self._check_positions(start_pos + seqlen)
        enc = encoder_hidden_state.to(table.dtype).contiguous() if extra else None
        hidden = self._embed(input_ids, start_pos, extra=enc)
        if extra:
            if attention_mask is not None:
                attention_mask = torch.cat([torch.ones((_bsz, 1), device=dev, dtype=attention_mask.dtype), attention_mask], dim=1)
        mask = None
        if seqlen > 1:
            mask = MaskSpec.from_attention_mask(attention_mask, causal=True, q_pos0=start_pos)
        hidden = hidden.view(_bsz, seqlen, -1)
        for layer in self.all_layer:
            hidden = layer(hidden, mask, freqs=self._rope, use_cache=use_cache, start_pos=start_pos)
        return origin, hidden

    @classmethod
    def from_config(cls, config, pos_embedding_type: Optional[str] = "absolute", attention_type: Optional[str] = None) -> nn.Module:
        return cls(config, pos_embedding_type, attention_type)


class VisionLanguageModel(nn.Module):
    """reference: models/multimodel.py:258-314"""

    def __init__(self, config, encoder, pos_embedding_type: Optional[str] = "absolute", attention_type: Optional[str] = None) -> None:
        super().__init__()
        self.is_gqa = True if attention_type == "gqa" else False
        self.encoder = encoder
        self.decoder = VisionLanguageDecoderModel(config=config, pos_embedding_type=pos_embedding_type,
                                                  attention_type=attention_type)

    def forward(self, pixel_values=None, decoder_input_ids=None, decoder_attention_mask=None, encoder_output=None,
                use_cache: Optional[bool] = False, start_pos: Optional[int] = 0) -> DecoderOutput:
        if encoder_output is None:
            encoder_output = self.encoder(pixel_values=pixel_values).logits[:, 0, :]  # cls token information
        return self.decoder(input_ids=decoder_input_ids, attention_mask=decoder_attention_mask,
                            encoder_hidden_state=encoder_output, use_cache=use_cache, start_pos=start_pos)

    def forward_loss(self, pixel_values, decoder_input_ids, decoder_attention_mask, labels_full, ignore_index: int = -100):
        """Captioning loss of one batch (see VisionLanguageDecoderModel.forward_loss)."""
        encoder_output = self.encoder(pixel_values=pixel_values).logits[:, 0, :]
        return self.decoder.forward_loss(decoder_input_ids, decoder_attention_mask, encoder_output, labels_full, ignore_index)

    def get_decoder(self) -> nn.Module:
        return self.decoder

    def get_encoder_output(self, pixel_values: torch.Tensor) -> torch.Tensor:
        return self.encoder(pixel_values=pixel_values).logits[:, 0, :]

    def _setup_cache(self, config, cls: Optional[object] = StaticCache) -> None:
        """setup kv-cache hooks for every self-attention layer"""
        for layer in self.decoder.all_layer:
            layer.attention.cache = cls(config, is_gqa=self.is_gqa)

    def _clean_cache(self) -> None:
        """destroy kv-cache hooks for every self-attention layer"""
        for layer in self.decoder.all_layer:
            layer.attention.cache = None
