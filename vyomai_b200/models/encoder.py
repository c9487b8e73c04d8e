"""BERT-style encoder — host-side mirror of VyomAI/models/encoder.py (same classes, constructor
arguments, attribute / state_dict names, forward signatures and output dataclasses)."""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from ..functional import MaskSpec
from ..layers.attention import EncoderAttention, EncoderAttentionGqa
from ..layers.ffn import FeedForward
from ._common import LMHead, TextStem, back_to, ensure_cuda


@dataclass
class EncoderOutput(object):
    logits: torch.Tensor


@dataclass
class MLMOutput(object):
    hidden_state: torch.Tensor
    logits: torch.Tensor


class EncoderLayer(nn.Module):
    "encoder layer for encoder model (reference: models/encoder.py:30-64)"

    def __init__(self, config, layer_idx: int, attention_type: str = None) -> None:
        super().__init__()
        self.attention = (
            EncoderAttentionGqa(config, layer_idx=layer_idx) if attention_type == "gqa"
            else EncoderAttention(config, layer_idx=layer_idx)
        )
        if attention_type == "gqa" and layer_idx == 0:
            print("Encoder Using GQA Attention")
        self.feed_forward = FeedForward(config)
        self.layer_idx = layer_idx

    def forward(self, hidden_state: torch.Tensor, attention_mask: torch.Tensor, freqs: torch.Tensor = None) -> torch.Tensor:
        out = self.attention(hidden_state=hidden_state, attention_mask=attention_mask, freqs=freqs)
        out = self.feed_forward(out, hidden_state)  # residual = LAYER INPUT (quirk Q2)
        return out


class EncoderModel(nn.Module, TextStem):
    """reference: models/encoder.py:92-178"""

    def __init__(self, config, pos_embedding_type: Optional[str] = "absolute", attention_type: str = None) -> None:
        super().__init__()
        self._build_stem(config, pos_embedding_type, "Encoder")
        self.all_layer = nn.ModuleList(
            [EncoderLayer(config, layer_idx, attention_type) for layer_idx in range(config.num_hidden_layers)]
        )

    def forward(self, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> EncoderOutput:
        dev, origin, (input_ids, attention_mask) = ensure_cuda(self, input_ids, attention_mask)
        bsz, seqlen = input_ids.shape
        self._check_positions(seqlen)
        hidden = self._embed(input_ids, 0)
        # (1 - mask) * finfo.min of encoder.py:161-164, in factored form; None = all ones
        mask = MaskSpec.from_attention_mask(attention_mask, causal=False)
        hidden = hidden.view(bsz, seqlen, -1)
        for layer in self.all_layer:
            hidden = layer(hidden, mask, self._rope)
        return EncoderOutput(back_to(origin, hidden))

    @classmethod
    def from_config(cls, config, pos_embedding_type: Optional[str] = "absolute", attention_type: str = None) -> nn.Module:
        return cls(config, pos_embedding_type, attention_type)


class EncoderForMaskedLM(nn.Module):
    """reference: models/encoder.py:180-217"""

    def __init__(self, config, pos_embedding_type: Optional[str] = "absolute", attention_type: str = None) -> None:
        super().__init__()
        self.encoder = EncoderModel(config, pos_embedding_type=pos_embedding_type, attention_type=attention_type)
        self.lm_head = LMHead(config=config)

    def forward(self, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> MLMOutput:
        dev, origin, (input_ids, attention_mask) = ensure_cuda(self, input_ids, attention_mask)
        out = self.encoder(input_ids=input_ids, attention_mask=attention_mask)
        logits = self.lm_head(out.logits)
        return MLMOutput(hidden_state=back_to(origin, out.logits), logits=back_to(origin, logits))

    @classmethod
    def from_config(cls, config, pos_embedding_type: Optional[str] = "absolute", attention_type: str = None) -> nn.Module:
        return cls(config, pos_embedding_type, attention_type)
