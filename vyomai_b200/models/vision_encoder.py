"""ViT — host-side mirror of VyomAI/models/vision_encoder.py (same names, state_dict keys, outputs).

The stem is fused: vy_patchify turns NCHW pixels into patch rows, ONE vy_gemm against
`pixel_seq.weight` viewed as [hidden, C*p*p] applies the stride==kernel convolution, its epilogue
adds the conv bias and the position row, multiplies by 2 and scatters each image's patches behind
its cls row; the cls rows are written by vy_embed_fwd. The factor 2 reproduces the reference's
double position add (positional_embeddings.py:222-226 + vision_encoder.py:125-127; quirk Q1).
"""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib, ops
from ..functional import MaskSpec
from ..layers.attention import VisionAttention
from ..layers.ffn import FeedForward
from ..layers.positional_embeddings import VitAbsoluteEncoding
from ._common import back_to, ensure_cuda

_position_embeddings = {"absolute": VitAbsoluteEncoding}


@dataclass
class EncoderOutput(object):
    logits: torch.Tensor


class EncoderLayer(nn.Module):
    "encoder layer for the vision encoder (reference: models/vision_encoder.py:22-53)"

    def __init__(self, config, layer_idx: int, attention_type: Optional[str] = None) -> None:
        super().__init__()
        self.attention = VisionAttention(config, layer_idx=layer_idx)
        self.feed_forward = FeedForward(config)
        self.layer_idx = layer_idx

    def forward(self, hidden_state, attention_mask, freqs=None) -> torch.Tensor:
        out = self.attention(hidden_state=hidden_state, attention_mask=attention_mask, freqs=freqs)
        out = self.feed_forward(out, hidden_state)
        return out


class Vit(nn.Module):
    """reference: models/vision_encoder.py:56-153"""

    def __init__(self, config, pos_embedding_type: Optional[str] = "absolute") -> None:
        super().__init__()
        self.image_size = config.image_size
        self.patch_size = config.patch_size
        self.num_channels = config.num_channels
        self.num_patches = (self.image_size[0] // self.patch_size[0]) * (self.image_size[1] // self.patch_size[1])
        if _position_embeddings.get(pos_embedding_type, None) is not None:
            self.position_embeddings = _position_embeddings.get(pos_embedding_type)(config)
        else:
            self.position_embeddings = None
        self.all_layer = nn.ModuleList([EncoderLayer(config, layer_idx) for layer_idx in range(config.num_hidden_layers)])
        self.pixel_seq = nn.Conv2d(in_channels=self.num_channels, out_channels=config.hidden_size,
                                   kernel_size=self.patch_size, stride=self.patch_size)
        patch_dim = config.num_channels * self.patch_size[0] * self.patch_size[1]
        self.cls_token = nn.Parameter(torch.randn(1, 1, patch_dim))
        self._hidden = config.hidden_size

    def forward(self, pixel_values: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> EncoderOutput:
        dev, origin, (pixel_values, attention_mask) = ensure_cuda(self, pixel_values, attention_mask)
        if self.position_embeddings is None:
            raise _lib.VyomError("Vit without position embeddings reads an undefined emb_freq in the reference too "
                                 "(vision_encoder.py:129); only pos_embedding_type='absolute' exists")
        from ..autograd import vit_stem_fn
        bsz = pixel_values.shape[0]
        hidden = vit_stem_fn(self, pixel_values)  # (bsz * (nP + 1), hidden)
        seqlen = self.num_patches + 1
        mask = MaskSpec.from_attention_mask(attention_mask, causal=False)  # None = all ones
        hidden = hidden.view(bsz, seqlen, -1)
        for layer in self.all_layer:
            hidden = layer(hidden, mask, None)
        return EncoderOutput(back_to(origin, hidden))

    @classmethod
    def from_config(cls, config, pos_embedding_type: Optional[str] = "absolute") -> nn.Module:
        return cls(config, pos_embedding_type)
