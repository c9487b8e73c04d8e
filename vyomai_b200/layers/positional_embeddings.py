"""Positional embeddings — host-side mirror of VyomAI/layers/positional_embeddings.py.

Same class names, constructor arguments, parameter / buffer names and return shapes as the
reference. The tables are data (parameters / precomputed tensors); the arithmetic that consumes
them (the add into the hidden state, the rotation of q/k) happens inside the fused sm_100a
kernels (vy_embed_fwd, the QKV-RoPE GEMM epilogue, vy_attn_decode).
"""
from typing import Tuple

import torch
import torch.nn as nn

from .. import ops


class AbsoluteEncoding(nn.Module):
    """Learned absolute positions (reference: positional_embeddings.py:7-51)."""

    def __init__(self, config) -> None:
        super().__init__()
        self.pos_embeddings = nn.Embedding(
            config.max_position_embeddings, config.hidden_size, padding_idx=getattr(config, "pad_token_id", None)
        )
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)), persistent=False)
        self.max_size = config.max_position_embeddings

    def table(self) -> torch.Tensor:
        return self.pos_embeddings.weight

    def forward(self, size: int) -> torch.Tensor:
        if self.max_size < size:
            raise ValueError(
                f"The hidden size ({size }) is more than the config max_position_embeddings {self.max_size}"
            )
        return self.pos_embeddings.weight[None, :size]


class SinusoidalEncoding(nn.Module):
    """Fixed sin/cos table, sin on even and cos on odd channels (reference: :54-106). Like the
    reference the table is a plain tensor attribute, not a buffer; a device /
    dtype copy is cached for the fused embedding kernel."""

    def __init__(self, config) -> None:
        super().__init__()
        if config.hidden_size % 2 != 0:
            raise ValueError(f"Cannot use SinusoidalEncoding with odd hidden dim got dim {config.hidden_size}")
        self.positional_encoding = torch.zeros(1, config.max_position_embeddings, config.hidden_size)
        self.position = torch.arange(0, config.max_position_embeddings).unsqueeze(1)
        self.div_term = torch.exp(
            torch.arange(0, config.hidden_size, 2, dtype=torch.float) * -(torch.log(torch.tensor(10000.0)) / config.hidden_size)
        )
        self.positional_encoding[:, :, 0::2] = torch.sin(self.position.float() * self.div_term)
        self.positional_encoding[:, :, 1::2] = torch.cos(self.position.float() * self.div_term)
        self._dev_cache = {}

    def table(self, device=None, dtype=None) -> torch.Tensor:
        key = (device, dtype)
        if key not in self._dev_cache:
            self._dev_cache[key] = self.positional_encoding[0].to(device=device, dtype=dtype).contiguous()
        return self._dev_cache[key]

    def forward(self, seq_len: int) -> torch.Tensor:
        return self.positional_encoding[:, :seq_len]


class RotaryEmbedding(nn.Module):
    """Angle table theta[p, i] = p * 10000^(-2i/d), returned as (1, seq_len, d/2) (reference: :109-137)."""

    def __init__(self, config):
        super().__init__()
        dim = int(config.hidden_size // config.num_attention_heads)
        inv_freq = 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq)

    def forward(self, seq_len):
        t = torch.arange(seq_len, device=self.inv_freq.device).type_as(self.inv_freq)
        freqs = torch.einsum("i, j -> i j", t, self.inv_freq)
        return freqs[None, :, :]


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """(-x2, x1) for x = (x1 | x2) (reference: :140-152). Layout-only helper (no arithmetic kernel)."""
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, freqs, unsqueeze_dim=1) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stand-alone RoPE with the reference's signature (:155-182) for callers that use the layer API
    directly: q [B,h,S,d], k [B,h_kv,S,d], freqs (1,S,d/2). Runs the vy_rope kernel (head_dim 64);
    inside the models the rotation is fused into the QKV GEMM epilogue instead."""
    return ops.rope_apply(q, freqs), ops.rope_apply(k, freqs)


class VitAbsoluteEncoding(nn.Module):
    """ViT position table (reference: :188-226). The reference adds it IN PLACE and returns its input,
    and the caller adds the result again, so the effective stem is 2 * (x + pos) (SURVEY.md quirk
    Q1); Vit.forward fuses exactly that into the patch-embedding GEMM epilogue."""

    def __init__(self, config) -> None:
        super().__init__()
        image_height, image_width = config.image_size
        patch_height, patch_width = config.patch_size
        assert (
            image_height % patch_height == 0 and image_width % patch_width == 0
        ), "Image dimensions must be divisible by the patch size."
        num_patches = (image_height // patch_height) * (image_width // patch_width)
        patch_dim = config.num_channels * patch_height * patch_width
        self.pos_embeddings = nn.Parameter(torch.randn(1, num_patches + 1, patch_dim))
        self.register_buffer("num_patches", torch.arange(num_patches + 1).expand((1, -1)), persistent=False)

    def forward(self, img_seq: torch.Tensor) -> torch.Tensor:
        raise RuntimeError(
            "VitAbsoluteEncoding is consumed by the fused ViT stem (vyomai_b200.models.vision_encoder.Vit); "
            "it is not callable on its own in the sm_100a build"
        )
