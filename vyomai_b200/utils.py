"""Config / helper surface of VyomAI/utils.py that the hot path and the tests touch."""
from dataclasses import dataclass
from typing import Dict

import torch.nn as nn


def model_size(model: nn.Module) -> float:
    """Parameter + buffer footprint in MiB (reference: utils.py:9-18)."""
    size = sum(p.nelement() * p.element_size() for p in model.parameters())
    size += sum(b.nelement() * b.element_size() for b in model.buffers())
    return size / 1024**2


def model_parameters(model: nn.Module) -> Dict[str, int]:
    """reference: utils.py:37-40"""
    return {
        "total_params": sum(p.numel() for p in model.parameters()),
        "trainable_params": sum(p.numel() for p in model.parameters() if p.requires_grad),
    }


def init_weights(module: nn.Module) -> None:
    """reference: utils.py:21-34"""
    if isinstance(module, nn.Linear):
        module.weight.data.normal_(mean=0.0, std=0.02)
        if module.bias is not None:
            module.bias.data.zero_()
    elif isinstance(module, nn.Embedding):
        module.weight.data.normal_(mean=0.0, std=0.02)
        if module.padding_idx is not None:
            module.weight.data[module.padding_idx].zero_()
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)


@dataclass
class EncoderConfig:
    """reference: utils.py:89-100 (public surface, kept verbatim)"""

    hidden_size: int = 768
    num_attention_heads: int = 12
    max_position_embeddings: int = 514
    num_hidden_layers: int = 4
    vocab_size: int = 50265
    hidden_dropout_prob: float = 0.1
    initializer_range: float = 0.02
    intermediate_size: int = 3072
    layer_norm_eps: float = 1e-05
    hidden_act: str = "gelu"
