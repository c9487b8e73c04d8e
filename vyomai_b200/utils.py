"""Config / helper surface of VyomAI/utils.py that the hot path and the tests touch."""
from dataclasses import dataclass


@dataclass
class EncoderConfig:
    """reference: utils.py:89-100 (public surface: same field names and defaults)"""

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
