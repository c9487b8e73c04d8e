"""Pieces shared by the model classes: device policy, LM head, embedding stem."""
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib
from .. import functional as F
from ..autograd import embed_fn, lm_head_fn, lm_head_loss_fn
from ..layers.positional_embeddings import AbsoluteEncoding, RotaryEmbedding, SinusoidalEncoding

_position_embeddings = {"absolute": AbsoluteEncoding, "sinusoidal": SinusoidalEncoding}


def ensure_cuda(module: nn.Module, *tensors):
    """Device policy of the sm_100a build. There is no CPU compute path, but the reference's tests
    build models and inputs on the CPU when they want to (tests/test_vision_encoder.py:22-27), so a
    module whose parameters still sit on the CPU is moved to the current CUDA device on first use,
    CPU inputs are copied over, and the caller is told which device the inputs came from so outputs
    can be returned there. Without a CUDA device this raises — nothing falls back."""
    if not torch.cuda.is_available():
        raise _lib.VyomError("vyomai_b200 needs a CUDA sm_100a device; there is no CPU fallback")
    p = next(module.parameters())
    if not p.is_cuda:
        module.to(torch.device("cuda", torch.cuda.current_device()))
        p = next(module.parameters())
    dev = p.device
    origin = None
    moved = []
    for t in tensors:
        if t is None:
            moved.append(None)
            continue
        if origin is None:
            origin = t.device
        moved.append(t if t.device == dev else t.to(dev))
    return dev, origin, moved


def back_to(origin, t: torch.Tensor) -> torch.Tensor:
    if origin is None or t.device == origin:
        return t
    return t.to(origin)


class LMHead(nn.Module):
    """dense -> GELU -> LayerNorm -> decoder (reference: models/decoder.py:253-275, encoder.py:67-90).
    `bias` and `decoder.bias` are one Parameter under two state_dict keys, as in the reference."""

    def __init__(self, config) -> None:
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.layer_norm = nn.LayerNorm(config.hidden_size, eps=getattr(config, "layer_norm_eps", 1e-6))
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))
        self.decoder.bias = self.bias

    def forward(self, hidden_state: torch.Tensor) -> torch.Tensor:
        shape = hidden_state.shape
        logits = lm_head_fn(self, hidden_state.reshape(-1, shape[-1]))
        return logits.view(*shape[:-1], logits.shape[-1])

    def loss(self, hidden_state: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
        """F.cross_entropy(self(hidden).view(-1, V), labels.view(-1), ignore_index=...) without materialising the
        logits more than once (training path; `labels` has one entry per hidden row)."""
        return lm_head_loss_fn(self, hidden_state.reshape(-1, hidden_state.shape[-1]), labels, ignore_index)


class TextStem:
    """Mixin: word embeddings + absolute / sinusoidal / RoPE positions (reference:
    models/encoder.py:100-117,146-154; models/decoder.py:288-307,343-354)."""

    def _build_stem(self, config, pos_embedding_type: Optional[str], who: str) -> None:
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size,
                                            padding_idx=getattr(config, "pad_token_id", None))
        if _position_embeddings.get(pos_embedding_type, None) is not None:
            self.position_embeddings = _position_embeddings.get(pos_embedding_type)(config)
        else:
            self.position_embeddings = None
        self._rope = None
        if pos_embedding_type == "rope":
            self.emb_freq = RotaryEmbedding(config)(config.max_position_embeddings)  # plain tensor (quirk Q5)
            self._rope = F.RopeTables(self.emb_freq)
            print(f"{who} Ignoring sinusoidal or absolute position embeddings because rope,is enable")

    def _pos_table(self, device, dtype) -> Optional[torch.Tensor]:
        pe = self.position_embeddings
        if pe is None:
            return None
        if isinstance(pe, AbsoluteEncoding):
            return pe.table()
        return pe.table(device, dtype)

    def _check_positions(self, end: int) -> None:
        pe = self.position_embeddings
        limit = None
        if isinstance(pe, AbsoluteEncoding):
            limit = pe.max_size
        elif isinstance(pe, SinusoidalEncoding):
            limit = pe.positional_encoding.shape[1]
        elif self._rope is not None:
            limit = self.emb_freq.shape[1]
        if limit is not None and end > limit:
            raise ValueError(f"The hidden size ({end}) is more than the config max_position_embeddings {limit}")

    def _embed(self, input_ids: torch.Tensor, start_pos: int, extra: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B * (S + (extra is not None)), H) hidden rows: word embeddings (+ absolute / sinusoidal position
        rows starting at start_pos), optionally behind one leading row per sequence taken from `extra`."""
        table = self.word_embeddings.weight
        pos = self._pos_table(table.device, table.dtype)
        pe = self.position_embeddings
        pos_pad = pe.pos_embeddings.padding_idx if isinstance(pe, AbsoluteEncoding) else None
        return embed_fn(input_ids, table, pos, start_pos, input_ids.shape[1], extra,
                        padding_idx=self.word_embeddings.padding_idx, pos_padding_idx=pos_pad)
