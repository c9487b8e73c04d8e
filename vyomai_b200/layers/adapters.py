"""LoRA / DoRA wrappers — host-side mirror of VyomAI/layers/adapters.py (PEFT wrappers around
nn.Linear, outside the benchmarked hot path, kept so `from VyomAI import LoraLinear, DoraLinear` and the reference's
tests/test_adapters.py work). Same parameter names (`linear`, `lora_a`, `lora_b`; `dora_m`, `dora_a`, `dora_b`); every
contraction runs on vy_gemm through `linear_fn` (forward, dgrad, wgrad)."""
from typing import Optional

import torch
import torch.nn as nn

from ..autograd import linear_fn
from ..models._common import ensure_cuda, back_to


def _lin(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    shape = x.shape
    y = linear_fn(x.reshape(-1, shape[-1]).contiguous(), w, b)
    return y.view(*shape[:-1], w.shape[0])


class LoraLinear(nn.Module):
    """linear(x) + alpha * (x A^T) B^T (reference: adapters.py:7-47)"""

    def __init__(self, linear_layer, rank: Optional[int] = 32, alpha: Optional[int] = 1, lora_dropout: Optional[float] = 0.0) -> None:
        super().__init__()
        self.linear = linear_layer
        self.in_features = self.linear.in_features
        self.out_features = self.linear.out_features
        self.rank = rank
        self.alpha = alpha
        std_dev = 1 / torch.sqrt(torch.tensor(self.rank).float())
        self.lora_a = nn.Parameter(torch.randn(self.rank, self.in_features) * std_dev)
        self.lora_b = nn.Parameter(torch.zeros(self.out_features, self.rank))
        self.dropout = nn.Dropout(lora_dropout)

    def forward(self, x) -> torch.Tensor:
        _dev, origin, (x,) = ensure_cuda(self, x)
        linear_output = _lin(x, self.linear.weight, self.linear.bias)
        lora_output = self.alpha * _lin(_lin(x, self.lora_a, None), self.lora_b, None)
        lora_output = self.dropout(lora_output)
        return back_to(origin, linear_output + lora_output)


class DoraLinear(nn.Module):
    """x (m * (W + A B) / ||W + A B||_col)^T + b (reference: adapters.py:50-75)"""

    def __init__(self, linear_layer, rank: Optional[int] = 32):
        super().__init__()
        self.linear = linear_layer
        self.in_features = self.linear.in_features
        self.out_features = self.linear.out_features
        self.dora_m = nn.Parameter(self.linear.weight.norm(p=2, dim=0, keepdim=True))
        std_dev = 1 / torch.sqrt(torch.tensor(rank).float())
        self.dora_a = nn.Parameter(torch.randn(self.out_features, rank) * std_dev)
        self.dora_b = nn.Parameter(torch.zeros(rank, self.in_features))

    def forward(self, x) -> torch.Tensor:
        _dev, origin, (x,) = ensure_cuda(self, x)
        # weight-space bookkeeping (a rank-r update and a column normalisation of the weight itself): parameter-sized
        # elementwise work outside the token path
        adapted = self.linear.weight + torch.matmul(self.dora_a, self.dora_b)
        calc_weights = self.dora_m * (adapted / adapted.norm(p=2, dim=0, keepdim=True))
        return back_to(origin, _lin(x, calc_weights.contiguous(), getattr(self.linear, "bias", None)))
