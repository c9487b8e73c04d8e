"""Feed-forward block — host-side mirror of VyomAI/layers/ffn.py.

Same module / parameter names (`intermediate`, `out`, `layernorm`, `dropout`, `act_fn`) and the
same quirks: the inner width is `multiplier * hidden_size` (config.intermediate_size is ignored,
ffn.py:21-23) and the residual added before the LayerNorm is whatever the caller passes as
`input_tensor` — the layer INPUT in every model. The forward is two vy_gemm
calls (bias + activation epilogue; bias + residual epilogue) and one vy_add_layernorm_fwd.
"""
from typing import Union

import torch
import torch.nn as nn

from .. import _lib

# activation names of ffn.py:7-15 that the GEMM epilogue implements
_FUSED_ACT = {"gelu": "gelu"}
_REF_ACT = ("gelu", "leaky_relu", "relu6", "sigmoid", "silu", "swish", "tanh")


class FeedForward(nn.Module):
    def __init__(self, config, multiplier: Union[int, float] = 4) -> None:
        super().__init__()
        self.intermediate = nn.Linear(config.hidden_size, int(multiplier) * config.hidden_size)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        name = getattr(config, "hidden_act", None)
        if name not in _REF_ACT:
            name = "gelu"  # ffn.py:26-29: unknown / missing names fall back to exact GELU
        if name not in _FUSED_ACT:
            raise _lib.VyomError(
                f"hidden_act={name!r}: the sm_100a GEMM epilogue implements exact-erf GELU (the default of every "
                "reference config) and tanh-GELU; other `_ACT_` entries are not built in this round"
            )
        self._act_name = _FUSED_ACT[name]
        self.act_fn = nn.GELU()  # attribute kept for parity; the activation itself runs in the epilogue
        self.out = nn.Linear(int(multiplier) * config.hidden_size, config.hidden_size)

    def forward(self, hidden_state: torch.Tensor, input_tensor: torch.Tensor) -> torch.Tensor:
        from ..autograd import feed_forward_fn
        shape = input_tensor.shape
        H = shape[-1]
        y = feed_forward_fn(self, hidden_state.reshape(-1, hidden_state.shape[-1]), input_tensor.reshape(-1, H))
        return y.view(shape)
