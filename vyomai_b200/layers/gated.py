"""RMSNorm and the gated (SwiGLU) MLP of the reference's HF-style decoder, on the fused kernels.

Same class names, constructor arguments, parameter names and forward signatures as
VyomAI/models/custom_transformer.py (`RMSNorm(hidden_size, eps)`: :227-244; `MLP(config)` with bias-free
gate_proj / up_proj / down_proj: :76-89), so `state_dict`s are interchangeable.
  RMSNorm  one vy_add_layernorm_fwd launch in VY_NORM_RMS mode (bwd: the single-pass LayerNorm backward with the mean
           terms switched off)
  MLP      gate and up projections as ONE vy_gemm over the row-interleaved weight with the SwiGLU epilogue
           (VY_ACT_SWIGLU: out = silu(gate) * up, pre-activations saved for backward), then the down projection;
           backward = vy_swiglu_bwd + the ordinary dgrad / wgrad GEMMs against the interleaved weight.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class _RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, weight, eps, kind):
        y, _, _, rstd = ops.add_layernorm(x2d, None, weight, None, eps, save_stats=True, kind=kind)
        ctx.kind = kind
        ctx.save_for_backward(x2d, weight, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, weight, rstd = ctx.saved_tensors
        dx, dgamma, _ = ops.add_layernorm_bwd(dy.contiguous(), x2d, weight, None, rstd, kind=ctx.kind)
        return dx, dgamma.to(weight.dtype), None, None


class RMSNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def forward(self, hidden_states):
        shape = hidden_states.shape
        x2d = hidden_states.reshape(-1, shape[-1]).contiguous()
        return _RMSNormFn.apply(x2d, self.weight, self.variance_epsilon, "rmsnorm").view(shape)

    def extra_repr(self):
        return f"{tuple(self.weight.shape)}, eps={self.variance_epsilon}"


def interleave_gate_up(w_gate: torch.Tensor, w_up: torch.Tensor) -> torch.Tensor:
    """[2 I, K] with row 2j = gate_j, row 2j+1 = up_j: the operand layout of the SwiGLU epilogue."""
    return torch.stack((w_gate, w_up), dim=1).reshape(2 * w_gate.shape[0], w_gate.shape[1])


class _GatedMLPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, w_gate, w_up, w_down):
        w_int = interleave_gate_up(w_gate, w_up)
        z = torch.empty((x2d.shape[0], w_int.shape[0]), device=x2d.device, dtype=x2d.dtype)
        h = ops.gemm(x2d, w_int, act="swiglu", aux=z)
        y = ops.gemm(h, w_down)
        ctx.save_for_backward(x2d, z, h, w_int, w_down)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, z, h, w_int, w_down = ctx.saved_tensors
        dy = dy.contiguous()
        d_down = ops.gemm(dy.t(), h.t(), allow_split_k=True)          # [H, I]
        dh = ops.gemm(dy, w_down.t())                                  # [rows, I]
        dz = ops.swiglu_bwd(dh, z)                                     # [rows, 2 I]
        dx = ops.gemm(dz, w_int.t())                                   # [rows, H]
        d_int = ops.gemm(dz.t(), x2d.t(), allow_split_k=True)         # [2 I, H]
        return dx, d_int[0::2], d_int[1::2], d_down


class MLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.hidden_size = config.hidden_size
        self.intermediate_size = config.intermediate_size
        self.gate_proj = nn.Linear(self.hidden_size, self.intermediate_size, bias=False)
        self.up_proj = nn.Linear(self.hidden_size, self.intermediate_size, bias=False)
        self.down_proj = nn.Linear(self.intermediate_size, self.hidden_size, bias=False)
        act = getattr(config, "hidden_act", "silu")
        if act not in ("silu", "swish"):
            raise ValueError(f"MLP: the fused gated epilogue implements SiLU (got hidden_act={act!r})")

    def forward(self, x):
        shape = x.shape
        x2d = x.reshape(-1, shape[-1]).contiguous()
        y = _GatedMLPFn.apply(x2d, self.gate_proj.weight, self.up_proj.weight, self.down_proj.weight)
        return y.view(*shape[:-1], self.hidden_size)
