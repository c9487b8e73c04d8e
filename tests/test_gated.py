"""RMSNorm / gated MLP. The fixture tests/golden/gated_rmsnorm_mlp.npz holds outputs and
gradients of the REAL reference classes (models/custom_transformer.py RMSNorm, MLP; tests/golden/make_golden_gated.py).
CPU: the oracle restatement reproduces them. GPU: the fused kernels behind the same class names do.
Tolerances (rel-L2): fp32 path (tf32 GEMMs) 6e-3, bf16 path 2e-2 forward / 3e-2 gradients; RMSNorm alone (no GEMM)
fp32 1e-5."""
import os

import numpy as np
import pytest
import torch

from oracle import vyom_oracle as O
from tests.conftest import rel_l2

FIX = np.load(os.path.join(os.path.dirname(__file__), "golden", "gated_rmsnorm_mlp.npz"))
T = lambda k: torch.from_numpy(FIX[k])


@pytest.mark.parametrize("tag,dt,tol", [("f32", torch.float32, 1e-6), ("bf16", torch.bfloat16, 1e-2)])
def test_oracle_reproduces_the_reference_rmsnorm_and_gated_mlp(tag, dt, tol):
    x = T("x").to(dt)
    y = O.rms_norm(x, T("norm_weight").to(dt), float(FIX["eps"]))
    assert rel_l2(y.float(), T(f"norm_y_{tag}")) <= tol
    y = O.gated_mlp(x, T("gate").to(dt), T("up").to(dt), T("down").to(dt))
    assert rel_l2(y.float(), T(f"mlp_y_{tag}")) <= max(tol, 2e-6)
    # gemma flavour and shift are plain algebra on top of the pinned xhat
    w = T("norm_weight")
    assert torch.allclose(O.rms_norm(T("x"), w - 1.0, 1e-6, gemma=True), O.rms_norm(T("x"), w, 1e-6), atol=1e-6)
    assert torch.allclose(O.rms_norm(T("x"), w, 1e-6, shift=w), O.rms_norm(T("x"), w, 1e-6) + w)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,dt,tol_n,tol_f,tol_g", [("f32", torch.float32, 1e-5, 6e-3, 6e-3), ("bf16", torch.bfloat16, 1e-2, 2e-2, 3e-2)])
def test_fused_rmsnorm_and_gated_mlp_match_the_reference(tag, dt, tol_n, tol_f, tol_g):
    from vyomai_b200.layers.gated import MLP, RMSNorm

    class Cfg:
        hidden_size, intermediate_size, hidden_act = 128, 320, "silu"

    dev = "cuda"
    norm = RMSNorm(128, eps=float(FIX["eps"])).to(dev).to(dt)
    mlp = MLP(Cfg()).to(dev).to(dt)
    with torch.no_grad():
        norm.weight.copy_(T("norm_weight"))
        mlp.gate_proj.weight.copy_(T("gate"))
        mlp.up_proj.weight.copy_(T("up"))
        mlp.down_proj.weight.copy_(T("down"))
    x = T("x").to(dev).to(dt).requires_grad_(True)
    y = norm(x)
    y.backward(T("cot_norm").to(dev).to(dt))
    assert rel_l2(y.float().cpu(), T(f"norm_y_{tag}")) <= tol_n
    assert rel_l2(x.grad.float().cpu(), T(f"norm_dx_{tag}")) <= max(tol_n, 2e-5) * 2
    assert rel_l2(norm.weight.grad.float().cpu(), T(f"norm_dw_{tag}")) <= max(tol_n, 2e-5) * 2
    x = T("x").to(dev).to(dt).requires_grad_(True)
    y = mlp(x)
    y.backward(T("cot_mlp").to(dev).to(dt))
    assert rel_l2(y.float().cpu(), T(f"mlp_y_{tag}")) <= tol_f
    assert rel_l2(x.grad.float().cpu(), T(f"mlp_dx_{tag}")) <= tol_g
    assert rel_l2(mlp.gate_proj.weight.grad.float().cpu(), T(f"mlp_dgate_{tag}")) <= tol_g
    assert rel_l2(mlp.up_proj.weight.grad.float().cpu(), T(f"mlp_dup_{tag}")) <= tol_g
    assert rel_l2(mlp.down_proj.weight.grad.float().cpu(), T(f"mlp_ddown_{tag}")) <= tol_g


@pytest.mark.gpu
def test_rmsnorm_kinds_residual_and_shift_against_the_oracle():
    from vyomai_b200 import ops
    torch.manual_seed(1)
    x = torch.randn(37, 768)
    r = torch.randn(37, 768)
    w = 1.0 + 0.2 * torch.randn(768)
    b = 0.1 * torch.randn(768)
    for kind, gemma in (("rmsnorm", False), ("gemma_rmsnorm", True)):
        for shift in (None, b):
            y, s, _, rstd = ops.add_layernorm(x.cuda(), r.cuda(), w.cuda(), None if shift is None else shift.cuda(), 1e-6,
                                              save_stats=True, save_sum=True, kind=kind)
            ref = O.rms_norm(x + r, w, 1e-6, gemma=gemma, shift=shift)
            assert rel_l2(y.cpu(), ref) <= 1e-5, (kind, shift is not None)
            assert torch.allclose(s.cpu(), x + r)
            # backward against autograd through the oracle
            xs = (x + r).clone().requires_grad_(True)
            ws = w.clone().requires_grad_(True)
            cot = torch.randn(37, 768)
            O.rms_norm(xs, ws, 1e-6, gemma=gemma, shift=shift).backward(cot)
            dx, dg, db = ops.add_layernorm_bwd(cot.cuda(), s, w.cuda(), None, rstd, kind=kind)
            assert rel_l2(dx.cpu(), xs.grad) <= 2e-5 and rel_l2(dg.cpu(), ws.grad) <= 2e-5
            assert rel_l2(db.cpu(), cot.sum(0)) <= 2e-5  # the shift's gradient
