"""Golden fixture for the RMSNorm / gated-MLP family: runs the REAL reference classes
`RMSNorm` and `MLP` of /root/reference/VyomAI/models/custom_transformer.py (loaded by path, unmodified) on seeded
inputs, fp32 and bf16, CPU, forward and backward, and stores inputs, weights, outputs and gradients.

    python tests/golden/make_golden_gated.py      # needs /root/reference; run in the build container
"""
import importlib.util
import os

import numpy as np
import torch

REF = os.environ.get("VYOM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_custom_transformer", os.path.join(REF, "VyomAI/models/custom_transformer.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    class Cfg:
        hidden_size, intermediate_size, hidden_act = 128, 320, "silu"

    torch.manual_seed(0)
    bf = lambda t: t.to(torch.bfloat16).to(torch.float32)  # bf16-representable values: exact in both paths
    x = bf(torch.randn(3, 5, 128) * 1.5)
    cot_n = bf(torch.randn(3, 5, 128))
    cot_m = bf(torch.randn(3, 5, 128))
    norm = ref.RMSNorm(128, eps=1e-6)
    mlp = ref.MLP(Cfg())
    with torch.no_grad():
        norm.weight.copy_(bf(1.0 + 0.3 * torch.randn(128)))
        for p in mlp.parameters():
            p.copy_(bf(p * 2.0))
    out = {"x": x.numpy(), "cot_norm": cot_n.numpy(), "cot_mlp": cot_m.numpy(), "norm_weight": norm.weight.detach().numpy(),
           "gate": mlp.gate_proj.weight.detach().numpy(), "up": mlp.up_proj.weight.detach().numpy(),
           "down": mlp.down_proj.weight.detach().numpy(), "eps": np.float32(1e-6)}
    for tag, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        n2, m2 = ref.RMSNorm(128, eps=1e-6).to(dt), ref.MLP(Cfg()).to(dt)
        n2.load_state_dict({k: v.to(dt) for k, v in norm.state_dict().items()})
        m2.load_state_dict({k: v.to(dt) for k, v in mlp.state_dict().items()})
        xi = x.detach().to(dt).clone().requires_grad_(True)
        y = n2(xi)
        y.backward(cot_n.to(dt))
        out[f"norm_y_{tag}"] = y.detach().float().numpy()
        out[f"norm_dx_{tag}"] = xi.grad.float().numpy()
        out[f"norm_dw_{tag}"] = n2.weight.grad.float().numpy()
        xi = x.detach().to(dt).clone().requires_grad_(True)
        y = m2(xi)
        y.backward(cot_m.to(dt))
        out[f"mlp_y_{tag}"] = y.detach().float().numpy()
        out[f"mlp_dx_{tag}"] = xi.grad.float().numpy()
        out[f"mlp_dgate_{tag}"] = m2.gate_proj.weight.grad.float().numpy()
        out[f"mlp_dup_{tag}"] = m2.up_proj.weight.grad.float().numpy()
        out[f"mlp_ddown_{tag}"] = m2.down_proj.weight.grad.float().numpy()
    np.savez_compressed(os.path.join(HERE, "gated_rmsnorm_mlp.npz"), **out)
    print("wrote gated_rmsnorm_mlp.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
