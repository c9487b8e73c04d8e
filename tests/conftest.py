import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100a device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container; run with -m gpu on a B200")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Fixture:
    """A golden fixture written by tests/golden/make_golden.py (outputs of the REAL reference)."""

    def __init__(self, name: str):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.sd, self.inputs, self.outputs = {}, {}, {}
        for k in z.files:
            if k.startswith("w::"):
                bits = torch.from_numpy(z[k].astype(np.int16))
                self.sd[k[3:]] = bits.view(torch.bfloat16).float()
            elif k.startswith("wi::"):
                self.sd[k[4:]] = torch.from_numpy(z[k])
            elif k.startswith("in::"):
                self.inputs[k[4:]] = torch.from_numpy(z[k])
            elif k.startswith("out::"):
                self.outputs[k[5:]] = torch.from_numpy(z[k])

    def cfg(self):
        from oracle.vyom_oracle import Cfg
        m = self.meta
        return Cfg(
            hidden_size=m["hidden_size"], num_attention_heads=m["num_attention_heads"],
            num_key_value_heads=m.get("num_key_value_heads"),
            max_position_embeddings=m.get("max_position_embeddings", 514),
            num_hidden_layers=m["num_hidden_layers"], vocab_size=m.get("vocab_size", 0),
            layer_norm_eps=m["layer_norm_eps"], hidden_act=m["hidden_act"],
            image_size=tuple(m.get("image_size", (224, 224))), patch_size=tuple(m.get("patch_size", (16, 16))),
            num_channels=m.get("num_channels", 3),
        )


def load_fixture(name: str) -> Fixture:
    return Fixture(name)


def rel_l2(a: torch.Tensor, b: torch.Tensor, floor: float = 0.0) -> float:
    """||a-b|| / max(||b||, floor). `floor` keeps mathematically-zero tensors (e.g. the key-bias
    gradient, which softmax shift-invariance makes exactly 0) from comparing rounding noise."""
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / max(float(b.norm()), floor, 1e-30))


def seeded_state_dict(shapes: dict, seed: int) -> dict:
    """{key: shape} -> {key: tensor}: N(0, 0.02) matrices / embeddings / biases, 1 + N(0, 0.1) norm scales, drawn key by
    key in sorted order from one CPU generator and rounded to bf16-representable fp32. The real-width fixtures
    (tests/golden/make_golden_real.py) store no weights: generator and tests both rebuild them with this recipe."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        t = torch.randn(shp, generator=g)
        if ("layernorm.weight" in k or "layer_norm.weight" in k) and len(shp) == 1:
            t = 1.0 + 0.1 * t
        else:
            t = 0.02 * t
        out[k] = t.bfloat16().float()
    return out


def real_state_dict(model_or_shapes, seed: int) -> dict:
    """seeded_state_dict over the floating-point entries of a module's state_dict (or a {key: shape} dict), with the LM
    head's shared bias present under both of its keys."""
    shapes = model_or_shapes
    if not isinstance(shapes, dict):
        shapes = {k: tuple(v.shape) for k, v in model_or_shapes.state_dict().items() if v.dtype.is_floating_point}
    sd = seeded_state_dict(shapes, seed)
    if "lm_head.bias" in sd:
        sd["lm_head.decoder.bias"] = sd["lm_head.bias"]
    return sd
