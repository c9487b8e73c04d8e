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
