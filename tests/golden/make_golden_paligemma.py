"""Golden fixture for the PaliGemma-scale scratch model (BASELINE config 5): the model code of Examples/paligemma.ipynb
(cell 28 StaticCache, cells 9-17 SigLIP / Gemma / PaliGemmaForConditionalGeneration) is EXECUTED from the notebook where it
lies under /root/reference — nothing is copied into the repo — with tiny random-init configs that keep the two head dims of
the real model (SigLIP 72, Gemma 256 with ONE kv head), on the CPU in fp32:

    python tests/golden/make_golden_paligemma.py

Outputs: image features, the batched inference prefill (right padding), the notebook's own generation procedure (cell 30
`test_inference`: static cache of 24 slots, batch 1, greedy) as logits per step + ids, and the training-form logits
(token_type_ids + labels: prefix-LM mask) for the mask restatement. Weights rounded to bf16-representable values."""
import json
import math
import os
import sys
from dataclasses import dataclass

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, round_weights_, save  # noqa: E402

NOTEBOOK = os.path.join(REF, "Examples", "paligemma.ipynb")
IMAGE_TOKEN = 290
CFG = {
    "hidden_size": 256, "projection_dim": 256, "image_token_index": IMAGE_TOKEN, "pad_token_id": 0,
    "vision": {"hidden_size": 144, "intermediate_size": 288, "num_hidden_layers": 2, "num_attention_heads": 2, "num_channels": 3,
               "image_size": 28, "patch_size": 14, "layer_norm_eps": 1e-6},
    "text": {"vocab_size": 300, "hidden_size": 256, "intermediate_size": 512, "num_hidden_layers": 2, "num_attention_heads": 2,
             "num_key_value_heads": 1, "head_dim": 256, "max_position_embeddings": 64, "rms_norm_eps": 1e-6, "rope_theta": 10000.0},
}


def notebook_namespace():
    cells = json.load(open(NOTEBOOK))["cells"]
    import typing
    from einops import rearrange
    ns = {"math": math, "torch": torch, "nn": torch.nn, "rearrange": rearrange, "dataclass": dataclass}
    ns.update({k: getattr(typing, k) for k in ("List", "Optional", "Tuple", "Union", "Any", "Dict")})
    # cell 28 subclasses transformers' `Cache`, whose constructor took no arguments in the release the notebook was written
    # against (4.47) and demands layer classes in the installed 5.x: stand in the old, behaviour-free base class
    import types
    shim = types.ModuleType("transformers.cache_utils")
    shim.Cache = type("Cache", (), {"__init__": lambda self: None})
    real = sys.modules.get("transformers.cache_utils")
    sys.modules["transformers.cache_utils"] = shim
    try:
        exec(compile("".join(cells[28]["source"]), f"{NOTEBOOK}:cell28", "exec"), ns)
    finally:
        if real is not None:
            sys.modules["transformers.cache_utils"] = real
        else:
            del sys.modules["transformers.cache_utils"]
    ns["Cache"] = shim.Cache
    for i in (9, 11, 12, 13, 15, 16, 17):  # StaticCache first: cell 17 tests isinstance(past_key_values, StaticCache)
        exec(compile("".join(cells[i]["source"]), f"{NOTEBOOK}:cell{i}", "exec"), ns)
    return ns


def main():
    ns = notebook_namespace()
    torch.manual_seed(777)
    config = ns["PaliGemmaConfig"](vision_config=dict(CFG["vision"]), text_config=dict(CFG["text"]), image_token_index=IMAGE_TOKEN,
                                   vocab_size=CFG["text"]["vocab_size"], projection_dim=CFG["projection_dim"], hidden_size=CFG["hidden_size"],
                                   pad_token_id=CFG["pad_token_id"])
    model = ns["PaliGemmaForConditionalGeneration"](config).eval()
    with torch.no_grad():  # GemmaRMSNorm weights start at zero: give the (1 + w) scale something to do
        for n, p in model.named_parameters():
            if n.endswith("layernorm.weight") or n.endswith("model.norm.weight"):
                p.normal_(0.0, 0.1)
    round_weights_(model)
    n_img = (CFG["vision"]["image_size"] // CFG["vision"]["patch_size"]) ** 2  # 4
    g = torch.Generator().manual_seed(8)
    B, S = 3, 14
    ids = torch.full((B, S), CFG["pad_token_id"], dtype=torch.long)
    mask = torch.zeros(B, S, dtype=torch.long)
    tt = torch.zeros(B, S, dtype=torch.long)
    labels = torch.full((B, S), -100, dtype=torch.long)
    for b, (n_prompt, n_suffix) in enumerate(((5, 5), (3, 2), (6, 4))):  # <image> x 4 + bos + prompt | suffix, right-padded
        row = [IMAGE_TOKEN] * n_img + torch.randint(1, IMAGE_TOKEN, (n_prompt + n_suffix,), generator=g).tolist()
        ids[b, :len(row)] = torch.tensor(row)
        mask[b, :len(row)] = 1
        tt[b, n_img + n_prompt:len(row)] = 1
        labels[b, n_img + n_prompt:len(row)] = ids[b, n_img + n_prompt:len(row)]
    px = torch.rand(B, 3, CFG["vision"]["image_size"], CFG["vision"]["image_size"], generator=g)
    out = {}
    with torch.no_grad():
        out["image_features"] = model.get_image_features(px)
        out["siglip_last_hidden"] = model.vision_tower(px).last_hidden_state
        out["prefill_logits"] = model(input_ids=ids, pixel_values=px, attention_mask=mask).logits        # inference form, no cache
        tr = model(input_ids=ids, pixel_values=px, attention_mask=mask, token_type_ids=tt, labels=labels)  # training form
        out["train_logits"], out["train_loss"] = tr.logits, tr.loss.reshape(1)
        # cell 30 test_inference, verbatim procedure (batch 1, static cache, pixel_values passed on every step)
        L = int(mask[1].sum())
        cur_ids, cur_mask, pv = ids[1:2, :L], mask[1:2, :L], px[1:2]
        cache = ns["StaticCache"](config.text_config, batch_size=1, device="cpu", dtype=torch.float32, max_cache_len=24)
        toks, step_logits = [], []
        for _ in range(6):
            o = model(input_ids=cur_ids, pixel_values=pv, attention_mask=cur_mask, past_key_values=cache, use_cache=True)
            cache = o.past_key_values
            nl = o.logits[:, -1, :]
            step_logits.append(nl)
            nxt = torch.argmax(nl, dim=-1, keepdim=True)
            toks.append(nxt)
            cur_ids = nxt
            cur_mask = torch.cat([cur_mask, torch.ones((1, 1), dtype=cur_mask.dtype)], dim=-1)
        out["gen_step_logits"] = torch.stack(step_logits, 1)  # [1, 6, V]
        out["generate"] = torch.cat(toks, 1)
        top2 = out["gen_step_logits"][0].topk(2, dim=-1).values
        margins = (top2[:, 0] - top2[:, 1]).tolist()
        out["key_cache_l0"] = cache.key_cache[0].clone()
    blob = {}
    for k, v in model.state_dict().items():
        blob["w::" + k] = v.detach().bfloat16().view(torch.int16).numpy().view(np.uint16)
    for k, v in {"input_ids": ids, "attention_mask": mask, "token_type_ids": tt, "labels": labels, "pixel_values": px}.items():
        blob["in::" + k] = v.numpy()
    for k, v in out.items():
        blob["out::" + k] = v.detach().numpy()
    meta = dict(CFG, generate_margins=margins, cache_len=24, gen_row=1, tied_lm_head=False,
                hidden_size_=CFG["hidden_size"], num_attention_heads=2, num_hidden_layers=2, layer_norm_eps=1e-6, hidden_act="gelu")
    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    save("paligemma_tiny", blob)


if __name__ == "__main__":
    main()
