"""Golden fixture for the image-slot captioner ("Multimodal-II"): the model code of
Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1 is EXECUTED from the notebook where it lies under /root/reference
(nothing is copied into the repo: the class definitions between `class RotaryEmbedding` and `def build_string_from_input`
and `loss_fn` are exec'd with the handful of imports they need), on the CPU in fp32:

    python tests/golden/make_golden_slots.py

The notebook feeds an HF ViTModel's `last_hidden_state`; a checkpoint cannot be downloaded here, so the fixture's encoder is
the reference package's own Vit (all tokens of `.logits`) behind a two-line adapter exposing `.last_hidden_state` — the
captioner only ever sees the [B, n, H] feature tensor. Small widths; weights rounded to bf16-representable values."""
import io
import json
import math
import os
import sys
from contextlib import redirect_stdout
from dataclasses import dataclass
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, VitCfg, VlmTextCfg, cfg_meta, pack, round_weights_, save  # noqa: E402

NOTEBOOK = os.path.join(REF, "Examples", "vyom-ai-accelerate-multimodel-2t4.ipynb")
IMAGE_TOKEN = 100


@dataclass
class SlotTextCfg(VlmTextCfg):
    pad_token_id: int = 1
    max_position_embeddings: int = 64


def notebook_namespace():
    src = "".join(json.load(open(NOTEBOOK))["cells"][1]["source"])
    a, b = src.index("class RotaryEmbedding"), src.index("def build_string_from_input")
    import typing
    from einops import rearrange
    ns = {"math": math, "torch": torch, "nn": torch.nn, "rearrange": rearrange, "dataclass": dataclass}
    ns.update({k: getattr(typing, k) for k in ("List", "Optional", "Tuple", "Union", "Any", "Dict", "Generator")})
    exec(compile(src[a:b], NOTEBOOK, "exec"), ns)
    a, b = src.index("def loss_fn"), src.index("def prep")
    exec(compile(src[a:b], NOTEBOOK, "exec"), ns)
    return ns


def main():
    sys.path.insert(0, REF)
    from VyomAI import Vit
    ns = notebook_namespace()
    quiet = io.StringIO()
    torch.manual_seed(4242)
    cfg, vcfg = SlotTextCfg(), VitCfg()

    class AllTokens(torch.nn.Module):  # `.last_hidden_state` = every token of the reference package's Vit
        def __init__(self, vit):
            super().__init__()
            self.vit = vit

        def forward(self, pixel_values):
            return SimpleNamespace(last_hidden_state=self.vit(pixel_values=pixel_values).logits)

    with redirect_stdout(quiet):
        vit = Vit(vcfg)
        model = ns["VisionLanguageModel"](AllTokens(vit), cfg, decoder_pos_embedding_type="rope").eval()
    model.image_token_index = IMAGE_TOKEN
    round_weights_(model)
    n_img = (vcfg.image_size[0] // vcfg.patch_size[0]) * (vcfg.image_size[1] // vcfg.patch_size[1]) + 1  # 17
    B, S = 3, 32
    g = torch.Generator().manual_seed(5)
    ids = torch.full((B, S), cfg.pad_token_id, dtype=torch.long)
    mask = torch.zeros(B, S, dtype=torch.long)
    tt = torch.zeros(B, S, dtype=torch.long)
    for b, n_text in enumerate((13, 6, 9)):  # [bos, <Caption>] + 17 x <image> + caption ... + eos, right-padded
        row = [0, 5] + [IMAGE_TOKEN] * n_img + torch.randint(3, IMAGE_TOKEN, (n_text - 1,), generator=g).tolist() + [2]
        ids[b, :len(row)] = torch.tensor(row)
        mask[b, :len(row)] = 1
        tt[b, 2 + n_img:len(row)] = 1
    px = torch.rand(B, vcfg.num_channels, *vcfg.image_size, generator=g)

    # training form: causal x padding; loss_fn + labels of main()
    out = model(pixel_values=px, input_ids=ids, attention_mask=mask, token_type_ids=tt)
    labels = ids.masked_fill(ids == IMAGE_TOKEN, -100)
    labels = torch.where(ids == cfg.pad_token_id, -100, labels)
    loss = ns["loss_fn"](out.logits, labels, mask, cfg)
    loss.backward()
    keep = ("decoder.all_layer.0.attention.query.weight", "decoder.all_layer.1.feed_forward.out.weight", "lm_head.dense.weight",
            "encoder.vit.all_layer.1.attention.qkv.weight", "encoder.vit.all_layer.0.feed_forward.intermediate.weight")
    grads = {"grad::" + k: p.grad for k, p in model.named_parameters() if p.grad is not None and (p.dim() == 1 or k in keep)}
    grads["grad::decoder.word_embeddings.weight"] = model.decoder.word_embeddings.weight.grad
    with torch.no_grad():
        feats = model.get_encoder_output(px)
        infer = model(pixel_values=px, input_ids=ids, attention_mask=mask).logits  # inference prefill: whole prefix visible
        # cached path, batch 1 (the notebook's StaticCache is batch-1): unpadded prefill, then three single-token steps
        L = int(mask[1].sum())
        one_ids, one_mask = ids[1:2, :L], mask[1:2, :L]
        with redirect_stdout(quiet):
            model._setup_cache(cfg, cls=ns["StaticCache"])
        pre = model(pixel_values=px[1:2], input_ids=one_ids, attention_mask=one_mask, use_cache=True, start_pos=0).logits
        steps, toks = [], []
        nxt = pre[:, -1].argmax(-1, keepdim=True)
        am = one_mask
        for t in range(3):
            toks.append(nxt)
            am = torch.cat([am, torch.ones(1, 1, dtype=torch.long)], dim=-1)
            lg = model(input_ids=nxt, attention_mask=am, use_cache=True, start_pos=L + t).logits
            steps.append(lg)
            nxt = lg[:, -1].argmax(-1, keepdim=True)
        model._clean_cache()
    meta = cfg_meta(cfg, pos="rope", attn=None, image_token_index=IMAGE_TOKEN, n_image_tokens=n_img,
                    vit={k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg_meta(vcfg).items()})
    save("slot_captioner_rope_mha",
         pack(model, {"pixel_values": px, "input_ids": ids, "attention_mask": mask, "token_type_ids": tt,
                      "decode_tokens": torch.cat(toks, 1)},
              {"image_features": feats, "train_logits": out.logits, "loss": loss.detach().reshape(1), "infer_logits": infer,
               "cached_prefill_logits": pre, "cached_step_logits": torch.cat(steps, 1), **grads}, meta))


if __name__ == "__main__":
    main()
