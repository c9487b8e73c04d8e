"""Generates the golden fixtures in this directory by running the REAL reference
(/root/reference/VyomAI, unmodified, imported — nothing is copied) on seeded inputs, fp32, CPU.

    python tests/golden/make_golden.py        # needs /root/reference; run in the build container

The reference's own tests pin shapes only, so these fixtures are what pins numerical parity: the
oracle (oracle/vyom_oracle.py) is checked against them in tests/test_oracle_golden.py, and the
CUDA path is checked against them in the `-m gpu` tests. Weights are rounded to bf16-representable
values before the reference runs, and stored as bf16 bit patterns (uint16), so the same weights
are exact in both the fp32 and the bf16 paths and the files stay small.
"""
import io
import json
import os
import sys
from contextlib import redirect_stdout
from dataclasses import dataclass
from typing import Tuple

import numpy as np
import torch

REF = os.environ.get("VYOM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


@dataclass
class TextCfg:
    hidden_size: int = 128
    num_attention_heads: int = 2
    max_position_embeddings: int = 64
    num_hidden_layers: int = 2
    vocab_size: int = 101
    hidden_dropout_prob: float = 0.1
    initializer_range: float = 0.02
    intermediate_size: int = 512
    layer_norm_eps: float = 1e-05
    hidden_act: str = "gelu"


@dataclass
class TextCfgGqa(TextCfg):
    num_key_value_heads: int = 1


@dataclass
class VitCfg:
    hidden_size: int = 192
    num_attention_heads: int = 3
    image_size: Tuple[int, int] = (32, 32)
    patch_size: Tuple[int, int] = (8, 8)
    num_channels: int = 3
    num_hidden_layers: int = 2
    hidden_dropout_prob: float = 0.1
    initializer_range: float = 0.02
    intermediate_size: int = 768
    layer_norm_eps: float = 1e-05
    hidden_act: str = "gelu"


@dataclass
class VlmTextCfg(TextCfg):
    hidden_size: int = 192
    num_attention_heads: int = 3
    intermediate_size: int = 768


@dataclass
class VlmTextCfgGqa(VlmTextCfg):
    num_key_value_heads: int = 1


IDS = torch.tensor(
    [
        [0, 2387, 766, 16, 181, 967, 46035, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1],
        [0, 12196, 16, 110, 766, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1],
        [0, 37111, 1137, 162, 110, 766, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1],
    ],
    dtype=torch.long,
)  # tests/test_encoder.py:28-37 (reference fixture), folded into the small vocab below
MASK = torch.tensor(
    [
        [1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    ],
    dtype=torch.long,
)  # tests/test_encoder.py:39-46


def fold(ids, vocab):
    out = ids.clone()
    big = out >= vocab
    out[big] = 3 + out[big] % (vocab - 3)
    return out


def round_weights_(model):
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(p.bfloat16().float())


def pack(model, inputs, outputs, meta):
    blob = {}
    for k, v in model.state_dict().items():
        if v.dtype.is_floating_point:
            blob["w::" + k] = v.detach().bfloat16().view(torch.int16).numpy().view(np.uint16)
        else:
            blob["wi::" + k] = v.numpy()
    for k, v in inputs.items():
        blob["in::" + k] = v.detach().numpy()
    for k, v in outputs.items():
        blob["out::" + k] = v.detach().numpy()
    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return blob


def save(name, blob):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path}  ({os.path.getsize(path) / 1e6:.2f} MB)")


def cfg_meta(cfg, **kw):
    d = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.__dict__.items()}
    if not d:  # dataclass defaults live on the class
        d = {f: (list(getattr(cfg, f)) if isinstance(getattr(cfg, f), tuple) else getattr(cfg, f))
             for f in cfg.__dataclass_fields__}
    d.update(kw)
    return d


def main():
    sys.path.insert(0, REF)
    import VyomAI  # noqa: F401  the real reference
    from VyomAI import (DecoderModel, DynamicCache, EncoderForMaskedLM, EncoderModel, StaticCache,
                        VisionLanguageModel, Vit, generate_multimodel)
    from VyomAI.layers.kv_cache import StaticCacheOne

    assert os.path.realpath(VyomAI.__file__).startswith(os.path.realpath(REF)), VyomAI.__file__
    quiet = io.StringIO()

    # ---------------- encoder: forward + gradients -------------------------------------------
    for pos, attn, cfgcls in (("rope", "gqa", TextCfgGqa), ("absolute", None, TextCfg)):
        torch.manual_seed(1234)
        cfg = cfgcls()
        with redirect_stdout(quiet):
            model = EncoderModel(cfg, pos_embedding_type=pos, attention_type=attn).eval()
        round_weights_(model)
        ids = fold(IDS, cfg.vocab_size)
        out = model(ids, MASK).logits
        g = torch.Generator().manual_seed(7)
        cot = torch.randn(out.shape, generator=g) * MASK[..., None]
        (out * cot).sum().backward()
        grads = {"grad::" + k: p.grad for k, p in model.named_parameters()
                 if p.grad is not None and (p.dim() == 1 or k in (
                     "all_layer.0.attention.query.weight", "all_layer.1.attention.key.weight",
                     "all_layer.0.attention.out.dense.weight", "all_layer.1.feed_forward.out.weight",
                     "all_layer.0.feed_forward.intermediate.weight"))}
        emb_grad_rows = model.word_embeddings.weight.grad[ids.unique()]
        outs = {"logits": out, "emb_grad_rows": emb_grad_rows, **grads}
        save(f"encoder_{pos}_{attn or 'mha'}",
             pack(model, {"input_ids": ids, "attention_mask": MASK, "cotangent": cot, "emb_rows": ids.unique()}, outs,
                  cfg_meta(cfg, pos=pos, attn=attn)))

    # ---------------- masked-LM head ------------------------------------------------------------
    torch.manual_seed(99)
    cfg = TextCfg()
    with redirect_stdout(quiet):
        model = EncoderForMaskedLM(cfg, pos_embedding_type="sinusoidal", attention_type=None).eval()
    round_weights_(model)
    ids = fold(IDS, cfg.vocab_size)
    o = model(ids, MASK)
    save("encoder_mlm_sinusoidal_mha", pack(model, {"input_ids": ids, "attention_mask": MASK},
                                             {"hidden_state": o.hidden_state, "logits": o.logits},
                                             cfg_meta(cfg, pos="sinusoidal", attn=None)))

    # ---------------- decoder: forward, cached prefill+decode, greedy generate -----------------------
    for pos, attn, cfgcls in (("rope", "gqa", TextCfgGqa), ("absolute", None, TextCfg), ("rope", None, TextCfg)):
        torch.manual_seed(4321)
        cfg = cfgcls()
        with redirect_stdout(quiet):
            model = DecoderModel(cfg, pos_embedding_type=pos, attention_type=attn).eval()
        round_weights_(model)
        ids = fold(IDS, cfg.vocab_size)
        with torch.no_grad():
            full = model(ids, MASK)
            # prefill 4 tokens + 3 single-token decode steps through the static cache, batch 2
            prompt = fold(torch.tensor([[9226, 16, 5, 1296], [0, 2387, 766, 16]]), cfg.vocab_size)
            am = torch.ones(2, 4, dtype=torch.long)
            kv = StaticCacheOne(cfg, max_cache_len=12, batch_size=2)
            o0 = model(prompt, am, use_cache=True, kv_cache=kv, start_pos=0)
            steps = []
            nxt_tokens = []
            nxt = o0.logits[:, -1].argmax(-1, keepdim=True)
            for t in range(3):
                nxt_tokens.append(nxt)
                am = torch.cat([am, torch.ones(2, 1, dtype=torch.long)], dim=-1)
                ot = model(nxt, am, use_cache=True, kv_cache=kv, start_pos=4 + t)
                steps.append(ot.logits)
                nxt = ot.logits[:, -1].argmax(-1, keepdim=True)
            k0 = kv.key_cache[0].clone()
            v1 = kv.value_cache[1].clone()
            gen_prompt = fold(torch.tensor([[9226, 16, 5, 1296]]), cfg.vocab_size)  # tests/test_decoder.py:150-151
            gmask = torch.ones(1, 4, dtype=torch.long)
            g_nocache = model.generate(gen_prompt, gmask, max_len=6, use_cache=False)
            g_dynamic = model.generate(gen_prompt, gmask, max_len=6, use_cache=True)
            g_static = model.generate(gen_prompt, gmask, max_len=6, use_cache=True, use_static_cache=True)
            assert torch.equal(g_nocache, g_dynamic) and torch.equal(g_nocache, g_static)
            g_batch = model.generate(prompt, torch.ones(2, 4, dtype=torch.long), max_len=5, use_cache=True,
                                     use_static_cache=True)
        save(f"decoder_{pos}_{attn or 'mha'}",
             pack(model,
                  {"input_ids": ids, "attention_mask": MASK, "prompt": prompt, "decode_tokens": torch.cat(nxt_tokens, 1),
                   "gen_prompt": gen_prompt},
                  {"logits": full.logits, "hidden_state": full.hidden_state, "prefill_logits": o0.logits,
                   "decode_logits": torch.cat(steps, 1), "key_cache_l0": k0, "value_cache_l1": v1,
                   "generate": g_static, "generate_batch": g_batch},
                  cfg_meta(cfg, pos=pos, attn=attn)))

    # ---------------- ViT -------------------------------------------------------------------------
    torch.manual_seed(2024)
    vcfg = VitCfg()
    vit = Vit(vcfg).eval()
    round_weights_(vit)
    px = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        vout = vit(px).logits
    save("vit_small", pack(vit, {"pixel_values": px}, {"logits": vout}, cfg_meta(vcfg)))

    # ---------------- captioner (image-text fusion VLM) ------------------------------------------
    vcfg = VitCfg(num_hidden_layers=1)
    for pos, attn, cfgcls in (("rope", "gqa", VlmTextCfgGqa),):
        torch.manual_seed(777)
        cfg = cfgcls(num_hidden_layers=1)
        with redirect_stdout(quiet):
            vlm = VisionLanguageModel(cfg, encoder=Vit(vcfg), pos_embedding_type=pos, attention_type=attn).eval()
        round_weights_(vlm)
        ids = fold(IDS, cfg.vocab_size)
        px3 = torch.rand(3, 3, 32, 32, generator=torch.Generator().manual_seed(6))
        with torch.no_grad():
            lg = vlm(pixel_values=px3, decoder_input_ids=ids, decoder_attention_mask=MASK).logits
            enc = vlm.get_encoder_output(px3[:1])
            start = torch.tensor([[0]])
            gen_nc = generate_multimodel(vlm, enc, None, start, max_new_tokens=6, use_cache=False)
            vlm._setup_cache(cfg, cls=DynamicCache)
            gen_dc = generate_multimodel(vlm, enc, None, start, max_new_tokens=6, use_cache=True)
            vlm._clean_cache()
            vlm._setup_cache(cfg, cls=StaticCache)
            gen_sc = generate_multimodel(vlm, enc, None, start, max_new_tokens=6, use_cache=True)
            vlm._clean_cache()
            assert torch.equal(gen_nc, gen_dc) and torch.equal(gen_nc, gen_sc), (gen_nc, gen_dc, gen_sc)
        # training-step material: CE loss on shifted labels + a few gradients (dropout off)
        lg2 = vlm(pixel_values=px3, decoder_input_ids=ids, decoder_attention_mask=MASK).logits
        labels = ids.masked_fill(MASK == 0, -100)
        # logits have one extra leading position (the image token): position i+1 predicts token i+1
        loss = torch.nn.functional.cross_entropy(lg2[:, 1:-1].reshape(-1, cfg.vocab_size), labels[:, 1:].reshape(-1),
                                                 ignore_index=-100)
        loss.backward()
        keep = ("decoder.all_layer.0.attention.query.weight", "decoder.all_layer.0.feed_forward.intermediate.weight",
                "encoder.all_layer.0.attention.qkv.weight", "encoder.pixel_seq.weight", "encoder.cls_token",
                "decoder.lm_head.dense.weight")
        grads = {"grad::" + k: p.grad for k, p in vlm.named_parameters() if p.grad is not None and (p.dim() == 1 or k in keep)}
        save(f"vlm_{pos}_{attn or 'mha'}",
             pack(vlm, {"pixel_values": px3, "input_ids": ids, "attention_mask": MASK, "labels": labels, "gen_start": start},
                  {"logits": lg, "encoder_output": enc, "generate": gen_nc, "loss": loss.detach().reshape(1), **grads},
                  cfg_meta(cfg, pos=pos, attn=attn, vit=cfg_meta(vcfg))))


if __name__ == "__main__":
    main()
