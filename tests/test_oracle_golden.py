"""Pins the CPU oracle (oracle/vyom_oracle.py) against outputs of the REAL reference
(tests/golden/*.npz, written by tests/golden/make_golden.py). CPU only.

Tolerance: both sides are fp32 on the CPU and differ only in operation order (explicit softmax vs
SDPA's fused kernel, unfold-matmul vs conv), so outputs must agree to rel-L2 <= 2e-6 and
gradients to <= 2e-5; token ids and cache indexing must be bit-exact.
"""
import pytest
import torch

from oracle import vyom_oracle as O
from tests.conftest import load_fixture, rel_l2

FWD_TOL = 2e-6
GRAD_TOL = 2e-5


@pytest.mark.parametrize("name", ["encoder_rope_gqa", "encoder_absolute_mha"])
def test_encoder_forward_and_grads(name):
    fx = load_fixture(name)
    cfg, m = fx.cfg(), fx.meta
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in fx.sd.items()}
    out = O.encoder_forward(sd, cfg, fx.inputs["input_ids"], fx.inputs["attention_mask"], m["pos"], m["attn"])
    assert out.shape == fx.outputs["logits"].shape
    assert rel_l2(out, fx.outputs["logits"]) <= FWD_TOL
    (out * fx.inputs["cotangent"]).sum().backward()
    checked = 0
    for k, g in fx.outputs.items():
        if not k.startswith("grad::"):
            continue
        got = sd[k[6:]].grad
        assert got is not None, k
        if float(g.abs().max()) < 1e-6:  # key.bias without RoPE: exactly 0 (softmax shift invariance)
            assert float(got.abs().max()) < 1e-6, k
            continue
        assert rel_l2(got, g, floor=1e-4) <= GRAD_TOL, k
        checked += 1
    assert checked >= 10
    rows = fx.inputs["emb_rows"]
    assert rel_l2(sd["word_embeddings.weight"].grad[rows], fx.outputs["emb_grad_rows"]) <= GRAD_TOL


def test_encoder_mlm():
    fx = load_fixture("encoder_mlm_sinusoidal_mha")
    h, logits = O.encoder_mlm_forward(fx.sd, fx.cfg(), fx.inputs["input_ids"], fx.inputs["attention_mask"],
                                      fx.meta["pos"], fx.meta["attn"])
    assert rel_l2(h, fx.outputs["hidden_state"]) <= FWD_TOL
    assert rel_l2(logits, fx.outputs["logits"]) <= FWD_TOL


@pytest.mark.parametrize("name", ["decoder_rope_gqa", "decoder_absolute_mha", "decoder_rope_mha"])
def test_decoder_forward_cache_generate(name):
    fx = load_fixture(name)
    cfg, m = fx.cfg(), fx.meta
    pos, attn = m["pos"], m["attn"]
    h, logits = O.decoder_forward(fx.sd, cfg, fx.inputs["input_ids"], fx.inputs["attention_mask"], pos, attn)
    assert rel_l2(h, fx.outputs["hidden_state"]) <= FWD_TOL
    assert rel_l2(logits, fx.outputs["logits"]) <= FWD_TOL

    # prefill 4 + 3 decode steps through the static cache (batch 2, max_cache_len 12)
    prompt = fx.inputs["prompt"]
    cache = O.StaticCacheOneOracle(cfg.num_hidden_layers, 2, cfg.kv_heads(attn), 12, cfg.head_dim)
    am = torch.ones(2, 4, dtype=torch.long)
    _, l0 = O.decoder_forward(fx.sd, cfg, prompt, am, pos, attn, cache, 0)
    assert rel_l2(l0, fx.outputs["prefill_logits"]) <= FWD_TOL
    steps = []
    for t in range(3):
        tok = fx.inputs["decode_tokens"][:, t:t + 1]
        am = torch.cat([am, torch.ones(2, 1, dtype=torch.long)], dim=-1)
        _, lt = O.decoder_forward(fx.sd, cfg, tok, am, pos, attn, cache, 4 + t)
        steps.append(lt)
        # greedy ids bit-exact with the reference's next token
        if t < 2:
            assert torch.equal(lt[:, -1].argmax(-1), fx.inputs["decode_tokens"][:, t + 1])
    assert rel_l2(torch.cat(steps, 1), fx.outputs["decode_logits"]) <= FWD_TOL
    # cache contents: written slots match, untouched slots are still exactly zero
    k0, v1 = fx.outputs["key_cache_l0"], fx.outputs["value_cache_l1"]
    assert rel_l2(cache.key_cache[0], k0) <= FWD_TOL and rel_l2(cache.value_cache[1], v1) <= FWD_TOL
    assert torch.equal(cache.key_cache[0][:, :, 7:] == 0, k0[:, :, 7:] == 0)
    assert bool((cache.key_cache[0][:, :, 7:] == 0).all())

    for kind in (None, "dynamic", "static"):
        g = O.decoder_generate(fx.sd, cfg, fx.inputs["gen_prompt"], torch.ones(1, 4, dtype=torch.long), max_len=6,
                               pos_type=pos, attention_type=attn, cache_kind=kind)
        assert torch.equal(g, fx.outputs["generate"]), kind
    gb = O.decoder_generate(fx.sd, cfg, prompt, torch.ones(2, 4, dtype=torch.long), max_len=5, pos_type=pos,
                            attention_type=attn, cache_kind="static")
    assert torch.equal(gb, fx.outputs["generate_batch"])


def test_vit():
    fx = load_fixture("vit_small")
    out = O.vit_forward(fx.sd, fx.cfg(), fx.inputs["pixel_values"])
    assert out.shape == fx.outputs["logits"].shape
    assert rel_l2(out, fx.outputs["logits"]) <= 5e-6


def _vit_cfg(fx):
    v = fx.meta["vit"]
    return O.Cfg(hidden_size=v["hidden_size"], num_attention_heads=v["num_attention_heads"],
                 num_hidden_layers=v["num_hidden_layers"], layer_norm_eps=v["layer_norm_eps"],
                 hidden_act=v["hidden_act"], image_size=tuple(v["image_size"]), patch_size=tuple(v["patch_size"]),
                 num_channels=v["num_channels"])


def test_vlm_forward_generate_and_train_grads():
    fx = load_fixture("vlm_rope_gqa")
    cfg, vcfg, m = fx.cfg(), _vit_cfg(fx), fx.meta
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in fx.sd.items()}
    logits = O.vlm_forward(sd, cfg, vcfg, fx.inputs["pixel_values"], fx.inputs["input_ids"],
                           fx.inputs["attention_mask"], m["pos"], m["attn"])
    assert logits.shape == fx.outputs["logits"].shape  # (3, 18, V): one extra image position
    assert rel_l2(logits, fx.outputs["logits"]) <= 5e-6
    labels = fx.inputs["labels"]
    lg = logits[:, 1:]  # drop the image position, then the usual shift
    loss = O.cross_entropy_shifted(lg, labels)
    assert abs(float(loss) - float(fx.outputs["loss"][0])) <= 2e-6 * max(1.0, abs(float(loss)))
    loss.backward()
    n = 0
    for k, g in fx.outputs.items():
        if k.startswith("grad::"):
            if float(g.abs().max()) < 1e-6:
                assert float(sd[k[6:]].grad.abs().max()) < 1e-6, k
                continue
            assert rel_l2(sd[k[6:]].grad, g, floor=1e-4) <= 5e-5, k
            n += 1
    assert n >= 10
    with torch.no_grad():
        enc = O.vit_forward(fx.sd, vcfg, fx.inputs["pixel_values"][:1], pre="encoder.")[:, 0]
        assert rel_l2(enc, fx.outputs["encoder_output"]) <= 5e-6
        for use_cache in (False, True):
            g = O.vlm_generate(fx.sd, cfg, enc, fx.inputs["gen_start"], 6, m["pos"], m["attn"], use_cache)
            assert torch.equal(g, fx.outputs["generate"]), use_cache


def test_fully_masked_row_is_uniform_mean():
    """Quirk Q4: finfo.min masks (not -inf) make an all-masked row the uniform mean of V."""
    torch.manual_seed(0)
    q, k, v = torch.randn(1, 2, 3, 8), torch.randn(1, 2, 5, 8), torch.randn(1, 2, 5, 8)
    mask = O.encoder_mask(torch.zeros(1, 5), torch.float32)
    out = O.sdpa(q, k, v, mask)
    assert torch.allclose(out, v.mean(dim=2, keepdim=True).expand_as(out), atol=1e-6)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask)
    assert torch.allclose(out, ref, atol=1e-6)


def test_rope_matches_closed_form():
    f = O.rope_freqs(16, 8)
    q = torch.randn(1, 1, 16, 8)
    qr, _ = O.apply_rope(q, q, f)
    j = 1
    p = 5
    th = p * 10000 ** (-2 * j / 8)
    want = q[0, 0, p, j] * torch.cos(torch.tensor(th)) - q[0, 0, p, j + 4] * torch.sin(torch.tensor(th))
    assert abs(float(qr[0, 0, p, j] - want)) < 1e-5


# ---- real-width fixtures (tests/golden/make_golden_real.py): weights rebuilt from the seeded recipe ----
def _real_fixture(name, shapes_from):
    from tests.conftest import real_state_dict
    fx = load_fixture(name)
    return fx, real_state_dict(shapes_from(fx.meta), fx.meta["weight_seed"])


def _encoder_shapes(m, lm_head=False):
    H, V, kv = m["hidden_size"], m["vocab_size"], m["num_key_value_heads"] * (m["hidden_size"] // m["num_attention_heads"])
    s = {"word_embeddings.weight": (V, H)}
    for i in range(m["num_hidden_layers"]):
        p = f"all_layer.{i}."
        s.update({p + "attention.query.weight": (H, H), p + "attention.query.bias": (H,),
                  p + "attention.key.weight": (kv, H), p + "attention.key.bias": (kv,),
                  p + "attention.value.weight": (kv, H), p + "attention.value.bias": (kv,),
                  p + "attention.out.dense.weight": (H, H), p + "attention.out.dense.bias": (H,),
                  p + "attention.out.layernorm.weight": (H,), p + "attention.out.layernorm.bias": (H,),
                  p + "feed_forward.intermediate.weight": (4 * H, H), p + "feed_forward.intermediate.bias": (4 * H,),
                  p + "feed_forward.out.weight": (H, 4 * H), p + "feed_forward.out.bias": (H,),
                  p + "feed_forward.layernorm.weight": (H,), p + "feed_forward.layernorm.bias": (H,)})
    if lm_head:
        s.update({"lm_head.dense.weight": (H, H), "lm_head.dense.bias": (H,), "lm_head.layer_norm.weight": (H,),
                  "lm_head.layer_norm.bias": (H,), "lm_head.decoder.weight": (V, H), "lm_head.bias": (V,),
                  "lm_head.decoder.bias": (V,)})
    return s


def test_real_width_encoder_forward_and_grads():
    """C1 shape (8 x 128, H 768, 12 / 4 heads): the oracle against the reference's outputs and gradients."""
    fx, sd = _real_fixture("encoder_real_rope_gqa", _encoder_shapes)
    m = fx.meta
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ids, mask = fx.inputs["input_ids"], fx.inputs["attention_mask"]
    out = O.encoder_forward(sd, fx.cfg(), ids, mask, m["pos"], m["attn"])
    assert rel_l2(out[:, ::8], fx.outputs["logits_rows"]) <= FWD_TOL
    cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(m["cotangent_seed"])) * mask[..., None]
    (out * cot).sum().backward()
    n = 0
    for k, g in fx.outputs.items():
        if not k.startswith("grad::"):
            continue
        name = k[6:]
        got = sd[name].grad
        assert abs(float(got.norm()) - m["grad_norms"][name]) <= 1e-4 * max(m["grad_norms"][name], 1e-3), name
        if name == "word_embeddings.weight":
            got = got[fx.inputs["emb_rows"]]
        elif got.dim() == 2:
            got = got[:64, :64]
        assert rel_l2(got, g, floor=1e-4) <= 5 * GRAD_TOL, name
        n += 1
    assert n >= 17


def test_real_width_decoder_prefill_decode_generate():
    """200-token causal prefill + 3 decode steps through the static cache + greedy generate at H 768."""
    fx, sd = _real_fixture("decoder_real_rope_gqa", lambda m: _encoder_shapes(m, lm_head=True))
    m, cfg = fx.meta, fx.cfg()
    prompt = fx.inputs["prompt"]
    B, P = prompt.shape
    N = m["new_tokens"]
    cache = O.StaticCacheOneOracle(cfg.num_hidden_layers, B, cfg.kv_heads("gqa"), P + N, cfg.head_dim)
    h, logits = O.decoder_forward(sd, cfg, prompt, torch.ones(B, P, dtype=torch.long), "rope", "gqa", cache, 0)
    assert rel_l2(logits[:, -1], fx.outputs["prefill_last_logits"]) <= FWD_TOL
    assert rel_l2(h[:, ::25], fx.outputs["prefill_hidden_rows"]) <= FWD_TOL
    steps = []
    for t in range(3):
        _, lg = O.decoder_forward(sd, cfg, fx.inputs["decode_tokens"][:, t:t + 1], None, "rope", "gqa", cache, P + t)
        steps.append(lg)
    assert rel_l2(torch.cat(steps, 1), fx.outputs["decode_logits"]) <= FWD_TOL
    assert rel_l2(cache.key_cache[0][:, :, ::16], fx.outputs["key_cache_l0_s16"]) <= FWD_TOL
    assert rel_l2(cache.value_cache[0][:, :, ::16], fx.outputs["value_cache_l0_s16"]) <= FWD_TOL
    ids = O.decoder_generate(sd, cfg, prompt, torch.ones(B, P, dtype=torch.long), max_len=N, pos_type="rope",
                             attention_type="gqa", cache_kind="static")
    ref = fx.outputs["generate"]
    for i, mg in enumerate(m["generate_margins"]):  # fp32 vs fp32: identical wherever the margin is not rounding noise
        if mg <= 1e-4:
            break
        assert torch.equal(ids[:, P + i], ref[:, P + i]), (i, mg)


def test_philox_known_answers_and_dropout_mask():
    """Philox4x32-10 against the Random123 known-answer vectors; the keep mask has the stated rate and depends on every
    coordinate of its identity (seed, offset, step)."""
    import numpy as np
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = O.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(x) for x in got) == want
    a = O.dropout_keep_mask(256, 768, 0.1, 99, 1, 0)
    assert abs(float(a.float().mean()) - 0.9) < 3e-3
    for other in (O.dropout_keep_mask(256, 768, 0.1, 100, 1, 0), O.dropout_keep_mask(256, 768, 0.1, 99, 2, 0),
                  O.dropout_keep_mask(256, 768, 0.1, 99, 1, 1)):
        assert 0.1 < float((a != other).float().mean()) < 0.25  # independent masks differ on ~2 p (1 - p) of the elements
    assert torch.equal(a, O.dropout_keep_mask(256, 768, 0.1, 99, 1, 0))
    assert bool(O.dropout_keep_mask(8, 64, 0.0, 5, 0, 0).all())
