"""The image-slot captioner ("Multimodal-II", SURVEY.md §8 config C4 notebook-II form; reference:
Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1 — VisionLanguageModel / _update_causal_mask / loss_fn / main()).

CPU: the oracle (slot_vlm_forward / slot_loss) against outputs of the notebook's own code (tests/golden/
slot_captioner_rope_mha.npz, written by make_golden_slots.py by executing the notebook cell).
-m gpu: ImageSlotVisionLanguageModel on the sm_100a path against the same fixture — training logits (causal x padding),
loss, gradients (incl. the ViT, reached through all 17 image tokens, and the embedding table, whose <image> and pad rows get
none), the inference prefill (whole prefix visible), the cached prefill + three single-token steps, and a Trainer step."""
import pytest
import torch

from oracle import vyom_oracle as O
from tests.conftest import load_fixture, rel_l2

NAME = "slot_captioner_rope_mha"


def _vit_cfg(fx):
    v = fx.meta["vit"]
    return O.Cfg(hidden_size=v["hidden_size"], num_attention_heads=v["num_attention_heads"], num_key_value_heads=None,
                 max_position_embeddings=0, num_hidden_layers=v["num_hidden_layers"], vocab_size=0, layer_norm_eps=v["layer_norm_eps"],
                 hidden_act=v["hidden_act"], image_size=tuple(v["image_size"]), patch_size=tuple(v["patch_size"]),
                 num_channels=v["num_channels"])


def test_oracle_slot_captioner_matches_the_notebook():
    fx = load_fixture(NAME)
    cfg, m = fx.cfg(), fx.meta
    tok = m["image_token_index"]
    ids, mask = fx.inputs["input_ids"], fx.inputs["attention_mask"]
    feats = O.vit_forward(fx.sd, _vit_cfg(fx), fx.inputs["pixel_values"], pre="encoder.vit.")
    assert rel_l2(feats, fx.outputs["image_features"]) <= 2e-6
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in fx.sd.items()}
    feats_g = O.vit_forward(sd, _vit_cfg(fx), fx.inputs["pixel_values"], pre="encoder.vit.")
    logits = O.slot_vlm_forward(sd, cfg, feats_g, ids, mask, True, tok)
    assert rel_l2(logits, fx.outputs["train_logits"]) <= 2e-6
    loss = O.slot_loss(logits, ids, mask, m["pad_token_id"], tok)
    assert abs(float(loss) - float(fx.outputs["loss"])) <= 1e-5
    loss.backward()
    n = 0
    for k, g in fx.outputs.items():
        if not k.startswith("grad::"):
            continue
        key = "lm_head.bias" if k[6:] == "lm_head.vocab.bias" else k[6:]
        got = sd[key].grad
        assert got is not None, k
        if float(g.abs().max()) < 1e-7:
            assert float(got.abs().max()) < 1e-7, k
            continue
        assert rel_l2(got, g, floor=1e-5) <= 5e-5, k
        n += 1
    assert n >= 30
    gemb = sd["decoder.word_embeddings.weight"].grad
    assert float(gemb[tok].abs().max()) == 0.0  # masked_scatter overwrote every <image> embedding: no gradient reaches that row
    with torch.no_grad():
        infer = O.slot_vlm_forward(fx.sd, cfg, fx.outputs["image_features"], ids, mask, False, tok)
        assert rel_l2(infer, fx.outputs["infer_logits"]) <= 2e-6
        assert rel_l2(infer, fx.outputs["train_logits"]) > 1e-2  # the two forms really differ (bidirectional prefix)
        L = int(mask[1].sum())
        cache = O.StaticCacheOneOracle(cfg.num_hidden_layers, 1, cfg.num_attention_heads, L + 3, cfg.head_dim)
        pre = O.slot_vlm_forward(fx.sd, cfg, fx.outputs["image_features"][1:2], ids[1:2, :L], mask[1:2, :L], False, tok, cache=cache)
        assert rel_l2(pre, fx.outputs["cached_prefill_logits"]) <= 2e-6
        am = mask[1:2, :L]
        for t in range(3):
            am = torch.cat([am, torch.ones(1, 1, dtype=am.dtype)], dim=-1)
            lg = O.slot_vlm_forward(fx.sd, cfg, None, fx.inputs["decode_tokens"][:, t:t + 1], am, False, tok, cache=cache, start_pos=L + t)
            assert rel_l2(lg, fx.outputs["cached_step_logits"][:, t:t + 1]) <= 2e-6


def _build(fx, dtype):
    from tests.test_gpu_models import _cfg_obj, _load
    from vyomai_b200 import Vit
    from vyomai_b200.models.multimodel_slots import ImageSlotVisionLanguageModel
    m = fx.meta
    cfg = _cfg_obj(m, drop=("pos", "attn", "vit", "image_token_index", "n_image_tokens"))
    vcfg = _cfg_obj(m["vit"])
    model = ImageSlotVisionLanguageModel(Vit(vcfg), cfg, decoder_pos_embedding_type="rope")
    model.image_token_index = m["image_token_index"]
    sd = {k.replace("encoder.vit.", "encoder."): v for k, v in fx.sd.items()}
    return _load(model, sd, dtype), cfg


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_slot_captioner_matches_the_notebook(dtype):
    from tests.test_gpu_models import TOL
    from tests.test_gpu_train import GTOL
    from vyomai_b200 import StaticCache
    from vyomai_b200.models.multimodel_slots import slot_caption_labels
    fx = load_fixture(NAME)
    m = fx.meta
    model, cfg = _build(fx, dtype)
    model.train()
    px = fx.inputs["pixel_values"].cuda().to(dtype)
    ids, mask, tt = fx.inputs["input_ids"].cuda(), fx.inputs["attention_mask"].cuda(), fx.inputs["token_type_ids"].cuda()
    logits = model(pixel_values=px, input_ids=ids, attention_mask=mask, token_type_ids=tt).logits
    valid = fx.inputs["attention_mask"].bool()  # pad QUERY rows are arbitrary in both (they attend to garbage-free but unused keys)
    assert rel_l2(logits.float().cpu()[valid], fx.outputs["train_logits"][valid]) <= TOL[dtype]
    labels = slot_caption_labels(ids, mask, m["pad_token_id"], m["image_token_index"])
    loss = model.forward_loss(px, ids, mask, labels)
    assert abs(float(loss) - float(fx.outputs["loss"])) <= (2e-3 if dtype == torch.float32 else 2e-2)
    loss.backward()
    params = dict(model.named_parameters())
    n, worst = 0, (0.0, None)
    for k, g in fx.outputs.items():
        if not k.startswith("grad::"):
            continue
        key = k[6:].replace("encoder.vit.", "encoder.")
        key = "lm_head.bias" if key == "lm_head.vocab.bias" else key
        got = params[key].grad
        assert got is not None, k
        got = got.float().cpu()
        if float(g.abs().max()) < 1e-7:
            assert float(got.abs().max()) < 2e-3, k
            continue
        r = rel_l2(got, g, floor=1e-4)
        worst = max(worst, (r, k))
        assert r <= GTOL[dtype], (k, r)
        n += 1
    assert n >= 30
    gemb = params["decoder.word_embeddings.weight"].grad.float().cpu()
    assert float(gemb[m["image_token_index"]].abs().max()) == 0.0 and float(gemb[m["pad_token_id"]].abs().max()) == 0.0
    print(f"slot captioner {dtype}: {n} gradient tensors within {GTOL[dtype]}, worst {worst}")

    model.eval()
    with torch.no_grad():
        infer = model(pixel_values=px, input_ids=ids, attention_mask=mask).logits
        assert rel_l2(infer.float().cpu()[valid], fx.outputs["infer_logits"][valid]) <= TOL[dtype]
        L = int(fx.inputs["attention_mask"][1].sum())
        model._setup_cache(cfg, cls=StaticCache)
        pre = model(pixel_values=px[1:2], input_ids=ids[1:2, :L], attention_mask=mask[1:2, :L], use_cache=True, start_pos=0).logits
        assert rel_l2(pre.float().cpu(), fx.outputs["cached_prefill_logits"]) <= TOL[dtype]
        am = mask[1:2, :L]
        for t in range(3):
            am = torch.cat([am, torch.ones(1, 1, dtype=am.dtype, device="cuda")], dim=-1)
            tok = fx.inputs["decode_tokens"][:, t:t + 1].cuda()
            lg = model(input_ids=tok, attention_mask=am, use_cache=True, start_pos=L + t).logits
            assert rel_l2(lg.float().cpu(), fx.outputs["cached_step_logits"][:, t:t + 1]) <= TOL[dtype]
            lg2 = model(input_ids=tok, attention_mask=None, use_cache=True, start_pos=L + t).logits  # no padding: the decode kernel
            assert rel_l2(lg2.float().cpu(), fx.outputs["cached_step_logits"][:, t:t + 1]) <= TOL[dtype]
        model._clean_cache()


@pytest.mark.gpu
def test_slot_merge_kernels_and_trainer_step():
    from vyomai_b200 import ops
    from vyomai_b200.models.multimodel_slots import image_slots, slot_caption_labels
    from vyomai_b200.trainer import Trainer
    g = torch.Generator().manual_seed(3)
    for dtype, H in ((torch.bfloat16, 768), (torch.float32, 100)):
        a = torch.randn(40, H, generator=g).to(dtype).cuda()
        b = torch.randn(11, H, generator=g).to(dtype).cuda()
        ids = torch.randint(0, 5, (4, 10), generator=g)
        ids[ids == 4] = 3
        ids[0, 2:6] = 4
        ids[2, 0:5] = 4
        ids[3, 9] = 4
        slot = image_slots(ids.cuda(), 4)
        ref = a.clone()
        ref[(ids == 4).reshape(-1).cuda()] = b[:10]  # masked_scatter order: row-major over the batch
        out = ops.slot_merge(a, b, slot)
        assert torch.equal(out, ref)
        dout = torch.randn(40, H, generator=g).to(dtype).cuda()
        da, db = ops.slot_merge_bwd(dout, slot, 11)
        is_img = (ids == 4).reshape(-1).cuda()
        assert torch.equal(da[~is_img], dout[~is_img]) and float(da[is_img].abs().max()) == 0.0
        assert torch.equal(db[:10], dout[is_img]) and float(db[10].abs().max()) == 0.0
    fx = load_fixture(NAME)
    m = fx.meta
    model, cfg = _build(fx, torch.bfloat16)
    model.train()
    tr = Trainer(model, lr=1e-3, weight_decay=0.0, max_grad_norm=1.0, use_graph=False)
    px = fx.inputs["pixel_values"].cuda().to(torch.bfloat16)
    ids, mask = fx.inputs["input_ids"].cuda(), fx.inputs["attention_mask"].cuda()
    labels = slot_caption_labels(ids, mask, m["pad_token_id"], m["image_token_index"])
    losses = [float(tr.caption_step(px, ids, mask, labels)) for _ in range(8)]
    assert abs(losses[0] - float(fx.outputs["loss"])) < 3e-2 and losses[-1] < losses[0] - 0.3, losses
    assert tr.grad_overwrite
