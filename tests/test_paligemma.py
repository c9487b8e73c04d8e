"""PaliGemma-scale scratch model (BASELINE config 5; reference: Examples/paligemma.ipynb cells 9-17, 28, 30).

CPU: the oracle (siglip_forward / gemma_layers / paligemma_mask / paligemma_forward) against outputs of the notebook's own
classes (tests/golden/paligemma_tiny.npz, written by make_golden_paligemma.py by executing the notebook cells): SigLIP
features, inference prefill, training-form (prefix-LM mask) logits, and cell 30's static-cache generation step by step.
-m gpu: the kernels that config needs beyond head_dim 64 — vy_attn_fwd on the mma.sync kernel (head dims 72 / 128 / 256 / 64,
key padding, causal, prefix-LM, packed-head decode), RoPE for head_dim 256, the GeGLU epilogue, padded patch rows — against
the oracle, and the model (bf16) against the fixture: features, prefill logits, generation logits and greedy ids (margin rule)."""
import pytest
import torch

from oracle import vyom_oracle as O
from tests.conftest import load_fixture, rel_l2

NAME = "paligemma_tiny"


def _new_cache(m, batch=1):
    t = m["text"]
    shape = (batch, t["num_key_value_heads"], m["cache_len"], t["head_dim"])
    return ([torch.zeros(shape) for _ in range(t["num_hidden_layers"])], [torch.zeros(shape) for _ in range(t["num_hidden_layers"])])


def test_oracle_paligemma_matches_the_notebook():
    fx = load_fixture(NAME)
    m = fx.meta
    ids, mask, px = fx.inputs["input_ids"], fx.inputs["attention_mask"], fx.inputs["pixel_values"]
    v = m["vision"]
    feats = O.siglip_forward(fx.sd, "vision_tower.vision_model.", px, v["patch_size"], v["num_hidden_layers"], v["num_attention_heads"],
                             v["layer_norm_eps"])
    assert rel_l2(feats, fx.outputs["siglip_last_hidden"]) <= 2e-6
    proj = O.linear(feats, fx.sd["multi_modal_projector.linear.weight"], fx.sd["multi_modal_projector.linear.bias"]) / m["hidden_size"] ** 0.5
    assert rel_l2(proj, fx.outputs["image_features"]) <= 2e-6
    assert rel_l2(O.paligemma_forward(fx.sd, m, ids, px, mask), fx.outputs["prefill_logits"]) <= 2e-6
    train = O.paligemma_forward(fx.sd, m, ids, px, mask, token_type_ids=fx.inputs["token_type_ids"])
    assert rel_l2(train, fx.outputs["train_logits"]) <= 2e-6
    assert rel_l2(train, fx.outputs["prefill_logits"]) > 1e-2  # the prefix-LM mask really differs from the inference one
    row = m["gen_row"]
    L = int(mask[row].sum())
    cache = _new_cache(m)
    cur, cm, seen, toks = ids[row:row + 1, :L], mask[row:row + 1, :L], 0, []
    for s in range(fx.outputs["generate"].shape[1]):
        lg = O.paligemma_forward(fx.sd, m, cur, px[row:row + 1], cm, cache=cache, seen=seen, cache_len=m["cache_len"])
        assert rel_l2(lg[:, -1], fx.outputs["gen_step_logits"][:, s]) <= 2e-6
        seen += cur.shape[1]
        cur = lg[:, -1].argmax(-1, keepdim=True)
        toks.append(cur)
        cm = torch.cat([cm, torch.ones(1, 1, dtype=cm.dtype)], -1)
    assert torch.equal(torch.cat(toks, 1), fx.outputs["generate"])
    assert rel_l2(cache[0][0], fx.outputs["key_cache_l0"]) <= 2e-6


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _r(shape, seed, scale=1.0):
    return (torch.randn(shape, generator=_gen(seed)) * scale).bfloat16().float()


def _ref_attention(q, k, v, mask_vis, n_rep):
    """softmax(q k^T / sqrt(d) + (1 - vis) * finfo.min) v in fp32 with the oracle's sdpa (vis: [B, 1|Hq, Sq, Skv] bool)."""
    add = torch.where(mask_vis, 0.0, torch.finfo(torch.float32).min)
    return O.merge_heads(O.sdpa(q, O.repeat_kv(k, n_rep), O.repeat_kv(v, n_rep), add))


@pytest.mark.gpu
@pytest.mark.parametrize("case", [
    # B, Hq, Hkv, Sq, Skv, D, mask kind
    (2, 16, 16, 256, 256, 72, "none"), (3, 4, 4, 50, 50, 72, "pad"), (2, 8, 1, 264, 264, 256, "pad"), (2, 8, 1, 70, 70, 256, "causal"),
    (2, 8, 1, 1, 300, 256, "decode"), (4, 8, 1, 1, 37, 256, "decode_pad"), (2, 4, 2, 130, 130, 128, "prefix"), (2, 12, 4, 100, 100, 64, "prefix"),
    (1, 2, 2, 3, 200, 128, "causal_offset"), (2, 6, 3, 1, 65, 64, "decode"), (2, 4, 4, 40, 40, 72, "allpad_row"),
], ids=lambda c: "B%d_h%d_%d_Sq%d_Skv%d_d%d_%s" % c)
def test_attn_fwd_other_head_dims_and_prefix_lm_vs_oracle(case):
    from vyomai_b200 import ops
    B, Hq, Hkv, Sq, Skv, D, kind = case
    q, k, v = _r((B, Hq, Sq, D), 1, 0.7), _r((B, Hkv, Skv, D), 2, 0.7), _r((B, Hkv, Skv, D), 3)
    vis = torch.ones(B, 1, Sq, Skv, dtype=torch.bool)
    kpm = prefix = None
    causal, q_pos0 = False, 0
    kk, ll = torch.arange(Skv)[None, :], torch.arange(Sq)[:, None]
    if kind in ("pad", "decode_pad", "allpad_row"):
        lens = torch.randint(max(1, Skv // 3), Skv + 1, (B,), generator=_gen(4))
        if kind == "allpad_row":
            lens[0] = 0  # a sequence with no visible key at all: uniform mean over every key (quirk Q4)
        kpm = (torch.arange(Skv)[None, :] < lens[:, None])
        vis = vis & kpm[:, None, None, :]
    if kind in ("causal", "prefix", "causal_offset"):
        causal = True
        q_pos0 = Skv - Sq
        cm = kk <= q_pos0 + ll
        if kind == "prefix":
            prefix = torch.tensor([Skv // 3, 0][:B] + [5] * max(0, B - 2), dtype=torch.int32)
            cm = cm[None] | (kk[None] < prefix[:, None, None].long())
            vis = vis & cm[:, None]
        else:
            vis = vis & cm[None, None]
    ref = _ref_attention(q, k, v, vis, Hq // Hkv)
    # operands as strided views of packed buffers (the way the models hand them over)
    qp = torch.zeros(B, Sq, Hq + 1, D, dtype=torch.bfloat16, device="cuda")
    qp[:, :, :Hq] = q.permute(0, 2, 1, 3).bfloat16().cuda()
    out, lse = ops.attn_fwd(qp[:, :, :Hq].permute(0, 2, 1, 3), k.bfloat16().cuda(), v.bfloat16().cuda(), causal=causal, q_pos0=q_pos0,
                            key_padding_mask=None if kpm is None else kpm.to(torch.uint8).cuda(),
                            prefix_len=None if prefix is None else prefix.cuda(), out_dtype=torch.float32, need_lse=True)
    assert list(out.shape) == [B, Sq, Hq * D]
    assert rel_l2(out.cpu(), ref) <= 8e-3, kind
    # log2-domain logsumexp of the visible scores
    sc = torch.einsum("bhid,bhjd->bhij", q, O.repeat_kv(k, Hq // Hkv)) / D ** 0.5
    sc = torch.where(vis.expand_as(sc), sc, torch.full_like(sc, -1e30))
    want = torch.logsumexp(sc, -1) * 1.4426950408889634
    rows_ok = vis.expand_as(sc).any(-1)
    assert float((lse.cpu() - want)[rows_ok].abs().max()) <= 2e-2


@pytest.mark.gpu
def test_rope_256_geglu_and_padded_patches_vs_oracle():
    from vyomai_b200 import ops
    # RoPE, head_dim 256, 1-indexed positions, written into a cache slot range
    B, Hh, S, D, start = 2, 3, 5, 256, 7
    x = _r((B, Hh, S, D), 5)
    pos = (torch.arange(start, start + S) + 1)[None].expand(B, -1)
    want, _ = O.gemma_rope(x, x, pos, 10000.0)
    inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
    ang = torch.arange(64, dtype=torch.float32)[:, None] * inv[None]
    cos, sin = ang.cos().contiguous().cuda(), ang.sin().contiguous().cuda()
    cache = torch.zeros(B, Hh, 32, D, dtype=torch.bfloat16, device="cuda")
    ops.rope_into(x.bfloat16().cuda(), cache[:, :, start:start + S], cos, sin, start + 1)
    assert rel_l2(cache[:, :, start:start + S].float().cpu(), want) <= 4e-3
    assert float(cache[:, :, :start].abs().max()) == 0.0 and float(cache[:, :, start + S:].abs().max()) == 0.0
    # the same as ONE launch over a packed q|k|v projection: q rotated in place, k rotated into the cache, v copied into it
    nq, nkv, S2 = 4, 1, 3
    packed = _r((B, S2, (nq + 2 * nkv) * D), 11).bfloat16().cuda()
    v4 = packed.view(B, S2, nq + 2 * nkv, D).permute(0, 2, 1, 3)
    ref4 = v4.float().cpu().clone()
    pos2 = (torch.arange(start, start + S2) + 1)[None].expand(B, -1)
    q_want, k_want = O.gemma_rope(ref4[:, :nq], ref4[:, nq:nq + nkv], pos2, 10000.0)
    kc2 = torch.zeros(B + 1, nkv, 32, D, dtype=torch.bfloat16, device="cuda")
    vc2 = torch.zeros_like(kc2)
    posd = torch.tensor([start], dtype=torch.int32, device="cuda")
    ops.rope_append(v4, nq, nkv, kc2, vc2, cos, sin, 1, 0, pos_dev=posd)  # host part 1 / 0, the rest from the device scalar
    assert rel_l2(v4[:, :nq].float().cpu(), q_want) <= 4e-3
    assert rel_l2(kc2[:B, :, start:start + S2].float().cpu(), k_want) <= 4e-3
    assert torch.equal(vc2[:B, :, start:start + S2].float().cpu(), ref4[:, nq + nkv:])
    assert float(kc2[B].abs().max()) == 0.0 and float(kc2[:, :, :start].abs().max()) == 0.0 and float(vc2[:, :, start + S2:].abs().max()) == 0.0
    # GeGLU: gelu_tanh(x Wg^T) * (x Wu^T) from ONE GEMM over interleaved gate / up rows
    M, H, inter = 300, 256, 512
    xx, wg, wu = _r((M, H), 6), _r((inter, H), 7, H ** -0.5), _r((inter, H), 8, H ** -0.5)
    w = torch.stack([wg, wu], dim=1).reshape(2 * inter, H)
    got = ops.gemm(xx.bfloat16().cuda(), w.bfloat16().cuda(), act="geglu_tanh")
    ref = O.gelu_tanh(xx @ wg.t()) * (xx @ wu.t())
    assert list(got.shape) == [M, inter] and rel_l2(got.float().cpu(), ref) <= 8e-3
    got1 = ops.gemm(xx[:1].bfloat16().cuda(), w.bfloat16().cuda(), act="geglu_tanh")  # a single decode row
    assert rel_l2(got1.float().cpu(), ref[:1]) <= 8e-3
    # 14 x 14 patches: 588 columns padded to 592 with zeros
    px = torch.rand(2, 3, 28, 28, generator=_gen(9))
    rows = ops.patchify(px.cuda(), (14, 14), torch.bfloat16, pad_to=8)
    assert list(rows.shape) == [8, 592] and float(rows[:, 588:].abs().max()) == 0.0
    want_rows = torch.nn.functional.unfold(px, kernel_size=14, stride=14).transpose(1, 2).reshape(8, 588)
    assert torch.equal(rows[:, :588].float().cpu(), want_rows.bfloat16().float())


def _build(fx):
    from vyomai_b200.models.paligemma import PaliGemmaConfig, PaliGemmaForConditionalGeneration
    m = fx.meta
    cfg = PaliGemmaConfig(vision_config=dict(m["vision"]), text_config=dict(m["text"]), image_token_index=m["image_token_index"],
                          vocab_size=m["text"]["vocab_size"], projection_dim=m["projection_dim"], hidden_size=m["hidden_size"],
                          pad_token_id=m["pad_token_id"])
    model = PaliGemmaForConditionalGeneration(cfg)
    missing, unexpected = model.load_state_dict(fx.sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    return model.cuda().to(torch.bfloat16).eval(), cfg


@pytest.mark.gpu
def test_paligemma_model_matches_the_notebook():
    from vyomai_b200.models.paligemma import StaticCache, paligemma_generate
    fx = load_fixture(NAME)
    m = fx.meta
    model, cfg = _build(fx)
    ids, mask, px = fx.inputs["input_ids"].cuda(), fx.inputs["attention_mask"].cuda(), fx.inputs["pixel_values"].cuda().bfloat16()
    TOLB = 2e-2  # bf16 weights, activations and kv-cache against the fp32 notebook run (tests/test_gpu_models.py TOL[bf16])
    last = model.vision_tower(px).last_hidden_state
    assert rel_l2(last.float().cpu(), fx.outputs["siglip_last_hidden"]) <= TOLB
    assert rel_l2(model.get_image_features(px).float().cpu(), fx.outputs["image_features"]) <= TOLB
    out = model(input_ids=ids, pixel_values=px, attention_mask=mask)
    valid = fx.inputs["attention_mask"].bool()
    assert rel_l2(out.logits.float().cpu()[valid], fx.outputs["prefill_logits"][valid]) <= TOLB
    # cell 30's procedure, teacher-forced on the reference's own ids: logits of every step, cache contents
    row = m["gen_row"]
    L = int(fx.inputs["attention_mask"][row].sum())
    cache = StaticCache(cfg.text_config, batch_size=1, device="cuda", dtype=torch.bfloat16, max_cache_len=m["cache_len"])
    cur, cm = ids[row:row + 1, :L], mask[row:row + 1, :L]
    ref_ids = fx.outputs["generate"]
    for s in range(ref_ids.shape[1]):
        o = model(input_ids=cur, pixel_values=px[row:row + 1], attention_mask=cm, past_key_values=cache, use_cache=True)
        assert rel_l2(o.logits[:, -1].float().cpu(), fx.outputs["gen_step_logits"][:, s]) <= TOLB, s
        cur = ref_ids[:, s:s + 1].cuda()
        cm = torch.cat([cm, torch.ones(1, 1, dtype=cm.dtype, device="cuda")], -1)
    kc = cache.key_cache[0].float().cpu()
    assert rel_l2(kc, fx.outputs["key_cache_l0"]) <= TOLB
    used = L + ref_ids.shape[1] - 1
    assert float(kc[:, :, used:].abs().max()) == 0.0  # slots beyond the written ones stay zero (bit-exact indexing)
    assert int(cache.get_seq_length()) == used
    # free-running greedy generation (batch 1 like the notebook, and the three rows as one right-padded batch)
    got = paligemma_generate(model, ids[row:row + 1, :L], px[row:row + 1], mask[row:row + 1, :L], max_tokens_to_generate=ref_ids.shape[1],
                             max_cache_len=m["cache_len"])
    n_ok = 0
    for s, mg in enumerate(m["generate_margins"]):  # margin rule: ids must agree until the reference's own top-1/top-2 gap is inside bf16 noise
        if mg < 0.2:
            break
        assert int(got[0, s]) == int(ref_ids[0, s]), (s, got.tolist(), ref_ids.tolist())
        n_ok += 1
    print(f"paligemma greedy ids equal to the notebook's for {n_ok}/{ref_ids.shape[1]} steps (margins {[round(x, 3) for x in m['generate_margins']]})")
    # the CUDA-graph path (default, device-side position, split packed-head attention) against the notebook-style eager loop
    eager = paligemma_generate(model, ids[row:row + 1, :L], px[row:row + 1], mask[row:row + 1, :L], max_tokens_to_generate=ref_ids.shape[1],
                               max_cache_len=m["cache_len"], use_graph=False)
    for s, mg in enumerate(m["generate_margins"]):
        if mg < 0.2:
            break
        assert int(eager[0, s]) == int(ref_ids[0, s]) == int(got[0, s])
    gb = paligemma_generate(model, ids, px, mask, max_tokens_to_generate=3, max_cache_len=m["cache_len"])
    ge = paligemma_generate(model, ids, px, mask, max_tokens_to_generate=3, max_cache_len=m["cache_len"], use_graph=False)
    assert list(gb.shape) == [3, 3] and torch.equal(gb[:, 0], ge[:, 0])  # (the prefill token: identical code path)


@pytest.mark.gpu
def test_paligemma_graph_decode_step_matches_eager_step_logits():
    """One replayed step of PaliGemmaDecodeGraph against the eager single-token forward on copies of the same cache: the token
    chosen, the cache row written at the device-side position, and nothing else touched."""
    import copy
    from vyomai_b200 import ops
    from vyomai_b200.models.paligemma import PaliGemmaDecodeGraph, StaticCache
    fx = load_fixture(NAME)
    m = fx.meta
    model, cfg = _build(fx)
    ids, mask, px = fx.inputs["input_ids"].cuda(), fx.inputs["attention_mask"].cuda(), fx.inputs["pixel_values"].cuda().bfloat16()
    B, S0 = ids.shape
    cache = StaticCache(cfg.text_config, batch_size=B, device="cuda", dtype=torch.bfloat16, max_cache_len=m["cache_len"])
    o = model(input_ids=ids, pixel_values=px, attention_mask=mask, past_key_values=cache, use_cache=True, logits_last_only=True)
    first = ops.argmax_rows(o.logits[:, -1])
    cache2 = copy.deepcopy(cache)
    am = torch.cat([mask, torch.ones(B, 1, dtype=mask.dtype, device="cuda")], -1)
    e = model(input_ids=first.view(B, 1), attention_mask=am, past_key_values=cache2, use_cache=True, logits_last_only=True)
    want = ops.argmax_rows(e.logits[:, -1])
    g = PaliGemmaDecodeGraph(model, cache, mask)
    out = torch.empty((B, 2), dtype=torch.long, device="cuda")
    g.run(first, S0, 1, out)
    top2 = e.logits[:, -1].float().topk(2, -1).values
    sure = (top2[:, 0] - top2[:, 1]) > 0.1
    assert torch.equal(out[sure, 0], want[sure])
    for li in range(len(cache.key_cache)):
        assert rel_l2(cache.key_cache[li][:, :, :S0 + 1].float().cpu(), cache2.key_cache[li][:, :, :S0 + 1].float().cpu()) <= 1e-2
        assert rel_l2(cache.value_cache[li][:, :, :S0 + 1].float().cpu(), cache2.value_cache[li][:, :, :S0 + 1].float().cpu()) <= 1e-2
        assert float(cache.key_cache[li][:, :, S0 + 1:].abs().max()) == 0.0


@pytest.mark.gpu
def test_paligemma_real_width_layers_vs_oracle():
    """The REAL layer widths of config 5 (SigLIP 1152 / 16 heads of 72 / MLP 4304 / 14-pixel patches of a 224 image; Gemma 2048 /
    8 query heads + 1 kv head of 256 / GeGLU 16384) with two layers each and a 32 000-token vocabulary, so that the kernels run
    the shapes of the benchmark (K = 16384 small-batch GEMM, 1-row gated GEMM, 256-wide packed-head decode tile split over keys,
    592-column patch rows) — against the CPU oracle on the same seeded bf16-representable weights: image features, the last-position
    prefill logits and three teacher-forced cached steps (eager and CUDA-graph), greedy choice where the oracle's margin allows."""
    from vyomai_b200 import ops
    from vyomai_b200.models.paligemma import PaliGemmaConfig, PaliGemmaDecodeGraph, PaliGemmaForConditionalGeneration, StaticCache
    vis = dict(hidden_size=1152, intermediate_size=4304, num_hidden_layers=2, num_attention_heads=16, num_channels=3, image_size=224,
               patch_size=14, layer_norm_eps=1e-6)
    txt = dict(vocab_size=32000, hidden_size=2048, intermediate_size=16384, num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=1,
               head_dim=256, max_position_embeddings=512, rms_norm_eps=1e-6, rope_theta=10000.0)
    meta = {"hidden_size": 2048, "image_token_index": 31999, "vision": vis, "text": txt}
    cfg = PaliGemmaConfig(vision_config=dict(vis), text_config=dict(txt), image_token_index=31999, vocab_size=32000, projection_dim=2048,
                          hidden_size=2048, pad_token_id=0)
    torch.manual_seed(5)
    model = PaliGemmaForConditionalGeneration(cfg)
    g = torch.Generator().manual_seed(6)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("layernorm.weight") or n.endswith("model.norm.weight"):
                p.copy_((0.1 * torch.randn(p.shape, generator=g)).bfloat16().float())
            elif p.dim() >= 2:
                p.copy_((0.02 * torch.randn(p.shape, generator=g)).bfloat16().float())
            else:
                p.copy_((0.02 * torch.randn(p.shape, generator=g) + (1.0 if "layer_norm" in n or "post_layernorm" in n else 0.0) * (".weight" in n)).bfloat16().float())
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().to(torch.bfloat16).eval()
    B, n_img, n_txt = 2, 256, 6
    ids = torch.cat([torch.full((B, n_img), 31999, dtype=torch.long), torch.randint(1, 31000, (B, n_txt), generator=g)], dim=1)
    mask = torch.ones(B, n_img + n_txt, dtype=torch.long)
    mask[1, -2:] = 0  # a right-padded row
    ids[1, -2:] = 0
    px = torch.rand(B, 3, 224, 224, generator=g).bfloat16().float()
    TOLB = 2e-2
    with torch.no_grad():
        feats = O.siglip_forward(sd, "vision_tower.vision_model.", px, 14, 2, 16, 1e-6)
    got_f = model.vision_tower(px.cuda().bfloat16()).last_hidden_state
    assert rel_l2(got_f.float().cpu(), feats) <= TOLB
    S0, CL = n_img + n_txt, 320
    ocache = ([torch.zeros(B, 1, CL, 256) for _ in range(2)], [torch.zeros(B, 1, CL, 256) for _ in range(2)])
    with torch.no_grad():
        ref0 = O.paligemma_forward(sd, meta, ids, px, mask, cache=ocache, seen=0, cache_len=CL)[:, -1]
    cache = StaticCache(cfg.text_config, batch_size=B, device="cuda", dtype=torch.bfloat16, max_cache_len=CL)
    out = model(input_ids=ids.cuda(), pixel_values=px.cuda().bfloat16(), attention_mask=mask.cuda(), past_key_values=cache, use_cache=True,
                logits_last_only=True)
    assert rel_l2(out.logits[:, -1].float().cpu()[:1], ref0[:1]) <= TOLB  # (row 1 ends in padding: its last position is a pad query)
    graph = PaliGemmaDecodeGraph(model, cache, mask.cuda())
    cur = ref0.argmax(-1)  # teacher forcing on the oracle's own choices
    am = mask
    toks = torch.empty((B, 1), dtype=torch.long, device="cuda")
    for s in range(3):
        am = torch.cat([am, torch.ones(B, 1, dtype=am.dtype)], -1)
        with torch.no_grad():
            ref = O.paligemma_forward(sd, meta, cur.view(B, 1), None, am, cache=ocache, seen=S0 + s, cache_len=CL)[:, -1]
        if s < 2:  # eager cached step
            o = model(input_ids=cur.view(B, 1).cuda(), attention_mask=am.cuda(), past_key_values=cache, use_cache=True, logits_last_only=True)
            assert rel_l2(o.logits[:, -1].float().cpu(), ref) <= TOLB, s
        else:      # the same step as one CUDA-graph replay: compare the choice it makes where the oracle is sure
            graph.run(cur.cuda(), S0 + s, 1, toks)
            top2 = ref.topk(2, -1).values
            sure = (top2[:, 0] - top2[:, 1]) > 0.15
            assert torch.equal(toks[:, 0].cpu()[sure], ref.argmax(-1)[sure])
        cur = ref.argmax(-1)
    for li in range(2):
        assert rel_l2(cache.key_cache[li][:, :, :S0 + 3].float().cpu(), ocache[0][li][:, :, :S0 + 3]) <= TOLB
        assert float(cache.key_cache[li][:, :, S0 + 3:].abs().max()) == 0.0
