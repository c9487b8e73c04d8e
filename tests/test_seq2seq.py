"""Cross-attention and the seq2seq model (SURVEY.md §8(a) row `EncoderDecoderAttention[Gqa]`; reference:
layers/attention.py:382-573, models/encoder_decoder.py, generation_utils.py:54-125).

CPU: the oracle against the REAL reference's outputs (tests/golden/seq2seq_*.npz, written by make_golden_seq2seq.py).
-m gpu: the sm_100a path (EncoderDecoderModel / generate_seq2seq from this package) against those fixtures — logits,
encoder states, a sample of gradients (incl. every cross-attention projection), the three cache modes and the greedy ids
(margin rule). Tolerances: those of tests/test_gpu_models.py / tests/test_gpu_train.py.
"""
import pytest
import torch

from oracle import vyom_oracle as O
from tests.conftest import load_fixture, rel_l2

NAMES = ["seq2seq_rope_gqa", "seq2seq_absolute_mha"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_seq2seq_matches_reference(name):
    fx = load_fixture(name)
    cfg, m = fx.cfg(), fx.meta
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in fx.sd.items()}
    logits, enc = O.seq2seq_forward(sd, cfg, cfg, fx.inputs["input_ids"], fx.inputs["attention_mask"], fx.inputs["decoder_input_ids"],
                                    fx.inputs["decoder_attention_mask"], m["pos"], m["attn"], m["pos"], m["attn"])
    assert rel_l2(logits, fx.outputs["logits"]) <= 2e-6
    assert rel_l2(enc, fx.outputs["key_value_states"]) <= 2e-6
    (logits * fx.inputs["cotangent"]).sum().backward()
    n = 0
    for k, g in fx.outputs.items():
        if not k.startswith("grad::"):
            continue
        key = k[6:]
        if key == "lm_head.vocab.bias":
            key = "lm_head.bias"
        got = sd[key].grad
        assert got is not None, k
        if float(g.abs().max()) < 1e-6:
            assert float(got.abs().max()) < 1e-6, k
            continue
        assert rel_l2(got, g, floor=1e-4) <= 2e-5, k
        n += 1
    assert n >= 20
    for use_cache in (False, True):
        ids = O.seq2seq_generate(fx.sd, cfg, cfg, fx.outputs["gen_encoder_output"], fx.inputs["attention_mask"][:1], fx.inputs["gen_start"], 6,
                                 m["pos"], m["attn"], use_cache=use_cache)
        assert torch.equal(ids, fx.outputs["generate"])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", NAMES)
def test_seq2seq_forward_backward_generate_match_reference(name, dtype):
    from tests.test_gpu_models import MARGIN, TOL, _cfg_obj, _ids_match_up_to_ambiguity, _load
    from tests.test_gpu_train import GTOL
    from vyomai_b200 import DynamicCache, EncoderDecoderModel, StaticCache, generate_seq2seq
    fx = load_fixture(name)
    m = fx.meta
    cfg = _cfg_obj(m)
    model = _load(EncoderDecoderModel(cfg, cfg, encoder_pos_embedding_type=m["pos"], encoder_attention_type=m["attn"],
                                      decoder_pos_embedding_type=m["pos"], decoder_attention_type=m["attn"]), fx.sd, dtype).train()
    ids, mask = fx.inputs["input_ids"].cuda(), fx.inputs["attention_mask"].cuda()
    dids, dmask = fx.inputs["decoder_input_ids"].cuda(), fx.inputs["decoder_attention_mask"].cuda()
    out = model(input_ids=ids, attention_mask=mask, decoder_input_ids=dids, decoder_attention_mask=dmask)
    assert list(out.logits.shape) == list(fx.outputs["logits"].shape)
    assert rel_l2(out.logits.float().cpu(), fx.outputs["logits"]) <= TOL[dtype]
    assert rel_l2(out.key_value_states.float().cpu(), fx.outputs["key_value_states"]) <= TOL[dtype]
    (out.logits.float() * fx.inputs["cotangent"].cuda()).sum().backward()
    params = dict(model.named_parameters())
    n, worst = 0, (0.0, None)
    for k, g in fx.outputs.items():
        if not k.startswith("grad::"):
            continue
        key = k[6:]
        if key == "lm_head.vocab.bias":
            key = "lm_head.bias"
        got = params[key].grad
        assert got is not None, k
        got = got.float().cpu()
        if float(g.abs().max()) < 1e-6:
            assert float(got.abs().max()) < 2e-3, k
            continue
        r = rel_l2(got, g, floor=1e-4)
        worst = max(worst, (r, k))
        assert r <= GTOL[dtype], (k, r)
        n += 1
    assert n >= 20
    print(f"{name} {dtype}: {n} gradient tensors within {GTOL[dtype]}, worst {worst}")

    model.eval()
    with torch.no_grad():
        enc = model.get_encoder_output(ids[:1], mask[:1]).logits
        assert rel_l2(enc.float().cpu(), fx.outputs["gen_encoder_output"]) <= TOL[dtype]
        start = fx.inputs["gen_start"].cuda()
        g0 = generate_seq2seq(model, enc, mask[:1], start, max_new_tokens=6, use_cache=False)
        model._setup_cache(cfg, cls=DynamicCache)
        g1 = generate_seq2seq(model, enc, mask[:1], start, max_new_tokens=6, use_cache=True)
        model._clean_cache()
        model._setup_cache(cfg, cls=StaticCache)
        g2 = generate_seq2seq(model, enc, mask[:1], start, max_new_tokens=6, use_cache=True)
        model._clean_cache()
    assert torch.equal(g1, g2)  # both caches run the same kernels on the same values
    ref = fx.outputs["generate"]
    assert list(g0.shape) == list(ref.shape)
    margins = []
    for cur in range(1, ref.shape[1]):
        lg, _ = O.seq2seq_forward(fx.sd, fx.cfg(), fx.cfg(), None, fx.inputs["attention_mask"][:1], ref[:, :cur], None, dec_pos=m["pos"],
                                  dec_attn=m["attn"], encoder_output=fx.outputs["gen_encoder_output"])
        top2 = lg[:, -1].topk(2, dim=-1).values
        margins.append(float((top2[:, 0] - top2[:, 1]).min()))
    for g in (g0, g1):
        n_ok = _ids_match_up_to_ambiguity(g.cpu(), ref, margins, 1, MARGIN[dtype])
    print(f"{name} {dtype}: greedy ids bit-exact for {n_ok}/{len(margins)} steps; identical to reference: {torch.equal(g0.cpu(), ref)}")


@pytest.mark.gpu
def test_adapters_run_on_the_gemm_kernels():
    """LoraLinear / DoraLinear (reference: layers/adapters.py) against their defining formulas in fp32 torch on the CPU."""
    import torch.nn as nn
    from vyomai_b200 import DoraLinear, LoraLinear
    torch.manual_seed(0)
    lin = nn.Linear(768, 3072)
    x = torch.rand(4, 32, 768)
    lora = LoraLinear(lin, rank=32, alpha=2)
    with torch.no_grad():
        lora.lora_b.normal_(0, 0.05)
    ref = lin(x) + 2 * torch.nn.functional.linear(torch.nn.functional.linear(x, lora.lora_a), lora.lora_b)
    got = lora.cuda()(x.cuda())
    assert list(got.shape) == [4, 32, 3072] and rel_l2(got.float().cpu(), ref) <= 6e-3
    got.float().pow(2).sum().backward()
    assert lora.lora_a.grad is not None and bool(torch.isfinite(lora.lora_a.grad).all()) and float(lora.lora_a.grad.abs().max()) > 0
    lin2 = nn.Linear(768, 768)
    dora = DoraLinear(lin2, rank=16)
    with torch.no_grad():
        dora.dora_b.normal_(0, 0.05)
    adapted = lin2.weight + dora.dora_a @ dora.dora_b
    refd = torch.nn.functional.linear(x, dora.dora_m * adapted / adapted.norm(p=2, dim=0, keepdim=True), lin2.bias)
    gotd = dora.cuda()(x.cuda())
    assert rel_l2(gotd.float().cpu(), refd) <= 6e-3
