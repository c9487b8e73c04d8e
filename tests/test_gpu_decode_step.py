"""-m gpu: the one-launch decode step (vy_decode_step, csrc/decode_step.cu) — embedding, all layers, LM head, argmax, token
write-back in ONE persistent kernel — against (a) the REAL reference's decode logits / cache contents / greedy ids at real
width (tests/golden/decoder_real_rope_gqa.npz), (b) the CPU oracle on config 3's shape family (B = 32, L = 4, MHA and GQA),
(c) the per-op kernel path it replaces.

Tolerances: bf16 module tolerance 2e-2 on logits (tests/test_gpu_models.py), cache slots 2e-2, which slots are written and
the position counter: bit-exact; greedy ids by the margin rule.
"""
import io
from contextlib import redirect_stdout
from dataclasses import make_dataclass

import pytest
import torch

from oracle import vyom_oracle as O
from tests.conftest import load_fixture, real_state_dict, rel_l2
from tests.test_gpu_models import MARGIN
from tests.test_gpu_real_shapes import _build

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _enable_fused_step(monkeypatch):
    """The one-launch step is opt-in (VY_DECODE_FUSED=1): these tests switch it on."""
    from vyomai_b200 import decode_step
    monkeypatch.setattr(decode_step, "ENABLED", True)


def _fused(model, cache, B, tok, pos, tokens, logits):
    from vyomai_b200 import decode_step
    assert decode_step.supported(model, cache, B)
    return decode_step.FusedDecodeStep(model, cache, B, tok, pos, tokens, pos_bound=cache.key_cache[0].shape[2] - 1, logits=logits)


def test_fused_step_matches_reference_decode_logits_and_cache():
    from vyomai_b200 import DecoderModel, StaticCacheOne
    fx = load_fixture("decoder_real_rope_gqa")
    m = fx.meta
    model = _build(DecoderModel, fx, torch.bfloat16).eval()
    cfg = model.config
    prompt = fx.inputs["prompt"].cuda()
    B, P = prompt.shape
    N = m["new_tokens"]
    V = cfg.vocab_size
    with torch.no_grad():
        kv = StaticCacheOne(cfg, max_cache_len=P + N, batch_size=B, dtype=torch.bfloat16)
        model(prompt, torch.ones(B, P, dtype=torch.long, device="cuda"), use_cache=True, kv_cache=kv, start_pos=0)
        tok = torch.zeros(B, dtype=torch.long, device="cuda")
        pos = torch.full((1,), P, dtype=torch.int32, device="cuda")
        tokens = torch.zeros(B, P + N, dtype=torch.long, device="cuda")
        logits = torch.zeros(B, V, dtype=torch.bfloat16, device="cuda")
        step = _fused(model, kv, B, tok, pos, tokens, logits)
        got = []
        for t in range(3):
            tok.copy_(fx.inputs["decode_tokens"][:, t].cuda())  # teacher-forced on the reference's own tokens
            step.launch()
            torch.cuda.synchronize()
            got.append(logits.float().cpu().clone())
            assert int(pos) == P + t + 1                                      # position counter
            assert torch.equal(tokens[:, P + t + 1].cpu(), tok.cpu())         # token write-back
        step.check()
    ref = fx.outputs["decode_logits"]
    assert rel_l2(torch.stack(got, 1), ref) <= 2e-2
    k0 = kv.key_cache[0][:, :, ::16].float().cpu()
    assert torch.equal(k0 == 0, fx.outputs["key_cache_l0_s16"] == 0)  # exactly the reference's slots were written
    assert rel_l2(k0, fx.outputs["key_cache_l0_s16"]) <= 2e-2
    assert rel_l2(kv.value_cache[0][:, :, ::16].float().cpu(), fx.outputs["value_cache_l0_s16"]) <= 2e-2
    # the token the kernel picked = first index of its own logits' row maximum (torch.topk(k=1) rule)
    last = got[-1]
    assert torch.equal(tok.cpu(), torch.topk(last, 1, dim=-1)[1].reshape(-1))


Cfg3 = make_dataclass("Cfg3", [("hidden_size", int, 768), ("num_attention_heads", int, 12), ("max_position_embeddings", int, 1024),
                               ("num_hidden_layers", int, 4), ("vocab_size", int, 50265), ("hidden_dropout_prob", float, 0.0),
                               ("layer_norm_eps", float, 1e-5), ("hidden_act", str, "gelu"), ("pad_token_id", int, 1),
                               ("eos_token_id", int, 2)])
Cfg3Gqa = make_dataclass("Cfg3Gqa", [("num_key_value_heads", int, 4)], bases=(Cfg3,))


@pytest.mark.parametrize("attn,B", [("gqa", 32), (None, 32), ("gqa", 5), (None, 1)])
def test_fused_step_vs_oracle_and_per_op_path_config3(attn, B):
    """BASELINE config 3's decode shape (L 4, H 768, V 50265, batch 32; also ragged batches): after an eager prefill, three
    fused steps against the CPU oracle's cached decode and against the per-op kernels on a twin cache."""
    from vyomai_b200 import DecoderModel, StaticCacheOne
    cfg = Cfg3Gqa() if attn == "gqa" else Cfg3()
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        model = DecoderModel(cfg, "rope", attn)
    sd = real_state_dict(model, 77)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().to(torch.bfloat16).eval()
    P, N = 70, 4
    V = cfg.vocab_size
    ids = torch.randint(3, V, (B, P), generator=torch.Generator().manual_seed(5))
    hkv = 4 if attn == "gqa" else 12
    ocfg = O.Cfg(768, 12, 4 if attn == "gqa" else None, 1024, 4, V, 1e-5, "gelu")
    ocache = O.StaticCacheOneOracle(4, B, hkv, P + N, 64)
    with torch.no_grad():
        _, lg = O.decoder_forward(sd, ocfg, ids, torch.ones(B, P, dtype=torch.long), "rope", attn, ocache, 0)
        kv_a = StaticCacheOne(cfg, max_cache_len=P + N, batch_size=B, dtype=torch.bfloat16)
        kv_b = StaticCacheOne(cfg, max_cache_len=P + N, batch_size=B, dtype=torch.bfloat16)
        for kv in (kv_a, kv_b):
            o0 = model(ids.cuda(), torch.ones(B, P, dtype=torch.long, device="cuda"), use_cache=True, kv_cache=kv, start_pos=0,
                       _logits_last_only=True)
        nxt = lg[:, -1].argmax(-1)
        tok = nxt.clone().cuda()
        pos = torch.full((1,), P, dtype=torch.int32, device="cuda")
        tokens = torch.zeros(B, P + N, dtype=torch.long, device="cuda")
        logits = torch.zeros(B, (V + 7) // 8 * 8, dtype=torch.bfloat16, device="cuda")[:, :V]
        step = _fused(model, kv_a, B, tok, pos, tokens, logits)
        for t in range(3):
            _, ref = O.decoder_forward(sd, ocfg, nxt[:, None], None, "rope", attn, ocache, P + t)   # oracle, fp32
            tok.copy_(nxt.cuda())
            step.launch()
            eager = model(nxt[:, None].cuda(), None, use_cache=True, kv_cache=kv_b, start_pos=P + t, _logits_last_only=True).logits[:, -1]
            torch.cuda.synchronize()
            assert rel_l2(logits.float().cpu(), ref[:, -1]) <= 2e-2, t
            assert rel_l2(logits.float().cpu(), eager.float().cpu()) <= 2e-2, t
            top2 = ref[:, -1].topk(2, dim=-1)
            clear = (top2.values[:, 0] - top2.values[:, 1]) > MARGIN[torch.bfloat16]
            assert torch.equal(tok.cpu()[clear], top2.indices[:, 0][clear]), t      # greedy ids wherever the margin is clear
            assert torch.equal(tok.cpu(), torch.topk(logits.float().cpu(), 1, dim=-1)[1].reshape(-1))  # first-index argmax of its own logits
            nxt = ref[:, -1].argmax(-1)
        step.check()
        for li in range(4):
            a, b = kv_a.key_cache[li].float().cpu(), kv_b.key_cache[li].float().cpu()
            assert torch.equal((a != 0).any(-1), (b != 0).any(-1))  # the same cache slots were written
            assert rel_l2(a, b) <= 2e-2 and rel_l2(a[:, :, :P + 3], ocache.key_cache[li][:, :, :P + 3]) <= 2e-2
            assert rel_l2(kv_a.value_cache[li].float().cpu(), kv_b.value_cache[li].float().cpu()) <= 2e-2


def test_generate_uses_the_fused_step_and_follows_the_python_loop():
    """DecoderModel.generate(static cache) replays the one-kernel step; its ids must equal the reference-style Python loop's
    wherever that loop's own top-1 / top-2 margin is clear (the two paths round differently, so a near-tie may flip)."""
    from vyomai_b200 import DecoderModel, decode_step
    fx = load_fixture("decoder_real_rope_gqa")
    model = _build(DecoderModel, fx, torch.bfloat16).eval()
    prompt = fx.inputs["prompt"].cuda()
    B, P = prompt.shape
    mask = torch.ones(B, P, dtype=torch.long, device="cuda")
    N = 8
    try:
        DecoderModel.use_decode_graph = False
        loop = model.generate(prompt, mask, max_len=N, use_cache=True, use_static_cache=True).cpu()
    finally:
        DecoderModel.use_decode_graph = True
    got = model.generate(prompt, mask, max_len=N, use_cache=True, use_static_cache=True).cpu()
    assert model._decode_graph is not None and model._decode_graph.fused is not None and decode_step.ENABLED
    margins = fx.meta["generate_margins"]
    for i, mg in enumerate(margins[:N]):
        if mg <= MARGIN[torch.bfloat16]:
            break
        assert torch.equal(got[:, P + i], loop[:, P + i]) and torch.equal(got[:, P + i], fx.outputs["generate"][:, P + i]), i
    # a second call reuses the captured graph and is deterministic
    again = model.generate(prompt, mask, max_len=N, use_cache=True, use_static_cache=True).cpu()
    assert torch.equal(got, again)
