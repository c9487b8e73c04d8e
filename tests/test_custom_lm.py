"""The RMSNorm / SwiGLU / RoPE decoder of VyomAI/models/custom_transformer.py (`ModelForCausalLM`).

CPU: the oracle (custom_lm_forward) against the REAL reference class's logits and greedy ids (tests/golden/custom_lm_d64.npz,
custom_lm_d128.npz, written by make_golden_custom.py). -m gpu: vyomai_b200's ModelForCausalLM (bf16) against the same fixtures —
logits of the right-padded batch, greedy ids through the static cache + CUDA-graph decode (margin rule), eager cached steps
against the uncached forward; head_dim 64 exercises the tensor-memory attention kernel, 128 the mma.sync one."""
import pytest
import torch

from oracle import vyom_oracle as O
from tests.conftest import load_fixture, rel_l2

NAMES = ["custom_lm_d64", "custom_lm_d128"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_custom_lm_matches_reference(name):
    fx = load_fixture(name)
    m = fx.meta
    lg = O.custom_lm_forward(fx.sd, m, fx.inputs["input_ids"], fx.inputs["attention_mask"])
    assert rel_l2(lg, fx.outputs["logits"]) <= 2e-6
    cur = fx.inputs["prompt"]
    for _ in range(fx.outputs["generate"].shape[1] - cur.shape[1]):
        nxt = O.custom_lm_forward(fx.sd, m, cur, None)[:, -1].argmax(-1, keepdim=True)
        cur = torch.cat([cur, nxt], dim=1)
    assert torch.equal(cur, fx.outputs["generate"])


def _build(fx):
    from vyomai_b200.models.custom_transformer import Config, ModelForCausalLM
    m = fx.meta
    cfg = Config(**{k: m[k] for k in ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                                      "num_key_value_heads", "max_position_embeddings", "rms_norm_eps", "rope_theta", "tie_word_embeddings",
                                      "pad_token_id")})
    model = ModelForCausalLM(cfg)
    sd = dict(fx.sd)
    sd["layers.0.mlp.up_proj.weight"] = torch.zeros(1)  # a dead inherited entry of the reference's state_dict: accepted and dropped
    missing, unexpected = model.load_state_dict(sd)
    assert not missing and not unexpected
    return model.cuda().to(torch.bfloat16).eval(), cfg


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_custom_lm_matches_reference(name):
    from vyomai_b200.models.paligemma import StaticCache
    fx = load_fixture(name)
    m = fx.meta
    model, cfg = _build(fx)
    ids, mask = fx.inputs["input_ids"].cuda(), fx.inputs["attention_mask"].cuda()
    TOLB = 2e-2  # bf16 model against the fp32 reference run (tests/test_gpu_models.py TOL[bf16])
    out = model(input_ids=ids, attention_mask=mask)
    valid = fx.inputs["attention_mask"].bool()
    assert list(out.logits.shape) == list(fx.outputs["logits"].shape)
    assert rel_l2(out.logits.float().cpu()[valid], fx.outputs["logits"][valid]) <= TOLB
    # cached: prefill of the first 8 columns of row 0 (unpadded), then the next 3 tokens one at a time == the uncached forward
    cache = StaticCache(cfg, batch_size=1, device="cuda", dtype=torch.bfloat16, max_cache_len=16)
    pre = model(input_ids=ids[:1, :8], attention_mask=mask[:1, :8], past_key_values=cache, use_cache=True)
    assert rel_l2(pre.logits.float().cpu(), fx.outputs["logits"][:1, :8]) <= TOLB
    for t in range(8, 11):
        step = model(input_ids=ids[:1, t:t + 1], attention_mask=mask[:1, :t + 1], past_key_values=cache, use_cache=True)
        assert rel_l2(step.logits[:, -1].float().cpu(), fx.outputs["logits"][:1, t]) <= TOLB, t
    assert int(cache.get_seq_length()) == 11 and float(cache.key_cache[0][:, :, 11:].abs().max()) == 0.0
    # greedy ids: static cache + one CUDA-graph replay per token, margin rule against the reference's own margins
    ref = fx.outputs["generate"]
    prompt = fx.inputs["prompt"].cuda()
    got = model.generate_greedy(prompt, max_new_tokens=ref.shape[1] - prompt.shape[1])
    assert list(got.shape) == list(ref.shape) and torch.equal(got[:, :prompt.shape[1]].cpu(), fx.inputs["prompt"])
    n_ok = 0
    for s, mg in enumerate(m["generate_margins"]):
        if mg < 0.1:
            break
        assert int(got[0, prompt.shape[1] + s]) == int(ref[0, prompt.shape[1] + s]), (s, got.tolist(), ref.tolist())
        n_ok += 1
    assert n_ok >= 3
    print(f"{name}: greedy ids equal to the reference's for {n_ok}/{len(m['generate_margins'])} steps")
    gb = model.generate_greedy(ids, mask, max_new_tokens=4)  # right-padded batch through the same path
    assert list(gb.shape) == [3, ids.shape[1] + 4]
