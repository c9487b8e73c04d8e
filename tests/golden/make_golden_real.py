"""Golden fixtures at REAL width (the shapes the bench's kernels actually run: H = 768, 12 q-heads / 4 kv-heads,
S = 128 encoder rows, a 200-token causal prefill), produced by running the REAL reference (/root/reference/VyomAI,
imported, nothing copied) on the CPU in fp32.

    python tests/golden/make_golden_real.py      # needs /root/reference; run in the build container

A 768-wide model does not fit a "small fixture", so the weights are NOT stored: they come from `seeded_state_dict`
in tests/conftest.py (a torch.Generator recipe over the sorted state_dict keys, rounded to bf16-representable values), which the
tests call again to rebuild exactly the same tensors. What is stored: the inputs, a strided sample of the reference's
outputs / gradients, whole 1-D gradients, every gradient's norm, greedy ids and the reference's top-1/top-2 margins.
"""
import io
import json
import os
import sys
from contextlib import redirect_stdout
from dataclasses import dataclass

import numpy as np
import torch

REF = os.environ.get("VYOM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.conftest import seeded_state_dict  # noqa: E402  (the weight recipe shared with the tests)


@dataclass
class RealCfg:
    hidden_size: int = 768
    num_attention_heads: int = 12
    num_key_value_heads: int = 4
    max_position_embeddings: int = 514
    num_hidden_layers: int = 1
    vocab_size: int = 4096
    hidden_dropout_prob: float = 0.1
    initializer_range: float = 0.02
    intermediate_size: int = 3072
    layer_norm_eps: float = 1e-05
    hidden_act: str = "gelu"
    pad_token_id: int = 1
    eos_token_id: int = 2


def _float_shapes(model) -> dict:
    return {k: tuple(v.shape) for k, v in model.state_dict().items() if v.dtype.is_floating_point}


def _load_seeded(model, seed):
    sd = seeded_state_dict(_float_shapes(model), seed)
    if "lm_head.bias" in sd:
        sd["lm_head.decoder.bias"] = sd["lm_head.bias"]  # one Parameter under two keys
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in m for m in missing), (missing, unexpected)


def right_padded(batch, seqlen, vocab, seed, min_len):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, vocab, (batch, seqlen), generator=g)
    lens = torch.randint(min_len, seqlen + 1, (batch,), generator=g)
    lens[0] = seqlen
    mask = (torch.arange(seqlen)[None, :] < lens[:, None]).long()
    return torch.where(mask.bool(), ids, torch.ones_like(ids)), mask


def main():
    sys.path.insert(0, REF)
    import VyomAI
    from VyomAI import DecoderModel, EncoderModel
    from VyomAI.layers.kv_cache import StaticCacheOne
    assert os.path.realpath(VyomAI.__file__).startswith(os.path.realpath(REF)), VyomAI.__file__
    quiet = io.StringIO()
    cfg = RealCfg()
    meta = {f: getattr(cfg, f) for f in cfg.__dataclass_fields__}

    # ---- C1 shape: EncoderModel RoPE + GQA, 8 x 128 right-padded, forward + backward (random cotangent) ----
    with redirect_stdout(quiet):
        enc = EncoderModel(cfg, pos_embedding_type="rope", attention_type="gqa").eval()
    _load_seeded(enc, 2025)
    ids, mask = right_padded(8, 128, cfg.vocab_size, 11, 16)
    out = enc(ids, mask).logits
    cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(12)) * mask[..., None]
    (out * cot).sum().backward()
    blob = {"in::input_ids": ids.numpy(), "in::attention_mask": mask.numpy(),
            "out::logits_rows": out[:, ::8].detach().numpy()}
    norms = {}
    for k, p in enc.named_parameters():
        if p.grad is None:
            continue
        norms[k] = float(p.grad.norm())
        if k == "word_embeddings.weight":
            rows = ids.unique()[:64]
            blob["in::emb_rows"] = rows.numpy()
            blob["out::grad::" + k] = p.grad[rows].numpy()
        elif p.dim() == 1:
            blob["out::grad::" + k] = p.grad.numpy()
        else:
            blob["out::grad::" + k] = p.grad[:64, :64].contiguous().numpy()
    m1 = dict(meta, pos="rope", attn="gqa", weight_seed=2025, cotangent_seed=12, grad_norms=norms)
    blob["meta"] = np.frombuffer(json.dumps(m1).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "encoder_real_rope_gqa.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")

    # ---- C3 shape (scaled): DecoderModel RoPE + GQA, 200-token causal prefill (two 128-row query tiles) through the
    # static cache, 3 teacher-forced decode steps, greedy generate with the reference's logit margins ----
    with redirect_stdout(quiet):
        dec = DecoderModel(cfg, pos_embedding_type="rope", attention_type="gqa").eval()
    _load_seeded(dec, 2026)
    B, P, N = 4, 200, 8
    pids = torch.randint(3, cfg.vocab_size, (B, P), generator=torch.Generator().manual_seed(21))
    am = torch.ones(B, P, dtype=torch.long)
    with torch.no_grad():
        kv = StaticCacheOne(cfg, max_cache_len=P + N, batch_size=B)
        o0 = dec(pids, am, use_cache=True, kv_cache=kv, start_pos=0)
        last = o0.logits[:, -1]
        steps, toks, margins = [], [], []
        nxt = last.argmax(-1, keepdim=True)
        t2 = last.topk(2, dim=-1).values
        margins.append(float((t2[:, 0] - t2[:, 1]).min()))
        for t in range(3):
            toks.append(nxt)
            am = torch.cat([am, torch.ones(B, 1, dtype=torch.long)], dim=-1)
            ot = dec(nxt, am, use_cache=True, kv_cache=kv, start_pos=P + t)
            steps.append(ot.logits)
            nxt = ot.logits[:, -1].argmax(-1, keepdim=True)
        k0 = kv.key_cache[0][:, :, ::16].clone()   # every 16th slot of layer 0's keys (rotated), incl. untouched zeros
        v0 = kv.value_cache[0][:, :, ::16].clone()
        gen = dec.generate(pids, torch.ones(B, P, dtype=torch.long), max_len=N, use_cache=True, use_static_cache=True)
        # the reference's margin at every generated position, teacher-forced on its own ids without a cache
        gmargins = []
        for cur in range(P, P + N):
            lg = dec(gen[:, :cur], torch.ones(B, cur, dtype=torch.long)).logits[:, -1]
            t2 = lg.topk(2, dim=-1).values
            gmargins.append(float((t2[:, 0] - t2[:, 1]).min()))
            assert torch.equal(lg.argmax(-1), gen[:, cur]) or gmargins[-1] < 1e-4
    blob = {"in::prompt": pids.numpy(), "in::decode_tokens": torch.cat(toks, 1).numpy(),
            "out::prefill_last_logits": last.numpy(), "out::prefill_hidden_rows": o0.hidden_state[:, ::25].numpy(),
            "out::decode_logits": torch.cat(steps, 1).numpy(), "out::key_cache_l0_s16": k0.numpy(),
            "out::value_cache_l0_s16": v0.numpy(), "out::generate": gen.numpy()}
    m2 = dict(meta, pos="rope", attn="gqa", weight_seed=2026, generate_margins=gmargins, new_tokens=N)
    blob["meta"] = np.frombuffer(json.dumps(m2).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "decoder_real_rope_gqa.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
