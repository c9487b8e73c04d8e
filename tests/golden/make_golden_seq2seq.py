"""Golden fixture for the seq2seq path (cross-attention, Seq2SeqDecoderLayer, EncoderDecoderModel, generate_seq2seq), produced
by running the REAL reference (/root/reference/VyomAI, imported) on the CPU in fp32:

    python tests/golden/make_golden_seq2seq.py

Same conventions as make_golden.py: small widths, weights rounded to bf16-representable values and stored as bf16 bits."""
import io
import os
import sys
from contextlib import redirect_stdout

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import IDS, MASK, REF, TextCfg, TextCfgGqa, cfg_meta, fold, pack, round_weights_, save  # noqa: E402


def main():
    sys.path.insert(0, REF)
    import VyomAI
    from VyomAI import DynamicCache, EncoderDecoderModel, StaticCache, generate_seq2seq
    assert os.path.realpath(VyomAI.__file__).startswith(os.path.realpath(REF)), VyomAI.__file__
    quiet = io.StringIO()
    for pos, attn, cfgcls in (("rope", "gqa", TextCfgGqa), ("absolute", None, TextCfg)):
        torch.manual_seed(31337)
        cfg = cfgcls()
        with redirect_stdout(quiet):
            model = EncoderDecoderModel(cfg, cfg, encoder_pos_embedding_type=pos, encoder_attention_type=attn,
                                        decoder_pos_embedding_type=pos, decoder_attention_type=attn).eval()
        round_weights_(model)
        ids = fold(IDS, cfg.vocab_size)
        dec_ids = ids[:, :11].clone()          # decoder sequence shorter than the encoder's: Sq != Skv in cross-attention
        dec_mask = MASK[:, :11].clone()
        out = model(input_ids=ids, attention_mask=MASK, decoder_input_ids=dec_ids, decoder_attention_mask=dec_mask)
        g = torch.Generator().manual_seed(9)
        cot = torch.randn(out.logits.shape, generator=g) * dec_mask[..., None]
        (out.logits * cot).sum().backward()
        keep = ("decoder.all_layer.0.cross_attention.query.weight", "decoder.all_layer.1.cross_attention.key.weight",
                "decoder.all_layer.0.cross_attention.value.weight", "decoder.all_layer.1.cross_attention.out.dense.weight",
                "decoder.all_layer.0.attention.query.weight", "encoder.all_layer.1.feed_forward.out.weight", "lm_head.dense.weight")
        grads = {"grad::" + k: p.grad for k, p in model.named_parameters() if p.grad is not None and (p.dim() == 1 or k in keep)}
        with torch.no_grad():
            enc = model.get_encoder_output(ids[:1], MASK[:1]).logits
            start = torch.tensor([[0]])
            g_nc = generate_seq2seq(model, enc, MASK[:1], start, max_new_tokens=6, use_cache=False)
            model._setup_cache(cfg, cls=DynamicCache)
            g_dc = generate_seq2seq(model, enc, MASK[:1], start, max_new_tokens=6, use_cache=True)
            model._clean_cache()
            model._setup_cache(cfg, cls=StaticCache)
            g_sc = generate_seq2seq(model, enc, MASK[:1], start, max_new_tokens=6, use_cache=True)
            model._clean_cache()
            assert torch.equal(g_nc, g_dc) and torch.equal(g_nc, g_sc), (g_nc, g_dc, g_sc)
        save(f"seq2seq_{pos}_{attn or 'mha'}",
             pack(model, {"input_ids": ids, "attention_mask": MASK, "decoder_input_ids": dec_ids, "decoder_attention_mask": dec_mask,
                          "cotangent": cot, "gen_start": start},
                  {"logits": out.logits, "key_value_states": out.key_value_states, "gen_encoder_output": enc, "generate": g_nc, **grads},
                  cfg_meta(cfg, pos=pos, attn=attn)))


if __name__ == "__main__":
    main()
