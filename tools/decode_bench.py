"""Config 3 of BASELINE.json: DecoderModel (GPT-style CLM, RoPE) prefill 512 + greedy decode 256, batch 32, bf16,
StaticCacheOne. Prints prefill tok/s, decode tok/s (device-timed forward + argmax per step), the HBM roofline of a
decode step (weights + kv-cache bytes over the measured copy bandwidth) and generate()'s own wall clock.
    python tools/decode_bench.py [--attn gqa|mha] [--layers 4] [--batch 32]"""
import argparse
import io
import json
import os
import sys
import time
from contextlib import redirect_stdout

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vyomai_b200 import DecoderModel, StaticCacheOne, ops  # noqa: E402


class Cfg:
    hidden_size = 768
    num_attention_heads = 12
    num_key_value_heads = 4
    max_position_embeddings = 1024
    num_hidden_layers = 4
    vocab_size = 50265
    hidden_dropout_prob = 0.0
    initializer_range = 0.02
    intermediate_size = 3072
    layer_norm_eps = 1e-05
    hidden_act = "gelu"
    pad_token_id = 1
    eos_token_id = 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--attn", default="gqa")
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--prefill", type=int, default=512)
    ap.add_argument("--decode", type=int, default=256)
    ap.add_argument("--trace", action="store_true", help="stage-by-stage %globaltimer trace of one fused decode step (stderr)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = Cfg()
    cfg.num_hidden_layers = args.layers
    if args.attn != "gqa":
        del Cfg.num_key_value_heads
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        model = DecoderModel(cfg, "rope", "gqa" if args.attn == "gqa" else None)
    model = model.to(dev).to(torch.bfloat16).eval()
    B, P, N = args.batch, args.prefill, args.decode
    ids = torch.randint(3, cfg.vocab_size, (B, P), device=dev)
    mask = torch.ones((B, P), device=dev, dtype=torch.long)
    hkv = 4 if args.attn == "gqa" else 12

    def run(timed):
        cache = StaticCacheOne(cfg, max_cache_len=P + N, batch_size=B, dtype=torch.bfloat16)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        with torch.no_grad():
            ev[0].record()
            out = model(ids, mask, use_cache=True, kv_cache=cache, start_pos=0, _logits_last_only=True)
            tok = ops.argmax_rows(out.logits[:, -1])
            ev[1].record()
            toks = [tok]
            for t in range(N - 1):
                out = model(tok.view(B, 1), None, use_cache=True, kv_cache=out.kv_cache, start_pos=P + t, _logits_last_only=True)
                tok = ops.argmax_rows(out.logits[:, -1])
                toks.append(tok)
            ev[2].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), torch.stack(toks, 1)

    run(False)
    pre_ms, dec_ms, toks = run(True)
    n_params = sum(p.numel() for n, p in model.named_parameters() if "word_embeddings" not in n)
    mean_ctx = P + N / 2
    step_bytes = 2.0 * n_params + 2.0 * B * args.layers * 2 * hkv * mean_ctx * 64
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    step_us = dec_ms * 1e3 / (N - 1)
    model.generate(ids, mask, max_len=N, use_cache=True, use_static_cache=True)  # captures the decode graph
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    full = model.generate(ids, mask, max_len=N, use_cache=True, use_static_cache=True)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    g = model._decode_graph  # device time of the replayed steps alone
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.pos.fill_(P)
    e0.record()
    for _ in range(N - 1):
        g.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    graph_step_us = e0.elapsed_time(e1) * 1e3 / (N - 1)
    if os.environ.get("VY_PROFILE_SEQUENCE"):
        g.pos.fill_(P + N // 2)
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            g.graph.replay()
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        for e in evs:
            print(f"  {e.device_time_total:8.1f}  {e.name[:100]}", file=sys.stderr)
    if args.trace and getattr(g, "fused", None) is not None:
        from vyomai_b200 import decode_step
        L = args.layers
        nb = 5 * L + 2
        tr = torch.zeros(256, dtype=torch.int64, device=dev)
        tok = g.tok.clone()
        pos = torch.full((1,), P + N // 2, dtype=torch.int32, device=dev)
        st = decode_step.FusedDecodeStep(model, g.cache, B, tok, pos, None, pos_bound=g.cache.key_cache[0].shape[2] - 1, trace=tr)
        for _ in range(3):
            pos.fill_(P + N // 2)
            st.launch()
        torch.cuda.synchronize()
        t = tr.cpu().tolist()
        names = []
        for l in range(L):
            names += [f"L{l}.qkv", f"L{l}.attn", f"L{l}.out", f"L{l}.ffn1", f"L{l}.ffn2"]
        names += ["lm.dense", "lm.vocab+argmax"]
        print(f"fused step trace (CTA 0): total {(t[2 * nb] - t[0]) / 1e3:.1f} us", file=sys.stderr)
        prev = t[0]
        for k, nm in enumerate(names):
            own, wait = t[1 + 2 * k] - prev, t[2 + 2 * k] - t[1 + 2 * k]
            print(f"  {nm:16s} own {own / 1e3:7.2f} us   wait+barrier {wait / 1e3:7.2f} us", file=sys.stderr)
            prev = t[2 + 2 * k]
        def rel(slot, ref):
            return (t[128 + slot] - ref) / 1e3
        q0 = t[2 + 2 * 4]  # start of L1.qkv = barrier 4 opened
        print(f"  L1.qkv detail: prologue {rel(0, q0):.2f}  weights consumed {rel(1, q0):.2f}  partials in smem {rel(2, q0):.2f}  done {rel(3, q0):.2f} us", file=sys.stderr)
        print(f"  L1.attn detail (warp 0's first item): prologue {rel(16, t[2 + 2 * 5]):.2f}  keys done {rel(17, t[2 + 2 * 5]):.2f}  merged {rel(18, t[2 + 2 * 5]):.2f} us", file=sys.stderr)
        f0 = t[2 + 2 * 8]  # start of L1.ffn2 = barrier 8 opened
        print(f"  L1.ffn2 detail: prologue {rel(8, f0):.2f}  weights consumed {rel(9, f0):.2f}  partials in smem {rel(10, f0):.2f}  done {rel(11, f0):.2f}  last-CTA LN start {rel(12, f0):.2f} end {rel(13, f0):.2f} us", file=sys.stderr)
    res = {
        "workload": f"decoder_clm_L{args.layers}_{args.attn}_B{B}_prefill{P}_decode{N}_bf16_staticcache",
        "prefill_tok_per_s": B * P / (pre_ms / 1e3), "prefill_ms": pre_ms,
        "decode_tok_per_s": B * (N - 1) / (dec_ms / 1e3), "decode_us_per_step": step_us,
        "decode_step_algorithmic_MB": step_bytes / 1e6,
        "decode_hbm_frac_of_measured": step_bytes / (step_us * 1e-6) / 1e9 / peaks["hbm_gbs"],
        "generate_wall_s": gen_s, "generate_tok_per_s": B * N / gen_s,
        "graph_decode_us_per_step": graph_step_us, "graph_decode_tok_per_s": B / (graph_step_us * 1e-6),
        "graph_decode_hbm_frac_of_measured": step_bytes / (graph_step_us * 1e-6) / 1e9 / peaks["hbm_gbs"],
        "ids_match_generate": bool((full[:, P:P + N] == toks).all().item()),
    }
    print(json.dumps(res))


if __name__ == "__main__":
    main()
