"""Same-box GPU baseline ("the bar is the PyTorch eager/SDPA path on the same B200"): the UNMODIFIED reference
(baseline/_ref/VyomAI, staged by __graft_entry__.build()) running bench.py's workload on the B200 with its own code path —
nn.Linear / F.scaled_dot_product_attention / nn.LayerNorm in eager PyTorch, bf16 autocast, torch.optim.AdamW(fused) +
clip_grad_norm_(1.0), the notebook's shifted cross-entropy — next to this repo's Trainer on the same batch.

    python tools/ref_gpu_bench.py [--steps 20] [--dropout 0.0]      -> one JSON line (gpurun_out/ref_gpu_bench.json)
"""
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
import bench  # noqa: E402


def time_steps(step, steps, warmup):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, float(loss)


def reference_arm(args, dev, batch):
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "VyomAI")):
        return None
    for k in [k for k in sys.modules if k == "VyomAI" or k.startswith("VyomAI.")]:
        del sys.modules[k]
    sys.path.insert(0, ref_root)
    import VyomAI as R  # the reference itself
    assert os.path.realpath(R.__file__).startswith(os.path.realpath(ref_root)), R.__file__
    tcfg, vcfg = bench.TextCfg(), bench.VitCfg()
    tcfg.hidden_dropout_prob = vcfg.hidden_dropout_prob = args.dropout
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        model = R.VisionLanguageModel(tcfg, encoder=R.Vit(vcfg), pos_embedding_type="rope", attention_type="gqa")
    model = model.to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, weight_decay=0.01, fused=True)
    px, ids, mask = batch
    labels = ids.masked_fill(mask == 0, -100)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = model(pixel_values=px, decoder_input_ids=ids, decoder_attention_mask=mask).logits
        # the notebook's loss_fn: position i + 1 of the logits (after the image token) predicts token i + 1
        loss = torch.nn.functional.cross_entropy(logits[:, 1:-1].reshape(-1, logits.shape[-1]).float(), labels[:, 1:].reshape(-1),
                                                 ignore_index=-100)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return loss.detach()

    ms, loss = time_steps(step, args.steps, args.warmup)
    sys.path.remove(ref_root)
    for k in [k for k in sys.modules if k == "VyomAI" or k.startswith("VyomAI.")]:
        del sys.modules[k]
    del model, opt
    torch.cuda.empty_cache()
    return {"ms_per_step": ms, "samples_per_s": px.shape[0] / (ms / 1e3), "loss": loss,
            "what": "unmodified reference, eager PyTorch, bf16 autocast, fused AdamW, clip 1.0"}


def ours_arm(args, dev, batch):
    from vyomai_b200 import VisionLanguageModel, Vit
    from vyomai_b200.trainer import Trainer, caption_labels
    tcfg, vcfg = bench.TextCfg(), bench.VitCfg()
    tcfg.hidden_dropout_prob = vcfg.hidden_dropout_prob = args.dropout
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        model = VisionLanguageModel(tcfg, encoder=Vit(vcfg), pos_embedding_type="rope", attention_type="gqa")
    model = model.to(dev).to(torch.bfloat16).train()
    trainer = Trainer(model, lr=1e-5, weight_decay=0.01, max_grad_norm=1.0, use_graph=True)
    px, ids, mask = batch
    labels = caption_labels(ids, mask)
    ms, loss = time_steps(lambda: trainer.caption_step(px, ids, mask, labels), args.steps, max(args.warmup, 3))
    return {"ms_per_step": ms, "samples_per_s": px.shape[0] / (ms / 1e3), "loss": loss,
            "what": "vyomai_b200 Trainer (sm_100a kernels, CUDA graph)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dropout", type=float, default=0.0)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    batch = tuple(t.to(dev) for t in bench.synth_batch(bench.PER_GPU_BATCH, 17, False))
    ref = reference_arm(args, dev, batch)
    ours = ours_arm(args, dev, batch)
    line = {"workload": bench.WORKLOAD, "per_gpu_batch": bench.PER_GPU_BATCH, "dropout": args.dropout, "steps": args.steps,
            "reference_gpu": ref, "ours": ours,
            "speedup_vs_reference_gpu": None if ref is None else ref["ms_per_step"] / ours["ms_per_step"]}
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(line, open(os.path.join(ROOT, "gpurun_out", f"ref_gpu_bench_p{args.dropout}.json"), "w"), indent=1)


if __name__ == "__main__":
    t0 = time.time()
    main()
