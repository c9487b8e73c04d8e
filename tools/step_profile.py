"""In-situ kernel times of the captured training step (CUDA graph replays under torch.profiler / CUPTI):
    python tools/step_profile.py [out.md]
Unlike the ncu launch list these are warm-cache, real-clock durations of the kernels as they run inside the step."""
import io
import os
import sys
from collections import defaultdict
from contextlib import redirect_stdout

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vyomai_b200 import VisionLanguageModel, Vit  # noqa: E402
from vyomai_b200.trainer import Trainer, caption_labels  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        model = VisionLanguageModel(bench.TextCfg(), encoder=Vit(bench.VitCfg()), pos_embedding_type="rope", attention_type="gqa")
    model = model.to(dev).to(torch.bfloat16).train()
    trainer = Trainer(model, lr=1e-5, weight_decay=0.01, max_grad_norm=1.0, use_graph=True)
    px, ids, mask = [t.to(dev) for t in bench.synth_batch(bench.PER_GPU_BATCH, 17, False)]
    labels = caption_labels(ids, mask)
    for _ in range(3):
        trainer.caption_step(px, ids, mask, labels)
    torch.cuda.synchronize()
    steps = 3
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            trainer.caption_step(px, ids, mask, labels)
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0])
    t0, t1 = None, None
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = ev.name
        for a, b in (("void ", ""), ("vy::", ""), ("__nv_bfloat16", "bf16"), ("(int)", ""), ("(bool)", "")):
            name = name.replace(a, b)
        name = name.split("(CUtensorMap")[0][:90]
        agg[name][0] += 1
        agg[name][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    if os.environ.get("VY_PROFILE_SEQUENCE"):
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        n = len(evs) // steps
        print("kernel sequence of the last step (us):")
        for e in evs[-n:]:
            nm = e.name.replace("void ", "").replace("vy::", "").replace("__nv_bfloat16", "bf16").split("(")[0][:60]
            print(f"  {e.device_time_total:8.1f}  {nm}")
    total = sum(v[1] for v in agg.values())
    lines = [f"in-situ kernel time per training step (graph replay, {steps} steps averaged): {total / steps / 1e3:.2f} ms of kernels",
             "", "| kernel | launches/step | us/step | mean us | share |", "|---|---:|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n / steps:.0f} | {t / steps:.1f} | {t / n:.1f} | {100 * t / total:.1f}% |")
    text = "\n".join(lines) + "\n"
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)


if __name__ == "__main__":
    main()
