"""Per-kernel device time of one PaliGemma-scale decode step (B = 1 by default): torch.profiler over a few eager steps."""
import sys, os, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vyomai_b200 import ops
from vyomai_b200.models.paligemma import PaliGemmaConfig, PaliGemmaForConditionalGeneration, StaticCache

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
cfg = PaliGemmaConfig(vision_config=dict(bench.C5_VISION), text_config=dict(bench.C5_TEXT), image_token_index=257152, vocab_size=257216,
                      projection_dim=2048, hidden_size=2048, pad_token_id=0)
torch.set_default_dtype(torch.bfloat16)
with torch.device(dev):
    model = PaliGemmaForConditionalGeneration(cfg)
torch.set_default_dtype(torch.float32)
model.tie_weights(); model.eval()
with torch.no_grad():
    for n, p in model.named_parameters():
        if p.dim() >= 2:
            p.normal_(0.0, 0.02)
ids = torch.cat([torch.full((B, 256), 257152, dtype=torch.long), torch.randint(2, 250000, (B, 8))], dim=1).to(dev)
mask = torch.ones((B, 264), dtype=torch.long, device=dev)
px = torch.rand((B, 3, 224, 224)).to(dev).to(torch.bfloat16)
cache = StaticCache(cfg.text_config, batch_size=B, device=dev, dtype=torch.bfloat16, max_cache_len=384)
o = model(input_ids=ids, pixel_values=px, attention_mask=mask, past_key_values=cache, use_cache=True, logits_last_only=True)
nxt = ops.argmax_rows(o.logits[:, -1]).view(B, 1)
from vyomai_b200.models.paligemma import PaliGemmaDecodeGraph
g = PaliGemmaDecodeGraph(model, cache, mask)
g.tok.copy_(nxt.view(-1)); g.pos.fill_(264)
g.capture()
def step():
    g.graph.replay()
for _ in range(3):
    step()
torch.cuda.synchronize()
N = 4
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        agg[ev.name[:90]][0] += 1
        agg[ev.name[:90]][1] += ev.device_time_total
tot = sum(v[1] for v in agg.values())
print(f"B={B}: {tot / N / 1e3:.3f} ms of kernel time per decode step (graph replay, PDL: kernels overlap, per-kernel times include waiting)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{t / N:9.1f} us/step  {n / N:6.1f} calls  {t / n:8.1f} us/call  {k}")
sk = sorted(ev.device_time_total for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA and "skinny" in ev.name)
n = len(sk) // 3
print("small-batch GEMM calls, us (sorted thirds = q|k|v / o_proj / down_proj by size):",
      [round(sum(sk[i * n:(i + 1) * n]) / max(1, n), 1) for i in range(3)], "min", round(sk[0], 1), "max", round(sk[-1], 1))
