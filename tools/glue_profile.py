"""Which torch (non-vy) kernels run inside the captured training step, and how long (CUPTI over graph replays)."""
import collections, io, os, sys
from contextlib import redirect_stdout
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vyomai_b200.trainer import Trainer
dev = torch.device("cuda", 0)
wl = bench.workload_of(sys.argv[1] if len(sys.argv) > 1 else "package")
torch.manual_seed(0)
with redirect_stdout(io.StringIO()):
    model = wl["build"]()
model = model.to(dev).to(torch.bfloat16).train()
tr = Trainer(model, lr=1e-5, weight_decay=wl["wd"], max_grad_norm=1.0, use_graph=True)
px, ids, mask = [t.to(dev) for t in wl["synth"](wl["batch"], 23, False)]
labels = wl["labels"](ids, mask)
for _ in range(5):
    tr.caption_step(px, ids, mask, labels)
torch.cuda.synchronize()
N = 3
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        tr.caption_step(px, ids, mask, labels)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and "vy::" not in ev.name:
        agg[ev.name[:110]][0] += 1
        agg[ev.name[:110]][1] += ev.device_time_total
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{t / N:8.1f} us/step {n / N:5.1f} calls  {k}")
