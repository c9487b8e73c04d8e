import torch, sys
sys.path.insert(0, "/root/repo")
from vyomai_b200 import ops
torch.manual_seed(0)
ok = True
for (R, V, dt) in [(32, 50265, torch.bfloat16), (3, 50265, torch.float32), (64, 1000, torch.bfloat16), (1, 7, torch.float32), (5, 4096, torch.bfloat16)]:
    ld = (V + 7) // 8 * 8
    buf = torch.randn(R, ld, device="cuda").to(dt)
    x = buf[:, :V]
    # plant ties: the maximum appears twice, first index must win
    for r in range(R):
        i, j = sorted(torch.randint(0, V, (2,)).tolist())
        x[r, i] = 9.0; x[r, j] = 9.0
    got = ops.argmax_rows(x)
    ref = torch.topk(x.float(), 1, dim=-1).indices[:, 0]
    ref2 = (x.float() == x.float().max(dim=1, keepdim=True).values).float().argmax(dim=1)
    good = bool((got == ref2).all())
    ok &= good
    print(R, V, dt, "OK" if good else "FAIL", got[:4].tolist(), ref2[:4].tolist())
print("ALL", "PASS" if ok else "FAIL")
