"""Does flash_attn's paged kv-cache kernel (the one Examples/simple_vllm.ipynb calls) run on this GPU? Prints the outcome."""
import torch
try:
    from flash_attn import flash_attn_with_kvcache
    B, H, Hk, D, bs, nb = 2, 8, 2, 64, 16, 8
    q = torch.randn(B, 1, H, D, device="cuda", dtype=torch.bfloat16)
    kc = torch.randn(nb, bs, Hk, D, device="cuda", dtype=torch.bfloat16)
    vc = torch.randn(nb, bs, Hk, D, device="cuda", dtype=torch.bfloat16)
    bt = torch.arange(nb, device="cuda", dtype=torch.int32).view(B, nb // B)
    sl = torch.tensor([20, 33], device="cuda", dtype=torch.int32)
    o = flash_attn_with_kvcache(q, kc, vc, cache_seqlens=sl, block_table=bt, causal=True)
    torch.cuda.synchronize()
    print("flash_attn_with_kvcache OK", tuple(o.shape), float(o.float().abs().mean()))
except Exception as e:  # noqa: BLE001
    print("flash_attn_with_kvcache FAILED:", type(e).__name__, str(e)[:300])
