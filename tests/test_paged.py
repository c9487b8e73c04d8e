"""Paged kv-cache / continuous batching (reference: Examples/simple_vllm.ipynb cell 2).

CPU: the oracle's paged decode equals the (golden-pinned) contiguous decode on the same logical cache; the block
manager's bookkeeping follows the notebook. GPU: vy_attn_decode in paged / per-row mode against the oracle, against its
own contiguous mode, and (when flash-attn runs on the box) against flash_attn_with_kvcache — the call the notebook
makes; the engine's greedy tokens against DecoderModel.generate.
Tolerances: fp32 caches, bf16 q/k/v products are not involved (the decode kernel is fp32 FMA) -> 2e-5 relative;
bf16 caches -> 2e-2 (storage rounding)."""
from dataclasses import make_dataclass

import pytest
import torch

from oracle import vyom_oracle as O

CFG = make_dataclass("CFG", [("hidden_size", int, 128), ("num_attention_heads", int, 2), ("num_key_value_heads", int, 1),
                             ("max_position_embeddings", int, 64), ("num_hidden_layers", int, 2), ("vocab_size", int, 101),
                             ("hidden_dropout_prob", float, 0.0), ("layer_norm_eps", float, 1e-5), ("hidden_act", str, "gelu")])


def _paged_case(seed, B, Hq, Hkv, bs, max_blocks, pool_blocks, ctx, dtype):
    g = torch.Generator().manual_seed(seed)
    perm = torch.randperm(pool_blocks, generator=g)
    table = perm[: B * max_blocks].view(B, max_blocks).to(torch.int32)
    k_pool = torch.zeros(pool_blocks, bs, Hkv, 64, dtype=dtype)
    v_pool = torch.zeros_like(k_pool)
    L = max_blocks * bs
    kc = torch.zeros(B, Hkv, L, 64, dtype=dtype)  # the same logical cache, contiguous
    vc = torch.zeros_like(kc)
    for b in range(B):
        n = ctx[b]
        if n <= 0:
            continue
        k = torch.randn(n, Hkv, 64, generator=g).to(dtype)
        v = torch.randn(n, Hkv, 64, generator=g).to(dtype)
        slots = O.paged_slots(table[b], 0, n, bs)
        k_pool.view(-1, Hkv, 64)[slots] = k
        v_pool.view(-1, Hkv, 64)[slots] = v
        kc[b, :, :n] = k.transpose(0, 1)
        vc[b, :, :n] = v.transpose(0, 1)
    qkv = torch.randn(B, (Hq + 2 * Hkv) * 64, generator=g)
    return table, k_pool, v_pool, kc, vc, qkv


def test_oracle_paged_decode_equals_contiguous_decode():
    B, Hq, Hkv, bs = 4, 6, 2, 4
    ctx = [0, 3, 8, 13]
    table, k_pool, v_pool, kc, vc, qkv = _paged_case(0, B, Hq, Hkv, bs, 5, 32, ctx, torch.float32)
    q = qkv[:, : Hq * 64].view(B, Hq, 64)
    kn = qkv[:, Hq * 64:(Hq + Hkv) * 64].view(B, Hkv, 64)
    vn = qkv[:, (Hq + Hkv) * 64:].view(B, Hkv, 64)
    out = O.paged_decode_attention(q, kn, vn, k_pool, v_pool, table, torch.tensor(ctx), bs)
    for b in range(B):
        n = ctx[b]
        k = torch.cat([kc[b, :, :n], kn[b][:, None]], dim=1)[None]
        v = torch.cat([vc[b, :, :n], vn[b][:, None]], dim=1)[None]
        ref = O.sdpa(q[b].view(1, Hq, 1, 64), O.repeat_kv(k, Hq // Hkv), O.repeat_kv(v, Hq // Hkv), None)[0, :, 0]
        assert torch.allclose(out[b], ref, atol=1e-6)
        # the new token landed in the block the table names for position n
        blk, off = int(table[b][n // bs]), n % bs
        assert torch.equal(k_pool[blk, off], kn[b]) and torch.equal(v_pool[blk, off], vn[b])


def test_paged_manager_bookkeeping_follows_the_notebook():
    import io
    from contextlib import redirect_stdout
    from vyomai_b200 import DecoderModel
    from vyomai_b200.paged import PagedKVManager, SequenceState
    with redirect_stdout(io.StringIO()):
        model = DecoderModel(CFG(), "rope", "gqa")
    mgr = PagedKVManager(model, max_blocks=6, block_size=4, dtype=torch.float32, device="cpu")
    assert mgr.k_cache[0].shape == (6, 4, 1, 64) and len(mgr.k_cache) == 2
    a = SequenceState(0, list(range(9)), max_gen_len=3, block_size=4, device="cpu")  # 9 tokens -> 3 blocks, room for 12
    assert a.block_table.numel() == 3 and mgr.can_allocate(9)
    mgr.allocate(a)
    assert a.block_count == 3 and a.block_table.tolist() == [0, 1, 2] and len(mgr.free_blocks) == 3
    assert a.slots(3, 6).tolist() == [3, 4, 5]
    b = SequenceState(1, list(range(14)), max_gen_len=2, block_size=4, device="cpu")  # needs 4 blocks, 3 are free
    assert not mgr.can_allocate(14)
    with pytest.raises(RuntimeError, match="KV Cache full"):
        mgr.allocate(b)
    mgr.free(b)
    mgr.free(a)
    assert a.block_count == 0 and sorted(mgr.free_blocks) == [0, 1, 2, 3, 4, 5]


@pytest.mark.gpu
@pytest.mark.parametrize("bs,max_blocks,pool,ctx", [(16, 8, 64, [0, 5, 16, 47, 100, 127]),
                                                     (256, 2, 16, [0, 5, 255, 256, 300, 511])])  # 256: what flash-attn pages by
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_paged_decode_kernel_matches_oracle_and_contiguous_mode(dtype, tol, bs, max_blocks, pool, ctx):
    from vyomai_b200 import ops
    B, Hq, Hkv = 6, 12, 4
    table, k_pool, v_pool, kc, vc, qkv = _paged_case(1, B, Hq, Hkv, bs, max_blocks, pool, ctx, dtype)
    freqs = O.rope_freqs(max_blocks * bs, 64)[0]  # [L, 32]
    cos, sin = freqs.cos().contiguous(), freqs.sin().contiguous()
    # oracle: rotate q / k of the new token at its position, then the paged step
    q = qkv[:, : Hq * 64].view(B, Hq, 64)
    kn = qkv[:, Hq * 64:(Hq + Hkv) * 64].view(B, Hkv, 64)
    vn = qkv[:, (Hq + Hkv) * 64:].view(B, Hkv, 64)
    qr, kr = torch.empty_like(q), torch.empty_like(kn)
    for b in range(B):
        f = O.rope_freqs(max_blocks * bs, 64)[:, ctx[b]:ctx[b] + 1]
        a, c = O.apply_rope(q[b].view(1, Hq, 1, 64), kn[b].view(1, Hkv, 1, 64), f)
        qr[b], kr[b] = a[0, :, 0], c[0, :, 0]
    kp_o, vp_o = k_pool.clone(), v_pool.clone()
    ref = O.paged_decode_attention(qr, kr, vn, kp_o, vp_o, table, torch.tensor(ctx), bs).reshape(B, Hq * 64)

    dev = "cuda"
    kp, vp = k_pool.to(dev), v_pool.to(dev)
    seq = torch.tensor(ctx, dtype=torch.int32, device=dev)
    seq_idle = seq.clone()
    seq_idle[2] = -1  # an idle batch slot is skipped
    out = torch.full((B, Hq * 64), 7.0, device=dev)
    ops.attn_decode(qkv.to(dev), kp.clone(), vp.clone(), max(ctx), Hq, Hkv, cos.to(dev), sin.to(dev), out=out, seqlens=seq_idle,
                    block_table=table.to(dev))
    assert bool((out[2] == 7.0).all())
    out = ops.attn_decode(qkv.to(dev), kp, vp, max(ctx), Hq, Hkv, cos.to(dev), sin.to(dev), out_dtype=torch.float32, seqlens=seq,
                          block_table=table.to(dev))
    err = float((out.cpu() - ref).norm() / ref.norm())
    assert err < tol, err
    # the append: pools now equal the oracle's pools (new rows written, nothing else touched)
    assert torch.allclose(kp.cpu().float(), kp_o.float(), atol=tol) and torch.allclose(vp.cpu().float(), vp_o.float(), atol=tol)
    untouched = torch.ones(pool, dtype=torch.bool)
    untouched[table.flatten().long()] = False
    assert bool((kp.cpu()[untouched] == 0).all())
    # contiguous mode, one row at a time, computes the same numbers
    kcd, vcd = kc.to(dev), vc.to(dev)
    for b in range(B):
        o1 = ops.attn_decode(qkv[b:b + 1].to(dev), kcd[b:b + 1], vcd[b:b + 1], ctx[b], Hq, Hkv, cos.to(dev), sin.to(dev),
                             out_dtype=torch.float32)
        assert float((o1[0] - out[b]).abs().max()) < 1e-4 * max(1.0, float(out[b].abs().max())), b
    # The call the notebook makes (Examples/simple_vllm.ipynb cell 2: flash_attn_with_kvcache over the block table) — an
    # implementation independent of this repo and of the oracle. flash-attn 2.8 pages by multiples of 256 slots, so the pin
    # runs in the block-size-256 cases; it is mandatory there whenever the package imports (no except around the call).
    if bs % 256 == 0:
        try:
            from flash_attn import flash_attn_with_kvcache
        except ImportError:
            flash_attn_with_kvcache = None
        if flash_attn_with_kvcache is not None:
            fa = flash_attn_with_kvcache(qr.to(dev).to(torch.bfloat16).unsqueeze(1), kp.to(torch.bfloat16), vp.to(torch.bfloat16),
                                         cache_seqlens=seq + 1, block_table=table.to(dev), causal=True)
            fa_err = float((fa.float().reshape(B, -1).cpu() - ref).norm() / ref.norm())
            ours_vs_fa = float((fa.float().reshape(B, -1) - out).norm() / out.norm())
            assert fa_err < 2e-2 and ours_vs_fa < 2e-2, (fa_err, ours_vs_fa)  # flash-attn computes in bf16
            print(f"paged decode pinned by flash_attn_with_kvcache: oracle vs flash-attn {fa_err:.2e}, kernel vs flash-attn {ours_vs_fa:.2e}")


@pytest.mark.gpu
def test_continuous_batch_engine_matches_generate():
    import io
    from contextlib import redirect_stdout
    from vyomai_b200 import DecoderModel
    from vyomai_b200.paged import ContinuousBatchEngine, PagedKVManager
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        model = DecoderModel(CFG(), "rope", "gqa").cuda().eval()
    g = torch.Generator().manual_seed(3)
    prompts = [torch.randint(3, 101, (n,), generator=g).tolist() for n in (3, 7, 12, 5, 9)]
    want = {}
    for i, p in enumerate(prompts):
        ids = torch.tensor([p], device="cuda")
        want[i] = model.generate(ids, torch.ones_like(ids), max_len=6, use_cache=True, use_static_cache=True)[0].tolist()
    # 14 blocks of 4 slots hold the three largest sequences at full length (5 + 4 + 4 blocks) but not all five requests,
    # and at most three are active at once: the waiting room is used
    mgr = PagedKVManager(model, max_blocks=14, block_size=4)
    eng = ContinuousBatchEngine(model, mgr, max_batch_size=3, eos_token_ids=[2])  # generate()'s default stop token
    for p in prompts:
        eng.add_sequence(p, max_gen_len=6)
    got, steps = {}, 0
    while eng.waiting_room or eng.active:
        got.update(eng.step())
        steps += 1
        assert steps < 200
    assert sorted(got) == [0, 1, 2, 3, 4]
    assert len(mgr.free_blocks) == 14  # every block came back
    for i in range(5):  # same tokens up to and including the stop token (generate() pads the rest of its row)
        assert got[i] == want[i][: len(got[i])], (i, got[i], want[i])
        assert len(got[i]) == len(prompts[i]) + 6 or got[i][-1] == 2


def test_engine_scheduling_and_retirement_on_the_host():
    """ContinuousBatchEngine's bookkeeping with the two model calls stubbed out (no GPU): at most max_batch_size
    sequences are active, a request waits while the pool cannot hold its prompt, block tables grow one block at a time
    during decoding, eos / length retire a sequence and return its blocks (notebook: ContinuousBatchEngine.step)."""
    import io
    from contextlib import redirect_stdout
    from vyomai_b200 import DecoderModel
    from vyomai_b200.paged import ContinuousBatchEngine, PagedKVManager
    with redirect_stdout(io.StringIO()):
        model = DecoderModel(CFG(), "rope", "gqa")
    mgr = PagedKVManager(model, max_blocks=8, block_size=4, dtype=torch.float32, device="cpu")

    class Stub(ContinuousBatchEngine):
        def _prefill(self, s):      # next token = 10 + sequence id
            return 10 + s.id

        def _decode(self, states):  # sequence 1 emits eos (2) on its second decode step, the others count up
            return [2 if (s.id == 1 and s.num_tokens == 8) else 20 + s.num_tokens for s in states]

    eng = Stub(model, mgr, max_batch_size=2, eos_token_ids=[2])
    sid0 = eng.add_sequence([5, 6, 7], max_gen_len=3)             # 3 + 3 tokens: 2 blocks at most
    sid1 = eng.add_sequence([5, 6, 7, 8, 9, 3], max_gen_len=6)    # stops at eos after 3 generated tokens
    sid2 = eng.add_sequence(list(range(3, 16)), max_gen_len=2)    # 13-token prompt: 4 blocks
    assert (sid0, sid1, sid2) == (0, 1, 2)
    done, max_active, steps = {}, 0, 0
    while eng.waiting_room or eng.active:
        done.update(eng.step())
        max_active = max(max_active, len(eng.active))
        steps += 1
        assert len(mgr.free_blocks) + sum(s.block_count for s in eng.active.values()) == 8  # no block is lost
        assert steps < 50
    assert max_active <= 2
    assert done[0] == [5, 6, 7, 10, 24, 25]                       # prefill token, then two decode tokens -> length 6
    assert done[1] == [5, 6, 7, 8, 9, 3, 11, 27, 2]               # ends with the eos token
    assert done[2][:13] == list(range(3, 16)) and len(done[2]) == 15 and done[2][13] == 12
    assert len(mgr.free_blocks) == 8 and not eng.active and not eng.waiting_room
