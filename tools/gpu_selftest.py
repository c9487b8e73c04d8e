"""Quick on-GPU kernel checks against torch's own CUDA ops (development aid, not the parity
suite — that lives in tests/ and checks against oracle/). Usage:
    python tools/gpu_selftest.py            # runs every group, each in its own subprocess
    python tools/gpu_selftest.py --group gemm_bf16
"""
import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GROUPS = {}


def group(fn):
    GROUPS[fn.__name__] = fn
    return fn


def rel_err(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), (a - b).abs().max().item()


def report(name, got, ref, tol):
    r, m = rel_err(got, ref)
    ok = r <= tol and bool(got.isfinite().all().item())
    print(f"  [{'PASS' if ok else 'FAIL'}] {name}: rel={r:.3e} maxabs={m:.3e} tol={tol:.1e}", flush=True)
    from vyomai_b200 import _lib, gemm_tune
    if _lib.lib().vy_gemm_poisoned() != 0:  # a wait inside a GEMM kernel timed out (also during the tuning runs of this case)
        print(f"  [FAIL] {name}: vy_gemm_poisoned; tuned so far: {gemm_tune.LOG[-3:]}", flush=True)
        ok = False
    return ok


def _gemm_cases(dtype, tol):
    import torch
    from vyomai_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    ok = True
    shapes = [(128, 128, 64), (256, 256, 128), (384, 768, 768), (1024, 3072, 768), (1000, 520, 264),
              (8192, 768, 3072), (300, 50265 if dtype == torch.bfloat16 else 1000, 768)]
    for (M, N, K) in shapes:
        a = torch.randn(M, K, device=dev, dtype=dtype)
        b = torch.randn(N, K, device=dev, dtype=dtype) / K ** 0.5
        bias = torch.randn(N, device=dev, dtype=dtype)
        ref = (a.float() @ b.float().t()) + bias.float()
        ldo = (N + 7) // 8 * 8
        outbuf = torch.empty(M, ldo, device=dev, dtype=dtype)
        out = ops.gemm(a, b, bias=bias, out=outbuf[:, :N])
        ok &= report(f"NT {M}x{N}x{K} bias", out, ref, tol)
        at = a.t().contiguous().t()  # logical (M,K) stored transposed -> MN-major
        bt = b.t().contiguous().t()
        if M % 8 == 0 and N % 8 == 0:
            ok &= report(f"A_MN {M}x{N}x{K}", ops.gemm(at, b, bias=bias, out=outbuf[:, :N]), ref, tol)
            ok &= report(f"B_MN {M}x{N}x{K}", ops.gemm(a, bt, bias=bias, out=outbuf[:, :N]), ref, tol)
            ok &= report(f"AB_MN {M}x{N}x{K}", ops.gemm(at, bt, bias=bias, out=outbuf[:, :N]), ref, tol)
    # narrow tiles walked many times by each CTA (pinned through the tuning hook; the time model never picks them here)
    from vyomai_b200 import _lib
    M, N, K = 2048, 3072, 256
    a = torch.randn(M, K, device=dev, dtype=dtype)
    b = torch.randn(N, K, device=dev, dtype=dtype) / K ** 0.5
    bias = torch.randn(N, device=dev, dtype=dtype)
    ref = (a.float() @ b.float().t()) + bias.float()
    for bn in (32, 64):
        _lib.lib().vy_gemm_tune_override(0, bn, 0)
        ok &= report(f"NT {M}x{N}x{K} pinned BN={bn}", ops.gemm(a, b, bias=bias), ref, tol)
        ok &= report(f"NT {M}x{N}x{K} pinned BN={bn} fp32 out", ops.gemm(a, b, bias=bias, out_dtype=torch.float32), ref, tol)
    _lib.lib().vy_gemm_tune_override(-1, 0, 0)
    # epilogues
    M, N, K = 512, 768, 768
    a = torch.randn(M, K, device=dev, dtype=dtype)
    b = torch.randn(N, K, device=dev, dtype=dtype) / K ** 0.5
    bias = torch.randn(N, device=dev, dtype=dtype)
    res = torch.randn(M, N, device=dev, dtype=dtype)
    z = a.float() @ b.float().t() + bias.float()
    aux = torch.empty(M, N, device=dev, dtype=dtype)
    out = ops.gemm(a, b, bias=bias, act="gelu", aux=aux)
    ok &= report("gelu_erf", out, torch.nn.functional.gelu(z), tol)
    ok &= report("gelu aux(pre-act)", aux, z, tol)
    out = ops.gemm(a, b, bias=bias, act="gelu_tanh")
    ok &= report("gelu_tanh", out, torch.nn.functional.gelu(z, approximate="tanh"), tol)
    out = ops.gemm(a, b, bias=bias, addend=res)
    ok &= report("residual addend", out, z + res.float(), tol)
    zz = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    out = ops.gemm(a, b, act="dgelu", aux=aux)
    ok &= report("dgelu", out, (a.float() @ b.float().t()) * zz.grad, tol * 2)
    other = torch.float32 if dtype == torch.bfloat16 else torch.bfloat16
    ok &= report("cross out dtype", ops.gemm(a, b, bias=bias, out_dtype=other), z, max(tol, 4e-3))
    # gated MLP epilogue: interleaved gate / up rows -> silu(gate) * up; staged fast path (no save) and general path (save z)
    if dtype == torch.bfloat16:
        for (Mg, Kg, Ig) in [(300, 256, 320), (1024, 1024, 3072), (257, 128, 104)]:
            xg = torch.randn(Mg, Kg, device=dev, dtype=dtype)
            wg = torch.randn(Ig, Kg, device=dev, dtype=dtype) / Kg ** 0.5
            wu = torch.randn(Ig, Kg, device=dev, dtype=dtype) / Kg ** 0.5
            bg = torch.randn(2 * Ig, device=dev, dtype=dtype)
            wi = torch.stack((wg, wu), 1).reshape(2 * Ig, Kg)
            for bb in (None, bg):
                zg = xg.float() @ wg.float().t() + (0 if bb is None else bb[0::2].float())
                zu = xg.float() @ wu.float().t() + (0 if bb is None else bb[1::2].float())
                refg = zg * torch.sigmoid(zg) * zu
                ok &= report(f"swiglu {Mg}x{2 * Ig}x{Kg} bias={bb is not None}", ops.gemm(xg, wi, bias=bb, act="swiglu"), refg, tol * 2)
                zsave = torch.empty(Mg, 2 * Ig, device=dev, dtype=dtype)
                ok &= report("   + save z", ops.gemm(xg, wi, bias=bb, act="swiglu", aux=zsave), refg, tol * 2)
                zi = torch.stack((zg, zu), 2).reshape(Mg, 2 * Ig)
                ok &= report("   z", zsave, zi, tol)
    # swap-AB (decode shapes)
    for (M, N, K) in [(32, 768, 768), (32, 3072, 768), (32, 768, 3072), (3, 2304, 768), (64, 50265, 768), (1, 768, 768)]:
        a = torch.randn(M, K, device=dev, dtype=dtype)
        b = torch.randn(N, K, device=dev, dtype=dtype) / K ** 0.5
        bias = torch.randn(N, device=dev, dtype=dtype)
        res = torch.randn(M, N, device=dev, dtype=dtype)
        ref = torch.nn.functional.gelu(a.float() @ b.float().t() + bias.float()) + res.float()
        out = ops.gemm(a, b, bias=bias, act="gelu", addend=res, swap_ab=True)
        ok &= report(f"swapAB {M}x{N}x{K} gelu+res", out, ref, tol)
    # patch-embed style remap: rows grouped 196 -> 197 with offset 1, addend row mod, scale 2
    Bn, P, Hd, Kp = 4, 196, 768, 768
    a = torch.randn(Bn * P, Kp, device=dev, dtype=dtype)
    w = torch.randn(Hd, Kp, device=dev, dtype=dtype) / Kp ** 0.5
    bias = torch.randn(Hd, device=dev, dtype=dtype)
    pos = torch.randn(P + 1, Hd, device=dev, dtype=dtype)
    outb = torch.zeros(Bn * (P + 1), Hd, device=dev, dtype=dtype)
    ops.gemm(a, w, bias=bias, addend=pos, addend_row_mod=P, addend_row_off=1, out=outb, out_scale=2.0,
             out_row_group=P, out_row_group_stride=P + 1, out_row_off=1)
    ref = 2 * ((a.float() @ w.float().t() + bias.float()).view(Bn, P, Hd) + pos.float()[1:])
    ok &= report("patch remap", outb.view(Bn, P + 1, Hd)[:, 1:], ref, tol)
    ok &= bool((outb.view(Bn, P + 1, Hd)[:, 0] == 0).all().item())
    # split-K (weight-gradient shapes: few output tiles, K = tokens), accumulating into the output buffer
    for (M, N, K) in [(768, 768, 8192), (1280, 768, 8192), (2304, 768, 12608), (768, 3072, 4096), (520, 264, 5000)]:
        K8 = (K + 7) // 8 * 8
        at = torch.randn(K8, M, device=dev, dtype=dtype) / K8 ** 0.5   # logical A = at.t() (MN-major)
        bt = torch.randn(K8, N, device=dev, dtype=dtype)
        acc0 = torch.randn(M, N, device=dev, dtype=dtype)
        ref = acc0.float() + 0.5 * (at.float().t() @ bt.float())
        out = acc0.clone()
        ops.gemm(at.t(), bt.t(), out=out, addend=None, out_scale=1.0, allow_split_k=True)
        ok &= report(f"splitK {M}x{N}x{K8} plain", out, at.float().t() @ bt.float(), tol * 2)
        out = (2 * acc0).clone()
        ops.gemm(at.t(), bt.t(), out=out, addend=out, out_scale=0.5, allow_split_k=True)
        ok &= report(f"splitK {M}x{N}x{K8} accum*0.5", out, ref, tol * 2)
    return ok


@group
def gemm_bf16():
    import torch
    return _gemm_cases(torch.bfloat16, 6e-3)


@group
def gemm_tf32():
    import torch
    return _gemm_cases(torch.float32, 2e-3)


@group
def qkv_rope():
    import torch
    from vyomai_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    ok = True
    for dtype, tol in ((torch.bfloat16, 8e-3), (torch.float32, 2e-3)):
        for (B, S, hq, hkv, start, smax) in [(8, 128, 12, 4, 0, 128), (3, 17, 12, 12, 0, 32), (2, 5, 12, 4, 7, 40)]:
            d, H = 64, 768
            N = (hq + 2 * hkv) * d
            x = torch.randn(B * S, H, device=dev, dtype=dtype)
            w = torch.randn(N, H, device=dev, dtype=dtype) / H ** 0.5
            bias = torch.randn(N, device=dev, dtype=dtype)
            inv = 1.0 / (10000 ** (torch.arange(0, d, 2, device=dev).float() / d))
            ang = torch.arange(smax, device=dev).float()[:, None] * inv[None]
            cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
            q = torch.zeros(B, hq, S, d, device=dev, dtype=dtype)
            kc = torch.zeros(B, hkv, smax, d, device=dev, dtype=dtype)
            vc = torch.zeros(B, hkv, smax, d, device=dev, dtype=dtype)
            ops.qkv_rope_gemm(x, w, bias, tokens_per_seq=S, start_pos=start, n_q_heads=hq, n_kv_heads=hkv,
                              head_dim=d, rope_cos=cos, rope_sin=sin, q_out=q, k_out=kc, v_out=vc)
            y = (x.float() @ w.float().t() + bias.float()).view(B, S, hq + 2 * hkv, d).permute(0, 2, 1, 3)
            c = torch.cat([cos, cos], -1)[start:start + S][None, None]
            s_ = torch.cat([sin, sin], -1)[start:start + S][None, None]

            def rot(t):
                t1, t2 = t.chunk(2, -1)
                return t * c + torch.cat([-t2, t1], -1) * s_

            ok &= report(f"q {dtype} B{B} S{S} start{start}", q, rot(y[:, :hq]), tol)
            ok &= report("k cache", kc[:, :, start:start + S], rot(y[:, hq:hq + hkv]), tol)
            ok &= report("v cache", vc[:, :, start:start + S], y[:, hq + hkv:], tol)
            ok &= bool((kc[:, :, :start] == 0).all().item() and (kc[:, :, start + S:] == 0).all().item())
    return ok


@group
def layernorm():
    import torch
    from vyomai_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    ok = True
    for dtype, tol in ((torch.float32, 2e-6), (torch.bfloat16, 4e-3)):
        for rows, H in ((1024, 768), (51, 768), (333, 1152), (64, 2048), (7, 64)):
            x = torch.randn(rows, H, device=dev, dtype=dtype)
            r = torch.randn(rows, H, device=dev, dtype=dtype)
            g = torch.randn(H, device=dev, dtype=dtype)
            b = torch.randn(H, device=dev, dtype=dtype)
            y, s, mean, rstd = ops.add_layernorm(x, r, g, b, 1e-5, save_stats=True, save_sum=True)
            sref = s.float().clone().requires_grad_(True)  # grads at the (rounded) saved sum
            gref = g.float().clone().requires_grad_(True)
            bref = b.float().clone().requires_grad_(True)
            yref = torch.nn.functional.layer_norm(sref, (H,), gref, bref, 1e-5)
            ok &= report(f"ln sum {dtype} {rows}x{H}", s, x.float() + r.float(), tol)
            ok &= report(f"ln fwd {dtype} {rows}x{H}", y, yref, tol)
            dy = torch.randn(rows, H, device=dev, dtype=dtype)
            yref.backward(dy.float())
            dx, dg, db = ops.add_layernorm_bwd(dy, s, g, mean, rstd)
            btol = tol if dtype == torch.bfloat16 else 2e-5
            ok &= report("ln bwd dx", dx, sref.grad, btol)
            ok &= report("ln bwd dgamma", dg, gref.grad, 2e-5 if dtype == torch.float32 else 2e-3)
            ok &= report("ln bwd dbeta", db, bref.grad, 2e-5)
        y2, _, _, _ = ops.add_layernorm(x, None, g, b, 1e-5)
        ok &= report("ln no residual", y2, torch.nn.functional.layer_norm(x.float(), (H,), g.float(), b.float(), 1e-5), tol)
    return ok


def _ref_attn(q, k, v, causal, q_pos0, kpm):
    """fp32 reference with the reference's additive finfo.min masks."""
    import torch
    B, Hq, Sq, D = q.shape
    Hkv, Skv = k.shape[1], k.shape[2]
    n_rep = Hq // Hkv
    kf = k.float().repeat_interleave(n_rep, dim=1)
    vf = v.float().repeat_interleave(n_rep, dim=1)
    s = q.float() @ kf.transpose(-1, -2) / 8.0
    vis = torch.ones(B, 1, Sq, Skv, device=q.device)
    if causal:
        kk = torch.arange(Skv, device=q.device)[None, :]
        ll = torch.arange(Sq, device=q.device)[:, None]
        vis = vis * (kk <= q_pos0 + ll).float()[None, None]
    if kpm is not None:
        vis = vis * kpm.float()[:, None, None, :]
    s = s + (1.0 - vis) * torch.finfo(torch.float32).min
    o = torch.softmax(s, dim=-1) @ vf
    return o.permute(0, 2, 1, 3).reshape(B, Sq, Hq * D)


@group
def attn_fwd():
    import torch
    from vyomai_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    ok = True
    cases = [
        # B, Hq, Hkv, Sq, Skv, causal, q_pos0, pad
        (2, 2, 2, 128, 128, False, 0, None),
        (3, 12, 4, 17, 17, False, 0, "right"),
        (3, 12, 12, 17, 17, True, 0, "right"),
        (8, 12, 4, 128, 128, False, 0, "right"),
        (4, 12, 12, 197, 197, False, 0, None),
        (2, 12, 4, 248, 248, True, 0, "right"),
        (2, 12, 12, 512, 512, True, 0, None),
        (2, 4, 2, 5, 12, True, 7, None),
        (2, 4, 4, 300, 700, False, 0, "right"),
        (2, 4, 2, 130, 130, True, 0, "left"),
        (1, 2, 1, 40, 40, False, 0, "all"),
    ]
    for (B, Hq, Hkv, Sq, Skv, causal, qp, pad) in cases:
        q = torch.randn(B, Hq, Sq, 64, device=dev).bfloat16()
        kbuf = torch.randn(B, Hkv, Skv + 9, 64, device=dev).bfloat16()  # strided like a cache view
        vbuf = torch.randn(B, Hkv, Skv + 9, 64, device=dev).bfloat16()
        k, v = kbuf[:, :, :Skv], vbuf[:, :, :Skv]
        kpm = None
        if pad is not None:
            kpm = torch.ones(B, Skv, device=dev, dtype=torch.uint8)
            for bi in range(B):
                n = max(1, Skv - 3 - 5 * bi)
                if pad == "right":
                    kpm[bi, n:] = 0
                elif pad == "left":
                    kpm[bi, : Skv - n] = 0
                else:
                    kpm[bi, :] = 0
        out, lse = ops.attn_fwd(q, k, v, causal=causal, q_pos0=qp, key_padding_mask=kpm, need_lse=True)
        ref = _ref_attn(q, k, v, causal, qp, kpm)
        ok &= report(f"attn B{B} Hq{Hq} Hkv{Hkv} Sq{Sq} Skv{Skv} causal{int(causal)} pos{qp} pad={pad}", out, ref, 1e-2)
        out32, _ = ops.attn_fwd(q, k, v, causal=causal, q_pos0=qp, key_padding_mask=kpm, out_dtype=torch.float32)
        ok &= report("   fp32 out", out32, ref, 6e-3)
    return ok


@group
def attn_decode():
    import torch
    from vyomai_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    ok = True
    for cdt, tol in ((torch.bfloat16, 8e-3), (torch.float32, 2e-5)):
        for (B, Hq, Hkv, start, clen, splits) in [(32, 12, 12, 640, 768, 0), (32, 12, 4, 513, 768, 0), (3, 12, 4, 0, 16, 0),
                                                   (2, 12, 12, 5, 16, 1), (1, 8, 1, 300, 384, 0), (4, 12, 4, 100, 128, 3)]:
            d = 64
            N = (Hq + 2 * Hkv) * d
            qdt = cdt
            qkv = torch.randn(B, N, device=dev, dtype=qdt)
            kc = torch.zeros(B + 1, Hkv, clen, d, device=dev, dtype=cdt)
            vc = torch.zeros(B + 1, Hkv, clen, d, device=dev, dtype=cdt)
            kc[:, :, :start] = torch.randn(B + 1, Hkv, start, d, device=dev).to(cdt)
            vc[:, :, :start] = torch.randn(B + 1, Hkv, start, d, device=dev).to(cdt)
            inv = 1.0 / (10000 ** (torch.arange(0, d, 2, device=dev).float() / d))
            ang = torch.arange(clen, device=dev).float()[:, None] * inv[None]
            cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
            kref, vref = kc.clone(), vc.clone()
            out = ops.attn_decode(qkv, kc, vc, start, Hq, Hkv, cos, sin, splits=splits)
            y = qkv.float().view(B, Hq + 2 * Hkv, d)
            c = torch.cat([cos, cos], -1)[start]
            s_ = torch.cat([sin, sin], -1)[start]

            def rot(t):
                t1, t2 = t.chunk(2, -1)
                return t * c + torch.cat([-t2, t1], -1) * s_

            qh = rot(y[:, :Hq])
            knew = rot(y[:, Hq:Hq + Hkv])
            vnew = y[:, Hq + Hkv:]
            kref[:B, :, start] = knew.to(cdt)
            vref[:B, :, start] = vnew.to(cdt)
            n_rep = Hq // Hkv
            kk = kref[:B, :, :start + 1].float()
            vv = vref[:B, :, :start + 1].float()
            kk[:, :, start] = knew
            vv[:, :, start] = vnew
            kk = kk.repeat_interleave(n_rep, 1)
            vv = vv.repeat_interleave(n_rep, 1)
            sc = torch.einsum("bhd,bhkd->bhk", qh, kk) / 8.0
            ref = torch.einsum("bhk,bhkd->bhd", torch.softmax(sc, -1), vv).reshape(B, Hq * d)
            ok &= report(f"decode {cdt} B{B} Hq{Hq} Hkv{Hkv} start{start} splits{splits}", out, ref, tol)
            ok &= report("   k cache append", kc, kref, 1e-6 if cdt == torch.float32 else 4e-3)
            ok &= report("   v cache append", vc, vref, 1e-6 if cdt == torch.float32 else 4e-3)
            ok &= bool((kc[B] == kref[B]).all().item())
    return ok


@group
def attn_bwd():
    import torch
    from vyomai_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    ok = True
    cases = [
        (2, 2, 2, 128, False, None, False),
        (3, 12, 4, 17, False, "right", True),
        (2, 12, 12, 248, True, "right", True),
        (2, 12, 4, 128, True, None, True),
        (2, 4, 2, 300, False, "right", False),
        (1, 6, 2, 197, False, None, False),
        (2, 4, 4, 130, True, "left", True),
        (2, 12, 12, 197, False, None, False),
        (2, 3, 3, 256, True, "right", True),
        (3, 6, 2, 100, True, "right", True),
    ]
    for (B, Hq, Hkv, S, causal, pad, use_rope) in cases:
        d = 64
        n_rep = Hq // Hkv
        # leaf tensors are the PRE-rope projections so the test also covers the fused inverse rotation
        qp = torch.randn(B, Hq, S, d, device=dev).bfloat16().float().requires_grad_(True)
        kp = torch.randn(B, Hkv, S, d, device=dev).bfloat16().float().requires_grad_(True)
        vp = torch.randn(B, Hkv, S, d, device=dev).bfloat16().float().requires_grad_(True)
        inv = 1.0 / (10000 ** (torch.arange(0, d, 2, device=dev).float() / d))
        ang = torch.arange(S, device=dev).float()[:, None] * inv[None]
        cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
        c = torch.cat([cos, cos], -1)[None, None]
        s_ = torch.cat([sin, sin], -1)[None, None]

        def rot(t):
            if not use_rope:
                return t
            t1, t2 = t.chunk(2, -1)
            return t * c + torch.cat([-t2, t1], -1) * s_

        qr, kr = rot(qp), rot(kp)
        kpm = None
        if pad is not None:
            kpm = torch.ones(B, S, device=dev, dtype=torch.uint8)
            for bi in range(B):
                n = max(1, S - 3 - 5 * bi)
                if pad == "right":
                    kpm[bi, n:] = 0
                else:
                    kpm[bi, : S - n] = 0
        kf = kr.repeat_interleave(n_rep, 1)
        vf = vp.repeat_interleave(n_rep, 1)
        sc = qr @ kf.transpose(-1, -2) / 8.0
        vis = torch.ones(B, 1, S, S, device=dev)
        if causal:
            vis = vis * torch.tril(torch.ones(S, S, device=dev))[None, None]
        if kpm is not None:
            vis = vis * kpm.float()[:, None, None, :]
        sc = sc + (1.0 - vis) * torch.finfo(torch.float32).min
        o_ref = (torch.softmax(sc, -1) @ vf).permute(0, 2, 1, 3).reshape(B, S, Hq * d)
        dout = torch.randn(B, S, Hq * d, device=dev).bfloat16()
        o_ref.backward(dout.float())

        q16, k16, v16 = qr.detach().bfloat16(), kr.detach().bfloat16(), vp.detach().bfloat16()
        o, lse = ops.attn_fwd(q16, k16, v16, causal=causal, q_pos0=0, key_padding_mask=kpm, need_lse=True)
        N = (Hq + 2 * Hkv) * d
        dqkv = torch.zeros(B * S, N, device=dev, dtype=torch.float32)
        ops.attn_bwd(q16, k16, v16, o, dout, lse, causal=causal, q_pos0=0, key_padding_mask=kpm,
                     rope_cos=cos if use_rope else None, rope_sin=sin if use_rope else None,
                     dq=dqkv[:, : Hq * d], dk=dqkv[:, Hq * d:(Hq + Hkv) * d], dv=dqkv[:, (Hq + Hkv) * d:])
        g = dqkv.view(B, S, Hq + 2 * Hkv, d).permute(0, 2, 1, 3)
        tag = f"B{B} Hq{Hq} Hkv{Hkv} S{S} causal{int(causal)} pad={pad} rope{int(use_rope)}"
        ok &= report(f"dq {tag}", g[:, :Hq], qp.grad, 2e-2)
        ok &= report("   dk", g[:, Hq:Hq + Hkv], kp.grad, 2e-2)
        ok &= report("   dv", g[:, Hq + Hkv:], vp.grad, 2e-2)
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default=None)
    ap.add_argument("--only", default=None, help="comma separated groups for the driver mode")
    ap.add_argument("--timeout", type=int, default=240)
    args = ap.parse_args()
    if args.group:
        import torch
        t0 = time.time()
        ok = GROUPS[args.group]()
        torch.cuda.synchronize()
        from vyomai_b200 import _lib
        poisoned = _lib.lib().vy_gemm_poisoned()
        if poisoned != 0:
            print(f"  [FAIL] vy_gemm_poisoned() = {poisoned}", flush=True)
            ok = False
        print(f"GROUP {args.group}: {'PASS' if ok else 'FAIL'} ({time.time() - t0:.1f}s)", flush=True)
        sys.exit(0 if ok else 1)
    names = args.only.split(",") if args.only else list(GROUPS)
    results = {}
    for n in names:
        print(f"=== {n}", flush=True)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", n], timeout=args.timeout)
            results[n] = p.returncode
        except subprocess.TimeoutExpired:
            results[n] = "timeout"
            print(f"GROUP {n}: TIMEOUT", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(results, open("gpurun_out/selftest.json", "w"), indent=1)
    print("SUMMARY", json.dumps(results))
    sys.exit(0 if all(v == 0 for v in results.values()) else 1)


if __name__ == "__main__":
    main()
