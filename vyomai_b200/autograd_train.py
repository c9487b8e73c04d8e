"""Backward passes of the fused blocks (training path).

Every gradient below is a sequence of C-ABI kernel calls:
  * input gradients  : vy_gemm with the weight read MN-major (no transposed copy), residual
                       gradients folded in through the epilogue's addend / addend2;
  * weight gradients : vy_gemm with BOTH operands MN-major (dW = dY^T X straight from row-major dY, X);
  * bias gradients   : vy_colsum;   LayerNorm: vy_add_layernorm_bwd;   GELU': dgrad epilogue / vy_act_bwd;
  * attention        : vy_attn_bwd (dq/dk/dv with RoPE undone, written into the packed dqkv);
  * embeddings       : vy_embed_bwd scatter-add.
torch.autograd only orders the calls and accumulates `.grad`.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from . import functional as F
from .functional import MaskSpec


def _wgrad(dy2d: torch.Tensor, x2d: torch.Tensor, like: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """dW[out, in] = dY^T X: both operands are read MN-major from their row-major storage."""
    return ops.gemm(dy2d.t(), x2d.t(), out_dtype=like.dtype, out_scale=scale, allow_split_k=True)


def _dgrad(dy2d: torch.Tensor, w: torch.Tensor, **kw) -> torch.Tensor:
    """dX = dY W: the nn.Linear weight [out, in] is the MN-major B operand. A long reduction (the LM head: K = vocabulary)
    may be split over K when the output has too few tiles for the SMs."""
    return ops.gemm(dy2d, w.t(), allow_split_k=dy2d.shape[1] >= 4096, **kw)


def _bgrad(dy2d: torch.Tensor, like: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if like is None:
        return None
    return ops.colsum(dy2d, out_dtype=like.dtype)


# ---- gradient accumulation fused into the kernels -----------------------------------------------
# A trainer that owns preallocated gradient buffers (trainer.FlatParams) marks its parameters with
# `_vy_direct_grad`; weight / bias gradients are then accumulated straight into `param.grad` by the
# wgrad GEMM epilogue (addend = the buffer itself) or vy_colsum(accumulate), autograd receives None for
# them, and `_vy_grad_ready` (the trainer's bucket hook) is called by hand. Saves one read-modify-write
# pass and one launch per parameter.
def _direct(p: Optional[torch.Tensor]) -> bool:
    return p is not None and getattr(p, "_vy_direct_grad", False) and p.grad is not None


def _overwrite(p: Optional[torch.Tensor]) -> bool:
    """The trainer guarantees this parameter's gradient is produced exactly once per step by one of the kernels below,
    so it may be WRITTEN instead of accumulated (no zero fill of the buffer, no read-modify-write in the epilogue)."""
    return p is not None and getattr(p, "_vy_grad_overwrite", False)


def _ready(p) -> None:
    cb = getattr(p, "_vy_grad_ready", None)
    if cb is not None:
        cb(p)


def _emit_wgrad(p: torch.Tensor, dy2d: torch.Tensor, x2d: torch.Tensor) -> Optional[torch.Tensor]:
    if _direct(p):
        g2 = p.grad.view(p.grad.shape[0], -1)
        ops.gemm(dy2d.t(), x2d.t(), out=g2, addend=None if _overwrite(p) else g2, allow_split_k=True)
        _ready(p)
        return None
    return _wgrad(dy2d, x2d, p).view(p.shape)


def _side_ok(p) -> bool:
    """Side-stream column sums need a consumer that joins before it reads the gradient: the trainer's optimizer step does
    (Trainer.optimizer_step); the NCCL bucket hooks fire at `_ready`, i.e. too early, so that mode keeps the main stream."""
    return getattr(p, "_vy_side_bgrad", False)


def _emit_bgrad(p: Optional[torch.Tensor], dy2d: torch.Tensor, after=None) -> Optional[torch.Tensor]:
    """`after`: ops.side_fork_point() taken when dy2d was complete — the sum then runs on the side stream (the caller joins
    before it returns)."""
    if p is None:
        return None
    if _direct(p):
        if after is not None and _side_ok(p):
            ops.colsum_side(after, dy2d, out=p.grad, accumulate=not _overwrite(p))
        else:
            ops.colsum(dy2d, out=p.grad, accumulate=not _overwrite(p))
        _ready(p)
        return None
    return ops.colsum(dy2d, out_dtype=p.dtype)


def _ln_bwd(dy: torch.Tensor, s: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, mean, rstd,
            bias: Optional[torch.Tensor] = None, dropout=None):
    """LayerNorm backward. Returns (ds, d_gamma, d_beta, d_bias); `bias` is the bias of the Linear whose output was
    normalised (its gradient = column sums of that Linear's output gradient, produced by the same kernels). With
    trainer-owned gradient buffers the parameter gradients are accumulated in place by the reduce kernel and None is
    returned for them. With `dropout` (the forward's DropoutState) ds is the pair (gradient of the pre-norm sum = of the
    residual, gradient of the dropped Linear output)."""
    direct = _direct(gamma) and _direct(beta) and gamma.grad.dtype == beta.grad.dtype
    if direct and (bias is None or (_direct(bias) and bias.grad.dtype == gamma.grad.dtype)):
        over = _overwrite(gamma) and _overwrite(beta) and (bias is None or _overwrite(bias))
        if not over and (_overwrite(gamma) or _overwrite(beta) or _overwrite(bias)):
            raise RuntimeError("a LayerNorm's weight / bias and the preceding Linear's bias must share the gradient write mode")
        ds = ops.add_layernorm_bwd(dy, s, gamma, mean, rstd, dgamma_out=gamma.grad, dbeta_out=beta.grad,
                                   dbias_out=bias.grad if bias is not None else None, accumulate=not over, dropout=dropout)[0]
        _ready(gamma)
        _ready(beta)
        if bias is not None:
            _ready(bias)
        return ds, None, None, None
    if bias is not None:
        ds, dgamma, dbeta, dbias = ops.add_layernorm_bwd(dy, s, gamma, mean, rstd, want_dbias=True, dropout=dropout)
        return ds, dgamma.to(gamma.dtype), dbeta.to(beta.dtype), dbias.to(bias.dtype)
    ds, dgamma, dbeta = ops.add_layernorm_bwd(dy, s, gamma, mean, rstd, dropout=dropout)
    return ds, dgamma.to(gamma.dtype), dbeta.to(beta.dtype), None


def _split_ds(ds):
    """(gradient of the residual, gradient of the Linear output) from _ln_bwd's first result."""
    return ds if isinstance(ds, tuple) else (ds, ds)


def _packed_grads(params) -> Optional[torch.Tensor]:
    """One 2-D (or 1-D) view over the adjacent .grad buffers of several parameters, or None."""
    if not all(_direct(p) for p in params):
        return None
    gs = [p.grad for p in params]
    if not F._adjacent(gs):
        return None
    g0 = gs[0]
    n = sum(g.shape[0] for g in gs)
    if g0.dim() == 2:
        return torch.as_strided(g0, (n, g0.shape[1]), (g0.stride(0), 1), g0.storage_offset())
    return torch.as_strided(g0, (n,), (1,), g0.storage_offset())


class AttentionBlockFn(torch.autograd.Function):
    """y = LN(dense(attention(x)) + x). Inputs after the non-tensor arguments: x2d, the 1 or 3
    projection weights, their biases (if any), dense.weight, [dense.bias], ln.weight, ln.bias."""

    @staticmethod
    def forward(ctx, mod, B, S, mask: MaskSpec, rope, start_pos, x2d, *params):
        lin = mod._packed()
        dense, ln = mod.out.dense, mod.out.layernorm
        w_qkv, b_qkv = F.pack_linears(lin)
        attn, saved = F.attention_core(x2d, B, S, w_qkv, b_qkv, mod.num_attention_heads, mod._kv_heads, mask, rope,
                                       None, False, need_lse=True, pos0=start_pos)
        q, k, v, lse = saved
        ctx.drop = F.dropout_state(mod.out, mod.out.dropout.p)
        y, (s, mean, rstd) = F.self_output(attn, x2d, dense, ln, save=True, dropout=ctx.drop)
        ctx.mod, ctx.B, ctx.S, ctx.mask, ctx.rope, ctx.start_pos = mod, B, S, mask, rope, start_pos
        ctx.n_w = len(lin)
        ctx.has_qkv_bias = lin[0].bias is not None
        ctx.has_dense_bias = dense.bias is not None
        ctx.save_for_backward(x2d, q, k, v, attn, lse, s, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, q, k, v, attn, lse, s, mean, rstd = ctx.saved_tensors
        mod, B, S, mask, rope = ctx.mod, ctx.B, ctx.S, ctx.mask, ctx.rope
        lin = mod._packed()
        dense, ln = mod.out.dense, mod.out.layernorm
        w_qkv, _ = F.pack_linears(lin)
        Hq, Hkv, d = mod.num_attention_heads, mod._kv_heads, F.HEAD_DIM
        dy = dy.contiguous()
        ds, dgamma, dbeta, d_bo = _ln_bwd(dy, s, ln.weight, ln.bias, mean, rstd, bias=dense.bias, dropout=ctx.drop)
        ds, dxo = _split_ds(ds)  # residual gradient, gradient of dense(attn) (the dropped branch)
        d_wo = _emit_wgrad(dense.weight, dxo, attn)
        d_attn = _dgrad(dxo, dense.weight, out_dtype=torch.bfloat16)  # bf16: MMA operand of the attention backward
        dqkv = torch.empty((B * S, (Hq + 2 * Hkv) * d), device=dy.device, dtype=x2d.dtype)
        cos = sin = None
        rope_pos0 = 0
        if rope is not None:
            cos, sin, base = rope
            rope_pos0 = 0 if base is None else ctx.start_pos - base  # table row of the first token
        ops.attn_bwd(q, k, v, attn, d_attn, lse, causal=mask.causal, q_pos0=mask.q_pos0, key_padding_mask=mask.key_padding,
                     rope_cos=cos, rope_sin=sin, rope_pos0=rope_pos0, dq=dqkv[:, : Hq * d],
                     dk=dqkv[:, Hq * d:(Hq + Hkv) * d], dv=dqkv[:, (Hq + Hkv) * d:])
        fork = ops.side_fork_point()
        dx = _dgrad(dqkv, w_qkv, addend=ds)  # + the residual branch of LN(dense(.) + x)
        grads = [dx]
        gw = _packed_grads([l.weight for l in lin])
        if gw is not None:  # one wgrad GEMM into the adjacent q|k|v gradient buffers
            ops.gemm(dqkv.t(), x2d.t(), out=gw, addend=None if all(_overwrite(l.weight) for l in lin) else gw, allow_split_k=True)
            for l in lin:
                _ready(l.weight)
                grads.append(None)
        else:
            d_wqkv = _wgrad(dqkv, x2d, w_qkv)
            r = 0
            for l in lin:
                n = l.weight.shape[0]
                grads.append(d_wqkv[r:r + n])
                r += n
        if ctx.has_qkv_bias:
            gb = _packed_grads([l.bias for l in lin])
            if gb is not None:
                acc = not all(_overwrite(l.bias) for l in lin)
                if fork is not None and all(_side_ok(l.bias) for l in lin):
                    ops.colsum_side(fork, dqkv, out=gb, accumulate=acc)
                else:
                    ops.colsum(dqkv, out=gb, accumulate=acc)
                for l in lin:
                    _ready(l.bias)
                    grads.append(None)
            else:
                d_bqkv = ops.colsum(dqkv, out_dtype=lin[0].bias.dtype)
                r = 0
                for l in lin:
                    n = l.bias.shape[0]
                    grads.append(d_bqkv[r:r + n])
                    r += n
        grads.append(d_wo)
        if ctx.has_dense_bias:
            grads.append(d_bo)
        grads += [dgamma, dbeta]
        ops.side_join()
        return (None, None, None, None, None, None, *grads)


class CrossAttentionBlockFn(torch.autograd.Function):
    """y = LN(dropout(dense(attention(q(x), k(enc), v(enc)))) + x) — cross-attention (layers/attention.py:382-573). Inputs
    after the non-tensor arguments: x2d, enc2d, Wq, bq, Wk, bk, Wv, bv, Wo, bo, ln.weight, ln.bias."""

    @staticmethod
    def forward(ctx, mod, B, Sq, Skv, mask, drop, x2d, enc2d, wq, bq, wk, bk, wv, bv, wo, bo, gamma, beta):
        attn, (q, k, v, lse) = F.cross_attention_core(x2d, B, Sq, enc2d, Skv, mod.query, mod.key, mod.value, mod.num_attention_heads,
                                                       mod._kv_heads, mask, None, need_lse=True)
        y, (s, mean, rstd) = F.self_output(attn, x2d, mod.out.dense, mod.out.layernorm, save=True, dropout=drop)
        ctx.mod, ctx.B, ctx.Sq, ctx.Skv, ctx.mask, ctx.drop = mod, B, Sq, Skv, mask, drop
        ctx.save_for_backward(x2d, enc2d, q, k, v, attn, lse, s, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, enc2d, q, k, v, attn, lse, s, mean, rstd = ctx.saved_tensors
        mod, B, Sq, Skv, mask = ctx.mod, ctx.B, ctx.Sq, ctx.Skv, ctx.mask
        dense, ln = mod.out.dense, mod.out.layernorm
        Hq, Hkv, d = mod.num_attention_heads, mod._kv_heads, F.HEAD_DIM
        ds, dgamma, dbeta, d_bo = _ln_bwd(dy.contiguous(), s, ln.weight, ln.bias, mean, rstd, bias=dense.bias, dropout=ctx.drop)
        ds, dxo = _split_ds(ds)
        d_wo = _emit_wgrad(dense.weight, dxo, attn)
        d_attn = _dgrad(dxo, dense.weight, out_dtype=torch.bfloat16)
        dq = torch.empty((B * Sq, Hq * d), device=dy.device, dtype=x2d.dtype)
        dkv = torch.empty((B * Skv, 2 * Hkv * d), device=dy.device, dtype=x2d.dtype)
        ops.attn_bwd(q, k, v, attn, d_attn, lse, causal=False, q_pos0=0, key_padding_mask=mask.key_padding, rope_cos=None,
                     rope_sin=None, dq=dq, dk=dkv[:, : Hkv * d], dv=dkv[:, Hkv * d:])
        dx = _dgrad(dq, mod.query.weight, addend=ds)  # + the residual branch
        w_kv, _ = F.pack_linears([mod.key, mod.value])
        d_enc = _dgrad(dkv, w_kv)
        d_wq = _wgrad(dq, x2d, mod.query.weight)
        d_wkv = _wgrad(dkv, enc2d, w_kv)
        n = Hkv * d
        d_bq = _bgrad(dq, mod.query.bias)
        d_bkv = _bgrad(dkv, mod.key.bias)
        d_bk = d_bkv[:n] if d_bkv is not None else None
        d_bv = d_bkv[n:] if d_bkv is not None else None
        return (None, None, None, None, None, None, dx, d_enc, d_wq, d_bq, d_wkv[:n], d_bk, d_wkv[n:], d_bv, d_wo, d_bo, dgamma, dbeta)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b (nn.Linear) on vy_gemm with its dgrad / wgrad / bias-gradient kernels."""

    @staticmethod
    def forward(ctx, x2d, w, b):
        ctx.save_for_backward(x2d, w)
        ctx.bdt = None if b is None else b.dtype
        return F._lin(x2d, w, b)

    @staticmethod
    def backward(ctx, dy):
        x2d, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = _dgrad(dy, w) if ctx.needs_input_grad[0] else None
        dw = _wgrad(dy, x2d, w) if ctx.needs_input_grad[1] else None
        db = ops.colsum(dy, out_dtype=ctx.bdt) if (ctx.bdt is not None and ctx.needs_input_grad[2]) else None
        return dx, dw, db


class SelfOutputFn(torch.autograd.Function):
    """y = LN(dense(attn) + residual) (AttentionSelfOutput used stand-alone)."""

    @staticmethod
    def forward(ctx, attn2d, residual2d, w, b, gamma, beta, eps, dropout=None):
        if dropout is None:
            s = F._lin(attn2d, w, b, addend=residual2d)
            y, _, mean, rstd = ops.add_layernorm(s, None, gamma, beta, eps, save_stats=True)
        else:
            y, s, mean, rstd = ops.add_layernorm(F._lin(attn2d, w, b), residual2d.contiguous(), gamma, beta, eps, save_stats=True,
                                                 save_sum=True, dropout=dropout)
        ctx.drop = dropout
        ctx.has_bias = b is not None
        ctx.save_for_backward(attn2d, w, gamma, s, mean, rstd)
        ctx.bdt = b.dtype if b is not None else None
        return y

    @staticmethod
    def backward(ctx, dy):
        attn2d, w, gamma, s, mean, rstd = ctx.saved_tensors
        ds, dgamma, dbeta = ops.add_layernorm_bwd(dy.contiguous(), s, gamma, mean, rstd, dropout=ctx.drop)
        ds, dxo = _split_ds(ds)
        d_w = _wgrad(dxo, attn2d, w)
        d_b = ops.colsum(dxo, out_dtype=ctx.bdt) if ctx.has_bias else None
        d_attn = _dgrad(dxo, w)
        return d_attn, ds, d_w, d_b, dgamma.to(gamma.dtype), dbeta.to(gamma.dtype), None, None


class FeedForwardFn(torch.autograd.Function):
    """y = LN(W2 act(W1 h + b1) + b2 + input)."""

    @staticmethod
    def forward(ctx, act, eps, dropout, h2d, input2d, w1, b1, w2, b2, gamma, beta):
        z = torch.empty((h2d.shape[0], w1.shape[0]), device=h2d.device, dtype=h2d.dtype)
        a = F._lin(h2d, w1, b1, act=act, aux=z)
        if dropout is None:
            s = F._lin(a, w2, b2, addend=input2d)
            y, _, mean, rstd = ops.add_layernorm(s, None, gamma, beta, eps, save_stats=True)
        else:
            y, s, mean, rstd = ops.add_layernorm(F._lin(a, w2, b2), input2d.contiguous(), gamma, beta, eps, save_stats=True,
                                                 save_sum=True, dropout=dropout)
        ctx.drop = dropout
        ctx.act = act
        ctx.params = (w1, b1, w2, b2, gamma, beta)  # the Parameter objects (direct-gradient attributes live on them)
        ctx.save_for_backward(h2d, z, a, s, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        h2d, z, a, s, mean, rstd = ctx.saved_tensors
        w1, b1, w2, b2, gamma, beta = ctx.params
        ds, dgamma, dbeta, d_b2 = _ln_bwd(dy.contiguous(), s, gamma, beta, mean, rstd, bias=b2, dropout=ctx.drop)
        ds, dxo = _split_ds(ds)  # gradient of input_tensor, gradient of the second Linear's output
        d_w2 = _emit_wgrad(w2, dxo, a)
        dz = _dgrad(dxo, w2, act="d" + ctx.act, aux=z)  # (dS W2) * act'(z) in the dgrad epilogue
        fork = ops.side_fork_point()
        d_w1 = _emit_wgrad(w1, dz, h2d)
        d_b1 = _emit_bgrad(b1, dz, after=fork)  # column sums of dz under the weight-gradient GEMM
        dh = _dgrad(dz, w1)
        ops.side_join()
        return None, None, None, dh, ds, d_w1, d_b1, d_w2, d_b2, dgamma, dbeta


def _padded_logits(rows: int, V: int, like: torch.Tensor):
    """[rows, V] view of a buffer whose row stride is V rounded up to 8 elements (16-byte aligned rows)."""
    ld = (V + 7) // 8 * 8
    buf = torch.empty((rows, ld), device=like.device, dtype=like.dtype)
    return buf[:, :V]


def _lm_head_backward(ctx_saved, params, dlogits, bias_part=None, gout=None):
    """`bias_part`: the per-CTA column sums of dlogits vy_softmax_xent took while writing it (for an upstream gradient of 1;
    `gout`, a device scalar, rescales them) — the 0.8 GB gradient is then not read a third time for the bias."""
    h2d, z, a, n, mean, rstd = ctx_saved
    wd, bd, gamma, beta, wv, bv = params
    fork = ops.side_fork_point()
    d_wv = _emit_wgrad(wv, dlogits, n)
    if bv is not None and bias_part is not None:
        if _direct(bv):
            ops.colsum_finish(bias_part, out=bv.grad, accumulate=not _overwrite(bv), scale_ptr=gout)
            _ready(bv)
            d_bv = None
        else:
            d_bv = ops.colsum_finish(bias_part, out_dtype=bv.dtype, scale_ptr=gout)
    else:
        d_bv = _emit_bgrad(bv, dlogits, after=fork)  # 0.8 GB of dlogits summed under the 0.5 ms vocabulary wgrad GEMM
    dn = _dgrad(dlogits, wv)
    da, dgamma, dbeta, _ = _ln_bwd(dn, a, gamma, beta, mean, rstd)
    dz = ops.act_bwd(da, z, "gelu")
    d_wd = _emit_wgrad(wd, dz, h2d)
    d_bd = _emit_bgrad(bd, dz)
    dh = _dgrad(dz, wd)
    ops.side_join()
    return dh, d_wd, d_bd, dgamma, dbeta, d_wv, d_bv


class LMHeadFn(torch.autograd.Function):
    """logits = decoder(LN(gelu(dense(h))))."""

    @staticmethod
    def forward(ctx, eps, h2d, wd, bd, gamma, beta, wv, bv):
        z = torch.empty((h2d.shape[0], wd.shape[0]), device=h2d.device, dtype=h2d.dtype)
        a = F._lin(h2d, wd, bd, act="gelu", aux=z)
        n, _, mean, rstd = ops.add_layernorm(a, None, gamma, beta, eps, save_stats=True)
        logits = F._lin(n, wv, bv, out=_padded_logits(h2d.shape[0], wv.shape[0], h2d))
        ctx.params = (wd, bd, gamma, beta, wv, bv)
        ctx.save_for_backward(h2d, z, a, n, mean, rstd)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        if dlogits.stride(1) != 1 or (dlogits.stride(0) * dlogits.element_size()) % 16 != 0:
            buf = _padded_logits(dlogits.shape[0], dlogits.shape[1], dlogits)
            buf.copy_(dlogits)
            dlogits = buf
        return (None, *_lm_head_backward(ctx.saved_tensors, ctx.params, dlogits))


class LMHeadLossFn(torch.autograd.Function):
    """loss = mean token cross-entropy(decoder(LN(gelu(dense(h)))), labels) over labels != ignore_index: the LM
    head (models/decoder.py:267-275) and the notebooks' `loss_fn` as ONE autograd node, so the [rows, V] logits
    exist exactly once: vy_softmax_xent computes the row losses and overwrites the logits with d loss / d logits in
    the same launch (the upstream gradient of a training loss is 1; backward rescales only if it is not), and the
    backward GEMMs read that buffer in place — no slice / as_strided gradient copies of an 800 MB tensor."""

    @staticmethod
    def forward(ctx, eps, ignore_index, h2d, wd, bd, gamma, beta, wv, bv, labels):
        z = torch.empty((h2d.shape[0], wd.shape[0]), device=h2d.device, dtype=h2d.dtype)
        a = F._lin(h2d, wd, bd, act="gelu", aux=z)
        n, _, mean, rstd = ops.add_layernorm(a, None, gamma, beta, eps, save_stats=True)
        logits = F._lin(n, wv, bv, out=_padded_logits(h2d.shape[0], wv.shape[0], h2d))
        labels = labels.reshape(-1).contiguous()
        n_valid = (labels != ignore_index).sum().clamp(min=1).to(torch.float32)
        inv = (1.0 / n_valid).reshape(1)
        part = ops.xent_colsum_part(logits) if (bv is not None and ctx.needs_input_grad[8]) else None
        loss_rows = ops.softmax_xent(logits, labels, ignore_index=ignore_index, grad_scale_ptr=inv, write_grad=True, colsum_part=part)
        ctx.params = (wd, bd, gamma, beta, wv, bv)
        ctx.has_part = part is not None
        ctx.save_for_backward(h2d, z, a, n, mean, rstd, logits, *([part] if part is not None else []))
        return loss_rows.sum() * inv[0]

    @staticmethod
    def backward(ctx, gout):
        saved = list(ctx.saved_tensors)
        part = saved.pop() if ctx.has_part else None
        dlogits = saved.pop()
        # dlogits already holds d loss / d logits for an upstream gradient of 1 (what loss.backward() passes); any other
        # upstream gradient (a scaled / accumulated loss) is applied on the device — the kernel returns at once for 1
        ld = dlogits.stride(0)  # the row-padded buffer behind the [:, :V] view
        g32 = gout.to(torch.float32).reshape(1)
        ops.scale_by_ptr(torch.as_strided(dlogits, (dlogits.shape[0], ld), (ld, 1), dlogits.storage_offset()), g32)
        grads = _lm_head_backward(saved, ctx.params, dlogits, bias_part=part, gout=g32)
        return (None, None, *grads, None)


class CrossEntropyFn(torch.autograd.Function):
    """mean token cross-entropy over rows whose label != ignore_index. The backward recomputes the
    softmax and OVERWRITES the logits buffer with d loss / d logits (the buffer is dead by then), so no
    second [rows, V] tensor is ever allocated."""

    @staticmethod
    def forward(ctx, logits2d, labels, ignore_index):
        labels = labels.contiguous()
        loss_rows = ops.softmax_xent(logits2d, labels, ignore_index=ignore_index, write_grad=False)
        n_valid = (labels != ignore_index).sum().clamp(min=1).to(torch.float32)
        ctx.save_for_backward(logits2d, labels, n_valid)
        ctx.ignore_index = ignore_index
        return loss_rows.sum() / n_valid

    @staticmethod
    def backward(ctx, gout):
        logits2d, labels, n_valid = ctx.saved_tensors
        scale = (gout.to(torch.float32) / n_valid).reshape(1).contiguous()
        ops.softmax_xent(logits2d, labels, ignore_index=ctx.ignore_index, grad_scale_ptr=scale, write_grad=True)
        return logits2d, None, None


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """F.cross_entropy(logits.view(-1, V), labels.view(-1), ignore_index=...) on the fused kernel. `logits`
    must be the (possibly row-padded) tensor returned by the LM head; it is consumed by backward."""
    V = logits.shape[-1]
    if logits.dim() == 3:
        B, S, _ = logits.shape
        if logits.stride(2) != 1 or logits.stride(0) != S * logits.stride(1):
            raise ValueError("cross_entropy: logits must be the LM head's output (rows with one stride)")
        logits2d = torch.as_strided(logits, (B * S, V), (logits.stride(1), 1), logits.storage_offset())
    else:
        logits2d = logits
    return CrossEntropyFn.apply(logits2d, labels.reshape(-1), ignore_index)


class EmbedFn(torch.autograd.Function):
    """hidden rows = table[ids] (+ pos rows), optionally with one extra leading row per sequence taken
    from `extra` (the captioner's image vector, models/multimodel.py:163-166)."""

    @staticmethod
    def forward(ctx, ids, table, pos_table, pos_row_off, tokens_per_seq, extra, padding_idx=None, pos_padding_idx=None):
        ctx.pad = (padding_idx, pos_padding_idx)
        B = ids.numel() // tokens_per_seq
        e = 0 if extra is None else 1
        seq = tokens_per_seq + e
        out = torch.empty((B * seq, table.shape[1]), device=table.device, dtype=table.dtype)
        idsf = ids.reshape(-1).contiguous()
        ops.embed(idsf, table, out=out, tokens_per_seq=tokens_per_seq, out_group_stride=seq, out_row_off=e, pos=pos_table,
                  pos_row_off=pos_row_off + e)
        if extra is not None:
            ops.embed(None, extra, out=out, rows=B, tokens_per_seq=1, out_group_stride=seq, out_row_off=0, pos=pos_table,
                      pos_row_off=pos_row_off)
        ctx.save_for_backward(idsf, table, pos_table if pos_table is not None else table.new_empty(0))
        ctx.table_param = table
        ctx.pos_param = pos_table
        ctx.has_pos = pos_table is not None
        ctx.args = (pos_row_off, tokens_per_seq, e, seq, B)
        ctx.extra_shape = None if extra is None else (extra.shape, extra.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        idsf, table, pos_table = ctx.saved_tensors
        pos_row_off, tps, e, seq, B = ctx.args
        dout = dout.contiguous()
        H = table.shape[1]
        d_table = d_pos = d_extra = None
        need_pos = ctx.has_pos and ctx.needs_input_grad[2]
        tab_direct = ctx.needs_input_grad[1] and _direct(ctx.table_param)
        pos_direct = need_pos and _direct(ctx.pos_param)
        if ctx.needs_input_grad[1]:
            d_table = ctx.table_param.grad if tab_direct else torch.zeros_like(table)  # scatter-add straight into .grad
        if need_pos:
            d_pos = ctx.pos_param.grad.view(pos_table.shape) if pos_direct else torch.zeros_like(pos_table)
        if d_table is not None or d_pos is not None:
            ops.embed_bwd(idsf, dout, rows=idsf.numel(), H=H, tokens_per_seq=tps, out_group_stride=seq, out_row_off=e,
                          dtable=d_table, dpos=d_pos, pos_row_off=pos_row_off + e, padding_idx=ctx.pad[0],
                          pos_padding_idx=ctx.pad[1])
        if e:
            if d_pos is not None:  # the extra row's position gradient
                ops.embed_bwd(None, dout, rows=B, H=H, tokens_per_seq=1, out_group_stride=seq, out_row_off=0,
                              dtable=None, dpos=d_pos, pos_row_off=pos_row_off, pos_padding_idx=ctx.pad[1])
            if ctx.needs_input_grad[5]:
                shape, dt = ctx.extra_shape
                d_extra = torch.empty(shape, device=dout.device, dtype=dout.dtype)
                ops.embed(None, dout, out=d_extra, rows=B, tokens_per_seq=1, out_group_stride=1, out_row_off=0,
                          src_row_stride=seq * dout.stride(0))
                d_extra = d_extra.to(dt)
        if tab_direct:
            _ready(ctx.table_param)
            d_table = None
        if pos_direct:
            _ready(ctx.pos_param)
            d_pos = None
        return None, d_table, d_pos, None, None, d_extra, None, None


class VitStemFn(torch.autograd.Function):
    """hidden = 2 * (cat(cls, conv(pixels)) + pos) (quirk Q1); gradients for the conv weight / bias,
    cls_token and the position table (pixels get none — they are data)."""

    @staticmethod
    def forward(ctx, mod, pixels, w4, bias, cls, pos):
        from .autograd import vit_stem
        hidden, patches = vit_stem(mod, pixels)
        ctx.mod = mod
        ctx.B = pixels.shape[0]
        ctx.save_for_backward(patches, w4, bias, cls, pos)
        return hidden

    @staticmethod
    def backward(ctx, dh):
        patches, w4, bias, cls, pos = ctx.saved_tensors
        mod, B = ctx.mod, ctx.B
        nP, H = mod.num_patches, w4.shape[0]
        dh = dh.contiguous()
        dh3 = dh.view(B, nP + 1, H)
        dp = ops.cast4d(dh3[:, 1:, :].unsqueeze(0), dh.dtype).view(B * nP, H)  # patch rows, contiguous
        d_w = _wgrad(dp, patches, w4, scale=2.0).view(w4.shape)
        d_b = ops.colsum(dp, out_dtype=bias.dtype, scale=2.0) if bias is not None else None
        d_pos = ops.colsum(dh.view(B, (nP + 1) * H), out_dtype=pos.dtype, scale=2.0).view(pos.shape)
        d_cls = d_pos.view(nP + 1, H)[0].to(cls.dtype).view(cls.shape).clone()
        return None, None, d_w, d_b, d_cls, d_pos
