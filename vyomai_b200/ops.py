"""Tensor-level wrappers over the C ABI (include/vyom_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; every function below
passes raw pointers / sizes / strides to libvyom_b200.so through ctypes. Nothing in this module
computes with torch ops, and nothing falls back to them.
"""
from __future__ import annotations

from typing import Optional, Tuple

import ctypes
import os

import torch

from . import _lib, gemm_tune
from ._lib import CONSTS as C

_DT = {torch.float32: C["VY_F32"], torch.bfloat16: C["VY_BF16"]}

NORM_KIND = {"layernorm": C["VY_NORM_LAYER"], "rmsnorm": C["VY_NORM_RMS"], "gemma_rmsnorm": C["VY_NORM_RMS_GEMMA"]}

ACT = {
    None: C["VY_ACT_NONE"],
    "none": C["VY_ACT_NONE"],
    "gelu": C["VY_ACT_GELU_ERF"],
    "swiglu": C["VY_ACT_SWIGLU"],
    "geglu_tanh": C["VY_ACT_GEGLU_TANH"],
    "gelu_erf": C["VY_ACT_GELU_ERF"],
    "gelu_tanh": C["VY_ACT_GELU_TANH"],
    "dgelu": C["VY_ACT_DGELU_ERF"],
    "dgelu_erf": C["VY_ACT_DGELU_ERF"],
    "dgelu_tanh": C["VY_ACT_DGELU_TANH"],
}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise _lib.VyomError(f"unsupported dtype {t.dtype}: the sm_100a path takes float32 or bfloat16") from None


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.VyomError(
                "vyomai_b200 has no CPU path: tensors must live on a CUDA (sm_100a) device; got a "
                f"{t.device} tensor"
            )


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _major(t: torch.Tensor, what: str):
    """(is_mn_major, leading stride) of a logical (rows, k) operand, from its torch strides."""
    assert t.dim() == 2, what
    if t.stride(1) == 1:
        return 0, t.stride(0)
    if t.stride(0) == 1:
        return 1, t.stride(1)
    raise _lib.VyomError(f"{what}: operand must have one unit stride, got strides {t.stride()}")


_SPLITK_WS = {}
SPLITK_WORKSPACE_BYTES = 128 << 20


def _splitk_workspace(device: torch.device) -> torch.Tensor:
    """Per-device fp32 scratch lent to vy_gemm for split-K (persistent, so captured CUDA graphs stay valid).
    All GEMMs of this process run on one stream at a time, which is what sharing one buffer requires."""
    t = _SPLITK_WS.get(device)
    if t is None:
        t = torch.empty(SPLITK_WORKSPACE_BYTES // 4, device=device, dtype=torch.float32)
        _SPLITK_WS[device] = t
    return t


def _small_batch(kw: dict) -> bool:
    """True when vy_gemm will route this call to the small-batch weight-streaming kernel (nothing to tune there)."""
    if not kw.get("transposed_out") or kw["N"] > 32:
        return False
    st = _lib.STRUCTS["VyGemm"]()
    for k, v in kw.items():
        if v is not None:
            setattr(st, k, v)
    return bool(_lib.lib().vy_gemm_is_small_batch(ctypes.byref(st)))


def gemm(
    a: torch.Tensor,
    b: torch.Tensor,
    *,
    bias: Optional[torch.Tensor] = None,
    act: Optional[str] = None,
    aux: Optional[torch.Tensor] = None,
    addend: Optional[torch.Tensor] = None,
    addend_row_mod: int = 0,
    addend_row_off: int = 0,
    addend2: Optional[torch.Tensor] = None,
    out: Optional[torch.Tensor] = None,
    out_dtype: Optional[torch.dtype] = None,
    out_scale: float = 1.0,
    out_row_group: int = 0,
    out_row_group_stride: int = 0,
    out_row_off: int = 0,
    swap_ab: bool = False,
    allow_split_k: bool = False,
) -> torch.Tensor:
    """out = epilogue(a @ b.T). `a` is logical (M, K), `b` logical (N, K); each may be stored with
    either index contiguous (K-major or MN-major — read off the torch strides, no copies).

    swap_ab=True computes the same logical result by feeding `b` as the 128-row MMA operand and
    writing the tile transposed — the decode-time path where M (tokens) is tiny and N large.
    """
    _need_cuda(a, b, bias, aux, addend, addend2, out)
    M, K = a.shape
    N, Kb = b.shape
    if K != Kb:
        raise _lib.VyomError(f"gemm: inner dimensions differ ({K} vs {Kb})")
    if a.dtype != b.dtype:
        raise _lib.VyomError("gemm: a and b must share a dtype")
    if out is None:
        odt = out_dtype or a.dtype
        rows = M if out_row_group == 0 else None
        if rows is None:
            raise _lib.VyomError("gemm: pass `out` when using a row-group remap")
        out = torch.empty((M, N // 2 if act in ("swiglu", "geglu_tanh") else N), device=a.device, dtype=odt)
    if act in ("swiglu", "geglu_tanh") and (swap_ab or N % 2 or out.shape[-1] != N // 2):
        raise _lib.VyomError("gemm: act='swiglu' / 'geglu_tanh' takes interleaved gate/up rows (even N), writes N / 2 columns, no swap_ab")
    if out.stride(-1) != 1:
        raise _lib.VyomError("gemm: out must be row-major")
    a_mn, lda = _major(a, "gemm a")
    b_mn, ldb = _major(b, "gemm b")
    kw = dict(
        in_dtype=_dt(a),
        epi=C["VY_EPI_LINEAR"],
        act=ACT[act],
        bias=_ptr(bias),
        bias_dtype=_dt(bias) if bias is not None else 0,
        addend=_ptr(addend),
        ld_addend=addend.stride(0) if addend is not None else 0,
        addend_dtype=_dt(addend) if addend is not None else 0,
        addend_row_mod=addend_row_mod,
        addend_row_off=addend_row_off,
        addend2=_ptr(addend2),
        ld_addend2=addend2.stride(0) if addend2 is not None else 0,
        addend2_dtype=_dt(addend2) if addend2 is not None else 0,
        aux=_ptr(aux),
        ld_aux=aux.stride(0) if aux is not None else 0,
        aux_dtype=_dt(aux) if aux is not None else 0,
        out_scale=float(out_scale),
        out=out.data_ptr(),
        ld_out=out.stride(0),
        out_dtype=_dt(out),
        out_row_group=out_row_group,
        out_row_group_stride=out_row_group_stride,
        out_row_off=out_row_off,
        stream=_stream(),
    )
    if allow_split_k:
        ws = _splitk_workspace(a.device)
        kw.update(workspace=ws.data_ptr(), workspace_bytes=ws.numel() * 4)
    if not swap_ab:
        kw.update(M=M, N=N, K=K, A=a.data_ptr(), lda=lda, a_mn_major=a_mn, B=b.data_ptr(), ldb=ldb,
                  b_mn_major=b_mn, transposed_out=0)
    else:
        # inference (no autograd): nothing in flight writes the weights, so the small-batch kernel may prefetch them under PDL
        kw.update(M=N, N=M, K=K, A=b.data_ptr(), lda=ldb, a_mn_major=b_mn, B=a.data_ptr(), ldb=lda,
                  b_mn_major=a_mn, transposed_out=1, weights_static=0 if torch.is_grad_enabled() else 1)
    if gemm_tune.ENABLED and not _small_batch(kw):
        key = ("gemm", kw["M"], kw["N"], K, kw["in_dtype"], kw["a_mn_major"], kw["b_mn_major"], kw["transposed_out"], act,
               bias is not None, addend is not None, addend is not None and addend.data_ptr() == out.data_ptr(),
               addend2 is not None, aux is not None, out.dtype, allow_split_k, out_row_group, addend_row_mod, out_scale != 1.0)
        saves_aux = aux is not None and act in ("gelu", "gelu_tanh", "swiglu", "geglu_tanh")
        kw.update(gemm_tune.hints(key, kw, [("out", out), ("aux", aux if saves_aux else None)], a.device))
    _lib.call("vy_gemm", "VyGemm", **kw)
    return out


def qkv_rope_gemm(
    x: torch.Tensor,
    w: torch.Tensor,
    bias: Optional[torch.Tensor],
    *,
    tokens_per_seq: int,
    start_pos: int,
    n_q_heads: int,
    n_kv_heads: int,
    head_dim: int,
    rope_cos: Optional[torch.Tensor],
    rope_sin: Optional[torch.Tensor],
    q_out: Optional[torch.Tensor],
    k_out: Optional[torch.Tensor],
    v_out: Optional[torch.Tensor],
    kv_dst_pos0: Optional[int] = None,
) -> None:
    """Fused q/k/v projection: bias + in-register RoPE + "b l (h d) -> b h l d" scatter + kv-cache
    append. q_out/k_out/v_out are 4-D [B, heads, tokens, head_dim] tensors (any batch/head/token
    strides, head_dim contiguous); k/v rows land at token index start_pos + l. With n_kv_heads == 0 only q is projected
    (k_out / v_out None), with n_q_heads == 0 only k and v — the two halves of cross-attention, whose queries and keys come
    from different sequences (layers/attention.py:431-451).
    """
    _need_cuda(x, w, bias, rope_cos, rope_sin, q_out, k_out, v_out)
    M, K = x.shape
    N = w.shape[0]
    a_mn, lda = _major(x, "qkv x")
    b_mn, ldb = _major(w, "qkv w")
    if (n_q_heads > 0 and q_out is None) or (n_kv_heads > 0 and (k_out is None or v_out is None)):
        raise _lib.VyomError("qkv_rope_gemm: missing output tensor")
    for t in (q_out, k_out, v_out):
        if t is not None and (t.dim() != 4 or t.stride(3) != 1):
            raise _lib.VyomError("qkv_rope_gemm: outputs must be [B,h,S,d] with contiguous d")
    if k_out is not None and k_out.dtype != v_out.dtype:
        raise _lib.VyomError("qkv_rope_gemm: k_out and v_out must share a dtype")
    ref_out = q_out if q_out is not None else k_out
    if q_out is None:
        q_out = k_out  # placeholders for the struct's strides / dtype (never written when the head count is 0)
    if k_out is None:
        k_out = v_out = q_out
    if rope_cos is not None and (rope_cos.dtype != torch.float32 or not rope_cos.is_contiguous()):
        raise _lib.VyomError("qkv_rope_gemm: rope tables must be contiguous float32")
    kw = dict(
        M=M, N=N, K=K, in_dtype=_dt(x), A=x.data_ptr(), lda=lda, a_mn_major=a_mn,
        B=w.data_ptr(), ldb=ldb, b_mn_major=b_mn,
        epi=C["VY_EPI_QKV_ROPE"], bias=_ptr(bias), bias_dtype=_dt(bias) if bias is not None else 0,
        out_dtype=_dt(q_out), tokens_per_seq=tokens_per_seq, start_pos=start_pos,
        kv_dst_pos0=start_pos if kv_dst_pos0 is None else kv_dst_pos0, head_dim=head_dim,
        n_q_heads=n_q_heads, n_kv_heads=n_kv_heads, rope_cos=_ptr(rope_cos), rope_sin=_ptr(rope_sin),
        q_out=q_out.data_ptr(), q_sb=q_out.stride(0), q_sh=q_out.stride(1), q_sl=q_out.stride(2),
        k_out=k_out.data_ptr(), k_sb=k_out.stride(0), k_sh=k_out.stride(1), k_sl=k_out.stride(2),
        v_out=v_out.data_ptr(), v_sb=v_out.stride(0), v_sh=v_out.stride(1), v_sl=v_out.stride(2),
        kv_out_dtype=_dt(k_out), kv_cap=k_out.shape[2] if n_kv_heads > 0 else 0, rope_rows=rope_cos.shape[0] if rope_cos is not None else 0,
        stream=_stream(),
    )
    if gemm_tune.ENABLED:
        key = ("qkv", M, N, K, kw["in_dtype"], a_mn, b_mn, bias is not None, rope_cos is not None, n_q_heads, n_kv_heads,
               tokens_per_seq, q_out.dtype, k_out.dtype)
        written = ([("q_out", q_out)] if n_q_heads > 0 else []) + ([("k_out", k_out), ("v_out", v_out)] if n_kv_heads > 0 else [])
        kw.update(gemm_tune.hints(key, kw, written, x.device))
    _lib.call("vy_gemm", "VyGemm", **kw)


def add_layernorm(
    x: torch.Tensor,
    residual: Optional[torch.Tensor],
    gamma: torch.Tensor,
    beta: torch.Tensor,
    eps: float,
    *,
    save_stats: bool = False,
    save_sum: bool = False,
    kind: str = "layernorm",
    dropout: Optional["DropoutState"] = None,
):
    """y = LayerNorm(dropout(x) + residual). Returns (y, sum_or_None, mean_or_None, rstd_or_None). `dropout` (a
    DropoutState with p > 0) drops x before the residual add — attention.py:70 / ffn.py:38 in .train().
    kind="rmsnorm": y = gamma * xhat with xhat = s * rsqrt(mean(s^2) + eps) (custom_transformer.py:227-241; `beta`, the
    optional shift of simple_vllm.ipynb's RMSNorm, may be None); kind="gemma_rmsnorm": y = (1 + gamma) * xhat."""
    _need_cuda(x, residual, gamma, beta)
    H = x.shape[-1]
    x2 = x.reshape(-1, H)
    if not x2.is_contiguous():
        raise _lib.VyomError("add_layernorm: x must be contiguous")
    rows = x2.shape[0]
    r2 = None
    if residual is not None:
        r2 = residual.reshape(-1, H)
        if not r2.is_contiguous() or r2.dtype != x2.dtype:
            raise _lib.VyomError("add_layernorm: residual must be contiguous and match x's dtype")
    y = torch.empty_like(x2)
    s = torch.empty_like(x2) if (save_sum and residual is not None) else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    if beta is None and kind == "layernorm":
        raise _lib.VyomError("add_layernorm: LayerNorm needs beta")
    _lib.call(
        "vy_add_layernorm_fwd", "VyNorm",
        rows=rows, H=H, x=x2.data_ptr(), residual=_ptr(r2), io_dtype=_dt(x2), gamma=gamma.data_ptr(),
        beta=_ptr(beta), param_dtype=_dt(gamma), eps=float(eps), y=y.data_ptr(), sum_out=_ptr(s),
        mean=_ptr(mean), rstd=_ptr(rstd), kind=NORM_KIND[kind], stream=_stream(), **(dropout.kwargs() if dropout else {}),
    )
    if save_sum and residual is None:
        s = x2
    return y.view(x.shape), s, mean, rstd


class DropoutState:
    """Identifies one dropout mask: (p, seed, offset, step counter). The mask itself is never stored — the forward and
    the backward kernels regenerate it from these four (include/vyom_b200.h, VyNorm.dropout_*)."""

    _next_offset = 0
    seed = 0x5DEECE66D
    step_ptr: Optional[torch.Tensor] = None  # device int32 step counter shared by every mask (trainer.Trainer sets it, so
    #                                          a replayed CUDA graph draws fresh masks each step); None = step 0

    def __init__(self, p: float, step_ptr: Optional[torch.Tensor] = None):
        self.p = float(p)
        self.seed = DropoutState.seed
        DropoutState._next_offset = (DropoutState._next_offset + 1) & 0xFFFFFFFF
        self.offset = DropoutState._next_offset
        self.step_ptr = step_ptr if step_ptr is not None else DropoutState.step_ptr

    @staticmethod
    def manual_seed(seed: int) -> None:
        DropoutState.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        DropoutState._next_offset = 0

    def kwargs(self) -> dict:
        return dict(dropout_p=self.p, dropout_seed=self.seed, dropout_offset=self.offset, dropout_step_ptr=_ptr(self.step_ptr))


def add_layernorm_bwd(dy, s, gamma, mean, rstd, dgamma_out: Optional[torch.Tensor] = None,
                      dbeta_out: Optional[torch.Tensor] = None, dbias_out: Optional[torch.Tensor] = None,
                      want_dbias: bool = False, accumulate: bool = True, kind: str = "layernorm",
                      dropout: Optional[DropoutState] = None):
    """Returns (dx, dgamma, dbeta[, dbias]) for y = LayerNorm(s) (or the RMSNorm kinds of add_layernorm; `mean` may then
    be None and dbeta is the gradient of the optional shift). Without dgamma_out/dbeta_out the parameter
    gradients come back as fresh fp32 tensors; with them (same dtype, e.g. the parameters' .grad views) the kernel
    ACCUMULATES into those buffers (or, with accumulate=False, overwrites them) and returns them. dbias (want_dbias / dbias_out) = column sums of dx: the bias
    gradient of the Linear that produced the normalised sum."""
    _need_cuda(dy, s, gamma, mean, rstd, dgamma_out, dbeta_out, dbias_out)
    H = dy.shape[-1]
    dy2 = dy.reshape(-1, H)
    s2 = s.reshape(-1, H)
    if not dy2.is_contiguous() or not s2.is_contiguous():
        raise _lib.VyomError("add_layernorm_bwd: dy and s must be contiguous")
    rows = dy2.shape[0]
    dx = torch.empty_like(dy2)
    dx_drop = torch.empty_like(dy2) if dropout else None
    acc = dgamma_out is not None
    if acc:
        if dbeta_out is None or dgamma_out.dtype != dbeta_out.dtype or not (dgamma_out.is_contiguous() and dbeta_out.is_contiguous()):
            raise _lib.VyomError("add_layernorm_bwd: dgamma_out / dbeta_out must both be given, contiguous, same dtype")
        dgamma, dbeta = dgamma_out, dbeta_out
        dbias = dbias_out
        if dbias is not None and (dbias.dtype != dgamma.dtype or not dbias.is_contiguous()):
            raise _lib.VyomError("add_layernorm_bwd: dbias_out must be contiguous and share dgamma_out's dtype")
    else:
        if dbias_out is not None:
            raise _lib.VyomError("add_layernorm_bwd: dbias_out needs dgamma_out / dbeta_out as well")
        dgamma = torch.empty(H, device=dy.device, dtype=torch.float32)
        dbeta = torch.empty(H, device=dy.device, dtype=torch.float32)
        dbias = torch.empty(H, device=dy.device, dtype=torch.float32) if want_dbias else None
    nparts = _lib.lib().vy_norm_bwd_partial_rows()
    partials = torch.empty(2 * nparts * H, device=dy.device, dtype=torch.float32)
    _lib.call(
        "vy_add_layernorm_bwd", "VyNorm",
        rows=rows, H=H, io_dtype=_dt(dy2), gamma=gamma.data_ptr(), param_dtype=_dt(gamma),
        mean=_ptr(mean), rstd=rstd.data_ptr(), dy=dy2.data_ptr(), s=s2.data_ptr(), dx=dx.data_ptr(),
        dgamma=dgamma.data_ptr(), dbeta=dbeta.data_ptr(), dbias=_ptr(dbias), dparam_dtype=_dt(dgamma),
        dparam_accumulate=int(acc and accumulate), partials=partials.data_ptr(), kind=NORM_KIND[kind], stream=_stream(),
        dx_drop=_ptr(dx_drop), **(dropout.kwargs() if dropout else {}),
    )
    # with dropout: dx is the gradient of the pre-norm sum (= of the residual), dx_drop that of the dropped input x
    # (and dbias its column sums); returned as a pair in dx's place
    dx_ret = dx.view(dy.shape) if not dropout else (dx.view(dy.shape), dx_drop.view(dy.shape))
    if want_dbias or dbias_out is not None:
        return dx_ret, dgamma, dbeta, dbias
    return dx_ret, dgamma, dbeta


def attn_fwd(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    *,
    causal: bool = False,
    q_pos0: int = 0,
    key_padding_mask: Optional[torch.Tensor] = None,
    out: Optional[torch.Tensor] = None,
    out_dtype: Optional[torch.dtype] = None,
    need_lse: bool = False,
    prefix_len: Optional[torch.Tensor] = None,
    pos_dev: Optional[torch.Tensor] = None,
):
    """Flash attention forward. q [B,Hq,Sq,D], k/v [B,Hkv,Skv,D] (bf16, any batch/head/token strides; D = 64 on the
    tensor-memory kernel, any multiple of 8 up to 256 on the mma.sync one). `prefix_len` (int32 [B], with causal): keys below
    it are visible to every query (prefix-LM). Returns (out [B,Sq,Hq*D], lse [B,Hq,Sq] or None)."""
    _need_cuda(q, k, v, key_padding_mask, out, prefix_len, pos_dev)
    if prefix_len is not None and (prefix_len.dtype != torch.int32 or not prefix_len.is_contiguous() or prefix_len.numel() != q.shape[0]):
        raise _lib.VyomError("attn_fwd: prefix_len must be a contiguous int32 tensor with one entry per batch row")
    B, Hq, Sq, D = q.shape
    _, Hkv, Skv, _ = k.shape
    for t in (q, k, v):
        if t.stride(3) != 1:
            raise _lib.VyomError("attn_fwd: head_dim must be contiguous")
    if out is None:
        out = torch.empty((B, Sq, Hq * D), device=q.device, dtype=out_dtype or q.dtype)
    lse = torch.empty((B, Hq, Sq), device=q.device, dtype=torch.float32) if need_lse else None
    kpm = None
    if key_padding_mask is not None:
        kpm = key_padding_mask
        if kpm.dtype != torch.uint8 or kpm.stride(1) != 1:
            raise _lib.VyomError("attn_fwd: key_padding_mask must be uint8 [B, Skv] with unit inner stride")
    _lib.call(
        "vy_attn_fwd", "VyAttn",
        B=B, n_q_heads=Hq, n_kv_heads=Hkv, head_dim=D, Sq=Sq, Skv=Skv, qkv_dtype=_dt(q),
        q=q.data_ptr(), q_sb=q.stride(0), q_sh=q.stride(1), q_sl=q.stride(2),
        k=k.data_ptr(), k_sb=k.stride(0), k_sh=k.stride(1), k_sl=k.stride(2),
        v=v.data_ptr(), v_sb=v.stride(0), v_sh=v.stride(1), v_sl=v.stride(2),
        causal=int(causal), q_pos0=q_pos0, key_padding_mask=_ptr(kpm),
        kpm_stride=kpm.stride(0) if kpm is not None else 0,
        out=out.data_ptr(), o_sb=out.stride(0), o_sl=out.stride(1), out_dtype=_dt(out), lse=_ptr(lse),
        prefix_len=_ptr(prefix_len), pos_ptr=_ptr(pos_dev), stream=_stream(),
    )
    return out, lse


_TICKETS = {}


def _tickets(device: torch.device, n: int) -> torch.Tensor:
    t = _TICKETS.get(device)
    if t is None or t.numel() < n:
        t = torch.zeros(max(n, 4096), device=device, dtype=torch.int32)
        _TICKETS[device] = t
    return t


def attn_decode(
    qkv: torch.Tensor,
    k_cache: torch.Tensor,
    v_cache: torch.Tensor,
    start_pos: int,
    n_q_heads: int,
    n_kv_heads: int,
    rope_cos: Optional[torch.Tensor],
    rope_sin: Optional[torch.Tensor],
    *,
    out: Optional[torch.Tensor] = None,
    out_dtype: Optional[torch.dtype] = None,
    splits: int = 0,
    start_pos_dev: Optional[torch.Tensor] = None,
    seqlens: Optional[torch.Tensor] = None,
    block_table: Optional[torch.Tensor] = None,
    rope_pos_off: int = 0,
) -> torch.Tensor:
    """Single-token attention with fused RoPE and kv-cache append. The new token's angles are row
    (position + rope_pos_off) of the tables (0 when row 0 is position 0). `seqlens` (int32 [B], device) gives every row its
    own context length (continuous batching; negative = idle slot); `block_table` (int32 [B, max_blocks], device) makes
    the caches paged pools [num_blocks, block_size, Hkv, 64] (Examples/simple_vllm.ipynb) — `start_pos` is then the
    largest context length of the batch. With `start_pos_dev` (int32 [1] on the device) the
    kernel takes the position from device memory and `start_pos` is only the upper bound that sizes the kv-split —
    the form a captured decode step uses. qkv [B, (Hq+2Hkv)*64] packed
    projections; caches [>=B, Hkv, cache_len, 64]; returns out [B, Hq*64]."""
    _need_cuda(qkv, k_cache, v_cache, rope_cos, rope_sin, out)
    B = qkv.shape[0]
    D = 64
    if k_cache.stride() != v_cache.stride() or k_cache.dtype != v_cache.dtype or k_cache.stride(3) != 1:
        raise _lib.VyomError("attn_decode: k/v caches must share dtype and strides, head_dim contiguous")
    _need_cuda(seqlens, block_table)
    for t, name in ((seqlens, "seqlens"), (block_table, "block_table")):
        if t is not None and (t.dtype != torch.int32 or not t.is_contiguous()):
            raise _lib.VyomError(f"attn_decode: {name} must be a contiguous int32 tensor")
    paged = block_table is not None
    if paged:  # [num_blocks, block_size, Hkv, 64]
        block_size = k_cache.shape[1]
        cache_len = block_table.shape[1] * block_size
        c_sb, c_sl, c_sh = k_cache.stride(0), k_cache.stride(1), k_cache.stride(2)
    else:      # [>=B, Hkv, cache_len, 64]
        block_size = 0
        cache_len = k_cache.shape[2]
        c_sb, c_sh, c_sl = k_cache.stride(0), k_cache.stride(1), k_cache.stride(2)
    if out is None:
        out = torch.empty((B, n_q_heads * D), device=qkv.device, dtype=out_dtype or qkv.dtype)
    kw = dict(
        B=B, n_q_heads=n_q_heads, n_kv_heads=n_kv_heads, head_dim=D, start_pos=start_pos,
        cache_len=cache_len, qkv=qkv.data_ptr(), ld_qkv=qkv.stride(0), qkv_dtype=_dt(qkv),
        rope_cos=_ptr(rope_cos), rope_sin=_ptr(rope_sin), rope_pos_off=rope_pos_off,
        rope_rows=rope_cos.shape[0] if rope_cos is not None else 0, k_cache=k_cache.data_ptr(), v_cache=v_cache.data_ptr(),
        cache_sb=c_sb, cache_sh=c_sh, cache_sl=c_sl, cache_dtype=_dt(k_cache),
        seqlens=_ptr(seqlens), block_table=_ptr(block_table),
        max_blocks_per_seq=block_table.shape[1] if paged else 0, block_size=block_size,
        out=out.data_ptr(), ld_out=out.stride(0), out_dtype=_dt(out), start_pos_ptr=_ptr(start_pos_dev), stream=_stream(),
    )
    if splits <= 0:  # the library's own choice for this layout (copy-engine kernel or LDG kernel)
        st = _lib.STRUCTS["VyDecode"]()
        for k, v in kw.items():
            if v is not None:
                setattr(st, k, v)
        splits = _lib.lib().vy_attn_decode_plan(ctypes.byref(st))
    ws = tk = None
    if splits > 1:
        ws = torch.empty(B * n_kv_heads * splits * (n_q_heads // n_kv_heads) * 66, device=qkv.device, dtype=torch.float32)
        tk = _tickets(qkv.device, B * n_kv_heads)
    _lib.call("vy_attn_decode", "VyDecode", splits=splits, workspace=_ptr(ws), tickets=_ptr(tk), **kw)
    return out


def cast4d(src: torch.Tensor, dtype: torch.dtype, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Strided 4-D copy with dtype conversion (inner dim contiguous)."""
    _need_cuda(src, out)
    if src.dim() != 4 or src.stride(3) != 1:
        raise _lib.VyomError("cast4d: expects a 4-D tensor with a contiguous inner dim")
    if out is None:
        out = torch.empty(src.shape, device=src.device, dtype=dtype)
    _lib.call(
        "vy_cast4d", "VyCast4d",
        n0=src.shape[0], n1=src.shape[1], n2=src.shape[2], n3=src.shape[3], src=src.data_ptr(), src_dtype=_dt(src),
        s0=src.stride(0), s1=src.stride(1), s2=src.stride(2), dst=out.data_ptr(), dst_dtype=_dt(out),
        d0=out.stride(0), d1=out.stride(1), d2=out.stride(2), stream=_stream(),
    )
    return out


def embed(ids: Optional[torch.Tensor], src: torch.Tensor, *, out: torch.Tensor, tokens_per_seq: int,
          out_group_stride: int = 0, out_row_off: int = 0, pos: Optional[torch.Tensor] = None, pos_row_off: int = 0,
          out_scale: float = 1.0, rows: Optional[int] = None, broadcast_src: bool = False,
          src_row_stride: Optional[int] = None) -> torch.Tensor:
    """Row gather + positional add + scale + row remap (see vy_embed_fwd). `out` is 2-D [rows_out, H]."""
    _need_cuda(ids, src, out, pos)
    H = src.shape[-1]
    n = rows if rows is not None else (ids.numel() if ids is not None else src.shape[0])
    if ids is not None and (ids.dtype != torch.int64 or not ids.is_contiguous()):
        raise _lib.VyomError("embed: ids must be contiguous int64")
    if pos is not None and (pos.dtype != src.dtype or not pos.is_contiguous()):
        raise _lib.VyomError("embed: pos must be contiguous and share the table dtype")
    _lib.call(
        "vy_embed_fwd", "VyEmbed",
        rows=n, H=H, ids=_ptr(ids), src=src.data_ptr(),
        ld_src=0 if broadcast_src else (src_row_stride if src_row_stride is not None else src.stride(-2)), dtype=_dt(src),
        vocab=src.shape[0] if src.dim() == 2 else 1, tokens_per_seq=tokens_per_seq, out_group_stride=out_group_stride,
        out_row_off=out_row_off, pos=_ptr(pos), pos_row_off=pos_row_off, out_scale=float(out_scale), out=out.data_ptr(),
        ld_out=out.stride(0), stream=_stream(),
    )
    return out


def embed_bwd(ids: Optional[torch.Tensor], dout: torch.Tensor, *, rows: int, H: int, tokens_per_seq: int,
              out_group_stride: int = 0, out_row_off: int = 0, dtable: Optional[torch.Tensor] = None,
              dpos: Optional[torch.Tensor] = None, pos_row_off: int = 0, out_scale: float = 1.0,
              padding_idx: Optional[int] = None, pos_padding_idx: Optional[int] = None) -> None:
    """Scatter-add of `dout` rows into the table / position gradients. `padding_idx` / `pos_padding_idx`: the table /
    position row that nn.Embedding(padding_idx=...) never updates (skipped here too)."""
    _need_cuda(ids, dout, dtable, dpos)
    _lib.call(
        "vy_embed_bwd", "VyEmbed",
        rows=rows, H=H, ids=_ptr(ids), dtype=_dt(dout), vocab=dtable.shape[0] if dtable is not None else rows,
        tokens_per_seq=tokens_per_seq, out_group_stride=out_group_stride, out_row_off=out_row_off,
        pos_row_off=pos_row_off, out_scale=float(out_scale), ld_out=dout.stride(0), dout=dout.data_ptr(),
        dtable=_ptr(dtable), ld_src=dtable.stride(0) if dtable is not None else 0, dpos=_ptr(dpos),
        padding_idx_plus1=0 if padding_idx is None else padding_idx + 1,
        pos_padding_idx_plus1=0 if pos_padding_idx is None else pos_padding_idx + 1, stream=_stream(),
    )


def patchify(pixels: torch.Tensor, patch: Tuple[int, int], dtype: torch.dtype, pad_to: int = 1) -> torch.Tensor:
    """NCHW pixels -> [B * nP, C*ph*pw] patch rows in `dtype`. pad_to > 1: rows are zero-padded to a multiple of `pad_to`
    columns (a K the GEMM can take: 16-byte rows)."""
    _need_cuda(pixels)
    if not pixels.is_contiguous():
        raise _lib.VyomError("patchify: pixels must be contiguous NCHW")
    B, Cc, Hh, Ww = pixels.shape
    ph, pw = patch
    K = Cc * ph * pw
    Kp = (K + pad_to - 1) // pad_to * pad_to
    rows = B * (Hh // ph) * (Ww // pw)
    out = torch.empty((rows, K), device=pixels.device, dtype=dtype) if Kp == K else torch.zeros((rows, Kp), device=pixels.device, dtype=dtype)
    _lib.call("vy_patchify", "VyPatchify", B=B, C=Cc, H=Hh, W=Ww, patch_h=ph, patch_w=pw, pixels=pixels.data_ptr(),
              in_dtype=_dt(pixels), out=out.data_ptr(), out_dtype=_dt(out), ld_out=Kp if Kp != K else 0, stream=_stream())
    return out


def argmax_rows(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Greedy token ids: first index of each row's maximum. x is 2-D with unit inner stride."""
    _need_cuda(x, out)
    if out is None:
        out = torch.empty(x.shape[0], device=x.device, dtype=torch.int64)
    elif out.dtype != torch.int64 or not out.is_contiguous() or out.numel() != x.shape[0]:
        raise _lib.VyomError("argmax_rows: out must be a contiguous int64 tensor with one entry per row")
    _lib.check(_lib.lib().vy_argmax_rows(x.shape[0], x.shape[1], x.data_ptr(), x.stride(0), _dt(x), out.data_ptr(), _stream()),
               "vy_argmax_rows")
    return out


def argmax_advance(x: torch.Tensor, out: torch.Tensor, pos: torch.Tensor, tokens: Optional[torch.Tensor] = None) -> torch.Tensor:
    """argmax_rows into `out`, then pos += 1 (device int32 scalar) and tokens[:, pos] = out (tokens: [rows, n] int64, unit inner
    stride) — one launch sequence for the tail of a captured greedy step."""
    _need_cuda(x, out, pos, tokens)
    if out.dtype != torch.int64 or not out.is_contiguous() or out.numel() != x.shape[0] or pos.dtype != torch.int32 or pos.numel() != 1:
        raise _lib.VyomError("argmax_advance: out must be contiguous int64 [rows], pos one int32 element")
    if tokens is not None and (tokens.dtype != torch.int64 or tokens.dim() != 2 or tokens.stride(1) != 1 or tokens.shape[0] != x.shape[0]):
        raise _lib.VyomError("argmax_advance: tokens must be int64 [rows, n] with a unit inner stride")
    _lib.check(_lib.lib().vy_argmax_advance(x.shape[0], x.shape[1], x.data_ptr(), x.stride(0), _dt(x), out.data_ptr(), pos.data_ptr(),
                                           _ptr(tokens), tokens.stride(0) if tokens is not None else 0,
                                           tokens.shape[1] if tokens is not None else 0, _stream()), "vy_argmax_advance")
    return out


def colsum(x: torch.Tensor, out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
           accumulate: bool = False, scale: float = 1.0) -> torch.Tensor:
    """out[c] (+)= sum_r x[r, c] (bias gradients)."""
    _need_cuda(x, out)
    R, Cn = x.shape
    if out is None:
        out = torch.empty(Cn, device=x.device, dtype=out_dtype or x.dtype)
    L = _lib.lib()
    ws = torch.empty(L.vy_colsum_workspace_floats(Cn), device=x.device, dtype=torch.float32)
    _lib.check(L.vy_colsum(R, Cn, x.data_ptr(), x.stride(0), _dt(x), out.data_ptr(), _dt(out), int(accumulate),
                           float(scale), ws.data_ptr(), _stream()), "vy_colsum")
    return out


def softmax_xent(logits: torch.Tensor, labels: torch.Tensor, *, ignore_index: int = -100, grad_scale: float = 1.0,
                 grad_scale_ptr: Optional[torch.Tensor] = None, write_grad: bool = True, colsum_part: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-row cross-entropy (fp32 [rows]); overwrites `logits` with d loss/d logits when write_grad. `colsum_part`
    (xent_colsum_part(logits)): the written gradient's column sums are left there as partial sums for colsum_finish."""
    _need_cuda(logits, labels, grad_scale_ptr, colsum_part)
    rows, V = logits.shape
    loss = torch.empty(rows, device=logits.device, dtype=torch.float32)
    _lib.call("vy_softmax_xent", "VyXent", rows=rows, V=V, logits=logits.data_ptr(), ld=logits.stride(0), dtype=_dt(logits),
              labels=labels.data_ptr(), ignore_index=ignore_index, grad_scale_ptr=_ptr(grad_scale_ptr),
              grad_scale=float(grad_scale), loss_rows=loss.data_ptr(), write_grad=int(write_grad), colsum_part=_ptr(colsum_part),
              stream=_stream())
    return loss


def xent_colsum_part(logits: torch.Tensor) -> Optional[torch.Tensor]:
    """The fp32 [chunks, V] buffer softmax_xent(colsum_part=...) fills for these logits, or None when the fused column
    sums do not cover the shape / dtype (the caller then sums the gradient with `colsum`)."""
    rows, V = logits.shape
    Vp = (V + 7) // 8 * 8  # a last partial vector of a row is read and written whole: the row stride must cover it
    if logits.dtype != torch.bfloat16 or logits.stride(1) != 1 or logits.stride(0) % 8 or logits.stride(0) < Vp or logits.data_ptr() % 16:
        return None
    chunks = _lib.lib().vy_xent_colsum_chunks(rows, V, _dt(logits))
    return torch.empty((chunks, Vp), device=logits.device, dtype=torch.float32)[:, :V] if chunks > 0 else None


def colsum_finish(part: torch.Tensor, out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
                  accumulate: bool = False, scale: float = 1.0, scale_ptr: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[c] (+)= scale * (*scale_ptr) * sum_k part[k, c], k in index order."""
    _need_cuda(part, out, scale_ptr)
    chunks, Cn = part.shape
    if out is None:
        out = torch.empty(Cn, device=part.device, dtype=out_dtype or torch.float32)
    _lib.check(_lib.lib().vy_colsum_finish(Cn, chunks, part.data_ptr(), part.stride(0), out.data_ptr(), _dt(out), int(accumulate), float(scale),
                                           _ptr(scale_ptr), _stream()), "vy_colsum_finish")
    return out


def attn_bwd(q, k, v, o, dout, lse, *, causal: bool, q_pos0: int, key_padding_mask, rope_cos, rope_sin,
             dq: torch.Tensor, dk: torch.Tensor, dv: torch.Tensor, rope_pos0: int = 0) -> None:
    """Flash-attention backward. q [B,Hq,Sq,64], k/v [B,Hkv,Skv,64] bf16; o / dout [B,Sq,Hq*64] (dout bf16);
    dq [B*Sq, >=Hq*64], dk/dv [B*Skv, >=Hkv*64] 2-D views (typically column slices of the packed dqkv)."""
    _need_cuda(q, k, v, o, dout, lse, key_padding_mask, rope_cos, rope_sin, dq, dk, dv)
    B, Hq, Sq, D = q.shape
    Hkv, Skv = k.shape[1], k.shape[2]
    if dout.dtype != torch.bfloat16 or q.dtype != torch.bfloat16:
        raise _lib.VyomError("attn_bwd: q/k/v/dout must be bf16")
    o3 = o.view(B, Sq, Hq * D)
    do3 = dout.view(B, Sq, Hq * D)
    dsum = torch.empty((B, Hq, Sq), device=q.device, dtype=torch.float32)
    _lib.call(
        "vy_attn_bwd", "VyAttnBwd",
        B=B, n_q_heads=Hq, n_kv_heads=Hkv, head_dim=D, Sq=Sq, Skv=Skv,
        q=q.data_ptr(), q_sb=q.stride(0), q_sh=q.stride(1), q_sl=q.stride(2),
        k=k.data_ptr(), k_sb=k.stride(0), k_sh=k.stride(1), k_sl=k.stride(2),
        v=v.data_ptr(), v_sb=v.stride(0), v_sh=v.stride(1), v_sl=v.stride(2),
        o=o3.data_ptr(), o_sb=o3.stride(0), o_sl=o3.stride(1), o_dtype=_dt(o3),
        dout=do3.data_ptr(), do_sb=do3.stride(0), do_sl=do3.stride(1), lse=lse.data_ptr(), dsum=dsum.data_ptr(),
        causal=int(causal), q_pos0=q_pos0, key_padding_mask=_ptr(key_padding_mask),
        kpm_stride=key_padding_mask.stride(0) if key_padding_mask is not None else 0,
        rope_cos=_ptr(rope_cos), rope_sin=_ptr(rope_sin), rope_pos0=rope_pos0,
        dq=dq.data_ptr(), ld_dq=dq.stride(0), dk=dk.data_ptr(), ld_dk=dk.stride(0), dv=dv.data_ptr(), ld_dv=dv.stride(0),
        out_dtype=_dt(dq), stream=_stream(),
    )


def rope_apply(x: torch.Tensor, freqs: torch.Tensor, inverse: bool = False) -> torch.Tensor:
    """Half-split RoPE of x [B,h,S,64] with the reference's angle slice freqs (1,S,32); cos/sin are
    rounded to x.dtype first like apply_rotary_pos_emb does."""
    _need_cuda(x)
    f = freqs[0].to(torch.float32).cpu()
    cos = f.cos().to(x.dtype).to(torch.float32).contiguous().to(x.device)
    sin = f.sin().to(x.dtype).to(torch.float32).contiguous().to(x.device)
    if x.stride(3) != 1:
        x = x.contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=x.dtype)
    _lib.call("vy_rope_apply", "VyRope", B=x.shape[0], H=x.shape[1], S=x.shape[2], head_dim=x.shape[3], x=x.data_ptr(),
              x_sb=x.stride(0), x_sh=x.stride(1), x_sl=x.stride(2), dtype=_dt(x), cos=cos.data_ptr(), sin=sin.data_ptr(),
              pos0=0, inverse=int(inverse), out=out.data_ptr(), o_sb=out.stride(0), o_sh=out.stride(1), o_sl=out.stride(2),
              stream=_stream())
    return out


def rope_into(x: torch.Tensor, out: torch.Tensor, cos: Optional[torch.Tensor], sin: Optional[torch.Tensor], pos0: int, inverse: bool = False,
              pos_dev: Optional[torch.Tensor] = None, out_follows_pos: bool = False, copy_only: bool = False) -> torch.Tensor:
    """Half-split RoPE of x [B, h, S, D] (any batch / head / token strides, D contiguous and even) written to `out` (same shape,
    its own strides — e.g. a kv-cache slot range; out may be x itself). cos / sin: fp32 [rows, D / 2] tables, row pos0 + l is
    used for token l. `pos_dev` (device int32 scalar) is added to pos0 on the device; with out_follows_pos the rows land at
    out token index l + pos_dev (out = the base of a kv-cache); copy_only skips the rotation (values)."""
    _need_cuda(x, out, cos, sin, pos_dev)
    B, Hh, S, D = x.shape
    if x.stride(3) != 1 or out.stride(3) != 1 or tuple(out.shape) != tuple(x.shape) or out.dtype != x.dtype:
        raise _lib.VyomError("rope_into: x and out must share shape and dtype with a contiguous head_dim")
    if not copy_only and (cos.dtype != torch.float32 or cos.shape[1] != D // 2 or not cos.is_contiguous() or not sin.is_contiguous()
                          or cos.shape[0] < pos0 + S):
        raise _lib.VyomError("rope_into: cos / sin must be contiguous fp32 [>= pos0 + S, D / 2] tables")
    _lib.call("vy_rope_apply", "VyRope", B=B, H=Hh, S=S, head_dim=D, x=x.data_ptr(), x_sb=x.stride(0), x_sh=x.stride(1),
              x_sl=x.stride(2), dtype=_dt(x), cos=_ptr(cos), sin=_ptr(sin), pos0=pos0, inverse=int(inverse),
              out=out.data_ptr(), o_sb=out.stride(0), o_sh=out.stride(1), o_sl=out.stride(2), pos_ptr=_ptr(pos_dev),
              out_follows_pos=int(out_follows_pos), copy_only=int(copy_only), stream=_stream())
    return out


def rope_append(qkv4: torch.Tensor, n_q_heads: int, n_kv_heads: int, k_cache: torch.Tensor, v_cache: torch.Tensor, cos: torch.Tensor,
                sin: torch.Tensor, pos0: int, slot0: int, pos_dev: Optional[torch.Tensor] = None) -> None:
    """RoPE of the query heads (in place) and key heads (into k_cache) and the value copy (into v_cache) of a packed projection
    viewed as [B, n_q + 2 n_kv, S, D]; caches [>= B, n_kv, slots, D]; token l -> table row pos0 + l, cache slot slot0 + l (both
    + pos_dev on the device)."""
    _need_cuda(qkv4, k_cache, v_cache, cos, sin, pos_dev)
    B, Hh, S, D = qkv4.shape
    if Hh != n_q_heads + 2 * n_kv_heads or qkv4.stride(3) != 1 or k_cache.stride(3) != 1 or k_cache.stride() != v_cache.stride():
        raise _lib.VyomError("rope_append: qkv must be [B, n_q + 2 n_kv, S, D] with contiguous D; k / v caches must share strides")
    if k_cache.dtype != qkv4.dtype or v_cache.dtype != qkv4.dtype or k_cache.shape[1] != n_kv_heads or k_cache.shape[3] != D:
        raise _lib.VyomError("rope_append: caches must be [B', n_kv, slots, D] in the projection's dtype")
    if cos.dtype != torch.float32 or cos.shape[1] != D // 2 or not cos.is_contiguous() or not sin.is_contiguous():
        raise _lib.VyomError("rope_append: cos / sin must be contiguous fp32 [rows, D / 2] tables")
    _lib.call("vy_rope_append", "VyRopeAppend", B=B, S=S, n_q_heads=n_q_heads, n_kv_heads=n_kv_heads, head_dim=D, qkv=qkv4.data_ptr(),
              sb=qkv4.stride(0), sh=qkv4.stride(1), sl=qkv4.stride(2), dtype=_dt(qkv4), cos=cos.data_ptr(), sin=sin.data_ptr(),
              pos0=pos0, slot0=slot0, pos_ptr=_ptr(pos_dev), k_cache=k_cache.data_ptr(), v_cache=v_cache.data_ptr(),
              c_sb=k_cache.stride(0), c_sh=k_cache.stride(1), c_sl=k_cache.stride(2), cache_slots=k_cache.shape[2],
              rope_rows=cos.shape[0], stream=_stream())


def act_bwd(dy: torch.Tensor, z: torch.Tensor, act: str = "gelu") -> torch.Tensor:
    """dy * act'(z), elementwise (contiguous, same dtype)."""
    _need_cuda(dy, z)
    if not (dy.is_contiguous() and z.is_contiguous()) or dy.dtype != z.dtype:
        raise _lib.VyomError("act_bwd: dy and z must be contiguous and share a dtype")
    out = torch.empty_like(dy)
    _lib.check(_lib.lib().vy_act_bwd(dy.numel(), dy.data_ptr(), z.data_ptr(), _dt(dy), ACT[act], out.data_ptr(), _stream()),
               "vy_act_bwd")
    return out


def scale_by_ptr(x: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """x *= scale (fp32 device scalar) in place; free when the scalar is 1."""
    _need_cuda(x, scale)
    if scale.dtype != torch.float32 or scale.numel() != 1:
        raise _lib.VyomError("scale_by_ptr: scale must be one float32 element on the device")
    if not x.is_contiguous():
        raise _lib.VyomError("scale_by_ptr: x must be contiguous")
    _lib.check(_lib.lib().vy_scale_by_ptr(x.numel(), x.data_ptr(), _dt(x), scale.data_ptr(), _stream()), "vy_scale_by_ptr")
    return x


def slot_merge(a: torch.Tensor, b: torch.Tensor, slot: torch.Tensor) -> torch.Tensor:
    """out[r] = b[slot[r]] where slot[r] >= 0, else a[r] (masked_scatter of image-feature rows into embedding rows)."""
    _need_cuda(a, b, slot)
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1] or a.dtype != b.dtype or not (a.is_contiguous() and b.is_contiguous()):
        raise _lib.VyomError("slot_merge: a [rows, H] and b [n, H] must be contiguous 2-D tensors of one dtype")
    if slot.dtype != torch.int32 or not slot.is_contiguous() or slot.numel() != a.shape[0]:
        raise _lib.VyomError("slot_merge: slot must be a contiguous int32 tensor with one entry per row of a")
    out = torch.empty_like(a)
    _lib.check(_lib.lib().vy_slot_merge_fwd(a.shape[0], a.shape[1], _dt(a), a.data_ptr(), b.data_ptr(), b.shape[0], slot.data_ptr(), out.data_ptr(),
                                           _stream()), "vy_slot_merge_fwd")
    return out


def slot_merge_bwd(dout: torch.Tensor, slot: torch.Tensor, n_b: int, need_a: bool = True, need_b: bool = True):
    """(da, db) for slot_merge: da = dout with the slot rows zeroed, db[slot[r]] = dout[r] (rows no slot points at are zero)."""
    _need_cuda(dout, slot)
    if dout.dim() != 2 or not dout.is_contiguous():
        raise _lib.VyomError("slot_merge_bwd: dout must be a contiguous 2-D tensor")
    da = torch.empty_like(dout) if need_a else None
    db = torch.zeros((n_b, dout.shape[1]), device=dout.device, dtype=dout.dtype) if need_b else None
    if da is None and db is None:
        return None, None
    _lib.check(_lib.lib().vy_slot_merge_bwd(dout.shape[0], dout.shape[1], _dt(dout), dout.data_ptr(), slot.data_ptr(), _ptr(da), _ptr(db),
                                           n_b, _stream()), "vy_slot_merge_bwd")
    return da, db


# ---- bias-gradient column sums on a side stream ------------------------------------------------------------------------
# A bias gradient is a column sum over the same dY a weight-gradient GEMM reads. The GEMM is tensor-bound and leaves HBM
# idle; the column sum is HBM-bound and needs few SM resources — so it is launched on a side stream AFTER the GEMM (whose
# persistent CTAs are placed first) and runs in the registers / shared memory the GEMM leaves free. `side_fork_point()` marks
# the moment dY is complete on the current stream, `colsum_side(...)` runs the sum on the side stream after that point, and
# `side_join()` makes the current stream wait for everything forked so far (before dY can be released).
# Measured on the captioner step (same box, back to back): 10.31 -> 10.13 ms/step (+1.8 % samples/s), bit-identical losses; the
# GEMMs that share their SMs with a column sum take longer individually (6.41 -> 6.73 ms summed), so the per-kernel table of
# bench.py no longer adds up to the step and vy_gemm's own roofline fraction reads lower. Opt-in for that reason:
# VY_COLSUM_SIDE=1.
_SIDE_ON = os.environ.get("VY_COLSUM_SIDE", "0") != "0"
_SIDE_STREAMS = {}
_SIDE_PENDING = {}


def side_fork_point() -> Optional[torch.cuda.Event]:
    if not _SIDE_ON:
        return None
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream())
    return ev


def colsum_side(after: Optional[torch.cuda.Event], x: torch.Tensor, **kw) -> None:
    """vy_colsum into kw['out'] on the side stream once `after` has passed (None: plain call on the current stream)."""
    if after is None:
        colsum(x, **kw)
        return
    dev = x.device
    side = _SIDE_STREAMS.get(dev)
    if side is None:
        side = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
    side.wait_event(after)
    with torch.cuda.stream(side):
        colsum(x, **kw)
        done = torch.cuda.Event()
        done.record(side)
    _SIDE_PENDING[dev] = done


def side_join(device: Optional[torch.device] = None) -> None:
    """The current stream waits for the side-stream sums forked so far (no-op when there are none)."""
    if device is None:
        if not _SIDE_PENDING:
            return
        device = torch.device("cuda", torch.cuda.current_device())
    ev = _SIDE_PENDING.pop(device, None)
    if ev is not None:
        torch.cuda.current_stream().wait_event(ev)


def swiglu_bwd(dh: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """dz for h = silu(z[:, 0::2]) * z[:, 1::2] (the act="swiglu" epilogue of gemm): dh [rows, I], z [rows, 2 I]."""
    _need_cuda(dh, z)
    if not (dh.is_contiguous() and z.is_contiguous()) or dh.dtype != z.dtype or z.shape[-1] != 2 * dh.shape[-1]:
        raise _lib.VyomError("swiglu_bwd: dh [rows, I] and z [rows, 2 I] must be contiguous and share a dtype")
    dz = torch.empty_like(z)
    inter = dh.shape[-1]
    _lib.check(_lib.lib().vy_swiglu_bwd(dh.numel() // inter, inter, dh.data_ptr(), z.data_ptr(), _dt(dh), dz.data_ptr(), _stream()),
               "vy_swiglu_bwd")
    return dz


def sqnorm(g: torch.Tensor, out: torch.Tensor) -> None:
    """out (fp32 scalar tensor, pre-zeroed) += sum(g^2)."""
    _need_cuda(g, out)
    _lib.check(_lib.lib().vy_sqnorm(g.numel(), g.data_ptr(), _dt(g), out.data_ptr(), _stream()), "vy_sqnorm")


def adamw(param, grad, exp_avg, exp_avg_sq, *, lr, beta1, beta2, eps, weight_decay, step, master=None,
          grad_sqnorm=None, max_grad_norm=0.0, grad_div=1.0, step_ptr=None) -> None:
    """Fused AdamW over flat buffers (see vy_adamw)."""
    _need_cuda(param, grad, exp_avg, exp_avg_sq, master, grad_sqnorm)
    _lib.call("vy_adamw", "VyAdamW", n=param.numel(), param=param.data_ptr(), param_dtype=_dt(param), grad=grad.data_ptr(),
              grad_dtype=_dt(grad), exp_avg=exp_avg.data_ptr(), exp_avg_sq=exp_avg_sq.data_ptr(), master=_ptr(master),
              lr=float(lr), beta1=float(beta1), beta2=float(beta2), eps=float(eps), weight_decay=float(weight_decay),
              step=int(step), step_ptr=_ptr(step_ptr), grad_sqnorm=_ptr(grad_sqnorm), max_grad_norm=float(max_grad_norm), grad_div=float(grad_div),
              stream=_stream())
