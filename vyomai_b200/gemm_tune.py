"""Per-shape tiling choice for vy_gemm, measured instead of modelled.

vy_gemm picks its kernel flavour (single-CTA 128 x BN tiles or CTA-pair 256 x BN tiles), tile width and K split from a
small time model (csrc/gemm.cu:choose_tiling). The model is within a few percent on average but off by 20-30 % on
individual shapes, so the host side times the candidates once per distinct call signature — on the caller's own
operands, with every buffer the GEMM WRITES redirected to scratch so that an accumulating epilogue is not applied
twice — and from then on passes the winner through VyGemm.hint_*. Nothing is tuned while a CUDA graph is being
captured (the capture gets the cached winner or, for a signature never seen eagerly, the model's choice).

VY_GEMM_AUTOTUNE=0 turns this off. VY_GEMM_TUNE_CACHE=<file.json> keeps the winners across processes (read at import,
appended to after every tuned signature): a second run — a profiler pass, a restarted job — launches no tuning kernels.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

ENABLED = os.environ.get("VY_GEMM_AUTOTUNE", "1") != "0"
CACHE_FILE = os.environ.get("VY_GEMM_TUNE_CACHE") or None
_CACHE: Dict[tuple, Dict[str, int]] = {}
_FILE_CACHE: Dict[str, Dict[str, int]] = {}
if CACHE_FILE and os.path.exists(CACHE_FILE):
    try:
        with open(CACHE_FILE) as _f:
            _FILE_CACHE = json.load(_f)
    except (OSError, ValueError):
        _FILE_CACHE = {}
_FLUSH: Dict[torch.device, torch.Tensor] = {}
MAX_SCRATCH_BYTES = 4 << 30
LOG: List[tuple] = []  # (key, winner, us, model_us) of every signature tuned in this process


def clear() -> None:
    _CACHE.clear()
    LOG.clear()


def _flush_l2(device: torch.device) -> None:
    buf = _FLUSH.get(device)
    if buf is None:
        buf = _FLUSH[device] = torch.empty(192 << 20, device=device, dtype=torch.uint8)
    buf.zero_()


def _candidates(kw: dict) -> List[Dict[str, int]]:
    M, N, K = kw["M"], kw["N"], kw["K"]
    bf16 = kw["in_dtype"] == _lib.CONSTS["VY_BF16"]
    mn_major = bool(kw.get("a_mn_major")) or bool(kw.get("b_mn_major"))
    qkv = kw["epi"] == _lib.CONSTS["VY_EPI_QKV_ROPE"]
    num_kb = (K + (63 if bf16 else 31)) // (64 if bf16 else 32)
    splits = [1]
    if kw.get("workspace"):
        per = M * N * 4
        for sp in (2, 3, 4, 6, 8):
            kb_per = (num_kb + sp - 1) // sp
            if sp * per <= kw["workspace_bytes"] and num_kb // sp >= 16 and (sp - 1) * kb_per < num_kb:
                splits.append(sp)
    out = []
    flavours = [1] + ([2] if bf16 and M > 128 and N >= 128 and not kw.get("transposed_out") else [])
    for fl in flavours:
        for bn in (256, 192, 128, 64, 32):
            if bn < 128 and (mn_major or fl == 2):
                continue
            if bn < 64 and qkv:
                continue
            if fl == 2 and kw.get("b_mn_major") and bn == 192:
                continue
            if bn > 32 and bn // 2 >= N:
                continue
            for sp in splits:
                out.append(dict(hint_flavour=fl, hint_bn=bn, hint_splits=sp))
    return out


def _time(kw: dict, device: torch.device, runs: int = 3) -> float:
    """Median CUDA-event time (us) of vy_gemm(kw) with cold L2."""
    stream = torch.cuda.current_stream(device)
    _lib.call("vy_gemm", "VyGemm", **kw)  # warm-up: tensor maps, function attributes, instruction cache
    evs = []
    for _ in range(runs):
        _flush_l2(device)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _lib.call("vy_gemm", "VyGemm", **kw)
        e1.record(stream)
        evs.append((e0, e1))
    torch.cuda.synchronize(device)
    ts = sorted(e0.elapsed_time(e1) * 1e3 for e0, e1 in evs)
    return ts[len(ts) // 2]


def hints(key: tuple, kw: dict, written: List[Tuple[str, Optional[torch.Tensor]]], device: torch.device) -> Dict[str, int]:
    """Tiling hints for the vy_gemm call described by `kw`. `written` lists (pointer field, tensor) for every buffer
    the call writes; during tuning each is replaced by a scratch tensor of the same shape and strides."""
    if not ENABLED:
        return {}
    hit = _CACHE.get(key)
    if hit is not None:
        return hit
    if repr(key) in _FILE_CACHE:
        _CACHE[key] = _FILE_CACHE[repr(key)]
        return _CACHE[key]
    if torch.cuda.is_current_stream_capturing() or _lib.TIMER is not None:
        return {}
    cands = _candidates(kw)
    if len(cands) <= 1:
        _CACHE[key] = {}
        return {}
    scratch_bytes = 0
    for _, t in written:
        if t is not None and t.numel():
            scratch_bytes += (sum((n - 1) * st for n, st in zip(t.size(), t.stride())) + 1) * t.element_size()
    if scratch_bytes > MAX_SCRATCH_BYTES:
        _CACHE[key] = {}
        return {}
    trial = dict(kw)
    keep = []
    for field, t in written:
        if t is None:
            continue
        s = torch.empty_strided(t.size(), t.stride(), dtype=t.dtype, device=t.device)
        keep.append(s)
        trial[field] = s.data_ptr()
    # pass 1: every candidate, 3 runs each; pass 2: the model's own choice and the three fastest candidates again with 7
    # runs — single measurements of 20-100 us kernels scatter by several percent, and a wrong pick costs every step
    model_us = _time(trial, device)
    timed = []
    for c in cands:
        t = dict(trial)
        t.update(c)
        timed.append((_time(t, device), c))
    timed.sort(key=lambda e: e[0])
    model_us = min(model_us, _time(trial, device, runs=7))
    best, best_us = {}, model_us * 0.98  # a hint has to beat the model's own choice by more than the residual noise
    for _, c in timed[:3]:
        t = dict(trial)
        t.update(c)
        us = _time(t, device, runs=7)
        if us < best_us:
            best, best_us = c, us
    del keep
    _CACHE[key] = best
    LOG.append((key, best, best_us if best else model_us, model_us))
    if CACHE_FILE:
        _FILE_CACHE[repr(key)] = best
        try:
            with open(CACHE_FILE, "w") as f:
                json.dump(_FILE_CACHE, f, indent=0)
        except OSError:
            pass
    return best
