"""Data-parallel training of the captioner on the fused sm_100a path.

What the reference's notebooks get from `accelerate` (DDP + AdamW + clip_grad_norm_(1.0);
Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1 `main()`) is rebuilt B200-first:

  * one process per GPU; every parameter lives in ONE flat buffer and every gradient in ONE flat
    buffer of the same layout (q|k|v projection weights adjacent, so the packed-QKV GEMM needs no
    repacking), which makes the optimizer a single fused kernel and the gradient exchange a
    handful of large NCCL all-reduces over NVLink / NVSwitch instead of one per tensor;
  * the all-reduce runs bucket by bucket on a side stream as soon as backward has produced a
    bucket (post-accumulate hooks), so the transfer overlaps the remaining backward kernels;
  * AdamW (fp32 master weights + fp32 moments, bf16 or fp32 model weights), the global-norm clip
    and the 1/world_size mean are fused into vy_sqnorm + vy_adamw; the clip coefficient never
    leaves the device.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Set, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from .autograd_train import cross_entropy
from .layers.attention import _SelfAttentionBase


def _param_order(model: nn.Module) -> List[nn.Parameter]:
    """Parameters in flat-buffer order: projection weights of an attention module adjacent (then their
    biases adjacent), everything else in module order; shared Parameters appear once."""
    seen, order = set(), []

    def add(p):
        if p is not None and id(p) not in seen:
            seen.add(id(p))
            order.append(p)

    for mod in model.modules():
        if isinstance(mod, _SelfAttentionBase):
            lin = mod._packed()
            for l in lin:
                add(l.weight)
            for l in lin:
                add(l.bias)
    for p in model.parameters():
        add(p)
    return order


class FlatParams:
    """Re-homes every parameter (and its .grad) of `model` inside two flat buffers."""

    ALIGN = 8  # elements; keeps every tensor 16-byte aligned for bf16 and fp32

    def __init__(self, model: nn.Module, alloc=None):
        """`alloc(numel, dtype) -> zero-filled 1-D tensor` places the two buffers (default: ordinary device memory; the
        peer-memory data-parallel step passes a symmetric-memory allocator)."""
        params = _param_order(model)
        dtypes = {p.dtype for p in params}
        if len(dtypes) != 1:
            raise ValueError(f"FlatParams needs one parameter dtype, got {dtypes}")
        self.dtype = params[0].dtype
        dev = params[0].device
        self.params = params
        self.offsets = []
        off = 0
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.numel = off
        self.zero_ranges = None  # None: zero_grad clears everything; else the [start, end) slices it must clear
        if alloc is None:
            self.flat = torch.zeros(off, device=dev, dtype=self.dtype)
            self.grad = torch.zeros(off, device=dev, dtype=self.dtype)
        else:
            self.flat = alloc(off, self.dtype)
            self.grad = alloc(off, self.dtype)
        for p, o in zip(params, self.offsets):
            view = self.flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def zero_grad(self) -> None:
        if self.zero_ranges is None:
            self.grad.zero_()
        else:  # overwrite-mode parameters need no clearing: zero only the slices that are accumulated into
            for a, b in self.zero_ranges:
                self.grad[a:b].zero_()
        es = self.grad.element_size()
        base = self.grad.data_ptr()
        for p, o in zip(self.params, self.offsets):  # re-attach in case autograd replaced a .grad
            if p.grad is None or p.grad.data_ptr() != base + o * es:
                p.grad = self.grad[o:o + p.numel()].view(p.shape)


_DP_TRACE = os.environ.get("VY_DP_TRACE", "0") != "0"  # development: log every gradient-ready event and bucket launch


class GradExchange:
    """Bucketed sum all-reduce of one flat gradient buffer. Buckets are contiguous slices walked from
    the END of the buffer (the order backward fills it); on CUDA the collectives run on a side
    stream so they overlap the rest of backward. Works with any torch.distributed backend (NCCL on
    the B200 box, gloo in the CPU tests)."""

    def __init__(self, flat_grad: torch.Tensor, bucket_elems: int, param_starts: Optional[List[int]] = None, enabled: bool = True):
        """`enabled=False`: the gradients are exchanged elsewhere (dp_shard.ShardedStep) and every call here is a no-op.
        `param_starts`: offsets at which parameters begin in the flat buffer. Bucket boundaries then snap to them (a
        bucket = whole parameters, at least `bucket_elems` elements unless a single parameter is larger), so that
        "every parameter that starts in the bucket is ready" means every element of the bucket has been written — with
        free-running boundaries the tail of a parameter lying in the next bucket would be reduced before backward has
        produced it."""
        self.grad = flat_grad
        self.world = dist.get_world_size() if enabled and dist.is_available() and dist.is_initialized() else 1
        self.buckets: List[Tuple[int, int]] = []
        cuts = sorted(set(param_starts)) if param_starts else None
        end = flat_grad.numel()
        while end > 0:
            start = max(0, end - bucket_elems)
            if cuts is not None:
                import bisect
                i = bisect.bisect_right(cuts, start) - 1  # the last parameter start at or before the free-running boundary
                start = cuts[i] if i >= 0 else 0
                assert start < end
            self.buckets.append((start, end))
            end = start
        self._launched: Set[int] = set()
        self._stream = torch.cuda.Stream(device=flat_grad.device) if (flat_grad.is_cuda and self.world > 1) else None

    def bucket_of(self, offset: int) -> int:
        return next(i for i, (s, e) in enumerate(self.buckets) if s <= offset < e)

    def begin_step(self) -> None:
        self._launched = set()

    def launch(self, b: int) -> None:
        if self.world == 1 or b in self._launched:
            return
        s, e = self.buckets[b]
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                dist.all_reduce(self.grad[s:e], op=dist.ReduceOp.SUM)
        else:
            dist.all_reduce(self.grad[s:e], op=dist.ReduceOp.SUM)
        self._launched.add(b)

    def finish(self) -> None:
        if self.world == 1:
            return
        for b in range(len(self.buckets)):
            self.launch(b)
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)


class Trainer:
    """AdamW + clip + data-parallel all-reduce around a model built from vyomai_b200 modules."""

    def __init__(self, model: nn.Module, lr: float = 1e-5, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                 max_grad_norm: float = 1.0, bucket_mb: float = 64.0, overlap: bool = True, use_graph: bool = False,
                 grad_overwrite: bool = True, dp_mode: Optional[str] = None):
        """dp_mode (world size > 1): "p2p" — the sharded optimizer step over NVLink peer memory (dp_shard.py: no NCCL kernel
        in the step, optimizer state and its traffic divided by the world size); "nccl" — bucketed, overlapped NCCL
        all-reduce + the full AdamW on every rank. Default: p2p when symmetric memory is available (VY_DP_MODE overrides)."""
        self.model = model
        self.use_graph = use_graph
        self._graph = None
        self._graph_key = None
        self.graph_kernels = 0       # kernels of this library inside the captured step
        self.replayed_kernels = 0    # ... launched so far through graph replays (vy_launch_count only sees eager calls)
        from . import dp_shard
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if dp_mode is None:
            dp_mode = os.environ.get("VY_DP_MODE", "p2p")
        self.dp_mode = dp_mode if (world > 1 and dp_mode == "p2p" and dp_shard.available()) else ("nccl" if world > 1 else "single")
        self.shard = None
        if self.dp_mode == "p2p":
            dev0 = next(model.parameters()).device
            alloc = dp_shard.SymmetricAllocator(dev0)
            ptrs = {}

            def symm_zeros(numel, dtype):
                t, pp = alloc.zeros(numel, dtype)
                ptrs[t.data_ptr()] = pp
                return t

            self.fp = FlatParams(model, alloc=symm_zeros)
            self.shard = dp_shard.ShardedStep.from_process_group(
                alloc, self.fp.flat, ptrs[self.fp.flat.data_ptr()], self.fp.grad, ptrs[self.fp.grad.data_ptr()],
                lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
            overlap = False
        else:
            self.fp = FlatParams(model)
        n = self.fp.numel
        dev = self.fp.flat.device
        if self.shard is None:
            self.master = self.fp.flat.to(torch.float32).clone() if self.fp.dtype != torch.float32 else None
            self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
            self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        else:  # optimizer state exists for this rank's shard only (dp_shard.ShardedStep)
            self.master = self.exp_avg = self.exp_avg_sq = None
        self.sqnorm = torch.zeros(1, device=dev, dtype=torch.float32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        ops.DropoutState.step_ptr = self.step_dev  # dropout masks follow the device step counter (fresh per graph replay)
        self.lr, self.betas, self.eps, self.wd, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.step_count = 0
        per = max(1, int(bucket_mb * 1024 * 1024 / self.fp.grad.element_size()))
        self.exchange = GradExchange(self.fp.grad, per, param_starts=list(self.fp.offsets), enabled=self.shard is None)
        self.world = world
        self.overlap = overlap and self.world > 1
        self._pending: Dict[int, int] = {}
        self._ready_ids: Set[int] = set()
        self._bucket_of: Dict[int, int] = {}
        self._bucket_size: List[int] = [0] * len(self.exchange.buckets)
        for p, o in zip(self.fp.params, self.fp.offsets):
            # the backward kernels accumulate straight into the flat gradient buffer (autograd_train._direct);
            # parameters whose gradient still arrives through autograd get the same bucket hook
            p._vy_direct_grad = True
            if self.overlap:
                b = self.exchange.bucket_of(o)
                self._bucket_of[id(p)] = b
                self._bucket_size[b] += 1
                p._vy_grad_ready = self._on_grad
                p.register_post_accumulate_grad_hook(self._on_grad)
        if not self.overlap:  # (the overlapped NCCL exchange launches a bucket at `_ready`, before a side-stream sum would be done)
            for p in self.fp.params:
                if p.dim() == 1:
                    p._vy_side_bgrad = True
        self.grad_overwrite = False
        if grad_overwrite:
            self._enable_grad_overwrite()

    def _enable_grad_overwrite(self) -> None:
        """Linear / LayerNorm parameters of the fused blocks get their gradient from exactly one kernel per step
        (wgrad GEMM, column sum, LayerNorm reduce): those kernels may write instead of accumulate, which removes the
        zero fill of ~3/4 of the gradient buffer and the read-modify-write in the wgrad epilogues. Everything else
        (embeddings: scatter-add; ViT stem: autograd accumulation) keeps zero + accumulate. `_verify_grad_overwrite`
        checks the "exactly once" premise on the first step by poisoning the buffer."""
        from .layers.ffn import FeedForward
        from .models._common import LMHead
        from .models.encoder_decoder import LMHead as LMHeadVocab  # same head, projection named `vocab`
        marked = set()

        def mark(*ps):
            for q in ps:
                if q is not None:
                    q._vy_grad_overwrite = True
                    marked.add(id(q))

        for mod in self.model.modules():
            if isinstance(mod, _SelfAttentionBase):
                for l in mod._packed():
                    mark(l.weight, l.bias)
                mark(mod.out.dense.weight, mod.out.dense.bias, mod.out.layernorm.weight, mod.out.layernorm.bias)
            elif isinstance(mod, FeedForward):
                mark(mod.intermediate.weight, mod.intermediate.bias, mod.out.weight, mod.out.bias, mod.layernorm.weight,
                     mod.layernorm.bias)
            elif isinstance(mod, (LMHead, LMHeadVocab)):
                mark(mod.dense.weight, mod.dense.bias, mod.layer_norm.weight, mod.layer_norm.bias, mod.decoder.weight, mod.bias)
        ranges = []
        for p, o in zip(self.fp.params, self.fp.offsets):
            end = o + (p.numel() + FlatParams.ALIGN - 1) // FlatParams.ALIGN * FlatParams.ALIGN
            start = o + p.numel() if id(p) in marked else o  # a written parameter still has its alignment padding cleared
            if start == end:
                continue
            if ranges and ranges[-1][1] == start:
                ranges[-1][1] = end
            else:
                ranges.append([start, end])
        self.fp.zero_ranges = [tuple(r) for r in ranges]
        self.grad_overwrite = True
        self._overwrite_verified = False

    def _disable_grad_overwrite(self) -> None:
        for p in self.fp.params:
            if getattr(p, "_vy_grad_overwrite", False):
                p._vy_grad_overwrite = False
        self.fp.zero_ranges = None
        self.grad_overwrite = False

    def _on_grad(self, p: nn.Parameter) -> None:
        # A parameter is counted once per step: the kernels that write a gradient straight into the flat buffer report
        # it by hand (autograd_train._ready), and autograd's post-accumulate hook ALSO fires for such a parameter when
        # its Function returns None for it — counted twice, a bucket would be reduced before its last writer has run.
        if id(p) in self._ready_ids:
            return
        self._ready_ids.add(id(p))
        b = self._bucket_of[id(p)]
        self._pending[b] = self._pending.get(b, 0) + 1
        if _DP_TRACE:
            names = getattr(self, "_names", None) or {id(q): n for n, q in self.model.named_parameters()}
            self._names = names
            print(f"[dp-trace] ready {names.get(id(p), '?')} bucket {b} {self._pending[b]}/{self._bucket_size[b]}"
                  f"{' -> all-reduce' if self._pending[b] == self._bucket_size[b] else ''}", flush=True)
        if self._pending[b] == self._bucket_size[b]:
            self.exchange.launch(b)

    def zero_grad(self) -> None:
        self.fp.zero_grad()
        self._pending = {}
        self._ready_ids = set()
        self.exchange.begin_step()

    def _check_gemm_health(self) -> None:
        """A GEMM whose barrier wait timed out finishes with undefined results and raises a per-device flag; its host
        mirror is read here without synchronising, so a corrupted step stops the run instead of training on."""
        from . import _lib
        if _lib.lib().vy_gemm_poison_peek() != 0:
            raise _lib.VyomError("a barrier wait inside a vy_gemm kernel timed out on this device: the parameters updated since "
                                 "then are invalid (vy_gemm_poisoned() acknowledges the flag)")

    def optimizer_step(self) -> None:
        ops.side_join()  # bias-gradient column sums still running on the side stream
        self._check_gemm_health()
        if self.shard is not None:
            if self.step_count % 64 == 63:  # a device-side barrier that timed out left the replicas inconsistent: stop
                self.shard.check()           # (one host sync every 64 steps)
            self.step_count += 1
            self.step_dev.add_(1)
            self.shard.step(self.step_count, self.step_dev)
            return
        self.exchange.finish()
        self.step_count += 1
        self.step_dev.add_(1)  # device-side step counter: the whole step can live in a replayed CUDA graph
        self.sqnorm.zero_()
        ops.sqnorm(self.fp.grad, self.sqnorm)
        ops.adamw(self.fp.flat, self.fp.grad, self.exp_avg, self.exp_avg_sq, lr=self.lr, beta1=self.betas[0],
                  beta2=self.betas[1], eps=self.eps, weight_decay=self.wd, step=self.step_count, step_ptr=self.step_dev,
                  master=self.master, grad_sqnorm=self.sqnorm, max_grad_norm=self.max_grad_norm, grad_div=float(self.world))

    def _caption_body(self, pixel_values, input_ids, attention_mask, labels_full) -> torch.Tensor:
        if self.grad_overwrite and not self._overwrite_verified:
            # first step: poison the buffer; any overwrite-mode slice that no kernel wrote would stay NaN
            self._overwrite_verified = True
            if not torch.cuda.is_current_stream_capturing():
                self.fp.grad.fill_(float("nan"))
                self.zero_grad()
                loss = self._forward_backward(pixel_values, input_ids, attention_mask, labels_full)
                if not bool(torch.isfinite(self.fp.grad).all()):
                    import warnings
                    names = {id(p): n for n, p in self.model.named_parameters()}
                    bad = [f"{names.get(id(p), '?')}[{int((~torch.isfinite(p.grad)).sum())}/{p.numel()}]"
                           for p in self.fp.params if not bool(torch.isfinite(p.grad).all())]
                    warnings.warn("gradient overwrite mode switched off: after the poisoned first step these gradients still hold "
                                  f"non-finite values: {bad[:6]}{' ...' if len(bad) > 6 else ''} (none listed = alignment padding)")
                    self._disable_grad_overwrite()  # some parameter is not written exactly once per step: accumulate instead
                    self.zero_grad()
                    loss = self._forward_backward(pixel_values, input_ids, attention_mask, labels_full)
                self.optimizer_step()
                return loss.detach()
        self.zero_grad()
        loss = self._forward_backward(pixel_values, input_ids, attention_mask, labels_full)
        self.optimizer_step()
        return loss.detach()

    def _forward_backward(self, pixel_values, input_ids, attention_mask, labels_full) -> torch.Tensor:
        if hasattr(self.model, "forward_loss"):  # LM head + cross-entropy as one autograd node
            loss = self.model.forward_loss(pixel_values, input_ids, attention_mask, labels_full, ignore_index=-100)
        else:
            logits = self.model(pixel_values=pixel_values, decoder_input_ids=input_ids, decoder_attention_mask=attention_mask).logits
            loss = cross_entropy(logits, labels_full, ignore_index=-100)
        loss.backward()
        return loss

    def caption_step(self, pixel_values: torch.Tensor, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                     labels_full: torch.Tensor) -> torch.Tensor:
        """One captioner training step (VisionLanguageModel): forward, shifted token cross-entropy,
        backward, gradient all-reduce, clip, AdamW. `labels_full` is [B, S+1] aligned with the logits rows
        (image position and the last position carry ignore_index). Returns the (local) loss tensor.

        With use_graph the first call of a given input shape runs three eager warm-up steps and captures
        the whole step (every kernel, the NCCL all-reduces on the side stream, the optimizer) into one CUDA
        graph; later calls copy the inputs into the graph's static buffers and replay it, so the ~400
        launches of a step cost one cudaGraphLaunch."""
        if not self.use_graph:
            return self._caption_body(pixel_values, input_ids, attention_mask, labels_full)
        key = (tuple(pixel_values.shape), tuple(input_ids.shape))
        if self._graph_key != key:
            self._capture(key, pixel_values, input_ids, attention_mask, labels_full)
        for dst, src in zip(self._static_in, (pixel_values, input_ids, attention_mask, labels_full)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._check_gemm_health()  # the replayed kernels make no host-side vy_gemm call that would notice the flag
        self._replays = getattr(self, "_replays", 0) + 1
        if self.shard is not None and self._replays % 64 == 0:
            self.shard.check()
        self._graph.replay()
        self.replayed_kernels += self.graph_kernels
        return self._static_loss

    def _capture(self, key, pixel_values, input_ids, attention_mask, labels_full) -> None:
        self._static_in = [t.clone() for t in (pixel_values, input_ids, attention_mask, labels_full)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):  # these are real optimisation steps (the caller's warm-up)
                self._caption_body(*self._static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        n0 = _lib.lib().vy_launch_count()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss = self._caption_body(*self._static_in)
        self.graph_kernels = int(_lib.lib().vy_launch_count() - n0)
        self._graph_key = key


class HostPrefetcher:
    """Feeds pinned host batches to the device one step ahead: the host->device copy of batch i+1 runs on a copy
    stream while step i computes (what a DataLoader with pin_memory + non_blocking copies gives the reference's
    notebooks). `next()` returns device tensors of the batch whose copy was issued by the previous call and
    immediately issues the copy of the following batch; the compute stream waits on the copy's event, never the host."""

    def __init__(self, batches, device: torch.device):
        self.batches = batches  # indexable / cyclic source of tuples of pinned CPU tensors
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self._i = 0
        self._pending = self._issue()

    def _issue(self):
        host = self.batches[self._i % len(self.batches)]
        self._i += 1
        with torch.cuda.stream(self.stream):
            dev = tuple(t.to(self.device, non_blocking=True) for t in host)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ev

    def next(self):
        dev, ev = self._pending
        torch.cuda.current_stream().wait_event(ev)
        for t in dev:
            t.record_stream(torch.cuda.current_stream())
        self._pending = self._issue()
        return dev


def caption_labels(input_ids: torch.Tensor, attention_mask: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """labels aligned with the captioner's logits rows: the logits have one extra leading (image)
    position, so row t (1 <= t <= S-1) predicts text token t; pads, row 0 and row S are ignored — the
    `loss_fn` of Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1 (shifted CE over non-pad labels)."""
    B, S = input_ids.shape
    full = torch.full((B, S + 1), ignore_index, dtype=torch.long, device=input_ids.device)
    lab = input_ids.masked_fill(attention_mask == 0, ignore_index)
    full[:, 1:S] = lab[:, 1:]
    return full
