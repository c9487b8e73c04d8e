"""Data-parallel optimizer step over NVLink peer memory (csrc/dp_step.cu): every rank reduces and updates ONE shard of the
flat parameter buffer by reading the other ranks' gradients and writing the other ranks' parameters directly — the
replacement, inside one node, for "NCCL all-reduce of the whole gradient + the whole AdamW on every rank" that
accelerate / DDP give the reference's notebooks (Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1 main()).

The flat parameter and gradient buffers live in symmetric memory (torch.distributed._symmetric_memory: CUDA VMM allocations
mapped into every rank of the node), so a pointer table per buffer is all the kernels need. torch.distributed stays the
plumbing: rendezvous and handle exchange; the step itself launches no NCCL kernel.
"""
from __future__ import annotations

import ctypes
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib

MAX_WORLD = 8


def available() -> bool:
    """Symmetric memory needs CUDA, an initialised process group of <= 8 ranks and a torch build that ships it."""
    if not (torch.cuda.is_available() and dist.is_available() and dist.is_initialized()):
        return False
    if dist.get_world_size() < 2 or dist.get_world_size() > MAX_WORLD:
        return False
    try:
        import importlib
        importlib.import_module("torch.distributed._symmetric_memory")
    except Exception:
        return False
    if dist.get_backend() != "nccl":
        return False
    return _probe()


_PROBE = None


def _probe() -> bool:
    """Collective, once per process: can every rank allocate and rendezvous a small symmetric buffer? All ranks get the same
    answer (MIN over ranks), so the trainers agree on the data-parallel mode; a failure (no peer access, allocator not
    supported) selects the NCCL all-reduce path and says so on stderr."""
    global _PROBE
    if _PROBE is not None:
        return _PROBE
    import sys
    import torch.distributed._symmetric_memory as symm
    dev = torch.device("cuda", torch.cuda.current_device())
    ok, why = 1, ""
    try:
        t = symm.empty(64, dtype=torch.float32, device=dev)
        h = symm.rendezvous(t, dist.group.WORLD)
        ok = 1 if len(h.buffer_ptrs) == dist.get_world_size() else 0
    except Exception as e:  # noqa: BLE001
        ok, why = 0, f"{type(e).__name__}: {str(e)[:200]}"
    flag = torch.tensor([ok], device=dev, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    _PROBE = bool(int(flag.item()))
    if not _PROBE and dist.get_rank() == 0:
        print(f"[vyomai_b200] symmetric memory unavailable ({why or 'on another rank'}): data-parallel step falls back to the NCCL "
              "all-reduce path", file=sys.stderr, flush=True)
    return _PROBE


class SymmetricAllocator:
    """Allocates tensors that every rank of the group can address: `empty(n, dtype)` returns (local tensor, [pointer of rank
    0's copy, rank 1's, ...] as seen from this rank). Collective: every rank must make the same calls in the same order."""

    def __init__(self, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm
        self._symm = symm
        self.device = device
        self.group = group if group is not None else dist.group.WORLD
        self._handles = []  # keep the rendezvous handles (and with them the peer mappings) alive

    def empty(self, numel: int, dtype: torch.dtype) -> Tuple[torch.Tensor, List[int]]:
        t = self._symm.empty(numel, dtype=dtype, device=self.device)
        h = self._symm.rendezvous(t, self.group)
        self._handles.append(h)
        return t, [int(p) for p in h.buffer_ptrs]

    def zeros(self, numel: int, dtype: torch.dtype) -> Tuple[torch.Tensor, List[int]]:
        t, ptrs = self.empty(numel, dtype)
        t.zero_()
        return t, ptrs


def shard_bounds(numel: int, world: int, rank: int) -> Tuple[int, int]:
    """Rank `rank`'s contiguous shard of a flat buffer of `numel` elements: equal parts rounded up to 8 elements (one 16-byte
    bf16 vector), the last one shorter."""
    per = (numel + world - 1) // world
    per = (per + 7) // 8 * 8
    lo = min(numel, rank * per)
    return lo, min(numel, lo + per)


def _table(ptrs: List[int]):
    return (ctypes.c_void_p * len(ptrs))(*ptrs)


class ShardedStep:
    """One rank's state of the sharded optimizer step: the peer pointer tables, its shard [lo, hi) of the flat buffers and
    the fp32 master weights / AdamW moments of that shard only."""

    def __init__(self, world: int, rank: int, flat: torch.Tensor, flat_ptrs: List[int], grad: torch.Tensor, grad_ptrs: List[int],
                 flags: torch.Tensor, flag_ptrs: List[int], scalars: torch.Tensor, scalar_ptrs: List[int],
                 lr: float, betas, eps: float, weight_decay: float, max_grad_norm: float):
        """`*_ptrs[r]`: address of rank r's copy of that buffer as seen from THIS rank (symmetric memory; in the single-GPU
        tests simply r's own tensors on the same device). flags: >= world int32, scalars: >= world fp32, zero-filled."""
        self.world, self.rank = world, rank
        dev = flat.device
        n = flat.numel()
        self.lo, self.hi = shard_bounds(n, world, rank)
        if self.hi <= self.lo or n % 8:
            raise _lib.VyomError("sharded data-parallel step: flat buffer too small / not a multiple of 8 elements")
        self.flat, self.grad, self.flags, self.scalars = flat, grad, flags, scalars
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.error = torch.zeros(1, dtype=torch.int32, device=dev)
        m = self.hi - self.lo
        self.gshard = torch.zeros(m, dtype=torch.float32, device=dev)
        self.sq_local = torch.zeros(1, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(m, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(m, dtype=torch.float32, device=dev)
        self.master = flat[self.lo:self.hi].to(torch.float32).clone() if flat.dtype != torch.float32 else None
        self.lr, self.betas, self.eps, self.wd, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self._tables = [_table(grad_ptrs), _table(flat_ptrs), _table(flag_ptrs), _table(scalar_ptrs)]
        g = _lib.STRUCTS["VyDpGroup"]()
        g.world, g.rank = self.world, self.rank
        g.grads = ctypes.cast(self._tables[0], ctypes.c_void_p)
        g.params = ctypes.cast(self._tables[1], ctypes.c_void_p)
        g.flags = ctypes.cast(self._tables[2], ctypes.c_void_p)
        g.scalars = ctypes.cast(self._tables[3], ctypes.c_void_p)
        g.epoch, g.error_flag = self.epoch.data_ptr(), self.error.data_ptr()
        self._group = g

    @classmethod
    def from_process_group(cls, alloc: SymmetricAllocator, flat, flat_ptrs, grad, grad_ptrs, **hyper) -> "ShardedStep":
        flags, flag_ptrs = alloc.zeros(64, torch.int32)
        scalars, scalar_ptrs = alloc.zeros(64, torch.float32)
        self = cls(dist.get_world_size(), dist.get_rank(), flat, flat_ptrs, grad, grad_ptrs, flags, flag_ptrs, scalars, scalar_ptrs, **hyper)
        torch.cuda.synchronize()
        dist.barrier()  # every rank's flags / scalars are zeroed before anyone starts signalling
        torch.cuda.synchronize()
        return self

    def _stream(self) -> int:
        return torch.cuda.current_stream().cuda_stream

    def barrier(self) -> None:
        L = _lib.lib()
        _lib.check(L.vy_dp_barrier(ctypes.byref(self._group), self._stream()), "vy_dp_barrier")

    def reduce(self) -> None:
        """gshard = sum over ranks of their gradients on [lo, hi) (fp32); the shard's sum of squares goes to every rank."""
        L = _lib.lib()
        r = _lib.STRUCTS["VyDpReduce"]()
        r.lo, r.hi = self.lo, self.hi
        r.dtype = _lib.CONSTS["VY_BF16"] if self.grad.dtype == torch.bfloat16 else _lib.CONSTS["VY_F32"]
        r.gshard, r.sq_local, r.stream = self.gshard.data_ptr(), self.sq_local.data_ptr(), self._stream()
        _lib.check(L.vy_dp_reduce_shard(ctypes.byref(self._group), ctypes.byref(r)), "vy_dp_reduce_shard")

    def adamw(self, step_count: int, step_ptr: Optional[torch.Tensor] = None) -> None:
        """Clip by the global norm of the mean gradient, AdamW on the shard, new parameters stored into every rank."""
        L = _lib.lib()
        a = _lib.STRUCTS["VyDpAdamW"]()
        a.lo, a.hi = self.lo, self.hi
        a.param_dtype = _lib.CONSTS["VY_BF16"] if self.flat.dtype == torch.bfloat16 else _lib.CONSTS["VY_F32"]
        a.gshard, a.exp_avg, a.exp_avg_sq = self.gshard.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        a.master = None if self.master is None else self.master.data_ptr()
        a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = self.lr, self.betas[0], self.betas[1], self.eps, self.wd
        a.max_grad_norm = self.max_grad_norm
        a.step = int(step_count)
        a.step_ptr = None if step_ptr is None else step_ptr.data_ptr()
        a.stream = self._stream()
        _lib.check(L.vy_dp_adamw_shard(ctypes.byref(self._group), ctypes.byref(a)), "vy_dp_adamw_shard")

    def step(self, step_count: int, step_ptr: Optional[torch.Tensor] = None) -> None:
        """After backward (this rank's local gradient complete in its flat buffer): reduce, clip, AdamW, broadcast."""
        self.barrier()  # all ranks' gradients are final
        self.reduce()
        self.barrier()  # every shard's partial norm has been published; nobody still reads gradients
        self.adamw(step_count, step_ptr)
        self.barrier()  # every rank's parameters are complete before the next forward reads them

    def check(self) -> None:
        """Synchronising health check: raises if a device-side barrier timed out (a rank that never arrived)."""
        if int(self.error.item()) != 0:
            raise _lib.VyomError("data-parallel barrier timed out on the device: a rank did not arrive; parameters are invalid")
