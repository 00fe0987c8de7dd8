"""FusedAdam: torch.optim.Adam(lr, weight_decay) semantics (reference train.py:79) as ONE kernel per step.

Parameters and gradients are flattened into two contiguous fp32 buffers at construction (each `param.data` / `param.grad`
becomes a view into them), so autograd accumulates straight into the flat gradient buffer, a data-parallel all-reduce is
a single collective over it, and the update is one pass of `hk_adam_step`.  All 21,797,672 parameters are updated,
including the 996 dead fc rows, which receive zero gradient but are still decayed -- as in the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable

import torch

from ._lib import check, lib, note_raw_parameter_write, ptr, stream_ptr


class FusedAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdam got no parameters")
        dev = self.params[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise RuntimeError("FusedAdam needs fp32 CUDA parameters on one device (no CPU fallback)")
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        # every tensor starts on a 16-byte boundary inside the flat buffers
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.numel = total
        self.flat_param = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                n = p.numel()
                self.flat_param[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.flat_param[o:o + n].view_as(p.data)
                p.grad = self.flat_grad[o:o + n].view_as(p.data)

    def adopt_grad_buffer(self, flat_grad: torch.Tensor) -> None:
        """Use an external flat gradient buffer with this optimiser's layout (TrainEngine.flat_grad): the engine's kernels then write
        the gradients exactly where the all-reduce and hk_adam_step read them -- no copy."""
        if flat_grad.shape != self.flat_grad.shape or flat_grad.dtype != torch.float32 or flat_grad.device != self.flat_grad.device:
            raise ValueError("gradient buffer layout mismatch")
        self.flat_grad = flat_grad
        o = 0
        for p in self.params:
            n = p.numel()
            p.grad = flat_grad[o:o + n].view_as(p.data)
            o += (n + 3) // 4 * 4

    def zero_grad(self, set_to_none: bool = False) -> None:
        """Keeps the gradient views alive (set_to_none would detach them from the flat buffer)."""
        self.flat_grad.zero_()

    def all_reduce_grads(self) -> None:
        """Data-parallel exchange: ONE sum all-reduce over the flat gradient buffer, then the 1/world average."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
            self.flat_grad.div_(dist.get_world_size())

    def all_reduce_range_async(self, begin: int, end: int):
        """Start the sum all-reduce of flat_grad[begin:end] on NCCL's stream (it waits for the work already enqueued on the current
        stream); returns the work handle, or None outside a multi-rank job.  finish_all_reduce() waits and averages."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or end <= begin:
            return None
        return dist.all_reduce(self.flat_grad[begin:end], op=dist.ReduceOp.SUM, async_op=True)

    def finish_all_reduce(self, works) -> None:
        import torch.distributed as dist
        works = [w for w in works if w is not None]
        for w in works:
            w.wait()
        if works:
            self.flat_grad.div_(dist.get_world_size())

    @torch.no_grad()
    def step(self) -> None:
        self.step_count += 1
        check(lib().hk_adam_step(ptr(self.flat_param), ptr(self.flat_grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                 C.c_longlong(self.numel), C.c_float(self.lr), C.c_float(self.betas[0]), C.c_float(self.betas[1]),
                                 C.c_float(self.eps), C.c_float(self.weight_decay), self.step_count, stream_ptr()), "hk_adam_step")
        note_raw_parameter_write()   # raw-pointer write: tensor versions do not move, eval engines must refold their weights
