"""Data-parallel plumbing: collocation points are sharded over ranks (one process per GPU), the
~3 KB parameter gradient is averaged with ONE flat all-reduce per step (NCCL over NVLink on a
B200 box, gloo in the CPU tests).  The reference has no distributed code;
this is the K10 row of the kernel table.

The all-reduce is issued from an autograd-engine callback queued by a tensor hook on the first
gradient that arrives, i.e. it runs once, after the LAST gradient of ``loss.backward()`` has been
accumulated and before the trainer's ``clip_grad_norm_`` -- so an unmodified trainer stays correct.
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist
from torch.autograd import Variable


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*); returns
    (rank, world, local_rank).  No-op when WORLD_SIZE is absent or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


class GradientAverager:
    """Averages ``.grad`` of every trainable parameter of ``module`` across the process group."""

    def __init__(self, module: torch.nn.Module, process_group=None, extra=0):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        # one persistent flat buffer: [grads..., `extra` scalar slots (loss terms)]
        self.flat = torch.zeros(self.numel + extra, dtype=p0.dtype, device=p0.device)
        self.extra = extra
        self._queued = False
        self.enabled = True
        self.calls = 0
        self._handles = [p.register_hook(self._on_grad) for p in self.params]

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []

    def _on_grad(self, grad):
        if self.enabled and not self._queued:
            self._queued = True
            Variable._execution_engine.queue_callback(self._finalize)
        return grad

    def _finalize(self):
        self._queued = False
        self.average()

    @torch.no_grad()
    def average(self, extras=None):
        """Pack grads (+ optional scalar tensors) -> all-reduce(sum) -> scale 1/W -> unpack.
        Returns the averaged extras (tensor view) when given."""
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        if extras is not None:
            for i, e in enumerate(extras):
                self.flat[self.numel + i].copy_(e.detach().reshape(()))
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / self.world)
        self.calls += 1
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = self.flat[off:off + n].view_as(p).clone()
            else:
                p.grad.copy_(self.flat[off:off + n].view_as(p))
            off += n
        if extras is not None:
            return self.flat[self.numel:self.numel + len(extras)]
        return None
