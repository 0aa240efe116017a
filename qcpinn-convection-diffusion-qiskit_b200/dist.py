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


class PeerAllReduce:
    """Average + clip of the flat float32 gradient over the ranks of ONE node in a single kernel per
    rank (``qcp_peer_allreduce_clip``: P2P stores into every rank's symmetric buffer, flag exchange,
    rank-ordered sum, 1 / world, clip) instead of NCCL all-reduce + clip kernel -- the exchange is
    3 KB, i.e. pure latency.  Buffers come from ``torch.distributed._symmetric_memory`` (peer-mapped
    over NVLink / NVSwitch).  ``PeerAllReduce.create`` returns None when anything it needs is missing
    (no CUDA, one rank, symmetric memory unavailable, self test against NCCL fails): callers then
    keep the NCCL route.  ``QCP_PEER_ALLREDUCE=0`` disables it.
    """

    TIMEOUT_S = 5.0

    def __init__(self, n_values, device, group):
        import ctypes

        import torch.distributed._symmetric_memory as symm

        from . import _lib

        self.lib = _lib.require_cuda()
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = torch.device(device)
        self.n_values = int(n_values)
        floats = self.lib.qcp_peer_allreduce_floats(self.n_values, self.world)
        if floats <= 0:
            raise RuntimeError(f"peer all-reduce supports up to 16 ranks, got {self.world}")
        self.buf = symm.empty(int(floats), dtype=torch.float32, device=self.device)
        self.handle = symm.rendezvous(self.buf, self.group)
        self.buf.zero_()
        self.seq = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)                       # nobody pushes into a buffer not yet zeroed
        ptrs = list(self.handle.buffer_ptrs)
        if len(ptrs) != self.world or not all(ptrs):
            raise RuntimeError("symmetric memory rendezvous returned no peer pointers")
        self._ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self.calls = 0

    @classmethod
    def create(cls, n_values, device, group=None, max_norm=1.0):
        if os.environ.get("QCP_PEER_ALLREDUCE", "1") == "0":
            return None
        if not (dist.is_available() and dist.is_initialized()) or torch.device(device).type != "cuda":
            return None
        if dist.get_world_size(group) < 2 or dist.get_backend(group) != "nccl":
            return None
        ok = torch.ones(1, dtype=torch.int32, device=device)
        peer = None
        try:
            peer = cls(n_values, device, group)
            ok.fill_(1 if peer.self_test(max_norm) else 0)
        except Exception as exc:                            # symmetric memory unavailable on this box
            ok.fill_(0)
            peer = None
            cls.last_failure = f"{type(exc).__name__}: {exc}"
        # all ranks must agree: a mixed NCCL / peer exchange would deadlock
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        return peer if int(ok.item()) == 1 else None

    last_failure = None

    def reduce_clip(self, flat, n_grad, n_extra, max_norm):
        """In place on ``flat[: n_grad + n_extra]`` (float32): mean over the ranks, then the first
        ``n_grad`` values clipped to ``max_norm``.  Stream-ordered, CUDA-graph capturable."""
        import ctypes

        from . import _lib

        if n_grad + n_extra != self.n_values or flat.dtype != torch.float32 or not flat.is_contiguous():
            raise ValueError("PeerAllReduce.reduce_clip: buffer does not match the exchange plan")
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            rc = self.lib.qcp_peer_allreduce_clip(
                ctypes.c_void_p(flat.data_ptr()), int(n_grad), int(n_extra), self._ptrs, self.rank,
                self.world, ctypes.c_void_p(self.seq.data_ptr()), float(max_norm),
                ctypes.c_void_p(self.error.data_ptr()), float(self.TIMEOUT_S), stream)
        _lib.check(rc, "qcp_peer_allreduce_clip")
        self.calls += 1

    def close(self):
        """Release the symmetric buffer (collective-free; call before destroy_process_group)."""
        if self.buf is not None:
            torch.cuda.synchronize(self.device)
            self.handle = None
            self.buf = None
            self._ptrs = None

    def check(self):
        """Raises if a peer ever failed to arrive (synchronises)."""
        code = int(self.error.item())
        if code:
            raise RuntimeError(f"peer all-reduce: rank {code - 1} did not arrive within "
                               f"{self.TIMEOUT_S} s (rank {self.rank})")

    def self_test(self, max_norm=1.0):
        """Two exchanges of rank-dependent vectors (both slot parities) against NCCL + the clip
        formula; True when they agree to float32 rounding and no peer timed out."""
        n = self.n_values
        n_grad = max(n - 1, 1) if n > 1 else 1
        n_extra = n - n_grad
        good = True
        for trial in range(2):
            g = torch.Generator(device="cpu").manual_seed(1000 * trial + self.rank)
            v = (torch.randn(n, generator=g) * (3.0 if trial else 0.01)).to(self.device)
            want = v.clone()
            dist.all_reduce(want, op=dist.ReduceOp.SUM, group=self.group)
            want /= self.world
            norm = want[:n_grad].double().norm().float()
            coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
            want[:n_grad] *= coef
            got = v.clone()
            self.reduce_clip(got, n_grad, n_extra, max_norm)
            torch.cuda.synchronize(self.device)
            err = float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))
            good = good and err < 1e-5 and int(self.error.item()) == 0
        return good
