"""Batch data-parallel plumbing for the deformable-conv op (one process per GPU).

The op shards over the batch: forward, grad_input, grad_offset and grad_mask are per-sample
independent; only grad_weight and grad_bias sum over samples (reference: the per-chunk `+=` into
grad_weight / grad_bias, src/mdeformable_conv.cu:436-444).  So the only exchange step is ONE
all-reduce(sum) of the flat buffer [grad_weight || grad_bias] per backward -- NCCL over NVLink on
GPUs, gloo in the CPU tests.  The reference has no multi-device code at all.
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n samples owned by `rank`; the first n % world ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors: Sequence[Optional[torch.Tensor]], rank: int, world: int) -> List[Optional[torch.Tensor]]:
    """Slice every per-sample tensor (input, offset, mask, grad_output ...) along dim 0."""
    out = []
    for t in tensors:
        if t is None:
            out.append(None)
            continue
        lo, hi = shard_range(t.shape[0], rank, world)
        out.append(t[lo:hi].contiguous())
    return out


class GradBucket:
    """Flat fp32 bucket for the parameter gradients of one or more deformable-conv modules.

    pack() copies the grads in, all_reduce() issues a single collective (optionally on a side
    stream so it overlaps whatever the caller enqueues next), unpack() writes the sums back in the
    gradients' own dtype.  Sum, not mean: identical to running the op on the concatenated batch.
    """

    def __init__(self, params: Sequence[torch.Tensor]):
        self.params = [p for p in params if p is not None]
        self.sizes = [p.numel() for p in self.params]
        dev = self.params[0].device
        self.flat = torch.zeros(sum(self.sizes), dtype=torch.float32, device=dev)
        self._stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._work = None

    def pack(self):
        off = 0
        for p, n in zip(self.params, self.sizes):
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            self.flat[off:off + n].copy_(g.reshape(-1))
            off += n

    def all_reduce(self, group=None, async_op: bool = False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self._stream):
                self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
            if not async_op:
                self.wait()
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self._stream is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._stream)

    def unpack(self):
        off = 0
        for p, n in zip(self.params, self.sizes):
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(self.flat[off:off + n].reshape(p.shape))
            off += n


class GradSync:
    """Overlapped all-reduce of one module's parameter gradients, driven from inside the op's backward.

    The tcgen05 backward computes grad_weight / grad_bias FIRST and records a CUDA event when they are final
    (``mdconv_backward_saved(..., weight_ready_event)``).  ``reduce`` then runs ONE all-reduce(sum) of the two
    tensors on a communication stream that waits for that event only, so the collective overlaps the
    grad_offset / grad_mask / grad_input kernels; the compute stream re-joins before the backward returns, which
    keeps autograd's accumulation into ``.grad`` ordered after the reduction.  Attach with ``attach(module, group)``;
    the reduction is in the gradients' own dtype and in place (as DDP does), sum not mean.
    On CPU tensors / a non-initialised process group it degrades to a plain (or no) all-reduce.
    """

    def __init__(self, group=None, device: Optional[torch.device] = None):
        self.group = group
        self._stream = None
        self.event = None
        if device is not None and device.type == "cuda":
            self._stream = torch.cuda.Stream(device=device)
            self.event = torch.cuda.Event()
            self.event.record(torch.cuda.current_stream(device))   # force creation of the CUDA event handle

    def active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def reduce(self, tensors: Sequence[Optional[torch.Tensor]]) -> None:
        ts = [t for t in tensors if t is not None and t.numel() > 0]
        if not ts or not self.active():
            return
        if self._stream is None:
            for t in ts:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            return
        main = torch.cuda.current_stream(ts[0].device)
        self._stream.wait_event(self.event)          # only grad_weight / grad_bias, not the whole backward
        with torch.cuda.stream(self._stream):
            flat = ts[0] if len(ts) == 1 else torch.cat([t.reshape(-1) for t in ts])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            if len(ts) > 1:
                off = 0
                for t in ts:
                    t.copy_(flat[off:off + t.numel()].reshape(t.shape))
                    off += t.numel()
            for t in ts:
                t.record_stream(self._stream)
        main.wait_stream(self._stream)


def attach(module, group=None) -> "GradSync":
    """Make ``module`` (a DeformConv*/ModulatedDeformConv* module of this package) all-reduce its weight / bias
    gradients inside its own backward, overlapped with the data gradients.  Returns the GradSync."""
    p = next(module.parameters())
    module._grad_sync = GradSync(group, p.device)
    return module._grad_sync


def all_reduce_param_grads(params: Sequence[torch.Tensor], group=None) -> None:
    """One-shot helper: sum the gradients of `params` (weight, bias) over all ranks."""
    b = GradBucket(params)
    b.pack()
    b.all_reduce(group)
    b.unpack()
