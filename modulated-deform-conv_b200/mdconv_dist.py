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


def all_reduce_param_grads(params: Sequence[torch.Tensor], group=None) -> None:
    """One-shot helper: sum the gradients of `params` (weight, bias) over all ranks."""
    b = GradBucket(params)
    b.pack()
    b.all_reduce(group)
    b.unpack()
