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
    (``mdconv_backward_ex(..., weight_ready_event)``).  ``reduce`` then runs ONE all-reduce(sum) on a communication
    stream that waits for that event only, so the collective overlaps the grad_offset / grad_mask / grad_input kernels;
    the compute stream re-joins before the backward returns, which keeps autograd's accumulation into ``.grad``
    ordered after the reduction.

    What is reduced: for fp16 / bf16 parameters the library also hands out the two gradients in fp32 BEFORE rounding
    (one flat buffer [grad_weight || grad_bias], ``mdconv_bwd_extra_t.grad_weight_f32``); the ranks exchange that and
    round once afterwards, so G GPUs give what one GPU gives on the concatenated batch instead of a sum of G rounded
    partial sums.  fp32 / fp64 parameters are reduced in place in their own dtype.
    ``average=True`` divides by the world size (DistributedDataParallel's convention); the default is the op's own
    single-GPU semantics, a sum.  On CPU tensors / a non-initialised process group it degrades to a plain (or no)
    all-reduce.
    """

    def __init__(self, group=None, device: Optional[torch.device] = None, average: bool = False):
        self.group = group
        self.average = average
        self._stream = None
        self.event = None
        self.trace = None            # optional list of (start, stop) CUDA events around each collective (bench.py)
        self.keep_last = False       # keep a reference to the last reduced fp32 buffer in .last_flat32 (checks)
        self.last_flat32 = None
        if device is not None and device.type == "cuda":
            self._stream = torch.cuda.Stream(device=device)
            self.event = torch.cuda.Event()
            self.event.record(torch.cuda.current_stream(device))   # force creation of the CUDA event handle

    def active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def wants_f32(self, weight: torch.Tensor) -> bool:
        return weight.is_cuda and weight.dtype in (torch.float16, torch.bfloat16)

    def _scale(self, t: torch.Tensor) -> None:
        if self.average:
            t.div_(dist.get_world_size(self.group))

    def reduce(self, grads, with_bias: bool = True) -> None:
        """``grads``: the dict MDCONV_CUDA.backward_ex returned (grad_weight, grad_bias, optionally flat_f32), or a
        plain sequence of tensors to reduce in place."""
        if not self.active():
            return
        if not isinstance(grads, dict):
            ts = [t for t in grads if t is not None and t.numel() > 0]
            grads = {"_list": ts}
            flat32, outs = None, ts
        else:
            flat32 = grads.get("flat_f32")
            outs = [grads["grad_weight"]] + ([grads["grad_bias"]] if with_bias and grads.get("grad_bias") is not None
                                              and grads["grad_bias"].numel() else [])
        if not outs:
            return
        if self._stream is None:
            for t in outs:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                self._scale(t)
            return
        # Host cost matters here (this runs inside every backward, once per rank, and eight launch loops share one
        # host): set_stream instead of the torch.cuda.stream context manager, one rounding copy when the gradients
        # are views of one buffer, and no Tensor.record_stream -- the main stream re-joins the communication stream
        # below, before this call returns, so whatever frees or re-uses these tensors afterwards is already ordered
        # behind the collective and the copies.
        main = torch.cuda.current_stream(outs[0].device)
        self._stream.wait_event(self.event)          # only grad_weight / grad_bias, not the whole backward
        torch.cuda.set_stream(self._stream)
        try:
            if self.trace is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(self._stream)
            if flat32 is not None:
                dist.all_reduce(flat32, op=dist.ReduceOp.SUM, group=self.group)
                self._scale(flat32)
                both = grads.get("flat_storage")
                if both is not None and both.numel() == flat32.numel():
                    both.copy_(flat32)               # the single rounding to the storage dtype, both gradients at once
                else:
                    off = 0
                    for t in outs:
                        t.copy_(flat32[off:off + t.numel()].view(t.shape))
                        off += t.numel()
                if self.keep_last:
                    self.last_flat32 = flat32
            else:
                flat = outs[0] if len(outs) == 1 else torch.cat([t.reshape(-1) for t in outs])
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                self._scale(flat)
                if len(outs) > 1:
                    off = 0
                    for t in outs:
                        t.copy_(flat[off:off + t.numel()].reshape(t.shape))
                        off += t.numel()
            if self.trace is not None:
                e1.record(self._stream)
                self.trace.append((e0, e1))
        finally:
            torch.cuda.set_stream(main)
        main.wait_stream(self._stream)


_comm_groups = {}


def comm_group(max_ctas: int = 4, cga_cluster_size: int = 1):
    """A dedicated NCCL communicator for the parameter-gradient all-reduce, limited to ``max_ctas`` CTAs that are
    launched WITHOUT thread-block clusters (``cga_cluster_size`` 1; NCCL's default on sm_90+ is clusters of 4).
    The collective runs BESIDE the persistent data-gradient kernels, which own every SM they are given, so it can
    only use the SMs those kernels leave free (mdconv_set_comm_sms): a communicator that wants 16+ CTAs leaves most
    of its channels waiting, and a cluster of 4 CTAs needs 4 free SMs of ONE GPC at the same moment, which the
    scattered spare SMs never offer before the neighbouring kernel retires.  Collective call (every rank, same
    order); cached per setting.  Returns None when the default backend is not NCCL."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend() != "nccl":
        return None
    key = (int(max_ctas), int(cga_cluster_size))
    if key not in _comm_groups:
        opts = dist.ProcessGroupNCCL.Options()
        try:
            opts.config.max_ctas = key[0]
            opts.config.min_ctas = 1
            if key[1] >= 0:
                opts.config.cga_cluster_size = key[1]
        except Exception:
            pass
        _comm_groups[key] = dist.new_group(backend="nccl", pg_options=opts)
    return _comm_groups[key]


def _is_deform_module(m) -> bool:
    return hasattr(m, "deformable_groups") and hasattr(m, "_conv_args")


def attach(module, group=None, average: bool = False, max_ctas: Optional[int] = None,
           cga_cluster_size: int = 1) -> "GradSync":
    """Make ``module`` (a DeformConv*/ModulatedDeformConv*[Pack] module of this package) all-reduce its weight / bias
    gradients inside its own backward, overlapped with the data gradients.  Every OTHER parameter of the module --
    the conv_offset / conv_mask producers of a Pack -- gets a post-accumulate-grad hook that all-reduces its
    gradient the same way (sum, or mean with ``average``), so a data-parallel user of Pack modules has no silently
    diverging replicas.  ``max_ctas``: run the collectives on a dedicated small NCCL communicator (``comm_group``)
    and leave exactly that many SMs free beside the data-gradient kernels.  Returns the GradSync."""
    p = next(module.parameters())
    if max_ctas is not None and group is None and p.is_cuda:
        group = comm_group(max_ctas, cga_cluster_size)
        import MDCONV_CUDA
        MDCONV_CUDA.set_comm_sms(max_ctas)
    sync = GradSync(group, p.device, average=average)
    module._grad_sync = sync
    own = {id(module.weight)} | ({id(module.bias)} if getattr(module, "bias", None) is not None else set())

    def _reduce_other(param):
        if sync.active() and param.grad is not None:
            dist.all_reduce(param.grad, op=dist.ReduceOp.SUM, group=sync.group)
            sync._scale(param.grad)

    handles = []
    for q in module.parameters():
        if id(q) not in own and q.requires_grad:
            handles.append(q.register_post_accumulate_grad_hook(_reduce_other))
    module._grad_sync_handles = handles
    return sync


def detach(module) -> None:
    """Undo ``attach``: the module returns LOCAL gradients again (what DistributedDataParallel expects to reduce)."""
    for h in getattr(module, "_grad_sync_handles", []):
        h.remove()
    module._grad_sync_handles = []
    module._grad_sync = None


def prepare_ddp(model, group=None, max_ctas: Optional[int] = None):
    """Use the in-backward overlapped reduction together with ``torch.nn.parallel.DistributedDataParallel``.

    Two modes exist for the parameter gradients of the deformable convolutions under DDP:
      * local-grad mode (default, nothing to call): the op returns local gradients and DDP's bucketed all-reduce
        averages them like every other parameter's;
      * this function: every deformable-conv module in ``model`` is attached with ``average=True`` (DDP's
        convention) and its parameters are put on DDP's ignore list, so DDP neither re-reduces the already reduced
        gradients nor waits for them -- the only cross-sample reduction of the op (reference
        src/mdeformable_conv.cu:436-444) then overlaps the op's own data-gradient kernels.
    Call it BEFORE wrapping ``model`` in DistributedDataParallel.  Returns the list of ignored parameter names."""
    ignored = []
    for name, m in model.named_modules():
        if _is_deform_module(m):
            attach(m, group, average=True, max_ctas=max_ctas)
            ignored += [f"{name}.{n}" if name else n for n, _ in m.named_parameters()]
    from torch.nn.parallel import DistributedDataParallel as DDP
    prev = list(getattr(model, "_ddp_params_and_buffers_to_ignore", []))
    DDP._set_params_and_buffers_to_ignore_for_model(model, prev + ignored)
    return ignored


def all_reduce_param_grads(params: Sequence[torch.Tensor], group=None) -> None:
    """One-shot helper: sum the gradients of `params` (weight, bias) over all ranks."""
    b = GradBucket(params)
    b.pack()
    b.all_reduce(group)
    b.unpack()
