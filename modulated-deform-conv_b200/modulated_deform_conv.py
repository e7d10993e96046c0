"""modulated_deform_conv -- drop-in for the reference's Python module of the same name
(reference: modulated_deform_conv.py; Functions :9-347, aliases :349-352, Modules :354-537,
Pack modules :730-839).

Same public names, constructor / forward signatures, parameter names and state-dict keys, so a
model written against the reference imports this file unchanged.  All compute goes through
``MDCONV_CUDA`` (this repo's sm_100a kernels behind the C ABI); CPU tensors raise
NotImplementedError exactly as in the reference (:22-23, :106-107).

Deliberate differences (reference defects):
  * all four ops work (only the modulated 2D pair is registered on the reference's HEAD);
  * the out-param forwards return the correct result for every ``in_step`` (B-1);
  * tensors are saved for backward whenever ANY differentiable input needs a gradient, including
    the bias-only case the reference forgets (B-7);
  * ``in_step`` is accepted and ignored: there is no column buffer to chunk.
"""
import math

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair, _triple

import MDCONV_CUDA

__all__ = [
    "DeformConv2dFunction", "ModulatedDeformConv2dFunction", "DeformConv3dFunction",
    "ModulatedDeformConv3dFunction", "deform_conv2d", "modulated_deform_conv2d", "deform_conv3d",
    "modulated_deform_conv3d", "DeformConv2d", "ModulatedDeformConv2d", "DeformConv3d",
    "ModulatedDeformConv3d", "DeformConv2dPack", "ModulatedDeformConv2dPack", "DeformConv3dPack",
    "ModulatedDeformConv3dPack",
]


def _setup(ctx, ntuple, input, bias, stride, padding, dilation, groups, deformable_groups, in_step):
    ctx.stride = ntuple(stride)
    ctx.padding = ntuple(padding)
    ctx.dilation = ntuple(dilation)
    ctx.groups = groups
    ctx.deformable_groups = deformable_groups
    ctx.in_step = in_step
    ctx.with_bias = bias is not None
    if not ctx.with_bias:
        bias = input.new_empty(0)  # the extension takes a 0-element tensor for "no bias"
    if not input.is_cuda:
        raise NotImplementedError
    return bias


def _needs_grad(*tensors):
    return any(t is not None and t.requires_grad for t in tensors)


# The Functions below go through MDCONV_CUDA.forward_ex / backward_ex -- the same C entry points the eight
# reference-named functions bind (MDCONV_CUDA.*_forward_cuda / *_backward_cuda stay available and equivalent) --
# because that form can hand the forward's internal layouts (channels-last input, re-ordered weight) to the
# backward, which the reference's positional signatures (:28-34, :57-64) have no slot for.  The hand-off buffer
# travels through ctx.save_for_backward like every other saved activation, so saved-tensor hooks (CPU offload,
# checkpointing) see it; MDCONV_CUDA.set_keep_layouts(False) / MDCONV_B200_KEEP_LAYOUTS=0 drops it (the backward then
# rebuilds the layouts: two more kernels, ~4 % of a cfg2 step, for one input-sized buffer less between the calls).
def _forward_common(ctx, ntuple, modulated, input, offset, mask, weight, bias, stride, padding, dilation, groups,
                    deformable_groups, in_step, grad_sync, prelaid=None):
    ctx.grad_sync = grad_sync
    ctx.modulated = modulated
    with_bias_tensor = bias
    bias = _setup(ctx, ntuple, input, bias, stride, padding, dilation, groups, deformable_groups, in_step)
    keep = _needs_grad(input, offset, mask, weight, with_bias_tensor)
    out = MDCONV_CUDA.forward_ex(input, weight, bias, offset, mask, None, weight.shape[2:], ctx.stride, ctx.padding,
                                 ctx.dilation, ctx.groups, ctx.deformable_groups, ctx.with_bias, keep=keep,
                                 prelaid=prelaid)
    if not keep:
        return out
    out, layouts = out
    ctx.layout_meta = None if layouts is None else layouts[1:]
    saved = [input, offset, weight, bias] + ([mask] if modulated else [])
    if layouts is not None:
        saved.append(layouts[0])
    ctx.save_for_backward(*saved)
    return out


def _backward_common(ctx, grad_output):
    """-> dict of gradients"""
    if not grad_output.is_cuda:
        raise NotImplementedError
    saved = list(ctx.saved_tensors)
    layouts = None
    if ctx.layout_meta is not None:
        layouts = (saved.pop(),) + tuple(ctx.layout_meta)
    input, offset, weight, bias = saved[:4]
    mask = saved[4] if ctx.modulated else None
    sync = getattr(ctx, "grad_sync", None)
    if sync is not None and not sync.active():
        sync = None
    g = MDCONV_CUDA.backward_ex(input, weight, bias, offset, mask, grad_output, weight.shape[2:], ctx.stride,
                                ctx.padding, ctx.dilation, ctx.groups, ctx.deformable_groups, ctx.with_bias,
                                saved=layouts, weight_ready=sync.event if sync is not None else None,
                                want_f32=sync is not None and sync.wants_f32(weight))
    if sync is not None:
        # data-parallel: ONE all-reduce of [grad_weight, grad_bias], overlapped with the data-gradient kernels
        sync.reduce(g, ctx.with_bias)
    return g


def _plain_grads(ctx, g):
    return (g["grad_input"], g["grad_offset"], g["grad_weight"], g["grad_bias"] if ctx.with_bias else None,
            None, None, None, None, None, None, None, None)


def _modulated_grads(ctx, g):
    return (g["grad_input"], g["grad_offset"], g["grad_mask"], g["grad_weight"],
            g["grad_bias"] if ctx.with_bias else None, None, None, None, None, None, None, None, None)


class DeformConv2dFunction(Function):
    @staticmethod
    def forward(ctx, input, offset, weight, bias=None, stride=1, padding=0, dilation=1,
                groups=1, deformable_groups=1, in_step=64, grad_sync=None, prelaid=None):
        return _forward_common(ctx, _pair, False, input, offset, None, weight, bias, stride, padding, dilation,
                               groups, deformable_groups, in_step, grad_sync, prelaid)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return _plain_grads(ctx, _backward_common(ctx, grad_output))


class ModulatedDeformConv2dFunction(Function):
    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1,
                groups=1, deformable_groups=1, in_step=64, grad_sync=None, prelaid=None):
        return _forward_common(ctx, _pair, True, input, offset, mask, weight, bias, stride, padding, dilation,
                               groups, deformable_groups, in_step, grad_sync, prelaid)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return _modulated_grads(ctx, _backward_common(ctx, grad_output))


class DeformConv3dFunction(Function):
    @staticmethod
    def forward(ctx, input, offset, weight, bias=None, stride=1, padding=0, dilation=1,
                groups=1, deformable_groups=1, in_step=64, grad_sync=None, prelaid=None):
        return _forward_common(ctx, _triple, False, input, offset, None, weight, bias, stride, padding, dilation,
                               groups, deformable_groups, in_step, grad_sync, prelaid)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return _plain_grads(ctx, _backward_common(ctx, grad_output))


class ModulatedDeformConv3dFunction(Function):
    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1,
                groups=1, deformable_groups=1, in_step=64, grad_sync=None, prelaid=None):
        return _forward_common(ctx, _triple, True, input, offset, mask, weight, bias, stride, padding, dilation,
                               groups, deformable_groups, in_step, grad_sync, prelaid)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return _modulated_grads(ctx, _backward_common(ctx, grad_output))


def _amp(*tensors):
    """Under ``torch.autocast('cuda')`` the op runs in the autocast dtype like the convolutions around it
    (the reference has no custom_fwd/custom_bwd, so mixed fp32 offsets / fp16 activations hit its
    same-dtype check).  The casts are ordinary differentiable ``.to()`` calls; outside autocast nothing changes."""
    if not torch.is_autocast_enabled("cuda"):
        return tensors
    dt = torch.get_autocast_dtype("cuda")
    return tuple(t if (t is None or not t.is_floating_point() or t.dtype == dt) else t.to(dt) for t in tensors)


# The reference binds these four names to ``Function.apply`` (modulated_deform_conv.py:349-352); same call
# signatures here, plus the autocast handling above.
def deform_conv2d(input, offset, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                  deformable_groups=1, in_step=64, grad_sync=None, prelaid=None):
    input, offset, weight, bias = _amp(input, offset, weight, bias)
    return DeformConv2dFunction.apply(input, offset, weight, bias, stride, padding, dilation, groups,
                                      deformable_groups, in_step, grad_sync, prelaid)


def modulated_deform_conv2d(input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                            deformable_groups=1, in_step=64, grad_sync=None, prelaid=None):
    input, offset, mask, weight, bias = _amp(input, offset, mask, weight, bias)
    return ModulatedDeformConv2dFunction.apply(input, offset, mask, weight, bias, stride, padding, dilation, groups,
                                               deformable_groups, in_step, grad_sync, prelaid)


def deform_conv3d(input, offset, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                  deformable_groups=1, in_step=64, grad_sync=None, prelaid=None):
    input, offset, weight, bias = _amp(input, offset, weight, bias)
    return DeformConv3dFunction.apply(input, offset, weight, bias, stride, padding, dilation, groups,
                                      deformable_groups, in_step, grad_sync, prelaid)


def modulated_deform_conv3d(input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                            deformable_groups=1, in_step=64, grad_sync=None, prelaid=None):
    input, offset, mask, weight, bias = _amp(input, offset, mask, weight, bias)
    return ModulatedDeformConv3dFunction.apply(input, offset, mask, weight, bias, stride, padding, dilation, groups,
                                               deformable_groups, in_step, grad_sync, prelaid)


class _DeformConvNd(nn.Module):
    """Shared body of the four Modules (reference: modulated_deform_conv.py:354-537).  Parameters:
    ``weight`` [out, in/groups, *kernel] and optional ``bias`` [out]; init U(-1/sqrt(in*prod(k)), +)
    with zero bias (:385-392)."""
    _ntuple = staticmethod(_pair)

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, deformable_groups=1, bias=False, in_step=64):
        super().__init__()
        assert in_channels % groups == 0, \
            'in_channels {} cannot be divisible by groups {}'.format(in_channels, groups)
        assert out_channels % groups == 0, \
            'out_channels {} cannot be divisible by groups {}'.format(out_channels, groups)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = self._ntuple(kernel_size)
        self.stride = self._ntuple(stride)
        self.padding = self._ntuple(padding)
        self.dilation = self._ntuple(dilation)
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.in_step = in_step
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // self.groups, *self.kernel_size))
        self.with_bias = bias
        if self.with_bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.bias = None
        self.reset_parameters()

    def _fan(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        return n

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self._fan())
        self.weight.data.uniform_(-stdv, stdv)
        if self.with_bias:
            self.bias.data.fill_(0)

    def _conv_args(self):
        # the trailing element is the data-parallel gradient reducer (mdconv_dist.attach), None by default
        return (self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups,
                self.deformable_groups, self.in_step, getattr(self, "_grad_sync", None))

    def extra_repr(self):
        return ('{in_channels}, {out_channels}, kernel_size={kernel_size}, stride={stride}, padding={padding}, '
                'dilation={dilation}, groups={groups}, deformable_groups={deformable_groups}, '
                'bias={with_bias}').format(**self.__dict__)


class DeformConv2d(_DeformConvNd):
    def forward(self, x, offset):
        return deform_conv2d(x, offset, *self._conv_args())


class ModulatedDeformConv2d(_DeformConvNd):
    def forward(self, x, offset, mask):
        return modulated_deform_conv2d(x, offset, mask, *self._conv_args())


class DeformConv3d(_DeformConvNd):
    _ntuple = staticmethod(_triple)

    def forward(self, x, offset):
        return deform_conv3d(x, offset, *self._conv_args())


class ModulatedDeformConv3d(_DeformConvNd):
    _ntuple = staticmethod(_triple)

    def forward(self, x, offset, mask):
        return modulated_deform_conv3d(x, offset, mask, *self._conv_args())


def _make_producer(mod, channels_per_tap):
    """The offset / mask producing convolution of the Pack modules: same kernel, stride and padding
    as the deformable conv, no dilation, with bias (reference: modulated_deform_conv.py:734-738,
    :759-768, :791-794, :815-822); weights U(-stdv, stdv), zero bias (:741-748)."""
    taps = 1
    for k in mod.kernel_size:
        taps *= k
    conv_cls = nn.Conv2d if len(mod.kernel_size) == 2 else nn.Conv3d
    conv = conv_cls(mod.in_channels, mod.deformable_groups * channels_per_tap * taps,
                    kernel_size=mod.kernel_size, stride=mod.stride, padding=mod.padding, bias=True)
    stdv = 1. / math.sqrt(mod._fan())
    conv.weight.data.uniform_(-stdv, stdv)
    conv.bias.data.zero_()
    return conv


def _plain_producer(conv):
    """True when ``conv`` is an unmodified producer as ``_make_producer`` builds it: a stock nn.Conv2d / nn.Conv3d with
    default dilation / groups / zero padding mode and no hooks.  Only then may its forward() be bypassed."""
    if type(conv) not in (nn.Conv2d, nn.Conv3d):
        return False
    if any(d != 1 for d in conv.dilation) or conv.groups != 1 or conv.padding_mode != "zeros" or conv.bias is None:
        return False
    if isinstance(conv.padding, str):
        return False
    hooks = (conv._forward_hooks, conv._forward_pre_hooks, conv._backward_hooks, conv._backward_pre_hooks,
             getattr(conv, "_forward_hooks_with_kwargs", {}), getattr(conv, "_forward_pre_hooks_with_kwargs", {}))
    if any(len(h) for h in hooks):
        return False
    # torch.nn.utils.prune / parametrizations re-compute .weight in a forward pre-hook or a property: both show up
    # as a missing plain Parameter
    return isinstance(conv._parameters.get("weight"), nn.Parameter) and isinstance(conv._parameters.get("bias"), nn.Parameter)


def _offset_and_mask(mod, x):
    """conv_offset(x), conv_mask(x) of a modulated Pack (reference: modulated_deform_conv.py:781-785, :835-839).
    When both producers are plain convolutions with identical geometry they run as ONE convolution launch over the
    concatenated filters (x is read once); anything else -- hooks, pruning, a replaced / wrapped producer, different
    stride or padding -- goes through the submodules' own forward() exactly as the reference does."""
    co, cm = mod.conv_offset, mod.conv_mask
    fused = (_plain_producer(co) and _plain_producer(cm) and type(co) is type(cm) and co.stride == cm.stride
             and co.padding == cm.padding and co.kernel_size == cm.kernel_size and co.in_channels == cm.in_channels
             and not torch.jit.is_scripting() and getattr(mod, "fuse_producers", True))
    if not fused:
        return co(x), cm(x)
    f = torch.nn.functional.conv2d if x.dim() == 4 else torch.nn.functional.conv3d
    both = f(x, torch.cat([co.weight, cm.weight], 0), torch.cat([co.bias, cm.bias], 0), co.stride, co.padding)
    n = co.out_channels
    return both[:, :n].contiguous(), both[:, n:].contiguous()


class _ProducerConv(Function):
    """The offset (+ mask) producer of a Pack module as ONE convolution over the concatenated filter banks, run on
    the deformable forward kernel in regular-grid mode (MDCONV_CUDA.producer_conv) from the SAME channels-last copy
    of x the deformable convolution then gathers from (``prelaid``).  Backward: the ordinary convolution gradients
    (aten::convolution_backward), split back onto the two banks."""

    @staticmethod
    def forward(ctx, x, w_off, b_off, w_mask, b_mask, stride, padding, prelaid):
        has_mask = w_mask is not None
        w = torch.cat([w_off, w_mask], 0) if has_mask else w_off
        b = torch.cat([b_off, b_mask], 0) if has_mask else b_off
        ctx.stride, ctx.padding, ctx.n_off, ctx.has_mask = tuple(stride), tuple(padding), w_off.shape[0], has_mask
        ctx.save_for_backward(x, w)
        off, msk = MDCONV_CUDA.producer_conv(x, w, b, w_off.shape[0], tuple(w.shape[2:]), ctx.stride, ctx.padding,
                                             prelaid=prelaid)
        if has_mask:
            return off, msk
        return off

    @staticmethod
    @once_differentiable
    def backward(ctx, g_off, g_mask=None):
        x, w = ctx.saved_tensors
        g = torch.cat([g_off, g_mask], 1) if ctx.has_mask else g_off
        nd = x.dim() - 2
        gx, gw, gb = torch.ops.aten.convolution_backward(
            g.contiguous(), x, w, [w.shape[0]], list(ctx.stride), list(ctx.padding), [1] * nd, False, [0] * nd, 1,
            [ctx.needs_input_grad[0], True, True])
        n = ctx.n_off
        if ctx.has_mask:
            return gx, gw[:n], gb[:n], gw[n:], gb[n:], None, None, None
        return gx, gw, gb, None, None, None, None, None


def _fused_pack_ok(mod, x, producers):
    """The fused route needs the tcgen05 kernels for BOTH the deformable convolution and the producer (CUDA, fp16 /
    bf16, NCHW-contiguous x, channels multiples of 64 ...), plain producers with identical geometry, and no autocast
    re-typing in between."""
    if not getattr(mod, "fuse_producers", True) or torch.is_autocast_enabled("cuda") or torch.jit.is_scripting():
        return False
    if not (x.is_cuda and x.is_contiguous() and x.dtype in (torch.float16, torch.bfloat16)):
        return False
    if any(not _plain_producer(c) or c.weight.dtype != x.dtype for c in producers):
        return False
    c0 = producers[0]
    if any(c.stride != c0.stride or c.padding != c0.padding or c.kernel_size != c0.kernel_size
           or c.in_channels != c0.in_channels for c in producers[1:]):
        return False
    if mod.weight.dtype != x.dtype:
        return False
    n_out = sum(c.out_channels for c in producers)
    return MDCONV_CUDA.producer_conv_supported(x, n_out, c0.kernel_size, c0.stride, c0.padding)


def _pack_forward(mod, x, modulated):
    """Forward of the four Pack modules (reference: modulated_deform_conv.py:750-753, :781-785, :806-809, :835-839):
    offset (and mask) = conv_offset(x) (, conv_mask(x)), then the deformable convolution.

    Fused route: x is laid out channels-last ONCE; the producers run as one regular-grid pass of the deformable
    forward kernel over that copy (27 / 18 / 81 ... output channels, padded to a multiple of 32) and the deformable
    convolution gathers from the same copy.  Anything the fused route does not cover -- hooks, pruning, a replaced
    producer, different stride / padding between the two, fp32, CPU -- goes through the submodules' own forward()
    exactly as the reference does."""
    producers = [mod.conv_offset] + ([mod.conv_mask] if modulated else [])
    fn = {(2, False): deform_conv2d, (2, True): modulated_deform_conv2d,
          (3, False): deform_conv3d, (3, True): modulated_deform_conv3d}[(len(mod.kernel_size), modulated)]
    prelaid = None
    if _fused_pack_ok(mod, x, producers):
        prelaid = MDCONV_CUDA.to_channels_last(x, mod.weight, mod.kernel_size, mod.stride, mod.padding, mod.dilation,
                                               mod.groups, mod.deformable_groups, modulated, mod.bias is not None)
    if prelaid is not None:
        co = mod.conv_offset
        cm = mod.conv_mask if modulated else None
        out = _ProducerConv.apply(x, co.weight, co.bias, cm.weight if cm is not None else None,
                                  cm.bias if cm is not None else None, co.stride, co.padding, prelaid)
        offset, mask = out if modulated else (out, None)
    elif modulated:
        offset, mask = _offset_and_mask(mod, x)
    else:
        offset, mask = mod.conv_offset(x), None
    args = (x, offset, mask) if modulated else (x, offset)
    return fn(*args, *mod._conv_args(), prelaid=prelaid)


class DeformConv2dPack(DeformConv2d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset = _make_producer(self, 2)

    def forward(self, x):
        return _pack_forward(self, x, False)


class ModulatedDeformConv2dPack(ModulatedDeformConv2d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset = _make_producer(self, 2)
        self.conv_mask = _make_producer(self, 1)

    def forward(self, x):
        # as in the reference the mask is the raw conv output (no sigmoid), :781-785
        return _pack_forward(self, x, True)


class DeformConv3dPack(DeformConv3d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset = _make_producer(self, 3)

    def forward(self, x):
        return _pack_forward(self, x, False)


class ModulatedDeformConv3dPack(ModulatedDeformConv3d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset = _make_producer(self, 3)
        self.conv_mask = _make_producer(self, 1)

    def forward(self, x):
        return _pack_forward(self, x, True)
