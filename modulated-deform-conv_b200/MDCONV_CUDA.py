"""MDCONV_CUDA -- drop-in for the reference's pybind11 extension module of the same name.

The reference builds ``MDCONV_CUDA`` from src/*.cu with torch's CUDAExtension (setup.py:36-40) and
registers ``*_forward_cuda`` / ``*_backward_cuda`` (src/mdeformable_conv.cu:460-465; the other six
registrations are commented out in src/warp.cpp:1-19 on HEAD).  This module exposes all eight names
with the reference's positional signatures, but is a thin ctypes binding over the C ABI of
``lib/libmdconv_b200.so`` (include/mdconv_b200.h): torch supplies device memory and the current
stream, every kernel is this repo's sm_100a code.

There is NO fallback: if the shared library is missing the import fails, and CPU tensors are
rejected -- the product path never computes on the host.

Error behaviour mirrors the reference: non-contiguous tensors and kernel/channel mismatches raise
RuntimeError (TORCH_CHECK / AT_ERROR in src/mdeformable_conv.cu:127-148); additionally dtype/device
mismatches and CUDA launch failures raise RuntimeError instead of being printed and ignored
(src/mdeformable_conv.cu:113-117).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("MDCONV_B200_LIB", os.path.join(_HERE, "lib", "libmdconv_b200.so"))
if not os.path.exists(_LIB_PATH):
    raise ImportError(
        f"MDCONV_CUDA: {_LIB_PATH} not found -- build it with `python __graft_entry__.py build` "
        "(or `make -C modulated-deform-conv_b200/csrc`).  There is no CPU/eager fallback.")
_lib = ctypes.CDLL(_LIB_PATH)

_vp, _i, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t


class _Geom(ctypes.Structure):
    """struct mdconv_geom (include/mdconv_b200.h)."""
    _fields_ = [("nd", ctypes.c_int32), ("batch", ctypes.c_int32), ("c_in", ctypes.c_int32),
                ("c_out", ctypes.c_int32), ("in_size", ctypes.c_int32 * 3), ("kernel", ctypes.c_int32 * 3),
                ("stride", ctypes.c_int32 * 3), ("pad", ctypes.c_int32 * 3), ("dilation", ctypes.c_int32 * 3),
                ("groups", ctypes.c_int32), ("deformable_groups", ctypes.c_int32), ("dtype", ctypes.c_int32),
                ("modulated", ctypes.c_int32), ("grad_rule", ctypes.c_int32), ("with_bias", ctypes.c_int32),
                ("path", ctypes.c_int32), ("io_layout", ctypes.c_int32)]


class _FwdExtra(ctypes.Structure):
    """struct mdconv_fwd_extra (include/mdconv_b200.h)."""
    _fields_ = [("input_channels_last", ctypes.c_void_p), ("regular_grid", ctypes.c_int32),
                ("c_out_valid", ctypes.c_int32), ("output2", ctypes.c_void_p), ("split", ctypes.c_int32)]


class _BwdExtra(ctypes.Structure):
    """struct mdconv_bwd_extra (include/mdconv_b200.h)."""
    _fields_ = [("saved", ctypes.c_void_p), ("saved_bytes", ctypes.c_size_t), ("weight_ready_event", ctypes.c_void_p),
                ("grad_weight_f32", ctypes.c_void_p), ("grad_bias_f32", ctypes.c_void_p)]


_lib.mdconv_abi_version.restype = _i
_lib.mdconv_status_string.restype = ctypes.c_char_p
_lib.mdconv_status_string.argtypes = [_i]
_lib.mdconv_last_cuda_error.restype = ctypes.c_char_p
_lib.mdconv_last_launch_count.restype = _i
_lib.mdconv_workspace_bytes.restype = _sz
_lib.mdconv_workspace_bytes.argtypes = [ctypes.POINTER(_Geom), _i]
_lib.mdconv_selected_path.restype = _i
_lib.mdconv_selected_path.argtypes = [ctypes.POINTER(_Geom), _i]
_lib.mdconv_out_size.restype = _i
_lib.mdconv_out_size.argtypes = [ctypes.POINTER(_Geom), ctypes.POINTER(ctypes.c_int32)]
_lib.mdconv_forward.restype = _i
_lib.mdconv_forward.argtypes = [ctypes.POINTER(_Geom)] + [_vp] * 7 + [_sz, _vp]
_lib.mdconv_backward.restype = _i
_lib.mdconv_backward.argtypes = [ctypes.POINTER(_Geom)] + [_vp] * 11 + [_sz, _vp]
_lib.mdconv_saved_bytes.restype = _sz
_lib.mdconv_saved_bytes.argtypes = [ctypes.POINTER(_Geom)]
_lib.mdconv_backward_saved.restype = _i
_lib.mdconv_backward_saved.argtypes = [ctypes.POINTER(_Geom)] + [_vp] * 11 + [_sz, _vp, _sz, _vp, _vp]
_lib.mdconv_forward_ex.restype = _i
_lib.mdconv_forward_ex.argtypes = [ctypes.POINTER(_Geom)] + [_vp] * 7 + [_sz, ctypes.POINTER(_FwdExtra), _vp]
_lib.mdconv_channels_last_bytes.restype = _sz
_lib.mdconv_channels_last_bytes.argtypes = [ctypes.POINTER(_Geom)]
_lib.mdconv_to_channels_last.restype = _i
_lib.mdconv_to_channels_last.argtypes = [ctypes.POINTER(_Geom), _vp, _vp, _vp]
_lib.mdconv_backward_ex.restype = _i
_lib.mdconv_backward_ex.argtypes = [ctypes.POINTER(_Geom)] + [_vp] * 11 + [_sz, ctypes.POINTER(_BwdExtra), _vp]
_lib.mdconv_set_comm_sms.restype = _i
_lib.mdconv_set_comm_sms.argtypes = [_i]

_TAIL = [_i, _vp, _sz, _vp]  # dtype, workspace, workspace_bytes, stream
_G2 = [_i] * 17              # batch, channels, H, W, Cout, kh,kw, sh,sw, ph,pw, dh,dw, g, dg, in_step, with_bias
_G3 = [_i] * 22              # batch, channels, H, W, L, Cout, 12 geometry ints, g, dg, in_step, with_bias
for _name, _nptr, _geo in (("deform_conv2d_forward_cuda", 5, _G2), ("deform_conv2d_backward_cuda", 9, _G2),
                           ("modulated_deform_conv2d_forward_cuda", 6, _G2),
                           ("modulated_deform_conv2d_backward_cuda", 11, _G2),
                           ("deform_conv3d_forward_cuda", 5, _G3), ("deform_conv3d_backward_cuda", 9, _G3),
                           ("modulated_deform_conv3d_forward_cuda", 6, _G3),
                           ("modulated_deform_conv3d_backward_cuda", 11, _G3)):
    _f = getattr(_lib, _name)
    _f.restype = _i
    _f.argtypes = [_vp] * _nptr + _geo + _TAIL
_lib.mdconv_legacy_workspace_bytes.restype = _sz
_lib.mdconv_legacy_workspace_bytes.argtypes = [_i] * 25

if _lib.mdconv_abi_version() != 3:
    raise ImportError("MDCONV_CUDA: libmdconv_b200.so ABI version mismatch")

_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.float64: 3}

# Kernel-family override for tests / benchmarks: 0 auto, 1 generic SIMT, 2 tcgen05 (error if n/a).
_path_override = int(os.environ.get("MDCONV_B200_PATH", "0"))


def set_path(path: int) -> None:
    """0 = auto (default), 1 = force the generic SIMT kernels, 2 = force the tcgen05 path."""
    global _path_override
    _path_override = int(path)


def last_launch_count() -> int:
    """Kernel launches enqueued by the most recent forward/backward call on this thread."""
    return int(_lib.mdconv_last_launch_count())


_lib.mdconv_total_launch_count.restype = ctypes.c_longlong


def total_launch_count() -> int:
    """Kernel launches enqueued by this library in this process so far (forward and backward)."""
    return int(_lib.mdconv_total_launch_count())


def library_path() -> str:
    return _LIB_PATH


_lib.mdconv_set_fp32_tensor_cores.restype = _i
_lib.mdconv_set_fp32_tensor_cores.argtypes = [_i]


def set_fp32_tensor_cores(on: bool) -> bool:
    """fp32 calls on the split-bf16 tensor-core path (default) or on the exact-fp32 SIMT kernels.  Returns the
    previous setting."""
    return bool(_lib.mdconv_set_fp32_tensor_cores(1 if on else 0))


_lib.mdconv_set_gcol_budget_mb.restype = ctypes.c_longlong
_lib.mdconv_set_gcol_budget_mb.argtypes = [ctypes.c_longlong]


def set_gcol_budget_mb(megabytes: int) -> int:
    """Cap (MiB) of the grad_columns buffer behind the col2im-gather grad_input of 16-bit calls (default 1024; larger
    batches are walked in chunks; 0: always the third contraction).  Returns the previous cap (-1 if the default had
    not been resolved yet)."""
    return int(_lib.mdconv_set_gcol_budget_mb(int(megabytes)))


def set_comm_sms(sms: int) -> int:
    """SMs the persistent data-gradient kernels leave free for a concurrent collective once the parameter
    gradients are final (see mdconv_set_comm_sms).  Returns the previous value."""
    return int(_lib.mdconv_set_comm_sms(int(sms)))


_lib.mdconv_profile_enable.restype = _i
_lib.mdconv_profile_enable.argtypes = [_i]
_lib.mdconv_side_stream_enable.restype = _i
_lib.mdconv_side_stream_enable.argtypes = [_i]
_lib.mdconv_profile_read.restype = _i
_lib.mdconv_profile_read.argtypes = [ctypes.c_char_p, _sz]


def profile_enable(on: bool, keep_side_stream: bool = False) -> None:
    """Bracket every kernel this library launches with CUDA events (benchmark use).  While profiling,
    the backward's side stream is switched off so each kernel's time is its own (``keep_side_stream`` leaves it
    on: for timelines, where the overlap is the point)."""
    _lib.mdconv_profile_enable(1 if on else 0)
    _lib.mdconv_side_stream_enable(1 if (not on or keep_side_stream) else 0)


def profile_read(with_start: bool = False):
    """-> list of (kernel_name, milliseconds) in launch order since the last read (synchronises); with
    ``with_start`` each entry is (name, milliseconds, start_ms relative to the first launch of the batch)."""
    buf = ctypes.create_string_buffer(1 << 18)
    _lib.mdconv_profile_read(buf, len(buf))
    out = []
    for line in buf.value.decode().splitlines():
        name, _, rest = line.partition("=")
        ms, _, at = rest.partition("@")
        out.append((name, float(ms), float(at or 0.0)) if with_start else (name, float(ms)))
    return out


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = _lib.mdconv_status_string(rc).decode()
        if rc == 5:
            msg += ": " + _lib.mdconv_last_cuda_error().decode()
        raise RuntimeError(f"{what}: {msg}")


def _raw_stream() -> int:
    """cudaStream_t of the current device's current stream (the cheap form of torch.cuda.current_stream().cuda_stream:
    this sits on the launch path of every call)."""
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


class _NoCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NO_CTX = _NoCtx()


def _on_device(device):
    """Context that makes ``device`` current -- nothing at all when it already is (the usual case: a device-guard
    context manager costs ~10 us of host time per call)."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_CTX
    return torch.cuda.device(device)


def _ptr(t):
    return None if t is None or t.numel() == 0 else ctypes.c_void_p(t.data_ptr())


_ACTIVATIONS = ("input", "output", "grad_output", "grad_input")


def _cl_format(nd):
    return torch.channels_last if nd == 2 else torch.channels_last_3d


def _is_channels_last(t) -> bool:
    """Dense channels-last (NHWC / NDHWC strides) and NOT also NCHW-contiguous (size-1 dims make both true)."""
    if t is None or t.dim() not in (4, 5) or t.is_contiguous():
        return False
    return t.is_contiguous(memory_format=_cl_format(t.dim() - 2))


def _validate(named, weight, kernel, group, nd, cl=False):
    """The reference's entry checks (src/mdeformable_conv.cu:127-148 and twins).  With ``cl`` the activation
    tensors (input / output / grad_output / grad_input) must be dense channels-last instead of NCHW-contiguous."""
    first = named[0][1]
    for name, t in named:
        if cl and name in _ACTIVATIONS:
            if not t.is_contiguous(memory_format=_cl_format(nd)):
                raise RuntimeError(f"{name} tensor has to be channels-last like the input")
        elif not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
        if not t.is_cuda:
            raise RuntimeError(f"{name} tensor has to be a CUDA tensor (no CPU path)")
        if t.numel() and (t.dtype != first.dtype or t.device != first.device):
            raise RuntimeError(f"{name} tensor dtype/device differs from input ({t.dtype}, {t.device})")
    if first.dtype not in _DTYPES:
        raise RuntimeError(f"unsupported dtype {first.dtype}")
    if first.dim() != nd + 2 or weight.dim() != nd + 2:
        raise RuntimeError(f"expected {nd + 2}-d input and weight")
    if tuple(weight.shape[2:]) != tuple(kernel):
        raise RuntimeError("Input shape and kernel shape wont match: ({} vs {}).".format(
            " x ".join(map(str, kernel)), " x ".join(map(str, weight.shape[2:]))))
    if first.shape[1] != weight.shape[1] * group:
        raise RuntimeError("Input shape and kernel channels wont match: ({} vs {}).".format(
            first.shape[1], weight.shape[1] * group))


def _geom(input, weight, kernel, stride, pad, dil, group, dgroup, modulated, with_bias, cl=False):
    nd = input.dim() - 2
    g = _Geom()
    g.nd, g.batch, g.c_in, g.c_out = nd, input.shape[0], input.shape[1], weight.shape[0]
    for a in range(3):
        g.in_size[a] = input.shape[2 + a] if a < nd else 1
        g.kernel[a] = kernel[a] if a < nd else 1
        g.stride[a] = stride[a] if a < nd else 1
        g.pad[a] = pad[a] if a < nd else 0
        g.dilation[a] = dil[a] if a < nd else 1
    g.groups, g.deformable_groups = group, dgroup
    g.dtype = _DTYPES[input.dtype]
    g.modulated = 1 if modulated else 0
    g.grad_rule = 0 if (nd == 2 and modulated) else 1
    g.with_bias = 1 if with_bias else 0
    g.path = _path_override
    g.io_layout = 1 if cl else 0
    return g


def _channels_last_ok(input, weight, kernel, stride, pad, dil, group, dgroup, modulated, with_bias, backward):
    """Does the library take this call with channels-last activations?  (tcgen05 kernels only: fp16 / bf16,
    channels per group multiples of 64, ...)"""
    if input.dtype not in (torch.float16, torch.bfloat16):
        return False
    g = _geom(input, weight, kernel, stride, pad, dil, group, dgroup, modulated, with_bias, cl=True)
    return int(_lib.mdconv_selected_path(ctypes.byref(g), 1 if backward else 0)) == 2


def _out_shape(g, input, weight):
    out = (ctypes.c_int32 * 3)()
    _check(_lib.mdconv_out_size(ctypes.byref(g), out), "output size")
    return (input.shape[0], weight.shape[0]) + tuple(out[a] for a in range(g.nd))


def _expect(t, shape, name):
    if tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")


def _workspace(g, backward, device):
    n = int(_lib.mdconv_workspace_bytes(ctypes.byref(g), 1 if backward else 0))
    if n == 0:
        return None, 0
    return torch.empty(n, dtype=torch.uint8, device=device), n


# Forward -> backward hand-off of the internal layouts (mdconv_saved_bytes / mdconv_backward_saved): the autograd
# Functions of modulated_deform_conv.py keep the forward's workspace alive in their ctx and give it back to the
# backward, which then skips the channels-last / weight re-order passes.  Costs the workspace's memory between
# the two calls; MDCONV_B200_KEEP_LAYOUTS=0 (or set_keep_layouts(False)) turns it off.
_keep_layouts = os.environ.get("MDCONV_B200_KEEP_LAYOUTS", "1") != "0"


def set_keep_layouts(on: bool) -> None:
    global _keep_layouts
    _keep_layouts = bool(on)


def _forward(input, weight, bias, offset, mask, output, kernel, stride, pad, dil, group, dgroup, with_bias,
             keep=False, prelaid=None):
    """-> output, or (output, saved) when ``keep``: ``saved`` is the workspace tensor whose head holds the
    layouts the matching backward can reuse (None when the tensor path does not run both directions)."""
    nd = len(kernel)
    modulated = mask is not None
    # channels-last activations (torch.channels_last / channels_last_3d) are taken as they are -- the kernels gather
    # from that layout anyway -- and the output comes back channels-last; shapes / dtypes the tcgen05 kernels do
    # not take fall back to an NCHW copy
    cl = _is_channels_last(input) and (output is None or _is_channels_last(output))
    if cl and not _channels_last_ok(input, weight, kernel, stride, pad, dil, group, dgroup, modulated, with_bias,
                                    False):
        cl = False
    if not cl and _is_channels_last(input) and output is None:
        input = input.contiguous()
    named = [("input", input), ("weight", weight), ("bias", bias), ("offset", offset)]
    if modulated:
        named.append(("mask", mask))
    if output is not None:
        named.append(("output", output))
    _validate(named, weight, kernel, group, nd, cl)
    g = _geom(input, weight, kernel, stride, pad, dil, group, dgroup, modulated, with_bias, cl)
    shape = _out_shape(g, input, weight)
    K = 1
    for k in kernel:
        K *= k
    _expect(offset, (input.shape[0], dgroup * nd * K) + shape[2:], "offset")
    if modulated:
        _expect(mask, (input.shape[0], dgroup * K) + shape[2:], "mask")
    if with_bias:
        _expect(bias, (weight.shape[0],), "bias")
    with _on_device(input.device):
        if output is None:
            output = torch.empty(shape, dtype=input.dtype, device=input.device,
                                 memory_format=_cl_format(nd) if cl else torch.contiguous_format)
        else:
            _expect(output, shape, "output")
        stream = _raw_stream()
        ex = None
        if prelaid is not None and not cl and prelaid.device == input.device and \
                prelaid.numel() >= int(_lib.mdconv_workspace_bytes(ctypes.byref(g), 0)) > 0 and \
                int(_lib.mdconv_channels_last_bytes(ctypes.byref(g))) > 0:
            # `prelaid` is a forward workspace whose head already holds the channels-last copy of THIS input
            # (to_channels_last): use it as the workspace, skip the re-layout pass
            ws, n = prelaid, prelaid.numel()
            ex = _FwdExtra()
            ex.input_channels_last = ws.data_ptr()
        else:
            ws, n = _workspace(g, False, input.device)
        rc = _lib.mdconv_forward_ex(ctypes.byref(g), _ptr(input), _ptr(weight), _ptr(bias) if with_bias else None,
                                    _ptr(offset), _ptr(mask), _ptr(output), _ptr(ws), n,
                                    ctypes.byref(ex) if ex is not None else None, ctypes.c_void_p(stream))
    _check(rc, "mdconv_forward")
    if keep:
        nsave = int(_lib.mdconv_saved_bytes(ctypes.byref(g))) if (_keep_layouts and ws is not None) else 0
        return output, ((ws, nsave, g.path, cl) if nsave > 0 else None)
    return output


def _backward(input, weight, bias, offset, mask, grad_output, grads, kernel, stride, pad, dil, group, dgroup,
              with_bias, saved=None, weight_ready=None, want_f32=False):
    """grads: None (allocate and return) or dict of caller-owned tensors (out-param style).
    saved: what ``_forward(..., keep=True)`` returned for the same input / weight / geometry, or None.
    weight_ready: a ``torch.cuda.Event`` recorded on the current stream once grad_weight / grad_bias are final
    (before the data gradients are computed on the tensor path), or None.
    want_f32: also return ``flat_f32`` = fp32 [grad_weight || grad_bias] before rounding to the storage dtype (what a
    batch-sharded caller all-reduces)."""
    nd = len(kernel)
    modulated = mask is not None
    # channels-last input -> channels-last grad_output in, channels-last grad_input out (when the tcgen05 backward
    # takes the shape; otherwise NCHW copies)
    cl = _is_channels_last(input) and grads is None and \
        _channels_last_ok(input, weight, kernel, stride, pad, dil, group, dgroup, modulated, with_bias, True)
    if cl:
        grad_output = grad_output.contiguous(memory_format=_cl_format(nd))
    else:
        if _is_channels_last(input) and grads is None:
            input = input.contiguous()
        if grads is None:
            grad_output = grad_output.contiguous()
    named = [("input", input), ("weight", weight), ("bias", bias), ("offset", offset)]
    if modulated:
        named.append(("mask", mask))
    named.append(("grad_output", grad_output))
    if grads is not None:
        named += [(k, v) for k, v in grads.items() if v is not None]
    _validate(named, weight, kernel, group, nd, cl)
    g = _geom(input, weight, kernel, stride, pad, dil, group, dgroup, modulated, with_bias, cl)
    _expect(grad_output, _out_shape(g, input, weight), "grad_output")
    with _on_device(input.device):
        if grads is None:
            grads = {"grad_input": torch.empty_like(input, memory_format=torch.preserve_format),
                     "grad_offset": torch.empty_like(offset),
                     "grad_mask": torch.empty_like(mask) if modulated else None}
            if want_f32 and with_bias and input.dtype != torch.float64:
                # the two parameter gradients as views of ONE buffer, [grad_weight || grad_bias] like flat_f32: a
                # batch-sharded caller rounds the reduced fp32 sum into both with a single copy
                nw = weight.numel()
                both = torch.empty(nw + bias.numel(), dtype=weight.dtype, device=input.device)
                grads["grad_weight"] = both[:nw].view(weight.shape)
                grads["grad_bias"] = both[nw:]
                grads["flat_storage"] = both
            else:
                grads["grad_weight"] = torch.empty_like(weight)
                grads["grad_bias"] = torch.empty_like(bias) if with_bias else torch.zeros_like(bias)
        else:
            _expect(grads["grad_input"], input.shape, "grad_input")
            _expect(grads["grad_offset"], offset.shape, "grad_offset")
            _expect(grads["grad_weight"], weight.shape, "grad_weight")
            if modulated:
                _expect(grads["grad_mask"], mask.shape, "grad_mask")
        ws, n = _workspace(g, True, input.device)
        stream = _raw_stream()
        sv_ptr, sv_n = None, 0
        if saved is not None and saved[2] == g.path and saved[0].device == input.device and \
                (len(saved) < 4 or saved[3] == cl) and int(_lib.mdconv_saved_bytes(ctypes.byref(g))) == saved[1]:
            sv_ptr, sv_n = _ptr(saved[0]), saved[1]
        ex = _BwdExtra()
        ex.saved = sv_ptr.value if sv_ptr is not None else None
        ex.saved_bytes = sv_n
        ex.weight_ready_event = weight_ready.cuda_event if weight_ready is not None else None
        if want_f32 and input.dtype != torch.float64:
            nw = weight.numel()
            flat = torch.empty(nw + (bias.numel() if with_bias else 0), dtype=torch.float32, device=input.device)
            ex.grad_weight_f32 = flat.data_ptr()
            ex.grad_bias_f32 = flat.data_ptr() + 4 * nw if with_bias else None
            grads["flat_f32"] = flat
        rc = _lib.mdconv_backward_ex(ctypes.byref(g), _ptr(input), _ptr(weight), _ptr(offset), _ptr(mask),
                                     _ptr(grad_output), _ptr(grads["grad_input"]), _ptr(grads["grad_weight"]),
                                     _ptr(grads["grad_bias"]) if with_bias else None, _ptr(grads["grad_offset"]),
                                     _ptr(grads.get("grad_mask")) if modulated else None, _ptr(ws), n,
                                     ctypes.byref(ex), ctypes.c_void_p(stream))
    _check(rc, "mdconv_backward")
    return grads


def forward_ex(input, weight, bias, offset, mask, output, kernel, stride, pad, dil, group, deformable_group,
               with_bias, keep=False, prelaid=None):
    """Generic form of the four forwards (mask None = plain op; output None = allocate).  With ``keep`` it
    returns (output, saved) for ``backward_ex``; used by the autograd Functions.  ``prelaid``: what
    ``to_channels_last`` returned for this input and geometry (the re-layout pass is then skipped)."""
    return _forward(input, weight, bias, offset, mask, output, tuple(kernel), tuple(stride), tuple(pad), tuple(dil),
                    group, deformable_group, with_bias, keep=keep, prelaid=prelaid)


def to_channels_last(input, weight, kernel, stride, pad, dil, group, deformable_group, modulated, with_bias):
    """Run the forward's re-layout pass on its own: -> a forward WORKSPACE tensor for this geometry whose head holds
    the channels-last copy of ``input`` (pass it as ``prelaid`` to forward_ex / producer_conv), or None when the
    tcgen05 path does not take the call (then every op simply does its own thing)."""
    if not input.is_cuda or input.dtype not in (torch.float16, torch.bfloat16) or not input.is_contiguous():
        return None
    g = _geom(input, weight, tuple(kernel), tuple(stride), tuple(pad), tuple(dil), group, deformable_group, modulated,
              with_bias)
    if int(_lib.mdconv_channels_last_bytes(ctypes.byref(g))) == 0:
        return None
    with _on_device(input.device):
        ws, n = _workspace(g, False, input.device)
        rc = _lib.mdconv_to_channels_last(ctypes.byref(g), _ptr(input), _ptr(ws),
                                          ctypes.c_void_p(_raw_stream()))
    _check(rc, "mdconv_to_channels_last")
    return ws


def producer_conv(input, weight, bias, n_first, kernel, stride, pad, prelaid=None):
    """Plain convolution on the deformable forward kernel in regular-grid mode (mdconv_fwd_extra_t.regular_grid):
    the offset / mask producer of the Pack modules.  weight [n, C, *kernel], bias [n]; the first ``n_first`` output
    channels come back as one tensor and the rest as a second one (None when n_first == n).  The filter bank is
    padded with zero rows to a multiple of 32 internally.  fp16 / bf16, C a multiple of 64, dilation 1, groups 1."""
    nd = len(kernel)
    n = weight.shape[0]
    npad = (n + 31) // 32 * 32
    if npad != n:
        weight = torch.cat([weight, weight.new_zeros((npad - n,) + tuple(weight.shape[1:]))], 0)
        bias = torch.cat([bias, bias.new_zeros(npad - n)], 0)
    weight, bias = weight.contiguous(), bias.contiguous()
    _validate([("input", input), ("weight", weight), ("bias", bias)], weight, kernel, 1, nd)
    g = _geom(input, weight, kernel, stride, pad, (1,) * nd, 1, 1, False, True)
    g.path = 2
    shape = _out_shape(g, input, weight)
    with _on_device(input.device):
        out1 = torch.empty((shape[0], n_first) + shape[2:], dtype=input.dtype, device=input.device)
        out2 = torch.empty((shape[0], n - n_first) + shape[2:], dtype=input.dtype, device=input.device) \
            if n_first < n else None
        need = int(_lib.mdconv_workspace_bytes(ctypes.byref(g), 0))
        ex = _FwdExtra()
        ex.regular_grid = 1
        ex.c_out_valid = n
        ex.output2 = out2.data_ptr() if out2 is not None else None
        ex.split = n_first if out2 is not None else 0
        if prelaid is not None and prelaid.device == input.device:
            # the channels-last copy at the head of the deformable conv's workspace is shared; the re-ordered
            # producer filters go to a workspace of their own (same head offset, so only the tail is written)
            ex.input_channels_last = prelaid.data_ptr()
            need = (weight.numel() * weight.element_size() + 255) // 256 * 256
        ws = torch.empty(need, dtype=torch.uint8, device=input.device)
        rc = _lib.mdconv_forward_ex(ctypes.byref(g), _ptr(input), _ptr(weight), _ptr(bias), None, None, _ptr(out1),
                                    _ptr(ws), need, ctypes.byref(ex),
                                    ctypes.c_void_p(_raw_stream()))
    _check(rc, "mdconv_forward_ex(regular_grid)")
    return out1, out2


def producer_conv_supported(input, n_out, kernel, stride, pad) -> bool:
    if not input.is_cuda or input.dtype not in (torch.float16, torch.bfloat16) or not input.is_contiguous():
        return False
    npad = (n_out + 31) // 32 * 32
    w = torch.empty((npad, input.shape[1]) + tuple(kernel), dtype=input.dtype, device="meta")
    g = _geom(input, w, tuple(kernel), tuple(stride), tuple(pad), (1,) * len(kernel), 1, 1, False, True)
    g.path = 0
    return int(_lib.mdconv_selected_path(ctypes.byref(g), 0)) == 2


def backward_ex(input, weight, bias, offset, mask, grad_output, kernel, stride, pad, dil, group, deformable_group,
                with_bias, saved=None, weight_ready=None, want_f32=False):
    """Generic form of the four backwards -> dict grad_input / grad_offset / grad_mask / grad_weight / grad_bias
    (+ flat_f32 with ``want_f32``)."""
    return _backward(input, weight, bias, offset, mask, grad_output, None, tuple(kernel), tuple(stride), tuple(pad),
                     tuple(dil), group, deformable_group, with_bias, saved=saved, weight_ready=weight_ready,
                     want_f32=want_f32)


# ---- the eight reference entry points ------------------------------------------------------------

def deform_conv2d_forward_cuda(input, weight, bias, offset, output, kernel_h, kernel_w, stride_h, stride_w,
                               pad_h, pad_w, dilation_h, dilation_w, group, deformable_group, in_step, with_bias):
    """reference: src/deformable_conv.cu:117-196 (out-param, returns 0).  Unlike the reference for
    gcd(batch, in_step) > 1 (SURVEY B-1), the caller's ``output`` always receives the correct result."""
    _forward(input, weight, bias, offset, None, output, (kernel_h, kernel_w), (stride_h, stride_w),
             (pad_h, pad_w), (dilation_h, dilation_w), group, deformable_group, with_bias)
    return 0


def deform_conv2d_backward_cuda(input, weight, bias, offset, grad_input, grad_weight, grad_bias, grad_offset,
                                grad_output, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                                dilation_w, group, deformable_group, in_step, with_bias):
    """reference: src/deformable_conv.cu:327-431."""
    _backward(input, weight, bias, offset, None, grad_output,
              {"grad_input": grad_input, "grad_weight": grad_weight, "grad_bias": grad_bias,
               "grad_offset": grad_offset},
              (kernel_h, kernel_w), (stride_h, stride_w), (pad_h, pad_w), (dilation_h, dilation_w), group,
              deformable_group, with_bias)
    return 0


def modulated_deform_conv2d_forward_cuda(input, weight, bias, offset, mask, kernel_h, kernel_w, stride_h,
                                         stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group,
                                         in_step, with_bias):
    """reference: src/mdeformable_conv.cu:120-194 (returns the output tensor)."""
    return _forward(input, weight, bias, offset, mask, None, (kernel_h, kernel_w), (stride_h, stride_w),
                    (pad_h, pad_w), (dilation_h, dilation_w), group, deformable_group, with_bias)


def modulated_deform_conv2d_backward_cuda(input, weight, bias, offset, mask, grad_output, kernel_h, kernel_w,
                                          stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                                          deformable_group, in_step, with_bias):
    """reference: src/mdeformable_conv.cu:361-458 (returns the 5-tuple
    grad_input, grad_offset, grad_mask, grad_weight, grad_bias)."""
    g = _backward(input, weight, bias, offset, mask, grad_output, None, (kernel_h, kernel_w),
                  (stride_h, stride_w), (pad_h, pad_w), (dilation_h, dilation_w), group, deformable_group,
                  with_bias)
    return g["grad_input"], g["grad_offset"], g["grad_mask"], g["grad_weight"], g["grad_bias"]


def deform_conv3d_forward_cuda(input, weight, bias, offset, output, kernel_h, kernel_w, kernel_l, stride_h,
                               stride_w, stride_l, pad_h, pad_w, pad_l, dilation_h, dilation_w, dilation_l,
                               group, deformable_group, in_step, with_bias):
    """reference: src/deformable_conv3d.cu:160-256."""
    _forward(input, weight, bias, offset, None, output, (kernel_h, kernel_w, kernel_l),
             (stride_h, stride_w, stride_l), (pad_h, pad_w, pad_l), (dilation_h, dilation_w, dilation_l), group,
             deformable_group, with_bias)
    return 0


def deform_conv3d_backward_cuda(input, weight, bias, offset, grad_input, grad_weight, grad_bias, grad_offset,
                                grad_output, kernel_h, kernel_w, kernel_l, stride_h, stride_w, stride_l, pad_h,
                                pad_w, pad_l, dilation_h, dilation_w, dilation_l, group, deformable_group,
                                in_step, with_bias):
    """reference: src/deformable_conv3d.cu:434-561."""
    _backward(input, weight, bias, offset, None, grad_output,
              {"grad_input": grad_input, "grad_weight": grad_weight, "grad_bias": grad_bias,
               "grad_offset": grad_offset},
              (kernel_h, kernel_w, kernel_l), (stride_h, stride_w, stride_l), (pad_h, pad_w, pad_l),
              (dilation_h, dilation_w, dilation_l), group, deformable_group, with_bias)
    return 0


def modulated_deform_conv3d_forward_cuda(input, weight, bias, offset, mask, output, kernel_h, kernel_w, kernel_l,
                                         stride_h, stride_w, stride_l, pad_h, pad_w, pad_l, dilation_h,
                                         dilation_w, dilation_l, group, deformable_group, in_step, with_bias):
    """reference: src/mdeformable_conv3d.cu:170-262."""
    _forward(input, weight, bias, offset, mask, output, (kernel_h, kernel_w, kernel_l),
             (stride_h, stride_w, stride_l), (pad_h, pad_w, pad_l), (dilation_h, dilation_w, dilation_l), group,
             deformable_group, with_bias)
    return 0


def modulated_deform_conv3d_backward_cuda(input, weight, bias, offset, mask, grad_input, grad_weight, grad_bias,
                                          grad_offset, grad_mask, grad_output, kernel_h, kernel_w, kernel_l,
                                          stride_h, stride_w, stride_l, pad_h, pad_w, pad_l, dilation_h,
                                          dilation_w, dilation_l, group, deformable_group, in_step, with_bias):
    """reference: src/mdeformable_conv3d.cu:443-586."""
    _backward(input, weight, bias, offset, mask, grad_output,
              {"grad_input": grad_input, "grad_weight": grad_weight, "grad_bias": grad_bias,
               "grad_offset": grad_offset, "grad_mask": grad_mask},
              (kernel_h, kernel_w, kernel_l), (stride_h, stride_w, stride_l), (pad_h, pad_w, pad_l),
              (dilation_h, dilation_w, dilation_l), group, deformable_group, with_bias)
    return 0
