"""CPU oracle for the deformable / modulated-deformable convolution path.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs; the product package never imports it.

The arithmetic lives in ``mdconv_oracle.c`` (a plain-C restatement of the reference's algorithm,
see the header of that file for the reference file:line map and the parity-pin status); this
module is the ctypes binding plus dtype plumbing.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

# indices into the flat geometry block, mirror of the enum in mdconv_oracle.c
_G_COUNT = 23


def build(force: bool = False) -> str:
    """Compile liboracle.so (gcc, seconds).  Returns the library path."""
    src = os.path.join(_HERE, "mdconv_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, env=env, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _tuple_n(v, nd: int) -> Tuple[int, ...]:
    if isinstance(v, int):
        return (v,) * nd
    v = tuple(int(i) for i in v)
    assert len(v) == nd, (v, nd)
    return v


@dataclass
class Geometry:
    """Shape/geometry of one call; field meaning as in the reference's argument lists
    (mdeformable_conv.cu:120-126, mdeformable_conv3d.cu:170-177)."""
    nd: int
    batch: int
    c_in: int
    c_out: int
    in_size: Tuple[int, ...]
    kernel: Tuple[int, ...]
    stride: Tuple[int, ...]
    padding: Tuple[int, ...]
    dilation: Tuple[int, ...]
    groups: int = 1
    deformable_groups: int = 1

    @property
    def out_size(self) -> Tuple[int, ...]:
        return tuple((i + 2 * p - (d * (k - 1) + 1)) // s + 1
                     for i, k, s, p, d in zip(self.in_size, self.kernel, self.stride, self.padding, self.dilation))

    @property
    def taps(self) -> int:
        return int(np.prod(self.kernel))

    def block(self, variant: int, coord_f32: bool) -> np.ndarray:
        pad3 = lambda t, fill: list(t) + [fill] * (3 - len(t))
        g = [self.nd, self.batch, self.c_in, self.c_out]
        g += pad3(self.in_size, 1) + pad3(self.kernel, 1) + pad3(self.stride, 1)
        g += pad3(self.padding, 0) + pad3(self.dilation, 1)
        g += [self.groups, self.deformable_groups, int(variant), int(bool(coord_f32))]
        assert len(g) == _G_COUNT
        return np.asarray(g, dtype=np.int32)


def geometry_from_tensors(x, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1) -> Geometry:
    nd = x.ndim - 2
    return Geometry(nd=nd, batch=int(x.shape[0]), c_in=int(x.shape[1]), c_out=int(weight.shape[0]),
                    in_size=tuple(int(s) for s in x.shape[2:]), kernel=tuple(int(s) for s in weight.shape[2:]),
                    stride=_tuple_n(stride, nd), padding=_tuple_n(padding, nd), dilation=_tuple_n(dilation, nd),
                    groups=int(groups), deformable_groups=int(deformable_groups))


def _np(a, dt):
    if a is None:
        return None
    if hasattr(a, "detach"):  # torch tensor
        a = a.detach().cpu()
        if str(a.dtype) in ("torch.float16", "torch.bfloat16"):
            a = a.float()
        a = a.numpy()
    return np.ascontiguousarray(a, dtype=dt)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def default_variant(nd: int, modulated: bool) -> int:
    """Rule A (0) for modulated 2D, rule B (1) for plain 2D and both 3D ops."""
    return 0 if (nd == 2 and modulated) else 1


def forward(x, weight, bias, offset, mask=None, stride=1, padding=0, dilation=1, groups=1,
            deformable_groups=1, precision: str = "f64", coord_f32: Optional[bool] = None) -> np.ndarray:
    """Oracle forward.  Inputs: numpy arrays or torch tensors of any float dtype.
    ``precision`` picks the arithmetic of the restatement ("f64" for checking, "f32" for the timed
    CPU baseline).  ``coord_f32`` forms sampling coordinates in fp32 (what the reference's fp32
    kernels do); default: True iff the offset tensor is fp32."""
    geo = geometry_from_tensors(x, weight, stride, padding, dilation, groups, deformable_groups)
    if coord_f32 is None:
        coord_f32 = "float32" in str(offset.dtype)
    dt = np.float64 if precision == "f64" else np.float32
    sfx = "f64" if precision == "f64" else "f32"
    xa, wa, ba, oa, ma = (_np(t, dt) for t in (x, weight, bias, offset, mask))
    if ba is not None and ba.size == 0:
        ba = None
    y = np.empty((geo.batch, geo.c_out) + geo.out_size, dtype=dt)
    blk = geo.block(default_variant(geo.nd, mask is not None), coord_f32)
    rc = getattr(_load(), "mdo_forward_" + sfx)(_ptr(blk), _ptr(xa), _ptr(wa), _ptr(ba), _ptr(oa), _ptr(ma), _ptr(y))
    if rc != 0:
        raise RuntimeError(f"oracle forward failed rc={rc}")
    return y


def backward(x, weight, offset, mask, grad_output, stride=1, padding=0, dilation=1, groups=1,
             deformable_groups=1, variant: Optional[int] = None, precision: str = "f64",
             coord_f32: Optional[bool] = None, with_bias: bool = True):
    """Oracle backward.  Returns dict(grad_input, grad_offset, grad_mask, grad_weight, grad_bias)."""
    geo = geometry_from_tensors(x, weight, stride, padding, dilation, groups, deformable_groups)
    if coord_f32 is None:
        coord_f32 = "float32" in str(offset.dtype)
    if variant is None:
        variant = default_variant(geo.nd, mask is not None)
    dt = np.float64 if precision == "f64" else np.float32
    sfx = "f64" if precision == "f64" else "f32"
    xa, wa, oa, ma, ga = (_np(t, dt) for t in (x, weight, offset, mask, grad_output))
    gx = np.empty_like(xa)
    gw = np.empty_like(wa)
    gb = np.empty((geo.c_out,), dtype=dt) if with_bias else None
    goff = np.empty_like(oa)
    gmask = np.empty_like(ma) if ma is not None else None
    blk = geo.block(variant, coord_f32)
    rc = getattr(_load(), "mdo_backward_" + sfx)(_ptr(blk), _ptr(xa), _ptr(wa), _ptr(oa), _ptr(ma), _ptr(ga),
                                                 _ptr(gx), _ptr(gw), _ptr(gb), _ptr(goff), _ptr(gmask))
    if rc != 0:
        raise RuntimeError(f"oracle backward failed rc={rc}")
    return {"grad_input": gx, "grad_offset": goff, "grad_mask": gmask, "grad_weight": gw, "grad_bias": gb}
