"""Uniform Python front-end to the UNMODIFIED reference kernels in oracle/_ref/ (built by
oracle/build_ref.py).  TEST INFRASTRUCTURE: used by the -m gpu tests, oracle/gen_ref_golden.py and
bench.py's reference-CUDA timing; never by the product package.

Reference defect handled here (SURVEY B-1): the three out-param FORWARD functions leave the caller's
buffer scrambled unless gcd(batch, in_step) == 1, so forwards default to in_step=1 for parity
(timing callers pass the reference's default in_step=64 explicitly).
"""
import importlib.util
import os

import torch

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mods = {}


def available() -> bool:
    return all(os.path.exists(os.path.join(_REF_DIR, n + ".so")) for n in ("MDCONV_REF_M2D", "MDCONV_REF_X6"))


def _mod(name):
    if name not in _mods:
        spec = importlib.util.spec_from_file_location(name, os.path.join(_REF_DIR, name + ".so"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mods[name] = m
    return _mods[name]


def _geo(weight, stride, padding, dilation):
    return (*weight.shape[2:], *stride, *padding, *dilation)


def forward(x, weight, bias, offset, mask, stride, padding, dilation, groups, dg, in_step=1):
    nd = x.dim() - 2
    with_bias = bias is not None
    b = bias if with_bias else x.new_empty(0)
    geo = _geo(weight, stride, padding, dilation)
    if nd == 2 and mask is not None:
        return _mod("MDCONV_REF_M2D").modulated_deform_conv2d_forward_cuda(
            x, weight, b, offset, mask, *geo, groups, dg, in_step, with_bias)
    out_sp = tuple((i + 2 * p - (d * (k - 1) + 1)) // s + 1
                   for i, k, s, p, d in zip(x.shape[2:], weight.shape[2:], stride, padding, dilation))
    out = x.new_empty((x.shape[0], weight.shape[0]) + out_sp)
    m = _mod("MDCONV_REF_X6")
    if nd == 2:
        m.deform_conv2d_forward_cuda(x, weight, b, offset, out, *geo, groups, dg, in_step, with_bias)
    elif mask is None:
        m.deform_conv3d_forward_cuda(x, weight, b, offset, out, *geo, groups, dg, in_step, with_bias)
    else:
        m.modulated_deform_conv3d_forward_cuda(x, weight, b, offset, mask, out, *geo, groups, dg, in_step, with_bias)
    return out


def backward(x, weight, bias, offset, mask, grad_output, stride, padding, dilation, groups, dg, in_step=64):
    nd = x.dim() - 2
    with_bias = bias is not None
    b = bias if with_bias else x.new_empty(0)
    geo = _geo(weight, stride, padding, dilation)
    go = grad_output.contiguous().clone()  # the reference transposes grad_output in place
    if nd == 2 and mask is not None:
        gi, goff, gm, gw, gb = _mod("MDCONV_REF_M2D").modulated_deform_conv2d_backward_cuda(
            x, weight, b, offset, mask, go, *geo, groups, dg, in_step, with_bias)
        return {"grad_input": gi, "grad_offset": goff, "grad_mask": gm, "grad_weight": gw,
                "grad_bias": gb if with_bias else None}
    gi, gw, gb, goff = (torch.zeros_like(t) for t in (x, weight, b, offset))
    m = _mod("MDCONV_REF_X6")
    if nd == 2:
        m.deform_conv2d_backward_cuda(x, weight, b, offset, gi, gw, gb, goff, go, *geo, groups, dg, in_step, with_bias)
        gm = None
    elif mask is None:
        m.deform_conv3d_backward_cuda(x, weight, b, offset, gi, gw, gb, goff, go, *geo, groups, dg, in_step, with_bias)
        gm = None
    else:
        gm = torch.zeros_like(mask)
        m.modulated_deform_conv3d_backward_cuda(x, weight, b, offset, mask, gi, gw, gb, goff, gm, go, *geo, groups,
                                                dg, in_step, with_bias)
    return {"grad_input": gi, "grad_offset": goff, "grad_mask": gm, "grad_weight": gw,
            "grad_bias": gb if with_bias else None}
