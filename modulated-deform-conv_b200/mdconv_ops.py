"""torch.library registration of the deformable-conv path.

The reference exposes raw pybind functions plus ``autograd.Function`` subclasses: opaque to ``torch.compile``,
fake-tensor tracing and ``torch.export``.  This module registers the same computation as two dispatcher ops,

    torch.ops.mdconv_b200.deform_conv(input, weight, bias?, offset, mask?, stride, padding, dilation,
                                      groups, deformable_groups) -> Tensor
    torch.ops.mdconv_b200.deform_conv_backward(input, weight, bias?, offset, mask?, grad_output, ...)
        -> (grad_input, grad_offset, grad_mask, grad_weight, grad_bias)

with fake (meta) implementations and an autograd formula, so a model using them traces with
``torch.compile(fullgraph=True)``.  ``mask=None`` selects the plain op, a 5-d input the 3D one; the gradient
rule at integer coordinates follows the reference op of that kind (A for modulated 2D, B otherwise).
The kernels are the ones behind ``MDCONV_CUDA`` (no second implementation); CPU tensors raise.
"""
from typing import List, Optional, Tuple

import torch

import MDCONV_CUDA

__all__ = ["deform_conv", "deform_conv_backward"]


def _out_spatial(in_sp, k_sp, stride, padding, dilation):
    return [(i + 2 * p - (d * (k - 1) + 1)) // s + 1 for i, k, s, p, d in zip(in_sp, k_sp, stride, padding, dilation)]


@torch.library.custom_op("mdconv_b200::deform_conv", mutates_args=())
def deform_conv(input: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], offset: torch.Tensor,
                mask: Optional[torch.Tensor], stride: List[int], padding: List[int], dilation: List[int],
                groups: int, deformable_groups: int) -> torch.Tensor:
    if not input.is_cuda:
        raise NotImplementedError("mdconv_b200::deform_conv has no CPU path")
    b = bias if bias is not None else input.new_empty(0)
    return MDCONV_CUDA.forward_ex(input.contiguous(), weight.contiguous(), b, offset.contiguous(),
                                  None if mask is None else mask.contiguous(), None, tuple(weight.shape[2:]),
                                  tuple(stride), tuple(padding), tuple(dilation), groups, deformable_groups,
                                  bias is not None)


@deform_conv.register_fake
def _(input, weight, bias, offset, mask, stride, padding, dilation, groups, deformable_groups):
    nd = input.dim() - 2
    torch._check(weight.dim() == nd + 2, lambda: "weight rank must match input rank")
    torch._check(input.shape[1] == weight.shape[1] * groups, lambda: "input channels != weight channels * groups")
    sp = _out_spatial(input.shape[2:], weight.shape[2:], stride, padding, dilation)
    return input.new_empty((input.shape[0], weight.shape[0], *sp))


@torch.library.custom_op("mdconv_b200::deform_conv_backward", mutates_args=())
def deform_conv_backward(input: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                         offset: torch.Tensor, mask: Optional[torch.Tensor], grad_output: torch.Tensor,
                         stride: List[int], padding: List[int], dilation: List[int], groups: int,
                         deformable_groups: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor,
                                                          torch.Tensor]:
    if not input.is_cuda:
        raise NotImplementedError("mdconv_b200::deform_conv_backward has no CPU path")
    b = bias if bias is not None else input.new_empty(0)
    g = MDCONV_CUDA.backward_ex(input.contiguous(), weight.contiguous(), b, offset.contiguous(),
                                None if mask is None else mask.contiguous(), grad_output.contiguous(),
                                tuple(weight.shape[2:]), tuple(stride), tuple(padding), tuple(dilation), groups,
                                deformable_groups, bias is not None)
    # the dispatcher wants tensors in every slot: absent gradients are 0-element tensors
    gm = g["grad_mask"] if mask is not None else input.new_empty(0)
    gb = g["grad_bias"] if bias is not None else input.new_empty(0)
    return g["grad_input"], g["grad_offset"], gm, g["grad_weight"], gb


@deform_conv_backward.register_fake
def _(input, weight, bias, offset, mask, grad_output, stride, padding, dilation, groups, deformable_groups):
    return (torch.empty_like(input), torch.empty_like(offset),
            torch.empty_like(mask) if mask is not None else input.new_empty(0), torch.empty_like(weight),
            torch.empty_like(bias) if bias is not None else input.new_empty(0))


def _setup_context(ctx, inputs, output):
    input, weight, bias, offset, mask, stride, padding, dilation, groups, deformable_groups = inputs
    ctx.save_for_backward(input, weight, bias, offset, mask)
    ctx.geo = (list(stride), list(padding), list(dilation), groups, deformable_groups)


def _backward(ctx, grad_output):
    input, weight, bias, offset, mask = ctx.saved_tensors
    gi, goff, gm, gw, gb = deform_conv_backward(input, weight, bias, offset, mask, grad_output, *ctx.geo)
    return (gi, gw, gb if bias is not None else None, goff, gm if mask is not None else None,
            None, None, None, None, None)


deform_conv.register_autograd(_backward, setup_context=_setup_context)
