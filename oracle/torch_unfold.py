"""Independent pure-PyTorch (autograd) N-d deformable "unfold + matmul" used as a second opinion
on the C oracle (tests and the cpu_baseline leg of bench.py only).  Written from SURVEY.md appendix A, not from the oracle's code:
columns are built with explicit corner gathers, gradients come from autograd, so at generic
(non-integer) sampling coordinates they must equal the reference's hand-written gradient rules.
"""
import itertools

import torch


def _tup(v, nd):
    return (v,) * nd if isinstance(v, int) else tuple(v)


def deform_unfold(x, offset, mask, kernel, stride, padding, dilation, dg):
    """-> columns [B, C, K, P] (C-major then tap, the reference's column row order c*K+tap)."""
    nd = x.dim() - 2
    B, C = x.shape[:2]
    in_size = x.shape[2:]
    out_size = offset.shape[2:]
    K = 1
    for k in kernel:
        K *= k
    P = 1
    for o in out_size:
        P *= o
    Cdg = C // dg
    xf = x.reshape(B, dg, Cdg, -1)
    off = offset.reshape(B, dg, K, nd, P)
    grids = torch.meshgrid(*[torch.arange(o, dtype=x.dtype, device=x.device) for o in out_size], indexing="ij")
    out_idx = [g.reshape(-1) for g in grids]  # each [P]
    cols = []
    for tap, t in enumerate(itertools.product(*[range(k) for k in kernel])):
        pos, lo, fr = [], [], []
        in_range = torch.ones(B, dg, P, dtype=torch.bool, device=x.device)
        for a in range(nd):
            p_a = out_idx[a] * stride[a] - padding[a] + t[a] * dilation[a] + off[:, :, tap, a]
            l_a = torch.floor(p_a)
            pos.append(p_a)
            lo.append(l_a.detach())
            fr.append(p_a - l_a.detach())
            in_range &= (p_a > -1) & (p_a < in_size[a])
        val = torch.zeros(B, dg, Cdg, P, dtype=x.dtype, device=x.device)
        for corner in itertools.product((0, 1), repeat=nd):
            w = torch.ones(B, dg, P, dtype=x.dtype, device=x.device)
            ok = in_range.clone()
            lin = torch.zeros(B, dg, P, dtype=torch.long, device=x.device)
            for a in range(nd):
                ia = lo[a].long() + corner[a]
                ok &= (ia >= 0) & (ia <= in_size[a] - 1)
                w = w * (fr[a] if corner[a] else (1 - fr[a]))
                lin = lin * in_size[a] + ia.clamp(0, in_size[a] - 1)
            v = torch.gather(xf, 3, lin[:, :, None, :].expand(B, dg, Cdg, P))
            val = val + (w * ok)[:, :, None, :] * v
        if mask is not None:
            val = val * mask.reshape(B, dg, K, P)[:, :, tap][:, :, None, :]
        cols.append(val.reshape(B, C, P))
    return torch.stack(cols, dim=2)  # [B, C, K, P]


def deform_conv_nd(x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                   deformable_groups=1):
    nd = x.dim() - 2
    stride, padding, dilation = _tup(stride, nd), _tup(padding, nd), _tup(dilation, nd)
    kernel = tuple(weight.shape[2:])
    B, C = x.shape[:2]
    Co = weight.shape[0]
    cols = deform_unfold(x, offset, mask, kernel, stride, padding, dilation, deformable_groups)
    K, P = cols.shape[2], cols.shape[3]
    cols = cols.reshape(B, groups, (C // groups) * K, P)
    w = weight.reshape(groups, Co // groups, (C // groups) * K)
    out = torch.einsum("gok,bgkp->bgop", w, cols).reshape(B, Co, *offset.shape[2:])
    if bias is not None:
        out = out + bias.reshape(1, Co, *([1] * nd))
    return out
