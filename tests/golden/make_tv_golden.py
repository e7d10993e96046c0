"""Generates tests/golden/tv_*.npz: torchvision.ops.deform_conv2d (CPU, fp64) forward and all
gradients on seeded inputs.  torchvision's op is an independent implementation of the same 2D
math (identical to the reference at non-integer sampling coordinates), so these are
golden vectors for the 2D oracle.  Run from the repo root:  python tests/golden/make_tv_golden.py
"""
import os
import sys

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from cases import Case, make_inputs  # noqa: E402

CASES = [
    Case("tv_m2d_g2dg2", 2, True, g=2, dg=2, seed=100),
    Case("tv_m2d_s2d2", 2, True, size=(11, 9), s=(2, 1), p=(2, 1), d=(2, 1), seed=101),
    Case("tv_d2d_dg4", 2, False, C=8, Co=8, g=1, dg=4, seed=102),
    Case("tv_d2d_k1x3", 2, False, k=(1, 3), p=(0, 1), B=3, seed=103),
]

for c in CASES:
    t = make_inputs(c, torch.float64)
    leaves = {k: (v.clone().requires_grad_() if v is not None and k != "grad_output" else v) for k, v in t.items()}
    y = torchvision.ops.deform_conv2d(leaves["x"], leaves["offset"], leaves["weight"], leaves["bias"],
                                      stride=c.s, padding=c.p, dilation=c.d, mask=leaves["mask"])
    y.backward(t["grad_output"])
    out = {k: v.numpy() for k, v in t.items() if v is not None}
    out.update(y=y.detach().numpy(), grad_input=leaves["x"].grad.numpy(), grad_offset=leaves["offset"].grad.numpy(),
               grad_weight=leaves["weight"].grad.numpy(), grad_bias=leaves["bias"].grad.numpy(),
               stride=np.array(c.s), padding=np.array(c.p), dilation=np.array(c.d), groups=np.array(c.g),
               dg=np.array(c.dg), torchvision=np.array(torchvision.__version__))
    if c.modulated:
        out["grad_mask"] = leaves["mask"].grad.numpy()
    np.savez_compressed(os.path.join(HERE, c.name + ".npz"), **out)
    print("wrote", c.name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim > 1})
