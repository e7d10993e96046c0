#!/usr/bin/env python
"""Kernel timeline of consecutive fwd+bwd steps of a workload with the side stream ON (library-side CUDA events:
start offset and duration of every launch).  usage: python scripts/timeline.py [workload] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "modulated-deform-conv_b200"), ROOT]
import torch

import bench
import MDCONV_CUDA
import modulated_deform_conv as mdc

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
dtype = bench.DTYPES[w["dtype"]]
cls = {(2, True): mdc.ModulatedDeformConv2d, (2, False): mdc.DeformConv2d,
       (3, True): mdc.ModulatedDeformConv3d, (3, False): mdc.DeformConv3d}[(w["nd"], w["modulated"])]
torch.manual_seed(0)
mod = cls(w["C"], w["Co"], w["k"], stride=w["s"], padding=w["p"], dilation=w["d"], groups=w["g"],
          deformable_groups=w["dg"], bias=False).to(dev).to(dtype)
h = bench.make_host_inputs(w, w["B"], 100, dtype, pin=False)
d = {k: (None if v is None else v.to(dev)) for k, v in h.items()}
lv = [d[k].clone().requires_grad_() for k in ("x", "offset") + (("mask",) if w["modulated"] else ())]


def step():
    for t in lv:
        t.grad = None
    mod.weight.grad = None
    mod(*lv).backward(d["grad_output"])


for _ in range(5):
    step()
torch.cuda.synchronize()
MDCONV_CUDA.profile_enable(True, keep_side_stream=True)
for _ in range(steps):
    step()
torch.cuda.synchronize()
tl = MDCONV_CUDA.profile_read(with_start=True)
MDCONV_CUDA.profile_enable(False)
print("name                 start_us  +dur_us   end_us")
for name, ms, at in tl:
    print(f"{name:20s} {at * 1e3:9.1f} +{ms * 1e3:7.1f} {(at + ms) * 1e3:9.1f}")
