#!/usr/bin/env python
"""Per-kernel times of a few fwd+bwd steps of a bench workload (library-side CUDA events).
usage: python scripts/kernel_times.py [cfg2|cfg3|cfg4] [steps] [fwd_only]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "modulated-deform-conv_b200"), ROOT]
import torch
import bench
import MDCONV_CUDA
import modulated_deform_conv as mdc
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
fwd_only = len(sys.argv) > 3 and sys.argv[3] == "fwd_only"
dev = torch.device("cuda", 0)
dtype = bench.DTYPES[w["dtype"]]
cls = {(2, True): mdc.ModulatedDeformConv2d, (2, False): mdc.DeformConv2d,
       (3, True): mdc.ModulatedDeformConv3d, (3, False): mdc.DeformConv3d}[(w["nd"], w["modulated"])]
torch.manual_seed(0)
mod = cls(w["C"], w["Co"], w["k"], stride=w["s"], padding=w["p"], dilation=w["d"], groups=w["g"],
          deformable_groups=w["dg"], bias=False).to(dev).to(dtype)
h = bench.make_host_inputs(w, w["B"], 100, dtype, pin=False)
d = {k: (None if v is None else v.to(dev)) for k, v in h.items()}
lv = [d[k].clone().requires_grad_(not fwd_only) for k in ("x", "offset") + (("mask",) if w["modulated"] else ())]
def step():
    y = mod(*lv)
    if not fwd_only:
        y.backward(d["grad_output"])
for _ in range(3):
    step()
torch.cuda.synchronize()
MDCONV_CUDA.profile_enable(True)
for _ in range(steps):
    step()
kt = {}
for name, ms in MDCONV_CUDA.profile_read():
    kt.setdefault(name, []).append(ms)
MDCONV_CUDA.profile_enable(False)
tot = 0.0
for k, v in kt.items():
    per = sum(v) / steps
    tot += per
    print(f"  {k:22s} {per*1e3:9.1f} us/step  ({len(v)//steps} launches)")
print(f"  {'sum':22s} {tot*1e3:9.1f} us/step   MDCONV_DBG={os.environ.get('MDCONV_DBG','0')}")
