#!/usr/bin/env python
"""Host-side cost of ENQUEUEING one fwd+bwd step of cfg2 (cProfile of the launch loop; the GPU runs behind).
usage: python scripts/host_profile.py [steps]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "modulated-deform-conv_b200"), ROOT]
import torch

import bench
import modulated_deform_conv as mdc

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
w = bench.WORKLOADS["cfg2"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
mod = mdc.ModulatedDeformConv2d(w["C"], w["Co"], w["k"], stride=w["s"], padding=w["p"], bias=False).to(dev).half()
h = bench.make_host_inputs(w, w["B"], 100, torch.float16, pin=False)
d = {k: v.to(dev) for k, v in h.items()}
lv = [d[k].clone().requires_grad_() for k in ("x", "offset", "mask")]


def step():
    for t in lv:
        t.grad = None
    mod.weight.grad = None
    mod(*lv).backward(d["grad_output"])


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / steps:.3f} ms/step; wall incl. drain {1e3 * (t2 - t0) / steps:.3f} ms/step")
# forward only / with torch.no_grad
with torch.no_grad():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        mod(*lv)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
print(f"forward-only host enqueue {1e3 * (t1 - t0) / steps:.3f} ms/call")
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for _ in range(steps):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
