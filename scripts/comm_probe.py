#!/usr/bin/env python
"""Where does the gradient all-reduce sit inside a data-parallel step?  (torchrun, N ranks.)
 1. the collective alone: 2.36 MB fp32 all-reduce on the default communicator and on the small-CTA one;
 2. a timeline of one backward: start -> weight-ready -> collective start/end -> backward end (CUDA events).
usage: python -m torch.distributed.run --nproc-per-node N scripts/comm_probe.py [comm_ctas ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "modulated-deform-conv_b200"), ROOT]
import torch
import torch.distributed as dist

import bench
import MDCONV_CUDA
import mdconv_dist
import modulated_deform_conv as mdc

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
# arguments: "ctas" or "ctas:cga" (cga -1 = NCCL's default cluster size)
ctas_list = [tuple(int(v) for v in (a.split(":") + ["1"])[:2]) for a in sys.argv[1:]] or [(4, 1)]


def ev():
    return torch.cuda.Event(enable_timing=True)


def time_allreduce(group, label, n=590080, iters=30):
    buf = torch.zeros(n, dtype=torch.float32, device=dev)
    for _ in range(5):
        dist.all_reduce(buf, group=group)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = ev(), ev()
    a.record()
    for _ in range(iters):
        dist.all_reduce(buf, group=group)
    b.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"allreduce alone [{label}] {n * 4 / 1e6:.2f} MB fp32: {a.elapsed_time(b) / iters * 1e3:.1f} us", flush=True)


time_allreduce(None, "default communicator")
for c, cga in ctas_list:
    time_allreduce(mdconv_dist.comm_group(c, cga), f"max_ctas={c} cga={cga}")

w = bench.WORKLOADS["cfg2"]
torch.manual_seed(0)
mod = mdc.ModulatedDeformConv2d(w["C"], w["Co"], w["k"], stride=w["s"], padding=w["p"], bias=False).to(dev).half()
h = bench.make_host_inputs(w, w["B"], 100 + rank, torch.float16, pin=False)
d = {k: v.to(dev) for k, v in h.items()}
lv = [d[k].clone().requires_grad_() for k in ("x", "offset", "mask")]

for c, cga in ctas_list + [(0, -1)]:
    mdconv_dist.detach(mod)
    sync = mdconv_dist.attach(mod, None, max_ctas=c if c > 0 else None, cga_cluster_size=cga)
    if c == 0:
        MDCONV_CUDA.set_comm_sms(8)

    def step(trace=False):
        for t in lv:
            t.grad = None
        mod.weight.grad = None
        y = mod(*lv)
        t0, t1 = ev(), ev()
        t0.record()
        y.backward(d["grad_output"])
        t1.record()
        return t0, t1
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sync.trace = []
    marks = []
    for _ in range(8):
        marks.append(step())
    torch.cuda.synchronize()
    rows = []
    for (t0, t1), (c0, c1) in zip(marks, sync.trace):
        rows.append((t0.elapsed_time(c0), t0.elapsed_time(c1), t0.elapsed_time(t1)))
    sync.trace = None
    m = [sum(r[i] for r in rows[2:]) / len(rows[2:]) * 1e3 for i in range(3)]
    print(f"rank {rank} comm_ctas={c} cga={cga}: backward start -> weight-ready {m[0]:.0f} us -> collective done {m[1]:.0f} us "
          f"(collective {m[1] - m[0]:.0f} us) -> backward end {m[2]:.0f} us", flush=True)
# 3. kernel timeline of two consecutive steps (library-side events, side stream left on) + the collective
mdconv_dist.detach(mod)
c, cga = ctas_list[0]
sync = mdconv_dist.attach(mod, None, max_ctas=c, cga_cluster_size=cga)
for _ in range(5):
    step()
torch.cuda.synchronize()
dist.barrier()
MDCONV_CUDA.profile_enable(True, keep_side_stream=True)
sync.trace = []
base = ev()
base.record()
for _ in range(3):
    step()
torch.cuda.synchronize()
tl = MDCONV_CUDA.profile_read(with_start=True)
MDCONV_CUDA.profile_enable(False)
if rank in (0, 1):
    lines = [f"rank {rank} timeline (us since the first launch; name start +duration)"]
    for name, ms, at in tl:
        lines.append(f"  {name:20s} {at * 1e3:9.1f} +{ms * 1e3:7.1f}")
    for c0, c1 in sync.trace:
        lines.append(f"  collective           {base.elapsed_time(c0) * 1e3:9.1f} +{c0.elapsed_time(c1) * 1e3:7.1f}   (relative to base event)")
    print("\n".join(lines), flush=True)
dist.barrier()
dist.destroy_process_group()
