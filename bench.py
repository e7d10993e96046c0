#!/usr/bin/env python
"""bench.py -- headline benchmark of the deformable-conv hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4]

A step = one forward + backward of ModulatedDeformConv2d over one batch of synthetic input
(cfg2: N=32 per GPU, C=256->256, 56x56, k3 s1 p1 dg1, fp16 -- the configuration the metric is quoted
on).  N>1 is launched by torchrun, one rank per GPU: the batch is sharded per rank (weak scaling,
32 samples per GPU) and the only exchange is ONE NCCL all-reduce of [grad_weight || grad_bias].
Rank 0 prints ONE JSON line.

`--impl reference`: the reference publishes no CPU implementation, so this arm times the CPU
restatement of its algorithm named by BASELINE.json's north_star -- the pure-PyTorch deformable
unfold + matmul path (oracle/torch_unfold.py, all host threads) -- on a bounded sample of the same
workload (one sample per step).  The reference's own CUDA op on the same GPU is timed inside the
default run and reported as `reference_cuda`.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "modulated-deform-conv_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "ModulatedDeformConv2d fwd+bwd Msamples/s"

WORKLOADS = {
    # name: (nd, modulated, per-GPU batch, C, Co, size, k, s, p, d, g, dg, dtype)
    "cfg2": dict(nd=2, modulated=True, B=32, C=256, Co=256, size=(56, 56), k=(3, 3), s=(1, 1), p=(1, 1), d=(1, 1),
                 g=1, dg=1, dtype="f16",
                 desc="cfg2: ModulatedDeformConv2d N=32/GPU C=256->256 56x56 k3 s1 p1 dg=1 fp16"),
    "cfg2_bf16": dict(nd=2, modulated=True, B=32, C=256, Co=256, size=(56, 56), k=(3, 3), s=(1, 1), p=(1, 1),
                      d=(1, 1), g=1, dg=1, dtype="bf16",
                      desc="cfg2 shape in bf16: ModulatedDeformConv2d N=32/GPU C=256->256 56x56 k3 s1 p1 dg=1"),
    "cfg3": dict(nd=2, modulated=False, B=16, C=512, Co=512, size=(28, 28), k=(3, 3), s=(2, 2), p=(1, 1), d=(1, 1),
                 g=4, dg=4, dtype="f32", desc="cfg3: DeformConv2d N=16 C=512->512 28x28 k3 s2 p1 dg=4 g=4 fp32"),
    "cfg4": dict(nd=3, modulated=True, B=8, C=64, Co=128, size=(16, 32, 32), k=(3, 3, 3), s=(1, 1, 1), p=(1, 1, 1),
                 d=(1, 1, 1), g=1, dg=1, dtype="f16",
                 desc="cfg4: ModulatedDeformConv3d N=8 C=64->128 16x32x32 k3 s1 p1 fp16"),
}
DTYPES = {"f16": torch.float16, "f32": torch.float32, "bf16": torch.bfloat16, "f64": torch.float64}


def out_size(w):
    return tuple((i + 2 * p - (d * (k - 1) + 1)) // s + 1 for i, k, s, p, d in zip(w["size"], w["k"], w["s"], w["p"], w["d"]))


def make_host_inputs(w, batch, seed, dtype, pin):
    """Synthetic inputs per SURVEY 8(d): x~N(0,1), offset~N(0,1) px, mask=sigmoid(N(0,1)), grad_out~N(0,1)."""
    gen = torch.Generator().manual_seed(seed)
    K = 1
    for k in w["k"]:
        K *= k
    o = out_size(w)
    x = torch.randn(batch, w["C"], *w["size"], generator=gen)
    off = torch.randn(batch, w["dg"] * w["nd"] * K, *o, generator=gen)
    msk = torch.sigmoid(torch.randn(batch, w["dg"] * K, *o, generator=gen)) if w["modulated"] else None
    go = torch.randn(batch, w["Co"], *o, generator=gen)
    t = {"x": x, "offset": off, "mask": msk, "grad_output": go}
    t = {k: (None if v is None else v.to(dtype).contiguous()) for k, v in t.items()}
    if pin:
        t = {k: (None if v is None else v.pin_memory()) for k, v in t.items()}
    return t


def algorithmic(w, batch, es):
    """Compulsory HBM bytes and dense-contraction FLOPs of one fwd / bwd."""
    K = 1
    for k in w["k"]:
        K *= k
    P = 1
    for o in out_size(w):
        P *= o
    IN = 1
    for i in w["size"]:
        IN *= i
    nx = batch * w["C"] * IN
    noff = batch * w["dg"] * w["nd"] * K * P
    nm = batch * w["dg"] * K * P if w["modulated"] else 0
    nw = w["Co"] * (w["C"] // w["g"]) * K
    ny = batch * w["Co"] * P
    bytes_fwd = es * (nx + noff + nm + nw + ny)
    bytes_bwd = es * (2 * nx + 2 * noff + 2 * nm + 2 * nw + ny)
    flops_fwd = 2.0 * batch * P * w["Co"] * (w["C"] // w["g"]) * K
    return bytes_fwd, bytes_bwd, flops_fwd, 2.0 * flops_fwd


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index, period_s=0.1):
        # NVML queries take driver-level locks that kernel launches of EVERY process on the box also need: keep the
        # polling rate modest (the profiling recipe uses 200 ms) and sample from one process per box
        self.period_s = period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(self.period_s)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "tflops_burst": p["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_sustained": 1400.0, "tflops_burst": 1590.0,
            "source": "fallback (B200_PROFILING.md)"}


def pick_tensor_peak(peaks, timed_region_s, clocks):
    """Burst cuBLAS figure unless the timed region is a seconds-long run whose SM clock settled well below max
    (the conditions MEASURED_PEAKS.json's sustained figure was taken under)."""
    mhz, mx = clocks.get("sm_mhz"), clocks.get("sm_max_mhz")
    sustained = timed_region_s >= 2.0 and mhz is not None and mx and mhz < 0.9 * mx
    if sustained:
        return peaks["tflops_sustained"], "sustained bf16 cuBLAS (timed region %.1f s, SM clock %s of %s MHz)" % (
            timed_region_s, mhz, mx)
    return peaks["tflops_burst"], "burst bf16 cuBLAS (timed region %.3f s, SM clock %s of %s MHz)" % (
        timed_region_s, mhz, mx)


# ---------------------------------------------------------------------------------------------
def _cpu_inputs(w, batch):
    t = make_host_inputs(w, batch, 1234, torch.float32, pin=False)
    gen = torch.Generator().manual_seed(5)
    K = 1
    for k in w["k"]:
        K *= k
    wt = (torch.rand(w["Co"], w["C"] // w["g"], *w["k"], generator=gen) * 2 - 1) / (w["C"] * K) ** 0.5
    return t, wt


def cpu_oracle_rate(w, threads, reps, warmup, batch=1):
    """fwd+bwd of the CPU oracle (fp32 build, OpenMP) on `batch` samples of the workload per step
    -> (Msamples/s, s/step)."""
    import oracle
    os.environ["OMP_NUM_THREADS"] = str(threads)
    t, wt = _cpu_inputs(w, batch)
    a = (w["s"], w["p"], w["d"], w["g"], w["dg"])

    def step():
        oracle.forward(t["x"], wt, None, t["offset"], t["mask"], *a, precision="f32")
        oracle.backward(t["x"], wt, t["offset"], t["mask"], t["grad_output"], *a, precision="f32", with_bias=False)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    return batch / dt / 1e6, dt


def cpu_unfold_rate(w, threads, reps, warmup, batch=1):
    """fwd + autograd bwd of the pure-PyTorch deformable unfold + matmul path (oracle/torch_unfold.py, fp32,
    MKL GEMM on `threads` host threads) on `batch` samples per step -> (Msamples/s, s/step).  This is the
    "pure-CPU PyTorch unfold+matmul path" BASELINE.json's north_star names as the CPU comparison."""
    from oracle.torch_unfold import deform_conv_nd
    torch.set_num_threads(threads)
    t, wt = _cpu_inputs(w, batch)
    lv = [t[k].clone().requires_grad_() for k in ("x", "offset")]
    mk = t["mask"].clone().requires_grad_() if t["mask"] is not None else None
    wl = wt.clone().requires_grad_()

    def step():
        for v in lv + [wl] + ([mk] if mk is not None else []):
            v.grad = None
        y = deform_conv_nd(lv[0], lv[1], mk, wl, None, w["s"], w["p"], w["d"], w["g"], w["dg"])
        y.backward(t["grad_output"])
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    return batch / dt / 1e6, dt


def cpu_baseline_leg(w, budget_s=10.0):
    """The CPU numbers reported beside the GPU one.  Headline `value`: the pure-PyTorch deformable
    unfold + matmul path (the fastest CPU path we have; the reference itself has none), sized to ~`budget_s`
    seconds of work.  `also`: the C+OpenMP restatement of the reference kernels' algorithm and torchvision's CPU
    deform_conv2d, on smaller samples."""
    threads = os.cpu_count() or 1
    _, dt1 = cpu_unfold_rate(w, threads, 1, 1, batch=1)             # calibration
    batch = 2
    reps = max(2, min(20, int(budget_s / max(2 * dt1, 1e-3))))
    v, dt = cpu_unfold_rate(w, threads, reps, 0, batch=batch)
    base = {"value": v, "unit": "Msamples/s", "cores": threads, "kind": "port",
            "sample": f"{batch} samples of the workload shape per step, fwd + autograd bwd, {reps} reps ({dt:.2f} s each), "
                      f"fp32 PyTorch deformable unfold+matmul (oracle/torch_unfold.py); the reference has no CPU path"}
    also = {}
    try:
        vo, dto = cpu_oracle_rate(w, threads, 2, 1, batch=1)
        also["c_openmp_restatement"] = {"value": vo, "unit": "Msamples/s", "cores": threads,
                                        "sample": f"1 sample per step, fwd+bwd, 2 reps ({dto:.2f} s each), oracle/mdconv_oracle.c fp32"}
    except Exception as e:
        also["c_openmp_restatement"] = {"unavailable": f"{type(e).__name__}: {e}"}
    if w["nd"] == 2:
        try:
            import torchvision
            t, wt = _cpu_inputs(w, 1)
            lv = [t[k].clone().requires_grad_() for k in ("x", "offset")]
            mk = t["mask"].clone().requires_grad_() if t["mask"] is not None else None
            wl = wt.clone().requires_grad_()

            def step_tv():
                y = torchvision.ops.deform_conv2d(lv[0], lv[1], wl, None, stride=w["s"], padding=w["p"],
                                                  dilation=w["d"], mask=mk)
                y.backward(t["grad_output"])
            step_tv()
            t0 = time.perf_counter()
            step_tv()
            dtv = time.perf_counter() - t0
            also["torchvision_cpu"] = {"value": 1.0 / dtv / 1e6, "unit": "Msamples/s", "cores": threads,
                                       "sample": f"1 sample, torchvision.ops.deform_conv2d CPU fwd+bwd, fp32 ({dtv:.2f} s)"}
        except Exception as e:
            also["torchvision_cpu"] = {"unavailable": f"{type(e).__name__}: {e}"}
    base["also"] = also
    return base


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    threads = os.cpu_count() or 1
    steps = max(1, args.steps)
    warm = max(1, min(args.warmup, 2))
    val, dt = cpu_unfold_rate(w, threads, steps, warm, batch=1)
    sample = ("1 sample of the workload shape per step: fwd + autograd bwd of the pure-PyTorch deformable "
              "unfold+matmul path (oracle/torch_unfold.py), fp32, all host threads")
    line = {"metric": METRIC, "value": val, "unit": "Msamples/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"],
                       "note": "the reference ships CUDA kernels only (no CPU path): this arm times the CPU unfold+matmul "
                               "restatement of its algorithm; the reference's own CUDA op on the same GPU is in the "
                               "main line's `reference_cuda`"},
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


# ---------------------------------------------------------------------------------------------
def _emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else this process (or a library under it, e.g. NCCL's
    version banner) writes to fd 1 has been routed to stderr by _quiet_stdout()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


_REAL_STDOUT = None


def _quiet_stdout() -> None:
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 generic SIMT, 2 tcgen05")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's per-GPU batch on every rank (default); strong: a FIXED global batch of "
                         "--global-batch samples split over the ranks (SURVEY 8(e): cfg5, N=256)")
    ap.add_argument("--global-batch", type=int, default=256)
    ap.add_argument("--comm-ctas", type=int, default=4,
                    help="N>1: CTAs of the dedicated NCCL communicator for the gradient all-reduce = SMs the "
                         "data-gradient kernels leave free (0: default communicator, library default of 8 SMs)")
    ap.add_argument("--channels-last", action="store_true",
                    help="x and grad_output in torch.channels_last / channels_last_3d (SURVEY 8(f) rank 4): no re-layout "
                         "pass of x, TMA-store epilogues; NOT the BASELINE configuration (which is NCHW)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference_arm(args, w)
        return

    import torch.distributed as dist
    import MDCONV_CUDA
    import modulated_deform_conv as mdc
    import mdconv_dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    MDCONV_CUDA.set_path(args.path)
    W = max(3, args.warmup)
    K_steps = max(1, args.steps)
    dtype = DTYPES[w["dtype"]]
    es = torch.tensor([], dtype=dtype).element_size()
    B = w["B"]
    if args.scaling == "strong":
        if args.global_batch % world:
            raise SystemExit("--global-batch must be a multiple of the number of ranks")
        B = args.global_batch // world

    cls = {(2, True): mdc.ModulatedDeformConv2d, (2, False): mdc.DeformConv2d,
           (3, True): mdc.ModulatedDeformConv3d, (3, False): mdc.DeformConv3d}[(w["nd"], w["modulated"])]
    torch.manual_seed(0)
    mod = cls(w["C"], w["Co"], w["k"], stride=w["s"], padding=w["p"], dilation=w["d"], groups=w["g"],
              deformable_groups=w["dg"], bias=False).to(dev).to(dtype)
    # data-parallel: the module all-reduces [grad_weight, grad_bias] inside its own backward, on a communication
    # stream that waits only for the weight-gradient kernels (overlaps grad_offset / grad_mask / grad_input)
    grad_sync = None
    if world > 1:
        grad_sync = mdconv_dist.attach(mod, None, max_ctas=args.comm_ctas if args.comm_ctas > 0 else None)

    host = make_host_inputs(w, B, 100 + rank, dtype, pin=True)
    d = {k: (None if v is None else v.to(dev)) for k, v in host.items()}
    if args.channels_last:
        fmt = torch.channels_last if w["nd"] == 2 else torch.channels_last_3d
        d["x"] = d["x"].contiguous(memory_format=fmt)
        d["grad_output"] = d["grad_output"].contiguous(memory_format=fmt)
    leaves = {k: d[k].clone().requires_grad_() for k in ("x", "offset") + (("mask",) if w["modulated"] else ())}

    launches = [0]

    def fwd_bwd(x, off, msk, go):
        args_ = (x, off, msk) if w["modulated"] else (x, off)
        y = mod(*args_)
        y.backward(go)
        return y

    def step_resident():
        for t in leaves.values():
            t.grad = None
        mod.weight.grad = None
        return fwd_bwd(leaves["x"], leaves["offset"], leaves.get("mask"), d["grad_output"])

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(W):
        step_resident()
    sync_all()

    # every rank watches its own GPU (a straggler GPU paces the whole data-parallel step), at a low rate for N > 1
    clocks = ClockSampler(local_rank, period_s=0.02 if world == 1 else 0.25)
    clocks.start()

    # ---- timed region 1: inputs resident in HBM ("value") ----
    launches[0] = MDCONV_CUDA.total_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    h0 = time.perf_counter()
    for _ in range(K_steps):
        step_resident()
    host_enqueue_ms = (time.perf_counter() - h0) * 1e3 / K_steps   # host time to ENQUEUE a step (no sync inside)
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    n_launch = MDCONV_CUDA.total_launch_count() - launches[0]
    per_rank_ms = None
    if world > 1:
        mine = torch.tensor([ms_total / K_steps, host_enqueue_ms], device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank_ms = {"device_ms_per_step": [round(float(t[0]), 4) for t in allr],
                       "host_enqueue_ms_per_step": [round(float(t[1]), 4) for t in allr]}
        tmax = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
    ms_per_step = ms_total / K_steps
    value = (B * world) / (ms_per_step * 1e-3) / 1e6

    # ---- timed region 2: end to end through the public Module API with HOST buffers ("e2e") ----
    keys = [k for k in ("x", "offset", "mask", "grad_output") if host[k] is not None]
    h2d = sum(host[k].numel() * es for k in keys)
    out_host = None

    # Pipelined over three streams with double-buffered device inputs: the H2D copy of step i+1 and the
    # D2H read-back of step i-1 overlap the kernels of step i (PCIe is full duplex).  Every step still
    # copies ITS inputs from pinned host memory and ITS results (output + all gradients) back to pinned
    # host memory inside the timed region; the region ends when the last read-back has landed.
    main = torch.cuda.current_stream()
    s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
    NBUF = 2
    dbuf = [{k: torch.empty_like(host[k], device=dev) for k in keys} for _ in range(NBUF)]
    ev_h2d = [torch.cuda.Event() for _ in range(NBUF)]
    ev_comp = [torch.cuda.Event() for _ in range(NBUF)]
    ev_d2h = [torch.cuda.Event() for _ in range(NBUF)]
    out_host = [None] * NBUF
    keep = [None] * NBUF
    d2h_bytes = [0]

    def issue_h2d(i):
        b = i % NBUF
        with torch.cuda.stream(s_h2d):
            s_h2d.wait_event(ev_comp[b])       # the previous user of this buffer has finished computing
            for k in keys:
                dbuf[b][k].copy_(host[k], non_blocking=True)
            ev_h2d[b].record(s_h2d)

    def issue_step(i):
        b = i % NBUF
        main.wait_event(ev_h2d[b])
        # the results of the step that used this slot last (two steps ago) have been read back: only now may their
        # device memory go back to the allocator (explicit events instead of Tensor.record_stream, whose deferred
        # frees make the caching allocator grow -- cudaMalloc inside the timed region -- on an unlucky interleaving)
        main.wait_event(ev_d2h[b])
        keep[b] = None
        lv = {k: dbuf[b][k].detach().requires_grad_() for k in keys if k != "grad_output"}
        mod.weight.grad = None
        y = fwd_bwd(lv["x"], lv["offset"], lv.get("mask"), dbuf[b]["grad_output"])
        res = [y.detach(), lv["x"].grad, lv["offset"].grad, mod.weight.grad] + ([lv["mask"].grad] if "mask" in lv else [])
        ev_comp[b].record(main)
        if out_host[b] is None:
            out_host[b] = [torch.empty(r.shape, dtype=r.dtype).pin_memory() for r in res]
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(ev_comp[b])
            for hbuf, r in zip(out_host[b], res):
                hbuf.copy_(r, non_blocking=True)
            ev_d2h[b].record(s_d2h)
        keep[b] = res
        d2h_bytes[0] = sum(r.numel() * r.element_size() for r in res)

    def run_e2e(n):
        issue_h2d(0)
        for i in range(n):
            if i + 1 < n:
                issue_h2d(i + 1)
            issue_step(i)
        main.wait_stream(s_d2h)

    run_e2e(3)
    sync_all()
    d2h = d2h_bytes[0]
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ke = max(3, K_steps)
    seg0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    e2.record()
    run_e2e(ke)
    e3.record()
    sync_all()
    e2e_device_allocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - seg0   # cudaMalloc calls inside the region
    ms_e2e = e2.elapsed_time(e3)
    if world > 1:
        tmax = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_e2e = float(tmax.item())
    e2e_value = (B * world) / (ms_e2e / ke * 1e-3) / 1e6
    clocks.stop()

    # copy-only probe: the same per-step H2D and D2H transfers with no kernels in between, both directions at once
    # (what the link of THIS box gives; a slow lease shows up here, not as a regression of the code)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe_n = 5
    sync_all()
    pe0.record()
    for i in range(probe_n):
        b_ = i % NBUF
        with torch.cuda.stream(s_h2d):
            for k in keys:
                dbuf[b_][k].copy_(host[k], non_blocking=True)
        with torch.cuda.stream(s_d2h):
            for hbuf, r in zip(out_host[b_], keep[b_]):
                hbuf.copy_(r, non_blocking=True)
    main.wait_stream(s_h2d)
    main.wait_stream(s_d2h)
    pe1.record()
    sync_all()
    ms_probe = pe0.elapsed_time(pe1) / probe_n
    copy_probe = {"ms_per_step": ms_probe, "h2d_gbs": h2d / (ms_probe * 1e-3) / 1e9,
                  "d2h_gbs": d2h / (ms_probe * 1e-3) / 1e9,
                  "note": "same pinned buffers, H2D and D2H streams running concurrently, no kernels"}

    # ---- N > 1: the in-backward all-reduce must equal an explicit all-reduce of the local gradients ----
    allreduce_check = None
    comm_trace = None
    if world > 1:
        sync_obj = getattr(mod, "_grad_sync", None)
        # explicit path: this rank's un-rounded fp32 parameter gradients straight from the C ABI, all-reduced on the
        # DEFAULT communicator; the in-backward path must give the same fp32 values (before the single rounding)
        wk = tuple(mod.weight.shape[2:])
        gl = MDCONV_CUDA.backward_ex(d["x"], mod.weight.detach(), d["x"].new_empty(0), d["offset"], d.get("mask"),
                                     d["grad_output"], wk, w["s"], w["p"], w["d"], w["g"], w["dg"], False, want_f32=True)
        if "flat_f32" in gl:
            expect = gl["flat_f32"].double()
        else:
            expect = gl["grad_weight"].double().reshape(-1)
        dist.all_reduce(expect, op=dist.ReduceOp.SUM)
        sync_obj.keep_last = True
        step_resident()
        torch.cuda.synchronize()
        got32 = sync_obj.last_flat32
        got = got32.double() if got32 is not None else mod.weight.grad.detach().double().reshape(-1)
        sync_obj.keep_last = False
        sync_obj.last_flat32 = None
        err = float((got - expect).abs().max() / expect.abs().max())
        rounded = float((mod.weight.grad.detach().double().reshape(-1) - expect).abs().max() / expect.abs().max())
        allreduce_check = {"rel_err_vs_explicit_allreduce": err, "ok": bool(err < 2e-4),
                           "rel_err_after_rounding_to_storage_dtype": rounded,
                           "reduced_in": "fp32 (un-rounded parameter gradients), rounded once after the exchange"}
        if not allreduce_check["ok"]:
            raise SystemExit(f"in-backward all-reduce differs from the explicit one: {err}")
        # where the collective sits inside a step: time from the weight-ready event to the end of the collective,
        # next to the whole backward (events on the communication stream; outside the timed regions)
        sync_obj.trace = []
        tb0, tb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ntr = 5
        sync_all()
        tb0.record()
        for _ in range(ntr):
            step_resident()
        tb1.record()
        sync_all()
        coll = [a.elapsed_time(b) for a, b in sync_obj.trace]
        sync_obj.trace = None
        nb = mod.weight.numel() * 4
        comm_trace = {"collective_ms_mean": sum(coll) / max(1, len(coll)), "collective_ms_max": max(coll) if coll else None,
                      "step_ms": tb0.elapsed_time(tb1) / ntr, "allreduce_bytes_per_rank": nb,
                      "nvlink_bytes_per_step_per_rank": 2 * nb * (world - 1) / world,
                      "comm_ctas": args.comm_ctas, "note": "collective = fp32 all-reduce + the rounding copy, on the "
                      "communication stream, measured from the moment the weight-ready event has fired"}

    # ---- per-kernel durations (CUDA events on the launching stream, inside the library) ----
    MDCONV_CUDA.profile_enable(True)
    kt = {}
    nprof = max(3, min(K_steps, 10))
    for _ in range(nprof):
        step_resident()
    for name, ms in MDCONV_CUDA.profile_read():
        kt.setdefault(name, []).append(ms)
    MDCONV_CUDA.profile_enable(False)
    kernels = {k: sum(v) / nprof for k, v in kt.items()}  # ms per step, summed over launches of that kernel
    calls = {k: len(v) // nprof for k, v in kt.items()}
    if world > 1:
        # per-rank view: the sum of this rank's kernel times and its SM clock (the slowest GPU paces every step)
        cs = clocks.summary()
        mine = torch.tensor([sum(kernels.values()), float(cs.get("sm_mhz") or 0)], device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank_ms["kernel_sum_ms_per_step_serialised"] = [round(float(t[0]), 4) for t in allr]
        per_rank_ms["sm_mhz_median"] = [int(t[1]) for t in allr]

    peaks = load_peaks()
    clk = clocks.summary()
    tpeak, tpeak_kind = pick_tensor_peak(peaks, ms_total * 1e-3, clk)
    bytes_fwd, bytes_bwd, flops_fwd, flops_bwd = algorithmic(w, B, es)
    # dominant kernel and the algorithmic work of ONE launch of it.  Each of the four GEMM kernels of the
    # tensor path (forward, grad_weight, grad_sample = grad_offset+grad_mask, grad_input) performs one
    # dense contraction of 2*N*P*Co*(C/g)*K FLOPs over the whole per-GPU batch per launch.
    gemm_kernels = {"tc_forward": flops_fwd, "tc_grad_weight": flops_fwd, "tc_grad_sample": flops_fwd,
                    "tc_grad_input": flops_fwd, "generic_forward": flops_fwd, "generic_bwd_data": flops_fwd,
                    "generic_bwd_weight": flops_fwd}
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # dram bytes per launch from `ncu --set full`
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload, {})
        except Exception:
            traffic = {}
    cand = {k: v for k, v in kernels.items() if k in gemm_kernels}
    dom = max(cand, key=cand.get) if cand else (max(kernels, key=kernels.get) if kernels else None)
    roofline = None
    roofline_kernels = {}
    tensor_path = any(k.startswith("tc_") for k in kernels)
    for k, v in cand.items():
        dur = v / max(1, calls[k]) * 1e-3
        roofline_kernels[k] = {"launch_ms": dur * 1e3, "tflops": gemm_kernels[k] / dur / 1e12,
                               "frac_of_tensor_peak": gemm_kernels[k] / dur / 1e12 / tpeak}
    if "col2im_gather" in kernels:
        # grad_input as a gather over grad_columns (HBM / L1-bound, no contraction): bytes it must move once --
        # grad_columns (B*P*K*C), the inverse-index entries (8 bytes per corner) and grad_input itself
        P_out = 1
        for o in out_size(w):
            P_out *= o
        n_in = 1
        for o in w["size"]:
            n_in *= o
        taps = 1
        for o in w["k"]:
            taps *= o
        gbytes = B * P_out * taps * w["C"] * es + 8 * B * w["dg"] * taps * P_out * (2 ** w["nd"]) + B * w["C"] * n_in * es
        dur = kernels["col2im_gather"] / max(1, calls["col2im_gather"]) * 1e-3
        roofline_kernels["col2im_gather"] = {"launch_ms": dur * 1e3, "bound": "hbm", "algorithmic_bytes": gbytes,
                                             "gbs": gbytes / dur / 1e9, "frac_of_hbm_peak": gbytes / dur / 1e9 / peaks["hbm_gbs"],
                                             "traffic": traffic.get("col2im_gather")}
    if dom is not None and tensor_path:
        dur = kernels[dom] / max(1, calls[dom]) * 1e-3
        ach = gemm_kernels.get(dom, flops_fwd) / dur / 1e12
        roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s",
                    "frac": ach / tpeak, "traffic": traffic.get(dom), "launch_ms": dur * 1e3,
                    "peak_source": peaks["source"] + ", " + tpeak_kind,
                    "share_of_step": kernels[dom] / max(1e-9, sum(kernels.values()))}
    elif dom is not None:
        # SIMT path (fp32 / fp64): no tensor cores; report the dominant kernel against the HBM roof with the
        # algorithmic bytes of its direction (it is FFMA-bound in practice: see DESIGN.md)
        dur = kernels[dom] / max(1, calls[dom]) * 1e-3
        nbytes = bytes_fwd if "forward" in dom else bytes_bwd
        ach = nbytes / dur / 1e9
        roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "traffic": traffic.get(dom), "launch_ms": dur * 1e3,
                    "peak_source": peaks["source"] + ", copy bandwidth",
                    "share_of_step": kernels[dom] / max(1e-9, sum(kernels.values())),
                    "note": "SIMT fp32 path: FFMA-bound, not HBM-bound"}
    hbm_ach = (bytes_fwd + bytes_bwd) / (ms_per_step * 1e-3) / 1e9
    roofline_hbm = {"scope": "whole fwd+bwd step, algorithmic bytes", "achieved": hbm_ach, "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": hbm_ach / peaks["hbm_gbs"], "frac_of_8TBs_spec": hbm_ach / 8000.0,
                    "algorithmic_bytes_fwd": bytes_fwd, "algorithmic_bytes_bwd": bytes_bwd}

    line = None
    if rank == 0:
        cpu_base = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu_base = cpu_baseline_leg(w)
            except Exception as e:  # the baseline is informative; never fail the bench for it
                cpu_base = {"value": None, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "port",
                            "sample": f"failed: {e}"}
        ref_cuda = None
        if world == 1 and not args.no_ref_cuda:
            ref_cuda = time_reference_cuda(w, d, mod, B)
        line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": K_steps, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": w["dtype"], "data": "synthetic",
                "config": {"workload": w["desc"] + (" [channels-last activations]" if args.channels_last else ""),
                           "global_batch": B * world, "per_gpu_batch": B,
                           "parallelism": f"batch-shard x{world}",
                           "l2": "working set (277 MB at cfg2) exceeds the 126 MB L2; no explicit flush",
                           "path": {0: "auto", 1: "generic", 2: "tensor"}[args.path],
                           "collective": "1 NCCL all-reduce(sum) of [grad_weight||grad_bias] per step, issued inside the backward on a "
                                         "side stream once the weight-gradient kernels finish" if world > 1 else "none"},
                "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / ke, "steps": ke, "cuda_mallocs_in_region": e2e_device_allocs,
                        "achieved_h2d_gbs": h2d / (ms_e2e / ke * 1e-3) / 1e9,
                        "achieved_d2h_gbs": d2h / (ms_e2e / ke * 1e-3) / 1e9,
                        "copy_only_probe": copy_probe,
                        "how": "public Module API; per step: H2D of x/offset/mask/grad_output from pinned host memory, "
                               "fwd+bwd, D2H of output and all gradients; copies of adjacent steps overlap the kernels "
                               "(3 streams, double-buffered); fill and drain included"},
                "gpu_launches": n_launch, "clocks": clk, "roofline": roofline, "roofline_kernels": roofline_kernels,
                "roofline_hbm": roofline_hbm,
                "kernels_ms_per_step": kernels, "cpu_baseline": cpu_base, "reference_cuda": ref_cuda,
                "allreduce_check": allreduce_check, "comm_trace": comm_trace,
                "host_enqueue_ms_per_step": host_enqueue_ms, "per_rank": per_rank_ms}
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def time_reference_cuda(w, d, mod, B):
    """The UNMODIFIED reference kernels (oracle/_ref) on the same GPU and tensors: reported beside our
    number, never part of it."""
    try:
        from oracle import ref_runner
        if not ref_runner.available():
            return {"unavailable": "oracle/_ref not built"}
        a = (w["s"], w["p"], w["d"], w["g"], w["dg"])
        wt = mod.weight.detach()

        def step():
            y = ref_runner.forward(d["x"], wt, None, d["offset"], d["mask"], *a, in_step=64)
            g = ref_runner.backward(d["x"], wt, None, d["offset"], d["mask"], d["grad_output"], *a, in_step=64)
            return y, g
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        ef, eb, ee = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        n = 5
        tf = tb = 0.0
        for _ in range(n):
            ef.record()
            ref_runner.forward(d["x"], wt, None, d["offset"], d["mask"], *a, in_step=64)
            eb.record()
            ref_runner.backward(d["x"], wt, None, d["offset"], d["mask"], d["grad_output"], *a, in_step=64)
            ee.record()
            torch.cuda.synchronize()
            tf += ef.elapsed_time(eb)
            tb += eb.elapsed_time(ee)
        tf, tb = tf / n, tb / n
        return {"value": B / ((tf + tb) * 1e-3) / 1e6, "unit": "Msamples/s", "fwd_ms": tf, "bwd_ms": tb,
                "in_step": 64, "note": "unmodified reference kernels + cuBLAS, same GPU, same tensors"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}


if __name__ == "__main__":
    main()
