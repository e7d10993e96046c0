"""Generates tests/golden/ref_*.npz on a GPU box: forward + all gradients of the UNMODIFIED
reference kernels (oracle/_ref, built from /root/reference by oracle/build_ref.py) on the seeded
inputs of tests/cases.py, in fp32 and fp64.  These fixtures pin the CPU oracle (tests/test_oracle.py).

    gpurun -- python oracle/gen_ref_golden.py gpurun_out/golden      # then copy into tests/golden/

Forwards use in_step=1 (SURVEY B-1); backwards use the reference default in_step=64.

It also writes tests/golden/reftc_*.npz: the reference's fp32 kernels on fp16-VALUED inputs of shapes the
tcgen05 path accepts (inputs stored as fp16, reference outputs as fp32) -- the fixtures the fp16 tensor path
is checked against directly (1e-2 rel), without the oracle in between.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import ALL_CASES, TENSOR_CASES, make_inputs  # noqa: E402
from oracle import ref_runner  # noqa: E402

PICK = ["m2d_base", "m2d_stride2", "m2d_dil2_pad2", "m2d_k1x3_s12", "m2d_integer_grid", "m2d_half_grid",
        "m2d_nobias_b3", "d2d_base", "d2d_stride2_dg4", "d2d_integer_grid", "d2d_half_grid", "d2d_cfg1",
        "m3d_base", "m3d_nobias_k132", "m3d_integer_grid", "d3d_base", "d3d_stride2", "d3d_half_grid"]


TC_PICK = ["tc_m2d_c64", "tc_d2d_c64", "tc_m2d_g2", "tc_m3d_c64"]


def main_tc(out_dir):
    by_name = {c.name: c for c in TENSOR_CASES}
    for name in TC_PICK:
        c = by_name[name]
        t16 = make_inputs(c, torch.float16)
        d = {k: (v.float().cuda() if v is not None else None) for k, v in t16.items()}
        y = ref_runner.forward(d["x"], d["weight"], d["bias"], d["offset"], d["mask"], c.s, c.p, c.d, c.g, c.dg,
                               in_step=1)
        g = ref_runner.backward(d["x"], d["weight"], d["bias"], d["offset"], d["mask"], d["grad_output"], c.s,
                                c.p, c.d, c.g, c.dg, in_step=64)
        torch.cuda.synchronize()
        out = {k: v.numpy() for k, v in t16.items() if v is not None}          # fp16 inputs
        if t16["bias"] is None:
            out["bias"] = np.zeros((0,), dtype=np.float16)
        out["y"] = y.cpu().numpy()
        for k, v in g.items():
            if v is not None:
                out[k] = v.cpu().numpy()
        out.update(stride=np.array(c.s), padding=np.array(c.p), dilation=np.array(c.d), groups=np.array(c.g),
                   dg=np.array(c.dg), rule=np.array(c.rule))
        path = os.path.join(out_dir, f"reftc_{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path)


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    assert ref_runner.available(), "oracle/_ref/*.so missing: run oracle/build_ref.py in the build container"
    torch.backends.cuda.matmul.allow_tf32 = False
    by_name = {c.name: c for c in ALL_CASES}
    main_tc(out_dir)
    for name in PICK:
        if os.environ.get("MDCONV_GOLDEN_TC_ONLY") == "1":
            break
        c = by_name[name]
        for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            t = make_inputs(c, dtype)
            d = {k: (v.cuda() if v is not None else None) for k, v in t.items()}
            y = ref_runner.forward(d["x"], d["weight"], d["bias"], d["offset"], d["mask"], c.s, c.p, c.d, c.g, c.dg,
                                   in_step=1)
            g = ref_runner.backward(d["x"], d["weight"], d["bias"], d["offset"], d["mask"], d["grad_output"], c.s,
                                    c.p, c.d, c.g, c.dg, in_step=64)
            torch.cuda.synchronize()
            out = {k: v.numpy() for k, v in t.items() if v is not None}
            if t["bias"] is None:
                out["bias"] = np.zeros((0,), dtype=out["x"].dtype)
            out["y"] = y.cpu().numpy()
            for k, v in g.items():
                if v is not None:
                    out[k] = v.cpu().numpy()
            out.update(stride=np.array(c.s), padding=np.array(c.p), dilation=np.array(c.d), groups=np.array(c.g),
                       dg=np.array(c.dg), rule=np.array(c.rule))
            path = os.path.join(out_dir, f"ref_{name}_{tag}.npz")
            np.savez_compressed(path, **out)
            print("wrote", path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
