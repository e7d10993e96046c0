"""Debug helper (GPU): tensor-path forward vs generic-path forward on one case; prints an error map."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "modulated-deform-conv_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch
import MDCONV_CUDA
import modulated_deform_conv as mdc
from cases import Case, make_inputs

def run(case, dtype, zero_off=False):
    t = make_inputs(case, dtype)
    d = {k: (v.cuda() if v is not None else None) for k, v in t.items()}
    if zero_off:
        d["offset"].zero_()
        if d["mask"] is not None: d["mask"].fill_(1)
    fn = {(2, False): mdc.deform_conv2d, (2, True): mdc.modulated_deform_conv2d,
          (3, False): mdc.deform_conv3d, (3, True): mdc.modulated_deform_conv3d}[(case.nd, case.modulated)]
    args = [d["x"], d["offset"]] + ([d["mask"]] if case.modulated else []) + [d["weight"], d["bias"], case.s, case.p, case.d, case.g, case.dg, 64]
    outs = {}
    for path in (1, 2):
        MDCONV_CUDA.set_path(path)
        outs[path] = fn(*args).float()
        torch.cuda.synchronize()
    MDCONV_CUDA.set_path(0)
    ref, out = outs[1], outs[2]
    err = (out - ref).abs()
    print(f"{case.name} {dtype} zero_off={zero_off}: max|ref|={ref.abs().max():.4f} max err={err.max():.5f} "
          f"rel={err.max()/ref.abs().max():.2e} nan={torch.isnan(out).sum().item()}")
    if err.max() / ref.abs().max() > 2e-2:
        B, Co = out.shape[:2]
        e2 = err.reshape(B, Co, -1)
        print("  per-sample max err:", [round(float(v), 3) for v in e2.amax(dim=(1, 2))])
        print("  per-16co max err:", [round(float(v), 3) for v in e2.reshape(B, -1, 16, e2.shape[-1]).amax(dim=(0, 2, 3))])
        pe = e2.amax(dim=(0, 1))
        print("  per-pixel(first 40) max err:", [round(float(v), 2) for v in pe[:40]])
        print("  out[0,0,:8]=", out.reshape(B, Co, -1)[0, 0, :8].tolist())
        print("  ref[0,0,:8]=", ref.reshape(B, Co, -1)[0, 0, :8].tolist())

if __name__ == "__main__":
    c = Case("dbg_c64", 2, True, B=2, C=64, Co=64, size=(12, 11), seed=50)
    run(c, torch.float16, zero_off=True)
    run(c, torch.float16)
    run(c, torch.bfloat16)
    run(Case("dbg_c128_co256", 2, True, B=3, C=128, Co=256, size=(9, 14), bias=False, seed=51), torch.float16)
    run(Case("dbg_big", 2, True, B=4, C=256, Co=256, size=(56, 56), bias=False, seed=52), torch.float16)
    run(Case("dbg_3d", 3, True, B=1, C=64, Co=64, size=(4, 6, 5), k=(3, 3, 3), s=(1, 1, 1), p=(1, 1, 1), d=(1, 1, 1), seed=70), torch.float16)
