"""GPU tests (-m gpu) that CALL every one of the eight reference-named entry points -- the Python functions of
MDCONV_CUDA with the reference's positional signatures, and the eight raw C symbols of libmdconv_b200.so through
ctypes -- against the CPU oracle.

Signatures followed (reference file:line):
  deform_conv2d_forward_cuda / _backward_cuda              src/deformable_conv.cu:117-123, 327-333
  modulated_deform_conv2d_forward_cuda / _backward_cuda    src/mdeformable_conv.cu:120-126, 361-366
  deform_conv3d_forward_cuda / _backward_cuda              src/deformable_conv3d.cu:160-167, 434-442
  modulated_deform_conv3d_forward_cuda / _backward_cuda    src/mdeformable_conv3d.cu:170-177, 443-451
  call sites                                               modulated_deform_conv.py:28-34,57-64,194-200,225-232,281-287,313-320

Out-param contract checked here: the callee owns the caller's buffers for the duration of the call, overwrites every
element and returns 0 -- whatever they held before.  (The reference zeroes `output` itself, src/deformable_conv.cu:155,
and ACCUMULATES into the gradient buffers, which its Python callers always pre-zero, modulated_deform_conv.py:53-56;
pre-zeroed and garbage-filled buffers are both exercised, the results must be identical.)
"""
import ctypes

import numpy as np
import pytest
import torch

import oracle
from cases import ALL_CASES, TENSOR_CASES, make_inputs, rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-4, torch.float16: 1e-2}

_BY_NAME = {c.name: c for c in ALL_CASES + TENSOR_CASES}
# (kind, small generic-path case, tensor-path case)
KINDS = {
    "d2d": ("d2d_base", "tc_d2d_c64"),
    "m2d": ("m2d_base", "tc_m2d_g2"),
    "d3d": ("d3d_base", "tc_d3d_c64_half_grid"),
    "m3d": ("m3d_base", "tc_m3d_c64"),
}
PARAMS = [(k, n, dt) for k, (small, tc) in KINDS.items() for n, dt in ((small, torch.float32), (tc, torch.float16))]


def _oracle(case, t):
    a = (case.s, case.p, case.d, case.g, case.dg)
    ref = oracle.backward(t["x"], t["weight"], t["offset"], t["mask"], t["grad_output"], *a, variant=case.rule,
                          with_bias=case.bias)
    ref["y"] = oracle.forward(t["x"], t["weight"], t["bias"], t["offset"], t["mask"], *a)
    return {k: v for k, v in ref.items() if v is not None}


def _garbage(shape, dtype, fill):
    """A caller-owned out-param with junk in it: NaN (never overwritten -> the comparison fails loudly) or 0."""
    t = torch.empty(shape, dtype=dtype, device="cuda")
    t.fill_(fill)
    return t


def _check(got, ref, tol, what):
    bad = {}
    for k, r in ref.items():
        e = rel_err(got[k].double().cpu().numpy(), r)
        if not e <= tol:
            bad[k] = e
    assert not bad, f"{what}: rel err over {tol}: {bad}"


def _dev(case, dtype):
    t = make_inputs(case, dtype)
    d = {k: (v.cuda() if v is not None else None) for k, v in t.items()}
    if d["bias"] is None:
        d["bias"] = d["x"].new_empty(0)      # the reference's "no bias" convention (modulated_deform_conv.py:20-21)
    return t, d


@pytest.mark.parametrize("fill", [0.0, float("nan")], ids=["zeroed", "garbage"])
@pytest.mark.parametrize("kind,name,dtype", PARAMS, ids=[f"{k}-{n}-{str(dt)[6:]}" for k, n, dt in PARAMS])
def test_python_entry_points_positional(kind, name, dtype, fill):
    """MDCONV_CUDA.<op>_forward_cuda / _backward_cuda with the reference's positional argument lists."""
    import MDCONV_CUDA as M
    case = _BY_NAME[name]
    t, d = _dev(case, dtype)
    geo = (*case.k, *case.s, *case.p, *case.d, case.g, case.dg, 64, case.bias)
    oshape = (case.B, case.Co, *case.out_size)
    got = {}
    if kind == "m2d":
        # return-value style (the only pair registered on the reference's HEAD)
        y = M.modulated_deform_conv2d_forward_cuda(d["x"], d["weight"], d["bias"], d["offset"], d["mask"], *geo)
        assert tuple(y.shape) == oshape
        g5 = M.modulated_deform_conv2d_backward_cuda(d["x"], d["weight"], d["bias"], d["offset"], d["mask"],
                                                     d["grad_output"], *geo)
        assert isinstance(g5, tuple) and len(g5) == 5
        got = {"y": y, "grad_input": g5[0], "grad_offset": g5[1], "grad_mask": g5[2], "grad_weight": g5[3],
               "grad_bias": g5[4]}
    else:
        y = _garbage(oshape, dtype, fill)
        gi, gw, goff = (_garbage(d[k].shape, dtype, fill) for k in ("x", "weight", "offset"))
        gb = _garbage(d["bias"].shape, dtype, fill)
        if kind == "d2d":
            rc = M.deform_conv2d_forward_cuda(d["x"], d["weight"], d["bias"], d["offset"], y, *geo)
            rb = M.deform_conv2d_backward_cuda(d["x"], d["weight"], d["bias"], d["offset"], gi, gw, gb, goff,
                                               d["grad_output"], *geo)
        elif kind == "d3d":
            rc = M.deform_conv3d_forward_cuda(d["x"], d["weight"], d["bias"], d["offset"], y, *geo)
            rb = M.deform_conv3d_backward_cuda(d["x"], d["weight"], d["bias"], d["offset"], gi, gw, gb, goff,
                                               d["grad_output"], *geo)
        else:
            gm = _garbage(d["mask"].shape, dtype, fill)
            rc = M.modulated_deform_conv3d_forward_cuda(d["x"], d["weight"], d["bias"], d["offset"], d["mask"], y, *geo)
            rb = M.modulated_deform_conv3d_backward_cuda(d["x"], d["weight"], d["bias"], d["offset"], d["mask"], gi,
                                                         gw, gb, goff, gm, d["grad_output"], *geo)
            got["grad_mask"] = gm
        assert rc == 0 and rb == 0
        got.update({"y": y, "grad_input": gi, "grad_weight": gw, "grad_offset": goff, "grad_bias": gb})
    torch.cuda.synchronize()
    ref = _oracle(case, t)
    if not case.bias:
        got.pop("grad_bias", None)
    _check(got, ref, TOL[dtype], f"{kind}/{name}/{dtype}/python")


def _cp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() else None


@pytest.mark.parametrize("fill", [0.0, float("nan")], ids=["zeroed", "garbage"])
@pytest.mark.parametrize("kind,name,dtype", PARAMS, ids=[f"{k}-{n}-{str(dt)[6:]}" for k, n, dt in PARAMS])
def test_raw_c_symbols(kind, name, dtype, fill):
    """The eight C symbols of include/mdconv_b200.h, called with raw device pointers (what a cgo / JNI / ctypes
    binder sees): pointer order as in the reference's argument lists, then sizes, geometry, dtype, workspace, stream."""
    import MDCONV_CUDA as M
    lib = M._lib
    case = _BY_NAME[name]
    t, d = _dev(case, dtype)
    nd, mod = case.nd, case.modulated
    dt = M._DTYPES[dtype]
    sizes = [case.B, case.C, *case.size, case.Co]
    geo = sizes + [*case.k, *case.s, *case.p, *case.d, case.g, case.dg, 64, 1 if case.bias else 0]
    pad3 = lambda v, fillv: list(v) + [fillv] * (3 - len(v))
    wsargs = [nd, 1 if mod else 0, 0, case.B, case.C, *pad3(case.size, 1), case.Co, *pad3(case.k, 1), *pad3(case.s, 1),
              *pad3(case.p, 0), *pad3(case.d, 1), case.g, case.dg, 1 if case.bias else 0, dt]

    def workspace(backward):
        a = list(wsargs)
        a[2] = backward
        n = int(lib.mdconv_legacy_workspace_bytes(*a))
        ws = torch.empty(max(n, 1), dtype=torch.uint8, device="cuda")
        return ws, n

    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    oshape = (case.B, case.Co, *case.out_size)
    y = _garbage(oshape, dtype, fill)
    gi, gw, goff = (_garbage(d[k].shape, dtype, fill) for k in ("x", "weight", "offset"))
    gb = _garbage(d["bias"].shape, dtype, fill)
    gm = _garbage(d["mask"].shape, dtype, fill) if mod else None
    x, w, b, off, m, go = (d[k] for k in ("x", "weight", "bias", "offset", "mask", "grad_output"))
    fwd = getattr(lib, {"d2d": "deform_conv2d_forward_cuda", "m2d": "modulated_deform_conv2d_forward_cuda",
                        "d3d": "deform_conv3d_forward_cuda", "m3d": "modulated_deform_conv3d_forward_cuda"}[kind])
    bwd = getattr(lib, {"d2d": "deform_conv2d_backward_cuda", "m2d": "modulated_deform_conv2d_backward_cuda",
                        "d3d": "deform_conv3d_backward_cuda", "m3d": "modulated_deform_conv3d_backward_cuda"}[kind])
    ws, n = workspace(0)
    if mod:
        rc = fwd(_cp(x), _cp(w), _cp(b), _cp(off), _cp(m), _cp(y), *geo, dt, _cp(ws) if n else None, n, stream)
    else:
        rc = fwd(_cp(x), _cp(w), _cp(b), _cp(off), _cp(y), *geo, dt, _cp(ws) if n else None, n, stream)
    assert rc == 0, M._lib.mdconv_status_string(rc)
    ws, n = workspace(1)
    wsp = _cp(ws) if n else None
    if kind == "m2d":     # src/mdeformable_conv.cu:361-366 + the returned tuple order (:456-457)
        rc = bwd(_cp(x), _cp(w), _cp(b), _cp(off), _cp(m), _cp(go), _cp(gi), _cp(goff), _cp(gm), _cp(gw), _cp(gb),
                 *geo, dt, wsp, n, stream)
    elif kind == "m3d":   # src/mdeformable_conv3d.cu:443-451
        rc = bwd(_cp(x), _cp(w), _cp(b), _cp(off), _cp(m), _cp(gi), _cp(gw), _cp(gb), _cp(goff), _cp(gm), _cp(go),
                 *geo, dt, wsp, n, stream)
    else:                 # src/deformable_conv.cu:327-333, src/deformable_conv3d.cu:434-442
        rc = bwd(_cp(x), _cp(w), _cp(b), _cp(off), _cp(gi), _cp(gw), _cp(gb), _cp(goff), _cp(go),
                 *geo, dt, wsp, n, stream)
    assert rc == 0, M._lib.mdconv_status_string(rc)
    torch.cuda.synchronize()
    got = {"y": y, "grad_input": gi, "grad_weight": gw, "grad_offset": goff}
    if mod:
        got["grad_mask"] = gm
    if case.bias:
        got["grad_bias"] = gb
    _check(got, _oracle(case, t), TOL[dtype], f"{kind}/{name}/{dtype}/C")


def test_python_and_c_entry_points_agree_bitwise():
    """Same tensors through the Python function and the raw C symbol of the same name: identical bits on the
    deterministic tensor path (the two bindings reach the same kernels)."""
    import MDCONV_CUDA as M
    case = _BY_NAME["tc_m3d_c64"]
    t, d = _dev(case, torch.float16)
    geo = (*case.k, *case.s, *case.p, *case.d, case.g, case.dg, 64, case.bias)
    y1 = torch.empty(case.B, case.Co, *case.out_size, dtype=torch.float16, device="cuda")
    M.modulated_deform_conv3d_forward_cuda(d["x"], d["weight"], d["bias"], d["offset"], d["mask"], y1, *geo)
    y2 = torch.full_like(y1, float("nan"))
    cgeo = [case.B, case.C, *case.size, case.Co, *case.k, *case.s, *case.p, *case.d, case.g, case.dg, 64, 1]
    n = int(M._lib.mdconv_legacy_workspace_bytes(3, 1, 0, case.B, case.C, *case.size, case.Co, *case.k, *case.s,
                                                 *case.p, *case.d, case.g, case.dg, 1, 1))
    ws = torch.empty(max(n, 1), dtype=torch.uint8, device="cuda")
    rc = M._lib.modulated_deform_conv3d_forward_cuda(_cp(d["x"]), _cp(d["weight"]), _cp(d["bias"]), _cp(d["offset"]),
                                                     _cp(d["mask"]), _cp(y2), *cgeo, 1, _cp(ws) if n else None, n,
                                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)


# ---- full-size configs, element-wise, against the live reference kernels ------------------------------------
def _margins_path():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, "r2_parity_margins.json")


def _record_margins(name, margins):
    import json
    import os
    path = _margins_path()
    data = {}
    if os.path.exists(path):
        try:
            data = json.load(open(path))
        except Exception:
            data = {}
    data[name] = margins
    with open(path, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


def _err_pair(a, r):
    """(max|a-r| / max|r|, ||a-r||_2 / ||r||_2) on the device, in fp64."""
    a, r = a.double(), r.double()
    den_m, den_2 = r.abs().max(), r.norm()
    diff = a - r
    return (float(diff.abs().max() / den_m) if float(den_m) > 0 else float(diff.abs().max()),
            float(diff.norm() / den_2) if float(den_2) > 0 else float(diff.norm()))


FULL = {
    # BASELINE.json configs[1], [3] at their full batch, and configs[4]'s global batch un-sharded on one GPU
    "cfg2_n32_f16": dict(nd=2, B=32, C=256, Co=256, size=(56, 56), dtype=torch.float16, seed=11),
    "cfg4_n8_f16": dict(nd=3, B=8, C=64, Co=128, size=(16, 32, 32), dtype=torch.float16, seed=12),
    "cfg5_n256_one_gpu_f16": dict(nd=2, B=256, C=256, Co=256, size=(56, 56), dtype=torch.float16, seed=13),
    "cfg2_n32_bf16": dict(nd=2, B=32, C=256, Co=256, size=(56, 56), dtype=torch.bfloat16, seed=14),
}


@pytest.mark.parametrize("name", sorted(FULL))
def test_full_size_config_vs_live_reference_elementwise(name):
    """Every element of the output and of all gradients at the benchmark's exact shapes, our fp16/bf16 tensor path
    against the UNMODIFIED reference fp32 kernels (oracle/_ref) on the same 16-bit-valued tensors, same GPU.
    Bar: BASELINE.json's 1e-2 rel (fp16; 3e-2 bf16) as max|a-b|/max|b|; the achieved max-rel and relative-L2 per
    tensor go to gpurun_out/r2_parity_margins.json (committed as profiles/r2_parity_margins.json)."""
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref not built on this box")
    import modulated_deform_conv as mdc
    cfg = FULL[name]
    nd, B, C, Co, size, dtype = (cfg[k] for k in ("nd", "B", "C", "Co", "size", "dtype"))
    K = 3 ** nd
    gen = torch.Generator(device="cuda").manual_seed(cfg["seed"])
    rn = lambda *s: torch.randn(*s, generator=gen, device="cuda")
    x = rn(B, C, *size).to(dtype)
    off = rn(B, nd * K, *size).to(dtype)
    msk = torch.sigmoid(rn(B, K, *size)).to(dtype)
    go = rn(B, Co, *size).to(dtype)
    w = ((torch.rand(Co, C, *([3] * nd), generator=gen, device="cuda") * 2 - 1) / (C * K) ** 0.5).to(dtype)
    one = (1,) * nd
    lv = [t.clone().requires_grad_() for t in (x, off, msk, w)]
    fn = mdc.modulated_deform_conv2d if nd == 2 else mdc.modulated_deform_conv3d
    y = fn(lv[0], lv[1], lv[2], lv[3], None, one, one, one, 1, 1)
    y.backward(go)
    ours = {"y": y.detach(), "grad_input": lv[0].grad, "grad_offset": lv[1].grad, "grad_mask": lv[2].grad,
            "grad_weight": lv[3].grad}
    torch.backends.cuda.matmul.allow_tf32 = False
    xf, of, mf, wf, gf = (t.float() for t in (x, off, msk, w, go))
    ry = ref_runner.forward(xf, wf, None, of, mf, one, one, one, 1, 1, in_step=1 if nd == 3 else 64)
    rg = ref_runner.backward(xf, wf, None, of, mf, gf, one, one, one, 1, 1, in_step=64)
    ref = {"y": ry, "grad_input": rg["grad_input"], "grad_offset": rg["grad_offset"], "grad_mask": rg["grad_mask"],
           "grad_weight": rg["grad_weight"]}
    torch.cuda.synchronize()
    tol = {torch.float16: 1e-2, torch.bfloat16: 3e-2}[dtype]
    margins, bad = {}, {}
    for k in ref:
        mx, l2 = _err_pair(ours[k], ref[k])
        margins[k] = {"max_rel": mx, "rel_l2": l2}
        if not (mx <= tol and l2 <= tol):
            bad[k] = (mx, l2)
    margins["_tolerance"] = tol
    margins["_shape"] = {"B": B, "C": C, "Co": Co, "size": list(size), "dtype": str(dtype)}
    _record_margins(name, margins)
    assert not bad, f"{name}: over {tol}: {bad}"
