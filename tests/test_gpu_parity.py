"""GPU parity tests proper (-m gpu): this repo's CUDA path, called through the drop-in Python
surface (modulated_deform_conv -> MDCONV_CUDA -> C ABI), against the CPU oracle on the same seeded
inputs.  Tolerances are BASELINE.json's: 1e-4 rel fp32, 1e-2 rel fp16 (rel = max|a-b| / max|b|),
for the forward and all five gradients.  fp64 (1e-10) and bf16 (3e-2) are extra."""
import glob
import math
import os

import numpy as np
import pytest
import torch

import oracle
from cases import ALL_CASES, TENSOR_CASES, Case, make_inputs, rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-10, torch.float32: 1e-4, torch.float16: 1e-2, torch.bfloat16: 3e-2}


def _run_ours(case, t, dev="cuda", fwd_path=None, bwd_path=None):
    import MDCONV_CUDA
    import modulated_deform_conv as mdc
    fn = {(2, False): mdc.deform_conv2d, (2, True): mdc.modulated_deform_conv2d,
          (3, False): mdc.deform_conv3d, (3, True): mdc.modulated_deform_conv3d}[(case.nd, case.modulated)]
    leaves = {k: (v.to(dev).requires_grad_() if v is not None and k != "grad_output" else v) for k, v in t.items()}
    args = [leaves["x"], leaves["offset"]] + ([leaves["mask"]] if case.modulated else []) + \
           [leaves["weight"], leaves["bias"], case.s, case.p, case.d, case.g, case.dg, 64]
    if fwd_path is not None:
        MDCONV_CUDA.set_path(fwd_path)
    y = fn(*args)
    if bwd_path is not None:
        MDCONV_CUDA.set_path(bwd_path)
    y.backward(t["grad_output"].to(dev))
    torch.cuda.synchronize()
    out = {"y": y.detach(), "grad_input": leaves["x"].grad, "grad_offset": leaves["offset"].grad,
           "grad_weight": leaves["weight"].grad}
    if case.modulated:
        out["grad_mask"] = leaves["mask"].grad
    if case.bias:
        out["grad_bias"] = leaves["bias"].grad
    return {k: v.float().cpu().numpy().astype(np.float64) if v.dtype != torch.float64 else v.cpu().numpy()
            for k, v in out.items()}


def _oracle(case, t):
    a = (case.s, case.p, case.d, case.g, case.dg)
    ref = oracle.backward(t["x"], t["weight"], t["offset"], t["mask"], t["grad_output"], *a, variant=case.rule,
                          with_bias=case.bias)
    ref["y"] = oracle.forward(t["x"], t["weight"], t["bias"], t["offset"], t["mask"], *a)
    return {k: v for k, v in ref.items() if v is not None}


def _compare(ours, ref, tol, what):
    bad = {}
    for k, r in ref.items():
        e = rel_err(ours[k], r)
        if not e <= tol:
            bad[k] = e
    assert not bad, f"{what}: rel err over {tol}: {bad}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.float64, torch.bfloat16],
                         ids=["f32", "f16", "f64", "bf16"])
@pytest.mark.parametrize("case", ALL_CASES, ids=lambda c: c.name)
def test_generic_path_matches_oracle(case, dtype):
    import MDCONV_CUDA
    MDCONV_CUDA.set_path(1)
    try:
        t = make_inputs(case, dtype)
        ours = _run_ours(case, t)
    finally:
        MDCONV_CUDA.set_path(0)
    _compare(ours, _oracle(case, t), TOL[dtype], f"{case.name}/{dtype}")


# ---- directly against the UNMODIFIED reference kernels (no oracle in between) ----------------------------
_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_REF_GOLDEN = sorted(glob.glob(os.path.join(_GOLDEN, "ref_*.npz")))
_REFTC_GOLDEN = sorted(glob.glob(os.path.join(_GOLDEN, "reftc_*.npz")))


def _run_ours_on_fixture(z, dtype, path):
    """Our CUDA path on the inputs stored in a golden fixture -> dict of float64 arrays."""
    import MDCONV_CUDA
    import modulated_deform_conv as mdc
    nd = z["x"].ndim - 2
    modulated = "mask" in z.files
    fn = {(2, False): mdc.deform_conv2d, (2, True): mdc.modulated_deform_conv2d,
          (3, False): mdc.deform_conv3d, (3, True): mdc.modulated_deform_conv3d}[(nd, modulated)]
    dev = lambda k: torch.from_numpy(z[k]).to(dtype).cuda()
    x, off, w = dev("x").requires_grad_(), dev("offset").requires_grad_(), dev("weight").requires_grad_()
    b = dev("bias").requires_grad_() if z["bias"].size else None
    m = dev("mask").requires_grad_() if modulated else None
    args = [x, off] + ([m] if modulated else []) + \
           [w, b, tuple(int(v) for v in z["stride"]), tuple(int(v) for v in z["padding"]),
            tuple(int(v) for v in z["dilation"]), int(z["groups"]), int(z["dg"]), 64]
    MDCONV_CUDA.set_path(path)
    try:
        y = fn(*args)
        if path == 2:
            MDCONV_CUDA.set_path(0)   # backward: the tensor path wherever it supports the shape, else generic
        y.backward(dev("grad_output"))
        torch.cuda.synchronize()
    finally:
        MDCONV_CUDA.set_path(0)
    out = {"y": y.detach(), "grad_input": x.grad, "grad_offset": off.grad, "grad_weight": w.grad}
    if modulated:
        out["grad_mask"] = m.grad
    if b is not None:
        out["grad_bias"] = b.grad
    return {k: v.double().cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("path", _REF_GOLDEN, ids=os.path.basename)
def test_generic_path_matches_reference_kernel_fixtures(path):
    """Our SIMT kernels vs the outputs the unmodified reference kernels produced on a B200 (fp32: 1e-4 rel as
    BASELINE.json asks; fp64: 1e-10), forward and every gradient."""
    z = np.load(path)
    dtype = torch.float32 if z["x"].dtype == np.float32 else torch.float64
    ours = _run_ours_on_fixture(z, dtype, 1)
    ref = {k: z[k] for k in ours}
    _compare(ours, ref, TOL[dtype], os.path.basename(path))


@pytest.mark.parametrize("path", _REFTC_GOLDEN, ids=os.path.basename)
def test_tensor_path_matches_reference_kernel_fixtures(path):
    """The fp16 tcgen05 path vs the reference's fp32 kernels run on the same fp16-valued inputs (1e-2 rel)."""
    z = np.load(path)
    ours = _run_ours_on_fixture(z, torch.float16, 2)
    ref = {k: z[k] for k in ours}
    _compare(ours, ref, TOL[torch.float16], os.path.basename(path))


def test_reference_fixtures_present():
    assert len(_REF_GOLDEN) >= 8, "tests/golden/ref_*.npz missing"


@pytest.mark.parametrize("name", ["m2d_base", "d2d_stride2_dg4", "m3d_base", "d3d_base"])
def test_live_reference_kernels_generic(name):
    """Same GPU, same tensors: our fp32 kernels next to the unmodified reference kernels (oracle/_ref, built by
    __graft_entry__.build() where the reference tree is mounted).  Skipped when that build is absent."""
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref not built on this box")
    case = {c.name: c for c in ALL_CASES}[name]
    t = make_inputs(case, torch.float32)
    d = {k: (v.cuda() if v is not None else None) for k, v in t.items()}
    torch.backends.cuda.matmul.allow_tf32 = False
    y = ref_runner.forward(d["x"], d["weight"], d["bias"], d["offset"], d["mask"], case.s, case.p, case.d, case.g,
                           case.dg, in_step=1)
    g = ref_runner.backward(d["x"], d["weight"], d["bias"], d["offset"], d["mask"], d["grad_output"], case.s, case.p,
                            case.d, case.g, case.dg, in_step=64)
    ref = {k: v.double().cpu().numpy() for k, v in g.items() if v is not None}
    ref["y"] = y.double().cpu().numpy()
    import MDCONV_CUDA
    MDCONV_CUDA.set_path(1)
    try:
        ours = _run_ours(case, t)
    finally:
        MDCONV_CUDA.set_path(0)
    _compare(ours, ref, TOL[torch.float32], f"live reference {name}")


def test_live_reference_kernels_tensor_path():
    """fp16 tcgen05 path vs the reference's own fp32 kernels on the same fp16-valued tensors, at a multi-tile
    shape (C=256, 24x20) no fixture holds.  Skipped when oracle/_ref is absent."""
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref not built on this box")
    case = {c.name: c for c in TENSOR_CASES}["tc_m2d_c256_multi_tile"]
    t = make_inputs(case, torch.float16)
    d = {k: (v.float().cuda() if v is not None else None) for k, v in t.items()}
    torch.backends.cuda.matmul.allow_tf32 = False
    y = ref_runner.forward(d["x"], d["weight"], d["bias"], d["offset"], d["mask"], case.s, case.p, case.d, case.g,
                           case.dg, in_step=1)
    g = ref_runner.backward(d["x"], d["weight"], d["bias"], d["offset"], d["mask"], d["grad_output"], case.s, case.p,
                            case.d, case.g, case.dg, in_step=64)
    ref = {k: v.double().cpu().numpy() for k, v in g.items() if v is not None}
    ref["y"] = y.double().cpu().numpy()
    import MDCONV_CUDA
    try:
        ours = _run_ours(case, t, fwd_path=2, bwd_path=2)
    finally:
        MDCONV_CUDA.set_path(0)
    _compare(ours, ref, TOL[torch.float16], "live reference tc_m2d_c256_multi_tile")


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
@pytest.mark.parametrize("case", TENSOR_CASES, ids=lambda c: c.name)
def test_tensor_path_matches_oracle(case, dtype):
    """tcgen05 path (forced; the test fails if the path refuses a shape it claims to support)."""
    import ctypes
    import MDCONV_CUDA
    t = make_inputs(case, dtype)
    # which directions does the tensor path take for this geometry?  (forward is mandatory here)
    g = MDCONV_CUDA._geom(t["x"], t["weight"], case.k, case.s, case.p, case.d, case.g, case.dg, case.modulated,
                          case.bias)
    g.path = 2
    assert MDCONV_CUDA._lib.mdconv_selected_path(ctypes.byref(g), 0) == 2, "tensor path refused the forward"
    bwd = 2 if MDCONV_CUDA._lib.mdconv_selected_path(ctypes.byref(g), 1) == 2 else 1
    try:
        ours = _run_ours(case, t, fwd_path=2, bwd_path=bwd)
    finally:
        MDCONV_CUDA.set_path(0)
    _compare(ours, _oracle(case, t), TOL[dtype], f"{case.name}/{dtype} (bwd path {bwd})")


def _data_grad_kernels(case, t, budget_mb=None):
    """Run one forced-tensor-path fwd+bwd with kernel profiling on -> (results, launched kernel names)."""
    import MDCONV_CUDA
    prev = None if budget_mb is None else MDCONV_CUDA.set_gcol_budget_mb(budget_mb)
    MDCONV_CUDA.profile_enable(True)
    try:
        ours = _run_ours(case, t, fwd_path=2, bwd_path=2)
        names = [n for n, _ in MDCONV_CUDA.profile_read()]
    finally:
        MDCONV_CUDA.profile_enable(False)
        MDCONV_CUDA.set_path(0)
        if prev is not None:
            MDCONV_CUDA.set_gcol_budget_mb(prev if prev >= 0 else 1024)
    return ours, names


_GCOL_CASES = ["tc_m2d_c64", "tc_m2d_stride2_dg2", "tc_m2d_g4_dg2", "tc_d2d_c192_dg3", "tc_m2d_c256_multi_tile",
               "tc_m2d_ragged", "tc_m2d_ragged_3chunks", "tc_d2d_c512_dg2", "tc_m2d_c1024_slabs", "tc_m3d_dg2_stride2", "tc_m2d_all_oob"]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
@pytest.mark.parametrize("name", _GCOL_CASES)
def test_grad_input_paths_col2im_gather_and_contraction(name, dtype):
    """16-bit grad_input has two routes over the same inverse index: the default col2im GATHER over the grad_columns
    k_tc_grad_sample streams out (reference src/mdeformable_conv.cu:282-293 + :396-419), and the third contraction
    (k_tc_grad_input, budget 0).  Both must hold the oracle's bar, and the chunked gather (a budget of 1 MiB: one
    sample per (grad_sample, gather) pair) must reproduce the one-pass gather bit for bit."""
    case = {c.name: c for c in TENSOR_CASES}[name]
    t = make_inputs(case, dtype)
    ref = _oracle(case, t)
    one, names = _data_grad_kernels(case, t, budget_mb=1024)
    assert "col2im_gather" in names and "tc_grad_input" not in names, names
    assert names.count("col2im_gather") == 1 and names.count("tc_grad_sample") == 1, names
    assert "transpose_weight" not in names, names          # no transposed filter bank, no channels-last grad_output
    _compare(one, ref, TOL[dtype], f"{name}/{dtype} col2im gather")
    chunked, names_c = _data_grad_kernels(case, t, budget_mb=1)
    per_sample = case.taps * case.C * 2 * int(np.prod(case.out_size))      # bytes of one sample's grad_columns
    step = min(case.B, max(1, 2 ** 20 // per_sample))
    chunks = -(-case.B // step)
    assert names_c.count("col2im_gather") == chunks == names_c.count("tc_grad_sample"), (names_c, chunks)
    for k in one:
        assert np.array_equal(one[k], chunked[k]), f"{name}: chunked gather changed {k}"
    old, names_o = _data_grad_kernels(case, t, budget_mb=0)
    assert "tc_grad_input" in names_o and "col2im_gather" not in names_o, names_o
    _compare(old, ref, TOL[dtype], f"{name}/{dtype} contraction")


@pytest.mark.parametrize("case", TENSOR_CASES, ids=lambda c: c.name)
def test_fp32_tensor_path_matches_oracle(case):
    """fp32 on the tensor cores (operands split into bf16 hi / lo, three products per K step, fp32 accumulation):
    forward, grad_weight and -- where the shape qualifies (output channels per group a multiple of 64) -- the data
    gradients on the tcgen05 kernels; BASELINE.json's fp32 bar, 1e-4 rel, against the fp64 oracle."""
    import ctypes
    import MDCONV_CUDA
    t = make_inputs(case, torch.float32)
    g = MDCONV_CUDA._geom(t["x"], t["weight"], case.k, case.s, case.p, case.d, case.g, case.dg, case.modulated,
                          case.bias)
    g.path = 2
    if MDCONV_CUDA._lib.mdconv_selected_path(ctypes.byref(g), 0) != 2:
        pytest.skip("shape not taken by the fp32 tensor path (output-channel tile)")
    MDCONV_CUDA.profile_enable(True)
    try:
        ours = _run_ours(case, t, fwd_path=2, bwd_path=2)
        names = {n for n, _ in MDCONV_CUDA.profile_read()}
    finally:
        MDCONV_CUDA.profile_enable(False)
        MDCONV_CUDA.set_path(0)
    assert {"tc_forward", "tc_grad_weight"} <= names, names
    assert {"tc_grad_sample", "tc_grad_input"} <= names or "generic_bwd_data" in names, names
    if (case.Co // case.g) % 64 == 0:
        assert "generic_bwd_data" not in names, names
    _compare(ours, _oracle(case, t), TOL[torch.float32], f"{case.name}/fp32 tensor path")


def test_fp32_tensor_cores_can_be_switched_off():
    import MDCONV_CUDA
    case = {c.name: c for c in TENSOR_CASES}["tc_m2d_c64"]
    t = make_inputs(case, torch.float32)
    prev = MDCONV_CUDA.set_fp32_tensor_cores(False)
    try:
        MDCONV_CUDA.profile_enable(True)
        ours = _run_ours(case, t)
        names = {n for n, _ in MDCONV_CUDA.profile_read()}
    finally:
        MDCONV_CUDA.profile_enable(False)
        MDCONV_CUDA.set_fp32_tensor_cores(prev)
    assert "generic_forward" in names and "tc_forward" not in names
    _compare(ours, _oracle(case, t), TOL[torch.float32], "fp32 SIMT")


def test_saved_layouts_hand_off_is_bit_identical_and_skips_two_passes():
    """Forward -> backward hand-off of the channels-last input / re-ordered weight (mdconv_backward_saved): same
    bits as the stateless backward, two launches fewer; the legacy-named stateless entry points stay equivalent."""
    import MDCONV_CUDA
    case = {c.name: c for c in TENSOR_CASES}["tc_m2d_c256_multi_tile"]
    t = make_inputs(case, torch.float16)
    res, launches = {}, {}
    for keep in (True, False):
        MDCONV_CUDA.set_keep_layouts(keep)
        try:
            n0 = MDCONV_CUDA.total_launch_count()
            res[keep] = _run_ours(case, t, fwd_path=2, bwd_path=2)
            launches[keep] = MDCONV_CUDA.total_launch_count() - n0
        finally:
            MDCONV_CUDA.set_keep_layouts(True)
            MDCONV_CUDA.set_path(0)
    for k in res[True]:
        assert np.array_equal(res[True][k], res[False][k]), k
    assert launches[False] - launches[True] == 2, launches
    # the reference-named stateless pair on the same tensors
    d = {k: (v.cuda() if v is not None else None) for k, v in t.items()}
    y = MDCONV_CUDA.modulated_deform_conv2d_forward_cuda(d["x"], d["weight"], d["x"].new_empty(0), d["offset"],
                                                         d["mask"], *case.k, *case.s, *case.p, *case.d, case.g,
                                                         case.dg, 64, False)
    g = MDCONV_CUDA.modulated_deform_conv2d_backward_cuda(d["x"], d["weight"], d["x"].new_empty(0), d["offset"],
                                                          d["mask"], d["grad_output"], *case.k, *case.s, *case.p,
                                                          *case.d, case.g, case.dg, 64, False)
    torch.cuda.synchronize()
    assert np.array_equal(y.double().cpu().numpy(), res[True]["y"])
    assert np.array_equal(g[0].double().cpu().numpy(), res[True]["grad_input"])
    assert np.array_equal(g[3].double().cpu().numpy(), res[True]["grad_weight"])


def test_dispatcher_ops_match_functions_and_compile():
    """mdconv_ops (torch.library): same bits as the autograd Functions, passes torch.library.opcheck, and a module
    using it compiles with fullgraph=True (aot_eager: fake tensors + the registered autograd formula)."""
    import mdconv_ops
    import modulated_deform_conv as mdc
    case = {c.name: c for c in TENSOR_CASES}["tc_m2d_c64"]
    t = make_inputs(case, torch.float16)
    d = {k: (v.cuda() if v is not None else None) for k, v in t.items()}
    geo = (list(case.s), list(case.p), list(case.d), case.g, case.dg)

    def run(fn):
        lv = {k: d[k].clone().requires_grad_() for k in ("x", "weight", "bias", "offset", "mask")}
        y = fn(lv)
        y.backward(d["grad_output"])
        return [y.detach()] + [lv[k].grad for k in ("x", "weight", "bias", "offset", "mask")]

    ref = run(lambda lv: mdc.modulated_deform_conv2d(lv["x"], lv["offset"], lv["mask"], lv["weight"], lv["bias"],
                                                     case.s, case.p, case.d, case.g, case.dg))
    op = run(lambda lv: mdconv_ops.deform_conv(lv["x"], lv["weight"], lv["bias"], lv["offset"], lv["mask"], *geo))
    assert torch.equal(ref[0], op[0])
    for a, b in zip(ref[1:], op[1:]):   # this shape's backward runs on the SIMT kernels (atomics: order-dependent bits)
        assert rel_err(b.float().cpu().numpy(), a.float().cpu().numpy()) < 2e-3
    compiled = torch.compile(lambda x, w, b, o, m: mdconv_ops.deform_conv(x, w, b, o, m, *geo) * 2.0,
                             backend="aot_eager", fullgraph=True)
    cp = run(lambda lv: compiled(lv["x"], lv["weight"], lv["bias"], lv["offset"], lv["mask"]) / 2.0)
    for a, b in zip(ref, cp):
        assert rel_err(b.float().cpu().numpy(), a.float().cpu().numpy()) < 2e-3
    # plain 3D op through the same entry (mask=None, 5-d input), small fp32 case on the SIMT kernels
    c3 = {c.name: c for c in ALL_CASES}["d3d_base"]
    t3 = make_inputs(c3, torch.float32)
    d3 = {k: (v.cuda() if v is not None else None) for k, v in t3.items()}
    y3 = mdconv_ops.deform_conv(d3["x"], d3["weight"], d3["bias"], d3["offset"], None, list(c3.s), list(c3.p),
                                list(c3.d), c3.g, c3.dg)
    assert rel_err(y3.double().cpu().numpy(), _oracle(c3, t3)["y"]) < 1e-4
    torch.library.opcheck(mdconv_ops.deform_conv, (d3["x"], d3["weight"], d3["bias"], d3["offset"], None, list(c3.s),
                                                    list(c3.p), list(c3.d), c3.g, c3.dg),
                          test_utils=("test_schema", "test_faketensor"))


def test_autocast_and_merged_pack_producers():
    """Under torch.autocast the op runs in the autocast dtype (fp32 offsets / weights are cast like a conv's);
    a modulated Pack's single merged offset+mask convolution equals the reference's two separate convs."""
    import modulated_deform_conv as mdc
    torch.manual_seed(3)
    m = mdc.ModulatedDeformConv2d(64, 64, 3, padding=1, bias=True).cuda()          # fp32 parameters
    x = torch.randn(2, 64, 9, 8, device="cuda")
    off = torch.randn(2, 18, 9, 8, device="cuda")
    msk = torch.sigmoid(torch.randn(2, 9, 9, 8, device="cuda"))
    y32 = m(x, off, msk)
    for dt in (torch.float16, torch.bfloat16):
        with torch.autocast("cuda", dtype=dt):
            y = m(x, off, msk)
        assert y.dtype == dt
        assert rel_err(y.detach().float().cpu().numpy(), y32.detach().cpu().numpy()) < TOL[dt]
    y.float().sum().backward()
    assert m.weight.grad is not None and m.weight.grad.dtype == torch.float32
    pack = mdc.ModulatedDeformConv2dPack(8, 8, 3, padding=1, deformable_groups=2).cuda()
    xs = torch.randn(2, 8, 7, 6, device="cuda")
    o, k = mdc._offset_and_mask(pack, xs)
    assert rel_err(o.detach().cpu().numpy(), pack.conv_offset(xs).detach().cpu().numpy()) < 1e-5
    assert rel_err(k.detach().cpu().numpy(), pack.conv_mask(xs).detach().cpu().numpy()) < 1e-5
    out = pack(xs)
    out.sum().backward()
    assert pack.conv_offset.weight.grad is not None and pack.conv_mask.bias.grad is not None
    assert out.shape == (2, 8, 7, 6)


def test_weight_ready_event_fires_before_the_data_gradients():
    """mdconv_backward_saved(weight_ready_event): the event is recorded once grad_weight / grad_bias are final; a
    stream that waits only for it must already see the finished parameter gradients, and the event completes
    before the whole backward does (it sits in front of the two data-gradient kernels)."""
    import MDCONV_CUDA
    case = {c.name: c for c in TENSOR_CASES}["tc_m2d_c256_multi_tile"]
    t = make_inputs(case, torch.float16)
    d = {k: (v.cuda() if v is not None else None) for k, v in t.items()}
    b0 = d["x"].new_empty(0)
    ref = MDCONV_CUDA.backward_ex(d["x"], d["weight"], b0, d["offset"], d["mask"], d["grad_output"], case.k, case.s,
                                  case.p, case.d, case.g, case.dg, False)
    torch.cuda.synchronize()
    ev, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start = torch.cuda.Event(enable_timing=True)
    ev.record()                       # creates the CUDA event handle
    side = torch.cuda.Stream()
    start.record()
    g = MDCONV_CUDA.backward_ex(d["x"], d["weight"], b0, d["offset"], d["mask"], d["grad_output"], case.k, case.s,
                                case.p, case.d, case.g, case.dg, False, weight_ready=ev)
    end.record()
    side.wait_event(ev)
    with torch.cuda.stream(side):
        gw_seen = g["grad_weight"].clone()   # ordered after the event only
    torch.cuda.synchronize()
    assert torch.equal(gw_seen, ref["grad_weight"])
    assert torch.equal(g["grad_input"], ref["grad_input"]) and torch.equal(g["grad_offset"], ref["grad_offset"])
    assert start.elapsed_time(ev) < start.elapsed_time(end)


def test_step_is_cuda_graph_capturable():
    """A forward + backward of the module captured into ONE CUDA graph (torch.cuda.graph) and replayed gives the
    eager bits: nothing in the library allocates, synchronises or touches another stream under capture (the
    inverse-index fork onto the side stream is skipped while the caller's stream is capturing)."""
    import modulated_deform_conv as mdc
    torch.manual_seed(5)
    m = mdc.ModulatedDeformConv2d(64, 64, 3, padding=1, bias=True).cuda().half()
    x = torch.randn(2, 64, 12, 16, device="cuda").half().requires_grad_()
    off = torch.randn(2, 18, 12, 16, device="cuda").half().requires_grad_()
    msk = torch.sigmoid(torch.randn(2, 9, 12, 16, device="cuda")).half().requires_grad_()
    go = torch.randn(2, 64, 12, 16, device="cuda").half()

    def step():
        for t in (x, off, msk, m.weight, m.bias):
            t.grad = None
        y = m(x, off, msk)
        y.backward(go)
        return y

    y_ref = step().detach().clone()
    ref = [t.grad.clone() for t in (x, off, msk, m.weight, m.bias)]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y_static = step()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(y_static, y_ref)
    for t, r in zip((x, off, msk, m.weight, m.bias), ref):
        assert torch.equal(t.grad, r)


def _fuzz_cases(n2d=28, n3d=8):
    """Seeded random geometries the tcgen05 path accepts (channels per group / deformable group multiples of 64)."""
    import random
    rng = random.Random(20261017)
    out = []
    for i in range(n2d + n3d):
        nd = 2 if i < n2d else 3
        g = rng.choice([1, 1, 2, 3])
        dg = rng.choice([1, 1, 2])
        # C: multiple of 64*g and of 64*dg
        lcm = g * dg // math.gcd(g, dg)
        C = 64 * lcm * rng.choice([1, 1, 2])
        Co = g * rng.choice([32, 64, 96, 128]) if nd == 2 else g * rng.choice([32, 64])
        if C > 384:
            C = 64 * lcm
        k = tuple(rng.choice([1, 3, 3]) for _ in range(nd)) if nd == 2 else tuple(rng.choice([1, 3]) for _ in range(nd))
        s_ = tuple(rng.choice([1, 1, 2]) for _ in range(nd))
        d_ = tuple(rng.choice([1, 1, 2]) for _ in range(nd))
        p_ = tuple(rng.choice([0, 1, 2]) for _ in range(nd))
        size = tuple(rng.randint(5, 14 if nd == 2 else 7) for _ in range(nd))
        c = Case(f"fuzz{i}_{nd}d", nd, rng.random() < 0.6, B=rng.randint(1, 3), C=C, Co=Co, size=size, k=k, s=s_, p=p_,
                 d=d_, g=g, dg=dg, sigma=rng.choice([0.0, 0.5, 1.0, 1.0, 3.0]), bias=rng.random() < 0.5, seed=500 + i)
        if all(o >= 1 for o in c.out_size):
            out.append(c)
    return out


@pytest.mark.parametrize("case", _fuzz_cases(), ids=lambda c: c.name)
def test_fuzzed_geometries_tensor_path_matches_oracle(case):
    """Random strides / dilations / paddings / kernel shapes / groups / deformable groups / ragged sizes: forward on
    the tcgen05 path (mandatory), backward on it wherever it supports the shape, against the CPU oracle (fp16, 1e-2)."""
    import ctypes
    import MDCONV_CUDA
    t = make_inputs(case, torch.float16)
    g = MDCONV_CUDA._geom(t["x"], t["weight"], case.k, case.s, case.p, case.d, case.g, case.dg, case.modulated,
                          case.bias)
    g.path = 0
    if MDCONV_CUDA._lib.mdconv_selected_path(ctypes.byref(g), 0) != 2:
        pytest.skip("geometry not eligible for the tensor path")
    try:
        ours = _run_ours(case, t, fwd_path=2, bwd_path=0)
    finally:
        MDCONV_CUDA.set_path(0)
    _compare(ours, _oracle(case, t), TOL[torch.float16], f"{case.name}: {case}")


def test_results_do_not_depend_on_in_step():
    import modulated_deform_conv as mdc
    case = ALL_CASES[1]
    t = make_inputs(case, torch.float32)
    outs = []
    for in_step in (1, 2, 64):
        x, off, m, w = (t[k].cuda() for k in ("x", "offset", "mask", "weight"))
        outs.append(mdc.modulated_deform_conv2d(x, off, m, w, None, case.s, case.p, case.d, case.g, case.dg, in_step))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_error_paths():
    import MDCONV_CUDA
    import modulated_deform_conv as mdc
    case = ALL_CASES[0]
    t = make_inputs(case, torch.float32)
    x, off, m, w, b = (t[k].cuda() for k in ("x", "offset", "mask", "weight", "bias"))
    with pytest.raises(RuntimeError, match="contiguous"):
        mdc.modulated_deform_conv2d(x.transpose(2, 3), off, m, w, b, case.s, case.p, case.d, case.g, case.dg)
    with pytest.raises(RuntimeError, match="kernel shape"):
        MDCONV_CUDA.modulated_deform_conv2d_forward_cuda(x, w, b, off, m, 5, 5, 1, 1, 1, 1, 1, 1, case.g, case.dg, 64, True)
    with pytest.raises(RuntimeError, match="channels"):
        MDCONV_CUDA.modulated_deform_conv2d_forward_cuda(x, w, b, off, m, 3, 3, 1, 1, 1, 1, 1, 1, 1, case.dg, 64, True)
    with pytest.raises(NotImplementedError):
        mdc.modulated_deform_conv2d(x.cpu(), off.cpu(), m.cpu(), w.cpu(), b.cpu(), case.s, case.p, case.d, case.g, case.dg)


def test_modules_and_packs_state_dict_and_run():
    import modulated_deform_conv as mdc
    m = mdc.ModulatedDeformConv2dPack(8, 12, 3, stride=1, padding=1, deformable_groups=2, bias=True).cuda()
    assert set(m.state_dict()) == {"weight", "bias", "conv_offset.weight", "conv_offset.bias",
                                   "conv_mask.weight", "conv_mask.bias"}
    x = torch.randn(2, 8, 9, 9, device="cuda", requires_grad=True)
    y = m(x)
    y.sum().backward()
    assert y.shape == (2, 12, 9, 9) and x.grad is not None and m.conv_offset.weight.grad is not None
    # the Pack equals the plain module fed with the producers' outputs, checked against the oracle
    y_ref = oracle.forward(x.detach().cpu(), m.weight.detach().cpu(), m.bias.detach().cpu(),
                           m.conv_offset(x).detach().cpu(), m.conv_mask(x).detach().cpu(), 1, 1, 1, 1, 2,
                           coord_f32=True)
    assert rel_err(y.detach().cpu().numpy(), y_ref) < 1e-4
    m3 = mdc.DeformConv3dPack(4, 4, (3, 3, 3), padding=1).cuda()
    y3 = m3(torch.randn(1, 4, 5, 5, 5, device="cuda"))
    assert y3.shape == (1, 4, 5, 5, 5) and set(m3.state_dict()) == {"weight", "conv_offset.weight", "conv_offset.bias"}


def test_legacy_named_c_entry_points():
    """The eight reference-named C symbols, called with raw device pointers (what a cgo/JNI binder sees)."""
    import ctypes
    import MDCONV_CUDA
    lib = MDCONV_CUDA._lib
    case = ALL_CASES[0]
    t = make_inputs(case, torch.float32)
    x, off, m, w, b, go = (t[k].cuda() for k in ("x", "offset", "mask", "weight", "bias", "grad_output"))
    y = torch.empty(case.B, case.Co, *case.out_size, device="cuda")
    geo = [case.B, case.C, *case.size, case.Co, *case.k, *case.s, *case.p, *case.d, case.g, case.dg, 64, 1]
    p = lambda tt: ctypes.c_void_p(tt.data_ptr())
    rc = lib.modulated_deform_conv2d_forward_cuda(p(x), p(w), p(b), p(off), p(m), p(y), *geo, 0, None, 0, None)
    assert rc == 0
    gx, goff, gm, gw, gb = (torch.empty_like(v) for v in (x, off, m, w, b))
    rc = lib.modulated_deform_conv2d_backward_cuda(p(x), p(w), p(b), p(off), p(m), p(go), p(gx), p(goff), p(gm),
                                                   p(gw), p(gb), *geo, 0, None, 0, None)
    assert rc == 0
    torch.cuda.synchronize()
    ref = _oracle(case, t)
    assert rel_err(y.cpu().numpy(), ref["y"]) < 1e-4
    assert rel_err(gx.cpu().numpy(), ref["grad_input"]) < 1e-4
    assert rel_err(gw.cpu().numpy(), ref["grad_weight"]) < 1e-4
    assert rel_err(goff.cpu().numpy(), ref["grad_offset"]) < 1e-4


def test_side_stream_and_device_guard():
    """Work is enqueued on torch's current stream (the reference uses legacy stream 0, SURVEY B-5)."""
    import modulated_deform_conv as mdc
    case = ALL_CASES[0]
    t = make_inputs(case, torch.float32)
    x, off, m, w, b = (t[k].cuda() for k in ("x", "offset", "mask", "weight", "bias"))
    y0 = mdc.modulated_deform_conv2d(x, off, m, w, b, case.s, case.p, case.d, case.g, case.dg)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        y1 = mdc.modulated_deform_conv2d(x, off, m, w, b, case.s, case.p, case.d, case.g, case.dg)
    s.synchronize()
    assert torch.equal(y0, y1)


# ---- full-size checks (BASELINE.json configs) ---------------------------------------------------------
def _full_case(name):
    return {
        # cfg2 at a quarter of the per-GPU batch (the oracle does ~6 samples/s here); every tile shape,
        # channel count and kernel path is the benchmark's
        "cfg2": (Case("cfg2_n8", 2, True, B=8, C=256, Co=256, size=(56, 56), bias=False, seed=90), torch.float16),
        "cfg3": (Case("cfg3", 2, False, B=16, C=512, Co=512, size=(28, 28), s=(2, 2), g=4, dg=4, bias=False, seed=91),
                 torch.float32),
        "cfg4": (Case("cfg4_n2", 3, True, B=2, C=64, Co=128, size=(16, 32, 32), k=(3, 3, 3), s=(1, 1, 1),
                      p=(1, 1, 1), d=(1, 1, 1), bias=False, seed=92), torch.float16),
    }[name]


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_baseline_config_matches_oracle(name):
    """The BASELINE.json shapes themselves (reduced batch where the CPU oracle needs it), auto path."""
    case, dtype = _full_case(name)
    t = make_inputs(case, dtype)
    ours = _run_ours(case, t)
    _compare(ours, _oracle(case, t), TOL[dtype], f"{name}/{dtype}")


def test_cfg2_full_batch_properties():
    """Full cfg2 (N=32, fp16) through size-independent properties:
    zero offsets + unit mask == a plain convolution (cuDNN, independent of this repo), and the backward is
    linear in grad_output."""
    import torch.nn.functional as F
    import modulated_deform_conv as mdc
    torch.manual_seed(7)
    B, C, Co, H = 32, 256, 256, 56
    dev = "cuda"
    x = torch.randn(B, C, H, H, device=dev).half().requires_grad_()
    w = ((torch.rand(Co, C, 3, 3, device=dev) * 2 - 1) / (C * 9) ** 0.5).half().requires_grad_()
    off = torch.zeros(B, 18, H, H, device=dev).half().requires_grad_()
    msk = torch.ones(B, 9, H, H, device=dev).half().requires_grad_()
    go = torch.randn(B, Co, H, H, device=dev).half()
    y = mdc.modulated_deform_conv2d(x, off, msk, w, None, 1, 1, 1, 1, 1)
    y.backward(go)
    xr = x.detach().float().requires_grad_()
    wr = w.detach().float().requires_grad_()
    yr = F.conv2d(xr, wr, None, 1, 1)
    yr.backward(go.float())
    rel = lambda a, b: float((a.float() - b).abs().max() / b.abs().max())
    assert rel(y, yr) < 1e-2
    assert rel(x.grad, xr.grad) < 1e-2
    assert rel(w.grad, wr.grad) < 1e-2
    # grad_mask at unit mask / zero offset: sum_c gcol * x over channels == per-tap correlation
    gm = msk.grad.float()
    assert torch.isfinite(gm).all() and torch.isfinite(off.grad.float()).all()
    # linearity of the backward in grad_output (random offsets)
    off2 = torch.randn(B, 18, H, H, device=dev).half()
    msk2 = torch.sigmoid(torch.randn(B, 9, H, H, device=dev)).half()
    go2 = torch.randn(B, Co, H, H, device=dev).half()

    def grads(g_):
        xx = x.detach().clone().requires_grad_()
        oo = off2.clone().requires_grad_()
        mm = msk2.clone().requires_grad_()
        ww = w.detach().clone().requires_grad_()
        mdc.modulated_deform_conv2d(xx, oo, mm, ww, None, 1, 1, 1, 1, 1).backward(g_)
        return [t_.grad.float() for t_ in (xx, oo, mm, ww)]
    ga, gb_, gs = grads(go), grads(go2), grads((go.float() + go2.float()).half())
    for a, b, c in zip(ga, gb_, gs):
        assert float((a + b - c).abs().max() / c.abs().max()) < 2e-2
    # determinism: the tensor path has no atomics
    g1, g2 = grads(go), grads(go)
    for a, b in zip(g1, g2):
        assert torch.equal(a, b)
