"""GPU robustness tests (-m gpu): absent corners never read the input, strongly clustered offsets (long inverse-index
bins), a second device in the same process, non-finite inputs."""
import numpy as np
import pytest
import torch

import oracle
from cases import Case, make_inputs, rel_err

pytestmark = pytest.mark.gpu


def _run(case, t, dev="cuda", path=0):
    import MDCONV_CUDA
    import modulated_deform_conv as mdc
    fn = {(2, False): mdc.deform_conv2d, (2, True): mdc.modulated_deform_conv2d,
          (3, False): mdc.deform_conv3d, (3, True): mdc.modulated_deform_conv3d}[(case.nd, case.modulated)]
    lv = {k: (v.to(dev).requires_grad_() if v is not None and k != "grad_output" else v) for k, v in t.items()}
    args = [lv["x"], lv["offset"]] + ([lv["mask"]] if case.modulated else []) + \
           [lv["weight"], lv["bias"], case.s, case.p, case.d, case.g, case.dg, 64]
    MDCONV_CUDA.set_path(path)
    try:
        y = fn(*args)
        y.backward(t["grad_output"].to(dev))
        torch.cuda.synchronize(dev)
    finally:
        MDCONV_CUDA.set_path(0)
    out = {"y": y.detach(), "grad_input": lv["x"].grad, "grad_offset": lv["offset"].grad,
           "grad_weight": lv["weight"].grad}
    if case.modulated:
        out["grad_mask"] = lv["mask"].grad
    return {k: v.double().cpu().numpy() for k, v in out.items()}


def _oracle(case, t):
    a = (case.s, case.p, case.d, case.g, case.dg)
    ref = oracle.backward(t["x"], t["weight"], t["offset"], t["mask"], t["grad_output"], *a, variant=case.rule,
                          with_bias=case.bias)
    ref["y"] = oracle.forward(t["x"], t["weight"], t["bias"], t["offset"], t["mask"], *a)
    return {k: v for k, v in ref.items() if v is not None and k != "grad_bias"}


@pytest.mark.parametrize("dtype,path", [(torch.float32, 1), (torch.float16, 2)], ids=["generic_f32", "tensor_f16"])
def test_absent_corners_never_read_the_input(dtype, path):
    """The reference returns an exact 0 for out-of-range samples and for corners outside the input WITHOUT reading
    the input (src/mdeformable_conv.cu:19-29,80).  An Inf at x[0, :, 0, 0] (the pixel a zero-weight placeholder
    index would hit) must therefore only reach outputs whose sampling footprint really contains that pixel: here
    every sample of batch element 1 is clean, and with all offsets pushed out of range everything is exactly 0."""
    case = Case("inf_pixel", 2, True, B=2, C=64, Co=64, size=(8, 8), bias=False, seed=301)
    t = make_inputs(case, dtype)
    t["x"][0, :, 0, 0] = float("inf")
    out = _run(case, t, path=path)
    assert np.isfinite(out["y"][1]).all(), "Inf of sample 0 leaked into sample 1 through an absent corner"
    assert np.isfinite(out["grad_offset"][1]).all() and np.isfinite(out["grad_mask"][1]).all()
    t2 = make_inputs(case, dtype)
    t2["x"][0, :, 0, 0] = float("inf")
    t2["offset"] += 1.0e4          # every sample out of range
    out2 = _run(case, t2, path=path)
    assert np.count_nonzero(out2["y"]) == 0 and np.isfinite(out2["y"]).all()
    assert np.count_nonzero(out2["grad_weight"]) == 0 and np.isfinite(out2["grad_weight"]).all()


def test_clustered_offsets_long_bins():
    """Every tap of every output position samples (almost) the same input pixel: the inverse index of the
    atomics-free grad_input then holds a handful of bins with thousands of entries (the block-wide sort path).
    Results must match the oracle and be run-to-run identical."""
    case = Case("clustered", 2, True, B=1, C=64, Co=64, size=(24, 24), bias=False, seed=302)
    t = make_inputs(case, torch.float16)
    H, W = case.size
    ii, jj = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    off = torch.zeros_like(t["offset"], dtype=torch.float32)
    for tap in range(9):
        ki, kj = tap // 3, tap % 3
        # sampling coordinate = out - pad + tap + offset  ->  force it to (5.25, 7.5) for every position and tap
        off[0, 2 * tap] = 5.25 - (ii - 1 + ki)
        off[0, 2 * tap + 1] = 7.5 - (jj - 1 + kj)
    t["offset"] = off.to(torch.float16)
    a = _run(case, t, path=0)
    b = _run(case, t, path=0)
    for k in a:
        assert np.array_equal(a[k], b[k]), f"{k} differs between two runs"
    ref = _oracle(case, t)
    for k, r in ref.items():
        assert rel_err(a[k], r) < 1e-2, (k, rel_err(a[k], r))


def test_second_device_in_the_same_process():
    """cudaFuncAttributeMaxDynamicSharedMemorySize and the SM count are per device: the tcgen05 path on cuda:1
    after cuda:0 in one process (nn.DataParallel, pipeline placement) must work and give the same bits."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    case = Case("two_dev", 2, True, B=2, C=64, Co=64, size=(16, 16), seed=303)
    t = make_inputs(case, torch.float16)
    a = _run(case, t, dev="cuda:0", path=2)
    b = _run(case, t, dev="cuda:1", path=2)
    c = _run(case, t, dev="cuda:0", path=2)
    for k in a:
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k]), k


@pytest.mark.parametrize("hw", [(14, 14), (7, 7), (9, 5)], ids=["14x14", "7x7", "9x5"])
def test_tensor_backward_takes_maps_with_any_position_count(hw):
    """Output maps whose position count is not a multiple of 8 (14x14 -> 196, 7x7 -> 49, 9x5 -> 45) stay on the
    tcgen05 backward (grad_weight / grad_sample, grad_input as the col2im gather over grad_sample's grad_columns), and
    match the oracle."""
    import ctypes
    import MDCONV_CUDA as M
    case = Case(f"p_{hw[0]}x{hw[1]}", 2, True, B=3, C=64, Co=64, size=hw, seed=304)
    t = make_inputs(case, torch.float16)
    g = M._geom(t["x"], t["weight"], case.k, case.s, case.p, case.d, case.g, case.dg, True, case.bias)
    g.path = 0
    assert M._lib.mdconv_selected_path(ctypes.byref(g), 1) == 2, "tensor path refused the backward"
    M.profile_enable(True)
    try:
        out = _run(case, t, path=0)
        names = {n for n, _ in M.profile_read()}
    finally:
        M.profile_enable(False)
    assert {"tc_forward", "tc_grad_weight", "tc_grad_sample", "col2im_gather"} <= names, names
    assert "generic_bwd_data" not in names, names
    ref = _oracle(case, t)
    for k, r in ref.items():
        assert rel_err(out[k], r) < 1e-2, (k, rel_err(out[k], r))


@pytest.mark.parametrize("name", ["tc_m2d_c256_multi_tile", "tc_m2d_ragged", "tc_m2d_g2", "tc_d2d_c64", "tc_m3d_c64",
                                  "tc_m2d_stride2_dg2"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_channels_last_io_matches_nchw_bitwise(name, dtype):
    """activations in torch.channels_last / channels_last_3d go through the SAME kernels with no
    re-layout pass of x (the gather reads the caller's tensor in place), the output and grad_input come back
    channels-last through TMA stores, and every result has the bits of the NCHW call."""
    import MDCONV_CUDA as M
    import modulated_deform_conv as mdc
    from cases import TENSOR_CASES
    case = {c.name: c for c in TENSOR_CASES}[name]
    t = make_inputs(case, dtype)
    fmt = torch.channels_last if case.nd == 2 else torch.channels_last_3d
    fn = {(2, False): mdc.deform_conv2d, (2, True): mdc.modulated_deform_conv2d,
          (3, False): mdc.deform_conv3d, (3, True): mdc.modulated_deform_conv3d}[(case.nd, case.modulated)]

    def run(cl):
        lv = {k: (v.cuda().requires_grad_() if v is not None and k != "grad_output" else v) for k, v in t.items()}
        x = lv["x"]
        if cl:
            x = lv["x"].detach().contiguous(memory_format=fmt).requires_grad_()
        args = [x, lv["offset"]] + ([lv["mask"]] if case.modulated else []) + \
               [lv["weight"], lv["bias"], case.s, case.p, case.d, case.g, case.dg, 64]
        M.profile_enable(True)
        try:
            y = fn(*args)
            go = t["grad_output"].cuda()
            y.backward(go.contiguous(memory_format=fmt) if cl else go)
            torch.cuda.synchronize()
            names = [n for n, _ in M.profile_read()]
        finally:
            M.profile_enable(False)
        out = {"y": y.detach(), "grad_input": x.grad, "grad_offset": lv["offset"].grad, "grad_weight": lv["weight"].grad}
        if case.modulated:
            out["grad_mask"] = lv["mask"].grad
        if case.bias:
            out["grad_bias"] = lv["bias"].grad
        return out, names

    ref, names_ref = run(False)
    got, names = run(True)
    assert got["y"].is_contiguous(memory_format=fmt) and got["grad_input"].is_contiguous(memory_format=fmt)
    assert "to_channels_last" in names_ref and "to_channels_last" not in names, names
    assert names.count("from_channels_last") == 1
    for k in ref:
        assert torch.equal(got[k].contiguous(), ref[k].contiguous()), k
    orc = _oracle(case, t)
    tol = 1e-2 if dtype == torch.float16 else 3e-2
    for k, r in orc.items():
        assert rel_err(got[k].contiguous().double().cpu().numpy(), r) < tol, k


def test_channels_last_input_on_an_unsupported_shape_falls_back_to_a_copy():
    """fp32 / small-channel calls have no channels-last kernels: the binding converts, results are the NCHW ones."""
    import modulated_deform_conv as mdc
    case = Case("cl_fallback", 2, True, B=2, C=8, Co=8, size=(9, 7), seed=305)
    t = make_inputs(case, torch.float32)
    d = {k: (v.cuda() if v is not None else None) for k, v in t.items()}
    y0 = mdc.modulated_deform_conv2d(d["x"], d["offset"], d["mask"], d["weight"], d["bias"], case.s, case.p, case.d)
    xcl = d["x"].contiguous(memory_format=torch.channels_last).requires_grad_()
    y1 = mdc.modulated_deform_conv2d(xcl, d["offset"], d["mask"], d["weight"], d["bias"], case.s, case.p, case.d)
    y1.sum().backward()
    assert torch.equal(y0, y1) and xcl.grad is not None and xcl.grad.shape == xcl.shape


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
@pytest.mark.parametrize("kind", ["m2d", "d2d", "m3d"])
def test_fused_pack_producer(kind, dtype):
    """the Pack modules' conv_offset / conv_mask run as ONE regular-grid pass of the deformable
    forward kernel over the channels-last copy of x that the deformable convolution then reuses (one re-layout pass,
    no cuDNN call in the forward).  The producer equals the two reference convolutions (fp32 cuDNN on the same
    values), the whole module equals the unfused route, and every parameter receives its gradient."""
    import torch.nn.functional as F
    import MDCONV_CUDA as M
    import modulated_deform_conv as mdc
    torch.manual_seed(11)
    if kind == "m3d":
        mod = mdc.ModulatedDeformConv3dPack(64, 64, 3, padding=1, deformable_groups=1, bias=True)
        x = torch.randn(1, 64, 6, 8, 8)
        conv = F.conv3d
    elif kind == "m2d":
        mod = mdc.ModulatedDeformConv2dPack(128, 64, 3, stride=1, padding=1, deformable_groups=2, bias=True)
        x = torch.randn(2, 128, 12, 10)
        conv = F.conv2d
    else:
        mod = mdc.DeformConv2dPack(64, 64, 3, stride=2, padding=1)
        x = torch.randn(2, 64, 13, 11)
        conv = F.conv2d
    mod = mod.cuda().to(dtype)
    with torch.no_grad():
        for nme, p_ in mod.named_parameters():
            p_.copy_(torch.randn_like(p_) * (0.01 if nme.startswith("conv_") else 0.05))   # offsets of ~1 px
    x = x.cuda().to(dtype)
    tol = 1e-2 if dtype == torch.float16 else 3e-2
    # 1. the producer alone against fp32 convolutions of the same 16-bit values
    producers = [mod.conv_offset] + ([mod.conv_mask] if kind != "d2d" else [])
    w = torch.cat([c.weight for c in producers], 0).detach()
    b = torch.cat([c.bias for c in producers], 0).detach()
    o1, o2 = M.producer_conv(x, w, b, mod.conv_offset.out_channels, mod.kernel_size, mod.conv_offset.stride,
                             mod.conv_offset.padding)
    ref = conv(x.float(), w.float(), b.float(), mod.conv_offset.stride, mod.conv_offset.padding)
    n = mod.conv_offset.out_channels
    assert rel_err(o1.float().cpu().numpy(), ref[:, :n].cpu().numpy()) < tol
    if o2 is not None:
        assert rel_err(o2.float().cpu().numpy(), ref[:, n:].cpu().numpy()) < tol

    # 2. the module: fused vs unfused route, forward and every gradient
    def run(fuse):
        mod.fuse_producers = fuse
        for p_ in mod.parameters():
            p_.grad = None
        xx = x.clone().requires_grad_()
        M.profile_enable(True)
        try:
            y = mod(xx)
            names = [nm for nm, _ in M.profile_read()]
        finally:
            M.profile_enable(False)
        y.float().square().mean().backward()
        torch.cuda.synchronize()
        return y.detach(), xx.grad, {k: v.grad.clone() for k, v in mod.named_parameters()}, names

    y1, gx1, gp1, names1 = run(True)
    y0, gx0, gp0, names0 = run(False)
    assert "tc_producer_conv" in names1 and names1.count("to_channels_last") == 1, names1
    assert "tc_producer_conv" not in names0
    loose = 5 * tol   # the two routes round the offsets differently; the deformable conv amplifies that slightly
    assert rel_err(y1.float().cpu().numpy(), y0.float().cpu().numpy()) < loose
    assert float((gx1.double() - gx0.double()).norm() / gx0.double().norm()) < 4 * loose
    # parameter gradients: relative L2 (the offset gradients are piecewise: an offset that rounds to the other side
    # of a pixel boundary in one route flips a few terms, which max-norm on a small tensor over-weights)
    # (the two routes round the offsets differently -- cuDNN vs this kernel -- and d(output)/d(offset) jumps at pixel
    # boundaries, so everything downstream of the offsets gets the looser bar)
    for k in gp0:
        assert gp1[k] is not None, k
        a, r = gp1[k].double(), gp0[k].double()
        bar = 4 * loose if k.startswith("conv_offset") else loose
        assert float((a - r).norm() / r.norm()) < bar, (k, float((a - r).norm() / r.norm()))


def test_pack_hooks_and_wrapped_producers_take_the_reference_route():
    """A forward hook on conv_offset must fire (it never did through the merged-filter shortcut), a pruned producer
    must use its recomputed weight, and a replaced producer must be called: each of these disables the fused route."""
    import torch.nn as nn
    import torch.nn.utils.prune as prune
    import modulated_deform_conv as mdc
    torch.manual_seed(12)
    mod = mdc.ModulatedDeformConv2dPack(64, 64, 3, padding=1).cuda().half()
    x = torch.randn(1, 64, 8, 8, device="cuda").half()
    seen = []
    h = mod.conv_offset.register_forward_hook(lambda m, i, o: seen.append(tuple(o.shape)))
    mod(x)
    h.remove()
    assert seen == [(1, 18, 8, 8)]
    prune.l1_unstructured(mod.conv_mask, "weight", amount=0.9)
    y_pruned = mod(x)
    ref_mask = mod.conv_mask(x)
    assert float((ref_mask == mod.conv_mask.bias.view(1, -1, 1, 1)).float().mean()) < 1.0   # still a convolution
    prune.remove(mod.conv_mask, "weight")
    y_again = mod(x)                       # plain again: fused route, same pruned weights
    assert rel_err(y_again.detach().float().cpu().numpy(), y_pruned.detach().float().cpu().numpy()) < 5e-2
    calls = []

    class Wrapped(nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, t):
            calls.append(1)
            return self.inner(t)

    mod.conv_offset = Wrapped(mod.conv_offset)
    mod(x)
    assert calls == [1]
