"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the header declares,
validates geometry without touching a GPU, and the Python drop-in modules expose the reference's
names.  No compute calls here."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    hdr = open(os.path.join(ROOT, "include", "mdconv_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+char\s*\*|int|size_t)\s+(\w+)\s*\(", hdr, flags=re.M)
    return sorted(set(names))


def test_header_symbols_are_exported():
    import MDCONV_CUDA
    names = _declared_functions()
    assert len(names) >= 18, names
    for n in ("mdconv_forward", "mdconv_backward", "deform_conv2d_forward_cuda", "deform_conv2d_backward_cuda",
              "modulated_deform_conv2d_forward_cuda", "modulated_deform_conv2d_backward_cuda",
              "deform_conv3d_forward_cuda", "deform_conv3d_backward_cuda", "modulated_deform_conv3d_forward_cuda",
              "modulated_deform_conv3d_backward_cuda"):
        assert n in names
    lib = ctypes.CDLL(MDCONV_CUDA.library_path())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mdconv_b200.h but not exported"


def test_geometry_validation_and_sizes():
    import MDCONV_CUDA as M
    x = torch.empty(32, 256, 56, 56, dtype=torch.float16)
    w = torch.empty(256, 256, 3, 3, dtype=torch.float16)
    g = M._geom(x, w, (3, 3), (1, 1), (1, 1), (1, 1), 1, 1, True, False)
    out = (ctypes.c_int32 * 3)()
    assert M._lib.mdconv_out_size(ctypes.byref(g), out) == 0 and (out[0], out[1]) == (56, 56)
    assert M._lib.mdconv_workspace_bytes(ctypes.byref(g), 1) > 0
    assert M._lib.mdconv_selected_path(ctypes.byref(g), 0) in (1, 2)
    # stride 2 / dilation 2 / 3D out sizes follow (in + 2p - (d(k-1)+1))/s + 1
    x3 = torch.empty(2, 4, 9, 8, 7)
    w3 = torch.empty(4, 4, 3, 3, 2)
    g3 = M._geom(x3, w3, (3, 3, 2), (2, 1, 1), (1, 0, 1), (1, 2, 1), 1, 1, False, False)
    assert M._lib.mdconv_out_size(ctypes.byref(g3), out) == 0 and tuple(out) == (5, 4, 8)
    # fp32 accumulates straight into the outputs: no workspace on the generic path
    g32 = M._geom(x3, w3, (3, 3, 2), (1, 1, 1), (1, 1, 1), (1, 1, 1), 1, 1, False, False)
    assert M._lib.mdconv_workspace_bytes(ctypes.byref(g32), 1) == 0
    # errors: groups not dividing channels, kernel larger than padded input, bad nd, bad dtype
    bad = M._geom(x3, w3, (3, 3, 2), (1, 1, 1), (1, 1, 1), (1, 1, 1), 3, 1, False, False)
    assert M._lib.mdconv_out_size(ctypes.byref(bad), out) == 1
    bad = M._geom(x3, w3, (30, 3, 2), (1, 1, 1), (0, 0, 0), (1, 1, 1), 1, 1, False, False)
    assert M._lib.mdconv_out_size(ctypes.byref(bad), out) == 1
    bad = M._geom(x3, w3, (3, 3, 2), (1, 1, 1), (1, 1, 1), (1, 1, 1), 1, 1, False, False)
    bad.nd = 4
    assert M._lib.mdconv_out_size(ctypes.byref(bad), out) == 1
    bad.nd, bad.dtype = 3, 9
    assert M._lib.mdconv_out_size(ctypes.byref(bad), out) == 4
    assert b"geometry" in M._lib.mdconv_status_string(1)
    # null pointers are rejected before any CUDA call
    assert M._lib.mdconv_forward(ctypes.byref(g3), None, None, None, None, None, None, None, 0, None) == 2


def test_grad_columns_budget_shapes_the_backward_workspace():
    """Host logic of the col2im-gather grad_input (no GPU needed): the backward workspace of a 16-bit tensor-path call
    holds ONE grad_columns buffer -- the whole batch's when it fits the cap, a whole number of samples otherwise -- and
    none at all with a cap of 0 (grad_input as a contraction) or for fp32 callers."""
    import MDCONV_CUDA as M
    x = torch.empty(32, 256, 56, 56, dtype=torch.float16)
    w = torch.empty(256, 256, 3, 3, dtype=torch.float16)
    g = M._geom(x, w, (3, 3), (1, 1), (1, 1), (1, 1), 1, 1, True, False)
    g.path = 2
    if M._lib.mdconv_selected_path(ctypes.byref(g), 1) != 2:
        pytest.skip("tensor path not selectable on this host")
    sample = 3136 * 9 * 256 * 2                       # bytes of one sample's grad_columns [K][P][C]
    prev = M.set_gcol_budget_mb(1024)
    try:
        full = M._lib.mdconv_workspace_bytes(ctypes.byref(g), 1)
        assert M.set_gcol_budget_mb(64) == 1024       # the setter returns the previous cap
        chunked = M._lib.mdconv_workspace_bytes(ctypes.byref(g), 1)
        M.set_gcol_budget_mb(0)
        none = M._lib.mdconv_workspace_bytes(ctypes.byref(g), 1)
        fwd = M._lib.mdconv_workspace_bytes(ctypes.byref(g), 0)
    finally:
        M.set_gcol_budget_mb(prev if prev >= 0 else 1024)
    assert full - chunked == (32 - (64 * 2 ** 20) // sample) * sample      # 32 samples vs the 4 that fit 64 MiB
    assert 0 < none < chunked < full
    got_wt = 32 * 256 * 3136 * 2 + 256 * 2304 * 2     # what the contraction route needs instead: got + wt
    assert abs((full - 32 * sample) - (none - got_wt)) < 4096              # same index / partials either way
    assert fwd < none                                  # the forward needs neither


def test_python_surface_matches_reference_names():
    import MDCONV_CUDA
    import modulated_deform_conv as mdc
    for n in ("deform_conv2d_forward_cuda", "deform_conv2d_backward_cuda", "modulated_deform_conv2d_forward_cuda",
              "modulated_deform_conv2d_backward_cuda", "deform_conv3d_forward_cuda", "deform_conv3d_backward_cuda",
              "modulated_deform_conv3d_forward_cuda", "modulated_deform_conv3d_backward_cuda"):
        assert callable(getattr(MDCONV_CUDA, n))
    for n in ("DeformConv2d", "ModulatedDeformConv2d", "DeformConv3d", "ModulatedDeformConv3d", "DeformConv2dPack",
              "ModulatedDeformConv2dPack", "DeformConv3dPack", "ModulatedDeformConv3dPack", "deform_conv2d",
              "modulated_deform_conv2d", "deform_conv3d", "modulated_deform_conv3d"):
        assert hasattr(mdc, n)
    m = mdc.ModulatedDeformConv3d(8, 16, 3, stride=2, padding=1, groups=2, deformable_groups=2, bias=True)
    assert m.weight.shape == (16, 4, 3, 3, 3) and m.bias.shape == (16,) and m.kernel_size == (3, 3, 3)
    assert float(m.bias.abs().sum()) == 0.0
    bound = 1.0 / (8 * 27) ** 0.5
    assert float(m.weight.abs().max()) <= bound
    with pytest.raises(AssertionError):
        mdc.DeformConv2d(6, 8, 3, groups=4)
    # the product path has no CPU fallback: CPU tensors are refused, as in the reference
    m2 = mdc.ModulatedDeformConv2d(4, 4, 3, padding=1)
    with pytest.raises(NotImplementedError):
        m2(torch.zeros(1, 4, 5, 5), torch.zeros(1, 18, 5, 5), torch.ones(1, 9, 5, 5))
    with pytest.raises(RuntimeError):
        MDCONV_CUDA.modulated_deform_conv2d_forward_cuda(
            torch.zeros(1, 4, 5, 5), m2.weight.data, torch.zeros(0), torch.zeros(1, 18, 5, 5),
            torch.ones(1, 9, 5, 5), 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 64, False)


def test_dispatcher_ops_trace_with_fake_tensors():
    """torch.library registration (mdconv_ops): shapes / dtypes propagate through FakeTensorMode without touching
    a GPU, for the four op kinds; CPU tensors are refused (no CPU path)."""
    import mdconv_ops  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        for nd, modulated in ((2, True), (2, False), (3, True), (3, False)):
            sp = (10, 12) if nd == 2 else (4, 6, 5)
            K = 3 ** nd
            x = torch.empty(2, 64, *sp, dtype=torch.float16, device="cuda")
            w = torch.empty(32, 32, *([3] * nd), dtype=torch.float16, device="cuda")
            b = torch.empty(32, dtype=torch.float16, device="cuda")
            out_sp = tuple((i + 2 - 3) // 2 + 1 for i in sp)
            off = torch.empty(2, 2 * nd * K, *out_sp, dtype=torch.float16, device="cuda")
            m = torch.empty(2, 2 * K, *out_sp, dtype=torch.float16, device="cuda") if modulated else None
            geo = ([2] * nd, [1] * nd, [1] * nd, 2, 2)
            y = torch.ops.mdconv_b200.deform_conv(x, w, b, off, m, *geo)
            assert tuple(y.shape) == (2, 32) + out_sp and y.dtype == torch.float16 and y.device.type == "cuda"
            g = torch.ops.mdconv_b200.deform_conv_backward(x, w, b, off, m, y, *geo)
            assert [tuple(t.shape) for t in g] == [tuple(x.shape), tuple(off.shape),
                                                   tuple(m.shape) if modulated else (0,), tuple(w.shape), (32,)]
    with pytest.raises(NotImplementedError):
        torch.ops.mdconv_b200.deform_conv(torch.zeros(1, 4, 5, 5), torch.zeros(4, 4, 3, 3), None,
                                          torch.zeros(1, 18, 5, 5), None, [1, 1], [1, 1], [1, 1], 1, 1)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "modulated-deform-conv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_bench_algorithmic_figures_match_the_survey():
    """bench.py's per-launch algorithmic bytes / FLOPs are SURVEY 8(d)'s figures (cfg2: fwd 109.36 MB, bwd 167.34 MB,
    118.38 GFLOP per contraction; cfg3 35.38 / 64.33 MB, 3.70 GFLOP; cfg4 79.09 / 124.62 MB, 57.98 GFLOP)."""
    sys.path.insert(0, ROOT)
    import bench
    want = {"cfg2": (109.36e6, 167.34e6, 118.38e9), "cfg3": (35.38e6, 64.33e6, 3.70e9), "cfg4": (79.09e6, 124.62e6, 57.98e9)}
    for name, (bf, bb, fl) in want.items():
        w = bench.WORKLOADS[name]
        es = {"f16": 2, "f32": 4}[w["dtype"]]
        got_bf, got_bb, got_ff, got_fb = bench.algorithmic(w, w["B"], es)
        assert abs(got_bf - bf) / bf < 2e-3 and abs(got_bb - bb) / bb < 2e-3, (name, got_bf, got_bb)
        assert abs(got_ff - fl) / fl < 2e-3 and got_fb == 2 * got_ff, (name, got_ff)
