"""CPU tests that pin the C oracle (no GPU needed).

Pins, strongest first:
  * tests/golden/ref_*.npz   -- outputs of the UNMODIFIED reference kernels compiled from
                                /root/reference and run on a B200 (oracle/gen_ref_golden.py);
  * tests/golden/tv_*.npz    -- torchvision.ops.deform_conv2d CPU fp64 fixtures
                                (tests/golden/make_tv_golden.py);
  * an independent pure-PyTorch autograd implementation (tests/torch_ref.py);
  * torch conv2d / conv3d at zero offset; the my_test.py known-answer case.
"""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from cases import ALL_CASES, GENERIC_OFFSET_CASES, Case, make_inputs, rel_err
from torch_ref import deform_conv_nd

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle_all(case, t, variant=None, precision="f64", coord_f32=False):
    a = (case.s, case.p, case.d, case.g, case.dg)
    y = oracle.forward(t["x"], t["weight"], t["bias"], t["offset"], t["mask"], *a, precision=precision,
                       coord_f32=coord_f32)
    g = oracle.backward(t["x"], t["weight"], t["offset"], t["mask"], t["grad_output"], *a, variant=variant,
                        precision=precision, coord_f32=coord_f32, with_bias=case.bias)
    return y, g


@pytest.mark.parametrize("case", GENERIC_OFFSET_CASES, ids=lambda c: c.name)
def test_oracle_matches_autograd_reference(case):
    """At generic offsets both gradient rules coincide with autograd of an independent forward."""
    t = make_inputs(case, torch.float64)
    leaves = {k: (v.clone().requires_grad_() if v is not None and k != "grad_output" else v) for k, v in t.items()}
    y = deform_conv_nd(leaves["x"], leaves["offset"], leaves["mask"], leaves["weight"], leaves["bias"],
                       case.s, case.p, case.d, case.g, case.dg)
    y.backward(t["grad_output"])
    for variant in (0, 1):
        yo, g = _oracle_all(case, t, variant=variant)
        assert rel_err(yo, y.detach().numpy()) < 1e-12
        assert rel_err(g["grad_input"], leaves["x"].grad.numpy()) < 1e-11
        assert rel_err(g["grad_offset"], leaves["offset"].grad.numpy()) < 1e-11
        assert rel_err(g["grad_weight"], leaves["weight"].grad.numpy()) < 1e-11
        if case.modulated:
            assert rel_err(g["grad_mask"], leaves["mask"].grad.numpy()) < 1e-11
        if case.bias:
            assert rel_err(g["grad_bias"], leaves["bias"].grad.numpy()) < 1e-11


@pytest.mark.parametrize("nd", [2, 3])
def test_zero_offset_equals_plain_convolution(nd):
    """offset = 0, mask = 1  ==  F.conv2d / F.conv3d (forward and input/weight/bias gradients)."""
    torch.manual_seed(3)
    size, k = ((9, 8), (3, 3)) if nd == 2 else ((5, 6, 5), (3, 2, 3))
    s, p, d, g = (2, 1, 1)[:nd], (1, 0, 1)[:nd], (1, 2, 1)[:nd], 2
    x = torch.randn(2, 4, *size, dtype=torch.float64, requires_grad=True)
    w = torch.randn(6, 2, *k, dtype=torch.float64, requires_grad=True)
    b = torch.randn(6, dtype=torch.float64, requires_grad=True)
    conv = torch.nn.functional.conv2d if nd == 2 else torch.nn.functional.conv3d
    y = conv(x, w, b, stride=s, padding=p, dilation=d, groups=g)
    go = torch.randn_like(y)
    y.backward(go)
    K = int(np.prod(k))
    off = torch.zeros(2, nd * K, *y.shape[2:], dtype=torch.float64)
    msk = torch.ones(2, K, *y.shape[2:], dtype=torch.float64)
    for mask in (msk, None):
        yo = oracle.forward(x, w, b, off, mask, s, p, d, g, 1)
        assert rel_err(yo, y.detach().numpy()) < 1e-12
        gr = oracle.backward(x, w, off, mask, go, s, p, d, g, 1)
        assert rel_err(gr["grad_input"], x.grad.numpy()) < 1e-12
        assert rel_err(gr["grad_weight"], w.grad.numpy()) < 1e-12
        assert rel_err(gr["grad_bias"], b.grad.numpy()) < 1e-12


def test_known_answer_my_test_inputs():
    """The reference's only 'test' (my_test.py:2-15): x=1 (1x1x5x5), offset=0, mask=1, W=1 (3x3),
    bias=0, s=p=d=1.  Expected values: SURVEY A.5 (analytic)."""
    x = np.ones((1, 1, 5, 5)); off = np.zeros((1, 18, 5, 5)); msk = np.ones((1, 9, 5, 5))
    w = np.ones((1, 1, 3, 3)); b = np.zeros((1,))
    y = oracle.forward(x, w, b, off, msk, 1, 1, 1, 1, 1)
    exp = np.array([[4, 6, 6, 6, 4], [6, 9, 9, 9, 6], [6, 9, 9, 9, 6], [6, 9, 9, 9, 6], [4, 6, 6, 6, 4]], float)
    assert np.array_equal(y[0, 0], exp) and y.sum() == 169
    go = np.ones_like(y)
    g = oracle.backward(x, w, off, msk, go, 1, 1, 1, 1, 1, variant=0)
    assert np.array_equal(g["grad_input"][0, 0], exp)
    assert np.array_equal(g["grad_weight"][0, 0], np.array([[16, 20, 16], [20, 25, 20], [16, 20, 16]], float))
    assert g["grad_bias"][0] == 25
    assert np.array_equal(g["grad_mask"].reshape(9, -1).sum(1), np.array([16, 20, 16, 20, 25, 20, 16, 20, 16], float))
    # rule A: -1 where the sampled point sits on the last row (h) / last column (w), 0 elsewhere --
    # including taps that land exactly on coordinate -1.
    goff = g["grad_offset"][0].reshape(9, 2, 5, 5)
    assert set(np.unique(goff)) <= {-1.0, 0.0}
    tap8_h = goff[8, 0]
    assert np.array_equal(tap8_h[3, :4], -np.ones(4)) and tap8_h[:3].sum() == 0 and tap8_h[4].sum() == 0
    assert goff[0].sum() == 0 or np.all(goff[0] <= 0)
    # rule B (plain op, same inputs): every in-range tap gives -1 on both components
    gB = oracle.backward(x, w, off, None, go, 1, 1, 1, 1, 1, variant=1)
    goffB = gB["grad_offset"][0].reshape(9, 2, 5, 5)
    in_range = g["grad_mask"][0].reshape(9, 5, 5) == 1
    assert np.array_equal(goffB[:, 0][in_range], -np.ones(in_range.sum()))
    assert np.array_equal(goffB[:, 1][in_range], -np.ones(in_range.sum()))
    assert goffB[:, 0][~in_range].sum() == 0


def test_gradient_rules_differ_only_at_integer_coordinates():
    case = [c for c in ALL_CASES if c.name == "m2d_half_grid"][0]
    t = make_inputs(case, torch.float64)
    _, gA = _oracle_all(case, t, variant=0)
    _, gB = _oracle_all(case, t, variant=1)
    for k in ("grad_input", "grad_mask", "grad_weight", "grad_bias"):
        assert rel_err(gA[k], gB[k]) < 1e-12
    assert rel_err(gA["grad_offset"], gB["grad_offset"]) > 1e-3  # the rules really are different here


def test_f32_build_of_oracle_tracks_f64():
    case = ALL_CASES[0]
    t = make_inputs(case, torch.float32)
    y64, g64 = _oracle_all(case, t, precision="f64", coord_f32=True)
    y32, g32 = _oracle_all(case, t, precision="f32", coord_f32=True)
    assert rel_err(y32, y64) < 1e-5
    for k in g64:
        assert rel_err(g32[k], g64[k]) < 1e-4


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "tv_*.npz"))), ids=os.path.basename)
def test_against_torchvision_fixtures(path):
    z = np.load(path)
    a = (tuple(z["stride"]), tuple(z["padding"]), tuple(z["dilation"]), int(z["groups"]), int(z["dg"]))
    mask = z["mask"] if "mask" in z.files else None
    y = oracle.forward(z["x"], z["weight"], z["bias"], z["offset"], mask, *a)
    assert rel_err(y, z["y"]) < 1e-12
    g = oracle.backward(z["x"], z["weight"], z["offset"], mask, z["grad_output"], *a, variant=0)
    for k in ("grad_input", "grad_offset", "grad_weight", "grad_bias"):
        assert rel_err(g[k], z[k]) < 1e-11, k
    if mask is not None:
        assert rel_err(g["grad_mask"], z["grad_mask"]) < 1e-11


_REF_GOLDEN = sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz")))


@pytest.mark.parametrize("path", _REF_GOLDEN, ids=os.path.basename)
def test_against_reference_kernel_outputs(path):
    """Golden outputs of the real reference (fp32 / fp64 kernels run on a B200)."""
    z = np.load(path)
    a = (tuple(z["stride"]), tuple(z["padding"]), tuple(z["dilation"]), int(z["groups"]), int(z["dg"]))
    mask = z["mask"] if "mask" in z.files else None
    is32 = z["x"].dtype == np.float32
    tol = 2e-5 if is32 else 1e-11
    y = oracle.forward(z["x"], z["weight"], z["bias"] if z["bias"].size else None, z["offset"], mask, *a,
                       coord_f32=is32)
    assert rel_err(y, z["y"]) < tol
    g = oracle.backward(z["x"], z["weight"], z["offset"], mask, z["grad_output"], *a, variant=int(z["rule"]),
                        coord_f32=is32, with_bias=bool(z["bias"].size))
    for k in ("grad_input", "grad_offset", "grad_weight"):
        assert rel_err(g[k], z[k]) < tol, k
    if mask is not None:
        assert rel_err(g["grad_mask"], z["grad_mask"]) < tol
    if z["bias"].size:
        assert rel_err(g["grad_bias"], z["grad_bias"]) < tol


def test_reference_goldens_present():
    """The reference-kernel fixtures are the oracle's primary pin; make their absence visible."""
    if not _REF_GOLDEN:
        pytest.skip("tests/golden/ref_*.npz not generated yet (needs one GPU run of oracle/gen_ref_golden.py)")
    assert len(_REF_GOLDEN) >= 8
