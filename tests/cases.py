"""Shared parity cases: seeded synthetic inputs for the four ops (tests only)."""
from dataclasses import dataclass, field
from typing import Tuple

import torch


@dataclass(frozen=True)
class Case:
    name: str
    nd: int
    modulated: bool
    B: int = 2
    C: int = 8
    Co: int = 12
    size: Tuple[int, ...] = (7, 6)
    k: Tuple[int, ...] = (3, 3)
    s: Tuple[int, ...] = (1, 1)
    p: Tuple[int, ...] = (1, 1)
    d: Tuple[int, ...] = (1, 1)
    g: int = 1
    dg: int = 1
    sigma: float = 1.0       # std of the offsets in pixels; 0 -> all-integer sampling grid
    offset_bias: float = 0.0  # constant added to every offset (e.g. 1e4 -> everything out of range)
    bias: bool = True
    seed: int = 0

    @property
    def out_size(self):
        return tuple((i + 2 * p - (d * (k - 1) + 1)) // s + 1
                     for i, k, s, p, d in zip(self.size, self.k, self.s, self.p, self.d))

    @property
    def taps(self):
        t = 1
        for k in self.k:
            t *= k
        return t

    @property
    def rule(self):
        return 0 if (self.nd == 2 and self.modulated) else 1


def make_inputs(case: Case, dtype=torch.float64):
    """-> dict of CPU tensors (generated in fp32-representable fp64, then cast to ``dtype``)."""
    gen = torch.Generator().manual_seed(1000 + case.seed)
    r = lambda *shape: torch.randn(*shape, generator=gen, dtype=torch.float32)
    x = r(case.B, case.C, *case.size)
    fan = (case.C // case.g) * case.taps
    w = (torch.rand(case.Co, case.C // case.g, *case.k, generator=gen) * 2 - 1) / fan ** 0.5
    b = r(case.Co) if case.bias else None
    off = r(case.B, case.dg * case.nd * case.taps, *case.out_size) * case.sigma + case.offset_bias
    if case.sigma == 0.5:  # half-integer grid flavour: snap to multiples of 0.5 (exact in every dtype)
        off = torch.round(off * 2) / 2
    msk = torch.sigmoid(r(case.B, case.dg * case.taps, *case.out_size)) if case.modulated else None
    go = r(case.B, case.Co, *case.out_size)
    cast = lambda t: None if t is None else t.to(dtype)
    return {"x": cast(x), "weight": cast(w), "bias": cast(b), "offset": cast(off), "mask": cast(msk),
            "grad_output": cast(go)}


CASES_2D = [
    Case("m2d_base", 2, True, g=2, dg=2),
    Case("m2d_nobias_b3", 2, True, B=3, bias=False, seed=1),
    Case("m2d_stride2", 2, True, size=(9, 8), s=(2, 2), seed=2),
    Case("m2d_dil2_pad2", 2, True, size=(9, 9), p=(2, 2), d=(2, 2), g=4, dg=4, C=8, Co=8, seed=3),
    Case("m2d_pad0_k3x1", 2, True, k=(3, 1), p=(0, 0), seed=4),
    Case("m2d_k1x3_s12", 2, True, k=(1, 3), s=(1, 2), p=(0, 1), dg=2, seed=5),
    Case("m2d_sigma4", 2, True, sigma=4.0, seed=6),
    Case("m2d_integer_grid", 2, True, sigma=0.0, seed=7),
    Case("m2d_half_grid", 2, True, sigma=0.5, seed=8),
    Case("m2d_all_oob", 2, True, offset_bias=1.0e4, seed=9),
    Case("m2d_b1_c64", 2, True, B=1, C=64, Co=64, size=(10, 12), seed=10),
    Case("m2d_odd_tiles", 2, True, B=2, C=6, Co=70, size=(11, 13), seed=11),
    Case("m2d_co130_wide_tile", 2, True, B=1, C=6, Co=130, size=(9, 7), seed=12),   # 128-wide SIMT forward tile + tail
    Case("d2d_base", 2, False, g=2, dg=2, seed=20),
    Case("d2d_nobias", 2, False, B=1, bias=False, seed=21),
    Case("d2d_stride2_dg4", 2, False, C=16, Co=16, size=(10, 10), s=(2, 2), g=4, dg=4, seed=22),
    Case("d2d_integer_grid", 2, False, sigma=0.0, seed=23),
    Case("d2d_half_grid", 2, False, sigma=0.5, seed=24),
    Case("d2d_all_oob", 2, False, offset_bias=-1.0e4, seed=25),
    Case("d2d_cfg1", 2, False, B=2, C=4, Co=8, size=(16, 16), bias=False, seed=26),  # BASELINE configs[0]
]

_t3 = dict(size=(5, 6, 4), k=(3, 3, 3), s=(1, 1, 1), p=(1, 1, 1), d=(1, 1, 1))
CASES_3D = [
    Case("m3d_base", 3, True, C=4, Co=6, g=2, dg=2, seed=30, **_t3),
    Case("m3d_nobias_k132", 3, True, C=4, Co=4, size=(5, 6, 4), k=(1, 3, 2), s=(1, 2, 1), p=(0, 1, 1),
         d=(1, 1, 2), bias=False, seed=31),
    Case("m3d_integer_grid", 3, True, C=4, Co=4, sigma=0.0, seed=32, **_t3),
    Case("m3d_all_oob", 3, True, C=4, Co=4, offset_bias=1.0e4, seed=33, **_t3),
    Case("d3d_base", 3, False, C=4, Co=6, g=2, dg=1, seed=40, **_t3),
    Case("d3d_stride2", 3, False, C=4, Co=4, size=(6, 6, 6), k=(3, 3, 3), s=(2, 2, 2), p=(1, 1, 1), d=(1, 1, 1),
         dg=2, seed=41),
    Case("d3d_integer_grid", 3, False, C=4, Co=4, sigma=0.0, seed=42, **_t3),
    Case("d3d_half_grid", 3, False, C=4, Co=4, sigma=0.5, seed=43, **_t3),
]

ALL_CASES = CASES_2D + CASES_3D

# shapes the tcgen05 path accepts (channels per group / deformable group multiples of 64)
TENSOR_CASES = [
    Case("tc_m2d_c64", 2, True, B=2, C=64, Co=64, size=(12, 11), seed=50),
    Case("tc_m2d_c128_co256_nobias", 2, True, B=3, C=128, Co=256, size=(8, 14), bias=False, seed=51),
    Case("tc_m2d_stride2_dg2", 2, True, B=2, C=128, Co=128, size=(15, 13), s=(2, 2), dg=2, seed=52),
    Case("tc_m2d_dil2_sigma4", 2, True, B=1, C=64, Co=64, size=(16, 16), p=(2, 2), d=(2, 2), sigma=4.0, seed=53),
    Case("tc_m2d_integer_grid", 2, True, B=2, C=64, Co=64, size=(8, 9), sigma=0.0, seed=54),
    Case("tc_m2d_half_grid", 2, True, B=2, C=64, Co=64, size=(8, 9), sigma=0.5, seed=55),
    Case("tc_m2d_all_oob", 2, True, B=1, C=64, Co=64, size=(8, 8), offset_bias=1e4, seed=56),
    Case("tc_m2d_g2", 2, True, B=2, C=128, Co=128, size=(10, 12), g=2, dg=2, seed=57),
    Case("tc_d2d_c64", 2, False, B=2, C=64, Co=64, size=(12, 12), seed=60),
    Case("tc_m2d_c256_multi_tile", 2, True, B=2, C=256, Co=256, size=(24, 20), bias=False, seed=63),
    Case("tc_d2d_integer_grid", 2, False, B=2, C=64, Co=64, size=(8, 9), sigma=0.0, seed=61),
    Case("tc_d2d_k1x3", 2, False, B=2, C=64, Co=64, size=(8, 9), k=(1, 3), p=(0, 1), seed=62),
    Case("tc_m3d_c64", 3, True, B=1, C=64, Co=64, size=(4, 6, 5), k=(3, 3, 3), s=(1, 1, 1), p=(1, 1, 1),
         d=(1, 1, 1), seed=70),
    Case("tc_d3d_c64_half_grid", 3, False, B=1, C=64, Co=64, size=(4, 5, 6), k=(3, 3, 3), s=(1, 1, 1),
         p=(1, 1, 1), d=(1, 1, 1), sigma=0.5, seed=71),
    # deformable-group / group nesting of the backward kernels
    Case("tc_m2d_dg4_c256", 2, True, B=2, C=256, Co=128, size=(8, 8), dg=4, seed=72),
    Case("tc_m2d_g2_dg1", 2, True, B=2, C=128, Co=128, size=(8, 8), g=2, dg=1, seed=73),
    Case("tc_m2d_g4_dg2", 2, True, B=1, C=256, Co=256, size=(8, 8), g=4, dg=2, seed=74),
    Case("tc_d2d_c192_dg3", 2, False, B=2, C=192, Co=64, size=(8, 8), dg=3, seed=75),
    Case("tc_m2d_c192_g3_dg1", 2, True, B=1, C=192, Co=192, size=(8, 8), g=3, dg=1, seed=76),
    # Cog = 32: forward + grad_offset/grad_mask on tensor cores, grad_input falls back
    Case("tc_m2d_co32", 2, True, B=2, C=64, Co=32, size=(8, 8), seed=77),
    # ragged tails: B*P not a multiple of the 128-row tile, P not a multiple of 64
    Case("tc_m2d_ragged", 2, True, B=3, C=64, Co=64, size=(10, 12), seed=78),
    Case("tc_m2d_ragged_3chunks", 2, True, B=3, C=128, Co=64, size=(22, 19), seed=86),   # P = 418: not a multiple of 8
    Case("tc_m2d_sigma8", 2, True, B=1, C=64, Co=64, size=(16, 16), sigma=8.0, seed=79),
    # more than one output-channel tile (Co = 512 > 256) and eight channel blocks (C = 512)
    Case("tc_m2d_co512", 2, True, B=1, C=64, Co=512, size=(8, 8), seed=81),
    Case("tc_d2d_c512_dg2", 2, False, B=1, C=512, Co=64, size=(8, 8), dg=2, seed=82),
    # 1024 input channels: two 512-channel slabs per position tile in the col2im gather of grad_input
    Case("tc_m2d_c1024_slabs", 2, True, B=2, C=1024, Co=64, size=(6, 7), dg=2, seed=85),
    # 1x1 kernel with stride 2 (a single tap), and a 5x5 kernel (25 taps, dilation 1, pad 2)
    Case("tc_m2d_k1_s2", 2, True, B=2, C=64, Co=64, size=(9, 10), k=(1, 1), s=(2, 2), p=(0, 0), seed=83),
    Case("tc_m2d_k5", 2, True, B=1, C=64, Co=64, size=(8, 8), k=(5, 5), p=(2, 2), seed=84),
    Case("tc_m3d_dg2_stride2", 3, True, B=1, C=128, Co=64, size=(6, 8, 8), k=(3, 3, 3), s=(2, 2, 2), p=(1, 1, 1),
         d=(1, 1, 1), dg=2, seed=80),
]

GENERIC_OFFSET_CASES = [c for c in ALL_CASES if c.sigma not in (0.0, 0.5)]


def rel_err(a, ref):
    """max|a - ref| / max|ref| (1 if ref is all-zero and a is not)."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.abs(ref).max() if ref.size else 0.0
    num = np.abs(a - ref).max() if ref.size else 0.0
    if den == 0.0:
        return 0.0 if num == 0.0 else float("inf")
    return float(num / den)
