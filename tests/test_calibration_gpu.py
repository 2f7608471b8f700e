"""Proves on the GPU which floating-point contraction torch's ATen binaries use for grid_sampler_2d and
upsample_bilinear2d, and that it is exactly the one compiled into libfuvs (csrc/fuvs_common.cuh `Nm`).

Every candidate contraction is run through the calibration entry points (include/fuvs_calib.h) on inputs that
exercise border clipping and near-integer coordinates and compared bit-for-bit with torch-CUDA."""
import itertools
import json
import os

import pytest
import torch
import torch.nn.functional as F

from flood_uav_video_segmentation_b200 import _lib
from flood_uav_video_segmentation_b200.synthetic import flow_grids, keyframe_logits

pytestmark = pytest.mark.gpu


def _bits_equal(a, b):
    return bool((a.view(torch.int32) == b.view(torch.int32)).all())


def _mismatch(a, b):
    return int((a.view(torch.int32) != b.view(torch.int32)).sum())


GS_CASES = [  # (C, Hin, Win, grid mode, H for grid, W for grid, align)
    (5, 433, 433, "block", 433, 433, False),
    (5, 270, 480, "dense", 270, 480, False),
    (3, 67, 120, "block", 1072, 1920, False),     # low-res source sampled by a 67x120 grid (chain step >= 2)
    (4, 135, 240, "block", 1072, 1920, True),      # default-grid resample of flow/model.py:157
]


def test_grid_sample_contraction(cuda, out_dir):
    lib = _lib.load()
    default = lib.fuvs_calib_default()
    d_un, d_tap = default & 1, (default >> 1) & 1
    report = {}
    matches_all = {(u, t): True for u in (0, 1) for t in (0, 1)}
    for ci, (C, Hin, Win, mode, Hg_src, Wg_src, align) in enumerate(GS_CASES):
        src = torch.randn(1, C, Hin, Win, generator=torch.Generator().manual_seed(ci), dtype=torch.float32).to(cuda)
        grid = flow_grids(Hg_src, Wg_src, 2, mode, clip=ci, jitter=0.08)[0].to(cuda)   # jitter crosses the border
        ref = F.grid_sample(src, grid, mode="bilinear", padding_mode="border", align_corners=align)
        Hg, Wg = grid.shape[1:3]
        for (u, t) in matches_all:
            dst = torch.empty(C, Hg, Wg, dtype=torch.float32, device=cuda)
            _lib.check(lib.fuvs_calib_grid_sample(src.data_ptr(), grid.data_ptr(), dst.data_ptr(), C, Hin, Win, Hg, Wg,
                                                  int(align), u, t, _lib.stream_ptr(cuda)))
            torch.cuda.synchronize()
            bad = _mismatch(dst, ref[0])
            report[f"gs case{ci} unnorm_fma={u} tap_fma={t}"] = bad
            matches_all[(u, t)] &= bad == 0
    with open(os.path.join(out_dir, "calibration_grid_sample.json"), "w") as f:
        json.dump({"default": [d_un, d_tap], "mismatching_elements": report}, f, indent=1)
    winners = [k for k, v in matches_all.items() if v]
    assert (d_un, d_tap) in winners, f"compiled-in grid_sample numerics {(d_un, d_tap)} do not match torch; matching: {winners}; {report}"


UP_CASES = [(5, 67, 120, 1072, 1920), (5, 27, 27, 433, 433), (3, 55, 55, 433, 433), (2, 135, 240, 1080, 1920), (4, 9, 7, 10, 8)]


def test_upsample_contraction(cuda, out_dir):
    lib = _lib.load()
    default = lib.fuvs_calib_default()
    d = ((default >> 2) & 1, (default >> 3) & 3, (default >> 5) & 3)
    combos = list(itertools.product((0, 1), (0, 1, 2), (0, 1, 2)))
    ok = {c: True for c in combos}
    report = {}
    for ci, (C, Hin, Win, Hout, Wout) in enumerate(UP_CASES):
        src = torch.randn(1, C, Hin, Win, generator=torch.Generator().manual_seed(100 + ci), dtype=torch.float32).to(cuda)
        ref = F.interpolate(src, size=(Hout, Wout), mode="bilinear", align_corners=True)
        for c in combos:
            dst = torch.empty(C, Hout, Wout, dtype=torch.float32, device=cuda)
            _lib.check(lib.fuvs_calib_upsample(src.data_ptr(), dst.data_ptr(), C, Hin, Win, Hout, Wout, c[0], c[1], c[2],
                                               _lib.stream_ptr(cuda)))
            torch.cuda.synchronize()
            bad = _mismatch(dst, ref[0])
            report[f"up case{ci} lambda_fma={c[0]} inner={c[1]} outer={c[2]}"] = bad
            ok[c] &= bad == 0
    with open(os.path.join(out_dir, "calibration_upsample.json"), "w") as f:
        json.dump({"default": list(d), "mismatching_elements": report}, f, indent=1)
    winners = [k for k, v in ok.items() if v]
    assert d in winners, f"compiled-in upsample numerics {d} do not match torch; matching: {winners}"


def test_mul_add_are_separately_rounded(cuda):
    """flow/model.py:234-236 on torch-CUDA: w0*a, w1*b and the add each round to fp32 (no FMA across launches)."""
    a = keyframe_logits(5, 64, 64, 0, 0).to(cuda)
    b = keyframe_logits(5, 64, 64, 0, 1).to(cuda)
    n, p = 5, 2
    ref = (n - p) / n * a + p / n * b
    w0 = torch.tensor((n - p) / n, dtype=torch.float32, device=cuda)
    w1 = torch.tensor(p / n, dtype=torch.float32, device=cuda)
    assert _bits_equal(ref, (a * w0) + (b * w1))
