"""Restatement of the reference's sliding-crop inference (model.no_cropping=False).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The arithmetic lives in numpy (float32 element-wise ops),
OpenCV (`cv2.resize(..., INTER_LINEAR)`, an un-vendored dependency of flow/transform.py:6) and torch (softmax,
fp64 adds): the functions below re-express the reference's call sequence over those same calls and cite the
lines they follow.  `resize_linear_np` additionally restates cv2's published 32F bilinear algorithm in plain
numpy so that the formula the CUDA kernel implements is pinned against cv2 itself (tests/test_crop_oracle.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def block_window(height, width, grid_h, grid_w, crop_h, crop_w, h_off, w_off):
    """flow/transform.py:223-235 — which grid blocks a crop covers (python double arithmetic, python round())."""
    ppb_h, ppb_w = height / grid_h, width / grid_w
    bh_off, bw_off = round(h_off / ppb_h), round(w_off / ppb_w)
    bh = round((h_off + crop_h) / ppb_h) - bh_off
    bw = round((w_off + crop_w) / ppb_w) - bw_off
    return dict(ppb_h=ppb_h, ppb_w=ppb_w, bh_off=bh_off, bw_off=bw_off, bh=bh, bw=bw, oh=crop_h // 16, ow=crop_w // 16)


def crop_grid_numpy(m, height, width, crop_h, crop_w, h_off, w_off, resize=None):
    """flow/transform.py:238-244 (_crop_motion_vector) on one [Hg,Wg,2] numpy grid; returns [crop_h//16, crop_w//16, 2].

    `resize` defaults to cv2.resize(INTER_LINEAR) exactly as the reference calls it."""
    g = block_window(height, width, m.shape[0], m.shape[1], crop_h, crop_w, h_off, w_off)
    m = m[g["bh_off"]:g["bh_off"] + g["bh"], g["bw_off"]:g["bw_off"] + g["bw"]].copy()   # the reference works on a host copy
    m[:, :, 0] = ((((m[:, :, 0] + 1) / 2) * width - w_off) / (g["bw"] * g["ppb_w"])) * 2 - 1
    m[:, :, 1] = ((((m[:, :, 1] + 1) / 2) * height - h_off) / (g["bh"] * g["ppb_h"])) * 2 - 1
    if resize is None:
        import cv2
        return cv2.resize(m, (g["ow"], g["oh"]), interpolation=cv2.INTER_LINEAR)
    return resize(m, g["ow"], g["oh"])


def resize_linear_np(m, ow, oh):
    """cv2.resize(m, (ow, oh), INTER_LINEAR) for float32 [h,w,c], restated from OpenCV's resize.cpp:
    fx = float((dx + 0.5) * scale - 0.5) with scale = 1 / (dsize / ssize) in double; sx = floor(fx); the fraction is
    zeroed where sx is clamped; rows are fl(fl(S0*(1-fx)) + fl(S1*fx)), then the same vertically."""
    ih, iw = m.shape[:2]

    def coords(o, i):
        scale = 1.0 / (o / i)
        idx = np.zeros(o, np.int64)
        frac = np.zeros(o, np.float32)
        for d in range(o):
            fx = np.float32((d + 0.5) * scale - 0.5)
            s = int(math.floor(fx))
            fx = np.float32(fx - np.float32(s))
            if s < 0:
                s, fx = 0, np.float32(0)
            if s >= i - 1:
                s, fx = i - 1, np.float32(0)
            idx[d], frac[d] = s, fx
        return idx, frac

    xi, xf = coords(ow, iw)
    yi, yf = coords(oh, ih)
    xi1, yi1 = np.minimum(xi + 1, iw - 1), np.minimum(yi + 1, ih - 1)
    m = m.astype(np.float32)
    a0, a1 = (np.float32(1) - xf)[None, :, None], xf[None, :, None]
    rows = (m[:, xi, :] * a0).astype(np.float32) + (m[:, xi1, :] * a1).astype(np.float32)
    b0, b1 = (np.float32(1) - yf)[:, None, None], yf[:, None, None]
    return ((rows[yi] * b0).astype(np.float32) + (rows[yi1] * b1).astype(np.float32)).astype(np.float32)


def crop_motion_vector(mvs_left, mvs_right, height, width, crop_h, crop_w, h_off, w_off, resize=None):
    """flow/transform.py:215-261 for lists of [1,Hg,Wg,2] tensors: every grid goes device -> host numpy -> cv2 -> device."""
    def is_grid_list(m):                                   # :216-221 — [B,1] dummies of no_warp clips pass through
        return m is not None and isinstance(m, list) and len(m) > 0 and len(m[0].shape) >= 3
    if not (is_grid_list(mvs_left) or is_grid_list(mvs_right)):
        return mvs_left, mvs_right

    def one(t):
        out = crop_grid_numpy(t.cpu().numpy()[0], height, width, crop_h, crop_w, h_off, w_off, resize)
        return torch.from_numpy(np.ascontiguousarray(out)).unsqueeze(0).to(t.device)
    return [one(t) for t in mvs_left], [one(t) for t in mvs_right]


def crop_windows(new_h, new_w, crop_h, crop_w, stride_rate=2 / 3):
    """flow/base.py:183-200 — the (s_h, e_h, s_w, e_w) windows in the order the reference visits them."""
    stride_h, stride_w = int(np.ceil(crop_h * stride_rate)), int(np.ceil(crop_w * stride_rate))
    grid_h = int(np.ceil(float(new_h - crop_h) / stride_h) + 1)
    grid_w = int(np.ceil(float(new_w - crop_w) / stride_w) + 1)
    for ih in range(grid_h):
        for iw in range(grid_w):
            e_h = min(ih * stride_h + crop_h, new_h)
            e_w = min(iw * stride_w + crop_w, new_w)
            yield e_h - crop_h, e_h, e_w - crop_w, e_w


def crop_softmax(output, h_i, w_i):
    """flow/base.py:215-221 / 229-234 — resize to the crop size if needed, then softmax over classes."""
    if output.shape[2] != h_i or output.shape[3] != w_i:
        output = F.interpolate(output, (h_i, w_i), mode="bilinear", align_corners=True)
    return F.softmax(output, dim=1)


def compute_output(n, function, frame_prev, frame_next, mvs_left, mvs_right, classes, crop_h, crop_w):
    """flow/base.py:182-209 — fp64 canvas of averaged class probabilities, [n,classes,H,W]."""
    _, _, new_h, new_w = frame_prev.shape
    prediction = torch.zeros((n, classes, new_h, new_w), dtype=float, device=frame_prev.device)
    count = torch.zeros((new_h, new_w), dtype=float, device=frame_prev.device)
    for s_h, e_h, s_w, e_w in crop_windows(new_h, new_w, crop_h, crop_w):
        prev_c = frame_prev[:, :, s_h:e_h, s_w:e_w].clone()
        next_c = frame_next[:, :, s_h:e_h, s_w:e_w].clone()
        ml, mr = crop_motion_vector(mvs_left, mvs_right, new_h, new_w, e_h - s_h, e_w - s_w, s_h, s_w)
        count[s_h:e_h, s_w:e_w] += 1
        prediction[:, :, s_h:e_h, s_w:e_w] += function(prev_c, next_c, ml, mr)
    prediction /= count.unsqueeze(0).unsqueeze(0)
    return prediction
