"""Restatement of the reference's interpolation call sequence over torch ops.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines it follows (paths relative to the reference repo root).  The
same code runs on torch-CPU (CPU baseline) and torch-CUDA (bit-exact
authority for the kernels on the same GPU).
"""
from __future__ import annotations

import contextlib

import numpy as np
import torch
import torch.nn.functional as F


class NullProfiler:
    """Stand-in for Lightning's profiler: the reference calls profiler.profile(name)
    as a context manager (flow/model.py:119,134,165,176,188,213,232)."""

    def __init__(self):
        self.regions = []

    @contextlib.contextmanager
    def profile(self, name):
        self.regions.append(name)
        yield


def default_grid(width=1920, height=1072, block=16):
    """flow/model.py:10-21 — identity grid at macro-block centres, fp64 [H/16, W/16, 2]."""
    bh, bw = height // block, width // block
    xs = (np.arange(bw, dtype=np.float64) * block + block // 2) / width * 2 - 1
    ys = (np.arange(bh, dtype=np.float64) * block + block // 2) / height * 2 - 1
    g = np.zeros((bh, bw, 2))
    g[:, :, 0] = xs[None, :]
    g[:, :, 1] = ys[:, None]
    return g


def warp(frame, grid, no_warp=False):
    """flow/model.py:244-249."""
    if no_warp:
        return frame
    if grid.dtype != torch.float32:
        grid = grid.float()
    return F.grid_sample(frame, grid, mode="bilinear", padding_mode="border", align_corners=False)


def _to_size(x, h, w):
    """The recurring `if shape != (h,w): interpolate(bilinear, align_corners=True)` of
    flow/model.py:41-42,67-68,85-86,138-139,149-150,158-159,178-179,192-193,205-206,217-218,227-228."""
    if x.shape[2] != h or x.shape[3] != w:
        x = F.interpolate(x, size=(h, w), mode="bilinear", align_corners=True)
    return x


def predict_segmentation(encoder, decoder, frame_prev, frame_next, mvs_left, mvs_right, n, no_warp=False):
    """flow/model.py:184-241 -> [n,C,h,w] logits ("pred")."""
    h, w = frame_prev.shape[2], frame_prev.shape[3]
    o = _to_size(decoder(encoder(frame_prev)), h, w)                      # :188-193
    maps = [o]
    if frame_next is not None:
        o_next = _to_size(decoder(encoder(frame_next)), h, w)              # :201-206
        fwd, bwd = [], []
        cur = o
        for m in mvs_left:                                                 # :212-219
            cur = warp(cur, m, no_warp)
            fwd.append(_to_size(cur, h, w))
        cur = o_next
        for m in mvs_right:                                                # :222-229
            cur = warp(cur, m, no_warp)
            bwd.append(_to_size(cur, h, w))
        for p in range(1, n):                                              # :233-237
            a = (n - p) / n * fwd[p - 1]
            b = p / n * bwd[n - p - 1]
            maps.append(a + b)
    return torch.cat(maps, 0)                                              # :239


def predict_feature(encoder, decoder, frame_prev, frame_next, mvs_left, mvs_right, n, default_mv, no_warp=False):
    """flow/model.py:116-181 -> [n,C,h,w] logits."""
    h, w = frame_prev.shape[2], frame_prev.shape[3]
    f = encoder(frame_prev)                                                # :120
    fh, fw = f.shape[2], f.shape[3]
    fwd, bwd = [], []
    f_next = None
    if frame_next is not None:
        f_next = encoder(frame_next)                                       # :129
        if not no_warp:
            cur = f
            for m in mvs_left:                                             # :133-140
                cur = warp(cur, m)
                fwd.append(_to_size(cur, fh, fw))
            cur = f_next
            for m in mvs_right:                                            # :144-151
                cur = warp(cur, m)
                bwd.append(_to_size(cur, fh, fw))
    if not no_warp:                                                        # :154-159 (align_corners=True, sic)
        f = F.grid_sample(f, default_mv.to(f.device), padding_mode="border", align_corners=True)
        f = _to_size(f, fh, fw)
    maps = [f]
    if frame_next is not None:
        for p in range(1, n):                                              # :166-171
            if not no_warp:
                maps.append((n - p) / n * fwd[p - 1] + p / n * bwd[n - p - 1])
            else:
                maps.append((n - p) / n * f + p / n * f_next)
    out = decoder(torch.cat(maps, 0))                                      # :173-177
    return _to_size(out, h, w)                                             # :178-179


def warp_batch(x, mvs, index_list, n_list, no_warp=False):
    """flow/model.py:92-106, including the :102 quirk (tests shape[1], shape[2] of a 4-D tensor)."""
    ih, iw = x.shape[2], x.shape[3]
    outs = []
    for i, index in enumerate(index_list):
        cur = x[i].unsqueeze(0)
        if not no_warp:
            for j in range(index):
                cur = warp(cur, mvs[j][i].unsqueeze(0))
            if cur.shape[1] != ih or cur.shape[2] != iw:
                cur = F.interpolate(cur, size=(ih, iw), mode="bilinear", align_corners=True)
        outs.append(cur * ((n_list[i] - index) / n_list[i]))
    return torch.cat(outs)


def forward_segmentation(encoder, decoder, frame_prev, frame_next, mvs_left, mvs_right, left_index, right_index,
                         no_warp=False):
    """flow/model.py:46-48 + 73-88 (eval branch of forward)."""
    li = [int(i) for i in left_index]
    ri = [int(i) for i in right_index]
    nl = [a + b for a, b in zip(li, ri)]
    h, w = frame_prev.shape[2], frame_prev.shape[3]
    o_prev = decoder(encoder(frame_prev))
    o_next = decoder(encoder(frame_next))
    o = warp_batch(o_prev, mvs_left, li, nl, no_warp) + warp_batch(o_next, mvs_right, ri, nl, no_warp)
    return _to_size(o, h, w)


def forward_feature(encoder, decoder, frame_prev, frame_next, mvs_left, mvs_right, left_index, right_index,
                    no_warp=False):
    """flow/model.py:46-48 + 55-70."""
    li = [int(i) for i in left_index]
    ri = [int(i) for i in right_index]
    nl = [a + b for a, b in zip(li, ri)]
    h, w = frame_prev.shape[2], frame_prev.shape[3]
    f = warp_batch(encoder(frame_prev), mvs_left, li, nl, no_warp) + warp_batch(encoder(frame_next), mvs_right, ri, nl,
                                                                               no_warp)
    return _to_size(decoder(f), h, w)


def argmax_labels(logits):
    """flow/base.py:147,167,276 — output.data.max(1)[1] (int64)."""
    return logits.data.max(1)[1]
