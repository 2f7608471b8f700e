"""ctypes wrapper of oracle/c/fuvs_oracle.c (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfuvs_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "c", "fuvs_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "c")], check=True)
        _lib = C.CDLL(_SO)
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.c_void_p)


def grid_sample(src, grid, align_corners=False):
    """src [C,Hin,Win], grid [Hg,Wg,2] -> [C,Hg,Wg]"""
    src, ps = _f(src)
    grid, pg = _f(grid)
    c, hin, win = src.shape
    hg, wg = grid.shape[:2]
    dst = np.empty((c, hg, wg), np.float32)
    lib().fo_grid_sample(ps, pg, dst.ctypes.data_as(C.c_void_p), c, hin, win, hg, wg, int(align_corners))
    return dst


def upsample_ac(src, size):
    src, ps = _f(src)
    hin, win = src.shape[-2:]
    planes = int(np.prod(src.shape[:-2])) if src.ndim > 2 else 1
    dst = np.empty(src.shape[:-2] + (size[0], size[1]), np.float32)
    lib().fo_upsample_ac(ps, dst.ctypes.data_as(C.c_void_p), C.c_longlong(planes), hin, win, size[0], size[1])
    return dst


def blend(a, b, wa, wb):
    a, pa = _f(a)
    out = np.empty_like(a)
    if b is not None:
        b, pb = _f(b)
    else:
        pb = None
    lib().fo_blend(pa, pb, C.c_double(wa), C.c_double(wb), out.ctypes.data_as(C.c_void_p), C.c_longlong(a.size))
    return out


def argmax(logits):
    """[F,C,H,W] -> uint8 [F,H,W]"""
    logits, pl = _f(logits)
    f, c, h, w = logits.shape
    out = np.empty((f, h, w), np.uint8)
    lib().fo_argmax(pl, f, c, C.c_longlong(h * w), out.ctypes.data_as(C.c_void_p))
    return out


def counts(pred, target, K, ignore_index=255, numpy_bins=False, out=None):
    p = np.ascontiguousarray(pred, dtype=np.int64).reshape(-1)
    t = np.ascontiguousarray(target, dtype=np.int64).reshape(-1)
    if out is None:
        out = np.zeros((3, K), np.int64)
    lib().fo_counts(p.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), C.c_longlong(p.size), K,
                    C.c_longlong(ignore_index), int(numpy_bins), out.ctypes.data_as(C.c_void_p))
    return out


def temporal(labels, K, ignore_index=255, tc_prev=None, out=None):
    labels = np.ascontiguousarray(labels, dtype=np.uint8)
    n, h, w = labels.shape
    if out is None:
        out = np.zeros((3, K), np.int64)
    pp = None
    if tc_prev is not None:
        tc_prev = np.ascontiguousarray(tc_prev, dtype=np.uint8)
        pp = tc_prev.ctypes.data_as(C.c_void_p)
    lib().fo_temporal(labels.ctypes.data_as(C.c_void_p), n, C.c_longlong(h * w), pp, K, C.c_longlong(ignore_index),
                      out.ctypes.data_as(C.c_void_p))
    return out


def interval(prev, nxt, grids_left, grids_right, n, warp, want_logits=True):
    """prev/next [C,H,W]; grids [n-1,Hg,Wg,2] (ignored when warp is False) -> (logits [n,C,H,W] | None, labels [n,H,W])"""
    prev, pp = _f(prev)
    c, h, w = prev.shape
    pn = pl = pr = None
    hg = wg = 0
    if n > 1:
        nxt, pn = _f(nxt)
        if warp:
            grids_left, pl = _f(grids_left)
            grids_right, pr = _f(grids_right)
            hg, wg = grids_left.shape[1:3]
    logits = np.empty((n, c, h, w), np.float32) if want_logits else None
    labels = np.empty((n, h, w), np.uint8)
    lib().fo_interval(pp, pn, pl, pr, c, h, w, hg, wg, n, int(bool(warp)),
                      logits.ctypes.data_as(C.c_void_p) if want_logits else None, labels.ctypes.data_as(C.c_void_p))
    return logits, labels
