"""Tensor-level wrappers over the C ABI (include/fuvs.h).

Each function validates shapes/dtypes, allocates outputs with torch (device
memory plumbing only), passes raw pointers + the current CUDA stream to
libfuvs.so and returns tensors.  Nothing here computes on the host.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import FuvsError, check, load, ptr, ptr_array, require_cuda, stream_ptr


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise FuvsError(f"{what}: expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def new_counts(K: int, device) -> torch.Tensor:
    """Device-resident int64 [3,K] accumulator: rows (intersection, union, target)."""
    return torch.zeros((3, K), dtype=torch.int64, device=device)


def _grid_list(mvs, n, shape, dev, what: str):
    """The reference's grid format — a python list of n-1 tensors [1,Hg,Wg,2] (flow/dataset.py:138-146), or one stacked
    [n-1,Hg,Wg,2] tensor — as n-1 contiguous fp32 CUDA tensors [Hg,Wg,2] WITHOUT copying: the *_ptrs entries take a
    host array of device pointers.  Only what the reference itself converts is converted (FlowModel.warp casts
    non-float grids with .float(), flow/model.py:246-247); a non-contiguous grid is made contiguous."""
    if isinstance(mvs, torch.Tensor):
        mvs = [mvs[j] for j in range(mvs.shape[0])]
    if len(mvs) != n - 1:
        raise FuvsError(f"{what}: n={n} needs {n - 1} grids per side, got {len(mvs)}")
    out = []
    for m in mvs:
        if not isinstance(m, torch.Tensor) or m.dim() < 3 or m.shape[-1] != 2:
            raise FuvsError(f"{what}: a grid must be a tensor [..,Hg,Wg,2], got {getattr(m, 'shape', type(m))}")
        g = m.reshape(m.shape[-3], m.shape[-2], 2)
        if shape is not None and tuple(g.shape[:2]) != tuple(shape):
            raise FuvsError(f"{what}: grids must be [{shape[0]},{shape[1]},2], got {tuple(g.shape)}")
        if g.dtype != torch.float32:
            g = g.float()
        if not g.is_contiguous():
            g = g.contiguous()
        out.append(g)
    require_cuda(*out, what=what)
    if out and out[0].device != dev:
        raise FuvsError(f"{what}: grids on {out[0].device}, key frames on {dev}")
    return out


class ScratchCache:
    """Per-owner workspace for the interval entries (chain states): grown on demand, reused across calls, so that an
    evaluation loop does not allocate 249 MB per dense interval.  Owned by the caller (FlowModel keeps one)."""

    def __init__(self):
        self.buf = None

    def get(self, floats: int, dev) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < floats or self.buf.device != dev:
            self.buf = torch.empty((max(int(floats), 1),), dtype=torch.float32, device=dev)
        return self.buf

    def clear(self):
        self.buf = None


def _scratch(scratch, need, dev):
    if isinstance(scratch, ScratchCache):
        return scratch.get(need, dev)
    if scratch is None or scratch.numel() < need:
        return torch.empty((max(need, 1),), dtype=torch.float32, device=dev)
    return scratch


def linear_blend_argmax(prev, nxt, n, *, want_labels=True, want_logits=False, tc_prev=None, counts=None,
                        ignore_index=255):
    """fuvs_linear_blend_argmax.  prev/nxt: [C,H,W] or [1,C,H,W]."""
    dev = require_cuda(prev, nxt, tc_prev, counts, what="linear_blend_argmax")
    prev = _f32c(prev, "prev")
    C, H, W = prev.shape[-3:]
    if n > 1:
        nxt = _f32c(nxt, "next")
        if nxt.shape[-3:] != prev.shape[-3:]:
            raise FuvsError(f"linear_blend_argmax: key frames differ in shape {tuple(prev.shape)} vs {tuple(nxt.shape)}")
    labels = torch.empty((n, H, W), dtype=torch.uint8, device=dev) if want_labels else None
    logits = torch.empty((n, C, H, W), dtype=torch.float32, device=dev) if want_logits else None
    _check_tc(tc_prev, counts, H, W, C)
    with torch.cuda.device(dev):
        check(load().fuvs_linear_blend_argmax(ptr(prev), ptr(nxt) if n > 1 else None, C, H, W, n, ptr(labels),
                                              ptr(logits), ptr(tc_prev), ptr(counts), ignore_index, stream_ptr(dev)))
    return labels, logits


def linear_lowres_blend_argmax(prev_lr, nxt_lr, size, n, *, want_labels=True, want_logits=False, tc_prev=None,
                               counts=None, ignore_index=255):
    """fuvs_linear_lowres_blend_argmax: key frames at decoder resolution [C,hl,wl] / [1,C,hl,wl], output size (H,W).

    Shapes the fused kernel does not take (C > 5, W % 4 != 0) go through fuvs_upsample_bilinear_ac +
    fuvs_linear_blend_argmax — the same arithmetic in two launches."""
    dev = require_cuda(prev_lr, nxt_lr, tc_prev, counts, what="linear_lowres_blend_argmax")
    prev_lr = _f32c(prev_lr, "prev_lr")
    C, hl, wl = prev_lr.shape[-3:]
    H, W = int(size[0]), int(size[1])
    if n > 1:
        nxt_lr = _f32c(nxt_lr, "next_lr")
        if nxt_lr.shape[-3:] != prev_lr.shape[-3:]:
            raise FuvsError("linear_lowres_blend_argmax: key frames differ in shape")
    if not load().fuvs_linear_lowres_supported(C, H, W) or (counts is not None and 0 <= ignore_index < C):
        up = upsample_bilinear_ac(prev_lr.reshape(1, C, hl, wl), (H, W))
        up_n = upsample_bilinear_ac(nxt_lr.reshape(1, C, hl, wl), (H, W)) if n > 1 else None
        return linear_blend_argmax(up, up_n, n, want_labels=want_labels, want_logits=want_logits, tc_prev=tc_prev,
                                   counts=counts, ignore_index=ignore_index)
    labels = torch.empty((n, H, W), dtype=torch.uint8, device=dev) if want_labels else None
    logits = torch.empty((n, C, H, W), dtype=torch.float32, device=dev) if want_logits else None
    _check_tc(tc_prev, counts, H, W, C)
    with torch.cuda.device(dev):
        check(load().fuvs_linear_lowres_blend_argmax(ptr(prev_lr), ptr(nxt_lr) if n > 1 else None, C, hl, wl, H, W, n,
                                                     ptr(labels), ptr(logits), ptr(tc_prev), ptr(counts), ignore_index,
                                                     stream_ptr(dev)))
    return labels, logits


def _check_tc(tc_prev, counts, H, W, K):
    if tc_prev is not None:
        if tc_prev.dtype != torch.uint8 or tuple(tc_prev.shape[-2:]) != (H, W) or not tc_prev.is_contiguous():
            raise FuvsError("tc_prev must be a contiguous uint8 [H,W] label map")
    if counts is not None:
        if counts.dtype != torch.int64 or tuple(counts.shape) != (3, K) or not counts.is_contiguous():
            raise FuvsError(f"counts must be a contiguous int64 [3,{K}] tensor")


def warp_step(src0, grid0, src1=None, grid1=None, *, align_corners=False):
    """fuvs_warp_step: F.grid_sample(bilinear, border) of one or two [C,Hin,Win] maps."""
    dev = require_cuda(src0, grid0, src1, grid1, what="warp_step")
    src0 = _f32c(src0, "src0")
    grid0 = _f32c(grid0 if grid0.dtype == torch.float32 else grid0.float(), "grid0")
    C, Hin, Win = src0.shape[-3:]
    Hg, Wg = grid0.shape[-3:-1]
    dst0 = torch.empty((C, Hg, Wg), dtype=torch.float32, device=dev)
    dst1 = None
    if src1 is not None:
        src1 = _f32c(src1, "src1")
        grid1 = _f32c(grid1 if grid1.dtype == torch.float32 else grid1.float(), "grid1")
        if src1.shape[-3:] != src0.shape[-3:] or grid1.shape[-3:] != grid0.shape[-3:]:
            raise FuvsError("warp_step: the two problems must have the same shapes")
        dst1 = torch.empty_like(dst0)
    with torch.cuda.device(dev):
        check(load().fuvs_warp_step(ptr(src0), ptr(grid0), ptr(dst0), ptr(src1), ptr(grid1), ptr(dst1), C, Hin, Win,
                                    Hg, Wg, 1 if align_corners else 0, stream_ptr(dev)))
    return dst0, dst1


def dense_interval(prev, nxt, grids_left, grids_right, n, *, want_labels=True, want_logits=False, tc_prev=None,
                   counts=None, ignore_index=255, scratch=None):
    """fuvs_dense_interval_ptrs.  grids_*: list of n-1 [1,H,W,2] tensors (the reference's format, passed as a pointer
    table: nothing is stacked) or a stacked [n-1,H,W,2].  scratch: tensor, ScratchCache or None (allocated per call)."""
    dev = require_cuda(prev, nxt, tc_prev, counts, what="dense_interval")
    prev = _f32c(prev, "prev")
    C, H, W = prev.shape[-3:]
    gl = gr = None
    if n > 1:
        nxt = _f32c(nxt, "next")
        gl = _grid_list(grids_left, n, (H, W), dev, "dense_interval(grids_left)")
        gr = _grid_list(grids_right, n, (H, W), dev, "dense_interval(grids_right)")
    scratch = _scratch(scratch, int(load().fuvs_dense_scratch_floats(C, H, W, n)), dev)
    labels = torch.empty((n, H, W), dtype=torch.uint8, device=dev) if (want_labels or counts is not None) else None
    logits = torch.empty((n, C, H, W), dtype=torch.float32, device=dev) if want_logits else None
    _check_tc(tc_prev, counts, H, W, C)
    with torch.cuda.device(dev):
        check(load().fuvs_dense_interval_ptrs(ptr(prev), ptr(nxt) if n > 1 else None, ptr_array(gl), ptr_array(gr), C, H, W,
                                              n, ptr(scratch), ptr(labels), ptr(logits), ptr(tc_prev), ptr(counts),
                                              ignore_index, stream_ptr(dev)))
    return labels, logits


class KeyFrameUps:
    """The two caller-owned buffers of fuvs_dense_lowres_interval_ptrs (up-sampled key frames, library-private layout) and
    the bookkeeping of which decoder-resolution key frame each one holds, so that an interval's `next` is reused as the
    following interval's `prev` (flow/model.py:189,202 would up-sample it twice)."""

    def __init__(self):
        self.bufs = [None, None]
        self.tags = [None, None]

    def _buffer(self, i, numel, dev):
        b = self.bufs[i]
        if b is None or b.numel() != numel or b.device != dev:
            self.bufs[i] = torch.empty((numel,), dtype=torch.float32, device=dev)
            self.tags[i] = None
        return self.bufs[i]

    def reset(self):
        self.tags = [None, None]


def _keyframe_tag(t, size, n, outputs):
    # the tensor OBJECT (kept alive, so its storage cannot be handed to another tensor) and its version counter;
    # inference tensors (torch.inference_mode, Lightning's predict loop) do not track one: identity alone then
    try:
        version = t._version
    except RuntimeError:
        version = -1
    return (t, version, tuple(size), n, outputs)


def _same_tag(a, b):
    return a is not None and b is not None and a[0] is b[0] and a[1:] == b[1:]


def dense_lowres_interval(prev_lr, nxt_lr, size, grids_left, grids_right, n, *, want_labels=True, want_logits=False,
                          tc_prev=None, counts=None, ignore_index=255, scratch=None, ups=None):
    """fuvs_dense_lowres_interval_ptrs: key frames at decoder resolution [C,hl,wl] / [1,C,hl,wl], frame size `size`.
    ups: a KeyFrameUps that lives across the intervals of a clip (None: buffers allocated per call, nothing reused): when
    prev_lr is the very tensor (same storage, same version) that was `nxt_lr` of the previous call, its up-sample is
    reused instead of recomputed."""
    dev = require_cuda(prev_lr, nxt_lr, tc_prev, counts, what="dense_lowres_interval")
    prev_obj, next_obj = prev_lr, nxt_lr
    prev_lr = _f32c(prev_lr, "prev_lr")
    C, hl, wl = prev_lr.shape[-3:]
    H, W = int(size[0]), int(size[1])
    gl = gr = None
    if n > 1:
        nxt_lr = _f32c(nxt_lr, "next_lr")
        if tuple(nxt_lr.shape[-3:]) != (C, hl, wl):
            raise FuvsError("dense_lowres_interval: the two key frames differ in shape")
        gl = _grid_list(grids_left, n, (H, W), dev, "dense_lowres_interval(grids_left)")
        gr = _grid_list(grids_right, n, (H, W), dev, "dense_lowres_interval(grids_right)")
    scratch = _scratch(scratch, int(load().fuvs_dense_scratch_floats(C, H, W, n)), dev)
    labels = torch.empty((n, H, W), dtype=torch.uint8, device=dev) if (want_labels or counts is not None) else None
    logits = torch.empty((n, C, H, W), dtype=torch.float32, device=dev) if want_logits else None
    _check_tc(tc_prev, counts, H, W, C)
    ups = ups if ups is not None else KeyFrameUps()
    outputs = (labels is not None, logits is not None)
    tag_prev = _keyframe_tag(prev_obj, (H, W), n, outputs)
    ip = 1 if _same_tag(ups.tags[1], tag_prev) else 0
    numel = C * H * W
    up_prev = ups._buffer(ip, numel, dev)                 # (a re-allocated buffer loses its tag)
    ready = 1 if _same_tag(ups.tags[ip], tag_prev) else 0
    up_next = ups._buffer(1 - ip, numel, dev)
    with torch.cuda.device(dev):
        check(load().fuvs_dense_lowres_interval_ptrs(ptr(prev_lr), ptr(nxt_lr) if n > 1 else None, hl, wl, ptr(up_prev), ready,
                                                     ptr(up_next), ptr_array(gl), ptr_array(gr), C, H, W, n, ptr(scratch),
                                                     ptr(labels), ptr(logits), ptr(tc_prev), ptr(counts), ignore_index,
                                                     stream_ptr(dev)))
    ups.tags[ip] = tag_prev
    ups.tags[1 - ip] = _keyframe_tag(next_obj, (H, W), n, outputs) if n > 1 else None
    return labels, logits


def block_interval(prev, nxt, grids_left, grids_right, n, *, want_labels=True, want_logits=False, tc_prev=None,
                   counts=None, ignore_index=255, scratch=None):
    """fuvs_block_interval_ptrs.  grids_*: list of n-1 [1,Hg,Wg,2] tensors or stacked [n-1,Hg,Wg,2]."""
    dev = require_cuda(prev, nxt, tc_prev, counts, what="block_interval")
    prev = _f32c(prev, "prev")
    C, H, W = prev.shape[-3:]
    gl = gr = None
    Hg = Wg = 0
    if n > 1:
        nxt = _f32c(nxt, "next")
        gl = _grid_list(grids_left, n, None, dev, "block_interval(grids_left)")
        Hg, Wg = gl[0].shape[:2]
        gl = _grid_list(gl, n, (Hg, Wg), dev, "block_interval(grids_left)")
        gr = _grid_list(grids_right, n, (Hg, Wg), dev, "block_interval(grids_right)")
    scratch = _scratch(scratch, int(load().fuvs_block_scratch_floats(C, Hg, Wg, n)), dev)
    labels = torch.empty((n, H, W), dtype=torch.uint8, device=dev) if (want_labels or counts is not None) else None
    logits = torch.empty((n, C, H, W), dtype=torch.float32, device=dev) if want_logits else None
    _check_tc(tc_prev, counts, H, W, C)
    with torch.cuda.device(dev):
        check(load().fuvs_block_interval_ptrs(ptr(prev), ptr(nxt) if n > 1 else None, ptr_array(gl), ptr_array(gr), C, H, W,
                                              Hg, Wg, n, ptr(scratch), ptr(labels), ptr(logits), ptr(tc_prev), ptr(counts),
                                              ignore_index, stream_ptr(dev)))
    return labels, logits


def block_lowres_interval(prev_lr, nxt_lr, size, grids_left, grids_right, n, *, want_labels=True, want_logits=False,
                          tc_prev=None, counts=None, ignore_index=255, scratch=None):
    """fuvs_block_lowres_interval_ptrs: key frames at decoder resolution [C,hl,wl] / [1,C,hl,wl], frame size `size`.
    Shapes the fused kernels do not take go through fuvs_upsample_bilinear_ac + fuvs_block_interval — the same arithmetic
    in three launches."""
    dev = require_cuda(prev_lr, nxt_lr, tc_prev, counts, what="block_lowres_interval")
    prev_lr = _f32c(prev_lr, "prev_lr")
    C, hl, wl = prev_lr.shape[-3:]
    H, W = int(size[0]), int(size[1])
    gl = _grid_list(grids_left, n, None, dev, "block_lowres_interval(grids_left)") if n > 1 else []
    Hg, Wg = (gl[0].shape[:2] if n > 1 else (0, 0))
    ok = n > 1 and load().fuvs_block_lowres_supported(C, hl, wl, H, W, Hg, Wg) and not (counts is not None and 0 <= ignore_index < C)
    if not ok or (hl, wl) == (H, W):
        up = upsample_bilinear_ac(prev_lr.reshape(1, C, hl, wl), (H, W))
        up_n = upsample_bilinear_ac(_f32c(nxt_lr, "next_lr").reshape(1, C, hl, wl), (H, W)) if n > 1 else None
        return block_interval(up, up_n, grids_left, grids_right, n, want_labels=want_labels, want_logits=want_logits,
                              tc_prev=tc_prev, counts=counts, ignore_index=ignore_index, scratch=scratch)
    nxt_lr = _f32c(nxt_lr, "next_lr")
    if nxt_lr.shape[-3:] != prev_lr.shape[-3:]:
        raise FuvsError("block_lowres_interval: key frames differ in shape")
    gl = _grid_list(gl, n, (Hg, Wg), dev, "block_lowres_interval(grids_left)")
    gr = _grid_list(grids_right, n, (Hg, Wg), dev, "block_lowres_interval(grids_right)")
    scratch = _scratch(scratch, int(load().fuvs_block_scratch_floats(C, Hg, Wg, n)), dev)
    labels = torch.empty((n, H, W), dtype=torch.uint8, device=dev) if (want_labels or counts is not None) else None
    logits = torch.empty((n, C, H, W), dtype=torch.float32, device=dev) if want_logits else None
    _check_tc(tc_prev, counts, H, W, C)
    with torch.cuda.device(dev):
        check(load().fuvs_block_lowres_interval_ptrs(ptr(prev_lr), ptr(nxt_lr), hl, wl, ptr_array(gl), ptr_array(gr), C, H, W,
                                                     Hg, Wg, n, ptr(scratch), ptr(labels), ptr(logits), ptr(tc_prev),
                                                     ptr(counts), ignore_index, stream_ptr(dev)))
    return labels, logits


def block_clip(keys, grids_left, grids_right, n, *, want_logits=False, tc_prev=None, counts=None, ignore_index=255,
               scratch=None):
    """fuvs_block_clip: m consecutive intervals of one clip in one call.  keys: m+1 key-frame logit maps [1,C,H,W] /
    [C,H,W]; grids_left / grids_right: per interval a list of n-1 grids.  Returns (labels [m,n,H,W] uint8, logits
    [m,n,C,H,W] or None); interval i's frame 0 is counted against the last map of interval i-1 (tc_prev for i = 0)."""
    m = len(keys) - 1
    if m < 1 or len(grids_left) != m or len(grids_right) != m:
        raise FuvsError(f"block_clip: {len(keys)} key frames need {len(keys) - 1} grid lists per side")
    dev = require_cuda(*keys, tc_prev, counts, what="block_clip")
    keys = [_f32c(k, "key frame") for k in keys]
    C, H, W = keys[0].shape[-3:]
    if any(tuple(k.shape[-3:]) != (C, H, W) for k in keys):
        raise FuvsError("block_clip: key frames differ in shape")
    gl, gr = [], []
    Hg = Wg = 0
    if n > 1:
        first = _grid_list(grids_left[0], n, None, dev, "block_clip(grids_left)")
        Hg, Wg = first[0].shape[:2]
        for i in range(m):
            gl += _grid_list(grids_left[i], n, (Hg, Wg), dev, "block_clip(grids_left)")
            gr += _grid_list(grids_right[i], n, (Hg, Wg), dev, "block_clip(grids_right)")
    scratch = _scratch(scratch, m * int(load().fuvs_block_scratch_floats(C, Hg, Wg, n)), dev)
    labels = torch.empty((m, n, H, W), dtype=torch.uint8, device=dev)
    logits = torch.empty((m, n, C, H, W), dtype=torch.float32, device=dev) if want_logits else None
    _check_tc(tc_prev, counts, H, W, C)
    with torch.cuda.device(dev):
        check(load().fuvs_block_clip(m, ptr_array(keys), ptr_array(gl) if n > 1 else None, ptr_array(gr) if n > 1 else None,
                                     C, H, W, Hg, Wg, n, ptr(scratch), ptr_array([labels[i] for i in range(m)]),
                                     ptr_array([logits[i] for i in range(m)]) if want_logits else None, ptr(tc_prev),
                                     ptr(counts), ignore_index, stream_ptr(dev)))
    return labels, logits


def upsample_bilinear_ac(src, size, out=None):
    """fuvs_upsample_bilinear_ac: F.interpolate(src, size, mode='bilinear', align_corners=True) for [...,Hin,Win]."""
    dev = require_cuda(src, out, what="upsample_bilinear_ac")
    src = _f32c(src, "src")
    Hin, Win = src.shape[-2:]
    Hout, Wout = int(size[0]), int(size[1])
    lead = tuple(src.shape[:-2])
    planes = 1
    for s in lead:
        planes *= s
    if out is None:
        out = torch.empty(lead + (Hout, Wout), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != lead + (Hout, Wout) or out.dtype != torch.float32 or not out.is_contiguous():
        raise FuvsError("upsample_bilinear_ac: bad `out` tensor")
    with torch.cuda.device(dev):
        check(load().fuvs_upsample_bilinear_ac(ptr(src), ptr(out), planes, Hin, Win, Hout, Wout, stream_ptr(dev)))
    return out


def upsample_argmax(logits, size, *, want_resized=False):
    """fuvs_upsample_argmax: F.interpolate(logits, size, bilinear, align_corners=True).max(1)[1] as uint8 (flow/base.py:
    275-277) -> (labels [F,Hout,Wout], resized logits or None)."""
    dev = require_cuda(logits, what="upsample_argmax")
    logits = _f32c(logits, "logits")
    if logits.dim() != 4:
        raise FuvsError("upsample_argmax: expected [F,C,H,W]")
    F_, Cc, Hin, Win = logits.shape
    Hout, Wout = int(size[0]), int(size[1])
    labels = torch.empty((F_, Hout, Wout), dtype=torch.uint8, device=dev)
    resized = torch.empty((F_, Cc, Hout, Wout), dtype=torch.float32, device=dev) if want_resized else None
    with torch.cuda.device(dev):
        check(load().fuvs_upsample_argmax(ptr(logits), F_, Cc, Hin, Win, Hout, Wout, ptr(labels), ptr(resized),
                                          stream_ptr(dev)))
    return labels, resized


def blend_argmax(a, b, wa, wb, *, want_out=True, want_labels=False, out=None):
    """fuvs_blend_argmax on [F,C,H,W] (or [C,H,W]): out = fl(fl(wa*a)+fl(wb*b)), labels = argmax_C."""
    dev = require_cuda(a, b, out, what="blend_argmax")
    a = _f32c(a, "a")
    if b is not None:
        b = _f32c(b, "b")
        if b.shape != a.shape:
            raise FuvsError(f"blend_argmax: shapes differ {tuple(a.shape)} vs {tuple(b.shape)}")
    if a.dim() == 3:
        frames, (Cc, H, W) = 1, a.shape
    elif a.dim() == 4:
        frames, Cc, H, W = a.shape
    else:
        raise FuvsError("blend_argmax: expected [C,H,W] or [F,C,H,W]")
    if want_out and out is None:
        out = torch.empty_like(a)
    if out is not None and (out.shape != a.shape or out.dtype != torch.float32 or not out.is_contiguous()):
        raise FuvsError("blend_argmax: bad `out` tensor")
    labels = torch.empty((frames, H, W), dtype=torch.uint8, device=dev) if want_labels else None
    with torch.cuda.device(dev):
        check(load().fuvs_blend_argmax(ptr(a), ptr(b), float(wa), float(wb), frames, Cc, H * W, ptr(out), ptr(labels),
                                       stream_ptr(dev)))
    return out, labels


def argmax(logits, *, dtype=torch.uint8):
    """fuvs_argmax: logits [F,C,H,W] -> labels [F,H,W] (uint8 or int64), torch.max(dim=1)[1] semantics."""
    dev = require_cuda(logits, what="argmax")
    logits = _f32c(logits, "logits")
    if logits.dim() != 4:
        raise FuvsError("argmax: expected [F,C,H,W]")
    F_, Cc, H, W = logits.shape
    if dtype == torch.uint8:
        u8, i64 = torch.empty((F_, H, W), dtype=torch.uint8, device=dev), None
    elif dtype == torch.int64:
        u8, i64 = None, torch.empty((F_, H, W), dtype=torch.int64, device=dev)
    else:
        raise FuvsError("argmax: dtype must be torch.uint8 or torch.int64")
    with torch.cuda.device(dev):
        check(load().fuvs_argmax(ptr(logits), F_, Cc, H * W, ptr(u8), ptr(i64), stream_ptr(dev)))
    return u8 if u8 is not None else i64


def confusion(pred, target, K, ignore_index=255, *, counts=None, numpy_bins=False, mutate_pred=False):
    """fuvs_confusion: accumulates (I,U,T) of util/util.py:52-63 (or :36-47 with numpy_bins) into counts [3,K]."""
    dev = require_cuda(pred, target, counts, what="confusion")
    if pred.shape != target.shape:
        raise FuvsError(f"confusion: shapes differ {tuple(pred.shape)} vs {tuple(target.shape)}")
    for t, name in ((pred, "pred"), (target, "target")):
        if t.dtype not in (torch.uint8, torch.int64):
            raise FuvsError(f"confusion: {name} must be uint8 or int64, got {t.dtype}")
    if not pred.is_contiguous():
        if mutate_pred:
            raise FuvsError("confusion: mutate_pred needs a contiguous pred")
        pred = pred.contiguous()
    target = target.contiguous()
    if counts is None:
        counts = new_counts(K, dev)
    _check_tc(None, counts, 0, 0, K)
    flags = (_lib.FUVS_BINS_NPHIST if numpy_bins else _lib.FUVS_BINS_HISTC) | (_lib.FUVS_MUTATE_PRED if mutate_pred else 0)
    with torch.cuda.device(dev):
        check(load().fuvs_confusion(ptr(pred), 1 if pred.dtype == torch.int64 else 0, ptr(target),
                                    1 if target.dtype == torch.int64 else 0, pred.numel(), K, int(ignore_index), flags,
                                    ptr(counts), stream_ptr(dev)))
    return counts


def temporal_counts(labels, K, ignore_index=255, *, tc_prev=None, counts=None):
    """fuvs_temporal_counts over uint8 labels [n,H,W] (flow/base.py:280-295)."""
    dev = require_cuda(labels, tc_prev, counts, what="temporal_counts")
    if labels.dtype != torch.uint8 or labels.dim() != 3 or not labels.is_contiguous():
        raise FuvsError("temporal_counts: labels must be contiguous uint8 [n,H,W]")
    n, H, W = labels.shape
    if counts is None:
        counts = new_counts(K, dev)
    _check_tc(tc_prev, counts, H, W, K)
    with torch.cuda.device(dev):
        check(load().fuvs_temporal_counts(ptr(labels), n, H * W, ptr(tc_prev), K, int(ignore_index), ptr(counts),
                                          stream_ptr(dev)))
    return counts


def crop_grid(grid, H, W, crop_h, crop_w, h_off, w_off):
    """fuvs_crop_grid: crop_motion_vector (flow/transform.py:215-261) for one grid [1,Hg,Wg,2] / [Hg,Wg,2] on the
    device -> [1, crop_h//16, crop_w//16, 2]."""
    dev = require_cuda(grid, what="crop_grid")
    if grid.dtype != torch.float32:
        raise FuvsError(f"crop_grid: expected a float32 grid (flow/dataset.py:240 loads them as float32), got {grid.dtype}")
    g = grid.reshape(grid.shape[-3], grid.shape[-2], 2).contiguous()
    out = torch.empty((1, crop_h // 16, crop_w // 16, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(load().fuvs_crop_grid(ptr(g), g.shape[0], g.shape[1], int(H), int(W), int(crop_h), int(crop_w), int(h_off),
                                    int(w_off), ptr(out), stream_ptr(dev)))
    return out


def crop_accumulate(logits, canvas, count, h_off, w_off):
    """fuvs_crop_accumulate: canvas[:, :, window] += softmax(logits, 1); count[window] += 1 (flow/base.py:206-207,220,233)."""
    dev = require_cuda(logits, canvas, count, what="crop_accumulate")
    logits = _f32c(logits, "logits")
    n, Cc, ch, cw = logits.shape
    if canvas.dtype != torch.float64 or count.dtype != torch.float64 or not canvas.is_contiguous() or not count.is_contiguous():
        raise FuvsError("crop_accumulate: canvas / count must be contiguous float64 tensors")
    if canvas.dim() != 4 or canvas.shape[0] != n or canvas.shape[1] != Cc or tuple(count.shape) != tuple(canvas.shape[2:]):
        raise FuvsError(f"crop_accumulate: canvas {tuple(canvas.shape)} / count {tuple(count.shape)} do not match logits {tuple(logits.shape)}")
    H, W = canvas.shape[2:]
    with torch.cuda.device(dev):
        check(load().fuvs_crop_accumulate(ptr(logits), ptr(canvas), ptr(count), n, Cc, ch, cw, H, W, int(h_off), int(w_off),
                                          stream_ptr(dev)))


def crop_finish(canvas, count, *, want_labels=True):
    """fuvs_crop_finish: canvas /= count in place; returns uint8 labels [n,H,W] = canvas.max(1)[1]."""
    dev = require_cuda(canvas, count, what="crop_finish")
    n, Cc, H, W = canvas.shape
    labels = torch.empty((n, H, W), dtype=torch.uint8, device=dev) if want_labels else None
    with torch.cuda.device(dev):
        check(load().fuvs_crop_finish(ptr(canvas), ptr(count), n, Cc, H * W, ptr(labels), stream_ptr(dev)))
    return labels


def feature_interval(f_prev, f_next, grids_left, grids_right, n, *, default_grid=None, scratch=None, out=None):
    """fuvs_feature_interval: predict_feature's warp / up-sample / blend / cat (flow/model.py:131-173) -> [n,C,fh,fw]."""
    dev = require_cuda(f_prev, f_next, default_grid, out, what="feature_interval")
    f_prev = _f32c(f_prev, "f_prev")
    C, fh, fw = f_prev.shape[-3:]
    gl = gr = None
    Hg = Wg = 0
    if n > 1:
        f_next = _f32c(f_next, "f_next")
        gl = _grid_list(grids_left, n, None, dev, "feature_interval(grids_left)")
        Hg, Wg = gl[0].shape[:2]
        gl = _grid_list(gl, n, (Hg, Wg), dev, "feature_interval(grids_left)")
        gr = _grid_list(grids_right, n, (Hg, Wg), dev, "feature_interval(grids_right)")
    Hd = Wd = 0
    if default_grid is not None:
        default_grid = _f32c(default_grid if default_grid.dtype == torch.float32 else default_grid.float(), "default_grid")
        Hd, Wd = default_grid.shape[-3:-1]
    scratch = _scratch(scratch, int(load().fuvs_feature_scratch_floats(C, Hg, Wg, Hd, Wd, n)), dev)
    if out is None:
        out = torch.empty((n, C, fh, fw), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (n, C, fh, fw) or out.dtype != torch.float32 or not out.is_contiguous():
        raise FuvsError("feature_interval: bad `out` tensor")
    with torch.cuda.device(dev):
        check(load().fuvs_feature_interval_ptrs(ptr(f_prev), ptr(f_next) if n > 1 else None, ptr_array(gl), ptr_array(gr), ptr(default_grid),
                                           Hd, Wd, C, fh, fw, Hg, Wg, n, ptr(scratch), ptr(out), stream_ptr(dev)))
    return out
