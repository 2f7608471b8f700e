"""Drop-in for the reference's flow/model.py (same names, arguments, return values).

FlowModel keeps the reference's constructor and methods (flow/model.py:23-249):
forward / forward_feature / forward_segmentation / warp_batch / predict /
predict_feature / predict_segmentation / warp.  Under inference (no autograd)
on CUDA tensors every warp, up-sample, blend and arg-max runs in libfuvs.so
(hand-written sm_100a kernels) and reproduces the eager ATen sequence bit for
bit.  `predict_labels` is the fused entry the evaluation loop uses
(FlowBaseModel.predict_step): it goes from key-frame logits straight to uint8
label maps and temporal-consistency counts without materialising [n,C,h,w]
logits.

Scope notes
  * The key-frame networks (model.encoder / model.decoder) stay the
    reference's PyTorch modules and are called as-is.
  * Training needs autograd through grid_sample; that is outside the B200
    path.  When gradients are required the methods compose
    the differentiable torch ops exactly as the reference does.  This is not a
    CPU fallback: inference on CPU tensors raises.
  * No parameters or buffers are added (checkpoint keys stay
    model_G.model.{encoder,decoder}.*; default_motion_vector is a plain
    attribute as in flow/model.py:32).
"""
from __future__ import annotations

import random

import numpy as np
import torch
from torch import nn

from .. import kernels
from .._lib import FuvsError


def get_default_grid():
    """flow/model.py:10-21 — identity grid of a 1920x1072 frame at 16x16 macro-block centres (fp64)."""
    width, height, block = 1920, 1072, 16
    bh, bw = height // block, width // block
    grid = np.zeros((bh, bw, 2))
    grid[:, :, 0] = ((np.arange(bw, dtype=np.float64) * block + block // 2) / width * 2 - 1)[None, :]
    grid[:, :, 1] = ((np.arange(bh, dtype=np.float64) * block + block // 2) / height * 2 - 1)[:, None]
    return grid


def _grad_needed(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def _interp_ac(x, h, w):
    """`if shape != (h,w): F.interpolate(bilinear, align_corners=True)` — kernel or autograd route."""
    if x.shape[2] == h and x.shape[3] == w:
        return x
    if _grad_needed(x):
        return torch.nn.functional.interpolate(x, size=(h, w), mode="bilinear", align_corners=True)
    return kernels.upsample_bilinear_ac(x, (h, w))


def _is_grid(m) -> bool:
    return isinstance(m, torch.Tensor) and m.dim() >= 3 and m.shape[-1] == 2


class FlowModel(nn.Module):
    def __init__(self, model, feature_based=True, no_warp=False, no_interpolation_percentage=0.0):
        super().__init__()
        self.model = model
        self.feature_based = feature_based
        self.no_warp = no_warp
        self.no_interpolation_percentage = no_interpolation_percentage
        # Default grid - no motion (plain attribute, not a buffer: flow/model.py:32)
        self.default_motion_vector = torch.from_numpy(get_default_grid()).float().unsqueeze(0)
        # Key-frame reuse (opt-in): in the reference every key frame goes through the network
        # twice, as `next` of interval i and as `prev` of interval i+1 (flow/model.py:189,202).  When the caller passes
        # frame ids the network output of `next` is kept for the following interval.  The entry is keyed on everything
        # that makes it valid — frame id, output size, device, resolution kind, and a version that every change of the
        # weights or of the mode bumps (train/eval, load_state_dict, .to()/.half()) — and dropped by
        # reset_keyframe_cache() (FlowBaseModel.on_predict_start: one video per predict run).  Plain attributes.
        self.reuse_keyframes = False
        self._kf_cache = None          # (key tuple, tensor)
        self._kf_version = 0
        # chain-state workspace of the interval kernels, grown on demand and reused across calls
        self._scratch = kernels.ScratchCache()
        self._dense_ups = kernels.KeyFrameUps()     # up-sampled key frames of the dense route (reused interval to interval)

    # ------------------------------------------------------------------ forward (train / val / test)
    def forward(self, frame_current, frame_prev, frame_next, mvs_left, mvs_right, left_index, right_index):
        """flow/model.py:35-52."""
        if self.training and frame_current is not None and random.random() < self.no_interpolation_percentage:
            h, w = frame_current.shape[2], frame_current.shape[3]
            output = self.model.decoder(self.model.encoder(frame_current))
            return {"pred": _interp_ac(output, h, w)}
        left_index = [int(i) for i in left_index]
        right_index = [int(i) for i in right_index]
        n_list = [sum(x) for x in zip(left_index, right_index)]
        if self.feature_based:
            return self.forward_feature(frame_prev, frame_next, mvs_left, mvs_right, left_index, right_index, n_list)
        return self.forward_segmentation(frame_prev, frame_next, mvs_left, mvs_right, left_index, right_index, n_list)

    def forward_feature(self, frame_prev, frame_next, mvs_left, mvs_right, left_index, right_index, n_list):
        """flow/model.py:55-70."""
        h, w = frame_prev.shape[2], frame_prev.shape[3]
        f_prev = self.model.encoder(frame_prev)
        f_next = self.model.encoder(frame_next)
        f = self._fuse_two_sides(f_prev, f_next, mvs_left, mvs_right, left_index, right_index, n_list)
        output = self.model.decoder(f)
        return {"pred": _interp_ac(output, h, w)}

    def forward_segmentation(self, frame_prev, frame_next, mvs_left, mvs_right, left_index, right_index, n_list):
        """flow/model.py:73-88."""
        h, w = frame_prev.shape[2], frame_prev.shape[3]
        o_prev = self.model.decoder(self.model.encoder(frame_prev))
        o_next = self.model.decoder(self.model.encoder(frame_next))
        o = self._fuse_two_sides(o_prev, o_next, mvs_left, mvs_right, left_index, right_index, n_list)
        return {"pred": _interp_ac(o, h, w)}

    def _fuse_two_sides(self, x_prev, x_next, mvs_left, mvs_right, left_index, right_index, n_list):
        """warp_batch(prev) + warp_batch(next) (flow/model.py:61-64, 81-84) in one blend launch per sample."""
        if _grad_needed(x_prev, x_next):
            return self.warp_batch(x_prev, mvs_left, left_index, n_list) + \
                self.warp_batch(x_next, mvs_right, right_index, n_list)
        dev = kernels.require_cuda(x_prev, x_next, what="FlowModel.forward")
        x_prev, x_next = x_prev.contiguous().float(), x_next.contiguous().float()
        out = torch.empty_like(x_prev)
        for i in range(len(left_index)):
            a = self._chain_restore(x_prev, mvs_left, i, left_index[i])
            b = self._chain_restore(x_next, mvs_right, i, right_index[i])
            n = n_list[i]
            kernels.blend_argmax(a, b, (n - left_index[i]) / n, (n - right_index[i]) / n, out=out[i])
        return out

    def _chain_restore(self, x, mvs, i, index):
        """`index` chained warps of sample i, then the size restore of flow/model.py:102-103 -> [C,ih,iw]."""
        i_h, i_w = x.shape[2], x.shape[3]
        cur = x[i]
        if self.no_warp:
            return cur
        for j in range(index):
            cur, _ = kernels.warp_step(cur, mvs[j][i])
        # the reference tests shape[1] (C) and shape[2] (H) of the 4-D tensor against (i_h, i_w)
        if cur.shape[0] != i_h or cur.shape[1] != i_w:
            if cur.shape[1] != i_h or cur.shape[2] != i_w:
                cur = kernels.upsample_bilinear_ac(cur, (i_h, i_w))
        elif cur.shape[1] != i_h or cur.shape[2] != i_w:
            raise FuvsError("warp_batch: warped map is not restored to the input size (reference would fail at the add)")
        return cur

    def warp_batch(self, input, mvs, index_list, n_list):
        """flow/model.py:92-106 — per-sample chained warps, size restore, weight (n-index)/n."""
        if _grad_needed(input):
            i_h, i_w = input.shape[2], input.shape[3]
            input_warped = []
            for i in range(len(index_list)):
                index = index_list[i]
                cur = input[i].unsqueeze(0)
                if not self.no_warp:
                    for j in range(index):
                        cur = self.warp(cur, mvs[j][i].unsqueeze(0))
                    if cur.shape[1] != i_h or cur.shape[2] != i_w:
                        cur = torch.nn.functional.interpolate(cur, size=(i_h, i_w), mode="bilinear", align_corners=True)
                input_warped.append(cur * ((n_list[i] - index) / n_list[i]))
            return torch.cat(input_warped)
        kernels.require_cuda(input, what="FlowModel.warp_batch")
        input = input.contiguous().float()
        out = torch.empty_like(input)
        for i in range(len(index_list)):
            cur = self._chain_restore(input, mvs, i, index_list[i])
            kernels.blend_argmax(cur, None, (n_list[i] - index_list[i]) / n_list[i], 0.0, out=out[i])
        return out

    # ------------------------------------------------------------------ predict (inference over a video)
    def predict(self, *args, **kwargs):
        """flow/model.py:109-113."""
        if self.feature_based:
            return self.predict_feature(*args, **kwargs)
        return self.predict_segmentation(*args, **kwargs)

    def _keyframe_logits(self, frame, h, w, profiler, keep_lowres=False):
        """decoder(encoder(frame)) brought to (h,w) (flow/model.py:188-193).  keep_lowres: hand back the decoder
        output itself — the linear route evaluates the up-sample inside the blend kernel."""
        with profiler.profile("predict_encoder"):
            f = self.model.encoder(frame)
        with profiler.profile("predict_decoder"):
            o = self.model.decoder(f)
            if not (keep_lowres and not _grad_needed(o) and o.is_cuda):
                o = _interp_ac(o, h, w)
        return o

    def _kf_key(self, frame_id, shape, device, lowres):
        return (int(frame_id), tuple(shape), str(device), bool(lowres), self._kf_version, bool(self.training))

    def _cached_keyframe(self, frame_id, shape, device, lowres):
        c = self._kf_cache
        if self.reuse_keyframes and frame_id is not None and c is not None and \
                c[0] == self._kf_key(frame_id, shape, device, lowres):
            return c[1]
        return None

    def reset_keyframe_cache(self):
        self._kf_cache = None
        self._kf_version += 1
        if hasattr(self, "_dense_ups"):
            self._dense_ups.reset()

    # anything that changes the weights or the mode invalidates cached network outputs
    def train(self, mode=True):
        self.reset_keyframe_cache()
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        self.reset_keyframe_cache()
        self._scratch.clear()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.reset_keyframe_cache()
        return super().load_state_dict(*args, **kwargs)

    def _keeps_lowres(self, mvs_left, h, w):
        """Routes whose entries take the key frames at decoder resolution: linear and block-grid
        evaluate the up-sample inside their kernels, the dense entry up-samples straight into the layout its first
        warp step reads and keeps the result for the next interval."""
        return self._interval_mode(mvs_left, h, w) in ("linear", "block", "dense")

    def _interval_mode(self, mvs_left, h, w):
        if self.no_warp or len(mvs_left) == 0 or not _is_grid(mvs_left[0]):
            return "linear"
        hg, wg = mvs_left[0].shape[-3], mvs_left[0].shape[-2]
        return "dense" if (hg == h and wg == w) else "block"

    def _run_interval(self, o, o_next, mvs_left, mvs_right, n, *, want_labels, want_logits, tc_prev=None, counts=None,
                      ignore_index=255, profiler=None, size=None):
        """Key-frame logits [1,C,h,w] x2 -> (labels [n,h,w] uint8, logits [n,C,h,w]) through one fused call."""
        kernels.require_cuda(o, o_next, what="FlowModel.predict")
        if o.shape[0] != 1:
            raise FuvsError("FlowModel.predict: the inference path takes one clip interval at a time (batch size 1, "
                            "as asserted by flow/base.py:263)")
        if size is None:
            size = (o.shape[2], o.shape[3])
        h, w = size
        if o_next is None:
            n = 1
        mode = self._interval_mode(mvs_left, h, w) if n > 1 else "linear"
        kw = dict(want_labels=want_labels, want_logits=want_logits, tc_prev=tc_prev, counts=counts,
                  ignore_index=ignore_index)
        if mode == "linear":
            if (o.shape[2], o.shape[3]) != (h, w):            # key frames still at decoder resolution
                return kernels.linear_lowres_blend_argmax(o, o_next, (h, w), n, **kw)
            return kernels.linear_blend_argmax(o, o_next, n, **kw)
        if len(mvs_left) != n - 1 or len(mvs_right) != n - 1:
            raise FuvsError(f"FlowModel.predict: n={n} needs {n - 1} grids per side, got {len(mvs_left)}/{len(mvs_right)}")
        if mode == "block" and (o.shape[2], o.shape[3]) != (h, w):
            # key frames still at decoder resolution: the block route evaluates their up-sample inside its kernels
            # (shapes it does not take are up-sampled by the wrapper: same arithmetic)
            return kernels.block_lowres_interval(o, o_next, (h, w), mvs_left, mvs_right, n, scratch=self._scratch, **kw)
        if mode == "dense" and (o.shape[2], o.shape[3]) != (h, w):
            return kernels.dense_lowres_interval(o, o_next, (h, w), mvs_left, mvs_right, n, scratch=self._scratch,
                                                 ups=self._dense_ups, **kw)
        if (o.shape[2], o.shape[3]) != (h, w):
            o = _interp_ac(o, h, w)
            o_next = _interp_ac(o_next, h, w) if o_next is not None else None
        if mode == "dense":
            return kernels.dense_interval(o, o_next, mvs_left, mvs_right, n, scratch=self._scratch, **kw)
        return kernels.block_interval(o, o_next, mvs_left, mvs_right, n, scratch=self._scratch, **kw)

    def predict_segmentation(self, frame_prev, frame_next, mvs_left, mvs_right, n, profiler):
        """flow/model.py:184-241 -> {"pred": [n,C,h,w]} (frame 0 = key-frame logits)."""
        h, w = frame_prev.shape[2], frame_prev.shape[3]
        lowres = self._keeps_lowres(mvs_left, h, w) and frame_next is not None
        o = self._keyframe_logits(frame_prev, h, w, profiler, keep_lowres=lowres)
        if frame_next is None:
            return {"pred": o}
        o_next = self._keyframe_logits(frame_next, h, w, profiler, keep_lowres=lowres)
        if _grad_needed(o, o_next):
            return {"pred": self._predict_segmentation_autograd(o, o_next, mvs_left, mvs_right, n, h, w)}
        with profiler.profile("predict_warp"):
            with profiler.profile("predict_fusion"):
                _, logits = self._run_interval(o, o_next, mvs_left, mvs_right, n, want_labels=False, want_logits=True,
                                               size=(h, w))
        return {"pred": logits}

    def predict_labels(self, frame_prev, frame_next, mvs_left, mvs_right, n, profiler, *, tc_prev=None, counts=None,
                       ignore_index=255, frame_id=None):
        """Fused evaluation entry: predict_segmentation + max(1)[1] + uint8 cast + temporal-consistency counts
        (flow/model.py:184-241, flow/base.py:276-277, 280-295) -> uint8 [n,h,w].  Segmentation-based models only."""
        if self.feature_based:
            logits = self.predict_feature(frame_prev, frame_next, mvs_left, mvs_right, n, profiler)["pred"]
            labels = kernels.argmax(logits)
            if counts is not None:
                kernels.temporal_counts(labels, logits.shape[1], ignore_index, tc_prev=tc_prev, counts=counts)
            return labels
        h, w = frame_prev.shape[2], frame_prev.shape[3]
        with torch.no_grad():
            lowres = self._keeps_lowres(mvs_left, h, w) and frame_next is not None
            o = self._cached_keyframe(frame_id, (h, w), frame_prev.device, lowres)
            if o is None:
                o = self._keyframe_logits(frame_prev, h, w, profiler, keep_lowres=lowres)
            o_next = self._keyframe_logits(frame_next, h, w, profiler, keep_lowres=lowres) if frame_next is not None else None
            if self.reuse_keyframes and frame_id is not None and o_next is not None:
                self._kf_cache = (self._kf_key(int(frame_id) + int(n), (h, w), frame_prev.device, lowres), o_next)
            with profiler.profile("predict_warp"):
                with profiler.profile("predict_fusion"):
                    labels, _ = self._run_interval(o, o_next, mvs_left, mvs_right, n, want_labels=True,
                                                   want_logits=False, tc_prev=tc_prev, counts=counts,
                                                   ignore_index=ignore_index, size=(h, w))
        return labels

    def predict_feature(self, frame_prev, frame_next, mvs_left, mvs_right, n, profiler):
        """flow/model.py:116-181 -> {"pred": [n,C,h,w]}."""
        h, w = frame_prev.shape[2], frame_prev.shape[3]
        with profiler.profile("predict_encoder"):
            f = self.model.encoder(frame_prev)
        f_next = None
        if frame_next is not None:
            with profiler.profile("predict_encoder"):
                f_next = self.model.encoder(frame_next)
        if _grad_needed(f, f_next):
            return {"pred": self._predict_feature_autograd(f, f_next, mvs_left, mvs_right, n, h, w, profiler)}
        dev = kernels.require_cuda(f, f_next, what="FlowModel.predict_feature")
        if f.shape[0] != 1:
            raise FuvsError("FlowModel.predict_feature: batch size must be 1 (flow/base.py:263)")
        f = f.contiguous().float()
        if f_next is not None:
            f_next = f_next.contiguous().float()
        cf, f_h, f_w = f.shape[1], f.shape[2], f.shape[3]
        frames = n if f_next is not None else 1
        if not self.no_warp:
            # one fused call: chains at grid resolution, up-sample + blend written straight into the decoder batch,
            # key frame through the default grid with align_corners=True (flow/model.py:131-173)
            if self.default_motion_vector.device != f.device:
                self.default_motion_vector = self.default_motion_vector.to(device=f.device)
            with profiler.profile("predict_warp"):
                with profiler.profile("predict_fusion"):
                    feature_maps = kernels.feature_interval(f[0], f_next[0] if f_next is not None else None, mvs_left,
                                                            mvs_right, frames, default_grid=self.default_motion_vector,
                                                            scratch=self._scratch)
        else:
            feature_maps = torch.empty((frames, cf, f_h, f_w), dtype=torch.float32, device=dev)
            feature_maps[0].copy_(f[0])
            if f_next is not None:
                with profiler.profile("predict_fusion"):
                    for p in range(1, n):
                        # the reference blends the *resampled* key frame only when warping; here f is untouched
                        kernels.blend_argmax(f[0], f_next[0], (n - p) / n, p / n, out=feature_maps[p])
        with profiler.profile("predict_decoder"):
            output = self.model.decoder(feature_maps)
            output = _interp_ac(output, h, w)
        return {"pred": output}

    @staticmethod
    def _restore(x, h, w):
        if x.shape[1] != h or x.shape[2] != w:
            return kernels.upsample_bilinear_ac(x, (h, w))
        return x

    # ------------------------------------------------------------------ warp
    def warp(self, frame, motion_vectors):
        """flow/model.py:244-249 — [B,C,H,W] x [B,Hg,Wg,2] -> [B,C,Hg,Wg]."""
        if self.no_warp:
            return frame
        if not isinstance(motion_vectors, torch.FloatTensor):
            motion_vectors = motion_vectors.float()
        if _grad_needed(frame, motion_vectors):
            return torch.nn.functional.grid_sample(frame, motion_vectors, mode="bilinear", padding_mode="border",
                                                   align_corners=False)
        kernels.require_cuda(frame, motion_vectors, what="FlowModel.warp")
        outs = [kernels.warp_step(frame[b], motion_vectors[b])[0] for b in range(frame.shape[0])]
        return torch.stack(outs, 0) if len(outs) != 1 else outs[0].unsqueeze(0)

    # ------------------------------------------------------------------ autograd routes (training only)
    def _predict_segmentation_autograd(self, o, o_next, mvs_left, mvs_right, n, h, w):
        fwd, bwd = [], []
        cur = o
        for m in mvs_left:
            cur = self.warp(cur, m)
            fwd.append(_interp_ac(cur, h, w))
        cur = o_next
        for m in mvs_right:
            cur = self.warp(cur, m)
            bwd.append(_interp_ac(cur, h, w))
        maps = [o]
        for p in range(1, n):
            maps.append((n - p) / n * fwd[p - 1] + p / n * bwd[n - p - 1])
        return torch.cat(maps, 0)

    def _predict_feature_autograd(self, f, f_next, mvs_left, mvs_right, n, h, w, profiler):
        f_h, f_w = f.shape[2], f.shape[3]
        fwd, bwd = [], []
        if f_next is not None and not self.no_warp:
            cur = f
            for m in mvs_left:
                cur = self.warp(cur, m)
                fwd.append(_interp_ac(cur, f_h, f_w))
            cur = f_next
            for m in mvs_right:
                cur = self.warp(cur, m)
                bwd.append(_interp_ac(cur, f_h, f_w))
        if not self.no_warp:
            if self.default_motion_vector.device != f.device:
                self.default_motion_vector = self.default_motion_vector.to(device=f.device)
            f = torch.nn.functional.grid_sample(f, self.default_motion_vector, padding_mode="border", align_corners=True)
            f = _interp_ac(f, f_h, f_w)
        maps = [f]
        if f_next is not None:
            for p in range(1, n):
                if not self.no_warp:
                    maps.append((n - p) / n * fwd[p - 1] + p / n * bwd[n - p - 1])
                else:
                    maps.append((n - p) / n * f + p / n * f_next)
        output = self.model.decoder(torch.cat(maps, 0))
        return _interp_ac(output, h, w)
