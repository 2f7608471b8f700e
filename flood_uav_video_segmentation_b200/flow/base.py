"""Drop-in for the evaluation / inference half of the reference's flow/base.py (FlowBaseModel, lines 51-343).

Same hook names and batch dictionaries: forward, validation_step, test_step, predict_step, on_predict_start,
on_predict_end.  What changes underneath:

  * predict_step (flow/base.py:259-312) calls FlowModel.predict_labels: key-frame logits -> uint8 label maps and
    temporal-consistency counts in fused kernels.  The label maps leave the device once (the reference's
    `.cpu().numpy().astype('uint8')`, :277) only if something consumes them (video / PNG sink, or the caller).
  * (I,U,T) counts stay on the GPU in DeviceMeter objects for the whole epoch instead of three blocking
    `.cpu().numpy()` per metric call (base/foundation.py:344); epoch-end hooks all-reduce them once over NCCL and
    apply the reference formulas in fp64 numpy.  The reference's per-rank AverageMeter attributes are filled from
    them at epoch end so downstream code that reads `*.sum` keeps working.
  * The key-frame networks are the reference's own PyTorch modules (model.pspnet.FlowPSPNet /
    model.deeplabv3.FlowDeepLabv3), imported from the reference tree when `arch` is given, or injected with
    `backbone=` (any module exposing .encoder and .decoder).

Lightning is optional: with pytorch_lightning installed the class is a LightningModule (drop-in for
FlowLightningCLI); without it the same methods work on a plain nn.Module driven by a loop (bench.py, tests).
  * The sliding-crop route (model.no_cropping=False, flow/base.py:182-234) runs entirely on the device:
    crop_motion_vector is a kernel (the reference moves every grid to the host, through numpy and cv2.resize, and
    back: 2(k-1) round trips per crop), soft-max and the fp64 canvas update are one kernel per crop, the final
    division and arg-max another.
"""
from __future__ import annotations

import contextlib
import time
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import kernels
from ..base.foundation import DeviceMeter, epoch_metrics, round_train
from ..util.util import AverageMeter
from .model import FlowModel

try:  # pragma: no cover - Lightning is absent in the build image
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # noqa: BLE001
    pl = None
    _Base = nn.Module


class SimpleProfiler:
    """Minimal stand-in for Lightning's SimpleProfiler: named wall-clock regions (flow/base.py:269, 321-328)."""

    def __init__(self):
        self.recorded_durations = {}

    @contextlib.contextmanager
    def profile(self, name):
        t0 = time.perf_counter()
        try:
            yield
        finally:
            self.recorded_durations.setdefault(name, []).append(time.perf_counter() - t0)


class FlowBaseModel(_Base):
    def __init__(self, classes: int = 5, ignore_index: int = 255, test_h: int = 873, test_w: int = 873,
                 arch: str = "pspnet", feature_based: bool = True, no_warp: bool = False, no_cropping: bool = False,
                 no_interpolation_percentage: float = 0.0, layers: int = 101, zoom_factor: int = 8,
                 compute_metrics: bool = True, save_images: bool = False, save_video: bool = True,
                 data_root: str = "dataset/flow/", predict_v_id: str = "florida-01", pretrained: bool = True,
                 backbone: nn.Module | None = None, output_size=(1072, 1920), reuse_keyframes: bool = False,
                 **kwargs):
        super().__init__()
        hp = dict(classes=classes, ignore_index=ignore_index, test_h=round_train(test_h, arch),
                  test_w=round_train(test_w, arch), arch=arch, feature_based=feature_based, no_warp=no_warp,
                  no_cropping=no_cropping, no_interpolation_percentage=no_interpolation_percentage, layers=layers,
                  zoom_factor=zoom_factor, compute_metrics=compute_metrics, save_images=save_images,
                  save_video=save_video, data_root=data_root, predict_v_id=predict_v_id, pretrained=pretrained,
                  output_size=tuple(output_size), reuse_keyframes=reuse_keyframes, **kwargs)
        if pl is not None:  # pragma: no cover
            self.save_hyperparameters(hp)
        else:
            self.hparams = SimpleNamespace(**hp)
        self._local_profiler = SimpleProfiler()
        self.frame_sink = None          # optional callable(frame_id:int, labels_uint8:np.ndarray[n,H,W])
        self.init_model(backbone)
        self.init_metrics_val()
        self.init_metrics_test()

    # ------------------------------------------------------------------ model
    def get_new_model_arch_G(self, backbone=None):
        """flow/base.py:87-108.  The segmentation networks are the reference's modules."""
        if backbone is None:
            arch = self.hparams.arch
            if arch == "pspnet":
                from model.pspnet import FlowPSPNet          # reference module (model/pspnet.py:113)
                backbone = FlowPSPNet(hparams=self.hparams)
            elif arch == "deeplabv3":
                from model.deeplabv3 import FlowDeepLabv3    # reference module (model/deeplabv3.py:47)
                backbone = FlowDeepLabv3(hparams=self.hparams)
            else:
                raise ValueError(f"flow models support arch in {{pspnet, deeplabv3}} (flow/base.py:94-103), got {arch!r}")
        return FlowModel(backbone, feature_based=self.hparams.feature_based, no_warp=self.hparams.no_warp,
                         no_interpolation_percentage=self.hparams.no_interpolation_percentage)

    def init_model(self, backbone=None):
        self.model_G = self.get_new_model_arch_G(backbone)
        self.model_G.reuse_keyframes = bool(self.hparams.reuse_keyframes)

    # ------------------------------------------------------------------ meters
    def _meter(self):
        dev = next((p.device for p in self.parameters()), None)
        if dev is None or dev.type != "cuda":
            dev = torch.device("cuda", torch.cuda.current_device())
        return DeviceMeter(self.hparams.classes, dev)

    def init_metrics_val(self):
        self.intersection_meter_val, self.union_meter_val, self.target_meter_val = (AverageMeter() for _ in range(3))
        self._val_counts = None

    def init_metrics_test(self):
        self.intersection_meter_test1, self.union_meter_test1, self.target_meter_test1 = (AverageMeter() for _ in range(3))
        self.intersection_meter_test2, self.union_meter_test2, self.target_meter_test2 = (AverageMeter() for _ in range(3))
        self._test_counts = [None, None]

    def _profiler(self):
        trainer = getattr(self, "_trainer", None) if pl is not None else getattr(self, "trainer", None)
        prof = getattr(trainer, "profiler", None) if trainer is not None else None
        return prof if prof is not None else self._local_profiler

    # ------------------------------------------------------------------ forward / val / test
    def forward(self, frame_prev, frame_next, mvs_left, mvs_right, left_index, right_index):
        """flow/base.py:134-135."""
        return self.model_G(None, frame_prev, frame_next, mvs_left, mvs_right, left_index, right_index)

    # ------------------------------------------------------------------ sliding-crop inference
    @staticmethod
    def crop_windows(new_h, new_w, crop_h, crop_w, stride_rate=2 / 3):
        """flow/base.py:183-200: the (s_h, e_h, s_w, e_w) windows in the order the reference visits them."""
        stride_h, stride_w = int(np.ceil(crop_h * stride_rate)), int(np.ceil(crop_w * stride_rate))
        grid_h = int(np.ceil(float(new_h - crop_h) / stride_h) + 1)
        grid_w = int(np.ceil(float(new_w - crop_w) / stride_w) + 1)
        for index_h in range(grid_h):
            for index_w in range(grid_w):
                e_h = min(index_h * stride_h + crop_h, new_h)
                e_w = min(index_w * stride_w + crop_w, new_w)
                yield e_h - crop_h, e_h, e_w - crop_w, e_w

    @staticmethod
    def crop_motion_vector(mvs_left, mvs_right, height, width, crop_h, crop_w, h_off, w_off):
        """flow/transform.py:215-261 on the device (kernels.crop_grid).  Grids without a [Hg,Wg,2] shape — the [B,1]
        dummies of no_warp clips, flow/dataset.py:200-205 — pass through unchanged (:216-221)."""
        def has_grid(m):
            return m is not None and isinstance(m, (list, tuple)) and len(m) > 0 and m[0].dim() >= 3
        if not (has_grid(mvs_left) or has_grid(mvs_right)):
            return mvs_left, mvs_right
        crop = lambda g: kernels.crop_grid(g, height, width, crop_h, crop_w, h_off, w_off)  # noqa: E731
        return ([crop(g) for g in mvs_left] if mvs_left is not None else None,
                [crop(g) for g in mvs_right] if mvs_right is not None else None)

    def compute_output(self, n, function, frame_prev, frame_next, mvs_left, mvs_right, *args, **kwargs):
        """flow/base.py:182-209 -> fp64 canvas [n,classes,H,W] of class probabilities averaged over the crops.

        `function` returns the crop's logits at crop size (compute_test_crop / compute_predict_crop); the soft-max of
        flow/base.py:220,233 is fused with the canvas update.  The arg-max of the finished canvas is kept in
        self.crop_labels (uint8 [n,H,W]) so that callers need not read the 8-byte canvas again."""
        hp = self.hparams
        crop_h, crop_w = hp.test_h, hp.test_w
        _, _, new_h, new_w = frame_prev.shape
        if crop_h > new_h or crop_w > new_w:
            raise ValueError(f"crop {crop_h}x{crop_w} exceeds the frame {new_h}x{new_w} (the reference slices out of range)")
        canvas = torch.zeros((n, hp.classes, new_h, new_w), dtype=torch.float64, device=frame_prev.device)
        count = torch.zeros((new_h, new_w), dtype=torch.float64, device=frame_prev.device)
        for s_h, e_h, s_w, e_w in self.crop_windows(new_h, new_w, crop_h, crop_w):
            prev_c = frame_prev[:, :, s_h:e_h, s_w:e_w].clone()
            next_c = frame_next[:, :, s_h:e_h, s_w:e_w].clone()
            ml, mr = self.crop_motion_vector(mvs_left, mvs_right, new_h, new_w, e_h - s_h, e_w - s_w, s_h, s_w)
            logits = function(prev_c, next_c, ml, mr, *args, **kwargs)
            kernels.crop_accumulate(logits, canvas, count, s_h, s_w)
        self.crop_labels = kernels.crop_finish(canvas, count)
        return canvas

    def _crop_logits(self, output, frame_prev):
        _, _, h_i, w_i = frame_prev.shape
        if output.shape[2] != h_i or output.shape[3] != w_i:                  # flow/base.py:217-219, 230-232
            output = kernels.upsample_bilinear_ac(output, (h_i, w_i))
        return output

    def compute_test_crop(self, frame_prev, frame_next, mvs_left, mvs_right, left_index, right_index):
        """flow/base.py:213-222 without its soft-max (fused into compute_output)."""
        return self._crop_logits(self.forward(frame_prev, frame_next, mvs_left, mvs_right, left_index, right_index)["pred"],
                                 frame_prev)

    def compute_predict_crop(self, frame_prev, *args, **kwargs):
        """flow/base.py:226-234 without its soft-max (fused into compute_output)."""
        return self._crop_logits(self.model_G.predict(frame_prev, *args, **kwargs)["pred"], frame_prev)

    def _labels_from_forward(self, batch):
        outs = self.forward(batch["frame_prev"], batch["frame_next"], batch["mvs_left"], batch["mvs_right"],
                            batch["left_index"], batch["right_index"])
        return kernels.argmax(outs["pred"])            # output.data.max(1)[1]  (flow/base.py:147,167)

    @torch.no_grad()
    def validation_step(self, batch, batch_idx):
        """flow/base.py:141-150."""
        output = self._labels_from_forward(batch)
        if self._val_counts is None:
            self._val_counts = self._meter()
        self._val_counts.update(output, batch["label"], self.hparams.ignore_index)

    @torch.no_grad()
    def test_step(self, batch, batch_idx):
        """flow/base.py:156-176."""
        batch_test, test_idx = batch
        assert batch_test["frame_prev"].shape[0] == 1 and batch_test["label"].shape[0] == 1
        with self._profiler().profile("test_interference"):
            if not self.hparams.no_cropping:                                  # flow/base.py:163-165
                self.compute_output(1, self.compute_test_crop, batch_test["frame_prev"], batch_test["frame_next"],
                                    batch_test["mvs_left"], batch_test["mvs_right"], batch_test["left_index"],
                                    batch_test["right_index"])
                output = self.crop_labels
            else:
                output = self._labels_from_forward(batch_test)
        which = 1 if int(test_idx) > 0 else 0              # Texas video -> test2, Florida -> test1
        if self._test_counts[which] is None:
            self._test_counts[which] = self._meter()
        self._test_counts[which].update(output, batch_test["label"], self.hparams.ignore_index)

    def validation_epoch_metrics(self):
        """base/foundation.py:160-172: one all-reduce of the counts, then the fp64 formulas."""
        if self._val_counts is None:
            return None
        m = self._val_counts.all_reduce()
        self.intersection_meter_val, self.union_meter_val, self.target_meter_val = m.to_meters()
        out = m.metrics()
        self._val_counts = None
        return out

    def test_epoch_metrics(self):
        """base/foundation.py:224-259."""
        res = {}
        for k, m in enumerate(self._test_counts):
            if m is None:
                continue
            m.all_reduce()
            meters = m.to_meters()
            if k == 0:
                self.intersection_meter_test1, self.union_meter_test1, self.target_meter_test1 = meters
            else:
                self.intersection_meter_test2, self.union_meter_test2, self.target_meter_test2 = meters
            res[f"test{k + 1}"] = m.metrics()
        if "test1" in res and "test2" in res:
            res["test"] = {k: (res["test1"][k] + res["test2"][k]) / 2 for k in ("miou", "macc", "accuracy")}
        self._test_counts = [None, None]
        return res

    # ------------------------------------------------------------------ predict
    def on_predict_start(self):
        """flow/base.py:236-255."""
        self.intersection_meter_predict, self.union_meter_predict, self.target_meter_predict = (AverageMeter() for _ in range(3))
        self._predict_counts = self._meter()
        self.last_output = None     # last label map of the previous interval (uint8 [H,W] on device)
        self._predict_intervals = 0
        self.model_G.reset_keyframe_cache()
        hp = self.hparams
        if (hp.save_video or hp.save_images) and self.frame_sink is None:
            # the reference writes an .avi (imageio) / PNGs (PIL) here (flow/base.py:249-253, 297-312): sinks are outside
            # the accelerated path, so the label maps go to `frame_sink` and nothing is written without one
            import warnings
            warnings.warn("FlowBaseModel: save_video / save_images are set but no frame_sink is attached; label maps are "
                          "returned by predict_step and not written anywhere (set model.frame_sink = callable(frame_id, "
                          "uint8 ndarray [n,H,W]))", stacklevel=2)

    @torch.no_grad()
    def predict_step(self, batch, batch_idx):
        """flow/base.py:259-312 -> uint8 label maps [n,H,W] on the device."""
        frame_prev, frame_next = batch["frame_prev"], batch["frame_next"]
        mvs_left, mvs_right = batch["mvs_left"], batch["mvs_right"]
        assert frame_prev.shape[0] == 1
        assert len(mvs_left) == len(mvs_right)
        n = len(mvs_left) + 1
        hp = self.hparams
        prof = self._profiler()
        want_counts = bool(hp.compute_metrics)
        with prof.profile("predict_interference"):
            out_h, out_w = hp.output_size                      # the hard-coded (1072, 1920) of flow/base.py:275
            if not hp.no_cropping:                             # flow/base.py:272-273
                canvas = self.compute_output(n, self.compute_predict_crop, frame_prev, frame_next, mvs_left, mvs_right,
                                             n, prof)
                if (canvas.shape[2], canvas.shape[3]) == (out_h, out_w):
                    output = self.crop_labels                  # the resize at :275 is an identity copy
                else:
                    # fp64 bilinear resize of the probability canvas: a torch op, outside the accelerated path
                    output = F.interpolate(canvas, (out_h, out_w), mode="bilinear", align_corners=True).max(1)[1].to(torch.uint8)
                if want_counts:
                    kernels.temporal_counts(output, hp.classes, hp.ignore_index, tc_prev=self.last_output,
                                            counts=self._predict_counts.counts)
            elif (frame_prev.shape[2], frame_prev.shape[3]) == (out_h, out_w) and not self.model_G.feature_based:
                # the resize at :275 is an identity copy -> fully fused route
                # the frame id is only needed (and only read: it costs a device-to-host sync when Lightning has moved
                # the batch to the GPU) for key-frame reuse
                fid = int(batch["frame_id"][0]) if (self.model_G.reuse_keyframes and "frame_id" in batch) else None
                output = self.model_G.predict_labels(
                    frame_prev, frame_next, mvs_left, mvs_right, n, prof, tc_prev=self.last_output,
                    counts=self._predict_counts.counts if want_counts else None, ignore_index=hp.ignore_index,
                    frame_id=fid)
            else:
                logits = self.model_G.predict(frame_prev, frame_next, mvs_left, mvs_right, n, prof)["pred"]
                output, _ = kernels.upsample_argmax(logits, (out_h, out_w))         # flow/base.py:275-277, one kernel
                if want_counts:
                    kernels.temporal_counts(output, hp.classes, hp.ignore_index, tc_prev=self.last_output,
                                            counts=self._predict_counts.counts)
            output_numpy = None
            if self.frame_sink is not None:
                output_numpy = output.cpu().numpy()                                 # flow/base.py:277 (already uint8)
        if want_counts:
            steps = output.shape[0] - (1 if self.last_output is None else 0)
            self._predict_counts.note_updates(steps)
            self.last_output = output[output.shape[0] - 1]                          # flow/base.py:295
        self._predict_intervals += 1
        if self.frame_sink is not None and output_numpy is not None:
            frame_id = int(batch["frame_id"][0]) if "frame_id" in batch else batch_idx * n
            self.frame_sink(frame_id, output_numpy)
        return output

    def on_predict_end(self):
        """flow/base.py:316-343 -> dict with predict_time_{mean,sum} and the temporal-consistency metrics."""
        res = {}
        d = self._profiler().recorded_durations.get("predict_interference", [])
        if len(d):
            res["predict_time_mean"] = float(np.mean(d))
            res["predict_time_sum"] = float(np.sum(d))
        m = self._predict_counts.all_reduce()
        self.intersection_meter_predict, self.union_meter_predict, self.target_meter_predict = m.to_meters()
        if self.intersection_meter_predict.count > 0:
            e = epoch_metrics(self.intersection_meter_predict.sum, self.union_meter_predict.sum,
                              self.target_meter_predict.sum)
            res.update(predict_miou1_epoch=e["miou"], predict_macc1_epoch=e["macc"],
                       predict_accuracy1_epoch=e["accuracy"], predict_miou1_epoch_classes=e["iou_class"],
                       predict_macc1_epoch_classes=e["accuracy_class"])
        return res
