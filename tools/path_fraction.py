"""BASELINE.json configs[2] / [3]: share of one interval's GPU time spent in the interpolation path vs the key-frame
network, for k = 2..10, random-init weights, 1072x1920:

  * DeepLabV3-ResNet101 (torchvision modules, split into encoder / decoder like model/deeplabv3.py:47-54)
  * FlowPSPNet-101 — the REFERENCE's own module (model/pspnet.py:113-141), imported from oracle/_ref (the unmodified copy
    oracle/make_ref.py makes; `pretrained=False`, so no checkpoint is read)
  * Segmenter ViT (model/vit.py) needs timm / mmcv, which are not installed: not measured.

Per model: segmentation-based routes (linear = no_warp, block-grid warp) through FlowModel._run_interval, and the
feature-based route (fuvs_feature_interval + the decoder on the [k,Cf,fh,fw] batch).  Key-frame reuse on: one network
pass per interval; the reference's schedule runs two (flow/model.py:189,202).

python tools/path_fraction.py [--out profiles/r02_path_fraction.json]"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flood_uav_video_segmentation_b200 import kernels  # noqa: E402
from flood_uav_video_segmentation_b200.flow.model import FlowModel  # noqa: E402
from flood_uav_video_segmentation_b200.flow.base import SimpleProfiler  # noqa: E402
from flood_uav_video_segmentation_b200.synthetic import flow_grids  # noqa: E402


class DeepLabParts(nn.Module):
    """encoder = ResNet-101 body (2048 ch, stride 8), decoder = DeepLabV3 head, like model/deeplabv3.py:47-54."""

    def __init__(self, classes=5):
        super().__init__()
        import torchvision
        torch.manual_seed(0)
        m = torchvision.models.segmentation.deeplabv3_resnet101(weights=None, weights_backbone=None, num_classes=classes)
        body = m.backbone

        class Enc(nn.Module):
            def forward(self, x):
                return body(x)["out"]

        self.encoder, self.decoder = Enc(), m.classifier
        self._body = body


def reference_pspnet(classes=5):
    from oracle import make_ref
    if make_ref.load() is None:
        raise RuntimeError("oracle/_ref is missing: run python oracle/make_ref.py where /root/reference exists")
    from model.pspnet import FlowPSPNet          # the reference's module, from oracle/_ref
    torch.manual_seed(0)
    return FlowPSPNet(hparams=SimpleNamespace(layers=101, pretrained=False, classes=classes))


def timed(fn, reps):
    fn()                                   # warm-up (allocator, cuDNN autotune, first-launch costs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def measure(name, bb, H, W, dev, ks):
    prof = SimpleProfiler()
    x = torch.randn(1, 3, H, W, device=dev)
    rows = []
    with torch.no_grad():
        t_key = timed(lambda: bb.decoder(bb.encoder(x)), 3)           # one key frame through the network (fp32)
        f0, f1 = bb.encoder(x), bb.encoder(x.flip(3))
        t_enc = timed(lambda: bb.encoder(x), 3)
        cf, fh, fw = f0.shape[1:]
        for mode in ("linear", "block"):
            fm = FlowModel(bb, feature_based=False, no_warp=(mode == "linear")).eval()
            o = fm._keyframe_logits(x, H, W, prof, keep_lowres=(mode == "linear"))
            o2 = fm._keyframe_logits(x.flip(3), H, W, prof, keep_lowres=(mode == "linear"))
            for k in ks:
                if mode == "linear":
                    gl = gr = [torch.zeros(1, 1, device=dev)] * (k - 1)
                else:
                    gl = [g.to(dev) for g in flow_grids(H, W, k, "block", clip=1, side=0)]
                    gr = [g.to(dev) for g in flow_grids(H, W, k, "block", clip=1, side=1)]
                counts = torch.zeros((3, 5), dtype=torch.int64, device=dev)
                t_int = timed(lambda: fm._run_interval(o, o2, gl, gr, k, want_labels=True, want_logits=False, counts=counts,
                                                       size=(H, W)), 10)
                rows.append({"route": f"segmentation-based, {mode}", "k": k, "keyframe_ms": t_key, "interpolation_ms": t_int,
                             "interpolation_share_reuse": t_int / (t_int + t_key),
                             "interpolation_share_reference_schedule": t_int / (t_int + 2 * t_key),
                             "frames_per_s_reuse": k / ((t_int + t_key) / 1e3)})
                print(name, rows[-1])
        # feature-based: warp / blend the encoder features, ONE decoder call on the [k,Cf,fh,fw] batch (flow/model.py:116-181)
        scratch = kernels.ScratchCache()
        dgrid = FlowModel(bb).default_motion_vector.to(dev)
        for k in ks:
            gl = [g.to(dev) for g in flow_grids(H, W, k, "block", clip=2, side=0)]
            gr = [g.to(dev) for g in flow_grids(H, W, k, "block", clip=2, side=1)]
            out = torch.empty((k, cf, fh, fw), device=dev)
            t_int = timed(lambda: kernels.feature_interval(f0[0], f1[0], gl, gr, k, default_grid=dgrid, scratch=scratch, out=out), 5)
            t_dec = timed(lambda: bb.decoder(out), 2)
            rows.append({"route": "feature-based, block", "k": k, "encoder_ms": t_enc, "decoder_batch_ms": t_dec,
                         "interpolation_ms": t_int, "feature_shape": [cf, fh, fw],
                         "interpolation_share_reuse": t_int / (t_int + t_enc + t_dec),
                         "interpolation_share_reference_schedule": t_int / (t_int + 2 * t_enc + t_dec),
                         "frames_per_s_reuse": k / ((t_int + t_enc + t_dec) / 1e3)})
            print(name, rows[-1])
            del out
            torch.cuda.empty_cache()
    return {"model": name, "rows": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--height", type=int, default=1072)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--kmax", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    H, W = a.height, a.width
    ks = list(range(2, a.kmax + 1))
    res = {"height": H, "width": W, "precision": "fp32, random init", "models": [],
           "not_measured": "Segmenter ViT (model/vit.py): timm and mmcv are not installed in this image"}
    for name, make in (("deeplabv3_resnet101 (torchvision modules)", lambda: DeepLabParts()),
                       ("FlowPSPNet-101 (reference model/pspnet.py from oracle/_ref)", reference_pspnet)):
        try:
            bb = make().to(dev).eval()
            res["models"].append(measure(name, bb, H, W, dev, ks))
            del bb
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001
            res["models"].append({"model": name, "error": repr(exc)})
            print(name, "failed:", repr(exc))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
