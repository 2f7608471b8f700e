"""BASELINE.json config 3/4 measurement: share of one interval's GPU time spent in the interpolation path vs the
key-frame network, DeepLabV3-ResNet101 (torchvision, random-init weights; the reference's PSPNet / Segmenter modules
live in the reference tree, which does not travel to the GPU box), 1072x1920, k = 2..10, key-frame reuse on.

python tools/path_fraction.py [--out profiles/r01_path_fraction.json]"""
import argparse
import json
import os
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--height", type=int, default=1072)
    ap.add_argument("--width", type=int, default=1920)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    H, W = a.height, a.width
    bb = DeepLabParts().to(dev).eval()
    prof = SimpleProfiler()
    x = torch.randn(1, 3, H, W, device=dev)
    rows = []
    with torch.no_grad():
        t_key = timed(lambda: bb.decoder(bb.encoder(x)), 3)           # one key frame through the network (bf16 off)
        for mode in ("linear", "block"):
            fm = FlowModel(bb, feature_based=False, no_warp=(mode == "linear")).eval()
            for k in range(2, 11):
                if mode == "linear":
                    gl = gr = [torch.zeros(1, 1, device=dev)] * (k - 1)
                else:
                    gl = [g.to(dev) for g in flow_grids(H, W, k, "block", clip=1, side=0)]
                    gr = [g.to(dev) for g in flow_grids(H, W, k, "block", clip=1, side=1)]
                o = fm._keyframe_logits(x, H, W, prof, keep_lowres=(mode == "linear"))
                o2 = fm._keyframe_logits(x.flip(3), H, W, prof, keep_lowres=(mode == "linear"))
                counts = torch.zeros((3, 5), dtype=torch.int64, device=dev)
                t_int = timed(lambda: fm._run_interval(o, o2, gl, gr, k, want_labels=True, want_logits=False, counts=counts,
                                                       size=(H, W)), 10)
                # with key-frame reuse one network pass per interval; the reference runs two (flow/model.py:189,202)
                rows.append({"mode": mode, "k": k, "keyframe_ms": t_key, "interpolation_ms": t_int,
                             "interpolation_share_reuse": t_int / (t_int + t_key),
                             "interpolation_share_reference_schedule": t_int / (t_int + 2 * t_key),
                             "frames_per_s_reuse": k / ((t_int + t_key) / 1e3)})
                print(rows[-1])
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"model": "deeplabv3_resnet101 (torchvision, random init, fp32)", "height": H, "width": W, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
