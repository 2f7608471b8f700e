"""ncu target: one fuvs_feature_interval at DeepLabv3 feature size (2048 x 135 x 240, 67x120 block grids, k = 5)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flood_uav_video_segmentation_b200 import kernels  # noqa: E402
from flood_uav_video_segmentation_b200.synthetic import flow_grids  # noqa: E402

dev = torch.device("cuda", 0)
Cf, fh, fw, n = 2048, 135, 240, 5
feats = [torch.randn(Cf, fh, fw, device=dev) for _ in range(2)]
gl = [g.to(dev) for g in flow_grids(1080, 1920, n, "block", clip=1, side=0)]
gr = [g.to(dev) for g in flow_grids(1080, 1920, n, "block", clip=1, side=1)]
out = torch.empty((n, Cf, fh, fw), device=dev)
for _ in range(2):
    kernels.feature_interval(feats[0], feats[1], gl, gr, n, out=out)
torch.cuda.synchronize()
print("ok")
