"""ncu target: fuvs_confusion on 5 x 1080p int64 predictions / int64 ground truth (the reference's dtypes) and uint8 / int64."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flood_uav_video_segmentation_b200 import kernels  # noqa: E402

dev = torch.device("cuda", 0)
N = 5 * 1080 * 1920
g = torch.Generator(device=dev).manual_seed(0)
tgt = torch.randint(0, 5, (N,), device=dev, generator=g)
tgt[torch.rand(N, device=dev, generator=g) < 0.05] = 255
counts = kernels.new_counts(5, dev)
for pdt in (torch.int64, torch.uint8):
    pred = torch.randint(0, 5, (N,), device=dev, generator=g).to(pdt)
    for _ in range(2):
        kernels.confusion(pred, tgt, 5, counts=counts)
torch.cuda.synchronize()
print("ok", counts.sum().item())
