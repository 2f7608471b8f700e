"""Small ncu target: a few intervals of one mode at 1080p through the C ABI (no bench bookkeeping)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from flood_uav_video_segmentation_b200 import kernels  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="dense")
ap.add_argument("--clips", type=int, default=2)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
clips = [bench.make_clip(a.mode, dev, i) for i in range(a.clips)]
counts = kernels.new_counts(bench.C, dev)
scratch = torch.empty((max(bench.scratch_floats(kernels, a.mode), 1),), dtype=torch.float32, device=dev)
for _ in range(a.reps):
    for c in clips:
        bench.run_clip(kernels, a.mode, c, counts, scratch)
torch.cuda.synchronize()
print("ok", counts.sum().item())
