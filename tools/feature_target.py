"""Developer target: fuvs_feature_interval at [cf,135,240] (default 2048) — timing line, or a plain run for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from flood_uav_video_segmentation_b200 import kernels  # noqa: E402

cf = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda", 0)
us, nbytes = bench.time_feature(kernels, cf, dev)
peak, _ = bench.measured_peak()
print(f"feature_interval [{cf},135,240] k=5: {us:.1f} us, {nbytes / us / 1e3:.0f} GB/s, frac {nbytes / us / 1e3 / peak:.3f}")
