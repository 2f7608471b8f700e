"""Developer tool: where does a dense strip CTA spend its time outside the block loop?
Needs the instrumented build:  FUVS_BUILD_TAG=tm FUVS_BUILD_DEFINES=-DFUVS_STRIP_TIMING python -m flood_uav_video_segmentation_b200.build
Runs a few 1080p intervals and prints, per step kind, SM-clock cycles (mean over CTAs): start -> wait passed (hostage time
behind the previous kernel), wait -> first block's window complete (ring warm-up), block loop per block, and the skew
between the first and the last warp of a CTA at the end."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from flood_uav_video_segmentation_b200 import _lib, kernels

_lib.use_library(os.path.join(ROOT, "flood_uav_video_segmentation_b200", "lib", "libfuvs_tm.so"))
mode = sys.argv[1] if len(sys.argv) > 1 else "dense"
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = kernels.load()
MAX_GRID, NW = 255, 16
buf = torch.zeros((64, MAX_GRID, NW, 5), dtype=torch.int64, device=dev)
lib.fuvs_dev_set_strip_timing.argtypes = [ctypes.c_void_p]
lib.fuvs_dev_set_strip_timing.restype = ctypes.c_int
assert lib.fuvs_dev_set_strip_timing(buf.data_ptr()) == 0
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # streams (clips alternate over them, like bench.py)
clips = [bench.make_clip(mode, dev, i) for i in range(S)]
scratch = [torch.empty((bench.scratch_floats(kernels, mode),), dtype=torch.float32, device=dev) for _ in range(S)]
counts = [kernels.new_counts(bench.C, dev) for _ in range(S)]
streams = [torch.cuda.Stream(dev) for _ in range(S)]
torch.cuda.synchronize()
for _ in range(4):                      # 3 intervals per clip x 4 steps = 12 launches per clip and pass
    for s in range(S):
        with torch.cuda.stream(streams[s]):
            bench.run_clip(kernels, mode, clips[s], counts[s], scratch[s])
torch.cuda.synchronize()
b = buf.cpu().numpy()
nl = 48 * S
for step in range(4):
    rows = []
    for seq in range(nl - 12 * S + step, nl, 4):
        d = b[seq % 64]
        d = d[d[:, 0, 4] > 0]
        start, wait, first, end, nb = (d[:, :, i] for i in range(5))
        cta_end = end.max(1)
        rows.append([(wait - start).mean(), (first - wait).mean(), ((cta_end[:, None] - first).mean(1) / nb[:, 0]).mean(),
                     (cta_end - end.min(1)).mean(), (cta_end - start[:, 0]).mean(), nb[:, 0].mean(), len(d)])
    r = np.mean(rows, 0)
    print(f"step {step + 1}: start->wait {r[0]:8.0f}  wait->first block {r[1]:8.0f}  per block {r[2]:8.0f}  end skew first..last warp {r[3]:8.0f}"
          f"  CTA total {r[4]:9.0f} cycles, {r[5]:.1f} blocks/CTA, {int(r[6])} CTAs")
