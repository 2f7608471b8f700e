"""Developer experiment: stream time of one fuvs_temporal_counts call (5 x 1080p label maps + tc_prev, L2-resident),
20 calls per graph replay.  FUVS_DEV_LIB selects an A/B build.  usage: python tools/exp_counts.py [n] [H] [W]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flood_uav_video_segmentation_b200 import _lib, kernels

if os.environ.get("FUVS_DEV_LIB"):
    _lib.use_library(os.environ["FUVS_DEV_LIB"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
g = torch.Generator(device=dev).manual_seed(0)
labels = torch.randint(0, 5, (n, H, W), device=dev, dtype=torch.uint8, generator=g)
prev = torch.randint(0, 5, (H, W), device=dev, dtype=torch.uint8, generator=g)
counts = kernels.new_counts(5, dev)
side = torch.cuda.Stream(dev)
with torch.cuda.stream(side):
    for _ in range(3):
        kernels.temporal_counts(labels, 5, tc_prev=prev, counts=counts)
    torch.cuda.synchronize(dev)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=side):
        for _ in range(20):
            kernels.temporal_counts(labels, 5, tc_prev=prev, counts=counts)
    for _ in range(3):
        gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        gr.replay()
    e1.record()
    torch.cuda.synchronize(dev)
print(f"temporal_counts n={n} {H}x{W}: {e0.elapsed_time(e1) * 1e3 / 400:.2f} us per call")
