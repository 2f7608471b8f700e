"""Interval kernels at the reference's crop size (433 x 433, data.train_w / test_w): HW is odd, so no plane but the
first is 16-byte aligned and the 128-bit paths do not apply.  Graph-replayed, 432 x 432 beside it for comparison."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flood_uav_video_segmentation_b200 import kernels  # noqa: E402
from flood_uav_video_segmentation_b200.synthetic import flow_grids  # noqa: E402

dev = torch.device("cuda", 0)
C, n = 5, 5
res = {}


def timed(fn, reps=20, inner=16):
    for i in range(inner):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(inner):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * inner)


for H, W in ((433, 433), (432, 432)):
    keys = [torch.randn(C, H, W, device=dev) for _ in range(17)]
    tc = torch.randint(0, C, (H, W), device=dev, dtype=torch.uint8)
    counts = kernels.new_counts(C, dev)
    r = {}
    r["linear"] = timed(lambda i: kernels.linear_blend_argmax(keys[i], keys[i + 1], n, tc_prev=tc, counts=counts))
    gl = [g.to(dev) for g in flow_grids(H, W, n, "block", clip=0, side=0)]
    gr = [g.to(dev) for g in flow_grids(H, W, n, "block", clip=0, side=1)]
    r["block"] = timed(lambda i: kernels.block_interval(keys[i], keys[i + 1], gl, gr, n, tc_prev=tc, counts=counts))
    dl = [g.to(dev) for g in flow_grids(H, W, n, "dense", clip=0, side=0)]
    dr = [g.to(dev) for g in flow_grids(H, W, n, "dense", clip=0, side=1)]
    r["dense"] = timed(lambda i: kernels.dense_interval(keys[i], keys[i + 1], dl, dr, n, tc_prev=tc, counts=counts))
    res[f"{H}x{W}"] = {k: round(v, 2) for k, v in r.items()}
print(json.dumps(res, indent=1))
