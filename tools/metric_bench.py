"""Throughput of fuvs_confusion (intersectionAndUnionGPU, util/util.py:52-63) at 5 x 1080p: uint8 / int64 predictions
against int64 ground truth, vs the measured HBM copy peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from flood_uav_video_segmentation_b200 import kernels  # noqa: E402

dev = torch.device("cuda", 0)
N = 5 * 1080 * 1920
g = torch.Generator(device=dev).manual_seed(0)
peak, _ = bench.measured_peak()
res = {}
for pdt in (torch.uint8, torch.int64):
    for tdt in (torch.int64, torch.uint8):
        sets = []
        for _ in range(16):                                   # 16 distinct input sets (> L2) cycled
            pred = torch.randint(0, 5, (N,), device=dev, generator=g).to(pdt)
            tgt = torch.randint(0, 5, (N,), device=dev, generator=g)
            tgt[torch.rand(N, device=dev, generator=g) < 0.05] = 255
            sets.append((pred, tgt.to(tdt)))
        counts = kernels.new_counts(5, dev)
        for p, t in sets[:3]:
            kernels.confusion(p, t, 5, counts=counts)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        gr = torch.cuda.CUDAGraph()                           # replayed: the Python call overhead is ~10 us per call
        with torch.cuda.graph(gr):
            for p, t in sets:
                kernels.confusion(p, t, 5, counts=counts)
        gr.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * len(sets))
        nbytes = N * (pred.element_size() + sets[0][1].element_size())
        res[f"pred_{str(pdt)[6:]}_target_{str(tdt)[6:]}"] = {"us_per_call": us, "GBps": nbytes / us / 1e3, "frac": nbytes / us / 1e3 / peak,
                                                          "bytes": nbytes}
        del sets
        torch.cuda.empty_cache()
print(json.dumps(res, indent=1))

# fuvs_temporal_counts (flow/base.py:280-295) over 5 x 1080p uint8 label maps + the previous interval's last map
HW = 1080 * 1920
sets = []
for _ in range(24):                                        # 24 x 12.4 MB > L2
    sets.append((torch.randint(0, 5, (5, 1080, 1920), device=dev, generator=g, dtype=torch.uint8),
                 torch.randint(0, 5, (1080, 1920), device=dev, generator=g, dtype=torch.uint8)))
counts = kernels.new_counts(5, dev)
for lab, last in sets[:3]:
    kernels.temporal_counts(lab, 5, 255, tc_prev=last, counts=counts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for lab, last in sets:
        kernels.temporal_counts(lab, 5, 255, tc_prev=last, counts=counts)
gr.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    gr.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (10 * len(sets))
print(json.dumps({"temporal_counts_5x1080p": {"us_per_call": us, "GBps": 6 * HW / us / 1e3, "frac": 6 * HW / us / 1e3 / peak}}))
