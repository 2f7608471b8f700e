"""Every C-ABI entry of the path that bench.py does not time on its own, at the shapes the reference uses, under
CUDA-graph replay (the Python call overhead is the same order as the short kernels): us per call, achieved GB/s on the
entry's algorithmic bytes against the measured HBM peak, and the stock torch op sequence it replaces beside it.
    python tools/abi_bench.py > profiles/r01_abi_bench.json"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from flood_uav_video_segmentation_b200 import kernels  # noqa: E402
from flood_uav_video_segmentation_b200.synthetic import flow_grids  # noqa: E402

dev = torch.device("cuda", 0)
peak, _ = bench.measured_peak()
res = {}


def timed(fn, reps=10, inner=8):
    """fn(i) is called `inner` times per graph with i = 0..inner-1 (distinct buffers), the graph replayed `reps` times."""
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


def record(name, us, nbytes, torch_us=None, note=None):
    res[name] = {"us_per_call": round(us, 2), "algorithmic_MB": round(nbytes / 1e6, 2), "GBps": round(nbytes / us / 1e3, 1),
                 "frac_of_hbm_peak": round(nbytes / us / 1e3 / peak, 3)}
    if torch_us is not None:
        res[name]["stock_torch_us"] = round(torch_us, 2)
    if note:
        res[name]["note"] = note


C, H, W, n = 5, 1080, 1920, 5
# ---- fuvs_argmax (flow/base.py:147,167,276): [n,C,H,W] logits -> labels
xs = [torch.randn(n, C, H, W, device=dev) for _ in range(4)]                       # 4 x 207 MB
us = timed(lambda i: kernels.argmax(xs[i % 4]))
ut = timed(lambda i: xs[i % 4].max(1)[1])
record("argmax_5x5x1080p_uint8", us, n * C * H * W * 4 + n * H * W, ut, "torch writes int64 labels")
us = timed(lambda i: kernels.argmax(xs[i % 4], dtype=torch.int64))
record("argmax_5x5x1080p_int64", us, n * C * H * W * 4 + n * H * W * 8, ut)

# ---- fuvs_blend_argmax (flow/model.py:104, val/test route): wa*a + wb*b (+ labels)
us = timed(lambda i: kernels.blend_argmax(xs[i % 4], xs[(i + 1) % 4], 0.6, 0.4, want_labels=True))
ut = timed(lambda i: (0.6 * xs[i % 4] + 0.4 * xs[(i + 1) % 4]).max(1)[1])
record("blend_argmax_5x5x1080p", us, 3 * n * C * H * W * 4 + n * H * W, ut)
del xs
torch.cuda.empty_cache()

# ---- fuvs_warp_step (flow/model.py:244-249), reference wire format: 67x120 block grid over a full-resolution key frame
key = [torch.randn(C, H, W, device=dev) for _ in range(4)]
grids = [g.to(dev) for g in flow_grids(H, W, 5, "block", clip=0, side=0)]
us = timed(lambda i: kernels.warp_step(key[i % 4], grids[i % 4]))
ut = timed(lambda i: F.grid_sample(key[i % 4][None], grids[i % 4], mode="bilinear", padding_mode="border", align_corners=False))
hg, wg = grids[0].shape[1:3]
record("warp_step_block_grid_5ch_1080p", us, hg * wg * (8 + 4 * C * 4 + C * 4), ut, "latency-bound: 8 040 points")
del key

# ---- feature-sized work (model.feature_based=True, DeepLabv3 features at 1080p)
Cf, fh, fw = 2048, 135, 240
feats = [torch.randn(Cf, fh, fw, device=dev) for _ in range(2)]                    # 265 MB each
us = timed(lambda i: kernels.warp_step(feats[i % 2], grids[i % 4]), inner=4)
ut = timed(lambda i: F.grid_sample(feats[i % 2][None], grids[i % 4], mode="bilinear", padding_mode="border", align_corners=False), inner=4)
record("warp_step_block_grid_2048ch_135x240", us, hg * wg * (8 + Cf * 4) + Cf * fh * fw * 4, ut,
       "bytes: source read once + output; taps touch the whole source")
gl = [g.to(dev) for g in flow_grids(H, W, n, "block", clip=1, side=0)]
gr = [g.to(dev) for g in flow_grids(H, W, n, "block", clip=1, side=1)]
out = torch.empty((n, Cf, fh, fw), device=dev)
need = int(kernels.load().fuvs_feature_scratch_floats(Cf, hg, wg, 0, 0, n))
scratch = torch.empty((need,), device=dev)
us = timed(lambda i: kernels.feature_interval(feats[0], feats[1], gl, gr, n, scratch=scratch, out=out), reps=5, inner=2)


def torch_feature(i):
    # flow/model.py:131-173 with no default grid: chains at grid resolution, per-step up-sample, blend, cat
    l, r = feats[0][None], feats[1][None]
    ls, rs = [], []
    for j in range(n - 1):
        l = F.grid_sample(l, gl[j], mode="bilinear", padding_mode="border", align_corners=False)
        r = F.grid_sample(r, gr[j], mode="bilinear", padding_mode="border", align_corners=False)
        ls.append(F.interpolate(l, size=(fh, fw), mode="bilinear", align_corners=True))
        rs.append(F.interpolate(r, size=(fh, fw), mode="bilinear", align_corners=True))
    frames = [feats[0][None]]
    for p in range(1, n):
        frames.append((n - p) / n * ls[p - 1] + p / n * rs[n - p - 1])
    return torch.cat(frames)


ut = timed(torch_feature, reps=3, inner=1)
record("feature_interval_2048x135x240_k5", us, (2 + n) * Cf * fh * fw * 4 + 2 * (n - 1) * 2 * Cf * hg * wg * 4, ut,
       "bytes: two key-frame feature maps in, n frames out, chain states written and read once")
print(json.dumps(res, indent=1))
