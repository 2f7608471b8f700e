"""Developer experiment: do independent clips on S streams fill each other's CTA tails?
usage: python tools/exp_streams.py <mode> <nstreams> [steps]   -> us per interval (graph replay, 4 clips x 3 intervals per step)
Granularity "clip": clip c runs on stream c % S.  FUVS_EXP_GRAN=interval: interval i of a clip on stream i % S with the
temporal-count dependency kept by events (labels of the previous interval).  FUVS_EXP_NOCOUNTS=1: no temporal counts."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from flood_uav_video_segmentation_b200 import _lib, kernels

if os.environ.get("FUVS_DEV_LIB"):          # developer A/B build (build.py FUVS_BUILD_TAG), this tool only
    _lib.use_library(os.environ["FUVS_DEV_LIB"])

mode = sys.argv[1] if len(sys.argv) > 1 else "dense"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = kernels.load()
clips = [bench.make_clip(mode, dev, i) for i in range(4)]
need = max(bench.scratch_floats(kernels, mode), 1)
scratch = [torch.empty((need,), dtype=torch.float32, device=dev) for _ in range(S)]
counts = [None if os.environ.get("FUVS_EXP_NOCOUNTS") else kernels.new_counts(bench.C, dev) for _ in range(S)]
streams = [torch.cuda.Stream(dev) for _ in range(S)]


def step():
    main = torch.cuda.current_stream(dev)
    fork = torch.cuda.Event()
    fork.record(main)
    for s in range(S):
        streams[s].wait_event(fork)
    for c, clip in enumerate(clips):
        s = c % S
        with torch.cuda.stream(streams[s]):
            bench.run_clip(kernels, mode, clip, counts[s], scratch[s])
    for s in range(S):
        e = torch.cuda.Event()
        e.record(streams[s])
        main.wait_event(e)


side = torch.cuda.Stream(dev)
with torch.cuda.stream(side):
    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize(dev)
ms = e0.elapsed_time(e1)
iv = steps * len(clips) * 3
print(f"{mode} streams={S}: "
      f"{ms * 1e3 / iv:.2f} us/interval, frac {bench.algorithmic_bytes(mode) * iv / (ms / 1e3) / 1e9 / 6548.8:.4f}")
