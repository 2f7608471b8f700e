#!/bin/bash
# Self-checking run of the dense strip kernel and the linear bulk kernel (compute-sanitizer is closed on the pool): builds libfuvs with
# -DFUVS_STRIP_ASSERT (device-side assert() on every ring address, slot number, global index and completion-counter value)
# and runs the dense parity tests and one bench-sized clip against it.  A failing assert aborts the process.
set -e
# usage: tools/strip_asserts.sh [extra nvcc defines, e.g. -DFUVS_STRIP_JITTER]
FUVS_BUILD_TAG=assert FUVS_BUILD_DEFINES="-DFUVS_STRIP_ASSERT $*" python -m flood_uav_video_segmentation_b200.build --force > /dev/null 2>&1
export FUVS_DEV_LIB=flood_uav_video_segmentation_b200/lib/libfuvs_assert.so
python - <<'PY'
import os, sys, subprocess
sys.path.insert(0, os.getcwd())
from flood_uav_video_segmentation_b200 import _lib
_lib.use_library(os.environ["FUVS_DEV_LIB"])
import pytest
rc = pytest.main(["tests/test_kernels_gpu.py", "tests/test_golden_gpu.py", "tests/test_flowmodel_gpu.py", "-q", "-m", "gpu", "-x",
                  "-k", "dense or linear or 1080p or golden or predict_matches or reuse", "-p", "no:cacheprovider"])
print("pytest under -DFUVS_STRIP_ASSERT rc =", int(rc))
import torch, bench
from flood_uav_video_segmentation_b200 import kernels
dev = torch.device("cuda", 0)
for mode in ("dense", "dense_smooth", "dense_lowres", "linear"):
    clip = bench.make_clip(mode, dev, 3)
    scratch = torch.empty((max(bench.scratch_floats(kernels, mode), 1),), dtype=torch.float32, device=dev)
    counts = kernels.new_counts(bench.C, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bench.run_clip(kernels, mode, clip, counts, scratch)
    e0.record()
    bench.run_clip(kernels, mode, clip, counts, scratch)
    e1.record()
    torch.cuda.synchronize()
    print(mode, "1080p clip x2 under asserts: ok, temporal intersections", int(counts[0].sum()),
          f"({e0.elapsed_time(e1) * 1e3 / 3:.0f} us per interval, eager launches)")
sys.exit(int(rc))
PY
