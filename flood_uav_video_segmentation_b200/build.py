"""Builds libfuvs.so (hand-written CUDA for sm_100a) in-tree.

`python -m flood_uav_video_segmentation_b200.build` or `build_library()`.
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
Parity-critical flag: -fmad=false (every FMA in the kernels is explicit, see
csrc/fuvs_common.cuh).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libfuvs.so")
SOURCES = ["abi.cu", "linear.cu", "warp.cu", "dense_tma.cu", "dense_strip.cu", "block.cu", "block_rows.cu", "pointwise.cu", "feature.cu", "crop.cu", "metric.cu", "calib.cu", "comm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v",
]
# developer A/B builds: FUVS_BUILD_TAG=x FUVS_BUILD_DEFINES="-DFOO -DBAR" writes lib/libfuvs_x.so (objects in build_x/);
# a tool selects it with _lib.use_library(path).  The shipped library is always the untagged default build.
_TAG = os.environ.get("FUVS_BUILD_TAG", "")
if _TAG:
    LIB = os.path.join(LIBDIR, f"libfuvs_{_TAG}.so")
    NVCC_FLAGS += os.environ.get("FUVS_BUILD_DEFINES", "").split()


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libfuvs.so cannot be built (there is no CPU fallback)")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build" + (f"_{_TAG}" if _TAG else ""))
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "fuvs.h"))
    headers.append(os.path.abspath(__file__))

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log = os.path.join(objdir, src + ".log")
            with open(log, "w") as f:
                f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
