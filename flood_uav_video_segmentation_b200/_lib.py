"""ctypes binding of libfuvs.so (C ABI declared in include/fuvs.h).

There is deliberately no CPU implementation behind this module: if the
library is missing, or the current device is not an sm_100 GPU, every op
raises.  Build the library with `python -m flood_uav_video_segmentation_b200.build`
(or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfuvs.so")

FUVS_ABI_VERSION = 3
FUVS_BINS_HISTC = 0
FUVS_BINS_NPHIST = 1
FUVS_MUTATE_PRED = 2
FUVS_MAX_FRAMES = 60

_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_d = C.c_double

# name -> (restype, argtypes); mirrors include/fuvs.h and include/fuvs_calib.h
SIGNATURES = {
    "fuvs_abi_version": (_i, []),
    "fuvs_last_error": (C.c_char_p, []),
    "fuvs_launch_count": (_ll, []),
    "fuvs_comm_unique_id": (_i, [_p]),
    "fuvs_comm_init": (_i, [_p, _i, _i]),
    "fuvs_comm_world_size": (_i, []),
    "fuvs_allreduce_counts": (_i, [_p, _ll, _p]),
    "fuvs_comm_destroy": (_i, []),
    "fuvs_device_ok": (_i, []),
    "fuvs_linear_blend_argmax": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p]),
    "fuvs_linear_lowres_supported": (_i, [_i, _i, _i]),
    "fuvs_linear_lowres_blend_argmax": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p]),
    "fuvs_warp_step": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "fuvs_dense_scratch_floats": (_ll, [_i, _i, _i, _i]),
    "fuvs_dense_interval": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p]),
    "fuvs_dense_interval_ptrs": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p]),
    "fuvs_dense_lowres_interval_ptrs": (_i, [_p, _p, _i, _i, _p, _i, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p]),
    "fuvs_block_scratch_floats": (_ll, [_i, _i, _i, _i]),
    "fuvs_block_interval": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p]),
    "fuvs_block_interval_ptrs": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p]),
    "fuvs_block_lowres_supported": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "fuvs_block_lowres_interval_ptrs": (_i, [_p, _p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p]),
    "fuvs_block_clip": (_i, [_i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p]),
    "fuvs_feature_scratch_floats": (_ll, [_i, _i, _i, _i, _i, _i]),
    "fuvs_feature_interval": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "fuvs_feature_interval_ptrs": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "fuvs_upsample_bilinear_ac": (_i, [_p, _p, _ll, _i, _i, _i, _i, _p]),
    "fuvs_upsample_argmax": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "fuvs_blend_argmax": (_i, [_p, _p, _d, _d, _i, _i, _ll, _p, _p, _p]),
    "fuvs_argmax": (_i, [_p, _i, _i, _ll, _p, _p, _p]),
    "fuvs_confusion": (_i, [_p, _i, _p, _i, _ll, _i, _i, _i, _p, _p]),
    "fuvs_temporal_counts": (_i, [_p, _i, _ll, _p, _i, _i, _p, _p]),
    "fuvs_crop_grid_shape": (_i, [_i, _i, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "fuvs_crop_grid": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "fuvs_crop_accumulate": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "fuvs_crop_finish": (_i, [_p, _p, _i, _i, _ll, _p, _p]),
    "fuvs_calib_grid_sample": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "fuvs_calib_upsample": (_i, [_p, _p, _ll, _i, _i, _i, _i, _i, _i, _i, _p]),
    "fuvs_calib_default": (_i, []),
}

_lock = threading.Lock()
_lib = None


class FuvsError(RuntimeError):
    """A libfuvs entry point returned a negative code."""


def use_library(path: str) -> None:
    """Developer hook (tools/ only): load another build of the library (build.py FUVS_BUILD_TAG) instead of the shipped
    one.  Must be called before the first op; the product never calls it and no environment variable selects a library."""
    global LIB_PATH, _lib
    if _lib is not None:
        raise FuvsError("use_library() must be called before the library is loaded")
    LIB_PATH = path


def ptr_array(tensors):
    """Host array of device pointers (const float* const*) for the *_ptrs / *_clip entries; None -> NULL."""
    if tensors is None:
        return None
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def load() -> C.CDLL:
    """Loads libfuvs.so once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise FuvsError(
                f"{LIB_PATH} is missing: build it with `python -m flood_uav_video_segmentation_b200.build` "
                "(nvcc, sm_100a). This package has no CPU or PyTorch fallback for the interpolation path."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError => header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.fuvs_abi_version() != FUVS_ABI_VERSION:
            raise FuvsError(f"libfuvs ABI {lib.fuvs_abi_version()} != binding {FUVS_ABI_VERSION}; rebuild the library")
        _lib = lib
    return _lib


def last_error() -> str:
    return load().fuvs_last_error().decode("utf-8", "replace")


def check(code: int) -> None:
    if code != 0:
        raise FuvsError(f"libfuvs error {code}: {last_error()}")


def launch_count() -> int:
    return int(load().fuvs_launch_count())


def ptr(t) -> int | None:
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors, what: str) -> torch.device:
    """All tensors must live on one CUDA device; returns it."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise FuvsError(
                f"{what}: expected CUDA tensors (got {type(t).__name__} on "
                f"{getattr(t, 'device', 'host')}); the B200 interpolation path has no CPU implementation"
            )
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise FuvsError(f"{what}: tensors on different devices ({dev} vs {t.device})")
    if dev is None:
        raise FuvsError(f"{what}: no tensor arguments")
    return dev
