"""B200-native (sm_100a) implementation of the inter-frame segmentation interpolation path of
lenke182/flood-uav-video-segmentation.

Layout mirrors the reference modules it stands in for:

  flow/model.py        FlowModel                       <- reference flow/model.py
  flow/base.py         FlowBaseModel step logic        <- reference flow/base.py:134-343
  util/util.py         intersectionAndUnion[GPU], AverageMeter <- reference util/util.py
  base/foundation.py   compute_metrics, epoch formulas <- reference base/foundation.py:22-42,160-164,333-344
  dist.py              clip sharding + one NCCL all-reduce of the counts
  kernels.py / _lib.py ctypes binding of libfuvs.so (include/fuvs.h); csrc/ holds the CUDA kernels

The library is loaded lazily on the first op; it is never replaced by a CPU or PyTorch
implementation (see _lib.load()).
"""
from ._lib import FuvsError, LIB_PATH, launch_count, load  # noqa: F401

__all__ = ["FuvsError", "LIB_PATH", "launch_count", "load"]
