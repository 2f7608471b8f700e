"""Drop-in for the hot-path half of the reference's util/util.py (lines 10-24, 36-63).

intersectionAndUnionGPU / intersectionAndUnion keep their signatures and return values; both run the
warp-aggregated histogram kernel (csrc/metric.cu).  The two functions differ exactly where the reference's
do: torch.histc drops values outside [0, K-1] and mutates the caller's `output`; np.histogram counts a
value == K in class K-1 (closed last bin) and works on a copy.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import kernels


class AverageMeter(object):
    """util/util.py:10-24."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def _as_label_tensor(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if x.dtype not in (torch.uint8, torch.int64):
        x = x.to(torch.int64)
    return x.to(device=device, non_blocking=True)


def intersectionAndUnionGPU(output, target, K, ignore_index=255):
    """util/util.py:52-63 -> (area_intersection, area_union, area_target), int64 [K] CUDA tensors.

    Like the reference, `output[target == ignore_index] = ignore_index` is applied to the caller's tensor
    when `output` is contiguous (reshape(-1) is then a view)."""
    assert output.dim() in [1, 2, 3]
    assert output.shape == target.shape
    kernels.require_cuda(output, target, what="intersectionAndUnionGPU")
    mutate = output.is_contiguous() and output.dtype in (torch.uint8, torch.int64)
    if output.dtype not in (torch.uint8, torch.int64):
        output = output.to(torch.int64)
    if target.dtype not in (torch.uint8, torch.int64):
        target = target.to(torch.int64)
    counts = kernels.confusion(output, target, K, ignore_index, numpy_bins=False, mutate_pred=mutate)
    return counts[0], counts[1], counts[2]


def intersectionAndUnion(output, target, K, ignore_index=255, device=None):
    """util/util.py:36-47 (numpy semantics: copy of `output`, closed last histogram bin) -> 3 x int64 ndarray [K].

    The arrays are staged to the current CUDA device and counted there."""
    assert output.ndim in [1, 2, 3]
    assert output.shape == target.shape
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    o = _as_label_tensor(output, device)
    t = _as_label_tensor(target, device)
    counts = kernels.confusion(o, t, K, ignore_index, numpy_bins=True, mutate_pred=False).cpu().numpy()
    return counts[0], counts[1], counts[2]
