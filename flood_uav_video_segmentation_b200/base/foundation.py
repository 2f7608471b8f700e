"""Drop-in for the hot-path pieces of the reference's base/foundation.py.

is_cpu            <- base/foundation.py:22-23
round_train       <- base/foundation.py:34-42
compute_metrics   <- BaseModel.compute_metrics, base/foundation.py:333-344
epoch_metrics     <- validation_epoch_end / test_epoch_end formulas, base/foundation.py:162-164, 226-230
DeviceMeter       device-resident (I,U,T) int64 accumulator replacing the three per-step AverageMeter updates
                  (base/foundation.py:87-106); materialised once per epoch, after the NCCL all-reduce
"""
from __future__ import annotations

import numpy as np
import torch

from .. import kernels
from ..util.util import AverageMeter, intersectionAndUnion, intersectionAndUnionGPU


def is_cpu():
    return torch.cuda.device_count() == 0


def round_train(train, arch):
    if arch == "pspnet":
        return (train - 1) // 8 * 8 + 1
    elif arch == "vit":
        return train // 32 * 32
    elif arch == "deeplabv3":
        return (train - 1) // 8 * 8 + 1
    else:
        assert False


def compute_metrics(output, target, classes, ignore_index=255, cpu=False):
    """base/foundation.py:333-344 -> three numpy [K] arrays.

    The reference picks numpy when `cpu`, when the process has no CUDA device, or when both arguments are
    ndarrays, and torch.histc otherwise; the same rule selects the binning convention here.  Both are counted
    on the GPU — a process without a CUDA device cannot run this package."""
    if is_cpu():
        raise kernels.FuvsError("compute_metrics: no CUDA device in this process; the B200 path has no CPU implementation")
    if cpu or (isinstance(output, np.ndarray) and isinstance(target, np.ndarray)):
        if isinstance(output, torch.Tensor):
            output = output.detach()
        if isinstance(target, torch.Tensor):
            target = target.detach()
        return intersectionAndUnion(_np_or_tensor(output), _np_or_tensor(target), classes, ignore_index)
    if isinstance(output, np.ndarray):
        output = torch.from_numpy(output).cuda()
    if isinstance(target, np.ndarray):
        target = torch.from_numpy(target).to(output.device)
    i, u, t = intersectionAndUnionGPU(output, target, classes, ignore_index)
    host = torch.stack([i, u, t]).cpu().numpy()
    return host[0], host[1], host[2]


def _np_or_tensor(x):
    if isinstance(x, torch.Tensor):
        return x.cpu().numpy()
    return x


def epoch_metrics(intersection_sum, union_sum, target_sum):
    """base/foundation.py:162-164 in fp64 numpy."""
    intersection_sum = np.asarray(intersection_sum)
    union_sum = np.asarray(union_sum)
    target_sum = np.asarray(target_sum)
    iou_class = intersection_sum / (union_sum + 1e-10)
    accuracy_class = intersection_sum / (target_sum + 1e-10)
    return {
        "miou": np.mean(iou_class),
        "macc": np.mean(accuracy_class),
        "accuracy": sum(intersection_sum) / (sum(target_sum) + 1e-10),
        "iou_class": iou_class,
        "accuracy_class": accuracy_class,
    }


class DeviceMeter:
    """(intersection, union, target) int64 counts kept on the GPU for a whole epoch."""

    def __init__(self, classes, device=None):
        self.classes = classes
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.counts = kernels.new_counts(classes, self.device)
        self.updates = 0
        self._reduced = False

    def reset(self):
        self.counts.zero_()
        self.updates = 0
        self._reduced = False

    def update(self, output, target, ignore_index=255):
        """One compute_metrics(...) + three AverageMeter.update(...) of the reference, without leaving the device."""
        if getattr(self, "_reduced", False):
            raise kernels.FuvsError("DeviceMeter.update after all_reduce: reset() the meter at the start of the next epoch")
        kernels.confusion(output, target, self.classes, ignore_index, counts=self.counts,
                          mutate_pred=output.is_contiguous())
        self.updates += 1

    def note_updates(self, steps):
        """Metric calls that were fused into an interval kernel (FlowBaseModel.predict_step) count like update()."""
        self.updates += max(int(steps), 0)

    def all_reduce(self):
        """Single NCCL all-reduce (sum, int64, 3K + 1 values) over the job: the counts and the number
        of metric updates travel together, so a rank that saw no interval still reports the job's metrics.  Idempotent
        within an epoch: a second call returns the already reduced counts instead of summing them again."""
        import torch.distributed as dist
        if getattr(self, "_reduced", False):
            return self
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from ..dist import allreduce_counts
            buf = torch.cat([self.counts.reshape(-1), torch.tensor([self.updates], dtype=torch.int64, device=self.device)])
            allreduce_counts(buf)              # fuvs_allreduce_counts once dist.init_fuvs_comm() has run, else torch.distributed
            self.counts.copy_(buf[:-1].reshape(self.counts.shape))
            self.updates = int(buf[-1].item())
        self._reduced = True
        return self

    def to_meters(self):
        """Materialises the counts as the reference's three AverageMeter objects (sum = numpy [K])."""
        host = self.counts.cpu().numpy()
        meters = []
        for row in host:
            m = AverageMeter()
            if self.updates > 0:
                m.update(row.copy())
                m.count = self.updates
            meters.append(m)
        return tuple(meters)

    def metrics(self):
        host = self.counts.cpu().numpy()
        return epoch_metrics(host[0], host[1], host[2])
