"""Restatement of the reference's mIoU counting (TEST INFRASTRUCTURE, see oracle/__init__.py).

intersection_and_union_np     <- util/util.py:36-47  (numpy; np.histogram closes the last bin)
intersection_and_union_torch  <- util/util.py:52-63  (torch.histc; int64 histc exists only on CUDA)
intersection_and_union_histc_ints <- the same, with histc's published integer binning written out so the
                                 convention can be checked on a CPU-only box
temporal_consistency_counts   <- flow/base.py:280-295
epoch_metrics                 <- base/foundation.py:162-164, 226-230; flow/base.py:332-336
"""
from __future__ import annotations

import numpy as np
import torch


def intersection_and_union_np(output, target, K, ignore_index=255):
    """util/util.py:36-47."""
    assert output.ndim in (1, 2, 3) and output.shape == target.shape
    out = output.reshape(output.size).copy()
    tgt = target.reshape(target.size)
    out[np.where(tgt == ignore_index)[0]] = ignore_index
    inter = out[np.where(out == tgt)[0]]
    edges = np.arange(K + 1)
    a_i, _ = np.histogram(inter, bins=edges)
    a_o, _ = np.histogram(out, bins=edges)
    a_t, _ = np.histogram(tgt, bins=edges)
    return a_i, a_o + a_t - a_i, a_t


def intersection_and_union_torch(output, target, K, ignore_index=255):
    """util/util.py:52-63 — mutates `output` in place when it is contiguous, like the reference."""
    assert output.dim() in (1, 2, 3) and output.shape == target.shape
    out = output.reshape(-1)
    tgt = target.reshape(-1)
    out[tgt == ignore_index] = ignore_index
    inter = out[out == tgt]
    a_i = torch.histc(inter, bins=K, min=0, max=K - 1)
    a_o = torch.histc(out, bins=K, min=0, max=K - 1)
    a_t = torch.histc(tgt, bins=K, min=0, max=K - 1)
    return a_i, a_o + a_t - a_i, a_t


def _histc_ints(v, K):
    """torch.histc(bins=K, min=0, max=K-1) on integer data: ATen SummaryOps.cu getBin():
    bin = (v - min) * bins / (max - min) in integer arithmetic, bin == bins folded into the last bin, values
    outside [min, max] ignored.  K == 1 gives min == max == 0, which torch.histc documents as "use the data's
    min and max": every element lands in the single bin (confirmed against torch-CUDA on the B200)."""
    v = np.asarray(v).astype(np.int64).reshape(-1)
    if K == 1:
        return np.array([v.size], dtype=np.int64)
    lo, hi = 0, K - 1
    keep = (v >= lo) & (v <= hi)
    b = (v[keep] - lo) * K // (hi - lo)
    b[b == K] = K - 1
    return np.bincount(b, minlength=K).astype(np.int64)


def intersection_and_union_histc_ints(output, target, K, ignore_index=255):
    """util/util.py:52-63 with histc's integer binning spelled out (CPU-checkable)."""
    out = np.asarray(output).reshape(-1).copy()
    tgt = np.asarray(target).reshape(-1)
    out[tgt == ignore_index] = ignore_index
    inter = out[out == tgt]
    a_i, a_o, a_t = _histc_ints(inter, K), _histc_ints(out, K), _histc_ints(tgt, K)
    return a_i, a_o + a_t - a_i, a_t


def temporal_consistency_counts(labels, K, ignore_index=255, last_output=None, metric=intersection_and_union_np):
    """flow/base.py:280-295: labels [n,H,W]; returns summed (I,U,T) and the new last_output."""
    n = labels.shape[0]
    tot = [np.zeros(K, np.int64) for _ in range(3)]
    for p in range(n):
        if p > 0:
            cur, ref = labels[p], labels[p - 1]
        elif last_output is not None:
            cur, ref = labels[p], last_output
        else:
            continue
        for acc, v in zip(tot, metric(cur[None], ref[None], K, ignore_index)):
            acc += np.asarray(v.cpu() if hasattr(v, "cpu") else v).astype(np.int64)
    return tuple(tot), labels[n - 1]


def epoch_metrics(i_sum, u_sum, t_sum):
    """base/foundation.py:162-164 (fp64 numpy): mIoU, mAcc, allAcc and the per-class vectors."""
    i_sum, u_sum, t_sum = (np.asarray(x) for x in (i_sum, u_sum, t_sum))
    iou_class = i_sum / (u_sum + 1e-10)
    acc_class = i_sum / (t_sum + 1e-10)
    return {
        "miou": np.mean(iou_class),
        "macc": np.mean(acc_class),
        "accuracy": sum(i_sum) / (sum(t_sum) + 1e-10),
        "iou_class": iou_class,
        "accuracy_class": acc_class,
    }
