"""Multi-GPU plumbing for the interpolation path.

The path shards by clip (or by contiguous interval range of one video): intervals are independent given their two
key frames and grids, and the (I,U,T) counts are additive integers.  One process per GPU; no data-path
collective; exactly one NCCL all-reduce of 3K int64 counts per evaluation.  The reference never all-reduces
counts (it averages per-rank mIoU scalars, base/foundation.py:166-168); summing the counts reproduces the
*single-process* reference result exactly, which is the parity target.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialises torch.distributed from torchrun's RANK / WORLD_SIZE / MASTER_* (no-op for a single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def shard_clips(num_clips, rank, world):
    """Contiguous block of whole clips for this rank (any whole-clip assignment gives the same summed counts,
    because last_output resets at every clip boundary)."""
    base, rem = divmod(num_clips, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def shard_intervals(num_intervals, rank, world):
    """Contiguous interval range [start, stop) of ONE long video for this rank.  A rank with start > 0 needs the
    last label map of interval start-1 as its temporal-consistency halo (flow/base.py:285-287): recompute that
    one interval (halo = one key frame + its grids) or receive the 2 MB map from rank-1 (`exchange_halo`)."""
    base, rem = divmod(num_intervals, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def exchange_halo(last_label, rank, world, shape=None, device=None):
    """Sends this rank's final label map to rank+1 and returns the one received from rank-1 (None on rank 0).
    One point-to-point message of H*W bytes per rank boundary."""
    if world == 1 or not dist.is_initialized():
        return None
    recv = None
    ops = []
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, last_label.contiguous(), rank + 1))
    if rank > 0:
        recv = torch.empty(shape if shape is not None else last_label.shape, dtype=torch.uint8,
                           device=device if device is not None else last_label.device)
        ops.append(dist.P2POp(dist.irecv, recv, rank - 1))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return recv


def allreduce_counts(counts):
    """The path's only collective: sum of the int64 [3,K] (or [M,3,K]) count buffers over all ranks, in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts
