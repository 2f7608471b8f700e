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


def init_fuvs_comm():
    """Gives libfuvs its own NCCL communicator over the ranks of the initialised process group (one process per GPU):
    rank 0 draws the ncclUniqueId through the C ABI, torch.distributed carries its 128 bytes to the other ranks, every
    rank joins with fuvs_comm_init on its current device.  Returns True when `allreduce_counts` will go through
    fuvs_allreduce_counts, False for a single process or a CPU (gloo) group."""
    import ctypes

    from ._lib import check, load
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2 or not torch.cuda.is_available():
        return False
    lib = load()
    rank, world = dist.get_rank(), dist.get_world_size()
    if lib.fuvs_comm_world_size() == world:
        return True
    buf = (ctypes.c_ubyte * 128)()
    if rank == 0:
        check(lib.fuvs_comm_unique_id(buf))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    raw = (ctypes.c_ubyte * 128)(*t.cpu().tolist())
    check(lib.fuvs_comm_init(raw, rank, world))
    # NCCL connects lazily: the first collective of a fresh communicator sets the channels up (hundreds of ms)
    from ._lib import ptr, stream_ptr
    warm = torch.zeros(8, dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))
    check(lib.fuvs_allreduce_counts(ptr(warm), warm.numel(), stream_ptr(warm.device)))
    torch.cuda.synchronize()
    return True


def allreduce_counts(counts):
    """The path's only collective: sum of the int64 [3,K] (or [M,3,K]) count buffers over all ranks, in place —
    fuvs_allreduce_counts (NCCL behind the C ABI) once `init_fuvs_comm` has run, torch.distributed otherwise."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if counts.is_cuda and counts.dtype == torch.int64 and counts.is_contiguous():
            from ._lib import check, load, ptr, stream_ptr
            lib = load()
            if lib.fuvs_comm_world_size() == dist.get_world_size():
                with torch.cuda.device(counts.device):
                    check(lib.fuvs_allreduce_counts(ptr(counts), counts.numel(), stream_ptr(counts.device)))
                return counts
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts
