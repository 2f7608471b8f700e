"""CPU tier: the N>1 plumbing on world_size=2 with the gloo backend — clip sharding, the single all-reduce of the
(I,U,T) counts and the one-label halo exchange — against a single-process computation over all clips."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle as co


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _clip_labels(clip, n_intervals=3, n=5, H=24, W=32, K=5):
    g = torch.Generator().manual_seed(100 + clip)
    return torch.randint(0, K, (n_intervals, n, H, W), generator=g, dtype=torch.uint8).numpy()


def _clip_counts(clip, K=5):
    """Temporal-consistency counts of one clip (chain resets at the clip boundary, flow/base.py:247)."""
    lab = _clip_labels(clip)
    tot = np.zeros((3, K), np.int64)
    last = None
    for it in range(lab.shape[0]):
        co.temporal(lab[it], K, 255, tc_prev=last, out=tot)
        last = lab[it][-1]
    return tot


def _worker(rank, world, port, num_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from flood_uav_video_segmentation_b200 import dist as fdist
    r, _, w = fdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    mine = fdist.shard_clips(num_clips, rank, world)
    counts = torch.zeros((3, 5), dtype=torch.int64)
    for c in mine:
        counts += torch.from_numpy(_clip_counts(c))
    fdist.allreduce_counts(counts)
    # interval-range sharding of ONE video: the halo label travels rank -> rank+1
    lab = _clip_labels(0, n_intervals=6)
    a, b = fdist.shard_intervals(6, rank, world)
    halo = fdist.exchange_halo(torch.from_numpy(lab[b - 1][-1].copy()), rank, world)
    part = np.zeros((3, 5), np.int64)
    last = None if halo is None else halo.numpy()
    for it in range(a, b):
        co.temporal(lab[it], 5, 255, tc_prev=last, out=part)
        last = lab[it][-1]
    part = torch.from_numpy(part)
    fdist.allreduce_counts(part)
    q.put((rank, counts.numpy(), part.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_matches_single_process():
    world, num_clips = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = sum(_clip_counts(c) for c in range(num_clips))
    lab = _clip_labels(0, n_intervals=6)
    want_video = np.zeros((3, 5), np.int64)
    last = None
    for it in range(6):
        co.temporal(lab[it], 5, 255, tc_prev=last, out=want_video)
        last = lab[it][-1]
    for _, counts, part in res:
        assert np.array_equal(counts, want)
        assert np.array_equal(part, want_video)
