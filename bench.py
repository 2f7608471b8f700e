#!/usr/bin/env python
"""bench.py — interpolated frames/s at 1080p (k=5) on N B200s + HBM roofline of the fused interval kernels.

Contract (task prompt §④ + base contract):
  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode dense|block|linear|...]
  N>1 is launched by torchrun, one rank per GPU; rank 0 prints ONE JSON line.

Workload (BASELINE.json configs[1]): flow-warped logit interpolation (no_warp=False) on synthetic 1080x1920
clips, C=5, k=5, dense [H,W,2] flow grids.  A step = `--clips-per-step` 16-frame clips = 3 intervals each = 12
interpolated frames per clip.  `value`: inputs resident in HBM.  `e2e`: the same work through
FlowBaseModel.predict_step with HOST (pinned) inputs, H2D and D2H inside the timed region, beside the rate of plain
pinned copies of the same bytes (`e2e.pcie_peak_gbs`: the arm's own roofline).
`--impl reference`: the reference's own modules (oracle/_ref, copied unmodified by oracle/make_ref.py) on torch-CPU,
a bounded sample of the workload per step; both arms print the SAME `config`.
`other_modes`: the other routes at 1080p, BASELINE configs[0] (433x433, beside the CPU reference) and configs[2]
(feature-based interval at DeepLabV3 / PSPNet feature sizes).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, H, W, K_DELTA, CLIP_FRAMES = 5, 1080, 1920, 5, 16
METRIC = "interpolated_frames_per_sec_1080p_k5"
S_BYTES = C * H * W * 4
LB_BYTES = H * W


def set_shape(h, w):
    """Frame size of the synthetic clips (1080x1920: BASELINE configs[1]; 433x433: configs[0], data.train_w=433)."""
    global H, W, S_BYTES, LB_BYTES
    H, W = int(h), int(w)
    S_BYTES = C * H * W * 4
    LB_BYTES = H * W


def algorithmic_bytes(mode, k=K_DELTA):
    """algorithmic HBM bytes per interval (uint8 labels)."""
    if mode == "linear":
        return 2 * S_BYTES + k * LB_BYTES + LB_BYTES
    if mode == "linear_lowres":    # key frames read at decoder resolution (stride 8), SURVEY.md §8f rank 1
        return 2 * C * (H // 8) * (W // 8) * 4 + k * LB_BYTES + LB_BYTES
    if mode in ("block", "block_clip"):
        hg, wg = H // 16, W // 16
        return S_BYTES + 8 * C * hg * wg * 4 + 2 * (k - 1) * hg * wg * 8 + (k + 1) * LB_BYTES
    if mode == "block_lowres":     # key frames at decoder resolution: frame 0 and the first chain step up-sample on the fly
        hg, wg = H // 16, W // 16
        return 2 * C * (H // 8) * (W // 8) * 4 + 8 * C * hg * wg * 4 + 2 * (k - 1) * hg * wg * 8 + (k + 1) * LB_BYTES
    g = H * W * 8
    dense = (5 * k - (8 if k % 2 == 0 else 7)) * S_BYTES + 2 * (k - 1) * g + (k + 1) * LB_BYTES
    if mode == "dense_lowres":     # + the decoder-resolution key frames in and ONE up-sampled key frame written per interval
        return dense + S_BYTES + 2 * C * (H // 8) * (W // 8) * 4      # (the other one is the previous interval's `next`)
    return dense


def feature_bytes(cf, fh, fw, hg, wg, k=K_DELTA):
    """fuvs_feature_interval: two key-frame feature maps in, k frames out, chain states (grid resolution) and the
    default-grid pass of the key frame written and read once."""
    return (2 + k) * cf * fh * fw * 4 + 2 * (k - 1) * 2 * cf * hg * wg * 4 + 2 * cf * hg * wg * 4


def measured_traffic(mode):
    """dram__bytes_read.sum + dram__bytes_write.sum per interval from the committed ncu launch list (profiles/)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(mode)
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def bind_rank_to_local_cores(local, world):
    """N>1: every rank stages its own pinned buffers; keep its threads (and, by first touch, its pinned pages) on a
    private slice of the cores NVML reports as local to its GPU instead of letting N ranks share all of them."""
    if world <= 1 or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cores = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1]
        cores = [c for c in cores if c in os.sched_getaffinity(0)] or sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // world)
        mine = cores[(local * per) % len(cores):(local * per) % len(cores) + per] or cores
        os.sched_setaffinity(0, set(mine))
        return {"cores": [mine[0], mine[-1]], "n": len(mine)}
    except Exception as exc:  # noqa: BLE001
        return {"error": repr(exc)}


# ----------------------------------------------------------------------------- synthetic clips
SMOOTH_CELL = int(os.environ.get("FUVS_BENCH_SMOOTH_CELL", "120"))


def make_grids(n_grids, mode, device, gen):
    """identity + uniform(+-0.025) jitter per grid point.  mode "dense_smooth": the same jitter drawn
    on a coarse lattice (one node per SMOOTH_CELL = 120 pixels) and bilinearly up-sampled to pixel resolution — a
    spatially coherent field like optical flow or up-sampled H.264 motion vectors (same +-24 px amplitude, gradient
    below 0.4 px/px) instead of iid noise per pixel."""
    from flood_uav_video_segmentation_b200.synthetic import identity_grid
    if mode == "dense_smooth":
        base = identity_grid(H, W, "dense").to(device)
        low = (torch.rand((n_grids, 2, H // SMOOTH_CELL + 1, W // SMOOTH_CELL + 1), device=device, generator=gen) - 0.5) * 0.05
        jit = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
        return (base.unsqueeze(0) + jit).contiguous()
    base = identity_grid(H, W, mode).to(device)
    jit = (torch.rand((n_grids,) + tuple(base.shape), device=device, generator=gen) - 0.5) * 0.05
    return (base.unsqueeze(0) + jit).contiguous()


def make_clip(mode, device, seed):
    """One 16-frame clip: 4 key-frame logit maps [1,C,H,W] and, per interval, stacked grids [k-1,Hg,Wg,2] x2."""
    gen = torch.Generator(device=device).manual_seed(seed)
    lowres = mode in ("linear_lowres", "block_lowres", "dense_lowres")
    if mode in ("block_clip", "block_lowres"):       # same grids as "block": only the entry point / key-frame size differ
        mode = "block"
    if mode == "dense_lowres":
        mode = "dense"
    n_int = (CLIP_FRAMES - 1) // K_DELTA
    kh, kw = (H // 8, W // 8) if lowres else (H, W)
    keys = [torch.randn((1, C, kh, kw), device=device, generator=gen) for _ in range(n_int + 1)]
    grids = []
    for _ in range(n_int):
        if mode in ("linear", "linear_lowres"):
            grids.append((None, None))
        else:
            grids.append((make_grids(K_DELTA - 1, mode, device, gen), make_grids(K_DELTA - 1, mode, device, gen)))
    return keys, grids


def clip_bytes(mode):
    n_int = (CLIP_FRAMES - 1) // K_DELTA
    b = (n_int + 1) * S_BYTES
    if mode in ("dense", "dense_smooth", "dense_lowres"):
        b += n_int * 2 * (K_DELTA - 1) * H * W * 8
    elif mode in ("block", "block_clip"):
        b += n_int * 2 * (K_DELTA - 1) * (H // 16) * (W // 16) * 8
    return b


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------- device-resident arm
_UPS = {}


def run_interval(kernels, mode, keys, grids, it, tc_prev, counts, scratch=None):
    if mode == "linear":
        labels, _ = kernels.linear_blend_argmax(keys[it], keys[it + 1], K_DELTA, tc_prev=tc_prev, counts=counts)
    elif mode == "linear_lowres":
        labels, _ = kernels.linear_lowres_blend_argmax(keys[it], keys[it + 1], (H, W), K_DELTA, tc_prev=tc_prev, counts=counts)
    elif mode in ("dense", "dense_smooth"):
        labels, _ = kernels.dense_interval(keys[it], keys[it + 1], grids[it][0], grids[it][1], K_DELTA, tc_prev=tc_prev,
                                           counts=counts, scratch=scratch)
    elif mode == "dense_lowres":
        # one pair of up-sample buffers per stream (= per scratch buffer): an interval's `next` is the following one's `prev`
        ups = _UPS.setdefault(scratch.data_ptr() if scratch is not None else 0, kernels.KeyFrameUps())
        labels, _ = kernels.dense_lowres_interval(keys[it], keys[it + 1], (H, W), grids[it][0], grids[it][1], K_DELTA,
                                                  tc_prev=tc_prev, counts=counts, scratch=scratch, ups=ups)
    elif mode == "block_lowres":
        labels, _ = kernels.block_lowres_interval(keys[it], keys[it + 1], (H, W), grids[it][0], grids[it][1], K_DELTA,
                                                  tc_prev=tc_prev, counts=counts, scratch=scratch)
    else:
        labels, _ = kernels.block_interval(keys[it], keys[it + 1], grids[it][0], grids[it][1], K_DELTA, tc_prev=tc_prev,
                                           counts=counts, scratch=scratch)
    return labels


def run_clip(kernels, mode, clip, counts, scratch=None):
    keys, grids = clip
    if mode == "block_clip":       # the clip-level entry: chain steps batched over the clip's intervals
        labels, _ = kernels.block_clip(keys, [g[0] for g in grids], [g[1] for g in grids], K_DELTA, counts=counts,
                                       scratch=scratch)
        return labels[-1, K_DELTA - 1]
    last = None
    for it in range(len(keys) - 1):
        labels = run_interval(kernels, mode, keys, grids, it, last, counts, scratch)
        last = labels[K_DELTA - 1]
    return last


def scratch_floats(kernels, mode):
    lib = kernels.load()
    if mode in ("dense", "dense_smooth", "dense_lowres"):
        return int(lib.fuvs_dense_scratch_floats(C, H, W, K_DELTA))
    if mode in ("block", "block_clip", "block_lowres"):
        return 3 * int(lib.fuvs_block_scratch_floats(C, H // 16, W // 16, K_DELTA))
    return 1


def time_resident(kernels, dist_mod, mode, clips, steps, warmup, clips_per_step, device, world, sampler=None, streams=1):
    """K steps of `clips_per_step` clips with inputs resident in HBM.  Clips are independent (the path shards by clip),
    so with streams > 1 clip c runs on stream c % streams: the CTAs of one clip's kernels fill the tail of the other's
    (every interval kernel is a persistent one-CTA-per-SM grid).  Counts of all clips go to one buffer (integer
    atomics: the sum does not depend on the interleaving)."""
    from flood_uav_video_segmentation_b200 import launch_count
    counts = kernels.new_counts(C, device)
    nst = max(1, int(streams))
    scr = [torch.empty((max(scratch_floats(kernels, mode), 1),), dtype=torch.float32, device=device) for _ in range(nst)]
    side = [torch.cuda.Stream(device) for _ in range(nst)] if nst > 1 else []
    ci = 0

    def eager_step():
        nonlocal ci
        if nst == 1:
            for _ in range(clips_per_step):
                run_clip(kernels, mode, clips[ci % len(clips)], counts, scr[0])
                ci += 1
            return
        cur = torch.cuda.current_stream(device)
        fork = torch.cuda.Event()
        fork.record(cur)
        for sidx in range(nst):
            side[sidx].wait_event(fork)
        for j in range(clips_per_step):
            with torch.cuda.stream(side[j % nst]):
                run_clip(kernels, mode, clips[ci % len(clips)], counts, scr[j % nst])
            ci += 1
        for sidx in range(nst):
            done = torch.cuda.Event()
            done.record(side[sidx])
            cur.wait_event(done)

    for _ in range(warmup):
        eager_step()
    # A step is a fixed sequence of C-ABI calls on fixed buffers: capture it once per phase of the clip cycle in a CUDA
    # graph and replay it, so the timed region measures the GPU and not the Python/ctypes call overhead (~20 us per
    # interval, the same order as a linear-mode interval).  FUVS_BENCH_GRAPH=0 times eager launches instead.
    time_resident.launch = "eager"
    step = eager_step
    if os.environ.get("FUVS_BENCH_GRAPH", "1") != "0":
        try:
            nphase = (len(clips) * clips_per_step // math.gcd(len(clips), clips_per_step)) // clips_per_step
            torch.cuda.synchronize(device)
            graphs, per_graph = [], []
            ci = 0
            for _ in range(nphase):
                g = torch.cuda.CUDAGraph()
                l_before = launch_count()
                with torch.cuda.graph(g):
                    eager_step()
                per_graph.append(launch_count() - l_before)
                graphs.append(g)
            phase = 0

            def graph_step():
                nonlocal phase
                graphs[phase % nphase].replay()
                time_resident.replayed += per_graph[phase % nphase]
                phase += 1

            step = graph_step
            time_resident.launch = "cuda_graph"
            for _ in range(2):
                step()
        except Exception as exc:          # capture refused: fall back to eager launches and say so
            print(f"[bench] CUDA graph capture failed ({exc!r}); timing eager launches", file=sys.stderr)
            torch.cuda.synchronize(device)
            step = eager_step
            time_resident.launch = "eager"
    counts.zero_()
    torch.cuda.synchronize(device)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(device)
    l0 = launch_count()
    time_resident.replayed = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx = sampler if sampler is not None else _Null()
    with ctx:
        e0.record()
        for _ in range(steps):
            step()
        dist_mod.allreduce_counts(counts)       # the path's only collective
        e1.record()
        torch.cuda.synchronize(device)
    if world > 1:
        torch.distributed.barrier()
    ms = e0.elapsed_time(e1)
    launches = launch_count() - l0 + time_resident.replayed        # kernels replayed from the graphs count like launches
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return ms, launches, counts


time_resident.launch = "eager"
time_resident.replayed = 0


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def time_feature(kernels, cf, device, reps=6):
    """fuvs_feature_interval (model.feature_based=True, flow/model.py:116-181) at encoder-feature size [cf,135,240]
    (DeepLabV3-R101: 2048 channels, PSPNet: 4096) with the 67x120 block grids of a 1080p clip -> us per interval."""
    from flood_uav_video_segmentation_b200.synthetic import flow_grids, identity_grid
    fh, fw, hg, wg = H // 8, W // 8, H // 16, W // 16
    gen = torch.Generator(device=device).manual_seed(99 + cf)
    feats = [torch.randn((cf, fh, fw), device=device, generator=gen) for _ in range(3)]
    gl = [[g.to(device) for g in flow_grids(H, W, K_DELTA, "block", clip=20 + i, side=0)] for i in range(2)]
    gr = [[g.to(device) for g in flow_grids(H, W, K_DELTA, "block", clip=20 + i, side=1)] for i in range(2)]
    dgrid = identity_grid(H, W, "block").unsqueeze(0).to(device)
    out = torch.empty((K_DELTA, cf, fh, fw), device=device)
    scratch = kernels.ScratchCache()

    def call(i):
        kernels.feature_interval(feats[i % 2], feats[i % 2 + 1], gl[i % 2], gr[i % 2], K_DELTA, default_grid=dgrid,
                                 scratch=scratch, out=out)

    for i in range(2):
        call(i)
    torch.cuda.synchronize(device)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(2):
            call(i)
    g.replay()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize(device)
    del feats, out, scratch
    torch.cuda.empty_cache()
    return e0.elapsed_time(e1) * 1e3 / (2 * reps), feature_bytes(cf, fh, fw, hg, wg)


# ----------------------------------------------------------------------------- end-to-end arm (host buffers)
class Stride8Net(torch.nn.Module):
    """Stand-in key-frame network of the end-to-end arm: stride-8 "encoder" (8x8 average pooling of the RGB frame) and a
    random-init 1x1 "decoder" to C classes — the shapes the reference's networks produce (decoder output at 1/8 of the
    frame, flow/model.py:188-193) at a cost that leaves the arm measuring the interpolation path."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(1234)
        self.encoder = torch.nn.AvgPool2d(8)
        self.decoder = torch.nn.Conv2d(3, C, 1)


def time_e2e(mode, steps, warmup, clips_per_step, device, world, host_clips, inputs="logits"):
    """FlowBaseModel.predict_step with pinned HOST inputs.  Per interval the copy stream uploads the NEXT key frame
    and the 2(k-1) grids (the previous interval's `next` becomes this interval's `prev` on the device: each key frame
    crosses PCIe once per clip), double-buffered against compute; the uint8 label maps go back D2H per interval and
    the counts are read once at the end.  inputs = "logits": the key frames are full-resolution logit maps [1,C,H,W]
    and the network is the identity (r01's arm); "image": they are RGB frames [1,3,H,W] like the reference's batches
    (flow/dataset.py), a stride-8 stand-in network produces decoder-resolution logits and the interval entry takes them
    from there."""
    from flood_uav_video_segmentation_b200.flow.base import FlowBaseModel

    class Identity(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = torch.nn.Identity()
            self.decoder = torch.nn.Identity()

    image = inputs == "image"
    KC = 3 if image else C                       # channels of an uploaded key frame
    key_bytes = KC * H * W * 4
    backbone = (Stride8Net().to(device).eval() if image else Identity())
    model = FlowBaseModel(classes=C, arch="pspnet", feature_based=False, no_warp=(mode == "linear"), no_cropping=True,
                          backbone=backbone, output_size=(H, W), save_video=False).eval()
    copy_stream = torch.cuda.Stream(device)
    n_int = (CLIP_FRAMES - 1) // K_DELTA
    # key-frame ring of 4 device buffers handed out round-robin: while interval j computes on two of them the copy
    # stream fills the one or two (clip boundary) that interval j+1 needs; a buffer is rewritten 4 uploads later, when
    # the last interval that read it has long been enqueued (its key_free event is recorded before the wait is issued)
    NK = 4
    keys_dev = [torch.empty((1, KC, H, W), device=device) for _ in range(NK)]
    key_ready = [torch.cuda.Event() for _ in range(NK)]
    key_free = [torch.cuda.Event() for _ in range(NK)]
    gslots = []
    for _ in range(2):
        s = {"ready": torch.cuda.Event(), "free": torch.cuda.Event()}
        if mode != "linear":
            gshape = host_clips[0][1][0][0].shape
            s["gl"], s["gr"] = torch.empty(gshape, device=device), torch.empty(gshape, device=device)
        gslots.append(s)
    out_host = [torch.empty((K_DELTA, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
    gbytes = 0 if mode == "linear" else 2 * host_clips[0][1][0][0].numel() * 4
    d2h = K_DELTA * H * W
    state = {"ci": 0, "alloc": 0, "last_next": 0, "h2d": 0}

    def upload_key(host_key):
        k = state["alloc"] % NK
        state["alloc"] += 1
        copy_stream.wait_event(key_free[k])
        keys_dev[k].copy_(host_key, non_blocking=True)
        key_ready[k].record(copy_stream)
        state["h2d"] += key_bytes
        return k

    def stage(j, clip, it):
        """uploads what interval `it` of `clip` (the j-th interval of the step) needs and does not have yet"""
        keys, grids = clip
        slot = gslots[j % 2]
        with torch.cuda.stream(copy_stream):
            # first interval of a clip: both key frames; later ones: prev is the previous interval's next
            k0 = upload_key(keys[0]) if it == 0 else state["last_next"]
            k1 = upload_key(keys[it + 1])
            if mode != "linear":
                copy_stream.wait_event(slot["free"])
                slot["gl"].copy_(grids[it][0], non_blocking=True)
                slot["gr"].copy_(grids[it][1], non_blocking=True)
                slot["ready"].record(copy_stream)
                state["h2d"] += gbytes
        state["last_next"] = k1
        return (k0, k1)

    def step():
        items = []
        for _ in range(clips_per_step):
            clip = host_clips[state["ci"] % len(host_clips)]
            state["ci"] += 1
            items += [(clip, it) for it in range(n_int)]
        cur = torch.cuda.current_stream(device)
        pairs = {0: stage(0, *items[0])}
        for j, (clip, it) in enumerate(items):
            if j + 1 < len(items):
                pairs[j + 1] = stage(j + 1, *items[j + 1])
            kp, kn = pairs.pop(j)
            slot = gslots[j % 2]
            if it == 0:
                model.last_output = None          # temporal chain resets at clip boundaries
            cur.wait_event(key_ready[kp])
            cur.wait_event(key_ready[kn])
            if mode == "linear":
                dummy = [None] * (K_DELTA - 1)
                batch = {"frame_prev": keys_dev[kp], "frame_next": keys_dev[kn], "mvs_left": dummy, "mvs_right": dummy}
            else:
                cur.wait_event(slot["ready"])
                batch = {"frame_prev": keys_dev[kp], "frame_next": keys_dev[kn],
                         "mvs_left": _GridList(slot["gl"]), "mvs_right": _GridList(slot["gr"])}
            with torch.no_grad():
                labels = model.predict_step(batch, j)
            slot["free"].record(cur)
            key_free[kp].record(cur)              # prev is not needed after this interval (next stays for the following one)
            if it == n_int - 1:
                key_free[kn].record(cur)
            out_host[j % 2].copy_(labels, non_blocking=True)          # flow/base.py:277 (already uint8)
        return len(items)

    model.on_predict_start()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize(device)
    if world > 1:
        torch.distributed.barrier()
    model.on_predict_start()
    state["h2d"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    n_items = 0
    for _ in range(steps):
        n_items += step()
    res = model.on_predict_end()              # all-reduce + D2H of the counts, fp64 formulas
    e1.record()
    torch.cuda.synchronize(device)
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms = max(e0.elapsed_time(e1), wall_ms)    # host-side staging is part of the end-to-end cost
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    per_step_items = n_items // max(steps, 1)
    return ms, state["h2d"] // max(steps, 1), d2h * per_step_items, res


def _GridList(stacked):
    """The reference's python list of k-1 grids [1,Hg,Wg,2] (flow/dataset.py:138-146): views of the staging buffer, handed
    to the kernels as a pointer table."""
    return [stacked[j:j + 1] for j in range(stacked.shape[0])]


def time_h2d_peak(mode, steps, clips_per_step, device, world, host_clips, inputs="logits"):
    """The e2e arm's own roofline: the SAME host buffers copied to the device with one plain pinned cudaMemcpyAsync
    per buffer (Tensor.copy_(non_blocking=True) of a pinned tensor), same order, no kernels, at the same N (all ranks
    copy at once: they share the host's memory system and PCIe root) -> (GB/s per GPU, bytes per step)."""
    n_int = (CLIP_FRAMES - 1) // K_DELTA
    KC = 3 if inputs == "image" else C
    key_bytes = KC * H * W * 4
    key_dev = [torch.empty((1, KC, H, W), device=device) for _ in range(2)]
    gdev = None
    if mode != "linear":
        gshape = host_clips[0][1][0][0].shape
        gdev = [torch.empty(gshape, device=device) for _ in range(2)]
    stream = torch.cuda.Stream(device)

    def step():
        nbytes = 0
        for c in range(clips_per_step):
            keys, grids = host_clips[c % len(host_clips)]
            for it in range(n_int):
                if it == 0:
                    key_dev[0].copy_(keys[0], non_blocking=True)
                    nbytes += key_bytes
                key_dev[(it + 1) % 2].copy_(keys[it + 1], non_blocking=True)
                nbytes += key_bytes
                if gdev is not None:
                    gdev[0].copy_(grids[it][0], non_blocking=True)
                    gdev[1].copy_(grids[it][1], non_blocking=True)
                    nbytes += 2 * grids[it][0].numel() * 4
        return nbytes

    with torch.cuda.stream(stream):
        step()
        torch.cuda.synchronize(device)
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nbytes = 0
        for _ in range(steps):
            nbytes = step()
        e1.record()
        torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return nbytes * steps / (ms / 1e3) / 1e9, nbytes


def to_host_clip(clip, pin=True):
    keys, grids = clip
    f = (lambda t: t.cpu().pin_memory()) if pin else (lambda t: t.cpu())
    hk = [f(k) for k in keys]
    hg = [(None, None) if g[0] is None else (f(g[0]), f(g[1])) for g in grids]
    return hk, hg


# ----------------------------------------------------------------------------- stock torch ops on the same GPU
def stock_torch_interval(mode, keys, grids, it, last, want_counts=False):
    """The op sequence the reference issues for one interval, as stock ATen CUDA kernels on this GPU: grid_sample
    chains (flow/model.py:212-229, 244-249), scalar mul / add (233-237), cat (239), max(1)[1] (flow/base.py:276) and
    the temporal-consistency metric through intersectionAndUnionGPU with its three .cpu() reads per call
    (flow/base.py:280-295, util/util.py:52-63, base/foundation.py:344).  Written out here, not imported from oracle/."""
    import torch.nn.functional as F
    n = K_DELTA
    o, o_next = keys[it], keys[it + 1]
    h, w = o.shape[2], o.shape[3]

    def chain(x, gs):
        outs = []
        for g in gs:
            x = F.grid_sample(x, g, mode="bilinear", padding_mode="border", align_corners=False)
            outs.append(x if (x.shape[2], x.shape[3]) == (h, w) else
                        F.interpolate(x, size=(h, w), mode="bilinear", align_corners=True))
        return outs

    if mode == "linear":
        fwd, bwd = [o] * (n - 1), [o_next] * (n - 1)
    else:
        fwd = chain(o, [grids[it][0][j:j + 1] for j in range(n - 1)])
        bwd = chain(o_next, [grids[it][1][j:j + 1] for j in range(n - 1)])
    maps = [o]
    for p in range(1, n):
        maps.append((n - p) / n * fwd[p - 1] + p / n * bwd[n - p - 1])
    labels = torch.cat(maps, 0).max(1)[1]

    def metric(output, target):
        output = output.reshape(-1).clone()
        target = target.reshape(-1)
        output[target == 255] = 255
        inter = output[output == target]
        a_i = torch.histc(inter, bins=C, min=0, max=C - 1)
        a_o = torch.histc(output, bins=C, min=0, max=C - 1)
        a_t = torch.histc(target, bins=C, min=0, max=C - 1)
        return a_i.cpu().numpy(), (a_o + a_t - a_i).cpu().numpy(), a_t.cpu().numpy()

    tot = np.zeros((3, C), np.int64)
    for p in range(n):
        if p == 0 and last is None:
            continue
        i, u, t = metric(labels[p], labels[p - 1] if p > 0 else last)
        tot += np.stack([i, u, t]).astype(np.int64)
    return (labels, tot) if want_counts else labels


def time_stock_torch(mode, clip, device, intervals=6, rounds=3):
    """Best of `rounds` wall-clock measurements over `intervals` intervals each (the sequence syncs with the host in
    every metric call, so wall clock is what a caller sees)."""
    keys, grids = clip
    n_int = len(keys) - 1
    best, labels = None, None
    with torch.no_grad():
        for w in range(2):                                      # warm-up: allocator cache, first-launch costs
            stock_torch_interval(mode, keys, grids, w % n_int, None)
        for _ in range(rounds):
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            last = None
            for j in range(intervals):
                labels = stock_torch_interval(mode, keys, grids, j % n_int, last if j % n_int else None)
                last = labels[K_DELTA - 1]
            torch.cuda.synchronize(device)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return best, labels


# ----------------------------------------------------------------------------- CPU baseline (reference / oracle port)
class _NullProfiler:
    """Lightning's profiler as the reference uses it: profiler.profile(name) context managers (flow/model.py:119-232)."""

    def profile(self, name):
        import contextlib
        return contextlib.nullcontext()


def load_reference():
    """The reference's own modules, copied unmodified to oracle/_ref by oracle/make_ref.py (None when absent)."""
    if load_reference.cache is False:
        try:
            from oracle import make_ref
            load_reference.cache = make_ref.load()
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] oracle/_ref not usable ({exc!r}); CPU baseline falls back to the oracle port", file=sys.stderr)
            load_reference.cache = None
    return load_reference.cache


load_reference.cache = False


def cpu_kind():
    return "reference" if load_reference() is not None else "port"


def cpu_interval(mode, keys, grids, it, last):
    """One interval on torch-CPU: FlowModel.predict -> max(1)[1] -> uint8 -> temporal-consistency counts.

    With oracle/_ref present this executes the reference's own FlowModel.predict (flow/model.py:109-249) and
    intersectionAndUnion (util/util.py:36-47); the ~15 lines of flow/base.py:276-295 around them (arg-max, uint8 cast,
    the temporal loop through compute_metrics' CPU branch, base/foundation.py:333-339) are restated here because
    flow/base.py needs pytorch_lightning.  Without oracle/_ref the oracle port does the same (kind "port")."""
    n = K_DELTA
    if mode == "linear":
        gl = gr = [torch.zeros(1, 1)] * (n - 1)
    else:
        gl = [grids[it][0][j:j + 1] for j in range(n - 1)]
        gr = [grids[it][1][j:j + 1] for j in range(n - 1)]
    ref = load_reference()
    if ref is None:
        from oracle import flow_oracle as fo
        from oracle import metric_oracle as mo
        ident = torch.nn.Identity()
        with torch.no_grad():
            logits = fo.predict_segmentation(ident, ident, keys[it], keys[it + 1], gl, gr, n, no_warp=(mode == "linear"))
            labels = fo.argmax_labels(logits)
        lab_np = labels.numpy().astype("uint8")
        counts, new_last = mo.temporal_consistency_counts(labels.numpy(), keys[it].shape[1], 255, last)
        return lab_np, counts, new_last
    from types import SimpleNamespace
    model = cpu_interval.models.get(mode)
    if model is None:
        net = SimpleNamespace(encoder=torch.nn.Identity(), decoder=torch.nn.Identity())
        model = cpu_interval.models[mode] = ref["FlowModel"](net, feature_based=False, no_warp=(mode == "linear")).eval()
    classes = keys[it].shape[1]
    with torch.no_grad():
        output = model.predict(keys[it], keys[it + 1], gl, gr, n, _NullProfiler())["pred"]     # flow/base.py:271
        output = output.data.max(1)[1]                                                         # :276
        output_numpy = output.data.cpu().numpy().astype("uint8")                               # :277
    tot = [np.zeros(classes, np.int64) for _ in range(3)]
    for p in range(n):                                                                         # :280-295
        if p > 0:
            cur, nxt = output[p], output[p - 1]
        elif last is not None:
            cur, nxt = output[p], last
        else:
            continue
        i, u, t = ref["intersectionAndUnion"](cur.unsqueeze(0).detach().numpy(), nxt.unsqueeze(0).detach().numpy(),
                                              classes, 255)                                    # foundation.py:334-339
        for acc, v in zip(tot, (i, u, t)):
            acc += v
    return output_numpy, tuple(tot), output[n - 1]


cpu_interval.models = {}


def time_cpu(mode, host_clip, intervals, warm=1):
    keys, grids = host_clip
    torch.set_num_threads(os.cpu_count() or 1)
    n_int = len(keys) - 1
    for w in range(warm):
        cpu_interval(mode, keys, grids, 0, None)
    t0 = time.perf_counter()
    last = None
    lab = None
    for j in range(intervals):
        lab, _, last = cpu_interval(mode, keys, grids, j % n_int, last if j % n_int else None)
    dt = time.perf_counter() - t0
    return dt, lab


# ----------------------------------------------------------------------------- on-hardware parity of the collective
def allreduce_parity(kernels, dist_mod, mode, rank, world, device):
    """every rank runs ONE small clip (its own seed) through the interval kernels, the (I,U,T) counts
    are all-reduced over NCCL, and rank 0 recomputes all `world` clips single-process with the stock torch-CUDA op
    sequence (stock_torch_interval, written out in this file) -> the sums must be identical."""
    pmode = "dense" if mode.startswith("dense") else ("block" if mode.startswith("block") else "linear")
    full = (H, W)
    set_shape(272, 480)
    try:
        mine = make_clip(pmode, device, 777000 + rank)
        counts = kernels.new_counts(C, device)
        run_clip(kernels, pmode, mine, counts)
        dist_mod.allreduce_counts(counts)
        ok, detail = True, None
        if rank == 0:
            ref = np.zeros((3, C), np.int64)
            for r in range(world):
                keys, grids = make_clip(pmode, device, 777000 + r)
                last = None
                for it in range(len(keys) - 1):
                    labels, c = stock_torch_interval(pmode, keys, grids, it, last, want_counts=True)
                    ref += c
                    last = labels[K_DELTA - 1]
            got = counts.cpu().numpy()
            ok = bool(np.array_equal(got, ref))
            detail = {"mode": pmode, "shape": [C, H, W], "clips": world, "intersections_allreduced": int(got[0].sum()),
                      "intersections_single_process": int(ref[0].sum())}
        return ok, detail
    finally:
        set_shape(*full)


# ----------------------------------------------------------------------------- main
def workload_config(mode, args, world):
    """`config` of the JSON line: the workload the metric is quoted on.  Both arms (--impl ours / reference) print exactly
    this dictionary; what a step of the CPU arm actually times is said in its cpu_baseline.sample."""
    workload = (f"flow-warped logit interpolation (no_warp=False), {mode} flow grids, C={C}, {H}x{W}, k={K_DELTA}, "
                f"{CLIP_FRAMES}-frame clips" if not mode.startswith("linear") else
                f"linear logit interpolation (no_warp=True), C={C}, {H}x{W}, k={K_DELTA}, {CLIP_FRAMES}-frame clips")
    resident_mb = args.distinct_clips * clip_bytes(mode) / 1e6
    return {"workload": workload, "classes": C, "height": H, "width": W, "frame_delta": K_DELTA, "mode": mode,
            "clips_per_step": args.clips_per_step, "intervals_per_step": args.clips_per_step * 3,
            "interpolated_frames_per_step": args.clips_per_step * 3 * (K_DELTA - 1),
            "parallelism": f"clip-sharded x{world}", "streams_per_gpu": args.streams,
            "l2_policy": f"inputs larger than L2: {resident_mb:.0f} MB of distinct clips cycled (L2 = 126 MB)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="dense", choices=["dense", "block", "block_clip", "block_lowres", "linear", "dense_smooth", "linear_lowres", "dense_lowres"])
    ap.add_argument("--clips-per-step", type=int, default=40, help="40 clips = 120 intervals: a dense step is ~27 ms")
    ap.add_argument("--distinct-clips", type=int, default=4)
    ap.add_argument("--streams", type=int, default=2, help="CUDA streams the independent clips of a step alternate over")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-modes", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    # rank 0 prints exactly one line on stdout.  NCCL_DEBUG is left as the launcher set it; its log (version banner,
    # INFO lines) is routed to stderr unless the launcher chose a file itself.
    if "NCCL_DEBUG" in os.environ and "NCCL_DEBUG_FILE" not in os.environ:
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    mode = args.mode
    config = workload_config(mode, args, world)

    if args.impl == "reference":
        return main_reference(args, rank, world, mode, config)

    affinity = bind_rank_to_local_cores(int(os.environ.get("LOCAL_RANK", "0")), world)
    from flood_uav_video_segmentation_b200 import dist as fdist
    from flood_uav_video_segmentation_b200 import kernels
    rank, local, world = fdist.init_from_env("nccl")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    kernels.load()
    fuvs_comm = fdist.init_fuvs_comm()       # the counts' all-reduce goes through the C ABI (its own NCCL communicator)

    clips = [make_clip(mode, device, 1000 * rank + i) for i in range(args.distinct_clips)]
    sampler = ClockSampler(local)
    ms, launches, counts = time_resident(kernels, fdist, mode, clips, args.steps, args.warmup, args.clips_per_step,
                                         device, world, sampler, streams=args.streams)
    launch_mode = ("one CUDA graph replay per step (captured C-ABI calls)" if time_resident.launch == "cuda_graph"
                   else "eager C-ABI calls from Python")
    intervals = args.steps * args.clips_per_step * 3
    frames = intervals * (K_DELTA - 1)
    value = frames * world / (ms / 1e3)
    peak, peak_src = measured_peak()
    bytes_iv = algorithmic_bytes(mode)
    achieved = bytes_iv * intervals / (ms / 1e3) / 1e9
    out = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "impl": "ours",
           "launch": launch_mode, "clocks": sampler.summary(), "gpu_launches": int(launches),
           "output_frames_per_sec": intervals * K_DELTA * world / (ms / 1e3),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": measured_traffic(mode), "peak_source": peak_src, "kernel": f"fuvs_{mode}_interval",
                        "algorithmic_bytes_per_launch": bytes_iv, "us_per_launch": ms * 1e3 / intervals,
                        "launch": "one interval = one C-ABI call (dense: 4 dense_strip_kernel steps + 1 temporal_counts_v16_kernel; "
                                  "block: chain steps + frame kernel + temporal_counts_v16_kernel (fused into the frame kernel on the low-res route); linear: one kernel); duration = CUDA events over the "
                                  "timed region / intervals",
                        "frac_of_nominal_8000": achieved / 8000.0},
           # content-dependent: the all-reduced temporal-consistency intersections (sum over classes)
           "temporal_intersections": int(counts[0].sum().item()),
           "temporal_targets": int(counts[2].sum().item())}
    if affinity is not None:
        out["cpu_affinity"] = affinity

    if world > 1:
        ok, detail = allreduce_parity(kernels, fdist, mode, rank, world, device)
        out["allreduce_parity"] = ok
        out["collective"] = ("fuvs_allreduce_counts (NCCL all-reduce of the int64 counts behind the C ABI)" if fuvs_comm
                             else "torch.distributed.all_reduce")
        if detail:
            out["allreduce_parity_detail"] = detail

    if args.streams > 1:
        # the same work with every clip on ONE stream (kernels strictly back to back): what a single interval costs
        ms1, _, _ = time_resident(kernels, fdist, mode, clips, max(args.steps // 4, 5), 3, args.clips_per_step, device,
                                  world, streams=1)
        iv1 = max(args.steps // 4, 5) * args.clips_per_step * 3
        a1 = bytes_iv * iv1 / (ms1 / 1e3) / 1e9
        out["single_stream"] = {"us_per_interval": ms1 * 1e3 / iv1, "achieved_gbs": a1, "frac": a1 / peak,
                                "value": iv1 * (K_DELTA - 1) * world / (ms1 / 1e3), "unit": "frames/s"}

    if rank == 0 and world == 1 and not args.no_cpu and mode not in ("linear_lowres", "block_lowres", "dense_lowres"):
        # the reference's own op sequence as stock torch-CUDA kernels on the same GPU, inputs resident (context for
        # `value`: there is no Blackwell-specific reference kernel to compare with)
        n_st = 6
        dt_st, lab_st = time_stock_torch(mode, clips[0], device, n_st)
        lab_ours = run_interval(kernels, mode if mode != "block_clip" else "block", clips[0][0], clips[0][1], (n_st - 1) % 3, None, None)
        out["stock_torch_cuda"] = {"value": n_st * (K_DELTA - 1) / dt_st, "unit": "frames/s",
                                   "us_per_interval": dt_st * 1e6 / n_st,
                                   "sample": f"best of 3 x {n_st} intervals of the same {mode} workload, eager ATen kernels "
                                             "(grid_sample, mul, add, cat, max, histc) incl. the metric's host reads",
                                   "label_pixels_differing_from_ours": int((lab_st != lab_ours.long()).sum().item())}
        del lab_st
        torch.cuda.empty_cache()

    host_clips = None
    if not args.no_e2e:
        host_clips = [to_host_clip(c) for c in clips[:2]]
        e_steps = max(min(args.steps // 10, 20), 3)
        e_frames = e_steps * args.clips_per_step * 3 * (K_DELTA - 1)

        def e2e_arm(inputs, hclips, api):
            ems, h2d, d2h, res = time_e2e(mode, e_steps, 3, args.clips_per_step, device, world, hclips, inputs)
            pk_gbs, pk_bytes = time_h2d_peak(mode, 3, args.clips_per_step, device, world, hclips, inputs)
            e_gbs = h2d / (ems / e_steps / 1e3) / 1e9
            return {"value": e_frames * world / (ems / 1e3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e_steps, "ms_per_step": ems / e_steps, "api": api,
                    "temporal_miou": float(res.get("predict_miou1_epoch", float("nan"))),
                    # the arm is bound by the host-to-device copies: its roofline is the rate of plain pinned copies
                    # of the same buffers at the same number of ranks (per GPU)
                    "pcie_gbs": e_gbs, "pcie_peak_gbs": pk_gbs, "frac": e_gbs / pk_gbs if pk_gbs > 0 else None,
                    "pcie_peak_sample": f"3 steps of the same {pk_bytes} B per step, one pinned cudaMemcpyAsync per buffer, "
                                        f"all {world} rank(s) at once"}

        # the reference's batches carry RGB frames (flow/dataset.py): [1,3,H,W] fp32 per key frame, the network output is at
        # 1/8 of the frame and the interval entries take it from there
        gen = torch.Generator().manual_seed(77 + rank)
        img_clips = [([torch.randn((1, 3, H, W), generator=gen).pin_memory() for _ in hk], hg) for hk, hg in host_clips]
        out["e2e"] = e2e_arm("image", img_clips,
                             "FlowBaseModel.predict_step (pinned host RGB key frames [1,3,H,W] + grids -> stride-8 stand-in "
                             "network on the GPU -> interval entry with decoder-resolution key frames -> uint8 labels on host)")
        del img_clips
        # r01's arm for continuity: full-resolution key-frame LOGITS uploaded, identity network
        out["e2e_logits"] = e2e_arm("logits", host_clips,
                                    "FlowBaseModel.predict_step (pinned host key-frame logits [1,C,H,W] + grids -> uint8 labels on host)")
    del clips
    torch.cuda.empty_cache()

    if not args.no_modes:
        modes = {}
        m_steps = max(args.steps // 4, 5)

        def measure(m, label, with_cpu, with_stock):
            mclips = [make_clip(m, device, 5000 + 1000 * rank + i) for i in range(args.distinct_clips)]
            mms, _, _ = time_resident(kernels, fdist, m, mclips, m_steps, 3, args.clips_per_step, device, world,
                                      streams=args.streams)
            miv = m_steps * args.clips_per_step * 3
            mach = algorithmic_bytes(m) * miv / (mms / 1e3) / 1e9
            row = {"value": miv * (K_DELTA - 1) * world / (mms / 1e3), "unit": "frames/s", "achieved_gbs": mach,
                   "frac": mach / peak, "algorithmic_bytes_per_interval": algorithmic_bytes(m),
                   "us_per_interval": mms * 1e3 / miv, "shape": [C, H, W]}
            if rank == 0 and world == 1 and not args.no_cpu and with_stock:
                dt_st, _ = time_stock_torch(m, mclips[0], device, 6)
                row["stock_torch_cuda_us_per_interval"] = dt_st * 1e6 / 6
            if rank == 0 and world == 1 and not args.no_cpu and with_cpu:
                hc = to_host_clip(mclips[0], pin=False)
                n_samp = 60
                dt, lab_cpu = time_cpu(m, ([k.clone() for k in hc[0]], hc[1]), n_samp)
                lab_gpu = run_interval(kernels, m, mclips[0][0], mclips[0][1], (n_samp - 1) % 3, None, None)
                row["cpu_baseline"] = {"value": n_samp * (K_DELTA - 1) / dt, "unit": "frames/s",
                                       "cores": torch.get_num_threads(), "kind": cpu_kind(),
                                       "sample": f"{n_samp} intervals of the same {m} workload at {H}x{W} (1 clip) on torch-CPU",
                                       "seconds": dt,
                                       "label_pixels_differing_from_gpu": int((lab_gpu.cpu().numpy() != lab_cpu).sum()),
                                       "label_pixels": int(lab_cpu.size)}
            modes[label] = row
            del mclips
            torch.cuda.empty_cache()

        for m in ("linear", "linear_lowres", "block", "block_clip", "block_lowres", "dense", "dense_smooth", "dense_lowres"):
            if m != mode:
                measure(m, m, False, m in ("linear", "block", "dense_smooth"))
        # BASELINE.json configs[0]: data.train_w = 433 crops, beside the reference's CPU path
        set_shape(433, 433)
        for m in ("linear", "block", "dense"):
            measure(m, f"{m}_433", True, False)
        set_shape(1080, 1920)
        # BASELINE.json configs[2] / [3]: the feature-based interval at DeepLabV3-R101 and PSPNet feature sizes
        for cf in (2048, 4096):
            try:
                us, fb = time_feature(kernels, cf, device)
                modes[f"feature_{cf}"] = {"us_per_interval": us, "algorithmic_bytes_per_interval": fb,
                                          "achieved_gbs": fb / us / 1e3, "frac": fb / us / 1e3 / peak,
                                          "shape": [cf, H // 8, W // 8], "grid": [H // 16, W // 16],
                                          "entry": "fuvs_feature_interval_ptrs (warp chains + up-sample + blend -> decoder batch)"}
            except Exception as exc:  # noqa: BLE001
                modes[f"feature_{cf}"] = {"error": repr(exc)}
        out["other_modes"] = modes

    if rank == 0 and world == 1 and not args.no_cpu:
        mclip = make_clip(mode, device, 1000 * rank)
        hc = host_clips[0] if host_clips else to_host_clip(mclip, pin=False)
        n_samp = 30 if mode in ("dense", "dense_smooth") else 45   # ~10-15 s of host work on the box's cores
        dt, lab_cpu = time_cpu(mode, ([k.clone() for k in hc[0]], hc[1]), n_samp)
        # same interval on the GPU for an informational label comparison (near-ties may differ: torch-CPU != torch-CUDA)
        lab_gpu = run_interval(kernels, mode if mode != "block_clip" else "block", mclip[0], mclip[1], (n_samp - 1) % 3, None, None)
        mism = int((lab_gpu.cpu().numpy() != lab_cpu).sum())
        out["cpu_baseline"] = {"value": n_samp * (K_DELTA - 1) / dt, "unit": "frames/s", "cores": torch.get_num_threads(),
                               "kind": cpu_kind(), "ref_dir": "oracle/_ref" if cpu_kind() == "reference" else None,
                               "sample": f"{n_samp} intervals of the same {mode} workload (1 clip): the reference's "
                                         "FlowModel.predict -> max(1)[1] -> uint8 -> numpy intersectionAndUnion on torch-CPU",
                               "host_cpu_count": os.cpu_count(), "seconds": dt,
                               "label_pixels_differing_from_gpu": mism, "label_pixels": int(lab_cpu.size)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main_reference(args, rank, world, mode, config):
    """Reference arm: the reference's own CPU implementation of the path (oracle/_ref: its unmodified flow/model.py and
    util/util.py; the oracle port only if that copy is missing), all host threads.  A step is a bounded sample of the
    workload — ONE interval (4 interpolated frames) of the `intervals_per_step` a GPU step holds; the metric is
    per-frame, so the rates compare.  Rank 0 alone runs it at every N; the other ranks exit 0 without work."""
    if rank != 0:
        return
    torch.manual_seed(0)
    clip = make_clip(mode, torch.device("cpu"), 0)
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(args.warmup):
        cpu_interval(mode, clip[0], clip[1], 0, None)
    t0 = time.perf_counter()
    last = None
    for s in range(args.steps):
        _, _, last = cpu_interval(mode, clip[0], clip[1], s % 3, last if s % 3 else None)
    dt = time.perf_counter() - t0
    value = args.steps * (K_DELTA - 1) / dt
    sample = (f"one interval (4 interpolated frames) per step x {args.steps} steps of the {mode} workload on the host "
              f"cores, rank 0 only at every N")
    out = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt * 1e3 / max(args.steps, 1), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
           "impl": "reference", "gpu_launches": 0,
           "cpu_baseline": {"value": value, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": cpu_kind(),
                            "ref_dir": "oracle/_ref" if cpu_kind() == "reference" else None, "sample": sample},
           "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
