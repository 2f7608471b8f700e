"""Parity of every C-ABI entry point against the oracle (oracle/flow_oracle.py, oracle/metric_oracle.py) on the
same seeded inputs.  Bit-exact for label maps, counts and (against torch-CUDA on the same device) logits; the
torch-CPU oracle is matched within 1e-5 relative for warp modes (torch's CPU and CUDA grid_sample differ)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from flood_uav_video_segmentation_b200 import kernels
from flood_uav_video_segmentation_b200.synthetic import clip_keyframes, flow_grids, gt_labels, keyframe_logits
from oracle import flow_oracle as fo
from oracle import metric_oracle as mo

pytestmark = pytest.mark.gpu

ident = torch.nn.Identity()


def bits_equal(a, b):
    return a.shape == b.shape and bool((a.contiguous().view(torch.int32) == b.contiguous().view(torch.int32)).all())


def oracle_interval(o, o_next, gl, gr, n, no_warp, device):
    """Reference call sequence on `device` with identity encoder/decoder: logits [n,C,h,w], labels int64."""
    o, o_next = o.to(device), o_next.to(device)
    gl = [g.to(device) for g in gl]
    gr = [g.to(device) for g in gr]
    logits = fo.predict_segmentation(ident, ident, o, o_next, gl, gr, n, no_warp=no_warp)
    return logits, fo.argmax_labels(logits)


def oracle_temporal(labels_i64, K, last=None):
    lab = labels_i64.cpu().numpy()
    (i, u, t), new_last = mo.temporal_consistency_counts(lab, K, 255, last)
    return np.stack([i, u, t]), new_last


SHAPES = [(5, 433, 433), (2, 433, 433), (5, 64, 96), (5, 37, 53), (3, 270, 480), (7, 48, 64), (12, 40, 40)]


@pytest.mark.parametrize("C,H,W", SHAPES)
@pytest.mark.parametrize("n", [2, 5])
def test_linear_interval(cuda, C, H, W, n):
    o, o_next = keyframe_logits(C, H, W, 0, 0)[None], keyframe_logits(C, H, W, 0, 1)[None]
    dummy = [torch.zeros(1, 1)] * (n - 1)
    ref_logits, ref_labels = oracle_interval(o, o_next, dummy, dummy, n, True, cuda)
    cpu_logits, cpu_labels = oracle_interval(o, o_next, dummy, dummy, n, True, "cpu")
    tc_prev = torch.randint(0, C, (H, W), generator=torch.Generator().manual_seed(5), dtype=torch.uint8)
    counts = kernels.new_counts(C, cuda)
    labels, logits = kernels.linear_blend_argmax(o.to(cuda), o_next.to(cuda), n, want_labels=True, want_logits=True,
                                                 tc_prev=tc_prev.to(cuda), counts=counts)
    assert bits_equal(logits, ref_logits), "logits differ from torch-CUDA oracle"
    assert bits_equal(logits.cpu(), cpu_logits), "linear mode must also be bit-equal to torch-CPU"
    assert torch.equal(labels.long(), ref_labels)
    assert torch.equal(labels.cpu().long(), cpu_labels)
    ref_counts, _ = oracle_temporal(ref_labels, C, tc_prev.numpy().astype(np.int64))
    assert np.array_equal(counts.cpu().numpy(), ref_counts)
    # labels only / counts without tc_prev
    counts2 = kernels.new_counts(C, cuda)
    labels2, none = kernels.linear_blend_argmax(o.to(cuda), o_next.to(cuda), n, counts=counts2)
    assert none is None and torch.equal(labels2, labels)
    ref_counts2, _ = oracle_temporal(ref_labels, C, None)
    assert np.array_equal(counts2.cpu().numpy(), ref_counts2)


@pytest.mark.parametrize("C,hl,wl,H,W", [(5, 135, 240, 1080, 1920), (5, 55, 55, 433, 433), (2, 17, 25, 136, 200),
                                         (3, 34, 60, 270, 480), (5, 40, 40, 32, 64), (4, 1, 7, 9, 52), (7, 12, 16, 96, 128),
                                         (5, 48, 64, 48, 64)])
@pytest.mark.parametrize("n", [1, 2, 5])
def test_linear_lowres_interval(cuda, C, hl, wl, H, W, n):
    """Key frames at decoder resolution: up-sample (align_corners=True) + blend + arg-max +
    counts in one kernel vs F.interpolate followed by the reference sequence on torch-CUDA.  433x433 and C=7 take
    the two-launch route of the wrapper, the others the fused kernel; (5,48,64,48,64) forwards to the plain entry."""
    g = torch.Generator().manual_seed(C * 1000 + hl)
    o_lr = (torch.randn(1, C, hl, wl, generator=g) * 3).to(cuda)
    o_next_lr = (torch.randn(1, C, hl, wl, generator=g) * 3).to(cuda)
    up = (lambda t: F.interpolate(t, size=(H, W), mode="bilinear", align_corners=True)) if (hl, wl) != (H, W) else (lambda t: t)
    dummy = [torch.zeros(1, 1, device=cuda)] * (n - 1)
    ref_logits = fo.predict_segmentation(ident, ident, up(o_lr), up(o_next_lr) if n > 1 else None, dummy, dummy, n, no_warp=True)
    ref_labels = fo.argmax_labels(ref_logits)
    tc_prev = torch.randint(0, C, (H, W), generator=torch.Generator().manual_seed(5), dtype=torch.uint8)
    counts = kernels.new_counts(C, cuda)
    labels, logits = kernels.linear_lowres_blend_argmax(o_lr, o_next_lr if n > 1 else None, (H, W), n, want_labels=True,
                                                        want_logits=True, tc_prev=tc_prev.to(cuda), counts=counts)
    assert bits_equal(logits, ref_logits), f"{int((logits.view(torch.int32) != ref_logits.view(torch.int32)).sum())} logits differ"
    assert torch.equal(labels.long(), ref_labels)
    ref_counts, _ = oracle_temporal(ref_labels, C, tc_prev.numpy().astype(np.int64))
    assert np.array_equal(counts.cpu().numpy(), ref_counts)
    labels2, none = kernels.linear_lowres_blend_argmax(o_lr, o_next_lr if n > 1 else None, (H, W), n)
    assert none is None and torch.equal(labels2, labels)


@pytest.mark.parametrize("C,H,W", [(5, 433, 433), (2, 270, 480), (5, 64, 96), (5, 37, 53), (6, 48, 64)])
@pytest.mark.parametrize("n", [2, 3, 4, 5, 6])
def test_dense_interval(cuda, C, H, W, n):
    o, o_next = keyframe_logits(C, H, W, 1, 0)[None], keyframe_logits(C, H, W, 1, 1)[None]
    gl = flow_grids(H, W, n, "dense", clip=1, side=0)
    gr = flow_grids(H, W, n, "dense", clip=1, side=1)
    ref_logits, ref_labels = oracle_interval(o, o_next, gl, gr, n, False, cuda)
    counts = kernels.new_counts(C, cuda)
    labels, logits = kernels.dense_interval(o.to(cuda), o_next.to(cuda), [g.to(cuda) for g in gl],
                                            [g.to(cuda) for g in gr], n, want_labels=True, want_logits=True,
                                            counts=counts)
    bad = int((logits.view(torch.int32) != ref_logits.view(torch.int32)).sum())
    assert bad == 0, f"{bad} logits differ from torch-CUDA oracle"
    assert torch.equal(labels.long(), ref_labels)
    ref_counts, _ = oracle_temporal(ref_labels, C, None)
    assert np.array_equal(counts.cpu().numpy(), ref_counts)
    cpu_logits, _ = oracle_interval(o, o_next, gl, gr, n, False, "cpu")
    torch.testing.assert_close(logits.cpu(), cpu_logits, rtol=1e-4, atol=1e-4)   # torch-CPU vs torch-CUDA kernels differ


@pytest.mark.parametrize("C,H,W", [(5, 433, 433), (2, 272, 480), (5, 64, 96), (5, 37, 53), (6, 48, 64)])
@pytest.mark.parametrize("n", [2, 5, 7])
def test_block_interval(cuda, C, H, W, n):
    o, o_next = keyframe_logits(C, H, W, 2, 0)[None], keyframe_logits(C, H, W, 2, 1)[None]
    gl = flow_grids(H, W, n, "block", clip=2, side=0)
    gr = flow_grids(H, W, n, "block", clip=2, side=1)
    ref_logits, ref_labels = oracle_interval(o, o_next, gl, gr, n, False, cuda)
    counts = kernels.new_counts(C, cuda)
    tc_prev = torch.randint(0, C, (H, W), generator=torch.Generator().manual_seed(6), dtype=torch.uint8)
    labels, logits = kernels.block_interval(o.to(cuda), o_next.to(cuda), [g.to(cuda) for g in gl],
                                            [g.to(cuda) for g in gr], n, want_labels=True, want_logits=True,
                                            tc_prev=tc_prev.to(cuda), counts=counts)
    bad = int((logits.view(torch.int32) != ref_logits.view(torch.int32)).sum())
    assert bad == 0, f"{bad} logits differ from torch-CUDA oracle"
    assert torch.equal(labels.long(), ref_labels)
    ref_counts, _ = oracle_temporal(ref_labels, C, tc_prev.numpy().astype(np.int64))
    assert np.array_equal(counts.cpu().numpy(), ref_counts)
    cpu_logits, _ = oracle_interval(o, o_next, gl, gr, n, False, "cpu")
    torch.testing.assert_close(logits.cpu(), cpu_logits, rtol=1e-4, atol=1e-4)   # torch-CPU vs torch-CUDA kernels differ


@pytest.mark.parametrize("C,H,W,n", [(5, 272, 480, 5), (5, 272, 480, 3), (3, 64, 96, 5)])
@pytest.mark.parametrize("lowres", [False, True])
def test_block_interval_nan_and_inf(cuda, C, H, W, n, lowres):
    """Key frames with NaN, +Inf and -Inf class values: a frame's arg-max takes the float-domain path unless one of its
    class maxima is NaN (max.NaN), Inf - Inf in a blend makes NaN on the way.  torch.max semantics: the first NaN wins,
    else the first maximum (Inf included)."""
    if lowres:
        hl, wl = H // 8, W // 8
        o, o_next = keyframe_logits(C, hl, wl, 8, 0)[None], keyframe_logits(C, hl, wl, 8, 1)[None]
    else:
        o, o_next = keyframe_logits(C, H, W, 8, 0)[None], keyframe_logits(C, H, W, 8, 1)[None]
    for k, (t, v) in enumerate(((o, float("nan")), (o_next, float("inf")), (o, float("-inf")), (o_next, float("nan")),
                                (o, float("inf")), (o_next, float("-inf")))):
        t[0, k % C, (2 * k + 1)::7, (3 * k)::5] = v
    gl = flow_grids(H, W, n, "block", clip=8, side=0)
    gr = flow_grids(H, W, n, "block", clip=8, side=1)
    oc, onc = o.to(cuda), o_next.to(cuda)
    glc, grc = [g.to(cuda) for g in gl], [g.to(cuda) for g in gr]
    if lowres:
        up = lambda t: F.interpolate(t, size=(H, W), mode="bilinear", align_corners=True)      # noqa: E731
        ref_logits = fo.predict_segmentation(ident, ident, up(oc), up(onc), glc, grc, n, False)
        labels, logits = kernels.block_lowres_interval(oc, onc, (H, W), glc, grc, n, want_labels=True, want_logits=True)
    else:
        ref_logits = fo.predict_segmentation(ident, ident, oc, onc, glc, grc, n, False)
        labels, logits = kernels.block_interval(oc, onc, glc, grc, n, want_labels=True, want_logits=True)
    ref_labels = fo.argmax_labels(ref_logits)
    assert bool(torch.isnan(ref_logits).any()) and bool(torch.isinf(ref_logits).any())
    a, b = logits.reshape(-1), ref_logits.reshape(-1)
    bad = int(((a.view(torch.int32) != b.view(torch.int32)) & ~(torch.isnan(a) & torch.isnan(b))).sum())     # any NaN equals any NaN
    assert bad == 0, f"{bad} logits differ from torch-CUDA oracle"
    assert torch.equal(labels.long().reshape(-1), ref_labels.reshape(-1))


@pytest.mark.parametrize("jitter", [0.05, 0.3, 1.5])
@pytest.mark.parametrize("C,H,W,n", [(5, 272, 480, 5), (2, 96, 256, 4), (7, 64, 132, 3)])
def test_dense_interval_tma_path_and_large_motion(cuda, jitter, C, H, W, n):
    """W % 4 == 0 shapes take the TMA-staged kernel; large jitter pushes taps outside the staged box (global-gather
    branch of the same kernel) and far outside the image (border clip)."""
    o, o_next = keyframe_logits(C, H, W, 4, 0)[None], keyframe_logits(C, H, W, 4, 1)[None]
    gl = flow_grids(H, W, n, "dense", clip=4, side=0, jitter=jitter)
    gr = flow_grids(H, W, n, "dense", clip=4, side=1, jitter=jitter)
    ref_logits, ref_labels = oracle_interval(o, o_next, gl, gr, n, False, cuda)
    tc_prev = torch.randint(0, C, (H, W), generator=torch.Generator().manual_seed(2), dtype=torch.uint8)
    counts = kernels.new_counts(C, cuda)
    labels, logits = kernels.dense_interval(o.to(cuda), o_next.to(cuda), [g.to(cuda) for g in gl],
                                            [g.to(cuda) for g in gr], n, want_labels=True, want_logits=True,
                                            tc_prev=tc_prev.to(cuda), counts=counts)
    bad = int((logits.view(torch.int32) != ref_logits.view(torch.int32)).sum())
    assert bad == 0, f"{bad} logits differ from torch-CUDA oracle"
    assert torch.equal(labels.long(), ref_labels)
    ref_counts, _ = oracle_temporal(ref_labels, C, tc_prev.numpy().astype(np.int64))
    assert np.array_equal(counts.cpu().numpy(), ref_counts)
    # labels-only call (no logits materialised) gives the same maps
    labels2, _ = kernels.dense_interval(o.to(cuda), o_next.to(cuda), [g.to(cuda) for g in gl], [g.to(cuda) for g in gr], n)
    assert torch.equal(labels2, labels)


@pytest.mark.parametrize("C,fh,fw,hg,wg", [(48, 40, 64, 20, 32), (48, 40, 64, 40, 64), (7, 33, 50, 16, 25), (130, 17, 24, 8, 12)])
@pytest.mark.parametrize("n", [1, 2, 5])
def test_feature_interval(cuda, C, fh, fw, hg, wg, n):
    """fuvs_feature_interval vs the restated FlowModel.predict_feature (flow/model.py:116-181) with identity encoder /
    decoder on torch-CUDA: chains at grid resolution, up-sample + blend into the decoder batch, key frame through the
    default grid with align_corners=True."""
    g = torch.Generator().manual_seed(C + fh)
    f, f_next = torch.randn(1, C, fh, fw, generator=g).to(cuda), torch.randn(1, C, fh, fw, generator=g).to(cuda)
    base = torch.stack(torch.meshgrid((torch.arange(hg) + 0.5) / hg * 2 - 1, (torch.arange(wg) + 0.5) / wg * 2 - 1,
                                      indexing="ij")[::-1], -1)

    def grids():
        return [(base + (torch.rand(base.shape, generator=g) - 0.5) * 0.2).unsqueeze(0).to(cuda) for _ in range(n - 1)]

    gl, gr = grids(), grids()
    default = torch.stack(torch.meshgrid((torch.arange(9) + 0.5) / 9 * 2 - 1, (torch.arange(13) + 0.5) / 13 * 2 - 1,
                                         indexing="ij")[::-1], -1).unsqueeze(0).float().to(cuda)
    ref = fo.predict_feature(ident, ident, f, f_next if n > 1 else None, gl, gr, n, default)
    got = kernels.feature_interval(f, f_next if n > 1 else None, gl, gr, n, default_grid=default)
    assert got.shape == ref.shape
    assert bits_equal(got, ref), f"{int((got.view(torch.int32) != ref.view(torch.int32)).sum())} elements differ"
    # without a default grid frame 0 is the key frame's features themselves
    got2 = kernels.feature_interval(f, f_next if n > 1 else None, gl, gr, n)
    assert bits_equal(got2[0], f[0]) and bits_equal(got2[1:], ref[1:])


@pytest.mark.parametrize("prev_kind", ["none", "labels", "out_of_range"])
@pytest.mark.parametrize("C,H,W", [(5, 64, 256), (2, 96, 128), (5, 136, 384)])
@pytest.mark.parametrize("n", [3, 5, 6, 9])
def test_dense_interval_counts_with_foreign_labels(cuda, C, H, W, n, prev_kind):
    """The interval's temporal counts (bit-plane counters, temporal_fields.cuh) with and without tc_prev; out-of-range
    labels in tc_prev (a caller's own map: ignore_index, or neither a class nor ignore_index) take the per-label path;
    the entry accumulates into the caller's counts."""
    o, o_next = keyframe_logits(C, H, W, 5, 0)[None], keyframe_logits(C, H, W, 5, 1)[None]
    gl = flow_grids(H, W, n, "dense", clip=5, side=0)
    gr = flow_grids(H, W, n, "dense", clip=5, side=1)
    _, ref_labels = oracle_interval(o, o_next, gl, gr, n, False, cuda)
    tc_prev = None
    if prev_kind != "none":
        tc_prev = torch.randint(0, C, (H, W), generator=torch.Generator().manual_seed(n), dtype=torch.uint8)
        if prev_kind == "out_of_range":
            tc_prev[::3, ::5] = 255          # ignore_index
            tc_prev[1::7, 2::11] = C + 1     # neither a class nor ignore_index
    counts = kernels.new_counts(C, cuda)
    counts += 7
    labels, _ = kernels.dense_interval(o.to(cuda), o_next.to(cuda), [g.to(cuda) for g in gl], [g.to(cuda) for g in gr], n,
                                       tc_prev=None if tc_prev is None else tc_prev.to(cuda), counts=counts)
    assert torch.equal(labels.long(), ref_labels)
    ref_counts, _ = oracle_temporal(ref_labels, C, None if tc_prev is None else tc_prev.numpy().astype(np.int64))
    assert np.array_equal(counts.cpu().numpy() - 7, ref_counts)


def test_full_size_1080p_all_modes(cuda):
    """BASELINE.json configs at full size, one interval each, against the oracle on torch-CUDA."""
    C, H, W, n = 5, 1080, 1920, 5
    ks = clip_keyframes(C, H, W, n, frames=6, clip=3)
    o, o_next = ks[0].to(cuda), ks[1].to(cuda)
    for mode in ("linear", "block", "dense"):
        if mode == "linear":
            gl = gr = [torch.zeros(1, 1)] * (n - 1)
            labels, _ = kernels.linear_blend_argmax(o, o_next, n)
        else:
            gl = [g.to(cuda) for g in flow_grids(H, W, n, mode, clip=3, side=0)]
            gr = [g.to(cuda) for g in flow_grids(H, W, n, mode, clip=3, side=1)]
            fn = kernels.block_interval if mode == "block" else kernels.dense_interval
            tc_prev = torch.randint(0, C, (H, W), generator=torch.Generator().manual_seed(9), dtype=torch.uint8)
            counts = kernels.new_counts(C, cuda)
            labels, _ = fn(o, o_next, gl, gr, n, tc_prev=tc_prev.to(cuda), counts=counts)
        ref = fo.argmax_labels(fo.predict_segmentation(ident, ident, o, o_next, gl, gr, n, no_warp=(mode == "linear")))
        diff = int((labels.long() != ref).sum())
        assert diff == 0, f"{mode}: {diff} label pixels differ at 1080p"
        if mode != "linear":   # counts accumulated by the interval call itself (dense: fused into the last step)
            ref_counts, _ = oracle_temporal(ref, C, tc_prev.numpy().astype(np.int64))
            assert np.array_equal(counts.cpu().numpy(), ref_counts), f"{mode}: temporal counts differ at 1080p"
        del ref
        torch.cuda.empty_cache()


def test_argmax_ties_and_nan(cuda):
    C, H, W = 5, 8, 16
    x = torch.zeros(1, C, H, W)
    x[0, :, 0, 0] = torch.tensor([1.0, 3.0, 3.0, 2.0, 3.0])           # tie -> lowest index
    x[0, :, 0, 1] = torch.tensor([1.0, float("nan"), 5.0, float("nan"), 0.0])  # NaN beats numbers, first NaN
    x[0, :, 0, 2] = torch.tensor([-0.0, 0.0, -0.0, 0.0, -0.0])          # -0 == +0 -> index 0
    x[0, :, 0, 3] = torch.tensor([float("-inf")] * 5)
    x[0, :, 0, 4] = torch.tensor([0.0, float("inf"), float("inf"), 1.0, 2.0])
    # pixels 8..11 share a thread with no NaN / Inf in it: the float-domain arg-max of the fast path (pix4.cuh)
    x[0, :, 0, 8] = torch.tensor([-0.0, 0.0, -0.0, 0.0, -0.0])          # -0 == +0 -> index 0
    x[0, :, 0, 9] = torch.tensor([-1.0, -0.0, 0.0, -2.0, 0.0])          # -> index 1
    x[0, :, 0, 10] = torch.tensor([-3.0, -2.0, -1.0, -0.5, 7.0])        # last class
    x[0, :, 0, 11] = torch.tensor([1e-45, 1e-45, 0.0, -1e-45, 1e-45])   # denormal ties
    x[0, :, 1:, :] = torch.randn(C, H - 1, W, generator=torch.Generator().manual_seed(3)).round()  # many ties
    xc = x.to(cuda)
    ref = xc.max(1)[1]
    assert torch.equal(kernels.argmax(xc).long(), ref)
    assert torch.equal(kernels.argmax(xc, dtype=torch.int64), ref)
    labels, _ = kernels.linear_blend_argmax(xc, xc, 3)
    ref3 = fo.argmax_labels(fo.predict_segmentation(ident, ident, xc, xc, [torch.zeros(1, 1)] * 2, [torch.zeros(1, 1)] * 2, 3, True))
    assert torch.equal(labels.long(), ref3)


@pytest.mark.parametrize("shape", [((5, 67, 120), (1072, 1920)), ((3, 55, 55), (433, 433)), ((2, 9, 7), (10, 8)),
                                   ((4, 33, 33), (33, 33)), ((2, 1, 1), (5, 7)), ((2, 6, 6), (1, 1)),
                                   # the staged separable kernel (ratio >= 4, Wout % 4 == 0): ragged chunks, several plane chunks
                                   ((5, 135, 240), (1080, 1920)), ((8, 17, 25), (40, 52)), ((20, 34, 60), (68, 120)),
                                   ((2, 2, 3), (9, 8)), ((3, 7, 300), (33, 700)), ((1, 1, 6), (8, 12)),
                                   ((20, 6, 10), (25, 44)), ((9, 5, 70), (41, 520))])
def test_upsample(cuda, shape):
    (C, Hin, Win), (Ho, Wo) = shape
    x = torch.randn(2, C, Hin, Win, generator=torch.Generator().manual_seed(9)).to(cuda)
    ref = F.interpolate(x, size=(Ho, Wo), mode="bilinear", align_corners=True)
    assert bits_equal(kernels.upsample_bilinear_ac(x, (Ho, Wo)), ref)


@pytest.mark.parametrize("align", [False, True])
def test_warp_step(cuda, align):
    C, H, W = 6, 135, 240
    x = torch.randn(1, C, H, W, generator=torch.Generator().manual_seed(11)).to(cuda)
    g = flow_grids(1072, 1920, 2, "block", clip=4, jitter=0.1)[0].to(cuda)
    ref = F.grid_sample(x, g, mode="bilinear", padding_mode="border", align_corners=align)
    out, _ = kernels.warp_step(x[0], g, align_corners=align)
    assert bits_equal(out, ref[0])
    # two-sided launch, many channels (feature-map shape class)
    xf = torch.randn(2, 300, 34, 60, generator=torch.Generator().manual_seed(12)).to(cuda)
    g2 = flow_grids(272, 480, 3, "block", clip=5, jitter=0.2)
    a, b = kernels.warp_step(xf[0], g2[0].to(cuda), xf[1], g2[1].to(cuda), align_corners=align)
    assert bits_equal(a, F.grid_sample(xf[0:1], g2[0].to(cuda), mode="bilinear", padding_mode="border", align_corners=align)[0])
    assert bits_equal(b, F.grid_sample(xf[1:2], g2[1].to(cuda), mode="bilinear", padding_mode="border", align_corners=align)[0])


def test_warp_step_special_coordinates(cuda):
    """NaN / Inf / far out-of-range grid values and Inf source pixels at the border (skipped taps)."""
    C, H, W = 2, 5, 6
    x = torch.arange(C * H * W, dtype=torch.float32).reshape(1, C, H, W)
    x[0, 0, 0, W - 1] = float("inf")
    x[0, 1, H - 1, 0] = float("nan")
    g = torch.tensor([[-1.0, -1.0], [1.0, 1.0], [float("nan"), 0.0], [float("inf"), float("-inf")], [5.0, -7.0],
                      [1.0, -1.0], [-1.0, 1.0], [0.999, 0.3], [0.0, 0.0]]).reshape(1, 3, 3, 2)
    ref = F.grid_sample(x.to(cuda), g.to(cuda), mode="bilinear", padding_mode="border", align_corners=False)
    out, _ = kernels.warp_step(x[0].to(cuda), g.to(cuda))
    assert torch.equal(torch.isnan(out), torch.isnan(ref[0]))
    assert bits_equal(torch.nan_to_num(out, nan=7.0), torch.nan_to_num(ref[0], nan=7.0))


def test_blend_argmax(cuda):
    a = torch.randn(3, 5, 37, 53, generator=torch.Generator().manual_seed(1)).to(cuda)
    b = torch.randn(3, 5, 37, 53, generator=torch.Generator().manual_seed(2)).to(cuda)
    wa, wb = 3 / 7, 4 / 7
    ref = wa * a + wb * b
    out, labels = kernels.blend_argmax(a, b, wa, wb, want_labels=True)
    assert bits_equal(out, ref) and torch.equal(labels.long(), ref.max(1)[1])
    out1, _ = kernels.blend_argmax(a, None, wa, 0.0)
    assert bits_equal(out1, a * wa)


@pytest.mark.parametrize("K", [5, 2, 1, 19])
@pytest.mark.parametrize("pdt,tdt", [(torch.int64, torch.int64), (torch.uint8, torch.int64), (torch.uint8, torch.uint8)])
def test_confusion_vs_oracle(cuda, K, pdt, tdt):
    H, W = 211, 307
    g = torch.Generator().manual_seed(K)
    pred = torch.randint(0, K, (2, H, W), generator=g).to(pdt)
    target = torch.randint(0, K + 3, (2, H, W), generator=g)     # includes out-of-range classes K..K+2
    target[torch.rand(2, H, W, generator=g) < 0.05] = 255
    target = target.to(tdt)
    # numpy convention (util/util.py:36-47)
    i, u, t = mo.intersection_and_union_np(pred.numpy(), target.numpy(), K, 255)
    got = kernels.confusion(pred.to(cuda), target.to(cuda), K, 255, numpy_bins=True).cpu().numpy()
    assert np.array_equal(got, np.stack([i, u, t]))
    # torch.histc convention (util/util.py:52-63), restated integer binning
    i, u, t = mo.intersection_and_union_histc_ints(pred.numpy(), target.numpy(), K, 255)
    pc = pred.to(cuda)
    got = kernels.confusion(pc, target.to(cuda), K, 255, mutate_pred=True).cpu().numpy()
    assert np.array_equal(got, np.stack([i, u, t]))
    # in-place ignore substitution on the caller's tensor (util/util.py:57)
    exp = pred.clone()
    exp[target == 255] = 255
    assert torch.equal(pc.cpu(), exp)
    if pdt == torch.int64 and tdt == torch.int64:
        # the real thing: torch.histc on int64 exists only on CUDA
        o2 = pred.to(cuda).clone()
        ri, ru, rt = mo.intersection_and_union_torch(o2, target.to(cuda), K, 255)
        assert np.array_equal(got, torch.stack([ri, ru, rt]).cpu().numpy().astype(np.int64))


@pytest.mark.parametrize("K", [2, 3, 4, 5])
@pytest.mark.parametrize("pdt,tdt", [(torch.int64, torch.int64), (torch.uint8, torch.int64), (torch.uint8, torch.uint8),
                                     (torch.int64, torch.uint8)])
def test_confusion_fast_path_edges(cuda, K, pdt, tdt):
    """16-labels-per-thread path of fuvs_confusion (histc binning, 2 <= K <= 5): predictions outside [0, K) (negative
    for int64), targets K..K+2 and 255, a length that leaves a tail, views that start off a 16-byte boundary (general
    kernel), accumulation over calls, with and without the in-place ignore substitution."""
    N = 16 * 4099 + 7
    g = torch.Generator().manual_seed(100 + K)
    lo = -2 if pdt == torch.int64 else 0
    pred = torch.randint(lo, K + 2, (N + 3,), generator=g).to(pdt)
    target = torch.randint(0, K + 3, (N + 3,), generator=g)
    target[torch.rand(N + 3, generator=g) < 0.07] = 255
    target = target.to(tdt)
    counts = kernels.new_counts(K, cuda)
    tot = np.zeros((3, K), np.int64)
    for off, mutate in ((0, False), (0, True), (3, True), (1, False)):
        p, t = pred[off:off + N], target[off:off + N]
        pc, tc = pred.to(cuda)[off:off + N], target.to(cuda)[off:off + N]
        kernels.confusion(pc, tc, K, 255, counts=counts, mutate_pred=mutate)
        tot += np.stack(mo.intersection_and_union_histc_ints(p.numpy(), t.numpy(), K, 255))
        exp = p.clone()
        if mutate:
            exp[t == 255] = 255
        assert torch.equal(pc.cpu(), exp)
        assert np.array_equal(counts.cpu().numpy(), tot)


@pytest.mark.parametrize("ignore", [-1, 300, 7])
def test_confusion_ignore_outside_byte_range(cuda, ignore):
    """ignore_index values the byte-domain fast path cannot represent (negative, >= 256) or that sit just above the
    classes: int64 targets holding them must be counted exactly like util/util.py:52-63 does."""
    K, N = 5, 512 * 37 + 100
    g = torch.Generator().manual_seed(ignore + 10)
    pred = torch.randint(0, K, (N,), generator=g)
    target = torch.randint(0, K, (N,), generator=g)
    target[torch.rand(N, generator=g) < 0.1] = ignore
    for pdt in (torch.int64, torch.uint8):
        if pdt == torch.uint8 and not 0 <= ignore <= 255:
            continue                            # torch (and numpy) refuse to store such a value in a uint8 tensor
        p = pred.to(pdt)
        exp_counts = np.stack(mo.intersection_and_union_histc_ints(p.numpy(), target.numpy(), K, ignore))
        pc = p.to(cuda)
        got = kernels.confusion(pc, target.to(cuda), K, ignore, mutate_pred=True).cpu().numpy()
        assert np.array_equal(got, exp_counts), f"pred {pdt}"
        exp = p.clone()
        exp[target == ignore] = ignore
        assert torch.equal(pc.cpu(), exp)


def test_confusion_accumulates_and_ragged(cuda):
    K = 5
    counts = kernels.new_counts(K, cuda)
    tot = np.zeros((3, K), np.int64)
    for N in (0, 1, 31, 33, 1000, 4097):
        g = torch.Generator().manual_seed(N)
        pred = torch.randint(0, K, (N,), generator=g)
        target = torch.randint(0, K, (N,), generator=g)
        if N:
            kernels.confusion(pred.to(cuda), target.to(cuda), K, 255, counts=counts)
            tot += np.stack(mo.intersection_and_union_histc_ints(pred.numpy(), target.numpy(), K, 255))
    assert np.array_equal(counts.cpu().numpy(), tot)


@pytest.mark.parametrize("K,H,W", [(5, 433, 433), (5, 64, 96), (12, 37, 53)])
def test_temporal_counts(cuda, K, H, W):
    g = torch.Generator().manual_seed(K + H)
    labels = torch.randint(0, K, (6, H, W), generator=g, dtype=torch.uint8)
    last = torch.randint(0, K, (H, W), generator=g, dtype=torch.uint8)
    for prev in (None, last):
        ref, _ = oracle_temporal(labels.long(), K, None if prev is None else prev.numpy().astype(np.int64))
        got = kernels.temporal_counts(labels.to(cuda), K, 255, tc_prev=None if prev is None else prev.to(cuda))
        assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.parametrize("foreign", [False, True])
@pytest.mark.parametrize("n", [1, 2, 5, 7])
@pytest.mark.parametrize("K", [1, 2, 3, 4, 5])
def test_temporal_counts_bit_planes(cuda, K, n, foreign):
    """The 16-labels-per-thread path (K <= 5, HW % 16 == 0): bit-plane counters, n = 5 with all loads up front.  With
    `foreign`, label maps carry values that are no class: K+1 .. 7 (fit the planes), 8 .. 16 and 200 (do not), 255 =
    ignore_index — every 16-pixel group that holds one takes the per-label path.  (The value K itself is left out:
    np.histogram's closed last bin would count it as class K-1, the library counts a value iff it is < K — see
    include/fuvs.h; arg-max output never holds it.)"""
    H, W = 48, 80
    g = torch.Generator().manual_seed(100 * K + n)
    labels = torch.randint(0, K, (n, H, W), generator=g, dtype=torch.uint8)
    last = torch.randint(0, K, (H, W), generator=g, dtype=torch.uint8)
    if foreign:
        for i, v in enumerate((K + 1, 7, 8, 15, 16, 200, 255)):
            labels[i % n, (3 * i) % H::11, (5 * i) % W::13] = v
            last[(2 * i + 1) % H::9, (7 * i) % W::17] = v
    for prev in (None, last):
        ref, _ = oracle_temporal(labels.long(), K, None if prev is None else prev.numpy().astype(np.int64))
        got = kernels.temporal_counts(labels.to(cuda), K, 255, tc_prev=None if prev is None else prev.to(cuda))
        assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.parametrize("n", [16, 29, 30, 31, 45, 60])
def test_counter_fields_saturate(cuda, n):
    """Every pixel of every frame has the same label, so one 6-bit counter field (C = 5) takes all the load: the
    spill schedule of the field-packed counters (linear, block-rows and temporal-counts kernels) must keep it below 64
    between spills for any interval length, including the two extra pairs the block kernel counts after its frame loop."""
    C, H, W = 5, 64, 96
    g = torch.Generator().manual_seed(n)
    o, o_next = torch.randn(1, C, H, W, generator=g), torch.randn(1, C, H, W, generator=g)
    o[:, 2] += 30.0
    o_next[:, 2] += 30.0
    tc_prev = torch.full((H, W), 2, dtype=torch.uint8)
    gl = flow_grids(H, W, n, "block", clip=3, side=0)
    gr = flow_grids(H, W, n, "block", clip=3, side=1)
    for no_warp in (True, False):
        _, ref_labels = oracle_interval(o, o_next, gl, gr, n, no_warp, cuda)
        assert bool((ref_labels == 2).all())
        ref_counts, _ = oracle_temporal(ref_labels, C, tc_prev.numpy().astype(np.int64))
        counts = kernels.new_counts(C, cuda)
        if no_warp:
            labels, _ = kernels.linear_blend_argmax(o[0].to(cuda), o_next[0].to(cuda), n, tc_prev=tc_prev.to(cuda), counts=counts)
        else:
            labels, _ = kernels.block_interval(o.to(cuda), o_next.to(cuda), [x.to(cuda) for x in gl], [x.to(cuda) for x in gr],
                                               n, want_labels=True, tc_prev=tc_prev.to(cuda), counts=counts)
        assert torch.equal(labels.long(), ref_labels)
        assert np.array_equal(counts.cpu().numpy(), ref_counts), f"no_warp={no_warp}"
        got = kernels.temporal_counts(labels, C, 255, tc_prev=tc_prev.to(cuda))
        assert np.array_equal(got.cpu().numpy(), ref_counts)


@pytest.mark.parametrize("mode", ["linear", "linear_lowres", "dense", "block"])
def test_interval_entries_replay_from_a_cuda_graph(cuda, mode):
    """the entry points only enqueue work, so a captured call replays bit-identically.  The block
    route takes its per-step chain launches while the stream is being captured (the cooperative launch otherwise)."""
    C, H, W, n = 5, 272, 480, 5
    o, o_next = keyframe_logits(C, H, W, 3, 0).to(cuda), keyframe_logits(C, H, W, 3, 1).to(cuda)
    kind = "dense" if mode == "dense" else "block"
    gl = [g.to(cuda) for g in flow_grids(H, W, n, kind, clip=3, side=0)]
    gr = [g.to(cuda) for g in flow_grids(H, W, n, kind, clip=3, side=1)]
    tc_prev = torch.randint(0, C, (H, W), generator=torch.Generator().manual_seed(8), dtype=torch.uint8).to(cuda)
    lo, lo_next = o[:, ::8, ::8].contiguous(), o_next[:, ::8, ::8].contiguous()

    def call(counts):
        if mode == "linear":
            return kernels.linear_blend_argmax(o, o_next, n, tc_prev=tc_prev, counts=counts)[0]
        if mode == "linear_lowres":
            return kernels.linear_lowres_blend_argmax(lo, lo_next, (H, W), n, tc_prev=tc_prev, counts=counts)[0]
        if mode == "dense":
            return kernels.dense_interval(o, o_next, gl, gr, n, tc_prev=tc_prev, counts=counts)[0]
        return kernels.block_interval(o, o_next, gl, gr, n, tc_prev=tc_prev, counts=counts)[0]

    eager_counts = kernels.new_counts(C, cuda)
    eager_labels = call(eager_counts).clone()
    torch.cuda.synchronize()
    graph_counts = kernels.new_counts(C, cuda)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        graph_labels = call(graph_counts)
    graph_counts.zero_()
    graph_labels.zero_()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(graph_labels, eager_labels)
    assert torch.equal(graph_counts, 3 * eager_counts)


def test_errors_are_loud(cuda):
    with pytest.raises(kernels.FuvsError):
        kernels.linear_blend_argmax(torch.zeros(5, 8, 8), torch.zeros(5, 8, 8), 5)          # CPU tensors
    with pytest.raises(kernels.FuvsError):
        kernels.linear_blend_argmax(torch.zeros(5, 8, 8, device=cuda), torch.zeros(5, 8, 8, device=cuda), 500)
    with pytest.raises(kernels.FuvsError):
        kernels.confusion(torch.zeros(4, device=cuda), torch.zeros(4, dtype=torch.int64, device=cuda), 5)
    # fuvs_dense_lowres_interval_ptrs: the two up-sample buffers must be distinct, the key frames and buffers non-NULL
    from flood_uav_video_segmentation_b200._lib import check, load, ptr, ptr_array, stream_ptr
    lib = load()
    C, hl, wl, H, W, n = 5, 8, 16, 64, 128, 3
    lr = torch.zeros(C, hl, wl, device=cuda)
    up = torch.empty(C * H * W, device=cuda)
    up2 = torch.empty(C * H * W, device=cuda)
    grids = [torch.zeros(1, H, W, 2, device=cuda) for _ in range(n - 1)]
    scratch = torch.empty(int(lib.fuvs_dense_scratch_floats(C, H, W, n)), device=cuda)
    labels = torch.empty((n, H, W), dtype=torch.uint8, device=cuda)

    def call(prev_lr, next_lr, prev_up, next_up):
        check(lib.fuvs_dense_lowres_interval_ptrs(ptr(prev_lr), ptr(next_lr), hl, wl, ptr(prev_up), 0, ptr(next_up), ptr_array(grids),
                                                  ptr_array(grids), C, H, W, n, ptr(scratch), ptr(labels), None, None, None, 255,
                                                  stream_ptr(cuda)))
    call(lr, lr, up, up2)                                   # the valid call
    for bad in ((lr, lr, up, up), (None, lr, up, up2), (lr, None, up, up2), (lr, lr, None, up2), (lr, lr, up, None)):
        with pytest.raises(kernels.FuvsError):
            call(*bad)


def test_limits_and_degenerate_shapes(cuda):
    """Maximum interval length (60 frames), maximum class count for uint8 label maps (256), one-pixel frames, empty
    metric inputs, and the loud failures just beyond the limits."""
    g = torch.Generator().manual_seed(11)
    # n = FUVS_MAX_FRAMES through the linear entry (C = 3: field counters spill many times)
    C, H, W, n = 3, 24, 32, 60
    o, o_next = torch.randn(1, C, H, W, generator=g), torch.randn(1, C, H, W, generator=g)
    dummy = [torch.zeros(1, 1)] * (n - 1)
    ref = fo.argmax_labels(fo.predict_segmentation(ident, ident, o, o_next, dummy, dummy, n, no_warp=True))
    counts = kernels.new_counts(C, cuda)
    labels, _ = kernels.linear_blend_argmax(o.to(cuda), o_next.to(cuda), n, counts=counts)
    assert torch.equal(labels.cpu().long(), ref)
    ref_counts, _ = oracle_temporal(ref, C, None)
    assert np.array_equal(counts.cpu().numpy(), ref_counts)
    with pytest.raises(kernels.FuvsError):
        kernels.linear_blend_argmax(o.to(cuda), o_next.to(cuda), 61)
    # C = 256 classes: generic-C kernels, uint8 labels up to 255, shared-memory histogram counts
    C, H, W, n = 256, 9, 12, 3
    o, o_next = torch.randn(1, C, H, W, generator=g), torch.randn(1, C, H, W, generator=g)
    dummy = [torch.zeros(1, 1)] * (n - 1)
    ref = fo.argmax_labels(fo.predict_segmentation(ident, ident, o, o_next, dummy, dummy, n, no_warp=True))
    counts = kernels.new_counts(C, cuda)
    # ignore_index = 255 collides with class 255 here: the reference then moves ignored pixels INTO class 255
    labels, _ = kernels.linear_blend_argmax(o.to(cuda), o_next.to(cuda), n, counts=counts, ignore_index=255)
    assert torch.equal(labels.cpu().long(), ref) and int(ref.max()) > 200
    ref_counts, _ = oracle_temporal(ref, C, None)
    assert np.array_equal(counts.cpu().numpy(), ref_counts)
    with pytest.raises(kernels.FuvsError):
        kernels.linear_blend_argmax(torch.zeros(1, 257, 4, 4, device=cuda), torch.zeros(1, 257, 4, 4, device=cuda), 2)
    # one-pixel frame, every mode that accepts it
    o, o_next = torch.tensor([[[[1.0]], [[3.0]], [[2.0]]]]), torch.tensor([[[[5.0]], [[0.0]], [[4.0]]]])
    ref = fo.argmax_labels(fo.predict_segmentation(ident, ident, o, o_next, [torch.zeros(1, 1)] * 3, [torch.zeros(1, 1)] * 3, 4,
                                                   no_warp=True))
    labels, _ = kernels.linear_blend_argmax(o.to(cuda), o_next.to(cuda), 4)
    assert torch.equal(labels.cpu().long(), ref)
    grid = [torch.zeros(1, 1, 1, 2)] * 3
    ref = fo.argmax_labels(fo.predict_segmentation(ident, ident, o.to(cuda), o_next.to(cuda), [x.to(cuda) for x in grid],
                                                   [x.to(cuda) for x in grid], 4))
    labels, _ = kernels.dense_interval(o.to(cuda), o_next.to(cuda), [x.to(cuda) for x in grid], [x.to(cuda) for x in grid], 4)
    assert torch.equal(labels.long(), ref)
    # empty metric input: nothing is counted, nothing fails
    counts = kernels.confusion(torch.zeros(0, dtype=torch.uint8, device=cuda), torch.zeros(0, dtype=torch.int64, device=cuda), 5)
    assert int(counts.sum()) == 0


@pytest.mark.parametrize("C,H,W,n,m", [(5, 272, 480, 5, 3), (2, 64, 96, 4, 2), (5, 37, 53, 3, 4), (5, 48, 64, 1, 2),
                                       (6, 48, 64, 5, 9)])
def test_block_clip(cuda, C, H, W, n, m):
    """fuvs_block_clip (m consecutive intervals, chain steps batched over the clip) against m successive
    fuvs_block_interval calls and against the oracle: labels, logits and the chained temporal counts bit-exact."""
    keys = [keyframe_logits(C, H, W, 4, j)[None].to(cuda) for j in range(m + 1)]
    gl = [[g.to(cuda) for g in flow_grids(H, W, n, "block", clip=4, interval=i, side=0)] for i in range(m)]
    gr = [[g.to(cuda) for g in flow_grids(H, W, n, "block", clip=4, interval=i, side=1)] for i in range(m)]
    tc0 = torch.randint(0, C, (H, W), generator=torch.Generator().manual_seed(3), dtype=torch.uint8).to(cuda)
    counts = kernels.new_counts(C, cuda)
    labels, logits = kernels.block_clip(keys, gl, gr, n, want_logits=True, tc_prev=tc0, counts=counts)
    assert tuple(labels.shape) == (m, n, H, W) and tuple(logits.shape) == (m, n, C, H, W)
    counts_ref = kernels.new_counts(C, cuda)
    ref_counts = np.zeros((3, C), np.int64)
    last, last_np = tc0, tc0.cpu().numpy().astype(np.int64)
    for i in range(m):
        li, gi = kernels.block_interval(keys[i], keys[i + 1] if n > 1 else None, gl[i], gr[i], n, want_labels=True,
                                        want_logits=True, tc_prev=last, counts=counts_ref)
        assert torch.equal(labels[i], li), f"interval {i}: labels differ from the per-interval entry"
        assert bits_equal(logits[i], gi), f"interval {i}: logits differ from the per-interval entry"
        ref_logits, ref_labels = oracle_interval(keys[i], keys[i + 1] if n > 1 else keys[i], gl[i], gr[i], n, False, cuda)
        if n == 1:
            ref_logits, ref_labels = keys[i], fo.argmax_labels(keys[i])
        assert bits_equal(logits[i], ref_logits), f"interval {i}: logits differ from the oracle"
        assert torch.equal(labels[i].long(), ref_labels)
        c, last_np = oracle_temporal(ref_labels, C, last_np)
        ref_counts += c
        last = li[n - 1]
    assert torch.equal(counts, counts_ref)
    assert np.array_equal(counts.cpu().numpy(), ref_counts)


@pytest.mark.parametrize("F_,C,Hin,Win,Hout,Wout", [(5, 5, 1080, 1920, 1072, 1920), (2, 5, 433, 433, 1072, 1920),
                                                     (3, 2, 37, 53, 41, 59), (1, 7, 17, 25, 136, 201),
                                                     (2, 5, 48, 64, 48, 64), (1, 3, 1, 1, 5, 7)])
def test_upsample_argmax(cuda, F_, C, Hin, Win, Hout, Wout):
    """fuvs_upsample_argmax = predict_step's F.interpolate(.., (1072,1920), bilinear, align_corners=True) + max(1)[1] +
    uint8 (flow/base.py:275-277) in one kernel: labels and (optional) resized logits bit-exact against torch-CUDA."""
    g = torch.Generator().manual_seed(F_ * 100 + Hin)
    x = (torch.randn(F_, C, Hin, Win, generator=g) * 2).to(cuda)
    x[0, :, 0, 0] = 0.5            # a tie: the lowest class index must win
    if Hin > 2 and Win > 2:
        x[0, 1, 1, 1] = float("nan")
        x[0, C - 1, 2, 1] = float("inf")
    ref = F.interpolate(x, size=(Hout, Wout), mode="bilinear", align_corners=True)
    labels, resized = kernels.upsample_argmax(x, (Hout, Wout), want_resized=True)
    assert bits_equal(resized, ref)
    assert torch.equal(labels.long(), ref.max(1)[1])
    labels2, none = kernels.upsample_argmax(x, (Hout, Wout))
    assert none is None and torch.equal(labels2, labels)


def test_grid_lists_are_passed_by_pointer(cuda):
    """The reference hands the grids as a python list of separate tensors (flow/dataset.py:138-146): the interval entries
    take them as a pointer table, so grids living in unrelated allocations (and views at odd offsets of a bigger
    buffer) give the same result as one stacked tensor."""
    C, H, W, n = 5, 64, 96, 5
    o, o_next = keyframe_logits(C, H, W, 6, 0)[None].to(cuda), keyframe_logits(C, H, W, 6, 1)[None].to(cuda)
    for mode, fn in (("dense", kernels.dense_interval), ("block", kernels.block_interval)):
        gl = [g.to(cuda) for g in flow_grids(H, W, n, mode, clip=6, side=0)]
        gr = [g.to(cuda) for g in flow_grids(H, W, n, mode, clip=6, side=1)]
        big = torch.zeros(gl[0].numel() * 2 * n + 64, device=cuda)
        views = []
        for j, gsrc in enumerate(gl):                      # scattered views, 8-byte aligned, in reverse order
            off = (n - 2 - j) * 2 * gsrc.numel() + 2 * j
            v = big[off:off + gsrc.numel()].view(gsrc.shape)
            v.copy_(gsrc)
            views.append(v)
        a = fn(o, o_next, torch.cat(gl, 0), torch.cat(gr, 0), n, want_logits=True)
        b = fn(o, o_next, views, gr, n, want_logits=True)
        assert torch.equal(a[0], b[0]) and bits_equal(a[1], b[1]), mode


@pytest.mark.parametrize("C,hl,wl,H,W", [(5, 135, 240, 1080, 1920), (5, 134, 240, 1072, 1920), (2, 34, 60, 272, 480),
                                         (5, 17, 25, 136, 200), (3, 30, 40, 64, 96), (5, 55, 55, 433, 433),
                                         (7, 12, 16, 96, 128), (5, 64, 96, 64, 96)])
@pytest.mark.parametrize("n", [2, 5])
def test_block_lowres_interval(cuda, C, hl, wl, H, W, n):
    """Block-grid route with the key frames at decoder resolution: first chain step and frame 0
    evaluate F.interpolate(bilinear, align_corners=True) on the fly.  Labels, logits and counts bit-exact against
    F.interpolate + the reference sequence on torch-CUDA.  433x433 / C=7 / a too fine key frame take the wrapper's
    three-launch route, (5,64,96,64,96) forwards to the plain entry."""
    g = torch.Generator().manual_seed(C * 1000 + hl + n)
    o_lr = (torch.randn(1, C, hl, wl, generator=g) * 3).to(cuda)
    o_next_lr = (torch.randn(1, C, hl, wl, generator=g) * 3).to(cuda)
    up = (lambda t: F.interpolate(t, size=(H, W), mode="bilinear", align_corners=True)) if (hl, wl) != (H, W) else (lambda t: t)
    gl = [x.to(cuda) for x in flow_grids(H, W, n, "block", clip=7, side=0)]
    gr = [x.to(cuda) for x in flow_grids(H, W, n, "block", clip=7, side=1)]
    ref_logits = fo.predict_segmentation(ident, ident, up(o_lr), up(o_next_lr), gl, gr, n, no_warp=False)
    ref_labels = fo.argmax_labels(ref_logits)
    tc_prev = torch.randint(0, C, (H, W), generator=torch.Generator().manual_seed(5), dtype=torch.uint8)
    counts = kernels.new_counts(C, cuda)
    labels, logits = kernels.block_lowres_interval(o_lr, o_next_lr, (H, W), gl, gr, n, want_labels=True, want_logits=True,
                                                   tc_prev=tc_prev.to(cuda), counts=counts)
    bad = int((logits.view(torch.int32) != ref_logits.view(torch.int32)).sum())
    assert bad == 0, f"{bad} logits differ from F.interpolate + the reference sequence"
    assert torch.equal(labels.long(), ref_labels)
    ref_counts, _ = oracle_temporal(ref_labels, C, tc_prev.numpy().astype(np.int64))
    assert np.array_equal(counts.cpu().numpy(), ref_counts)
    labels2, none = kernels.block_lowres_interval(o_lr, o_next_lr, (H, W), gl, gr, n)
    assert none is None and torch.equal(labels2, labels)


@pytest.mark.parametrize("C,hl,wl,H,W,n", [(5, 135, 240, 1080, 1920, 5), (5, 34, 60, 272, 480, 5), (5, 17, 32, 136, 256, 3),
                                           (5, 17, 30, 136, 240, 4), (2, 17, 32, 136, 256, 5), (5, 20, 25, 77, 101, 3),
                                           (7, 12, 16, 96, 128, 3), (5, 17, 32, 136, 256, 1), (5, 64, 96, 64, 96, 3)])
def test_dense_lowres_interval(cuda, C, hl, wl, H, W, n):
    """Dense route with the key frames at decoder resolution: the up-sample runs inside the entry
    and, for C = 5 with odd n, writes the strip kernel's 4+1 layout that step 1 then reads.  Labels, logits and counts
    bit-exact against F.interpolate + the reference sequence on torch-CUDA; even n, C != 5, W % 4 != 0 and n = 1 take
    the planar route; the second interval of a clip reuses the up-sample of its `prev` (the first one's `next`)."""
    g = torch.Generator().manual_seed(C * 1000 + hl + n)
    keys_lr = [(torch.randn(1, C, hl, wl, generator=g) * 3).to(cuda) for _ in range(3)]
    up = lambda t: F.interpolate(t, size=(H, W), mode="bilinear", align_corners=True)      # noqa: E731
    ups = kernels.KeyFrameUps()
    scratch = kernels.ScratchCache()
    tc_prev = torch.randint(0, C, (H, W), generator=torch.Generator().manual_seed(5), dtype=torch.uint8).to(cuda)
    for it in range(2):
        o_lr, o_next_lr = keys_lr[it], keys_lr[it + 1]
        gl = [x.to(cuda) for x in flow_grids(H, W, n, "dense", clip=7 + it, side=0)]
        gr = [x.to(cuda) for x in flow_grids(H, W, n, "dense", clip=7 + it, side=1)]
        if n > 1:
            ref_logits = fo.predict_segmentation(ident, ident, up(o_lr), up(o_next_lr), gl, gr, n, no_warp=False)
        else:
            ref_logits = up(o_lr)
        ref_labels = fo.argmax_labels(ref_logits)
        counts = kernels.new_counts(C, cuda)
        labels, logits = kernels.dense_lowres_interval(o_lr, o_next_lr if n > 1 else None, (H, W), gl, gr, n, want_labels=True,
                                                       want_logits=True, tc_prev=tc_prev, counts=counts, ups=ups,
                                                       scratch=scratch)
        bad = int((logits.reshape(-1).view(torch.int32) != ref_logits.reshape(-1).view(torch.int32)).sum())
        assert bad == 0, f"interval {it}: {bad} logits differ from F.interpolate + the reference sequence"
        assert torch.equal(labels.long().reshape(-1), ref_labels.reshape(-1))
        ref_counts, _ = oracle_temporal(ref_labels.reshape(n, H, W), C, tc_prev.cpu().numpy().astype(np.int64))
        assert np.array_equal(counts.cpu().numpy(), ref_counts)
        if n > 1:
            assert ups.tags[0] is not None and ups.tags[1] is not None
            if it == 1:      # the first interval's `next` was found again as this interval's `prev`
                assert any(t[0] is o_lr for t in ups.tags)
        tc_prev = labels[n - 1]
    # a key frame changed in place is up-sampled again (version counter), labels-only calls agree
    if n > 1:
        keys_lr[2].mul_(-1.0)
        gl = [x.to(cuda) for x in flow_grids(H, W, n, "dense", clip=3, side=0)]
        gr = [x.to(cuda) for x in flow_grids(H, W, n, "dense", clip=3, side=1)]
        ref = fo.argmax_labels(fo.predict_segmentation(ident, ident, up(keys_lr[2]), up(keys_lr[0]), gl, gr, n, no_warp=False))
        labels, none = kernels.dense_lowres_interval(keys_lr[2], keys_lr[0], (H, W), gl, gr, n, ups=ups, scratch=scratch)
        assert none is None and torch.equal(labels.long().reshape(-1), ref.reshape(-1))


@pytest.mark.parametrize("jitter", [0.3, 1.5])
def test_dense_lowres_interval_large_motion(cuda, jitter):
    """Decoder-resolution key frames with taps far outside the strip kernel's window (and outside the image): step 1's
    warps then gather from the 4+1 key frames in global memory (block_from_global with both layouts 4+1)."""
    C, hl, wl, H, W, n = 5, 34, 64, 272, 512, 5
    g = torch.Generator().manual_seed(31)
    o_lr, o_next_lr = ((torch.randn(1, C, hl, wl, generator=g) * 3).to(cuda) for _ in range(2))
    up = lambda t: F.interpolate(t, size=(H, W), mode="bilinear", align_corners=True)      # noqa: E731
    gl = [x.to(cuda) for x in flow_grids(H, W, n, "dense", clip=12, side=0, jitter=jitter)]
    gr = [x.to(cuda) for x in flow_grids(H, W, n, "dense", clip=12, side=1, jitter=jitter)]
    ref_logits = fo.predict_segmentation(ident, ident, up(o_lr), up(o_next_lr), gl, gr, n, no_warp=False)
    labels, logits = kernels.dense_lowres_interval(o_lr, o_next_lr, (H, W), gl, gr, n, want_labels=True, want_logits=True)
    bad = int((logits.reshape(-1).view(torch.int32) != ref_logits.reshape(-1).view(torch.int32)).sum())
    assert bad == 0, f"{bad} logits differ from F.interpolate + the reference sequence"
    assert torch.equal(labels.long().reshape(-1), fo.argmax_labels(ref_logits).reshape(-1))


@pytest.mark.parametrize("C,H,W,n", [(5, 272, 512, 5), (5, 96, 132, 3), (2, 64, 128, 4)])
def test_dense_interval_nan_and_inf(cuda, C, H, W, n):
    """Key frames with NaN, +Inf and -Inf class values through the dense route (strip kernel with 4+1 and planar states,
    the direct kernel for the even-n middle step): taps with weight zero at the border must not pull a NaN in, a NaN
    class value wins the arg-max like in torch.max, Inf - Inf in a blend makes a NaN on the way."""
    o, o_next = keyframe_logits(C, H, W, 8, 0)[None], keyframe_logits(C, H, W, 8, 1)[None]
    for k, (t, v) in enumerate(((o, float("nan")), (o_next, float("inf")), (o, float("-inf")), (o_next, float("nan")),
                                (o, float("inf")), (o_next, float("-inf")))):
        t[0, k % C, (2 * k + 1)::7, (3 * k)::5] = v
    o[0, :, 0, :] = float("inf")                # image border rows / columns: the clipped taps
    o_next[0, :, :, W - 1] = float("nan")
    gl = [g.to(cuda) for g in flow_grids(H, W, n, "dense", clip=8, side=0, jitter=0.05)]
    gr = [g.to(cuda) for g in flow_grids(H, W, n, "dense", clip=8, side=1, jitter=0.05)]
    oc, onc = o.to(cuda), o_next.to(cuda)
    ref_logits = fo.predict_segmentation(ident, ident, oc, onc, gl, gr, n, False)
    labels, logits = kernels.dense_interval(oc, onc, gl, gr, n, want_labels=True, want_logits=True)
    assert bool(torch.isnan(ref_logits).any()) and bool(torch.isinf(ref_logits).any())
    a, b = logits.reshape(-1), ref_logits.reshape(-1)
    bad = int(((a.view(torch.int32) != b.view(torch.int32)) & ~(torch.isnan(a) & torch.isnan(b))).sum())     # any NaN equals any NaN
    assert bad == 0, f"{bad} logits differ from torch-CUDA oracle"
    assert torch.equal(labels.long().reshape(-1), fo.argmax_labels(ref_logits).reshape(-1))


def test_comm_single_rank_roundtrip(cuda):
    """fuvs_comm_* / fuvs_allreduce_counts with a one-rank communicator (the multi-rank sum is checked on hardware by
    bench.py --gpus N: `allreduce_parity`): unique id, init, an all-reduce that must leave the counts unchanged, destroy;
    without a communicator the entry refuses loudly."""
    import ctypes

    from flood_uav_video_segmentation_b200._lib import FuvsError, check, load, ptr, stream_ptr
    lib = load()
    counts = torch.arange(15, dtype=torch.int64, device=cuda).reshape(3, 5) * 1000003
    with pytest.raises(FuvsError):
        check(lib.fuvs_allreduce_counts(ptr(counts), counts.numel(), stream_ptr(cuda)))
    assert lib.fuvs_comm_world_size() == 0
    buf = (ctypes.c_ubyte * 128)()
    check(lib.fuvs_comm_unique_id(buf))
    assert any(buf)
    with torch.cuda.device(cuda):
        check(lib.fuvs_comm_init(buf, 0, 1))
        try:
            assert lib.fuvs_comm_world_size() == 1
            with pytest.raises(FuvsError):
                check(lib.fuvs_comm_init(buf, 0, 1))                      # one communicator per process
            before = counts.clone()
            check(lib.fuvs_allreduce_counts(ptr(counts), counts.numel(), stream_ptr(cuda)))
            torch.cuda.synchronize()
            assert torch.equal(counts, before)
        finally:
            check(lib.fuvs_comm_destroy())
    assert lib.fuvs_comm_world_size() == 0
