"""Drop-in surface (FlowModel / FlowBaseModel / compute_metrics) against the oracle's restatement of the reference
call sequence, with a small seeded encoder/decoder standing in for the (out-of-scope) key-frame network."""
import numpy as np
import pytest
import torch
from torch import nn

from flood_uav_video_segmentation_b200 import kernels
from flood_uav_video_segmentation_b200.base.foundation import compute_metrics, epoch_metrics
from flood_uav_video_segmentation_b200.flow.base import FlowBaseModel, SimpleProfiler
from flood_uav_video_segmentation_b200.flow.model import FlowModel
from flood_uav_video_segmentation_b200.synthetic import flow_grids, gt_labels
from oracle import flow_oracle as fo
from oracle import metric_oracle as mo

pytestmark = pytest.mark.gpu


class TinyBackbone(nn.Module):
    """encoder: stride-8 conv features; decoder: 1x1 conv to class logits (shape class of PSPNet/DeepLab heads)."""

    def __init__(self, classes=5, feat=24, stride=8):
        super().__init__()
        torch.manual_seed(0)
        self.encoder = nn.Sequential(nn.Conv2d(3, feat, 3, stride=stride, padding=1), nn.ReLU())
        self.decoder = nn.Conv2d(feat, classes, 1)


def bits_equal(a, b):
    return a.shape == b.shape and bool((a.contiguous().view(torch.int32) == b.contiguous().view(torch.int32)).all())


def frames(H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, 3, H, W, generator=g), torch.randn(1, 3, H, W, generator=g)


@pytest.mark.parametrize("feature_based", [False, True])
@pytest.mark.parametrize("mode", ["linear", "block", "dense"])
def test_predict_matches_oracle(cuda, feature_based, mode):
    H, W, n = 144, 208, 5
    bb = TinyBackbone().to(cuda).eval()
    prev, nxt = (t.to(cuda) for t in frames(H, W, 1))
    no_warp = mode == "linear"
    if no_warp:
        gl = gr = [torch.zeros(1, 1, device=cuda)] * (n - 1)
    else:
        gh, gw = (H, W)
        gl = [g.to(cuda) for g in flow_grids(gh, gw, n, mode, clip=7, side=0)]
        gr = [g.to(cuda) for g in flow_grids(gh, gw, n, mode, clip=7, side=1)]
    fm = FlowModel(bb, feature_based=feature_based, no_warp=no_warp).eval()
    prof = fo.NullProfiler()
    with torch.no_grad():
        got = fm.predict(prev, nxt, gl, gr, n, prof)["pred"]
        if feature_based:
            ref = fo.predict_feature(bb.encoder, bb.decoder, prev, nxt, gl, gr, n, fm.default_motion_vector, no_warp)
        else:
            ref = fo.predict_segmentation(bb.encoder, bb.decoder, prev, nxt, gl, gr, n, no_warp)
    assert got.shape == ref.shape == (n, 5, H, W)
    if feature_based:
        # blended features are bit-equal; the decoder conv then sees batch n instead of... the same batch n: equal too
        assert bits_equal(got, ref)
    else:
        assert bits_equal(got, ref)
    assert set(prof.regions) >= {"predict_encoder", "predict_decoder"}
    if not feature_based:
        with torch.no_grad():
            labels = fm.predict_labels(prev, nxt, gl, gr, n, prof)
        assert torch.equal(labels.long(), ref.max(1)[1])


@pytest.mark.parametrize("feature_based", [False, True])
@pytest.mark.parametrize("no_warp", [False, True])
def test_forward_matches_oracle(cuda, feature_based, no_warp):
    """validation/test route: FlowModel.forward with per-sample left/right indices (flow/model.py:35-106)."""
    H, W, k, B = 144, 208, 5, 2
    bb = TinyBackbone().to(cuda).eval()
    g = torch.Generator().manual_seed(3)
    prev, nxt = torch.randn(B, 3, H, W, generator=g).to(cuda), torch.randn(B, 3, H, W, generator=g).to(cuda)
    left, right = torch.tensor([2, 1]), torch.tensor([3, 4])
    if no_warp:
        gl = gr = [torch.zeros(B, 1, device=cuda)] * (k - 1)
    else:
        def batched(side):
            per = [flow_grids(H, W, k, "block", clip=b, side=side) for b in range(B)]
            return [torch.cat([per[b][j] for b in range(B)], 0).to(cuda) for j in range(k - 1)]
        gl, gr = batched(0), batched(1)
    fm = FlowModel(bb, feature_based=feature_based, no_warp=no_warp).eval()
    with torch.no_grad():
        got = fm(None, prev, nxt, gl, gr, left, right)["pred"]
        f = fo.forward_feature if feature_based else fo.forward_segmentation
        ref = f(bb.encoder, bb.decoder, prev, nxt, gl, gr, left, right, no_warp)
    assert bits_equal(got, ref)


def test_warp_and_warp_batch_methods(cuda):
    bb = TinyBackbone().to(cuda).eval()
    fm = FlowModel(bb, feature_based=False, no_warp=False).eval()
    x = torch.randn(2, 5, 64, 96, generator=torch.Generator().manual_seed(4)).to(cuda)
    gs = [torch.cat([flow_grids(64, 96, 3, "block", clip=b)[j] for b in range(2)], 0).to(cuda) for j in range(2)]
    with torch.no_grad():
        assert bits_equal(fm.warp(x, gs[0]), fo.warp(x, gs[0]))
        assert bits_equal(fm.warp_batch(x, gs, [2, 1], [3, 3]), fo.warp_batch(x, gs, [2, 1], [3, 3]))


def test_training_route_is_differentiable_and_cpu_inference_raises(cuda):
    bb = TinyBackbone().to(cuda).train()
    fm = FlowModel(bb, feature_based=False, no_warp=False).train()
    prev, nxt = (t.to(cuda) for t in frames(64, 96, 2))
    gl = [g.to(cuda) for g in flow_grids(64, 96, 3, "block", side=0)]
    gr = [g.to(cuda) for g in flow_grids(64, 96, 3, "block", side=1)]
    out = fm(None, prev, nxt, gl, gr, [1], [2])["pred"]
    out.sum().backward()
    assert bb.decoder.weight.grad is not None
    cpu_model = FlowModel(TinyBackbone().eval(), feature_based=False, no_warp=True).eval()
    with torch.no_grad(), pytest.raises(kernels.FuvsError):
        cpu_model.predict(*frames(32, 32, 0), [torch.zeros(1, 1)], [torch.zeros(1, 1)], 2, fo.NullProfiler())


def test_predict_step_temporal_metrics(cuda):
    """Three consecutive intervals through FlowBaseModel.predict_step == reference loop (flow/base.py:259-343)."""
    H, W, n, C = 96, 128, 5, 5
    bb = TinyBackbone(classes=C).to(cuda).eval()
    m = FlowBaseModel(classes=C, arch="pspnet", feature_based=False, no_warp=False, no_cropping=True,
                      backbone=bb, output_size=(H, W)).to(cuda).eval()
    m.on_predict_start()
    tot = [np.zeros(C, np.int64) for _ in range(3)]
    last = None
    g = torch.Generator().manual_seed(8)
    keys = [torch.randn(1, 3, H, W, generator=g).to(cuda) for _ in range(4)]
    for it in range(3):
        gl = [x.to(cuda) for x in flow_grids(H, W, n, "block", clip=9, interval=it, side=0)]
        gr = [x.to(cuda) for x in flow_grids(H, W, n, "block", clip=9, interval=it, side=1)]
        out = m.predict_step({"frame_prev": keys[it], "frame_next": keys[it + 1], "mvs_left": gl, "mvs_right": gr,
                              "frame_id": torch.tensor([it * n])}, it)
        with torch.no_grad():
            ref = fo.predict_segmentation(bb.encoder, bb.decoder, keys[it], keys[it + 1], gl, gr, n).max(1)[1]
        assert torch.equal(out.long(), ref)
        (i, u, t), last = mo.temporal_consistency_counts(ref.cpu().numpy(), C, 255, last)
        for acc, v in zip(tot, (i, u, t)):
            acc += v
    res = m.on_predict_end()
    assert np.array_equal(m.intersection_meter_predict.sum, tot[0])
    assert np.array_equal(m.union_meter_predict.sum, tot[1])
    assert np.array_equal(m.target_meter_predict.sum, tot[2])
    e = mo.epoch_metrics(*tot)
    assert res["predict_miou1_epoch"] == e["miou"] and res["predict_macc1_epoch"] == e["macc"]
    assert res["predict_accuracy1_epoch"] == e["accuracy"]


def test_validation_step_and_compute_metrics(cuda):
    H, W, C, k = 96, 128, 5, 5
    bb = TinyBackbone(classes=C).to(cuda).eval()
    m = FlowBaseModel(classes=C, arch="pspnet", feature_based=False, no_warp=False, no_cropping=True, backbone=bb).to(cuda).eval()
    tot = np.zeros((3, C), np.int64)
    for s in range(2):
        prev, nxt = (t.to(cuda) for t in frames(H, W, 20 + s))
        gl = [x.to(cuda) for x in flow_grids(H, W, k, "block", clip=s, side=0)]
        gr = [x.to(cuda) for x in flow_grids(H, W, k, "block", clip=s, side=1)]
        label = gt_labels(H, W, C, seed=s)[None].to(cuda)
        batch = {"frame_prev": prev, "frame_next": nxt, "mvs_left": gl, "mvs_right": gr,
                 "left_index": torch.tensor([2]), "right_index": torch.tensor([3]), "label": label}
        m.validation_step(batch, s)
        with torch.no_grad():
            ref = fo.forward_segmentation(bb.encoder, bb.decoder, prev, nxt, gl, gr, [2], [3]).max(1)[1]
        tot += np.stack(mo.intersection_and_union_histc_ints(ref.cpu().numpy(), label.cpu().numpy(), C, 255))
        # compute_metrics drop-in (base/foundation.py:333-344), both conventions
        got = compute_metrics(ref.clone(), label, C, 255)
        assert np.array_equal(np.stack(got), np.stack(mo.intersection_and_union_histc_ints(ref.cpu().numpy(), label.cpu().numpy(), C, 255)))
        got = compute_metrics(ref.cpu().numpy(), label.cpu().numpy(), C, 255)
        assert np.array_equal(np.stack(got), np.stack(mo.intersection_and_union_np(ref.cpu().numpy(), label.cpu().numpy(), C, 255)))
    res = m.validation_epoch_metrics()
    assert np.array_equal(m.intersection_meter_val.sum, tot[0]) and np.array_equal(m.union_meter_val.sum, tot[1])
    e = mo.epoch_metrics(*tot)
    assert res["miou"] == e["miou"] and res["macc"] == e["macc"] and res["accuracy"] == e["accuracy"]
    assert epoch_metrics(*tot)["miou"] == e["miou"]


class DeepLabParts(nn.Module):
    """torchvision DeepLabV3 split like the reference's FlowDeepLabv3 (model/deeplabv3.py:47-54): encoder = ResNet
    backbone -> 2048-channel stride-8 features, decoder = DeepLabHead (ASPP).  Random init, no network access."""

    def __init__(self, classes=5):
        super().__init__()
        from torchvision.models.segmentation import deeplabv3_resnet50
        torch.manual_seed(0)
        m = deeplabv3_resnet50(weights=None, weights_backbone=None, aux_loss=False, num_classes=classes)
        self.backbone = m.backbone
        self.decoder = m.classifier
        self.encoder = _OutOnly(self.backbone)


class _OutOnly(nn.Module):
    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, x):
        return self.m(x)["out"]


@pytest.mark.parametrize("feature_based", [False, True])
@pytest.mark.parametrize("mode", ["linear", "block"])
def test_deeplabv3_keyframes(cuda, feature_based, mode):
    """BASELINE.json config 3 shape class: DeepLabV3 key frames (2048-channel features at stride 8), random-init
    weights, segmentation- and feature-based interpolation, against the oracle with the same modules."""
    H, W, n = 193, 257, 4
    bb = DeepLabParts().to(cuda).eval()
    prev, nxt = (t.to(cuda) for t in frames(H, W, 5))
    no_warp = mode == "linear"
    if no_warp:
        gl = gr = [torch.zeros(1, 1, device=cuda)] * (n - 1)
    else:
        gl = [g.to(cuda) for g in flow_grids(H, W, n, "block", clip=11, side=0)]
        gr = [g.to(cuda) for g in flow_grids(H, W, n, "block", clip=11, side=1)]
    fm = FlowModel(bb, feature_based=feature_based, no_warp=no_warp).eval()
    with torch.no_grad():
        got = fm.predict(prev, nxt, gl, gr, n, fo.NullProfiler())["pred"]
        if feature_based:
            ref = fo.predict_feature(bb.encoder, bb.decoder, prev, nxt, gl, gr, n, fm.default_motion_vector, no_warp)
        else:
            ref = fo.predict_segmentation(bb.encoder, bb.decoder, prev, nxt, gl, gr, n, no_warp)
    assert got.shape == ref.shape == (n, 5, H, W)
    assert bits_equal(got, ref)
    assert torch.equal(kernels.argmax(got).long(), ref.max(1)[1])


@pytest.mark.parametrize("kind", ["block", "dense"])
def test_keyframe_reuse_gives_identical_labels(cuda, kind):
    """Caching the `next` key frame's logits for the following interval halves the backbone
    calls and must not change a single label or count.  On the dense route the cached decoder output is the very tensor
    the next interval passes as `prev`, so its up-sample (kernels.KeyFrameUps) is reused as well."""
    H, W, n, C = 96, 128, 5, 5
    bb = TinyBackbone(classes=C).to(cuda).eval()
    calls = {"n": 0}
    bb.encoder.register_forward_hook(lambda *a: calls.__setitem__("n", calls["n"] + 1))
    g = torch.Generator().manual_seed(8)
    keys = [torch.randn(1, 3, H, W, generator=g).to(cuda) for _ in range(4)]
    results = {}
    for reuse in (False, True, "inference_mode"):
        m = FlowBaseModel(classes=C, arch="pspnet", feature_based=False, no_warp=False, no_cropping=True, backbone=bb,
                          output_size=(H, W), reuse_keyframes=bool(reuse)).to(cuda).eval()
        m.on_predict_start()
        calls["n"] = 0
        outs = []
        # Lightning's predict loop runs under torch.inference_mode: its tensors carry no version counter
        ctx = torch.inference_mode() if reuse == "inference_mode" else torch.no_grad()
        with ctx:
          for it in range(3):
            gl = [x.to(cuda) for x in flow_grids(H, W, n, kind, clip=9, interval=it, side=0)]
            gr = [x.to(cuda) for x in flow_grids(H, W, n, kind, clip=9, interval=it, side=1)]
            outs.append(m.predict_step({"frame_prev": keys[it], "frame_next": keys[it + 1], "mvs_left": gl,
                                        "mvs_right": gr, "frame_id": torch.tensor([it * n])}, it).clone())
            if kind == "dense" and reuse and it > 0:      # this interval's prev was found in the up-sample buffers
                assert all(t is not None for t in m.model_G._dense_ups.tags)
        m.on_predict_end()
        results[reuse] = (outs, m.intersection_meter_predict.sum.copy(), calls["n"])
    assert results[False][2] == 6 and results[True][2] == 4 and results["inference_mode"][2] == 4
    for other in (True, "inference_mode"):
        for a, b in zip(results[False][0], results[other][0]):
            assert torch.equal(a, b)
        assert np.array_equal(results[False][1], results[other][1])


def test_keyframe_cache_is_invalidated_by_weight_and_mode_changes(cuda):
    """The opt-in key-frame cache must never hand back logits of other weights, another mode, another device or
    another size: every such change bumps the cache version."""
    H, W, n, C = 64, 96, 3, 5
    bb = TinyBackbone(classes=C).to(cuda).eval()
    fm = FlowModel(bb, feature_based=False, no_warp=True).to(cuda).eval()
    fm.reuse_keyframes = True
    prof = SimpleProfiler()
    x0, x1 = torch.randn(1, 3, H, W, device=cuda), torch.randn(1, 3, H, W, device=cuda)
    dummy = [torch.zeros(1, 1, device=cuda)] * (n - 1)
    fm.predict_labels(x0, x1, dummy, dummy, n, prof, frame_id=0)
    assert fm._kf_cache is not None and fm._cached_keyframe(n, (H, W), x0.device, True) is not None
    assert fm._cached_keyframe(n, (H + 8, W), x0.device, True) is None          # another size
    assert fm._cached_keyframe(n, (H, W), x0.device, False) is None             # full-resolution vs decoder-resolution
    for change in (lambda: fm.train(), lambda: fm.eval(), lambda: fm.load_state_dict(fm.state_dict()),
                   lambda: fm.float(), lambda: fm.reset_keyframe_cache()):
        fm.eval()
        fm.predict_labels(x0, x1, dummy, dummy, n, prof, frame_id=0)
        assert fm._cached_keyframe(n, (H, W), x0.device, True) is not None
        change()
        assert fm._cached_keyframe(n, (H, W), x0.device, True) is None
    # and with new weights the second interval really uses them
    fm.eval()
    a = fm.predict_labels(x1, x0, dummy, dummy, n, prof, frame_id=n)
    with torch.no_grad():
        for p in bb.parameters():
            p.mul_(-1.0)
    fm.load_state_dict(fm.state_dict())
    b = fm.predict_labels(x1, x0, dummy, dummy, n, prof, frame_id=n)
    fm.reuse_keyframes = False
    c = fm.predict_labels(x1, x0, dummy, dummy, n, prof)
    assert torch.equal(b, c) and not torch.equal(a, b)


@pytest.mark.parametrize("feature_based,no_warp,no_cropping", [(False, False, True), (False, True, True), (True, False, True),
                                                               (False, False, False)])
def test_predict_step_resizes_to_the_output_size(cuda, feature_based, no_warp, no_cropping):
    """Frames that are NOT at the hard-coded (1072, 1920) of flow/base.py:275: predict_step's
    F.interpolate(output, size, bilinear, align_corners=True) + max(1)[1] + uint8 (one kernel, fuvs_upsample_argmax; for
    the sliding-crop route the fp64 canvas goes through torch's interpolate as in the reference) and the temporal metric
    on the RESIZED label maps, against the restated reference sequence on torch-CUDA."""
    import torch.nn.functional as F
    H, W, n, C = 80, 112, 3, 5
    out_size = (96, 160)
    bb = TinyBackbone(classes=C).to(cuda).eval()
    m = FlowBaseModel(classes=C, arch="pspnet", feature_based=feature_based, no_warp=no_warp, no_cropping=no_cropping,
                      backbone=bb, output_size=out_size, test_h=49, test_w=65, save_video=False).to(cuda).eval()
    m.on_predict_start()
    g = torch.Generator().manual_seed(12)
    keys = [torch.randn(1, 3, H, W, generator=g).to(cuda) for _ in range(3)]
    tot = [np.zeros(C, np.int64) for _ in range(3)]
    last = None
    for it in range(2):
        if no_warp:
            gl = gr = [torch.zeros(1, 1, device=cuda)] * (n - 1)
        else:
            gl = [x.to(cuda) for x in flow_grids(H, W, n, "block", clip=13, interval=it, side=0)]
            gr = [x.to(cuda) for x in flow_grids(H, W, n, "block", clip=13, interval=it, side=1)]
        out = m.predict_step({"frame_prev": keys[it], "frame_next": keys[it + 1], "mvs_left": gl, "mvs_right": gr}, it)
        assert tuple(out.shape) == (n,) + out_size and out.dtype == torch.uint8
        if not no_cropping:
            continue              # reference canvas: covered bit-exactly at canvas size by tests/test_crop_gpu.py
        with torch.no_grad():
            if feature_based:
                logits = fo.predict_feature(bb.encoder, bb.decoder, keys[it], keys[it + 1], gl, gr, n,
                                            m.model_G.default_motion_vector, no_warp=no_warp)
            else:
                logits = fo.predict_segmentation(bb.encoder, bb.decoder, keys[it], keys[it + 1], gl, gr, n, no_warp=no_warp)
            ref = F.interpolate(logits, out_size, mode="bilinear", align_corners=True).max(1)[1]      # flow/base.py:275-276
        assert torch.equal(out.long(), ref)
        (i, u, t), last = mo.temporal_consistency_counts(ref.cpu().numpy(), C, 255, last)
        for acc, v in zip(tot, (i, u, t)):
            acc += v
    m.on_predict_end()
    if no_cropping:
        assert np.array_equal(m.intersection_meter_predict.sum, tot[0])
        assert np.array_equal(m.union_meter_predict.sum, tot[1])
        assert np.array_equal(m.target_meter_predict.sum, tot[2])
