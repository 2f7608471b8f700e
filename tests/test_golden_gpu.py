"""GPU tier against the REFERENCE-GENERATED fixtures (tests/golden/*.npz, written by oracle/make_golden.py from the
imported reference on torch-CPU): the fixtures are fed straight into libfuvs through the C ABI, without the oracle in
between.

Tolerances.  Linear mode and the metric are bit-exact against the reference on any device.  Warp modes: torch's CPU
and CUDA grid_sample kernels differ at the 1e-5 level (different unnormalisation / accumulation order),
the kernels reproduce torch-CUDA bit for bit, so against these CPU-generated vectors logits are compared with
atol 1e-4 and labels wherever the reference's own top-2 margin exceeds 1e-3 (north_star's 1e-5 relative holds against
torch-CUDA, see test_kernels_gpu.py)."""
import glob
import os

import numpy as np
import pytest
import torch
from torch import nn

from flood_uav_video_segmentation_b200 import kernels
from flood_uav_video_segmentation_b200.flow.base import SimpleProfiler
from flood_uav_video_segmentation_b200.flow.model import FlowModel, get_default_grid
from flood_uav_video_segmentation_b200.util.util import intersectionAndUnion, intersectionAndUnionGPU

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INTERVALS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*_n[0-9].npz")))
FORWARDS = ["forward_seg_warp", "forward_seg_nowarp", "forward_feat_warp", "forward_feat_nowarp"]
ATOL_CPU_VS_CUDA = 1e-4
MARGIN = 1e-3


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


class TinyBackbone(nn.Module):
    """Same seeded conv backbone as oracle/make_golden.py."""

    def __init__(self, classes=5, feat=12, stride=8):
        super().__init__()
        torch.manual_seed(0)
        self.encoder = nn.Sequential(nn.Conv2d(3, feat, 3, stride=stride, padding=1), nn.ReLU())
        self.decoder = nn.Conv2d(feat, classes, 1)


def test_inventory():
    assert len(INTERVALS) == 6, INTERVALS


@pytest.mark.parametrize("name", INTERVALS)
def test_interval_entries_vs_reference_golden(cuda, name):
    """fuvs_linear_blend_argmax / fuvs_block_interval / fuvs_dense_interval on the reference's own inputs."""
    g = load(name)
    n, mode = int(g["n"]), str(g["mode"])
    o, o_next = torch.from_numpy(g["prev"]).to(cuda), torch.from_numpy(g["next"]).to(cuda)
    C = o.shape[1]
    counts = kernels.new_counts(C, cuda)
    kw = dict(want_labels=True, want_logits=True, counts=counts)
    if mode == "linear":
        labels, logits = kernels.linear_blend_argmax(o, o_next, n, **kw)
    else:
        gl = [torch.from_numpy(x).to(cuda) for x in g["grids_left"]]
        gr = [torch.from_numpy(x).to(cuda) for x in g["grids_right"]]
        fn = kernels.dense_interval if mode == "dense" else kernels.block_interval
        labels, logits = fn(o, o_next, gl, gr, n, **kw)
    lab, ref_lab = labels.cpu().numpy(), g["labels"]
    if mode == "linear":
        assert np.array_equal(logits.cpu().numpy().view(np.int32), g["pred"].view(np.int32)), "logits not bit-equal"
        assert np.array_equal(lab, ref_lab)
        assert np.array_equal(counts.cpu().numpy(), g["counts"])
    else:
        np.testing.assert_allclose(logits.cpu().numpy(), g["pred"], rtol=0, atol=ATOL_CPU_VS_CUDA)
        sure = g["margin"] > MARGIN
        assert sure.mean() > 0.9
        assert np.array_equal(lab[sure], ref_lab[sure])
        # the counts follow the labels: recount on the device from the reference's own labels and compare with the fixture
        c2 = kernels.temporal_counts(torch.from_numpy(ref_lab).to(cuda), C)
        assert np.array_equal(c2.cpu().numpy(), g["counts"])


@pytest.mark.parametrize("name", FORWARDS)
def test_flowmodel_vs_reference_golden(cuda, name):
    """FlowModel.forward / FlowModel.predict (drop-in classes) on the fixtures of the reference's FlowModel with the
    same seeded conv backbone.  The conv layers run in cuDNN on the device, so even the no-warp cases carry the
    convolution's CPU/CUDA difference: atol 1e-4 throughout, bit-exactness is covered by the interval test above."""
    g = load(name)
    fb, nw, k = bool(g["feature_based"]), bool(g["no_warp"]), int(g["k"])
    bb = TinyBackbone().to(cuda).eval()
    m = FlowModel(bb, feature_based=fb, no_warp=nw).to(cuda).eval()
    prev, nxt = torch.from_numpy(g["prev"]).to(cuda), torch.from_numpy(g["next"]).to(cuda)
    B = prev.shape[0]
    if nw:
        gl = gr = [torch.zeros(B, 1, device=cuda)] * (k - 1)
    else:
        gl = [torch.from_numpy(x).to(cuda) for x in g["grids_left"]]
        gr = [torch.from_numpy(x).to(cuda) for x in g["grids_right"]]
    with torch.no_grad():
        prev_tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            fwd = m(None, prev, nxt, gl, gr, torch.from_numpy(g["left"]), torch.from_numpy(g["right"]))["pred"]
            pred = m.predict(prev[:1], nxt[:1], [x[:1] for x in gl], [x[:1] for x in gr], k, SimpleProfiler())["pred"]
        finally:
            torch.backends.cudnn.allow_tf32 = prev_tf32
    np.testing.assert_allclose(fwd.cpu().numpy(), g["forward"], rtol=0, atol=ATOL_CPU_VS_CUDA)
    np.testing.assert_allclose(pred.cpu().numpy(), g["predict"], rtol=0, atol=ATOL_CPU_VS_CUDA)


def _metric_keys():
    g = load("metric_cases")
    return sorted(k[:-5] for k in g.files if k.endswith("_pred"))


@pytest.mark.parametrize("key", _metric_keys())
def test_confusion_vs_reference_golden(cuda, key):
    """fuvs_confusion on the reference's metric fixtures (incl. the florida-01 label masks): the numpy convention
    (util/util.py:36-47) bit-exact for every dtype combination, and the histc convention (util/util.py:52-63) wherever
    the two conventions agree (no label equal to K)."""
    g = load("metric_cases")
    pred, target, iut = g[key + "_pred"], g[key + "_target"], g[key + "_iut"]
    K = iut.shape[1]
    for pd in (torch.uint8, torch.int64):
        for td in (torch.uint8, torch.int64):
            p, t = torch.from_numpy(pred).to(cuda, pd), torch.from_numpy(target).to(cuda, td)
            c = kernels.confusion(p, t, K, 255, numpy_bins=True)
            assert np.array_equal(c.cpu().numpy(), iut), (key, pd, td)
    i, u, t = intersectionAndUnion(torch.from_numpy(pred).to(cuda), torch.from_numpy(target).to(cuda), K, 255)
    assert np.array_equal(np.stack([i, u, t]), iut)
    if not ((pred == K).any() or (target == K).any()):
        gi, gu, gt = intersectionAndUnionGPU(torch.from_numpy(pred).to(cuda, torch.int64),
                                             torch.from_numpy(target).to(cuda, torch.int64), K, 255)
        assert np.array_equal(np.stack([x.cpu().numpy() for x in (gi, gu, gt)]), iut)


def test_default_grid_vs_reference_golden(cuda):
    g = load("default_grid")
    assert np.array_equal(get_default_grid(), g["grid"])
    m = FlowModel(TinyBackbone(), feature_based=True)
    assert np.array_equal(m.default_motion_vector[0].numpy(), g["grid"].astype(np.float32))
