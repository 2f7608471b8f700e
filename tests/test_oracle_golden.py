"""CPU tier: the oracle (torch restatement, numpy restatement, plain-C restatement) against the golden vectors the
REFERENCE ITSELF produced (oracle/make_golden.py, run in the authoring container by importing /root/reference)."""
import glob
import os

import numpy as np
import pytest
import torch
from torch import nn

from oracle import c_oracle as co
from oracle import flow_oracle as fo
from oracle import metric_oracle as mo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INTERVALS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*_n[0-9].npz")))
ident = nn.Identity()
# torch-CPU's grid_sample (the golden vectors) and the CUDA algorithm (C oracle / kernels) differ by ~1e-5 abs per
# warp; label comparisons are made where the top-2 logit margin exceeds this by an order of magnitude
ATOL_CPU_VS_CUDA = 1e-4
MARGIN = 1e-3


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


class TinyBackbone(nn.Module):
    def __init__(self, classes=5, feat=12, stride=8):
        super().__init__()
        torch.manual_seed(0)
        self.encoder = nn.Sequential(nn.Conv2d(3, feat, 3, stride=stride, padding=1), nn.ReLU())
        self.decoder = nn.Conv2d(feat, classes, 1)


def test_fixture_inventory():
    assert len(INTERVALS) == 6, INTERVALS


@pytest.mark.parametrize("name", INTERVALS)
def test_torch_oracle_matches_reference_golden(name):
    g = load(name)
    n, mode = int(g["n"]), str(g["mode"])
    o, o_next = torch.from_numpy(g["prev"]), torch.from_numpy(g["next"])
    if mode == "linear":
        gl = gr = [torch.zeros(1, 1)] * (n - 1)
    else:
        gl = [torch.from_numpy(x) for x in g["grids_left"]]
        gr = [torch.from_numpy(x) for x in g["grids_right"]]
    with torch.no_grad():
        pred = fo.predict_segmentation(ident, ident, o, o_next, gl, gr, n, no_warp=(mode == "linear"))
    if mode == "linear":
        assert np.array_equal(pred.numpy().view(np.int32), g["pred"].view(np.int32))
        assert np.array_equal(fo.argmax_labels(pred).numpy().astype(np.uint8), g["labels"])
    else:
        np.testing.assert_allclose(pred.numpy(), g["pred"], rtol=0, atol=1e-6)   # same torch-CPU kernels
        sure = g["margin"] > MARGIN
        assert np.array_equal(fo.argmax_labels(pred).numpy().astype(np.uint8)[sure], g["labels"][sure])
    (i, u, t), _ = mo.temporal_consistency_counts(g["labels"].astype(np.int64), int(g["prev"].shape[1]), 255, None)
    assert np.array_equal(np.stack([i, u, t]), g["counts"])


@pytest.mark.parametrize("name", INTERVALS)
def test_c_oracle_matches_reference_golden(name):
    g = load(name)
    n, mode = int(g["n"]), str(g["mode"])
    logits, labels = co.interval(g["prev"][0], g["next"][0], g["grids_left"][:, 0] if mode != "linear" else None,
                                 g["grids_right"][:, 0] if mode != "linear" else None, n, warp=(mode != "linear"))
    if mode == "linear":
        assert np.array_equal(logits.view(np.int32), g["pred"].view(np.int32))      # bit-exact
        assert np.array_equal(labels, g["labels"])
        assert np.array_equal(co.temporal(labels, g["prev"].shape[1]), g["counts"])
    else:
        np.testing.assert_allclose(logits, g["pred"], rtol=1e-5, atol=ATOL_CPU_VS_CUDA)
        sure = g["margin"] > MARGIN
        assert sure.mean() > 0.99
        assert np.array_equal(labels[sure], g["labels"][sure])
        unsure_diff = int((labels != g["labels"]).sum())
        assert unsure_diff <= (~sure).sum()


@pytest.mark.parametrize("name", ["forward_seg_warp", "forward_seg_nowarp", "forward_feat_warp", "forward_feat_nowarp"])
def test_forward_and_predict_with_backbone(name):
    """FlowModel.forward (val/test route) and .predict (feature- and segmentation-based) with a small conv
    backbone; conv results may differ in the last bits across CPUs (oneDNN ISA dispatch) -> 1e-5 tolerance."""
    g = load(name)
    fb, nw, k = bool(g["feature_based"]), bool(g["no_warp"]), int(g["k"])
    bb = TinyBackbone().eval()
    prev, nxt = torch.from_numpy(g["prev"]), torch.from_numpy(g["next"])
    B = prev.shape[0]
    if nw:
        gl = gr = [torch.zeros(B, 1)] * (k - 1)
    else:
        gl = [torch.from_numpy(x) for x in g["grids_left"]]
        gr = [torch.from_numpy(x) for x in g["grids_right"]]
    with torch.no_grad():
        f = fo.forward_feature if fb else fo.forward_segmentation
        fwd = f(bb.encoder, bb.decoder, prev, nxt, gl, gr, g["left"], g["right"], nw)
        g1 = [x[:1] for x in gl]
        g2 = [x[:1] for x in gr]
        if fb:
            dmv = torch.from_numpy(fo.default_grid()).float().unsqueeze(0)
            pred = fo.predict_feature(bb.encoder, bb.decoder, prev[:1], nxt[:1], g1, g2, k, dmv, nw)
        else:
            pred = fo.predict_segmentation(bb.encoder, bb.decoder, prev[:1], nxt[:1], g1, g2, k, nw)
    np.testing.assert_allclose(fwd.numpy(), g["forward"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(pred.numpy(), g["predict"], rtol=1e-5, atol=1e-5)


def test_metric_oracles_match_reference_golden():
    g = load("metric_cases")
    keys = sorted(k[:-4] for k in g.files if k.endswith("_iut"))
    assert "mask" in keys and len(keys) == 5
    for key in keys:
        pred, target, iut = g[key + "_pred"], g[key + "_target"], g[key + "_iut"]
        K = iut.shape[1]
        assert np.array_equal(np.stack(mo.intersection_and_union_np(pred, target, K, 255)), iut)
        assert np.array_equal(co.counts(pred, target, K, 255, numpy_bins=True), iut)
        # the torch.histc convention differs from numpy's exactly by the closed last bin (target == K)
        hist = np.stack(mo.intersection_and_union_histc_ints(pred, target, K, 255))
        assert np.array_equal(co.counts(pred, target, K, 255, numpy_bins=False), hist)
        extra = int((target == K).sum())
        assert hist[2, K - 1] + extra == iut[2, K - 1]


def test_default_grid_matches_reference():
    g = load("default_grid")["grid"]
    assert g.shape == (67, 120, 2) and g.dtype == np.float64
    assert np.array_equal(fo.default_grid(), g)
    from flood_uav_video_segmentation_b200.flow.model import get_default_grid
    assert np.array_equal(get_default_grid(), g)


def test_c_oracle_pieces_against_torch_cpu():
    """The C restatement of the CUDA formulas stays within the documented CPU/CUDA gap of torch-CPU's kernels."""
    import torch.nn.functional as F
    x = torch.randn(1, 4, 33, 47, generator=torch.Generator().manual_seed(0))
    from flood_uav_video_segmentation_b200.synthetic import flow_grids
    g = flow_grids(33 * 4, 47 * 4, 2, "block", jitter=0.2)[0]
    for ac in (False, True):
        ref = F.grid_sample(x, g, mode="bilinear", padding_mode="border", align_corners=ac)[0].numpy()
        np.testing.assert_allclose(co.grid_sample(x[0].numpy(), g[0].numpy(), ac), ref, rtol=1e-5, atol=ATOL_CPU_VS_CUDA)
    up = F.interpolate(x, size=(97, 131), mode="bilinear", align_corners=True)[0].numpy()
    np.testing.assert_allclose(co.upsample_ac(x[0].numpy(), (97, 131)), up, rtol=1e-5, atol=1e-5)
    a, b = x[0].numpy(), torch.randn(4, 33, 47, generator=torch.Generator().manual_seed(1)).numpy()
    ref = (3 / 7 * torch.from_numpy(a) + 4 / 7 * torch.from_numpy(b)).numpy()
    assert np.array_equal(co.blend(a, b, 3 / 7, 4 / 7).view(np.int32), ref.view(np.int32))
    t = torch.randn(2, 5, 9, 11, generator=torch.Generator().manual_seed(2)).round()
    t[0, 1, 0, 0] = float("nan")
    assert np.array_equal(co.argmax(t.numpy()), t.max(1)[1].numpy().astype(np.uint8))
