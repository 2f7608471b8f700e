"""CPU tier, authoring container only: re-imports the LIVE reference (/root/reference) and checks the oracle's
restatement against it on fresh seeded inputs.  Skipped where the reference tree is absent (the GPU box)."""
import contextlib
import os
import sys
import types

import numpy as np
import pytest
import torch
from torch import nn

REF = os.environ.get("FUVS_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "flow")), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    try:
        import flow.model as fm
        import util.util as uu
        yield types.SimpleNamespace(FlowModel=fm.FlowModel, iau=uu.intersectionAndUnion, grid=fm.get_default_grid)
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "flow" or k.startswith("flow.") or k == "util" or k.startswith("util.")]:
            del sys.modules[k]


class Prof:
    @contextlib.contextmanager
    def profile(self, name):
        yield


@pytest.mark.parametrize("mode", ["linear", "block", "dense"])
@pytest.mark.parametrize("seed", [0, 1])
def test_predict_segmentation_restatement(ref, mode, seed):
    from flood_uav_video_segmentation_b200.synthetic import flow_grids, keyframe_logits
    from oracle import flow_oracle as fo
    C, H, W, n = 4, 40 + seed * 9, 56, 3 + seed * 2
    o, o_next = keyframe_logits(C, H, W, seed, 0)[None], keyframe_logits(C, H, W, seed, 1)[None]
    if mode == "linear":
        gl = gr = [torch.zeros(1, 1)] * (n - 1)
    else:
        gl, gr = flow_grids(H, W, n, mode, clip=seed, side=0), flow_grids(H, W, n, mode, clip=seed, side=1)
    bb = types.SimpleNamespace(encoder=nn.Identity(), decoder=nn.Identity())
    m = ref.FlowModel(bb, feature_based=False, no_warp=(mode == "linear")).eval()
    with torch.no_grad():
        want = m.predict(o, o_next, gl, gr, n, Prof())["pred"]
        got = fo.predict_segmentation(bb.encoder, bb.decoder, o, o_next, gl, gr, n, no_warp=(mode == "linear"))
    assert torch.equal(want, got)


def test_metric_restatement(ref):
    from oracle import metric_oracle as mo
    g = torch.Generator().manual_seed(0)
    for K in (2, 5, 7):
        pred = torch.randint(0, K, (3, 31, 29), generator=g).numpy()
        target = torch.randint(0, K + 2, (3, 31, 29), generator=g).numpy()
        target[0, :3] = 255
        for a, b in zip(ref.iau(pred, target, K, 255), mo.intersection_and_union_np(pred, target, K, 255)):
            assert np.array_equal(a, b)
    assert np.array_equal(ref.grid(), __import__("oracle.flow_oracle", fromlist=["x"]).default_grid())
