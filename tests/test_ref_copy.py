"""oracle/_ref — the reference's own modules, copied unmodified by oracle/make_ref.py — against the oracle restatement,
through the very function bench.py times for `cpu_baseline` / `--impl reference` (bench.cpu_interval)."""
import os

import numpy as np
import pytest
import torch

import bench
from flood_uav_video_segmentation_b200.synthetic import flow_grids, keyframe_logits
from oracle import flow_oracle as fo
from oracle import make_ref
from oracle import metric_oracle as mo

needs_ref = pytest.mark.skipif(not make_ref.verify(), reason="oracle/_ref absent (python oracle/make_ref.py where /root/reference exists)")


def test_copy_matches_reference_tree_when_present():
    """Where the reference tree exists the copies are byte-identical to it (provenance hashes re-checked)."""
    if not os.path.isdir(make_ref.REF_SRC):
        pytest.skip("reference tree absent")
    assert make_ref.make(verbose=False) and make_ref.verify()
    for rel in make_ref.FILES:
        with open(os.path.join(make_ref.REF_SRC, rel), "rb") as a, open(os.path.join(make_ref.REF_DST, rel), "rb") as b:
            assert a.read() == b.read(), rel


@needs_ref
@pytest.mark.parametrize("mode", ["linear", "block", "dense"])
def test_cpu_arm_runs_the_reference_modules(mode):
    C, H, W, n = 5, 48, 64, bench.K_DELTA
    keys = [keyframe_logits(C, H, W, 3, j)[None] for j in range(3)]
    grids = []
    for it in range(2):
        if mode == "linear":
            grids.append((None, None))
        else:
            grids.append((torch.cat(flow_grids(H, W, n, mode, clip=3, interval=it, side=0), 0),
                          torch.cat(flow_grids(H, W, n, mode, clip=3, interval=it, side=1), 0)))
    assert bench.cpu_kind() == "reference"
    last = None
    ident = torch.nn.Identity()
    for it in range(2):
        lab, counts, new_last = bench.cpu_interval(mode, keys, grids, it, last)
        if mode == "linear":
            gl = gr = [torch.zeros(1, 1)] * (n - 1)
        else:
            gl = [grids[it][0][j:j + 1] for j in range(n - 1)]
            gr = [grids[it][1][j:j + 1] for j in range(n - 1)]
        ref = fo.argmax_labels(fo.predict_segmentation(ident, ident, keys[it], keys[it + 1], gl, gr, n, no_warp=(mode == "linear")))
        assert np.array_equal(lab, ref.numpy().astype(np.uint8))
        (i, u, t), _ = mo.temporal_consistency_counts(ref.numpy(), C, 255, None if last is None else last.numpy())
        assert np.array_equal(np.stack(counts), np.stack([i, u, t]))
        last = new_last
