"""CPU tier: host-side logic of the drop-in layer (no kernels are launched)."""
import os
import numpy as np
import pytest
import torch
from torch import nn

from flood_uav_video_segmentation_b200 import dist as fdist
from flood_uav_video_segmentation_b200 import kernels, synthetic
from flood_uav_video_segmentation_b200.base.foundation import epoch_metrics, round_train
from flood_uav_video_segmentation_b200.flow.model import FlowModel
from flood_uav_video_segmentation_b200.util.util import AverageMeter
from oracle import flow_oracle as fo
from oracle import metric_oracle as mo


def test_round_train_matches_reference_table():
    # base/foundation.py:34-42; SURVEY.md §5: 433->433, 1080->1073, 1920->1913
    assert [round_train(v, "pspnet") for v in (433, 1080, 1920, 873)] == [433, 1073, 1913, 873]
    assert round_train(433, "deeplabv3") == 433 and round_train(433, "vit") == 416
    with pytest.raises(AssertionError):
        round_train(433, "unet")


def test_average_meter_and_epoch_formulas():
    m = AverageMeter()
    m.update(np.array([1, 2, 3]))
    m.update(np.array([3, 2, 1]))
    assert np.array_equal(m.sum, [4, 4, 4]) and m.count == 2
    i, u, t = np.array([10, 0, 5]), np.array([20, 0, 10]), np.array([15, 0, 5])
    a, b = epoch_metrics(i, u, t), mo.epoch_metrics(i, u, t)
    for k in ("miou", "macc", "accuracy"):
        assert a[k] == b[k]
    assert a["miou"] == np.mean(i / (u + 1e-10))


def test_blend_weights_are_the_fp32_roundings_of_python_doubles():
    for n in range(2, 26):                      # experiments/frame_delta_example.yaml sweeps k = 2..25
        for p in range(1, n):
            w0 = torch.tensor((n - p) / n, dtype=torch.float64).float().item()
            assert np.float32((n - p) / n) == np.float32(w0)
            # ATen: tensor * python_float == tensor * float32(python_float)
            x = torch.tensor([1.2345678], dtype=torch.float32)
            assert torch.equal(x * ((n - p) / n), x * torch.tensor(np.float32((n - p) / n)))


def test_synthetic_grids_follow_the_reference_wire_format():
    g = synthetic.identity_grid(1072, 1920, "block")
    assert tuple(g.shape) == (67, 120, 2)
    assert np.allclose(g.numpy(), fo.default_grid().astype(np.float32))
    d = synthetic.identity_grid(8, 12, "dense")
    # a dense identity grid samples every pixel centre exactly: ((x+1)*W-1)/2 == j
    ix = ((d[..., 0].double() + 1) * 12 - 1) / 2
    assert torch.allclose(ix[0], torch.arange(12, dtype=torch.float64), atol=1e-6)
    gs = synthetic.flow_grids(64, 96, 5, "block", clip=1, interval=2, side=1)
    assert len(gs) == 4 and tuple(gs[0].shape) == (1, 4, 6, 2)
    assert torch.equal(gs[0], synthetic.flow_grids(64, 96, 5, "block", clip=1, interval=2, side=1)[0])
    ks = synthetic.clip_keyframes(5, 8, 8, 5, frames=16)
    assert len(ks) == 4          # key frames 0,5,10,15 -> 3 intervals / 12 interpolated frames


def test_clip_and_interval_sharding_cover_everything_once():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 9, 64):
            got = sum((fdist.shard_clips(n, r, world) for r in range(world)), [])
            assert got == list(range(n))
            spans = [fdist.shard_intervals(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_cpu_inference_raises_instead_of_falling_back():
    class BB(nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder, self.decoder = nn.Identity(), nn.Identity()
    m = FlowModel(BB(), feature_based=False, no_warp=True).eval()
    x = torch.zeros(1, 5, 8, 8)
    with torch.no_grad(), pytest.raises(kernels.FuvsError, match="no CPU implementation"):
        m.predict(x, x, [torch.zeros(1, 1)], [torch.zeros(1, 1)], 2, fo.NullProfiler())
    with torch.no_grad(), pytest.raises(kernels.FuvsError):
        m(None, x, x, [torch.zeros(1, 1)], [torch.zeros(1, 1)], [1], [1])
    # no parameters/buffers are added by the wrapper (checkpoint compatibility)
    assert list(m.state_dict().keys()) == []
    assert isinstance(m.default_motion_vector, torch.Tensor) and tuple(m.default_motion_vector.shape) == (1, 67, 120, 2)


def test_training_route_matches_oracle_on_cpu():
    """The autograd route (training only) composes the same torch ops as the reference."""
    class BB(nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.encoder = nn.Conv2d(3, 6, 3, stride=4, padding=1)
            self.decoder = nn.Conv2d(6, 5, 1)
    bb = BB()
    m = FlowModel(bb, feature_based=False, no_warp=False).train()
    g = torch.Generator().manual_seed(1)
    prev, nxt = torch.randn(2, 3, 32, 48, generator=g), torch.randn(2, 3, 32, 48, generator=g)
    per = [[synthetic.flow_grids(32, 48, 3, "block", clip=b, side=s) for b in range(2)] for s in range(2)]
    gl = [torch.cat([per[0][b][j] for b in range(2)], 0) for j in range(2)]
    gr = [torch.cat([per[1][b][j] for b in range(2)], 0) for j in range(2)]
    out = m(None, prev, nxt, gl, gr, [1, 2], [2, 1])["pred"]
    ref = fo.forward_segmentation(bb.encoder, bb.decoder, prev, nxt, gl, gr, [1, 2], [2, 1])
    assert torch.equal(out, ref) and out.requires_grad
    out.sum().backward()
    assert bb.encoder.weight.grad is not None


def test_gridpack_matches_the_reference_dataset_loading(tmp_path):
    """GridPack against the reference's own loading logic, restated from flow/dataset.py:138-146,
    231-240: per-frame float64 .npy files -> np.load(..).astype('float32') -> mvs_left in order, mvs_right reversed."""
    import numpy as np
    import torch
    from flood_uav_video_segmentation_b200.flow.gridpack import GridPack
    root, v_id, k = str(tmp_path), "florida-01", 5
    rng = np.random.default_rng(0)
    for name in ("grids", "inv_grids"):
        os.makedirs(os.path.join(root, "frames", v_id, name))
    ids = list(range(3, 20))
    for i in ids:
        for name in ("grids", "inv_grids"):
            np.save(os.path.join(root, "frames", v_id, name, f"{i}.npy"), rng.uniform(-1, 1, (67, 120, 2)))   # float64 like the reference
    np.save(os.path.join(root, "frames", v_id, "grids", "40.npy"), np.zeros((67, 120, 2)))                     # a gap: not part of the run

    def ref_load(i, name):                       # flow/dataset.py:239-240
        return np.load(os.path.join(root, "frames", v_id, name, f"{i}.npy")).astype("float32")

    pack = GridPack.from_directory(root, v_id)
    assert len(pack) == len(ids) and pack.first_id == 3 and pack.grids.dtype == torch.float32
    for f_index in (3, 7, 14):
        left, right = pack.interval(f_index, k)
        ref_left = [ref_load(f_index + i + 1, "grids") for i in range(k - 1)]                                  # :140-141
        ref_right = [ref_load(f_index + i + 1, "inv_grids") for i in range(k - 1)]                             # :143-144
        ref_right.reverse()                                                                                   # :146
        assert len(left) == len(right) == k - 1
        for a, b in zip(left, ref_left):
            assert tuple(a.shape) == (1, 67, 120, 2) and np.array_equal(a[0].numpy(), b)
        for a, b in zip(right, ref_right):
            assert np.array_equal(a[0].numpy(), b)
        l2, r2 = pack.interval_to("cpu", f_index, k)
        assert all(torch.equal(a, b) for a, b in zip(l2, left)) and all(torch.equal(a, b) for a, b in zip(r2, right))
    with pytest.raises(IndexError):
        pack.interval(16, k)
    p = os.path.join(root, "pack.pt")
    pack.save(p)
    again = GridPack.load(p)
    assert torch.equal(again.grids, pack.grids) and torch.equal(again.inv_grids, pack.inv_grids) and again.first_id == 3
