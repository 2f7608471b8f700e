"""CPU tier for the sliding-crop route (flow/base.py:182-234, flow/transform.py:215-261): the oracle's restatement
against golden vectors produced by the un-modified reference (oracle/make_golden.py), and the plain-numpy restatement
of cv2's INTER_LINEAR — the formula the CUDA kernel implements — against cv2 itself."""
import os

import numpy as np
import pytest
import torch
from torch import nn

from oracle import crop_oracle as co
from oracle import flow_oracle as fo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "crop_cases.npz")


class TinyBackbone(nn.Module):
    def __init__(self, classes=5, feat=12, stride=8):
        super().__init__()
        torch.manual_seed(0)
        self.encoder = nn.Sequential(nn.Conv2d(3, feat, 3, stride=stride, padding=1), nn.ReLU())
        self.decoder = nn.Conv2d(feat, classes, 1)


def test_crop_grid_oracle_matches_reference_golden():
    cv2 = pytest.importorskip("cv2")  # noqa: F841  the reference itself needs it (flow/transform.py:6)
    z = np.load(GOLD)
    for k in range(int(z["n_cases"])):
        H, W, ch, cw, ho, wo = (int(v) for v in z[f"c{k}_args"])
        grid = torch.from_numpy(z[f"c{k}_grid"])
        got, _ = co.crop_motion_vector([grid], [grid], H, W, ch, cw, ho, wo)
        assert got[0].shape == (1, ch // 16, cw // 16, 2)
        assert np.array_equal(got[0].numpy().view(np.int32), z[f"c{k}_out"].view(np.int32)), f"case {k}"


def test_numpy_restatement_of_cv2_linear_resize_is_bit_exact():
    """Every crop case through resize_linear_np instead of cv2 (no cv2 needed: compared with the reference's output)."""
    z = np.load(GOLD)
    for k in range(int(z["n_cases"])):
        H, W, ch, cw, ho, wo = (int(v) for v in z[f"c{k}_args"])
        grid = torch.from_numpy(z[f"c{k}_grid"])
        got, _ = co.crop_motion_vector([grid], [grid], H, W, ch, cw, ho, wo, resize=co.resize_linear_np)
        assert np.array_equal(got[0].numpy().view(np.int32), z[f"c{k}_out"].view(np.int32)), f"case {k}"


@pytest.mark.parametrize("shape", [(28, 28, 27, 27), (27, 28, 27, 27), (433, 433, 27, 27), (40, 33, 27, 20), (9, 9, 4, 4)])
def test_resize_linear_np_vs_cv2(shape):
    cv2 = pytest.importorskip("cv2")
    ih, iw, oh, ow = shape
    m = (np.random.default_rng(ih * 1000 + iw).random((ih, iw, 2), dtype=np.float32) * 2 - 1).astype(np.float32)
    ref = cv2.resize(m, (ow, oh), interpolation=cv2.INTER_LINEAR)
    got = co.resize_linear_np(m, ow, oh)
    assert np.array_equal(got.view(np.int32), ref.view(np.int32))


def test_crop_windows_known_answers():
    """flow/base.py:183-200 at the reference's sizes: 433x433 crops on 1072x1920 -> 4 x 7 windows."""
    w = list(co.crop_windows(1072, 1920, 433, 433))
    assert len(w) == 28
    assert w[0] == (0, 433, 0, 433) and w[-1] == (1072 - 433, 1072, 1920 - 433, 1920)
    assert all(e_h - s_h == 433 and e_w - s_w == 433 and s_h >= 0 and s_w >= 0 for s_h, e_h, s_w, e_w in w)
    covered = np.zeros((1072, 1920), bool)
    for s_h, e_h, s_w, e_w in w:
        covered[s_h:e_h, s_w:e_w] = True
    assert covered.all()
    assert list(co.crop_windows(433, 433, 433, 433)) == [(0, 433, 0, 433)]


def test_compute_output_oracle_matches_reference_golden():
    pytest.importorskip("cv2")
    z = np.load(GOLD)
    ch, cw, ncrops = (int(v) for v in z["full_crop"])
    prev, nxt = torch.from_numpy(z["full_prev"]), torch.from_numpy(z["full_next"])
    gl = [torch.from_numpy(g) for g in z["full_gl"]]
    gr = [torch.from_numpy(g) for g in z["full_gr"]]
    n = len(gl) + 1
    bb = TinyBackbone().eval()
    torch.set_num_threads(1)

    def fn(p, q, ml, mr):
        return co.crop_softmax(fo.predict_segmentation(bb.encoder, bb.decoder, p, q, ml, mr, n), p.shape[2], p.shape[3])

    with torch.no_grad():
        canvas = co.compute_output(n, fn, prev, nxt, gl, gr, 5, ch, cw)
    assert len(list(co.crop_windows(prev.shape[2], prev.shape[3], ch, cw))) == ncrops
    assert np.array_equal(canvas[:, :, ::3, ::3].numpy(), z["full_canvas_sub"])
    assert np.array_equal(canvas.max(1)[1].numpy().astype(np.uint8), z["full_labels"])


def test_crop_properties_random_shapes():
    """Property tests (hypothesis): the crop windows tile the frame for any frame / crop size, the host-side mirror
    enumerates the same windows as the oracle, and the numpy restatement of cv2's bilinear resize is bit-exact for
    arbitrary down-scales (the only direction crop_motion_vector uses)."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    from flood_uav_video_segmentation_b200.flow.base import FlowBaseModel

    @hyp.settings(max_examples=60, deadline=None)
    @hyp.given(st.integers(17, 400), st.integers(17, 400), st.integers(16, 200), st.integers(16, 200))
    def windows(new_h, new_w, crop_h, crop_w):
        crop_h, crop_w = min(crop_h, new_h), min(crop_w, new_w)
        w = list(co.crop_windows(new_h, new_w, crop_h, crop_w))
        assert w == list(FlowBaseModel.crop_windows(new_h, new_w, crop_h, crop_w))
        cover = np.zeros((new_h, new_w), bool)
        for s_h, e_h, s_w, e_w in w:
            assert 0 <= s_h and e_h <= new_h and 0 <= s_w and e_w <= new_w and e_h - s_h == crop_h and e_w - s_w == crop_w
            cover[s_h:e_h, s_w:e_w] = True
        assert cover.all()

    windows()
    cv2 = pytest.importorskip("cv2")

    @hyp.settings(max_examples=40, deadline=None)
    @hyp.given(st.integers(2, 60), st.integers(2, 60), st.integers(1, 30), st.integers(1, 30), st.integers(0, 2 ** 31 - 1))
    def resize(ih, iw, oh, ow, seed):
        oh, ow = min(oh, ih), min(ow, iw)                      # down-scale or identity
        m = (np.random.default_rng(seed).random((ih, iw, 2), dtype=np.float32) * 2 - 1).astype(np.float32)
        ref = cv2.resize(m, (ow, oh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(co.resize_linear_np(m, ow, oh).view(np.int32), ref.view(np.int32))

    resize()
