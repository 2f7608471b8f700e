"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF (imported from /root/reference, read-only).

Run in the authoring container only:  python oracle/make_golden.py
The reference holds no tests or golden vectors for this path, so these fixtures — outputs of the
unmodified reference modules flow/model.py and util/util.py on seeded inputs, torch-CPU — are what pins the oracle
and, through it, the kernels.  Inputs are regenerated from seeds by flood_uav_video_segmentation_b200.synthetic
(and stored too, so the fixtures stay valid if a generator changes).
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FUVS_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from flow.model import FlowModel as RefFlowModel, get_default_grid as ref_default_grid  # noqa: E402  (reference)
from util.util import intersectionAndUnion as ref_iau  # noqa: E402  (reference)

from flood_uav_video_segmentation_b200.synthetic import flow_grids, gt_labels, keyframe_logits  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class Prof:
    @contextlib.contextmanager
    def profile(self, name):
        yield


def identity_backbone():
    return types.SimpleNamespace(encoder=nn.Identity(), decoder=nn.Identity())


class TinyBackbone(nn.Module):
    def __init__(self, classes=5, feat=12, stride=8):
        super().__init__()
        torch.manual_seed(0)
        self.encoder = nn.Sequential(nn.Conv2d(3, feat, 3, stride=stride, padding=1), nn.ReLU())
        self.decoder = nn.Conv2d(feat, classes, 1)


def interval_case(name, C, H, W, n, mode):
    o, o_next = keyframe_logits(C, H, W, 0, 0)[None], keyframe_logits(C, H, W, 0, 1)[None]
    no_warp = mode == "linear"
    if no_warp:
        gl = gr = [torch.zeros(1, 1)] * (n - 1)
    else:
        gl, gr = flow_grids(H, W, n, mode, side=0), flow_grids(H, W, n, mode, side=1)
    m = RefFlowModel(identity_backbone(), feature_based=False, no_warp=no_warp).eval()
    with torch.no_grad():
        pred = m.predict(o, o_next, gl, gr, n, Prof())["pred"]
    labels = pred.data.max(1)[1]
    # temporal-consistency counts exactly as flow/base.py:280-295 does them (numpy path of compute_metrics)
    tot = [np.zeros(C, np.int64) for _ in range(3)]
    lab = labels.numpy()
    for p in range(1, n):
        for acc, v in zip(tot, ref_iau(lab[p][None], lab[p - 1][None], C, 255)):
            acc += v
    # margin between the two best logits: label comparisons against torch-CUDA-exact kernels are only
    # meaningful where the CPU/CUDA grid_sample difference (<= ~1e-4) cannot flip the arg-max
    top2 = pred.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1]).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), prev=o.numpy(), next=o_next.numpy(),
                        grids_left=np.stack([g.numpy() for g in gl]) if not no_warp else np.zeros(0, np.float32),
                        grids_right=np.stack([g.numpy() for g in gr]) if not no_warp else np.zeros(0, np.float32),
                        pred=pred.numpy(), labels=labels.numpy().astype(np.uint8), margin=margin.astype(np.float32),
                        counts=np.stack(tot), n=n, mode=mode)
    print(name, tuple(pred.shape))


def forward_case(name, feature_based, no_warp):
    H, W, k, B = 80, 112, 5, 2
    bb = TinyBackbone().eval()
    g = torch.Generator().manual_seed(3)
    prev, nxt = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 3, H, W, generator=g)
    left, right = torch.tensor([2, 1]), torch.tensor([3, 4])
    if no_warp:
        gl = gr = [torch.zeros(B, 1)] * (k - 1)
    else:
        def batched(side):
            per = [flow_grids(H, W, k, "block", clip=b, side=side) for b in range(B)]
            return [torch.cat([per[b][j] for b in range(B)], 0) for j in range(k - 1)]
        gl, gr = batched(0), batched(1)
    m = RefFlowModel(bb, feature_based=feature_based, no_warp=no_warp).eval()
    with torch.no_grad():
        fwd = m(None, prev, nxt, gl, gr, left, right)["pred"]
        pred = m.predict(prev[:1], nxt[:1], [x[:1] for x in gl], [x[:1] for x in gr], k, Prof())["pred"]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), prev=prev.numpy(), next=nxt.numpy(),
                        grids_left=np.stack([x.numpy() for x in gl]) if not no_warp else np.zeros(0, np.float32),
                        grids_right=np.stack([x.numpy() for x in gr]) if not no_warp else np.zeros(0, np.float32),
                        left=left.numpy(), right=right.numpy(), forward=fwd.numpy(), predict=pred.numpy(),
                        feature_based=feature_based, no_warp=no_warp, k=k)
    print(name, tuple(fwd.shape), tuple(pred.shape))


def metric_cases():
    rows = {}
    for K in (5, 2):
        for s, (H, W) in enumerate([(97, 131), (64, 64)]):
            g = torch.Generator().manual_seed(10 * K + s)
            pred = torch.randint(0, K, (H, W), generator=g).numpy()
            target = torch.randint(0, K + 2, (H, W), generator=g)          # classes K, K+1 are out of range
            target[torch.rand(H, W, generator=g) < 0.07] = 255
            target = target.numpy()
            i, u, t = ref_iau(pred, target, K, 255)
            key = f"K{K}_s{s}"
            rows[key + "_pred"], rows[key + "_target"] = pred.astype(np.uint8), target.astype(np.uint8)
            rows[key + "_iut"] = np.stack([i, u, t])
    # a real label mask from the reference's dataset (dataset/flow/masks), sub-sampled, against a shifted copy
    try:
        from PIL import Image
        mdir = os.path.join(REF, "dataset", "flow", "masks", "florida-01")
        files = sorted(os.listdir(mdir), key=lambda f: int(os.path.splitext(f)[0]))[:2]
        m0 = np.array(Image.open(os.path.join(mdir, files[0])))[::8, ::8]
        m1 = np.array(Image.open(os.path.join(mdir, files[1])))[::8, ::8]
        i, u, t = ref_iau(m0, m1, 5, 255)
        rows["mask_pred"], rows["mask_target"], rows["mask_iut"] = m0.astype(np.uint8), m1.astype(np.uint8), np.stack([i, u, t])
        print("real masks", files, m0.shape, np.unique(m0))
    except Exception as e:  # noqa: BLE001
        print("real masks skipped:", e)
    np.savez_compressed(os.path.join(OUT, "metric_cases.npz"), **rows)


def crop_cases():
    """crop_motion_vector (flow/transform.py:215-261, un-modified reference incl. its cv2.resize) on block and dense
    grids, and one full sliding-crop canvas: flow/base.py:182-209 restated (flow/base.py needs Lightning, which is
    absent) around the reference's own FlowModel.predict and crop_motion_vector."""
    from flow.transform import crop_motion_vector as ref_cmv  # noqa: E402  (reference; imports cv2)
    rows = {}
    H, W = 144, 208
    cases = [("block", 9, 13, 65, 65, 0, 0), ("block", 9, 13, 65, 65, 44, 88), ("block", 9, 13, 65, 65, 79, 143),
             ("dense", H, W, 65, 65, 44, 88), ("block", 9, 13, 50, 84, 37, 90)]
    for k, (mode, hg, wg, ch, cw, ho, wo) in enumerate(cases):
        g = flow_grids(H, W, 2, mode, clip=11 + k, side=0)[0]            # [1,Hg,Wg,2] float32
        assert tuple(g.shape[1:3]) == (hg, wg), g.shape
        out, _ = ref_cmv([g.clone()], [g.clone()], H, W, ch, cw, ho, wo)   # CPU tensors alias their numpy view: clone
        rows[f"c{k}_grid"], rows[f"c{k}_out"] = g.numpy(), out[0].numpy()
        rows[f"c{k}_args"] = np.array([H, W, ch, cw, ho, wo])
    rows["n_cases"] = np.array(len(cases))
    # full canvas
    n, C, ch, cw = 3, 5, 65, 65
    bb = TinyBackbone().eval()
    g = torch.Generator().manual_seed(21)
    prev, nxt = torch.randn(1, 3, H, W, generator=g), torch.randn(1, 3, H, W, generator=g)
    gl, gr = flow_grids(H, W, n, "block", clip=31, side=0), flow_grids(H, W, n, "block", clip=31, side=1)
    m = RefFlowModel(bb, feature_based=False, no_warp=False).eval()
    stride_h, stride_w = int(np.ceil(ch * 2 / 3)), int(np.ceil(cw * 2 / 3))
    grid_h, grid_w = int(np.ceil(float(H - ch) / stride_h) + 1), int(np.ceil(float(W - cw) / stride_w) + 1)
    canvas = torch.zeros((n, C, H, W), dtype=float)
    count = torch.zeros((H, W), dtype=float)
    with torch.no_grad():
        for ih in range(grid_h):
            for iw in range(grid_w):
                e_h, e_w = min(ih * stride_h + ch, H), min(iw * stride_w + cw, W)
                s_h, s_w = e_h - ch, e_w - cw
                ml, mr = ref_cmv([x.clone() for x in gl], [x.clone() for x in gr], H, W, ch, cw, s_h, s_w)
                out = m.predict(prev[:, :, s_h:e_h, s_w:e_w].clone(), nxt[:, :, s_h:e_h, s_w:e_w].clone(), ml, mr, n, Prof())["pred"]
                count[s_h:e_h, s_w:e_w] += 1
                canvas[:, :, s_h:e_h, s_w:e_w] += torch.nn.functional.softmax(out, dim=1)
    canvas /= count.unsqueeze(0).unsqueeze(0)
    rows.update(full_prev=prev.numpy(), full_next=nxt.numpy(), full_gl=np.stack([x.numpy() for x in gl]),
                full_gr=np.stack([x.numpy() for x in gr]), full_canvas_sub=canvas[:, :, ::3, ::3].numpy(),
                full_labels=canvas.max(1)[1].numpy().astype(np.uint8), full_crop=np.array([ch, cw, grid_h * grid_w]))
    np.savez_compressed(os.path.join(OUT, "crop_cases.npz"), **rows)
    print("crop cases", len(cases), "canvas", tuple(canvas.shape), "crops", grid_h * grid_w)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)        # fixtures must not depend on the thread count
    interval_case("linear_c5_48x64_n5", 5, 48, 64, 5, "linear")
    interval_case("linear_c2_37x53_n3", 2, 37, 53, 3, "linear")
    interval_case("block_c5_64x96_n5", 5, 64, 96, 5, "block")
    interval_case("block_c2_37x53_n4", 2, 37, 53, 4, "block")
    interval_case("dense_c5_48x64_n5", 5, 48, 64, 5, "dense")
    interval_case("dense_c3_37x53_n2", 3, 37, 53, 2, "dense")
    forward_case("forward_seg_warp", False, False)
    forward_case("forward_seg_nowarp", False, True)
    forward_case("forward_feat_warp", True, False)
    forward_case("forward_feat_nowarp", True, True)
    metric_cases()
    crop_cases()
    np.savez_compressed(os.path.join(OUT, "default_grid.npz"), grid=ref_default_grid())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "crop":      # regenerate only the crop fixtures
        torch.set_num_threads(1)
        crop_cases()
    else:
        main()
