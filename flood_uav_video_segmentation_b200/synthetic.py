"""Seeded synthetic inputs in the reference's wire format.

Key-frame logits are what `model.decoder(model.encoder(frame))` hands to the
interpolation path; grids follow `dataset/flow/extract_motion_vectors.py` /
`flow/model.py:10-21`: normalised (x, y) sample positions, one per 16x16
macro-block (block mode) or one per pixel (dense mode), identity + jitter.
Generated on the CPU generator so every machine sees the same bits.
"""
from __future__ import annotations

import torch

BLOCK = 16


def keyframe_logits(C, H, W, clip=0, keyframe=0):
    g = torch.Generator().manual_seed(1234 + clip * 1000 + keyframe)
    return torch.randn(C, H, W, generator=g, dtype=torch.float32)


def identity_grid(H, W, mode):
    """[Hg,Wg,2] fp32 identity sampling grid for an H x W frame."""
    if mode == "block":
        hg, wg = H // BLOCK, W // BLOCK
        xs = (torch.arange(wg, dtype=torch.float64) * BLOCK + BLOCK // 2) / W * 2 - 1
        ys = (torch.arange(hg, dtype=torch.float64) * BLOCK + BLOCK // 2) / H * 2 - 1
    elif mode == "dense":
        hg, wg = H, W
        xs = (torch.arange(wg, dtype=torch.float64) + 0.5) / W * 2 - 1
        ys = (torch.arange(hg, dtype=torch.float64) + 0.5) / H * 2 - 1
    else:
        raise ValueError(mode)
    g = torch.empty(hg, wg, 2, dtype=torch.float64)
    g[:, :, 0] = xs[None, :]
    g[:, :, 1] = ys[:, None]
    return g.float()


def flow_grids(H, W, k, mode, clip=0, interval=0, side=0, jitter=0.05):
    """k-1 grids [1,Hg,Wg,2] (the list format of flow/dataset.py:138-146)."""
    base = identity_grid(H, W, mode)
    g = torch.Generator().manual_seed(4321 + clip * 1000 + interval * 10 + side)
    out = []
    for _ in range(k - 1):
        out.append((base + (torch.rand(base.shape, generator=g) - 0.5) * jitter).unsqueeze(0))
    return out


def clip_keyframes(C, H, W, k, frames=16, clip=0):
    """Key frames 0, k, 2k, ... of a `frames`-long clip (flow/dataset.py:64,86): list of [1,C,H,W]."""
    n_int = (frames - 1) // k
    return [keyframe_logits(C, H, W, clip, j).unsqueeze(0) for j in range(n_int + 1)]


def gt_labels(H, W, C, seed=0, ignore_frac=0.05, ignore_index=255):
    g = torch.Generator().manual_seed(777 + seed)
    t = torch.randint(0, C, (H, W), generator=g, dtype=torch.int64)
    m = torch.rand(H, W, generator=g) < ignore_frac
    t[m] = ignore_index
    return t
