"""Packed per-video grid store.

The reference keeps one float64 `.npy` file per frame and direction (`frames/<v_id>/grids/<i>.npy`,
`frames/<v_id>/inv_grids/<i>.npy`, written by dataset/flow/extract_motion_vectors.py:101-104), loads 2(k-1) of them
per interval, casts each to float32 (flow/dataset.py:239-240) and hands them to the model as two python lists
(flow/dataset.py:138-146): 2(k-1) small files, casts and host-to-device copies per interval.

GridPack holds the same numbers once per video: two float32 tensors [F,Hg,Wg,2] (forward / inverse grids of frames
first_id .. first_id+F-1), optionally in pinned host memory or on the device.  `interval(f_index, k)` returns exactly
what the reference's predict branch builds — mvs_left = [grid(f+1) .. grid(f+k-1)], mvs_right = [inv_grid(f+k-1) ..
inv_grid(f+1)] — as lists of VIEWS of the pack (no copy; the interval kernels take them as a pointer table), and
`interval_to(device, ...)` stages the k-1 consecutive frames of both directions with two contiguous copies.
"""
from __future__ import annotations

import os

import numpy as np
import torch


class GridPack:
    def __init__(self, grids: torch.Tensor, inv_grids: torch.Tensor, first_id: int = 0):
        if grids.shape != inv_grids.shape or grids.dim() != 4 or grids.shape[-1] != 2:
            raise ValueError(f"GridPack: expected two [F,Hg,Wg,2] tensors, got {tuple(grids.shape)} / {tuple(inv_grids.shape)}")
        self.grids = grids.to(torch.float32).contiguous()          # the reference's .astype('float32'), flow/dataset.py:240
        self.inv_grids = inv_grids.to(torch.float32).contiguous()
        self.first_id = int(first_id)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_arrays(cls, grids, inv_grids, first_id=0):
        """grids / inv_grids: sequences of [Hg,Wg,2] arrays (any float dtype), one per consecutive frame id."""
        g = torch.from_numpy(np.stack([np.asarray(a).astype("float32") for a in grids]))
        ig = torch.from_numpy(np.stack([np.asarray(a).astype("float32") for a in inv_grids]))
        return cls(g, ig, first_id)

    @classmethod
    def from_directory(cls, data_root, v_id, first_id=None, last_id=None):
        """Reads the reference's per-frame files (flow/dataset.py:231-240) for the consecutive frame ids that have both
        a grid and an inverse grid."""
        gdir = os.path.join(data_root, "frames", v_id, "grids")
        idir = os.path.join(data_root, "frames", v_id, "inv_grids")
        ids = sorted(int(os.path.splitext(f)[0]) for f in os.listdir(gdir) if f.endswith(".npy")
                     and os.path.exists(os.path.join(idir, f)))
        if first_id is not None:
            ids = [i for i in ids if i >= first_id]
        if last_id is not None:
            ids = [i for i in ids if i <= last_id]
        if not ids:
            raise FileNotFoundError(f"GridPack: no grid files under {gdir}")
        run = [ids[0]]
        for i in ids[1:]:                       # the longest consecutive run starting at the first id
            if i != run[-1] + 1:
                break
            run.append(i)
        load = lambda d, i: np.load(os.path.join(d, f"{i}.npy")).astype("float32")      # noqa: E731
        return cls.from_arrays([load(gdir, i) for i in run], [load(idir, i) for i in run], run[0])

    def save(self, path):
        torch.save({"grids": self.grids.cpu(), "inv_grids": self.inv_grids.cpu(), "first_id": self.first_id}, path)

    @classmethod
    def load(cls, path):
        d = torch.load(path, map_location="cpu")
        return cls(d["grids"], d["inv_grids"], d["first_id"])

    # ------------------------------------------------------------------ placement
    def pin(self):
        self.grids, self.inv_grids = self.grids.pin_memory(), self.inv_grids.pin_memory()
        return self

    def to(self, device, non_blocking=True):
        return GridPack(self.grids.to(device, non_blocking=non_blocking), self.inv_grids.to(device, non_blocking=non_blocking),
                        self.first_id)

    def __len__(self):
        return self.grids.shape[0]

    @property
    def nbytes(self):
        return 2 * self.grids.numel() * 4

    # ------------------------------------------------------------------ access
    def _slice(self, f_index, k):
        lo = f_index + 1 - self.first_id
        hi = f_index + k - self.first_id            # exclusive: frames f+1 .. f+k-1
        if lo < 0 or hi > len(self):
            raise IndexError(f"GridPack: interval {f_index}..{f_index + k} outside frames {self.first_id}..{self.first_id + len(self) - 1}")
        return lo, hi

    def interval(self, f_index, k):
        """(mvs_left, mvs_right) of the interval starting at key frame f_index (flow/dataset.py:138-146): lists of k-1
        views [1,Hg,Wg,2]."""
        lo, hi = self._slice(f_index, k)
        left = [self.grids[j:j + 1] for j in range(lo, hi)]
        right = [self.inv_grids[j:j + 1] for j in range(hi - 1, lo - 1, -1)]      # mvs_right.reverse()
        return left, right

    def interval_to(self, device, f_index, k, out=None, stream=None):
        """Stages the interval's 2(k-1) grids on `device` with TWO contiguous copies (one per direction) and returns
        (mvs_left, mvs_right) as views of the staging buffer `out` ([2,k-1,Hg,Wg,2], allocated when None)."""
        lo, hi = self._slice(f_index, k)
        if out is None:
            out = torch.empty((2, k - 1) + tuple(self.grids.shape[1:]), dtype=torch.float32, device=device)
        ctx = torch.cuda.stream(stream) if stream is not None else _null()
        with ctx:
            out[0].copy_(self.grids[lo:hi], non_blocking=True)
            out[1].copy_(self.inv_grids[lo:hi], non_blocking=True)
        left = [out[0, j:j + 1] for j in range(k - 1)]
        right = [out[1, j:j + 1] for j in range(k - 2, -1, -1)]
        return left, right


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
