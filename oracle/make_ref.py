"""Copies the reference's own modules for this path, UNMODIFIED, into oracle/_ref/ (git-ignored: the sources never
enter this repository's history; the directory is not gpurun-ignored, so it travels to the GPU box).

    python oracle/make_ref.py            # run where /root/reference exists (the authoring container)

The reference is pure Python, so "building" it is a byte-for-byte copy.  Provenance (source path, size, sha256 of
every file) is written to oracle/_ref/PROVENANCE.json; `verify()` re-hashes the copies, and `load()` imports them:
bench.py's `--impl reference` arm and `cpu_baseline` leg time THESE modules (FlowModel.predict, flow/model.py:109-249;
intersectionAndUnion, util/util.py:36-47), tests/test_ref_copy.py checks the oracle restatement against them, and
tools/path_fraction.py takes FlowPSPNet / FlowDeepLabv3 (model/pspnet.py:113-141, model/deeplabv3.py:47-73) from here.
TEST / MEASUREMENT INFRASTRUCTURE: nothing under flood_uav_video_segmentation_b200/ imports it.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("FUVS_REFERENCE", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
# the arithmetic core of the path and the key-frame networks it is measured with (these import cleanly
# with torch / numpy / PIL / torchvision only; flow/base.py and base/foundation.py need Lightning and stay restated)
FILES = ["flow/model.py", "util/util.py", "model/pspnet.py", "model/resnet.py", "model/wrapper.py", "model/deeplabv3.py"]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def make(verbose=True):
    """Copy FILES from the reference tree; returns True if oracle/_ref is complete afterwards."""
    if not os.path.isdir(REF_SRC):
        return verify()
    prov = {"source_root": REF_SRC, "upstream": "lenke182/flood-uav-video-segmentation", "files": {}}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        prov["files"][rel] = {"bytes": os.path.getsize(dst), "sha256": _sha(dst)}
    with open(os.path.join(REF_DST, "PROVENANCE.json"), "w") as f:
        json.dump(prov, f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference files copied unmodified from {REF_SRC}")
    return True


def verify():
    """True if every file listed in PROVENANCE.json is present with the recorded hash."""
    p = os.path.join(REF_DST, "PROVENANCE.json")
    if not os.path.exists(p):
        return False
    with open(p) as f:
        prov = json.load(f)
    for rel, meta in prov["files"].items():
        dst = os.path.join(REF_DST, rel)
        if not os.path.exists(dst) or _sha(dst) != meta["sha256"]:
            return False
    return set(FILES) <= set(prov["files"])


def load():
    """Imports the copied reference modules -> dict(FlowModel, get_default_grid, intersectionAndUnion,
    intersectionAndUnionGPU, AverageMeter) or None when oracle/_ref is absent / incomplete."""
    if not verify():
        return None
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    for name in ("flow", "flow.model", "util", "util.util", "model"):      # drop modules imported from another tree
        m = sys.modules.get(name)
        f = getattr(m, "__file__", None) or (list(getattr(m, "__path__", [])) or [""])[0] if m is not None else None
        if m is not None and not str(f).startswith(REF_DST):
            del sys.modules[name]
    fm = importlib.import_module("flow.model")
    uu = importlib.import_module("util.util")
    return {"FlowModel": fm.FlowModel, "get_default_grid": fm.get_default_grid,
            "intersectionAndUnion": uu.intersectionAndUnion, "intersectionAndUnionGPU": uu.intersectionAndUnionGPU,
            "AverageMeter": uu.AverageMeter, "dir": REF_DST}


if __name__ == "__main__":
    ok = make()
    print("oracle/_ref complete:", ok)
    sys.exit(0 if ok else 1)
