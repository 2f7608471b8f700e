"""CPU tier: libfuvs.so loads, exports every symbol include/*.h declares, and refuses to compute without an sm_100
device (no CPU fallback)."""
import ctypes
import glob
import os
import re

import pytest
import torch

from flood_uav_video_segmentation_b200 import _lib, kernels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        src = open(h).read()
        names += re.findall(r"FUVS_API\s+[\w\s\*]+?\b(fuvs_\w+)\s*\(", src)
    return sorted(set(names))


def test_headers_declare_the_expected_surface():
    names = declared_symbols()
    for must in ("fuvs_linear_blend_argmax", "fuvs_warp_step", "fuvs_dense_interval", "fuvs_block_interval",
                 "fuvs_upsample_bilinear_ac", "fuvs_blend_argmax", "fuvs_argmax", "fuvs_confusion",
                 "fuvs_temporal_counts", "fuvs_abi_version", "fuvs_last_error", "fuvs_launch_count"):
        assert must in names
    assert len(names) >= 18


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} declared in include/ but not exported by libfuvs.so"


def test_binding_covers_every_declared_symbol():
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.fuvs_abi_version() == _lib.FUVS_ABI_VERSION
    assert lib.fuvs_launch_count() >= 0
    assert lib.fuvs_dense_scratch_floats(5, 1080, 1920, 5) == 2 * 3 * 5 * 1080 * 1920
    assert lib.fuvs_dense_scratch_floats(5, 1080, 1920, 2) == 0
    assert lib.fuvs_block_scratch_floats(5, 67, 120, 5) == 2 * 4 * 5 * 67 * 120


def test_library_is_sm100_only_and_has_device_code():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    lib = _lib.load()
    assert lib.fuvs_device_ok() == -4                                   # FUVS_ENODEV
    assert b"no CPU path" in lib.fuvs_last_error() or b"sm_100" in lib.fuvs_last_error()
    # compute entry points refuse (with NULL pointers: nothing is dereferenced before the device check)
    assert lib.fuvs_linear_blend_argmax(None, None, 5, 8, 8, 5, None, None, None, None, 255, None) == -4
    assert lib.fuvs_confusion(None, 0, None, 0, 10, 5, 255, 0, None, None) == -4
    with pytest.raises(kernels.FuvsError):
        kernels.linear_blend_argmax(torch.zeros(5, 8, 8), torch.zeros(5, 8, 8), 5)
    with pytest.raises(kernels.FuvsError):
        kernels.confusion(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.int64), 5)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libfuvs.so"))
    with pytest.raises(_lib.FuvsError, match="no CPU or PyTorch fallback"):
        _lib.load()
