"""The C-ABI library loads and exports every symbol include/rtb200.h declares (no compute calls: CPU only)."""
import ctypes
import os
import re

import pytest

from raytracer_b200 import capi


def header_functions(root):
    text = open(os.path.join(root, "include", "rtb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtb_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree(repo_root):
    assert header_functions(repo_root) == sorted(capi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(repo_root):
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in header_functions(repo_root):
        assert hasattr(lib, name), "librtb200.so does not export %s" % name


def test_struct_sizes_match_header():
    # sizes the C compiler produces for the POD structs (static_asserts in rtb_upload.h / rtb_scene_file.h rely on them)
    assert ctypes.sizeof(capi.RtbBvhNode) == 32
    assert ctypes.sizeof(capi.RtbMat4) == 64
    assert ctypes.sizeof(capi.RtbShape) == 64
    assert ctypes.sizeof(capi.RtbMaterial) == 64
    assert ctypes.sizeof(capi.RtbLight) == 176
    assert ctypes.sizeof(capi.RtbObject) == 144
    assert ctypes.sizeof(capi.RtbCamera) == 96
    assert capi.RAY_DTYPE.itemsize == 64 and capi.HIT_DTYPE.itemsize == 24


def test_no_cpu_fallback_without_device():
    """Without a CUDA device rtb_create must fail loudly (RTB_ERR_NO_DEVICE), never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.RtbError) as e:
        capi.Context()
    assert "no CUDA device" in str(e.value)
