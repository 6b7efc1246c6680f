"""CPU tests: the C-ABI library builds, loads and exports every symbol include/tepose_b200.h
declares (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

import tepose_b200._native as nv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tepose_b200.h")).read()
    return sorted(set(re.findall(r"^TP_API [^;(]*?\b(tp_[a-z0-9_]+)\(", text, flags=re.M)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(nv.EXPORTS)
    assert sorted(nv._SIGNATURES) == sorted(nv.EXPORTS)


def test_library_builds_and_exports_every_symbol():
    from tepose_b200 import build
    path = build.build()
    assert os.path.isfile(path)
    handle = ctypes.CDLL(path)
    for name in header_symbols():
        assert hasattr(handle, name), name
    handle.tp_version.restype = ctypes.c_int
    assert handle.tp_version() == 100


def test_sass_contains_blackwell_tensor_path():
    """The K1 kernel must be tcgen05 + TMA, not a recompiled legacy path."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    from tepose_b200 import build
    build.build()
    obj = os.path.join(ROOT, "tepose_b200", "build", "gemm_tc.o")
    sass = subprocess.run([cuobjdump, "-sass", obj], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass


def test_struct_sizes_match_c_layout():
    # natural alignment, no packing pragmas in the header
    assert ctypes.sizeof(nv.GemmSeg) == 56          # + flags, ldr, residual (epilogue options)
    assert ctypes.sizeof(nv.GruJob) == 108 + 4 + 16
    assert ctypes.sizeof(nv.IefWeights) == 56
    assert ctypes.sizeof(nv.SmplModel) == 104
