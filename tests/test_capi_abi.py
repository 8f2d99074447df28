"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""

import ctypes
import glob
import os
import re

from zyg_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for header in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        names.update(re.findall(r"\b((?:su|zyg|zygpu)_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_headers_declare_something():
    syms = declared_symbols()
    assert "zygpu_trace_batch" in syms and "zyg_mesh_build" in syms


def test_every_declared_symbol_is_exported():
    cdll = lib.load_library()
    missing = [s for s in declared_symbols() if not hasattr(cdll, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_struct_sizes_match_header():
    assert lib.RAY_DTYPE.itemsize == 32
    assert lib.HIT_DTYPE.itemsize == 16
    assert ctypes.sizeof(lib.TraceCounters) == 32
    assert ctypes.sizeof(lib.MeshInfo) == 8 * 4 + 6 * 4


def test_missing_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        return
    try:
        lib.Device(0)
    except RuntimeError as e:
        assert "zygpu_create" in str(e)
    else:
        raise AssertionError("Device() must raise without a GPU (no CPU fallback)")
