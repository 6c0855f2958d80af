"""CPU checks of the C-ABI boundary: the library loads, exports exactly the symbols include/abc_b200.h declares,
and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from abc_inference_transcription_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "abc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(abc_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/abc_b200.h but not exported"
    assert names == set(_lib.SYMBOLS), names ^ set(_lib.SYMBOLS)


def test_metadata_entry_points_work_without_gpu():
    lib = _lib.load()
    assert lib.abc_version() >= 100
    assert [lib.abc_n_params(m) for m in range(0, 7)] == [-1, 5, 5, 9, 9, 9, -1]
    assert [lib.abc_model_name(m).decode() for m in range(1, 6)] == ["const", "const_const", "kon", "alpha", "gamma"]
    assert lib.abc_model_name(0) is None


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    rc = lib.abc_create(0, ctypes.byref(ctx))
    assert rc == -2 and not ctx          # ABC_ERR_CUDA
    assert b"no CPU fallback" in lib.abc_last_error()
    from abc_inference_transcription_b200 import AbcEngine, AbcError
    with pytest.raises(AbcError):
        AbcEngine(0)


def test_design_struct_layout_matches_header():
    """field order/types of the ctypes mirror vs the C struct (sizes computed from the header text)"""
    assert ctypes.sizeof(_lib.AbcDesign) == 8 * (2 + 5 + 11 + 11 + 55 + 9) + 4 * 4 + (8 + 8 + 8) * 2 + 16
    assert ctypes.sizeof(_lib.AbcCounters) == 8 * 8
