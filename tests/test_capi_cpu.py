"""CPU checks of the C-ABI boundary: the library loads, exports exactly the symbols include/abc_b200.h declares,
and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os

import numpy as np
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


def header_prototypes():
    """{name: number of parameters} from include/abc_b200.h"""
    src = open(os.path.join(ROOT, "include", "abc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for name, params in re.findall(r"\b(abc_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", src):
        params = params.strip()
        protos[name] = 0 if params in ("", "void") else params.count(",") + 1
    return protos


def julia_ccalls():
    """(symbol, number of argument types) of every ccall in julia/*.jl (the Julia host cannot be executed here)"""
    out = []
    jdir = os.path.join(ROOT, "julia")
    for fn in sorted(os.listdir(jdir)):
        if not fn.endswith(".jl"):
            continue
        src = open(os.path.join(jdir, fn)).read()
        for mt in re.finditer(r"ccall\(\(:(abc_[a-z0-9_]+),\s*LIB\),\s*\w+,\s*\(", src):
            i, depth = mt.end(), 1
            start = i
            while depth:                       # matching parenthesis of the argument-type tuple
                depth += {"(": 1, ")": -1}.get(src[i], 0)
                i += 1
            types = src[start:i - 1].strip().rstrip(",")
            n, d = (0 if types == "" else 1), 0
            for ch in types:                   # top-level commas only (Ref{Ptr{Cvoid}} has none, but be safe)
                d += {"{": 1, "}": -1, "(": 1, ")": -1}.get(ch, 0)
                n += (ch == "," and d == 0)
            out.append((fn, mt.group(1), n))
    return out


def test_julia_binding_matches_the_header():
    """every ccall of the Julia host names a declared entry point and passes as many arguments as the prototype has"""
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 15
    for fn, name, nargs in calls:
        assert name in protos, f"{fn}: ccall of undeclared symbol {name}"
        assert nargs == protos[name], f"{fn}: {name} is called with {nargs} argument types, the header declares {protos[name]}"
    # the hot-path entry points are all bound
    bound = {name for _, name, _ in calls}
    for need in ("abc_create", "abc_set_design", "abc_set_data", "abc_fix_params", "abc_simulate", "abc_score",
                 "abc_simulate_score", "abc_accept_fetch", "abc_posterior_summary", "abc_set_option", "abc_host_alloc"):
        assert need in bound, need


def test_julia_host_never_includes_a_reference_script():
    """the drop-in must not pull the reference's scripts in (compute_errors.jl's tail runs the CPU scoring loop and writes
    to an HPC path, compute_errors.jl:72-81): every include of julia/*.jl resolves inside julia/"""
    jdir = os.path.join(ROOT, "julia")
    for fn in sorted(os.listdir(jdir)):
        if not fn.endswith(".jl"):
            continue
        for ln, line in enumerate(open(os.path.join(jdir, fn)), 1):
            code = line.split("#", 1)[0]
            for mt in re.finditer(r"\binclude\s*\((.*)\)", code):
                arg = mt.group(1)
                assert "scripts/" not in arg and "scripts\\" not in arg, f"{fn}:{ln} includes a reference script: {line.strip()}"
                assert "@__DIR__" in arg, f"{fn}:{ln}: include outside julia/: {line.strip()}"


def test_gene_ranges_cut_equal_tuple_mass():
    """abc_gene_ranges (host logic of the multi-GPU exchange): contiguous, covering, deterministic, balanced"""
    from abc_inference_transcription_b200 import gene_ranges
    rng = np.random.default_rng(0)
    counts = (rng.pareto(1.2, 3419) * 50).astype(np.int64)
    for n in (1, 2, 3, 8):
        b = gene_ranges(counts, n)
        assert b[0] == 0 and b[-1] == 3419 and np.all(np.diff(b) >= 0)
        mass = np.array([counts[b[k]:b[k + 1]].sum() for k in range(n)])
        assert mass.sum() == counts.sum()
        assert mass.max() <= counts.sum() / n + counts.max()          # within one gene of the ideal share
    z = gene_ranges(np.zeros(10, dtype=np.int64), 4)
    assert list(z) == [0, 2, 5, 7, 10]
    one = np.zeros(100, dtype=np.int64); one[37] = 5
    b = gene_ranges(one, 4)
    assert b[0] == 0 and b[-1] == 100 and sum(one[b[k]:b[k + 1]].sum() for k in range(4)) == 5
