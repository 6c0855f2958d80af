"""The reference's on-disk layouts written by the library's host-side writers (include/abc_b200.h "on-disk layouts",
csrc/abc_io.cu): the Julia host (and this mirror) hands over the arrays the compute entry points returned
instead of formatting ~65 KB of text per particle itself (compute_errors.jl:66-68, process_error_files.jl:3-7,
abc_simulation.jl:47-61, 89-95, accepted_particles.jl:19-30, recover_statistics.jl:49-68).  No device is needed."""
import ctypes
import os

import numpy as np

from . import _lib
from .model import ID_LABELS, model_name

MOMENT_FILES = ["mean_u", "mean_l", "var_u", "cov_ul", "var_l"]          # recover_statistics.jl:52-66


def format_float64(x):
    """Julia's print(io, x::Float64)"""
    buf = ctypes.create_string_buffer(40)
    n = _lib.load().abc_format_float64(float(x), buf, 40)
    if n < 0:
        _lib.check(n)
    return buf.value.decode()


def writedlm(path, a, append=True):
    """writedlm(io, A) of a Float64 matrix (1-D input = one value per row, like Julia)"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    _lib.check(_lib.load().abc_writedlm(os.fsencode(path), _lib.ptr(a), a.shape[0], a.shape[1], int(bool(append))))


def write_simulation(root, m, submit, theta, stats, first_trial=1):
    """the seven appends of abc_simulation.jl:47-61, 89-95 for a batch: <root>/data/simulations/<model>/..."""
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    stats = np.ascontiguousarray(stats, dtype=np.float64)
    assert theta.shape[0] == stats.shape[0] and stats.shape[1] == _lib.NSTATS
    d = os.path.join(root, "data", "simulations")
    _lib.check(_lib.load().abc_write_simulation(os.fsencode(d), int(m), int(submit), _lib.ptr(theta), _lib.ptr(stats),
                                                theta.shape[0], int(first_trial)))


def write_accepted(path, offsets, idx, append=True):
    """particles_<model>.txt: one line of 1-based indices per gene, "0" when none (accepted_particles.jl:19-30)"""
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    _lib.check(_lib.load().abc_write_accepted(os.fsencode(path), _lib.ptr(offsets), _lib.ptr(idx), len(offsets) - 1, int(bool(append))))


def write_error_columns(col_dir, err_gene_major, append=True):
    """err_gene_major: (G, n) = ERR_GENE_MAJOR output of score / simulate_score -> <col_dir>/x<g>.f64 (+ meta.txt)"""
    e = np.ascontiguousarray(err_gene_major, dtype=np.float64)
    _lib.check(_lib.load().abc_write_error_columns(os.fsencode(col_dir), _lib.ptr(e), e.shape[1], e.shape[1], e.shape[0], int(bool(append))))


def read_error_column(col_dir, g):
    """f["x<g>"] of the reference's JDFFile (accepted_particles.jl:14-18); g is 1-based"""
    n = ctypes.c_int64()
    lib = _lib.load()
    _lib.check(lib.abc_read_error_column(os.fsencode(col_dir), int(g), None, 0, ctypes.byref(n)))
    out = np.empty(n.value, dtype=np.float64)
    _lib.check(lib.abc_read_error_column(os.fsencode(col_dir), int(g), _lib.ptr(out), n.value, ctypes.byref(n)))
    return out


def recover_statistics(engine, m, maps, root="."):
    """scripts/recover_statistics.jl:49-68 (wrapper.jl:109-111): run_part_sim without downsampling on every row of `maps`
    (data/posterior_estimates/map_sets_<model>.txt) and append, per labelling condition, one row of the five ages to
    data/recovered_statistics/<model>/<condition>/{mean_u, mean_l, var_u, cov_ul, var_l}.txt.  The engine's design must be
    the recovery design (sim_kind = ODE, downsampling off, iv[1] = 1/2: recover_statistics.jl:33-42).  Returns the moments
    (n, 11, 5, 5)."""
    maps = np.atleast_2d(np.asarray(maps, dtype=np.float64))
    mom, _ = engine.simulate_moments(m, maps)
    name = model_name(m)
    for k, label in enumerate(ID_LABELS):
        d = os.path.join(root, "data", "recovered_statistics", name, label)
        os.makedirs(d, exist_ok=True)
        for q, stem in enumerate(MOMENT_FILES):
            writedlm(os.path.join(d, stem + ".txt"), mom[:, k, :, q], append=True)
    return mom
