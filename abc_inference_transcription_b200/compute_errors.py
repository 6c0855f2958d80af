"""Host-side mirror of scripts/compute_errors.jl and scripts/process_error_files.jl.

Reference                          here
get_mean_subset / get_ff_subset    same names                         compute_errors.jl:1-15
load_s_data(path, model, ext)      same                               compute_errors.jl:17-28
load_summary_stats(path, ext)      same (called at wrapper.jl:42, defined nowhere in the reference: SURVEY R6)
nlsqerror_part / compute_trunc_errors   compute_trunc_errors(engine, 14 data matrices, 7 sim matrices, model_name)
                                                                      compute_errors.jl:30-70  -> libabcb200 abc_score
process_error_files.jl:3-7         process_error_files(...): text matrix -> gene-major column store
"""
import os

import numpy as np

from . import _lib
from .jlfmt import readdlm, writedlm_rows

STAT_ORDER = ["pulse_mean", "pulse_ff", "chase_mean", "chase_ff", "ratio", "mean_corr", "corr_mean"]


def get_mean_subset(data):
    """odd rows (1-based) of the 2-rows-per-particle files: the means (compute_errors.jl:1-7)"""
    data = np.asarray(data)
    return data[0::2].copy()


def get_ff_subset(data):
    """even rows: the Fano factors (compute_errors.jl:9-15)"""
    data = np.asarray(data)
    return data[1::2].copy()


def load_s_data(path, model_name, ext):
    """compute_errors.jl:17-28.  The reference reads files named without the _<submit> suffix (concatenated by
    hand); pass ext = "_1.txt" to read one submit's files directly."""
    base = os.path.join(path, model_name)
    s_pulse = readdlm(os.path.join(base, f"s_pulse_{model_name}{ext}"))
    s_chase = readdlm(os.path.join(base, f"s_chase_{model_name}{ext}"))
    return (get_mean_subset(s_pulse), get_ff_subset(s_pulse), get_mean_subset(s_chase), get_ff_subset(s_chase),
            readdlm(os.path.join(base, f"s_ratios_{model_name}{ext}")),
            readdlm(os.path.join(base, f"s_mean_corr_{model_name}{ext}")),
            readdlm(os.path.join(base, f"s_corr_mean_{model_name}{ext}")))


def load_summary_stats(path, ext=".txt"):
    """the 14 data matrices in the order wrapper.jl:42 unpacks them"""
    def g(name):
        return readdlm(os.path.join(path, name + ext))

    return (g("pulse_mean"), g("pulse_ff"), g("pulse_mean_se"), g("pulse_ff_se"), g("chase_mean"), g("chase_ff"),
            g("chase_mean_se"), g("chase_ff_se"), g("ratio_data"), g("ratio_se"), g("mean_corr_data"), g("mean_corr_se"),
            g("corr_mean_data"), g("corr_mean_se"))


def pack_data(pulse_mean, pulse_mean_se, pulse_ff, pulse_ff_se, chase_mean, chase_mean_se, chase_ff, chase_ff_se,
              ratio_data, ratio_se, mean_corr_data, mean_corr_se, corr_mean_data, corr_mean_se):
    """14 matrices in compute_trunc_errors' argument order (compute_errors.jl:45-48) -> d (G,53), se (G,53)"""
    d = np.concatenate([pulse_mean, pulse_ff, chase_mean, chase_ff, ratio_data, mean_corr_data, corr_mean_data], axis=1)
    se = np.concatenate([pulse_mean_se, pulse_ff_se, chase_mean_se, chase_ff_se, ratio_se, mean_corr_se, corr_mean_se], axis=1)
    return np.ascontiguousarray(d, dtype=np.float64), np.ascontiguousarray(se, dtype=np.float64)


def pack_stats(s_pulse_mean, s_pulse_ff, s_chase_mean, s_chase_ff, s_ratios, s_mean_corr, s_corr_mean):
    return np.ascontiguousarray(np.concatenate([s_pulse_mean, s_pulse_ff, s_chase_mean, s_chase_ff, s_ratios,
                                                s_mean_corr, s_corr_mean], axis=1), dtype=np.float64)


def compute_trunc_errors(engine, pulse_mean, pulse_mean_se, pulse_ff, pulse_ff_se, chase_mean, chase_mean_se, chase_ff,
                         chase_ff_se, ratio_data, ratio_se, mean_corr_data, mean_corr_se, corr_mean_data, corr_mean_se,
                         s_pulse_mean, s_pulse_ff, s_chase_mean, s_chase_ff, s_ratios, s_mean_corr, s_corr_mean, model_name,
                         out_dir=None, eps=4.8, particle_offset=0):
    """compute_errors.jl:45-70, same argument order.  Returns the M x G error matrix (rows = particles); when
    out_dir is given appends it to <out_dir>/error_<model>.txt exactly like the reference's writedlm rows.
    The eps-acceptance is fused into the same kernel: fetch it with accepted_particles.accepted_from_engine."""
    d, se = pack_data(pulse_mean, pulse_mean_se, pulse_ff, pulse_ff_se, chase_mean, chase_mean_se, chase_ff, chase_ff_se,
                      ratio_data, ratio_se, mean_corr_data, mean_corr_se, corr_mean_data, corr_mean_se)
    engine.set_data(d, se)
    stats = pack_stats(s_pulse_mean, s_pulse_ff, s_chase_mean, s_chase_ff, s_ratios, s_mean_corr, s_corr_mean)
    err, counts, _ = engine.score(stats, eps=eps, particle_offset=particle_offset, err_layout=_lib.ERR_PARTICLE_MAJOR)
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, f"error_{model_name}.txt"), "a") as fh:
            writedlm_rows(fh, err)
    return err


def process_error_files(text_dir, out_dir, model_names=("const", "const_const", "kon", "alpha", "gamma")):
    """process_error_files.jl:3-7: text M x G matrix -> one column per gene (x1..xG).  The reference stores a
    JDF.jl directory (third-party binary format, unpinned: SURVEY 8c); here the same columns are one
    gene-major .npy (row g-1 == column x<g>), memory-mappable per gene."""
    os.makedirs(out_dir, exist_ok=True)
    out = {}
    for name in model_names:
        p = os.path.join(text_dir, f"error_{name}.txt")
        if not os.path.exists(p):
            continue
        cols = np.ascontiguousarray(readdlm(p).T)
        np.save(os.path.join(out_dir, f"error_{name}.npy"), cols)
        out[name] = cols.shape
    return out


def load_error_column(out_dir, model_name, g):
    """f["x<g>"] of the reference's JDFFile (accepted_particles.jl:14-18); g is 1-based"""
    return np.load(os.path.join(out_dir, f"error_{model_name}.npy"), mmap_mode="r")[g - 1]
