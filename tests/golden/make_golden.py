"""Builds the committed golden fixtures from the reference's shipped result files.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

Outputs (all numpy, float64 unless noted):
  ref_map_sets.npz        the 2503 MAP parameter vectors  data/posterior_estimates/map_sets_<model>.txt
                          + the gene index of each row     data/model_selection/<model>_genes.txt
  ref_recovered.npz       known answers of syntheticdata() (scripts/recover_statistics.jl:49-68) for ALL rows:
                          data/recovered_statistics/<model>/<cond>/<moment>.txt
                          -> array [row, cond(11), age(5), moment(5)] + the row indices used
  ref_summary_stats.npz   data/summary_stats/*.txt  -> d[G,53], se[G,53] in the scoring order
                          pulse_mean, pulse_ff, chase_mean, chase_ff, ratio, mean_corr, corr_mean
  ref_betas.npy           data/capture_efficiencies.txt (5422 cells; rows 1-2364 chase, 2365-5422 pulse)
"""
import os
import numpy as np

REF = "/root/reference/data/"
OUT = os.path.dirname(os.path.abspath(__file__))
MODELS = ["const", "const_const", "kon", "alpha", "gamma"]
LABELS = ["pulse_15", "pulse_30", "pulse_45", "pulse_60", "pulse_120", "pulse_180",
          "chase_0", "chase_60", "chase_120", "chase_240", "chase_360"]
MOMENTS = ["mean_u", "mean_l", "var_u", "cov_ul", "var_l"]
MAX_ROWS = 10**9  # per model: all rows (1844 / 478 / 101 / 15 / 65 = 2503)


def main():
    maps, rec = {}, {}
    for name in MODELS:
        th = np.atleast_2d(np.loadtxt(REF + f"posterior_estimates/map_sets_{name}.txt"))
        genes = np.loadtxt(REF + f"model_selection/{name}_genes.txt", dtype=np.int64).reshape(-1)
        assert len(genes) == len(th)
        maps[f"theta_{name}"] = th
        maps[f"genes_{name}"] = genes
        n = len(th)
        rows = np.unique(np.linspace(0, n - 1, min(n, MAX_ROWS)).round().astype(np.int64))
        g = np.stack([np.stack([np.atleast_2d(np.loadtxt(REF + f"recovered_statistics/{name}/{lab}/{f}.txt"))[rows]
                                for f in MOMENTS], axis=-1) for lab in LABELS], axis=1)
        rec[f"rows_{name}"] = rows
        rec[f"moments_{name}"] = g
    np.savez_compressed(os.path.join(OUT, "ref_map_sets.npz"), **maps)
    np.savez_compressed(os.path.join(OUT, "ref_recovered.npz"), **rec)

    order = ["pulse_mean", "pulse_ff", "chase_mean", "chase_ff", "ratio", "mean_corr", "corr_mean"]
    dfile = {"ratio": "ratio_data", "mean_corr": "mean_corr_data", "corr_mean": "corr_mean_data"}
    d = np.concatenate([np.loadtxt(REF + f"summary_stats/{dfile.get(k, k)}.txt") for k in order], axis=1)
    se = np.concatenate([np.loadtxt(REF + f"summary_stats/{k}_se.txt") for k in order], axis=1)
    assert d.shape == se.shape == (3419, 53)
    np.savez_compressed(os.path.join(OUT, "ref_summary_stats.npz"), d=d, se=se)
    np.save(os.path.join(OUT, "ref_betas.npy"), np.loadtxt(REF + "capture_efficiencies.txt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
