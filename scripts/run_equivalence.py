"""BASELINE configs[1]-style equivalence run: for every model draw n prior particles (fixed seed), simulate them
with the SSA (hybrid burn-in, n_cells per read-out) AND with the device moment-ODE path (what the reference
computes), score both against the 3419 genes and compare (i) the summary statistics (normalised differences),
(ii) the per-gene acceptance counts and accepted sets, (iii) the per-gene posterior means of the accepted
parameters.  Usage: python scripts/run_equivalence.py [n_per_model] [n_cells] [out.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abc_inference_transcription_b200 import AbcEngine, ERR_NONE, SIM_ODE, MODEL_NAMES, synthetic_design  # noqa: E402
from abc_inference_transcription_b200.posteriors import get_posterior_estimate  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    out = sys.argv[3] if len(sys.argv) > 3 else None
    gold = os.path.join(ROOT, "tests", "golden")
    betas = np.load(os.path.join(gold, "ref_betas.npy"))
    z = np.load(os.path.join(gold, "ref_summary_stats.npz"))
    d, se = z["d"], z["se"]
    ssa, ode = AbcEngine(0), AbcEngine(0)
    ssa.set_design(synthetic_design(betas, n_cells=n_cells, n_pre_cycles=10))
    ode.set_design(synthetic_design(betas, sim_kind=SIM_ODE))
    ssa.set_data(d, se)
    ode.set_data(d, se)
    report = {"n_per_model": n, "n_cells": n_cells, "eps": 4.8, "models": {}}
    for m, name in enumerate(MODEL_NAMES, start=1):
        t0 = time.time()
        theta, s_ssa, cnt = ssa.simulate(m, n_trials=n, particle_offset=0, seed=20240229)
        t_ssa = time.time() - t0
        _, s_ode, _ = ode.simulate(m, theta=theta, particle_offset=0, seed=20240229)
        ok = np.isfinite(s_ssa).all(1) & np.isfinite(s_ode).all(1)
        # (i) statistics: relative difference of the 20 mean/Fano statistics and absolute difference of the
        # ratio / correlation statistics
        rel = np.abs(s_ssa[ok, :20] - s_ode[ok, :20]) / np.maximum(np.abs(s_ode[ok, :20]), 1e-2)
        absd = np.abs(s_ssa[ok, 20:] - s_ode[ok, 20:])
        res = {"finite_fraction": float(ok.mean()), "ssa_particles_per_s": n / t_ssa,
               "events_per_particle": cnt["n_events"] / n,
               "median_rel_diff_mean_fano": float(np.median(rel)), "p90_rel_diff_mean_fano": float(np.quantile(rel, 0.9)),
               "median_abs_diff_ratio_corr": float(np.median(absd)), "p90_abs_diff_ratio_corr": float(np.quantile(absd, 0.9))}
        # (ii) acceptance
        accepted = {}
        for key, eng, st in (("ssa", ssa, s_ssa), ("ode", ode, s_ode)):
            eng.accept_reset()
            _, counts, _ = eng.score(st, eps=4.8, err_layout=ERR_NONE)
            off, idx, _ = eng.accept_fetch()
            accepted[key] = (counts, off, idx)
        c_s, c_o = accepted["ssa"][0].astype(float), accepted["ode"][0].astype(float)
        both = (c_s > 0) | (c_o > 0)
        res["genes_with_accepted_ssa"] = int((c_s > 0).sum())
        res["genes_with_accepted_ode"] = int((c_o > 0).sum())
        res["total_accepted_ssa"] = int(c_s.sum())
        res["total_accepted_ode"] = int(c_o.sum())
        if both.sum() > 10:
            rs, ro = np.argsort(np.argsort(c_s[both])), np.argsort(np.argsort(c_o[both]))
            res["spearman_counts"] = float(np.corrcoef(rs, ro)[0, 1])
            res["log_count_ratio_median"] = float(np.median(np.log10((c_s[both] + 1) / (c_o[both] + 1))))
        # (iii) posterior means of the accepted parameters, genes with >= 20 accepted particles in both
        rich = np.nonzero((c_s >= 20) & (c_o >= 20))[0] + 1
        if len(rich) > 0:
            pm_s = get_posterior_estimate(theta, accepted["ssa"][1], accepted["ssa"][2], rich, "mean")
            pm_o = get_posterior_estimate(theta, accepted["ode"][1], accepted["ode"][2], rich, "mean")
            sd = np.array([theta[accepted["ode"][2][accepted["ode"][1][g - 1]:accepted["ode"][1][g]] - 1].std(0) for g in rich])
            zz = np.abs(pm_s - pm_o) / np.maximum(sd, 1e-6)
            jac = []
            for g in rich:
                a = set(accepted["ssa"][2][accepted["ssa"][1][g - 1]:accepted["ssa"][1][g]])
                b = set(accepted["ode"][2][accepted["ode"][1][g - 1]:accepted["ode"][1][g]])
                jac.append(len(a & b) / len(a | b))
            res.update({"genes_compared": int(len(rich)), "posterior_mean_shift_in_sd_median": float(np.median(zz)),
                        "posterior_mean_shift_in_sd_p90": float(np.quantile(zz, 0.9)), "jaccard_median": float(np.median(jac))})
        report["models"][name] = res
        print(name, json.dumps(res), flush=True)
    if out:
        with open(out, "w") as f:
            json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
