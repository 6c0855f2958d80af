"""Diagnostic: z-scores of SSA sample moments against the moment ODEs (the computation of
tests/test_gpu_parity.py::test_ssa_moments_match_moment_odes) for every SSA mode and several Philox streams.
Usage: python scripts/diag_moments_z.py [n_cells] [n_streams]"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from abc_inference_transcription_b200 import AbcEngine, split_betas, synthetic_design  # noqa: E402
from test_gpu_parity import DEMO  # noqa: E402

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n_streams = int(sys.argv[2]) if len(sys.argv) > 2 else 6
betas = np.load(os.path.join(ROOT, "tests", "golden", "ref_betas.npy"))
eng = AbcEngine(0)
eng.set_design(synthetic_design(betas, n_cells=n_cells, n_pre_cycles=10))
READ = [(5, 0), (6, 2), (9, 4), (0, 3)]
NAMES = ["mean_u", "mean_l", "var_u", "cov_ul", "var_l"]
for m in (1, 3, 5):
    od = oracle.make_design(iv_index=1, downsampling=True, betas=split_betas(betas), rtol=1e-9)
    od_raw = oracle.make_design(iv_index=1, downsampling=False, rtol=1e-9)
    _, mom_ds = oracle.run_sim(DEMO[m], m, od)
    mom_raw, _ = oracle.run_part_sim(DEMO[m], m, od_raw)
    for mode in (0, 1, 2):
        eng.set_option("ssa_hybrid_burnin", mode)
        allz = []
        for pidx in range(5, 5 + n_streams):
            zs, lab = [], []
            for c, a in READ:
                x = eng.ssa_cells(m, DEMO[m], particle_index=pidx, cond=c, age=a, seed=3).astype(np.float64)
                for tag, (u, l), ref in [("raw", (x[0], x[1]), mom_raw[c, a]), ("ds", (x[2], x[3]), mom_ds[c, a])]:
                    n = len(u)
                    for k, (sample, target) in enumerate([(u, ref[0]), (l, ref[1])]):
                        zs.append((sample.mean() - target) / (sample.std(ddof=1) / np.sqrt(n) + 1e-12)); lab.append((c, a, tag, NAMES[k]))
                    for k, (xs, ys, target) in enumerate([(u, u, ref[2]), (u, l, ref[3]), (l, l, ref[4])]):
                        p = (xs - xs.mean()) * (ys - ys.mean())
                        zs.append((p.sum() / (n - 1) - target) / (p.std(ddof=1) / np.sqrt(n) + 1e-12)); lab.append((c, a, tag, NAMES[2 + k]))
            zs = np.array(zs)
            allz.append(zs)
            i = int(np.abs(zs).argmax())
            print(f"m={m} mode={mode} stream={pidx}: max|z|={abs(zs[i]):.2f} at {lab[i]}  mean z^2={np.mean(zs**2):.2f}", flush=True)
        allz = np.array(allz)
        print(f"m={m} mode={mode}: mean z over streams per statistic (bias if |.|*sqrt(k) large):")
        mz = allz.mean(0) * np.sqrt(len(allz))
        j = np.argsort(-np.abs(mz))[:4]
        print("   ", [(lab[t], round(float(mz[t]), 2)) for t in j], flush=True)
