"""End-to-end rate of abc_simulate_score (page-locked outputs, error matrix copied back) as a function of the pipelining
granularity `simulate_score_sub_batch`, all settings timed back to back on the same box.
Usage: python scripts/bench_e2e_sub.py [particles_per_call] [settings e.g. 8192,4096,2048]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abc_inference_transcription_b200 import AbcEngine, ERR_PARTICLE_MAJOR, PinnedArray, n_params, synthetic_design  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
subs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "8192,4096,2048").split(",")]
gold = os.path.join(ROOT, "tests", "golden")
betas = np.load(os.path.join(gold, "ref_betas.npy"))
z = np.load(os.path.join(gold, "ref_summary_stats.npz"))
eng = AbcEngine(0)
eng.set_design(synthetic_design(betas, n_cells=96, n_pre_cycles=10))
eng.set_data(z["d"], z["se"])
G = z["d"].shape[0]
err = PinnedArray((B, G)); st = PinnedArray((B, 53)); th = [PinnedArray((B, n_params(m))) for m in range(1, 6)]


def step(k):
    eng.accept_reset()
    for m in range(1, 6):
        eng.simulate_score(m, n_trials=B, particle_offset=k * B, seed=20240229, eps=4.8, err_layout=ERR_PARTICLE_MAJOR,
                           out=err.array, theta_out=th[m - 1].array, stats_out=st.array)
    eng.accept_fetch()


step(0)
for rep in range(2):
    for sub in subs:
        eng.set_option("simulate_score_sub_batch", sub)
        t0 = time.perf_counter()
        for k in range(1, 3):
            step(k)
        dt = time.perf_counter() - t0
        print(f"rep {rep} sub_batch>={sub}: {2 * 5 * B / dt:.0f} particles/s end to end", flush=True)
