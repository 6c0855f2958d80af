"""Where the end-to-end step goes: wall time of every host call of one e2e step (bench.py e2e_step) next to the device times
the library reports.  Usage: python scripts/diag_e2e.py [particles_per_call]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abc_inference_transcription_b200 import AbcEngine, ERR_PARTICLE_MAJOR, ERR_NONE, PinnedArray, n_params, synthetic_design  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
gold = os.path.join(ROOT, "tests", "golden")
betas = np.load(os.path.join(gold, "ref_betas.npy"))
z = np.load(os.path.join(gold, "ref_summary_stats.npz"))
eng = AbcEngine(0)
eng.set_design(synthetic_design(betas, n_cells=96, n_pre_cycles=10))
eng.set_data(z["d"], z["se"])
G = z["d"].shape[0]
err = PinnedArray((B, G)); st = PinnedArray((B, 53)); th = [PinnedArray((B, n_params(m))) for m in range(1, 6)]

for layout, name in ((ERR_PARTICLE_MAJOR, "matrix"), (ERR_NONE, "no matrix")):
    for k in range(3):
        t_fix = t_ss = t_fetch = 0.0
        dev_sim = dev_score = 0.0
        t0 = time.perf_counter()
        eng.accept_reset()
        for m in range(1, 6):
            a = time.perf_counter()
            thin = eng.fix_params(m, B, particle_offset=k * B, seed=1, out=th[m - 1].array)
            b = time.perf_counter()
            _, _, _, _, c = eng.simulate_score(m, theta=thin, particle_offset=k * B, seed=1, eps=4.8, err_layout=layout,
                                               out=err.array if layout else None, stats_out=st.array)
            d = time.perf_counter()
            t_fix += b - a; t_ss += d - b
            dev_sim += c["ms_simulate"]; dev_score += c["ms_score"]
        a = time.perf_counter()
        eng.accept_fetch()
        t_fetch = time.perf_counter() - a
        tot = time.perf_counter() - t0
        print(f"{name} step {k}: total {tot*1e3:.1f} ms = fix_params {t_fix*1e3:.1f} + simulate_score {t_ss*1e3:.1f} "
              f"(device: simulate {dev_sim:.1f} + score {dev_score:.1f}) + accept_fetch {t_fetch*1e3:.1f}; "
              f"{5*B/tot:.0f} particles/s", flush=True)
