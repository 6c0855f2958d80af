"""Full prior sweep of one model on one GPU (BASELINE configs[2] per-GPU share / SURVEY R5: M = 10^6 per model):
simulate M particles with the SSA in device-resident chunks, score every chunk against the 3419 genes with the fused
eps-acceptance (no error matrix), fetch the per-gene accepted lists.  Usage: python scripts/run_full_model.py [m] [M] [out.json]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abc_inference_transcription_b200 import AbcEngine, ERR_NONE, n_params, synthetic_design  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
out = sys.argv[3] if len(sys.argv) > 3 else None
gold = os.path.join(ROOT, "tests", "golden")
betas = np.load(os.path.join(gold, "ref_betas.npy"))
z = np.load(os.path.join(gold, "ref_summary_stats.npz"))
eng = AbcEngine(0)
eng.set_design(synthetic_design(betas, n_cells=96, n_pre_cycles=10))
eng.set_data(z["d"], z["se"])
dev = torch.device("cuda", 0)
chunk = 65536
th = torch.empty((chunk, n_params(m)), dtype=torch.float64, device=dev)
st = torch.empty((chunk, 53), dtype=torch.float64, device=dev)
eng.accept_reset()
torch.cuda.synchronize()
t0 = time.time()
events = 0
for b0 in range(0, M, chunk):
    nb = min(chunk, M - b0)
    eng.simulate_dev(m, nb, th.data_ptr(), st.data_ptr(), particle_offset=b0, seed=20240229)
    eng.score_dev(st.data_ptr(), nb, eps=4.8, particle_offset=b0, err_layout=ERR_NONE)
    events += eng.counters()["n_events"]
torch.cuda.synchronize()
t1 = time.time()
offsets, idx, errs = eng.accept_fetch()
t2 = time.time()
cnt = np.diff(offsets)
res = {"model": m, "M": M, "seconds_simulate_score": t1 - t0, "seconds_fetch_sort": t2 - t1,
       "particles_per_s": M / (t1 - t0), "events": int(events), "accepted_pairs": int(len(idx)),
       "genes_with_posterior": int((cnt > 0).sum()), "median_accepted_per_gene": float(np.median(cnt)),
       "max_accepted_per_gene": int(cnt.max()), "index_range": [int(idx.min()), int(idx.max())] if len(idx) else None}
print(json.dumps(res))
if out:
    json.dump(res, open(out, "w"), indent=1)
