"""Reproducibility probe: the same particles simulated twice (and in different batch splits) must give identical bits."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abc_inference_transcription_b200 import AbcEngine, n_params, synthetic_design
betas = np.load(os.path.join(ROOT, "tests", "golden", "ref_betas.npy"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
dev = torch.device("cuda", 0)
eng = AbcEngine(0)
eng.set_design(synthetic_design(betas, n_cells=96, n_pre_cycles=10))
for m in (1, 3, 5):
    P = n_params(m)
    outs = []
    for split in (1, 1, 2, 5):
        th = torch.empty((n, P), dtype=torch.float64, device=dev)
        st = torch.empty((n, 53), dtype=torch.float64, device=dev)
        step = -(-n // split)
        for c0 in range(0, n, step):
            nb = min(step, n - c0)
            eng.simulate_dev(m, nb, th[c0:].data_ptr(), st[c0:].data_ptr(), particle_offset=c0, seed=7, prior_supplied=False)
        torch.cuda.synchronize()
        outs.append(st.cpu().numpy().view(np.uint64))
    for k in range(1, len(outs)):
        diff = (outs[k] != outs[0]).any(1)
        print(f"m={m} run {k} vs 0: {int(diff.sum())} particles differ", np.nonzero(diff)[0][:8], flush=True)
        if diff.any():
            i = int(np.nonzero(diff)[0][0])
            cols = np.nonzero(outs[k][i] != outs[0][i])[0]
            print("   first differing particle", i, "stat columns", cols[:10], outs[k][i].view(np.float64)[cols[:4]], outs[0][i].view(np.float64)[cols[:4]])
