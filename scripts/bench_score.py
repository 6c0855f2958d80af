"""Scoring-kernel micro-benchmark: realistic statistics (prior draws through the device ODE path), one large
resident batch, CUDA events.
Usage: python scripts/bench_score.py [n_particles] [layout 0|1|2] [reference_kernel 0|1] [tile_kernel 1|0] [overlap 1|0] [sub_batches] [mma_filter 0|1|2]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abc_inference_transcription_b200 import AbcEngine, SIM_ODE, synthetic_design  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
layout = int(sys.argv[2]) if len(sys.argv) > 2 else 2
refk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
tilek = int(sys.argv[4]) if len(sys.argv) > 4 else 1
overlap = int(sys.argv[5]) if len(sys.argv) > 5 else 1
nsb = int(sys.argv[6]) if len(sys.argv) > 6 else 0
mma = int(sys.argv[7]) if len(sys.argv) > 7 else 0
gold = os.path.join(ROOT, "tests", "golden")
betas = np.load(os.path.join(gold, "ref_betas.npy"))
z = np.load(os.path.join(gold, "ref_summary_stats.npz"))
eng = AbcEngine(0)
eng.set_design(synthetic_design(betas, sim_kind=SIM_ODE))
eng.set_data(z["d"], z["se"])
eng.set_option("score_reference_kernel", refk)
eng.set_option("score_tile_kernel", tilek)
eng.set_option("score_overlap", overlap)
eng.set_option("score_sub_batches", nsb)
eng.set_option("score_mma_filter", mma)
G = z["d"].shape[0]
dev = torch.device("cuda", 0)
per = n // 5
stats = torch.empty((5 * per, 53), dtype=torch.float64, device=dev)
for m in range(1, 6):
    th = torch.empty((per, 9), dtype=torch.float64, device=dev)
    eng.simulate_dev(m, per, th.data_ptr(), stats[(m - 1) * per:].data_ptr(), particle_offset=0, seed=1)
torch.cuda.synchronize()
stats = stats[torch.randperm(5 * per, device=dev)].contiguous()
n = 5 * per
err = torch.empty((n, G), dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ms = []
for it in range(5):
    eng.accept_reset()
    flush.fill_(it)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.score_dev(stats.data_ptr(), n, eps=4.8, err_layout=layout, d_err_ptr=err.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
t = float(np.median(ms[1:])) / 1e3
print(f"n={n} layout={layout} ref_kernel={refk} tile_kernel={tilek} overlap={overlap} sub_batches={nsb} mma_filter={mma}: {t*1e3:.3f} ms  {n*27776/t/1e9:.1f} GB/s algorithmic  "
      f"{n*G/t/1e9:.2f} Gpairs/s  accepted={eng.accept_total()}  frac<10={(err < 10).float().mean().item():.4f}")
