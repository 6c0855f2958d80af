"""SSA-kernel micro-benchmark: one simulate_dev launch per model on a resident batch, events/s from the in-kernel
counters and the library's CUDA-event timing.  Usage: python scripts/bench_ssa.py [n_particles] [models e.g. 12345] [n_cells] [n_pre] [corner|prior] [mode] [adaptive]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abc_inference_transcription_b200 import AbcEngine, n_params, synthetic_design  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
models = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "12345")]
n_cells = int(sys.argv[3]) if len(sys.argv) > 3 else 96
n_pre = int(sys.argv[4]) if len(sys.argv) > 4 else 10
corner = len(sys.argv) > 5 and sys.argv[5] == "corner"
mode = int(sys.argv[6]) if len(sys.argv) > 6 else 2        # ssa_hybrid_burnin   # BASELINE configs[4]: kon,koff,alpha ~ U(2,3), gamma ~ U(1,2)
adaptive = int(sys.argv[7]) if len(sys.argv) > 7 else 2    # ssa_adaptive_burnin
betas = np.load(os.path.join(ROOT, "tests", "golden", "ref_betas.npy"))
eng = AbcEngine(0)
eng.set_design(synthetic_design(betas, n_cells=n_cells, n_pre_cycles=n_pre))
eng.set_option("ssa_hybrid_burnin", mode)
eng.set_option("ssa_adaptive_burnin", adaptive)
dev = torch.device("cuda", 0)
st = torch.empty((n, 53), dtype=torch.float64, device=dev)
tot_ev = tot_ms = tot_dr = 0.0
for rep in range(2):
    for m in models:
        th = torch.empty((n, n_params(m)), dtype=torch.float64, device=dev)
        if corner:
            P = n_params(m)
            g = torch.Generator(device=dev).manual_seed(1 + rep)
            u = torch.rand((n, P), dtype=torch.float64, device=dev, generator=g)
            lo = torch.full((P,), 2.0, dtype=torch.float64, device=dev)
            ngam = 5 if m == 5 else 1
            lo[P - 1 - ngam:P - 1] = 1.0
            th.copy_(lo + u)
            th[:, P - 1] = -0.7 + 0.7 * u[:, P - 1]
        eng.simulate_dev(m, n, th.data_ptr(), st.data_ptr(), particle_offset=rep * n, seed=20240229, prior_supplied=corner)
        c = eng.counters()
        if rep == 1:
            tot_ev += c["n_events"]; tot_ms += c["ms_simulate"]; tot_dr += c["n_draws"]
            print(f"m={m} n={n}: {c['ms_simulate']:.1f} ms  {c['n_events']/c['ms_simulate']/1e6:.1f} Gev/s  "
                  f"{n/c['ms_simulate']*1e3:.0f} particles/s  events/particle {c['n_events']/n:.3g}")
print(f"mode {mode} adaptive {adaptive} {'corner' if corner else 'prior'}: {len(models)*n/tot_ms*1e3:.0f} particles/s, draws/particle {tot_dr/len(models)/n:.4g}, "
      f"frac(24/draw) {tot_dr/tot_ms*1e3*24/37.225e12:.3f}")
print(f"total: {tot_ev/tot_ms/1e6:.1f} Gev/s  issue-roofline frac: {tot_ev/tot_ms*1e3*24/37.225e12:.3f} at 24 lane-instr per telegraph draw "
      f"(modes 1-2), {tot_ev/tot_ms*1e3*64/37.225e12:.3f} at 64 per six-channel event (mode 0)")
