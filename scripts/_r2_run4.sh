cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cat > /tmp/ode_probe.py <<'PY'
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.environ["GRAFT_REPO_ROOT"])
from abc_inference_transcription_b200 import AbcEngine, SIM_ODE, n_params, synthetic_design
betas = np.load("tests/golden/ref_betas.npy")
eng = AbcEngine(0); eng.set_design(synthetic_design(betas, sim_kind=SIM_ODE))
dev = torch.device("cuda", 0)
for n in (8192, 32768, 131072):
    st = torch.empty((n, 53), dtype=torch.float64, device=dev)
    for rep in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        for m in range(1, 6):
            th = torch.empty((n, n_params(m)), dtype=torch.float64, device=dev)
            eng.simulate_dev(m, n, th.data_ptr(), st.data_ptr(), particle_offset=rep * n, seed=3)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"n={n}: {5*n/dt:.0f} particles/s (5 models, {dt*1e3:.1f} ms)", flush=True)
PY
python /tmp/ode_probe.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:abc_ode -c 9 --csv --log-file gpurun_out/r2_ode_launches.csv python /tmp/ode_probe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_ode_launches.csv') if not l.startswith('=='))]
h=rows[0]
for r in rows[1:10]:
    print(r[h.index('Kernel Name')][:40], r[h.index('Grid Size')], r[-1], r[-2])
PY
timeout 600 python -m pytest tests/test_gpu_ode.py -m gpu -x -q 2>&1 | tail -2
