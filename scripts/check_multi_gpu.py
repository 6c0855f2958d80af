"""Multi-GPU parity check (run under torchrun, one rank per GPU): every rank simulates and scores its shard of one
global particle range; the library-owned NCCL exchange by gene range (abc_comm_accept_fetch via dist.gather_acceptance) must reproduce, bit for bit, what a single GPU gets
for the whole range (counts, per-gene lists incl. order).  Usage:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/check_multi_gpu.py [n_total]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abc_inference_transcription_b200 import AbcEngine, ERR_NONE, synthetic_design  # noqa: E402
from abc_inference_transcription_b200.dist import gather_acceptance, init_library_comm, shard_range  # noqa: E402

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
gold = os.path.join(ROOT, "tests", "golden")
betas = np.load(os.path.join(gold, "ref_betas.npy"))
z = np.load(os.path.join(gold, "ref_summary_stats.npz"))
eng = AbcEngine(local)
eng.set_design(synthetic_design(betas, n_cells=96, n_pre_cycles=10))
eng.set_data(z["d"], z["se"])
init_library_comm(eng, world, rank, dev)
m, seed = 1, 20240229
lo, hi = shard_range(n_total, rank, world)
theta, stats, _ = eng.simulate(m, n_trials=hi - lo, particle_offset=lo, seed=seed)
eng.accept_reset()
eng.score(stats, eps=4.8, particle_offset=lo, err_layout=ERR_NONE)
res = gather_acceptance(eng, world, dev)
mine = gather_acceptance(eng, world, dev, all_ranks=True)      # every rank keeps the lists of its own gene range
ok = None
if rank == 0:
    ref = AbcEngine(local)
    ref.set_design(synthetic_design(betas, n_cells=96, n_pre_cycles=10))
    ref.set_data(z["d"], z["se"])
    th_all, st_all, _ = ref.simulate(m, n_trials=n_total, particle_offset=0, seed=seed)
    ref.accept_reset()
    _, counts, _ = ref.score(st_all, eps=4.8, particle_offset=0, err_layout=ERR_NONE)
    off, idx, errs = ref.accept_fetch()
    same_stats = np.array_equal(st_all[lo:hi].view(np.uint64), stats.view(np.uint64))
    ok = (same_stats and np.array_equal(res["counts"], counts) and np.array_equal(res["offsets"], off)
          and np.array_equal(res["idx"], idx) and np.array_equal(res["errs"].view(np.uint64), errs.view(np.uint64)))
    g0, g1 = mine["gene_range"]
    a, b = off[g0], off[g1]
    ok = ok and np.array_equal(mine["idx"][a:b], idx[a:b]) and np.array_equal(mine["errs"][a:b].view(np.uint64), errs[a:b].view(np.uint64))
    print(json.dumps({"world": world, "gene_range_rank0": [int(g0), int(g1)], "n_total": n_total, "accepted_pairs": int(len(idx)), "shard_stats_bit_identical": bool(same_stats),
                      "merged_lists_bit_identical": bool(ok)}))
dist.barrier()
dist.destroy_process_group()
if rank == 0 and not ok:
    sys.exit(1)
