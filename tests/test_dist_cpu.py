"""world_size-2 gloo test of the multi-GPU host logic: ranks own contiguous particle shards, acceptance counts
are all-reduced, accepted tuples gathered and merged; the result must equal the single-process answer.
(The scoring itself is the CPU oracle here -- this exercises the partition / gather / merge plumbing only.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from abc_inference_transcription_b200.dist import csr_from_tuples, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_summary_stats.npz"))
    d, se = z["d"][:200], z["se"][:200]
    rng = np.random.default_rng(0)                                # same stream on every rank: global particle set
    stats = d[rng.integers(0, 200, n)] * np.exp(rng.normal(0, 0.2, (n, 53)))
    lo, hi = shard_range(n, rank, world)
    err = oracle.compute_trunc_errors(stats[lo:hi], d, se)
    pi, gi = np.nonzero(err <= 4.8)
    counts = torch.from_numpy((err <= 4.8).sum(0).astype(np.int64))
    dist.all_reduce(counts)
    mine = (gi.astype(np.int32), (pi + lo + 1).astype(np.int64), err[pi, gi])
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    gene = np.concatenate([g[0] for g in gathered])
    part = np.concatenate([g[1] for g in gathered])
    e = np.concatenate([g[2] for g in gathered])
    off, idx, _ = csr_from_tuples(gene, part, e, 200)
    if rank == 0:
        full = oracle.compute_trunc_errors(stats, d, se)
        ok = np.array_equal(counts.numpy(), (full <= 4.8).sum(0)) and off[-1] > 0
        for g in range(200):
            ok = ok and np.array_equal(idx[off[g]:off[g + 1]], oracle.accept_gene(full[:, g], 4.8))
        np.save(out, np.array([int(ok)]))
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(tmp_path):
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(2, 29533, 301, out), nprocs=2, join=True)
    assert np.load(out)[0] == 1
