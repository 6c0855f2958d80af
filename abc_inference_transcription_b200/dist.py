"""Multi-GPU plumbing: particles shard across ranks (one process per GPU, contiguous global index ranges,
Philox keyed by the global index so any partition gives identical bits); NCCL is used only at the end of a
batch to (i) all-reduce the per-gene acceptance counts and (ii) gather the accepted (gene, particle, err)
tuples -- SURVEY section 8e.  The reference's equivalent is "run several `submit` ids and concatenate the
files by hand" (wrapper.jl:62-63).
"""
import numpy as np


def shard_range(n_total, rank, world):
    """contiguous particle range [lo, hi) of this rank; the first n_total % world ranks get one extra"""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def csr_from_tuples(gene, particle, err, n_genes):
    """per gene: indices sorted by (err asc, index asc) == v[sortperm(err[v])] (accepted_particles.jl:20-24)"""
    order = np.lexsort((particle, err, gene))
    gene, particle, err = gene[order], particle[order], err[order]
    offsets = np.zeros(n_genes + 1, dtype=np.int64)
    np.add.at(offsets, gene.astype(np.int64) + 1, 1)
    return np.cumsum(offsets), particle.astype(np.int64), err


def gather_acceptance(eng, world, device=None, backend_group=None, all_ranks=False):
    """Returns {"counts": (G,) int64 summed over ranks, "offsets"/"idx"/"errs": merged CSR (rank 0; every rank if
    all_ranks), "bytes_d2h": bytes this rank copied to the host}.  world == 1 needs no torch.

    NCCL traffic: one all-reduce of G int64 and three padded all-gathers of the accepted tuples (acceptance rates are
    << 1 %, so this is latency bound).  The merge -- stable sort by (gene, error, particle) -- runs on rank 0's GPU."""
    G = eng.n_genes
    if world <= 1:
        offsets, idx, errs = eng.accept_fetch()           # per-gene order built on the device (csrc/abc_accept.cu)
        counts = np.diff(offsets)
        return {"counts": counts, "offsets": offsets, "idx": idx, "errs": errs,
                "bytes_d2h": offsets.nbytes + idx.nbytes + errs.nbytes}
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(backend_group)
    stream = torch.cuda.current_stream().cuda_stream
    counts = torch.zeros(G, dtype=torch.int64, device=device)
    eng.counts_dev(counts.data_ptr(), stream=stream)
    dist.all_reduce(counts, group=backend_group)                       # (i) G int64, latency bound
    n_local = torch.tensor([eng.accept_total()], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=backend_group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    g = torch.zeros(cap, dtype=torch.int32, device=device)
    p = torch.zeros(cap, dtype=torch.int64, device=device)
    e = torch.zeros(cap, dtype=torch.float64, device=device)
    eng.accept_tuples_dev(g.data_ptr(), p.data_ptr(), e.data_ptr(), cap, stream=stream)
    gl = [torch.empty_like(g) for _ in range(world)]
    pl = [torch.empty_like(p) for _ in range(world)]
    el = [torch.empty_like(e) for _ in range(world)]
    dist.all_gather(gl, g, group=backend_group)                        # (ii) gather-v as padded all-gathers
    dist.all_gather(pl, p, group=backend_group)
    dist.all_gather(el, e, group=backend_group)
    out = {"counts": counts.cpu().numpy(), "offsets": None, "idx": None, "errs": None, "bytes_d2h": G * 8}
    if rank == 0 or all_ranks:
        gene = torch.cat([t[:s] for t, s in zip(gl, sizes)])
        part = torch.cat([t[:s] for t, s in zip(pl, sizes)])
        err = torch.cat([t[:s] for t, s in zip(el, sizes)])
        # per gene: ascending error, ties by ascending particle index (accepted_particles.jl:20-24)
        o = torch.sort(part, stable=True).indices
        o = o[torch.sort(err[o], stable=True).indices]
        o = o[torch.sort(gene[o], stable=True).indices]
        gene, part, err = gene[o], part[o], err[o]
        offsets = torch.zeros(G + 1, dtype=torch.int64, device=device)
        offsets[1:] = torch.cumsum(torch.bincount(gene.to(torch.int64), minlength=G), 0)
        out.update({"offsets": offsets.cpu().numpy(), "idx": part.cpu().numpy(), "errs": err.cpu().numpy()})
        out["bytes_d2h"] += part.numel() * 16 + (G + 1) * 8
    return out
