"""Multi-GPU plumbing for ONE PROCESS PER GPU (torchrun): particles shard across ranks by contiguous global index ranges
(Philox is keyed by the global index, so any partition gives identical bits); after a batch the per-gene acceptance counts
are summed and the accepted (gene, particle, err) tuples are exchanged BY GENE RANGE so that every rank orders the lists of
G / world genes of equal tuple mass -- SURVEY section 8e.  The communicator, the exchange (ncclSend / ncclRecv all-to-all)
and the ordering (csrc/abc_accept.cu) live in the library (csrc/abc_multi.cu, abc_comm_*); torch.distributed is used only
to hand the 128-byte NCCL unique id from rank 0 to the other ranks.  The one-process front end over the same core is
AbcMulti (abc_multi_*).  The reference's equivalent is "run several `submit` ids and concatenate the files by hand"
(wrapper.jl:62-63).
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


def init_library_comm(eng, world, rank, device=None, group=None):
    """give the context a library-owned NCCL communicator: rank 0 draws the unique id, torch.distributed broadcasts it"""
    if world <= 1:
        eng.comm_init(b"", 1, 0)
        return
    import torch
    import torch.distributed as dist
    from .engine import comm_unique_id
    if rank == 0:
        uid = torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8).clone()
    else:
        uid = torch.zeros(128, dtype=torch.uint8)
    if device is not None:
        uid = uid.to(device)
    dist.broadcast(uid, src=0, group=group)
    eng.comm_init(bytes(uid.cpu().numpy().tobytes()), world, rank)


def gather_acceptance(eng, world, device=None, backend_group=None, all_ranks=False):
    """Returns {"counts": (G,) int64 summed over ranks, "offsets"/"idx"/"errs": merged CSR (rank 0; with all_ranks every rank
    holds its own gene range), "gene_range", "bytes_d2h": bytes this rank copied to the host}.  world == 1 needs no
    communicator.  The context must have been attached with init_library_comm."""
    if world <= 1:
        offsets, idx, errs = eng.accept_fetch()           # per-gene order built on the device (csrc/abc_accept.cu)
        counts = np.diff(offsets)
        return {"counts": counts, "offsets": offsets, "idx": idx, "errs": errs, "gene_range": (0, eng.n_genes),
                "bytes_d2h": offsets.nbytes + idx.nbytes + errs.nbytes}
    offsets, idx, errs, grange = eng.comm_accept_fetch(root=-1 if all_ranks else 0)
    counts = np.diff(offsets)
    held = 0 if idx is None else (int(offsets[grange[1]] - offsets[grange[0]]) if all_ranks else int(offsets[-1]))
    return {"counts": counts, "offsets": offsets, "idx": idx, "errs": errs, "gene_range": grange,
            "bytes_d2h": counts.nbytes + held * 16}
