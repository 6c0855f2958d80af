"""Host-side mirror of scripts/accepted_particles.jl: eps = 4.8; per gene and model the particles with
err <= eps ordered by ascending error (stable: ties by index), 1-based, one tab-separated line per gene in
data/posteriors/particles_<model>.txt; a gene without accepted particles gets the line "0"
(accepted_particles.jl:10-32).  The selection itself runs fused in the scoring kernel."""
import os

import numpy as np

EPS = 4.8                                                                   # accepted_particles.jl:10


def accepted_from_engine(engine):
    """CSR (offsets, idx) of the particles accepted by the abc_score calls since the last accept_reset"""
    offsets, idx, _ = engine.accept_fetch()
    return offsets, idx


def write_particles(root, model_name, offsets, idx):
    """append the G lines of particles_<model>.txt (accepted_particles.jl:23-30)"""
    d = os.path.join(root, "data", "posteriors")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, f"particles_{model_name}.txt"), "a") as fh:
        for g in range(len(offsets) - 1):
            v = idx[offsets[g]:offsets[g + 1]]
            fh.write(("\t".join(str(int(x)) for x in v) if len(v) else "0") + "\n")


def accepted_particles(engine, stats_by_model, root=".", eps=EPS, particle_offset=0):
    """For every model (dict name -> (n,53) statistics) score against the engine's data statistics and write
    particles_<model>.txt.  Returns {name: (offsets, idx, counts)}."""
    out = {}
    for name, stats in stats_by_model.items():
        engine.accept_reset()
        _, counts, _ = engine.score(stats, eps=eps, particle_offset=particle_offset, err_layout=0)
        offsets, idx = accepted_from_engine(engine)
        write_particles(root, name, offsets, idx)
        out[name] = (offsets, idx, counts)
    return out


def read_particles(path):
    """parse particles_<model>.txt back into a list of index arrays ([] for the "0" sentinel)"""
    res = []
    with open(path) as fh:
        for line in fh:
            v = np.array([int(t) for t in line.split()], dtype=np.int64)
            res.append(v[:0] if (len(v) == 1 and v[0] == 0) else v)
    return res
