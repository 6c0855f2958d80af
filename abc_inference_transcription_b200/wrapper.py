"""Driver mirroring wrapper.jl sections 2 and 3 (wrapper.jl:59-81): the entry-point variables m, n_trials, submit
stay, the per-particle work runs in libabcb200.

    python -m abc_inference_transcription_b200.wrapper --m 1 --n_trials 250000 --submit 1 --root . \
        --summary_stats data/summary_stats --betas data/capture_efficiencies.txt [--errors] [--sim ode]

Section 2 (`include("scripts/abc_simulation.jl")`): writes data/simulations/<model>/{progress,sets,s_pulse,s_chase,
s_ratios,s_mean_corr,s_corr_mean}_<model>_<submit>.txt.
Section 3 (`compute_errors.jl`, `process_error_files.jl`, `accepted_particles.jl`): scores the simulated statistics
against the data statistics, optionally writes data/errors/error_<model>.txt (+ the gene-major column store) and
appends data/posteriors/particles_<model>.txt.

The design constants the reference derives from the unshipped raw data (age clusters, age_id_distribution; SURVEY
R10) default to the synthetic stand-ins of design.synthetic_design; pass --age_dist / --age_pulse / --age_chase
(text files) to supply the real ones.
"""
import argparse
import os

import numpy as np

from . import SIM_ODE, SIM_SSA, AbcEngine, abc_simulation, accepted_particles, compute_errors
from .design import Design, split_betas
from .jlfmt import readdlm
from .model import model_name


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--m", type=int, required=True, help="model index 1..5 (wrapper.jl:59)")
    ap.add_argument("--n_trials", type=int, default=250000, help="wrapper.jl:61")
    ap.add_argument("--submit", type=int, default=1, help="wrapper.jl:63")
    ap.add_argument("--root", default=".")
    ap.add_argument("--summary_stats", default=None, help="directory with the 14 data/summary_stats/*.txt files")
    ap.add_argument("--betas", default=None, help="data/capture_efficiencies.txt")
    ap.add_argument("--age_dist", default=None, help="5 x 11 age_id_distribution (text); default uniform 0.2")
    ap.add_argument("--age_pulse", default=None)
    ap.add_argument("--age_chase", default=None)
    ap.add_argument("--sim", choices=["ssa", "ode"], default="ssa")
    ap.add_argument("--n_cells", type=int, default=96)
    ap.add_argument("--n_pre_cycles", type=int, default=10)
    ap.add_argument("--seed", type=int, default=20240229)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--errors", action="store_true", help="also write error_<model>.txt and the column store")
    ap.add_argument("--eps", type=float, default=4.8, help="accepted_particles.jl:10")
    ap.add_argument("--skip_simulation", action="store_true")
    args = ap.parse_args(argv)

    name = model_name(args.m)
    betas = np.loadtxt(args.betas) if args.betas else None
    des = Design(n_cells=args.n_cells, n_pre_cycles=args.n_pre_cycles, sim_kind=SIM_SSA if args.sim == "ssa" else SIM_ODE,
                 downsampling=betas is not None)
    if betas is not None:
        bp, ap_, bc, ac = split_betas(betas)
        des.betas_pulse, des.betas_chase = bp, bc
        des.age_pulse = np.loadtxt(args.age_pulse, dtype=np.int32) if args.age_pulse else ap_
        des.age_chase = np.loadtxt(args.age_chase, dtype=np.int32) if args.age_chase else ac
    if args.age_dist:
        des.age_dist = readdlm(args.age_dist)
    with AbcEngine(args.device) as eng:
        eng.set_design(des)
        # ---- section 2 ----
        if not args.skip_simulation:
            tot = abc_simulation.run(eng, args.m, args.n_trials, submit=args.submit, root=args.root, seed=args.seed)
            print(f"[wrapper] simulated {tot['n_particles']} particles of model '{name}' "
                  f"({tot['n_events']:.3g} SSA events, {tot['ms_simulate'] / 1e3:.1f} s on device)")
        # ---- section 3 ----
        if args.summary_stats:
            data14 = compute_errors.load_summary_stats(args.summary_stats, ".txt")
            # wrapper.jl:42 order -> compute_trunc_errors' argument order (compute_errors.jl:45-48)
            (pm, pf, pms, pfs, cm, cf, cms, cfs, rd, rs, md, ms, cd, cs) = data14
            sim7 = compute_errors.load_s_data(os.path.join(args.root, "data", "simulations"), name, f"_{args.submit}.txt")
            eng.accept_reset()
            first = (args.submit - 1) * args.n_trials
            err = compute_errors.compute_trunc_errors(
                eng, pm, pms, pf, pfs, cm, cms, cf, cfs, rd, rs, md, ms, cd, cs, *sim7, name,
                out_dir=os.path.join(args.root, "data", "errors") if args.errors else None, eps=args.eps, particle_offset=first)
            if args.errors:
                compute_errors.process_error_files(os.path.join(args.root, "data", "errors"),
                                                   os.path.join(args.root, "data", "errors"), model_names=(name,))
            offsets, idx = accepted_particles.accepted_from_engine(eng)
            accepted_particles.write_particles(args.root, name, offsets, idx)
            print(f"[wrapper] scored {err.shape[0]} particles x {err.shape[1]} genes; accepted pairs: {len(idx)}; "
                  f"genes with a posterior: {int((np.diff(offsets) > 0).sum())}")


if __name__ == "__main__":
    main()
