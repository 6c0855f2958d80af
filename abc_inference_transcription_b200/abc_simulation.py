"""Host-side mirror of scripts/abc_simulation.jl: same entry points (model index m, n_trials, submit), same
simulation file layout; the arithmetic is libabcb200 (CUDA).

Reference              here
fix_params(vary_map,N) fix_params(engine, m, N, ...)            abc_simulation.jl:3-11
abc_sim(theta, ...)    abc_sim(engine, theta, m, rate_name, n)  abc_simulation.jl:13-62 (appends the 5 statistic files)
run_sim(theta, ...)    run_sim(engine, theta, m)                model_realisation.jl:3-37 (returns the 5 arrays)
script tail            run(engine, m, n_trials, submit)         abc_simulation.jl:82-97

File layout under <root>/data/simulations/<model>/ (abc_simulation.jl:47-61, 89-95), tab separated, Julia
float text: progress_<model>_<submit>.txt, sets_<model>_<submit>.txt (1 x P per particle),
s_pulse_* / s_chase_* (2 rows x 5 per particle: means then Fano factors), s_ratios_*, s_mean_corr_*,
s_corr_mean_* (1 x 11 per particle).  Files are opened in append mode like the reference.
"""
import os

import numpy as np

from .jlfmt import writedlm_rows
from .model import _check_m, model_name, n_params, vary_map_for


def fix_params(engine, m, N, particle_offset=0, seed=20240229):
    """N prior draws in log10 units, columns [kon..., koff, alpha..., gamma..., lambda] (abc_simulation.jl:3-11).
    The reference's unseeded `rand(Uniform(a,b))` becomes counter-based Philox draws."""
    return engine.fix_params(_check_m(m), N, particle_offset=particle_offset, seed=seed)


def split_stats(stats):
    """(n, 53) -> s_pulse (n,2,5), s_chase (n,2,5), s_ratios (n,11), s_mean_corr (n,11), s_corr_mean (n,11)"""
    stats = np.asarray(stats).reshape(-1, 53)
    s_pulse = np.stack([stats[:, 0:5], stats[:, 5:10]], axis=1)
    s_chase = np.stack([stats[:, 10:15], stats[:, 15:20]], axis=1)
    return s_pulse, s_chase, stats[:, 20:31], stats[:, 31:42], stats[:, 42:53]


def run_sim(engine, theta, m, particle_offset=0, seed=20240229):
    """run_sim (model_realisation.jl:3-37) for one theta or a batch: returns the five statistic arrays"""
    _, stats, _ = engine.simulate(_check_m(m), theta=np.atleast_2d(theta), particle_offset=particle_offset, seed=seed)
    return split_stats(stats)


def _sim_dir(root, rate_name):
    d = os.path.join(root, "data", "simulations", rate_name)
    os.makedirs(d, exist_ok=True)
    return d


def write_stats(root, rate_name, n, stats):
    """the five appends of abc_sim (abc_simulation.jl:47-61) for a batch of particles"""
    d = _sim_dir(root, rate_name)
    s_pulse, s_chase, s_ratios, s_mean_corr, s_corr_mean = split_stats(stats)
    with open(os.path.join(d, f"s_pulse_{rate_name}_{n}.txt"), "a") as fh:
        writedlm_rows(fh, s_pulse.reshape(-1, 5))        # transpose(s_pulse): row 1 means, row 2 Fano
    with open(os.path.join(d, f"s_chase_{rate_name}_{n}.txt"), "a") as fh:
        writedlm_rows(fh, s_chase.reshape(-1, 5))
    with open(os.path.join(d, f"s_ratios_{rate_name}_{n}.txt"), "a") as fh:
        writedlm_rows(fh, s_ratios)
    with open(os.path.join(d, f"s_mean_corr_{rate_name}_{n}.txt"), "a") as fh:
        writedlm_rows(fh, s_mean_corr)
    with open(os.path.join(d, f"s_corr_mean_{rate_name}_{n}.txt"), "a") as fh:
        writedlm_rows(fh, s_corr_mean)


def abc_sim(engine, theta, m, rate_name=None, n=1, root=".", particle_offset=0, seed=20240229):
    """abc_sim (abc_simulation.jl:13-62): simulate theta (one vector or a batch), append the statistic files"""
    m = _check_m(m)
    rate_name = rate_name or model_name(m)
    _, stats, counters = engine.simulate(m, theta=np.atleast_2d(theta), particle_offset=particle_offset, seed=seed)
    write_stats(root, rate_name, n, stats)
    return stats, counters


def run(engine, m, n_trials, submit=1, root=".", seed=20240229, batch=65536, first_particle=None):
    """The trial loop of abc_simulation.jl:88-97 at batch granularity: for each batch draw the prior, simulate,
    append sets_/s_* and the progress file (last trial index of the batch).  Particle indices are global:
    submit k covers [(k-1)*n_trials, k*n_trials) unless first_particle is given, so several `submit` runs
    (wrapper.jl:62-63) never reuse a Philox stream and can be concatenated like the reference's."""
    m = _check_m(m)
    name = model_name(m)
    assert len(vary_map_for(m)) == 4 and n_params(m) in (5, 9)
    base = (int(submit) - 1) * int(n_trials) if first_particle is None else int(first_particle)
    d = _sim_dir(root, name)
    total = {"n_particles": 0, "n_events": 0, "n_lineages": 0, "ms_simulate": 0.0}
    done = 0
    while done < n_trials:
        nb = min(int(batch), int(n_trials) - done)
        theta, stats, cnt = engine.simulate(m, n_trials=nb, particle_offset=base + done, seed=seed)
        with open(os.path.join(d, f"sets_{name}_{submit}.txt"), "a") as fh:
            writedlm_rows(fh, theta)
        write_stats(root, name, submit, stats)
        done += nb
        with open(os.path.join(d, f"progress_{name}_{submit}.txt"), "a") as fh:
            writedlm_rows(fh, np.array([done], dtype=np.int64))
        for k in total:
            total[k] += cnt[k]
    return total
