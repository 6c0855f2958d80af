"""GPU tests of the device moment-ODE path (sim_kind = ABC_SIM_ODE): the computation scripts/model.jl performs,
checked against the reference's shipped recovered_statistics and against the CPU oracle."""
import os

import numpy as np
import pytest

import oracle
from abc_inference_transcription_b200 import AbcEngine, SIM_ODE, split_betas, synthetic_design

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = ["const", "const_const", "kon", "alpha", "gamma"]


@pytest.fixture(scope="module")
def betas():
    return np.load(os.path.join(GOLD, "ref_betas.npy"))


def test_device_odes_reproduce_reference_recovered_statistics(betas):
    """run_part_sim (recover_statistics.jl:1-11: iv[1] = 1/2, no downsampling) on every shipped MAP row of the
    fixture: same tolerance as the oracle-vs-golden test (the goldens carry CVODE 1e-3 noise), and the device
    path agrees with the tight-tolerance oracle to 1e-4."""
    maps = np.load(os.path.join(GOLD, "ref_map_sets.npz"))
    rec = np.load(os.path.join(GOLD, "ref_recovered.npz"))
    des = synthetic_design(betas, sim_kind=SIM_ODE, downsampling=False, iv=np.array([0.5, 0, 0, 0, 0, 0, 0, 0, 0.0]),
                           ode_rtol=1e-7, ode_atol=1e-10)
    od = oracle.make_design(iv_index=0, downsampling=False, rtol=1e-9)
    with AbcEngine(0) as eng:
        eng.set_design(des)
        for m, name in enumerate(MODELS, start=1):
            rows, gold = rec[f"rows_{name}"], rec[f"moments_{name}"]
            theta = maps[f"theta_{name}"][rows]
            mom, cnt = eng.simulate_moments(m, theta)
            assert cnt["n_ode_steps"] > 0
            rel = np.abs(mom - gold) / np.maximum(np.abs(gold), 1e-4)
            assert np.median(rel) < 1e-3, (name, np.median(rel))
            assert (rel < 2e-2).mean() > 0.97, (name, (rel < 2e-2).mean())
            assert rel.max() < 0.2, (name, rel.max())
            for i in range(0, len(theta), max(1, len(theta) // 6)):
                want, _ = oracle.run_part_sim(theta[i], m, od)
                r2 = np.abs(mom[i] - want) / np.maximum(np.abs(want), 1e-6)
                assert r2.max() < 1e-4, (name, i, r2.max())


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5])
def test_device_ode_statistics_match_oracle_run_sim(betas, m):
    """abc_sim through the ODE path with downsampling: 53 statistics vs the oracle's run_sim on prior draws"""
    bt = split_betas(betas)
    des = synthetic_design(betas, sim_kind=SIM_ODE, ode_rtol=1e-7, ode_atol=1e-10)
    od = oracle.make_design(iv_index=1, downsampling=True, betas=bt, rtol=1e-9)
    with AbcEngine(0) as eng:
        eng.set_design(des)
        theta, stats, cnt = eng.simulate(m, n_trials=40, particle_offset=500, seed=3)
        # S1 on the device's own moments is bit exact
        mom, _ = eng.simulate_moments(m, theta, particle_offset=500, seed=3)
        assert oracle.same_bits(stats, oracle.summary_stats(mom, des.age_dist))
        for i in range(0, 40, 5):
            want, wmom = oracle.run_sim(theta[i], m, od)
            assert np.allclose(mom[i], wmom, rtol=2e-4, atol=1e-9), (i, np.abs(mom[i] - wmom).max())
            ok = np.isfinite(want)
            assert np.array_equal(np.isfinite(stats[i]), ok)
            assert np.allclose(stats[i][ok], want[ok], rtol=2e-3, atol=1e-6), (i, np.abs(stats[i][ok] - want[ok]).max())


@pytest.mark.parametrize("m,n,min_rich,min_rho", [(1, 20000, 1200, 0.97), (2, 20000, 1200, 0.97), (3, 20000, 1000, 0.97),
                                                  (4, 20000, 400, 0.93), (5, 40000, 40, 0.88)])
def test_accepted_posteriors_agree_between_ssa_and_moment_odes(betas, m, n, min_rich, min_rho):
    """north-star part 2 (BASELINE configs[1]), all five models: on the 3419 real genes the accepted posteriors of the SSA
    path (product sampler: telegraph SSA + conditional Poisson read-out, start time per read-out) and of the moment-ODE path
    (what the reference computes) agree for the same fixed-seed parameter sets, 96 cells per read-out, eps = 4.8: per-gene
    acceptance counts rank-correlate, and the posterior means of the genes with >= 20 accepted particles under both differ
    by a fraction of the posterior SD.  (A 96-cell sample adds Monte-Carlo noise to the statistics, which lowers the
    acceptance rate of a hard threshold: 0.4-0.8 of the ODE path's.)  Reference numbers: profiles/r2_equivalence_*.json."""
    from abc_inference_transcription_b200 import ERR_NONE
    from abc_inference_transcription_b200.posteriors import get_posterior_estimate
    z = np.load(os.path.join(GOLD, "ref_summary_stats.npz"))
    d, se = z["d"], z["se"]
    with AbcEngine(0) as ssa, AbcEngine(0) as ode:
        ssa.set_design(synthetic_design(betas, n_cells=96, n_pre_cycles=10))
        ode.set_design(synthetic_design(betas, sim_kind=SIM_ODE))
        ssa.set_data(d, se)
        ode.set_data(d, se)
        theta, s_ssa, _ = ssa.simulate(m, n_trials=n, particle_offset=0, seed=4)
        _, s_ode, _ = ode.simulate(m, theta=theta)
        assert np.isfinite(s_ssa).all()
        acc = {}
        for key, eng, st in (("ssa", ssa, s_ssa), ("ode", ode, s_ode)):
            eng.accept_reset()
            _, counts, _ = eng.score(st, eps=4.8, err_layout=ERR_NONE)
            off, idx, _ = eng.accept_fetch()
            acc[key] = (counts.astype(float), off, idx)
    c_s, c_o = acc["ssa"][0], acc["ode"][0]
    both = (c_s > 0) | (c_o > 0)
    rs, ro = np.argsort(np.argsort(c_s[both])), np.argsort(np.argsort(c_o[both]))
    assert np.corrcoef(rs, ro)[0, 1] > min_rho, np.corrcoef(rs, ro)[0, 1]
    assert 0.4 < c_s.sum() / c_o.sum() < 1.1                      # MC noise of 96-cell samples lowers acceptance
    rich = np.nonzero((c_s >= 20) & (c_o >= 20))[0] + 1
    assert len(rich) >= min_rich, len(rich)
    pm_s = get_posterior_estimate(theta, acc["ssa"][1], acc["ssa"][2], rich, "mean")
    pm_o = get_posterior_estimate(theta, acc["ode"][1], acc["ode"][2], rich, "mean")
    sd = np.array([theta[acc["ode"][2][acc["ode"][1][g - 1]:acc["ode"][1][g]] - 1].std(0) for g in rich])
    shift = np.abs(pm_s - pm_o) / np.maximum(sd, 1e-6)
    assert np.median(shift) < 0.15 and np.quantile(shift, 0.9) < 0.45, (np.median(shift), np.quantile(shift, 0.9))
