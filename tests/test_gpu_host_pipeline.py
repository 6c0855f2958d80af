"""GPU test of the host-side mirror end to end: run(m, n_trials, submit) -> simulation files -> load_s_data ->
compute_trunc_errors -> error file -> accepted particles file, each stage checked against the oracle."""
import os

import numpy as np
import pytest

import oracle
from abc_inference_transcription_b200 import AbcEngine, abc_simulation, accepted_particles, compute_errors, synthetic_design
from abc_inference_transcription_b200.jlfmt import readdlm

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_wrapper_section_2_and_3(tmp_path):
    betas = np.load(os.path.join(GOLD, "ref_betas.npy"))
    z = np.load(os.path.join(GOLD, "ref_summary_stats.npz"))
    d, se = z["d"][:300], z["se"][:300]
    root = str(tmp_path)
    with AbcEngine(0) as eng:
        eng.set_design(synthetic_design(betas, n_cells=32, n_pre_cycles=8))
        # wrapper.jl:59-66: m = 3; n_trials = 40; submit = 2
        m, n_trials, submit = 3, 40, 2
        tot = abc_simulation.run(eng, m, n_trials, submit=submit, root=root, seed=9, batch=16)
        assert tot["n_particles"] == 40
        sim = os.path.join(root, "data", "simulations")
        assert readdlm(os.path.join(sim, "kon", "progress_kon_2.txt")).ravel().tolist() == [16, 32, 40]
        sets = readdlm(os.path.join(sim, "kon", "sets_kon_2.txt"))
        assert sets.shape == (40, 9)
        # the batches are windows of one global particle stream: submit 2 covers particles [40, 80)
        assert oracle.same_bits(sets, eng.fix_params(m, 40, particle_offset=40, seed=9))
        parts = compute_errors.load_s_data(sim, "kon", "_2.txt")
        stats = compute_errors.pack_stats(*parts)
        _, direct, _ = eng.simulate(m, theta=sets, particle_offset=40, seed=9)
        assert oracle.same_bits(stats, direct)                 # text round trip is bit exact
        # wrapper.jl:72-81
        cols = [d[:, 0:5], se[:, 0:5], d[:, 5:10], se[:, 5:10], d[:, 10:15], se[:, 10:15], d[:, 15:20], se[:, 15:20],
                d[:, 20:31], se[:, 20:31], d[:, 31:42], se[:, 31:42], d[:, 42:53], se[:, 42:53]]
        probe = np.concatenate([stats, d[:10] * 1.02])         # a few particles that do get accepted
        pp = [probe[:, 0:5], probe[:, 5:10], probe[:, 10:15], probe[:, 15:20], probe[:, 20:31], probe[:, 31:42], probe[:, 42:53]]
        eng.accept_reset()
        err = compute_errors.compute_trunc_errors(eng, *cols, *pp, "kon", out_dir=os.path.join(root, "errors"))
        ref = oracle.compute_trunc_errors(probe, d, se)
        assert oracle.same_bits(err, ref)
        assert oracle.same_bits(readdlm(os.path.join(root, "errors", "error_kon.txt")), ref)
        offsets, idx = accepted_particles.accepted_from_engine(eng)
        accepted_particles.write_particles(root, "kon", offsets, idx)
        lines = accepted_particles.read_particles(os.path.join(root, "data", "posteriors", "particles_kon.txt"))
        assert len(lines) == 300 and sum(len(v) for v in lines) > 0
        for g in range(300):
            assert np.array_equal(lines[g], oracle.accept_gene(ref[:, g], 4.8))


def test_wrapper_cli_sections_2_and_3(tmp_path):
    """python -m abc_inference_transcription_b200.wrapper --m 2 --n_trials 24 --submit 1: files of wrapper.jl sections 2-3"""
    from abc_inference_transcription_b200 import wrapper
    from abc_inference_transcription_b200.jlfmt import writedlm_rows
    betas = np.load(os.path.join(GOLD, "ref_betas.npy"))
    z = np.load(os.path.join(GOLD, "ref_summary_stats.npz"))
    d, se = z["d"][:120], z["se"][:120]
    ss = tmp_path / "summary_stats"
    ss.mkdir()
    cols = {"pulse_mean": (0, 5), "pulse_ff": (5, 10), "chase_mean": (10, 15), "chase_ff": (15, 20),
            "ratio": (20, 31), "mean_corr": (31, 42), "corr_mean": (42, 53)}
    for k, (a, b) in cols.items():
        dn = k if k in ("pulse_mean", "pulse_ff", "chase_mean", "chase_ff") else k + "_data"
        with open(ss / f"{dn}.txt", "w") as fh:
            writedlm_rows(fh, d[:, a:b])
        with open(ss / f"{k}_se.txt", "w") as fh:
            writedlm_rows(fh, se[:, a:b])
    np.savetxt(tmp_path / "betas.txt", betas, fmt="%.17g")
    wrapper.main(["--m", "2", "--n_trials", "24", "--submit", "1", "--root", str(tmp_path), "--summary_stats", str(ss),
                  "--betas", str(tmp_path / "betas.txt"), "--n_cells", "32", "--n_pre_cycles", "8", "--errors"])
    sim = tmp_path / "data" / "simulations" / "const_const"
    assert sorted(os.listdir(sim)) == ["progress_const_const_1.txt", "s_chase_const_const_1.txt", "s_corr_mean_const_const_1.txt",
                                       "s_mean_corr_const_const_1.txt", "s_pulse_const_const_1.txt", "s_ratios_const_const_1.txt",
                                       "sets_const_const_1.txt"]
    err = readdlm(str(tmp_path / "data" / "errors" / "error_const_const.txt"))
    assert err.shape == (24, 120)
    stats = compute_errors.pack_stats(*compute_errors.load_s_data(str(tmp_path / "data" / "simulations"), "const_const", "_1.txt"))
    assert oracle.same_bits(err, oracle.compute_trunc_errors(stats, d, se))
    assert np.array_equal(compute_errors.load_error_column(str(tmp_path / "data" / "errors"), "const_const", 7), err[:, 6])
    lines = accepted_particles.read_particles(str(tmp_path / "data" / "posteriors" / "particles_const_const.txt"))
    assert len(lines) == 120
    for g in range(120):
        assert np.array_equal(lines[g], oracle.accept_gene(err[:, g], 4.8))


def test_recover_statistics_writes_the_reference_layout(tmp_path):
    """wrapper.jl:109-111 -> scripts/recover_statistics.jl:49-68 through the device moment-ODE path and the library's
    writedlm: the 55 files of one model (11 conditions x 5 moments, one row of 5 ages per MAP row) reproduce the reference's
    shipped data/recovered_statistics files (golden fixture) to their own CVODE noise"""
    from abc_inference_transcription_b200 import SIM_ODE, io
    from abc_inference_transcription_b200.model import ID_LABELS
    betas = np.load(os.path.join(GOLD, "ref_betas.npy"))
    maps = np.load(os.path.join(GOLD, "ref_map_sets.npz"))
    rec = np.load(os.path.join(GOLD, "ref_recovered.npz"))
    root = str(tmp_path)
    with AbcEngine(0) as eng:
        eng.set_design(synthetic_design(betas, sim_kind=SIM_ODE, downsampling=False, iv=np.array([0.5, 0, 0, 0, 0, 0, 0, 0, 0.0]),
                                        ode_rtol=1e-7, ode_atol=1e-10))                 # recover_statistics.jl:33-42
        for m, name in ((5, "gamma"), (4, "alpha")):
            theta = maps[f"theta_{name}"]
            io.recover_statistics(eng, m, theta[:7], root=root)          # appended in two batches like a resumed run
            mom = io.recover_statistics(eng, m, theta[7:], root=root)
            gold = rec[f"moments_{name}"]
            for k, label in enumerate(ID_LABELS):
                for q, stem in enumerate(io.MOMENT_FILES):
                    got = readdlm(os.path.join(root, "data", "recovered_statistics", name, label, stem + ".txt"))
                    assert got.shape == (len(theta), 5)
                    assert oracle.same_bits(got[7:], mom[:, k, :, q])     # text round trip of the device result
                    rel = np.abs(got - gold[:, k, :, q]) / np.maximum(np.abs(gold[:, k, :, q]), 1e-4)
                    assert np.median(rel) < 5e-3 and rel.max() < 0.2, (name, label, stem, rel.max())
