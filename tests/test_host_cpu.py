"""CPU tests of the host-side mirror: model tables, Julia text formats, file layouts (no GPU, no compute)."""
import os

import numpy as np
import pytest

from abc_inference_transcription_b200 import get_vary_map, model_name, n_params, prior_bounds, scaling_for, vary_map_for
from abc_inference_transcription_b200 import abc_simulation, accepted_particles, compute_errors
from abc_inference_transcription_b200.dist import csr_from_tuples, shard_range
from abc_inference_transcription_b200.jlfmt import jl_float, readdlm, writedlm_rows


def test_model_tables_match_reference():
    assert [model_name(m) for m in range(1, 6)] == ["const", "const_const", "kon", "alpha", "gamma"]
    assert get_vary_map([0, 0, 0, 0], 5) == [1, 2, 3, 4]                       # model.jl:30-43
    assert get_vary_map([1, 0, 0, 0], 5) == [[1, 2, 3, 4, 5], 6, 7, 8]
    assert vary_map_for(5) == [1, 2, 3, [4, 5, 6, 7, 8]]
    assert [scaling_for(m) for m in range(1, 6)] == [1, 0, 1, 1, 1]            # abc_simulation.jl:85
    assert [n_params(m) for m in range(1, 6)] == [5, 5, 9, 9, 9]
    lo, hi = prior_bounds(5)
    assert list(lo) == [-3, -3, -3, -3, -3, -3, -3, -3, -0.7] and list(hi) == [3, 3, 3, 2, 2, 2, 2, 2, 0]
    with pytest.raises(ValueError):
        model_name(6)


@pytest.mark.parametrize("x,want", [(1.0, "1.0"), (0.1, "0.1"), (1e-5, "1.0e-5"), (0.0001, "0.0001"), (1e6, "1.0e6"),
                                    (123456.0, "123456.0"), (2.056424836200255, "2.056424836200255"), (10.0, "10.0"),
                                    (float("nan"), "NaN"), (float("-inf"), "-Inf"), (-1.5e-7, "-1.5e-7")])
def test_julia_float_text(x, want):
    assert jl_float(x) == want


def test_text_roundtrip_is_bit_exact(tmp_path):
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.lognormal(0, 5, (50, 11)), [[np.nan, np.inf, -np.inf, 0.0, 10.0, 4.8, 1e-300, 1e300, 5e-324, -0.0, 1 / 3]]])
    p = tmp_path / "a.txt"
    with open(p, "w") as fh:
        writedlm_rows(fh, a)
    b = readdlm(str(p))
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.array_equal(a[~np.isnan(a)].view(np.uint64), b[~np.isnan(b)].view(np.uint64))


def test_simulation_file_layout_roundtrip(tmp_path):
    """write_stats -> load_s_data recovers the 7 matrices (abc_simulation.jl:47-61 <-> compute_errors.jl:17-28)"""
    rng = np.random.default_rng(1)
    stats = rng.lognormal(0, 1, (7, 53))
    abc_simulation.write_stats(str(tmp_path), "kon", 3, stats[:4])
    abc_simulation.write_stats(str(tmp_path), "kon", 3, stats[4:])           # append mode like the reference
    d = tmp_path / "data" / "simulations" / "kon"
    assert sorted(os.listdir(d)) == ["s_chase_kon_3.txt", "s_corr_mean_kon_3.txt", "s_mean_corr_kon_3.txt",
                                     "s_pulse_kon_3.txt", "s_ratios_kon_3.txt"]
    assert len(open(d / "s_pulse_kon_3.txt").read().splitlines()) == 14      # 2 rows per particle
    parts = compute_errors.load_s_data(str(tmp_path / "data" / "simulations"), "kon", "_3.txt")
    assert [p.shape for p in parts] == [(7, 5)] * 4 + [(7, 11)] * 3
    assert np.array_equal(compute_errors.pack_stats(*parts), stats)
    raw = compute_errors.readdlm(str(d / "s_pulse_kon_3.txt"))
    assert np.array_equal(compute_errors.get_mean_subset(raw), stats[:, 0:5])
    assert np.array_equal(compute_errors.get_ff_subset(raw), stats[:, 5:10])


def test_particles_file_layout(tmp_path):
    offsets = np.array([0, 2, 2, 5])
    idx = np.array([7, 3, 1, 9, 4])
    accepted_particles.write_particles(str(tmp_path), "const", offsets, idx)
    text = open(tmp_path / "data" / "posteriors" / "particles_const.txt").read()
    assert text == "7\t3\n0\n1\t9\t4\n"                                       # accepted_particles.jl:23-30
    got = accepted_particles.read_particles(str(tmp_path / "data" / "posteriors" / "particles_const.txt"))
    assert [list(v) for v in got] == [[7, 3], [], [1, 9, 4]]


def test_error_column_store(tmp_path):
    err = np.arange(12.0).reshape(4, 3)
    with open(tmp_path / "error_const.txt", "w") as fh:
        writedlm_rows(fh, err)
    shapes = compute_errors.process_error_files(str(tmp_path), str(tmp_path / "cols"), model_names=("const",))
    assert shapes == {"const": (3, 4)}
    assert np.array_equal(compute_errors.load_error_column(str(tmp_path / "cols"), "const", 2), err[:, 1])


def test_shard_range_partitions_exactly():
    for n, w in [(10, 3), (5_000_000, 8), (7, 8), (0, 2)]:
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_csr_from_tuples_order():
    gene = np.array([2, 0, 2, 0, 2], dtype=np.int32)
    part = np.array([5, 9, 3, 1, 8], dtype=np.int64)
    err = np.array([1.0, 2.0, 1.0, 2.0, 0.5])
    off, idx, e = csr_from_tuples(gene, part, err, 4)
    assert list(off) == [0, 2, 2, 5, 5] and list(idx) == [1, 9, 8, 3, 5] and list(e) == [2.0, 2.0, 0.5, 1.0, 1.0]


def test_model_probs_and_posterior_summaries():
    from abc_inference_transcription_b200 import posteriors
    rng = np.random.default_rng(0)
    p, lb, ub = posteriors.get_model_probs([30, 10], rng=rng)
    assert np.allclose(p, [0.75, 0.25]) and (lb <= p).all() and (p <= ub).all() and ub[0] <= 1.0
    assert all((x == 0).all() for x in posteriors.get_model_probs([0, 0]))
    counts = np.array([[5, 0, 0, 2], [0, 0, 0, 1], [0, 0, 3, 1], [0, 0, 0, 0], [0, 0, 1, 0]])
    prob, lb, ub = posteriors.model_probs_for_genes(counts, [[0, 1], [2, 3, 4]], rng=rng)
    assert prob[0].tolist() == [1.0, 0.0] and prob[1].tolist() == [0.0, 0.0] and prob[2].tolist() == [0.0, 1.0]
    assert np.allclose(prob[3], [0.75, 0.25])                        # model_probs.jl:42-54
    sets = np.arange(40.0).reshape(10, 4)
    offsets, idx = np.array([0, 3, 3, 5]), np.array([7, 2, 9, 1, 4])
    assert posteriors.get_n_particles(offsets).tolist() == [3, 0, 2]
    m = posteriors.get_posterior_estimate(sets, offsets, idx, [1, 3], "map")
    assert np.array_equal(m, sets[[6, 0]])                           # first accepted index, 1-based (posterior_kinetics.jl:14)
    mean = posteriors.get_posterior_estimate(sets, offsets, idx, [1], "mean")
    assert np.allclose(mean[0], sets[[6, 1, 8]].mean(0))
    lbs, ubs = posteriors.get_posterior_ci(sets, offsets, idx, [1], 0.95)
    assert np.allclose(ubs[0], np.quantile(sets[[6, 1, 8]], 0.95, axis=0)) and (lbs <= ubs).all()
    with pytest.raises(ValueError):
        posteriors.get_posterior_estimate(sets, offsets, idx, [2], "map")


def test_engine_argument_plumbing_with_stub_library():
    """the numpy side of fix_params(out=...) -> simulate_score(theta=...) (what bench.py's end-to-end loop calls) against a
    stub of the C library: shapes, contiguity checks and that caller-provided (page-locked-like) buffers are passed through"""
    import ctypes
    from abc_inference_transcription_b200 import engine as eng_mod

    calls = []

    class Stub:
        def __getattr__(self, name):
            def f(*a):
                calls.append((name, a))
                return 0
            return f

    e = object.__new__(eng_mod.AbcEngine)
    e._lib, e._ctx, e.n_genes, e.design, e.device = Stub(), ctypes.c_void_p(1), 7, None, 0
    B, m, P = 16, 3, 9
    buf = (ctypes.c_char * (B * P * 8))()
    th_host = np.frombuffer(buf, dtype=np.float64, count=B * P).reshape(B, P)          # like PinnedArray.array
    th = e.fix_params(m, B, particle_offset=5, seed=1, out=th_host)
    assert th is th_host
    err = np.empty((B, 7)); st = np.empty((B, 53))
    theta, stats, err_o, counts, cnt = e.simulate_score(m, theta=th, particle_offset=5, seed=1, eps=4.8, err_layout=2, out=err,
                                                        stats_out=st)
    assert theta.ctypes.data == th_host.ctypes.data and stats is st and err_o is err and counts.shape == (7,)
    name, a = calls[-1]
    assert name == "abc_simulate_score" and a[1] == m and a[2] == B and a[3] == 5 and a[5] == 1        # prior supplied
    theta2, *_ = e.simulate_score(m, n_trials=B, theta_out=th_host, stats_out=st, out=err, err_layout=2)
    assert theta2 is th_host and calls[-1][1][5] == 0
    with pytest.raises(AssertionError):
        e.fix_params(m, B, out=np.empty((B, 5)))
    # the asynchronous pair bench.py's end-to-end loop uses: buffers passed through untouched, wait() returns counts
    e.simulate_score_async(m, th_host, st, err, prior_supplied=True, particle_offset=5, seed=1, eps=4.8, err_layout=2)
    name, a = calls[-1]
    assert name == "abc_simulate_score_async" and a[1] == m and a[2] == B and a[3] == 5 and a[5] == 1 and a[9] == 2
    assert a[6].value == th_host.ctypes.data and a[7].value == st.ctypes.data and a[10].value == err.ctypes.data
    with pytest.raises(AssertionError):
        e.simulate_score_async(m, th_host, st, np.empty((7, B)), err_layout=2)       # wrong orientation of the error matrix
    counts, cnt = e.wait()
    assert calls[-1][0] == "abc_wait" and counts.shape == (7,)
    e._ctx = None
