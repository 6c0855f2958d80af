"""CPU tests of the library's host-side writers (csrc/abc_io.cu) against the Python mirror of Julia's text
formatting (jlfmt.py, itself pinned on Julia spellings in test_host_cpu.py) and against the reference's file layouts."""
import os

import numpy as np
import pytest

from abc_inference_transcription_b200 import io
from abc_inference_transcription_b200.accepted_particles import read_particles, write_particles
from abc_inference_transcription_b200.abc_simulation import write_stats
from abc_inference_transcription_b200.jlfmt import jl_float, readdlm, writedlm_rows


@pytest.mark.parametrize("x,want", [
    (1.0, "1.0"), (0.1, "0.1"), (1e-5, "1.0e-5"), (0.0001, "0.0001"), (100000.0, "100000.0"), (1e6, "1.0e6"),
    (123456.7, "123456.7"), (1234567.0, "1.234567e6"), (-2.5e-7, "-2.5e-7"), (float("nan"), "NaN"), (float("inf"), "Inf"),
    (-float("inf"), "-Inf"), (0.0, "0.0"), (-0.0, "-0.0"), (5e-324, "5.0e-324"), (1.7976931348623157e308, "1.7976931348623157e308"),
    (10.0, "10.0"), (4.8, "4.8"), (0.010000000000000002, "0.010000000000000002"), (1 / 3, "0.3333333333333333"),
])
def test_format_float64_is_julias_print(x, want):
    assert io.format_float64(x) == want


def test_format_float64_matches_the_python_mirror_and_round_trips():
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.normal(size=20000) * 10.0 ** rng.integers(-30, 30, 20000),
                         np.frombuffer(rng.bytes(8 * 20000), dtype=np.float64)])
    for x in xs:
        s = io.format_float64(x)
        assert s == jl_float(x)
        if np.isfinite(x):
            assert float(s) == x                      # shortest digits that round-trip


def test_writedlm_equals_the_mirror_byte_for_byte(tmp_path):
    rng = np.random.default_rng(2)
    a = rng.lognormal(0, 3, size=(137, 53))
    a[5, 7] = np.nan; a[9, 0] = 10.0; a[11, 52] = np.inf
    p1, p2 = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
    io.writedlm(p1, a[:60], append=False)
    io.writedlm(p1, a[60:], append=True)
    with open(p2, "w") as fh:
        writedlm_rows(fh, a)
    assert open(p1, "rb").read() == open(p2, "rb").read()
    assert np.array_equal(readdlm(p1).view(np.uint64), a.view(np.uint64))
    big = rng.normal(size=(3000, 40))                 # large enough for the multi-threaded path
    io.writedlm(p1, big, append=False)
    with open(p2, "w") as fh:
        writedlm_rows(fh, big)
    assert open(p1, "rb").read() == open(p2, "rb").read()


def test_write_simulation_equals_the_reference_layout(tmp_path):
    """abc_simulation.jl:47-61, 89-95: sets_, s_pulse_ / s_chase_ (2 rows x 5 per trial), s_ratios_, s_mean_corr_, s_corr_mean_"""
    rng = np.random.default_rng(3)
    n, m, submit = 23, 4, 2
    theta, stats = rng.uniform(-3, 3, (n, 9)), rng.lognormal(0, 1, (n, 53))
    r1, r2 = str(tmp_path / "lib"), str(tmp_path / "py")
    io.write_simulation(r1, m, submit, theta[:10], stats[:10], first_trial=1)
    io.write_simulation(r1, m, submit, theta[10:], stats[10:], first_trial=11)
    write_stats(r2, "alpha", submit, stats)
    d1, d2 = os.path.join(r1, "data", "simulations", "alpha"), os.path.join(r2, "data", "simulations", "alpha")
    for stem in ("s_pulse", "s_chase", "s_ratios", "s_mean_corr", "s_corr_mean"):
        f = f"{stem}_alpha_{submit}.txt"
        assert open(os.path.join(d1, f), "rb").read() == open(os.path.join(d2, f), "rb").read(), stem
    assert np.array_equal(readdlm(os.path.join(d1, f"sets_alpha_{submit}.txt")).view(np.uint64), theta.view(np.uint64))
    assert [int(x) for x in open(os.path.join(d1, f"progress_alpha_{submit}.txt")).read().split()] == list(range(1, n + 1))
    sp = readdlm(os.path.join(d1, f"s_pulse_alpha_{submit}.txt"))
    assert sp.shape == (2 * n, 5) and np.array_equal(sp[0::2], stats[:, 0:5]) and np.array_equal(sp[1::2], stats[:, 5:10])


def test_write_accepted_and_error_columns(tmp_path):
    offsets = np.array([0, 3, 3, 4, 4, 9], dtype=np.int64)
    idx = np.array([7, 2, 1000000, 5, 1, 2, 3, 4, 5], dtype=np.int64)
    p = str(tmp_path / "posteriors" / "particles_kon.txt")
    io.write_accepted(p, offsets, idx, append=False)
    assert open(p).read() == "7\t2\t1000000\n0\n5\n0\n1\t2\t3\t4\t5\n"
    write_particles(str(tmp_path / "ref"), "kon", offsets, idx)
    assert open(p).read() == open(tmp_path / "ref" / "data" / "posteriors" / "particles_kon.txt").read()
    got = read_particles(p)
    assert [list(v) for v in got] == [[7, 2, 1000000], [], [5], [], [1, 2, 3, 4, 5]]
    rng = np.random.default_rng(4)
    e = rng.uniform(0, 10, (6, 50))                   # gene-major: 6 genes x 50 particles
    col = str(tmp_path / "error_kon.cols")
    io.write_error_columns(col, e[:, :20], append=False)
    io.write_error_columns(col, e[:, 20:], append=True)
    for g in range(1, 7):
        assert np.array_equal(io.read_error_column(col, g).view(np.uint64), e[g - 1].view(np.uint64))
    assert open(os.path.join(col, "meta.txt")).read().split() == ["6", "50"]
