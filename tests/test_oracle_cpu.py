"""CPU tests of the oracle itself (-m "not gpu"): pinned on the reference's golden vectors."""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = ["const", "const_const", "kon", "alpha", "gamma"]


# ------------------------------------------------------------------ Philox known answers (Random123 kat_vectors)
@pytest.mark.parametrize("ctr,key,want", [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
])
def test_philox4x32_10_known_answers(ctr, key, want):
    assert [int(x) for x in oracle.philox(ctr, key)] == want


def test_numpy_philox_equals_the_c_restatement():
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2 ** 32, size=(50, 4), dtype=np.uint64)
    key = [0xa4093822, 0x299f31d0]
    got = np.stack(oracle.philox_np(ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3], key), axis=1)
    for c, g in zip(ctr, got):
        assert [int(x) for x in oracle.philox([int(v) for v in c], key)] == [int(x) for x in g]


def test_model_probs_restatement_follows_the_reference_cases():
    """model_probs.jl:42-54 + get_model_probs: no hypothesis accepted -> zeros; one -> ones; else ratios with bootstrap bounds
    around them; compared with the host mirror posteriors.get_model_probs (multinomial) in distribution"""
    from abc_inference_transcription_b200.posteriors import model_probs_for_genes
    counts = np.array([[0, 5, 300, 40], [0, 0, 100, 40], [0, 0, 600, 0]], dtype=np.int64)       # K = 3 hypotheses x 4 genes
    prob, lb, ub = oracle.model_probs(counts, 100, 0.95, 7)
    assert not prob[0].any() and not lb[0].any() and not ub[0].any()
    assert list(prob[1]) == [1.0, 0.0, 0.0] and list(lb[1]) == [1.0, 0.0, 0.0] and list(ub[1]) == [1.0, 0.0, 0.0]
    assert np.allclose(prob[2], [0.3, 0.1, 0.6]) and np.all(lb[2] < prob[2]) and np.all(prob[2] < ub[2])
    assert prob[3][2] == 0.0 and ub[3][2] == 0.0 and abs(prob[3][0] - 0.5) < 1e-12
    p2, l2, u2 = model_probs_for_genes(counts, [[0], [1], [2]], rng=np.random.default_rng(3))
    assert np.allclose(p2, prob) and np.abs(l2 - lb).max() < 0.05 and np.abs(u2 - ub).max() < 0.05
    # the bootstrap spread is the binomial one: sd = sqrt(p (1 - p) / n), bounds ~ +-1.645 sd
    sd = np.sqrt(0.3 * 0.7 / 1000)
    assert abs((ub[2][0] - lb[2][0]) / (2 * 1.645 * sd) - 1.0) < 0.35


# ------------------------------------------------------------------ simulator vs data/recovered_statistics
def _pool_map(fn, items, threads=None):
    """the oracle's C calls release the GIL: run them on all host threads"""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=threads or (os.cpu_count() or 1)) as ex:
        return list(ex.map(fn, items))


@pytest.mark.parametrize("m,name,n_rows", [(1, "const", 1844), (2, "const_const", 478), (3, "kon", 101), (4, "alpha", 15), (5, "gamma", 65)])
def test_moment_odes_match_reference_recovered_statistics(m, name, n_rows):
    """orc_run_part_sim == run_part_sim (recover_statistics.jl:1-11) on ALL 2503 MAP parameter sets the reference ships
    (data/posterior_estimates/map_sets_*.txt vs data/recovered_statistics/**: 2503 x 275 = 688 325 values, SURVEY section 4
    test 1).  The goldens carry CVODE reltol 1e-3 noise + the 1 % transient criterion: tolerance 2e-2 relative with a 1e-4
    absolute floor on >= 97 % of the entries, 0.2 worst case, median < 1e-3."""
    maps = np.load(os.path.join(GOLD, "ref_map_sets.npz"))
    rec = np.load(os.path.join(GOLD, "ref_recovered.npz"))
    rows, gold = rec[f"rows_{name}"], rec[f"moments_{name}"]
    assert len(rows) == n_rows == len(maps[f"theta_{name}"]) and np.array_equal(rows, np.arange(n_rows))
    theta = maps[f"theta_{name}"][rows]
    d = oracle.make_design(iv_index=0, downsampling=False, rtol=1e-7)    # recover_statistics.jl:33-34: iv[1] = 1/2
    oracle.lib()
    res = _pool_map(lambda th: oracle.run_part_sim(th, m, d), theta)
    assert all(1 <= k <= 101 for _, k in res)
    got = np.array([g for g, _ in res])
    rel = np.abs(got - gold) / np.maximum(np.abs(gold), 1e-4)
    assert np.median(rel) < 1e-3, np.median(rel)
    assert (rel < 2e-2).mean() > 0.97, (rel < 2e-2).mean()
    assert rel.max() < 0.2, rel.max()
    # row by row: no single parameter set is off as a whole.  (Two gamma rows with switching times of ~130 h sit 2-3 %
    # off uniformly: transient_phase stops at 1 % change per cycle, model.jl:116, one iteration earlier or later under
    # CVODE noise -- hence 5e-2 here.)
    per_row = (rel < 5e-2).reshape(len(theta), -1).mean(1)
    assert per_row.min() > 0.95, (int(per_row.argmin()), per_row.min())


def test_integrator_is_converged():
    """tight vs tighter tolerance agree to 1e-7: the restatement's own error is far below the goldens' noise"""
    th = np.log10([2.0, 1.0, 85.0, 2.0, 3.0, 1.0, 1.0, 1.5, 0.7])       # model_realisation.jl:314, gamma == kon+koff resonance
    a, _ = oracle.run_part_sim(th, 5, oracle.make_design(iv_index=0, rtol=1e-8))
    b, _ = oracle.run_part_sim(th, 5, oracle.make_design(iv_index=0, rtol=1e-10))
    assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6)) < 1e-6


def test_rate_schedule_semantics():
    """get_rate / size_scaling / labelling corner cases (model.jl:1-27, 58-64)"""
    import ctypes
    lib = oracle.lib()
    assert lib.orc_size_scaling(20.0, 0.0) == 1.0 and lib.orc_size_scaling(20.0, 20.0) == 2.0
    assert lib.orc_size_scaling(20.0, 10.0) == 1.5 and lib.orc_size_scaling(20.0, -50.0) == 1.5   # floored mod
    assert lib.orc_size_scaling(20.0, 25.0) == 1.0                                                   # t > cycle
    assert lib.orc_labelling(-0.5, 1.0, 2.0, 1.0) == pytest.approx(10 ** -0.5) and lib.orc_labelling(-0.5, 1.0, 2.0, 3.0001) == 0.0
    th = np.array([0.1, 0.2, 0.3, 0.4, 0.5, -1.0, 1.0, -0.5, -0.3])     # m = 3: kon varies
    p = np.zeros(4)
    for t, want in [(0.0, 0.1), (4.0, 0.1), (4.0001, 0.2), (19.9, 0.5), (-60.0, 0.1), (-0.5, 0.5), (20.0, 0.1)]:
        lib.orc_get_rate(th.ctypes.data_as(ctypes.c_void_p), 3, ctypes.c_double(20.0), ctypes.c_double(t),
                         p.ctypes.data_as(ctypes.c_void_p))
        assert p[0] == want, (t, p[0])
    # scaling enters alpha only, and not for m = 2
    lib.orc_get_rate(th.ctypes.data_as(ctypes.c_void_p), 3, ctypes.c_double(20.0), ctypes.c_double(10.0), p.ctypes.data_as(ctypes.c_void_p))
    assert p[2] == pytest.approx(1.0 + np.log10(1.5)) and p[1] == -1.0 and p[3] == -0.5
    th5 = np.array([0.1, -1.0, 1.0, -0.5, -0.3])
    lib.orc_get_rate(th5.ctypes.data_as(ctypes.c_void_p), 2, ctypes.c_double(20.0), ctypes.c_double(10.0), p.ctypes.data_as(ctypes.c_void_p))
    assert p[2] == 1.0


def test_periodic_boundary_and_downsample_formulas():
    import ctypes
    lib = oracle.lib()
    e = np.arange(1.0, 10.0)
    v = np.zeros(9)
    lib.orc_periodic_boundary(e.ctypes.data_as(ctypes.c_void_p), v.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(v, [1, 1.0, 1.5, 4, 2.5, 3.0, 7 / 4 + 2 / 4, 8 / 4, 9 / 4 + 3 / 4])     # model.jl:98-111
    # downsample == law of total variance for binomial thinning with random beta (model.jl:221-239)
    rng = np.random.default_rng(0)
    betas = rng.uniform(0.02, 0.4, 500)
    cl = np.ones(500, dtype=np.int32)
    cl[:] = 1 + np.arange(500) % 5
    bm, b2, bv = oracle.beta_moments(betas, cl)
    s = np.tile(np.array([30.0, 12.0, 80.0, 9.0, 25.0]), (5, 1))
    o = np.zeros((5, 5))
    lib.orc_downsample(np.ascontiguousarray(s).ctypes.data_as(ctypes.c_void_p), bm.ctypes.data_as(ctypes.c_void_p),
                       b2.ctypes.data_as(ctypes.c_void_p), bv.ctypes.data_as(ctypes.c_void_p), o.ctypes.data_as(ctypes.c_void_p))
    for c in range(5):
        b = betas[cl == c + 1]
        m1, m2_, var = b.mean(), (b ** 2).mean(), b.var(ddof=1)
        assert o[c, 0] == pytest.approx(m1 * 30) and o[c, 1] == pytest.approx(m1 * 12)
        assert o[c, 2] == pytest.approx((m1 - m2_) * 30 + var * (900 + 80) + m1 ** 2 * 80)
        assert o[c, 3] == pytest.approx(var * (30 * 12 + 9) + m1 ** 2 * 9)


# ------------------------------------------------------------------ scoring / acceptance
def test_nlsqerror_matches_numpy_restatement_bitwise():
    """independent numpy restatement of compute_errors.jl:30-43,58-64 (elementwise IEEE ops, sequential sums)"""
    z = np.load(os.path.join(GOLD, "ref_summary_stats.npz"))
    d, se = z["d"], z["se"]
    rng = np.random.default_rng(3)
    s = d[rng.integers(0, len(d), 40)] * np.exp(rng.normal(0, 0.3, (40, 53)))
    got = oracle.compute_trunc_errors(s, d, se)
    sig2 = np.float64(0.1) * np.float64(0.1)
    eps = np.where(se + d != 0.0, 0.0, 0.0001)
    den = (se * se + sig2 * (d * d)) + eps
    off = [0, 5, 10, 15, 20, 31, 42, 53]
    want = np.zeros((40, len(d)))
    for i in range(40):
        q = ((d - s[i]) * (d - s[i])) / den
        err = np.zeros(len(d))
        for l in range(7):
            e = np.zeros(len(d))
            for t in range(off[l], off[l + 1]):
                e = e + q[:, t]
            err = err + e / 53.0
        want[i] = np.where(err > 10.0, 10.0, err)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    assert ((se + d) == 0).sum() >= 10      # the eps branch is present in the shipped data


def test_accept_gene_order_ties_and_sentinel():
    err = np.array([5.0, 1.0, 4.8, 1.0, np.nan, 0.5, 10.0, 4.800000000000001])
    assert list(oracle.accept_gene(err, 4.8)) == [6, 2, 4, 3]           # ascending error, ties by index, 1-based
    assert len(oracle.accept_gene(np.array([10.0, np.nan, 4.9]), 4.8)) == 0   # -> the reference writes "0"


def test_weak_end_to_end_kat_map_rows_are_accepted_by_their_gene():
    """SURVEY 8c KAT (3) on ALL 2503 MAP rows: a MAP row is the particle with the smallest error for its gene among the
    particles of its model (posterior_kinetics.jl:14), so (i) err(stats(MAP theta_i), gene_i) <= 4.8 and (ii) among the MAP
    rows of the same model -- all of them particles of the same run -- row i minimises the error of gene_i
    (data/model_selection/<model>_genes.txt maps rows to genes).  This pins downsample, the 53 statistics and the scoring
    order end to end on reference-held data.  Design constants are approximated (uniform age weights, round-robin age
    clusters: SURVEY R10), hence 'nearly all' rather than 'all'; measured: 99.7 % accepted, 86 % exact minima, 100 % within
    1.25 x the minimum."""
    from abc_inference_transcription_b200.design import split_betas
    maps = np.load(os.path.join(GOLD, "ref_map_sets.npz"))
    z = np.load(os.path.join(GOLD, "ref_summary_stats.npz"))
    betas = np.load(os.path.join(GOLD, "ref_betas.npy"))
    od = oracle.make_design(iv_index=1, downsampling=True, betas=split_betas(betas), rtol=1e-5)
    oracle.lib()
    acc = is_min = near_min = tot = 0
    for m, name in enumerate(MODELS, start=1):
        th, genes = maps[f"theta_{name}"], maps[f"genes_{name}"]
        stats = np.array(_pool_map(lambda x: oracle.run_sim(x, m, od)[0], th))
        g = genes - 1
        E = oracle.compute_trunc_errors(stats, z["d"][g], z["se"][g])      # rows: MAP particles, columns: their genes
        own, col_min = np.diag(E), np.nanmin(E, axis=0)
        acc += int((own <= 4.8).sum())
        is_min += int((own <= col_min).sum())
        near_min += int((own <= 1.25 * col_min).sum())
        tot += len(th)
        assert (own <= 4.8).mean() >= 0.8, (name, (own <= 4.8).mean())
    assert tot == 2503
    assert acc / tot > 0.99, (acc, tot)
    assert is_min / tot > 0.8, (is_min, tot)
    assert near_min / tot > 0.99, (near_min, tot)


# ------------------------------------------------------------------ oracle SSA vs the moment ODEs
def test_oracle_ssa_moments_agree_with_moment_odes():
    """z-tests: the CME simulated by the oracle SSA has the moments scripts/model.jl integrates"""
    from abc_inference_transcription_b200.design import split_betas
    betas = np.load(os.path.join(GOLD, "ref_betas.npy"))
    bt = split_betas(betas)
    th, m = np.log10([0.5, 1.0, 4.0, 0.3, 0.7]), 1
    n = 3000
    sd, keep = oracle.make_ssa_design(n, 9, True, bt)
    _, mom_ds = oracle.run_sim(th, m, oracle.make_design(iv_index=1, downsampling=True, betas=bt, rtol=1e-9))
    mom_raw, _ = oracle.run_part_sim(th, m, oracle.make_design(iv_index=1, downsampling=False, rtol=1e-9))
    zs = []
    for c, a in [(5, 1), (8, 3)]:
        x, ev = oracle.ssa_readout(th, m, sd, 3, 17, c, a, oracle.MATH_DET)
        x = x.astype(np.float64)
        for (u, l), ref in [((x[0], x[1]), mom_raw[c, a]), ((x[2], x[3]), mom_ds[c, a])]:
            for sample, target in [(u, ref[0]), (l, ref[1])]:
                zs.append((sample.mean() - target) / (sample.std(ddof=1) / np.sqrt(n)))
            for xs, ys, target in [(u, u, ref[2]), (u, l, ref[3]), (l, l, ref[4])]:
                p = (xs - xs.mean()) * (ys - ys.mean())
                zs.append((p.sum() / (n - 1) - target) / (p.std(ddof=1) / np.sqrt(n)))
    zs = np.array(zs)
    assert np.abs(zs).max() < 4.5, zs
    assert (zs ** 2).mean() < 2.5, zs


def test_oracle_ssa_math_modes_and_binomials():
    """deterministic vs libm math: same law (most cells identical); exact bitwise binomial has the right mean"""
    from abc_inference_transcription_b200.design import split_betas
    betas = np.load(os.path.join(GOLD, "ref_betas.npy"))
    sd, keep = oracle.make_ssa_design(400, 6, True, split_betas(betas))
    th = np.log10([0.5, 1.0, 20.0, 0.1, 0.7])
    a, _ = oracle.ssa_readout(th, 1, sd, 0, 5, 6, 2, oracle.MATH_DET)
    b, _ = oracle.ssa_readout(th, 1, sd, 0, 5, 6, 2, oracle.MATH_LIBM)
    assert (a == b).all(0).mean() > 0.5
    ratio = a[2:].sum() / a[:2].sum()              # chase cells: mean beta = 0.2 exactly
    assert abs(ratio - 0.2) < 0.01
    assert (a[2] <= a[0]).all() and (a[3] <= a[1]).all()


def test_exp10_det_accuracy():
    x = np.linspace(-3.2, 3.2, 2001)
    got = np.array([oracle.lib().orc_exp10_det(float(v)) for v in x])
    assert np.max(np.abs(got / 10.0 ** x - 1.0)) < 1e-15
