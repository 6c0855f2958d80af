"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

Bars (north_star): scoring / acceptance / statistics / prior draws bit-exact; SSA bit-exact against the
oracle in the deterministic-math variant and statistically equivalent (KS, z-tests) in the fast variant.
"""
import contextlib
import copy
import os

import numpy as np
import pytest

import oracle
from abc_inference_transcription_b200 import (AbcEngine, AbcError, ERR_GENE_MAJOR, ERR_NONE, ERR_PARTICLE_MAJOR, n_params,
                                              split_betas, synthetic_design)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def betas():
    return np.load(os.path.join(GOLD, "ref_betas.npy"))


@pytest.fixture(scope="module")
def data_stats():
    z = np.load(os.path.join(GOLD, "ref_summary_stats.npz"))
    return z["d"], z["se"]


@pytest.fixture(scope="module")
def eng(betas, data_stats):
    e = AbcEngine(0)
    e.set_design(synthetic_design(betas, n_cells=96, n_pre_cycles=10))
    e.set_data(*data_stats)
    yield e
    e.close()


@contextlib.contextmanager
def cells_per_readout(eng, n_cells):
    old = eng.design
    new = copy.copy(old)
    new.n_cells = n_cells
    eng.set_design(new)
    try:
        yield
    finally:
        eng.set_design(old)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def synth_stats(rng, d, n):
    """statistics near the data of random genes (so that errors straddle 4.8 and 10) + wild ones"""
    g = rng.integers(0, d.shape[0], size=n)
    s = d[g] * np.exp(rng.normal(0.0, 0.25, size=(n, 53)))
    wild = rng.random(n) < 0.3
    s[wild] = d[g[wild]] * np.exp(rng.normal(0.0, 2.0, size=(wild.sum(), 53)))
    return s


# ------------------------------------------------------------------------------------------ scoring
def test_score_bit_exact_both_layouts(eng, data_stats):
    d, se = data_stats
    rng = np.random.default_rng(1)
    n = 777                                      # ragged: not a multiple of the 128-particle tile
    s = synth_stats(rng, d, n)
    ref = oracle.compute_trunc_errors(s, d, se)
    assert (ref < 10.0).mean() > 0.001 and (ref <= 4.8).any()
    eng.accept_reset()
    err_p, counts, _ = eng.score(s, eps=4.8, err_layout=ERR_PARTICLE_MAJOR)
    assert oracle.same_bits(err_p, ref)
    eng.accept_reset()
    err_g, counts_g, _ = eng.score(s, eps=4.8, err_layout=ERR_GENE_MAJOR)
    assert oracle.same_bits(err_g, ref.T)
    assert np.array_equal(counts, (ref <= 4.8).sum(0)) and np.array_equal(counts, counts_g)


def test_score_special_values(eng, data_stats):
    """NaN passes through unclipped, +-Inf clips to 10, the se+d==0 branch (eps=1e-4) is exercised"""
    d, se = data_stats
    assert ((se + d) == 0.0).any(), "fixture must contain the se+d==0 entries (SURVEY section 4)"
    rng = np.random.default_rng(2)
    s = synth_stats(rng, d, 300)
    s[3, 7] = np.nan
    s[5, 40] = np.inf
    s[6, 0] = -np.inf
    s[7, :] = np.nan
    s[8, 25] = np.inf
    s[8, 30] = np.nan
    s[9] = d[11]                                  # exact match of gene 12: error 0
    s[10] = 0.0
    ref = oracle.compute_trunc_errors(s, d, se)
    assert np.isnan(ref[3]).all() and np.isnan(ref[7]).all() and (ref[5] == 10.0).all() and ref[9, 11] == 0.0
    eng.accept_reset()
    err, counts, _ = eng.score(s, eps=4.8, err_layout=ERR_PARTICLE_MAJOR)
    assert oracle.same_bits(err, ref)
    assert np.array_equal(counts, (ref <= 4.8).sum(0))


def test_score_late_nan_defeats_early_exit(eng, data_stats):
    """a NaN in the LAST group must give NaN even when the first groups already exceed 10 for every lane
    of the warp (the reference does not clip NaN, compute_errors.jl:62-64)"""
    d, se = data_stats
    rng = np.random.default_rng(22)
    s = d[rng.integers(0, d.shape[0], 256)] * 1e3       # wildly off: every pair saturates in group 1
    s[17, 52] = np.nan
    s[40, 31] = np.nan
    s[41, 45] = np.inf
    ref = oracle.compute_trunc_errors(s, d, se)
    assert np.isnan(ref[17]).all() and np.isnan(ref[40]).all() and (ref[41] == 10.0).all() and (ref[0] == 10.0).all()
    eng.accept_reset()
    err, counts, _ = eng.score(s, eps=4.8, err_layout=ERR_PARTICLE_MAJOR)
    assert oracle.same_bits(err, ref)
    assert counts.sum() == 0
    # non-finite DATA of a gene: that gene's column is evaluated in full
    d2, se2 = d.copy(), se.copy()
    d2[5, 50] = np.nan
    se2[9, 44] = np.inf
    eng.set_data(d2, se2)
    try:
        err2, _, _ = eng.score(s, eps=4.8, err_layout=ERR_PARTICLE_MAJOR)
    finally:
        eng.set_data(d, se)
    ref2 = oracle.compute_trunc_errors(s, d2, se2)
    assert np.isnan(ref2[:, 5]).all()
    assert oracle.same_bits(err2, ref2)


def test_accept_lists_match_reference_order(eng, data_stats):
    """v[sortperm(err[v])] per gene, 1-based, stable; empty -> the '0' line (accepted_particles.jl:19-30)"""
    d, se = data_stats
    rng = np.random.default_rng(3)
    n = 1500
    s = synth_stats(rng, d, n)
    s[100:110] = s[50:60]                         # exact ties in error -> order by index
    ref = oracle.compute_trunc_errors(s, d, se)
    eng.accept_reset()
    _, counts, _ = eng.score(s, eps=4.8, err_layout=ERR_NONE)
    offsets, idx, errs = eng.accept_fetch()
    assert offsets[-1] == (ref <= 4.8).sum() > 0
    n_empty = 0
    for g in range(d.shape[0]):
        want = oracle.accept_gene(ref[:, g], 4.8)
        got = idx[offsets[g]:offsets[g + 1]]
        assert np.array_equal(got, want), g
        assert oracle.same_bits(errs[offsets[g]:offsets[g + 1]], ref[want - 1, g])
        n_empty += len(want) == 0
    assert n_empty > 0
    # fused-only pass must agree with the matrix pass, and offsets shift the reported indices
    eng.accept_reset()
    eng.score(s[:700], eps=4.8, particle_offset=0, err_layout=ERR_NONE)
    eng.score(s[700:], eps=4.8, particle_offset=700, err_layout=ERR_NONE)
    o2, i2, _ = eng.accept_fetch()
    assert np.array_equal(o2, offsets) and np.array_equal(i2, idx)


def test_score_kernels_agree_and_eps_above_clip(eng, data_stats):
    """tile-pruned kernel == three-stage kernel == plain FP64 kernel == oracle; eps >= 10 (accepts the clipped
    pairs) takes the plain path"""
    d, se = data_stats
    rng = np.random.default_rng(8)
    s = synth_stats(rng, d, 600)
    s[5, 3] = np.nan
    ref = oracle.compute_trunc_errors(s, d, se)
    got = {}
    for key, (tile, plain) in {"tile": (1, 0), "three_stage": (0, 0), "plain": (0, 1)}.items():
        eng.set_option("score_tile_kernel", tile)
        eng.set_option("score_reference_kernel", plain)
        try:
            for layout, want in ((ERR_GENE_MAJOR, ref.T), (ERR_PARTICLE_MAJOR, ref)):
                eng.accept_reset()
                err, counts, _ = eng.score(s, eps=4.8, err_layout=layout)
                off, idx, errs = eng.accept_fetch()
                assert oracle.same_bits(err, want), (key, layout)
            got[key] = (counts, off, idx)
        finally:
            eng.set_option("score_tile_kernel", 1)
            eng.set_option("score_reference_kernel", 0)
    for key in ("three_stage", "plain"):
        assert all(np.array_equal(a, b) for a, b in zip(got["tile"], got[key])), key
    eng.accept_reset()
    err, counts, _ = eng.score(s[:40], eps=10.0, err_layout=ERR_PARTICLE_MAJOR)
    assert oracle.same_bits(err, ref[:40])
    assert np.array_equal(counts, (ref[:40] <= 10.0).sum(0)) and counts.sum() > 40 * 3000
    off, idx, _ = eng.accept_fetch()
    for g in (0, 1711, 3418):
        assert np.array_equal(idx[off[g]:off[g + 1]], oracle.accept_gene(ref[:40, g], 10.0))


def test_score_tile_pruning_wide_particles(eng, data_stats):
    """particles spread over twelve decades, negative values in statistics that are non-negative for simulated
    particles, several particle blocks, a ragged tail: the tile bounds must never drop a pair below 10"""
    d, se = data_stats
    rng = np.random.default_rng(31)
    n = 2 * 2048 + 333
    s = 10.0 ** rng.uniform(-6, 6, size=(n, 53)) * np.where(rng.random((n, 53)) < 0.1, -1.0, 1.0)
    near = rng.random(n) < 0.4                       # particles near one gene, rescaled as a whole
    g = rng.integers(0, d.shape[0], size=n)
    s[near] = d[g[near]] * np.exp(rng.normal(0.0, 0.15, size=(near.sum(), 53))) * 10.0 ** rng.choice([0, 0, 0.5, -0.5, 1], size=(near.sum(), 1))
    s[11, 20] = np.nan
    s[2050, 52] = np.nan
    s[4100, 0] = np.inf
    ref = oracle.compute_trunc_errors(s, d, se)
    assert 0.0005 < (ref < 10.0).mean() < 0.2 and np.isnan(ref[11]).all() and np.isnan(ref[2050]).all()
    for layout, want in ((ERR_PARTICLE_MAJOR, ref), (ERR_GENE_MAJOR, ref.T)):
        eng.accept_reset()
        err, counts, _ = eng.score(s, eps=4.8, err_layout=layout)
        assert oracle.same_bits(err, want), layout
        assert np.array_equal(counts, (ref <= 4.8).sum(0))
    eng.accept_reset()
    _, counts, _ = eng.score(s, eps=4.8, err_layout=ERR_NONE)
    assert np.array_equal(counts, (ref <= 4.8).sum(0))
    off, idx, errs = eng.accept_fetch()
    for gg in range(0, d.shape[0], 53):
        want = oracle.accept_gene(ref[:, gg], 4.8)
        assert np.array_equal(idx[off[gg]:off[gg + 1]], want)
        assert oracle.same_bits(errs[off[gg]:off[gg + 1]], ref[want - 1, gg])


def test_score_sub_batches_on_two_streams(eng, data_stats):
    """large calls are split into sub-batches that alternate between two internal streams (and, above the queue
    limit, into several launches): the result must equal the one-stream result and the oracle"""
    d, se = data_stats
    rng = np.random.default_rng(33)
    n = 70000                                        # > 2 x 4 x 2048: the overlapped path
    s = synth_stats(rng, d, n)
    s[12345, 7] = np.nan
    eng.accept_reset()
    err, counts, _ = eng.score(s, eps=4.8, err_layout=ERR_PARTICLE_MAJOR)
    off, idx, errs = eng.accept_fetch()
    sub = rng.choice(n, 150, replace=False)
    ref = oracle.compute_trunc_errors(s[sub], d, se)
    assert oracle.same_bits(err[sub], ref)
    assert np.isnan(err[12345]).all() and np.nanmax(err) <= 10.0
    for sb in (1, 5):                                # one stream; five sub-batches
        eng.set_option("score_overlap", 1 if sb > 1 else 0)
        eng.set_option("score_sub_batches", sb if sb > 1 else 0)
        try:
            eng.accept_reset()
            err2, counts2, _ = eng.score(s, eps=4.8, err_layout=ERR_PARTICLE_MAJOR)
            off2, idx2, errs2 = eng.accept_fetch()
        finally:
            eng.set_option("score_overlap", 1)
            eng.set_option("score_sub_batches", 0)
        assert oracle.same_bits(err2, err) and np.array_equal(counts2, counts)
        assert np.array_equal(off2, off) and np.array_equal(idx2, idx) and oracle.same_bits(errs2, errs)
    eng.accept_reset()
    _, counts3, _ = eng.score(s, eps=4.8, err_layout=ERR_NONE)      # acceptance only prunes at eps, same lists
    off3, idx3, errs3 = eng.accept_fetch()
    assert np.array_equal(counts3, counts) and np.array_equal(idx3, idx) and oracle.same_bits(errs3, errs)
    eng.accept_reset()
    errg, countsg, _ = eng.score(s[:40000], eps=4.8, err_layout=ERR_GENE_MAJOR)
    assert oracle.same_bits(errg, err[:40000].T)


def test_score_tile_pruning_signed_and_degenerate_data(eng, data_stats):
    """data with negative entries in the 'non-negative' statistics, zero rows, tiny and huge magnitudes and a
    gene count that is not a multiple of the tile size"""
    d, se = data_stats
    rng = np.random.default_rng(32)
    G = 200 + 7
    d2 = d[rng.choice(d.shape[0], G, replace=False)].copy()
    se2 = se[rng.choice(se.shape[0], G, replace=False)].copy()
    d2[3, :10] *= -1.0
    d2[40] = 0.0
    se2[40] = 0.0
    d2[77] *= 1e-12
    se2[77] *= 1e-12
    d2[120] *= 1e9
    d2[150, 12] = -np.inf
    s = np.concatenate([d2[rng.integers(0, G, 900)] * np.exp(rng.normal(0, 0.2, (900, 53))),
                        10.0 ** rng.uniform(-14, 10, size=(600, 53)), np.zeros((3, 53))])
    eng.set_data(d2, se2)
    try:
        ref = oracle.compute_trunc_errors(s, d2, se2)
        assert (ref < 10.0).mean() > 0.001
        for layout, want in ((ERR_PARTICLE_MAJOR, ref), (ERR_GENE_MAJOR, ref.T)):
            eng.accept_reset()
            err, counts, _ = eng.score(s, eps=4.8, err_layout=layout)
            assert oracle.same_bits(err, want), layout
            assert np.array_equal(counts, (ref <= 4.8).sum(0))
    finally:
        eng.set_data(d, se)


def test_score_near_matches_fill_the_queues(eng, data_stats):
    """every particle close to the data of every gene of a small gene set: nothing is 'surely > 10', so all
    pairs travel through both queues (queue drains mid-tile are exercised)"""
    d, se = data_stats
    rng = np.random.default_rng(9)
    sub = slice(100, 164)
    eng.set_data(np.tile(d[100:101], (64, 1)) * np.exp(rng.normal(0, 0.01, (64, 53))), np.tile(se[100:101], (64, 1)))
    try:
        dd = np.tile(d[100:101], (64, 1)) * np.exp(np.random.default_rng(9).normal(0, 0.01, (64, 53)))
        ss = np.tile(se[100:101], (64, 1))
        s = d[100] * np.exp(rng.normal(0, 0.05, (1000, 53)))
        ref = oracle.compute_trunc_errors(s, dd, ss)
        assert (ref < 10.0).mean() > 0.9
        eng.accept_reset()
        err, counts, _ = eng.score(s, eps=4.8, err_layout=ERR_PARTICLE_MAJOR)
        assert oracle.same_bits(err, ref)
        assert np.array_equal(counts, (ref <= 4.8).sum(0))
        off, idx, _ = eng.accept_fetch()
        for g in range(0, 64, 7):
            assert np.array_equal(idx[off[g]:off[g + 1]], oracle.accept_gene(ref[:, g], 4.8))
    finally:
        eng.set_data(d, se)


def test_score_empty_and_small(eng, data_stats):
    d, se = data_stats
    eng.accept_reset()
    err, counts, _ = eng.score(np.zeros((0, 53)), err_layout=ERR_PARTICLE_MAJOR)
    assert err.shape == (0, d.shape[0]) and counts.sum() == 0
    s = d[:1].copy()
    err, counts, _ = eng.score(s, err_layout=ERR_PARTICLE_MAJOR)
    assert oracle.same_bits(err, oracle.compute_trunc_errors(s, d, se))


def test_score_linearity_property_large(eng, data_stats):
    """size-independent property at a larger n: scoring a concatenation == concatenating the scores"""
    d, se = data_stats
    rng = np.random.default_rng(4)
    s = synth_stats(rng, d, 20000)
    eng.accept_reset()
    full, c_full, _ = eng.score(s, err_layout=ERR_PARTICLE_MAJOR)
    eng.accept_reset()
    a, ca, _ = eng.score(s[:7001], err_layout=ERR_PARTICLE_MAJOR)
    eng.accept_reset()
    b, cb, _ = eng.score(s[7001:], err_layout=ERR_PARTICLE_MAJOR)
    assert oracle.same_bits(full, np.concatenate([a, b]))
    assert np.array_equal(c_full, ca + cb)
    sub = rng.choice(20000, 200, replace=False)
    assert oracle.same_bits(full[sub], oracle.compute_trunc_errors(s[sub], d, se))
    assert full.max() <= 10.0 and full.min() >= 0.0


# ------------------------------------------------------------------------------------------ statistics, prior
def test_summary_stats_bit_exact(eng):
    rng = np.random.default_rng(5)
    mom = np.abs(rng.lognormal(0, 2, size=(500, 11, 5, 5)))
    mom[..., 3] *= rng.choice([-1, 1], size=mom[..., 3].shape) * 0.1
    mom[0] = 0.0                                  # zero totals -> eps branch, 0/0 -> NaN
    ad = rng.dirichlet(np.ones(5), size=11).T
    des = eng.design
    import copy
    d2 = copy.copy(des)
    d2.age_dist = ad
    eng.set_design(d2)
    mom[1, 3, :, 1] = 0.0                          # condition 4: no labelled molecule in any age group
    mom[1, 3, :, 3:] = 0.0
    try:
        eng.set_option("stats_sample_guards", 0)
        got = eng.summary_stats(mom)
        eng.set_option("stats_sample_guards", 1)
        got_s = eng.summary_stats(mom)
    finally:
        eng.set_option("stats_sample_guards", -1)
        eng.set_design(des)
    want = oracle.summary_stats(mom, ad)           # abc_simulation.jl:23-46 verbatim (moment-ODE inputs)
    assert oracle.same_bits(got, want)
    assert np.isnan(want[0, 20:]).all()
    want_s = oracle.summary_stats(mom, ad, sample_guards=True)   # finite-sample conventions (SSA inputs)
    assert oracle.same_bits(got_s, want_s)
    assert (want_s[0, 20:] == 0).all() and want_s[1, 20 + 3] == 0.0 and want_s[1, 31 + 3] == 0.0 and np.isfinite(want_s).all()


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5])
def test_fix_params_bit_exact_and_in_box(eng, m):
    from abc_inference_transcription_b200 import prior_bounds
    P = n_params(m)
    th = eng.fix_params(m, 1000, particle_offset=12345, seed=99)
    want = np.stack([oracle.prior(m, 12345 + i, 99, P) for i in range(1000)])
    assert oracle.same_bits(th, want)
    lo, hi = prior_bounds(m)
    assert (th >= lo).all() and (th < hi + 1e-12).all()
    # counter-based: a shifted window reproduces the overlap
    th2 = eng.fix_params(m, 10, particle_offset=12350, seed=99)
    assert oracle.same_bits(th2, th[5:15])


# ------------------------------------------------------------------------------------------ SSA
DEMO = {1: np.log10([0.5, 1.0, 20.0, 0.1, 0.7]), 2: np.log10([0.5, 1.0, 20.0, 0.1, 0.7]),
        3: np.log10([0.1, 0.1, 30.0, 0.1, 0.1, 1.0, 15.0, 0.1, 0.7]),
        4: np.log10([1.0, 1.0, 1.0, 1.0, 40.0, 1.0, 1.0, 0.1, 0.7]),
        5: np.log10([2.0, 1.0, 8.5, 2.0, 3.0, 1.0, 1.0, 1.5, 0.7])}    # model_realisation.jl:294-314 (alpha/10 for m=5)


@pytest.mark.parametrize("m,cond,age", [(1, 5, 0), (2, 6, 4), (3, 2, 2), (4, 10, 0), (5, 8, 3)])
def test_ssa_exact_math_is_bit_identical_to_oracle(eng, betas, m, cond, age):
    sd, keep = oracle.make_ssa_design(96, 10, True, split_betas(betas))
    got = eng.ssa_cells(m, DEMO[m], particle_index=4242, cond=cond, age=age, seed=7, exact_math=True)
    want, ev = oracle.ssa_readout(DEMO[m], m, sd, 4242, 7, cond, age, oracle.MATH_DET)
    assert ev > 1000
    assert np.array_equal(got, want)


def test_ssa_fast_math_statistically_equivalent(eng, betas):
    """fast (MUFU) variant vs deterministic variant of the SAME algorithm (full direct method from the first
    cycle): most lineages identical, marginals pass KS"""
    from scipy.stats import ks_2samp
    eng.set_option("ssa_hybrid_burnin", 0)
    try:
        with cells_per_readout(eng, 4096):
            m, cond, age = 1, 7, 2
            fast = eng.ssa_cells(m, DEMO[m], particle_index=1, cond=cond, age=age, seed=11, exact_math=False)
            det = eng.ssa_cells(m, DEMO[m], particle_index=1, cond=cond, age=age, seed=11, exact_math=True)
            other = eng.ssa_cells(m, DEMO[m], particle_index=2, cond=cond, age=age, seed=11, exact_math=True)
    finally:
        eng.set_option("ssa_hybrid_burnin", 2)
    same = (fast == det).all(0).mean()
    assert same > 0.5, same
    for row in range(4):
        assert ks_2samp(fast[row], other[row]).pvalue > 1e-4
        assert ks_2samp(det[row], other[row]).pvalue > 1e-4


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("m,cond,age", [(1, 5, 0), (1, 9, 3), (2, 0, 4), (3, 7, 1), (4, 10, 0), (5, 6, 2), (5, 3, 3)])
def test_ssa_hybrid_burnin_equals_full_direct_method(eng, betas, m, cond, age, mode):
    """telegraph SSA + conditional Poisson sampling (mode 1: burn-in only; mode 2, the default: to the read-out) vs
    the full six-channel SSA from the first cycle (which is bit-identical to the oracle): two-sample KS on U, L, U',
    L' and on U+L, plus mean/variance z-scores"""
    from scipy.stats import ks_2samp
    n = 16384
    with cells_per_readout(eng, n):
        eng.set_option("ssa_hybrid_burnin", mode)
        hyb = eng.ssa_cells(m, DEMO[m], particle_index=3, cond=cond, age=age, seed=21, exact_math=False).astype(np.float64)
        eng.set_option("ssa_hybrid_burnin", 0)
        try:
            full = eng.ssa_cells(m, DEMO[m], particle_index=4, cond=cond, age=age, seed=21, exact_math=False).astype(np.float64)
        finally:
            eng.set_option("ssa_hybrid_burnin", 2)
    assert not np.array_equal(hyb, full)
    rows = [hyb[0], hyb[1], hyb[2], hyb[3], hyb[0] + hyb[1]], [full[0], full[1], full[2], full[3], full[0] + full[1]]
    for a, b in zip(*rows):
        assert ks_2samp(a, b).pvalue > 1e-4, (ks_2samp(a, b), a.mean(), b.mean())
        z = (a.mean() - b.mean()) / np.sqrt(a.var() / n + b.var() / n + 1e-300)
        assert abs(z) < 4.5, z
    # the covariance structure (same beta for U and L, shared gene history) is preserved
    c_h = np.cov(hyb[2], hyb[3])[0, 1]
    c_f = np.cov(full[2], full[3])[0, 1]
    pr = (hyb[2] - hyb[2].mean()) * (hyb[3] - hyb[3].mean())
    qr = (full[2] - full[2].mean()) * (full[3] - full[3].mean())
    assert abs(c_h - c_f) / np.sqrt(pr.var() / n + qr.var() / n + 1e-300) < 4.5


CORNERS = [
    (4, np.array([2.5, 2.2, 1.0, 1.5, 2.0, 2.5, 3.0, 1.5, -0.1]), 9, 0),      # config 5: high-rate corner, alpha steps
    (5, np.array([2.0, 2.8, 2.6, 1.0, 1.3, 1.6, 1.9, 2.0, -0.3]), 4, 4),      # config 5: fast decay steps
    (1, np.array([-2.7, -2.9, 2.9, -2.8, -0.5]), 6, 1),                       # slow switch, long-lived mRNA: Poisson(1e4)
    (2, np.array([-1.0, 1.0, 2.0, -1.0, 0.0]), 8, 2),                         # bursty, no scaling, lambda = 1
    (3, np.array([-3.0, 3.0, -3.0, 3.0, 0.0, 0.5, 2.0, -1.5, -0.7]), 10, 3),  # kon jumps by 6 decades between steps
    (5, np.array([1.0, 1.5, 2.0, -2.5, -1.6, -1.1, 0.0, 1.5, -0.2]), 7, 3),   # decay steps on both sides of the series/closed-form split
    (1, np.array([-0.5, 2.9, 2.5, -3.0, -0.4]), 2, 4),                        # P_on = 4e-4, immortal mRNA: short on-stretches, slow decay
]


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("m,theta,cond,age", CORNERS)
def test_ssa_hybrid_burnin_corner_cases(eng, m, theta, cond, age, mode):
    from scipy.stats import ks_2samp
    n = 8192
    with cells_per_readout(eng, n):
        eng.set_option("ssa_hybrid_burnin", mode)
        hyb = eng.ssa_cells(m, theta, particle_index=11, cond=cond, age=age, seed=5, exact_math=False).astype(np.float64)
        eng.set_option("ssa_hybrid_burnin", 0)
        try:
            full = eng.ssa_cells(m, theta, particle_index=12, cond=cond, age=age, seed=5, exact_math=False).astype(np.float64)
        finally:
            eng.set_option("ssa_hybrid_burnin", 2)
    for a, b in zip([hyb[0], hyb[1], hyb[2], hyb[3], hyb[0] + hyb[1]], [full[0], full[1], full[2], full[3], full[0] + full[1]]):
        assert ks_2samp(a, b).pvalue > 1e-4, (ks_2samp(a, b), a.mean(), b.mean())
        assert abs(a.mean() - b.mean()) / np.sqrt(a.var() / n + b.var() / n + 1e-300) < 4.5


def test_ssa_hybrid_vs_full_over_prior_particles(eng):
    """all 55 read-outs of 16 prior particles per model family: hybrid vs full direct method, z-scores of the
    per-read-out means from the sample variances; calibrated (mean z^2 ~ 1) and without outliers"""
    n = 2048
    zs = []
    with cells_per_readout(eng, n):
        for m in (1, 4):
            th = eng.fix_params(m, 16, particle_offset=900, seed=77)
            mh, _ = eng.simulate_moments(m, th, particle_offset=900, seed=77)
            eng.set_option("ssa_hybrid_burnin", 0)
            try:
                mf, _ = eng.simulate_moments(m, th, particle_offset=5000, seed=78)
            finally:
                eng.set_option("ssa_hybrid_burnin", 2)
            for q, v in ((0, 2), (1, 4)):
                se = np.sqrt((mh[..., v] + mf[..., v]) / n)
                ok = se > 0
                zs.append(((mh[..., q] - mf[..., q])[ok] / se[ok]).ravel())
    zs = np.concatenate(zs)
    assert len(zs) > 2000
    assert np.abs(zs).max() < 5.5, np.abs(zs).max()
    assert 0.8 < (zs ** 2).mean() < 1.25, (zs ** 2).mean()


@pytest.mark.parametrize("m", [1, 3, 5])
def test_ssa_moments_match_moment_odes(eng, betas, m):
    """z-tests of SSA sample moments against the reference's moment ODEs (oracle pinned on the goldens)"""
    # 32 768 cells: the z-statistics of the (heavy-tailed) covariance products are then close to Gaussian.  With 8192
    # cells one of 54 (model, mode, Philox stream) runs of scripts/diag_moments_z.py gave |z| = 5.04 for a down-sampled
    # cov_ul while the other streams of the same statistic average to z = -0.05 (no bias).
    with cells_per_readout(eng, 32768):
        cells = {(c, a): eng.ssa_cells(m, DEMO[m], particle_index=5, cond=c, age=a, seed=3).astype(np.float64)
                 for c, a in [(5, 0), (6, 2), (9, 4), (0, 3)]}
    od = oracle.make_design(iv_index=1, downsampling=True, betas=split_betas(betas), rtol=1e-9)
    od_raw = oracle.make_design(iv_index=1, downsampling=False, rtol=1e-9)
    _, mom_ds = oracle.run_sim(DEMO[m], m, od)
    mom_raw, _ = oracle.run_part_sim(DEMO[m], m, od_raw)
    zs = []
    for (c, a), x in cells.items():
        for (u, l), ref in [((x[0], x[1]), mom_raw[c, a]), ((x[2], x[3]), mom_ds[c, a])]:
            n = len(u)
            for sample, target in [(u, ref[0]), (l, ref[1])]:
                zs.append((sample.mean() - target) / (sample.std(ddof=1) / np.sqrt(n) + 1e-12))
            for xs, ys, target in [(u, u, ref[2]), (u, l, ref[3]), (l, l, ref[4])]:
                p = (xs - xs.mean()) * (ys - ys.mean())
                zs.append((p.sum() / (n - 1) - target) / (p.std(ddof=1) / np.sqrt(n) + 1e-12))
    zs = np.array(zs)
    # burn-in bias bound 2^-10 and the ODE transient criterion (1 %) are far below these tolerances
    assert np.abs(zs).max() < 5.0, zs
    assert (zs ** 2).mean() < 2.5, zs


def test_simulate_is_partition_invariant_and_reproducible(eng):
    """identical bits for any split of the particle range (Philox keyed by global ids)"""
    m = 3
    th, st, cnt = eng.simulate(m, n_trials=24, particle_offset=1000, seed=5)
    assert cnt["n_lineages"] == 24 * 55 * 96 and cnt["n_events"] > 0
    th_a, st_a, _ = eng.simulate(m, n_trials=10, particle_offset=1000, seed=5)
    th_b, st_b, _ = eng.simulate(m, n_trials=14, particle_offset=1010, seed=5)
    assert oracle.same_bits(th, np.concatenate([th_a, th_b]))
    assert oracle.same_bits(st, np.concatenate([st_a, st_b]))
    # supplying theta reproduces the prior-drawn run
    _, st_c, _ = eng.simulate(m, theta=th, particle_offset=1000, seed=5)
    assert oracle.same_bits(st, st_c)
    assert np.isfinite(st[:, :20]).all()


def test_simulate_statistics_equal_oracle_ssa_pipeline(eng, betas):
    """whole pipeline for a few prior particles vs the oracle SSA (fast math => compare moments loosely,
    then S1 bit-exact on the GPU's own moments)"""
    m = 1
    th = eng.fix_params(m, 3, particle_offset=77, seed=1)
    mom, _ = eng.simulate_moments(m, th, particle_offset=77, seed=1)
    _, st, _ = eng.simulate(m, theta=th, particle_offset=77, seed=1)
    want = oracle.summary_stats(mom, eng.design.age_dist, sample_guards=True)
    assert oracle.same_bits(st, want)


def test_device_exports_for_multi_gpu_gather(eng, data_stats):
    """abc_counts_dev / abc_accept_tuples_dev (what dist.gather_acceptance feeds to NCCL) == host fetch"""
    import torch
    from abc_inference_transcription_b200.dist import csr_from_tuples, gather_acceptance
    d, se = data_stats
    rng = np.random.default_rng(31)
    s = synth_stats(rng, d, 900)
    eng.accept_reset()
    _, counts, _ = eng.score(s, eps=4.8, particle_offset=10_000, err_layout=ERR_NONE)
    off, idx, errs = eng.accept_fetch()
    dev = torch.device("cuda", 0)
    c = torch.zeros(d.shape[0], dtype=torch.int64, device=dev)
    eng.counts_dev(c.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(c.cpu().numpy(), counts)
    tot = eng.accept_total()
    g = torch.zeros(tot, dtype=torch.int32, device=dev)
    p = torch.zeros(tot, dtype=torch.int64, device=dev)
    e = torch.zeros(tot, dtype=torch.float64, device=dev)
    eng.accept_tuples_dev(g.data_ptr(), p.data_ptr(), e.data_ptr(), tot)
    torch.cuda.synchronize()
    o2, i2, e2 = csr_from_tuples(g.cpu().numpy(), p.cpu().numpy(), e.cpu().numpy(), d.shape[0])
    assert np.array_equal(o2, off) and np.array_equal(i2, idx) and oracle.same_bits(e2, errs)
    assert idx.min() > 10_000                      # global, 1-based particle indices
    res = gather_acceptance(eng, 1)
    assert np.array_equal(res["offsets"], off) and np.array_equal(res["idx"], idx) and np.array_equal(res["counts"], counts)


def test_pinned_output_buffer(eng, data_stats):
    """abc_host_alloc'ed (page-locked) output buffers give the same bits as pageable ones"""
    from abc_inference_transcription_b200 import PinnedArray
    d, se = data_stats
    s = synth_stats(np.random.default_rng(41), d, 300)
    pin = PinnedArray((300, d.shape[0]))
    eng.accept_reset()
    err_p, _, _ = eng.score(s, err_layout=ERR_PARTICLE_MAJOR, out=pin.array)
    assert err_p is pin.array
    eng.accept_reset()
    err, _, _ = eng.score(s, err_layout=ERR_PARTICLE_MAJOR)
    assert oracle.same_bits(err, err_p)


@pytest.mark.parametrize("m,cond,age", [(1, 8, 2), (4, 5, 4), (3, 10, 1)])
def test_ssa_hybrid_high_power(eng, m, cond, age):
    """262 144 cells per arm: means and variances of U, L, U', L' of the hybrid burn-in vs the full direct method
    agree within 4.5 standard errors (SE ~ 0.2 % of the CV), i.e. no bias above ~1 % survives"""
    n = 262144
    with cells_per_readout(eng, n):
        hyb = eng.ssa_cells(m, DEMO[m], particle_index=21, cond=cond, age=age, seed=99, exact_math=False).astype(np.float64)
        eng.set_option("ssa_hybrid_burnin", 0)
        try:
            full = eng.ssa_cells(m, DEMO[m], particle_index=22, cond=cond, age=age, seed=99, exact_math=False).astype(np.float64)
        finally:
            eng.set_option("ssa_hybrid_burnin", 2)
    for a, b in zip(hyb, full):
        z_mean = (a.mean() - b.mean()) / np.sqrt(a.var() / n + b.var() / n + 1e-300)
        da, db = (a - a.mean()) ** 2, (b - b.mean()) ** 2
        z_var = (da.mean() - db.mean()) / np.sqrt(da.var() / n + db.var() / n + 1e-300)
        assert abs(z_mean) < 4.5 and abs(z_var) < 4.5, (z_mean, z_var, a.mean(), b.mean())
    # U and L are drawn independently GIVEN the gene path; their covariance comes from the shared path (and beta)
    for i, j in ((0, 1), (2, 3)):
        pa = (hyb[i] - hyb[i].mean()) * (hyb[j] - hyb[j].mean())
        pb = (full[i] - full[i].mean()) * (full[j] - full[j].mean())
        z_cov = (pa.mean() - pb.mean()) / np.sqrt(pa.var() / n + pb.var() / n + 1e-300)
        assert abs(z_cov) < 4.5, (z_cov, pa.mean(), pb.mean())


@pytest.mark.parametrize("rule", [1, 2])
@pytest.mark.parametrize("m,theta,cond,age", [
    (1, np.array([1.0, 1.2, 2.0, 0.0, -0.3]), 8, 1),                               # gamma = 1/h: one pre-cycle (two for the chase)
    (4, np.array([0.5, 0.0, 1.0, 1.5, 2.0, 1.5, 1.0, -0.9, -0.2]), 5, 3),          # gamma = 0.126/h: three pre-cycles
    (5, np.array([0.0, 0.5, 2.0, -2.0, -1.0, 0.0, -1.5, -0.5, -0.1]), 2, 0),       # decay steps: sum over the cycle decides
    (3, np.array([1.0, 0.5, 0.0, 1.5, 2.0, 0.7, 1.5, 0.3, -0.4]), 9, 2),           # model 3: gene memory is part of the bound
    (4, np.array([2.5, 2.2, 2.0, 2.9, 2.4, 2.1, 2.7, 1.3, -0.05]), 6, 0),          # BASELINE configs[4] corner, 22 h pulse, lambda = 0.89
    (5, np.array([2.1, 2.8, 2.6, 1.0, 1.9, 1.4, 1.1, 1.7, -0.6]), 10, 3),          # corner, decay steps, 6 h chase
    (2, np.array([-1.0, -0.5, 1.5, -2.5, -0.3]), 3, 4),                            # slow gene, near-immortal mRNA: no shortening possible
])
def test_ssa_adaptive_burnin_keeps_the_bias_bound(eng, m, theta, cond, age, rule):
    """ssa_adaptive_burnin = 1 (whole cycles per particle) and = 2 (start time per (particle, read-out), default) against the
    full n_pre_cycles burn-in: fewer draws, means / variances / covariance of U, L, U', L' within 4.5 standard errors at
    262 144 cells (SE ~ 0.2-0.4 %)"""
    n = 262144
    with cells_per_readout(eng, n):
        eng.set_option("ssa_adaptive_burnin", rule)
        try:
            ada = eng.ssa_cells(m, theta, particle_index=31, cond=cond, age=age, seed=17, exact_math=False).astype(np.float64)
            ev_ada = eng.counters()["n_draws"]
            eng.set_option("ssa_adaptive_burnin", 0)
            full = eng.ssa_cells(m, theta, particle_index=32, cond=cond, age=age, seed=17, exact_math=False).astype(np.float64)
            ev_full = eng.counters()["n_draws"]
        finally:
            eng.set_option("ssa_adaptive_burnin", 2)
    if m != 2:
        assert ev_ada < 0.7 * ev_full, (ev_ada, ev_full)
    for a, b in zip(ada, full):
        z_mean = (a.mean() - b.mean()) / np.sqrt(a.var() / n + b.var() / n + 1e-300)
        da, db = (a - a.mean()) ** 2, (b - b.mean()) ** 2
        z_var = (da.mean() - db.mean()) / np.sqrt(da.var() / n + db.var() / n + 1e-300)
        assert abs(z_mean) < 4.5 and abs(z_var) < 4.5, (z_mean, z_var, a.mean(), b.mean())
    for i, j in ((0, 1), (2, 3)):
        pa = (ada[i] - ada[i].mean()) * (ada[j] - ada[j].mean())
        pb = (full[i] - full[i].mean()) * (full[j] - full[j].mean())
        assert abs(pa.mean() - pb.mean()) / np.sqrt(pa.var() / n + pb.var() / n + 1e-300) < 4.5


def test_ssa_start_times_match_the_quadrature_restatement(eng):
    """abc_window_kernel (closed forms per schedule piece + bisection) against tests/window_rule.py (quadrature on a 7 s grid)
    on prior draws of all five models and all 55 read-outs: same start time, the bias bound 2^-n_pre holds there for the
    unlabelled and the labelled Poisson mean, and it is tight (starting 2 % of the simulated span + 0.05 h later breaks it)"""
    import window_rule as wr
    rng = np.random.default_rng(11)
    eps = 2.0 ** -10
    n_checked = n_short = 0
    for m in range(1, 6):
        theta = eng.fix_params(m, 6, particle_offset=900 + m, seed=5)
        theta[0, -1] = 0.0                                  # lambda = 1: no unlabelled births inside the window
        for th in theta:
            starts, draws = eng.ssa_window(m, th)
            assert starts.shape == (11, 5) and draws > 0
            for cond in range(11):
                for age_i in rng.choice(5, size=2, replace=False):
                    s_dev = float(starts[cond, age_i])
                    age = wr.AGES[age_i]
                    assert -200.0 <= s_dev <= age
                    want = wr.burnin_window(th, m, cond, age_i)
                    span = age - min(s_dev, want)
                    assert abs(s_dev - want) <= 0.01 + 2e-3 * span, (m, cond, age_i, s_dev, want)
                    mu, ku, ml, kl = wr.missing_share(th, m, cond, age_i, s_dev)
                    # resolution of the two methods (grid 0.002 h, bisection 4 h / 4096): the share moves by exp(gamma ds)
                    slack = 1.02 * np.exp(wr.rates_of(th, m)[3].max() * 0.006)
                    assert mu <= max(eps * ku, 2.0 ** -30) * slack and ml <= max(eps * kl, 2.0 ** -30) * slack
                    if s_dev > -200.0 + 1e-3 and m != 3:    # tight (model 3 starts earlier: gene memory)
                        later = s_dev + 0.02 * (age - s_dev) + 0.05
                        if later < age:
                            mu, ku, ml, kl = wr.missing_share(th, m, cond, age_i, later)
                            assert mu > max(eps * ku, 2.0 ** -30) or ml > max(eps * kl, 2.0 ** -30), (m, cond, age_i, s_dev)
                        n_short += 1
                    n_checked += 1
    assert n_checked == 5 * 6 * 22 and n_short > 100


def test_ssa_refuses_absurd_rates_instead_of_hanging(eng):
    """rates far outside any prior box (10^33 switches per hour) would never finish a lineage: the particle is refused, its
    statistics are NaN and it is never accepted; its neighbours in the batch are unaffected"""
    theta = eng.fix_params(1, 4, particle_offset=0, seed=3)
    want = eng.simulate(1, theta=theta, particle_offset=0, seed=3)[1]
    bad = theta.copy()
    bad[2, 0] = 33.0
    bad[2, 1] = 33.0
    _, stats, _ = eng.simulate(1, theta=bad, particle_offset=0, seed=3)
    assert np.isnan(stats[2]).sum() >= 40            # (the sample guards turn the 11 ratios of an empty sample into 0)
    assert oracle.same_bits(stats[[0, 1, 3]], want[[0, 1, 3]])
    eng.accept_reset()
    err, counts, _ = eng.score(stats, eps=4.8, err_layout=ERR_PARTICLE_MAJOR)
    assert np.isnan(err[2]).all()                    # NaN error: neither clipped nor accepted (compute_errors.jl:62-64)
    _, idx, _ = eng.accept_fetch()
    assert 3 not in idx


@pytest.mark.parametrize("layout", [ERR_PARTICLE_MAJOR, ERR_GENE_MAJOR, ERR_NONE])
def test_simulate_score_pipelined_equals_separate_calls(eng, data_stats, layout):
    """abc_simulate_score (sub-batches pipelined over two streams) == abc_simulate followed by abc_score, bit for bit:
    theta, statistics, error matrix, counts and the accepted lists; prior drawn and prior supplied; ragged sizes"""
    for n, supplied in ((16400, False), (4097, True), (300, False)):
        eng.accept_reset()
        th0, st0, _ = eng.simulate(3, n_trials=n, particle_offset=777, seed=5)
        err0, cnt0, _ = eng.score(st0, eps=4.8, particle_offset=777, err_layout=layout)
        off0, idx0, e0 = eng.accept_fetch()
        eng.accept_reset()
        th1, st1, err1, cnt1, c = eng.simulate_score(3, n_trials=n, theta=th0 if supplied else None, particle_offset=777, seed=5,
                                                     eps=4.8, err_layout=layout)
        off1, idx1, e1 = eng.accept_fetch()
        assert c["n_particles"] == n and c["n_events"] > 0
        assert np.array_equal(th0, th1) and oracle.same_bits(st0, st1)
        if layout == ERR_NONE:
            assert err0 is None and err1 is None
        else:
            assert oracle.same_bits(err0, err1)
        assert np.array_equal(cnt0, cnt1) and np.array_equal(off0, off1) and np.array_equal(idx0, idx1) and oracle.same_bits(e0, e1)


def _julia_quantile(v, p):
    """Statistics.jl quantile(v, p) (alpha = beta = 1) on sorted v"""
    n = len(v)
    if n == 1:
        return v[0]
    aleph = n * p + (1.0 - p)
    j = min(max(int(np.trunc(aleph)), 1), n - 1)
    gam = min(max(aleph - j, 0.0), 1.0)
    return v[j - 1] + gam * (v[j] - v[j - 1])


@pytest.mark.parametrize("m,offset", [(1, 0), (4, 123456)])
def test_posterior_summary_on_device(eng, data_stats, m, offset):
    """SURVEY 8f-3 (posterior_kinetics.jl:10-33) on the device: MAP, mean, quantiles per gene over the device-ordered lists,
    bit-exact against a numpy restatement of the same arithmetic and within 1e-12 of posteriors.py (numpy mean / quantile)"""
    from abc_inference_transcription_b200 import posteriors
    n = 3000
    eng.accept_reset()
    theta, stats, _ = eng.simulate(m, n_trials=n, particle_offset=offset, seed=9)
    eng.score(stats, eps=4.8, particle_offset=offset, err_layout=ERR_NONE)
    offsets, idx, _ = eng.accept_fetch()
    got = eng.posterior_summary(theta, particle_offset=offset, q=0.95)
    assert np.array_equal(got["n"], np.diff(offsets)) and got["n"].sum() > 1000
    P = theta.shape[1]
    genes = np.nonzero(got["n"] > 0)[0]
    assert len(genes) > 100 and (got["n"] > 5).any()
    for g in genes:
        rows = theta[idx[offsets[g]:offsets[g + 1]] - 1 - offset]
        assert oracle.same_bits(got["map"][g], rows[0])
        assert oracle.same_bits(got["mean"][g], np.cumsum(rows, axis=0)[-1] / len(rows))      # summed in list order
        srt = np.sort(rows, axis=0)
        lo = np.array([_julia_quantile(srt[:, p], 1.0 - 0.95) for p in range(P)])
        hi = np.array([_julia_quantile(srt[:, p], 0.95) for p in range(P)])
        assert oracle.same_bits(got["lo"][g], lo) and oracle.same_bits(got["hi"][g], hi)
    empty = np.nonzero(got["n"] == 0)[0]
    assert all(np.isnan(got[k][empty]).all() for k in ("map", "mean", "lo", "hi"))
    # the host-side mirror (numpy mean / numpy quantile) agrees to rounding
    gv = genes[:200] + 1
    ref_map = posteriors.get_posterior_estimate(theta, offsets, idx - offset, gv, "map")
    ref_mean = posteriors.get_posterior_estimate(theta, offsets, idx - offset, gv, "mean")
    ref_lo, ref_hi = posteriors.get_posterior_ci(theta, offsets, idx - offset, gv, 0.95)
    assert np.array_equal(ref_map, got["map"][gv - 1])
    assert np.allclose(ref_mean, got["mean"][gv - 1], rtol=1e-12, atol=1e-12)
    assert np.allclose(ref_lo, got["lo"][gv - 1], rtol=1e-12, atol=1e-12) and np.allclose(ref_hi, got["hi"][gv - 1], rtol=1e-12, atol=1e-12)
    # a theta window that does not cover the accepted indices is an error, not a silent gather
    with pytest.raises(AbcError, match="outside"):
        eng.posterior_summary(theta[:10], particle_offset=offset, q=0.95)


# ------------------------------------------------------------------------------------------ multi-GPU behind the C ABI
def _multi_vs_single(eng, betas, data_stats, n_dev, n, m, layout):
    from abc_inference_transcription_b200 import AbcMulti
    d, se = data_stats
    with AbcMulti(n_dev=n_dev) as mg:
        mg.set_design(synthetic_design(betas, n_cells=96, n_pre_cycles=10))
        mg.set_data(d, se)
        mg.accept_reset()
        th, st, err, counts, cnt = mg.simulate_score(m, n_trials=n, particle_offset=4321, seed=7, eps=4.8, err_layout=layout)
        off, idx, errs = mg.accept_fetch()
    eng.accept_reset()
    th1, st1, err1, counts1, cnt1 = eng.simulate_score(m, n_trials=n, particle_offset=4321, seed=7, eps=4.8, err_layout=layout)
    off1, idx1, errs1 = eng.accept_fetch()
    assert oracle.same_bits(th, th1) and oracle.same_bits(st, st1)
    if layout != ERR_NONE:
        assert oracle.same_bits(err, err1)
    assert np.array_equal(counts, counts1) and cnt["n_lineages"] == cnt1["n_lineages"] and cnt["n_events"] == cnt1["n_events"]
    assert np.array_equal(off, off1) and np.array_equal(idx, idx1) and oracle.same_bits(errs, errs1)
    assert off[-1] > 0


@pytest.mark.parametrize("layout", [ERR_PARTICLE_MAJOR, ERR_NONE])
def test_multi_context_on_one_device_equals_the_engine(eng, betas, data_stats, layout):
    """abc_multi_* with a single device (no NCCL involved): same bits as abc_simulate_score / abc_accept_fetch"""
    _multi_vs_single(eng, betas, data_stats, 1, 301, 1, layout)


@pytest.mark.parametrize("layout,m", [(ERR_PARTICLE_MAJOR, 1), (ERR_GENE_MAJOR, 4), (ERR_NONE, 3)])
def test_multi_gpu_from_one_process_is_bit_identical(eng, betas, data_stats, layout, m):
    """abc_multi_create over all GPUs of the box (one host thread + one NCCL rank per device inside the library): theta,
    statistics, error matrix, counts and the merged per-gene lists equal the single-GPU result bit for bit (odd batch size:
    uneven shards; the lists go through the gene-range all-to-all and the per-range merge)"""
    import torch
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    _multi_vs_single(eng, betas, data_stats, n_dev, 1501, m, layout)


# ------------------------------------------------------------------------------------------ SURVEY 8f-2 on the device
def test_model_probs_on_device_equal_the_restatement(eng):
    """abc_model_probs (bootstrap of the per-gene acceptance counts, model_probs.jl:1-54) against the numpy restatement on
    the same Philox draws: probabilities, lower and upper bounds bit for bit; K = 2 (constant vs non-constant), 3 and 5"""
    rng = np.random.default_rng(5)
    for K, G in ((2, 40), (3, 25), (5, 12)):
        counts = (rng.pareto(1.0, size=(K, G)) * 20).astype(np.int64)
        counts[:, 0] = 0                               # nothing accepted
        counts[:, 1] = 0; counts[K - 1, 1] = 17        # a single hypothesis accepted
        counts[:, 2] = [1] + [0] * (K - 2) + [1]       # two labels only
        counts = np.minimum(counts, 400)
        got = eng.model_probs(counts, n_bootstraps=100, alpha=0.95, seed=20240229)
        want = oracle.model_probs(counts, 100, 0.95, 20240229)
        for a, b, name in zip(got, want, ("prob", "l_bound", "u_bound")):
            assert oracle.same_bits(a, b), (K, name, np.abs(a - b).max())
    # reproducible, and a different seed gives different bounds
    a = eng.model_probs(counts, seed=1)
    b = eng.model_probs(counts, seed=1)
    c = eng.model_probs(counts, seed=2)
    assert oracle.same_bits(a[1], b[1]) and not oracle.same_bits(a[1], c[1]) and oracle.same_bits(a[0], c[0])


def test_model_probs_from_scored_counts(eng, data_stats):
    """end to end: per-gene counts of two scored 'models' -> device bootstrap; large counts (no restatement: property checks)"""
    d, se = data_stats
    rng = np.random.default_rng(6)
    counts = []
    for k in range(2):
        eng.accept_reset()
        _, c, _ = eng.score(synth_stats(rng, d, 3000), eps=4.8, err_layout=ERR_NONE)
        counts.append(c)
    counts = np.stack(counts)
    prob, lb, ub = eng.model_probs(counts, n_bootstraps=100, alpha=0.95, seed=9)
    tot = counts.sum(0)
    both = (counts > 0).all(0)
    assert both.sum() > 50
    assert np.allclose(prob[both].sum(1), 1.0) and np.array_equal(prob[both][:, 0], (counts[0] / tot)[both])
    assert np.all(lb[both] <= prob[both] + 0.2) and np.all(ub[both] >= prob[both] - 0.2) and np.all(lb <= ub)
    none = tot == 0
    assert not prob[none].any() and not ub[none].any()


# ------------------------------------------------------------------------------------------ data-side statistics
def _synthetic_cells(rng, n_cells, G):
    """per-cell counts shaped like the experiment: 12 experiment ids (0 = unused control, 1..6 pulse, 7..11 chase)"""
    experiment = rng.integers(0, 12, n_cells).astype(np.int32)
    age = rng.integers(1, 6, n_cells).astype(np.int32)
    lam = rng.lognormal(0.0, 1.0, size=(G, 1))
    u = rng.poisson(lam * (1.0 + 0.2 * age), size=(G, n_cells)).astype(np.float64)
    l = rng.poisson(lam * 0.3 * (experiment > 0), size=(G, n_cells)).astype(np.float64)
    u[0] = 0; l[0] = 0                               # a silent gene: every guard of the formulas
    l[1] = 0
    cond_vec = np.arange(1, 12, dtype=np.int32)
    pulse_idx = (np.nonzero((experiment >= 0) & (experiment <= 6))[0] + 1).astype(np.int32)      # load_process_data.jl:72
    chase_idx = (np.nonzero(experiment >= 7)[0] + 1).astype(np.int32)
    age_id_dist = rng.dirichlet(np.ones(5), size=11).T / 11.0                                     # columns sum to 1/11 (R12)
    return u, l, age, experiment, cond_vec, pulse_idx, chase_idx, age_id_dist


def test_data_summary_stats_equal_the_restatement(eng):
    """abc_data_summary_stats (get_summary_stats of data_summary_statistics.jl for every gene, bootstrap SEs from Philox
    resamples) against the numpy restatement on the same draws: d and se bit for bit, incl. a silent gene, a gene without
    labelled counts, an unused experiment id and clusters that run empty in small resamples"""
    rng = np.random.default_rng(8)
    u, l, age, experiment, cond_vec, pulse_idx, chase_idx, ad = _synthetic_cells(rng, 150, 7)
    d, se = eng.data_summary_stats(u, l, age, experiment, cond_vec, pulse_idx, chase_idx, ad, n_bootstraps=12, seed=77)
    wd, wse = oracle.data_summary_stats(u, l, age, experiment, cond_vec, pulse_idx, chase_idx, ad, 12, 77)
    assert oracle.same_bits(d, wd), np.nonzero(bits(d) != bits(wd))
    assert oracle.same_bits(se, wse), np.nonzero(bits(se) != bits(wse))
    # the silent gene: zeros, except the correlations of a (condition, age) bin with a single cell (Julia's var of one value is NaN)
    assert not np.nan_to_num(d[0]).any() and not np.nan_to_num(se[0]).any() and not np.isnan(d[0][:31]).any()


def test_data_summary_stats_at_the_experiments_size(eng):
    """5422 cells x 64 genes x 100 bootstraps: point estimates against plain numpy means / variances, SEs against the
    textbook sigma / sqrt(n), and the result feeds abc_set_data"""
    rng = np.random.default_rng(9)
    u, l, age, experiment, cond_vec, pulse_idx, chase_idx, ad = _synthetic_cells(rng, 5422, 64)
    d, se = eng.data_summary_stats(u, l, age, experiment, cond_vec, pulse_idx, chase_idx, ad, n_bootstraps=100, seed=3)
    t = u + l
    for g in (2, 17, 63):
        for f, idx in ((0, pulse_idx), (1, chase_idx)):
            for c in range(5):
                x = t[g, idx - 1][age[idx - 1] == c + 1]
                assert d[g, 10 * f + c] == pytest.approx(x.mean(), rel=1e-13)
                assert d[g, 10 * f + 5 + c] == pytest.approx(x.var(ddof=1) / x.mean(), rel=1e-11)
                assert se[g, 10 * f + c] == pytest.approx(x.std(ddof=1) / np.sqrt(len(x)), rel=0.35)
        for j in range(11):
            sel = experiment == cond_vec[j]
            assert d[g, 20 + j] == pytest.approx(l[g, sel].mean() / (u[g, sel].mean() + l[g, sel].mean()), rel=1e-13)
    assert np.isfinite(d).all() and np.isfinite(se).all() and (se[2:] >= 0).all()
    from abc_inference_transcription_b200 import AbcEngine
    with AbcEngine(0) as e2:
        e2.set_data(d, se)                             # the scoring kernel accepts it as its data statistics
        assert e2.n_genes == 64


@pytest.mark.parametrize("layout", [ERR_PARTICLE_MAJOR, ERR_GENE_MAJOR, ERR_NONE])
def test_simulate_score_async_equals_the_blocking_call(eng, data_stats, layout):
    """abc_simulate_score_async + abc_wait over the five models with three batches in flight in turn (two buffer sets: the
    third call waits for the first) == abc_simulate_score per model: theta, statistics, error matrices, counts, accepted
    lists bit for bit; a blocking call issued while batches are in flight completes them first"""
    from abc_inference_transcription_b200 import PinnedArray
    G, n = eng.n_genes, 700
    shape = None if layout == ERR_NONE else ((n, G) if layout == ERR_PARTICLE_MAJOR else (G, n))
    eng.accept_reset()
    want = []
    for m in range(1, 6):
        th, st, err, counts, _ = eng.simulate_score(m, n_trials=n, particle_offset=100 * m, seed=5, eps=4.8, err_layout=layout)
        want.append((th, st, None if err is None else err.copy()))
    off_w, idx_w, errs_w = eng.accept_fetch()
    counts_w = counts
    eng.accept_reset()
    bufs = [(PinnedArray((n, n_params(m))), PinnedArray((n, 53)), PinnedArray(shape) if shape else None) for m in range(1, 6)]
    for m in range(1, 6):
        th, st, er = bufs[m - 1]
        eng.simulate_score_async(m, th.array, st.array, None if er is None else er.array, prior_supplied=False,
                                 particle_offset=100 * m, seed=5, eps=4.8, err_layout=layout)
    counts, cnt = eng.wait()
    assert cnt["n_particles"] == 5 * n and cnt["n_lineages"] == 5 * n * 55 * 96
    for m in range(1, 6):
        th, st, er = bufs[m - 1]
        assert oracle.same_bits(th.array, want[m - 1][0]) and oracle.same_bits(st.array, want[m - 1][1])
        if er is not None:
            assert oracle.same_bits(er.array, want[m - 1][2])
    off, idx, errs = eng.accept_fetch()
    assert np.array_equal(counts, counts_w) and np.array_equal(off, off_w) and np.array_equal(idx, idx_w) and oracle.same_bits(errs, errs_w)
    # theta supplied, and a blocking call while one batch is in flight
    eng.accept_reset()
    th, st, er = bufs[0]
    th.array[:] = want[0][0]
    st.array[:] = 0.0
    eng.simulate_score_async(1, th.array, st.array, None if er is None else er.array, prior_supplied=True, particle_offset=100,
                             seed=5, eps=4.8, err_layout=layout)
    th2, st2, _, _, _ = eng.simulate_score(2, n_trials=n, particle_offset=200, seed=5, eps=4.8, err_layout=ERR_NONE)
    assert oracle.same_bits(st.array, want[0][1]) and oracle.same_bits(st2, want[1][1])
    eng.wait()
