"""GPU tests (-m gpu) of the tensor-core filter of the scoring path (csrc/abc_score3.cu, "tensor-core filter"): a TF32
tcgen05 GEMM decides which (particle, gene) pairs can have an error <= 10 and reach the FP64 stage.

1. the raw accumulators equal a numpy restatement of the GEMM (operands rounded to TF32 the same way, float64 sums) within
   the accumulation bound of the soundness argument -- this pins the operand layouts and descriptors;
2. soundness: every pair whose oracle error is below 10 has a negative accumulator;
3. the whole scoring suite of test_gpu_parity.py (bit-exact against the oracle: layouts, special values, signed and
   degenerate data, sub-batches, queues) passes with the option switched on.
"""
import os

import numpy as np
import pytest
import torch

import oracle
import test_gpu_parity as P
from abc_inference_transcription_b200 import AbcEngine, ERR_PARTICLE_MAJOR, synthetic_design

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def betas():
    return np.load(os.path.join(GOLD, "ref_betas.npy"))


@pytest.fixture(scope="module")
def data_stats():
    z = np.load(os.path.join(GOLD, "ref_summary_stats.npz"))
    return z["d"], z["se"]


@pytest.fixture(scope="module", params=[1, 2], ids=["sign_bit_words", "queue"])
def eng_mma(request, betas, data_stats):
    """score_mma_filter = 1: sign-bit words + mask-driven stage 3; 2: the filter kernel writes the stage-3 queue"""
    e = AbcEngine(0)
    e.set_design(synthetic_design(betas, n_cells=96, n_pre_cycles=10))
    e.set_data(*data_stats)
    e.set_option("score_mma_filter", request.param)
    e.mma_mode = request.param
    yield e
    e.close()


def tf32_rn(x):
    f = np.asarray(x, dtype=np.float32).copy()
    u = f.view(np.uint32)
    fin = (u & np.uint32(0x7F800000)) != np.uint32(0x7F800000)
    r = (u + np.uint32(0x0FFF) + ((u >> np.uint32(13)) & np.uint32(1))) & np.uint32(0xFFFFE000)
    u[fin] = r[fin]
    return f


def tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest, ties away from zero (on the magnitude)"""
    f = np.asarray(x, dtype=np.float32).copy()
    u = f.view(np.uint32)
    fin = (u & np.uint32(0x7F800000)) != np.uint32(0x7F800000)
    r = (u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    u[fin] = r[fin]
    return f


def operands(d, se):
    """csrc/abc_score3.cu abc_score_mma_build for genes with usable constants"""
    eps = np.where(se + d != 0.0, 0.0, 1e-4)
    den = (se * se + 0.1 * 0.1 * (d * d)) + eps
    w = 1.0 / (53.0 * den)
    R = (w * d * d).sum(1)
    Q = (np.sqrt(10.0) + np.sqrt(R)) ** 2
    X = 2.0 * np.sqrt(R * Q)
    slack = 1.002 * (X + Q) * 2.0 ** -10 + (R + X + Q) * 2.0 ** -14
    p = tf32_rn(-2.0 * w * d).astype(np.float64)
    q = tf32_rn(w).astype(np.float64)
    return p, q, R, slack, w


def test_accumulators_match_the_tf32_restatement_and_are_sound(eng_mma, data_stats):
    d, se = data_stats
    rng = np.random.default_rng(5)
    n = 300                                          # ragged: 2 tiles of 128 + 44
    s = P.synth_stats(rng, d, n)
    s[3, 7] = np.nan                                 # NaN row: never queued (the whole row is NaN in the matrix)
    s[5, 40] = np.inf                                # enters as 0
    s[6, 2] = 1e25                                   # square overflows binary32: enters as 0
    dev = torch.device("cuda", 0)
    st = torch.from_numpy(np.ascontiguousarray(s)).to(dev)
    cols = eng_mma.score_mma_columns()
    assert cols % 64 == 0 and cols >= d.shape[0]
    out = torch.full((384, cols), 7.0, dtype=torch.float32, device=dev)
    gene = eng_mma.score_mma_debug(st.data_ptr(), n, out.data_ptr())
    V = out.cpu().numpy().astype(np.float64)
    real = gene >= 0
    assert real.sum() == d.shape[0] and np.array_equal(np.sort(gene[real]), np.arange(d.shape[0]))
    # padding columns and invalid rows: +1
    assert (V[:n][:, ~real] == 1.0).all()
    assert (V[3] == 1.0).all() and (V[n:] == 1.0).all()
    p, q, R, slack, w = operands(d, se)
    s1 = s.copy()
    s2 = s * s
    with np.errstate(over="ignore", invalid="ignore"):
        a1 = tf32_rna(s1.astype(np.float32)).astype(np.float64)
        a2 = tf32_rna(s2.astype(np.float32)).astype(np.float64)
    a1[~(np.abs(a1) <= 3.0e38)] = 0.0
    a2[~(np.abs(a2) <= 3.0e38)] = 0.0
    c = V[0, :][real] * 0.0                          # the constant column is read back from a zero-statistics row below
    zero = torch.zeros((1, 53), dtype=torch.float64, device=dev)
    out0 = torch.empty((128, cols), dtype=torch.float32, device=dev)
    eng_mma.score_mma_debug(zero.data_ptr(), 1, out0.data_ptr())
    c = out0[0].cpu().numpy().astype(np.float64)[real]
    g = gene[real]
    thr = 10.0 + slack + 0.002
    assert (c <= (R - thr)[g] + 1e-12).all() and (c >= (R - thr)[g] - 2.0 ** -10 * np.abs(R - thr)[g] - 1e-6).all()
    rows = [i for i in range(n) if i != 3]
    want = a1[rows] @ p[g].T + a2[rows] @ q[g].T + c[None, :]
    mag = np.abs(a1[rows]) @ np.abs(p[g]).T + a2[rows] @ q[g].T + np.abs(c)[None, :]
    got = V[rows][:, real]
    fin = np.isfinite(want) & (mag < 1e30)
    assert fin.mean() > 0.95
    assert (np.abs(got - want)[fin] <= 2.0 ** -14 * mag[fin] + 1e-6).all(), np.abs((got - want) / mag)[fin].max()
    # soundness against the oracle's errors
    ref = oracle.compute_trunc_errors(s, d, se)
    below = ref[rows][:, g] < 10.0
    assert below.mean() > 0.001
    assert (got[below] < 0.0).all()
    queued = (got < 0.0).mean()
    assert queued < 1.5 * below.mean() + 0.01, (queued, below.mean())


SUITE = [P.test_score_bit_exact_both_layouts, P.test_score_special_values, P.test_score_late_nan_defeats_early_exit,
         P.test_accept_lists_match_reference_order, P.test_score_tile_pruning_wide_particles,
         P.test_score_sub_batches_on_two_streams, P.test_score_tile_pruning_signed_and_degenerate_data,
         P.test_score_near_matches_fill_the_queues, P.test_score_empty_and_small, P.test_score_linearity_property_large]


@pytest.mark.parametrize("fn", SUITE, ids=[f.__name__ for f in SUITE])
def test_scoring_suite_with_the_tensor_core_filter(eng_mma, data_stats, fn):
    fn(eng_mma, data_stats)


def test_filter_on_and_off_agree_bit_for_bit(eng_mma, data_stats):
    d, se = data_stats
    rng = np.random.default_rng(77)
    s = P.synth_stats(rng, d, 9000)
    s[100, 3] = np.nan
    res = []
    for on in (eng_mma.mma_mode, 0):
        eng_mma.set_option("score_mma_filter", on)
        try:
            eng_mma.accept_reset()
            err, counts, _ = eng_mma.score(s, eps=4.8, err_layout=ERR_PARTICLE_MAJOR)
            res.append((err, counts, eng_mma.accept_fetch()))
        finally:
            eng_mma.set_option("score_mma_filter", eng_mma.mma_mode)
    assert oracle.same_bits(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    for x, y in zip(res[0][2], res[1][2]):
        assert np.array_equal(np.asarray(x).view(np.uint64) if np.asarray(x).dtype == np.float64 else x,
                              np.asarray(y).view(np.uint64) if np.asarray(y).dtype == np.float64 else y)
