"""CPU check of the soundness argument of the tensor-core filter (csrc/abc_score3.cu, "tensor-core filter").

The filter evaluates  sum_t w_t (d_t - s_t)^2  in its expanded form with every operand rounded to TF32 and sends a pair to the
FP64 stage iff the value is below a per-gene threshold 10 + slack_g.  Here the same quantities are restated in numpy on the
shipped data: for particles whose true error is <= 10 the TF32 value may exceed it by at most the part of slack_g that covers
the operand rounding, with either rounding mode the device and the host use; particles far above 10 are not needed.
"""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def tf32(x, ties_away):
    f = np.asarray(x, dtype=np.float32).copy()
    u = f.view(np.uint32)
    if ties_away:
        r = (u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    else:
        r = (u + np.uint32(0x0FFF) + ((u >> np.uint32(13)) & np.uint32(1))) & np.uint32(0xFFFFE000)
    fin = (u & np.uint32(0x7F800000)) != np.uint32(0x7F800000)
    u[fin] = r[fin]
    return f.astype(np.float64)


def test_tf32_value_stays_within_the_slack_for_pairs_below_ten():
    z = np.load(os.path.join(GOLD, "ref_summary_stats.npz"))
    d, se = z["d"], z["se"]
    G = d.shape[0]
    eps = np.where(se + d != 0.0, 0.0, 1e-4)
    den = (se * se + 0.01 * (d * d)) + eps
    w = 1.0 / (53.0 * den)
    R = (w * d * d).sum(1)
    assert R.max() <= 100.0 * (1 + 1e-12)                 # den >= 0.01 d^2: what bounds the slack
    Q = (np.sqrt(10.0) + np.sqrt(R)) ** 2
    X = 2.0 * np.sqrt(R * Q)
    round_slack = 1.002 * (X + Q) * 2.0 ** -10            # operand rounding; the accumulation term 2^-14 (R + X + Q) is on top
    assert 0.1 < np.median(round_slack) < 0.3 and round_slack.max() < 0.45
    p, q = -2.0 * w * d, w
    rng = np.random.default_rng(11)
    worst = 0.0
    n_below = 0
    for rep in range(6):
        # particles near the data of random genes, whole-particle rescalings, a few wild ones; signs flipped now and then
        n = 400
        g = rng.integers(0, G, n)
        s = d[g] * np.exp(rng.normal(0.0, [0.05, 0.15, 0.3][rep % 3], (n, 53)))
        s *= 10.0 ** rng.choice([0.0, 0.0, 0.0, 0.3, -0.3], size=(n, 1))
        s[rng.random((n, 53)) < 0.01] *= -1.0
        E = np.zeros((n, G))
        for t in range(53):
            E += w[None, :, t] * (d[None, :, t] - s[:, None, t]) ** 2
        below = E <= 10.0
        n_below += int(below.sum())
        for ties_away in (False, True):                   # cvt.rna on the device, round-to-nearest-even on the host
            s1 = tf32(s.astype(np.float32), ties_away)
            s2 = tf32((s * s).astype(np.float32), ties_away)
            V = R[None, :] + s1 @ tf32(p, False).T + s2 @ tf32(q, False).T
            excess = (V - E)[below]
            assert (excess <= round_slack[np.nonzero(below)[1]]).all()
            worst = max(worst, float(excess.max()))
        # the pairs the filter lets through are a small multiple of the pairs that are needed
        live = V <= (10.0 + round_slack + 2.0 ** -14 * (R + X + Q) + 0.002)[None, :]
        assert not (below & ~live).any()
        assert live.sum() <= 1.25 * below.sum() + 10
    assert n_below > 20000 and worst > 0.0                # the sample exercises the bound
    assert worst < 0.2                                    # and stays well inside it on real data
