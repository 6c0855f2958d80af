"""The algebra behind the telegraph phase of the SSA (csrc/abc_ssa.cu make_tseg / tseg_F), restated in
numpy with binary32 rounding and checked against quadrature: the Poisson mean of the transcripts alive at the end of a
sub-interval is  Lam_start exp(-gam len) + sum over "on" stretches [x1, x2] of F(x2) - F(x1),  F' = alpha(w) exp(-gam (len - w)),
with the closed form above gam * step >= 1/4 and the six-term series below.  Also the adaptive burn-in rule (burnin_cycles)."""
import math

import numpy as np
import pytest
from scipy.integrate import quad

f32 = np.float32


def make_tseg(len_, A0, A1, gam, step_len):
    dec = math.exp(-gam * len_)
    if gam * step_len < 0.25:
        c, flen = [], 0.0
        for j in range(1, 10):
            v = dec * (A0 * gam ** (j - 1) / math.factorial(j) + (A1 * gam ** (j - 2) / (math.factorial(j - 2) * j) if j >= 2 else 0.0))
            if j <= (4 if gam * step_len < 0.05 else 6):       # short series below gam * step = 1/20 (csrc/abc_tele.cu make_piece)
                c.append(f32(v))
            flen += v * len_ ** j
        return dict(small=True, c=c, Flen=f32(flen), F0=f32(0.0), dec=f32(dec), len=f32(len_))
    e = A1 / gam
    cc = (A0 - e) / gam
    return dict(small=False, k1=f32(gam * 1.4426950408889634), c=f32(cc), e=f32(e), Flen=f32(cc + e * len_), F0=f32(dec * cc),
                dec=f32(dec), len=f32(len_))


def tseg_F(k, x):
    x = f32(x)
    if k["small"]:
        c = k["c"]
        h = c[-1]
        for j in range(len(c) - 2, -1, -1):
            h = f32(f32(h * x) + c[j])
        return f32(h * x)
    d = f32(2.0 ** float(f32(k["k1"] * f32(x - k["len"]))))
    return f32(d * f32(f32(k["e"] * x) + k["c"]))


def exact(A0, A1, gam, len_, a, b):
    return quad(lambda w: (A0 + A1 * w) * math.exp(-gam * (len_ - w)), a, b, epsabs=0, epsrel=1e-12)[0]


@pytest.mark.parametrize("seed", range(4))
def test_accumulated_F_equals_the_integral_over_on_stretches(seed):
    rng = np.random.default_rng(seed)
    worst = 0.0
    for _ in range(150):
        gam = 10 ** rng.uniform(-3, 2)
        A = 10 ** rng.uniform(-3, 3)
        kon, koff = 10 ** rng.uniform(-2, 2.5), 10 ** rng.uniform(-2, 2.5)
        pos = float(rng.choice([0, 4, 8, 12, 16]))
        step = float(rng.choice([4.0, 20.0]))
        len_ = float(rng.choice([step, rng.uniform(0.05, step)]))
        A0, A1 = A * (1 + pos / 20), (A / 20 if rng.random() < 0.8 else 0.0)
        k = make_tseg(len_, A0, A1, gam, step)
        g = int(rng.random() < kon / (kon + koff))
        x, want, n = 0.0, 0.0, 0
        acc = f32(-g * k["F0"])
        while n < 400:
            xn = x + rng.exponential(1.0 / (koff if g else kon))
            if xn >= len_:
                break
            if g:
                want += exact(A0, A1, gam, len_, x, xn)
            fx = tseg_F(k, xn)
            acc = f32(acc + (fx if g else -fx))
            g ^= 1
            x = xn
            n += 1
        if n >= 400:
            continue
        if g:
            want += exact(A0, A1, gam, len_, x, len_)
        got = float(f32(acc + (k["Flen"] if g else f32(0.0))))
        scale = exact(A0, A1, gam, len_, 0.0, len_)          # the sub-interval's contribution with the gene always on
        worst = max(worst, abs(got - want) / scale)
    # binary32 rounding of up to 400 terms of size |F| <= ~8 x the segment's own scale (random walk): a few 1e-6; bound 5e-5
    assert worst < 5e-5, worst


def test_series_and_closed_form_agree_at_the_split():
    for gam in (0.0624, 0.0626):
        k = make_tseg(4.0, 3.0, 0.15, gam, 4.0)
        for x in (0.3, 1.7, 3.9):
            got = float(tseg_F(k, x)) - float(k["F0"])
            assert abs(got - exact(3.0, 0.15, gam, 4.0, 0.0, x)) < 2e-5 * exact(3.0, 0.15, gam, 4.0, 0.0, 4.0)
        assert abs(float(k["Flen"]) - float(k["F0"]) - exact(3.0, 0.15, gam, 4.0, 0.0, 4.0)) < 1e-5 * exact(3.0, 0.15, gam, 4.0, 0.0, 4.0)


def burnin_cycles(gam5, kon5, koff5, m, n_pre, cycle, tl0):
    """csrc/abc_ssa.cu burnin_cycles: smallest k with k (1 + log2(e) sum gam_s cycle/5) >= n_pre (model 3: gene memory too)"""
    step = cycle / 5.0
    bits = 1.0 + 1.4426950 * sum(gam5) * step
    need = float(n_pre)
    if m == 3:
        bits = min(bits, 1.4426950 * sum(a + b for a, b in zip(kon5, koff5)) * step)
        need += 6.0
    k = n_pre
    if bits * n_pre >= need:
        k = math.ceil(need / bits)
    k_win = math.ceil(-tl0 / cycle) if tl0 < 0 else 0          # cycles back to the one in which the label window opens
    return min(max(k + k_win, 1), n_pre)


def test_adaptive_burnin_keeps_the_bias_bound_of_the_full_burnin():
    rng = np.random.default_rng(7)
    for _ in range(2000):
        m = int(rng.integers(1, 6))
        gam5 = [10 ** rng.uniform(-3, 2)] * 5 if m != 5 else list(10 ** rng.uniform(-3, 2, 5))
        kon5 = [10 ** rng.uniform(-3, 3)] * 5 if m != 3 else list(10 ** rng.uniform(-3, 3, 5))
        koff5 = [10 ** rng.uniform(-3, 3)] * 5
        tl0 = float(rng.choice([1.0, -1.0, -20.0, -26.0]))
        k = burnin_cycles(gam5, kon5, koff5, m, 10, 20.0, tl0)
        assert 1 <= k <= 10 and k * 20.0 >= -tl0
        if k < 10:
            # complete cycles simulated before the cycle in which the label window opens
            kb = k - (math.ceil(-tl0 / 20.0) if tl0 < 0 else 0)
            assert kb >= 1
            # share of the pre-window Lam_U born before the simulated range: (1/2 exp(-sum gam_s 4))^kb <= 2^-10
            assert (0.5 * math.exp(-sum(gam5) * 4.0)) ** kb <= 2.0 ** -10 * (1 + 1e-6)
            if m == 3:      # the approximate start law of the gene has decayed as well
                assert math.exp(-sum(a + b for a, b in zip(kon5, koff5)) * 4.0 * kb) <= 2.0 ** -10
    assert burnin_cycles([1.0] * 5, [1.0] * 5, [1.0] * 5, 1, 10, 20.0, 1.0) == 1
    assert burnin_cycles([1e-3] * 5, [1.0] * 5, [1.0] * 5, 1, 10, 20.0, 1.0) == 10
