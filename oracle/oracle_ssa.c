/*
 * oracle_ssa.c -- CPU restatement (TEST ORACLE) of the SSA stage.  See oracle_ssa.h.
 *
 * The reference contains no SSA: its simulator integrates the moment equations of this
 * same CME (scripts/model.jl:74-86).  This file states the stochastic process those moments belong
 * to -- channels and rates from model.jl:1-27,58-64,74-86, binomial partitioning at division from
 * model.jl:98-111, binomial capture from model.jl:221-239 -- and simulates it with Gillespie's direct
 * method following DESIGN.md section 5 (Philox keying, per-cycle schedule, exact bitwise binomials).
 * Chain of trust: reference golden vectors -> orc_model (abc_oracle.c) -> moment z-tests of this
 * file -> draw-for-draw / KS comparison with the CUDA kernel.
 *
 * Compile with -ffp-contract=off: the DET math mode relies on unfused IEEE single precision ops.
 */
#include "oracle_ssa.h"
#include "oracle_philox.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- deterministic helpers */
double orc_exp10_det(double x) {
    const double HI = 3.321928094887362181708567732130177319049835205078125;
    const double LO = 1.66146163114990303432e-16;
    const double LN2 = 0.6931471805599453094172321214581765680755001343602552;
    static const double inv_fact[14] = {1.0, 1.0, 0.5, 1.0/6.0, 1.0/24.0, 1.0/120.0, 1.0/720.0, 1.0/5040.0,
        1.0/40320.0, 1.0/362880.0, 1.0/3628800.0, 1.0/39916800.0, 1.0/479001600.0, 1.0/6227020800.0};
    if (!(x > -300.0)) return (x != x) ? x : 0.0;
    if (x > 300.0) return INFINITY;
    double n = rint(x * HI);
    double r = fma(x, HI, -n);
    r = fma(x, LO, r);
    double z = r * LN2;
    double p = inv_fact[13];
    for (int k = 12; k >= 0; --k) p = fma(p, z, inv_fact[k]);
    return ldexp(p, (int)n);
}

static float log_det(float u) { /* ln(u), u in (0,1], IEEE single ops only */
    union { float f; uint32_t i; } v;
    v.f = u;
    uint32_t ix = v.i - 0x3f3504f3u;
    int e = (int32_t)ix >> 23;
    v.i = (ix & 0x007fffffu) + 0x3f3504f3u;
    float f = v.f + -1.0f;
    float s = f / (2.0f + f);
    float z = s * s;
    float p = fmaf(z, 1.0f / 9.0f, 1.0f / 7.0f);
    p = fmaf(z, p, 1.0f / 5.0f);
    p = fmaf(z, p, 1.0f / 3.0f);
    p = fmaf(z, p, 1.0f);
    float l1p = (s + s) * p;
    return fmaf((float)e, 0.693147182464599609375f, l1p);
}

/* ---------------------------------------------------------------- random stream of one lineage */
typedef struct {
    uint32_t ctr[4], key[2];
    uint32_t buf[4];
    int avail;
} stream_t;

static void stream_block(stream_t* s, uint32_t out[4]) {
    orc_philox4x32_10(s->ctr, s->key, out);
    s->ctr[0] += 1u;
}
static uint32_t stream_word(stream_t* s) {
    if (s->avail == 0) { stream_block(s, s->buf); s->avail = 4; }
    uint32_t w = s->buf[4 - s->avail];
    s->avail -= 1;
    return w;
}
static uint32_t popc(uint32_t x) { return (uint32_t)__builtin_popcount(x); }

/* Binomial(n, 1/2) = number of ones among n fresh bits (32 per word, low bits of the last word) */
static uint32_t bin_half(uint32_t n, stream_t* s) {
    uint32_t cnt = 0;
    while (n >= 32u) { cnt += popc(stream_word(s)); n -= 32u; }
    if (n > 0u) cnt += popc(stream_word(s) & ((1u << n) - 1u));
    return cnt;
}
/* Binomial(n, B/2^32): bitwise comparison of every molecule's uniform with B, MSB first */
static uint32_t bin_q32(uint32_t n, uint32_t B, stream_t* s) {
    uint32_t m = n, acc = 0;
    for (int bit = 31; bit >= 0 && m > 0u; --bit) {
        uint32_t h = bin_half(m, s);
        if ((B >> bit) & 1u) { acc += h; m -= h; } else m = h;
    }
    return acc;
}

/* ---------------------------------------------------------------- rates */
typedef struct { float kon[5], koff[5], alpha[5], gamma[5], lam; uint32_t pon_thr; } rates_t;

static void make_rates(const double* th, int m, rates_t* r) {
    int vary = (m == 3) ? 0 : (m == 4) ? 2 : (m == 5) ? 3 : -1; /* abc_simulation.jl:83 vary_flag */
    float* dst[4] = {r->kon, r->koff, r->alpha, r->gamma};
    int k = 0;
    for (int q = 0; q < 4; ++q) {
        if (q == vary) { for (int j = 0; j < 5; ++j) dst[q][j] = (float)orc_exp10_det(th[k + j]); k += 5; }
        else { float v = (float)orc_exp10_det(th[k]); for (int j = 0; j < 5; ++j) dst[q][j] = v; k += 1; }
    }
    double lam = orc_exp10_det(th[k]);
    if (!(lam <= 1.0)) lam = 1.0;
    if (!(lam >= 0.0)) lam = 0.0;
    r->lam = (float)lam;
    double kon = (double)r->kon[4], koff = (double)r->koff[4];
    double thr = kon / (kon + koff) * 4294967296.0;
    r->pon_thr = (thr >= 4294967295.0) ? 0xFFFFFFFFu : (thr > 0.0 ? (uint32_t)thr : 0u);
}

/* abc_simulation.jl:3-11 with the Philox keying of DESIGN.md 5.4 */
void orc_prior(int m, int64_t particle, uint64_t seed, double* theta) {
    int P = orc_n_params(m);
    int vary = (m == 3) ? 0 : (m == 4) ? 2 : (m == 5) ? 3 : -1;
    double lo[9], hi[9];
    int k = 0;
    for (int q = 0; q < 4; ++q) {
        int len = (q == vary) ? 5 : 1;
        for (int j = 0; j < len; ++j) { lo[k] = -3.0; hi[k] = (q == 3) ? 2.0 : 3.0; k++; }
    }
    lo[k] = -0.7; hi[k] = 0.0;
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int b = 0; b * 2 < P; ++b) {
        uint32_t ctr[4] = {(uint32_t)b, (uint32_t)(uint64_t)particle, (uint32_t)((uint64_t)particle >> 32),
                           ((uint32_t)(m - 1) << 26) | (1u << 29)};
        uint32_t w[4];
        orc_philox4x32_10(ctr, key, w);
        double u0 = (double)(((uint64_t)(w[0] >> 5) << 26) | (uint64_t)(w[1] >> 6)) * (1.0 / 9007199254740992.0);
        double u1 = (double)(((uint64_t)(w[2] >> 5) << 26) | (uint64_t)(w[3] >> 6)) * (1.0 / 9007199254740992.0);
        theta[2*b] = fma(hi[2*b] - lo[2*b], u0, lo[2*b]);
        if (2*b + 1 < P) theta[2*b+1] = fma(hi[2*b+1] - lo[2*b+1], u1, lo[2*b+1]);
    }
}

void orc_quantise_betas(const double* betas, const int* clusters, int n, uint32_t* q32, int* off5) {
    int k = 0;
    for (int c = 1; c <= ORC_NAGE; ++c) {
        off5[c - 1] = k;
        for (int i = 0; i < n; ++i) if (clusters[i] == c) {
            double q = floor(betas[i] * 4294967296.0 + 0.5);
            q32[k++] = (q >= 4294967295.0) ? 0xFFFFFFFFu : (uint32_t)q;
        }
    }
    off5[ORC_NAGE] = k;
}

/* ---------------------------------------------------------------- one lineage */
typedef struct { float len, kon, koff, A0, A1, gam, lamf; } seg_t;

static int build_cycle(const rates_t* r, const orc_ssa_design_t* d, int scaling, int cond, int age_i, int c, seg_t* out) {
    const double cycle = d->cycle, age = d->agevec[age_i];
    const double tl0 = age - d->pulse[cond] - d->chase[cond], tl1 = age - d->chase[cond];
    const double Tc = (double)(c - d->n_pre) * cycle;
    const double cyc_end = (c == d->n_pre) ? age : cycle;
    const double l0 = tl0 - Tc, l1 = tl1 - Tc, step_len = cycle / 5.0, sc = scaling ? 1.0 : 0.0;
    double pos = 0.0;
    int k = 0, n = 0;
    while (pos < cyc_end && n < 7) {
        double step_end = (double)(k + 1) * step_len;
        double nxt = step_end < cyc_end ? step_end : cyc_end;
        if (l0 > pos && l0 < nxt) nxt = l0;
        if (l1 > pos && l1 < nxt) nxt = l1;
        double mid = 0.5 * (pos + nxt);
        int lab = (mid >= l0) && (mid <= l1);
        seg_t* sg = &out[n++];
        sg->len = (float)(nxt - pos);
        sg->kon = r->kon[k]; sg->koff = r->koff[k]; sg->gam = r->gamma[k];
        sg->A0 = (float)((double)r->alpha[k] * (1.0 + sc * pos / cycle));
        sg->A1 = (float)((double)r->alpha[k] * sc / cycle);
        sg->lamf = lab ? r->lam : 0.0f;
        pos = nxt;
        if (!(pos < step_end)) k += 1;
        if (k > 4) k = 4;
    }
    return n;
}

typedef struct { uint32_t U, L; int g; uint64_t events; } cell_t;

/* one event draw in sub-interval sg at position *x; returns 1 when the boundary is crossed.
 * Channel layout on [0, tot): switch at the bottom, death at the top (U first), birth in between
 * (DESIGN.md 5.3). */
static int ssa_step(cell_t* s, float* x, const seg_t* sg, uint32_t wt, uint32_t wc, int math_mode) {
    int on = s->g != 0;
    float asw = on ? sg->koff : sg->kon;
    float n = (float)s->U + (float)s->L;
    float ad = sg->gam * n;
    float ab = on ? fmaf(sg->A1, *x, sg->A0) : 0.0f;
    float c1 = on ? sg->A1 : 0.0f;
    float base = asw + ad;
    float c0 = base + ab;
    float u = fmaf((float)wt, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    float E = (math_mode == ORC_MATH_DET) ? -log_det(u) : -logf(u);
    float disc = fmaf(c1 + c1, E, c0 * c0);
    float tau = (E + E) / (c0 + sqrtf(disc));
    float xn = *x + tau;
    if (!(xn < sg->len)) return 1;
    *x = xn;
    float abn = on ? fmaf(sg->A1, xn, sg->A0) : 0.0f;
    float tot = base + abn;
    float t32 = tot * 2.3283064365386963e-10f;
    float rs = (float)wc * t32;
    float rb = (float)(~wc) * t32;
    /* (float)wc rounds the top 128 words up to 2^32 (rs == tot): when the switch is the only channel it must still fire */
    int sw = (rs < asw) || !(asw < tot);
    int death = !sw && (rb < ad);
    int dU = rb < sg->gam * (float)s->U;
    int birth = !sw && !death;
    int lab = birth && ((rs + -asw) < sg->lamf * abn);
    if (sw) s->g ^= 1;
    if (birth && !lab) s->U += 1u;
    if (lab) s->L += 1u;
    if (death && dU) s->U -= 1u;
    if (death && !dU) s->L -= 1u;
    s->events += 1;
    return 0;
}

static void simulate_cell(const rates_t* r, const orc_ssa_design_t* d, int m, int64_t particle, uint64_t seed,
                          int cond, int age_i, int cell, int math_mode, uint32_t out[4], uint64_t* events) {
    stream_t st;
    uint64_t gp = (uint64_t)particle;
    int readout = cond * ORC_NAGE + age_i;
    st.ctr[0] = 0; st.ctr[1] = (uint32_t)gp; st.ctr[2] = (uint32_t)(gp >> 32);
    st.ctr[3] = ((uint32_t)cell & 0xFFFFFu) | ((uint32_t)readout << 20) | ((uint32_t)(m - 1) << 26);
    st.key[0] = (uint32_t)seed; st.key[1] = (uint32_t)(seed >> 32);
    st.avail = 0;
    cell_t s = {0u, 0u, 0, 0};
    uint32_t w[4];
    stream_block(&st, w);
    s.g = (w[0] < r->pon_thr) ? 1 : 0;
    for (int c = 0; c <= d->n_pre; ++c) {
        seg_t segs[8];
        int n_ent = build_cycle(r, d, m != 2, cond, age_i, c, segs);
        int e = 0;
        float x = 0.0f;
        while (e < n_ent) {
            stream_block(&st, w);
            if (ssa_step(&s, &x, &segs[e], w[0], w[1], math_mode)) { e += 1; x = 0.0f; }
            if (e < n_ent) {
                if (ssa_step(&s, &x, &segs[e], w[2], w[3], math_mode)) { e += 1; x = 0.0f; }
            }
        }
        if (c < d->n_pre) {
            st.avail = 0;
            s.U = bin_half(s.U, &st);
            s.L = bin_half(s.L, &st);
        }
    }
    out[0] = s.U; out[1] = s.L; out[2] = s.U; out[3] = s.L;
    if (d->downsampling) {
        st.avail = 0;
        int grp = (cond < 6 ? 0 : ORC_NAGE) + age_i;
        uint32_t off = (uint32_t)d->beta_off[grp], cnt = (uint32_t)d->beta_off[grp + 1] - off;
        uint32_t pick = (uint32_t)(((uint64_t)stream_word(&st) * cnt) >> 32);
        uint32_t B = d->beta_q32[off + pick];
        out[2] = bin_q32(s.U, B, &st);
        out[3] = bin_q32(s.L, B, &st);
    }
    *events += s.events;
}

void orc_ssa_readout(const double* theta, int m, const orc_ssa_design_t* d, int64_t particle, uint64_t seed,
                     int cond, int age, int math_mode, uint32_t* counts, uint64_t* n_events) {
    rates_t r;
    make_rates(theta, m, &r);
    uint64_t ev = 0;
    for (int cell = 0; cell < d->n_cells; ++cell) {
        uint32_t o[4];
        uint64_t e1 = 0;
        simulate_cell(&r, d, m, particle, seed, cond, age, cell, math_mode, o, &e1);
        for (int q = 0; q < 4; ++q) counts[(size_t)q * d->n_cells + cell] = o[q];
        ev += e1;
    }
    if (n_events) *n_events = ev;
}

static double u128_to_double(unsigned __int128 v) {
    uint64_t hi = (uint64_t)(v >> 64), lo = (uint64_t)v;
    return (double)hi * 18446744073709551616.0 + (double)lo;
}

/* corrected sample moments (var(), cov() of data_summary_statistics.jl:117-121) from exact integer sums */
void orc_moments_from_sums(const uint64_t s[5], int n_cells, double o[5]) {
    unsigned __int128 N = (unsigned __int128)(uint64_t)n_cells, a, b;
    double dn = (double)n_cells, dnn = dn * (double)(n_cells - 1);
    o[0] = (double)s[0] / dn;
    o[1] = (double)s[1] / dn;
    a = N * s[2]; b = (unsigned __int128)s[0] * s[0];
    o[2] = u128_to_double(a - b) / dnn;
    a = N * s[3]; b = (unsigned __int128)s[0] * s[1];
    o[3] = (a >= b) ? u128_to_double(a - b) / dnn : -(u128_to_double(b - a) / dnn);
    a = N * s[4]; b = (unsigned __int128)s[1] * s[1];
    o[4] = u128_to_double(a - b) / dnn;
}

void orc_ssa_moments(const double* theta, int m, const orc_ssa_design_t* d, int64_t particle, uint64_t seed,
                     int math_mode, double* moments, uint64_t* n_events) {
    rates_t r;
    make_rates(theta, m, &r);
    uint64_t ev = 0;
    for (int ro = 0; ro < ORC_NCOND * ORC_NAGE; ++ro) {
        uint64_t sums[5] = {0, 0, 0, 0, 0}, e1 = 0;
        for (int cell = 0; cell < d->n_cells; ++cell) {
            uint32_t o[4];
            simulate_cell(&r, d, m, particle, seed, ro / ORC_NAGE, ro % ORC_NAGE, cell, math_mode, o, &e1);
            uint64_t u = o[2], l = o[3];
            sums[0] += u; sums[1] += l; sums[2] += u * u; sums[3] += u * l; sums[4] += l * l;
        }
        orc_moments_from_sums(sums, d->n_cells, &moments[ro * 5]);
        ev += e1;
    }
    if (n_events) *n_events = ev;
}
