// abc_tele.cu -- the product sampler of the SSA stage (ssa_hybrid_burnin = 2): Gillespie's direct
// method on the gene switch, the Poisson means of the unlabelled / labelled transcripts carried in closed form along the
// gene path, U ~ Poisson(Lam_U), L ~ Poisson(Lam_L) at the read-out, then binomial capture-efficiency thinning.  sm_100a.
//
// The CME is the one whose moments scripts/model.jl:74-86 (f), :98-111 (periodic_boundary), :221-239 (downsample)
// describe; rate schedule scripts/model.jl:1-27, label window :58-64, :183.
//
// Two kernels:
//   abc_window_kernel  one thread per (particle, read-out): where the lineages of that read-out start.  Transcripts born
//                      before the start s0 are not simulated; s0 is the latest time for which their expected share of
//                      Lam_U and of Lam_L at the read-out is below 2^-n_pre_cycles (the bias bound of the configured
//                      burn-in), evaluated from the exact mean contributions of every piece of the schedule
//                      (ssa_adaptive_burnin = 2; = 1: whole cycles by the worst-case rule of abc_ssa.cu; = 0: always
//                      n_pre_cycles cycles).  Also the expected number of draws per particle (scheduling hint).
//   abc_tele_kernel    one thread per cell lineage, one warp per 32 cells of one (particle, read-out), persistent CTAs
//                      pulling items from an atomic queue in longest-processing-time order.  Per iteration one
//                      Philox4x32-10 block = four switch draws; a block in which no lane of the warp reaches the end of its
//                      sub-interval takes the branch-free path; otherwise the same expression is evaluated with the
//                      draws beyond the boundary masked, and only the lanes that cross run the boundary code (the
//                      rest of their block is discarded: the waiting times are memoryless and the words independent).
#include "abc_ssa_dev.cuh"
#include <type_traits>

#ifndef TELE_MIN_CTAS
#define TELE_MIN_CTAS 3      // 24 warps x 80 registers per SM: measured best (4 x 64: -4 %, 2 x 100: -8 %; computing the next
                             // Philox block one iteration ahead for more ILP: -8 %, the loop is bound by issue slots, not latency)
#endif
#define TELE_WARPS 8
#define TELE_MAX_SEG 80      // 5 rate steps x (n_pre_cycles + 1 <= 14 cycles) + label on/off + the cut at the start + slack
#define FULL 0xffffffffu

// one piece of the schedule between two consecutive cuts (rate steps, label on/off, read-out), 64 B.
//   closed form (k1 != 0):  F(x) = 2^(k1 x + e0) (p0 + p1 x),  e0 = -k1 len, p0 = A0/gam - A1/gam^2, p1 = A1/gam, p2 = F(0)
//   series (k1 == 0, gam step < 1/4):  F(x) = x (p0 + p1 x + ... + p5 x^5); e0 = 1 (gam step < 1/20): F(x) = x (p0 + ... + p3 x^3)
// F is the antiderivative of alpha(w) exp(-gam (len - w)): the Poisson mean advances by F(x2) - F(x1) over an "on"
// stretch [x1, x2] and decays by dec = exp(-gam len) over the piece.
struct __align__(16) TPiece {
    float len, qon, qoff, dec;      // q = -ln2 / rate: waiting time = lg2(u) * q
    float e0, k1, p0, p1;
    float p2, p3, p4, p5;
    float Flen, lamf;               // F(len); labelled share of the births
    int meta;                       // bit 0: a cell division follows, bit 1: last piece
    float pad;
};
#define TP_DIV 1
#define TP_LAST 2

// ------------------------------------------------------------------------------------------------
// The cuts of one read-out between a start time s and the read-out t_end (times in hours, 0 = start of the read-out
// cycle): every multiple of `step` and the label window's ends tl0 < tl1 where they fall strictly inside and off a step.
struct Cuts {
    double s, t_end, step, inv_step, tl0, tl1;
    int j_first;            // first multiple of step after s
    int nb, ins0, ins1, r0, r1, n;   // multiples of step inside, window ends inserted (and their ranks), number of pieces
};

// floor(t / step) with the division replaced by a multiplication and two exact fix-ups
__device__ __forceinline__ int floor_div(double t, double step, double inv_step) {
    int j = (int)floor(t * inv_step);
    if ((double)j * step > t) j -= 1;
    if ((double)(j + 1) * step <= t) j += 1;
    return j;
}
__device__ __forceinline__ int ceil_div(double t, double step, double inv_step) {
    const int j = floor_div(t, step, inv_step);
    return ((double)j * step == t) ? j : j + 1;
}
__device__ __forceinline__ bool on_grid(double t, double step, double inv_step) {
    return (double)floor_div(t, step, inv_step) * step == t;
}

__device__ __forceinline__ Cuts make_cuts(double s, double t_end, double step, double inv_step, double tl0, double tl1,
                                          bool window) {
    Cuts c;
    c.s = s; c.t_end = t_end; c.step = step; c.inv_step = inv_step; c.tl0 = tl0; c.tl1 = tl1;
    c.j_first = floor_div(s, step, inv_step) + 1;
    const int j_end = ceil_div(t_end, step, inv_step);          // multiples below t_end: j_first .. j_end - 1
    c.nb = (j_end > c.j_first) ? (j_end - c.j_first) : 0;
    c.ins0 = (window && tl0 > s && tl0 < t_end && !on_grid(tl0, step, inv_step)) ? 1 : 0;
    c.ins1 = (window && tl1 > s && tl1 < t_end && !on_grid(tl1, step, inv_step)) ? 1 : 0;
    const int q0 = ceil_div(tl0, step, inv_step) - c.j_first, q1 = ceil_div(tl1, step, inv_step) - c.j_first;
    c.r0 = q0 < 0 ? 0 : (q0 > c.nb ? c.nb : q0);                // multiples of step in (s, tl0)
    c.r1 = (q1 < 0 ? 0 : (q1 > c.nb ? c.nb : q1)) + c.ins0;
    c.n = (t_end > s) ? c.nb + c.ins0 + c.ins1 + 1 : 0;
    return c;
}

#define NO_GRID (-0x40000000)
// i-th interior cut, i in [0, n-1); *grid_index = multiple of step it sits on (NO_GRID for a window end)
__device__ __forceinline__ double interior_cut(const Cuts& c, int i, int* grid_index) {
    if (c.ins0 && i == c.r0) { *grid_index = NO_GRID; return c.tl0; }
    if (c.ins1 && i == c.r1) { *grid_index = NO_GRID; return c.tl1; }
    const int j = i - ((c.ins0 && i > c.r0) ? 1 : 0) - ((c.ins1 && i > c.r1) ? 1 : 0);
    *grid_index = c.j_first + j;
    return (double)(c.j_first + j) * c.step;
}

struct PieceDesc {
    double a, b;            // [a, b)
    int step;               // rate step 0..4 of scripts/model.jl:1-22
    double xa;              // position of a inside its cycle
    bool labelled, div_after, last;
};

// inv_cycle, inv_step5: 1 / cycle, 5 / cycle
__device__ __forceinline__ PieceDesc describe_piece(const Cuts& c, int k, double cycle, double inv_cycle, double inv_step5,
                                                    int steps_per_cycle, bool window) {
    PieceDesc d;
    int gi = 0, gdummy = 0;
    d.a = (k == 0) ? c.s : interior_cut(c, k - 1, &gdummy);
    d.last = (k == c.n - 1);
    d.b = d.last ? c.t_end : interior_cut(c, k, &gi);
    const double mid = 0.5 * (d.a + d.b);
    const double cyc0 = cycle * (double)floor_div(mid, cycle, inv_cycle);
    int st = (int)((mid - cyc0) * inv_step5);        // mid lies strictly inside a rate step
    d.step = st < 0 ? 0 : (st > 4 ? 4 : st);
    d.xa = d.a - cyc0;
    d.labelled = window && mid >= c.tl0 && mid <= c.tl1;
    // a division follows when the piece ends on a multiple of the cycle (never the read-out itself: 0 < age < cycle)
    d.div_after = !d.last && gi != NO_GRID && (steps_per_cycle == 1 || (gi % 5) == 0);
    return d;
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tp_F(float k1, float e0, float p0, float p1, float p2, float p3, float p4, float p5, float x) {
    if (k1 != 0.0f) {
        float D;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(D) : "f"(f_fma(k1, x, e0)));
        return f_mul(D, f_fma(p1, x, p0));
    }
    float h = f_fma(p5, x, p4);
    h = f_fma(h, x, p3);
    h = f_fma(h, x, p2);
    h = f_fma(h, x, p1);
    h = f_fma(h, x, p0);
    return f_mul(h, x);
}

// sc_inv_cycle = scaling / cycle (alpha(x) = alpha_step (1 + scaling x / cycle), scripts/model.jl:1-27)
__device__ __forceinline__ TPiece make_piece(const AbcRates& r, const PieceDesc& d, float sc_inv_cycle, float step_len) {
    const float len = (float)(d.b - d.a);
    const float gam = r.gamma[d.step];
    const float A1 = f_mul(r.alpha[d.step], sc_inv_cycle);
    const float A0 = f_fma(A1, (float)d.xa, r.alpha[d.step]);
    TPiece t;
    t.len = len;
    t.qon = __fdiv_rn(-0.693147182464599609375f, r.kon[d.step]);
    t.qoff = __fdiv_rn(-0.693147182464599609375f, r.koff[d.step]);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t.dec) : "f"(f_mul(f_mul(gam, len), -1.4426950408889634f)));
    t.lamf = d.labelled ? r.lam : 0.0f;
    t.meta = (d.div_after ? TP_DIV : 0) | (d.last ? TP_LAST : 0);
    t.pad = 0.0f;
    if (f_mul(gam, step_len) < 0.25f) {
        // F(x) = dec * sum_j x^j [A0 gam^(j-1)/j! + A1 gam^(j-2)/((j-2)! j)]: six terms in the kernel (the next one is below
        // 5e-8 relative), nine for the boundary value F(len)
        const float rj[9] = {1.0f, 1.0f / 2, 1.0f / 6, 1.0f / 24, 1.0f / 120, 1.0f / 720, 1.0f / 5040, 1.0f / 40320, 1.0f / 362880};
        const float sj[9] = {0.0f, 1.0f / 2, 1.0f / 3, 1.0f / 8, 1.0f / 30, 1.0f / 144, 1.0f / 840, 1.0f / 5760, 1.0f / 45360};
        float cj[9], gp = t.dec, gq = 0.0f;       // gp = dec gam^(j-1), gq = dec gam^(j-2)
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            cj[j] = f_fma(f_mul(A0, gp), rj[j], f_mul(f_mul(A1, gq), sj[j]));
            gq = gp;
            gp = f_mul(gp, gam);
        }
        float h = cj[8];
#pragma unroll
        for (int j = 7; j >= 0; --j) h = f_fma(h, len, cj[j]);
        t.Flen = f_mul(h, len);
        // gam step < 1/20: four terms reach the same 5e-8 (the fifth is (gam x)^4 / 120 of the first); flagged in e0, which the
        // series does not use
        t.k1 = 0.0f; t.e0 = (f_mul(gam, step_len) < 0.05f) ? 1.0f : 0.0f;
        t.p0 = cj[0]; t.p1 = cj[1]; t.p2 = cj[2]; t.p3 = cj[3]; t.p4 = cj[4]; t.p5 = cj[5];
    } else {
        const float e = __fdiv_rn(A1, gam), c = __fdiv_rn(f_add(A0, -e), gam);
        t.k1 = f_mul(gam, 1.4426950408889634f);
        t.e0 = -f_mul(t.k1, len);
        t.p0 = c; t.p1 = e;
        t.p3 = 0.0f; t.p4 = 0.0f; t.p5 = 0.0f;
        // both boundary values through the kernel's own expression: a rounding of e0 then scales the whole piece
        // consistently instead of separating the boundary terms from the switch terms
        t.p2 = tp_F(t.k1, t.e0, c, e, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f);
        t.Flen = tp_F(t.k1, t.e0, c, e, 0.0f, 0.0f, 0.0f, 0.0f, len);
    }
    return t;
}

// models 1, 2: no rate varies and alpha is linear over the whole cycle: one piece per cycle (measured: cutting them into 4 h
// pieces to use the short series costs more in boundary crossings than the two FFMAs per draw it saves)
__device__ __forceinline__ int tele_steps_per_cycle(int m, float gam0, float cycle) {
    (void)gam0; (void)cycle;
    return (m <= 2) ? 1 : 5;
}

// ------------------------------------------------------------------------------------------------
// whole-cycle rule of the first adaptive burn-in (ssa_adaptive_burnin = 1; the same arithmetic as burnin_cycles in abc_ssa.cu)
__device__ __forceinline__ int tele_burnin_cycles(const AbcRates& r, const AbcSsaParams& prm, int cond, int age_i) {
    const float step = (float)(prm.cycle / 5.0);
    float zg = 0.0f, zr = 0.0f;
#pragma unroll
    for (int j = 0; j < 5; ++j) { zg += r.gamma[j]; zr += r.kon[j] + r.koff[j]; }
    float bits = 1.0f + 1.4426950f * zg * step;
    float need = (float)prm.n_pre;
    if (prm.m == 3) { bits = fminf(bits, 1.4426950f * zr * step); need += 6.0f; }
    int k = prm.n_pre;
    if (bits * (float)prm.n_pre >= need) k = (int)ceilf(need / bits);
    const double tl0 = prm.agevec[age_i] - prm.pulse[cond] - prm.chase[cond];
    const int k_win = (tl0 < 0.0) ? (int)ceil(-tl0 / prm.cycle) : 0;
    k = max(k + k_win, 1);
    return min(k, prm.n_pre);
}

// integral over [0, L] of (A0 + A1 u) exp(-gam (L - u)) du, FP64, stable for every gam L > 0
__device__ __forceinline__ double piece_integral(double A0, double A1, double gam, double L) {
    const double E1 = -expm1(-gam * L) / gam;                 // int_0^L exp(-gam (L - u)) du
    const double z = gam * L;
    // int_0^L u exp(-gam (L - u)) du = (L - E1) / gam, by its series when gam L is small (cancellation)
    const double E2 = (z < 1e-3) ? L * L * (0.5 - z * (1.0 / 6.0 - z * (1.0 / 24.0))) : (L - E1) / gam;
    return A0 * E1 + A1 * E2;
}

#define WIN_MAX_PIECES 96

// Where the lineages of one (particle, read-out) start.  out_start[p*55 + r] (hours, <= 0 means before the read-out cycle),
// cost[p] = expected switch draws of one lineage summed over the 55 read-outs (deterministic: one warp per particle).
__global__ void __launch_bounds__(128)
abc_window_kernel(AbcRates* __restrict__ rates, const AbcSsaParams prm, float* __restrict__ out_start, long long n) {
    const int lane = threadIdx.x & 31;
    const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= n) return;
    const AbcRates r = rates[p];
    const double cycle = prm.cycle, step5 = cycle / 5.0, inv_cycle = 1.0 / cycle, inv_step5 = 5.0 / cycle;
    const double t_min = -(double)prm.n_pre * cycle;
    const double eps = exp2(-(double)prm.n_pre), floor_abs = 9.313225746154785e-10;   // 2^-30 molecules
    double cost = 0.0, rate_max = 0.0;      // rate_max >= alpha(t) P_on(t) at any time
#pragma unroll
    for (int j = 0; j < 5; ++j)
        rate_max = fmax(rate_max, 2.0 * (double)r.alpha[j] * (double)r.kon[j] / ((double)r.kon[j] + (double)r.koff[j]));
    // a whole rate step (4 h) contributes the same in every cycle: P_on x integral of alpha(w) exp(-gam (end - w)), and exp(-gam step)
    double step_int[5], step_dec[5], pon_step[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const double gam = (double)r.gamma[j], al = (double)r.alpha[j];
        const double A0 = al * (1.0 + (prm.scaling ? (double)j * 0.2 : 0.0)), A1 = prm.scaling ? al / cycle : 0.0;
        pon_step[j] = (double)r.kon[j] / ((double)r.kon[j] + (double)r.koff[j]);
        step_int[j] = pon_step[j] * piece_integral(A0, A1, gam, step5);
        step_dec[j] = exp(-gam * step5);
    }
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float q[4] = {r.kon[j], r.koff[j], r.alpha[j], r.gamma[j]};
#pragma unroll
        for (int i = 0; i < 4; ++i) bad = bad || !(q[i] > 0.0f && q[i] < 1e30f);
    }
    for (int readout = lane; readout < ABC_NREAD; readout += 32) {
        const int cond = readout / ABC_NAGE, age_i = readout % ABC_NAGE;
        const double age = prm.agevec[age_i];
        const double tl0 = age - prm.pulse[cond] - prm.chase[cond], tl1 = age - prm.chase[cond];
        const bool window = prm.pulse[cond] > 0.0;
        double s0 = t_min;
        if (prm.adaptive == 1) {
            s0 = -(double)tele_burnin_cycles(r, prm, cond, age_i) * cycle;
        } else if (prm.adaptive >= 2) {
            // mean contribution of the births of every piece to Lam_U and Lam_L at the read-out, newest piece first
            const Cuts c = make_cuts(t_min, age, step5, inv_step5, tl0, tl1, window);
            double cu[WIN_MAX_PIECES], cl[WIN_MAX_PIECES];
            double surv = 1.0, TU = 0.0, TL = 0.0;      // surv = 2^-B: survival from the newer end of the piece to the read-out
            const int np = min(c.n, WIN_MAX_PIECES);
            int k_old = 0;          // pieces older than this one are not evaluated: their sum is provably negligible
            for (int k = np - 1; k >= 0; --k) {
                const PieceDesc d = describe_piece(c, k, cycle, inv_cycle, inv_step5, 5, window);
                if (d.div_after) surv *= 0.5;
                const double L = d.b - d.a;
                double integral, decay;
                if (L == step5) {       // a whole rate step: the same integral and decay in every cycle (per-particle tables)
                    integral = step_int[d.step]; decay = step_dec[d.step];
                } else {
                    const double gam = (double)r.gamma[d.step], al = (double)r.alpha[d.step];
                    const double A0 = al * (1.0 + (prm.scaling ? d.xa / cycle : 0.0)), A1 = prm.scaling ? al / cycle : 0.0;
                    integral = pon_step[d.step] * piece_integral(A0, A1, gam, L);
                    decay = exp(-gam * L);
                }
                const double I = surv * integral;
                const double lf = d.labelled ? (double)r.lam : 0.0;
                cu[k] = I * (1.0 - lf); cl[k] = I * lf;
                TU += cu[k]; TL += cl[k];
                surv *= decay;
                // Everything older than this piece is born before the label window (unlabelled only) and survives with at
                // most `surv`: once even the largest possible birth rate over the remaining time stays below 2^-20 of the
                // bound itself, the older pieces cannot move the start time and the walk stops.
                if (k > 0 && (!window || d.a <= tl0) &&
                    rate_max * (d.a - t_min) * surv <= 9.5367431640625e-7 * fmax(eps * TU, floor_abs)) {
                    k_old = k;
                    break;
                }
            }
            for (int k = 0; k < k_old; ++k) { cu[k] = 0.0; cl[k] = 0.0; }
            // drop the oldest pieces while their sum stays below the bound for both species
            double MU = 0.0, ML = 0.0;
            int j = k_old;
            for (; j < np; ++j) {
                const double mu = MU + cu[j], ml = ML + cl[j];
                if (!(mu <= fmax(eps * (TU - mu), floor_abs) && ml <= fmax(eps * (TL - ml), floor_abs))) break;
                MU = mu; ML = ml;
            }
            if (j >= np) {
                s0 = age;          // nothing to simulate: every expected count is below 2^-30
            } else {
                // part of piece j can go as well: bisection on the cut inside it
                const PieceDesc d = describe_piece(c, j, cycle, inv_cycle, inv_step5, 5, window);
                const double L = d.b - d.a, gam = (double)r.gamma[d.step], al = (double)r.alpha[d.step];
                const double A0 = al * (1.0 + (prm.scaling ? d.xa / cycle : 0.0)), A1 = prm.scaling ? al / cycle : 0.0;
                const double full = piece_integral(A0, A1, gam, L);
                const double lf = d.labelled ? (double)r.lam : 0.0;
                double lo = 0.0, hi = L;       // dropping [a, a + lo) is fine, [a, a + hi) is not
                if (full > 0.0) {
                    for (int it = 0; it < 12; ++it) {
                        const double mid = 0.5 * (lo + hi);
                        const double frac = piece_integral(A0, A1, gam, mid) * exp(-gam * (L - mid)) / full;
                        const double mu = MU + cu[j] * frac, ml = ML + cl[j] * frac;
                        (void)lf;
                        if (mu <= fmax(eps * (TU - mu), floor_abs) && ml <= fmax(eps * (TL - ml), floor_abs)) lo = mid; else hi = mid;
                    }
                }
                s0 = d.a + lo;
                if (prm.m == 3) {
                    // kon varies: the gene starts in the stationary law of its step, which is only approximate; its memory
                    // exp(-(kon + koff) t) must have decayed by 2^-6 before the first transcript that counts is born
                    double need = 6.0, t = s0;
                    int k = j;
                    while (need > 0.0 && t > t_min) {
                        const PieceDesc q = describe_piece(c, k, cycle, inv_cycle, inv_step5, 5, window);
                        const double rate = ((double)r.kon[q.step] + (double)r.koff[q.step]) * 1.4426950408889634;
                        const double avail = (t - q.a) * rate;
                        if (avail >= need) { t -= need / rate; need = 0.0; }
                        else { need -= avail; t = q.a; k -= 1; if (k < 0) break; }
                    }
                    s0 = fmax(t, t_min);
                }
            }
        }
        if (s0 < t_min) s0 = t_min;
        out_start[p * ABC_NREAD + readout] = (float)s0;
        // expected draws of one lineage: switches + one discarded draw per piece
        {
            const int spc = tele_steps_per_cycle(prm.m, r.gamma[0], (float)cycle);
            const double stepc = (spc == 1) ? cycle : step5;
            const Cuts c = make_cuts((double)(float)s0, age, stepc, (spc == 1) ? inv_cycle : inv_step5, tl0, tl1, window);
            double lineage = 0.0;
            for (int k = 0; k < c.n; ++k) {
                const PieceDesc d = describe_piece(c, k, cycle, inv_cycle, inv_step5, spc, window);
                const double kon = (double)r.kon[d.step], koff = (double)r.koff[d.step];
                lineage += 2.0 * kon * koff / (kon + koff) * (d.b - d.a) + 1.0;
            }
            // a lineage of more than 5e8 draws (rates far outside any prior box; 4e5 at its corner) is refused
            bad = bad || !(lineage < 5.0e8);
            cost += lineage + 40.0;       // + set-up, Poisson read-out and thinning, in draws
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(FULL, cost, o);
    bad = __any_sync(FULL, bad);
    if (lane == 0) {
        rates[p].pad0 = (cost == cost && cost > 0.0 && !bad) ? (float)cost : 0.0f;
        rates[p].pad1 = bad ? 1.0f : 0.0f;       // refused: its moments become NaN, it is never accepted
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TELE_WARPS * 32, TELE_MIN_CTAS)
abc_tele_kernel(const AbcRates* __restrict__ rates, const AbcSsaParams prm, const float* __restrict__ win,
                const uint32_t* __restrict__ beta_q32, unsigned long long* __restrict__ sums,
                unsigned long long* __restrict__ counters, unsigned int* __restrict__ work, uint32_t* __restrict__ cells_out,
                const int* __restrict__ order) {
    __shared__ TPiece tabs[TELE_WARPS][TELE_MAX_SEG];
    __shared__ AbcRates srates[TELE_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TPiece* tab = tabs[warp];
    const unsigned long long per_particle = (unsigned long long)ABC_NREAD * prm.chunks;
    const unsigned long long total = (prm.single_readout >= 0)
                                         ? (unsigned long long)prm.chunks
                                         : (unsigned long long)prm.n_particles * per_particle;
    unsigned long long acc_lineages = 0, acc_events = 0, acc_draws = 0;
    const double inv_cycle = 1.0 / prm.cycle, inv_step5 = 5.0 / prm.cycle, step5 = prm.cycle / 5.0;
    const float sc_inv_cycle = prm.scaling ? (float)inv_cycle : 0.0f;

    for (;;) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(work, 1u);
        item = __shfl_sync(FULL, item, 0);
        if ((unsigned long long)item >= total) break;
        long long p;
        int readout, chunk;
        if (prm.single_readout >= 0) {
            p = 0; readout = prm.single_readout; chunk = (int)item;
        } else {
            p = (long long)(item / per_particle);
            if (order != nullptr) p = order[p];      // heaviest predicted particles first
            unsigned int rem = (unsigned int)(item % per_particle);
            readout = (int)(rem / prm.chunks);
            chunk = (int)(rem % prm.chunks);
        }
        const int cond = readout / ABC_NAGE, age_i = readout % ABC_NAGE;

        // stage this particle's rates, then build the pieces of the read-out: one lane per piece
        __syncwarp();
        if (lane < (int)(sizeof(AbcRates) / 4)) ((uint32_t*)&srates[warp])[lane] = ((const uint32_t*)&rates[p])[lane];
        __syncwarp();
        if (srates[warp].pad1 != 0.0f) {
            // refused particle (non-finite or absurd rates): flag the read-out, simulate nothing
            if (lane == 0 && chunk == 0 && sums != nullptr) sums[((unsigned long long)p * ABC_NREAD + readout) * 5ull] = ~0ull;
            continue;
        }
        const double age = prm.agevec[age_i];
        const double tl0 = age - prm.pulse[cond] - prm.chase[cond], tl1 = age - prm.chase[cond];
        const bool window = prm.pulse[cond] > 0.0;
        const double s0 = (double)win[p * ABC_NREAD + readout];
        const int steps_per_cycle = tele_steps_per_cycle(prm.m, srates[warp].gamma[0], (float)prm.cycle);
        const double step_cut = (steps_per_cycle == 1) ? prm.cycle : step5, inv_step_cut = (steps_per_cycle == 1) ? inv_cycle : inv_step5;
        const float step_len = (float)step_cut;
        const Cuts cuts = make_cuts(s0, age, step_cut, inv_step_cut, tl0, tl1, window);
        const int n_seg = min(cuts.n, TELE_MAX_SEG);
        unsigned int kinds = 0u;         // forms of F among the pieces: bit 1 closed form, bit 2 six-term series, bit 3 four-term
        for (int k = lane; k < n_seg; k += 32) {
            const PieceDesc d = describe_piece(cuts, k, prm.cycle, inv_cycle, inv_step5, steps_per_cycle, window);
            TPiece t = make_piece(srates[warp], d, sc_inv_cycle, step_len);
            if (k == n_seg - 1) t.meta |= TP_LAST;
            tab[k] = t;
            kinds |= (t.k1 != 0.0f) ? 2u : ((t.e0 != 0.0f) ? 8u : 4u);
        }
        kinds = __reduce_or_sync(FULL, kinds);
        __syncwarp();

        const int cell = chunk * 32 + lane;
        const bool live = cell < prm.n_cells;
        const unsigned long long gp = (unsigned long long)(prm.particle_offset + p);
        Lineage s;
        s.c1 = (uint32_t)gp; s.c2 = (uint32_t)(gp >> 32);
        s.c3 = abc_tag((uint32_t)cell, (uint32_t)readout, (uint32_t)prm.m, ABC_DOM_SSA);
        s.k0 = prm.seed_lo; s.k1 = prm.seed_hi;
        s.ctr = 0u; s.U = 0.0f; s.L = 0.0f; s.g = 0; s.n_events = 0u;
        uint32_t Ud = 0u, Ld = 0u, n_cross = 0u;

        // initial gene state ~ the stationary law of the rate step the lineage starts in (exact for constant kon, koff)
        {
            const double mid0 = (n_seg > 0) ? s0 + 0.5 * (double)tab[0].len : s0;
            int st = (int)((mid0 - prm.cycle * (double)floor_div(mid0, prm.cycle, inv_cycle)) * inv_step5);
            st = st < 0 ? 0 : (st > 4 ? 4 : st);
            const float kon = srates[warp].kon[st], koff = srates[warp].koff[st];
            const uint4 b = next_block(s);
            // the smaller of P_on, P_off is compared with the low end of the uniform (full binary32 resolution near 0)
            const float u = f_fma((float)b.x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
            const float rs = f_add(kon, koff);
            s.g = (kon <= koff) ? ((u < __fdiv_rn(kon, rs)) ? 1 : 0) : ((u < __fdiv_rn(koff, rs)) ? 0 : 1);
        }

        float lam = 0.0f, lamL = 0.0f;      // Poisson means of U and L given the gene path
        {
            const TPiece* tp = tab;
            const bool done = !live || n_seg == 0;
            float x = 0.0f, acc = 0.0f;
            float sgn = s.g ? 1.0f : -1.0f;                 // +1 while the gene is on
            float len = INFINITY, e0 = 0.0f, k1 = 0.0f, p0 = 0.0f, p1 = 0.0f;
            uint32_t qsum = 0u, qb = 0u;                    // bit patterns: qon + qoff, q of the current state (0, 0: finished)
            if (!done) {
                const float4 a = *reinterpret_cast<const float4*>(&tp->len);
                const float4 f = *reinterpret_cast<const float4*>(&tp->e0);
                len = a.x; e0 = f.x; k1 = f.y; p0 = f.z; p1 = f.w;
                qsum = __float_as_uint(a.y) + __float_as_uint(a.z);
                qb = __float_as_uint(s.g ? a.z : a.y);
                acc = (s.g && k1 != 0.0f) ? -tp->p2 : 0.0f;
            }
            uint32_t invalid = 0u, ctr_done = s.ctr;
            const uint32_t ctr0 = s.ctr;
            // The loop, instantiated per form of F: when every piece of the read-out has the same form (always for models 1-4,
            // where the decay rate is constant along the lineage) the loop body carries no branch on it (KIND 1 closed form,
            // 2 six-term series, 3 four-term series); mixed schedules (model 5 with decay steps on both sides of a split) take
            // the generic body (KIND 0), which selects per lane and block.
            auto telegraph = [&](auto kind_tag) {
            constexpr int KIND = decltype(kind_tag)::value;
            for (;;) {
                const uint4 b = next_block(s);
                float l0, l1, l2, l3;
                {
                    const float u0 = f_fma((float)b.x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
                    const float u1 = f_fma((float)b.y, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
                    const float u2 = f_fma((float)b.z, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
                    const float u3 = f_fma((float)b.w, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(u0));
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(u1));
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u2));
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l3) : "f"(u3));
                }
                // four switch times: the waiting-time factor alternates between the two gene states
                const float qc = __uint_as_float(qb), qo = __uint_as_float(qsum - qb);
                const float x1 = f_fma(l0, qc, x), x2 = f_fma(l1, qo, x1);
                const float x3 = f_fma(l2, qc, x2), x4 = f_fma(l3, qo, x3);
                float F1, F2, F3, F4;
                if (KIND == 1 || (KIND == 0 && k1 != 0.0f)) {
                    float d1, d2, d3, d4;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(d1) : "f"(f_fma(k1, x1, e0)));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(d2) : "f"(f_fma(k1, x2, e0)));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(d3) : "f"(f_fma(k1, x3, e0)));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(d4) : "f"(f_fma(k1, x4, e0)));
                    F1 = f_mul(d1, f_fma(p1, x1, p0)); F2 = f_mul(d2, f_fma(p1, x2, p0));
                    F3 = f_mul(d3, f_fma(p1, x3, p0)); F4 = f_mul(d4, f_fma(p1, x4, p0));
                } else if (KIND == 3 || (KIND == 0 && e0 != 0.0f)) {
                    const float2 c = *reinterpret_cast<const float2*>(&tp->p2);
                    float h1 = f_fma(c.y, x1, c.x), h2 = f_fma(c.y, x2, c.x), h3 = f_fma(c.y, x3, c.x), h4 = f_fma(c.y, x4, c.x);
                    h1 = f_fma(h1, x1, p1); h2 = f_fma(h2, x2, p1); h3 = f_fma(h3, x3, p1); h4 = f_fma(h4, x4, p1);
                    h1 = f_fma(h1, x1, p0); h2 = f_fma(h2, x2, p0); h3 = f_fma(h3, x3, p0); h4 = f_fma(h4, x4, p0);
                    F1 = f_mul(h1, x1); F2 = f_mul(h2, x2); F3 = f_mul(h3, x3); F4 = f_mul(h4, x4);
                } else {
                    const float4 c = *reinterpret_cast<const float4*>(&tp->p2);
                    float h1 = f_fma(c.w, x1, c.z), h2 = f_fma(c.w, x2, c.z), h3 = f_fma(c.w, x3, c.z), h4 = f_fma(c.w, x4, c.z);
                    h1 = f_fma(h1, x1, c.y); h2 = f_fma(h2, x2, c.y); h3 = f_fma(h3, x3, c.y); h4 = f_fma(h4, x4, c.y);
                    h1 = f_fma(h1, x1, c.x); h2 = f_fma(h2, x2, c.x); h3 = f_fma(h3, x3, c.x); h4 = f_fma(h4, x4, c.x);
                    h1 = f_fma(h1, x1, p1); h2 = f_fma(h2, x2, p1); h3 = f_fma(h3, x3, p1); h4 = f_fma(h4, x4, p1);
                    h1 = f_fma(h1, x1, p0); h2 = f_fma(h2, x2, p0); h3 = f_fma(h3, x3, p0); h4 = f_fma(h4, x4, p0);
                    F1 = f_mul(h1, x1); F2 = f_mul(h2, x2); F3 = f_mul(h3, x3); F4 = f_mul(h4, x4);
                }
                const bool v4 = x4 < len;
                if (__all_sync(FULL, v4)) {
                    // no lane reaches the end of its piece within this block
                    acc = f_fma(sgn, f_add(f_add(F1, -F2), f_add(F3, -F4)), acc);
                    x = x4;
                    continue;
                }
                // some lane crosses: the same expression with the draws beyond the boundary masked (identical arithmetic
                // for the lanes that stay inside, so a result never depends on which lanes share the warp)
                const bool v1 = x1 < len, v2 = x2 < len, v3 = x3 < len;
                F1 = v1 ? F1 : 0.0f; F2 = v2 ? F2 : 0.0f; F3 = v3 ? F3 : 0.0f; F4 = v4 ? F4 : 0.0f;
                acc = f_fma(sgn, f_add(f_add(F1, -F2), f_add(F3, -F4)), acc);
                x = v4 ? x4 : (v3 ? x3 : (v2 ? x2 : (v1 ? x1 : x)));
                if ((v1 != v2) || (v3 != v4)) {         // an odd number of switches: the gene state changed
                    sgn = -sgn;
                    qb = qsum - qb;
                }
                if (!v4) {
                    // end of the piece (memoryless: the crossing draw and the rest of the block are discarded)
                    invalid += 4u - ((uint32_t)v1 + (uint32_t)v2 + (uint32_t)v3);
                    const float gs = f_fma(sgn, 0.5f, 0.5f);
                    const float2 fl = *reinterpret_cast<const float2*>(&tp->Flen);
                    const int meta = tp->meta;
                    const float dec = tp->dec;
                    const float inc = f_fma(gs, fl.x, acc), incL = f_mul(fl.y, inc);
                    lam = f_fma(lam, dec, f_add(inc, -incL));
                    lamL = f_fma(lamL, dec, incL);
                    n_cross += 1u;
                    x = 0.0f;
                    if (meta & TP_DIV) {                // cell division: a Poisson count thins to half its mean
                        lam = f_mul(lam, 0.5f);
                        lamL = f_mul(lamL, 0.5f);
                    }
                    if (meta & TP_LAST) {
                        ctr_done = s.ctr;
                        len = INFINITY; qb = 0u; qsum = 0u;      // a finished lane idles: x stays, nothing crosses
                        k1 = 0.0f; p0 = 0.0f; p1 = 0.0f; e0 = 0.0f;
                    } else {
                        tp += 1;
                        const float4 a = *reinterpret_cast<const float4*>(&tp->len);
                        const float4 f = *reinterpret_cast<const float4*>(&tp->e0);
                        len = a.x; e0 = f.x; k1 = f.y; p0 = f.z; p1 = f.w;
                        qsum = __float_as_uint(a.y) + __float_as_uint(a.z);
                        qb = __float_as_uint(sgn > 0.0f ? a.z : a.y);
                        acc = (k1 != 0.0f) ? f_mul(-gs, tp->p2) : 0.0f;
                    }
                }
                if (__all_sync(FULL, !(len < INFINITY))) break;      // every lane has finished (len = inf marks it)
            }
            };
            if (!__all_sync(FULL, done)) {
                const int kind = (kinds == 2u) ? 1 : (kinds == 4u) ? 2 : (kinds == 8u) ? 3 : 0;    // warp uniform
                if (kind == 1) telegraph(std::integral_constant<int, 1>{});
                else if (kind == 2) telegraph(std::integral_constant<int, 2>{});
                else if (kind == 3) telegraph(std::integral_constant<int, 3>{});
                else telegraph(std::integral_constant<int, 0>{});
            }
            if (live) {
                // draws = switches + boundary crossings; the words discarded after a crossing are not counted
                s.n_events = 4u * (ctr_done - ctr0) - invalid;
                s.ctr = ctr_done;                       // the read-out continues the lane's own stream
                s.g = (sgn > 0.0f) ? 1 : 0;
            }
        }
        if (live) {
            // Read-out.  U ~ Poisson(Lam_U), L ~ Poisson(Lam_L) given the gene path, then U' ~ Bin(U, beta), L' ~ Bin(L, beta)
            // with the cell's capture efficiency (scripts/model.jl:221-239).  A binomially thinned Poisson variable is
            // Poisson again and independent of the part removed, so (U', L') ~ Poisson(beta Lam_U) x Poisson(beta Lam_L)
            // is drawn directly -- the same joint law without simulating the molecules that are thrown away; the
            // per-cell debug output also draws the removed part and reports U = U' + U''.
            WordSrc ws; ws.avail = 0;
            float beta = 1.0f;
            if (prm.downsampling) {
                const int grp = (cond < 6 ? 0 : ABC_NAGE) + age_i;
                const uint32_t off = (uint32_t)prm.beta_off[grp];
                const uint32_t cnt = (uint32_t)prm.beta_off[grp + 1] - off;
                beta = f_mul((float)beta_q32[off + __umulhi(next_word(ws, s), cnt)], 2.3283064365386963e-10f);
            }
            const float fu = poisson_draw(f_mul(lam, beta), ws, s);
            const float fl = poisson_draw(f_mul(lamL, beta), ws, s);
            Ud = (uint32_t)fu; Ld = (uint32_t)fl;
            if (cells_out != nullptr) {
                uint32_t Uc = Ud, Lc = Ld;
                if (prm.downsampling) {
                    const float rest = f_add(1.0f, -beta);
                    Uc += (uint32_t)poisson_draw(f_mul(lam, rest), ws, s);
                    Lc += (uint32_t)poisson_draw(f_mul(lamL, rest), ws, s);
                }
                cells_out[0 * prm.n_cells + cell] = Uc;
                cells_out[1 * prm.n_cells + cell] = Lc;
                cells_out[2 * prm.n_cells + cell] = Ud;
                cells_out[3 * prm.n_cells + cell] = Ld;
            }
        }
        __syncwarp();
        // per read-out moment sums (exact integers: order independent => deterministic)
        const unsigned long long u = Ud, l = Ld;
        unsigned long long su = warp_sum_u64(u), sl = warp_sum_u64(l);
        unsigned long long suu = warp_sum_u64(u * u), sul = warp_sum_u64(u * l), sll = warp_sum_u64(l * l);
        if (lane == 0 && sums != nullptr) {
            unsigned long long* dst = sums + ((unsigned long long)p * ABC_NREAD + readout) * 5ull;
            atomicAdd(dst + 0, su); atomicAdd(dst + 1, sl); atomicAdd(dst + 2, suu);
            atomicAdd(dst + 3, sul); atomicAdd(dst + 4, sll);
        }
        acc_lineages += live ? 1ull : 0ull;
        acc_events += s.n_events;
        acc_draws += (unsigned long long)s.n_events + n_cross;
    }
    acc_lineages = warp_sum_u64(acc_lineages);
    acc_events = warp_sum_u64(acc_events);
    acc_draws = warp_sum_u64(acc_draws);
    if (lane == 0) {
        atomicAdd(counters + 0, acc_lineages);
        atomicAdd(counters + 1, acc_events);
        atomicAdd(counters + 2, acc_draws);
    }
}

// ------------------------------------------------------------------------------------------------
int abc_launch_window(AbcRates* d_rates, const AbcSsaParams& prm, float* d_win, int64_t n, cudaStream_t st) {
    if (n <= 0) return ABC_OK;
    if ((prm.n_pre + 1) * 5 + 3 > TELE_MAX_SEG) {
        abc_set_error("n_pre_cycles too large for the telegraph kernel's schedule table");
        return ABC_ERR_ARG;
    }
    const int threads = 128;
    const long long blocks = (n * 32 + threads - 1) / threads;
    abc_window_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_rates, prm, d_win, (long long)n);
    ABC_CUDA_CHECK(cudaGetLastError());
    return ABC_OK;
}

int abc_launch_tele(const AbcRates* d_rates, const AbcSsaParams& prm, const float* d_win, const uint32_t* d_beta_q32,
                    unsigned long long* d_sums, unsigned long long* d_counters, unsigned int* d_work,
                    uint32_t* d_cells_out, const int* d_order, int sm_count, cudaStream_t st) {
    ABC_CUDA_CHECK(cudaMemsetAsync(d_work, 0, sizeof(unsigned int), st));
    const void* fn = (const void*)abc_tele_kernel;
    ABC_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    ABC_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, TELE_WARPS * 32, 0));
    if (per_sm < 1) per_sm = 1;
    unsigned long long items = (prm.single_readout >= 0) ? (unsigned long long)prm.chunks
                               : (unsigned long long)prm.n_particles * ABC_NREAD * prm.chunks;
    if (items > 0xFFFFFFF0ull - 65536ull) {
        abc_set_error("too many work items in one SSA launch (%llu); split the batch", items);
        return ABC_ERR_ARG;
    }
    unsigned long long want = (items + TELE_WARPS - 1) / TELE_WARPS;
    unsigned long long grid = (unsigned long long)sm_count * per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    void* args[] = {(void*)&d_rates, (void*)&prm, (void*)&d_win, (void*)&d_beta_q32, (void*)&d_sums, (void*)&d_counters,
                    (void*)&d_work, (void*)&d_cells_out, (void*)&d_order};
    ABC_CUDA_CHECK(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(TELE_WARPS * 32), args, 0, st));
    ABC_CUDA_CHECK(cudaGetLastError());
    return ABC_OK;
}
