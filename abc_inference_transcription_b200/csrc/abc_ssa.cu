// abc_ssa.cu -- Gillespie direct-method SSA of the 5 transcription models over all six channels (ssa_hybrid_burnin = 0, the
// exact-math variant that is bit-identical to the CPU oracle, and = 1 with a telegraph burn-in), capture-efficiency thinning,
// per read-out moment sums, prior draws, rates, moments and the 53 summary statistics.  The product sampler
// (ssa_hybrid_burnin = 2) is abc_tele.cu.  sm_100a.
//
// The stochastic process is the CME whose moments scripts/model.jl:74-86 (f), :98-111
// (periodic_boundary) and :221-239 (downsample) describe -- SURVEY.md section 8a-CME:
//   state (g in {0,1}, U, L);  off->on kon(t)(1-g);  on->off koff*g;  U birth alpha(t)(1-lam(t))g;
//   L birth alpha(t)lam(t)g;  U death gamma(t)U;  L death gamma(t)L;
//   alpha(t) = 10^theta_alpha(step) * (1 + scaling*mod(t,cycle)/cycle)      (model.jl:1-27)
//   lam(t) = 10^theta_lambda on [age-pulse-chase, age-chase], else 0            (model.jl:58-64, :183)
//   at every multiple of the cycle: U <- Bin(U,1/2), L <- Bin(L,1/2)            (model.jl:98-111)
//   read-out at t = age, then U' ~ Bin(U,beta), L' ~ Bin(L,beta), beta drawn from the empirical
//   capture efficiencies of the (pulse|chase, age cluster) group                (model.jl:221-239)
//
// Mapping: one thread per cell lineage, one warp per 32 cells of one (particle, condition, age)
// read-out, persistent CTAs pulling (particle, read-out, chunk) items from a global atomic queue.
// The piecewise-constant rate schedule of the read-out is staged per warp in shared memory.
// Random numbers: Philox4x32-10, counter = (block index, particle lo, particle hi, tag(cell, read-out,
// model)), key = seed: results do not depend on the launch geometry or on the number of GPUs.
#include "abc_ssa_dev.cuh"
#include <cub/cub.cuh>

// one sub-interval of the rate schedule: constant kon, koff, gamma, lam; alpha(x) = A0 + A1*x
struct __align__(8) Seg {
    float len, kon, koff, A0;
    float A1, gam, lamf, pad;
};

#define SSA_MAX_CYCLES 14
#define SSA_SEG_PER_CYCLE 7
#define SSA_WARPS 8

// the same sub-interval as the telegraph phase of the hybrid burn-in sees it (56 B): waiting-time factors and the
// antiderivative F of alpha(w) exp(-gam (len - w)), so that the Poisson mean advances by F(x2) - F(x1) over an "on"
// stretch [x1, x2].
//   k1 != 0: F(x) = 2^(k1 (x - len)) (p0 + p1 x),  p0 = A0/gam - A1/gam^2, p1 = A1/gam, p2 = F(0)
//   k1 == 0 (gam * step < 1/4): F(x) = x (p0 + p1 x + ... + p5 x^5), the series of the same integral from 0
struct __align__(8) TSeg {
    float len, qon, qoff, dec;     // q = -ln2 / rate (waiting time = lg2(u) * q), dec = exp(-gam len)
    float Flen, k1, p0, p1;
    float p2, p3, p4, p5;
    float lamf;                    // labelled share of the births in this sub-interval
    int meta;                      // bit 0: a cell division follows, bit 1: last sub-interval of the telegraph phase,
                                   // bits 8..: byte offset to the next sub-interval's slot
};
#define TSEG_DIV 1
#define TSEG_LAST 2
union Slot {
    Seg s;
    TSeg t;
};

struct WarpTable {
    Slot slot[SSA_MAX_CYCLES][SSA_SEG_PER_CYCLE];
    int n_ent[SSA_MAX_CYCLES];
    int c_star, e_star;     // cycle / entry at which the label window opens (hybrid burn-in hands over here)
};

// one event draw.  returns true if the sub-interval boundary was crossed (no reaction fired).
// Channel layout on [0, tot): switch at the bottom (u*tot < a_sw), death at the top ((1-u)*tot < a_d, U first),
// birth in between (labelled iff u*tot - a_sw < lam*a_b).  Both ends of the uniform keep full binary32
// resolution, and an empty death channel (a_d == 0) or an empty species can never be selected.
template <bool EXACT>
__device__ __forceinline__ bool ssa_step(Lineage& s, float& x, const Seg& sg, uint32_t wt, uint32_t wc) {
    const int g = s.g;
    const float asw = pick(g, sg.kon, sg.koff);
    const float n = f_add(s.U, s.L);
    const float ad = f_mul(sg.gam, n);
    const float ab = gate(g, f_fma(sg.A1, x, sg.A0));
    const float c1 = gate(g, sg.A1);
    const float base = f_add(asw, ad);
    const float c0 = f_add(base, ab);
    const float E = exp_variate<EXACT>(wt);
    const float tau = wait_time<EXACT>(c0, c1, E);
    const float xn = f_add(x, tau);
    if (!(xn < sg.len)) return true;
    x = xn;
    const float abn = gate(g, f_fma(sg.A1, xn, sg.A0));
    const float tot = f_add(base, abn);
    const float t32 = f_mul(tot, 2.3283064365386963e-10f);
    const float rs = f_mul((float)wc, t32);
    const float rb = f_mul((float)(~wc), t32);
    // (float)wc rounds the top 128 words up to 2^32 (rs == tot): when the switch is the only channel it must still fire
    const bool sw = (rs < asw) || !(asw < tot);
    const bool death = !sw && (rb < ad);
    const bool dU = rb < f_mul(sg.gam, s.U);
    const bool birth = !sw && !death;
    const bool lab = birth && (f_add(rs, -asw) < f_mul(sg.lamf, abn));
    s.g = g ^ (int)sw;
    if (birth && !lab) s.U = f_add(s.U, 1.0f);
    if (lab) s.L = f_add(s.L, 1.0f);
    if (death && dU) s.U = f_add(s.U, -1.0f);
    if (death && !dU) s.L = f_add(s.L, -1.0f);
    s.n_events += 1u;
    return false;
}


// ------------------------------------------------------------------------------------------------
// Hybrid burn-in (exact).  Before the label window opens the only species are g and U, and the gene switches
// autonomously (kon, koff do not depend on the mRNA counts).  Conditional on the gene path, transcripts are
// born as an inhomogeneous Poisson process of rate alpha(t) g(t) and survive independently (death hazard
// gamma(t), probability 1/2 at every division), so the number alive at time t* is Poisson(Lam) with
//     Lam(t*) = int alpha(s) g(s) exp(-int_s^t* gamma) 2^-(divisions in (s,t*]) ds .
// Phase A simulates only the telegraph process with Gillespie's method and integrates Lam in closed form on
// every constant-g stretch (alpha is linear in time there); at t* it draws U ~ Poisson(Lam), L = 0, and the
// full direct-method SSA of all six channels takes over.  The law of (g, U) at t* is exactly the one the
// full SSA would have produced from the same start (U = 0 at the first simulated cycle).
//
// Lam over one sub-interval: Lam_end = Lam_start exp(-gam len) + sum over "on" stretches [x1, x2] of F(x2) - F(x1),
// F' = alpha(w) exp(-gam (len - w)).  A switch at x therefore only adds +F(x) (the gene goes off) or -F(x) (it goes on) to
// an accumulator: one MUFU and four FP32 operations per draw; the boundary terms F(0), F(len) are per-segment constants.
__device__ __forceinline__ float tseg_F(const TSeg* tp, float len, float k1, float p0, float p1, float x) {
    if (k1 != 0.0f) {
        float D;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(D) : "f"(f_mul(k1, f_add(x, -len))));
        return f_mul(D, f_fma(p1, x, p0));
    }
    float h = f_fma(tp->p5, x, tp->p4);
    h = f_fma(h, x, tp->p3);
    h = f_fma(h, x, tp->p2);
    h = f_fma(h, x, p1);
    h = f_fma(h, x, p0);
    return f_mul(h, x);
}

// Seg -> TSeg, once per sub-interval and read-out.  step_len = cycle / 5 bounds len.
__device__ __forceinline__ TSeg make_tseg(const Seg& sg, float step_len) {
    const float len = sg.len, gam = sg.gam, A0 = sg.A0, A1 = sg.A1;
    float dec;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(dec) : "f"(f_mul(f_mul(gam, len), -1.4426950408889634f)));
    TSeg t;
    t.len = len;
    t.lamf = sg.lamf; t.meta = 0;
    t.qon = __fdiv_rn(-0.693147182464599609375f, sg.kon);
    t.qoff = __fdiv_rn(-0.693147182464599609375f, sg.koff);
    t.dec = dec;
    if (f_mul(gam, step_len) < 0.25f) {
        // F(x) = dec * sum_j x^j [A0 gam^(j-1)/j! + A1 gam^(j-2)/((j-2)! j)]: six terms in the kernel (the next one is
        // below 5e-8 relative), nine for the boundary value F(len)
        const float rj[9] = {1.0f, 1.0f / 2, 1.0f / 6, 1.0f / 24, 1.0f / 120, 1.0f / 720, 1.0f / 5040, 1.0f / 40320, 1.0f / 362880};
        const float sj[9] = {0.0f, 1.0f / 2, 1.0f / 3, 1.0f / 8, 1.0f / 30, 1.0f / 144, 1.0f / 840, 1.0f / 5760, 1.0f / 45360};
        float cj[9], gp = dec, gq = 0.0f;       // gp = dec gam^(j-1), gq = dec gam^(j-2)
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
        t.k1 = 0.0f;
        t.p0 = cj[0]; t.p1 = cj[1]; t.p2 = cj[2]; t.p3 = cj[3]; t.p4 = cj[4]; t.p5 = cj[5];
    } else {
        const float e = __fdiv_rn(A1, gam), c = __fdiv_rn(f_add(A0, -e), gam);
        t.k1 = f_mul(gam, 1.4426950408889634f);
        t.p0 = c; t.p1 = e;
        t.p2 = f_mul(dec, c);               // F(0)
        t.p3 = 0.0f; t.p4 = 0.0f; t.p5 = 0.0f;
        t.Flen = f_fma(e, len, c);
    }
    return t;
}

// Adaptive burn-in (modes 1 and 2).  A lineage that starts k complete cycles before the read-out cycle with U = 0 misses
// the transcripts born earlier; their share of Lam at the read-out is at most (1/2 exp(-sum_s gam_s cycle/5))^k (dilution
// and decay per cycle), so the smallest k with k (1 + log2(e) gam cycle) >= n_pre has the truncation bias bound 2^-n_pre of
// the configured n_pre cycles.  The gene starts in its stationary law, which is exact for constant kon, koff; for model 3
// (kon varies over the cycle) the start law is only approximate and its memory, exp(-(kon + koff) t), must have
// decayed as well.  The reference's own burn-in is adaptive too
// (transient_phase iterates until 1 % change, scripts/model.jl:114-142).
__device__ __forceinline__ int burnin_cycles(const AbcRates& r, const AbcSsaParams& prm, int cond, int age_i) {
    const float step = (float)(prm.cycle / 5.0);
    float zg = 0.0f, zr = 0.0f;
#pragma unroll
    for (int j = 0; j < 5; ++j) { zg += r.gamma[j]; zr += r.kon[j] + r.koff[j]; }
    float bits = 1.0f + 1.4426950f * zg * step;
    float need = (float)prm.n_pre;
    if (prm.m == 3) { bits = fminf(bits, 1.4426950f * zr * step); need += 6.0f; }
    int k = prm.n_pre;
    if (bits * (float)prm.n_pre >= need) k = (int)ceilf(need / bits);
    // The k cycles are counted back from the cycle in which the label window opens, not from the read-out cycle: the
    // unlabelled transcripts seen after a long pulse are all older than the window, and correlations normalise their
    // magnitude away, so the bound must hold relative to that history too.
    const double tl0 = prm.agevec[age_i] - prm.pulse[cond] - prm.chase[cond];
    const int k_win = (tl0 < 0.0) ? (int)ceil(-tl0 / prm.cycle) : 0;
    k = max(k + k_win, 1);
    return min(k, prm.n_pre);
}

// build the schedule of one read-out: lane c builds cycle c (c >= c0)
// n_steps: rate steps per cycle (5 = scripts/model.jl:1-22; 1 when no rate varies and the caller does not need the
// step boundaries, i.e. the telegraph-only mode of models 1 and 2)
__device__ __forceinline__ void build_table(WarpTable& tab, const AbcRates& r, const AbcSsaParams& prm,
                                            int cond, int age_i, int lane, int n_steps, int c0) {
    const int c = lane;
    if (c >= c0 && c <= prm.n_pre) {
        const double cycle = prm.cycle;
        const double age = prm.agevec[age_i];
        const double tl0 = age - prm.pulse[cond] - prm.chase[cond];
        const double tl1 = age - prm.chase[cond];
        const double Tc = (double)(c - prm.n_pre) * cycle;
        const double cyc_end = (c == prm.n_pre) ? age : cycle;
        const double l0 = tl0 - Tc, l1 = tl1 - Tc;
        const double step_len = cycle / (double)n_steps;
        const double sc = prm.scaling ? 1.0 : 0.0;
        double pos = 0.0;
        int k = 0, n = 0;
        while (pos < cyc_end && n < SSA_SEG_PER_CYCLE) {
            double step_end = (double)(k + 1) * step_len;
            double nxt = step_end < cyc_end ? step_end : cyc_end;
            if (l0 > pos && l0 < nxt) nxt = l0;
            if (l1 > pos && l1 < nxt) nxt = l1;
            const double mid = 0.5 * (pos + nxt);
            const bool lab = (mid >= l0) && (mid <= l1);
            Seg sg;
            sg.len = (float)(nxt - pos);
            sg.kon = r.kon[k];
            sg.koff = r.koff[k];
            sg.gam = r.gamma[k];
            sg.A0 = (float)((double)r.alpha[k] * (1.0 + sc * pos / cycle));
            sg.A1 = (float)((double)r.alpha[k] * sc / cycle);
            sg.lamf = lab ? r.lam : 0.0f;
            sg.pad = 0.0f;
            if (lab && pos == l0) { tab.c_star = c; tab.e_star = n; }
            tab.slot[c][n].s = sg;
            n += 1;
            pos = nxt;
            if (!(pos < step_end)) k += 1;
            if (k > n_steps - 1) k = n_steps - 1;
        }
        tab.n_ent[c] = n;
    }
}

template <bool EXACT, int HYBRID>      // HYBRID: 0 = six-channel direct method throughout, 1 = telegraph burn-in before the label
                                       // window (mode 2, telegraph to the read-out, is abc_tele.cu)
__global__ void __launch_bounds__(SSA_WARPS * 32, 4)
abc_ssa_kernel(const AbcRates* __restrict__ rates, const AbcSsaParams prm, const uint32_t* __restrict__ beta_q32,
               unsigned long long* __restrict__ sums, unsigned long long* __restrict__ counters,
               unsigned int* __restrict__ work, uint32_t* __restrict__ cells_out, const int* __restrict__ order) {
    __shared__ WarpTable tabs[SSA_WARPS];
    __shared__ AbcRates srates[SSA_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpTable& tab = tabs[warp];
    const unsigned long long per_particle = (unsigned long long)ABC_NREAD * prm.chunks;
    const unsigned long long total = (prm.single_readout >= 0)
                                         ? (unsigned long long)prm.chunks
                                         : (unsigned long long)prm.n_particles * per_particle;
    unsigned long long acc_lineages = 0, acc_events = 0, acc_draws = 0;

    for (;;) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(work, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
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

        // stage this particle's rates and the read-out's schedule in shared memory
        __syncwarp();
        if (lane < 24) ((uint32_t*)&srates[warp])[lane] = ((const uint32_t*)&rates[p])[lane];
        __syncwarp();
        if (lane == 0) { tab.c_star = -1; tab.e_star = 0; }
        __syncwarp();
        const int n_steps = 5;
        const int c0 = (HYBRID && prm.adaptive) ? prm.n_pre - burnin_cycles(srates[warp], prm, cond, age_i) : 0;   // first simulated cycle
        build_table(tab, srates[warp], prm, cond, age_i, lane, n_steps, c0);
        __syncwarp();
        if (lane == 0 && tab.c_star < 0) {
            // empty window: the telegraph phase runs to the read-out
            tab.c_star = prm.n_pre; tab.e_star = tab.n_ent[prm.n_pre];
        }
        __syncwarp();
        if (HYBRID) {      // sub-intervals before the hand-over: telegraph-phase form, one slot per lane and round
            const int c_star = tab.c_star, e_star = tab.e_star;
            const int n_conv = c_star * SSA_SEG_PER_CYCLE + e_star;
            const float step_len = (float)(prm.cycle / (double)n_steps);
            for (int k = c0 * SSA_SEG_PER_CYCLE + lane; k < n_conv; k += 32) {
                const int cc = k / SSA_SEG_PER_CYCLE, ee = k % SSA_SEG_PER_CYCLE;
                if (ee < tab.n_ent[cc]) {
                    const Seg sg = tab.slot[cc][ee].s;
                    TSeg t = make_tseg(sg, step_len);
                    int c2 = cc, e2 = ee + 1, meta = 0;
                    if (cc < c_star && e2 == tab.n_ent[cc]) { meta |= TSEG_DIV; c2 = cc + 1; e2 = 0; }
                    if (c2 == c_star && e2 == e_star) meta |= TSEG_LAST;
                    meta |= (((c2 - cc) * SSA_SEG_PER_CYCLE + (e2 - ee)) * (int)sizeof(Slot)) << 8;
                    t.meta = meta;
                    tab.slot[cc][ee].t = t;
                }
            }
        }
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

        if (live) {
            {   // initial gene state ~ Bernoulli(P_on)
                uint4 b = next_block(s);
                s.g = (b.x < srates[warp].pon_thr) ? 1 : 0;
            }
            int c_first = 0, e_first = 0;               // where the six-channel direct method starts
            if (HYBRID) {
                // phase A: telegraph process + closed-form Lam, one random word per draw.  Divisions only halve
                // Lam, so the lanes run through all phase-A cycles without waiting for each other.
                float x = 0.0f, lam = 0.0f, lamL = 0.0f;   // Poisson means of U and L given the gene path
                const TSeg* tp = &tab.slot[c0][0].t;
                bool done = (tab.c_star == c0) && (tab.e_star == 0);
                float len = -INFINITY, k1 = 0.0f, p0 = 0.0f, p1 = 0.0f, acc = 0.0f;
                float sgn = s.g ? 1.0f : -1.0f;         // +1 while the gene is on
                uint32_t qsum = 0u, qb = 0u;            // bit patterns: q of the current state, qon + qoff
                if (!done) {
                    len = tp->len; k1 = tp->k1; p0 = tp->p0; p1 = tp->p1;
                    qsum = __float_as_uint(tp->qon) + __float_as_uint(tp->qoff);
                    qb = __float_as_uint(s.g ? tp->qoff : tp->qon);
                    acc = (s.g && k1 != 0.0f) ? -tp->p2 : 0.0f;
                }
                const uint32_t ctr0 = s.ctr;
                uint32_t unused = 0u;                   // words of the last block that were not drawn
                while (!done) {
                    const uint4 b = next_block(s);
                    const uint32_t w4[4] = {b.x, b.y, b.z, b.w};
                    float l4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float u = f_fma((float)w4[j], 2.3283064365386963e-10f, 1.1641532182693481e-10f);
                        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l4[j]) : "f"(u));
                    }
                    // fast path: all four draws of every lane of the warp are switches inside the current sub-interval
                    // (the waiting-time factor alternates between the two gene states); same arithmetic as the
                    // draw-by-draw path below
                    const float qc = __uint_as_float(qb), qo = __uint_as_float(qsum - qb);
                    const float x1 = f_fma(l4[0], qc, x), x2 = f_fma(l4[1], qo, x1);
                    const float x3 = f_fma(l4[2], qc, x2), x4 = f_fma(l4[3], qo, x3);
                    if (__all_sync(__activemask(), x4 < len)) {
                        // the same four sequential FMAs as the draw-by-draw path below: a lineage's result does not depend
                        // on which lanes happen to be converged at the vote
                        acc = f_fma(sgn, tseg_F(tp, len, k1, p0, p1, x1), acc);
                        acc = f_fma(-sgn, tseg_F(tp, len, k1, p0, p1, x2), acc);
                        acc = f_fma(sgn, tseg_F(tp, len, k1, p0, p1, x3), acc);
                        acc = f_fma(-sgn, tseg_F(tp, len, k1, p0, p1, x4), acc);
                        x = x4;
                        continue;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float xn = f_fma(l4[j], __uint_as_float(qb), x);
                        if (xn < len) {                 // the gene switches at xn
                            acc = f_fma(sgn, tseg_F(tp, len, k1, p0, p1, xn), acc);
                            sgn = -sgn;
                            qb = qsum - qb;
                            x = xn;
                        } else if (!done) {             // sub-interval boundary (memoryless: the draw is discarded)
                            const float gs = f_fma(sgn, 0.5f, 0.5f);
                            const float inc = f_fma(gs, tp->Flen, acc), incL = f_mul(tp->lamf, inc), dec = tp->dec;
                            const int meta = tp->meta;
                            lam = f_fma(lam, dec, f_add(inc, -incL));
                            lamL = f_fma(lamL, dec, incL);
                            x = 0.0f; n_cross += 1u;
                            if (meta & TSEG_DIV) {      // cell division: a Poisson count thins to half its mean
                                lam = f_mul(lam, 0.5f);
                                lamL = f_mul(lamL, 0.5f);
                            }
                            if (meta & TSEG_LAST) {
                                done = true;
                                len = -INFINITY;        // a finished lane draws no further event in this block
                                unused = (uint32_t)(3 - j);
                            } else {
                                tp = (const TSeg*)((const char*)tp + (meta >> 8));
                                len = tp->len; k1 = tp->k1; p0 = tp->p0; p1 = tp->p1;
                                qsum = __float_as_uint(tp->qon) + __float_as_uint(tp->qoff);
                                qb = __float_as_uint(sgn > 0.0f ? tp->qoff : tp->qon);
                                acc = (k1 != 0.0f) ? f_mul(-gs, tp->p2) : 0.0f;
                            }
                        }
                    }
                }
                // draws = words consumed; every draw is either a switch or a boundary crossing
                s.n_events += 4u * (s.ctr - ctr0) - unused - n_cross;
                s.g = (sgn > 0.0f) ? 1 : 0;
                WordSrc ws; ws.avail = 0;               // hand over: U ~ Poisson(Lam); L = 0 unless the window was included
                s.U = poisson_draw(lam, ws, s);
                s.L = poisson_draw(lamL, ws, s);
                c_first = tab.c_star; e_first = tab.e_star;
            }
            for (int c = c_first; c <= prm.n_pre; ++c) {
                const int n_ent = tab.n_ent[c];
                int e = (c == c_first) ? e_first : 0;
                n_cross += (uint32_t)(n_ent - e);
                float x = 0.0f;
                Seg sg = tab.slot[c][e < n_ent ? e : 0].s;
                while (e < n_ent) {
                    const uint4 b = next_block(s);
                    if (ssa_step<EXACT>(s, x, sg, b.x, b.y)) {
                        e += 1; x = 0.0f;
                        if (e < n_ent) sg = tab.slot[c][e].s;
                    }
                    if (e < n_ent) {
                        if (ssa_step<EXACT>(s, x, sg, b.z, b.w)) {
                            e += 1; x = 0.0f;
                            if (e < n_ent) sg = tab.slot[c][e].s;
                        }
                    }
                }
                if (c < prm.n_pre) {   // cell division: binomial partitioning, gene state kept
                    WordSrc ws; ws.avail = 0;
                    s.U = (float)binhalf((uint32_t)s.U, ws, s);
                    s.L = (float)binhalf((uint32_t)s.L, ws, s);
                }
            }
            const uint32_t Uc = (uint32_t)s.U, Lc = (uint32_t)s.L;
            Ud = Uc; Ld = Lc;
            if (prm.downsampling) {
                WordSrc ws; ws.avail = 0;
                const int grp = (cond < 6 ? 0 : ABC_NAGE) + age_i;
                const uint32_t off = (uint32_t)prm.beta_off[grp];
                const uint32_t cnt = (uint32_t)prm.beta_off[grp + 1] - off;
                const uint32_t B = beta_q32[off + __umulhi(next_word(ws, s), cnt)];
                Ud = binom_q32(Uc, B, ws, s);
                Ld = binom_q32(Lc, B, ws, s);
            }
            if (cells_out != nullptr) {
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

int abc_launch_ssa(const AbcRates* d_rates, const AbcSsaParams& prm, const uint32_t* d_beta_q32,
                   unsigned long long* d_sums, unsigned long long* d_counters, unsigned int* d_work,
                   uint32_t* d_cells_out, const int* d_order, int exact_math, int sm_count, cudaStream_t st) {
    if (prm.n_pre + 1 > SSA_MAX_CYCLES) {
        abc_set_error("n_pre_cycles must be <= %d", SSA_MAX_CYCLES - 1);
        return ABC_ERR_ARG;
    }
    ABC_CUDA_CHECK(cudaMemsetAsync(d_work, 0, sizeof(unsigned int), st));
    // mode 2 (telegraph SSA to the read-out) is abc_tele.cu's kernel: the caller dispatches to abc_launch_tele
    const int hybrid = exact_math ? 0 : (prm.hybrid == 1 ? 1 : 0);
    const void* fn = exact_math    ? (const void*)abc_ssa_kernel<true, 0>
                     : hybrid == 1 ? (const void*)abc_ssa_kernel<false, 1>
                                   : (const void*)abc_ssa_kernel<false, 0>;
    // 4 CTAs x 45 KB of schedule tables per SM; the kernel does not use L1 (per device: set at every launch)
    ABC_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    ABC_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, SSA_WARPS * 32, 0));
    if (per_sm < 1) per_sm = 1;
    unsigned long long items = (prm.single_readout >= 0) ? (unsigned long long)prm.chunks
                               : (unsigned long long)prm.n_particles * ABC_NREAD * prm.chunks;
    if (items > 0xFFFFFFF0ull - 65536ull) {
        abc_set_error("too many work items in one SSA launch (%llu); split the batch", items);
        return ABC_ERR_ARG;
    }
    unsigned long long want = (items + SSA_WARPS - 1) / SSA_WARPS;
    unsigned long long grid = (unsigned long long)sm_count * per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    void* args[] = {(void*)&d_rates, (void*)&prm, (void*)&d_beta_q32, (void*)&d_sums, (void*)&d_counters, (void*)&d_work,
                    (void*)&d_cells_out, (void*)&d_order};
    ABC_CUDA_CHECK(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(SSA_WARPS * 32), args, 0, st));
    ABC_CUDA_CHECK(cudaGetLastError());
    return ABC_OK;
}

// ------------------------------------------------------------------------------------------------
// theta (log10, [n][P]) -> linear float rates.  vary_map of model.jl:30-43 / abc_simulation.jl:82-85.
__global__ void abc_rates_kernel(const double* __restrict__ theta, int m, long long n, AbcRates* __restrict__ out,
                                 int ssa_hybrid, int n_pre_adaptive, float cycle) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int P = (m <= 2) ? 5 : 9;
    const double* th = theta + i * P;
    const int vary = (m == 3) ? 0 : (m == 4) ? 2 : (m == 5) ? 3 : -1;
    AbcRates r;
    int k = 0;
    float* dst[4] = {r.kon, r.koff, r.alpha, r.gamma};
    for (int q = 0; q < 4; ++q) {
        if (q == vary) {
            for (int j = 0; j < 5; ++j) dst[q][j] = (float)abc_exp10_det(th[k + j]);
            k += 5;
        } else {
            float v = (float)abc_exp10_det(th[k]);
            for (int j = 0; j < 5; ++j) dst[q][j] = v;
            k += 1;
        }
    }
    double lam = abc_exp10_det(th[k]);
    if (!(lam <= 1.0)) lam = 1.0;
    if (!(lam >= 0.0)) lam = 0.0;
    r.lam = (float)lam;
    // gene state at the start of the first simulated cycle ~ stationary of the last rate step
    double kon = (double)r.kon[4], koff = (double)r.koff[4];
    double pon = kon / (kon + koff);
    double thr = pon * 4294967296.0;
    r.pon_thr = (thr >= 4294967295.0) ? 0xFFFFFFFFu : (thr > 0.0 ? (uint32_t)thr : 0u);
    // predicted SSA work per simulated hour (in telegraph draws), used only to schedule the heaviest particles first
    // (longest-processing-time order); it never influences a result.  Births and deaths are simulated over the whole
    // lineage (mode 0), inside the label window only (mode 1: ~12 h of ~210 h, at about twice the instructions of a
    // telegraph draw) or not at all (mode 2).
    const float wb = (ssa_hybrid == 0) ? 2.0f : (ssa_hybrid == 1) ? 0.2f : 0.0f;
    float cost = 0.0f;
    for (int j = 0; j < 5; ++j) {
        const float s2 = r.kon[j] + r.koff[j];
        const float sw = 2.0f * r.kon[j] * r.koff[j] / s2;
        const float br = r.alpha[j] * (m != 2 ? 1.5f : 1.0f) * r.kon[j] / s2;
        cost += 0.2f * (sw + wb * br);
    }
    if (n_pre_adaptive > 0) {   // shorter burn-in for short-lived transcripts (burnin_cycles)
        float zg = 0.0f, zr = 0.0f;
        for (int j = 0; j < 5; ++j) { zg += r.gamma[j]; zr += r.kon[j] + r.koff[j]; }
        float bits = 1.0f + 1.4426950f * zg * cycle * 0.2f;
        if (m == 3) bits = fminf(bits, 1.4426950f * zr * cycle * 0.2f);
        const float k = fminf((float)n_pre_adaptive, fmaxf(1.0f, (float)n_pre_adaptive / bits) + 0.6f);   // + mean k_win
        cost *= (k + 0.5f) / ((float)n_pre_adaptive + 0.5f);
    }
    r.pad0 = (cost == cost && cost > 0.0f) ? cost : 0.0f;
    r.pad1 = 0.0f;
    out[i] = r;
}

// particle order by predicted cost, descending (CUB radix sort on the float bit patterns; plumbing only)
__global__ void abc_cost_keys_kernel(const AbcRates* __restrict__ rates, int n, unsigned int* __restrict__ keys,
                                     int* __restrict__ idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = __float_as_uint(rates[i].pad0);
    idx[i] = i;
}

size_t abc_order_temp_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const unsigned int*)nullptr, (unsigned int*)nullptr,
                                              (const int*)nullptr, (int*)nullptr, n);
    return bytes;
}

int abc_launch_order(const AbcRates* d_rates, int n, unsigned int* d_keys_in, unsigned int* d_keys_out, int* d_idx_in,
                     int* d_order, void* d_temp, size_t temp_bytes, cudaStream_t st) {
    if (n <= 0) return ABC_OK;
    abc_cost_keys_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_rates, n, d_keys_in, d_idx_in);
    ABC_CUDA_CHECK(cudaGetLastError());
    ABC_CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(d_temp, temp_bytes, d_keys_in, d_keys_out, d_idx_in, d_order,
                                                             n, 0, 32, st));
    return ABC_OK;
}

int abc_launch_rates(const double* d_theta, int m, int64_t n, AbcRates* d_rates, int ssa_hybrid, int n_pre_adaptive,
                     double cycle, cudaStream_t st) {
    if (n <= 0) return ABC_OK;
    int threads = 128;
    long long blocks = (n + threads - 1) / threads;
    abc_rates_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_theta, m, (long long)n, d_rates, ssa_hybrid, n_pre_adaptive,
                                                          (float)cycle);
    ABC_CUDA_CHECK(cudaGetLastError());
    return ABC_OK;
}

// ------------------------------------------------------------------------------------------------
// P1: fix_params (abc_simulation.jl:3-11): log10 kon,koff,alpha ~ U(-3,3), gamma ~ U(-3,2), lambda ~ U(-0.7,0)
__global__ void abc_prior_kernel(double* __restrict__ theta, int m, long long n, long long offset,
                                 uint32_t k0, uint32_t k1) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int P = (m <= 2) ? 5 : 9;
    const int vary = (m == 3) ? 0 : (m == 4) ? 2 : (m == 5) ? 3 : -1;
    const unsigned long long gp = (unsigned long long)(offset + i);
    const uint32_t tag = abc_tag(0u, 0u, (uint32_t)m, ABC_DOM_PRIOR);
    double lo[ABC_MAXP], hi[ABC_MAXP];
    int k = 0;
    for (int q = 0; q < 4; ++q) {
        int len = (q == vary) ? 5 : 1;
        for (int j = 0; j < len; ++j) { lo[k] = -3.0; hi[k] = (q == 3) ? 2.0 : 3.0; k++; }
    }
    lo[k] = -0.7; hi[k] = 0.0;
    for (int b = 0; b * 2 < P; ++b) {
        uint4 w = philox4x32_10((uint32_t)b, (uint32_t)gp, (uint32_t)(gp >> 32), tag, k0, k1);
        // 53-bit uniforms in [0,1)
        double u0 = (double)(((unsigned long long)(w.x >> 5) << 26) | (unsigned long long)(w.y >> 6)) * (1.0 / 9007199254740992.0);
        double u1 = (double)(((unsigned long long)(w.z >> 5) << 26) | (unsigned long long)(w.w >> 6)) * (1.0 / 9007199254740992.0);
        int j0 = 2 * b, j1 = 2 * b + 1;
        theta[i * P + j0] = abc_fma(hi[j0] - lo[j0], u0, lo[j0]);
        if (j1 < P) theta[i * P + j1] = abc_fma(hi[j1] - lo[j1], u1, lo[j1]);
    }
}

int abc_launch_prior(double* d_theta, int m, int64_t n, int64_t offset, uint64_t seed, cudaStream_t st) {
    if (n <= 0) return ABC_OK;
    int threads = 128;
    long long blocks = (n + threads - 1) / threads;
    abc_prior_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_theta, m, (long long)n, (long long)offset,
                                                          (uint32_t)seed, (uint32_t)(seed >> 32));
    ABC_CUDA_CHECK(cudaGetLastError());
    return ABC_OK;
}

// ------------------------------------------------------------------------------------------------
// integer sums -> sample moments per (particle, read-out): mean_u, mean_l, var_u, cov_ul, var_l
// (corrected, n-1, like var()/cov() of scripts/data_summary_statistics.jl:117-121)
__device__ __forceinline__ double u128_to_double(unsigned __int128 v) {
    unsigned long long hi = (unsigned long long)(v >> 64), lo = (unsigned long long)v;
    return __dadd_rn(__dmul_rn((double)hi, 18446744073709551616.0), (double)lo);
}

__global__ void abc_moments_kernel(const unsigned long long* __restrict__ sums, long long n_items, int n_cells,
                                   double* __restrict__ mom) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const unsigned long long* s = sums + i * 5;
    const unsigned long long su = s[0], sl = s[1], suu = s[2], sul = s[3], sll = s[4];
    if (su == ~0ull) {          // read-out of a refused particle (abc_window_kernel): NaN moments
        double* o = mom + i * 5;
        o[0] = o[1] = o[2] = o[3] = o[4] = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    const unsigned __int128 N = (unsigned __int128)(unsigned long long)n_cells;
    const double dn = (double)n_cells;
    const double dnn = __dmul_rn(dn, (double)(n_cells - 1));
    unsigned __int128 a, b;
    double* o = mom + i * 5;
    o[0] = __ddiv_rn((double)su, dn);
    o[1] = __ddiv_rn((double)sl, dn);
    a = N * suu; b = (unsigned __int128)su * su;
    o[2] = __ddiv_rn(u128_to_double(a - b), dnn);
    a = N * sul; b = (unsigned __int128)su * sl;
    o[3] = (a >= b) ? __ddiv_rn(u128_to_double(a - b), dnn) : -__ddiv_rn(u128_to_double(b - a), dnn);
    a = N * sll; b = (unsigned __int128)sl * sl;
    o[4] = __ddiv_rn(u128_to_double(a - b), dnn);
}

int abc_launch_moments_from_sums(const unsigned long long* d_sums, int64_t n, int n_cells, double* d_moments,
                                 cudaStream_t st) {
    long long items = (long long)n * ABC_NREAD;
    if (items <= 0) return ABC_OK;
    int threads = 128;
    long long blocks = (items + threads - 1) / threads;
    abc_moments_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_sums, items, n_cells, d_moments);
    ABC_CUDA_CHECK(cudaGetLastError());
    return ABC_OK;
}

// ------------------------------------------------------------------------------------------------
// S1: the 53 summary statistics of abc_sim (abc_simulation.jl:23-46) with weighted_cov
// (data_summary_statistics.jl:179-181).  FP64, explicit rounding, the reference's operation order.
__device__ __forceinline__ double d_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double d_mul(double a, double b) { return __dmul_rn(a, b); }

__device__ __forceinline__ double wsum5(const double* w, const double* x) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) s = d_add(s, d_mul(w[i], x[i]));
    return s;
}
__device__ __forceinline__ double wcov5(const double* x, const double* y, const double* w) {
    const double mx = wsum5(w, x), my = wsum5(w, y);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) s = d_add(s, d_mul(w[i], d_mul(d_add(x[i], -mx), d_add(y[i], -my))));
    return s;
}

// sample_guards: the moments come from a finite sample of cells (SSA).  Degenerate samples are then treated as
// the reference treats the real cells (scripts/data_summary_statistics.jl:64-71, 138-147): ratio = 0 when no
// molecule was counted, both correlations = 0 when a total variance vanishes (and mean_corr = 0 when all
// covariances vanish); the moment-ODE path never meets these cases and follows abc_simulation.jl:23-46 verbatim.
__global__ void abc_stats_kernel(const double* __restrict__ mom, const double* __restrict__ age_dist, long long n,
                                 double* __restrict__ stats, int sample_guards) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* mp = mom + i * (ABC_NREAD * 5);
    double* st = stats + i * ABC_NSTATS;
    for (int j = 0; j < ABC_NCOND; ++j) {
        double w[5], c[5][5];
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            w[a] = age_dist[j * 5 + a];
#pragma unroll
            for (int q = 0; q < 5; ++q) c[q][a] = mp[(j * 5 + a) * 5 + q];
        }
        const double s1 = wsum5(w, c[0]), s2 = wsum5(w, c[1]);
        const double v1 = d_add(wsum5(w, c[2]), wcov5(c[0], c[0], w));
        const double v2 = d_add(wsum5(w, c[4]), wcov5(c[1], c[1], w));
        const double stds = __dsqrt_rn(fabs(d_mul(v1, v2)));
        double ratio = __ddiv_rn(s2, d_add(s1, s2));
        double mcorr = __ddiv_rn(wsum5(w, c[3]), stds);
        double cmean = __ddiv_rn(wcov5(c[0], c[1], w), stds);
        if (sample_guards) {
            bool cov_all_zero = true;
#pragma unroll
            for (int a = 0; a < 5; ++a) cov_all_zero = cov_all_zero && (c[3][a] == 0.0);
            if (!(d_add(s1, s2) > 0.0)) ratio = 0.0;
            if (v1 == 0.0 || v2 == 0.0) { mcorr = 0.0; cmean = 0.0; }
            if (cov_all_zero) mcorr = 0.0;
        }
        st[20 + j] = ratio;
        st[31 + j] = mcorr;
        st[42 + j] = cmean;
        if (j == 5 || j == 6) {
            double* mo = st + (j == 5 ? 0 : 10);
            double* ff = st + (j == 5 ? 5 : 15);
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                const double tot = d_add(c[0][a], c[1][a]);
                const double eps = (tot == 0.0) ? 0.0001 : 0.0;
                mo[a] = tot;
                ff[a] = __ddiv_rn(d_add(d_add(c[2][a], d_mul(2.0, c[3][a])), c[4][a]), d_add(tot, eps));
            }
        }
    }
}

int abc_launch_summary_stats(const double* d_moments, const double* d_age_dist, int64_t n, double* d_stats,
                             int sample_guards, cudaStream_t st) {
    if (n <= 0) return ABC_OK;
    int threads = 64;
    long long blocks = (n + threads - 1) / threads;
    abc_stats_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_moments, d_age_dist, (long long)n, d_stats, sample_guards);
    ABC_CUDA_CHECK(cudaGetLastError());
    return ABC_OK;
}
