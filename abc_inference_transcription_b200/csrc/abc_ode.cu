// abc_ode.cu -- device moment-ODE simulator: what scripts/model.jl actually integrates.  sm_100a, FP64.
//
// Reference call stack: abc_sim -> get_steady_state_iv -> transient_phase (model.jl:114-142),
// then per condition get_synthetic_data -> syntheticdata -> trajectories (model.jl:146-187): 4 integrations
// [-60,-40],[-40,-20],[-20,0],[0,age] of the 9 moment equations f (model.jl:74-86) with periodic_boundary
// (model.jl:98-111) in between, then downsample (model.jl:221-239).
//
// The reference hands the RHS to Sundials CVODE_BDF (reltol 1e-3).  Here each piece between two
// discontinuities of the RHS (rate step every cycle/5, label on/off, division) is integrated with an
// adaptive 3-stage Radau IIA collocation method (order 5, L-stable; step-doubling error control).  The
// system is linear and lower triangular in the order (y1,y4,y2,y3,y5,y6,y7,y8,y9) with constant diagonal on
// a piece, so the implicit stage equations reduce to nine 3x3 solves by forward substitution.
//
// Kernel 1: one thread per particle  -> cyclo-stationary post-division state ss_iv (transient_phase).
// Kernel 2: one thread per particle -> the unlabelled path shared by all read-outs, recorded at the 55 window starts.
// Kernel 3: one thread per (particle, condition, age) -> label window + chase from the recorded state (+ downsample).
#include "abc_common.cuh"
#include "abc_internal.h"

struct OdeRates {
    double kon[5], koff[5], alpha[5], gamma[5];
    double lam;
};

struct OdePiece {
    double kon, koff, gam, lam, a_step, cyc_start, inv_cycle;
};

__device__ __forceinline__ void ode_make_rates(const double* __restrict__ th, int m, OdeRates& r) {
    const int vary = (m == 3) ? 0 : (m == 4) ? 2 : (m == 5) ? 3 : -1;   // abc_simulation.jl:83
    int k = 0;
    double* dst[4] = {r.kon, r.koff, r.alpha, r.gamma};
    for (int q = 0; q < 4; ++q) {
        if (q == vary) {
            for (int j = 0; j < 5; ++j) dst[q][j] = abc_exp10_det(th[k + j]);
            k += 5;
        } else {
            const double v = abc_exp10_det(th[k]);
            for (int j = 0; j < 5; ++j) dst[q][j] = v;
            k += 1;
        }
    }
    r.lam = abc_exp10_det(th[k]);                                        // labelling(): 10^lambda, model.jl:60
}

// solve (I - mu*A) z = h * A * w for the Radau IIA matrix A (3x3), mu = h*d
__device__ __forceinline__ void radau_solve3(double mu, double h, const double w[3], double z[3]) {
    const double S6 = 2.449489742783178;
    const double a00 = (88.0 - 7.0 * S6) / 360.0, a01 = (296.0 - 169.0 * S6) / 1800.0, a02 = (-2.0 + 3.0 * S6) / 225.0;
    const double a10 = (296.0 + 169.0 * S6) / 1800.0, a11 = (88.0 + 7.0 * S6) / 360.0, a12 = (-2.0 - 3.0 * S6) / 225.0;
    const double a20 = (16.0 - S6) / 36.0, a21 = (16.0 + S6) / 36.0, a22 = 1.0 / 9.0;
    const double r0 = h * fma(a00, w[0], fma(a01, w[1], a02 * w[2]));
    const double r1 = h * fma(a10, w[0], fma(a11, w[1], a12 * w[2]));
    const double r2 = h * fma(a20, w[0], fma(a21, w[1], a22 * w[2]));
    const double m00 = 1.0 - mu * a00, m01 = -mu * a01, m02 = -mu * a02;
    const double m10 = -mu * a10, m11 = 1.0 - mu * a11, m12 = -mu * a12;
    const double m20 = -mu * a20, m21 = -mu * a21, m22 = 1.0 - mu * a22;
    const double c00 = m11 * m22 - m12 * m21, c01 = m12 * m20 - m10 * m22, c02 = m10 * m21 - m11 * m20;
    const double det = fma(m00, c00, fma(m01, c01, m02 * c02));
    const double inv = 1.0 / det;
    const double c10 = m02 * m21 - m01 * m22, c11 = m00 * m22 - m02 * m20, c12 = m01 * m20 - m00 * m21;
    const double c20 = m01 * m12 - m02 * m11, c21 = m02 * m10 - m00 * m12, c22 = m00 * m11 - m01 * m10;
    z[0] = inv * fma(c00, r0, fma(c10, r1, c20 * r2));
    z[1] = inv * fma(c01, r0, fma(c11, r1, c21 * r2));
    z[2] = inv * fma(c02, r0, fma(c12, r1, c22 * r2));
}

// one Radau IIA step of the moment equations (model.jl:74-86) on a piece with constant kon, koff, gamma, lam
__device__ void radau_step(const OdePiece& pc, double t, double h, const double* __restrict__ y, double* __restrict__ yn) {
    const double S6 = 2.449489742783178;
    const double cj[3] = {(4.0 - S6) / 10.0, (4.0 + S6) / 10.0, 1.0};
    double al[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) al[j] = pc.a_step * fma((t + cj[j] * h) - pc.cyc_start, pc.inv_cycle, 1.0);
    const double kon = pc.kon, koff = pc.koff, g = pc.gam, l = pc.lam, s = kon + koff, ul = 1.0 - l;
    double w[3], z0[3], z1[3], z2[3], z3[3], z4[3], z5[3], z[3];
    // y1 = E[g]:            dy = kon - s*y1
#pragma unroll
    for (int j = 0; j < 3; ++j) w[j] = fma(-s, y[0], kon);
    radau_solve3(-s * h, h, w, z0);
    // y4 = Var g:           dy = kon + (koff-kon)*y1 - 2s*y4
#pragma unroll
    for (int j = 0; j < 3; ++j) w[j] = fma(-2.0 * s, y[3], fma(koff - kon, y[0] + z0[j], kon));
    radau_solve3(-2.0 * s * h, h, w, z3);
    // y2 = E[U], y3 = E[L]: dy = (1-l)|l * alpha*y1 - g*y
#pragma unroll
    for (int j = 0; j < 3; ++j) w[j] = fma(-g, y[1], ul * al[j] * (y[0] + z0[j]));
    radau_solve3(-g * h, h, w, z1);
#pragma unroll
    for (int j = 0; j < 3; ++j) w[j] = fma(-g, y[2], l * al[j] * (y[0] + z0[j]));
    radau_solve3(-g * h, h, w, z2);
    // y5 = Cov(g,U), y6 = Cov(g,L): dy = alpha*(1-l)|l * y4 - (s+g)*y
#pragma unroll
    for (int j = 0; j < 3; ++j) w[j] = fma(-(s + g), y[4], al[j] * ul * (y[3] + z3[j]));
    radau_solve3(-(s + g) * h, h, w, z4);
#pragma unroll
    for (int j = 0; j < 3; ++j) w[j] = fma(-(s + g), y[5], al[j] * l * (y[3] + z3[j]));
    radau_solve3(-(s + g) * h, h, w, z5);
    yn[0] = y[0] + z0[2]; yn[3] = y[3] + z3[2]; yn[1] = y[1] + z1[2]; yn[2] = y[2] + z2[2];
    yn[4] = y[4] + z4[2]; yn[5] = y[5] + z5[2];
    // y7 = Var U:  dy = alpha(1-l) y1 + g y2 + 2 alpha (1-l) y5 - 2g y7
#pragma unroll
    for (int j = 0; j < 3; ++j)
        w[j] = fma(-2.0 * g, y[6], fma(al[j] * ul, (y[0] + z0[j]) + 2.0 * (y[4] + z4[j]), g * (y[1] + z1[j])));
    radau_solve3(-2.0 * g * h, h, w, z);
    yn[6] = y[6] + z[2];
    // y8 = Cov(U,L): dy = alpha l y5 + alpha (1-l) y6 - 2g y8
#pragma unroll
    for (int j = 0; j < 3; ++j)
        w[j] = fma(-2.0 * g, y[7], al[j] * fma(l, y[4] + z4[j], ul * (y[5] + z5[j])));
    radau_solve3(-2.0 * g * h, h, w, z);
    yn[7] = y[7] + z[2];
    // y9 = Var L:  dy = alpha l y1 + g y3 + 2 alpha l y6 - 2g y9
#pragma unroll
    for (int j = 0; j < 3; ++j)
        w[j] = fma(-2.0 * g, y[8], fma(al[j] * l, (y[0] + z0[j]) + 2.0 * (y[5] + z5[j]), g * (y[2] + z2[j])));
    radau_solve3(-2.0 * g * h, h, w, z);
    yn[8] = y[8] + z[2];
}

__device__ unsigned int integrate_piece(const OdePiece& pc, double ta, double tb, double* y, double rtol, double atol) {
    unsigned int nsteps = 0;
    double t = ta, h = (tb - ta) * 0.25;
    while (t < tb) {
        bool last = false;
        if (t + h >= tb || (tb - (t + h)) < 1e-10 * (tb - ta)) { h = tb - t; last = true; }
        double y1[9], y2[9], yh[9];
        radau_step(pc, t, h, y, y1);
        radau_step(pc, t, 0.5 * h, y, yh);
        radau_step(pc, t + 0.5 * h, 0.5 * h, yh, y2);
        double err = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const double sc = atol + rtol * fmax(fabs(y[i]), fabs(y2[i]));
            err = fmax(err, fabs(y2[i] - y1[i]) / (31.0 * sc));
        }
        const bool ok = (err <= 1.0) || (h < 1e-13 * fmax(1.0, fabs(t)));
        if (!(err == err)) {             // NaN (non-finite parameters): give up on this piece, propagate NaN
#pragma unroll
            for (int i = 0; i < 9; ++i) y[i] = err;
            return nsteps;
        }
        if (ok) {
#pragma unroll
            for (int i = 0; i < 9; ++i) y[i] = y2[i] + (y2[i] - y1[i]) / 31.0;
            t = last ? tb : t + h;
            nsteps += 1;
        }
        double fac = (err > 0.0) ? 0.9 * exp2(-log2(err) / 6.0) : 4.0;
        fac = fmin(4.0, fmax(0.2, fac));
        h *= fac;
        if (nsteps > 2000000u && t < tb) {      // step limit: a truncated integration must never pass for a result
#pragma unroll
            for (int i = 0; i < 9; ++i) y[i] = __longlong_as_double(0x7ff8000000000000ll);
            return nsteps;
        }
    }
    return nsteps;
}

// model(): integrate from tmin to tmax with stops at every RHS discontinuity (model.jl:89-96 + :1-27, :58-64)
__device__ unsigned int ode_model(const OdeRates& r, int scaling, double* y, double tmin, double tmax, double cycle,
                                  double texp, double pulse, double rtol, double atol) {
    const double step_len = cycle / 5.0, tl1 = texp + pulse;
    unsigned int nsteps = 0;
    double pos = tmin;
    while (pos < tmax) {
        double nxt = (floor(pos / step_len + 1e-9) + 1.0) * step_len;       // next multiple of cycle/5 after pos
        if (nxt > tmax) nxt = tmax;
        if (texp > pos && texp < nxt) nxt = texp;
        if (tl1 > pos && tl1 < nxt) nxt = tl1;
        const double mid = 0.5 * (pos + nxt);
        const double cyc_start = cycle * floor(mid / cycle);
        int j = (int)floor((mid - cyc_start) / step_len);
        j = j < 0 ? 0 : (j > 4 ? 4 : j);
        OdePiece pc;
        pc.kon = r.kon[j]; pc.koff = r.koff[j]; pc.gam = r.gamma[j]; pc.a_step = r.alpha[j];
        pc.lam = (mid >= texp && mid <= tl1) ? r.lam : 0.0;
        pc.cyc_start = cyc_start;
        pc.inv_cycle = scaling ? 1.0 / cycle : 0.0;
        nsteps += integrate_piece(pc, pos, nxt, y, rtol, atol);
        pos = nxt;
    }
    return nsteps;
}

// periodic_boundary (model.jl:98-111)
__device__ __forceinline__ void ode_divide(double* y) {
    y[4] *= 0.5; y[5] *= 0.5;
    y[6] = y[6] * 0.25 + y[1] * 0.25;
    y[8] = y[8] * 0.25 + y[2] * 0.25;
    y[7] *= 0.25;
    y[1] *= 0.5; y[2] *= 0.5;
}

struct AbcOdeParams {
    long long n;
    int m, scaling, downsampling;
    double cycle, t0, rtol, atol;
    double agevec[5], pulse[11], chase[11], iv[9];
    int order[ABC_NREAD];     // read-outs sorted by their label-window start
};

// transient_phase (model.jl:114-142): one thread per particle
__global__ void abc_ode_transient_kernel(const double* __restrict__ theta, const AbcOdeParams prm, double* __restrict__ ss_iv,
                                         unsigned long long* __restrict__ counters) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.n) return;
    const int P = (prm.m <= 2) ? 5 : 9;
    OdeRates r;
    ode_make_rates(theta + i * P, prm.m, r);
    double e1[9], e2[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) e1[k] = prm.iv[k];
    unsigned int steps = ode_model(r, prm.scaling, e1, 0.0, prm.cycle, prm.cycle, -1.0, 0.1, prm.rtol, prm.atol);
#pragma unroll
    for (int k = 0; k < 9; ++k) e2[k] = e1[k];
    int it = 0;
    bool conv = false;
    while (!conv && it <= 100) {
        it += 1;
#pragma unroll
        for (int k = 0; k < 9; ++k) e2[k] = e1[k];
        ode_divide(e2);
        steps += ode_model(r, prm.scaling, e2, 0.0, prm.cycle, prm.cycle, -1.0, 0.1, prm.rtol, prm.atol);
        const int chk[5] = {0, 1, 3, 4, 6};
        bool all_ok = true;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const double a = e1[chk[q]], b = e2[chk[q]];
            if (a > 0.0 && !(fabs((a - b) / a) <= 0.01)) all_ok = false;
        }
        conv = all_ok;
#pragma unroll
        for (int k = 0; k < 9; ++k) e1[k] = e2[k];
    }
    ode_divide(e2);
#pragma unroll
    for (int k = 0; k < 9; ++k) ss_iv[i * 9 + k] = e2[k];
    atomicAdd(counters + 4, (unsigned long long)steps);
}

// trajectories (model.jl:146-174), shared prefix.  Until its label window opens, every (condition, age) read-out
// of a particle follows the same unlabelled path (lambda = 0: the labelled moments stay 0) through the solves
// [t0,t0+cycle], ..., [-cycle,0], [0,age]; the state at time t < age does not depend on age.  One thread per
// particle integrates that path once and records the state at each of the 55 window starts texp = age-pulse-chase
// (visited in increasing order, prm.order).
__global__ void abc_ode_prefix_kernel(const double* __restrict__ theta, const double* __restrict__ ss_iv,
                                      const AbcOdeParams prm, double* __restrict__ prefix,
                                      unsigned long long* __restrict__ counters) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.n) return;
    const int P = (prm.m <= 2) ? 5 : 9;
    OdeRates r;
    ode_make_rates(theta + i * P, prm.m, r);
    double y[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) y[k] = ss_iv[i * 9 + k];
    unsigned int steps = 0;
    double pos = prm.t0;
    int next = 0;                                           // next read-out (in texp order) to record
    while (next < ABC_NREAD) {
        const int ro = prm.order[next];
        const double texp = prm.agevec[ro % ABC_NAGE] - prm.pulse[ro / ABC_NAGE] - prm.chase[ro / ABC_NAGE];
        // advance to texp, applying periodic_boundary at every multiple of the cycle that is crossed
        while (pos < texp) {
            const double nb = prm.cycle * (floor(pos / prm.cycle + 1e-12) + 1.0);
            const double stop = fmin(nb, texp);
            steps += ode_model(r, prm.scaling, y, pos, stop, prm.cycle, -1e30, 0.0, prm.rtol, prm.atol);
            pos = stop;
            if (pos == nb && pos <= texp) ode_divide(y);   // the next solve starts from the divided state
        }
        double* o = prefix + (i * ABC_NREAD + ro) * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) o[k] = y[k];
        next += 1;
    }
    atomicAdd(counters + 4, (unsigned long long)steps);
}

// label window + chase: one thread per (particle, read-out), from the recorded prefix state at texp to the read-out
// age; then downsample (model.jl:221-239)
__global__ void abc_ode_readout_kernel(const double* __restrict__ theta, const double* __restrict__ prefix,
                                       const AbcOdeParams prm, const double* __restrict__ beta_mom,
                                       double* __restrict__ mom, unsigned long long* __restrict__ counters) {
    // read-out major: the 32 lanes of a warp integrate the SAME read-out (same pieces, same breakpoints, same loop
    // structure) for 32 different particles, so they diverge only where the step-size control does; particle-major
    // (32 read-outs of one particle per warp, windows of 0.25 h to 28 h side by side) ran at 13 of 32 active threads
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= prm.n * ABC_NREAD) return;
    const int ro = (int)(tid / prm.n), cond = ro / ABC_NAGE, a = ro % ABC_NAGE;
    const long long i = tid % prm.n;
    const long long idx = i * ABC_NREAD + ro;
    const int P = (prm.m <= 2) ? 5 : 9;
    OdeRates r;
    ode_make_rates(theta + i * P, prm.m, r);
    const double age = prm.agevec[a], pulse = prm.pulse[cond], texp = age - pulse - prm.chase[cond];
    double y[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) y[k] = prefix[idx * 9 + k];
    unsigned int steps = 0;
    double pos = texp;
    while (pos < age) {
        const double nb = prm.cycle * (floor(pos / prm.cycle + 1e-12) + 1.0);
        const double stop = fmin(nb, age);
        steps += ode_model(r, prm.scaling, y, pos, stop, prm.cycle, texp, pulse, prm.rtol, prm.atol);
        pos = stop;
        if (pos == nb && pos < age) ode_divide(y);
    }
    double mu = y[1], ml = y[2], vu = y[6], cv = y[7], vl = y[8];
    if (prm.downsampling) {
        const int grp = (cond < 6 ? 0 : ABC_NAGE) + a;
        const double bm = beta_mom[grp], b2 = beta_mom[10 + grp], bv = beta_mom[20 + grp];
        const double nvu = ((bm - b2) * mu + bv * (mu * mu + vu)) + (bm * bm) * vu;
        const double nvl = ((bm - b2) * ml + bv * (ml * ml + vl)) + (bm * bm) * vl;
        const double ncv = bv * (mu * ml + cv) + (bm * bm) * cv;
        vu = nvu; vl = nvl; cv = ncv;
        mu = mu * bm; ml = ml * bm;
    }
    double* o = mom + idx * 5;
    o[0] = mu; o[1] = ml; o[2] = vu; o[3] = cv; o[4] = vl;
    atomicAdd(counters + 4, (unsigned long long)steps);
}

int abc_launch_ode(const double* d_theta, const abc_design_t& des, int m, int64_t n, const double* d_beta_mom,
                   double* d_ss_iv, double* d_prefix, double* d_moments, unsigned long long* d_counters, cudaStream_t st) {
    if (n <= 0) return ABC_OK;
    AbcOdeParams prm;
    prm.n = n; prm.m = m; prm.scaling = (m != 2) ? 1 : 0; prm.downsampling = des.downsampling;
    prm.cycle = des.cycle; prm.t0 = des.t0; prm.rtol = des.ode_rtol; prm.atol = des.ode_atol;
    for (int a = 0; a < 5; ++a) prm.agevec[a] = des.agevec[a];
    for (int j = 0; j < 11; ++j) { prm.pulse[j] = des.pulse[j]; prm.chase[j] = des.chase[j]; }
    for (int k = 0; k < 9; ++k) prm.iv[k] = des.iv[k];
    // read-outs by increasing window start
    double texp[ABC_NREAD];
    for (int ro = 0; ro < ABC_NREAD; ++ro) {
        prm.order[ro] = ro;
        texp[ro] = des.agevec[ro % ABC_NAGE] - des.pulse[ro / ABC_NAGE] - des.chase[ro / ABC_NAGE];
        if (texp[ro] < des.t0) { abc_set_error("label window of condition %d starts before t0", ro / ABC_NAGE + 1); return ABC_ERR_ARG; }
    }
    for (int x = 1; x < ABC_NREAD; ++x)
        for (int yy = x; yy > 0 && texp[prm.order[yy]] < texp[prm.order[yy - 1]]; --yy) {
            const int t = prm.order[yy]; prm.order[yy] = prm.order[yy - 1]; prm.order[yy - 1] = t;
        }
    const int threads = 64;
    // one thread per particle: single-warp blocks spread a small batch over all SMs (8192 particles = 256 warps for 592 schedulers)
    abc_ode_transient_kernel<<<(unsigned)((n + 31) / 32), 32, 0, st>>>(d_theta, prm, d_ss_iv, d_counters);
    ABC_CUDA_CHECK(cudaGetLastError());
    abc_ode_prefix_kernel<<<(unsigned)((n + 31) / 32), 32, 0, st>>>(d_theta, d_ss_iv, prm, d_prefix, d_counters);
    ABC_CUDA_CHECK(cudaGetLastError());
    const long long items = (long long)n * ABC_NREAD;
    abc_ode_readout_kernel<<<(unsigned)((items + threads - 1) / threads), threads, 0, st>>>(d_theta, d_prefix, prm, d_beta_mom,
                                                                                          d_moments, d_counters);
    ABC_CUDA_CHECK(cudaGetLastError());
    return ABC_OK;
}
