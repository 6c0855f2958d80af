/*
 * abc_oracle.c -- CPU restatement (TEST ORACLE) of the reference's moment-ODE simulator,
 * summary statistics, error scoring and acceptance.  See abc_oracle.h for the rules on who
 * may use this file.  Compile with -O2 -ffp-contract=off (no FMA contraction: the scoring
 * and statistics functions define the bit-exact operation order).
 *
 * The reference integrates the 9 moment ODEs with Sundials CVODE_BDF (third-party, unpinned,
 * absent from /root/reference; call site scripts/model.jl:94).  CVODE is not restated; the ODE
 * system it is applied to IS (model.jl:74-86), and it is integrated here with an adaptive
 * 3-stage Radau IIA collocation method (order 5, L-stable) with step-doubling error control
 * and exact stops at every discontinuity of the right-hand side.  Because the system is
 * linear and lower-triangular in the order (y1,y4,y2,y3,y5,y6,y7,y8,y9), the implicit stage
 * equations are solved exactly by forward substitution with 3x3 solves.
 */
#include "abc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ model.jl:30-43 + abc_simulation.jl:82-85 */
static const int VARY_FLAG[5][4] = {{0,0,0,0},{0,0,0,0},{1,0,0,0},{0,0,1,0},{0,0,0,1}};

int orc_n_params(int m) { return (m <= 2) ? 5 : 9; }

/* vary_map: first theta index (0-based) of each of kon,koff,alpha,gamma and its length */
static void vary_map(int m, int start[4], int len[4]) {
    int k = 0;
    for (int i = 0; i < 4; ++i) {
        start[i] = k;
        len[i] = VARY_FLAG[m - 1][i] ? ORC_NAGE : 1;
        k += len[i];
    }
}

static double jl_mod(double t, double c) { /* Julia mod(): result has the sign of c */
    double r = fmod(t, c);
    if (r != 0.0 && ((r < 0.0) != (c < 0.0))) r += c;
    return r;
}

/* model.jl:25-27 */
double orc_size_scaling(double cycle, double t) {
    return 1.0 + (jl_mod(t, cycle) / cycle) * (t < cycle ? 1.0 : 0.0) + (t == cycle ? 1.0 : 0.0);
}

/* model.jl:1-22.  t_steps = 0:cycle/5:cycle */
void orc_get_rate(const double* theta, int m, double cycle, double t, double p[4]) {
    int start[4], len[4];
    int scaling = (m != 2);
    vary_map(m, start, len);
    for (int i = 0; i < 4; ++i) {
        double txn = (i == 2) ? scaling * log10(orc_size_scaling(cycle, t)) : 0.0;
        if (len[i] == 1) {
            p[i] = theta[start[i]] + txn;
        } else {
            double mt = jl_mod(t, cycle);
            p[i] = NAN;
            for (int j = 0; j < ORC_NAGE; ++j) {
                double lo = j * (cycle / ORC_NAGE), hi = (j + 1) * (cycle / ORC_NAGE);
                if (mt > lo && mt <= hi) p[i] = theta[start[i] + j] + txn;
                else if (mt == 0.0) p[i] = theta[start[i]] + txn;
            }
        }
    }
}

/* model.jl:58-64 */
double orc_labelling(double lambda, double texp, double pulse, double t) {
    return (t >= texp && t <= texp + pulse) ? pow(10.0, lambda) : 0.0;
}

/* model.jl:74-86 */
void orc_f(const double y[9], const double p[4], double l, double dy[9]) {
    dy[0] = p[0]*(1-y[0]) - p[1]*y[0];
    dy[1] = (1-l)*p[2]*y[0] - p[3]*y[1];
    dy[2] = l*p[2]*y[0] - p[3]*y[2];
    dy[3] = p[1]*y[0] + p[0]*(1-y[0]) - 2*(p[0]+p[1])*y[3];
    dy[4] = p[2]*(1-l)*y[3] - (p[0]+p[1]+p[3])*y[4];
    dy[5] = p[2]*l*y[3] - (p[0]+p[1]+p[3])*y[5];
    dy[6] = p[2]*(1-l)*y[0] + p[3]*y[1] + 2*p[2]*(1-l)*y[4] - 2*p[3]*y[6];
    dy[7] = p[2]*l*y[4] + p[2]*(1-l)*y[5] - 2*p[3]*y[7];
    dy[8] = p[2]*l*y[0] + p[3]*y[2] + 2*p[2]*l*y[5] - 2*p[3]*y[8];
}

/* ------------------------------------------------------------------ integrator */
typedef struct {
    double kon, koff, gam, lam; /* constant on the piece */
    double a_step;              /* 10^theta_alpha of the rate step */
    double cyc_start, cycle;
    int scaling;
} piece_t;

static double piece_alpha(const piece_t* pc, double t) {
    return pc->a_step * (1.0 + (pc->scaling ? (t - pc->cyc_start) / pc->cycle : 0.0));
}

/* A(t) y + b(t) of model.jl:74-86 written as a matrix (row-major 9x9) */
static void piece_Ab(const piece_t* pc, double t, double A[81], double b[9]) {
    double kon = pc->kon, koff = pc->koff, g = pc->gam, l = pc->lam, al = piece_alpha(pc, t);
    double s = kon + koff;
    memset(A, 0, 81 * sizeof(double));
    memset(b, 0, 9 * sizeof(double));
#define AA(i,j) A[(i)*9+(j)]
    AA(0,0) = -s;            b[0] = kon;
    AA(1,0) = (1-l)*al;      AA(1,1) = -g;
    AA(2,0) = l*al;          AA(2,2) = -g;
    AA(3,0) = koff - kon;    AA(3,3) = -2*s;   b[3] = kon;
    AA(4,3) = al*(1-l);      AA(4,4) = -(s+g);
    AA(5,3) = al*l;          AA(5,5) = -(s+g);
    AA(6,0) = al*(1-l);      AA(6,1) = g;      AA(6,4) = 2*al*(1-l);  AA(6,6) = -2*g;
    AA(7,4) = al*l;          AA(7,5) = al*(1-l);                      AA(7,7) = -2*g;
    AA(8,0) = al*l;          AA(8,2) = g;      AA(8,5) = 2*al*l;      AA(8,8) = -2*g;
#undef AA
}

static const int ORD[9] = {0, 3, 1, 2, 4, 5, 6, 7, 8}; /* triangular order y1,y4,y2,y3,y5..y9 */

static void solve3(double M[9], double r[3]) { /* Gaussian elimination, partial pivoting */
    for (int c = 0; c < 3; ++c) {
        int piv = c;
        for (int i = c + 1; i < 3; ++i) if (fabs(M[i*3+c]) > fabs(M[piv*3+c])) piv = i;
        if (piv != c) {
            for (int j = 0; j < 3; ++j) { double t = M[c*3+j]; M[c*3+j] = M[piv*3+j]; M[piv*3+j] = t; }
            double t = r[c]; r[c] = r[piv]; r[piv] = t;
        }
        for (int i = c + 1; i < 3; ++i) {
            double f = M[i*3+c] / M[c*3+c];
            for (int j = c; j < 3; ++j) M[i*3+j] -= f * M[c*3+j];
            r[i] -= f * r[c];
        }
    }
    for (int i = 2; i >= 0; --i) {
        double s = r[i];
        for (int j = i + 1; j < 3; ++j) s -= M[i*3+j] * r[j];
        r[i] = s / M[i*3+i];
    }
}

/* one Radau IIA (3-stage, order 5) step of size h from (t, y) */
static void radau_step(const piece_t* pc, double t, double h, const double y[9], double ynew[9]) {
    static const double S6 = 2.449489742783178; /* sqrt(6) */
    const double c[3] = {(4.0 - S6) / 10.0, (4.0 + S6) / 10.0, 1.0};
    const double a[3][3] = {
        {(88.0 - 7.0*S6) / 360.0, (296.0 - 169.0*S6) / 1800.0, (-2.0 + 3.0*S6) / 225.0},
        {(296.0 + 169.0*S6) / 1800.0, (88.0 + 7.0*S6) / 360.0, (-2.0 - 3.0*S6) / 225.0},
        {(16.0 - S6) / 36.0, (16.0 + S6) / 36.0, 1.0 / 9.0}};
    double A[3][81], b[3][9], Z[3][9];
    for (int j = 0; j < 3; ++j) piece_Ab(pc, t + c[j] * h, A[j], b[j]);
    for (int q = 0; q < 9; ++q) {
        int v = ORD[q];
        double G[3], M[9], r[3];
        for (int j = 0; j < 3; ++j) {
            double s = b[j][v];
            for (int p = 0; p < q; ++p) { int w = ORD[p]; s += A[j][v*9+w] * (y[w] + Z[j][w]); }
            G[j] = s + A[j][v*9+v] * y[v];
        }
        for (int i = 0; i < 3; ++i) {
            double s = 0.0;
            for (int j = 0; j < 3; ++j) {
                M[i*3+j] = (i == j ? 1.0 : 0.0) - h * a[i][j] * A[j][v*9+v];
                s += a[i][j] * G[j];
            }
            r[i] = h * s;
        }
        solve3(M, r);
        for (int i = 0; i < 3; ++i) Z[i][v] = r[i];
    }
    for (int v = 0; v < 9; ++v) ynew[v] = y[v] + Z[2][v];
}

static long integrate_piece(const piece_t* pc, double ta, double tb, double y[9], double rtol, double atol) {
    long nsteps = 0;
    double t = ta, h = (tb - ta) / 4.0;
    while (t < tb) {
        int last = 0;
        if (t + h >= tb || (tb - (t + h)) < 1e-10 * (tb - ta)) { h = tb - t; last = 1; }
        double y1[9], yh[9], y2[9];
        radau_step(pc, t, h, y, y1);
        radau_step(pc, t, 0.5 * h, y, yh);
        radau_step(pc, t + 0.5 * h, 0.5 * h, yh, y2);
        nsteps += 3;
        double err = 0.0;
        for (int i = 0; i < 9; ++i) {
            double sc = atol + rtol * fmax(fabs(y[i]), fabs(y2[i]));
            double e = fabs(y2[i] - y1[i]) / 31.0 / sc;
            if (e > err) err = e;
        }
        if (err <= 1.0 || h < 1e-13 * fmax(1.0, fabs(t))) {
            /* accept the two half steps with local extrapolation */
            for (int i = 0; i < 9; ++i) y[i] = y2[i] + (y2[i] - y1[i]) / 31.0;
            t = last ? tb : t + h;
        }
        double fac = (err > 0.0) ? 0.9 * pow(err, -1.0 / 6.0) : 4.0;
        if (fac > 4.0) fac = 4.0;
        if (fac < 0.2) fac = 0.2;
        h *= fac;
        if (nsteps > 30000000L) break;
    }
    return nsteps;
}

static int cmp_dbl(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* model.jl:89-96 (with momentodes :67-71, get_rate :1-22, labelling :58-64) */
long orc_model(const double* theta, int m, const double iv[9], double tmin, double tmax, double cycle,
               double texp, double pulse, double rtol, double atol, double out[9]) {
    double bp[64];
    int nb = 0;
    int start[4], len[4];
    vary_map(m, start, len);
    int P = orc_n_params(m);
    double step = cycle / ORC_NAGE;
    bp[nb++] = tmin;
    bp[nb++] = tmax;
    for (double k = ceil(tmin / step); k * step < tmax && nb < 60; k += 1.0)
        if (k * step > tmin) bp[nb++] = k * step;
    if (texp > tmin && texp < tmax) bp[nb++] = texp;
    if (texp + pulse > tmin && texp + pulse < tmax) bp[nb++] = texp + pulse;
    qsort(bp, nb, sizeof(double), cmp_dbl);
    double y[9];
    memcpy(y, iv, sizeof(y));
    long nsteps = 0;
    for (int i = 0; i + 1 < nb; ++i) {
        double ta = bp[i], tb = bp[i + 1];
        if (!(tb > ta)) continue;
        double mid = 0.5 * (ta + tb);
        double mt = jl_mod(mid, cycle);
        int j = (int)floor(mt / step);
        if (j > ORC_NAGE - 1) j = ORC_NAGE - 1;
        piece_t pc;
        pc.kon  = pow(10.0, theta[start[0] + (len[0] > 1 ? j : 0)]);
        pc.koff = pow(10.0, theta[start[1] + (len[1] > 1 ? j : 0)]);
        pc.a_step = pow(10.0, theta[start[2] + (len[2] > 1 ? j : 0)]);
        pc.gam  = pow(10.0, theta[start[3] + (len[3] > 1 ? j : 0)]);
        pc.lam  = orc_labelling(theta[P - 1], texp, pulse, mid);
        pc.cycle = cycle;
        pc.cyc_start = cycle * floor(mid / cycle);
        pc.scaling = (m != 2);
        nsteps += integrate_piece(&pc, ta, tb, y, rtol, atol);
    }
    memcpy(out, y, sizeof(y));
    return nsteps;
}

/* model.jl:98-111 */
void orc_periodic_boundary(const double e[9], double v[9]) {
    memcpy(v, e, 9 * sizeof(double));
    v[4] = e[4] / 2; v[5] = e[5] / 2;
    v[6] = e[6] / 4 + e[1] / 4;
    v[8] = e[8] / 4 + e[2] / 4;
    v[7] = e[7] / 4;
    v[1] = e[1] / 2; v[2] = e[2] / 2;
}

/* model.jl:114-142.  Returns the number of while-loop iterations k. */
int orc_transient_phase(const double* theta, int m, const double iv_in[9], double cycle,
                        double rtol, double atol, double ss_iv[9]) {
    const double eps = 0.01, texp = -1.0, pulse = 0.1;
    static const int check_idx[5] = {0, 1, 3, 4, 6};
    double iv[9], e1[9], e2[9];
    int k = 0, convergence = 0;
    const int max_iter = 100;
    memcpy(iv, iv_in, sizeof(iv));
    orc_model(theta, m, iv, 0.0, cycle, cycle, texp, pulse, rtol, atol, e1);
    memcpy(e2, e1, sizeof(e2)); /* reference leaves endpoint_2 undefined if the loop never runs; it always runs */
    while (!convergence && k <= max_iter) {
        k += 1;
        orc_periodic_boundary(e1, iv);
        orc_model(theta, m, iv, 0.0, cycle, cycle, texp, pulse, rtol, atol, e2);
        int nz = 0, ok = 0;
        for (int q = 0; q < 5; ++q) {
            int i = check_idx[q];
            if (e1[i] > 0.0) {
                nz++;
                if (fabs((e1[i] - e2[i]) / e1[i]) <= eps) ok++;
            }
        }
        if (ok == nz) convergence = 1;
        memcpy(e1, e2, sizeof(e1));
    }
    orc_periodic_boundary(e2, ss_iv);
    return k;
}

/* model.jl:146-174 */
void orc_trajectories(const double* theta, int m, const double iv[9], double age, double cycle, double pulse,
                      double t0, double texp, double rtol, double atol, double mean2[2], double cov3[3]) {
    int nsols;
    if (age > 0 && age < cycle) nsols = (int)floor((age - t0) / cycle) + 1;
    else nsols = (int)floor((age - t0) / cycle);
    double tau = t0, endpoint[9], start[9];
    memcpy(endpoint, iv, sizeof(endpoint));
    for (int k = 1; k <= nsols; ++k) {
        double tf = (tau < 0) ? tau + cycle : age;
        if (tau == t0) memcpy(start, iv, sizeof(start));
        else orc_periodic_boundary(endpoint, start);
        orc_model(theta, m, start, tau, tf, cycle, texp, pulse, rtol, atol, endpoint);
        tau = tf;
    }
    mean2[0] = endpoint[1]; mean2[1] = endpoint[2];
    cov3[0] = endpoint[6]; cov3[1] = endpoint[7]; cov3[2] = endpoint[8];
}

/* model.jl:176-187 */
void orc_syntheticdata(const double* theta, int m, const double ss_iv[9], const orc_design_t* d,
                       double pulse, double chase, double s[ORC_NAGE * 5]) {
    for (int i = 0; i < ORC_NAGE; ++i) {
        double age = d->agevec[i];
        double texp = age - pulse - chase;
        orc_trajectories(theta, m, ss_iv, age, d->cycle, pulse, d->t0, texp, d->rtol, d->atol, &s[i*5], &s[i*5+2]);
    }
}

/* model.jl:226-235: per-cluster mean(beta), mean(beta^2), var(beta) (corrected) */
void orc_beta_moments(const double* betas, const int* clusters, int n, double bmean[ORC_NAGE],
                      double bm2[ORC_NAGE], double bvar[ORC_NAGE]) {
    for (int c = 1; c <= ORC_NAGE; ++c) {
        double s = 0.0, s2 = 0.0; long cnt = 0;
        for (int i = 0; i < n; ++i) if (clusters[i] == c) { s += betas[i]; s2 += betas[i] * betas[i]; cnt++; }
        double mu = s / (double)cnt, ss = 0.0;
        for (int i = 0; i < n; ++i) if (clusters[i] == c) ss += (betas[i] - mu) * (betas[i] - mu);
        bmean[c-1] = mu; bm2[c-1] = s2 / (double)cnt; bvar[c-1] = ss / (double)(cnt - 1);
    }
}

/* model.jl:221-239 (columns: 0 mean_u, 1 mean_l, 2 var_u, 3 cov_ul, 4 var_l) */
void orc_downsample(const double s[ORC_NAGE * 5], const double* bm, const double* bm2, const double* bv,
                    double o[ORC_NAGE * 5]) {
    for (int c = 0; c < ORC_NAGE; ++c) {
        const double* r = &s[c*5];
        double mu = r[0], ml = r[1], vu = r[2], cv = r[3], vl = r[4];
        o[c*5+0] = mu * bm[c];
        o[c*5+1] = ml * bm[c];
        o[c*5+2] = ((bm[c] - bm2[c]) * mu + bv[c] * (mu*mu + vu)) + (bm[c]*bm[c]) * vu;
        o[c*5+4] = ((bm[c] - bm2[c]) * ml + bv[c] * (ml*ml + vl)) + (bm[c]*bm[c]) * vl;
        o[c*5+3] = bv[c] * (mu*ml + cv) + (bm[c]*bm[c]) * cv;
    }
}

/* data_summary_statistics.jl:179-181 */
double orc_weighted_cov(const double* x, const double* y, const double* w, int n) {
    double mx = 0.0, my = 0.0, s = 0.0;
    for (int i = 0; i < n; ++i) mx += w[i] * x[i];
    for (int i = 0; i < n; ++i) my += w[i] * y[i];
    for (int i = 0; i < n; ++i) s += w[i] * ((x[i] - mx) * (y[i] - my));
    return s;
}

static double wsum(const double* w, const double* x, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += w[i] * x[i];
    return s;
}

static void summary_stats_impl(const double* mom, const double* age_dist, double* st, int sample_guards);

/* abc_simulation.jl:23-46 */
void orc_summary_stats(const double mom[ORC_NCOND * ORC_NAGE * 5], const double* age_dist, double st[ORC_NSTATS]) {
    summary_stats_impl(mom, age_dist, st, 0);
}

/* the same statistics for moments estimated from a finite sample of cells: degenerate samples are handled as
 * the reference handles the real cells, data_summary_statistics.jl:64-71 (ratio) and :138-147 (correlations) */
void orc_summary_stats_sample(const double mom[ORC_NCOND * ORC_NAGE * 5], const double* age_dist, double st[ORC_NSTATS]) {
    summary_stats_impl(mom, age_dist, st, 1);
}

static void summary_stats_impl(const double* mom, const double* age_dist, double* st, int sample_guards) {
    double* pulse_mean = st, *pulse_ff = st + 5, *chase_mean = st + 10, *chase_ff = st + 15;
    double* ratio = st + 20, *mean_corr = st + 31, *corr_mean = st + 42;
    for (int j = 0; j < ORC_NCOND; ++j) {
        const double* w = &age_dist[j * ORC_NAGE];
        double c[5][ORC_NAGE];
        for (int a = 0; a < ORC_NAGE; ++a) for (int q = 0; q < 5; ++q) c[q][a] = mom[(j*ORC_NAGE + a)*5 + q];
        ratio[j] = wsum(w, c[1], 5) / (wsum(w, c[0], 5) + wsum(w, c[1], 5));
        double tv1 = wsum(w, c[2], 5) + orc_weighted_cov(c[0], c[0], w, 5);
        double tv2 = wsum(w, c[4], 5) + orc_weighted_cov(c[1], c[1], w, 5);
        double stds = sqrt(fabs(tv1 * tv2));
        mean_corr[j] = wsum(w, c[3], 5) / stds;
        corr_mean[j] = orc_weighted_cov(c[0], c[1], w, 5) / stds;
        if (sample_guards) {
            int cov_all_zero = 1;
            for (int a = 0; a < ORC_NAGE; ++a) if (c[3][a] != 0.0) cov_all_zero = 0;
            if (!(wsum(w, c[0], 5) + wsum(w, c[1], 5) > 0.0)) ratio[j] = 0.0;       /* :64-71 */
            if (tv1 == 0.0 || tv2 == 0.0) { mean_corr[j] = 0.0; corr_mean[j] = 0.0; } /* :138-147 */
            if (cov_all_zero) mean_corr[j] = 0.0;
        }
        if (j == 5 || j == 6) {
            double* mo = (j == 5) ? pulse_mean : chase_mean;
            double* ff = (j == 5) ? pulse_ff : chase_ff;
            for (int a = 0; a < ORC_NAGE; ++a) {
                double tot = c[0][a] + c[1][a];
                double eps = (tot > 0.0 ? 0.0 : 0.0) + (tot == 0.0 ? 0.0001 : 0.0);
                mo[a] = tot;
                ff[a] = ((c[2][a] + 2 * c[3][a]) + c[4][a]) / (tot + eps);
            }
        }
    }
}

int orc_run_part_sim(const double* theta, int m, const orc_design_t* d, double mom[ORC_NCOND * ORC_NAGE * 5]) {
    double ss[9];
    int k = orc_transient_phase(theta, m, d->iv, d->cycle, d->rtol, d->atol, ss);
    for (int j = 0; j < ORC_NCOND; ++j)
        orc_syntheticdata(theta, m, ss, d, d->pulse[j], d->chase[j], &mom[j * ORC_NAGE * 5]);
    return k;
}

int orc_run_sim(const double* theta, int m, const orc_design_t* d, double stats[ORC_NSTATS], double* moments_out) {
    double mom[ORC_NCOND * ORC_NAGE * 5];
    int k = orc_run_part_sim(theta, m, d, mom);
    if (d->downsampling) {
        for (int j = 0; j < ORC_NCOND; ++j) {
            int off = (j < 6) ? 0 : ORC_NAGE; /* pulse cells' betas for conditions 1..6, chase cells' for 7..11 */
            double o[ORC_NAGE * 5];
            orc_downsample(&mom[j * ORC_NAGE * 5], d->beta_mean + off, d->beta_m2 + off, d->beta_var + off, o);
            memcpy(&mom[j * ORC_NAGE * 5], o, sizeof(o));
        }
    }
    orc_summary_stats(mom, d->age_dist, stats);
    if (moments_out) memcpy(moments_out, mom, sizeof(mom));
    return k;
}

/* ------------------------------------------------------------------ compute_errors.jl:30-43 */
double orc_nlsqerror_part(const double* data, const double* se, const double* s, int n, int n_summary_stats) {
    const double sigma = 0.1;
    double err = 0.0;
    for (int i = 0; i < n; ++i) {
        double eps = (se[i] + data[i] != 0.0) ? 0.0 : 0.0001;
        double diff = data[i] - s[i];
        double num = diff * diff;                         /* ^2 is x*x */
        double den = (se[i] * se[i] + (sigma * sigma) * (data[i] * data[i])) + eps; /* σ^2 * data^2: literal_pow on both */
        err += num / den;
    }
    return err / (double)n_summary_stats;
}

/* compute_errors.jl:45-70 */
void orc_compute_trunc_errors(const double* stats, int64_t n, const double* d, const double* se, int G, double* err) {
    static const int off[8] = {0, 5, 10, 15, 20, 31, 42, 53};
    for (int64_t i = 0; i < n; ++i) {
        const double* s = stats + i * ORC_NSTATS;
        for (int j = 0; j < G; ++j) {
            const double* dj = d + (int64_t)j * ORC_NSTATS;
            const double* sj = se + (int64_t)j * ORC_NSTATS;
            double e = 0.0;
            for (int l = 0; l < 7; ++l)
                e += orc_nlsqerror_part(dj + off[l], sj + off[l], s + off[l], off[l+1] - off[l], ORC_NSTATS);
            if (e > 10.0) e = 10.0;
            err[i * G + j] = e;
        }
    }
}

/* ------------------------------------------------------------------ accepted_particles.jl:19-30 */
typedef struct { double e; int64_t i; } acc_t;
static int cmp_acc(const void* a, const void* b) {
    const acc_t* x = (const acc_t*)a; const acc_t* y = (const acc_t*)b;
    if (x->e < y->e) return -1;
    if (x->e > y->e) return 1;
    return (x->i > y->i) - (x->i < y->i); /* stable sortperm == ties by ascending index */
}
int64_t orc_accept_gene(const double* err, int64_t n, int64_t stride, double eps, int64_t* idx) {
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) if (err[i * stride] <= eps) cnt++;
    if (cnt == 0) return 0;
    acc_t* v = (acc_t*)malloc(sizeof(acc_t) * (size_t)cnt);
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i) if (err[i * stride] <= eps) { v[k].e = err[i * stride]; v[k].i = i + 1; k++; }
    qsort(v, (size_t)cnt, sizeof(acc_t), cmp_acc);
    for (k = 0; k < cnt; ++k) idx[k] = v[k].i;
    free(v);
    return cnt;
}
