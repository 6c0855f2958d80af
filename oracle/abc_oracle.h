/*
 * abc_oracle.h -- CPU restatement (test oracle) of the reference's ABC hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (abc_inference_transcription_b200/ + libabcb200.so) never links or imports it.
 *
 * Reference: pthomaslab/abc_inference_transcription (Julia).  Each function cites the
 * reference file:line it restates.  The reference cannot be run here (no Julia, no
 * Sundials): parity is pinned on the reference's shipped result files
 *   data/recovered_statistics/ ** (= syntheticdata() at data/posterior_estimates/map_sets_*.txt)
 * for the simulator (M1..M9) and on the source text for downsample / the 53 statistics /
 * scoring / acceptance, for which no shipped file exists ("parity unpinned" rows, see
 * DESIGN.md section 3).
 */
#ifndef ABC_ORACLE_H
#define ABC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NAGE 5
#define ORC_NCOND 11
#define ORC_NSTATS 53

/* Experimental design (scripts/abc_simulation.jl:65-79 + the globals of
 * scripts/load_process_data.jl:59-83 that the reference does not ship). */
typedef struct {
    double cycle;                 /* 20.0                      abc_simulation.jl:72 */
    double t0;                    /* -3*cycle                  abc_simulation.jl:73 */
    double agevec[ORC_NAGE];      /* tau_ * cycle = 2,6,...,18 abc_simulation.jl:74 */
    double pulse[ORC_NCOND];      /* condition_id[:,1]         abc_simulation.jl:65 */
    double chase[ORC_NCOND];      /* condition_id[:,2] */
    double age_dist[ORC_NAGE * ORC_NCOND]; /* column-major 5x11, used as given (never renormalised) */
    double iv[9];                 /* zeros, iv[2]=1/2 (abc_simulation.jl:70-71) or iv[1]=1/2 (recover_statistics.jl:33-34) */
    int downsampling;             /* abc_simulation.jl:79 */
    /* beta moments per age cluster for pulse cells [0..4] and chase cells [5..9]:
     * mean(beta), mean(beta^2), var(beta) (n-1)       model.jl:229-234 */
    double beta_mean[2 * ORC_NAGE];
    double beta_m2[2 * ORC_NAGE];
    double beta_var[2 * ORC_NAGE];
    /* integrator tolerances of the restatement (CVODE defaults are 1e-3 / 1e-6) */
    double rtol, atol;
} orc_design_t;

/* ---- simulator: scripts/model.jl ---- */
int  orc_n_params(int m);                                                    /* 5 or 9 */
void orc_get_rate(const double* theta, int m, double cycle, double t, double p_log10[4]); /* model.jl:1-22 */
double orc_size_scaling(double cycle, double t);                             /* model.jl:25-27 */
double orc_labelling(double lambda_log10, double texp, double pulse, double t); /* model.jl:58-64 */
void orc_f(const double y[9], const double p[4], double lam, double dy[9]);  /* model.jl:74-86 */
long orc_model(const double* theta, int m, const double iv[9], double tmin, double tmax,
               double cycle, double texp, double pulse, double rtol, double atol, double out[9]); /* model.jl:89-96 */
void orc_periodic_boundary(const double e[9], double v[9]);                  /* model.jl:98-111 */
int  orc_transient_phase(const double* theta, int m, const double iv[9], double cycle,
                         double rtol, double atol, double ss_iv[9]);         /* model.jl:114-142 */
void orc_trajectories(const double* theta, int m, const double iv[9], double age, double cycle,
                      double pulse, double t0, double texp, double rtol, double atol,
                      double mean2[2], double cov3[3]);                      /* model.jl:146-174 */
void orc_syntheticdata(const double* theta, int m, const double ss_iv[9], const orc_design_t* d,
                       double pulse, double chase, double s[ORC_NAGE * 5]);  /* model.jl:176-187; s row-major [age][5] */
void orc_downsample(const double s[ORC_NAGE * 5], const double* bmean, const double* bm2,
                    const double* bvar, double out[ORC_NAGE * 5]);           /* model.jl:221-239 */
void orc_beta_moments(const double* betas, const int* clusters, int n, double bmean[ORC_NAGE],
                      double bm2[ORC_NAGE], double bvar[ORC_NAGE]);          /* model.jl:226-235 */

/* ---- statistics: scripts/abc_simulation.jl:22-46 (run_sim, model_realisation.jl:3-37) ---- */
double orc_weighted_cov(const double* x, const double* y, const double* w, int n); /* data_summary_statistics.jl:179-181 */
/* moments: [cond][age][5] (mean_u, mean_l, var_u, cov_ul, var_l) -> 53 statistics in the order
 * pulse_mean[5], pulse_ff[5], chase_mean[5], chase_ff[5], ratio[11], mean_corr[11], corr_mean[11] */
void orc_summary_stats(const double moments[ORC_NCOND * ORC_NAGE * 5], const double* age_dist,
                       double stats[ORC_NSTATS]);
/* same, for moments estimated from a finite sample of cells (SSA): degenerate samples as on the data side,
 * data_summary_statistics.jl:64-71, 138-147 */
void orc_summary_stats_sample(const double moments[ORC_NCOND * ORC_NAGE * 5], const double* age_dist,
                              double stats[ORC_NSTATS]);
/* full per-particle path: abc_sim / run_sim */
int  orc_run_sim(const double* theta, int m, const orc_design_t* d, double stats[ORC_NSTATS],
                 double* moments_out /* nullable, [11][5][5] after optional downsampling */);
/* pre-downsampling moments for all 11 conditions: recover_statistics.jl:1-11 (run_part_sim) */
int  orc_run_part_sim(const double* theta, int m, const orc_design_t* d, double moments[ORC_NCOND * ORC_NAGE * 5]);

/* ---- scoring: scripts/compute_errors.jl ---- */
double orc_nlsqerror_part(const double* data, const double* se, const double* s, int n, int n_summary_stats); /* :30-43 */
/* stats: n x 53 row-major (particle-major); d, se: G x 53 row-major (gene-major), same 53-order.
 * err: n x G row-major (one row per particle, like error_<model>.txt). compute_errors.jl:45-70 */
void orc_compute_trunc_errors(const double* stats, int64_t n, const double* d, const double* se, int G, double* err);

/* ---- acceptance: scripts/accepted_particles.jl:10-32 ---- */
/* err column for one gene (stride between particles given). Writes 1-based indices sorted by
 * (err asc, index asc) into idx (capacity n); returns the count (0 => the reference writes "0"). */
int64_t orc_accept_gene(const double* err, int64_t n, int64_t stride, double eps, int64_t* idx);

#ifdef __cplusplus
}
#endif
#endif
