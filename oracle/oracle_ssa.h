/* oracle_ssa.h -- CPU restatement (TEST ORACLE) of the SSA stage: the CME of SURVEY 8a-CME
 * simulated with the direct method, one lineage at a time, with the Philox keying, schedule
 * construction and exact binomial thinning specified in DESIGN.md section 5.  Test infrastructure
 * only -- see abc_oracle.h. */
#ifndef ORACLE_SSA_H
#define ORACLE_SSA_H
#include <stdint.h>
#include "abc_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    double cycle;
    double agevec[ORC_NAGE];
    double pulse[ORC_NCOND], chase[ORC_NCOND];
    int n_cells, n_pre, downsampling;
    /* capture efficiencies quantised to 32 bits, grouped: pulse clusters 1..5 then chase clusters 1..5 */
    const uint32_t* beta_q32;
    int beta_off[11];
} orc_ssa_design_t;

/* math mode of the event step: 0 = libm (logf, sqrtf, /), 1 = the deterministic IEEE-only
 * formulation (bit-identical to the GPU kernel's exact_math variant) */
#define ORC_MATH_LIBM 0
#define ORC_MATH_DET  1

double orc_exp10_det(double x);
void orc_prior(int m, int64_t particle, uint64_t seed, double* theta);          /* abc_simulation.jl:3-11 */
void orc_quantise_betas(const double* betas, const int* clusters, int n, uint32_t* q32, int* off5); /* one cell class */
/* one read-out (condition j, age a) of one particle: counts = 4 x n_cells (U, L, U', L') */
void orc_ssa_readout(const double* theta, int m, const orc_ssa_design_t* d, int64_t particle, uint64_t seed,
                     int cond, int age, int math_mode, uint32_t* counts, uint64_t* n_events);
/* all 55 read-outs -> sample moments [cond][age][5] */
void orc_ssa_moments(const double* theta, int m, const orc_ssa_design_t* d, int64_t particle, uint64_t seed,
                     int math_mode, double* moments, uint64_t* n_events);
void orc_moments_from_sums(const uint64_t sums[5], int n_cells, double mom[5]);

#ifdef __cplusplus
}
#endif
#endif
