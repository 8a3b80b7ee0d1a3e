/*
 * flexlmm_oracle.h — CPU restatement ("oracle") of the per-SNP LRT loop of birneylab/flexlmm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (flexlmm_b200/, include/) may link,
 * import or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (R script modules/local/r/fit_model.nf) holds no golden vectors
 * or numerical tests, and its arithmetic lives in un-vendored third-party code (R stats/base,
 * LINPACK dqrdc2/dqrsl, nmath, pgenlibr/pgenlib) that cannot run in this image.  The functions
 * below restate those published algorithms; they are pinned only by the known answers listed in
 * SURVEY.md Appendix B (R RNG vectors, test-asset pgen bytes, test-asset fit) and by independent
 * cross-checks (numpy lstsq, scipy chi2.sf, mpmath) in tests/.
 *
 * Every function cites the reference call site (relative to /root/reference) it restates.
 */
#ifndef FLEXLMM_ORACLE_H
#define FLEXLMM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- R RNG: set.seed(seed); sample(m)  (fit_model.nf:97-98; R RNG.c, R_unif_index, do_sample) ---- */
typedef struct { uint32_t mt[624]; int mti; } orc_rng;
void   orc_set_seed(orc_rng* g, int32_t seed);
double orc_unif_rand(orc_rng* g);
/* out[i] = 1-based permutation of 1..m exactly as `sample.int(m)` after set.seed(seed). */
void   orc_sample_perm(int32_t seed, int64_t m, int32_t* out);

/* ---- pgen decode (fit_model.nf:39-40,54,123; pgenlibr ReadHardcalls) ---- */
/* Parse a .pgen held in memory. Supports storage mode 0x02 (fixed-width 2-bit) and 0x10
 * (variable-width; vrtypes 0,1,2,3,4,6,7 with difflists; no multiallelic/phase/dosage tracks).
 * Decodes variant `vidx0` (0-based) into codes[0..sample_ct) with values 0,1,2,3 (3 = missing).
 * Returns 0 on success, negative on error. */
int orc_pgen_info(const uint8_t* file, int64_t len, uint32_t* variant_ct, uint32_t* sample_ct, int* mode);
int orc_pgen_read_hardcalls(const uint8_t* file, int64_t len, uint32_t vidx0, uint8_t* codes);
/* pgenlibr::Read (use_dosage = true, the reference's default: nextflow.config:59): hard calls overlaid with dosage_uint16 / 16384
 * where the record's dosage track (vrtype bits 5-6) has an entry; NA (NaN) for a missing call without dosage. */
int orc_pgen_read_dosage(const uint8_t* file, int64_t len, uint32_t vidx0, double* buf);
/* pgenlibr buffer semantics: 0,1,2 -> 0.0,1.0,2.0 ; 3 -> NA (R NA_real_, a NaN payload 1954). */
void orc_codes_to_buf(const uint8_t* codes, int64_t n, double* buf);

/* ---- forwardsolve(L, X) (fit_model.nf:128; reference-BLAS dtrsm Left/Lower/NoTrans/NonUnit) ---- */
void orc_forwardsolve(const double* L, int64_t n, int64_t ldl, double* X, int64_t k, int64_t ldx);

/* ---- .lm.fit (fit_model.nf:100,133; Cdqrls -> dqrls -> dqrdc2 + dqrsl, tol 1e-7) ---- */
/* x: n*p col-major, overwritten with the QR; y: n. Outputs: coef[p] (PIVOTED order, zeros past rank),
 * resid[n], effects[n], pivot[p] (1-based), qraux[p], returns rank. work: 2*p doubles. */
int orc_lm_fit(double* x, int64_t n, int p, const double* y, double tol,
               double* coef, double* resid, double* effects, int* pivot, double* qraux);

/* ---- stats:::logLik.lm (fit_model.nf:101,134) ---- */
double orc_loglik_lm(const double* resid, int64_t n);

/* ---- pchisq(q, df, lower.tail=FALSE) (fit_model.nf:138), integer df closed forms ---- */
double orc_pchisq_upper(double q, int df);

/* ---- one FIT_MODEL task, literal loop (fit_model.nf:96-181) ----
 * Design rows are taken from M0/M1/M2 (model.matrix evaluated with x==0,1,2; n*k col-major each),
 * which is exactly what model.frame/model.matrix (:126-127) produce for a hard-call x.
 * geno: M variants * n codes (already gathered to model order, :124), values 0..2 (3 => returns -2,
 *       the reference aborts on a missing genotype).
 * perm_seed <= 0 : ORIG mode, uses ll_null/df_null as given.
 * perm_seed  > 0 : PERM mode, builds y from y_pred + U1 %*% sample(e_p), refits the null.
 * Outputs per SNP (ORIG and PERM): chisq[M], df[M], pval[M] (NaN for NA), rank[M], pivot[M*k], coef[M*k];
 * any of them may be NULL.  min_pval / nvars_tested as in :141-143 (PERM), also filled in ORIG mode.
 */
typedef struct {
    int64_t n; int k; int c_null;
    const double* L; const double* y_mm; const double* X_null;
    double ll_null; int df_null;
    const double* M0; const double* M1; const double* M2;
    const double* y_pred; const double* U1; const double* e_p; int64_t m;
} orc_task;
int orc_fit_model(const orc_task* t, const uint8_t* geno, int64_t M, int32_t perm_seed,
                  double* chisq, int* df, double* pval, int* rank, int* pivot, double* coef,
                  double* min_pval, int64_t* nvars_tested);

#ifdef __cplusplus
}
#endif
#endif
