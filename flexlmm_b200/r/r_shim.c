/*
 * r_shim.c — the .Call binding of libflexlmm_cuda.so for the FIT_MODEL R script
 * (/root/reference/modules/local/r/fit_model.nf).  Thin by design: every SEXP is unpacked to the plain pointers of
 * include/flexlmm_cuda.h, one flx_* call per reference step, results copied into freshly allocated R vectors.
 *
 * Build inside the pgenlibr container (needs R's headers, not available in the B200 build image):
 *     R CMD SHLIB -o flexlmm_rshim.so r_shim.c -I<repo>/include -L<repo>/flexlmm_b200/lib -lflexlmm_cuda
 * In this repo it is syntax-checked against flexlmm_b200/r/stub/Rinternals.h (tests/test_host.py).
 *
 * Ownership: R owns every input (read-only for the duration of the call); outputs are PROTECTed until return; the
 * device state lives in a flx_handle destroyed before returning (also on error, before Rf_error longjmps).
 * Errors: any non-zero status becomes Rf_error(flx_last_error()) -> R stop() -> exit status 1 -> Nextflow errorStrategy
 * 'finish' (conf/base.config:13-18), the reference's failure semantics.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "flexlmm_cuda.h"

static void fail(flx_multi* h, const char* where) {
    char msg[1100];
    snprintf(msg, sizeof msg, "%s: %s", where, flx_last_error());
    if (h) flx_multi_destroy(h);
    Rf_error("%s", msg);
}

/* .Call("flexlmm_fit", pgen, var_idx, pgen_order, L, y_mm, X_null, ll_null, df_null, M0, M1, M2, M05, M15,
 *                       y_pred, U1, e_p, seeds, devices, use_dosage)
 * ORIG mode: seeds = integer(0) and y_pred/U1/e_p may be NULL.
 * M05 / M15: model.matrix with x = 0.5 / 1.5 on every row — how the SNP columns continue between hard calls (use_dosage).
 * devices: integer vector of CUDA ordinals; the task's SNPs are split over them (one R process drives the whole box: L and U1
 *          cross the host link once and are broadcast over NVLink, only the length-P min-p vector is reduced).
 * returns list(lrt_chisq, lrt_df, pval, rank, pivot (k x M), coef (k x M), min_pval (P), nvars_tested) */
SEXP flexlmm_fit(SEXP pgen, SEXP var_idx, SEXP pgen_order, SEXP L, SEXP y_mm, SEXP X_null, SEXP ll_null, SEXP df_null,
                 SEXP M0, SEXP M1, SEXP M2, SEXP M05, SEXP M15, SEXP y_pred, SEXP U1, SEXP e_p, SEXP seeds, SEXP devices, SEXP use_dosage) {
    const R_xlen_t n = XLENGTH(y_mm);
    const R_xlen_t M = XLENGTH(var_idx);
    const int c = (int)(XLENGTH(X_null) / n);
    const int k = (int)(XLENGTH(M0) / n);
    const int P = (int)XLENGTH(seeds);
    flx_config cfg;
    memset(&cfg, 0, sizeof cfg);
    flx_multi* h = NULL;
    if (flx_multi_create(&cfg, INTEGER(devices), (int32_t)XLENGTH(devices), &h)) fail(NULL, "flx_multi_create");
    if (flx_multi_set_whitening(h, REAL(L), (int64_t)n, (int64_t)n, 0)) fail(h, "flx_set_whitening");             /* fit_model.nf:61,128 */
    if (flx_multi_set_null(h, REAL(X_null), c, REAL(y_mm), Rf_asReal(ll_null), Rf_asInteger(df_null))) fail(h, "flx_set_null"); /* :59-60,63 */
    if (flx_multi_set_design(h, REAL(M0), REAL(M1), REAL(M2), k)) fail(h, "flx_set_design");                          /* :124-127 */
    if (flx_multi_set_design_probes(h, REAL(M05), REAL(M15))) fail(h, "flx_set_design_probes");
    if (flx_multi_set_use_dosage(h, Rf_asLogical(use_dosage) == TRUE)) fail(h, "flx_set_use_dosage");                 /* :26 Read / ReadHardcalls */
    if (P > 0 && flx_multi_set_perm(h, REAL(y_pred), REAL(U1), REAL(e_p), (int64_t)XLENGTH(e_p), INTEGER(seeds), P)) fail(h, "flx_set_perm"); /* :96-101 */
    if (flx_multi_open_pgen(h, CHAR(STRING_ELT(pgen, 0)), INTEGER(pgen_order), (int64_t)n)) fail(h, "flx_open_pgen"); /* :39-40,70 */

    SEXP out = PROTECT(Rf_allocVector(VECSXP, 8));
    SEXP chisq = PROTECT(Rf_allocVector(REALSXP, M)), pval = PROTECT(Rf_allocVector(REALSXP, M));
    SEXP df = PROTECT(Rf_allocVector(INTSXP, M)), rank = PROTECT(Rf_allocVector(INTSXP, M));
    SEXP pivot = PROTECT(Rf_allocMatrix(INTSXP, k, (int)M)), coef = PROTECT(Rf_allocMatrix(REALSXP, k, (int)M));
    SEXP minp = PROTECT(Rf_allocVector(REALSXP, P));
    flx_snp_out so;
    so.lrt_chisq = REAL(chisq); so.pval = REAL(pval); so.lrt_df = INTEGER(df); so.rank = INTEGER(rank);
    so.pivot = INTEGER(pivot); so.coef = REAL(coef);
    int64_t nv = 0;
    if (flx_multi_run(h, INTEGER(var_idx), (int64_t)M, &so, P > 0 ? REAL(minp) : NULL, &nv)) {                       /* :120-171 */
        UNPROTECT(8);
        fail(h, "flx_run");
    }
    flx_multi_destroy(h);
    for (R_xlen_t i = 0; i < M; i++)
        if (ISNAN(REAL(pval)[i])) REAL(pval)[i] = NA_REAL;                                                   /* :138 NA */
    SET_VECTOR_ELT(out, 0, chisq); SET_VECTOR_ELT(out, 1, df); SET_VECTOR_ELT(out, 2, pval); SET_VECTOR_ELT(out, 3, rank);
    SET_VECTOR_ELT(out, 4, pivot); SET_VECTOR_ELT(out, 5, coef); SET_VECTOR_ELT(out, 6, minp);
    SET_VECTOR_ELT(out, 7, Rf_ScalarReal((double)nv));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 8));
    const char* nm[8] = {"lrt_chisq", "lrt_df", "pval", "rank", "pivot", "coef", "min_pval", "nvars_tested"};
    for (int i = 0; i < 8; i++) SET_STRING_ELT(names, i, Rf_mkChar(nm[i]));
    Rf_setAttrib(out, R_NamesSymbol, names);
    UNPROTECT(9);
    return out;
}

/* .Call("flexlmm_sample", seed, m): set.seed(seed); sample.int(m) — lets the R side assert index parity */
SEXP flexlmm_sample(SEXP seed, SEXP m) {
    const R_xlen_t mm = (R_xlen_t)Rf_asReal(m);
    SEXP out = PROTECT(Rf_allocVector(INTSXP, mm));
    if (flx_perm_indices(Rf_asInteger(seed), (int64_t)mm, INTEGER(out))) { UNPROTECT(1); Rf_error("%s", flx_last_error()); }
    UNPROTECT(1);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"flexlmm_fit", (DL_FUNC)&flexlmm_fit, 19},
    {"flexlmm_sample", (DL_FUNC)&flexlmm_sample, 2},
    {NULL, NULL, 0}};

void R_init_flexlmm_rshim(DllInfo* dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
