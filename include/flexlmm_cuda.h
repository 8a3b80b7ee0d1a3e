/*
 * flexlmm_cuda.h — C ABI of the B200-native per-SNP likelihood-ratio-test engine.
 *
 * This is the drop-in boundary for the hot loop of birneylab/flexlmm's FIT_MODEL process
 * (/root/reference/modules/local/r/fit_model.nf:96-181).  The R script of that process keeps all
 * file I/O (readRDS, pvar/psam parsing, TSV writing) and calls these entry points through a thin
 * .Call shim (flexlmm_b200/r/r_shim.c); INTEGRATION.md shows the binding.  Everything here is
 * plain C: host pointers, sizes, column-major FP64 matrices exactly as R stores them, 1-based
 * indices where R uses them.  No torch types, no C++ in the signatures.
 *
 * Status codes: 0 = ok; negative = error, message available from flx_last_error().
 * Error semantics follow the reference: a missing hard call in a tested variant is fatal
 * (fit_model.nf:126-128 aborts in forwardsolve after model.frame drops the NA row).
 *
 * Thread model: one caller thread per handle; a handle owns one CUDA device, its streams and
 * (optionally) one NCCL communicator.
 */
#ifndef FLEXLMM_CUDA_H
#define FLEXLMM_CUDA_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FLX_OK                    0
#define FLX_ERR_BAD_ARG          -1
#define FLX_ERR_MISSING_GENOTYPE -2   /* NA hard call in a tested variant: the reference aborts */
#define FLX_ERR_CUDA             -3
#define FLX_ERR_NCCL             -4
#define FLX_ERR_OOM              -5
#define FLX_ERR_PGEN             -6   /* unreadable / unsupported .pgen */
#define FLX_ERR_STATE            -7   /* call order violated (e.g. run before set_design) */
#define FLX_ERR_NO_DEVICE        -8   /* no CUDA device: there is NO CPU fallback */

/* flx_config.flags */
#define FLX_FLAG_NO_INT8   1  /* never use the exact int8 tensor-core route (K2q/K4q); always whiten in FP64 (K2) */
#define FLX_FLAG_PLANES_5   4  /* int8 route with 5 digit planes (40 bits, ~1e-12 relative in x'V^-1x) instead of 6 (48 bits, ~4e-15)        */


#define FLX_MAX_K   24   /* design columns of the real model (k)            */
#define FLX_MAX_S    8   /* SNP-dependent design columns (s)                */

typedef struct flx_handle flx_handle;

typedef struct {
    int32_t device;            /* CUDA device ordinal                                                  */
    int32_t batch_tiles;       /* column tiles (128 design columns each) per batch; 0 = auto (SM count)*/
    int32_t verbose;           /* 1: timing lines on stderr                                            */
    int32_t flags;             /* FLX_FLAG_* bits                                                      */
} flx_config;

/* Per-SNP results in var_idx order, structure-of-arrays (what the R shim hands back as vectors / a k x M
 * matrix).  Mirrors the ORIG-mode TSV columns lrt_chisq, lrt_df, pval, beta (fit_model.nf:112-116,145-169):
 *   coef are the .lm.fit coefficients in PIVOTED order exactly as the reference pastes them against the
 *   unpivoted column names (fit_model.nf:150-155); pivot is 1-based like R's fit$pivot.
 * Any pointer may be NULL (that output is skipped).  pivot / coef: k values per SNP, SNP-major (R: k x M). */
typedef struct {
    double*  lrt_chisq;        /* [M]                                                                  */
    double*  pval;             /* [M] NaN when lrt_df <= 0 (R NA, fit_model.nf:138)                    */
    int32_t* lrt_df;           /* [M]                                                                  */
    int32_t* rank;             /* [M]                                                                  */
    int32_t* pivot;            /* [M*k]                                                                */
    double*  coef;             /* [M*k]                                                                */
} flx_snp_out;

/* Timings and counters of the last flx_run*, for bench.py (device times from CUDA events, ms). */
typedef struct {
    double ms_total;           /* whole flx_run* call on the device stream                             */
    double ms_decode;          /* K0/K1 pgen_expand + decode_2bit_tile / decode_i8_tile                */
    double ms_whiten;          /* K2  whiten_tile (L^-1 . X)  or  K2q int8 quadratic form                      */
    double ms_fit;             /* K3  snp_gram + snp_solve; int8 route: K4q (u, b, b*) + snp_solve     */
    double ms_perm;            /* K4  perm_contract_minp; int8 route: perm_reduce                      */
    double ms_h2d;             /* host->device copies inside the call                                  */
    double ms_d2h;             /* device->host copies inside the call                                  */
    int64_t launches;          /* kernels launched by this library inside the call                     */
    int64_t launches_whiten;
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    int64_t snps;
    int64_t n_pad;
    int64_t cols_per_tile;     /* design columns actually used per 128-wide tile                       */
    int64_t int8_path;         /* 1: the batches ran K2q (int8 tensor cores), 0: K2 (FP64 DMMA)        */
    int64_t refit_snps;        /* SNPs re-fitted on the FP64 path (near-collinear with constant columns)*/
    int64_t int8_passes;       /* K2q work per batch in triangular-pass units (6 n^2 int8 ops per SNP): 1 for one SNP term, 5 for x + I(x==1) + x:cov */
    int64_t fallback_reason;   /* FLX_FALLBACK_*: why the task runs on the FP64 route (0 when it runs on the int8 route)                         */
    double ms_wall;            /* host wall clock of the whole flx_run* call (entry to return), ms                                              */
    double ms_batches;         /* device time of the batch loop(s) alone (ms_total minus the final copies / collective)                          */
} flx_stats;

/* flx_stats.fallback_reason: the int8 tensor-core route (K2q/K4q) is ~14x the FP64 route, so a task that cannot use it says why */
#define FLX_FALLBACK_NONE            0
#define FLX_FALLBACK_FLAG            1   /* FLX_FLAG_NO_INT8 was set                                                            */
#define FLX_FALLBACK_NOT_SEPARABLE   2   /* a SNP term is not (integer genotype triple) x (per-sample scale), e.g. exp(x * cov)  */
#define FLX_FALLBACK_TOO_MANY_FORMS  3   /* more distinct triples / scales / terms than the int8 plan holds                      */
#define FLX_FALLBACK_NO_DENSE_INVERSE 4  /* design / permutations re-set after the dense L^-1 was released (n^2 * 8 B did not fit next to the planes): call flx_set_whitening again */
#define FLX_FALLBACK_PLANES_DONT_FIT 5   /* the digit planes of the design's scale pairs do not fit in HBM next to L^-1          */
#define FLX_FALLBACK_DOSAGE          6   /* real-valued dosages (pgenlibr::Read): genotypes are not small integers               */

const char* flx_last_error(void);
const char* flx_version(void);
int  flx_device_count(void);

/* fit_model.nf:30-32,56-66 — create / destroy the per-task state. */
int  flx_create(const flx_config* cfg, flx_handle** out);
void flx_destroy(flx_handle* h);

/* L of l.mm[["L"]] (fit_model.nf:61; decorrelate.nf:57,77): n x n column-major, lower triangle used,
 * exactly what forwardsolve(L, X) reads (fit_model.nf:128).  is_inverse != 0: the matrix already is L^-1. */
int  flx_set_whitening(flx_handle* h, const double* L, int64_t n, int64_t ld, int is_inverse);

/* The same in pieces, for factors that should not exist twice (n = 100,000: 80 GB): _begin allocates the n x n device matrix
 * (after releasing the previous factor's), _cols copies columns [j0, j0 + ncols) from `cols` (column-major, leading dimension ld;
 * HOST or DEVICE memory), _end inverts in place and packs exactly like flx_set_whitening.  flx_set_whitening(L) is
 * _begin(n); _cols(0, n, L, ld); _end(is_inverse). */
int  flx_whitening_begin(flx_handle* h, int64_t n);
int  flx_whitening_cols(flx_handle* h, int64_t j0, int64_t ncols, const double* cols, int64_t ld);
int  flx_whitening_end(flx_handle* h, int is_inverse);

/* DECORRELATE on the GPU (/root/reference/modules/local/r/decorrelate.nf:50-77) — the producer of the hot path's largest input:
 *   V = s2g*K + s2e*I (:54);  L = t(chol(V)) (:57);  if the Cholesky fails, eigen-decompose V, require every eigenvalue
 *   >= min_allowed_eig (:65), replace those <= 0 by eig_replacement, V <- U |lambda| U' and factor again (:58-69);
 *   y.mm / X.mm = forwardsolve(L, .) (:71-72).
 * K: n x n column-major symmetric GRM;  X: n x c.  Outputs are host arrays and may be NULL (L_out is n x n, lower triangle,
 * zeros above).  The factor also becomes this handle's whitening matrix, exactly as if flx_set_whitening(L_out) had been
 * called, without the n^2 round trip through the host.  *clipped_eigs (may be NULL): number of eigenvalues replaced, 0 when
 * the plain Cholesky succeeded.  Defaults of the reference: min_allowed_eig = -1e-6, eig_replacement = 1e-10 (nextflow.config:54-56). */
int  flx_decorrelate(flx_handle* h, const double* K, int64_t n, double s2g, double s2e, const double* y, const double* X, int32_t c,
                     double min_allowed_eig, double eig_replacement, double* L_out, double* y_mm_out, double* X_mm_out, int32_t* clipped_eigs);

/* y.mm, X.mm (null design, whitened) and the null log-likelihood with its df attribute
 * (fit_model.nf:59-60,63; fit_null_model.nf:33-34,72).  df_null = attr(ll, "df") = rank + 1. */
int  flx_set_null(flx_handle* h, const double* X_null, int32_t c, const double* y_mm, double ll_null, int32_t df_null);

/* The real model's design as a genotype lookup (fit_model.nf:124-127): M0/M1/M2 are
 * model.matrix(model[-2], model.frame(...)) evaluated with x == 0, 1, 2 on every row, n x k column-major,
 * columns in model.matrix order.  Columns identical in all three are SNP-independent. */
int  flx_set_design(flx_handle* h, const double* M0, const double* M1, const double* M2, int32_t k);

/* use_dosage (fit_model.nf:12,26 — true is the reference's default, nextflow.config:59): read variants with pgenlibr::Read instead of
 * ReadHardcalls, i.e. x = dosage_uint16 / 16384 for every sample the record's dosage track lists, the hard call elsewhere.  A file
 * without dosage tracks gives the same x either way (and keeps the int8 route).  With dosages x is real-valued, so the three
 * hard-call lookups no longer determine a SNP column: flx_set_design_probes hands over model.matrix evaluated with x = 0.5 and
 * x = 1.5 on every row (n x k, same column order), from which every SNP column is recognised as affine in x (x, x:cov — evaluated as
 * v(0) + x (v(1) - v(0))) or as the indicator of one hard-call value (I(x == 1)); anything else is refused when the run starts.
 * Such a task runs on the FP64 route (flx_stats.fallback_reason = FLX_FALLBACK_DOSAGE). */
int  flx_set_design_probes(flx_handle* h, const double* M05, const double* M15);
int  flx_set_use_dosage(flx_handle* h, int use_dosage);

/* Permutation inputs (fit_model.nf:64-66,96-101): y.mm.pred (n), U1 (n x m col-major), e.p (m) and the
 * permutation seeds (workflows/flexlmm.nf:67: 1..P).  Index order is R's set.seed(seed); sample(e.p). */
int  flx_set_perm(flx_handle* h, const double* y_pred, const double* U1, const double* e_p, int64_t m,
                  const int32_t* seeds, int32_t P);

/* FIT_NULL_MODEL without the n x n eigen-decomposition (/root/reference/modules/local/r/fit_null_model.nf:33-72).
 * The reference needs SOME orthonormal basis U1 of the orthogonal complement of the null design to decorrelate the
 * residuals (e.p = U1' e, :65) and takes the LAPACK eigenvectors of I - QQ' (:52-62) — an O(n^3) step with three n x n
 * matrices, and an arbitrary basis of a repeated eigenvalue.  Here the basis is the implicit Householder complement
 * Q[:, rank:] of the QR of X.mm: O(nc) storage, never materialised.
 *   outputs (host, any may be NULL): ll (+ df = rank + 1), y_pred = fitted (n), resid (n), e_p (m = n - rank),
 *   U1 (n x m column-major, explicit copy of the implicit basis — for tests / for feeding the reference's own PERM path).
 * Host-only; no GPU needed. */
int  flx_fit_null_model(const double* X_mm, int32_t c, const double* y_mm, int64_t n, double* ll, int32_t* df,
                        double* y_pred, double* resid, double* e_p, double* U1, int64_t* m);

/* Permutations from the implicit basis: same as flx_set_perm(y_pred, U1, e_p, ...) with the y_pred / U1 / e_p that
 * flx_fit_null_model derives from the null model already given to flx_set_null, but U1 (n x m, 80 GB at n = 10^5) is never
 * formed: y* = y_pred + Q [0; e_p[idx]] costs O(ncP).  Index order is R's set.seed(seed); sample(e.p) as always. */
int  flx_set_perm_implicit(flx_handle* h, const int32_t* seeds, int32_t P);

/* Genotype source A: a .pgen file (fit_model.nf:39-40,54,123).  pgen_order[i] = 1-based raw-sample index of
 * model sample i (fit_model.nf:70,124: match(samples, psam$IID)). Modes 0x02 and 0x10 (hard calls). */
int  flx_open_pgen(flx_handle* h, const char* path, const int32_t* pgen_order, int64_t n);

/* pgenlibr::NewPgen metadata of the open file: raw variant / sample counts and whether any record carries a dosage
 * track (vrtype bits 5-6).  The engine reads hard calls (ReadHardcalls, fit_model.nf:26 with use_dosage = false) unless
 * flx_set_use_dosage(h, 1) was called. */
int  flx_pgen_info(flx_handle* h, int64_t* variant_ct, int64_t* sample_ct, int32_t* has_dosage);

/* Genotype source B: packed 2-bit hard-call records already in host memory (pgen mode-0x02 body:
 * variant v at packed + v*bytes_per_variant, sample i at bits 2*(i&3) of byte i>>2; 3 = missing). */
int  flx_set_packed(flx_handle* h, const uint8_t* packed, int64_t n_variants, int64_t bytes_per_variant,
                    int64_t raw_sample_ct, const int32_t* pgen_order, int64_t n);

/* Copy the packed records of source B into HBM once (bench "value": inputs resident before timing). */
int  flx_make_resident(flx_handle* h);

/* The hot loop (fit_model.nf:120-171) for var_idx[0..M) (1-based variant numbers, get_var_idx.nf:34-35).
 *   out  : per-SNP outputs (may be NULL to skip the per-SNP D2H copy, PERM-only use)
 *   minp : P minima over the M variants, one per seed (fit_model.nf:141-143); NULL if no permutations set
 *   nvars_tested: M on success (fit_model.nf:142,178)
 * With an attached communicator minp is min-allreduced and nvars_tested sum-allreduced over the ranks. */
int  flx_run(flx_handle* h, const int32_t* var_idx, int64_t M, const flx_snp_out* out, double* minp, int64_t* nvars_tested);

/* Derive the per-task operands (whitened constant columns, permutation prologue fit_model.nf:96-101 for every seed, digit
 * planes) NOW instead of inside the first flx_run.  The LOCO driver (flexlmm_b200/loco.py; lmm.nf:27-46 gives every chromosome
 * its own factor) calls it for chromosome k+1 on a second handle of the same device, from a host thread, while flx_run of
 * chromosome k is in flight: handles share no mutable state. */
int  flx_prepare(flx_handle* h);

int  flx_get_stats(flx_handle* h, flx_stats* out);

/* R-exact permutation indices (set.seed(seed); sample.int(m)), 1-based. Host-only, bit-exact. */
int  flx_perm_indices(int32_t seed, int64_t m, int32_t* out);

/* Test hook for K1: decode var_idx[0..M) and return the design columns X (n x M*s, column-major; the
 * s SNP-dependent columns of variant j at columns j*s .. j*s+s-1) and, if codes != NULL, the gathered
 * 2-bit codes (M x n, row-major, value 3 = missing; no error is raised for them here). */
int  flx_decode_tile(flx_handle* h, const int32_t* var_idx, int64_t M, double* X, uint8_t* codes);

/* Test hook for K2: whiten arbitrary columns, W = L^-1 . X  (n x ncols, column-major both). */
int  flx_whiten(flx_handle* h, const double* X, int64_t ncols, double* W);

/* Multi-GPU (SNP-block sharding, one handle per GPU): NCCL communicator for the min-p allreduce.
 * flx_comm_unique_id fills 128 bytes on rank 0; broadcast them by any means, then attach on every rank. */
int  flx_comm_unique_id(uint8_t id[128]);
int  flx_comm_attach(flx_handle* h, const uint8_t id[128], int32_t rank, int32_t nranks);

/* Replicate-once forms of the two n^2-sized inputs: collective over the attached communicator.  Only `root` reads L /
 * U1 from the host (other ranks may pass NULL), uploads it and — for L — inverts it; L^-1 / U1 then reach the other GPUs with
 * ncclBroadcast over NVLink: one host upload and one inversion per box instead of one per GPU.  Without a communicator they are
 * flx_set_whitening / flx_set_perm.  A failure on any rank is reported on every rank before the broadcast starts. */
int  flx_set_whitening_bcast(flx_handle* h, const double* L, int64_t n, int64_t ld, int is_inverse, int32_t root);
int  flx_set_perm_bcast(flx_handle* h, const double* y_pred, const double* U1, const double* e_p, int64_t m, const int32_t* seeds, int32_t P, int32_t root);

/* One caller thread, several GPUs of one box (R is single-threaded; "one process drives all 8 GPUs").  A flx_multi owns
 * one handle per listed device and the NCCL communicator between them.  Each call below does what its single-GPU namesake does,
 * on every device (one host thread per device inside the call); flx_multi_set_whitening / _set_perm upload once and broadcast;
 * flx_multi_run splits var_idx into one contiguous SNP block per device (fit_model.nf:120 loop, lmm.nf:47-71 scatter), writes the
 * per-SNP outputs into disjoint slices of the caller's arrays in var_idx order, and returns the min over all devices of the
 * length-P min-p vector (ncclAllReduce(min) inside) and the summed count.  Results are bit-identical to one device's. */
typedef struct flx_multi flx_multi;
int  flx_multi_create(const flx_config* cfg /* device field ignored */, const int32_t* devices, int32_t ndev, flx_multi** out);
void flx_multi_destroy(flx_multi* m);
int  flx_multi_device_count(flx_multi* m);
int  flx_multi_set_whitening(flx_multi* m, const double* L, int64_t n, int64_t ld, int is_inverse);
int  flx_multi_set_null(flx_multi* m, const double* X_null, int32_t c, const double* y_mm, double ll_null, int32_t df_null);
int  flx_multi_set_design(flx_multi* m, const double* M0, const double* M1, const double* M2, int32_t k);
int  flx_multi_set_design_probes(flx_multi* m, const double* M05, const double* M15);
int  flx_multi_set_use_dosage(flx_multi* m, int use_dosage);
int  flx_multi_set_perm(flx_multi* m, const double* y_pred, const double* U1, const double* e_p, int64_t mm, const int32_t* seeds, int32_t P);
int  flx_multi_set_perm_implicit(flx_multi* m, const int32_t* seeds, int32_t P);
int  flx_multi_open_pgen(flx_multi* m, const char* path, const int32_t* pgen_order, int64_t n);
int  flx_multi_set_packed(flx_multi* m, const uint8_t* packed, int64_t n_variants, int64_t bytes_per_variant, int64_t raw_sample_ct,
                          const int32_t* pgen_order, int64_t n);
int  flx_multi_make_resident(flx_multi* m);
int  flx_multi_prepare(flx_multi* m);
int  flx_multi_run(flx_multi* m, const int32_t* var_idx, int64_t M, const flx_snp_out* out, double* minp, int64_t* nvars_tested);
int  flx_multi_get_stats(flx_multi* m, int32_t device_index, flx_stats* out);

/* Downstream definition of the minP threshold (get_null_p_dist.nf:77-79; make_plots.nf:31):
 * quantile(min_pval, p_thr), R type 7. Host-only. */
double flx_minp_threshold(const double* min_pval, int64_t P, double p_thr);

#ifdef __cplusplus
}
#endif
#endif
