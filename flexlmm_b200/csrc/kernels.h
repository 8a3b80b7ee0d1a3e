// kernels.h — launch interfaces of the sm_100a kernels (host side of csrc/*.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace flx {

// ---------------------------------------------------------------- K2 / K4 (panel_gemm.cu)
struct GemmArgs {
    const double* A;          // packed panels (K2: triangular L^-1 by row tile; K4: W tiles)
    const double* B;          // packed panels (K2: X tiles; K4: R permutation tiles)
    double* Out;              // K2: packed W tiles
    int64_t n_items;
    int T;                    // K2: row tiles
    int NJ;                   // design-column tiles in the batch
    int NP;                   // K4: permutation tiles
    int KC;                   // k-chunks (n_pad / 16)
    int KC_live;              // K2: k-chunks that hold real samples = ceil(n / 16)
    int last_row_blocks;      // K2: 8-row blocks of the last row tile that hold real samples (1..16)
    // K4 epilogue
    const int32_t* snp_df;    // per SNP in batch: rank bucket 1..s = lrt_df - (cr + 1 - df_null) (0 => skip)
    const double* snp_sm;     // per SNP: s(s+1)/2 packed symmetric matrix Sm, dSS = b' Sm b
    const double* perm_inv_rss;  // per permutation column: 1 / RSS_f,p (0 for padding)
    double* perm_maxratio;    // [s][P_pad] running max of dSS / RSS_f,p per (df, permutation)
    int64_t P_pad;
};
int launch_whiten(const GemmArgs& g, int num_sms, cudaStream_t st);
int launch_perm_contract(const GemmArgs& g, int s, int num_sms, cudaStream_t st);

// ---------------------------------------------------------------- K2q (qf.cu): x' V^-1 x on the int8 tensor cores
struct QfArgs {
    const uint32_t* X2;       // [tile pair][NKB][tile 2][chunk 4][SNP 128] 2-bit code words of the MMA operand g' (qf.cu, K1q)
    uint32_t vt_mma;          // its code -> int8 value table (byte c = value of code c; tile_code_map)
    const uint32_t* Xd2[2];   // code tiles of the left vectors g of the forms g' T g' (same layout; Xd2[0] == X2 for a plain quadratic form)
    uint32_t vt_dot[2];
    int ndot;                 // 1 or 2 left vectors per pass
    const int8_t* T8;         // [row tile I][kb < 4(I+1)][Q planes][16 KB] digit planes of T = tril(C_a V^-1 C_a) (half diagonal), or,
                              // full != 0, [row tile I][kb < NKB][Q planes][16 KB] planes of the full matrix C_a V^-1 C_b
    int full;
    const double* rowscale;   // [n_pad256] power-of-two scale of each row of T (one value per 16 consecutive rows)
    double plane_scale[8];    // 2^-8(q+1)
    double* part[2];          // per left vector: [T256][Q][nsnp] partial forms
    int64_t n_items;
    int64_t nsnp;
    int64_t n;                // model samples (rows of T beyond n are padding)
    int T256, NJ, Q;
    int GP;                   // SNP tile pairs per L2 group (qf_item); >= ceil(NJ / 2): no grouping
    int NKB;                  // k-blocks per tile (n_pad256 / 64)
    int NKB_live;             // k-blocks that hold real samples
    long long* prof;          // optional [16 * grid] cycle counters of the three roles (qf.cu, PROF builds): FitModel(verbose=2)
};
int launch_qf(const QfArgs& g, int num_sms, cudaStream_t st);

// K4q (qf.cu): the code tiles against digit planes of a column block of G = L^-T [Qc | r | R_f]
struct XgArgs {
    const uint32_t* X2;       // [tile pair][NKB][tile 2][chunk 4][SNP 128] 2-bit code words
    uint32_t vt;              // code -> int8 value table of the tile set
    const int8_t* G8;         // [col block][NKB][Q][NB*64 B]
    const double* colscale;   // [ntiles*NB] power-of-two column scales
    double pair_scale[3];     // 2^-16, 2^-32, 2^-48
    int64_t n_items;          // column blocks * ceil(NJ / 2): a work item is a pair of SNP tiles against one column block
    int64_t nsnp;
    int NJ, NB, NC, Q, NKB, NKB_live;
    int NCT;                  // column blocks
    double* out; int64_t ld_out;   // out[snp * ld_out + col], ld_out = NCT * NB (a multiple of 8)
};
int launch_xg(const XgArgs& g, int num_sms, cudaStream_t st);
int xg_block_cols(int Q);     // widest column block: Q * NB <= 256 TMEM columns per SNP tile
// columns 0..c1-1 come from G (ld), columns c1..NC-1 from G2 (ld2) — [Qc | r] and R_f live in different buffers.
// rowmul (optional, [n]): the planes are those of diag(rowmul) [G | G2]
int launch_xg_slice(const double* G, int64_t ld, int c1, const double* G2, int64_t ld2, int64_t n, int NC, int NB, int ntiles, int Q, int NKB, const double* rowmul,
                    double* colscale, int8_t* G8, cudaStream_t st);
// Vtile: rows 256 I .. 256 I + 255 of V^-1 (column-major, ld), columns 0 .. row (full: all n columns).  ca, cb (optional, [n]):
// the planes are those of tril(diag(ca) V^-1 diag(cb)) with half diagonal, or of the full matrix diag(ca) V^-1 diag(cb)
int launch_qf_slice(const double* Vtile, int64_t n, int64_t ld, int I, int T256, int Q, const double* ca, const double* cb, int full, double* rowscale, int8_t* T8,
                    cudaStream_t st);
int launch_zero_upper(double* Z, int64_t n, int64_t ld, cudaStream_t st);

// ---------------------------------------------------------------- K0 (pgen_expand.cu)
// .pgen mode-0x10 records -> packed 2-bit rows.  Entry e: record bytes blob[rec_off[e] .. +rec_len[e]), vrtype[e]; row codes:
// r >= 0 -> out + r * bpv, r < 0 -> scratch + (-r - 1) * bpv (LD bases that are not themselves part of the batch).
struct PgenExpandArgs {
    const uint8_t* blob; const uint32_t* rec_off; const uint32_t* rec_len; const uint8_t* vrtype;
    const int32_t* base_row;  // row of the LD base (vrtype & 7 in {2, 3}), unused otherwise
    const int32_t* out_row;
    int E; uint32_t sample_ct; int64_t bpv;
    uint8_t* out; uint8_t* scratch;
    int* err;                 // 0, or (entry << 4) | code of the first malformed record
    int any_ld;               // host hint: skip the second launch when no record is LD-compressed
    int pass, row_words;      // set by the launcher
    // dosage track (pgenlibr::Read, vrtype bits 5-6), optional
    uint32_t* main_end;       // out [E]: bytes of the record taken by the main genotype track (where the next track starts)
    uint16_t* dos_out;        // out: row r >= 0 at dos_out + r * dos_ld, one uint16 per raw sample (0..32768 = dosage 0..2); rows
    int64_t dos_ld;           //      must be pre-filled with 0xFFFF = "no dosage for this sample: use the hard call"
};
int launch_pgen_expand(const PgenExpandArgs& a, cudaStream_t st);
// after launch_pgen_expand (needs main_end): the dosage tracks of the batch's own records -> dos_out
int launch_pgen_dosage(const PgenExpandArgs& a, cudaStream_t st);

// ---------------------------------------------------------------- K1 (decode.cu)
struct DecodeArgs {
    const uint8_t* packed;    // device: 2-bit records of the batch's variants, record v at packed + v*bpv
    int64_t bpv;              // bytes per variant record
    const int32_t* vmap;      // device: record index of batch SNP q (nullptr => q)
    const int32_t* order;     // device: 0-based raw sample index of model sample i (nullptr => identity)
    int64_t n;                // model samples
    int KC;                   // k-chunks (n_pad/16)
    int64_t nsnp;             // SNPs in this batch
    int64_t snp_offset;       // index of the batch's first SNP in the run (for missing_flag)
    int s;                    // SNP-dependent terms
    const double* lut;        // device: [s][3][n_pad] term value by genotype code and sample (nullptr if all uniform)
    double uni[8][3];         // per term: value by code when the term is sample-independent
    uint32_t uniform_mask;    // bit t set: term t uses uni[t][code]
    int64_t n_pad;
    double* X;                // out: packed X tiles of the batch
    int NJ;                   // tiles in batch
    int* missing_flag;        // out: atomicMin(run-wide index of a SNP holding a missing call); INT_MAX = none
    uint8_t* codes_out;       // optional (tests): nsnp x n gathered codes
    const uint8_t* smask;     // K1q only: per model sample 0xFF keep / 0x00 zero (an x:factor indicator folded into the tile set); nullptr: none
    uint32_t gmap;            // K1q only: byte g = 2-bit code of genotype g in this tile set (tile_code_map)
    // pgenlibr::Read: x = dosage / 16384 where the record has one for the sample (row = the SNP's record row, like `packed`)
    const uint16_t* dosage;   // nullptr: hard calls only
    int64_t dos_ld;
    uint32_t dos_indicator;   // bit t: term t is base + (peak - base) [x == g_t] (I(x == 1)); else affine in x: v(0) + x (v(1) - v(0))
    int dos_g[8];             // g_t of the indicator terms
};
int launch_decode(const DecodeArgs& a, cudaStream_t st);
int launch_decode_i8(const DecodeArgs& a, uint32_t* X2, int NJ, int NKB, cudaStream_t st);
void tile_code_map(const int a[3], uint32_t* gmap, uint32_t* vt);   // integer triple of a tile set -> genotype -> code map, code -> int8 value table

// ---------------------------------------------------------------- K3 (snp_fit.cu)
struct FitConst {
    int n;                    // samples (N of logLik)
    int s, k, cr;             // SNP terms, design columns, rank of the constant columns
    int col_kind[24];         // model column j: >=0 -> constant column index (into Rc columns); <0 -> -(t+1) SNP term t
    int ncf;                  // number of constant columns
    double Rc[24 * 24];       // cr x ncf (column-major, ld = 24): coordinates of constant columns in the Qc basis
    double qcy[24];           // Qc' y
    double rss_f;             // || y - Qc Qc' y ||^2
    double lrs0;              // log(RSS0 / rss_f), RSS0 from the given ll_null
    int df_null;
};
struct GramArgs {
    const double* W;          // packed W tiles of the batch
    const double* Qc;         // device [n_pad][crp] row-major, crp = padded basis width (gram_padded_width), zero padded
    const double* r;          // device [n_pad] residual of y on Qc (zero padded)
    int KC; int cr; int crp; int s;
    int KS;                   // sample splits
    int cps;                  // k-chunks per split
    int64_t nsnp;
    // per split, per SNP partial sums (summed in fixed order by the consumers)
    double* n_part;           // [KS][nsnp][s]        ||w_t||^2
    double* u_part;           // [KS][nsnp][s][crp]   Qc' w_t
    double* a_part;           // [KS][nsnp][s(s+1)/2] residualised Gram
    double* b_part;           // [KS][nsnp][s]        w_r' r
};
int gram_padded_width(int cr);
int launch_snp_gram(const GramArgs& a, cudaStream_t st);

struct SolveArgs {
    const double* n_part; const double* u_part; const double* a_part; const double* b_part;
    int KS; int crp;
    int64_t nsnp;
    int64_t out_offset;       // first SNP of this batch in the run-wide outputs
    // outputs (run-wide arrays)
    double* chisq; double* pval; int32_t* df; int32_t* rank; int32_t* pivot; double* coef;  // pivot/coef: [M][k]
    // batch-local outputs for K4
    int32_t* snp_df; double* snp_sm;
    int* df_count;            // [9] run-wide count of SNPs per rank bucket 0..s (lrt_df = cr + 1 - df_null + bucket, counted when lrt_df > 0)
    const int32_t* out_map;   // optional: run-wide output slot of batch SNP q (refit pass), else out_offset + q
    // fast (int8 quadratic form) mode: A_raw[t][v] = S[slot_a[t][v]] + S[slot_b[t][v]], S[slot] = sum of the qf_nparts partials
    // of K2q pass `slot` (qf_part: [slot][qf_nparts][nsnp]); u_t (cr values) and b_t from K4q: u_part[t * u_term_stride + snp * u_ld + j]
    int64_t u_term_stride, u_ld;
    int fast;
    double refit_thr;         // re-fit in FP64 when a Cholesky pivot <= refit_thr * A_raw[t][t]
    const double* qf_part; int qf_nparts;
    int8_t slot_a[8][8], slot_b[8][8];   // slot_b < 0: single term
    int32_t* refit_flag;      // [nsnp] set to 1 when the SNP must be re-fitted on the FP64 path (near-collinear with the constant columns)
};
int launch_snp_solve(const SolveArgs& a, const FitConst& c, cudaStream_t st);

// ---------------------------------------------------------------- setup kernels (setup.cu)
// dense column-major (n x ncols, ld) -> packed panel with k = row index, j = column index (tiles of 128 columns)
int launch_pack_cols(const double* dense, int64_t n, int64_t ld, int64_t ncols, int KC, double* packed, cudaStream_t st);
// packed panel tiles -> dense column-major
int launch_unpack_cols(const double* packed, int64_t n, int64_t ld, int64_t ncols, int KC, double* dense, cudaStream_t st);
// dense lower-triangular inverse Z (n x n col-major) -> triangular packed A panels (k = column of Z, j = row of Z)
int launch_pack_tri(const double* Z, int64_t n, int64_t ld, int T, double* packed, cudaStream_t st);
// Y[:, p] = y_pred + UE[:, p]  then  Rf[:, p] = Y - Qc Qc' Y ; rss_f[p] = ||Rf||^2 ; rss0[p] = || Y - Qn Qn' Y ||^2
struct PermPrepArgs {
    double* UE;               // in: U1 * E (n x P col-major, ld = n_pad) ; out: residual Rf (in place)
    const double* y_pred;     // [n]
    const double* Qc; int cr; int crp; // [n_pad][crp] row-major
    const double* Qn; int cn; int cnp; // [n_pad][cnp] row-major (null model basis)
    int64_t n, n_pad; int P;
    double* rss_f; double* rss0;  // [P]
};
int launch_perm_prep(const PermPrepArgs& a, cudaStream_t st);
// Y[:, p] = Q [0 (r rows); e_p[idx_p]]  with Q = H_1 ... H_r given by reflectors V (n x r) and beta (implicit U1)
int launch_implicit_u1(const double* e_p, const int32_t* idx, int64_t m, int P, const double* V, const double* beta, int r, int64_t n, int64_t ldy, double* Y, cudaStream_t st);
// E[j, p] = e_p[idx[p*m + j] - 1]
int launch_gather_ep(const double* e_p, const int32_t* idx, int64_t m, int P, double* E, cudaStream_t st);
// V = a*V + b*I in place (decorrelate.nf:54)
int launch_scale_add_diag(double* V, int64_t n, int64_t ld, double a, double b, cudaStream_t st);
// list[0] = number of non-zero flags, list[1..] = their indices (any order); list holds M + 1 entries
int launch_compact_flags(const int32_t* flag, int64_t M, int32_t* list, cudaStream_t st);
// fill identity
int launch_set_identity(double* Z, int64_t n, int64_t ld, cudaStream_t st);
// finalize min-p: minp[p] = min over df of pchisq_upper(N*(lrs0p[p] - log1p(-maxratio[df][p])), df)
// int8 route, s > 1: bst [s][nsnp_ld][ldp] = b*_t(snp, perm) -> running max over SNPs of b*' Sm b* / RSS_f per (perm, df)
struct PermReduceArgs {
    const double* bst; int64_t term_stride; int64_t ldp;
    const int32_t* snp_df; const double* snp_sm; const double* perm_inv_rss; double* perm_maxratio;
    int64_t nsnp; int P; int64_t P_pad; int s;
};
int launch_perm_reduce(const PermReduceArgs& a, cudaStream_t st);
int launch_minp_finalize(const double* maxratio, int64_t P_pad, int P, int s, int df_off, const double* lrs0p, double N, const int* df_count, double* minp, cudaStream_t st);

}  // namespace flx
