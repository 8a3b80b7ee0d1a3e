// snp_fit.cu — K3: the per-SNP fit, LRT and chi-square tail (sm_100a).
//
// Replaces .lm.fit + stats:::logLik.lm + pchisq for every SNP (fit_model.nf:133-139) without ever
// forming the n x k whitened design:
//
//   K3a snp_gram   (HBM-bound, one CTA per 8-SNP group, quad-per-SNP shuffle reductions)
//        u_t = Qc' w_t,  ||w_t||^2                       (pass 1)
//        w_r,t = w_t - Qc u_t ;  A = W_r' W_r ;  b = W_r' r   (pass 2, explicit residualisation)
//        where Qc is an orthonormal basis of the whitened SNP-independent design columns and r the
//        residual of y.mm on them.  Algorithmic bytes: 8 n s per SNP per pass.
//   K3b snp_solve  (one thread per SNP, register/local k x k work)
//        Every whitened design column is a vector of coordinates in the orthonormal basis [Qc, Qs]
//        (Qs from the Cholesky of A), so the n x k least-squares problem equals a (cr+s) x k one.
//        On that small matrix the kernel runs the SAME algorithm the reference runs on the n x k
//        matrix: LINPACK dqrdc2 as modified for R (limited column pivoting, tol 1e-7, norm
//        down-dating) and dqrsl.  Column norms are preserved by the change of basis, so rank, pivot
//        and the pivoted coefficients follow the reference's rule, and
//            lrt_chisq = 2 (ll - ll.null) = N [ log(RSS0/RSS_f) - log1p(-dSS/RSS_f) ]
//        with dSS the drop in residual sum of squares due to the SNP columns.
//        It also emits, per SNP, the symmetric matrix Sm with dSS(y*) = b*' Sm b* for ANY phenotype
//        y* — what the permutation kernel K4 needs.
#include <cmath>
#include "common.cuh"
#include "kernels.h"

namespace flx {

// The per-task constants (FitConst, ~5 KB) travel as a __grid_constant__ kernel parameter (constant bank, like a __constant__
// symbol, but private to the launch): two handles on one device — the LOCO driver sets the next chromosome's factor up while
// the current one runs — never share mutable device state.

// ------------------------------------------------------------------------------------------------
// K3a  — two sweeps over the whitened tile, split along the sample axis for parallelism.
//   grid = (column tiles, KS sample splits); block = one warp per 8-SNP group of the tile.
//   lane = (SNP within group = lane >> 2, sample phase = lane & 3): the packed panel stores, for one k-step, the 8 SNPs x 4
//   samples of a group contiguously (256 B), so every warp load is two full 128-byte lines, and all warps of a CTA read
//   the same Qc / r rows (L1 hits).  Partial sums are written per split and summed in fixed order by the consumer
//   (deterministic, no atomics).
// ------------------------------------------------------------------------------------------------
template <int S, int CRP>
struct GramCfg {
    static constexpr bool UREG = (S * CRP <= 64);   // Qc' w coordinates live in registers
    static constexpr int GROUPS = 16 / S;
    static constexpr int THREADS = GROUPS * 32;
    static constexpr int NSYM = S * (S + 1) / 2;
};

// sweep 1: u = Qc' w (per split), ||w||^2
template <int S, int CRP>
__global__ void __launch_bounds__(GramCfg<S, CRP>::THREADS) snp_gram1_kernel(const GramArgs a) {
    using C = GramCfg<S, CRP>;
    const int J = blockIdx.x, ks = blockIdx.y;
    const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 2, kp = lane & 3;
    const int kc0 = ks * a.cps, kc1 = min(kc0 + a.cps, a.KC);
    const double* wt = a.W + (int64_t)J * a.KC * CHUNK + ((S * g) << 5) + lane;
    const int64_t snp = (int64_t)J * (8 * C::GROUPS) + 8 * g + grp;

    double nrm[S];
#pragma unroll
    for (int t = 0; t < S; t++) nrm[t] = 0.0;
    constexpr int JB = C::UREG ? CRP : 8;
    for (int jb = 0; jb < CRP; jb += JB) {
        double acc[S][JB];
#pragma unroll
        for (int t = 0; t < S; t++)
#pragma unroll
            for (int j = 0; j < JB; j++) acc[t][j] = 0.0;
#pragma unroll 2
        for (int kc = kc0; kc < kc1; kc++) {
            const double* wc = wt + (int64_t)kc * CHUNK;
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) {
                const int64_t k = (int64_t)kc * KCH + 4 * k4 + kp;
                double w[S];
#pragma unroll
                for (int t = 0; t < S; t++) w[t] = wc[(k4 * 16 + t) << 5];
                if (jb == 0) {
#pragma unroll
                    for (int t = 0; t < S; t++) nrm[t] = fma(w[t], w[t], nrm[t]);
                }
                const double2* q = reinterpret_cast<const double2*>(a.Qc + k * CRP + jb);
#pragma unroll
                for (int j = 0; j < JB; j += 2) {
                    if (jb + j < CRP) {
                        const double2 qq = __ldg(q + (j >> 1));
#pragma unroll
                        for (int t = 0; t < S; t++) {
                            acc[t][j] = fma(w[t], qq.x, acc[t][j]);
                            acc[t][j + 1] = fma(w[t], qq.y, acc[t][j + 1]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < S; t++)
#pragma unroll
            for (int j = 0; j < JB; j++) {
                double v = acc[t][j];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (kp == 0 && jb + j < CRP && snp < a.nsnp) a.u_part[(((int64_t)ks * a.nsnp + snp) * S + t) * CRP + jb + j] = v;
            }
    }
#pragma unroll
    for (int t = 0; t < S; t++) {
        double v = nrm[t];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (kp == 0 && snp < a.nsnp) a.n_part[((int64_t)ks * a.nsnp + snp) * S + t] = v;
    }
}

// sweep 2: w_r = w - Qc u (explicit residualisation), A = W_r' W_r, b = W_r' r  (per split)
template <int S, int CRP>
__global__ void __launch_bounds__(GramCfg<S, CRP>::THREADS) snp_gram2_kernel(const GramArgs a) {
    using C = GramCfg<S, CRP>;
    constexpr int NSYM = C::NSYM;
    const int J = blockIdx.x, ks = blockIdx.y;
    const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 2, kp = lane & 3;
    const int kc0 = ks * a.cps, kc1 = min(kc0 + a.cps, a.KC);
    const double* wt = a.W + (int64_t)J * a.KC * CHUNK + ((S * g) << 5) + lane;
    const int64_t snp = (int64_t)J * (8 * C::GROUPS) + 8 * g + grp;
    const bool live = snp < a.nsnp;

    extern __shared__ __align__(16) double u_sm[];   // only when !UREG: [GROUPS][8][S][CRP]
    double u[C::UREG ? S : 1][C::UREG ? CRP : 1];
    // total u over the splits, fixed order
    if constexpr (C::UREG) {
#pragma unroll
        for (int t = 0; t < S; t++)
#pragma unroll
            for (int j = 0; j < CRP; j++) {
                double v = 0.0;
                if (live)
                    for (int q = 0; q < a.KS; q++) v += a.u_part[(((int64_t)q * a.nsnp + snp) * S + t) * CRP + j];
                u[t][j] = v;
            }
    } else {
        if (kp == 0)
            for (int e = 0; e < S * CRP; e++) {
                double v = 0.0;
                if (live)
                    for (int q = 0; q < a.KS; q++) v += a.u_part[((int64_t)q * a.nsnp + snp) * S * CRP + e];
                u_sm[((g * 8 + grp) * S) * CRP + e] = v;
            }
        __syncwarp();
    }
    const double* ug = u_sm + ((g * 8 + grp) * S) * CRP;

    double ag[NSYM], bb[S];
#pragma unroll
    for (int q = 0; q < NSYM; q++) ag[q] = 0.0;
#pragma unroll
    for (int t = 0; t < S; t++) bb[t] = 0.0;
#pragma unroll 2
    for (int kc = kc0; kc < kc1; kc++) {
        const double* wc = wt + (int64_t)kc * CHUNK;
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) {
            const int64_t k = (int64_t)kc * KCH + 4 * k4 + kp;
            double w[S];
#pragma unroll
            for (int t = 0; t < S; t++) w[t] = wc[(k4 * 16 + t) << 5];
            const double2* q = reinterpret_cast<const double2*>(a.Qc + k * CRP);
#pragma unroll
            for (int j = 0; j < CRP; j += 2) {
                const double2 qq = __ldg(q + (j >> 1));
#pragma unroll
                for (int t = 0; t < S; t++) {
                    if constexpr (C::UREG) {
                        w[t] = fma(-qq.x, u[t][j], w[t]);
                        w[t] = fma(-qq.y, u[t][j + 1], w[t]);
                    } else {
                        const double2 uu = *reinterpret_cast<const double2*>(ug + t * CRP + j);
                        w[t] = fma(-qq.x, uu.x, w[t]);
                        w[t] = fma(-qq.y, uu.y, w[t]);
                    }
                }
            }
            const double rk = __ldg(a.r + k);
            int qi = 0;
#pragma unroll
            for (int t = 0; t < S; t++) {
                bb[t] = fma(w[t], rk, bb[t]);
#pragma unroll
                for (int v = t; v < S; v++) { ag[qi] = fma(w[t], w[v], ag[qi]); qi++; }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NSYM; q++) {
        double v = ag[q];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (kp == 0 && live) a.a_part[((int64_t)ks * a.nsnp + snp) * NSYM + q] = v;
    }
#pragma unroll
    for (int t = 0; t < S; t++) {
        double v = bb[t];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (kp == 0 && live) a.b_part[((int64_t)ks * a.nsnp + snp) * S + t] = v;
    }
}

template <int S, int CRP>
static int launch_gram_sc(const GramArgs& a, cudaStream_t st) {
    using C = GramCfg<S, CRP>;
    const int64_t NJ = (a.nsnp + 8 * C::GROUPS - 1) / (8 * C::GROUPS);
    dim3 grid((unsigned)NJ, (unsigned)a.KS);
    snp_gram1_kernel<S, CRP><<<grid, C::THREADS, 0, st>>>(a);
    const size_t smem = C::UREG ? 0 : sizeof(double) * C::GROUPS * 8 * S * CRP;
    snp_gram2_kernel<S, CRP><<<grid, C::THREADS, smem, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("snp_gram launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

template <int S>
static int launch_gram_s(const GramArgs& a, cudaStream_t st) {
    switch (a.crp) {
        case 4: return launch_gram_sc<S, 4>(a, st);
        case 8: return launch_gram_sc<S, 8>(a, st);
        case 16: return launch_gram_sc<S, 16>(a, st);
        case 24: return launch_gram_sc<S, 24>(a, st);
    }
    set_error("snp_gram: unsupported padded basis width %d", a.crp);
    return -1;
}

int gram_padded_width(int cr) { return cr <= 4 ? 4 : cr <= 8 ? 8 : cr <= 16 ? 16 : 24; }

int launch_snp_gram(const GramArgs& a, cudaStream_t st) {
    if (a.nsnp <= 0) return 0;
    switch (a.s) {
        case 1: return launch_gram_s<1>(a, st);
        case 2: return launch_gram_s<2>(a, st);
        case 3: return launch_gram_s<3>(a, st);
        case 4: return launch_gram_s<4>(a, st);
        case 5: return launch_gram_s<5>(a, st);
        case 6: return launch_gram_s<6>(a, st);
        case 7: return launch_gram_s<7>(a, st);
        case 8: return launch_gram_s<8>(a, st);
    }
    set_error("snp_gram: unsupported s=%d", a.s);
    return -1;
}

// ------------------------------------------------------------------------------------------------
// chi-square upper tail for integer df (pchisq(q, df, lower.tail = FALSE), fit_model.nf:138)
// ------------------------------------------------------------------------------------------------
__device__ double chisq_upper_tail(double q, int df) {
    if (df <= 0 || isnan(q)) return nan("");
    if (q <= 0.0) return 1.0;
    const double h = 0.5 * q;
    if ((df & 1) == 0) {
        double term = 1.0, sum = 1.0;
        for (int i = 1; i < df / 2; i++) { term *= h / i; sum += term; }
        return exp(-h) * sum;
    }
    const double tail = erfc(sqrt(h));
    const int j = (df - 1) / 2;
    if (j == 0) return tail;
    double term = 1.0, sum = 1.0;
    for (int i = 1; i < j; i++) { term *= q / (2 * i + 1); sum += term; }
    return tail + exp(-h) * sqrt(2.0 * q / 3.14159265358979323846) * sum;
}

// ------------------------------------------------------------------------------------------------
// K3b
// ------------------------------------------------------------------------------------------------
constexpr int MAXK = 24;
constexpr int MAXS = 8;
constexpr int LDT = MAXK + 1;

// Apply Householder reflector l (stored LINPACK style: v = column l of T from row l, with T[l][l] replaced by qraux[l])
__device__ __forceinline__ void apply_reflector(const double* T, const double* qraux, int l, int nrow, double* v) {
    if (qraux[l] == 0.0) return;
    const double* x = T + l * LDT;
    double t = qraux[l] * v[l];
    for (int i = l + 1; i < nrow; i++) t += x[i] * v[i];
    t = -t / qraux[l];
    v[l] += t * qraux[l];
    for (int i = l + 1; i < nrow; i++) v[i] += t * x[i];
}

__global__ void __launch_bounds__(64) snp_solve_kernel(const __grid_constant__ SolveArgs a, const __grid_constant__ FitConst c) {
    const int64_t snp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (snp >= a.nsnp) return;
    const int s = c.s, k = c.k, cr = c.cr;
    const int nr = cr + s;
    const int nrow = nr + 1;  // one zero row: the small problem must behave like n > p (dqrdc2's l == n exit never fires)
    const int nsym = s * (s + 1) / 2;

    double A[MAXS][MAXS], Rs[MAXS][MAXS], Mq[MAXS][MAXS];
    double bq[MAXS], qsr[MAXS], nrm2[MAXS];
    bool alive[MAXS];
    bool refit = false;
    if (a.fast) {
        // A_raw[t][v] = g_t' V^-1 g_v from the int8 passes (K2q), u_t = (L^-T Qc)' g_t and b_t = (L^-T r)' g_t from K4q
        double S[16];    // at most 12 forms: two scales x two triples on either side
        int nslots = 0;
        for (int t = 0; t < s; t++)
            for (int v = t; v < s; v++) { nslots = max(nslots, (int)a.slot_a[t][v] + 1); nslots = max(nslots, (int)a.slot_b[t][v] + 1); }
        for (int sl = 0; sl < nslots; sl++) {
            // fixed summation order, eight loads in flight per thread (the serial version is latency bound)
            const double* __restrict__ qp = a.qf_part + (int64_t)sl * a.qf_nparts * a.nsnp + snp;
            double acc = 0.0;
            int p = 0;
            for (; p + 8 <= a.qf_nparts; p += 8) {
                double v[8];
#pragma unroll
                for (int e = 0; e < 8; e++) v[e] = __ldg(qp + (int64_t)(p + e) * a.nsnp);
#pragma unroll
                for (int e = 0; e < 8; e++) acc += v[e];
            }
            for (; p < a.qf_nparts; p++) acc += __ldg(qp + (int64_t)p * a.nsnp);
            S[sl] = acc;
        }
        const double* __restrict__ ub = a.u_part + snp * a.u_ld;        // term t at + t * u_term_stride: u_t (cr values) then b_t
        for (int t = 0; t < s; t++) {
            bq[t] = __ldg(ub + t * a.u_term_stride + cr);
            for (int v = t; v < s; v++) {
                double uu = 0.0;
                for (int j = 0; j < cr; j++) uu += __ldg(ub + t * a.u_term_stride + j) * __ldg(ub + v * a.u_term_stride + j);
                const double araw = S[a.slot_a[t][v]] + (a.slot_b[t][v] >= 0 ? S[a.slot_b[t][v]] : 0.0);
                if (v == t) nrm2[t] = araw;
                A[t][v] = A[v][t] = araw - uu;
            }
        }
    } else {
        // partial sums of the sample splits, fixed order
        int q = 0;
        for (int t = 0; t < s; t++)
            for (int v = t; v < s; v++) {
                double acc = 0.0;
                for (int p = 0; p < a.KS; p++) acc += a.a_part[((int64_t)p * a.nsnp + snp) * nsym + q];
                A[t][v] = A[v][t] = acc;
                q++;
            }
        for (int t = 0; t < s; t++) {
            double ab = 0.0, an = 0.0;
            for (int p = 0; p < a.KS; p++) { ab += a.b_part[((int64_t)p * a.nsnp + snp) * s + t]; an += a.n_part[((int64_t)p * a.nsnp + snp) * s + t]; }
            bq[t] = ab; nrm2[t] = an;
        }
    }
    // guarded Cholesky A = Rs' Rs in term order; a vanished pivot leaves a zero row (direction dropped)
    for (int l = 0; l < s; l++) {
        double d = A[l][l];
        for (int i = 0; i < l; i++) d -= Rs[i][l] * Rs[i][l];
        alive[l] = (d > 0.0) && (d > 1e-22 * nrm2[l]);
        if (a.fast) {
            // A_raw - U U' and the pivots lose log2(A_raw / d) bits of the 2^-48 the digit planes resolve: hand near-collinear
            // SNPs (monomorphic, ...) to the FP64 path.  A term that DUPLICATES an earlier one (x and I(x==1) of a SNP without
            // hom-alt calls) has bit-identical int8 sums, its pivot is rounding noise of the Cholesky itself: dropped here.
            if (l > 0 && A[l][l] > a.refit_thr * nrm2[l] && d <= 1e-15 * A[l][l]) alive[l] = false;
            else if (!(d > a.refit_thr * nrm2[l])) refit = true;
        }
        for (int j = 0; j < s; j++) Rs[l][j] = 0.0;
        if (alive[l]) {
            const double rll = sqrt(d);
            Rs[l][l] = rll;
            for (int j = l + 1; j < s; j++) {
                double v = A[l][j];
                for (int i = 0; i < l; i++) v -= Rs[i][l] * Rs[i][j];
                Rs[l][j] = v / rll;
            }
        }
    }
    if (a.refit_flag) a.refit_flag[snp] = refit ? 1 : 0;
    // Mq: q = Rs^-T b as a linear map (lower triangular), qsr = Mq b
    for (int l = 0; l < s; l++) {
        for (int cidx = 0; cidx < s; cidx++) Mq[l][cidx] = 0.0;
        if (alive[l]) {
            for (int cidx = 0; cidx <= l; cidx++) {
                double v = (cidx == l) ? 1.0 : 0.0;
                for (int i = cidx; i < l; i++) v -= Rs[i][l] * Mq[i][cidx];
                Mq[l][cidx] = v / Rs[l][l];
            }
        }
    }
    double qq = 0.0;
    for (int l = 0; l < s; l++) {
        double v = 0.0;
        for (int cidx = 0; cidx <= l; cidx++) v += Mq[l][cidx] * bq[cidx];
        qsr[l] = v;
        qq += v * v;
    }

    // ---- the small design T (nrow x k, column-major, ld LDT) and response ty
    double T[MAXK * LDT];
    double ty[LDT];
    for (int j = 0; j < k; j++) {
        double* col = T + j * LDT;
        for (int i = 0; i < nrow; i++) col[i] = 0.0;
        const int kind = c.col_kind[j];
        if (kind >= 0) {
            for (int i = 0; i < cr; i++) col[i] = c.Rc[i + kind * MAXK];
        } else {
            const int t = -kind - 1;
            for (int i = 0; i < cr; i++) {
                double acc = 0.0;
                if (a.fast) acc = a.u_part[t * a.u_term_stride + snp * a.u_ld + i];
                else for (int p = 0; p < a.KS; p++) acc += a.u_part[(((int64_t)p * a.nsnp + snp) * s + t) * a.crp + i];
                col[i] = acc;
            }
            for (int i = 0; i <= t; i++) col[cr + i] = Rs[i][t];
        }
    }
    for (int i = 0; i < cr; i++) ty[i] = c.qcy[i];
    for (int i = 0; i < s; i++) ty[cr + i] = qsr[i];
    ty[nr] = 0.0;

    // ---- dqrdc2 (R's LINPACK variant): limited column pivoting with tol = 1e-7
    const double tol = 1e-7;
    double qraux[MAXK], w1[MAXK], w2[MAXK];
    int jpvt[MAXK];
    for (int j = 0; j < k; j++) {
        double ss = 0.0;
        for (int i = 0; i < nrow; i++) ss += T[i + j * LDT] * T[i + j * LDT];
        qraux[j] = sqrt(ss);
        w1[j] = qraux[j];
        w2[j] = (qraux[j] == 0.0) ? 1.0 : qraux[j];
        jpvt[j] = j + 1;
    }
    const int lup = (nrow < k) ? nrow : k;
    int kk = k + 1;
    for (int l = 1; l <= lup; l++) {
        while (l < kk && qraux[l - 1] < w2[l - 1] * tol) {
            // rotate column l to the end
            for (int i = 0; i < nrow; i++) {
                const double t = T[i + (l - 1) * LDT];
                for (int j = l; j < k; j++) T[i + (j - 1) * LDT] = T[i + j * LDT];
                T[i + (k - 1) * LDT] = t;
            }
            const int ii = jpvt[l - 1];
            const double t0 = qraux[l - 1], t1 = w1[l - 1], t2 = w2[l - 1];
            for (int j = l; j < k; j++) { jpvt[j - 1] = jpvt[j]; qraux[j - 1] = qraux[j]; w1[j - 1] = w1[j]; w2[j - 1] = w2[j]; }
            jpvt[k - 1] = ii; qraux[k - 1] = t0; w1[k - 1] = t1; w2[k - 1] = t2;
            kk--;
        }
        if (l == nrow) break;
        double* xl = T + (l - 1) * LDT;
        double ss = 0.0;
        for (int i = l - 1; i < nrow; i++) ss += xl[i] * xl[i];
        double nrmxl = sqrt(ss);
        if (nrmxl == 0.0) continue;
        if (xl[l - 1] != 0.0) nrmxl = copysign(nrmxl, xl[l - 1]);
        const double sc = 1.0 / nrmxl;
        for (int i = l - 1; i < nrow; i++) xl[i] *= sc;
        xl[l - 1] += 1.0;
        for (int j = l + 1; j <= k; j++) {
            double* xj = T + (j - 1) * LDT;
            double dot = 0.0;
            for (int i = l - 1; i < nrow; i++) dot += xl[i] * xj[i];
            const double t = -dot / xl[l - 1];
            for (int i = l - 1; i < nrow; i++) xj[i] += t * xl[i];
            if (qraux[j - 1] != 0.0) {
                const double r = fabs(xj[l - 1]) / qraux[j - 1];
                double tt = 1.0 - r * r;
                if (tt < 0.0) tt = 0.0;
                if (fabs(tt) < 1e-6) {
                    double s2 = 0.0;
                    for (int i = l; i < nrow; i++) s2 += xj[i] * xj[i];
                    qraux[j - 1] = sqrt(s2);
                    w1[j - 1] = qraux[j - 1];
                } else {
                    qraux[j - 1] *= sqrt(tt);
                }
            }
        }
        qraux[l - 1] = xl[l - 1];
        xl[l - 1] = -nrmxl;
    }
    int rank = kk - 1;
    if (rank > nrow) rank = nrow;

    // ---- dqrsl: qty, coefficients (pivoted order)
    const int ju = (rank < nrow - 1) ? rank : nrow - 1;
    for (int l = 0; l < ju; l++) apply_reflector(T, qraux, l, nrow, ty);
    double coef[MAXK];
    for (int j = 0; j < k; j++) coef[j] = (j < rank) ? ty[j] : 0.0;
    for (int j = rank - 1; j >= 0; j--) {
        const double d = T[j + j * LDT];
        if (d == 0.0) break;
        coef[j] /= d;
        const double t = -coef[j];
        for (int i = 0; i < j; i++) coef[i] += t * T[i + j * LDT];
    }
    double rss_small = 0.0;
    for (int i = rank; i < nr; i++) rss_small += ty[i] * ty[i];

    // ---- LRT
    double dss = qq - rss_small;
    if (dss < 0.0) dss = 0.0;
    const double ratio = dss / c.rss_f;
    const double N = (double)c.n;
    const double chisq = N * (c.lrs0 - log1p(-ratio));
    const int df = rank + 1 - c.df_null;
    const double p = (df > 0) ? chisq_upper_tail(chisq, df) : nan("");
    // permutation maxima are kept per lrt_df; the SNP columns add bucket = 0..s to the rank cr of the constant columns, so
    // lrt_df = (cr + 1 - df_null) + bucket whatever the null model is (a null nested more than s columns deep has lrt_df > s).
    // bucket 0 (no SNP direction survives) has dSS = 0 for every phenotype: counted, never reduced.
    const int bucket = min(max(rank - cr, 0), s);

    const int64_t o = a.out_map ? (int64_t)a.out_map[snp] : a.out_offset + snp;
    a.chisq[o] = chisq;
    a.pval[o] = p;
    a.df[o] = df;
    a.rank[o] = rank;
    for (int j = 0; j < k; j++) {
        a.pivot[o * k + j] = jpvt[j];
        a.coef[o * k + j] = coef[j];
    }

    // ---- Sm for the permutation kernel: dSS(y*) = q*' (I - N N') q*, q* = Mq b*, N = rows >= rank of H' restricted to the Qs block
    if (a.snp_sm) {
        double G[MAXS][MAXS];
        for (int i = 0; i < s; i++)
            for (int j = 0; j < s; j++) G[i][j] = (i == j) ? 1.0 : 0.0;
        if (rank < nr) {
            double Nm[MAXS][MAXS + MAXK];  // [i][row - rank]
            const int nn = nr - rank;
            for (int i = 0; i < s; i++) {
                double v[LDT];
                for (int r2 = 0; r2 < nrow; r2++) v[r2] = 0.0;
                v[cr + i] = 1.0;
                for (int l = 0; l < ju; l++) apply_reflector(T, qraux, l, nrow, v);
                for (int r2 = 0; r2 < nn; r2++) Nm[i][r2] = v[rank + r2];
            }
            for (int i = 0; i < s; i++)
                for (int j = 0; j < s; j++) {
                    double v = 0.0;
                    for (int r2 = 0; r2 < nn; r2++) v += Nm[i][r2] * Nm[j][r2];
                    G[i][j] -= v;
                }
        }
        // Sm = Mq' G Mq
        double GM[MAXS][MAXS];
        for (int i = 0; i < s; i++)
            for (int j = 0; j < s; j++) {
                double v = 0.0;
                for (int l = 0; l < s; l++) v += G[i][l] * Mq[l][j];
                GM[i][j] = v;
            }
        int q = 0;
        for (int t = 0; t < s; t++)
            for (int v2 = t; v2 < s; v2++) {
                double v = 0.0;
                for (int l = 0; l < s; l++) v += Mq[l][t] * GM[l][v2];
                a.snp_sm[snp * nsym + q] = (df > 0 && !refit) ? v : 0.0;
                q++;
            }
        a.snp_df[snp] = (df > 0 && !refit) ? bucket : 0;
    }
    if (a.df_count && df > 0 && !refit) atomicAdd(&a.df_count[bucket], 1);
}

int launch_snp_solve(const SolveArgs& a, const FitConst& c, cudaStream_t st) {
    if (a.nsnp <= 0) return 0;
    const int threads = 64;
    snp_solve_kernel<<<(unsigned)((a.nsnp + threads - 1) / threads), threads, 0, st>>>(a, c);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("snp_solve launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// min-p finalisation: one thread per permutation
// ------------------------------------------------------------------------------------------------
// int8 route with s > 1 terms: K4q stores b*_t = g_t' (L^-T r*_p) per term; this pass forms b*' Sm b* / RSS_f per (SNP, permutation)
// and keeps the running max over SNPs per (df, permutation).  HBM bound: one read of s * nsnp * P doubles.
constexpr int PR_LANES = 8;        // SNP lanes per block (x 128 permutation lanes)
constexpr int PR_SNPS = 512;       // SNPs per block
template <int S>
__global__ void __launch_bounds__(128 * PR_LANES) perm_reduce_kernel(const PermReduceArgs a) {
    const int pl = threadIdx.x & 127, y = threadIdx.x >> 7;
    const int p = blockIdx.x * 128 + pl;
    const int64_t q0 = (int64_t)blockIdx.y * PR_SNPS;
    const int64_t q1 = min(q0 + PR_SNPS, a.nsnp);
    constexpr int s = S, nsym = S * (S + 1) / 2;      // compile-time: z, mx and the quadratic form stay in registers
    __shared__ double red[PR_LANES][128];
    const bool live = p < a.P;
    const double inv = live ? a.perm_inv_rss[p] : 0.0;
    double mx[S];
#pragma unroll
    for (int t = 0; t < s; t++) mx[t] = 0.0;
    // a warp reads one contiguous run of permutations of one SNP; Sm and df are warp-uniform (broadcast loads); no
    // data-dependent branch, so the loads of several SNPs are in flight at once
    if (live) {
#pragma unroll 4
        for (int64_t q = q0 + y; q < q1; q += PR_LANES) {
            const int df = __ldg(a.snp_df + q);
            double z[S];
#pragma unroll
            for (int t = 0; t < s; t++) z[t] = __ldg(a.bst + (int64_t)t * a.term_stride + q * a.ldp + p);
            double quad = 0.0;
            int c = 0;
#pragma unroll
            for (int t = 0; t < s; t++)
#pragma unroll
                for (int u = t; u < s; u++) { quad += ((u == t) ? 1.0 : 2.0) * __ldg(a.snp_sm + q * nsym + c) * z[t] * z[u]; c++; }
            const double ratio = fmax(quad * inv, 0.0);      // Sm is zero for df = 0 and for SNPs handed to the FP64 refit
#pragma unroll
            for (int t = 0; t < s; t++) mx[t] = (df == t + 1) ? fmax(mx[t], ratio) : mx[t];
        }
    }
    // one atomic per (permutation, df) per block: with 148 blocks per batch the hot addresses see little contention
#pragma unroll
    for (int t = 0; t < s; t++) {
        red[y][pl] = mx[t];
        __syncthreads();
        if (y == 0 && live) {
            double v = red[0][pl];
            for (int l = 1; l < PR_LANES; l++) v = fmax(v, red[l][pl]);
            if (v > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(&a.perm_maxratio[(int64_t)t * a.P_pad + p]), (unsigned long long)__double_as_longlong(v));
        }
        __syncthreads();
    }
}

int launch_perm_reduce(const PermReduceArgs& a, cudaStream_t st) {
    if (a.P <= 0 || a.nsnp <= 0) return 0;
    const dim3 grid((unsigned)((a.P + 127) / 128), (unsigned)((a.nsnp + PR_SNPS - 1) / PR_SNPS));
    switch (a.s) {
        case 1: perm_reduce_kernel<1><<<grid, 128 * PR_LANES, 0, st>>>(a); break;
        case 2: perm_reduce_kernel<2><<<grid, 128 * PR_LANES, 0, st>>>(a); break;
        case 3: perm_reduce_kernel<3><<<grid, 128 * PR_LANES, 0, st>>>(a); break;
        case 4: perm_reduce_kernel<4><<<grid, 128 * PR_LANES, 0, st>>>(a); break;
        default: set_error("perm_reduce: s = %d terms not supported on the int8 route", a.s); return -1;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("perm_reduce launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

// bucket b = 0..s <-> lrt_df = df_off + b (snp_solve_kernel); maxratio rows are the buckets 1..s, bucket 0 has ratio 0
__global__ void minp_finalize_kernel(const double* maxratio, int64_t P_pad, int P, int s, int df_off, const double* lrs0p, double N, const int* df_count, double* minp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double best = 2.0;  // "an impossibly large value" (fit_model.nf:104)
    for (int b = 0; b <= s; b++) {
        if (df_count[b] == 0 || df_off + b <= 0) continue;  // no SNP with this df in the run
        const double mr = (b == 0) ? 0.0 : maxratio[(int64_t)(b - 1) * P_pad + p];
        const double chisq = N * (lrs0p[p] - log1p(-mr));
        const double pv = chisq_upper_tail(chisq, df_off + b);
        if (pv < best) best = pv;
    }
    minp[p] = best;
}

int launch_minp_finalize(const double* maxratio, int64_t P_pad, int P, int s, int df_off, const double* lrs0p, double N, const int* df_count, double* minp, cudaStream_t st) {
    if (P <= 0) return 0;
    minp_finalize_kernel<<<(P + 127) / 128, 128, 0, st>>>(maxratio, P_pad, P, s, df_off, lrs0p, N, df_count, minp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("minp_finalize launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

}  // namespace flx
