// setup.cu — once-per-task kernels (outside the timed hot loop): panel packing and the permutation prologue
// (fit_model.nf:96-101: y.mm <- y.mm.pred + U1 %*% sample(e.p); null refit).
#include "common.cuh"
#include "kernels.h"

namespace flx {

// ---- dense (n x ncols, col-major) -> packed panel tiles, k = row, j = column
__global__ void pack_cols_kernel(const double* __restrict__ dense, int64_t n, int64_t ld, int64_t ncols, int KC, double* __restrict__ packed) {
    const int64_t J = blockIdx.y;
    const int kc = blockIdx.x;
    double* out = packed + (J * KC + kc) * CHUNK;
    for (int e = threadIdx.x; e < CHUNK; e += blockDim.x) {
        const int ks = e >> 9, blk = (e >> 5) & 15, lane = e & 31;
        const int64_t k = (int64_t)kc * KCH + 4 * ks + (lane & 3);
        const int64_t j = J * TJ + 8 * blk + (lane >> 2);
        out[e] = (k < n && j < ncols) ? dense[k + j * ld] : 0.0;
    }
}
int launch_pack_cols(const double* dense, int64_t n, int64_t ld, int64_t ncols, int KC, double* packed, cudaStream_t st) {
    const int64_t NJ = (ncols + TJ - 1) / TJ;
    if (NJ <= 0) return 0;
    pack_cols_kernel<<<dim3((unsigned)KC, (unsigned)NJ), 256, 0, st>>>(dense, n, ld, ncols, KC, packed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("pack_cols launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

__global__ void unpack_cols_kernel(const double* __restrict__ packed, int64_t n, int64_t ld, int64_t ncols, int KC, double* __restrict__ dense) {
    const int64_t J = blockIdx.y;
    const int kc = blockIdx.x;
    const double* in = packed + (J * KC + kc) * CHUNK;
    for (int e = threadIdx.x; e < CHUNK; e += blockDim.x) {
        const int ks = e >> 9, blk = (e >> 5) & 15, lane = e & 31;
        const int64_t k = (int64_t)kc * KCH + 4 * ks + (lane & 3);
        const int64_t j = J * TJ + 8 * blk + (lane >> 2);
        if (k < n && j < ncols) dense[k + j * ld] = in[e];
    }
}
int launch_unpack_cols(const double* packed, int64_t n, int64_t ld, int64_t ncols, int KC, double* dense, cudaStream_t st) {
    const int64_t NJ = (ncols + TJ - 1) / TJ;
    if (NJ <= 0) return 0;
    unpack_cols_kernel<<<dim3((unsigned)KC, (unsigned)NJ), 256, 0, st>>>(packed, n, ld, ncols, KC, dense);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("unpack_cols launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

// ---- dense lower-triangular Z = L^-1 -> triangular A panels: P[k][j] = Z[128 I + j][k], k <= row
__global__ void pack_tri_kernel(const double* __restrict__ Z, int64_t n, int64_t ld, double* __restrict__ packed) {
    const int I = blockIdx.y;
    const int kc = blockIdx.x;
    if (kc >= tri_chunks_of(I)) return;
    double* out = packed + (tri_chunks_before(I) + kc) * CHUNK;
    for (int e = threadIdx.x; e < CHUNK; e += blockDim.x) {
        const int ks = e >> 9, blk = (e >> 5) & 15, lane = e & 31;
        const int64_t k = (int64_t)kc * KCH + 4 * ks + (lane & 3);
        const int64_t row = (int64_t)I * TJ + 8 * blk + (lane >> 2);
        out[e] = (row < n && k <= row) ? Z[row + k * ld] : 0.0;
    }
}
int launch_pack_tri(const double* Z, int64_t n, int64_t ld, int T, double* packed, cudaStream_t st) {
    if (T <= 0) return 0;
    pack_tri_kernel<<<dim3((unsigned)(8 * T), (unsigned)T), 256, 0, st>>>(Z, n, ld, packed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("pack_tri launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

__global__ void set_identity_kernel(double* Z, int64_t n, int64_t ld) {
    const int64_t j = blockIdx.x;
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.y * blockDim.x) Z[i + j * ld] = (i == j) ? 1.0 : 0.0;
}
int launch_set_identity(double* Z, int64_t n, int64_t ld, cudaStream_t st) {
    if (n <= 0) return 0;
    const unsigned gx = (unsigned)((n + 255) / 256 > 64 ? 64 : (n + 255) / 256);
    set_identity_kernel<<<dim3((unsigned)n, gx), 256, 0, st>>>(Z, n, ld);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("set_identity launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

__global__ void scale_add_diag_kernel(double* V, int64_t n, int64_t ld, double a, double b) {
    const int64_t j = blockIdx.x;
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.y * blockDim.x) {
        const double v = a * V[i + j * ld];
        V[i + j * ld] = (i == j) ? v + b : v;
    }
}
int launch_scale_add_diag(double* V, int64_t n, int64_t ld, double a, double b, cudaStream_t st) {
    if (n <= 0) return 0;
    const unsigned gy = (unsigned)((n + 255) / 256 > 64 ? 64 : (n + 255) / 256);
    scale_add_diag_kernel<<<dim3((unsigned)n, gy), 256, 0, st>>>(V, n, ld, a, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("scale_add_diag launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

// ---- E[j, p] = e_p[idx[p*m + j] - 1]   (sample(e.p), fit_model.nf:98)
__global__ void gather_ep_kernel(const double* __restrict__ e_p, const int32_t* __restrict__ idx, int64_t m, double* __restrict__ E) {
    const int64_t p = blockIdx.x;
    for (int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; j < m; j += (int64_t)gridDim.y * blockDim.x)
        E[j + p * m] = e_p[idx[p * m + j] - 1];
}
int launch_gather_ep(const double* e_p, const int32_t* idx, int64_t m, int P, double* E, cudaStream_t st) {
    if (P <= 0 || m <= 0) return 0;
    const unsigned gx = (unsigned)((m + 255) / 256 > 128 ? 128 : (m + 255) / 256);
    gather_ep_kernel<<<dim3((unsigned)P, gx), 256, 0, st>>>(e_p, idx, m, E);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("gather_ep launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

// ---- implicit U1: Y[:, p] = H_1 ... H_r [0; e_p[idx_p]]   (one CTA per permutation column)
__global__ void __launch_bounds__(256) implicit_u1_kernel(const double* __restrict__ e_p, const int32_t* __restrict__ idx, int64_t m,
                                                          const double* __restrict__ V, const double* __restrict__ beta, int r, int64_t n, int64_t ldy,
                                                          double* __restrict__ Y) {
    __shared__ double sh[8];
    const int64_t p = blockIdx.x;
    double* col = Y + p * ldy;
    for (int64_t i = threadIdx.x; i < n; i += 256) col[i] = (i < r) ? 0.0 : e_p[idx[p * m + (i - r)] - 1];
    __syncthreads();
    for (int j = r - 1; j >= 0; --j) {
        const double* v = V + (int64_t)j * n;
        double t = 0.0;
        for (int64_t i = j + threadIdx.x; i < n; i += 256) t += v[i] * col[i];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
        __syncthreads();
        double tot = 0.0;
        for (int w = 0; w < 8; w++) tot += sh[w];
        tot *= beta[j];
        for (int64_t i = j + threadIdx.x; i < n; i += 256) col[i] -= tot * v[i];
        __syncthreads();
    }
}
int launch_implicit_u1(const double* e_p, const int32_t* idx, int64_t m, int P, const double* V, const double* beta, int r, int64_t n, int64_t ldy, double* Y, cudaStream_t st) {
    if (P <= 0) return 0;
    implicit_u1_kernel<<<(unsigned)P, 256, 0, st>>>(e_p, idx, m, V, beta, r, n, ldy, Y);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("implicit_u1 launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

// ---- permutation prologue, one CTA per permutation column
//   y* = y_pred + U1 E[:, p]                                       (fit_model.nf:98)
//   rss0[p]  = || y* - Qn Qn' y* ||^2   (null refit, fit_model.nf:100-101)
//   Rf[:, p] = y* - Qc Qc' y*, rss_f[p] = || Rf ||^2   (residual w.r.t. the real model's constant columns)
template <int NT>
__device__ double block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < NT / 32; w++) t += sh[w];
    return t;
}

__global__ void __launch_bounds__(256) perm_prep_kernel(const PermPrepArgs a) {
    __shared__ double sh[8];
    __shared__ double coef[24];
    const int p = blockIdx.x;
    double* col = a.UE + (int64_t)p * a.n_pad;
    const int crp = a.crp, cnp = a.cnp;
    // y*
    for (int64_t i = threadIdx.x; i < a.n_pad; i += 256) col[i] = (i < a.n) ? a.y_pred[i] + col[i] : 0.0;
    __syncthreads();
    // null refit residual sum of squares (explicit residual, two passes)
    for (int j = 0; j < a.cn; j++) {
        double v = 0.0;
        for (int64_t i = threadIdx.x; i < a.n; i += 256) v += a.Qn[i * cnp + j] * col[i];
        v = block_sum<256>(v, sh);
        if (threadIdx.x == 0) coef[j] = v;
    }
    __syncthreads();
    {
        double v = 0.0;
        for (int64_t i = threadIdx.x; i < a.n; i += 256) {
            double r = col[i];
            for (int j = 0; j < a.cn; j++) r -= a.Qn[i * cnp + j] * coef[j];
            v += r * r;
        }
        v = block_sum<256>(v, sh);
        if (threadIdx.x == 0) a.rss0[p] = v;
    }
    __syncthreads();
    for (int j = 0; j < a.cr; j++) {
        double v = 0.0;
        for (int64_t i = threadIdx.x; i < a.n; i += 256) v += a.Qc[i * crp + j] * col[i];
        v = block_sum<256>(v, sh);
        if (threadIdx.x == 0) coef[j] = v;
    }
    __syncthreads();
    {
        double v = 0.0;
        for (int64_t i = threadIdx.x; i < a.n; i += 256) {
            double r = col[i];
            for (int j = 0; j < a.cr; j++) r -= a.Qc[i * crp + j] * coef[j];
            col[i] = r;
            v += r * r;
        }
        v = block_sum<256>(v, sh);
        if (threadIdx.x == 0) a.rss_f[p] = v;
    }
}
int launch_perm_prep(const PermPrepArgs& a, cudaStream_t st) {
    if (a.P <= 0) return 0;
    perm_prep_kernel<<<(unsigned)a.P, 256, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("perm_prep launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

// ---- refit flags of a run -> list[0] = count, list[1 + i] = run-wide index of a flagged SNP (order arbitrary; the host sorts
//      the short list).  Replaces an M-element D2H copy and a host scan per flx_run.
__global__ void compact_flags_kernel(const int32_t* __restrict__ flag, int64_t M, int32_t* __restrict__ list) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = i < M && flag[i] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(list, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (on) list[1 + base + __popc(m & ((1u << lane) - 1))] = (int32_t)i;
}
int launch_compact_flags(const int32_t* flag, int64_t M, int32_t* list, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(list, 0, sizeof(int32_t), st);
    if (e == cudaSuccess && M > 0) {
        compact_flags_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(flag, M, list);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { set_error("compact_flags launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

}  // namespace flx
