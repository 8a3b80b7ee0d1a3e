// panel_gemm.cu — the two FP64 tensor-pipe kernels of the hot path (sm_100a):
//
//   K2  whiten_tile         W = L^-1 . X          replaces forwardsolve(L, X)   (fit_model.nf:128)
//   K4  perm_contract_minp  D = W^T . R, fused    replaces the FIT_MODEL_PERM fan-out
//                           min-p reduction       (lmm.nf:56-71, fit_model.nf:96-109,141-143)
//
// Both are persistent, warp-specialised kernels over PACKED PANELS (common.cuh): one producer warp
// streams 16 KB operand chunks into a 5-stage shared-memory ring with 1-D TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx), eight consumer warps issue DMMA (mma.sync m8n8k4 f64 — the
// only FP64 tensor instruction sm_100a has; tcgen05 has no f64 kind) with 128x16 warp tiles, fragments
// read with conflict-free 8-byte LDS.  L^-1 is lower triangular: row tile I only visits its 8(I+1)
// k-chunks and, inside the diagonal block, only the 8x4 sub-blocks on or below the diagonal.
#include "common.cuh"
#include "kernels.h"

namespace flx {

constexpr int STAGES = 5;
constexpr int NCONS_WARPS = 8;
constexpr int GEMM_THREADS = (NCONS_WARPS + 4) * 32;  // 2 consumer warpgroups + 1 producer warpgroup (1 active lane)
constexpr int STAGE_BYTES = 2 * CHUNK_BYTES;
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
    // v >= 0: IEEE order == unsigned integer order of the bit pattern
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// MODE 0: whiten (triangular A = packed L^-1, B = packed X, out = packed W)
// MODE 1: permutation contraction (A = packed W, B = packed R), S = SNP-dependent columns per SNP
template <int MODE, int S>
__global__ void __launch_bounds__(GEMM_THREADS, 1) panel_gemm_kernel(const GemmArgs g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NCONS_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t n_items = g.n_items;

    if (warp >= NCONS_WARPS) {
        // ===================== producer warpgroup: gives its registers away; one elected lane drives TMA =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == NCONS_WARPS && lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                const double* a_src;
                const double* b_src;
                int nk;
                if (MODE == 0) {
                    const int I = g.T - 1 - (int)(item / g.NJ);
                    const int J = (int)(item % g.NJ);
                    a_src = g.A + tri_chunks_before(I) * CHUNK;
                    b_src = g.B + (int64_t)J * g.KC * CHUNK;
                    nk = tri_chunks_of(I);
                    if (nk > g.KC_live) nk = g.KC_live;
                } else {
                    const int J = (int)(item / g.NP);
                    const int PJ = (int)(item % g.NP);
                    a_src = g.A + (int64_t)J * g.KC * CHUNK;
                    b_src = g.B + (int64_t)PJ * g.KC * CHUNK;
                    nk = g.KC;
                }
                for (int kc = 0; kc < nk; kc++) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    double* dstA = stage_base + (size_t)stage * (2 * CHUNK);
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tma_bulk_g2s(dstA, a_src + (int64_t)kc * CHUNK, CHUNK_BYTES, &full_bar[stage]);
                    tma_bulk_g2s(dstA + CHUNK, b_src + (int64_t)kc * CHUNK, CHUNK_BYTES, &full_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    // ===================== consumer warps: DMMA, warp tile = 128 (all 16 j-blocks of A) x 16 =====================
    // The register file is split per SM sub-partition (16384 regs each): 2 consumer warps x 232 + 1 producer warp x 40.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int grp = lane >> 2;   // fragment row (A) / column (B) index
    const int qid = lane & 3;
    uint32_t stage = 0, phase = 0;

    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        int I = 0, J = 0, PJ = 0, nk;
        if (MODE == 0) {
            I = g.T - 1 - (int)(item / g.NJ);
            J = (int)(item % g.NJ);
            nk = tri_chunks_of(I);
            if (nk > g.KC_live) nk = g.KC_live;
        } else {
            J = (int)(item / g.NP);
            PJ = (int)(item % g.NP);
            nk = g.KC;
        }
        double acc[16][2][2];
#pragma unroll
        for (int i = 0; i < 16; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

        const int diag_start = (MODE == 0) ? 8 * I : (1 << 30);
        // row blocks (8 rows each) of this tile that hold real samples: only the last row tile can be partial
        const int i_hi = (MODE == 0 && I == g.T - 1) ? g.last_row_blocks : 16;

        // Software-pipelined main loop: the fragments of k-step t+1 are fetched from shared memory while the DMMAs of
        // k-step t issue (each A register is reloaded right after its last use), so the LDS latency and the mbarrier
        // wait of the next stage hide behind tensor-pipe work.
        mbar_wait(&full_bar[stage], phase);
        const double* sA = stage_base + (size_t)stage * (2 * CHUNK) + lane;
        double a[16], b[2];
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = sA[i * 32];
#pragma unroll
        for (int j = 0; j < 2; j++) b[j] = sA[CHUNK + (warp * 2 + j) * 32];

        for (int kc = 0; kc < nk; kc++) {
            const int d = kc - diag_start;
            const bool plain = (d < 0) && (i_hi == 16);
            uint32_t nstage = stage + 1, nphase = phase;
            if (nstage == STAGES) { nstage = 0; nphase ^= 1; }
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const double* nA;
                if (ks < 3) {
                    nA = sA + (ks + 1) * 16 * 32;
                } else if (kc + 1 < nk) {
                    mbar_wait(&full_bar[nstage], nphase);
                    nA = stage_base + (size_t)nstage * (2 * CHUNK) + lane;
                } else {
                    nA = sA;  // nothing follows in this item: harmless re-read of the current stage
                }
                double bn[2];
#pragma unroll
                for (int j = 0; j < 2; j++) bn[j] = nA[CHUNK + (warp * 2 + j) * 32];
                if (plain) {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        dmma884(acc[i][0][0], acc[i][0][1], a[i], b[0]);
                        dmma884(acc[i][1][0], acc[i][1][1], a[i], b[1]);
                        a[i] = nA[i * 32];
                    }
                } else {
                    // (a) diagonal block of L^-1: k = 128 I + 16 d + 4 ks + (0..3); row block i (rows 8i..8i+7) is non-zero
                    //     only when 8 i + 7 >= 16 d + 4 ks;  (b) partial last row tile: row blocks >= i_hi are padding.
                    const int i_lo = (d >= 0) ? ((16 * d + 4 * ks) >> 3) : 0;
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        if (i >= i_lo && i < i_hi) {
                            dmma884(acc[i][0][0], acc[i][0][1], a[i], b[0]);
                            dmma884(acc[i][1][0], acc[i][1][1], a[i], b[1]);
                        }
                        a[i] = nA[i * 32];
                    }
                }
                b[0] = bn[0];
                b[1] = bn[1];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            stage = nstage;
            phase = nphase;
            sA = stage_base + (size_t)stage * (2 * CHUNK) + lane;
        }

        if (MODE == 0) {
            // ---- epilogue K2: W tile J, rows 128 I .. 128 I + 127, stored in the packed panel layout (k = sample row) ----
            double* wout = g.Out + ((int64_t)J * g.KC + 8 * (int64_t)I) * CHUNK;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int row = 8 * i + grp;  // row within the 128-row tile
                if (8 * I + (row >> 4) < g.KC) {
                    double* wc = wout + (int64_t)(row >> 4) * CHUNK;
#pragma unroll
                    for (int j = 0; j < 2; j++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int col = 16 * warp + 8 * j + 2 * qid + e;
                            wc[chunk_off(row & 15, col)] = acc[i][j][e];
                        }
                }
            }
        } else {
            // ---- epilogue K4: per (SNP, permutation) quadratic form b' Sm b / RSS, running max per (permutation, df) ----
            constexpr int GROUPS = 16 / S;
            constexpr int NSYM = S * (S + 1) / 2;
            double mx[2][2][S];
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int e = 0; e < 2; e++)
#pragma unroll
                    for (int t = 0; t < S; t++) mx[j][e][t] = 0.0;
            double inv_rss[2][2];
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int e = 0; e < 2; e++) inv_rss[j][e] = g.perm_inv_rss[PJ * TJ + 16 * warp + 8 * j + 2 * qid + e];
#pragma unroll
            for (int gi = 0; gi < GROUPS; gi++) {
                const int64_t snp = (int64_t)J * (8 * GROUPS) + 8 * gi + grp;  // SNP index inside the batch
                const int df = g.snp_df[snp];
                if (df > 0) {
                    double sm[NSYM];
#pragma unroll
                    for (int q = 0; q < NSYM; q++) sm[q] = g.snp_sm[snp * NSYM + q];
#pragma unroll
                    for (int j = 0; j < 2; j++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            double quad = 0.0;
                            int q = 0;
#pragma unroll
                            for (int t = 0; t < S; t++) {
                                const double bt = acc[S * gi + t][j][e];
#pragma unroll
                                for (int u = t; u < S; u++) {
                                    const double bu = acc[S * gi + u][j][e];
                                    quad += ((u == t) ? 1.0 : 2.0) * sm[q] * bt * bu;
                                    q++;
                                }
                            }
                            const double ratio = fmax(quad * inv_rss[j][e], 0.0);
#pragma unroll
                            for (int t = 0; t < S; t++)
                                if (df == t + 1) mx[j][e][t] = fmax(mx[j][e][t], ratio);
                        }
                }
            }
            // reduce over the 8 SNP lanes (grp), then one atomic per (permutation, df)
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int e = 0; e < 2; e++)
#pragma unroll
                    for (int t = 0; t < S; t++) {
                        double v = mx[j][e][t];
                        v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 4));
                        v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 8));
                        v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 16));
                        if (grp == 0 && v > 0.0) {
                            const int pcol = PJ * TJ + 16 * warp + 8 * j + 2 * qid + e;
                            atomic_max_nonneg(&g.perm_maxratio[(int64_t)t * g.P_pad + pcol], v);
                        }
                    }
        }
    }
}

template <int MODE, int S>
static int launch_one(const GemmArgs& g, int num_sms, cudaStream_t st) {
    static DeviceOnce configured;
    if (configured.need()) {
        cudaError_t e = cudaFuncSetAttribute(panel_gemm_kernel<MODE, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(panel_gemm): %s", cudaGetErrorString(e)); return -3; }
        configured.set();
    }
    int64_t grid = g.n_items < num_sms ? g.n_items : num_sms;
    if (grid <= 0) return 0;
    panel_gemm_kernel<MODE, S><<<(unsigned)grid, GEMM_THREADS, GEMM_SMEM, st>>>(g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("panel_gemm launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

int launch_whiten(const GemmArgs& g, int num_sms, cudaStream_t st) { return launch_one<0, 1>(g, num_sms, st); }

int launch_perm_contract(const GemmArgs& g, int s, int num_sms, cudaStream_t st) {
    switch (s) {
        case 1: return launch_one<1, 1>(g, num_sms, st);
        case 2: return launch_one<1, 2>(g, num_sms, st);
        case 3: return launch_one<1, 3>(g, num_sms, st);
        case 4: return launch_one<1, 4>(g, num_sms, st);
        case 5: return launch_one<1, 5>(g, num_sms, st);
        case 6: return launch_one<1, 6>(g, num_sms, st);
        case 7: return launch_one<1, 7>(g, num_sms, st);
        case 8: return launch_one<1, 8>(g, num_sms, st);
    }
    set_error("perm_contract: unsupported s=%d", s);
    return -1;
}

}  // namespace flx
