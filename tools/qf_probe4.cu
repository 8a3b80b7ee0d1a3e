// qf_probe.cu — scratch probe for the int8 tcgen05 path: correctness of descriptors / TMEM layout and raw throughput.
// D_q[128 x 256] (s32, TMEM) = A[128 x K] (s8, K-major, no swizzle) * B_q[256 x K]^T (s8, K-major), q = 0,1
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s at line %d\n",cudaGetErrorString(e),__LINE__); exit(1);}}while(0)

constexpr int KB = 64;                 // k bytes per stage
constexpr int A_BYTES = 128 * KB;      // 8 KB
constexpr int B_BYTES = 128 * KB;      // 8 KB per slice: THIS CTA's half (128 of the 256 B rows)
constexpr int QC = 2;
constexpr int STAGE_BYTES = A_BYTES + QC * B_BYTES;  // 40 KB
constexpr int STAGES = 6;
constexpr int SMEM = STAGES * STAGE_BYTES + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (sm_100 "version 1")
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version for Blackwell
    return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void remote_arrive(uint64_t* bar, uint32_t cta) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}

// idesc: c_format S32 (2) @4, a_format INT8 (1) @7, b_format INT8 (1) @10, K-major both, N>>3 @17, M>>4 @24
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);   // M = 256 over the CTA pair

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(224, 1) qf_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int nkb, int reps, int32_t* __restrict__ Dout) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* peer_full = empty_bar + STAGES;     // leader only: the peer CTA's stage has landed
    uint64_t* acc_bar = peer_full + STAGES;       // accumulator ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
    const uint32_t crank = cluster_ctarank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&peer_full[s], 1); }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;

    const int total = nkb * reps;
    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < total; it++) {
                const int kb = it % nkb;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                unsigned char* dst = smem + (size_t)stage * STAGE_BYTES;
                // both CTAs' loads complete on the LEADER's barrier (peer bit cleared): cta_group::2 tensor TMA
                const uint32_t lbar = smem_u32(&full_bar[stage]) & 0xFEFFFFFFu;
                if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                const int rowA = (int)(((size_t)blockIdx.x * nkb + kb) * 64);
                asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(&tmA), "r"(lbar), "r"(0), "r"(rowA) : "memory");
                const int rowB = (int)(((size_t)kb * 2 + crank) * QC * 64);
                asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst + A_BYTES)), "l"(&tmB), "r"(lbar), "r"(0), "r"(rowB) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst + A_BYTES + B_BYTES)), "l"(&tmB), "r"(lbar), "r"(0), "r"(rowB + 64) : "memory");
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) {
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < total; it++) {
                mbar_wait(&full_bar[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
#pragma unroll
                for (int q = 0; q < QC; q++) {
                    const uint32_t sb = sa + A_BYTES + q * B_BYTES;
#pragma unroll
                    for (int j = 0; j < KB / 32; j++) {
                        const uint64_t ad = make_desc(sa + j * 2 * 2048, 2048, 128);
                        const uint64_t bd = make_desc(sb + j * 2 * 2048, 2048, 128);     // this CTA's 128-row half of the plane
                        umma_i8(tmem_base + q * 256, ad, bd, IDESC, (it > 0 || j > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[stage]);      // frees the smem stage when the MMAs above have read it
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(acc_bar);
        }
    } else if (warp == 6) {
        // peer CTA: tell the leader when each of MY stages has landed
        if (false) {
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < total; it++) {
                mbar_wait(&full_bar[stage], phase);
                remote_arrive(&peer_full[stage], 0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // epilogue warps 2..5: TMEM lane quadrant = warp % 4
        mbar_wait(acc_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        for (int c0 = 0; c0 < 512; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                           "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            if (Dout && blockIdx.x < 2)
                for (int j = 0; j < 32; j++) Dout[((size_t)blockIdx.x * 128 + row) * 512 + c0 + j] = (int32_t)v[j];
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

int main(int argc, char** argv) {
    const int nkb = 8;  // K = 512
    const int K = nkb * KB;
    std::vector<int8_t> A(2 * 128 * K), B((size_t)QC * 256 * K);   // two distinct A tiles for the CTA pair
    srand(1);
    for (auto& v : A) v = (int8_t)(rand() % 3);
    for (auto& v : B) v = (int8_t)((rand() % 256) - 128);
    // pack: A [kb][c 4][r8 16][8][16], B [kb][q][c 4][n8 32][8][16]
    int nblk = 148;
    std::vector<int8_t> Ap((size_t)nblk * nkb * A_BYTES), Bp((size_t)nkb * 2 * QC * B_BYTES);
    for (int blk = 0; blk < nblk; blk++)
        for (int r = 0; r < 128; r++)
            for (int k = 0; k < K; k++) {
                const int kb = k / KB, b = k % KB, c = b / 16;
                Ap[((size_t)blk * nkb + kb) * A_BYTES + ((c * 16 + r / 8) * 8 + r % 8) * 16 + b % 16] = A[((blk & 1) * 128 + r) * K + k];
            }
    for (int q = 0; q < QC; q++)
        for (int n = 0; n < 256; n++)
            for (int k = 0; k < K; k++) {
                const int kb = k / KB, b = k % KB, c = b / 16;
                const int half = n / 128, nl = n % 128;
                Bp[(((size_t)kb * 2 + half) * QC + q) * B_BYTES + ((c * 16 + nl / 8) * 8 + nl % 8) * 16 + b % 16] = B[((size_t)q * 256 + n) * K + k];
            }
    int8_t *dA, *dB; int32_t* dD;
    CK(cudaMalloc(&dA, Ap.size())); CK(cudaMalloc(&dB, Bp.size())); CK(cudaMalloc(&dD, 256 * 512 * 4));
    CK(cudaMemcpy(dA, Ap.data(), Ap.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bp.data(), Bp.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, 256 * 512 * 4));
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeFn encode = nullptr; cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    auto make_map = [&](void* base, size_t bytes) {
        CUtensorMap m; cuuint64_t dims[2] = {128, bytes / 128}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {128, 64}; cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
        return m;
    };
    CUtensorMap tmA = make_map(dA, Ap.size()), tmB = make_map(dB, Bp.size());
    CK(cudaFuncSetAttribute(qf_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    qf_probe_kernel<<<2, 224, SMEM>>>(tmA, tmB, nkb, 1, dD);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> D(256 * 512);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int q = 0; q < QC; q++)
        for (int r = 0; r < 256; r++)
            for (int n = 0; n < 256; n++) {
                int32_t ref = 0;
                for (int k = 0; k < K; k++) ref += (int32_t)A[r * K + k] * (int32_t)B[((size_t)q * 256 + n) * K + k];
                const int32_t got = D[r * 512 + q * 256 + n];
                if (got != ref) { if (bad < 5) printf("mismatch q=%d r=%d n=%d got %d ref %d\n", q, r, n, got, ref); bad++; }
            }
    printf("correctness: %ld mismatches of %d\n", bad, QC * 256 * 256);
    // throughput: 148 CTAs x reps
    for (int reps : {50, 400}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        qf_probe_kernel<<<nblk, 224, SMEM>>>(tmA, tmB, nkb, reps, nullptr);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        qf_probe_kernel<<<nblk, 224, SMEM>>>(tmA, tmB, nkb, reps, nullptr);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double ops = 2.0 * 128 * 256 * QC * (double)K * reps * nblk;
        printf("reps %d: %.3f ms  %.1f TOPS int8 (%.1f B/clk/SM operand fill)\n", reps, ms, ops / ms * 1e-9, (double)STAGE_BYTES * nkb * reps / (ms * 1e-3 * 1.965e9));
    }
    return 0;
}
