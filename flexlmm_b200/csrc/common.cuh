// common.cuh — shared layout definitions for the sm_100a kernels of the per-SNP LRT engine.
//
// PACKED PANEL LAYOUT (the one operand format every tensor-pipe kernel here consumes)
// A panel stores a matrix P[k][j] (k = reduction index, j = the other index) in 16 KB chunks of
// 16 k x 128 j, each chunk laid out so that the DMMA m8n8k4 fragment of lane l for k-step ks and
// 8-wide j-block jb is ONE conflict-free 8-byte shared-memory load:
//     offset_in_chunk(k, j) = ((k%16)/4 * 16 + (j%128)/8) * 32 + (j%8)*4 + (k%4)
// i.e. [kstep 4][jblock 16][lane 32] with lane = (j%8)*4 + (k%4)  (A frag: row=lane>>2, col=lane&3;
// B frag: row=lane&3, col=lane>>2 — the same element mapping serves both operands).
// Chunks are contiguous, so the producer warp moves one chunk with one cp.async.bulk (TMA, 1-D).
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace flx {

constexpr int KCH = 16;            // k per chunk
constexpr int TJ = 128;            // j per chunk (tile width)
constexpr int CHUNK = KCH * TJ;    // 2048 doubles = 16 KB
constexpr int CHUNK_BYTES = CHUNK * 8;

__host__ __device__ __forceinline__ int chunk_off(int k16, int j128) {
    return (((k16 >> 2) * 16 + (j128 >> 3)) << 5) + ((j128 & 7) << 2) + (k16 & 3);
}

// Number of k-chunks of the lower-triangular L^-1 panel of row tile I: columns 0 .. 128*(I+1)-1
__host__ __device__ __forceinline__ int64_t tri_chunks_before(int64_t I) { return 4 * I * (I + 1); }
__host__ __device__ __forceinline__ int tri_chunks_of(int I) { return 8 * (I + 1); }

// Column layout of a design tile: s SNP-dependent terms, SNPs in groups of 8, term-major inside a group:
//   j-block (8 columns) b  ->  group b / s, term b % s ;  SNP within tile = 8*(b/s) + (j%8)
// so that in the permutation GEMM all s rows of one SNP land in the same thread's C fragments.
struct TileShape {
    int s;          // SNP-dependent columns per SNP
    int groups;     // 8-SNP groups per tile = 16 / s
    int snps;       // SNPs per tile = 8 * groups
    int blocks;     // used j-blocks = s * groups (<= 16)
};
__host__ __device__ __forceinline__ TileShape make_tile_shape(int s) {
    TileShape t; t.s = s; t.groups = 16 / s; t.snps = 8 * t.groups; t.blocks = s * t.groups; return t;
}

}  // namespace flx

#define FLX_CUDA_TRY(expr)                                                                        \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            flx::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? FLX_ERR_OOM : FLX_ERR_CUDA;                \
        }                                                                                         \
    } while (0)

namespace flx {
void set_error(const char* fmt, ...);

// cudaFuncSetAttribute applies to the CURRENT device only, and one process may hold handles on several GPUs (one host
// thread each): one flag per device ordinal instead of one per process.
struct DeviceOnce {
    std::atomic<unsigned long long> done{0};
    static int dev() { int d = 0; cudaGetDevice(&d); return d & 63; }
    bool need() const { return !((done.load(std::memory_order_acquire) >> dev()) & 1ull); }
    void set() { done.fetch_or(1ull << dev(), std::memory_order_release); }
};
}
