// qf.cu — the int8 tensor-core route (tcgen05.mma kind::i8, TMEM accumulators): K2q (quadratic forms), K4q (x against
// whitened matrices), K1q (pgen 2-bit records -> 2-bit code tiles) and the once-per-task digit-plane packers.
//
// For a design whose SNP terms are integer genotype codes g(x) (x, I(x==1): values in {-2..2}) times a per-sample scale c
// (1, or the covariate of an x:cov interaction) the only O(n^2)-per-SNP quantities of the fit are the forms
//   g' C_a V^-1 C_b g'  =  g' T g' + g'' T' g      (same scale: T = strict lower triangle of C V^-1 C plus half its diagonal)
//                       =  g' F g'                 (two scales: the full matrix F = C_a V^-1 C_b)
// (everything else — Qc' L^-1 x, r' L^-1 x, R_f' L^-1 x — is g against n x (c+1+P) matrices that are whitened once: K4q).
// g is an exact small integer vector, so T / F are split ONCE into Q = 6 balanced base-256 digit planes ("Ozaki slices",
// 48 bits below a power-of-two scale per 16 rows): every int8 x int8 product and every s32 accumulation is exact, and the
// result differs from the FP64 value only by the 2^-48 truncation of the matrix — measured 4e-15 relative in x' V^-1 x at
// n = 2000, better than the rounding a length-n FP64 dot product accumulates.  Throughput of the i8 pipe (4.5 POP/s
// nominal, 4.17 POP/s measured by tools/qf_probe.cu) over 6 slices is ~19x the FP64 DMMA pipe.
//
// The genotype operand travels as 2-BIT CODES (a quarter of the int8 bytes) from K1q through HBM and L2 into shared memory
// and is expanded to the int8 UMMA layout inside the SM: with two MMA issuers the L2 -> SM stream was the bound (32 KB per 512
// tensor clocks per SM = 16 TB/s over the chip; profiles/r02_k2q_issuer_probes.txt), the codes make it 20 KB.  A tile set's
// code c stands for the int8 value vt[c] (vt[0] = 0: missing / masked samples), at most three non-zero values per set.
//
// K2q kernel: persistent, warp-specialised, one CTA of 16 warps per SM (warp ids in scheduling-priority order, lowest first).
//   warps 0-3  : expanders — 2-bit codes -> int8 in the UMMA layout, warp w owns ring slot w of a 4-slot x 16 KB ring and expands every
//                fourth stage (conflict-free LDS.32 / STS.128: the code word of (tile, k16 chunk, SNP) expands to the 16 bytes of exactly
//                that core-matrix row), fence.proxy.async.
//   warps 4-11 : epilogue (4 per accumulator) — tcgen05.ld an accumulator 16 rows at a time, multiply by the left vector(s) at those rows,
//                reduce the 16 rows exactly in s32 (they share one power-of-two scale), scale and sum over the 256 rows
//                IN THE THREAD (thread = TMEM lane = SNP): no cross-lane reduction.  One partial per (row tile, plane, SNP).
//   warp 12    : TMA producer — 1-D bulk copies of pre-packed blocks into an 8-stage x 20 KB ring: one digit-plane block (16 KB,
//                UMMA canonical K-major, no swizzle: 8 x 16 B core matrices) + the code block of a PAIR of SNP tiles (4 KB).
//   warps 13-14: MMA issuers, one per SNP tile of the pair (tcgen05.mma.cta_group::1.kind::i8, M = 128 SNPs, N = 256 rows of T,
//                K = 32 per instruction, one 128 x 256 s32 accumulator each; a tcgen05.mma costs its issuing thread ~85 clk whatever
//                its size, one thread cannot feed the pipe); tcgen05.commit frees ring stages and slots.  Warp 13 owns the 512
//                TMEM columns.  (warp 15 idle: setmaxnreg works on whole warpgroups.)
// All single-thread roles run their loops warp-wide and elect only the issuing instruction (uniform registers, no R2UR waterfall).
// Work item = (row tile I of 256 rows, pair of SNP tiles, digit plane q); k runs over the lower-triangular range (all k
// for a full matrix).
#include "common.cuh"
#include "kernels.h"

namespace flx {

constexpr int QF_KB = 64;                      // samples (k) per stage
constexpr int QF_A_BYTES = 128 * QF_KB;        // 8 KB: one SNP tile block expanded to int8
constexpr int QF_A2_WORDS = 128 * QF_KB / 16;  // 512 words = 2 KB: the same block as 2-bit codes, [k16 chunk 4][SNP 128] words of 16 samples
constexpr int QF_P2_BYTES = 2 * QF_A2_WORDS * 4;   // 4 KB: the code blocks of a tile PAIR, [tile 2][chunk 4][SNP 128]
constexpr int QF_B_BYTES = 256 * QF_KB;        // 16 KB: one digit plane block
constexpr int QF_STAGE_BYTES = QF_B_BYTES + QF_P2_BYTES;      // 20 KB, filled by TMA
#ifndef QF_RING_STAGES
#define QF_RING_STAGES 8
#define QF_RING_SLOTS 4
#endif
constexpr int QF_STAGES = QF_RING_STAGES;
constexpr int QF_SLOTS = QF_RING_SLOTS;        // ring of expanded pairs, one expander warp per slot
constexpr int QF_SLOT_BYTES = 2 * QF_A_BYTES;  // 16 KB
constexpr int QF_SLOT0 = QF_STAGES * QF_STAGE_BYTES;
constexpr int QF_BAR0 = QF_SLOT0 + QF_SLOTS * QF_SLOT_BYTES;
constexpr int QF_SMEM = QF_BAR0 + 256;         // <= 232,448 B a CTA may use
static_assert(QF_SMEM <= 232448, "ring does not fit");
// warps, lowest scheduling priority first (the arbiter prefers the highest warp id): expanders, then the eight epilogue warps, then the
// producer warpgroup (TMA producer, the two MMA issuers, one idle warp) whose few instructions per stage must never wait for a slot
constexpr int QF_EXP_WG = (QF_SLOTS + 3) / 4;
constexpr int QF_EPI_WARP0 = 4 * QF_EXP_WG;
constexpr int QF_PROD_WARP0 = QF_EPI_WARP0 + 8;
constexpr int QF_THREADS = (QF_PROD_WARP0 + 4) * 32;
// register pool = what the launch allocates (QF_THREADS x the per-thread count ptxas settles on under the launch bound, a multiple of 8):
// expanders and the producer warpgroup shrink to 40, the epilogue warps take the rest.  One register too many and setmaxnreg.inc never returns.
constexpr int QF_LAUNCH_REGS = (65536 / QF_THREADS) / 8 * 8;
constexpr int QF_EPI_REGS_RAW = ((QF_THREADS * QF_LAUNCH_REGS - (QF_EXP_WG + 1) * 128 * 40) / 256) / 8 * 8;
constexpr int QF_EPI_REGS = QF_EPI_REGS_RAW > 208 ? 208 : QF_EPI_REGS_RAW;

__device__ __forceinline__ uint32_t qsmem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void qbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(qsmem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void qbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(qsmem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool qbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(qsmem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void qbar_wait(uint64_t* bar, uint32_t parity) { while (!qbar_try_wait(bar, parity)) {} }
// instruction descriptor: D = S32 (2 @bit4), A = INT8 (1 @7), B = INT8 (1 @10), both K-major, N>>3 @17, M>>4 @24
constexpr uint32_t QF_IDESC_M128 = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);      // | (N >> 3) << 17

#define XG_TMEM_LD16(v, taddr)                                                                                                  \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"       \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),   \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                      \
                 : "r"(taddr))

// sign-extended byte b of w in one PRMT (selector nibble 8|b replicates the byte's sign bit)
__device__ __forceinline__ int sext_byte(uint32_t w, int b) {
    uint32_t d;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(d) : "r"(w), "r"(0x8880u + 0x1111u * (uint32_t)b));
    return (int)d;
}

// ---- 2-bit code words.  A word holds 16 consecutive samples of one SNP; sample j sits at bit 2 (j / 4) + 8 (j % 4), so that
// (w >> 2 m) & 0x03030303 is samples 4 m .. 4 m + 3 as four bytes: one shift and one mask per int8 word when the codes are the
// values (vt = 0, 1, 2, 3: x, I(x == g), 2 - x ...).  Other value sets (negative entries) go through three byte-wise selects.
constexpr uint32_t QF_VT_IDENTITY = 0x03020100u;
template <bool IDENT>
__device__ __forceinline__ uint32_t code_expand(uint32_t w, int m, uint32_t vt) {
    if (IDENT) return (w >> (2 * m)) & 0x03030303u;
    const uint32_t lo = (w >> (2 * m)) & 0x01010101u, hi = (w >> (2 * m + 1)) & 0x01010101u, both = lo & hi;
    // bytes are 0 / 1 and at most one of the three selects is set per byte: the products cannot carry
    return (lo ^ both) * ((vt >> 8) & 0xffu) + (hi ^ both) * ((vt >> 16) & 0xffu) + both * (vt >> 24);
}

// item -> (row tile I, SNP tile pair JP, digit plane q).  Tile pairs are taken in groups of GP: a group goes through ALL row tiles
// (long ones first) before the next group starts, so the group's genotype tiles (sized for L2 by the host) are read from HBM once
// per launch instead of once per row tile; planes innermost, so the Q CTAs that work on one tile pair at the same time stream the
// same code blocks.  GP >= the pair count reproduces the plain (I, JP, q) order.
__device__ __forceinline__ void qf_item(const QfArgs& g, int64_t item, int NJ2, int& I, int& JP, int& q) {
    const int64_t per_group = (int64_t)g.T256 * g.GP * g.Q;
    const int G = (int)(item / per_group);
    int64_t rem = item - (int64_t)G * per_group;
    const int pairs = min(g.GP, NJ2 - G * g.GP);
    const int per_I = pairs * g.Q;
    I = g.T256 - 1 - (int)(rem / per_I);
    rem %= per_I;
    JP = G * g.GP + (int)(rem / g.Q);
    q = (int)(rem % g.Q);
}

// ---- lean forms for the single-thread roles (TMA producer, MMA issuers): at 512 tensor clocks per stage a dependent chain of
// ~100 instructions IS the critical path (measured: the loops took 450-650 clk per stage), so barrier and descriptor words are
// 32-bit shared-window values advanced by adds, nothing is recomputed per stage
__device__ __forceinline__ bool qbar_try_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void qbar_wait_a(uint32_t bar, uint32_t parity) { while (!qbar_try_wait_a(bar, parity)) {} }
__device__ __forceinline__ void qbar_arrive_a(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void qbar_expect_tx_a(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void qtma_bulk_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
// K-major, SWIZZLE_NONE shared-memory matrix descriptors (sm_100 descriptor version 1; canonical layout ((8, n), 2) : ((16 B, SBO), LBO),
// 8-row x 16-byte core matrices) as (lo, hi) words: lo = (addr >> 4) | (LBO >> 4) << 16, hi = (SBO >> 4) | version 1 << 14 (bit 46)
__device__ __forceinline__ void umma_i8_w(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n\t}"
                 ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
}
// one lane of a converged warp (the same lane every time): the single-thread roles run their loops WARP-WIDE and elect only the
// issuing instruction, so that every loop variable stays warp-uniform and ptxas keeps barrier addresses and descriptors in uniform
// registers; under `if (lane == 0)` it wraps each UTCIMMA / UBLKCP in an ELECT + R2UR.BROADCAST waterfall loop (~25 instructions per MMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- the ring shared by K2q and K4q.  Shared memory: QF_STAGES x [plane block 16 KB | code blocks of the tile pair 4 KB], QF_SLOTS x
// [expanded tile 0 | expanded tile 1], then the barriers:
//   full[s]     TMA landed stage s (tx bytes)                  -> expanders, issuers
//   empty[s]    both issuers' MMAs have read stage s (commits) -> producer          (the expanders finished with it long before)
//   xfull[k]    expander warp k wrote slot k                   -> issuers
//   xempty[k]   both issuers' MMAs have read slot k (commits)  -> expanders
//   acc_full[t] / acc_empty[t]   accumulator of tile t: issuer t <-> the four epilogue warps of tile t
struct QfBars {
    uint32_t full, empty, xfull, xempty, acc_full, acc_empty;      // shared-window addresses of the first barrier of each kind
};
__device__ __forceinline__ QfBars qf_bars(unsigned char* smem) {
    const uint32_t b = qsmem_u32(smem + QF_BAR0);
    return QfBars{b, b + 8 * QF_STAGES, b + 16 * QF_STAGES, b + 16 * QF_STAGES + 8 * QF_SLOTS, b + 16 * QF_STAGES + 16 * QF_SLOTS, b + 16 * QF_STAGES + 16 * QF_SLOTS + 16};
}
__device__ __forceinline__ void qf_bars_init(unsigned char* smem) {
    uint64_t* b = reinterpret_cast<uint64_t*>(smem + QF_BAR0);
    for (int s = 0; s < QF_STAGES; s++) { qbar_init(b + s, 1); qbar_init(b + QF_STAGES + s, 2); }
    for (int k = 0; k < QF_SLOTS; k++) { qbar_init(b + 2 * QF_STAGES + k, 1); qbar_init(b + 2 * QF_STAGES + QF_SLOTS + k, 2); }
    for (int t = 0; t < 2; t++) { qbar_init(b + 2 * QF_STAGES + 2 * QF_SLOTS + t, 1); qbar_init(b + 2 * QF_STAGES + 2 * QF_SLOTS + 2 + t, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ uint32_t* qf_tmem_slot(unsigned char* smem) { return reinterpret_cast<uint32_t*>(smem + QF_BAR0 + 8 * (2 * QF_STAGES + 2 * QF_SLOTS + 4)); }

// expanders (warps 4-7): warp w expands the stages i = w, w + 4, ... of this CTA's ring into slot w — four stages in flight, a
// stage's barrier round trips, fence and 32 words per lane are one warp's business (4 x 512 tensor clocks to do it in)
template <bool IDENT>
__device__ __forceinline__ void qf_expand_loop(unsigned char* smem, const QfBars& bars, int w, int lane, int64_t nstages, uint32_t vt) {
    uint4* dst = reinterpret_cast<uint4*>(smem + QF_SLOT0 + w * QF_SLOT_BYTES) + lane;
    const uint32_t xfull = bars.xfull + 8 * w, xempty = bars.xempty + 8 * w;
    uint32_t sphase = 0;
    for (int64_t i = w; i < nstages; i += QF_SLOTS) {
        const uint32_t stage = (uint32_t)(i % QF_STAGES), phase = (uint32_t)((i / QF_STAGES) & 1);
        qbar_wait_a(bars.full + 8 * stage, phase);
        qbar_wait_a(xempty, sphase ^ 1);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(smem + stage * QF_STAGE_BYTES + QF_B_BYTES) + lane;
#pragma unroll
        for (int k0 = 0; k0 < 32; k0 += 8) {
            uint32_t wd[8];
#pragma unroll
            for (int k = 0; k < 8; k++) wd[k] = src[(k0 + k) * 32];      // word index = (tile, chunk, SNP) = byte offset / 16 of the expanded row
#pragma unroll
            for (int k = 0; k < 8; k++)
                dst[(k0 + k) * 32] = make_uint4(code_expand<IDENT>(wd[k], 0, vt), code_expand<IDENT>(wd[k], 1, vt), code_expand<IDENT>(wd[k], 2, vt), code_expand<IDENT>(wd[k], 3, vt));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) qbar_arrive_a(xfull);
        sphase ^= 1;
    }
}
__device__ __forceinline__ void qf_expand_stages(unsigned char* smem, const QfBars& bars, int w, int lane, int64_t nstages, uint32_t vt) {
    if (vt == QF_VT_IDENTITY) qf_expand_loop<true>(smem, bars, w, lane, nstages, vt);
    else qf_expand_loop<false>(smem, bars, w, lane, nstages, vt);
}

// prof (PROF builds only) [16 * grid] cycle counters — issuer of tile 0: 0 total, 1 waiting for the epilogue, 2 waiting for operands (TMA +
// expansion), 3 of which in the first ring of an item, 4-5 waits > 200 clk (sum, count), 6 issuing; producer: 8 total, 9 waiting for a free
// stage, 10-11 waits > 200 clk; first epilogue warp: 12 waiting for the accumulator, 13 draining it
template <int NDOT, bool PROF>
__global__ void __launch_bounds__(QF_THREADS, 1) qf_i8_kernel(const QfArgs g) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint32_t* tmem_slot = qf_tmem_slot(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) qf_bars_init(smem);
    if (warp == QF_PROD_WARP0 + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(qsmem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = qsmem_u32(smem);
    const QfBars bars = qf_bars(smem);
    const int NJ2 = (g.NJ + 1) >> 1;

    // the register file is partitioned per SM sub-partition, so the split must be explicit
    if (warp >= QF_PROD_WARP0) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
      if (warp == QF_PROD_WARP0) {
        // ---------------- TMA producer (warp-wide loop, one elected lane issues)
        uint32_t stage = 0, phase = 0;
        long long p_empty = 0, p_long = 0, p_nlong = 0;
        const long long p_begin = PROF ? clock64() : 0;
        const int64_t b_step = (int64_t)g.Q * QF_B_BYTES;
        for (int64_t item = blockIdx.x; item < g.n_items; item += gridDim.x) {
            int I, JP, q;
            qf_item(g, item, NJ2, I, JP, q);
            int nkb = g.full ? g.NKB_live : 4 * (I + 1);      // lower-triangular matrices: k runs to the diagonal block
            if (nkb > g.NKB_live) nkb = g.NKB_live;
            const uint32_t* ap = g.X2 + (int64_t)JP * g.NKB * (2 * QF_A2_WORDS);
            // triangular: 4*I*(I+1)/2 blocks precede tile I; full: NKB blocks per tile
            const int8_t* bp = g.T8 + ((g.full ? (int64_t)I * g.NKB : (int64_t)2 * I * (I + 1)) * g.Q + q) * QF_B_BYTES;
            for (int kb = 0; kb < nkb; kb++) {
                const uint32_t fb = bars.full + 8 * stage, dst = smem0 + stage * QF_STAGE_BYTES;
                if (PROF) {
                    const long long t0 = clock64();
                    qbar_wait_a(bars.empty + 8 * stage, phase ^ 1);
                    const long long dt = clock64() - t0;
                    p_empty += dt; if (dt > 200) { p_long += dt; p_nlong++; }
                } else qbar_wait_a(bars.empty + 8 * stage, phase ^ 1);
                if (elect_one()) {
                    qbar_expect_tx_a(fb, QF_STAGE_BYTES);
                    qtma_bulk_a(dst, bp, QF_B_BYTES, fb);
                    qtma_bulk_a(dst + QF_B_BYTES, ap, QF_P2_BYTES, fb);
                }
                bp += b_step; ap += 2 * QF_A2_WORDS;
                if (++stage == QF_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        if (PROF && lane == 0) { long long* pp = g.prof + 16 * blockIdx.x; pp[8] = clock64() - p_begin; pp[9] = p_empty; pp[10] = p_long; pp[11] = p_nlong; }
      } else if (warp <= QF_PROD_WARP0 + 2) {
        // ---------------- MMA issuers: the first drives SNP tile 0 of the pair (accumulator columns 0-255), the second tile 1 (256-511).
        // A tcgen05.mma costs its issuing thread ~85 clk whatever its size (measured with N = 32: same kernel time), so ONE thread
        // issuing a stage's four MMAs plus the barrier round trip needs > 600 clk for 512 clk of tensor work — the kernel was bound
        // by that thread, not by the pipe.  The two accumulators are independent, so each gets its own issuer; a stage and a slot
        // are released when both have committed.
        const int t = warp - (QF_PROD_WARP0 + 1);
        uint32_t stage = 0, phase = 0, slot = 0, sphase = 0, aphase = 0;
        long long t_acc = 0, t_full = 0, t_head = 0, t0 = 0, t_long = 0, n_long = 0, t_issue = 0;
        const long long t_begin = PROF ? clock64() : 0;
        // A (expanded tile): [k16 chunk 4][row group 16][8 rows][16 B] -> LBO 2048 B, SBO 128 B;  B: [k16 chunk 4][row group 32][8 rows][16 B] -> LBO 4096 B
        const uint32_t d_hi = (128u >> 4) | (1u << 14);
        const uint32_t a_lo0 = (((smem0 + QF_SLOT0 + t * QF_A_BYTES) >> 4) & 0x3FFFu) | ((2048u >> 4) << 16);
        const uint32_t b_lo0 = ((smem0 >> 4) & 0x3FFFu) | ((4096u >> 4) << 16);
        const uint32_t acc_full_a = bars.acc_full + 8 * t, acc_empty_a = bars.acc_empty + 8 * t;
        const uint32_t tacc = tmem_base + t * 256;
        for (int64_t item = blockIdx.x; item < g.n_items; item += gridDim.x) {
            int I, JP, q;
            qf_item(g, item, NJ2, I, JP, q);
            const bool mine = 2 * JP + t < g.NJ;      // an odd tile count leaves the last pair a single tile: the second issuer only keeps step
            int nkb = g.full ? g.NKB_live : 4 * (I + 1);
            if (nkb > g.NKB_live) nkb = g.NKB_live;
            // rows of T the MMAs must cover: the last tile's padding rows are zero planes (N shrinks to the live rows, a multiple
            // of 16), and on the diagonal tile of a triangular matrix the rows above the k range of a K step are zero as well
            int nrows = 256;
            if (g.n - 256 * (int64_t)I < 256) nrows = (int)((g.n - 256 * (int64_t)I + 15) & ~(int64_t)15);
            const int kb_plain = g.full ? nkb : min(nkb, 4 * I);      // k-blocks below the diagonal tile: all rows
            const uint32_t idesc_full = QF_IDESC_M128 | (((uint32_t)nrows >> 3) << 17);
            if (PROF) t0 = clock64();
            qbar_wait_a(acc_empty_a, aphase ^ 1);           // the epilogue has drained this accumulator of the previous item
            if (PROF) t_acc += clock64() - t0;
            asm volatile("tcgen05.fence::after_thread_sync;");
            for (int kb = 0; kb < nkb; kb++) {
                if (PROF) t0 = clock64();
                qbar_wait_a(bars.full + 8 * stage, phase);          // the plane block (TMA)
                qbar_wait_a(bars.xfull + 8 * slot, sphase);         // the expanded tiles
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (PROF) { const long long dt = clock64() - t0; t_full += dt; if (kb < QF_STAGES) t_head += dt; if (dt > 200) { t_long += dt; n_long++; } t0 = clock64(); }
                const uint32_t al = a_lo0 + slot * (QF_SLOT_BYTES >> 4), bl = b_lo0 + stage * (QF_STAGE_BYTES >> 4);
                if (mine) {
                    if (kb < kb_plain) {
                        if (elect_one()) {
                            umma_i8_w(tacc, al, bl, d_hi, idesc_full, kb > 0 ? 1u : 0u);
                            umma_i8_w(tacc, al + (4096 >> 4), bl + (8192 >> 4), d_hi, idesc_full, 1u);
                        }
                    } else {
                        // diagonal tile: K step j of k-block kb only meets rows r >= r0 = 64 (kb - 4 I) + 32 j
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const int r0 = 64 * (kb - 4 * I) + 32 * j;
                            const int N = nrows - r0;
                            if (N > 0) {
                                const uint32_t id = QF_IDESC_M128 | (((uint32_t)N >> 3) << 17);
                                if (elect_one()) umma_i8_w(tacc + r0, al + ((j * 4096) >> 4), bl + ((j * 8192) >> 4) + (uint32_t)(r0 >> 3) * (128 >> 4), d_hi, id, (kb > 0 || j > 0) ? 1u : 0u);
                            }
                        }
                    }
                }
                if (elect_one()) { umma_commit_a(bars.empty + 8 * stage); umma_commit_a(bars.xempty + 8 * slot); }
                if (PROF) t_issue += clock64() - t0;
                if (++stage == QF_STAGES) { stage = 0; phase ^= 1; }
                if (++slot == QF_SLOTS) { slot = 0; sphase ^= 1; }
            }
            if (elect_one()) umma_commit_a(acc_full_a);
            aphase ^= 1;
        }
        if (PROF && t == 0 && lane == 0) { long long* pp = g.prof + 16 * blockIdx.x; pp[0] = clock64() - t_begin; pp[1] = t_acc; pp[2] = t_full; pp[3] = t_head; pp[4] = t_long; pp[5] = n_long; pp[6] = t_issue; }
      }
    } else if (warp < QF_EPI_WARP0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        // ---------------- expanders: the stage count of this CTA's items, then the ring in order
        int64_t nstages = 0;
        for (int64_t item = blockIdx.x; item < g.n_items; item += gridDim.x) {
            int I, JP, q;
            qf_item(g, item, NJ2, I, JP, q);
            int nkb = g.full ? g.NKB_live : 4 * (I + 1);
            if (nkb > g.NKB_live) nkb = g.NKB_live;
            nstages += nkb;
        }
        if (warp < QF_SLOTS) qf_expand_stages(smem, bars, warp, lane, nstages, g.vt_mma);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(QF_EPI_REGS));
        // ---------------- epilogue: thread = TMEM lane = SNP; the first four warps own the first SNP tile of the pair, the other four the second
        const int quad = warp & 3;
        const int t = (warp - QF_EPI_WARP0) >> 2;
        const int snp_l = quad * 32 + lane;
        uint64_t* acc_full = reinterpret_cast<uint64_t*>(smem + QF_BAR0) + 2 * QF_STAGES + 2 * QF_SLOTS + t;
        uint64_t* acc_empty = acc_full + 2;
        uint32_t aphase = 0;
        long long e_wait = 0, e_work = 0;
        for (int64_t item = blockIdx.x; item < g.n_items; item += gridDim.x) {
            int I, JP, q;
            qf_item(g, item, NJ2, I, JP, q);
            const int J = 2 * JP + t;
            const bool live = J < g.NJ;
            // the left vectors at the rows of the tile (for a plain quadratic form the SNP's own genotypes): k = 256 I + 16 c + j lies in
            // block 4 I + c / 4, chunk c % 4 — sixteen code words per left vector, expanded chunk by chunk below
            const int64_t xoff = (((int64_t)JP * g.NKB + 4 * I) * 2 + t) * QF_A2_WORDS + snp_l;
            // the chunks' power-of-two scales: the high words of 16 doubles (the low words are zero), fetched with the left vectors
            // BEFORE the accumulator wait — a load inside the chunk loop would put a global-memory round trip into every chunk.
            // Chunks beyond the live rows of the last tile were not written by the (trimmed) MMAs: scale 0 discards whatever TMEM holds
            const int* rs_hi = reinterpret_cast<const int*>(g.rowscale + 256 * (int64_t)I) + 1;
            const int nch = (g.n - 256 * (int64_t)I < 256) ? (int)((g.n - 256 * (int64_t)I + 15) >> 4) : 16;
            uint32_t rsp[8];            // sign + exponent (the top 12 bits of a power of two) of two chunks per register
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
                const uint32_t lo16 = (c < nch) ? ((uint32_t)__ldg(rs_hi + 32 * c) >> 20) : 0u, hi16 = (c + 1 < nch) ? ((uint32_t)__ldg(rs_hi + 32 * (c + 1)) >> 20) : 0u;
                rsp[c >> 1] = lo16 | (hi16 << 16);
            }
            // one left vector: its sixteen code words are expanded to int8 words here, off the critical path (64 registers);
            // two: the code words stay packed (32 registers) and each chunk expands its own
            bool ident = g.vt_dot[0] == QF_VT_IDENTITY;
            if (NDOT == 2) ident = ident && g.vt_dot[1] == QF_VT_IDENTITY;
            uint32_t xp[NDOT == 1 ? 64 : 32];
#pragma unroll
            for (int d = 0; d < NDOT; d++)
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    const uint32_t wd = live ? __ldg(g.Xd2[d] + xoff + (int64_t)(c >> 2) * (2 * QF_A2_WORDS) + (c & 3) * 128) : 0u;
                    if (NDOT == 1) {
#pragma unroll
                        for (int m = 0; m < 4; m++) xp[4 * c + m] = ident ? code_expand<true>(wd, m, 0) : code_expand<false>(wd, m, g.vt_dot[0]);
                    } else xp[16 * d + c] = wd;
                }
            if (NDOT == 1) {
                // pin the expanded words here: the compiler would otherwise sink their (pure ALU) computation below the wait, into the drain
#pragma unroll
                for (int i = 0; i < 64; i++) asm volatile("" : "+r"(xp[i]));
            }
            const long long e0 = PROF ? clock64() : 0;
            qbar_wait(acc_full, aphase);
            aphase ^= 1;
            const long long e1 = PROF ? clock64() : 0;
            asm volatile("tcgen05.fence::after_thread_sync;");
            double sum[NDOT];
#pragma unroll
            for (int d = 0; d < NDOT; d++) sum[d] = 0.0;
            if (live) {
                const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + t * 256;
                uint32_t acc[2][16];
                XG_TMEM_LD16(acc[0], trow);
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;");
                    if (c + 1 < 16) XG_TMEM_LD16(acc[(c + 1) & 1], trow + 16 * (c + 1));     // next 16 rows in flight while these are reduced
                    // the 16 rows of a chunk share one power-of-two scale, so the chunk is reduced exactly in s32
                    // (|acc| <= 128 n |x|max, |x|max <= 2 on this route: launch_qf refuses n that could overflow) and
                    // converted once, without I2F (16/clk/SM): 2^52 + 2^51 + v has v in its mantissa, one DADD removes the bias.
#pragma unroll
                    for (int d = 0; d < NDOT; d++) {
                        int cx32 = 0;
#pragma unroll
                        for (int m = 0; m < 4; m++) {
                            const uint32_t xw = NDOT == 1 ? xp[4 * c + m] : (ident ? code_expand<true>(xp[16 * d + c], m, 0) : code_expand<false>(xp[16 * d + c], m, g.vt_dot[d]));
#pragma unroll
                            for (int b = 0; b < 4; b++) cx32 += (int)acc[c & 1][4 * m + b] * sext_byte(xw, b);
                        }
                        // a chunk the MMAs did not write may hold anything: its scale is 0 and 0 * (finite integer) = 0
                        const long long cx = cx32;
                        const double dv = __longlong_as_double(0x4330000000000000LL + (cx + (1LL << 51))) - 6755399441055744.0;
                        sum[d] = fma(dv, __hiloint2double((int)(((rsp[c >> 1] >> (16 * (c & 1))) & 0xffffu) << 20), 0), sum[d]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) qbar_arrive(acc_empty);
            if (PROF && warp == QF_EPI_WARP0 && lane == 0) { e_wait += e1 - e0; e_work += clock64() - e1; }
            const int64_t snp = (int64_t)J * 128 + snp_l;
            if (live && snp < g.nsnp) {
#pragma unroll
                for (int d = 0; d < NDOT; d++) g.part[d][((int64_t)I * g.Q + q) * g.nsnp + snp] = sum[d] * g.plane_scale[q];
            }
        }
        if (PROF && warp == QF_EPI_WARP0 && lane == 0) { long long* pp = g.prof + 16 * blockIdx.x; pp[12] = e_wait; pp[13] = e_work; }
    }
    __syncthreads();
    if (warp == QF_PROD_WARP0 + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

template <int NDOT, bool PROF>
static cudaError_t qf_launch_one(const QfArgs& g, unsigned grid, cudaStream_t st) {
    static DeviceOnce configured;
    if (configured.need()) {
        cudaError_t e = cudaFuncSetAttribute(qf_i8_kernel<NDOT, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, QF_SMEM);
        if (e != cudaSuccess) return e;
        configured.set();
    }
    qf_i8_kernel<NDOT, PROF><<<grid, QF_THREADS, QF_SMEM, st>>>(g);
    return cudaGetLastError();
}

int launch_qf(const QfArgs& g, int num_sms, cudaStream_t st) {
    if (g.n_items <= 0) return 0;
    if (g.GP < 1) { set_error("qf: GP must be >= 1"); return -1; }
    // the epilogue reduces 16 rows in s32: 16 * (128 * n_pad * xmax) * xmax with xmax = 2 must stay below 2^31
    if ((int64_t)g.NKB * QF_KB * 16 * 128 * 4 >= ((int64_t)1 << 31)) { set_error("qf: n too large for the s32 chunk reduction"); return -1; }
    const unsigned grid = (unsigned)(g.n_items < num_sms ? g.n_items : num_sms);
    cudaError_t e;
    if (g.ndot == 1) e = g.prof ? qf_launch_one<1, true>(g, grid, st) : qf_launch_one<1, false>(g, grid, st);
    else if (g.ndot == 2) e = g.prof ? qf_launch_one<2, true>(g, grid, st) : qf_launch_one<2, false>(g, grid, st);
    else { set_error("qf: ndot must be 1 or 2"); return -1; }
    if (e != cudaSuccess) { set_error("qf launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K1q: pgen 2-bit hard calls -> 2-bit CODE tiles of one tile set (integer triple x sample mask), the genotype operand of K2q / K4q:
//   X2[tile pair J / 2][kb][tile J % 2][k16 chunk 4][SNP 128]  words of 16 samples, sample j at bit 2 (j / 4) + 8 (j % 4)
// code = gmap[genotype] (a.gmap byte g), 0 for missing calls (flagged: the reference aborts on them) and masked samples.  With the
// identity sample order a record's own 32-bit words are the input words: the kernel is a 4-byte-granular transpose plus a
// permutation of the 2-bit fields inside each word.
// ------------------------------------------------------------------------------------------------
constexpr int DEC8_KBS = 8;   // k-blocks per CTA: 512 samples = 128 packed bytes per record, one cache line
template <bool IDG>      // IDG: the codes are the genotypes (gmap = 0, 1, 2: x itself) — no remapping
__global__ void __launch_bounds__(512, 3) decode_2b_tile_kernel(const DecodeArgs a, uint32_t* __restrict__ X2, int NKB) {
    // k-block groups fastest: consecutive CTAs read consecutive 128-byte segments of the SAME 128 records (DRAM rows and the sectors of a
    // misaligned record stay hot), not the same segment of records 160 KB apart
    const int J = blockIdx.y, kb0 = blockIdx.x * DEC8_KBS;
    const int r = threadIdx.x & 127, ch = threadIdx.x >> 7;
    const int64_t snp = (int64_t)J * 128 + r;
    const bool live = snp < a.nsnp;
    const uint8_t* rec = a.packed + (live ? (a.vmap ? (int64_t)a.vmap[snp] : snp) : 0) * a.bpv;
    const int kend = min(kb0 + DEC8_KBS, NKB);
    // genotype -> code by bit planes: code bit 0 / bit 1 set for the genotypes whose gmap entry has it
    const uint32_t E = 0x55555555u;
    uint32_t ml[3], mh[3];
#pragma unroll
    for (int gq = 0; gq < 3; gq++) { const uint32_t c = (a.gmap >> (8 * gq)) & 3u; ml[gq] = (c & 1u) ? E : 0u; mh[gq] = (c & 2u) ? E : 0u; }
    // identity sample order: the CTA's 128-byte segment of every record is staged in shared memory with coalesced loads
    // (a warp reads one record's segment), then each thread picks its 4 packed bytes as one word; rows padded to 33 words
    __shared__ uint32_t seg[128][33];
    const bool staged = (a.order == nullptr);
    if (staged) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int64_t b0 = (int64_t)kb0 * (QF_KB / 4) + 4 * lane;          // first of this lane's 4 bytes within a record
        // the lane's four bytes of a record through ALIGNED 32-bit loads and a funnel shift (records start at any byte: bpv = ceil(N_raw / 4));
        // the second word is read only when bytes of this record live in it, so nothing beyond the record's last aligned word is touched.
        // Eight records per warp, all of their loads in flight before the first use (the loop is latency-bound otherwise)
        uint32_t w0v[8], w1v[8], shv[8];
        int64_t leftv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int rr = warp + 16 * i;
            const int64_t s2 = (int64_t)J * 128 + rr;
            w0v[i] = w1v[i] = 0u; shv[i] = 0u; leftv[i] = 0;
            if (s2 < a.nsnp && b0 < a.bpv) {
                const uint8_t* p = a.packed + (a.vmap ? (int64_t)a.vmap[s2] : s2) * a.bpv + b0;
                const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3);
                const uint32_t* pw = reinterpret_cast<const uint32_t*>(p - sh);
                const int64_t left = a.bpv - b0;                                    // bytes of the record from p on (>= 1)
                w0v[i] = __ldg(pw);
                if (sh != 0 && left > (int64_t)(4 - sh)) w1v[i] = __ldg(pw + 1);
                shv[i] = sh; leftv[i] = left;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t word = __funnelshift_r(w0v[i], w1v[i], 8 * shv[i]);
            if (leftv[i] < 4) word = leftv[i] > 0 ? (word & ((1u << (8 * (int)leftv[i])) - 1u)) : 0u;
            seg[warp + 16 * i][lane] = word;
        }
        __syncthreads();
    }
    // this thread's words: X2[pair][kb][tile][ch][r], 1024 words from one k-block to the next
    uint32_t* dst = X2 + ((((int64_t)(J >> 1) * NKB + kb0) * 2 + (J & 1)) * 4 + ch) * 128 + r;
    for (int kb = kb0; kb < kend; kb++, dst += 1024) {
        uint32_t w = 0;              // genotypes of the 16 samples, sample j at bits 2j (the pgen order)
        uint32_t keep = 0xffffffffu; // fields of real, unmasked samples
        const int64_t k0 = (int64_t)kb * QF_KB + ch * 16;
        if (live && k0 < a.n) {
            if (staged) w = seg[r][(kb - kb0) * 4 + ch];
            else {
#pragma unroll
                for (int b = 0; b < 16; b++) {
                    const int64_t k = k0 + b;
                    if (k < a.n) {
                        const int64_t ro = (int64_t)a.order[k];
                        w |= (uint32_t)((__ldg(rec + (ro >> 2)) >> (2 * (int)(ro & 3))) & 3) << (2 * b);
                    }
                }
            }
            if (k0 + 16 > a.n) keep = (1u << (2 * (int)(a.n - k0))) - 1u;        // the record's bits beyond the last sample are not genotypes
            if (a.smask) {
                // sixteen 0xFF / 0x00 bytes -> sixteen 11 / 00 fields (k0 is a multiple of 16; the mask array is padded to 256)
                const uint4 mk = __ldg(reinterpret_cast<const uint4*>(a.smask + k0));
                const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
                uint32_t f = 0;
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    uint32_t nib = (((mw[m] & 0x01010101u) * 0x01020408u) >> 24) & 0xFu;
                    nib = (nib | (nib << 2)) & 0x33u;
                    nib = (nib | (nib << 1)) & 0x55u;
                    f |= (nib * 3u) << (8 * m);
                }
                keep &= f;
            }
        } else keep = 0;
        w &= keep;
        const uint32_t lo = w & E, hi = (w >> 1) & E, miss = lo & hi;
        if (miss) atomicMin(a.missing_flag, (int)(a.snp_offset + snp));           // code 3 inside the kept fields: a missing call
        uint32_t x;
        if (IDG) x = w & ~(miss * 3u);
        else {
            const uint32_t is0 = ~(lo | hi) & (keep & E), is1 = lo & ~hi, is2 = hi & ~lo;
            x = ((is0 & ml[0]) | (is1 & ml[1]) | (is2 & ml[2])) | (((is0 & mh[0]) | (is1 & mh[1]) | (is2 & mh[2])) << 1);
        }
        // field j: bit 2j -> bit 2 (j / 4) + 8 (j % 4): the 4 x 4 transpose of the 2-bit fields
        uint32_t tq = (x ^ (x >> 6)) & 0x00CC00CCu; x ^= tq ^ (tq << 6);
        tq = (x ^ (x >> 12)) & 0x0000F0F0u; x ^= tq ^ (tq << 12);
        *dst = x;
    }
}

int launch_decode_i8(const DecodeArgs& a, uint32_t* X2, int NJ, int NKB, cudaStream_t st) {
    if (NJ <= 0) return 0;
    const dim3 grid((unsigned)((NKB + DEC8_KBS - 1) / DEC8_KBS), (unsigned)NJ);
    if ((a.gmap & 0xffffffu) == 0x020100u) decode_2b_tile_kernel<true><<<grid, 512, 0, st>>>(a, X2, NKB);
    else decode_2b_tile_kernel<false><<<grid, 512, 0, st>>>(a, X2, NKB);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("decode_i8 launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

// code map of a tile set with integer triple a[3] (value by genotype): code 0 is value 0 (missing, masked and zero-valued genotypes);
// when every value lies in 0..3 the code IS the value (vt = 0, 1, 2, 3: the expanders' one-shift path), else the non-zero values take
// codes 1.. in order of appearance.  gmap byte g = code of genotype g, vt byte c = int8 value of code c.
void tile_code_map(const int a[3], uint32_t* gmap, uint32_t* vt) {
    bool small = true;
    for (int gq = 0; gq < 3; gq++) small = small && a[gq] >= 0 && a[gq] <= 3;
    uint32_t gm = 0, v = 0;
    if (small) {
        v = QF_VT_IDENTITY;
        for (int gq = 0; gq < 3; gq++) gm |= (uint32_t)a[gq] << (8 * gq);
    } else {
        int vals[4] = {0, 0, 0, 0}, nv = 1;
        for (int gq = 0; gq < 3; gq++) {
            int c = 0;
            if (a[gq] != 0) {
                for (int k = 1; k < nv; k++) if (vals[k] == a[gq]) c = k;
                if (c == 0) { c = nv; vals[nv++] = a[gq]; }
            }
            gm |= (uint32_t)c << (8 * gq);
        }
        for (int k = 1; k < 4; k++) v |= ((uint32_t)vals[k] & 0xffu) << (8 * k);
    }
    *gmap = gm; *vt = v;
}

// ------------------------------------------------------------------------------------------------
// setup: T = tril(V^-1) with half diagonal  ->  per-row power-of-two scale + Q balanced base-256 digit planes,
//        packed as the B operand:  T8[row tile I][kb < 4(I+1)][plane q][k16 chunk 4][row group 32][8 rows][16 B]
// ------------------------------------------------------------------------------------------------
// power-of-two scale of a group of values with largest magnitude m: the leading balanced base-256 digit of v / scale is
// round(256 v / scale) and must fit int8, i.e. stay <= 127 — m / scale < 0.5 is NOT enough (it allows 128).
__device__ __forceinline__ double digit_scale(double m) {
    if (!(m > 0.0)) return 1.0;
    int e = 0;
    frexp(m, &e);                              // m = f * 2^e, f in [0.5, 1)
    double sc = ldexp(1.0, e + 1);             // m / sc = f / 2 in [0.25, 0.5)
    if (m * 256.0 > 127.0 * sc) sc *= 2.0;     // f > 127/128: one more bit of headroom
    return sc;
}

// V: rows row0 .. row0+255 of V^-1 (column-major, leading dimension ld), columns 0 .. row
__global__ void qf_rowscale_kernel(const double* __restrict__ V, int64_t n, int64_t ld, int64_t row0, const double* __restrict__ ca,
                                   const double* __restrict__ cb, int full, double* __restrict__ rowscale) {
    const int64_t il = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;    // 256 threads in all: one per row of the tile
    const int64_t i = row0 + il;
    double m = 0.0;
    if (i < n) {
        const double ai = ca ? ca[i] : 1.0;
        if (full) {
            for (int64_t j = 0; j < n; j++) m = fmax(m, fabs(ai * V[il + j * ld] * (cb ? cb[j] : 1.0)));
        } else {
            for (int64_t j = 0; j < i; j++) m = fmax(m, fabs(ai * V[il + j * ld] * (cb ? cb[j] : 1.0)));
            m = fmax(m, 0.5 * fabs(ai * V[il + i * ld] * (cb ? cb[i] : 1.0)));
        }
    }
    // one scale per 16 consecutive rows (= one tcgen05.ld chunk of the K2q epilogue, which then reduces the chunk in int64)
    for (int off = 8; off >= 1; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    rowscale[i] = digit_scale(m);
}

__global__ void __launch_bounds__(256) qf_slice_pack_kernel(const double* __restrict__ V, int64_t n, int64_t ld, const double* __restrict__ ca,
                                                           const double* __restrict__ cb, const double* __restrict__ rowscale, int Q, int I,
                                                           int full, int NKB, int8_t* __restrict__ T8) {
    const int kb = blockIdx.x;
    const int rl = threadIdx.x;                         // row within the 256-row tile
    const int64_t i = (int64_t)I * 256 + rl;
    const double inv = (i < n) ? 1.0 / rowscale[i] : 0.0;
    const double ai = (ca && i < n) ? ca[i] : 1.0;
    const double up = ldexp(1.0, 8 * Q);
    int8_t* blk = T8 + ((full ? (int64_t)I * NKB : (int64_t)2 * I * (I + 1)) + kb) * Q * QF_B_BYTES;
    for (int ch = 0; ch < 4; ch++) {
        uint32_t w[7][4];
        for (int q = 0; q < Q; q++) w[q][0] = w[q][1] = w[q][2] = w[q][3] = 0;
        for (int b = 0; b < 16; b++) {
            const int64_t j = (int64_t)kb * QF_KB + ch * 16 + b;
            double t = 0.0;
            if (full) { if (i < n && j < n) t = ai * V[rl + j * ld] * (cb ? cb[j] : 1.0); }
            else if (i < n && j <= i) t = ((j == i) ? 0.5 : 1.0) * (ai * V[rl + j * ld] * (cb ? cb[j] : 1.0));
            long long ri = __double2ll_rn(t * inv * up);
            for (int q = Q - 1; q >= 0; q--) {
                const long long d = ((ri + 128) & 255) - 128;
                ri = (ri - d) >> 8;
                w[q][b >> 2] |= (uint32_t)((int)d & 0xff) << (8 * (b & 3));
            }
        }
        for (int q = 0; q < Q; q++) {
            uint4* dst = reinterpret_cast<uint4*>(blk + (int64_t)q * QF_B_BYTES + ((ch * 32 + (rl >> 3)) * 8 + (rl & 7)) * 16);
            *dst = make_uint4(w[q][0], w[q][1], w[q][2], w[q][3]);
        }
    }
}

int launch_qf_slice(const double* Vtile, int64_t n, int64_t ld, int I, int T256, int Q, const double* ca, const double* cb, int full, double* rowscale, int8_t* T8,
                    cudaStream_t st) {
    qf_rowscale_kernel<<<2, 128, 0, st>>>(Vtile, n, ld, (int64_t)I * 256, ca, cb, full, rowscale);
    qf_slice_pack_kernel<<<(unsigned)(full ? 4 * T256 : 4 * (I + 1)), 256, 0, st>>>(Vtile, n, ld, ca, cb, rowscale, Q, I, full, 4 * T256, T8);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("qf_slice launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

// zero the strict upper triangle of a dense column-major matrix (caller-supplied L^-1 may carry anything there)
__global__ void zero_upper_kernel(double* Z, int64_t n, int64_t ld) {
    const int64_t j = blockIdx.x;
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < j; i += (int64_t)gridDim.y * blockDim.x) Z[i + j * ld] = 0.0;
}
int launch_zero_upper(double* Z, int64_t n, int64_t ld, cudaStream_t st) {
    if (n <= 1) return 0;
    zero_upper_kernel<<<dim3((unsigned)n, 32), 256, 0, st>>>(Z, n, ld);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("zero_upper launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

// ================================================================================================
// K4q: a tile pair (2 x 128 SNPs x n, 2-bit codes) against a column block of G = C L^-T [Qc | r | R_f] split into Q digit planes:
//      out[snp][col] = g_snp' G_col  exactly (up to the 2^-48 truncation of G below each column's power-of-two scale):
//      u = Qc' w (cr columns), b = r' w, and b* for every permutation seed in ONE pass over the genotype tile per column block.
// Same ring, roles and operand economy as K2q (producer, four expander warps, one MMA issuer per SNP tile): a k-block costs 4 KB of
// codes + Q * NB * 64 B of planes for 2 x 128 x (Q NB) x 64 MACs.  The Q planes of the block sit side by side along N in shared
// memory and in TMEM (plane q of tile t at column 256 t + q NB, Q NB <= 256), so one instruction covers all planes and the A tile
// is read from shared memory once per K step.
// Warps as in K2q: 0-3 expanders, 4-11 epilogue (4 per SNP tile; thread = TMEM lane = SNP), 12 TMA producer, 13-14 MMA issuers.
// ================================================================================================
#define XG_TMEM_LD8(v, taddr)                                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                          \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) \
                 : "r"(taddr))

__global__ void __launch_bounds__(QF_THREADS, 1) xg_i8_kernel(const XgArgs g) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint32_t* tmem_slot = qf_tmem_slot(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) qf_bars_init(smem);
    if (warp == QF_PROD_WARP0 + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(qsmem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = qsmem_u32(smem);
    const QfBars bars = qf_bars(smem);
    const int NB = g.NB, Q = g.Q;
    const int n_total = Q * NB;                                  // MMA N: all planes of the column block
    const uint32_t b_bytes = (uint32_t)n_total * QF_KB;          // all planes of one k-block (<= 16 KB)
    const int nkb = g.NKB_live;
    const int64_t my_items = g.n_items > (int64_t)blockIdx.x ? (g.n_items - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (warp == QF_PROD_WARP0) {
        uint32_t stage = 0, phase = 0;
        for (int64_t item = blockIdx.x; item < g.n_items; item += gridDim.x) {
            const int JP = (int)(item / g.NCT), ct = (int)(item % g.NCT);   // column blocks of one tile pair back to back: its codes stay L2-hot
            const uint32_t* ap = g.X2 + (int64_t)JP * g.NKB * (2 * QF_A2_WORDS);
            const int8_t* bp = g.G8 + (int64_t)ct * g.NKB * b_bytes;
            for (int kb = 0; kb < nkb; kb++) {
                const uint32_t fb = bars.full + 8 * stage, dst = smem0 + stage * QF_STAGE_BYTES;
                qbar_wait_a(bars.empty + 8 * stage, phase ^ 1);
                if (elect_one()) {
                    qbar_expect_tx_a(fb, b_bytes + QF_P2_BYTES);
                    qtma_bulk_a(dst, bp, b_bytes, fb);
                    qtma_bulk_a(dst + QF_B_BYTES, ap, QF_P2_BYTES, fb);
                }
                bp += b_bytes; ap += 2 * QF_A2_WORDS;
                if (++stage == QF_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp > QF_PROD_WARP0 && warp <= QF_PROD_WARP0 + 2) {
        const int t = warp - (QF_PROD_WARP0 + 1);
        // B block: [k16 chunk 4][n_total / 8 groups][8 columns][16 B]  ->  LBO = n_total / 8 * 128 B between k16 chunks, SBO 128 B
        const uint32_t lbo_b = (uint32_t)(n_total / 8) * 128;
        const uint32_t idesc = QF_IDESC_M128 | (((uint32_t)n_total >> 3) << 17);
        const uint32_t d_hi = (128u >> 4) | (1u << 14);
        const uint32_t a_lo0 = (((smem0 + QF_SLOT0 + t * QF_A_BYTES) >> 4) & 0x3FFFu) | ((2048u >> 4) << 16);
        const uint32_t b_lo0 = ((smem0 >> 4) & 0x3FFFu) | ((lbo_b >> 4) << 16);
        const uint32_t acc_full_a = bars.acc_full + 8 * t, acc_empty_a = bars.acc_empty + 8 * t;
        const uint32_t tacc = tmem_base + t * 256;
        uint32_t stage = 0, phase = 0, slot = 0, sphase = 0, aphase = 0;
        for (int64_t item = blockIdx.x; item < g.n_items; item += gridDim.x) {
            const int JP = (int)(item / g.NCT);
            const bool mine = 2 * JP + t < g.NJ;
            qbar_wait_a(acc_empty_a, aphase ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            for (int kb = 0; kb < nkb; kb++) {
                qbar_wait_a(bars.full + 8 * stage, phase);
                qbar_wait_a(bars.xfull + 8 * slot, sphase);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t al = a_lo0 + slot * (QF_SLOT_BYTES >> 4), bl = b_lo0 + stage * (QF_STAGE_BYTES >> 4);
                if (mine && elect_one()) {
                    umma_i8_w(tacc, al, bl, d_hi, idesc, kb > 0 ? 1u : 0u);
                    umma_i8_w(tacc, al + (4096 >> 4), bl + ((2 * lbo_b) >> 4), d_hi, idesc, 1u);
                }
                if (elect_one()) { umma_commit_a(bars.empty + 8 * stage); umma_commit_a(bars.xempty + 8 * slot); }
                if (++stage == QF_STAGES) { stage = 0; phase ^= 1; }
                if (++slot == QF_SLOTS) { slot = 0; sphase ^= 1; }
            }
            if (elect_one()) umma_commit_a(acc_full_a);
            aphase ^= 1;
        }
    } else if (warp < QF_SLOTS) {
        qf_expand_stages(smem, bars, warp, lane, my_items * nkb, g.vt);
    } else if (warp >= QF_EPI_WARP0 && warp < QF_PROD_WARP0) {
        const int quad = warp & 3;
        const int t = (warp - QF_EPI_WARP0) >> 2;
        const int snp_l = quad * 32 + lane;
        uint64_t* acc_full = reinterpret_cast<uint64_t*>(smem + QF_BAR0) + 2 * QF_STAGES + 2 * QF_SLOTS + t;
        uint64_t* acc_empty = acc_full + 2;
        uint32_t aphase = 0;
        for (int64_t item = blockIdx.x; item < g.n_items; item += gridDim.x) {
            const int JP = (int)(item / g.NCT), ct = (int)(item % g.NCT);
            const int J = 2 * JP + t;
            const int64_t snp = (int64_t)J * 128 + snp_l;
            const bool live = J < g.NJ && snp < g.nsnp;
            qbar_wait(acc_full, aphase);
            aphase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (J < g.NJ) {
                const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + t * 256;
                double* orow = g.out + snp * g.ld_out + (int64_t)ct * NB;
                for (int c0 = 0; c0 < NB; c0 += 8) {
                    uint32_t v[6][8];
#pragma unroll
                    for (int q = 0; q < 6; q++)
                        if (q < Q) XG_TMEM_LD8(v[q], trow + q * NB + c0);
                    asm volatile("tcgen05.wait::ld.sync.aligned;");
                    double val[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        // planes pairwise exact in int64, then FP64: sum_q d_q 256^-(q+1)
                        double acc = 0.0;
#pragma unroll
                        for (int qq = 2; qq >= 0; qq--) {
                            if (2 * qq < Q) {
                                const long long comb = (long long)(int)v[2 * qq][j] * 256 + ((2 * qq + 1 < Q) ? (long long)(int)v[2 * qq + 1][j] : 0);
                                // |comb| < 2^40: bias trick instead of I2F (16/clk/SM)
                                acc += (__longlong_as_double(0x4330000000000000LL + (comb + (1LL << 51))) - 6755399441055744.0) * g.pair_scale[qq];
                            }
                        }
                        val[j] = acc * __ldg(g.colscale + ct * NB + c0 + j);
                    }
                    if (live) {
                        // NB and ld_out are multiples of 8: 16-byte stores, eight consecutive columns of the SNP's row
                        double2* o2 = reinterpret_cast<double2*>(orow + c0);
#pragma unroll
                        for (int j = 0; j < 4; j++) o2[j] = make_double2(val[2 * j], val[2 * j + 1]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) qbar_arrive(acc_empty);
        }
    }
    __syncthreads();
    if (warp == QF_PROD_WARP0 + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// column block width of K4q: the largest NB with Q * NB <= 256 that keeps the MMA N = Q * NB a multiple of 16
int xg_block_cols(int Q) { return (Q % 2 == 0) ? (256 / Q) / 8 * 8 : (256 / Q) / 16 * 16; }

int launch_xg(const XgArgs& g, int num_sms, cudaStream_t st) {
    static DeviceOnce configured;
    if (configured.need()) {
        cudaError_t e = cudaFuncSetAttribute(xg_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, QF_SMEM);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(xg): %s", cudaGetErrorString(e)); return -3; }
        configured.set();
    }
    if (g.NB % 8 != 0 || g.NB < 8 || g.Q < 1 || g.Q > 6 || g.Q * g.NB > 256 || (g.Q * g.NB) % 16 != 0 || g.ld_out % 8 != 0) { set_error("xg: bad column block %d / planes %d", g.NB, g.Q); return -1; }
    const int64_t grid = g.n_items < num_sms ? g.n_items : num_sms;
    if (grid <= 0) return 0;
    xg_i8_kernel<<<(unsigned)grid, QF_THREADS, QF_SMEM, st>>>(g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("xg launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

// ---- setup: G (n x NC dense, column-major) -> per-column power-of-two scale + Q digit planes, packed per column block:
//      G8[col block ct][kb][plane q][k16 chunk 4][col group NB/8][8 cols][16 B]
__global__ void xg_colscale_kernel(const double* __restrict__ G, int64_t ld, int c1, const double* __restrict__ G2, int64_t ld2, int64_t n, int NC,
                                   const double* __restrict__ rowmul, double* __restrict__ colscale) {
    const int col = blockIdx.x;
    __shared__ double sh[8];
    double m = 0.0;
    if (col < NC) {
        const double* src = (col < c1) ? G + (int64_t)col * ld : G2 + (int64_t)(col - c1) * ld2;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, fabs(src[i] * (rowmul ? rowmul[i] : 1.0)));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) m = fmax(m, sh[w]);
        colscale[col] = digit_scale(m);
    }
}

__global__ void __launch_bounds__(256) xg_slice_pack_kernel(const double* __restrict__ G, int64_t ld, int c1, const double* __restrict__ G2, int64_t ld2, int64_t n,
                                                           int NC, int NB, int Q, int NKB, const double* __restrict__ rowmul,
                                                           const double* __restrict__ colscale, int8_t* __restrict__ G8) {
    const int ct = blockIdx.y, kb = blockIdx.x;
    const int plane_bytes = NB * QF_KB;
    int8_t* blk = G8 + ((int64_t)ct * NKB + kb) * Q * plane_bytes;
    const double up = ldexp(1.0, 8 * Q);
    // one thread per (column of the block, k16 chunk)
    for (int e = threadIdx.x; e < NB * 4; e += blockDim.x) {
        const int cl = e % NB, ch = e / NB;
        const int col = ct * NB + cl;
        const double inv = (col < NC) ? 1.0 / colscale[col] : 0.0;
        const double* src = (col < c1) ? G + (int64_t)col * ld : G2 + (int64_t)(col - c1) * ld2;
        uint32_t w[6][4];
        for (int q = 0; q < 6; q++) w[q][0] = w[q][1] = w[q][2] = w[q][3] = 0;
        for (int b = 0; b < 16; b++) {
            const int64_t k = (int64_t)kb * QF_KB + ch * 16 + b;
            const double t = (col < NC && k < n) ? src[k] * (rowmul ? rowmul[k] : 1.0) : 0.0;
            long long ri = __double2ll_rn(t * inv * up);
            for (int q = Q - 1; q >= 0; q--) {
                const long long d = ((ri + 128) & 255) - 128;
                ri = (ri - d) >> 8;
                w[q][b >> 2] |= (uint32_t)((int)d & 0xff) << (8 * (b & 3));
            }
        }
        for (int q = 0; q < Q; q++) {
            const int nidx = q * NB + cl;       // planes side by side along N
            uint4* dst = reinterpret_cast<uint4*>(blk + ((ch * (Q * NB / 8) + (nidx >> 3)) * 8 + (nidx & 7)) * 16);
            *dst = make_uint4(w[q][0], w[q][1], w[q][2], w[q][3]);
        }
    }
}

int launch_xg_slice(const double* G, int64_t ld, int c1, const double* G2, int64_t ld2, int64_t n, int NC, int NB, int ntiles, int Q, int NKB, const double* rowmul,
                    double* colscale, int8_t* G8, cudaStream_t st) {
    xg_colscale_kernel<<<(unsigned)(NB * ntiles), 256, 0, st>>>(G, ld, c1, G2, ld2, n, NC, rowmul, colscale);
    xg_slice_pack_kernel<<<dim3((unsigned)NKB, (unsigned)ntiles), 256, 0, st>>>(G, ld, c1, G2, ld2, n, NC, NB, Q, NKB, rowmul, colscale, G8);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("xg_slice launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

}  // namespace flx
