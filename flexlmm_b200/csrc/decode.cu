// decode.cu — K1 decode_2bit_tile (sm_100a).
//
// Replaces pgenlibr::ReadHardcalls + buf[pgen_order] + model.frame/model.matrix for the SNP-dependent
// columns (fit_model.nf:123-127).  Input: packed 2-bit hard-call records (pgen mode-0x02 record layout:
// sample i in bits 2*(i&3) of byte i>>2; 0/1/2 = ALT count, 3 = missing).  Output: the design columns of
// every SNP of the batch, FP64, written directly in the PACKED PANEL layout (common.cuh) that the
// whitening GEMM consumes as its B operand — the store index of a thread is its linear element index, so
// every warp store is one contiguous 256-byte segment.
//
// Missing-genotype rule: a code 3 in a tested variant is fatal in the reference; the kernel
// records the first offending SNP of the batch in *missing_flag and the host turns that into
// FLX_ERR_MISSING_GENOTYPE.
//
// HBM-bound: algorithmic bytes per SNP = ceil(N_raw/4) read + 8 * n_pad * s written.
#include "common.cuh"
#include "kernels.h"

namespace flx {

constexpr int DEC_THREADS = 256;
constexpr int DEC_CPB = 8;  // k-chunks per CTA (128 samples)

__global__ void __launch_bounds__(DEC_THREADS) decode_2bit_tile_kernel(const DecodeArgs a) {
    const int J = blockIdx.x;
    const int kc0 = blockIdx.y * DEC_CPB;
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const TileShape ts = make_tile_shape(a.s);

    // the two j-blocks this thread serves: w and w + 8
    int64_t snp[2], rec[2];
    int term[2];
    bool valid[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int blk = w + 8 * h;
        const int gi = blk / ts.s;
        term[h] = blk - gi * ts.s;
        snp[h] = (int64_t)J * ts.snps + 8 * gi + (lane >> 2);
        valid[h] = (blk < ts.blocks) && (snp[h] < a.nsnp);
        rec[h] = (valid[h] && a.vmap) ? (int64_t)a.vmap[snp[h]] : snp[h];
    }
    double* xt = a.X + ((int64_t)J * a.KC) * CHUNK;
    const int kend = min(kc0 + DEC_CPB, a.KC);
    for (int kc = kc0; kc < kend; kc++) {
        double* xc = xt + (int64_t)kc * CHUNK;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            const int64_t k = (int64_t)kc * KCH + 4 * ks + (lane & 3);
            const bool in_n = k < a.n;
            int64_t ro = 0;
            if (in_n) ro = a.order ? (int64_t)a.order[k] : k;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                double v = 0.0;
                if (in_n && valid[h]) {
                    const uint8_t byte = __ldg(a.packed + rec[h] * a.bpv + (ro >> 2));
                    const int code = (byte >> (2 * (int)(ro & 3))) & 3;
                    const uint32_t d16 = a.dosage ? (uint32_t)__ldg(a.dosage + rec[h] * a.dos_ld + ro) : 0xffffu;
                    if (d16 != 0xffffu) {
                        // pgenlibr::Read: x = dosage / 16384 (exact); the term as model.matrix evaluates it on a real x
                        const double x = (double)d16 * 0.00006103515625;
                        const int tt = term[h];
                        const bool un = (a.uniform_mask >> tt) & 1u;
                        const double* lt = a.lut + (int64_t)(tt * 3) * a.n_pad + k;
                        if ((a.dos_indicator >> tt) & 1u) {
                            const int g = a.dos_g[tt], o = (g == 0) ? 1 : 0;
                            v = (x == (double)g) ? (un ? a.uni[tt][g] : __ldg(lt + (int64_t)g * a.n_pad)) : (un ? a.uni[tt][o] : __ldg(lt + (int64_t)o * a.n_pad));
                        } else {
                            const double v0 = un ? a.uni[tt][0] : __ldg(lt), v1 = un ? a.uni[tt][1] : __ldg(lt + a.n_pad);
                            v = (v0 == 0.0) ? x * v1 : v0 + x * (v1 - v0);      // x, x:cov exactly as R forms them; general affine terms in one rounding more
                        }
                    } else if (code == 3) {
                        atomicMin(a.missing_flag, (int)(a.snp_offset + snp[h]));
                    } else if ((a.uniform_mask >> term[h]) & 1u) {
                        v = a.uni[term[h]][code];
                    } else {
                        v = __ldg(a.lut + ((int64_t)(term[h] * 3 + code)) * a.n_pad + k);
                    }
                    if (a.codes_out && term[h] == 0) a.codes_out[snp[h] * a.n + k] = (uint8_t)code;
                }
                xc[((ks * 16 + w + 8 * h) << 5) + lane] = v;
            }
        }
    }
}

int launch_decode(const DecodeArgs& a, cudaStream_t st) {
    if (a.NJ <= 0) return 0;
    dim3 grid((unsigned)a.NJ, (unsigned)((a.KC + DEC_CPB - 1) / DEC_CPB));
    decode_2bit_tile_kernel<<<grid, DEC_THREADS, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("decode launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

}  // namespace flx
