// pgen_expand.cu — K0: .pgen mode-0x10 variant records -> packed 2-bit hard calls, on the device.
//
// Replaces the decompression half of pgenlibr::ReadHardcalls (fit_model.nf:123; plink-ng pgenlib PgrGet1: the main
// genotype track of a record, vrtype & 7):
//   0      raw 2-bit
//   1      1-bit array over the two commonest genotypes + difflist of the others
//   2, 3   difflist against the nearest earlier non-LD variant (3: 0 <-> 2 swapped afterwards)
//   4-7    difflist against an all-(vrtype & 3) row
// A difflist is: varint length, per group of 64 entries the first sample id (1-4 bytes), ngroups-1 group byte counts
// (bytes of the group's delta stream minus 63), the 2-bit codes of all entries, then the varint sample-id deltas.
// One warp per record: the row is staged in shared memory, lanes decode difflist groups in parallel (the group byte
// counts give every group its start in the delta stream: a warp scan), 2-bit fields are set with shared-memory atomics.
// Two launches per batch: records that stand alone (pass 0), then LD records on top of their finished base rows (pass 1).
// HBM bound and tiny next to the rest of the path: its point is that only COMPRESSED bytes cross PCIe and no host thread
// touches a genotype.
#include "common.cuh"
#include "kernels.h"

namespace flx {

constexpr int PX_WARPS = 4;

__device__ __forceinline__ bool px_varint(const uint8_t*& p, const uint8_t* end, uint32_t& out) {
    out = 0;
    for (int sh = 0; sh <= 28 && p < end; sh += 7) {
        const uint8_t b = *p++;
        out |= (uint32_t)(b & 127) << sh;
        if (!(b & 128)) return true;
    }
    return false;
}

__device__ __forceinline__ void px_set2(uint32_t* row, uint32_t sid, uint32_t v) {
    const int sh = 2 * (sid & 15);
    atomicAnd(&row[sid >> 4], ~(3u << sh));
    atomicOr(&row[sid >> 4], v << sh);
}

__device__ __forceinline__ void px_fail(int* err, int entry, int code) { atomicCAS(err, 0, (entry << 4) | code); }

__global__ void __launch_bounds__(PX_WARPS * 32) pgen_expand_kernel(const PgenExpandArgs a) {
    extern __shared__ uint32_t px_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * PX_WARPS + warp;
    if (e >= a.E) return;
    const int t = a.vrtype[e] & 7;
    const bool ld = (t == 2 || t == 3);
    if (ld != (a.pass == 1)) return;
    const int64_t bpv = a.bpv;
    const int nwords = (int)((bpv + 3) / 4);
    uint32_t* row = px_smem + (size_t)warp * a.row_words;
    const uint8_t* p = a.blob + a.rec_off[e];
    const uint8_t* const rec0 = p;
    const uint8_t* end = p + a.rec_len[e];
    const uint32_t n = a.sample_ct;
    uint32_t track_end = 0;     // bytes of the main genotype track (the dosage track, if any, follows it)

    // ---- base row
    if (t == 0) {
        if ((int64_t)a.rec_len[e] < bpv) { if (lane == 0) px_fail(a.err, e, 1); return; }
        uint8_t* rb = reinterpret_cast<uint8_t*>(row);
        for (int64_t i = lane; i < (int64_t)nwords * 4; i += 32) rb[i] = (i < bpv) ? p[i] : 0;
        track_end = (uint32_t)bpv;
    } else if (t == 1) {
        if (p + 1 + (n + 7) / 8 > end) { if (lane == 0) px_fail(a.err, e, 2); return; }
        const uint32_t c2 = p[0];
        const uint32_t lo = c2 >> 2, hi = (c2 >> 2) + (c2 & 3);
        const uint8_t* bits = p + 1;
        for (int w = lane; w < nwords; w += 32) {
            // 16 samples per output word = 2 bytes of the bit array
            const uint32_t nb = (n + 7) / 8;
            uint32_t b16 = ((uint32_t)(2 * w) < nb ? bits[2 * w] : 0u) | (((uint32_t)(2 * w + 1) < nb ? bits[2 * w + 1] : 0u) << 8);
            uint32_t outw = 0;
#pragma unroll
            for (int k = 0; k < 16; k++) outw |= (((b16 >> k) & 1u) ? hi : lo) << (2 * k);
            row[w] = outw;
        }
        p += 1 + (n + 7) / 8;
    } else if (!ld) {
        for (int w = lane; w < nwords; w += 32) row[w] = (uint32_t)(t & 3) * 0x55555555u;
    } else {
        const int br = a.base_row[e];
        const uint8_t* src = (br >= 0) ? a.out + (int64_t)br * bpv : a.scratch + (int64_t)(-br - 1) * bpv;
        uint8_t* rb = reinterpret_cast<uint8_t*>(row);
        for (int64_t i = lane; i < (int64_t)nwords * 4; i += 32) rb[i] = (i < bpv) ? src[i] : 0;
    }
    __syncwarp();

    // ---- difflist
    if (t != 0) {
        uint32_t dl = 0;
        bool ok = px_varint(p, end, dl);          // every lane decodes the (tiny) header redundantly
        if (!ok || dl > n) { if (lane == 0) px_fail(a.err, e, 3); return; }
        track_end = (uint32_t)(p - rec0);
        if (dl > 0) {
            const uint32_t ngroups = (dl + 63) / 64;
            const int idb = n < 256u ? 1 : n < 65536u ? 2 : n < 16777216u ? 3 : 4;   // pgenlib BytesToRepresentNzU32(raw_sample_ct) = bsr / 8 + 1
            const uint8_t* first_ids = p;
            const uint8_t* cnts = first_ids + (uint64_t)ngroups * idb;
            const uint8_t* rare = cnts + (ngroups - 1);
            const uint8_t* deltas = rare + (dl + 3) / 4;
            if (deltas > end) { if (lane == 0) px_fail(a.err, e, 4); return; }
            uint32_t carry = 0;
            bool bad = false;
            for (uint32_t g0 = 0; g0 < ngroups; g0 += 32) {
                const uint32_t g = g0 + lane;
                const uint32_t c = (g + 1 < ngroups) ? (uint32_t)cnts[g] + 63u : 0u;   // bytes of group g's deltas (last group: open ended)
                uint32_t incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const uint32_t start = carry + incl - c;
                if (g < ngroups) {
                    uint32_t sid = 0;
                    for (int b = 0; b < idb; b++) sid |= (uint32_t)first_ids[(uint64_t)g * idb + b] << (8 * b);
                    const uint8_t* q = deltas + start;
                    const uint32_t e0 = g * 64, e1 = min(e0 + 64, dl);
                    for (uint32_t en = e0; en < e1; en++) {
                        if (en > e0) {
                            uint32_t d;
                            if (!px_varint(q, end, d)) { bad = true; break; }
                            sid += d;
                        }
                        if (sid >= n) { bad = true; break; }
                        px_set2(row, sid, (rare[en >> 2] >> (2 * (en & 3))) & 3u);
                    }
                    // the group byte counts must agree with what was parsed
                    if (!bad && g + 1 < ngroups && (uint32_t)(q - deltas) != start + c) bad = true;
                    if (g + 1 == ngroups) track_end = (uint32_t)(q - rec0);      // the last group ends the track
                }
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (__any_sync(0xffffffffu, bad)) { if (lane == 0) px_fail(a.err, e, 5); return; }
            track_end = __reduce_max_sync(0xffffffffu, track_end);
        }
    }
    if (a.main_end && lane == 0) a.main_end[e] = track_end;
    __syncwarp();
    if (t == 3) {
        // inverted LD: 0 <-> 2, het and missing unchanged: flip the high bit where the low bit is 0
        for (int w = lane; w < nwords; w += 32) { const uint32_t gq = row[w]; row[w] = gq ^ ((~gq & 0x55555555u) << 1); }
        __syncwarp();
    }
    // ---- pad bits of the last byte are zero; write the row out
    const int r = a.out_row[e];
    uint8_t* dst = (r >= 0) ? a.out + (int64_t)r * bpv : a.scratch + (int64_t)(-r - 1) * bpv;
    const uint8_t* rb = reinterpret_cast<const uint8_t*>(row);
    const uint32_t tail = n & 3u;
    for (int64_t i = lane; i < bpv; i += 32) {
        uint8_t v = rb[i];
        if (tail && i == bpv - 1) v &= (uint8_t)((1u << (2 * tail)) - 1);
        dst[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Dosage track (pgenlibr::Read, fit_model.nf:26,123; pgenlib PgrGetD): vrtype bits 5-6
//   01  delta list of sample ids (a difflist without genotypes), then one uint16 per listed sample
//   11  bit array over all samples, then one uint16 per set bit
//   10  one uint16 for every sample (65535 = none)
// values 0..32768 = ALT dosage 0..2.  One warp per record; rows were pre-filled with 0xFFFF.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t px_u16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

__global__ void __launch_bounds__(PX_WARPS * 32) pgen_dosage_kernel(const PgenExpandArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * PX_WARPS + warp;
    if (e >= a.E) return;
    const uint32_t vt = a.vrtype[e];
    const int dm = (vt >> 5) & 3;
    const int r = a.out_row[e];
    if (dm == 0 || r < 0) return;
    if (vt & 0x90u) { if (lane == 0) px_fail(a.err, e, 6); return; }       // phased hard calls / phased dosages: not produced by the pipeline
    const uint32_t n = a.sample_ct;
    const uint8_t* p = a.blob + a.rec_off[e] + a.main_end[e];
    const uint8_t* end = a.blob + a.rec_off[e] + a.rec_len[e];
    uint16_t* dst = a.dos_out + (int64_t)r * a.dos_ld;
    if (dm == 2) {
        if (p + 2ull * n > end) { if (lane == 0) px_fail(a.err, e, 7); return; }
        for (uint32_t i = lane; i < n; i += 32) dst[i] = (uint16_t)px_u16(p + 2ull * i);
        return;
    }
    if (dm == 3) {
        const uint32_t nb = (n + 7) / 8;
        const uint8_t* vals = p + nb;
        if (vals > end) { if (lane == 0) px_fail(a.err, e, 7); return; }
        uint32_t carry = 0;
        bool bad = false;
        for (uint32_t w0 = 0; w0 * 32 < n; w0 += 32) {
            const uint32_t w = w0 + lane;
            uint32_t word = 0;
            for (int b = 0; b < 4; b++) if (4ull * w + b < nb) word |= (uint32_t)p[4ull * w + b] << (8 * b);
            if (32ull * w + 32 > n) word &= (32ull * w < n) ? ((1u << (n - 32 * w)) - 1u) : 0u;     // pad bits of the last byte
            const uint32_t cnt = __popc(word);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            uint32_t k = carry + incl - cnt;
            while (word) {
                const int b = __ffs(word) - 1;
                word &= word - 1;
                if (vals + 2ull * k + 2 > end) { bad = true; break; }
                dst[32ull * w + b] = (uint16_t)px_u16(vals + 2ull * k);
                k++;
            }
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (__any_sync(0xffffffffu, bad) && lane == 0) px_fail(a.err, e, 7);
        return;
    }
    // dm == 1: delta list
    uint32_t dl = 0;
    if (!px_varint(p, end, dl) || dl > n) { if (lane == 0) px_fail(a.err, e, 7); return; }
    if (dl == 0) return;
    const uint32_t ngroups = (dl + 63) / 64;
    const int idb = n < 256u ? 1 : n < 65536u ? 2 : n < 16777216u ? 3 : 4;
    const uint8_t* first_ids = p;
    const uint8_t* cnts = first_ids + (uint64_t)ngroups * idb;
    const uint8_t* deltas = cnts + (ngroups - 1);
    if (deltas > end) { if (lane == 0) px_fail(a.err, e, 7); return; }
    // the values follow the delta stream, whose last group is open ended: one lane walks that group first
    uint32_t pre = 0;
    for (uint32_t g = lane; g + 1 < ngroups; g += 32) pre += (uint32_t)cnts[g] + 63u;
    pre = __reduce_add_sync(0xffffffffu, pre);
    uint32_t vals_off = 0;
    if (lane == 0) {
        const uint8_t* q = deltas + pre;
        bool ok = q <= end;
        for (uint32_t en = (ngroups - 1) * 64 + 1; ok && en < dl; en++) { uint32_t d; ok = px_varint(q, end, d); }
        vals_off = ok ? (uint32_t)(q - deltas) : 0xffffffffu;
    }
    vals_off = __shfl_sync(0xffffffffu, vals_off, 0);
    const uint8_t* vals = deltas + vals_off;
    if (vals_off == 0xffffffffu || vals + 2ull * dl > end) { if (lane == 0) px_fail(a.err, e, 7); return; }
    uint32_t carry = 0;
    bool bad = false;
    for (uint32_t g0 = 0; g0 < ngroups; g0 += 32) {
        const uint32_t g = g0 + lane;
        const uint32_t c = (g + 1 < ngroups) ? (uint32_t)cnts[g] + 63u : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const uint32_t start = carry + incl - c;
        if (g < ngroups) {
            uint32_t sid = 0;
            for (int b = 0; b < idb; b++) sid |= (uint32_t)first_ids[(uint64_t)g * idb + b] << (8 * b);
            const uint8_t* q = deltas + start;
            const uint32_t e0 = g * 64, e1 = min(e0 + 64, dl);
            for (uint32_t en = e0; en < e1; en++) {
                if (en > e0) { uint32_t d; if (!px_varint(q, end, d)) { bad = true; break; } sid += d; }
                if (sid >= n) { bad = true; break; }
                dst[sid] = (uint16_t)px_u16(vals + 2ull * en);
            }
            if (!bad && g + 1 < ngroups && (uint32_t)(q - deltas) != start + c) bad = true;
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) px_fail(a.err, e, 7);
}

int launch_pgen_dosage(const PgenExpandArgs& a, cudaStream_t st) {
    if (a.E <= 0 || !a.dos_out || !a.main_end) return 0;
    pgen_dosage_kernel<<<(unsigned)((a.E + PX_WARPS - 1) / PX_WARPS), PX_WARPS * 32, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("pgen_dosage launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

int launch_pgen_expand(const PgenExpandArgs& a0, cudaStream_t st) {
    if (a0.E <= 0) return 0;
    PgenExpandArgs a = a0;
    a.row_words = (int)(((a.bpv + 3) / 4 + 3) / 4 * 4);
    const size_t smem = (size_t)PX_WARPS * a.row_words * sizeof(uint32_t);
    static std::atomic<size_t> configured_dev[64];          // per device: the attribute belongs to the current device
    std::atomic<size_t>& configured = configured_dev[DeviceOnce::dev()];
    if (smem > 48 * 1024 && smem > configured.load()) {
        if (smem > 227 * 1024) { set_error("pgen_expand: %u samples per record exceed the shared-memory row buffers", a.sample_ct); return -1; }
        cudaError_t e = cudaFuncSetAttribute(pgen_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(pgen_expand): %s", cudaGetErrorString(e)); return -3; }
        configured = smem;
    }
    const unsigned grid = (unsigned)((a.E + PX_WARPS - 1) / PX_WARPS);
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1 && !a.any_ld) break;
        a.pass = pass;
        pgen_expand_kernel<<<grid, PX_WARPS * 32, smem, st>>>(a);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("pgen_expand launch: %s", cudaGetErrorString(e)); return -3; }
    return 0;
}

}  // namespace flx
