/*
 * flexlmm_oracle.c — CPU restatement of flexlmm's FIT_MODEL hot loop.  TEST INFRASTRUCTURE ONLY
 * (see flexlmm_oracle.h).  PARITY UNPINNED at the R / pgenlibr boundary (no golden vectors exist in
 * the reference); pinned against SURVEY.md Appendix B known answers in tests/test_oracle.py.
 *
 * Plain C, single thread, written for fidelity to the reference's arithmetic ORDER, not for speed:
 *   - forwardsolve   : reference-BLAS dtrsm loop order (R's bundled BLAS), fit_model.nf:128
 *   - .lm.fit        : LINPACK dqrdc2 (R-modified, limited pivoting, tol 1e-7) + dqrsl, fit_model.nf:133
 *   - logLik.lm      : long-double sum like R's rsum(), fit_model.nf:134
 *   - pchisq         : closed forms for integer df (== pgamma upper tail), fit_model.nf:138
 *   - set.seed/sample: MT19937 + R initial scrambling + rejection sampling, fit_model.nf:97-98
 *   - ReadHardcalls  : pgen 2-bit decode, fit_model.nf:123
 */
#include "flexlmm_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* R RNG (src/main/RNG.c): Mersenne-Twister, Inversion, Rejection                               */
/* ------------------------------------------------------------------------------------------ */
void orc_set_seed(orc_rng* g, int32_t seed_in) {
    uint32_t seed = (uint32_t)seed_in;
    for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;          /* initial scrambling */
    uint32_t i_seed[625];
    for (int j = 0; j < 625; j++) { seed = 69069u * seed + 1u; i_seed[j] = seed; }
    /* FixupSeeds: dummy[0] = mti = N (624) => the table is regenerated on first draw */
    for (int j = 0; j < 624; j++) g->mt[j] = i_seed[j + 1];
    g->mti = 624;
}

static uint32_t mt_next(orc_rng* g) {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
    uint32_t y;
    if (g->mti >= 624) {
        int kk;
        uint32_t* mt = g->mt;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
        }
        for (; kk < 623; kk++) {
            y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
        }
        y = (mt[623] & UPPER) | (mt[0] & LOWER);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
        g->mti = 0;
    }
    y = g->mt[g->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

double orc_unif_rand(orc_rng* g) {
    const double i2_32m1 = 2.328306437080797e-10; /* 1/(2^32 - 1) */
    double v = (double)mt_next(g) * 2.3283064365386963e-10; /* [0,1) */
    if (v <= 0.0) return 0.5 * i2_32m1;                       /* fixup() */
    if ((1.0 - v) <= 0.0) return 1.0 - 0.5 * i2_32m1;
    return v;
}

static double r_rbits(orc_rng* g, int bits) {
    int64_t v = 0;
    for (int n = 0; n <= bits; n += 16) {
        int v1 = (int)floor(orc_unif_rand(g) * 65536);
        v = 65536 * v + v1;
    }
    if (bits < 64) v &= (((int64_t)1 << bits) - 1);
    return (double)v;
}

static double r_unif_index(orc_rng* g, double dn) {
    if (dn <= 0) return 0.0;
    int bits = (int)ceil(log2(dn));
    double dv;
    do { dv = r_rbits(g, bits); } while (dn <= dv);
    return dv;
}

void orc_sample_perm(int32_t seed, int64_t m, int32_t* out) {
    orc_rng g;
    orc_set_seed(&g, seed);
    int32_t* x = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1));
    int64_t n = m;
    for (int64_t i = 0; i < m; i++) x[i] = (int32_t)i;
    for (int64_t i = 0; i < m; i++) {
        int64_t j = (int64_t)r_unif_index(&g, (double)n);
        out[i] = x[j] + 1;
        x[j] = x[--n];
    }
    free(x);
}

/* ------------------------------------------------------------------------------------------ */
/* pgen (plink-ng pgenlib format as consumed by pgenlibr::ReadHardcalls)                        */
/* ------------------------------------------------------------------------------------------ */
static uint32_t rd_u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint64_t rd_le(const uint8_t* p, int nbytes) { uint64_t v = 0; for (int i = 0; i < nbytes; i++) v |= (uint64_t)p[i] << (8 * i); return v; }

int orc_pgen_info(const uint8_t* f, int64_t len, uint32_t* variant_ct, uint32_t* sample_ct, int* mode) {
    if (len < 12 || f[0] != 0x6c || f[1] != 0x1b) return -1;
    *mode = f[2];
    if (f[2] != 0x02 && f[2] != 0x10) return -3;
    *variant_ct = rd_u32(f + 3);
    *sample_ct = rd_u32(f + 7);
    return 0;
}

typedef struct { uint8_t vrtype; uint64_t off; uint64_t len; } vrec;

/* locate record `vidx` of a mode-0x10 file, and (through *ldbase) the nearest preceding non-LD record */
static int locate10(const uint8_t* f, int64_t flen, uint32_t variant_ct, uint32_t vidx, vrec* rec, vrec* ldbase) {
    uint8_t ctrl = f[11];
    int enc = ctrl & 15;
    if (enc > 7) return -4;
    int vrtype_bits = (enc < 4) ? 4 : 8;
    int len_bytes = (enc & 3) + 1;
    int allele_bytes = (ctrl >> 4) & 3;
    int nonref_mode = (ctrl >> 6) & 3;
    uint32_t nblocks = (variant_ct + 65535u) / 65536u;
    const uint8_t* p = f + 12 + 8ull * nblocks;
    uint32_t blk = vidx >> 16;
    /* walk the per-block header sections up to the block of interest */
    for (uint32_t b = 0; b < blk; b++) {
        uint32_t cnt = 65536u;
        p += (vrtype_bits == 4) ? (cnt + 1) / 2 : cnt;
        p += (uint64_t)cnt * len_bytes + (uint64_t)cnt * allele_bytes;
        if (nonref_mode == 3) p += (cnt + 7) / 8;
    }
    uint32_t cnt = variant_ct - blk * 65536u; if (cnt > 65536u) cnt = 65536u;
    const uint8_t* vt = p;
    const uint8_t* vl = p + ((vrtype_bits == 4) ? (cnt + 1) / 2 : cnt);
    uint64_t off = rd_le(f + 12 + 8ull * blk, 8);
    uint32_t in_blk = vidx & 65535u;
    int have_base = 0;
    for (uint32_t i = 0; i <= in_blk; i++) {
        uint8_t t = (vrtype_bits == 4) ? ((vt[i >> 1] >> (4 * (i & 1))) & 15) : vt[i];
        uint64_t l = rd_le(vl + (uint64_t)i * len_bytes, len_bytes);
        if ((int64_t)(off + l) > flen) return -5;
        if ((t & 6) != 2) { ldbase->vrtype = t; ldbase->off = off; ldbase->len = l; have_base = 1; }
        if (i == in_blk) { rec->vrtype = t; rec->off = off; rec->len = l; }
        off += l;
    }
    if (((rec->vrtype & 6) == 2) && !have_base) return -6; /* LD record with no base in its block */
    return 0;
}

static int rd_varint(const uint8_t** pp, const uint8_t* end, uint32_t* out) {
    uint32_t v = 0; int shift = 0;
    while (*pp < end) {
        uint8_t b = *(*pp)++;
        v |= (uint32_t)(b & 127) << shift;
        if (!(b & 128)) { *out = v; return 0; }
        shift += 7;
        if (shift > 28) return -1;
    }
    return -1;
}

/* apply a difflist at *pp onto codes[] */
static int apply_difflist(const uint8_t** pp, const uint8_t* end, uint32_t sample_ct, uint8_t* codes) {
    uint32_t dl;
    if (rd_varint(pp, end, &dl)) return -7;
    if (dl == 0) return 0;
    if (dl > sample_ct) return -7;
    uint32_t group_ct = (dl + 63) / 64;
    int idb = (sample_ct < 256u) ? 1 : (sample_ct < 65536u) ? 2 : (sample_ct < 16777216u) ? 3 : 4;   /* pgenlib BytesToRepresentNzU32(raw_sample_ct) = bsr / 8 + 1 */
    const uint8_t* group_first = *pp;
    const uint8_t* p = group_first + (uint64_t)group_ct * idb;
    p += group_ct - 1;                       /* per-group byte counts (only needed for random access) */
    const uint8_t* raregeno = p;
    p += (dl + 3) / 4;
    if (p > end) return -7;
    uint32_t sid = 0;
    for (uint32_t e = 0; e < dl; e++) {
        if ((e & 63) == 0) sid = (uint32_t)rd_le(group_first + (uint64_t)(e >> 6) * idb, idb);
        else { uint32_t d; if (rd_varint(&p, end, &d)) return -7; sid += d; }
        if (sid >= sample_ct) return -7;
        codes[sid] = (raregeno[e >> 2] >> (2 * (e & 3))) & 3;
    }
    *pp = p;
    return 0;
}

/* main genotype track of a non-LD record; *track_end (optional) = first byte after it (where the next track starts) */
static int decode_record(const uint8_t* f, const vrec* r, uint32_t sample_ct, uint8_t* codes, const uint8_t** track_end) {
    const uint8_t* p = f + r->off;
    const uint8_t* end = p + r->len;
    int t = r->vrtype & 7;
    int rc = -9; /* LD types handled by caller */
    if (t == 0) {
        if (r->len < (sample_ct + 3) / 4) return -8;
        for (uint32_t i = 0; i < sample_ct; i++) codes[i] = (p[i >> 2] >> (2 * (i & 3))) & 3;
        p += (sample_ct + 3) / 4;
        rc = 0;
    } else if (t == 1) {
        uint8_t c2 = *p++;
        uint8_t lo = c2 >> 2, delta = c2 & 3;
        if (p + (sample_ct + 7) / 8 > end) return -8;
        for (uint32_t i = 0; i < sample_ct; i++) codes[i] = lo + (((p[i >> 3] >> (i & 7)) & 1) ? delta : 0);
        p += (sample_ct + 7) / 8;
        rc = apply_difflist(&p, end, sample_ct, codes);
    } else if (t >= 4) {
        memset(codes, t & 3, sample_ct);
        rc = apply_difflist(&p, end, sample_ct, codes);
    }
    if (track_end) *track_end = p;
    return rc;
}

static int read_hardcalls10(const uint8_t* f, int64_t len, uint32_t vct, uint32_t sct, uint32_t vidx, uint8_t* codes, vrec* rec_out, const uint8_t** track_end) {
    vrec rec = {0, 0, 0}, base = {0, 0, 0};
    int rc = locate10(f, len, vct, vidx, &rec, &base);
    if (rc) return rc;
    if (rec_out) *rec_out = rec;
    if (rec.vrtype & 0x08) return -10; /* multiallelic track: pipeline guarantees biallelic (geno_to_pgen.nf:32-42) */
    if ((rec.vrtype & 6) == 2) {
        rc = decode_record(f, &base, sct, codes, NULL);
        if (rc) return rc;
        const uint8_t* p = f + rec.off;
        rc = apply_difflist(&p, p + rec.len, sct, codes);
        if (rc) return rc;
        if ((rec.vrtype & 7) == 3)
            for (uint32_t i = 0; i < sct; i++) if (codes[i] != 1 && codes[i] != 3) codes[i] = 2 - codes[i];
        if (track_end) *track_end = p;
        return 0;
    }
    return decode_record(f, &rec, sct, codes, track_end);
}

int orc_pgen_read_hardcalls(const uint8_t* f, int64_t len, uint32_t vidx, uint8_t* codes) {
    uint32_t vct, sct; int mode;
    int rc = orc_pgen_info(f, len, &vct, &sct, &mode);
    if (rc) return rc;
    if (vidx >= vct) return -2;
    if (mode == 0x02) {
        uint64_t bpv = (sct + 3) / 4;
        if ((int64_t)(12 + bpv * ((uint64_t)vidx + 1)) > len) return -5;
        const uint8_t* p = f + 12 + bpv * vidx;
        for (uint32_t i = 0; i < sct; i++) codes[i] = (p[i >> 2] >> (2 * (i & 3))) & 3;
        return 0;
    }
    return read_hardcalls10(f, len, vct, sct, vidx, codes, NULL, NULL);
}

/* pgenlibr::Read (fit_model.nf:26 with use_dosage = true, :123): the hard calls of ReadHardcalls, overlaid with dosage / 16384
 * for every sample that has an entry in the record's dosage track (vrtype bits 5-6: 01 = sample-id delta list, 11 = bitarray over
 * all samples, 10 = a value for every sample with 65535 = none; values uint16, 0..32768 = ALT dosage 0..2; pgenlib PgrGetD +
 * Dosage16ToDoubles).  buf[i] = NA for a missing hard call without a dosage.  Phased records (bit 4) and phased dosages (bit 7)
 * are not produced by the pipeline's conversion (geno_to_pgen.nf:30-42) and are refused. */
int orc_pgen_read_dosage(const uint8_t* f, int64_t len, uint32_t vidx, double* buf) {
    uint32_t vct, sct; int mode;
    int rc = orc_pgen_info(f, len, &vct, &sct, &mode);
    if (rc) return rc;
    if (vidx >= vct) return -2;
    uint8_t* codes = (uint8_t*)malloc(sct ? sct : 1);
    if (mode == 0x02) {
        rc = orc_pgen_read_hardcalls(f, len, vidx, codes);
        if (!rc) orc_codes_to_buf(codes, sct, buf);
        free(codes);
        return rc;
    }
    vrec rec = {0, 0, 0};
    const uint8_t* p = NULL;
    rc = read_hardcalls10(f, len, vct, sct, vidx, codes, &rec, &p);
    if (rc) { free(codes); return rc; }
    orc_codes_to_buf(codes, sct, buf);
    free(codes);
    int dmode = (rec.vrtype >> 5) & 3;
    if (dmode == 0) return 0;
    if (rec.vrtype & 0x90) return -11;
    const uint8_t* end = f + rec.off + rec.len;
    const double scale = 0.00006103515625;   /* 1 / 16384 */
    if (dmode == 2) {
        if (p + 2ull * sct > end) return -12;
        for (uint32_t i = 0; i < sct; i++) { uint32_t d = p[2 * i] | ((uint32_t)p[2 * i + 1] << 8); if (d != 65535u) buf[i] = d * scale; }
        return 0;
    }
    if (dmode == 3) {
        const uint8_t* bits = p;
        const uint8_t* vals = p + (sct + 7) / 8;
        if (vals > end) return -12;
        for (uint32_t i = 0; i < sct; i++)
            if ((bits[i >> 3] >> (i & 7)) & 1) {
                if (vals + 2 > end) return -12;
                buf[i] = (vals[0] | ((uint32_t)vals[1] << 8)) * scale;
                vals += 2;
            }
        return 0;
    }
    /* dmode == 1: delta list of sample ids (a difflist without genotypes), then one uint16 per entry */
    uint32_t dl;
    if (rd_varint(&p, end, &dl)) return -12;
    if (dl == 0) return 0;
    if (dl > sct) return -12;
    uint32_t group_ct = (dl + 63) / 64;
    int idb = (sct < 256u) ? 1 : (sct < 65536u) ? 2 : (sct < 16777216u) ? 3 : 4;
    const uint8_t* group_first = p;
    p += (uint64_t)group_ct * idb + (group_ct - 1);
    if (p > end) return -12;
    uint32_t* ids = (uint32_t*)malloc(sizeof(uint32_t) * dl);
    uint32_t sid = 0;
    for (uint32_t e = 0; e < dl; e++) {
        if ((e & 63) == 0) sid = (uint32_t)rd_le(group_first + (uint64_t)(e >> 6) * idb, idb);
        else { uint32_t d; if (rd_varint(&p, end, &d)) { free(ids); return -12; } sid += d; }
        if (sid >= sct) { free(ids); return -12; }
        ids[e] = sid;
    }
    if (p + 2ull * dl > end) { free(ids); return -12; }
    for (uint32_t e = 0; e < dl; e++) buf[ids[e]] = (p[2 * e] | ((uint32_t)p[2 * e + 1] << 8)) * scale;
    free(ids);
    return 0;
}

void orc_codes_to_buf(const uint8_t* codes, int64_t n, double* buf) {
    /* R's NA_real_: quiet NaN with low word 1954 */
    union { double d; uint64_t u; } na; na.u = 0x7FF00000000007A2ull;
    for (int64_t i = 0; i < n; i++) buf[i] = (codes[i] == 3) ? na.d : (double)codes[i];
}

/* ------------------------------------------------------------------------------------------ */
/* forwardsolve: reference BLAS dtrsm, SIDE=L UPLO=L TRANS=N DIAG=N, alpha=1                    */
/* ------------------------------------------------------------------------------------------ */
void orc_forwardsolve(const double* L, int64_t n, int64_t ldl, double* X, int64_t k, int64_t ldx) {
    for (int64_t j = 0; j < k; j++) {
        double* b = X + j * ldx;
        for (int64_t kk = 0; kk < n; kk++) {
            if (b[kk] != 0.0) {
                const double* a = L + kk * ldl;
                b[kk] /= a[kk];
                double t = b[kk];
                for (int64_t i = kk + 1; i < n; i++) b[i] -= t * a[i];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* LINPACK dqrdc2 / dqrsl as modified for R (src/appl/dqrdc2.f, dqrsl.f, dqrls.f)               */
/* ------------------------------------------------------------------------------------------ */
static double dnrm2_(int64_t n, const double* x) {
    /* reference BLAS dnrm2: scaled sum of squares */
    if (n < 1) return 0.0;
    if (n == 1) return fabs(x[0]);
    double scale = 0.0, ssq = 1.0;
    for (int64_t i = 0; i < n; i++) {
        if (x[i] != 0.0) {
            double a = fabs(x[i]);
            if (scale < a) { ssq = 1.0 + ssq * (scale / a) * (scale / a); scale = a; }
            else ssq += (a / scale) * (a / scale);
        }
    }
    return scale * sqrt(ssq);
}
static double ddot_(int64_t n, const double* x, const double* y) { double s = 0.0; for (int64_t i = 0; i < n; i++) s += x[i] * y[i]; return s; }
static void daxpy_(int64_t n, double a, const double* x, double* y) { if (a == 0.0) return; for (int64_t i = 0; i < n; i++) y[i] += a * x[i]; }

static void dqrdc2(double* x, int64_t ldx, int64_t n, int p, double tol, int* k_out, double* qraux, int* jpvt, double* work /* 2p */) {
    double* w1 = work; double* w2 = work + p;
    for (int j = 0; j < p; j++) {
        qraux[j] = dnrm2_(n, x + j * ldx);
        w1[j] = qraux[j]; w2[j] = qraux[j];
        if (w2[j] == 0.0) w2[j] = 1.0;
    }
    int lup = (n < p) ? (int)n : p;
    int k = p + 1;                                  /* 1-based like the Fortran */
    for (int l = 1; l <= lup; l++) {
        /* cycle negligible columns to the end */
        while (l < k && qraux[l - 1] < w2[l - 1] * tol) {
            for (int64_t i = 0; i < n; i++) {
                double t = x[i + (l - 1) * ldx];
                for (int j = l + 1; j <= p; j++) x[i + (j - 2) * ldx] = x[i + (j - 1) * ldx];
                x[i + (p - 1) * ldx] = t;
            }
            int ii = jpvt[l - 1]; double t = qraux[l - 1], tt = w1[l - 1], ttt = w2[l - 1];
            for (int j = l + 1; j <= p; j++) {
                jpvt[j - 2] = jpvt[j - 1]; qraux[j - 2] = qraux[j - 1]; w1[j - 2] = w1[j - 1]; w2[j - 2] = w2[j - 1];
            }
            jpvt[p - 1] = ii; qraux[p - 1] = t; w1[p - 1] = tt; w2[p - 1] = ttt;
            k--;
        }
        if (l == n) break;
        double* xl = x + (l - 1) + (l - 1) * ldx;
        double nrmxl = dnrm2_(n - l + 1, xl);
        if (nrmxl == 0.0) continue;
        if (xl[0] != 0.0) nrmxl = copysign(nrmxl, xl[0]);
        { double s = 1.0 / nrmxl; for (int64_t i = 0; i < n - l + 1; i++) xl[i] *= s; }
        xl[0] = 1.0 + xl[0];
        for (int j = l + 1; j <= p; j++) {
            double* xj = x + (l - 1) + (j - 1) * ldx;
            double t = -ddot_(n - l + 1, xl, xj) / xl[0];
            daxpy_(n - l + 1, t, xl, xj);
            if (qraux[j - 1] != 0.0) {
                double r = fabs(xj[0]) / qraux[j - 1];
                double tt = 1.0 - r * r;
                if (tt < 0.0) tt = 0.0;
                t = tt;
                if (fabs(t) < 1e-6) {
                    qraux[j - 1] = dnrm2_(n - l, xj + 1);
                    w1[j - 1] = qraux[j - 1];
                } else {
                    qraux[j - 1] = qraux[j - 1] * sqrt(t);
                }
            }
        }
        qraux[l - 1] = xl[0];
        xl[0] = -nrmxl;
    }
    k = k - 1; if (k > n) k = (int)n;
    *k_out = k;
}

/* dqrsl with job = 1110 (qty, b, rsd) */
static void dqrsl_1110(double* x, int64_t ldx, int64_t n, int k, const double* qraux, const double* y,
                       double* qty, double* b, double* rsd) {
    int ju = (k < n - 1) ? k : (int)(n - 1);
    memcpy(qty, y, sizeof(double) * (size_t)n);
    for (int j = 1; j <= ju; j++) {
        if (qraux[j - 1] != 0.0) {
            double* xj = x + (j - 1) + (j - 1) * ldx;
            double temp = xj[0]; xj[0] = qraux[j - 1];
            double t = -ddot_(n - j + 1, xj, qty + j - 1) / xj[0];
            daxpy_(n - j + 1, t, xj, qty + j - 1);
            xj[0] = temp;
        }
    }
    for (int i = 0; i < k; i++) b[i] = qty[i];
    for (int64_t i = 0; i < n; i++) rsd[i] = (i < k) ? 0.0 : qty[i];
    for (int jj = 1; jj <= k; jj++) {
        int j = k - jj + 1;
        double d = x[(j - 1) + (j - 1) * ldx];
        if (d == 0.0) break; /* info = j: exactly singular */
        b[j - 1] /= d;
        if (j != 1) { double t = -b[j - 1]; daxpy_(j - 1, t, x + (j - 1) * ldx, b); }
    }
    for (int jj = 1; jj <= ju; jj++) {
        int j = ju - jj + 1;
        if (qraux[j - 1] != 0.0) {
            double* xj = x + (j - 1) + (j - 1) * ldx;
            double temp = xj[0]; xj[0] = qraux[j - 1];
            double t = -ddot_(n - j + 1, xj, rsd + j - 1) / xj[0];
            daxpy_(n - j + 1, t, xj, rsd + j - 1);
            xj[0] = temp;
        }
    }
}

int orc_lm_fit(double* x, int64_t n, int p, const double* y, double tol,
               double* coef, double* resid, double* effects, int* pivot, double* qraux) {
    double* work = (double*)malloc(sizeof(double) * 2 * (size_t)(p > 0 ? p : 1));
    double* qty = effects ? effects : (double*)malloc(sizeof(double) * (size_t)n);
    int k = 0;
    for (int j = 0; j < p; j++) pivot[j] = j + 1;
    dqrdc2(x, n, n, p, tol, &k, qraux, pivot, work);
    if (k > 0) {
        dqrsl_1110(x, n, n, k, qraux, y, qty, coef, resid);
    } else {
        memcpy(resid, y, sizeof(double) * (size_t)n);
        memcpy(qty, y, sizeof(double) * (size_t)n);
    }
    for (int j = k; j < p; j++) coef[j] = 0.0;
    if (!effects) free(qty);
    free(work);
    return k;
}

/* stats:::logLik.lm, no weights: 0.5*(0 - N*(log(2*pi) + 1 - log(N) + log(sum(res^2)))) ; R's sum() is long double */
double orc_loglik_lm(const double* resid, int64_t n) {
    long double s = 0.0L;
    for (int64_t i = 0; i < n; i++) s += (long double)(resid[i] * resid[i]);
    double ss = (double)s;
    double N = (double)n;
    return 0.5 * (0.0 - N * (log(2 * M_PI) + 1 - log(N) + log(ss)));
}

/* pchisq(q, df, lower.tail = FALSE) for integer df >= 1 */
double orc_pchisq_upper(double q, int df) {
    if (isnan(q)) return q;
    if (df <= 0) return NAN;
    if (q <= 0.0) return 1.0;
    double h = 0.5 * q;
    if ((df & 1) == 0) {
        int j = df / 2;
        double term = 1.0, sum = 1.0;
        for (int i = 1; i < j; i++) { term *= h / i; sum += term; }
        return exp(-h) * sum;
    } else {
        int j = (df - 1) / 2;
        double tail = erfc(sqrt(h));
        if (j == 0) return tail;
        double term = 1.0, sum = 1.0;            /* sum_{i<j} q^i / (1*3*...*(2i+1)) */
        for (int i = 1; i < j; i++) { term *= q / (2 * i + 1); sum += term; }
        return tail + exp(-h) * sqrt(2.0 * q / M_PI) * sum;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* One FIT_MODEL task (fit_model.nf:96-181)                                                     */
/* ------------------------------------------------------------------------------------------ */
int orc_fit_model(const orc_task* t, const uint8_t* geno, int64_t M, int32_t perm_seed,
                  double* chisq, int* df, double* pval, int* rank, int* pivot, double* coef,
                  double* min_pval_out, int64_t* nvars_tested_out) {
    const int64_t n = t->n; const int k = t->k;
    double* y = (double*)malloc(sizeof(double) * (size_t)n);
    double* X = (double*)malloc(sizeof(double) * (size_t)n * (size_t)(k > t->c_null ? k : t->c_null));
    double* res = (double*)malloc(sizeof(double) * (size_t)n);
    double* cf = (double*)malloc(sizeof(double) * (size_t)(k + t->c_null + 1));
    double* qraux = (double*)malloc(sizeof(double) * (size_t)(k + t->c_null + 1));
    int* pv = (int*)malloc(sizeof(int) * (size_t)(k + t->c_null + 1));
    double ll_null = t->ll_null; int df_null = t->df_null;
    memcpy(y, t->y_mm, sizeof(double) * (size_t)n);
    int rc = 0;

    if (perm_seed > 0) {
        /* :97-98  set.seed(seed); y.mm <- y.mm.pred + drop(U1 %*% sample(e.p)) */
        int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)t->m);
        orc_sample_perm(perm_seed, t->m, idx);
        for (int64_t i = 0; i < n; i++) y[i] = 0.0;
        for (int64_t j = 0; j < t->m; j++) {       /* reference dgemv: column axpy order */
            double e = t->e_p[idx[j] - 1];
            const double* u = t->U1 + j * n;
            for (int64_t i = 0; i < n; i++) y[i] += e * u[i];
        }
        for (int64_t i = 0; i < n; i++) y[i] = t->y_pred[i] + y[i];
        free(idx);
        /* :100-101 null refit */
        memcpy(X, t->X_null, sizeof(double) * (size_t)n * (size_t)t->c_null);
        int r0 = orc_lm_fit(X, n, t->c_null, y, 1e-7, cf, res, NULL, pv, qraux);
        ll_null = orc_loglik_lm(res, n);
        df_null = r0 + 1;
    }

    double min_pval = 2.0; int64_t nvars_tested = 0;
    for (int64_t v = 0; v < M; v++) {
        const uint8_t* g = geno + v * n;
        /* :124-127 design rows by genotype lookup */
        for (int j = 0; j < k; j++) {
            const double* m0 = t->M0 + (int64_t)j * n; const double* m1 = t->M1 + (int64_t)j * n; const double* m2 = t->M2 + (int64_t)j * n;
            double* xj = X + (int64_t)j * n;
            for (int64_t i = 0; i < n; i++) {
                uint8_t c = g[i];
                if (c > 2) { rc = -2; goto done; }       /* NA genotype: reference aborts */
                xj[i] = (c == 0) ? m0[i] : (c == 1) ? m1[i] : m2[i];
            }
        }
        orc_forwardsolve(t->L, n, n, X, k, n);           /* :128 */
        int r1 = orc_lm_fit(X, n, k, y, 1e-7, cf, res, NULL, pv, qraux);   /* :133 */
        double ll = orc_loglik_lm(res, n);               /* :134 */
        int lrt_df = (r1 + 1) - df_null;                 /* :135 */
        double lrt_chisq = 2.0 * (ll - ll_null);         /* :136 */
        double p = (lrt_df > 0) ? orc_pchisq_upper(lrt_chisq, lrt_df) : NAN;   /* :137-139 */
        nvars_tested++;
        if (!isnan(p) && p < min_pval) min_pval = p;     /* :143 min(..., na.rm = TRUE) */
        if (chisq) chisq[v] = lrt_chisq;
        if (df) df[v] = lrt_df;
        if (pval) pval[v] = p;
        if (rank) rank[v] = r1;
        if (pivot) for (int j = 0; j < k; j++) pivot[v * k + j] = pv[j];
        if (coef) for (int j = 0; j < k; j++) coef[v * k + j] = cf[j];
    }
    if (min_pval_out) *min_pval_out = min_pval;
    if (nvars_tested_out) *nvars_tested_out = nvars_tested;
done:
    free(y); free(X); free(res); free(cf); free(qraux); free(pv);
    return rc;
}
