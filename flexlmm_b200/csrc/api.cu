// api.cu — the C ABI (include/flexlmm_cuda.h): per-task state and the batch pipeline
//   K1 decode -> K2 whiten -> K3a gram -> K3b solve -> K4 permutation contraction (+ min-p finalise)
// over SNP batches of (SM count) column tiles, on one CUDA stream with double-buffered H2D staging.
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/flexlmm_cuda.h"
#include "common.cuh"
#include "host_math.h"
#include "kernels.h"

namespace flx {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace flx

using namespace flx;

#define FLX_CUBLAS_TRY(expr)                                                             \
    do {                                                                                 \
        cublasStatus_t _s = (expr);                                                      \
        if (_s != CUBLAS_STATUS_SUCCESS) {                                               \
            set_error("%s:%d: %s failed: cublas status %d", __FILE__, __LINE__, #expr, (int)_s); \
            return FLX_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)
#define FLX_TRY(expr)              \
    do {                           \
        int _r = (expr);           \
        if (_r != 0) return _r;    \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    void swap(DevBuf& o) { std::swap(p, o.p); std::swap(cap, o.cap); }
    int ensure(size_t count) {
        if (count <= cap && p) return 0;
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; set_error("cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e)); return FLX_ERR_OOM; }
        cap = count;
        return 0;
    }
};
template <typename T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    int ensure(size_t count) {
        if (count <= cap && p) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        if (count == 0) count = 1;
        cudaError_t e = cudaMallocHost(&p, count * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; set_error("cudaMallocHost(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e)); return FLX_ERR_OOM; }
        cap = count;
        return 0;
    }
};

struct flx_handle {
    flx_config cfg{};
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t d2h_stream = nullptr;   // per-batch result copies into pinned staging, overlapped with the next batch
    cudaEvent_t ev_res = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    std::vector<cudaEvent_t> ev_pool;    // stage stopwatch events, created once and reused by every run
    cublasHandle_t cublas = nullptr;
    ncclComm_t comm = nullptr;
    int nranks = 1, comm_rank = 0;

    // whitening
    int64_t n = 0, n_pad = 0;
    int KC = 0, T = 0;
    DevBuf<double> linv;     // triangular packed panels
    DevBuf<double> Zdense;   // dense L^-1 (kept until finalize when the int8 path may be used)
    DevBuf<double> Lstage;   // flx_whitening_begin / _cols / _end: the factor being assembled on the device
    int64_t Lstage_n = 0;
    bool have_Z = false;
    // null model
    bool have_null = false;
    std::vector<double> X_null, y_mm;
    int c_null = 0, df_null = 0;
    double ll_null = 0;
    // design
    bool have_design = false;
    int k = 0;
    std::vector<double> M0, M1, M2;
    // pgenlibr::Read (use_dosage): how every SNP column continues between the hard-call values, from two more probes of model.matrix
    bool use_dosage = false, have_probes = false, dos_ok = false, run_dosage = false;
    std::vector<double> M05, M15;
    uint32_t dos_indicator = 0;
    int dos_g[8]{};
    std::string dos_why;
    DevBuf<uint16_t> dos_d[2];
    DevBuf<uint32_t> main_end_d[2];
    // permutations (raw inputs kept until finalisation)
    bool have_perm = false;
    bool perm_implicit = false;   // U1 is the implicit Householder complement of the null design
    NullFit nullfit;
    std::vector<double> y_pred, e_p;
    DevBuf<double> U1_d;              // explicit U1 (n x m), device copy made by flx_set_perm, freed after finalisation
    int64_t m = 0;
    std::vector<int32_t> seeds;
    int P = 0;

    // derived
    bool dirty = true;
    int s = 0, ncf = 0, cr = 0, cn = 0;
    FitConst fc{};
    DevBuf<double> Qc, Qn, r, lut;
    double uni[8][3]{};
    uint32_t uniform_mask = 0;
    int64_t P_pad = 0;
    int NP = 0;
    DevBuf<double> Rpk, inv_rss, lrs0p, maxratio, minp_d;
    // int8 quadratic-form path (K2q)
    bool fast = false;
    int Q = 6, T256 = 0, NKB = 0;
    // every SNP term is g_t(x) * c_t[sample] with an integer triple g_t (|g| <= 2) and a per-sample scale c_t.  A 0/1 scale (the
    // indicator column of an x:factor interaction) is folded into the genotype code tile as a sample mask, so such terms share the
    // digit planes of the unscaled matrix: a tile set ("alpha") = (triple, mask); a real-valued scale is a "group" with its own planes
    struct QfPass { int mat, amma, ndot, adot[2], slot[2]; };
    static constexpr int MAX_NA = 4;
    struct FastPlan {
        int NG = 0, NA = 0, nslots = 0, nmask = 0;
        int grp[8]{}, alp[8]{};
        int aval[MAX_NA][3]{};
        int amask[MAX_NA]{};            // sample mask of the tile set (index into mask_d), -1: none
        bool ones[2]{};                 // group scale is identically 1
        std::vector<QfPass> passes;     // one K2q launch each
        int8_t slot_a[8][8]{}, slot_b[8][8]{};
    } plan;
    DevBuf<uint32_t> X2;                           // 2-bit code tiles of the batch, one set per tile set of the plan (K1q -> K2q / K4q)
    DevBuf<int8_t> T8[4], G8[2];                   // T8[a * NG + b]: digit planes of tril(C_a V^-1 C_b) / of C_a V^-1 C_b; G8[a]: of C_a L^-T [Qc | r | R_f]
    DevBuf<double> rowscale[4], qf_part, G_scale[2], cvec_d, ub, Hd;
    int NB_G = 16, ntilesG = 1;                    // column blocks of G: ntilesG x NB_G >= cr + 1 + P
    DevBuf<int32_t> refit_flag, refit_list, out_map_d;
    DevBuf<uint8_t> mask_d;                        // [nmask][n_pad256]: 0xFF keep / 0x00 zero, per model sample (tile-set masks of the int8 plan)
    DevBuf<int> df_count, missing;
    DevBuf<long long> coll_d;                      // [nvars, status] of the end-of-run collective
    int fallback_reason = 0;                       // FLX_FALLBACK_*: why `fast` is false
    int tiles_auto[2] = {0, 0};                    // cached default batch size per route (one cudaMemGetInfo per setup, not per run)

    // genotype source
    int src = 0;  // 0 none, 1 pgen file, 2 host packed
    PgenFile* pgen = nullptr;
    const uint8_t* packed_host = nullptr;
    bool host_registered = false;
    bool host_pinned_by_owner = false;   // a flx_multi pinned packed_host once for all of its handles
    int64_t src_variants = 0, bpv = 0, raw_n = 0;
    std::vector<int32_t> order0;
    bool identity_order = true;
    DevBuf<int32_t> order_d;
    DevBuf<uint8_t> resident;
    bool is_resident = false;

    // batch buffers
    DevBuf<uint8_t> packed_d[2];
    PinBuf<uint8_t> staging[2];
    DevBuf<uint8_t> blob_d[2], px_scratch[2];   // K0: compressed .pgen records of a batch + descriptors; rows of out-of-batch LD bases
    DevBuf<int> px_err;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr};
    cudaEvent_t ev_used[2] = {nullptr, nullptr};
    DevBuf<int32_t> vmap_d;
    DevBuf<double> X, W, n_part, u_part, a_part, b_part, snp_sm;
    DevBuf<int32_t> snp_df;
    // run-wide results
    DevBuf<double> chisq, pval, coef;
    DevBuf<int32_t> df, rank, pivot;
    PinBuf<double> chisq_h, pval_h, coef_h;       // pinned staging of the per-SNP results
    PinBuf<int32_t> df_h, rank_h, pivot_h;

    flx_stats stats{};

    ~flx_handle() {
        if (host_registered && packed_host) cudaHostUnregister(const_cast<uint8_t*>(packed_host));
        if (pgen) delete pgen;
        for (int i = 0; i < 2; i++) {
            if (ev_h2d[i]) cudaEventDestroy(ev_h2d[i]);
            if (ev_used[i]) cudaEventDestroy(ev_used[i]);
        }
        if (comm) ncclCommDestroy(comm);
        if (cublas) cublasDestroy(cublas);
        if (stream) cudaStreamDestroy(stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (d2h_stream) cudaStreamDestroy(d2h_stream);
        if (ev_res) cudaEventDestroy(ev_res);
        if (ev_t0) cudaEventDestroy(ev_t0);
        if (ev_t1) cudaEventDestroy(ev_t1);
        for (auto e : ev_pool) cudaEventDestroy(e);
    }
};

extern "C" const char* flx_last_error(void) { return g_err; }
extern "C" const char* flx_version(void) { return "flexlmm_b200 0.1.0 (sm_100a)"; }
extern "C" int flx_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
    return c;
}

extern "C" int flx_create(const flx_config* cfg, flx_handle** out) {
    if (!out) { set_error("flx_create: out is NULL"); return FLX_ERR_BAD_ARG; }
    *out = nullptr;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        set_error("flx_create: no CUDA device available (this library has no CPU fallback)");
        return FLX_ERR_NO_DEVICE;
    }
    flx_handle* h = new flx_handle();
    if (cfg) h->cfg = *cfg;
    h->Q = (h->cfg.flags & FLX_FLAG_PLANES_5) ? 5 : 6;
    if (h->cfg.device < 0 || h->cfg.device >= cnt) { set_error("flx_create: device %d out of range (%d devices)", h->cfg.device, cnt); delete h; return FLX_ERR_BAD_ARG; }
    cudaError_t e = cudaSetDevice(h->cfg.device);
    if (e != cudaSuccess) { set_error("cudaSetDevice: %s", cudaGetErrorString(e)); delete h; return FLX_ERR_CUDA; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, h->cfg.device);
    if (prop.major != 10) {
        set_error("flx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", h->cfg.device, prop.major, prop.minor);
        delete h;
        return FLX_ERR_NO_DEVICE;
    }
    h->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_res, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&h->ev_t0) != cudaSuccess || cudaEventCreate(&h->ev_t1) != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        delete h;
        return FLX_ERR_CUDA;
    }
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->ev_used[i], cudaEventDisableTiming);
    }
    if (cublasCreate(&h->cublas) != CUBLAS_STATUS_SUCCESS) { set_error("cublasCreate failed"); delete h; return FLX_ERR_CUDA; }
    cublasSetStream(h->cublas, h->stream);
    *out = h;
    return FLX_OK;
}

extern "C" void flx_destroy(flx_handle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    delete h;
}

// ------------------------------------------------------------------------------------------------
// A = L dense on the device (n x n column-major, ld = n) -> L^-1 in place: LAPACK dtrtri's blocked lower-triangular sweep (last
// block column first) on plain library trmm / trsm calls — setup only, outside the hot loop.  (cusolverDnXtrtri refuses n = 10^5.)
static int invert_lower_in_place(flx_handle* h, double* A, int64_t n) {
    std::vector<double> dg(n);
    FLX_CUDA_TRY(cudaMemcpy2DAsync(dg.data(), sizeof(double), A, (n + 1) * sizeof(double), sizeof(double), n, cudaMemcpyDeviceToHost, h->stream));
    FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int64_t i = 0; i < n; i++)
        if (!(dg[i] != 0.0) || !std::isfinite(dg[i])) { set_error("flx_set_whitening: L is singular or not finite (diagonal entry %lld)", (long long)i + 1); return FLX_ERR_BAD_ARG; }
    const int64_t nb = 1024;
    const double one = 1.0, mone = -1.0;
    DevBuf<double> Tb;
    FLX_TRY(Tb.ensure((size_t)nb * nb));
    for (int64_t j = ((n - 1) / nb) * nb; j >= 0; j -= nb) {
        const int64_t jb = std::min(nb, n - j), rest = n - j - jb;
        if (rest > 0) {
            FLX_CUBLAS_TRY(cublasDtrmm_64(h->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, rest, jb, &one,
                                          A + (j + jb) + (j + jb) * n, n, A + (j + jb) + j * n, n, A + (j + jb) + j * n, n));
            FLX_CUBLAS_TRY(cublasDtrsm_64(h->cublas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, rest, jb, &mone,
                                          A + j + j * n, n, A + (j + jb) + j * n, n));
        }
        // diagonal block: solve against the identity, copy the lower triangle back
        FLX_TRY(launch_set_identity(Tb.p, jb, jb, h->stream));
        FLX_CUBLAS_TRY(cublasDtrsm_64(h->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, jb, jb, &one, A + j + j * n, n, Tb.p, jb));
        FLX_CUDA_TRY(cudaMemcpy2DAsync(A + j + j * n, n * sizeof(double), Tb.p, jb * sizeof(double), jb * sizeof(double), jb, cudaMemcpyDeviceToDevice, h->stream));
    }
    FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return FLX_OK;
}

// src: L (or L^-1) dense on the device, n x n column-major with leading dimension n.  The buffer is CONSUMED: it becomes
// h->Zdense and is inverted in place, so that the setup of n = 10^5 holds one dense matrix (80 GB), not two.
static int whitening_from_device(flx_handle* h, DevBuf<double>& src, int64_t n, int is_inverse) {
    if (h->n > 0 && n != h->n) {
        // null model, design and permutation inputs are vectors of the OLD n: they must be given again
        h->have_null = h->have_design = h->have_perm = false;
        h->P = 0;
    }
    h->n = n;
    h->n_pad = (n + TJ - 1) / TJ * TJ;
    h->KC = (int)(h->n_pad / KCH);
    h->T = (int)(h->n_pad / TJ);
    h->dirty = true;
    h->Zdense.release();
    h->Zdense.swap(src);
    if (!is_inverse) FLX_TRY(invert_lower_in_place(h, h->Zdense.p, n));
    FLX_TRY(launch_zero_upper(h->Zdense.p, n, n, h->stream));      // the strict upper triangle of the caller's matrix may carry anything
    h->have_Z = true;
    FLX_TRY(h->linv.ensure((size_t)tri_chunks_before(h->T) * CHUNK));
    FLX_TRY(launch_pack_tri(h->Zdense.p, n, n, h->T, h->linv.p, h->stream));
    FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return FLX_OK;
}

extern "C" int flx_whitening_begin(flx_handle* h, int64_t n) {
    if (!h || n <= 0) { set_error("flx_whitening_begin: bad arguments"); return FLX_ERR_BAD_ARG; }
    FLX_CUDA_TRY(cudaSetDevice(h->cfg.device));
    h->Zdense.release(); h->have_Z = false;          // the previous factor's dense inverse goes first: one n x n matrix at a time
    h->Lstage.release();
    FLX_TRY(h->Lstage.ensure((size_t)n * n));
    h->Lstage_n = n;
    return FLX_OK;
}

extern "C" int flx_whitening_cols(flx_handle* h, int64_t j0, int64_t ncols, const double* cols, int64_t ld) {
    if (!h || !cols || j0 < 0 || ncols <= 0) { set_error("flx_whitening_cols: bad arguments"); return FLX_ERR_BAD_ARG; }
    const int64_t n = h->Lstage_n;
    if (n <= 0 || !h->Lstage.p) { set_error("flx_whitening_cols: call flx_whitening_begin first"); return FLX_ERR_STATE; }
    if (j0 + ncols > n || ld < n) { set_error("flx_whitening_cols: columns %lld..%lld / ld %lld do not fit n = %lld", (long long)j0, (long long)(j0 + ncols), (long long)ld, (long long)n); return FLX_ERR_BAD_ARG; }
    FLX_CUDA_TRY(cudaSetDevice(h->cfg.device));
    // cudaMemcpyDefault: `cols` may be host memory (R's matrix) or device memory (a factor produced on the GPU)
    FLX_CUDA_TRY(cudaMemcpy2DAsync(h->Lstage.p + j0 * n, n * sizeof(double), cols, ld * sizeof(double), n * sizeof(double), ncols, cudaMemcpyDefault, h->stream));
    FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));      // the caller may reuse `cols` as soon as this returns
    return FLX_OK;
}

extern "C" int flx_whitening_end(flx_handle* h, int is_inverse) {
    if (!h) { set_error("flx_whitening_end: NULL handle"); return FLX_ERR_BAD_ARG; }
    if (h->Lstage_n <= 0 || !h->Lstage.p) { set_error("flx_whitening_end: call flx_whitening_begin first"); return FLX_ERR_STATE; }
    FLX_CUDA_TRY(cudaSetDevice(h->cfg.device));
    const int64_t n = h->Lstage_n;
    h->Lstage_n = 0;
    return whitening_from_device(h, h->Lstage, n, is_inverse);
}

extern "C" int flx_set_whitening(flx_handle* h, const double* L, int64_t n, int64_t ld, int is_inverse) {
    if (!h || !L || n <= 0 || ld < n) { set_error("flx_set_whitening: bad arguments"); return FLX_ERR_BAD_ARG; }
    FLX_TRY(flx_whitening_begin(h, n));
    FLX_TRY(flx_whitening_cols(h, 0, n, L, ld));
    return flx_whitening_end(h, is_inverse);
}

// min over the ranks of a status word (0 = ok, negative = error): nobody enters a large collective after a local failure elsewhere
static int comm_min_status(flx_handle* h, int rc) {
    if (!h->comm || h->nranks < 2) return rc;
    long long w = rc;
    if (h->coll_d.ensure(2) != 0 || cudaMemcpyAsync(h->coll_d.p, &w, sizeof(w), cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
        ncclAllReduce(h->coll_d.p, h->coll_d.p, 1, ncclInt64, ncclMin, h->comm, h->stream) != ncclSuccess ||
        cudaMemcpyAsync(&w, h->coll_d.p, sizeof(w), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess) {
        if (rc == FLX_OK) { set_error("status exchange over the communicator failed"); return FLX_ERR_NCCL; }
        return rc;
    }
    if (rc == FLX_OK && w != FLX_OK) { set_error("another rank of the communicator failed with status %lld", w); return (int)w; }
    return rc;
}

extern "C" int flx_set_whitening_bcast(flx_handle* h, const double* L, int64_t n, int64_t ld, int is_inverse, int32_t root) {
    if (!h || n <= 0) { set_error("flx_set_whitening_bcast: bad arguments"); return FLX_ERR_BAD_ARG; }
    if (!h->comm || h->nranks < 2) return flx_set_whitening(h, L, n, ld, is_inverse);
    if (root < 0 || root >= h->nranks) { set_error("flx_set_whitening_bcast: root %d outside the communicator (%d ranks)", root, h->nranks); return FLX_ERR_BAD_ARG; }
    FLX_CUDA_TRY(cudaSetDevice(h->cfg.device));
    int rc = FLX_OK;
    if (h->comm_rank == root) {
        // the one rank that holds the factor uploads it (once per box) and inverts it
        if (!L || ld < n) { set_error("flx_set_whitening_bcast: the root rank must pass L"); rc = FLX_ERR_BAD_ARG; }
        if (rc == FLX_OK) rc = flx_whitening_begin(h, n);
        if (rc == FLX_OK) rc = flx_whitening_cols(h, 0, n, L, ld);
        if (rc == FLX_OK && !is_inverse) rc = invert_lower_in_place(h, h->Lstage.p, n);
    } else {
        h->Zdense.release(); h->have_Z = false;
        h->Lstage.release();
        rc = h->Lstage.ensure((size_t)n * n);
    }
    rc = comm_min_status(h, rc);
    if (rc != FLX_OK) { h->Lstage.release(); h->Lstage_n = 0; return rc; }
    // L^-1 travels GPU to GPU over NVLink / NVSwitch: n^2 * 8 bytes once, instead of one host upload and one inversion per GPU
    ncclResult_t r = ncclBroadcast(h->Lstage.p, h->Lstage.p, (size_t)n * (size_t)n, ncclDouble, root, h->comm, h->stream);
    if (r != ncclSuccess) { set_error("ncclBroadcast of L^-1 failed: %s", ncclGetErrorString(r)); return FLX_ERR_NCCL; }
    FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->Lstage_n = 0;
    return whitening_from_device(h, h->Lstage, n, 1);
}

#define FLX_CUSOLVER_TRY(expr)                                                             \
    do {                                                                                   \
        cusolverStatus_t _s = (expr);                                                      \
        if (_s != CUSOLVER_STATUS_SUCCESS) {                                               \
            set_error("%s:%d: %s failed: cusolver status %d", __FILE__, __LINE__, #expr, (int)_s); \
            return FLX_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

extern "C" int flx_decorrelate(flx_handle* h, const double* K, int64_t n, double s2g, double s2e, const double* y, const double* X, int32_t c,
                               double min_allowed_eig, double eig_replacement, double* L_out, double* y_mm_out, double* X_mm_out, int32_t* clipped_eigs) {
    if (!h || !K || n <= 0 || c < 0 || (c > 0 && !X)) { set_error("flx_decorrelate: bad arguments"); return FLX_ERR_BAD_ARG; }
    FLX_CUDA_TRY(cudaSetDevice(h->cfg.device));
    struct CsGuard { cusolverDnHandle_t h = nullptr; ~CsGuard() { if (h) cusolverDnDestroy(h); } } csg;    // released on every return path
    cusolverDnHandle_t& cs = csg.h;
    FLX_CUSOLVER_TRY(cusolverDnCreate(&cs));
    FLX_CUSOLVER_TRY(cusolverDnSetStream(cs, h->stream));
    DevBuf<double> V, work;
    DevBuf<int> info_d;
    FLX_TRY(V.ensure((size_t)n * n));
    FLX_TRY(info_d.ensure(1));
    auto build_V = [&]() -> int {
        FLX_CUDA_TRY(cudaMemcpyAsync(V.p, K, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        return launch_scale_add_diag(V.p, n, n, s2g, s2e, h->stream);       // V = s2g*K + s2e*I   (decorrelate.nf:54)
    };
    if (int rc = build_V()) return rc;
    int lwork = 0, info = 0, nclip = 0;
    FLX_CUSOLVER_TRY(cusolverDnDpotrf_bufferSize(cs, CUBLAS_FILL_MODE_LOWER, (int)n, V.p, (int)n, &lwork));
    FLX_TRY(work.ensure((size_t)lwork));
    FLX_CUSOLVER_TRY(cusolverDnDpotrf(cs, CUBLAS_FILL_MODE_LOWER, (int)n, V.p, (int)n, work.p, lwork, info_d.p));   // L = t(chol(V)) (:57)
    FLX_CUDA_TRY(cudaMemcpyAsync(&info, info_d.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (info != 0) {
        // not positive definite: force it by replacing the non-positive eigenvalues (:58-69)
        if (int rc = build_V()) return rc;
        DevBuf<double> lam, U, T2;
        FLX_TRY(lam.ensure(n)); FLX_TRY(U.ensure((size_t)n * n)); FLX_TRY(T2.ensure((size_t)n * n));
        FLX_CUSOLVER_TRY(cusolverDnDsyevd_bufferSize(cs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, V.p, (int)n, lam.p, &lwork));
        FLX_TRY(work.ensure((size_t)lwork));
        FLX_CUSOLVER_TRY(cusolverDnDsyevd(cs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, V.p, (int)n, lam.p, work.p, lwork, info_d.p));
        std::vector<double> lh(n);
        FLX_CUDA_TRY(cudaMemcpyAsync(lh.data(), lam.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        FLX_CUDA_TRY(cudaMemcpyAsync(&info, info_d.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
        if (info != 0) { set_error("flx_decorrelate: eigen-decomposition failed (info %d)", info); return FLX_ERR_CUDA; }
        for (int64_t i = 0; i < n; i++) {
            if (!(lh[i] >= min_allowed_eig)) {
                set_error("flx_decorrelate: eigenvalue %.6g below min_allowed_eig %.6g (decorrelate.nf:65 stopifnot)", lh[i], min_allowed_eig);
                return FLX_ERR_BAD_ARG;
            }
            if (lh[i] <= 0.0) { lh[i] = eig_replacement; nclip++; }
            lh[i] = std::fabs(lh[i]);
        }
        FLX_CUDA_TRY(cudaMemcpyAsync(lam.p, lh.data(), n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        // V <- U diag(|lambda|) U'
        FLX_CUDA_TRY(cudaMemcpyAsync(U.p, V.p, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        FLX_CUBLAS_TRY(cublasDdgmm_64(h->cublas, CUBLAS_SIDE_RIGHT, (int)n, (int)n, U.p, (int)n, lam.p, 1, T2.p, (int)n));
        const double one = 1.0, zero = 0.0;
        FLX_CUBLAS_TRY(cublasDgemm_64(h->cublas, CUBLAS_OP_N, CUBLAS_OP_T, (int)n, (int)n, (int)n, &one, T2.p, (int)n, U.p, (int)n, &zero, V.p, (int)n));
        FLX_CUSOLVER_TRY(cusolverDnDpotrf_bufferSize(cs, CUBLAS_FILL_MODE_LOWER, (int)n, V.p, (int)n, &lwork));
        FLX_TRY(work.ensure((size_t)lwork));
        FLX_CUSOLVER_TRY(cusolverDnDpotrf(cs, CUBLAS_FILL_MODE_LOWER, (int)n, V.p, (int)n, work.p, lwork, info_d.p));
        FLX_CUDA_TRY(cudaMemcpyAsync(&info, info_d.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
        if (info != 0) { set_error("flx_decorrelate: V is not positive definite even after eigenvalue replacement (info %d)", info); return FLX_ERR_BAD_ARG; }
    }
    cusolverDnDestroy(cs);
    cs = nullptr;                                       // the workspace-heavy handle goes before the factor is inverted
    if (clipped_eigs) *clipped_eigs = nclip;
    FLX_TRY(launch_zero_upper(V.p, n, n, h->stream));     // potrf leaves the input above the diagonal
    // y.mm, X.mm = forwardsolve(L, .)  (:71-72)
    const int nrhs = c + (y ? 1 : 0);
    if (nrhs > 0 && (y_mm_out || X_mm_out)) {
        DevBuf<double> B;
        FLX_TRY(B.ensure((size_t)n * nrhs));
        if (y) FLX_CUDA_TRY(cudaMemcpyAsync(B.p, y, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        if (c > 0) FLX_CUDA_TRY(cudaMemcpyAsync(B.p + (y ? n : 0), X, (size_t)n * c * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        const double one = 1.0;
        FLX_CUBLAS_TRY(cublasDtrsm_64(h->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, (int)n, nrhs, &one, V.p, (int)n, B.p, (int)n));
        if (y && y_mm_out) FLX_CUDA_TRY(cudaMemcpyAsync(y_mm_out, B.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (c > 0 && X_mm_out) FLX_CUDA_TRY(cudaMemcpyAsync(X_mm_out, B.p + (y ? n : 0), (size_t)n * c * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    if (L_out) FLX_CUDA_TRY(cudaMemcpyAsync(L_out, V.p, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return whitening_from_device(h, V, n, 0);
}

extern "C" int flx_set_null(flx_handle* h, const double* X_null, int32_t c, const double* y_mm, double ll_null, int32_t df_null) {
    if (!h || !y_mm || c < 0 || (c > 0 && !X_null) || c > FLX_MAX_K) { set_error("flx_set_null: bad arguments"); return FLX_ERR_BAD_ARG; }
    if (h->n <= 0) { set_error("flx_set_null: call flx_set_whitening first"); return FLX_ERR_STATE; }
    h->X_null.assign(X_null, X_null + (size_t)h->n * c);
    h->y_mm.assign(y_mm, y_mm + h->n);
    h->c_null = c; h->ll_null = ll_null; h->df_null = df_null;
    h->have_null = true; h->dirty = true;
    return FLX_OK;
}

extern "C" int flx_set_design(flx_handle* h, const double* M0, const double* M1, const double* M2, int32_t k) {
    if (!h || !M0 || !M1 || !M2 || k <= 0) { set_error("flx_set_design: bad arguments"); return FLX_ERR_BAD_ARG; }
    if (k > FLX_MAX_K) { set_error("flx_set_design: k=%d exceeds FLX_MAX_K=%d", k, FLX_MAX_K); return FLX_ERR_BAD_ARG; }
    if (h->n <= 0) { set_error("flx_set_design: call flx_set_whitening first"); return FLX_ERR_STATE; }
    const size_t sz = (size_t)h->n * k;
    h->M0.assign(M0, M0 + sz); h->M1.assign(M1, M1 + sz); h->M2.assign(M2, M2 + sz);
    h->k = k; h->have_design = true; h->dirty = true;
    h->have_probes = false;      // probes belong to the design they were evaluated with
    return FLX_OK;
}

extern "C" int flx_set_design_probes(flx_handle* h, const double* M05, const double* M15) {
    if (!h || !M05 || !M15) { set_error("flx_set_design_probes: bad arguments"); return FLX_ERR_BAD_ARG; }
    if (!h->have_design) { set_error("flx_set_design_probes: call flx_set_design first"); return FLX_ERR_STATE; }
    const size_t sz = (size_t)h->n * h->k;
    h->M05.assign(M05, M05 + sz); h->M15.assign(M15, M15 + sz);
    h->have_probes = true; h->dirty = true;
    return FLX_OK;
}

extern "C" int flx_set_use_dosage(flx_handle* h, int use_dosage) {
    if (!h) { set_error("flx_set_use_dosage: NULL handle"); return FLX_ERR_BAD_ARG; }
    h->use_dosage = use_dosage != 0;
    return FLX_OK;
}

extern "C" int flx_set_perm(flx_handle* h, const double* y_pred, const double* U1, const double* e_p, int64_t m, const int32_t* seeds, int32_t P) {
    if (!h) { set_error("flx_set_perm: NULL handle"); return FLX_ERR_BAD_ARG; }
    if (P == 0) { h->have_perm = false; h->P = 0; h->dirty = true; return FLX_OK; }
    if (!y_pred || !U1 || !e_p || !seeds || m <= 0 || P < 0) { set_error("flx_set_perm: bad arguments"); return FLX_ERR_BAD_ARG; }
    if (h->n <= 0) { set_error("flx_set_perm: call flx_set_whitening first"); return FLX_ERR_STATE; }
    h->y_pred.assign(y_pred, y_pred + h->n);
    h->e_p.assign(e_p, e_p + m);
    // R owns U1 only for the duration of the call: it goes straight to the device (no host copy of an n x m matrix)
    FLX_CUDA_TRY(cudaSetDevice(h->cfg.device));
    FLX_TRY(h->U1_d.ensure((size_t)h->n * m));
    FLX_CUDA_TRY(cudaMemcpy(h->U1_d.p, U1, (size_t)h->n * m * sizeof(double), cudaMemcpyHostToDevice));
    h->m = m;
    h->seeds.assign(seeds, seeds + P);
    h->P = P; h->have_perm = true; h->perm_implicit = false; h->dirty = true;
    return FLX_OK;
}

extern "C" int flx_set_perm_bcast(flx_handle* h, const double* y_pred, const double* U1, const double* e_p, int64_t m, const int32_t* seeds, int32_t P, int32_t root) {
    if (!h) { set_error("flx_set_perm_bcast: NULL handle"); return FLX_ERR_BAD_ARG; }
    if (!h->comm || h->nranks < 2 || P == 0) return flx_set_perm(h, y_pred, U1, e_p, m, seeds, P);
    if (root < 0 || root >= h->nranks) { set_error("flx_set_perm_bcast: root %d outside the communicator (%d ranks)", root, h->nranks); return FLX_ERR_BAD_ARG; }
    int rc = FLX_OK;
    if (!y_pred || !e_p || !seeds || m <= 0 || P < 0 || (h->comm_rank == root && !U1)) { set_error("flx_set_perm_bcast: bad arguments"); rc = FLX_ERR_BAD_ARG; }
    else if (h->n <= 0) { set_error("flx_set_perm_bcast: call flx_set_whitening first"); rc = FLX_ERR_STATE; }
    if (rc == FLX_OK && cudaSetDevice(h->cfg.device) != cudaSuccess) rc = FLX_ERR_CUDA;
    if (rc == FLX_OK) rc = h->U1_d.ensure((size_t)h->n * m);
    if (rc == FLX_OK && h->comm_rank == root && cudaMemcpy(h->U1_d.p, U1, (size_t)h->n * m * sizeof(double), cudaMemcpyDefault) != cudaSuccess) { set_error("flx_set_perm_bcast: upload of U1 failed"); rc = FLX_ERR_CUDA; }
    rc = comm_min_status(h, rc);
    if (rc != FLX_OK) return rc;
    // U1 (n x m, as large as L) crosses the host link once per box; the other GPUs receive it over NVLink
    ncclResult_t r = ncclBroadcast(h->U1_d.p, h->U1_d.p, (size_t)h->n * (size_t)m, ncclDouble, root, h->comm, h->stream);
    if (r != ncclSuccess) { set_error("ncclBroadcast of U1 failed: %s", ncclGetErrorString(r)); return FLX_ERR_NCCL; }
    FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->y_pred.assign(y_pred, y_pred + h->n);
    h->e_p.assign(e_p, e_p + m);
    h->m = m;
    h->seeds.assign(seeds, seeds + P);
    h->P = P; h->have_perm = true; h->perm_implicit = false; h->dirty = true;
    return FLX_OK;
}

extern "C" int flx_set_perm_implicit(flx_handle* h, const int32_t* seeds, int32_t P) {
    if (!h || P < 0 || (P > 0 && !seeds)) { set_error("flx_set_perm_implicit: bad arguments"); return FLX_ERR_BAD_ARG; }
    if (P == 0) { h->have_perm = false; h->P = 0; h->dirty = true; return FLX_OK; }
    if (!h->have_null) { set_error("flx_set_perm_implicit: call flx_set_null first"); return FLX_ERR_STATE; }
    null_fit_householder(h->X_null.data(), h->n, h->c_null, h->y_mm.data(), h->nullfit);
    h->y_pred = h->nullfit.fitted;
    h->e_p = h->nullfit.e_p;
    h->m = (int64_t)h->e_p.size();
    h->U1_d.release();
    h->seeds.assign(seeds, seeds + P);
    h->P = P; h->have_perm = true; h->perm_implicit = true; h->dirty = true;
    return FLX_OK;
}

extern "C" int flx_fit_null_model(const double* X_mm, int32_t c, const double* y_mm, int64_t n, double* ll, int32_t* df,
                                  double* y_pred, double* resid, double* e_p, double* U1, int64_t* m) {
    if (!y_mm || n <= 0 || c < 0 || (c > 0 && !X_mm)) { set_error("flx_fit_null_model: bad arguments"); return FLX_ERR_BAD_ARG; }
    NullFit nf;
    null_fit_householder(X_mm, n, c, y_mm, nf);
    const int64_t mm = n - nf.rank;
    if (ll) *ll = nf.ll;
    if (df) *df = nf.rank + 1;
    if (m) *m = mm;
    if (y_pred) memcpy(y_pred, nf.fitted.data(), n * sizeof(double));
    if (resid) memcpy(resid, nf.resid.data(), n * sizeof(double));
    if (e_p) memcpy(e_p, nf.e_p.data(), mm * sizeof(double));
    if (U1)
        for (int64_t j = 0; j < mm; j++) {
            double* col = U1 + j * n;
            for (int64_t i = 0; i < n; i++) col[i] = 0.0;
            col[nf.rank + j] = 1.0;
            nf.apply_Q(col);
        }
    return FLX_OK;
}

static int set_order(flx_handle* h, const int32_t* pgen_order, int64_t n, int64_t raw_n) {
    if (h->n > 0 && n != h->n) { set_error("genotype source: n=%lld does not match the whitening matrix (%lld)", (long long)n, (long long)h->n); return FLX_ERR_BAD_ARG; }
    h->order0.resize(n);
    h->identity_order = true;
    for (int64_t i = 0; i < n; i++) {
        const int64_t o = pgen_order ? (int64_t)pgen_order[i] - 1 : i;
        if (o < 0 || o >= raw_n) { set_error("pgen_order[%lld]=%lld out of range 1..%lld", (long long)i, (long long)o + 1, (long long)raw_n); return FLX_ERR_BAD_ARG; }
        h->order0[i] = (int32_t)o;
        if (o != i) h->identity_order = false;
    }
    FLX_CUDA_TRY(cudaSetDevice(h->cfg.device));
    FLX_TRY(h->order_d.ensure(n));
    FLX_CUDA_TRY(cudaMemcpy(h->order_d.p, h->order0.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
    h->is_resident = false;
    return FLX_OK;
}

extern "C" int flx_open_pgen(flx_handle* h, const char* path, const int32_t* pgen_order, int64_t n) {
    if (!h || !path || n <= 0) { set_error("flx_open_pgen: bad arguments"); return FLX_ERR_BAD_ARG; }
    if (h->pgen) { delete h->pgen; h->pgen = nullptr; }
    h->pgen = new PgenFile();
    std::string e = h->pgen->open(path);
    if (!e.empty()) { set_error("flx_open_pgen(%s): %s", path, e.c_str()); delete h->pgen; h->pgen = nullptr; return FLX_ERR_PGEN; }
    h->src = 1;
    h->src_variants = h->pgen->variant_ct();
    h->raw_n = h->pgen->sample_ct();
    h->bpv = h->pgen->bytes_per_variant();
    return set_order(h, pgen_order, n, h->raw_n);
}

extern "C" int flx_pgen_info(flx_handle* h, int64_t* variant_ct, int64_t* sample_ct, int32_t* has_dosage) {
    if (!h || h->src != 1 || !h->pgen) { set_error("flx_pgen_info: no .pgen is open"); return FLX_ERR_STATE; }
    if (variant_ct) *variant_ct = h->pgen->variant_ct();
    if (sample_ct) *sample_ct = h->pgen->sample_ct();
    if (has_dosage) *has_dosage = h->pgen->has_dosage() ? 1 : 0;
    return FLX_OK;
}

// pinned_by_owner: a flx_multi registered the buffer once (portable) for all of its handles
static int set_packed_impl(flx_handle* h, const uint8_t* packed, int64_t n_variants, int64_t bytes_per_variant, int64_t raw_sample_ct, const int32_t* pgen_order, int64_t n, bool pinned_by_owner) {
    if (!h || !packed || n_variants < 0 || raw_sample_ct <= 0 || bytes_per_variant < (raw_sample_ct + 3) / 4 || n <= 0) { set_error("flx_set_packed: bad arguments"); return FLX_ERR_BAD_ARG; }
    if (h->host_registered && h->packed_host) { cudaHostUnregister(const_cast<uint8_t*>(h->packed_host)); h->host_registered = false; }
    h->src = 2;
    h->packed_host = packed;
    h->host_pinned_by_owner = pinned_by_owner;
    // pin the caller's buffer in place so batches go H2D by DMA straight from it (no staging copy); harmless if it fails
    if (!pinned_by_owner && cudaSetDevice(h->cfg.device) == cudaSuccess && n_variants * bytes_per_variant >= (1 << 20)) {
        if (cudaHostRegister(const_cast<uint8_t*>(packed), (size_t)n_variants * bytes_per_variant, cudaHostRegisterReadOnly) == cudaSuccess ||
            (cudaGetLastError(), cudaHostRegister(const_cast<uint8_t*>(packed), (size_t)n_variants * bytes_per_variant, cudaHostRegisterDefault) == cudaSuccess))
            h->host_registered = true;
        else
            cudaGetLastError();
    }
    h->src_variants = n_variants;
    h->raw_n = raw_sample_ct;
    h->bpv = bytes_per_variant;
    return set_order(h, pgen_order, n, raw_sample_ct);
}

extern "C" int flx_set_packed(flx_handle* h, const uint8_t* packed, int64_t n_variants, int64_t bytes_per_variant, int64_t raw_sample_ct, const int32_t* pgen_order, int64_t n) {
    return set_packed_impl(h, packed, n_variants, bytes_per_variant, raw_sample_ct, pgen_order, n, false);
}

extern "C" int flx_make_resident(flx_handle* h) {
    if (!h || h->src != 2) { set_error("flx_make_resident: needs flx_set_packed first"); return FLX_ERR_STATE; }
    FLX_CUDA_TRY(cudaSetDevice(h->cfg.device));
    FLX_TRY(h->resident.ensure((size_t)h->src_variants * h->bpv));
    FLX_CUDA_TRY(cudaMemcpy(h->resident.p, h->packed_host, (size_t)h->src_variants * h->bpv, cudaMemcpyHostToDevice));
    h->is_resident = true;
    return FLX_OK;
}

// ------------------------------------------------------------------------------------------------
// whiten `ncols` dense columns (device, n x ncols, ld) -> dense (device). Used for the constant design columns.
static int whiten_dense(flx_handle* h, const double* Xd, int64_t ld, int64_t ncols, double* Wd, int64_t ldw) {
    const int NJ = (int)((ncols + TJ - 1) / TJ);
    DevBuf<double> Xp, Wp;
    FLX_TRY(Xp.ensure((size_t)NJ * h->KC * CHUNK));
    FLX_TRY(Wp.ensure((size_t)NJ * h->KC * CHUNK));
    FLX_TRY(launch_pack_cols(Xd, h->n, ld, ncols, h->KC, Xp.p, h->stream));
    GemmArgs g{};
    g.A = h->linv.p; g.B = Xp.p; g.Out = Wp.p;
    g.T = h->T; g.NJ = NJ; g.KC = h->KC; g.n_items = (int64_t)h->T * NJ;
    g.KC_live = (int)((h->n + KCH - 1) / KCH); g.last_row_blocks = (int)((h->n - (int64_t)(h->T - 1) * TJ + 7) / 8);
    FLX_TRY(launch_whiten(g, h->num_sms, h->stream));
    FLX_TRY(launch_unpack_cols(Wp.p, h->n, ldw, ncols, h->KC, Wd, h->stream));
    FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return FLX_OK;
}

extern "C" int flx_whiten(flx_handle* h, const double* X, int64_t ncols, double* W) {
    if (!h || !X || !W || ncols <= 0) { set_error("flx_whiten: bad arguments"); return FLX_ERR_BAD_ARG; }
    if (!h->linv.p) { set_error("flx_whiten: call flx_set_whitening first"); return FLX_ERR_STATE; }
    FLX_CUDA_TRY(cudaSetDevice(h->cfg.device));
    DevBuf<double> Xd, Wd;
    FLX_TRY(Xd.ensure((size_t)h->n * ncols));
    FLX_TRY(Wd.ensure((size_t)h->n * ncols));
    FLX_CUDA_TRY(cudaMemcpyAsync(Xd.p, X, (size_t)h->n * ncols * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    FLX_TRY(whiten_dense(h, Xd.p, h->n, ncols, Wd.p, h->n));
    FLX_CUDA_TRY(cudaMemcpy(W, Wd.p, (size_t)h->n * ncols * sizeof(double), cudaMemcpyDeviceToHost));
    return FLX_OK;
}

static void upload_rowmajor_padded(const std::vector<double>& Q, int64_t n, int r, int64_t n_pad, std::vector<double>& out) {
    const int rp = gram_padded_width(r);
    out.assign((size_t)n_pad * rp + 8, 0.0);
    for (int j = 0; j < r; j++)
        for (int64_t i = 0; i < n; i++) out[i * rp + j] = Q[i + (int64_t)j * n];
}

// derive everything that depends on (L, null, design, perm): once per task
static int finalize_setup(flx_handle* h) {
    if (!h->linv.p || !h->have_null || !h->have_design) { set_error("flx_run: whitening, null model and design must be set first"); return FLX_ERR_STATE; }
    FLX_CUDA_TRY(cudaSetDevice(h->cfg.device));
    const int64_t n = h->n, n_pad = h->n_pad;
    const int k = h->k;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (h->cfg.verbose < 1) return;
        cudaStreamSynchronize(h->stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[flx] setup: %-34s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    // ---- split the design into SNP-independent columns and SNP-dependent terms (fit_model.nf:124-127)
    std::vector<int> const_cols, snp_cols;
    for (int j = 0; j < k; j++) {
        bool same = true;
        const double *a = h->M0.data() + (size_t)j * n, *b = h->M1.data() + (size_t)j * n, *c = h->M2.data() + (size_t)j * n;
        for (int64_t i = 0; i < n && same; i++) same = (a[i] == b[i]) && (a[i] == c[i]);
        (same ? const_cols : snp_cols).push_back(j);
    }
    h->s = (int)snp_cols.size();
    h->ncf = (int)const_cols.size();
    if (h->s < 1) { set_error("flx_set_design: no design column depends on the genotype x"); return FLX_ERR_BAD_ARG; }
    if (h->s > FLX_MAX_S) { set_error("flx_set_design: %d SNP-dependent columns exceed FLX_MAX_S=%d", h->s, FLX_MAX_S); return FLX_ERR_BAD_ARG; }
    FitConst& fc = h->fc;
    memset(&fc, 0, sizeof(fc));
    for (int ci = 0; ci < h->ncf; ci++) fc.col_kind[const_cols[ci]] = ci;
    for (int t = 0; t < h->s; t++) fc.col_kind[snp_cols[t]] = -(t + 1);
    // ---- lookup tables of the SNP terms
    h->uniform_mask = 0;
    std::vector<double> lut((size_t)h->s * 3 * n_pad, 0.0);
    for (int t = 0; t < h->s; t++) {
        const double* src[3] = {h->M0.data() + (size_t)snp_cols[t] * n, h->M1.data() + (size_t)snp_cols[t] * n, h->M2.data() + (size_t)snp_cols[t] * n};
        bool uniform = true;
        for (int g = 0; g < 3; g++) {
            h->uni[t][g] = src[g][0];
            for (int64_t i = 0; i < n; i++) {
                lut[((size_t)t * 3 + g) * n_pad + i] = src[g][i];
                if (src[g][i] != src[g][0]) uniform = false;
            }
        }
        if (uniform) h->uniform_mask |= 1u << t;
    }
    FLX_TRY(h->lut.ensure(lut.size()));
    FLX_CUDA_TRY(cudaMemcpy(h->lut.p, lut.data(), lut.size() * sizeof(double), cudaMemcpyHostToDevice));
    // ---- dosages (pgenlibr::Read): a SNP column must be affine in x (x, x:cov: v(x) = v(0) + x (v(1) - v(0))) or the indicator of
    // one hard-call value (I(x == 1): v(x) = v(g) if x == g else the common value elsewhere); the probes at x = 0.5, 1.5 tell which
    h->dos_ok = false; h->dos_indicator = 0; h->dos_why.clear();
    if (h->have_probes) {
        bool all_ok = true;
        for (int t = 0; t < h->s && all_ok; t++) {
            const size_t o = (size_t)snp_cols[t] * n;
            const double *v0 = h->M0.data() + o, *v1 = h->M1.data() + o, *v2 = h->M2.data() + o, *h05 = h->M05.data() + o, *h15 = h->M15.data() + o;
            bool affine = true;
            for (int64_t i = 0; i < n && affine; i++) {
                const double sl = v1[i] - v0[i];
                affine = (v2[i] == v0[i] + 2.0 * sl) && (h05[i] == v0[i] + 0.5 * sl) && (h15[i] == v0[i] + 1.5 * sl);
            }
            if (affine) continue;
            int gfound = -1;
            for (int g = 0; g < 3 && gfound < 0; g++) {
                const double* vg[3] = {v0, v1, v2};
                const double *va = vg[(g + 1) % 3], *vb = vg[(g + 2) % 3];
                bool ind = true;
                for (int64_t i = 0; i < n && ind; i++) ind = (va[i] == vb[i]) && (h05[i] == va[i]) && (h15[i] == va[i]);
                if (ind) gfound = g;
            }
            if (gfound >= 0) { h->dos_indicator |= 1u << t; h->dos_g[t] = gfound; continue; }
            all_ok = false;
            char buf[200];
            snprintf(buf, sizeof(buf), "design column %d is neither affine in x nor the indicator of one genotype value: it cannot be evaluated on dosages", snp_cols[t] + 1);
            h->dos_why = buf;
        }
        h->dos_ok = all_ok;
    } else {
        h->dos_why = "use_dosage on a .pgen with dosage tracks needs flx_set_design_probes (model.matrix at x = 0.5 and 1.5)";
    }
    lap("design split, lookup tables");
    // ---- whiten the constant columns of the real model, orthonormal basis Qc
    std::vector<double> Cw((size_t)n * std::max(h->ncf, 1));
    if (h->ncf > 0) {
        std::vector<double> C((size_t)n * h->ncf);
        for (int ci = 0; ci < h->ncf; ci++) memcpy(C.data() + (size_t)ci * n, h->M0.data() + (size_t)const_cols[ci] * n, n * sizeof(double));
        DevBuf<double> Cd, Wd;
        FLX_TRY(Cd.ensure(C.size()));
        FLX_TRY(Wd.ensure(C.size()));
        FLX_CUDA_TRY(cudaMemcpy(Cd.p, C.data(), C.size() * sizeof(double), cudaMemcpyHostToDevice));
        FLX_TRY(whiten_dense(h, Cd.p, n, h->ncf, Wd.p, n));
        FLX_CUDA_TRY(cudaMemcpy(Cw.data(), Wd.p, C.size() * sizeof(double), cudaMemcpyDeviceToHost));
    }
    std::vector<double> Qc, Rc;
    h->cr = orth_basis(Cw.data(), n, h->ncf, Qc, Rc);
    const int cr = h->cr;
    for (int ci = 0; ci < h->ncf; ci++)
        for (int i = 0; i < cr; i++) fc.Rc[i + ci * 24] = Rc[i + (size_t)ci * cr];
    // qcy, residual r (two projection passes), rss_f
    std::vector<double> r(h->y_mm);
    std::vector<long double> qcy(std::max(cr, 1), 0.0L);
    for (int pass = 0; pass < 2; pass++)
        for (int j = 0; j < cr; j++) {
            const double* q = Qc.data() + (size_t)j * n;
            long double t = 0.0L;
            for (int64_t i = 0; i < n; i++) t += (long double)q[i] * r[i];
            qcy[j] += t;
            for (int64_t i = 0; i < n; i++) r[i] -= (double)t * q[i];
        }
    long double rss_f = 0.0L;
    for (int64_t i = 0; i < n; i++) rss_f += (long double)r[i] * r[i];
    for (int j = 0; j < cr; j++) fc.qcy[j] = (double)qcy[j];
    fc.rss_f = (double)rss_f;
    fc.n = (int)n; fc.s = h->s; fc.k = k; fc.cr = cr; fc.ncf = h->ncf; fc.df_null = h->df_null;
    // lrt_chisq = 2 (ll - ll.null):  log RSS0 := -2 ll.null / N - (log 2 pi + 1 - log N)   (stats:::logLik.lm)
    {
        const long double N = (long double)n;
        const long double log_rss0 = -2.0L * (long double)h->ll_null / N - (logl(2.0L * 3.14159265358979323846264338327950288L) + 1.0L - logl(N));
        fc.lrs0 = (double)(log_rss0 - logl(rss_f));
    }
    std::vector<double> tmp;
    upload_rowmajor_padded(Qc, n, cr, n_pad, tmp);
    FLX_TRY(h->Qc.ensure(tmp.size()));
    FLX_CUDA_TRY(cudaMemcpy(h->Qc.p, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice));
    r.resize(n_pad, 0.0);
    FLX_TRY(h->r.ensure(n_pad));
    FLX_CUDA_TRY(cudaMemcpy(h->r.p, r.data(), n_pad * sizeof(double), cudaMemcpyHostToDevice));
    // null-model basis (for the permutation null refit, fit_model.nf:100-101)
    std::vector<double> Qn, Rn;
    h->cn = orth_basis(h->X_null.data(), n, h->c_null, Qn, Rn);
    upload_rowmajor_padded(Qn, n, h->cn, n_pad, tmp);
    FLX_TRY(h->Qn.ensure(tmp.size()));
    FLX_CUDA_TRY(cudaMemcpy(h->Qn.p, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice));

    lap("constant columns, Qc, null basis");
    // ---- int8 quadratic-form path (K2q / K4q): every SNP term must factor as g_t(x) * c_t[sample] with an integer triple
    // |g_t| <= 2 (x itself, I(x==1), ...) and a per-sample scale c_t (1 for main effects, the covariate for x:cov).  Then
    //   A_raw[t][v] = g_t' C_t V^-1 C_v g_v = g_t' T^(t,v) g_v + g_v' T^(v,t) g_t,   T^(a,b) = tril(C_a V^-1 C_b), half diagonal,
    // each form an exact int8 contraction against the digit planes of a matrix prepared once per task.
    h->fast = false;
    h->fallback_reason = FLX_FALLBACK_NONE;
    h->tiles_auto[0] = h->tiles_auto[1] = 0;
    flx_handle::FastPlan& pl = h->plan;
    pl = flx_handle::FastPlan();
    std::vector<std::vector<double>> cvecs;
    std::vector<std::vector<uint8_t>> masks;
    int slot_of[2][2][flx_handle::MAX_NA][flx_handle::MAX_NA];
    memset(slot_of, -1, sizeof(slot_of));
    if (h->cfg.flags & FLX_FLAG_NO_INT8) h->fallback_reason = FLX_FALLBACK_FLAG;
    else if (!h->have_Z) h->fallback_reason = FLX_FALLBACK_NO_DENSE_INVERSE;
    else if (h->s > 4 || cr + 1 > 24) h->fallback_reason = FLX_FALLBACK_TOO_MANY_FORMS;
    else {
        // Two ways to lay a design out, tried in turn; the cheaper feasible one (in K2q pass units) wins:
        //   plain  — every distinct per-sample scale is a group with its own digit planes (x + I(x==1) + x:cov: 5 units);
        //   folded — 0/1 scales (the indicator columns of an x:factor interaction) become sample masks of the genotype code tiles,
        //            so those terms share the unscaled planes (x + I(x==1) + x:factor with three levels: 8 units instead of the
        //            FP64 route, which costs ~14).
        struct Trial { int why = FLX_FALLBACK_NONE; flx_handle::FastPlan pl; std::vector<std::vector<double>> cvecs; std::vector<std::vector<uint8_t>> masks;
                       int slot_of[2][2][flx_handle::MAX_NA][flx_handle::MAX_NA]; int units = 0; };
        auto build = [&](bool fold, Trial& T) {
            flx_handle::FastPlan& pl = T.pl;
            std::vector<std::vector<double>>& cvecs = T.cvecs;
            std::vector<std::vector<uint8_t>>& masks = T.masks;
            int& why = T.why;
            auto& slot_of = T.slot_of;
            for (int t = 0; t < h->s && !why; t++) {
                const double* src[3] = {h->M0.data() + (size_t)snp_cols[t] * n, h->M1.data() + (size_t)snp_cols[t] * n, h->M2.data() + (size_t)snp_cols[t] * n};
                // the sample with the largest entry fixes the integer triple
                int64_t ibest = 0; double vbest = 0.0;
                for (int64_t i = 0; i < n; i++)
                    for (int g = 0; g < 3; g++) if (std::fabs(src[g][i]) > vbest) { vbest = std::fabs(src[g][i]); ibest = i; }
                if (!(vbest > 0.0) || !std::isfinite(vbest)) { why = FLX_FALLBACK_NOT_SEPARABLE; break; }
                int a[3] = {0, 0, 0};
                bool found = false;
                for (int mlt = 1; mlt <= 2 && !found; mlt++) {
                    bool ints = true;
                    for (int g = 0; g < 3; g++) {
                        const double v = src[g][ibest] * mlt / vbest;
                        if (v != std::floor(v)) ints = false; else a[g] = (int)v;
                    }
                    found = ints;
                }
                if (!found) { why = FLX_FALLBACK_NOT_SEPARABLE; break; }
                int g0 = 0; while (g0 < 3 && a[g0] == 0) g0++;
                if (a[g0] < 0) for (int g = 0; g < 3; g++) a[g] = -a[g];
                int gs = 0; for (int g = 1; g < 3; g++) if (std::abs(a[g]) > std::abs(a[gs])) gs = g;
                std::vector<double> c(n);
                for (int64_t i = 0; i < n && !why; i++) {
                    c[i] = src[gs][i] / a[gs];                       // exact: a in {+-1, +-2}
                    for (int g = 0; g < 3; g++) if (src[g][i] != a[g] * c[i]) why = FLX_FALLBACK_NOT_SEPARABLE;
                }
                if (why) break;
                // a 0/1 scale becomes a sample mask of the tile set and the term joins the unscaled group
                bool binary = true, all_one = true;
                for (int64_t i = 0; i < n; i++) { binary = binary && (c[i] == 0.0 || c[i] == 1.0); all_one = all_one && (c[i] == 1.0); }
                int mk = -1;
                if (fold && binary && !all_one) {
                    std::vector<uint8_t> mv(n);
                    for (int64_t i = 0; i < n; i++) mv[i] = c[i] != 0.0 ? 0xFF : 0x00;
                    for (int q = 0; q < (int)masks.size(); q++) if (masks[q] == mv) mk = q;
                    if (mk < 0) { mk = (int)masks.size(); masks.push_back(std::move(mv)); }
                    std::fill(c.begin(), c.end(), 1.0);
                }
                int al = -1, gr = -1;
                for (int q = 0; q < pl.NA; q++) if (pl.aval[q][0] == a[0] && pl.aval[q][1] == a[1] && pl.aval[q][2] == a[2] && pl.amask[q] == mk) al = q;
                if (al < 0) {
                    if (pl.NA == flx_handle::MAX_NA) { why = FLX_FALLBACK_TOO_MANY_FORMS; break; }
                    al = pl.NA++;
                    for (int g = 0; g < 3; g++) pl.aval[al][g] = a[g];
                    pl.amask[al] = mk;
                }
                for (int q = 0; q < pl.NG; q++) if (memcmp(cvecs[q].data(), c.data(), n * sizeof(double)) == 0) gr = q;
                if (gr < 0) {
                    if (pl.NG == 2) { why = FLX_FALLBACK_TOO_MANY_FORMS; break; }
                    gr = pl.NG++;
                    bool ones = true;
                    for (int64_t i = 0; i < n; i++) ones = ones && (c[i] == 1.0);
                    pl.ones[gr] = ones;
                    cvecs.push_back(std::move(c));
                }
                pl.alp[t] = al; pl.grp[t] = gr;
            }
            if (why) return;
            constexpr int MA = flx_handle::MAX_NA;
            memset(slot_of, -1, sizeof(slot_of));
            bool used[2][MA] = {};
            for (int t = 0; t < h->s; t++) used[pl.grp[t]][pl.alp[t]] = true;
            for (int a = 0; a < pl.NG; a++)
                for (int b = a; b < pl.NG; b++)
                    for (int ar = 0; ar < pl.NA; ar++) {
                        if (!used[b][ar]) continue;
                        // the epilogue of a pass dots with at most two left vectors: more tile sets on the left take more passes
                        flx_handle::QfPass ps{};
                        ps.mat = a * pl.NG + b; ps.amma = ar; ps.ndot = 0;
                        for (int al = 0; al < pl.NA; al++) {
                            if (!used[a][al]) continue;
                            ps.adot[ps.ndot] = al; ps.slot[ps.ndot] = pl.nslots; slot_of[a][b][al][ar] = pl.nslots++; ps.ndot++;
                            if (ps.ndot == 2) { pl.passes.push_back(ps); ps.ndot = 0; }
                        }
                        if (ps.ndot > 0) pl.passes.push_back(ps);
                    }
            if (pl.nslots > 16) { why = FLX_FALLBACK_TOO_MANY_FORMS; return; }
            for (const auto& ps : pl.passes) T.units += (ps.mat / pl.NG != ps.mat % pl.NG) ? 2 : 1;
        };
        Trial plain, folded;
        build(false, plain);
        build(true, folded);
        Trial* best = nullptr;
        if (!plain.why) best = &plain;
        if (!folded.why && (!best || folded.units < best->units)) best = &folded;
        h->fallback_reason = best ? FLX_FALLBACK_NONE : (plain.why == FLX_FALLBACK_NOT_SEPARABLE ? plain.why : folded.why);
        h->fast = best != nullptr;
        if (best) { pl = best->pl; cvecs = best->cvecs; masks = best->masks; memcpy(slot_of, best->slot_of, sizeof(slot_of)); }
    }
    if (h->fast) {
        // passes: one K2q launch per (matrix, tile set on the right), its epilogue dots with the tile sets used on the left.
        //   same scale a:      T^(a,a) = tril(C_a V^-1 C_a), half diagonal: A_raw[t][v] = g_t' T g_v + g_v' T g_t
        //   scales a < b:      the FULL matrix F^(a,b) = C_a V^-1 C_b:        A_raw[t][v] = g_t' F g_v  (t in a, v in b)
        // (a full pass costs two triangular ones, but one full pass with two left vectors replaces three triangular passes
        //  of the x + I(x==1) + x:cov model: 5 pass units instead of 6)
        pl.nmask = (int)masks.size();
        if (pl.nmask > 0) {
            const size_t np256 = (size_t)((n + 255) / 256) * 256;
            std::vector<uint8_t> mh((size_t)pl.nmask * np256, 0);
            for (int q = 0; q < pl.nmask; q++) memcpy(mh.data() + (size_t)q * np256, masks[q].data(), n);
            FLX_TRY(h->mask_d.ensure(mh.size()));
            FLX_CUDA_TRY(cudaMemcpy(h->mask_d.p, mh.data(), mh.size(), cudaMemcpyHostToDevice));
        }
        for (int t = 0; t < h->s; t++)
            for (int v = 0; v < h->s; v++) {
                const int a = pl.grp[t], b = pl.grp[v];
                if (a == b) {
                    pl.slot_a[t][v] = (int8_t)slot_of[a][a][pl.alp[t]][pl.alp[v]];
                    pl.slot_b[t][v] = (int8_t)slot_of[a][a][pl.alp[v]][pl.alp[t]];
                } else {
                    pl.slot_a[t][v] = (int8_t)(a < b ? slot_of[a][b][pl.alp[t]][pl.alp[v]] : slot_of[b][a][pl.alp[v]][pl.alp[t]]);
                    pl.slot_b[t][v] = -1;
                }
            }
        const double one = 1.0, zero = 0.0;
        h->T256 = (int)((n + 255) / 256);
        h->NKB = h->T256 * 4;
        std::vector<double> cpad((size_t)pl.NG * n_pad, 0.0);
        for (int gq = 0; gq < pl.NG; gq++) memcpy(cpad.data() + (size_t)gq * n_pad, cvecs[gq].data(), n * sizeof(double));
        FLX_TRY(h->cvec_d.ensure(cpad.size()));
        FLX_CUDA_TRY(cudaMemcpy(h->cvec_d.p, cpad.data(), cpad.size() * sizeof(double), cudaMemcpyHostToDevice));
        auto cptr = [&](int gq) -> const double* { return pl.ones[gq] ? nullptr : h->cvec_d.p + (size_t)gq * n_pad; };
        // H = L^-T [Qc | r]  (n x (cr+1))
        std::vector<double> Hc((size_t)n * (cr + 1));
        memcpy(Hc.data(), Qc.data(), (size_t)n * cr * sizeof(double));
        memcpy(Hc.data() + (size_t)n * cr, r.data(), (size_t)n * sizeof(double));
        FLX_TRY(h->Hd.ensure(Hc.size()));
        FLX_CUDA_TRY(cudaMemcpyAsync(h->Hd.p, Hc.data(), Hc.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        FLX_CUBLAS_TRY(cublasDtrmm_64(h->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, (int)n, cr + 1, &one, h->Zdense.p, (int)n, h->Hd.p, (int)n, h->Hd.p, (int)n));
        FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    // digit planes of G = C_g [H | L^-T R_f] for K4q, one set per scale group; without permutations G = C_g H
    auto slice_G = [&](const double* G2, int64_t ld2, int P) -> int {
        const flx_handle::FastPlan& pl = h->plan;
        const int NC = cr + 1 + P;
        const int nb_max = xg_block_cols(h->Q), gran = (h->Q % 2 == 0) ? 8 : 16;
        h->ntilesG = (NC + nb_max - 1) / nb_max;
        h->NB_G = (((NC + h->ntilesG - 1) / h->ntilesG + gran - 1) / gran) * gran;
        for (int gq = 0; gq < pl.NG; gq++) {
            FLX_TRY(h->G_scale[gq].ensure((size_t)h->ntilesG * h->NB_G));
            FLX_TRY(h->G8[gq].ensure((size_t)h->ntilesG * h->NKB * h->Q * h->NB_G * 64));
            FLX_TRY(launch_xg_slice(h->Hd.p, n, cr + 1, G2, ld2, n, NC, h->NB_G, h->ntilesG, h->Q, h->NKB, pl.ones[gq] ? nullptr : h->cvec_d.p + (size_t)gq * n_pad,
                                    h->G_scale[gq].p, h->G8[gq].p, h->stream));
        }
        return 0;
    };

    lap("int8 plan, H = L^-T [Qc | r]");
    // ---- permutation prologue (fit_model.nf:96-101) for all seeds at once
    h->P_pad = 0; h->NP = 0;
    if (h->have_perm && h->P > 0) {
        const int P = h->P;
        const int64_t m = h->m;
        h->NP = (P + TJ - 1) / TJ;
        h->P_pad = (int64_t)h->NP * TJ;
        std::vector<int32_t> idx((size_t)P * m);
        {
            const unsigned hw = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
            const unsigned nt = std::min<unsigned>(hw, (unsigned)P);
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; t++)
                th.emplace_back([&, t]() { for (int p = (int)t; p < P; p += (int)nt) r_sample_int(h->seeds[p], m, idx.data() + (size_t)p * m); });
            for (auto& x : th) x.join();
        }
        DevBuf<int32_t> idx_d;
        DevBuf<double> ep_d, E, UE, ypred_d, rssf_d, rss0_d;
        FLX_TRY(idx_d.ensure(idx.size()));
        FLX_TRY(ep_d.ensure(m));
        FLX_TRY(E.ensure((size_t)m * P));
        FLX_TRY(UE.ensure((size_t)n_pad * h->P_pad));
        FLX_TRY(ypred_d.ensure(n));
        FLX_TRY(rssf_d.ensure(P));
        FLX_TRY(rss0_d.ensure(P));
        FLX_CUDA_TRY(cudaMemcpyAsync(idx_d.p, idx.data(), idx.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        FLX_CUDA_TRY(cudaMemcpyAsync(ep_d.p, h->e_p.data(), m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        if (!h->perm_implicit && !h->U1_d.p) { set_error("flx_run: the permutation inputs were consumed by an earlier setup; call flx_set_perm again"); return FLX_ERR_STATE; }
        FLX_CUDA_TRY(cudaMemcpyAsync(ypred_d.p, h->y_pred.data(), n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        FLX_CUDA_TRY(cudaMemsetAsync(UE.p, 0, (size_t)n_pad * h->P_pad * sizeof(double), h->stream));
        const double one = 1.0, zero = 0.0;
        if (h->perm_implicit) {
            // U1 = Q[:, rank:] of the Householder QR of the null design, applied as reflectors: O(n c P), no n x m matrix
            DevBuf<double> Vd, bd;
            const int rk = h->nullfit.rank;
            FLX_TRY(Vd.ensure((size_t)n * std::max(rk, 1)));
            FLX_TRY(bd.ensure(std::max(rk, 1)));
            FLX_CUDA_TRY(cudaMemcpyAsync(Vd.p, h->nullfit.V.data(), (size_t)n * std::max(rk, 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            FLX_CUDA_TRY(cudaMemcpyAsync(bd.p, h->nullfit.beta.data(), std::max(rk, 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            FLX_TRY(launch_implicit_u1(ep_d.p, idx_d.p, m, P, Vd.p, bd.p, rk, n, n_pad, UE.p, h->stream));
            FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
        } else {
            FLX_TRY(launch_gather_ep(ep_d.p, idx_d.p, m, P, E.p, h->stream));
            // U1 %*% sample(e.p) for every seed: one plain library GEMM (setup, outside the hot loop)
            FLX_CUBLAS_TRY(cublasDgemm_64(h->cublas, CUBLAS_OP_N, CUBLAS_OP_N, (int)n, P, (int)m, &one, h->U1_d.p, (int)n, E.p, (int)m, &zero, UE.p, (int)n_pad));
        }
        PermPrepArgs pa{};
        pa.UE = UE.p; pa.y_pred = ypred_d.p; pa.Qc = h->Qc.p; pa.cr = cr; pa.crp = gram_padded_width(cr); pa.Qn = h->Qn.p; pa.cn = h->cn; pa.cnp = gram_padded_width(h->cn);
        pa.n = n; pa.n_pad = n_pad; pa.P = P; pa.rss_f = rssf_d.p; pa.rss0 = rss0_d.p;
        FLX_TRY(launch_perm_prep(pa, h->stream));
        FLX_TRY(h->Rpk.ensure((size_t)h->NP * h->KC * CHUNK));
        FLX_TRY(launch_pack_cols(UE.p, n, n_pad, h->P_pad, h->KC, h->Rpk.p, h->stream));
        int64_t inv_len = h->P_pad;
        if (h->fast) {
            // int8 path: the contraction runs against the raw genotypes, b* = x' (L^-T r*): digit planes of L^-T R_f
            FLX_CUBLAS_TRY(cublasDtrmm_64(h->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, (int)n, P, &one, h->Zdense.p, (int)n, UE.p, (int)n_pad, UE.p, (int)n_pad));
            FLX_TRY(slice_G(UE.p, n_pad, P));
        }
        std::vector<double> rssf(P), rss0(P);
        FLX_CUDA_TRY(cudaMemcpyAsync(rssf.data(), rssf_d.p, P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        FLX_CUDA_TRY(cudaMemcpyAsync(rss0.data(), rss0_d.p, P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
        std::vector<double> inv(inv_len, 0.0), lr(P);
        for (int p = 0; p < P; p++) {
            inv[p] = 1.0 / rssf[p];
            lr[p] = (double)(logl((long double)rss0[p]) - logl((long double)rssf[p]));
        }
        FLX_TRY(h->inv_rss.ensure(inv_len));
        FLX_TRY(h->lrs0p.ensure(P));
        FLX_TRY(h->maxratio.ensure((size_t)h->s * std::max<int64_t>(h->P_pad, inv_len)));
        FLX_TRY(h->minp_d.ensure(P));
        FLX_CUDA_TRY(cudaMemcpy(h->inv_rss.p, inv.data(), inv_len * sizeof(double), cudaMemcpyHostToDevice));
        FLX_CUDA_TRY(cudaMemcpy(h->lrs0p.p, lr.data(), P * sizeof(double), cudaMemcpyHostToDevice));
        FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->U1_d.release();
    }
    lap("permutation prologue, G planes");
    if (h->fast && !(h->have_perm && h->P > 0)) FLX_TRY(slice_G(nullptr, 0, 0));
    if (h->fast) {
        // V^-1 = L^-T L^-1 one 256-row tile at a time (a plain library GEMM over the non-zero k range of the triangular
        // factor), each tile cut into the digit planes of tril(C_a V^-1 C_b): no second dense n x n matrix at any n
        const flx_handle::FastPlan& pl = h->plan;
        // the plane sets must fit next to the dense inverse and the batch buffers; if they do not (two-scale models at
        // n ~ 10^5), the task runs on the FP64 route instead of failing
        size_t free_b = 0, total_b = 0;
        FLX_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const size_t planes_b = (size_t)pl.NG * pl.NG * 2 * h->T256 * (h->T256 + 1) * h->Q * 16384;   // a full set = two triangular ones
        const size_t batch_b = (size_t)h->num_sms * h->NKB * 2048 * pl.NA + (size_t)pl.nslots * h->T256 * h->Q * h->num_sms * 128 * 8 + ((size_t)2 << 30);   // one tile per SM at least
        if (planes_b + batch_b + (size_t)1024 * n * 8 > free_b) {
            h->fast = false;
            h->fallback_reason = FLX_FALLBACK_PLANES_DONT_FIT;
        }
    }
    if (h->fast) {
        const flx_handle::FastPlan& pl = h->plan;
        const double one = 1.0, zero = 0.0;
        // row tiles are produced four at a time (1024 rows per library GEMM: the 256-row products ran at ~60 % of the DGEMM rate)
        constexpr int64_t RT = 1024;
        DevBuf<double> Vt;
        FLX_TRY(Vt.ensure((size_t)RT * n));
        const size_t tri_b = (size_t)2 * h->T256 * (h->T256 + 1) * h->Q * 16384, full_b = (size_t)h->T256 * h->NKB * h->Q * 16384;
        for (int a = 0; a < pl.NG; a++)
            for (int b = a; b < pl.NG; b++) {
                FLX_TRY(h->rowscale[a * pl.NG + b].ensure((size_t)h->T256 * 256));
                FLX_TRY(h->T8[a * pl.NG + b].ensure(a == b ? tri_b : full_b));
            }
        for (int64_t r0 = 0; r0 < n; r0 += RT) {
            const int64_t m = std::min<int64_t>(RT, n - r0), ncols = (pl.NG > 1) ? n : r0 + m, K = n - r0;
            FLX_CUBLAS_TRY(cublasDgemm_64(h->cublas, CUBLAS_OP_T, CUBLAS_OP_N, m, ncols, K, &one, h->Zdense.p + r0 + r0 * n, n, h->Zdense.p + r0, n, &zero, Vt.p, RT));
            for (int64_t sub = 0; sub < m; sub += 256) {
                const int I = (int)((r0 + sub) / 256);
                for (int a = 0; a < pl.NG; a++)
                    for (int b = a; b < pl.NG; b++) {
                        const int mi = a * pl.NG + b;
                        const double* ca = pl.ones[a] ? nullptr : h->cvec_d.p + (size_t)a * n_pad;
                        const double* cb = pl.ones[b] ? nullptr : h->cvec_d.p + (size_t)b * n_pad;
                        FLX_TRY(launch_qf_slice(Vt.p + sub, n, RT, I, h->T256, h->Q, ca, cb, a == b ? 0 : 1, h->rowscale[mi].p, h->T8[mi].p, h->stream));
                    }
            }
        }
        FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    lap("V^-1 row tiles -> digit planes");
    FLX_TRY(h->df_count.ensure(16));
    FLX_TRY(h->missing.ensure(1));
    // a task that cannot use the int8 tensor-core route is ~14x slower: never silently (flx_stats.fallback_reason carries the code)
    if (h->fallback_reason > FLX_FALLBACK_FLAG) {
        static const char* why[] = {"", "", "a SNP term is not (integer genotype triple) x (per-sample scale)", "more distinct triples / scales / SNP terms than the int8 plan holds",
                                    "the dense L^-1 was released after an earlier setup (call flx_set_whitening again before re-setting the design / permutations)",
                                    "the digit planes do not fit in device memory", "real-valued dosages"};
        fprintf(stderr, "[flx] warning: this task runs on the FP64 route (~14x slower than the int8 tensor-core route): %s\n", why[std::min(h->fallback_reason, 6)]);
    }
    // the dense inverse is only needed again if the null model / design / permutations are re-set: keep it while the device has
    // room for it next to a default batch (n = 50,000: 20 GB, kept; n = 100,000: 80 GB next to 70 GB of panels and planes, released)
    if (h->have_Z) {
        size_t free_b = 0, total_b = 0;
        FLX_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const size_t batch_b = h->fast ? (size_t)4 * h->num_sms * ((size_t)h->plan.NA * h->NKB * 2048 + (size_t)h->plan.nslots * h->T256 * h->Q * 128 * 8 + (size_t)h->s * 128 * h->ntilesG * h->NB_G * 8)
                                       : (size_t)2 * h->num_sms * h->KC * CHUNK * 8;
        if (free_b < batch_b + ((size_t)8 << 30)) { h->Zdense.release(); h->have_Z = false; }
    }
    h->dirty = false;
    return FLX_OK;
}

// ------------------------------------------------------------------------------------------------
// use_fast: K2q (int8 quadratic form) instead of K2 + K3a sweep 2.  accumulate: second pass over a subset (refit of flagged
// SNPs on the FP64 path) — run-wide accumulators are kept and results go to the slots out_map_host[q].
static int run_batches(flx_handle* h, const int32_t* var_idx, int64_t M, bool want_results, bool do_perm,
                       double* X_out_dense, uint8_t* codes_out, bool use_fast = false, bool accumulate = false,
                       const int32_t* out_map_host = nullptr, int64_t M_total = 0) {
    if (M_total < M) M_total = M;
    TileShape ts = make_tile_shape(h->s);
    if (use_fast) ts.snps = 128;          // code tiles hold 128 SNPs whatever the number of terms
    // default batch: one column tile per SM on the FP64 route (the X / W tiles are 8 n_pad bytes per column); four on the int8
    // route, where a tile is 128 n bytes and K1q / K3b / K4q want more SNPs in flight than one tile per SM (K2q is indifferent),
    // capped so that the batch buffers stay below a quarter of the free memory
    int tiles_per_batch = h->cfg.batch_tiles > 0 ? h->cfg.batch_tiles : h->num_sms;
    if (h->cfg.batch_tiles <= 0 && h->tiles_auto[use_fast ? 1 : 0] > 0) tiles_per_batch = h->tiles_auto[use_fast ? 1 : 0];
    else if (h->cfg.batch_tiles <= 0 && use_fast) {
        size_t free_b = 0, total_b = 0;
        FLX_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const size_t per_tile = (size_t)h->plan.NA * h->NKB * 2048 + (size_t)h->plan.nslots * h->T256 * h->Q * 128 * 8 +
                                (size_t)h->s * 128 * h->ntilesG * h->NB_G * 8 + (size_t)128 * 40 * 8 * (h->s + 2);
        const size_t have = free_b + h->X2.cap * 4 + h->qf_part.cap * 8 + h->ub.cap * 8;     // what this handle already holds is reusable
        int mult = 4;
        while (mult > 1 && (size_t)mult * h->num_sms * per_tile > have / 4) mult--;
        tiles_per_batch = mult * h->num_sms;
        h->tiles_auto[1] = tiles_per_batch;
    }
    // a run smaller than one batch sizes every batch buffer for itself, not for the default batch
    tiles_per_batch = (int)std::max<int64_t>(1, std::min<int64_t>(tiles_per_batch, (M + ts.snps - 1) / ts.snps));
    const int64_t snps_per_batch = (int64_t)tiles_per_batch * ts.snps;
    // batch boundaries; when the genotypes come from the host, the first batches ramp up (1/4, 1/2, then full size) so that the
    // first H2D, which nothing can overlap, is short
    std::vector<int64_t> bstart;
    {
        const int64_t unit = (int64_t)h->num_sms * ts.snps;
        int64_t pos = 0, step = (!h->is_resident && snps_per_batch >= 4 * unit) ? snps_per_batch / 4 : snps_per_batch;
        while (pos < M) { bstart.push_back(pos); pos += step; step = std::min(snps_per_batch, 2 * step); }
        bstart.push_back(M);
    }
    const int64_t nbatch = (int64_t)bstart.size() - 1;
    const int64_t bpv = h->bpv;
    const int nsym = h->s * (h->s + 1) / 2;
    const int cr = h->cr;
    const int crp = gram_padded_width(cr);
    constexpr int KS_MAX = 32;
    const bool decode_only = (X_out_dense != nullptr) || (codes_out != nullptr);

    // validate indices; contiguity lets the resident path skip the index map
    bool contiguous = true;
    for (int64_t i = 0; i < M; i++) {
        if (var_idx[i] < 1 || var_idx[i] > h->src_variants) { set_error("var_idx[%lld]=%d out of range 1..%lld", (long long)i, var_idx[i], (long long)h->src_variants); return FLX_ERR_BAD_ARG; }
        if (i > 0 && var_idx[i] != var_idx[i - 1] + 1) contiguous = false;
    }
    const size_t tile_elems = (size_t)h->KC * CHUNK;
    if (!use_fast || decode_only) FLX_TRY(h->X.ensure((size_t)tiles_per_batch * tile_elems));
    const size_t x2_stride = (size_t)((tiles_per_batch + 1) / 2) * h->NKB * 1024;     // words per tile set: tile pairs x k-blocks x (2 x 512)
    const int64_t ldg = (int64_t)h->ntilesG * h->NB_G;     // row length of the K4q output: u (cr), b, b* (P), padding
    if (!decode_only) {
        if (!use_fast) FLX_TRY(h->W.ensure((size_t)tiles_per_batch * tile_elems));
        if (use_fast) {
            FLX_TRY(h->X2.ensure(x2_stride * h->plan.NA));
            FLX_TRY(h->qf_part.ensure((size_t)h->plan.nslots * h->T256 * h->Q * snps_per_batch));
            FLX_TRY(h->ub.ensure((size_t)h->s * snps_per_batch * ldg));
            FLX_TRY(h->refit_flag.ensure(M_total));
        }
        FLX_TRY(h->n_part.ensure((size_t)KS_MAX * snps_per_batch * h->s));
        FLX_TRY(h->u_part.ensure((size_t)KS_MAX * snps_per_batch * h->s * std::max(crp, 32)));
        FLX_TRY(h->a_part.ensure((size_t)KS_MAX * snps_per_batch * nsym));
        FLX_TRY(h->b_part.ensure((size_t)KS_MAX * snps_per_batch * h->s));
        FLX_TRY(h->snp_sm.ensure((size_t)snps_per_batch * nsym));
        FLX_TRY(h->snp_df.ensure((size_t)snps_per_batch));
        FLX_TRY(h->chisq.ensure(M_total)); FLX_TRY(h->pval.ensure(M_total)); FLX_TRY(h->df.ensure(M_total)); FLX_TRY(h->rank.ensure(M_total));
        FLX_TRY(h->pivot.ensure((size_t)M_total * h->k)); FLX_TRY(h->coef.ensure((size_t)M_total * h->k));
        if (want_results) {
            FLX_TRY(h->chisq_h.ensure(M_total)); FLX_TRY(h->pval_h.ensure(M_total)); FLX_TRY(h->df_h.ensure(M_total)); FLX_TRY(h->rank_h.ensure(M_total));
            FLX_TRY(h->pivot_h.ensure((size_t)M_total * h->k)); FLX_TRY(h->coef_h.ensure((size_t)M_total * h->k));
        }
        if (out_map_host) {
            FLX_TRY(h->out_map_d.ensure(M));
            FLX_CUDA_TRY(cudaMemcpyAsync(h->out_map_d.p, out_map_host, M * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
            FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
        }
    }
    DevBuf<uint8_t> codes_d;
    if (codes_out) FLX_TRY(codes_d.ensure((size_t)snps_per_batch * h->n));
    const bool use_resident = h->is_resident;
    if (use_resident && !contiguous) {
        std::vector<int32_t> v0(M);
        for (int64_t i = 0; i < M; i++) v0[i] = var_idx[i] - 1;
        FLX_TRY(h->vmap_d.ensure(M));
        FLX_CUDA_TRY(cudaMemcpyAsync(h->vmap_d.p, v0.data(), M * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    if (!use_resident)
        for (int i = 0; i < 2; i++) {
            FLX_TRY(h->packed_d[i].ensure((size_t)snps_per_batch * bpv));
        }
    const int int_max = INT_MAX;
    FLX_TRY(h->px_err.ensure(1));
    FLX_CUDA_TRY(cudaMemsetAsync(h->px_err.p, 0, sizeof(int), h->stream));
    if (!accumulate) {
        FLX_CUDA_TRY(cudaMemcpyAsync(h->missing.p, &int_max, sizeof(int), cudaMemcpyHostToDevice, h->stream));
        FLX_CUDA_TRY(cudaMemsetAsync(h->df_count.p, 0, 16 * sizeof(int), h->stream));
        if (do_perm) FLX_CUDA_TRY(cudaMemsetAsync(h->maxratio.p, 0, (size_t)h->s * h->P_pad * sizeof(double), h->stream));
    }

    // stage stopwatch: events come from the handle's pool (created once, reused by every run; nothing to leak on an early return)
    size_t nev = 0;
    auto mark = [&]() {
        if (nev == h->ev_pool.size()) { cudaEvent_t e = nullptr; cudaEventCreate(&e); h->ev_pool.push_back(e); }
        cudaEventRecord(h->ev_pool[nev++], h->stream);
    };
    std::vector<cudaEvent_t>& evs = h->ev_pool;
    std::vector<uint8_t> xhost;

    for (int64_t bi = 0; bi < nbatch; bi++) {
        const int64_t s0 = bstart[bi];
        const int64_t nsnp = bstart[bi + 1] - s0;
        const int NJ = (int)((nsnp + ts.snps - 1) / ts.snps);
        const int buf = (int)(bi & 1);
        const uint8_t* packed_dev;
        const int32_t* vmap = nullptr;
        mark();  // [0] batch start / h2d start
        if (use_resident) {
            if (contiguous) packed_dev = h->resident.p + (size_t)(var_idx[s0] - 1) * bpv;
            else { packed_dev = h->resident.p; vmap = h->vmap_d.p + s0; }
        } else {
            // H2D on the copy stream so that the transfer of batch b+1 overlaps the kernels of batch b
            FLX_CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_used[buf], 0));   // the decode of batch b-2 has read this buffer
            const uint8_t* src_ptr = nullptr;
            bool expanded_on_device = false;
            if (h->src == 2 && contiguous && (h->host_registered || h->host_pinned_by_owner)) {
                src_ptr = h->packed_host + (size_t)(var_idx[s0] - 1) * bpv;          // pinned in place: no staging copy
            } else {
                if (bi >= 2) FLX_CUDA_TRY(cudaEventSynchronize(h->ev_h2d[buf]));      // the staging buffer's previous H2D is done
                if (h->src == 2 || h->pgen->mode() == 0x02) FLX_TRY(h->staging[buf].ensure((size_t)snps_per_batch * bpv));    // pinned, allocated only where used
                uint8_t* st = h->staging[buf].p;
                if (h->src == 2) {
                    if (contiguous) memcpy(st, h->packed_host + (size_t)(var_idx[s0] - 1) * bpv, (size_t)nsnp * bpv);
                    else for (int64_t i = 0; i < nsnp; i++) memcpy(st + i * bpv, h->packed_host + (size_t)(var_idx[s0 + i] - 1) * bpv, bpv);
                } else if (h->src == 1 && h->pgen->mode() == 0x02) {
                    for (int64_t i = 0; i < nsnp; i++) memcpy(st + i * bpv, h->pgen->fixed_record((uint32_t)(var_idx[s0 + i] - 1)), bpv);
                } else if (h->src == 1) {
                    // K0: ship the COMPRESSED records (plus LD bases that are not part of the batch) and expand on the device
                    struct Ent { uint64_t off; uint32_t len; uint8_t vt; int32_t base_row, out_row; };
                    std::vector<Ent> ents((size_t)nsnp);
                    std::unordered_map<uint32_t, int32_t> row_of;
                    bool any_ld = false;
                    for (int64_t i = 0; i < nsnp; i++) {
                        const uint32_t v = (uint32_t)(var_idx[s0 + i] - 1);
                        const PgenFile::RecInfo ri = h->pgen->rec_info(v);
                        if (ri.vrtype & 0x08) { set_error("pgen variant %d: multiallelic record (the pipeline converts with --max-alleles 2)", var_idx[s0 + i]); return FLX_ERR_PGEN; }
                    if (h->run_dosage && (ri.vrtype & 0x60) && (ri.vrtype & 0x90)) { set_error("pgen variant %d: phased record with a dosage track is not supported", var_idx[s0 + i]); return FLX_ERR_PGEN; }
                        ents[i] = Ent{ri.off, ri.len, ri.vrtype, 0, (int32_t)i};
                        if ((ri.vrtype & 6) != 2) row_of.emplace(v, (int32_t)i); else any_ld = true;
                    }
                    int32_t nbase = 0;
                    if (any_ld)
                        for (int64_t i = 0; i < nsnp; i++) {
                            if ((ents[i].vt & 6) != 2) continue;
                            const uint32_t b = h->pgen->rec_info((uint32_t)(var_idx[s0 + i] - 1)).ldbase;
                            auto it = row_of.find(b);
                            if (it == row_of.end()) {
                                const PgenFile::RecInfo rb = h->pgen->rec_info(b);
                                nbase++;
                                ents.push_back(Ent{rb.off, rb.len, rb.vrtype, 0, -nbase});
                                it = row_of.emplace(b, -nbase).first;
                            }
                            ents[i].base_row = it->second;
                        }
                    const size_t E = ents.size();
                    std::vector<uint32_t> roff(E);
                    size_t blob_bytes = 0;
                    for (size_t e = 0; e < E; e++) { roff[e] = (uint32_t)blob_bytes; blob_bytes += ents[e].len; }
                    if (blob_bytes >= ((size_t)1 << 32)) { set_error("pgen batch of %lld records exceeds 4 GB of compressed bytes: lower batch_tiles", (long long)nsnp); return FLX_ERR_BAD_ARG; }
                    const size_t o_off = (blob_bytes + 15) / 16 * 16, o_len = o_off + 4 * E, o_base = o_len + 4 * E, o_out = o_base + 4 * E, o_vt = o_out + 4 * E;
                    const size_t total = (o_vt + E + 15) / 16 * 16;
                    FLX_TRY(h->staging[buf].ensure(total));
                    st = h->staging[buf].p;
                    {
                        const uint8_t* file = h->pgen->data();
                        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(std::min(std::thread::hardware_concurrency(), 16u), E / 1024));
                        auto work = [&](unsigned ti) {
                            for (size_t e = E * ti / nt; e < E * (ti + 1) / nt; e++) memcpy(st + roff[e], file + ents[e].off, ents[e].len);
                        };
                        if (nt == 1) work(0);
                        else {
                            std::vector<std::thread> th;
                            for (unsigned ti = 0; ti < nt; ti++) th.emplace_back(work, ti);
                            for (auto& x : th) x.join();
                        }
                    }
                    for (size_t e = 0; e < E; e++) {
                        memcpy(st + o_off + 4 * e, &roff[e], 4); memcpy(st + o_len + 4 * e, &ents[e].len, 4);
                        memcpy(st + o_base + 4 * e, &ents[e].base_row, 4); memcpy(st + o_out + 4 * e, &ents[e].out_row, 4);
                        st[o_vt + e] = ents[e].vt;
                    }
                    FLX_TRY(h->blob_d[buf].ensure(total));
                    FLX_TRY(h->px_scratch[buf].ensure((size_t)std::max(nbase, 1) * bpv));
                    FLX_CUDA_TRY(cudaMemcpyAsync(h->blob_d[buf].p, st, total, cudaMemcpyHostToDevice, h->copy_stream));
                    FLX_CUDA_TRY(cudaEventRecord(h->ev_h2d[buf], h->copy_stream));
                    FLX_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_h2d[buf], 0));
                    h->stats.h2d_bytes += (int64_t)total;
                    PgenExpandArgs px{};
                    const uint8_t* bd = h->blob_d[buf].p;
                    px.blob = bd; px.rec_off = reinterpret_cast<const uint32_t*>(bd + o_off); px.rec_len = reinterpret_cast<const uint32_t*>(bd + o_len);
                    px.base_row = reinterpret_cast<const int32_t*>(bd + o_base); px.out_row = reinterpret_cast<const int32_t*>(bd + o_out); px.vrtype = bd + o_vt;
                    px.E = (int)E; px.sample_ct = (uint32_t)h->raw_n; px.bpv = bpv; px.out = h->packed_d[buf].p; px.scratch = h->px_scratch[buf].p;
                    px.err = h->px_err.p; px.any_ld = any_ld ? 1 : 0;
                    if (h->run_dosage) {
                        FLX_TRY(h->main_end_d[buf].ensure(E));
                        FLX_TRY(h->dos_d[buf].ensure((size_t)snps_per_batch * h->raw_n));
                        FLX_CUDA_TRY(cudaMemsetAsync(h->dos_d[buf].p, 0xFF, (size_t)nsnp * h->raw_n * sizeof(uint16_t), h->stream));
                        px.main_end = h->main_end_d[buf].p; px.dos_out = h->dos_d[buf].p; px.dos_ld = h->raw_n;
                    }
                    FLX_TRY(launch_pgen_expand(px, h->stream));
                    h->stats.launches += any_ld ? 2 : 1;
                    if (h->run_dosage) { FLX_TRY(launch_pgen_dosage(px, h->stream)); h->stats.launches++; }
                    expanded_on_device = true;
                } else { set_error("flx_run: no genotype source (flx_open_pgen / flx_set_packed)"); return FLX_ERR_STATE; }
                src_ptr = st;
            }
            if (!expanded_on_device) {
                FLX_CUDA_TRY(cudaMemcpyAsync(h->packed_d[buf].p, src_ptr, (size_t)nsnp * bpv, cudaMemcpyHostToDevice, h->copy_stream));
                FLX_CUDA_TRY(cudaEventRecord(h->ev_h2d[buf], h->copy_stream));
                FLX_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_h2d[buf], 0));
                h->stats.h2d_bytes += nsnp * bpv;
            }
            packed_dev = h->packed_d[buf].p;
        }
        mark();  // [1] decode start
        DecodeArgs da{};
        da.packed = packed_dev; da.bpv = bpv; da.vmap = vmap;
        da.order = h->identity_order ? nullptr : h->order_d.p;
        da.n = h->n; da.KC = h->KC; da.nsnp = nsnp; da.snp_offset = s0; da.s = h->s;
        da.lut = h->lut.p; memcpy(da.uni, h->uni, sizeof(da.uni)); da.uniform_mask = h->uniform_mask;
        da.n_pad = h->n_pad; da.X = h->X.p; da.NJ = NJ; da.missing_flag = h->missing.p;
        da.codes_out = codes_out ? codes_d.p : nullptr;
        if (h->run_dosage && !use_resident && h->src == 1 && h->pgen->mode() == 0x10) {
            da.dosage = h->dos_d[buf].p; da.dos_ld = h->raw_n; da.dos_indicator = h->dos_indicator;
            memcpy(da.dos_g, h->dos_g, sizeof(da.dos_g));
        }
        if (use_fast && !decode_only) {
            // K1q: one 2-bit code tile set per (integer triple, sample mask) of the design (x itself, I(x==1), ...)
            for (int al = 0; al < h->plan.NA; al++) {
                uint32_t vt_unused;
                tile_code_map(h->plan.aval[al], &da.gmap, &vt_unused);
                da.smask = h->plan.amask[al] >= 0 ? h->mask_d.p + (size_t)h->plan.amask[al] * ((size_t)h->T256 * 256) : nullptr;
                FLX_TRY(launch_decode_i8(da, h->X2.p + al * x2_stride, NJ, h->NKB, h->stream));
                h->stats.launches++;
            }
        } else {
            FLX_TRY(launch_decode(da, h->stream));
            h->stats.launches++;
        }
        if (!use_resident) FLX_CUDA_TRY(cudaEventRecord(h->ev_used[buf], h->stream));
        mark();  // [2] whiten start
        if (decode_only) {
            if (X_out_dense) {
                // test hook: unpack the design tile on the host
                std::vector<double> xp((size_t)NJ * tile_elems);
                FLX_CUDA_TRY(cudaMemcpyAsync(xp.data(), h->X.p, xp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
                FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
                for (int64_t q = 0; q < nsnp; q++) {
                    const int64_t J = q / ts.snps; const int ql = (int)(q % ts.snps);
                    const int gi = ql / 8, g8 = ql % 8;
                    for (int t = 0; t < h->s; t++) {
                        const int col = 8 * (h->s * gi + t) + g8;
                        double* dst = X_out_dense + ((size_t)(s0 + q) * h->s + t) * h->n;
                        for (int64_t i = 0; i < h->n; i++) dst[i] = xp[(size_t)(J * h->KC + i / KCH) * CHUNK + chunk_off((int)(i % KCH), col)];
                    }
                }
            }
            if (codes_out) {
                FLX_CUDA_TRY(cudaMemcpyAsync(codes_out + (size_t)s0 * h->n, codes_d.p, (size_t)nsnp * h->n, cudaMemcpyDeviceToHost, h->stream));
                FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
            }
            mark(); mark(); mark();
            continue;
        }
        GramArgs ga{};
        ga.KC = h->KC; ga.cr = cr; ga.crp = crp; ga.s = h->s; ga.nsnp = nsnp;
        // sample splits: a function of n only, so a SNP's result never depends on which other SNPs share its batch
        ga.cps = std::max(64, (h->KC + KS_MAX - 1) / KS_MAX);
        ga.KS = (h->KC + ga.cps - 1) / ga.cps;
        ga.n_part = h->n_part.p; ga.u_part = h->u_part.p; ga.a_part = h->a_part.p; ga.b_part = h->b_part.p;
        if (use_fast) {
            // K2q: the exact forms g' C_a V^-1 C_b g' on the i8 tensor pipe, one launch per pass of the plan
            const int nkb_live = (int)((h->n + 63) / 64);
            const int64_t nparts = (int64_t)h->T256 * h->Q;
            for (size_t pi = 0; pi < h->plan.passes.size(); pi++) {
                const flx_handle::QfPass& ps = h->plan.passes[pi];
                QfArgs q{};
                uint32_t gm_unused;
                q.X2 = h->X2.p + ps.amma * x2_stride; tile_code_map(h->plan.aval[ps.amma], &gm_unused, &q.vt_mma);
                q.T8 = h->T8[ps.mat].p; q.rowscale = h->rowscale[ps.mat].p;
                q.full = (ps.mat / h->plan.NG != ps.mat % h->plan.NG) ? 1 : 0;
                q.ndot = ps.ndot;
                for (int d = 0; d < ps.ndot; d++) {
                    q.Xd2[d] = h->X2.p + ps.adot[d] * x2_stride; tile_code_map(h->plan.aval[ps.adot[d]], &gm_unused, &q.vt_dot[d]);
                    q.part[d] = h->qf_part.p + (int64_t)ps.slot[d] * nparts * nsnp;
                }
                for (int i = 0; i < 8; i++) q.plane_scale[i] = std::ldexp(1.0, -8 * (i + 1));
                q.nsnp = nsnp; q.n = h->n; q.T256 = h->T256; q.NJ = NJ; q.Q = h->Q; q.NKB = h->NKB;
                q.NKB_live = nkb_live;
                q.n_items = (int64_t)h->T256 * h->Q * ((NJ + 1) / 2);
                {
                    // L2 grouping of the tile pairs (qf_item) is OFF by default: measured (tools/qf_group_sweep.py, profiles/r02_qf_group_sweep.txt)
                    // it buys nothing at n = 5,000 (K2q is not HBM bound: 13 % DRAM utilisation) and costs 25 % at n = 20,000, where groups
                    // small enough for L2 leave too few CTAs sharing a digit-plane block.  FLX_QF_GROUP_PAIRS=k keeps the experiment reachable.
                    const char* env = getenv("FLX_QF_GROUP_PAIRS");
                    const int all = (NJ + 1) / 2;
                    q.GP = (env && atoi(env) > 0) ? std::min(atoi(env), all) : all;
                }
                DevBuf<long long> prof_d;
                const bool prof = h->cfg.verbose >= 2 && bi == 0 && pi == 0;
                if (prof) { FLX_TRY(prof_d.ensure(16 * (size_t)h->num_sms)); FLX_CUDA_TRY(cudaMemsetAsync(prof_d.p, 0, 16 * (size_t)h->num_sms * sizeof(long long), h->stream)); q.prof = prof_d.p; }
                FLX_TRY(launch_qf(q, h->num_sms, h->stream));
                if (prof) {
                    std::vector<long long> pv(16 * (size_t)h->num_sms);
                    FLX_CUDA_TRY(cudaMemcpy(pv.data(), prof_d.p, pv.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                    double a[16] = {};
                    for (int i = 0; i < h->num_sms; i++) for (int j = 0; j < 16; j++) a[j] += (double)pv[16 * i + j] / h->num_sms;
                    const double items = (double)q.n_items / h->num_sms;
                    fprintf(stderr, "[flx] K2q MMA issuer: %.0f cycles/CTA, %.1f%% waiting for the epilogue, %.1f%% waiting for TMA (%.1f%% in the first ring of an item; waits > 200 clk: %.1f%%, %.1f per item), %.1f%% issuing, items/CTA %.1f\n",
                            a[0], 100 * a[1] / a[0], 100 * a[2] / a[0], 100 * a[3] / a[0], 100 * a[4] / a[0], a[5] / items, 100 * a[6] / a[0], items);
                    fprintf(stderr, "[flx] K2q producer: %.1f%% waiting for a free stage (waits > 200 clk: %.1f%%, %.1f per item); epilogue warp: %.0f clk per item draining, %.0f waiting\n",
                            100 * a[9] / a[8], 100 * a[10] / a[8], a[11] / items, a[13] / items, a[12] / items);
                }
                h->stats.launches += 1; h->stats.launches_whiten++;
            }
            mark();  // [3] fit start
            // K4q per term: u_t = (C_t L^-T Qc)' g_t, b_t = (C_t L^-T r)' g_t and, for every seed, b*_t = (C_t L^-T r*)' g_t in one
            // launch against the digit planes of C_t [H | L^-T R_f]
            for (int t = 0; t < h->s; t++) {
                XgArgs xa{};
                uint32_t gm_unused;
                xa.X2 = h->X2.p + h->plan.alp[t] * x2_stride; tile_code_map(h->plan.aval[h->plan.alp[t]], &gm_unused, &xa.vt);
                xa.G8 = h->G8[h->plan.grp[t]].p; xa.colscale = h->G_scale[h->plan.grp[t]].p;
                for (int i = 0; i < 3; i++) xa.pair_scale[i] = std::ldexp(1.0, -16 * (i + 1));
                xa.n_items = (int64_t)h->ntilesG * ((NJ + 1) / 2); xa.NCT = h->ntilesG; xa.nsnp = nsnp; xa.NJ = NJ; xa.NB = h->NB_G; xa.NC = cr + 1 + (do_perm ? h->P : 0);
                xa.Q = h->Q; xa.NKB = h->NKB; xa.NKB_live = nkb_live;
                xa.out = h->ub.p + (int64_t)t * nsnp * ldg; xa.ld_out = ldg;
                FLX_TRY(launch_xg(xa, h->num_sms, h->stream));
                h->stats.launches += 1;
            }
            ga.KS = 1;
        } else {
            GemmArgs g{};
            g.A = h->linv.p; g.B = h->X.p; g.Out = h->W.p;
            g.T = h->T; g.NJ = NJ; g.KC = h->KC; g.n_items = (int64_t)h->T * NJ;
            g.KC_live = (int)((h->n + KCH - 1) / KCH); g.last_row_blocks = (int)((h->n - (int64_t)(h->T - 1) * TJ + 7) / 8);
            FLX_TRY(launch_whiten(g, h->num_sms, h->stream));
            h->stats.launches++; h->stats.launches_whiten++;
            mark();  // [3] fit start
            ga.W = h->W.p; ga.Qc = h->Qc.p; ga.r = h->r.p;
            FLX_TRY(launch_snp_gram(ga, h->stream));
            h->stats.launches += 2;
        }
        if (do_perm) {
            FLX_CUDA_TRY(cudaMemsetAsync(h->snp_df.p, 0, (size_t)snps_per_batch * sizeof(int32_t), h->stream));
            FLX_CUDA_TRY(cudaMemsetAsync(h->snp_sm.p, 0, (size_t)snps_per_batch * nsym * sizeof(double), h->stream));
        }
        SolveArgs sa{};
        sa.n_part = h->n_part.p; sa.u_part = use_fast ? h->ub.p : h->u_part.p; sa.a_part = h->a_part.p; sa.b_part = h->b_part.p; sa.KS = ga.KS; sa.crp = crp;
        sa.u_term_stride = nsnp * ldg; sa.u_ld = ldg;
        sa.nsnp = nsnp; sa.out_offset = s0;
        sa.out_map = out_map_host ? h->out_map_d.p + s0 : nullptr;
        sa.chisq = h->chisq.p; sa.pval = h->pval.p; sa.df = h->df.p; sa.rank = h->rank.p; sa.pivot = h->pivot.p; sa.coef = h->coef.p;
        sa.snp_df = do_perm ? h->snp_df.p : nullptr; sa.snp_sm = do_perm ? h->snp_sm.p : nullptr;
        sa.df_count = h->df_count.p;
        sa.refit_thr = std::ldexp(1e-5, 8 * (6 - h->Q));   // the digit planes resolve 2^-8Q of x'V^-1x: keep ~9 digits of A
        sa.fast = use_fast ? 1 : 0; sa.qf_part = h->qf_part.p; sa.qf_nparts = h->T256 * h->Q;
        memcpy(sa.slot_a, h->plan.slot_a, sizeof(sa.slot_a)); memcpy(sa.slot_b, h->plan.slot_b, sizeof(sa.slot_b));
        sa.refit_flag = use_fast ? h->refit_flag.p + s0 : nullptr;
        FLX_TRY(launch_snp_solve(sa, h->fc, h->stream));
        h->stats.launches += 1;
        if (want_results && !out_map_host) {
            // this batch's per-SNP results go to pinned staging while the permutation pass and the next batch run
            FLX_CUDA_TRY(cudaEventRecord(h->ev_res, h->stream));
            FLX_CUDA_TRY(cudaStreamWaitEvent(h->d2h_stream, h->ev_res, 0));
            const size_t k = (size_t)h->k;
            FLX_CUDA_TRY(cudaMemcpyAsync(h->chisq_h.p + s0, h->chisq.p + s0, nsnp * sizeof(double), cudaMemcpyDeviceToHost, h->d2h_stream));
            FLX_CUDA_TRY(cudaMemcpyAsync(h->pval_h.p + s0, h->pval.p + s0, nsnp * sizeof(double), cudaMemcpyDeviceToHost, h->d2h_stream));
            FLX_CUDA_TRY(cudaMemcpyAsync(h->df_h.p + s0, h->df.p + s0, nsnp * sizeof(int32_t), cudaMemcpyDeviceToHost, h->d2h_stream));
            FLX_CUDA_TRY(cudaMemcpyAsync(h->rank_h.p + s0, h->rank.p + s0, nsnp * sizeof(int32_t), cudaMemcpyDeviceToHost, h->d2h_stream));
            FLX_CUDA_TRY(cudaMemcpyAsync(h->pivot_h.p + s0 * k, h->pivot.p + s0 * k, nsnp * k * sizeof(int32_t), cudaMemcpyDeviceToHost, h->d2h_stream));
            FLX_CUDA_TRY(cudaMemcpyAsync(h->coef_h.p + s0 * k, h->coef.p + s0 * k, nsnp * k * sizeof(double), cudaMemcpyDeviceToHost, h->d2h_stream));
        }
        mark();  // [4] perm start
        if (do_perm && use_fast) {
            // b*' Sm b* / RSS_f* per (SNP, seed) from the stored b*_t, running max per (seed, df)
            PermReduceArgs pr{};
            pr.bst = h->ub.p + (cr + 1); pr.term_stride = nsnp * ldg; pr.ldp = ldg;
            pr.snp_df = h->snp_df.p; pr.snp_sm = h->snp_sm.p; pr.perm_inv_rss = h->inv_rss.p; pr.perm_maxratio = h->maxratio.p;
            pr.nsnp = nsnp; pr.P = h->P; pr.P_pad = h->P_pad; pr.s = h->s;
            FLX_TRY(launch_perm_reduce(pr, h->stream));
            h->stats.launches++;
        } else if (do_perm) {
            GemmArgs gp{};
            gp.A = h->W.p; gp.B = h->Rpk.p; gp.NJ = NJ; gp.NP = h->NP; gp.KC = h->KC; gp.n_items = (int64_t)NJ * h->NP;
            gp.snp_df = h->snp_df.p; gp.snp_sm = h->snp_sm.p; gp.perm_inv_rss = h->inv_rss.p; gp.perm_maxratio = h->maxratio.p; gp.P_pad = h->P_pad;
            FLX_TRY(launch_perm_contract(gp, h->s, h->num_sms, h->stream));
            h->stats.launches++;
        }
        mark();  // [5] batch end
    }
    mark();
    FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    // accumulate per-stage device times
    for (int64_t bi = 0; bi < nbatch; bi++) {
        float ms;
        cudaEvent_t* e = &evs[(size_t)bi * 6];
        cudaEventElapsedTime(&ms, e[0], e[1]); h->stats.ms_h2d += ms;
        cudaEventElapsedTime(&ms, e[1], e[2]); h->stats.ms_decode += ms;
        cudaEventElapsedTime(&ms, e[2], e[3]); h->stats.ms_whiten += ms;
        cudaEventElapsedTime(&ms, e[3], e[4]); h->stats.ms_fit += ms;
        cudaEventElapsedTime(&ms, e[4], e[5]); h->stats.ms_perm += ms;
    }
    if (nev > 0) { float ms; cudaEventElapsedTime(&ms, evs[0], evs[nev - 1]); h->stats.ms_batches += ms; }
    int pxe = 0;
    FLX_CUDA_TRY(cudaMemcpy(&pxe, h->px_err.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (pxe != 0) {
        static const char* what[] = {"", "short raw record", "short 1-bit record", "bad difflist length", "truncated difflist", "bad difflist entry"};
        set_error("pgen: malformed variant record (%s; record %d of its batch)", what[std::min(pxe & 15, 5)], pxe >> 4);
        return FLX_ERR_PGEN;
    }
    int miss = INT_MAX;
    FLX_CUDA_TRY(cudaMemcpy(&miss, h->missing.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (miss != INT_MAX && !decode_only) {
        set_error("missing genotype (NA hard call) in tested variant var_idx[%d]=%d: the reference aborts on it (no imputation anywhere in the pipeline)", miss, var_idx[miss]);
        return FLX_ERR_MISSING_GENOTYPE;
    }
    return FLX_OK;
}

extern "C" int flx_prepare(flx_handle* h) {
    if (!h) { set_error("flx_prepare: NULL handle"); return FLX_ERR_BAD_ARG; }
    FLX_CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (h->dirty) FLX_TRY(finalize_setup(h));
    return FLX_OK;
}

// The rank-local part of flx_run: batches, FP64 refits, min-p finalisation, per-SNP results to the caller's arrays.
static int run_local(flx_handle* h, const int32_t* var_idx, int64_t M, const flx_snp_out* out, bool do_perm) {
    // pgenlibr::Read instead of ReadHardcalls (fit_model.nf:26): only a .pgen can carry dosages; real-valued x runs on the FP64 route
    const bool dosage = h->use_dosage && h->src == 1 && h->pgen && h->pgen->mode() == 0x10 && h->pgen->has_dosage();
    if (dosage && !h->dos_ok) { set_error("flx_run: %s", h->dos_why.c_str()); return h->have_probes ? FLX_ERR_BAD_ARG : FLX_ERR_STATE; }
    h->run_dosage = dosage;
    if (dosage) h->stats.fallback_reason = FLX_FALLBACK_DOSAGE;
    if (M > 0) {
        const bool fast = h->fast && !dosage;
        h->stats.int8_path = fast ? 1 : 0;
        h->stats.int8_passes = 0;      // in units of one triangular pass (6 n^2 int8 ops per SNP); a full-matrix pass counts two
        if (fast) for (const auto& ps : h->plan.passes) h->stats.int8_passes += (ps.mat / h->plan.NG != ps.mat % h->plan.NG) ? 2 : 1;
        FLX_TRY(run_batches(h, var_idx, M, out != nullptr, do_perm, nullptr, nullptr, fast, false, nullptr, M));
        if (fast) {
            // SNPs whose x is (nearly) in the span of the constant columns lose digits in A = x'V^-1x - |u|^2: re-fit them in FP64.
            // The flags are compacted on the device; the host reads a count and, only if it is not zero, the short list.
            int nref = 0;
            FLX_TRY(h->refit_list.ensure((size_t)M + 1));
            FLX_TRY(launch_compact_flags(h->refit_flag.p, M, h->refit_list.p, h->stream));
            h->stats.launches++;
            FLX_CUDA_TRY(cudaMemcpyAsync(&nref, h->refit_list.p, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
            h->stats.refit_snps = nref;
            if (nref > 0) {
                std::vector<int32_t> slot(nref), sub(nref);
                FLX_CUDA_TRY(cudaMemcpy(slot.data(), h->refit_list.p + 1, (size_t)nref * sizeof(int32_t), cudaMemcpyDeviceToHost));
                std::sort(slot.begin(), slot.end());      // atomics hand out list positions in any order: keep the run deterministic
                for (int i = 0; i < nref; i++) sub[i] = var_idx[slot[i]];
                FLX_TRY(run_batches(h, sub.data(), nref, out != nullptr, do_perm, nullptr, nullptr, false, true, slot.data(), M));
            }
        }
        if (do_perm) {
            FLX_TRY(launch_minp_finalize(h->maxratio.p, h->P_pad, h->P, h->s, h->cr + 1 - h->df_null, h->lrs0p.p, (double)h->n, h->df_count.p, h->minp_d.p, h->stream));
            h->stats.launches++;
        }
    }
    if (out && M > 0) {
        const size_t k = (size_t)h->k;
        FLX_CUDA_TRY(cudaStreamSynchronize(h->d2h_stream));       // the per-batch copies have landed (and cannot overtake what follows)
        if (h->stats.refit_snps > 0) {
            // the FP64 refit pass rewrote scattered slots: fetch the arrays again
            FLX_CUDA_TRY(cudaMemcpyAsync(h->chisq_h.p, h->chisq.p, M * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            FLX_CUDA_TRY(cudaMemcpyAsync(h->pval_h.p, h->pval.p, M * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            FLX_CUDA_TRY(cudaMemcpyAsync(h->df_h.p, h->df.p, M * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            FLX_CUDA_TRY(cudaMemcpyAsync(h->rank_h.p, h->rank.p, M * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            FLX_CUDA_TRY(cudaMemcpyAsync(h->pivot_h.p, h->pivot.p, M * k * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            FLX_CUDA_TRY(cudaMemcpyAsync(h->coef_h.p, h->coef.p, M * k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
        }
        // pinned staging -> the caller's arrays, a slice per host thread
        struct Cp { void* dst; const void* src; size_t bytes; };
        const Cp cps[6] = {{out->lrt_chisq, h->chisq_h.p, M * sizeof(double)}, {out->pval, h->pval_h.p, M * sizeof(double)},
                           {out->lrt_df, h->df_h.p, M * sizeof(int32_t)},    {out->rank, h->rank_h.p, M * sizeof(int32_t)},
                           {out->pivot, h->pivot_h.p, M * k * sizeof(int32_t)}, {out->coef, h->coef_h.p, M * k * sizeof(double)}};
        const unsigned nt = (unsigned)std::max<int64_t>(1, std::min<int64_t>(std::min(std::thread::hardware_concurrency(), 8u), M / 65536));
        auto work = [&](unsigned ti) {
            for (const Cp& c : cps) {
                if (!c.dst) continue;
                const size_t lo = c.bytes * ti / nt, hi = c.bytes * (ti + 1) / nt;
                memcpy((char*)c.dst + lo, (const char*)c.src + lo, hi - lo);
            }
        };
        if (nt == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (unsigned ti = 0; ti < nt; ti++) th.emplace_back(work, ti);
            for (auto& x : th) x.join();
        }
        for (const Cp& c : cps) if (c.dst) h->stats.d2h_bytes += (int64_t)c.bytes;
    }
    if (do_perm && M == 0) {
        std::vector<double> two(h->P, 2.0);
        FLX_CUDA_TRY(cudaMemcpyAsync(h->minp_d.p, two.data(), h->P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        FLX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return FLX_OK;
}

extern "C" int flx_run(flx_handle* h, const int32_t* var_idx, int64_t M, const flx_snp_out* out, double* minp, int64_t* nvars_tested) {
    if (!h) { set_error("flx_run: NULL handle"); return FLX_ERR_BAD_ARG; }
    const auto wall0 = std::chrono::steady_clock::now();
    // Every rank of an attached communicator must reach the collectives below whatever happens locally (a missing genotype
    // or a malformed record in ONE shard must not leave the other ranks blocked in ncclAllReduce): local failures are
    // carried in `rc` to the status exchange instead of returning early.
    int rc = FLX_OK;
    bool do_perm = false;
    if ((M > 0 && !var_idx) || M < 0) { set_error("flx_run: bad arguments"); rc = FLX_ERR_BAD_ARG; }
    else if (M >= INT_MAX) { set_error("flx_run: M too large"); rc = FLX_ERR_BAD_ARG; }
    else if (cudaSetDevice(h->cfg.device) != cudaSuccess) { set_error("flx_run: cudaSetDevice(%d) failed", h->cfg.device); rc = FLX_ERR_CUDA; }
    else if (h->src == 0) { set_error("flx_run: no genotype source (flx_open_pgen / flx_set_packed)"); rc = FLX_ERR_STATE; }
    else if ((int64_t)h->order0.size() != h->n) { set_error("flx_run: the genotype source was set for n=%lld samples, the whitening matrix has n=%lld", (long long)h->order0.size(), (long long)h->n); rc = FLX_ERR_STATE; }
    if (rc == FLX_OK && h->dirty) rc = finalize_setup(h);
    memset(&h->stats, 0, sizeof(h->stats));
    h->stats.fallback_reason = h->fallback_reason;
    if (rc == FLX_OK) {
        do_perm = h->have_perm && h->P > 0;
        if (minp && !do_perm) { set_error("flx_run: minp requested but no permutations were set (flx_set_perm)"); rc = FLX_ERR_STATE; }
    }
    h->stats.snps = M; h->stats.n_pad = h->n_pad; h->stats.cols_per_tile = make_tile_shape(std::max(h->s, 1)).blocks * 8;
    cudaEventRecord(h->ev_t0, h->stream);
    if (rc == FLX_OK) rc = run_local(h, var_idx, M, out, do_perm);
    int64_t nv = M;
    if (h->comm && h->nranks > 1) {
        // the only collective of the design: min over SNP shards of the length-P vector (get_null_p_dist.nf:77), with the sum
        // of the tested-variant counts and the worst status of any rank riding along
        long long word[2] = {rc == FLX_OK ? (long long)M : 0, (long long)rc};
        ncclResult_t r0 = ncclSuccess, r1 = ncclSuccess, r2 = ncclSuccess;
        bool ok = h->coll_d.ensure(2) == 0 &&
                  cudaMemcpyAsync(h->coll_d.p, word, sizeof(word), cudaMemcpyHostToDevice, h->stream) == cudaSuccess;
        if (ok) {
            ncclGroupStart();
            if (h->P > 0 && h->minp_d.p) r0 = ncclAllReduce(h->minp_d.p, h->minp_d.p, (size_t)h->P, ncclDouble, ncclMin, h->comm, h->stream);
            r1 = ncclAllReduce(h->coll_d.p, h->coll_d.p, 1, ncclInt64, ncclSum, h->comm, h->stream);
            r2 = ncclAllReduce(h->coll_d.p + 1, h->coll_d.p + 1, 1, ncclInt64, ncclMin, h->comm, h->stream);
            ncclGroupEnd();
            ok = r0 == ncclSuccess && r1 == ncclSuccess && r2 == ncclSuccess &&
                 cudaMemcpyAsync(word, h->coll_d.p, sizeof(word), cudaMemcpyDeviceToHost, h->stream) == cudaSuccess &&
                 cudaStreamSynchronize(h->stream) == cudaSuccess;
        }
        if (!ok) {
            if (rc == FLX_OK) { set_error("flx_run: the min-p allreduce failed (%s)", ncclGetErrorString(r0 != ncclSuccess ? r0 : r1 != ncclSuccess ? r1 : r2)); rc = FLX_ERR_NCCL; }
        } else {
            nv = word[0];
            if (rc == FLX_OK && word[1] != FLX_OK) {
                set_error("flx_run: another rank of the communicator failed with status %lld; this shard's results are discarded with it", word[1]);
                rc = (int)word[1];
            }
        }
    }
    if (rc == FLX_OK && do_perm && minp) {
        if (cudaMemcpyAsync(minp, h->minp_d.p, h->P * sizeof(double), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) { set_error("flx_run: min-p copy failed"); rc = FLX_ERR_CUDA; }
        h->stats.d2h_bytes += h->P * (int64_t)sizeof(double);
    }
    cudaEventRecord(h->ev_t1, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess && rc == FLX_OK) { set_error("flx_run: %s", cudaGetErrorString(cudaGetLastError())); rc = FLX_ERR_CUDA; }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1) == cudaSuccess) h->stats.ms_total = ms;
    h->stats.ms_d2h = std::max(0.0, h->stats.ms_total - h->stats.ms_batches);
    h->stats.ms_wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    if (rc != FLX_OK) return rc;
    if (nvars_tested) *nvars_tested = nv;
    return FLX_OK;
}

extern "C" int flx_get_stats(flx_handle* h, flx_stats* out) {
    if (!h || !out) { set_error("flx_get_stats: bad arguments"); return FLX_ERR_BAD_ARG; }
    *out = h->stats;
    return FLX_OK;
}

extern "C" int flx_perm_indices(int32_t seed, int64_t m, int32_t* out) {
    if (!out || m < 0) { set_error("flx_perm_indices: bad arguments"); return FLX_ERR_BAD_ARG; }
    r_sample_int(seed, m, out);
    return FLX_OK;
}

extern "C" int flx_decode_tile(flx_handle* h, const int32_t* var_idx, int64_t M, double* X, uint8_t* codes) {
    if (!h || !var_idx || M <= 0 || (!X && !codes)) { set_error("flx_decode_tile: bad arguments"); return FLX_ERR_BAD_ARG; }
    FLX_CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (h->src == 0) { set_error("flx_decode_tile: no genotype source"); return FLX_ERR_STATE; }
    if (h->dirty) FLX_TRY(finalize_setup(h));
    if (codes) memset(codes, 3, (size_t)M * h->n);
    h->run_dosage = h->use_dosage && h->src == 1 && h->pgen && h->pgen->mode() == 0x10 && h->pgen->has_dosage();
    if (h->run_dosage && !h->dos_ok) { set_error("flx_decode_tile: %s", h->dos_why.c_str()); return FLX_ERR_STATE; }
    return run_batches(h, var_idx, M, false, false, X, codes);
}

extern "C" int flx_comm_unique_id(uint8_t id[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId u;
    ncclResult_t r = ncclGetUniqueId(&u);
    if (r != ncclSuccess) { set_error("ncclGetUniqueId: %s", ncclGetErrorString(r)); return FLX_ERR_NCCL; }
    memcpy(id, &u, 128);
    return FLX_OK;
}

extern "C" int flx_comm_attach(flx_handle* h, const uint8_t id[128], int32_t rank, int32_t nranks) {
    if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) { set_error("flx_comm_attach: bad arguments"); return FLX_ERR_BAD_ARG; }
    FLX_CUDA_TRY(cudaSetDevice(h->cfg.device));
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclResult_t r = ncclCommInitRank(&h->comm, nranks, u, rank);
    if (r != ncclSuccess) { set_error("ncclCommInitRank: %s", ncclGetErrorString(r)); h->comm = nullptr; return FLX_ERR_NCCL; }
    h->nranks = nranks;
    h->comm_rank = rank;
    return FLX_OK;
}


// ================================================================================================
// One caller, several GPUs (R is single-threaded, one process drives the box).  A flx_multi owns one handle per
// device plus the NCCL communicator between them; every collective step runs on one host thread per device.
// ================================================================================================
struct flx_multi {
    std::vector<flx_handle*> hs;
    const uint8_t* pinned = nullptr;
    ~flx_multi() {
        if (pinned) cudaHostUnregister(const_cast<uint8_t*>(pinned));
        for (auto h : hs) flx_destroy(h);
    }
};

// f(handle, device index) on one thread per device; first failure (lowest index) wins, its message becomes the caller's
template <typename F>
static int multi_each(flx_multi* m, F f) {
    const size_t nd = m->hs.size();
    std::vector<int> rc(nd, 0);
    std::vector<std::string> msg(nd);
    auto body = [&](size_t i) { rc[i] = f(m->hs[i], (int)i); if (rc[i] != 0) msg[i] = g_err; };
    if (nd == 1) body(0);
    else {
        std::vector<std::thread> th;
        for (size_t i = 0; i < nd; i++) th.emplace_back(body, i);
        for (auto& t : th) t.join();
    }
    for (size_t i = 0; i < nd; i++)
        if (rc[i] != 0) { set_error("device %d: %s", m->hs[i]->cfg.device, msg[i].c_str()); return rc[i]; }
    return FLX_OK;
}

extern "C" int flx_multi_create(const flx_config* cfg, const int32_t* devices, int32_t ndev, flx_multi** out) {
    if (!out || !devices || ndev < 1 || ndev > 64) { set_error("flx_multi_create: bad arguments"); return FLX_ERR_BAD_ARG; }
    *out = nullptr;
    for (int i = 0; i < ndev; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) { set_error("flx_multi_create: device %d listed twice", devices[i]); return FLX_ERR_BAD_ARG; }
    flx_multi* m = new flx_multi();
    for (int i = 0; i < ndev; i++) {
        flx_config c{};
        if (cfg) c = *cfg;
        c.device = devices[i];
        flx_handle* h = nullptr;
        const int rc = flx_create(&c, &h);
        if (rc != FLX_OK) { delete m; return rc; }
        m->hs.push_back(h);
    }
    if (ndev > 1) {
        std::vector<ncclComm_t> comms(ndev);
        std::vector<int> devs(devices, devices + ndev);
        ncclResult_t r = ncclCommInitAll(comms.data(), ndev, devs.data());
        if (r != ncclSuccess) { set_error("ncclCommInitAll: %s", ncclGetErrorString(r)); delete m; return FLX_ERR_NCCL; }
        for (int i = 0; i < ndev; i++) { m->hs[i]->comm = comms[i]; m->hs[i]->nranks = ndev; m->hs[i]->comm_rank = i; }
    }
    *out = m;
    return FLX_OK;
}

extern "C" void flx_multi_destroy(flx_multi* m) { delete m; }
extern "C" int flx_multi_device_count(flx_multi* m) { return m ? (int)m->hs.size() : 0; }

extern "C" int flx_multi_set_whitening(flx_multi* m, const double* L, int64_t n, int64_t ld, int is_inverse) {
    if (!m || !L) { set_error("flx_multi_set_whitening: bad arguments"); return FLX_ERR_BAD_ARG; }
    // device 0 uploads and inverts; L^-1 reaches the others by ncclBroadcast over NVLink
    return multi_each(m, [&](flx_handle* h, int i) { return flx_set_whitening_bcast(h, i == 0 ? L : nullptr, n, ld, is_inverse, 0); });
}
extern "C" int flx_multi_set_null(flx_multi* m, const double* X_null, int32_t c, const double* y_mm, double ll_null, int32_t df_null) {
    if (!m) { set_error("flx_multi_set_null: NULL handle"); return FLX_ERR_BAD_ARG; }
    return multi_each(m, [&](flx_handle* h, int) { return flx_set_null(h, X_null, c, y_mm, ll_null, df_null); });
}
extern "C" int flx_multi_set_design(flx_multi* m, const double* M0, const double* M1, const double* M2, int32_t k) {
    if (!m) { set_error("flx_multi_set_design: NULL handle"); return FLX_ERR_BAD_ARG; }
    return multi_each(m, [&](flx_handle* h, int) { return flx_set_design(h, M0, M1, M2, k); });
}
extern "C" int flx_multi_set_design_probes(flx_multi* m, const double* M05, const double* M15) {
    if (!m) { set_error("flx_multi_set_design_probes: NULL handle"); return FLX_ERR_BAD_ARG; }
    return multi_each(m, [&](flx_handle* h, int) { return flx_set_design_probes(h, M05, M15); });
}
extern "C" int flx_multi_set_use_dosage(flx_multi* m, int use_dosage) {
    if (!m) { set_error("flx_multi_set_use_dosage: NULL handle"); return FLX_ERR_BAD_ARG; }
    for (auto h : m->hs) FLX_TRY(flx_set_use_dosage(h, use_dosage));
    return FLX_OK;
}
extern "C" int flx_multi_set_perm(flx_multi* m, const double* y_pred, const double* U1, const double* e_p, int64_t mm, const int32_t* seeds, int32_t P) {
    if (!m) { set_error("flx_multi_set_perm: NULL handle"); return FLX_ERR_BAD_ARG; }
    return multi_each(m, [&](flx_handle* h, int i) { return flx_set_perm_bcast(h, y_pred, i == 0 ? U1 : nullptr, e_p, mm, seeds, P, 0); });
}
extern "C" int flx_multi_set_perm_implicit(flx_multi* m, const int32_t* seeds, int32_t P) {
    if (!m) { set_error("flx_multi_set_perm_implicit: NULL handle"); return FLX_ERR_BAD_ARG; }
    return multi_each(m, [&](flx_handle* h, int) { return flx_set_perm_implicit(h, seeds, P); });
}
extern "C" int flx_multi_open_pgen(flx_multi* m, const char* path, const int32_t* pgen_order, int64_t n) {
    if (!m) { set_error("flx_multi_open_pgen: NULL handle"); return FLX_ERR_BAD_ARG; }
    return multi_each(m, [&](flx_handle* h, int) { return flx_open_pgen(h, path, pgen_order, n); });
}
extern "C" int flx_multi_set_packed(flx_multi* m, const uint8_t* packed, int64_t n_variants, int64_t bytes_per_variant, int64_t raw_sample_ct, const int32_t* pgen_order, int64_t n) {
    if (!m || !packed) { set_error("flx_multi_set_packed: bad arguments"); return FLX_ERR_BAD_ARG; }
    if (m->pinned) { cudaHostUnregister(const_cast<uint8_t*>(m->pinned)); m->pinned = nullptr; }
    bool pinned = false;
    if (n_variants * bytes_per_variant >= (1 << 20) && cudaSetDevice(m->hs[0]->cfg.device) == cudaSuccess) {
        if (cudaHostRegister(const_cast<uint8_t*>(packed), (size_t)n_variants * bytes_per_variant, cudaHostRegisterPortable) == cudaSuccess) { pinned = true; m->pinned = packed; }
        else cudaGetLastError();
    }
    for (auto h : m->hs) FLX_TRY(set_packed_impl(h, packed, n_variants, bytes_per_variant, raw_sample_ct, pgen_order, n, pinned));
    return FLX_OK;
}
extern "C" int flx_multi_make_resident(flx_multi* m) {
    if (!m) { set_error("flx_multi_make_resident: NULL handle"); return FLX_ERR_BAD_ARG; }
    return multi_each(m, [&](flx_handle* h, int) { return flx_make_resident(h); });
}
extern "C" int flx_multi_prepare(flx_multi* m) {
    if (!m) { set_error("flx_multi_prepare: NULL handle"); return FLX_ERR_BAD_ARG; }
    return multi_each(m, [&](flx_handle* h, int) { return flx_prepare(h); });
}

extern "C" int flx_multi_run(flx_multi* m, const int32_t* var_idx, int64_t M, const flx_snp_out* out, double* minp, int64_t* nvars_tested) {
    if (!m || (M > 0 && !var_idx) || M < 0) { set_error("flx_multi_run: bad arguments"); return FLX_ERR_BAD_ARG; }
    const int64_t nd = (int64_t)m->hs.size();
    std::vector<int64_t> nv(nd, 0);
    // contiguous SNP blocks, one per device; per-SNP outputs land in disjoint slices of the caller's arrays (var_idx order is kept);
    // inside flx_run the length-P min-p vector is min-allreduced and the counts summed, so every device ends with the box's result
    const int rc = multi_each(m, [&](flx_handle* h, int i) {
        const int64_t base = M / nd, extra = M % nd;
        const int64_t lo = i * base + std::min<int64_t>(i, extra), cnt = base + (i < extra ? 1 : 0);
        flx_snp_out o{};
        const int64_t k = h->k;
        if (out) {
            o.lrt_chisq = out->lrt_chisq ? out->lrt_chisq + lo : nullptr; o.pval = out->pval ? out->pval + lo : nullptr;
            o.lrt_df = out->lrt_df ? out->lrt_df + lo : nullptr;          o.rank = out->rank ? out->rank + lo : nullptr;
            o.pivot = out->pivot ? out->pivot + lo * k : nullptr;         o.coef = out->coef ? out->coef + lo * k : nullptr;
        }
        std::vector<double> tmp;
        double* mp = nullptr;
        if (minp) { if (i == 0) mp = minp; else { tmp.resize((size_t)std::max(h->P, 1)); mp = tmp.data(); } }
        return flx_run(h, var_idx + lo, cnt, out ? &o : nullptr, mp, &nv[i]);
    });
    if (rc != FLX_OK) return rc;
    if (nvars_tested) *nvars_tested = nv[0];
    return FLX_OK;
}

extern "C" int flx_multi_get_stats(flx_multi* m, int32_t device_index, flx_stats* out) {
    if (!m || device_index < 0 || device_index >= (int)m->hs.size()) { set_error("flx_multi_get_stats: bad arguments"); return FLX_ERR_BAD_ARG; }
    return flx_get_stats(m->hs[device_index], out);
}

extern "C" double flx_minp_threshold(const double* min_pval, int64_t P, double p_thr) {
    if (!min_pval || P <= 0) return NAN;
    std::vector<double> x(min_pval, min_pval + P);
    std::sort(x.begin(), x.end());
    const double hh = (double)(P - 1) * p_thr;
    int64_t lo = (int64_t)std::floor(hh);
    if (lo < 0) lo = 0;
    if (lo > P - 1) lo = P - 1;
    const int64_t hi = std::min(lo + 1, P - 1);
    return x[lo] + (hh - (double)lo) * (x[hi] - x[lo]);
}
