// host_math.cpp — host-side pieces of the hot path (see host_math.h).
#include "host_math.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstring>

namespace flx {

// ------------------------------------------------------------------------------------------------
// R's default RNG stream (RNG.c): set.seed() scrambling, MT19937, unif_rand fix-up, R_unif_index by
// rejection on ceil(log2(n)) random bits, do_sample without replacement by swap-removal.
// fit_model.nf:97-98; workflows/flexlmm.nf:67.
// ------------------------------------------------------------------------------------------------
namespace {
class RMersenne {
public:
    explicit RMersenne(int32_t seed) {
        uint32_t s = static_cast<uint32_t>(seed);
        for (int i = 0; i < 50; ++i) s = 69069u * s + 1u;
        s = 69069u * s + 1u;  // this word lands in dummy[0] (= mti) and is overwritten with 624
        for (auto& w : state_) { s = 69069u * s + 1u; w = s; }
        pos_ = kN;
    }
    double unif() {
        const double v = word() * 2.3283064365386963e-10;
        constexpr double tiny = 2.328306437080797e-10;
        if (v <= 0.0) return 0.5 * tiny;
        if (1.0 - v <= 0.0) return 1.0 - 0.5 * tiny;
        return v;
    }
    // uniform integer in [0, dn)
    int64_t index(double dn) {
        if (dn <= 0) return 0;
        const int bits = static_cast<int>(std::ceil(std::log2(dn)));
        for (;;) {
            int64_t v = 0;
            for (int n = 0; n <= bits; n += 16) v = 65536 * v + static_cast<int>(std::floor(unif() * 65536));
            if (bits < 64) v &= ((int64_t{1} << bits) - 1);
            if (static_cast<double>(v) < dn) return v;
        }
    }

private:
    static constexpr int kN = 624, kM = 397;
    uint32_t word() {
        if (pos_ >= kN) refill();
        uint32_t y = state_[pos_++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    void refill() {
        auto twist = [](uint32_t hi, uint32_t lo) {
            const uint32_t y = (hi & 0x80000000u) | (lo & 0x7fffffffu);
            return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        };
        for (int i = 0; i < kN; ++i) state_[i] = state_[(i + kM) % kN] ^ twist(state_[i], state_[(i + 1) % kN]);
        pos_ = 0;
    }
    uint32_t state_[kN];
    int pos_;
};
}  // namespace

void r_sample_int(int32_t seed, int64_t m, int32_t* out) {
    RMersenne rng(seed);
    std::vector<int32_t> pool(static_cast<size_t>(m));
    for (int64_t i = 0; i < m; ++i) pool[i] = static_cast<int32_t>(i);
    int64_t left = m;
    for (int64_t i = 0; i < m; ++i) {
        const int64_t j = rng.index(static_cast<double>(left));
        out[i] = pool[j] + 1;
        pool[j] = pool[--left];
    }
}

// ------------------------------------------------------------------------------------------------
// Orthonormal basis of the constant design columns, with dqrdc2's acceptance rule so that the rank of
// the SNP-independent part matches what .lm.fit would find.
// Modified Gram-Schmidt with one re-orthogonalisation pass ("twice is enough"), long-double dots.
// ------------------------------------------------------------------------------------------------
static long double ldot(const double* a, const double* b, int64_t n) {
    long double s = 0.0L;
    for (int64_t i = 0; i < n; ++i) s += static_cast<long double>(a[i]) * b[i];
    return s;
}

int orth_basis(const double* C, int64_t n, int nc, std::vector<double>& Q, std::vector<double>& R) {
    std::vector<double> q(static_cast<size_t>(n) * (nc > 0 ? nc : 1));
    std::vector<double> v(static_cast<size_t>(n));
    int r = 0;
    for (int j = 0; j < nc; ++j) {
        const double* cj = C + static_cast<int64_t>(j) * n;
        const double orig = std::sqrt(static_cast<double>(ldot(cj, cj, n)));
        std::memcpy(v.data(), cj, sizeof(double) * n);
        for (int pass = 0; pass < 2; ++pass)
            for (int i = 0; i < r; ++i) {
                const double* qi = q.data() + static_cast<int64_t>(i) * n;
                const double t = static_cast<double>(ldot(qi, v.data(), n));
                for (int64_t k = 0; k < n; ++k) v[k] -= t * qi[k];
            }
        const double rem = std::sqrt(static_cast<double>(ldot(v.data(), v.data(), n)));
        const double ref = (orig == 0.0) ? 1.0 : orig;
        if (rem < ref * 1e-7) continue;  // negligible: dqrdc2 would cycle this column to the end
        double* qr = q.data() + static_cast<int64_t>(r) * n;
        for (int64_t k = 0; k < n; ++k) qr[k] = v[k] / rem;
        ++r;
    }
    Q.assign(q.begin(), q.begin() + static_cast<int64_t>(r) * n);
    R.assign(static_cast<size_t>(r > 0 ? r : 1) * (nc > 0 ? nc : 1), 0.0);
    for (int j = 0; j < nc; ++j)
        for (int i = 0; i < r; ++i)
            R[i + static_cast<int64_t>(j) * r] = static_cast<double>(ldot(Q.data() + static_cast<int64_t>(i) * n, C + static_cast<int64_t>(j) * n, n));
    return r;
}

// ------------------------------------------------------------------------------------------------
// Null fit with implicit complement basis
// ------------------------------------------------------------------------------------------------
void NullFit::apply_Q(double* z) const {
    for (int j = rank - 1; j >= 0; --j) {
        const double* v = V.data() + static_cast<int64_t>(j) * n;
        const long double t = ldot(v + j, z + j, n - j) * beta[j];
        for (int64_t i = j; i < n; ++i) z[i] -= static_cast<double>(t) * v[i];
    }
}
void NullFit::apply_Qt(double* z) const {
    for (int j = 0; j < rank; ++j) {
        const double* v = V.data() + static_cast<int64_t>(j) * n;
        const long double t = ldot(v + j, z + j, n - j) * beta[j];
        for (int64_t i = j; i < n; ++i) z[i] -= static_cast<double>(t) * v[i];
    }
}

void null_fit_householder(const double* X, int64_t n, int c, const double* y, NullFit& out) {
    out.n = n;
    // accepted columns by the in-order rank rule (same rule as orth_basis)
    std::vector<double> Qtmp, Rtmp;
    std::vector<int> keep;
    {
        std::vector<double> q(static_cast<size_t>(n) * (c > 0 ? c : 1)), v(static_cast<size_t>(n));
        int r = 0;
        for (int j = 0; j < c; ++j) {
            const double* cj = X + static_cast<int64_t>(j) * n;
            const double orig = std::sqrt(static_cast<double>(ldot(cj, cj, n)));
            std::memcpy(v.data(), cj, sizeof(double) * n);
            for (int pass = 0; pass < 2; ++pass)
                for (int i = 0; i < r; ++i) {
                    const double* qi = q.data() + static_cast<int64_t>(i) * n;
                    const double t = static_cast<double>(ldot(qi, v.data(), n));
                    for (int64_t k = 0; k < n; ++k) v[k] -= t * qi[k];
                }
            const double rem = std::sqrt(static_cast<double>(ldot(v.data(), v.data(), n)));
            if (rem < ((orig == 0.0) ? 1.0 : orig) * 1e-7) continue;
            double* qr = q.data() + static_cast<int64_t>(r) * n;
            for (int64_t k = 0; k < n; ++k) qr[k] = v[k] / rem;
            keep.push_back(j);
            ++r;
        }
    }
    const int r = static_cast<int>(keep.size());
    out.rank = r;
    out.V.assign(static_cast<size_t>(n) * (r > 0 ? r : 1), 0.0);
    out.beta.assign(r > 0 ? r : 1, 0.0);
    // Householder QR of the accepted columns
    std::vector<double> A(static_cast<size_t>(n) * (r > 0 ? r : 1));
    for (int j = 0; j < r; ++j) std::memcpy(A.data() + static_cast<int64_t>(j) * n, X + static_cast<int64_t>(keep[j]) * n, sizeof(double) * n);
    for (int j = 0; j < r; ++j) {
        double* a = A.data() + static_cast<int64_t>(j) * n;
        const double alpha = a[j];
        const double nrm = std::sqrt(static_cast<double>(ldot(a + j, a + j, n - j)));
        double* v = out.V.data() + static_cast<int64_t>(j) * n;
        if (nrm == 0.0) { out.beta[j] = 0.0; continue; }
        const double s = (alpha >= 0.0) ? -nrm : nrm;       // a_j -> s e_j
        for (int64_t i = j; i < n; ++i) v[i] = a[i];
        v[j] = alpha - s;
        const double vnorm2 = static_cast<double>(ldot(v + j, v + j, n - j));
        out.beta[j] = (vnorm2 > 0.0) ? 2.0 / vnorm2 : 0.0;
        for (int k = j; k < r; ++k) {
            double* ak = A.data() + static_cast<int64_t>(k) * n;
            const double t = static_cast<double>(ldot(v + j, ak + j, n - j)) * out.beta[j];
            for (int64_t i = j; i < n; ++i) ak[i] -= t * v[i];
        }
    }
    // lm.fit: Q'y, fitted = Q [Q'y](1:r; 0), residuals = Q (0; [Q'y](r+1:n))
    std::vector<double> qty(y, y + n);
    out.apply_Qt(qty.data());
    out.e_p.assign(qty.begin() + r, qty.end());
    out.fitted.assign(static_cast<size_t>(n), 0.0);
    for (int i = 0; i < r; ++i) out.fitted[i] = qty[i];
    out.apply_Q(out.fitted.data());
    out.resid.resize(static_cast<size_t>(n));
    long double rss = 0.0L;
    for (int64_t i = 0; i < n; ++i) { out.resid[i] = y[i] - out.fitted[i]; rss += static_cast<long double>(out.resid[i]) * out.resid[i]; }
    const double N = static_cast<double>(n);
    out.ll = 0.5 * (0.0 - N * (std::log(2.0 * 3.14159265358979323846) + 1.0 - std::log(N) + std::log(static_cast<double>(rss))));
}

// ------------------------------------------------------------------------------------------------
// .pgen (plink-ng pgenlib on-disk format): header and record index only — the records themselves are expanded on the
// device (pgen_expand.cu).
// ------------------------------------------------------------------------------------------------
PgenFile::~PgenFile() {
    if (mapped_ && data_) munmap(const_cast<uint8_t*>(data_), static_cast<size_t>(len_));
    if (fd_ >= 0) close(fd_);
}

std::string PgenFile::open(const char* path) {
    fd_ = ::open(path, O_RDONLY);
    if (fd_ < 0) return std::string("cannot open ") + path;
    struct stat st;
    if (fstat(fd_, &st) != 0) return "fstat failed";
    len_ = st.st_size;
    void* p = mmap(nullptr, static_cast<size_t>(len_), PROT_READ, MAP_PRIVATE, fd_, 0);
    if (p == MAP_FAILED) return "mmap failed";
    data_ = static_cast<const uint8_t*>(p);
    mapped_ = true;
    return parse();
}

std::string PgenFile::open_memory(const uint8_t* data, int64_t len) {
    data_ = data;
    len_ = len;
    mapped_ = false;
    return parse();
}

static inline uint64_t le_bytes(const uint8_t* p, int nb) {
    uint64_t v = 0;
    for (int i = 0; i < nb; ++i) v |= static_cast<uint64_t>(p[i]) << (8 * i);
    return v;
}

std::string PgenFile::parse() {
    if (len_ < 12 || data_[0] != 0x6c || data_[1] != 0x1b) return "not a .pgen file (bad magic)";
    mode_ = data_[2];
    variant_ct_ = static_cast<uint32_t>(le_bytes(data_ + 3, 4));
    sample_ct_ = static_cast<uint32_t>(le_bytes(data_ + 7, 4));
    if (mode_ == 0x02) {
        if (12 + bytes_per_variant() * static_cast<int64_t>(variant_ct_) > len_) return "truncated mode-0x02 .pgen";
        return "";
    }
    if (mode_ != 0x10) return "unsupported .pgen storage mode (only 0x02 and 0x10)";
    const uint8_t ctrl = data_[11];
    const int enc = ctrl & 15;
    if (enc > 7) return "unsupported .pgen record-length encoding";
    const int type_bits = enc < 4 ? 4 : 8;
    const int len_bytes = (enc & 3) + 1;
    const int allele_bytes = (ctrl >> 4) & 3;
    const bool nonref_flags = ((ctrl >> 6) & 3) == 3;
    const uint32_t nblk = (variant_ct_ + 65535u) >> 16;
    const uint8_t* hp = data_ + 12 + 8ull * nblk;
    recs_.resize(variant_ct_);
    ldbase_.assign(variant_ct_, 0);
    for (uint32_t b = 0; b < nblk; ++b) {
        const uint32_t first = b << 16;
        const uint32_t cnt = (variant_ct_ - first < 65536u) ? variant_ct_ - first : 65536u;
        const uint8_t* types = hp;
        hp += type_bits == 4 ? (cnt + 1) / 2 : cnt;
        const uint8_t* lens = hp;
        hp += static_cast<uint64_t>(cnt) * len_bytes;
        hp += static_cast<uint64_t>(cnt) * allele_bytes;
        if (nonref_flags) hp += (cnt + 7) / 8;
        if (hp - data_ > len_) return "truncated .pgen header";
        uint64_t off = le_bytes(data_ + 12 + 8ull * b, 8);
        uint32_t base = 0;
        bool have_base = false;
        for (uint32_t i = 0; i < cnt; ++i) {
            Rec& r = recs_[first + i];
            r.vrtype = type_bits == 4 ? ((types[i >> 1] >> (4 * (i & 1))) & 15) : types[i];
            r.len = static_cast<uint32_t>(le_bytes(lens + static_cast<uint64_t>(i) * len_bytes, len_bytes));
            r.off = off;
            off += r.len;
            if (static_cast<int64_t>(off) > len_) return "truncated .pgen record";
            if (r.vrtype & 0x60) has_dosage_ = true;
            if ((r.vrtype & 6) != 2) { base = first + i; have_base = true; }
            else if (!have_base) return "LD-compressed record without a base variant";
            ldbase_[first + i] = base;
        }
    }
    return "";
}

}  // namespace flx
