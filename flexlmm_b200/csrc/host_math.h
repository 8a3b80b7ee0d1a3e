// host_math.h — host-side (C++) pieces of the hot path that are inherently sequential or once-per-task:
// R-exact permutation indices, the QR of the SNP-independent design columns, .pgen record access.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace flx {

// set.seed(seed); sample.int(m) — R >= 3.6 defaults (Mersenne-Twister, Inversion, Rejection).
// fit_model.nf:97-98.  out: 1-based.
void r_sample_int(int32_t seed, int64_t m, int32_t* out);

// Orthonormal basis of the column space of C (n x nc, col-major, ld = n) with LINPACK dqrdc2's rank rule
// (limited pivoting, tol 1e-7).  Returns rank r; Q: n x r (col-major); R: r x nc coordinates of ALL original
// columns in that basis (col-major, ld = r).
int orth_basis(const double* C, int64_t n, int nc, std::vector<double>& Q, std::vector<double>& R);

// Null-model fit with an IMPLICIT orthonormal basis of the complement of the null design (FIT_NULL_MODEL without the n x n
// eigen-decomposition, fit_null_model.nf:33-72).  Householder QR of the accepted columns of X (dqrdc2's rank rule):
//   reflectors V (n x r, column j zero above row j), beta (r);  Q = H_1 ... H_r,  U1 = Q[:, r:]  (never formed).
struct NullFit {
    int64_t n = 0;
    int rank = 0;
    std::vector<double> V, beta;      // reflectors
    std::vector<double> fitted, resid;  // lm.fit
    std::vector<double> e_p;          // U1' e  (length n - rank)
    double ll = 0.0;                  // stats:::logLik.lm
    // z (n) <- Q z   /   z <- Q' z
    void apply_Q(double* z) const;
    void apply_Qt(double* z) const;
};
void null_fit_householder(const double* X, int64_t n, int c, const double* y, NullFit& out);

// .pgen access (pgenlibr::NewPgen, fit_model.nf:39-40): header, record index, raw record bytes.  ReadHardcalls (:123) is
// K0 (pgen_expand.cu) + K1 on the device.
class PgenFile {
public:
    ~PgenFile();
    // returns "" on success, else an error message
    std::string open(const char* path);
    std::string open_memory(const uint8_t* data, int64_t len);  // not owned
    uint32_t variant_ct() const { return variant_ct_; }
    uint32_t sample_ct() const { return sample_ct_; }
    int64_t bytes_per_variant() const { return ((int64_t)sample_ct_ + 3) / 4; }
    int mode() const { return mode_; }
    bool has_dosage() const { return has_dosage_; }
    // If the file is mode 0x02 the records are already packed 2-bit: pointer to record v (0-based).
    const uint8_t* fixed_record(uint32_t v) const { return data_ + 12 + (int64_t)v * bytes_per_variant(); }
    // mode 0x10: where variant v's record lives in the file, its type byte and the variant its LD compression refers to
    struct RecInfo { uint64_t off; uint32_t len; uint8_t vrtype; uint32_t ldbase; };
    RecInfo rec_info(uint32_t v) const { return RecInfo{recs_[v].off, recs_[v].len, recs_[v].vrtype, ldbase_[v]}; }
    const uint8_t* data() const { return data_; }
private:
    struct Rec { uint64_t off; uint32_t len; uint8_t vrtype; };
    std::string parse();
    const uint8_t* data_ = nullptr;
    int64_t len_ = 0;
    bool mapped_ = false;
    int fd_ = -1;
    uint32_t variant_ct_ = 0, sample_ct_ = 0;
    int mode_ = 0;
    bool has_dosage_ = false;
    std::vector<Rec> recs_;           // mode 0x10
    std::vector<uint32_t> ldbase_;    // mode 0x10: index of the LD base record for LD-compressed variants
};

}  // namespace flx
