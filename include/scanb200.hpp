// scanb200.hpp -- header-only C++ host layer over the C ABI (scanb200.h).
//
// The reference is a Rust workspace and no Rust toolchain exists in this image, so the host side
// above the C ABI is written in C++ and mirrors the reference's interface for this path: same
// module layout as namespaces (sqz, normalization, dim_red, snoop), same type and function names,
// same argument meaning, same error messages.  A Rust crate would wrap the same C symbols the same
// way.  Everything here is thin: all logic lives behind the C ABI.
//
//   reference                                                    here
//   sqz::AdaptiveMat<u32>            sqz/src/mat.rs:34-42         scanb200::sqz::AdaptiveMat
//   sqz::LowRankOffset               sqz/src/low_rank_offset.rs   scanb200::sqz::LowRankOffset
//   scan_rs::normalization::*        scan-rs/src/normalization.rs scanb200::normalization::*
//   scan_rs::dim_red::{Pca,BkSvd,..} scan-rs/src/dim_red/*.rs     scanb200::dim_red::*
//   snoop::{CancelProgress,NoOpSnoop} snoop/src/lib.rs            scanb200::snoop::*
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "scanb200.h"

namespace scanb200 {

// anyhow::Error carrying the library message; CancellationError mirrors snoop::CancellationError
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};
struct CancellationError : Error {
    using Error::Error;
};

inline void check(int rc) {
    if (rc == SB_OK) return;
    std::string msg = sb_last_error();
    if (rc == SB_ERR_CANCELLED) throw CancellationError(rc, msg);
    throw Error(rc, msg);
}

// ndarray::Array2<f64> in standard (row-major) layout
struct Array2 {
    size_t rows = 0, cols = 0;
    std::vector<double> data;
    Array2() {}
    Array2(size_t r, size_t c) : rows(r), cols(c), data(r * c, 0.0) {}
    double &operator()(size_t r, size_t c) { return data[r * cols + c]; }
    double operator()(size_t r, size_t c) const { return data[r * cols + c]; }
    std::array<size_t, 2> shape() const { return {rows, cols}; }
};

class Context {
  public:
    explicit Context(int device = 0) { check(sb_init(device, &h_)); }
    ~Context() {
        if (h_) sb_shutdown(h_);
    }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    static std::array<char, 128> unique_id() {
        std::array<char, 128> id{};
        check(sb_comm_unique_id(id.data()));
        return id;
    }
    void comm_init(int nranks, int rank, const std::array<char, 128> &id) { check(sb_comm_init(h_, nranks, rank, id.data())); }
    void set_option(const char *name, double v) { check(sb_set_option(h_, name, v)); }
    sb_ctx *raw() const { return h_; }

  private:
    sb_ctx *h_ = nullptr;
};

namespace snoop {
// snoop::CancelProgress (snoop/src/lib.rs:37-58): is_cancelled + set_progress; set_progress_check is
// implemented once in the callback trampoline below.
struct CancelProgress {
    virtual ~CancelProgress() {}
    virtual bool is_cancelled() const = 0;
    virtual void set_progress(double fraction) = 0;
};
struct NoOpSnoop : CancelProgress {  // snoop/src/lib.rs:60-85
    bool is_cancelled() const override { return false; }
    void set_progress(double) override {}
};
inline int trampoline(double fraction, void *user) {
    auto *s = static_cast<CancelProgress *>(user);
    if (s->is_cancelled()) return 1;  // Err(CancellationError) before the progress is recorded (:50-56)
    s->set_progress(fraction);
    return 0;
}
}  // namespace snoop

namespace sqz {

class LowRankOffset;

// AdaptiveMat<u32>: the count matrix, features (rows) x barcodes (cols), resident on the device.
class AdaptiveMat {
  public:
    using Partition = std::tuple<AdaptiveMat, AdaptiveMat, std::vector<size_t>, std::vector<size_t>>;

    // from_csmat on CSR arrays: indptr over genes (sqz/src/mat.rs:92-124)
    static AdaptiveMat from_csr(Context &ctx, uint32_t rows, uint64_t cols, const std::vector<uint64_t> &indptr, const std::vector<uint32_t> &idx,
                                const std::vector<uint32_t> &val) {
        sb_mat *h = nullptr;
        check(sb_upload(ctx.raw(), SB_GENE_MAJOR, rows, cols, indptr.data(), idx.data(), val.data(), &h));
        return AdaptiveMat(h);
    }
    // CSC arrays: indptr over cells
    static AdaptiveMat from_csc(Context &ctx, uint32_t rows, uint64_t cols, const std::vector<uint64_t> &indptr, const std::vector<uint32_t> &idx,
                                const std::vector<uint32_t> &val) {
        sb_mat *h = nullptr;
        check(sb_upload(ctx.raw(), SB_CELL_MAJOR, rows, cols, indptr.data(), idx.data(), val.data(), &h));
        return AdaptiveMat(h);
    }
    // CSC arrays in the narrow host form of sb_upload_compact (rows <= 65536): counts >= 255 go to the side list
    static AdaptiveMat from_csc_compact(Context &ctx, uint32_t rows, uint64_t cols, const std::vector<uint64_t> &indptr,
                                        const std::vector<uint32_t> &idx, const std::vector<uint32_t> &val) {
        std::vector<uint16_t> idx16(idx.size());
        std::vector<uint8_t> cnt8(val.size());
        std::vector<uint64_t> big_pos;
        std::vector<uint32_t> big_cnt;
        for (size_t i = 0; i < idx.size(); i++) {
            idx16[i] = (uint16_t)idx[i];
            cnt8[i] = (uint8_t)std::min<uint32_t>(val[i], 255u);
            if (val[i] >= 255u) {
                big_pos.push_back(i);
                big_cnt.push_back(val[i]);
            }
        }
        sb_mat *h = nullptr;
        check(sb_upload_compact(ctx.raw(), rows, cols, indptr.data(), idx16.data(), cnt8.data(), big_pos.size(), big_pos.data(), big_cnt.data(), &h));
        return AdaptiveMat(h);
    }
    // CSC arrays in the packed host form of sb_upload_packed (gene delta byte + count nibble + side lists), built by the
    // library's host encoder (sb_pack_csc_count / sb_pack_csc_fill)
    static AdaptiveMat from_csc_packed(Context &ctx, uint32_t rows, uint64_t cols, const std::vector<uint64_t> &indptr,
                                       const std::vector<uint32_t> &idx, const std::vector<uint32_t> &val, int threads = 0) {
        uint64_t n_esc = 0, n_big = 0;
        check(sb_pack_csc_count(cols, indptr.data(), idx.data(), val.data(), threads, &n_esc, &n_big));
        std::vector<uint8_t> dgene(idx.size()), cnt4((idx.size() + 1) / 2);
        std::vector<uint64_t> esc_pos(n_esc), big_pos(n_big);
        std::vector<uint32_t> esc_gene(n_esc), big_cnt(n_big);
        check(sb_pack_csc_fill(cols, indptr.data(), idx.data(), val.data(), threads, dgene.data(), cnt4.data(), esc_pos.data(), esc_gene.data(),
                               big_pos.data(), big_cnt.data()));
        sb_mat *h = nullptr;
        check(sb_upload_packed(ctx.raw(), rows, cols, indptr.data(), dgene.data(), cnt4.data(), n_esc, esc_pos.data(), esc_gene.data(), n_big,
                               big_pos.data(), big_cnt.data(), &h));
        return AdaptiveMat(h);
    }
    // from_dense (mat.rs:586-609); dense is row-major rows x cols
    static AdaptiveMat from_dense(Context &ctx, uint32_t rows, uint64_t cols, const std::vector<uint32_t> &dense) {
        std::vector<uint64_t> indptr{0};
        std::vector<uint32_t> idx, val;
        for (uint32_t r = 0; r < rows; r++) {
            for (uint64_t c = 0; c < cols; c++)
                if (dense[r * cols + c]) {
                    idx.push_back((uint32_t)c);
                    val.push_back(dense[r * cols + c]);
                }
            indptr.push_back(idx.size());
        }
        return from_csr(ctx, rows, cols, indptr, idx, val);
    }
    AdaptiveMat(AdaptiveMat &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    AdaptiveMat &operator=(AdaptiveMat &&o) noexcept {
        std::swap(h_, o.h_);
        return *this;
    }
    ~AdaptiveMat() {
        if (h_) sb_free_mat(h_);
    }
    size_t rows() const { return shape()[0]; }
    size_t cols() const { return shape()[1]; }
    std::array<size_t, 2> shape() const {
        uint32_t m;
        uint64_t n;
        check(sb_mat_shape(h_, &m, &n, nullptr, nullptr));
        return {m, (size_t)n};
    }
    size_t nnz() const {
        uint64_t z;
        check(sb_mat_shape(h_, nullptr, nullptr, nullptr, &z));
        return z;
    }
    // sum_axis::<u32>(Axis(0)): per-barcode totals (mat.rs:377-406)
    std::vector<uint32_t> sum_axis0_u32() const {
        std::vector<uint32_t> out(cols());
        check(sb_cell_totals(h_, out.data()));
        return out;
    }
    // sum_axis(Axis(1)) in u64 (hdf5-io/src/matrix.rs:106-114)
    std::vector<uint64_t> sum_axis1_u64() const {
        std::vector<uint64_t> out(rows());
        check(sb_gene_totals(h_, 0, out.data()));
        return out;
    }
    // mean_var_axis (mat.rs:285-329); size_factors (optional, one per cell): the SizeNormalized view of diff-exp (diff_exp.rs:340-358)
    std::pair<std::vector<double>, std::vector<double>> mean_var_axis(int axis, const std::vector<double> *size_factors = nullptr) const {
        std::vector<double> mean(axis == 0 ? cols() : rows()), var(mean.size());
        check(sb_mean_var_axis(h_, axis, size_factors ? size_factors->data() : nullptr, mean.data(), var.data()));
        return {std::move(mean), std::move(var)};
    }
    // mean_var_rows (mat.rs:332-374)
    std::pair<std::vector<double>, std::vector<double>> mean_var_rows(const std::vector<uint64_t> &cells, const std::vector<double> *size_factors = nullptr) const {
        std::vector<double> mean(rows()), var(rows());
        check(sb_mean_var_rows(h_, cells.data(), cells.size(), size_factors ? size_factors->data() : nullptr, mean.data(), var.data()));
        return {std::move(mean), std::move(var)};
    }
    // sum_rows_dual (mat.rs:484-583), exact u64
    std::pair<std::vector<uint64_t>, std::vector<uint64_t>> sum_rows_dual(const std::vector<uint64_t> &cols1, const std::vector<uint64_t> &cols2) const {
        std::vector<uint64_t> a(rows()), b(rows());
        check(sb_sum_rows_dual(h_, cols1.data(), cols1.size(), cols2.data(), cols2.size(), a.data(), b.data()));
        return {std::move(a), std::move(b)};
    }
    // size_factors (diff-exp/src/diff_exp.rs:314-334) over all cells
    std::vector<double> size_factors() const {
        std::vector<double> out(cols());
        check(sb_size_factors(h_, nullptr, 0, nullptr, out.data()));
        return out;
    }
    Partition partition_on_threshold(double threshold) const { return partition_on_thresholds(true, threshold, true, threshold); }  // mat.rs:766
    // partition_on_thresholds(Option<f64>, Option<f64>) (mat.rs:772-889)
    Partition partition_on_thresholds(bool has_row, double row_thr, bool has_col, double col_thr) const {
        auto sh = shape();
        std::vector<uint64_t> rows(sh[0]), cols(sh[1]);
        uint64_t nr = 0, nc = 0;
        sb_mat *kept = nullptr, *resid = nullptr;
        check(sb_partition(h_, has_row, row_thr, has_col, col_thr, &kept, &resid, rows.data(), &nr, cols.data(), &nc));
        return Partition(AdaptiveMat(kept), AdaptiveMat(resid), std::vector<size_t>(rows.begin(), rows.begin() + nr),
                         std::vector<size_t>(cols.begin(), cols.begin() + nc));
    }
    AdaptiveMat select_rows(const std::vector<uint32_t> &rows) const {  // mat.rs:1040-1071
        sb_mat *h = nullptr;
        check(sb_select_rows(h_, rows.data(), (uint32_t)rows.size(), &h));
        return AdaptiveMat(h);
    }
    AdaptiveMat select_cols(const std::vector<uint64_t> &cols) const {  // mat.rs:1004-1037
        sb_mat *h = nullptr;
        check(sb_select_cols(h_, cols.data(), cols.size(), &h));
        return AdaptiveMat(h);
    }
    sb_mat *raw() const { return h_; }

  private:
    explicit AdaptiveMat(sb_mat *h) : h_(h) {}
    sb_mat *h_ = nullptr;
    friend class LowRankOffset;
};

// LowRankOffset: A = map(counts) + u.v (low_rank_offset.rs:12-16).  Borrows its AdaptiveMat.
class LowRankOffset {
  public:
    explicit LowRankOffset(sb_nmat *h, const AdaptiveMat &m) : h_(h), mat_(&m) {}
    LowRankOffset(LowRankOffset &&o) noexcept : h_(o.h_), mat_(o.mat_) { o.h_ = nullptr; }
    ~LowRankOffset() {
        if (h_) sb_free_nmat(h_);
    }
    size_t rows() const { return mat_->rows(); }
    size_t cols() const { return mat_->cols(); }
    std::array<size_t, 2> shape() const { return mat_->shape(); }
    Array2 to_dense() const {  // :55-57
        Array2 out(rows(), cols());
        check(sb_nmat_to_dense(h_, out.data.data()));
        return out;
    }
    Array2 dot(const Array2 &rhs) const {  // A . rhs  (:68-81)
        if (rhs.rows != cols()) throw Error(SB_ERR_INVALID_ARG, "Dimension mismatch");
        Array2 out(rows(), rhs.cols);
        check(sb_nmat_dot(h_, rhs.data.data(), (uint32_t)rhs.cols, out.data.data()));
        return out;
    }
    Array2 dot_left(const Array2 &lhs) const {  // lhs . A  (:83-96)
        if (lhs.cols != rows()) throw Error(SB_ERR_INVALID_ARG, "Dimension mismatch");
        Array2 out(lhs.rows, cols());
        check(sb_nmat_rdot(h_, lhs.data.data(), (uint32_t)lhs.rows, out.data.data()));
        return out;
    }
    sb_nmat *raw() const { return h_; }

  private:
    sb_nmat *h_ = nullptr;
    const AdaptiveMat *mat_ = nullptr;
};
}  // namespace sqz

namespace normalization {
enum class Normalization {  // normalization.rs:11-28
    CellRanger = SB_NORM_CELLRANGER,
    CellRanger8 = SB_NORM_CELLRANGER8,
    SeuratLog = SB_NORM_SEURATLOG,
    BinomialDeviance = SB_NORM_BINOMIAL_DEVIANCE,
    BinomialPearson = SB_NORM_BINOMIAL_PEARSON,
    WithSizeFactors = SB_NORM_WITH_SIZE_FACTORS,
    LogTransform = SB_NORM_LOG_TRANSFORM
};
enum class LogBase { E = SB_LOG_E, Two = SB_LOG_TWO, Ten = SB_LOG_TEN };  // :105-112
struct FixedPointFormat {                                                    // :181-187
    uint32_t base, exponent;
};

inline Normalization from_str(const std::string &s) {  // impl FromStr (:30-43)
    if (s == "cellranger") return Normalization::CellRanger;
    if (s == "cellranger8") return Normalization::CellRanger8;
    if (s == "seuratlog") return Normalization::SeuratLog;
    if (s == "binomialdeviance") return Normalization::BinomialDeviance;
    if (s == "binomialpearson") return Normalization::BinomialPearson;
    throw Error(SB_ERR_INVALID_ARG, "Normalization not recognized: " + s);
}

// normalize (:46-69): CellRanger, CellRanger8, SeuratLog; the rest is the reference's panic!("not implemented")
inline sqz::LowRankOffset normalize(const sqz::AdaptiveMat &mat, Normalization norm) {
    if (norm != Normalization::CellRanger && norm != Normalization::CellRanger8 && norm != Normalization::SeuratLog)
        throw Error(SB_ERR_INVALID_ARG, "not implemented");
    sb_nmat *h = nullptr;
    check(sb_normalize(mat.raw(), (int)norm, nullptr, &h));
    return sqz::LowRankOffset(h, mat);
}
// normalize_with_size_factor (:72-102)
inline sqz::LowRankOffset normalize_with_size_factor(const sqz::AdaptiveMat &mat, Normalization norm, const std::vector<uint32_t> *size_factors) {
    if (norm == Normalization::BinomialDeviance || norm == Normalization::BinomialPearson) throw Error(SB_ERR_INVALID_ARG, "not implemented");
    const uint32_t *sf = nullptr;
    if (norm == Normalization::WithSizeFactors && size_factors) {
        if (size_factors->size() != mat.cols()) throw Error(SB_ERR_INVALID_ARG, "Size of the size factor and matrix columns dont match.");
        sf = size_factors->data();
    }
    sb_nmat *h = nullptr;
    check(sb_normalize(mat.raw(), (int)norm, sf, &h));
    return sqz::LowRankOffset(h, mat);
}
// log_normalize_with_size_factor (:138-178): umi_count_sum = nullptr -> median of the barcode totals
inline sqz::LowRankOffset log_normalize_with_size_factor(const sqz::AdaptiveMat &mat, const double *umi_count_sum, LogBase base,
                                                         const std::vector<uint32_t> *size_factors) {
    if (size_factors && size_factors->size() != mat.cols()) throw Error(SB_ERR_INVALID_ARG, "Size of the size factor and matrix columns dont match.");
    sb_nmat *h = nullptr;
    check(sb_log_normalize(mat.raw(), umi_count_sum != nullptr, umi_count_sum ? *umi_count_sum : 0.0, (int)base,
                           size_factors ? size_factors->data() : nullptr, 0, nullptr, &h));
    return sqz::LowRankOffset(h, mat);
}
inline sqz::LowRankOffset log1p_normalize_fixed_point(const sqz::AdaptiveMat &mat, LogBase base, FixedPointFormat fp) {  // :191-213
    sb_nmat *h = nullptr;
    check(sb_normalize_fixed_point(mat.raw(), (int)base, fp.base, fp.exponent, &h));
    return sqz::LowRankOffset(h, mat);
}
inline sqz::LowRankOffset binom_deviance_resid(const sqz::AdaptiveMat &mat) {  // :233-260
    sb_nmat *h = nullptr;
    check(sb_normalize(mat.raw(), SB_NORM_BINOMIAL_DEVIANCE, nullptr, &h));
    return sqz::LowRankOffset(h, mat);
}
inline sqz::LowRankOffset binom_pearson_resid(const sqz::AdaptiveMat &mat) {  // :307-323
    sb_nmat *h = nullptr;
    check(sb_normalize(mat.raw(), SB_NORM_BINOMIAL_PEARSON, nullptr, &h));
    return sqz::LowRankOffset(h, mat);
}
}  // namespace normalization

namespace dim_red {
// PcaResult = (Array2<f64>, Array1<f64>, Array2<f64>) = (u m x k, s k, v n x k)  (dim_red/mod.rs:47)
struct PcaResult {
    Array2 u;
    std::vector<double> s;
    Array2 v;
};

// trait Pca<T, f64> (dim_red/mod.rs:103-111)
struct Pca {
    virtual ~Pca() {}
    virtual PcaResult run_pca_cancellable(const sqz::LowRankOffset &matrix, size_t k, snoop::CancelProgress &c) const = 0;
    PcaResult run_pca(const sqz::LowRankOffset &matrix, size_t k) const {
        snoop::NoOpSnoop s;
        return run_pca_cancellable(matrix, k, s);
    }
};

// svd_bk (bk_svd.rs:57-146); returns (U, sigma, Va) with Va = k x n like the reference
inline std::tuple<Array2, std::vector<double>, Array2> svd_bk(const sqz::LowRankOffset &A, size_t k, size_t b, size_t n_iter, uint64_t seed,
                                                               snoop::CancelProgress &snoop_) {
    Array2 U(A.rows(), k), V(A.cols(), k);
    std::vector<double> S(k);
    check(sb_bksvd(A.raw(), (uint32_t)k, (uint32_t)b, (uint32_t)n_iter, seed, nullptr, snoop::trampoline, &snoop_, U.data.data(), S.data(), V.data.data()));
    Array2 Va(k, A.cols());
    for (size_t c = 0; c < V.rows; c++)
        for (size_t j = 0; j < k; j++) Va(j, c) = V(c, j);
    return {std::move(U), std::move(S), std::move(Va)};
}

struct BkSvd : Pca {  // bk_svd.rs:16-53
    double k_multiplier = 2.0;
    size_t n_iter = 5;
    PcaResult run_pca_cancellable(const sqz::LowRankOffset &array, size_t k, snoop::CancelProgress &c) const override {
        PcaResult r{Array2(array.rows(), k), std::vector<double>(k), Array2(array.cols(), k)};
        check(sb_bksvd_run_pca(array.raw(), (uint32_t)k, k_multiplier, (uint32_t)n_iter, snoop::trampoline, &c, r.u.data.data(), r.s.data(),
                               r.v.data.data()));
        return r;
    }
};

struct RandSvd : Pca {  // rand_svd.rs:13-50 (ignores the snoop, :44-45)
    double l_multiplier = 10.0;
    size_t n_iter = 2;
    PcaResult run_pca_cancellable(const sqz::LowRankOffset &array, size_t k, snoop::CancelProgress &) const override {
        PcaResult r{Array2(array.rows(), k), std::vector<double>(k), Array2(array.cols(), k)};
        check(sb_randsvd_run_pca(array.raw(), (uint32_t)k, l_multiplier, (uint32_t)n_iter, r.u.data.data(), r.s.data(), r.v.data.data()));
        return r;
    }
};

struct Irlba : Pca {  // irlba.rs:36-69 (the reference cannot run it on a LowRankOffset: no 1-D Dot; the device type can)
    double tol = 0.0001;
    size_t max_iter = 50;
    PcaResult run_pca_cancellable(const sqz::LowRankOffset &array, size_t k, snoop::CancelProgress &c) const override {
        PcaResult r{Array2(array.rows(), k), std::vector<double>(k), Array2(array.cols(), k)};
        check(sb_irlba(array.raw(), (uint32_t)k, tol, (uint32_t)max_iter, nullptr, snoop::trampoline, &c, r.u.data.data(), r.s.data(), r.v.data.data(),
                       nullptr, nullptr));
        return r;
    }
};

// frobenius (dim_red/mod.rs:114-122): sqrt(sum v^2) / (rows * cols)
inline double frobenius(const Array2 &a) {
    double acc = 0.0;
    for (double v : a.data) acc += v * v;
    return std::sqrt(acc) / (double)(a.rows * a.cols);
}
}  // namespace dim_red
}  // namespace scanb200
