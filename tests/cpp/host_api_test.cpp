// host_api_test.cpp -- the reference's own unit tests for this path, rewritten against the C++ host
// layer (include/scanb200.hpp).  Reads like scan-rs/src/normalization.rs:539-575 and
// scan-rs/src/dim_red/test.rs:58-110.  Exit code 0 = all passed.  Needs a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "scanb200.hpp"

using namespace scanb200;
using normalization::Normalization;

static int failures = 0;
#define EXPECT(cond)                                                  \
    do {                                                              \
        if (!(cond)) {                                                \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                               \
        }                                                             \
    } while (0)

static bool abs_diff_eq(const Array2 &a, const double *b, double eps) {
    for (size_t i = 0; i < a.data.size(); i++)
        if (std::fabs(a.data[i] - b[i]) > eps) return false;
    return true;
}

static void test_cellranger_normalisation(Context &ctx) {  // normalization.rs:539-575
    std::vector<uint32_t> dense = {136, 936, 0, 0, 264, 134, 682, 417, 8, 391, 0, 133, 780, 0, 0, 396, 76, 96, 198, 0};
    const double expected[] = {0.61392149,  0.95459951, -1.21707302, -1.21707302, 0.86562504,  -0.11878431, 0.54279925,
                               0.38607315,  -1.85660965, 1.04652156, -0.78758751, 0.76437149,  1.59839105,  -0.78758751,
                               -0.78758751, 0.88718256, -0.25584717, -0.01048423, 1.09574143,  -1.71659259};
    auto mtx = sqz::AdaptiveMat::from_dense(ctx, 4, 5, dense);
    auto norm_mat = normalization::normalize_with_size_factor(mtx, Normalization::CellRanger, nullptr);
    EXPECT(abs_diff_eq(norm_mat.to_dense(), expected, 1e-6));
    auto sums = mtx.sum_axis0_u32();
    EXPECT(sums[0] == 666 && sums[1] == 1827 && sums[4] == 655);
}

static void test_packed_constructor(Context &ctx) {  // the same golden matrix through sb_pack_csc_* + sb_upload_packed (cell-major arrays)
    const uint32_t dense[4][5] = {{136, 936, 0, 0, 264}, {134, 682, 417, 8, 391}, {0, 133, 780, 0, 0}, {396, 76, 96, 198, 0}};
    std::vector<uint64_t> indptr{0};
    std::vector<uint32_t> idx, val;
    for (uint32_t c = 0; c < 5; c++) {
        for (uint32_t g = 0; g < 4; g++)
            if (dense[g][c]) {
                idx.push_back(g);
                val.push_back(dense[g][c]);
            }
        indptr.push_back(idx.size());
    }
    auto mtx = sqz::AdaptiveMat::from_csc_packed(ctx, 4, 5, indptr, idx, val, 2);
    auto sums = mtx.sum_axis0_u32();
    EXPECT(sums[0] == 666 && sums[1] == 1827 && sums[2] == 1293 && sums[3] == 206 && sums[4] == 655);  // column sums of the golden matrix
}

struct Recorder : snoop::CancelProgress {
    std::vector<double> seen;
    size_t cancel_after = 1000;
    bool is_cancelled() const override { return seen.size() >= cancel_after; }
    void set_progress(double f) override { seen.push_back(f); }
};

static void test_bksvd(Context &ctx) {  // dim_red/test.rs:58-110 on a sparse count matrix
    const uint32_t m = 300;
    const uint64_t n = 900;
    std::vector<uint32_t> dense(m * n, 0);
    uint64_t s = 12345;
    for (auto &v : dense) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        uint32_t r = (uint32_t)(s >> 40);
        v = (r % 5 == 0) ? 1 + (r >> 8) % 7 : 0;
    }
    auto mtx = sqz::AdaptiveMat::from_dense(ctx, m, n, dense);
    auto a = normalization::normalize(mtx, Normalization::CellRanger);
    dim_red::BkSvd svd;
    auto res = svd.run_pca(a, 10);
    EXPECT(res.u.rows == m && res.u.cols == 10 && res.v.rows == n && res.v.cols == 10 && res.s.size() == 10);
    for (size_t i = 1; i < 10; i++) EXPECT(res.s[i] <= res.s[i - 1]);
    // n > m branch: T = Q^T A = U_T S V^T and U = Q U_T, so u^T A = s v^T is an identity of the method
    // (the reference's ||A v - u s|| bar, test.rs:69-75, needs a matrix with a decaying spectrum; this one is noise)
    Array2 ut(10, m);
    for (size_t r = 0; r < m; r++)
        for (size_t j = 0; j < 10; j++) ut(j, r) = res.u(r, j);
    Array2 uta = a.dot_left(ut);
    double worst = 0.0;
    for (size_t j = 0; j < 10; j++)
        for (size_t c = 0; c < n; c++) worst = std::fmax(worst, std::fabs(uta(j, c) - res.s[j] * res.v(c, j)));
    EXPECT(worst < 1e-9 * res.s[0]);
    Array2 av = a.dot(res.v);
    EXPECT(av.rows == m && av.cols == 10 && dim_red::frobenius(av) > 0.0);
    // error behaviour (bk_svd.rs:73-79)
    try {
        svd.run_pca(a, 301);
        EXPECT(false);
    } catch (const Error &e) {
        EXPECT(std::strcmp(e.what(), "invalid k") == 0 && e.code == SB_ERR_INVALID_K);
    }
    // progress milestones and cancellation (bk_svd.rs:125-143, snoop/src/lib.rs:45-57)
    Recorder rec;
    svd.run_pca_cancellable(a, 4, rec);
    EXPECT(rec.seen.size() == 8 && rec.seen[5] == 0.82 && rec.seen[6] == 0.93 && rec.seen[7] == 1.0);
    Recorder rec2;
    rec2.cancel_after = 2;
    try {
        svd.run_pca_cancellable(a, 4, rec2);
        EXPECT(false);
    } catch (const CancellationError &) {
        EXPECT(rec2.seen.size() == 2);
    }
    EXPECT(normalization::from_str("seuratlog") == Normalization::SeuratLog);
}

static void test_mean_var_and_sum_fns(Context &ctx) {  // sqz/src/mat.rs:1302-1370 on input_a
    std::vector<uint32_t> dense = {136, 936, 0, 0, 264, 134, 682, 417, 8, 391, 0, 133, 780, 885, 0, 396, 76, 96, 198, 0};
    auto mtx = sqz::AdaptiveMat::from_dense(ctx, 4, 5, dense);
    auto mv0 = mtx.mean_var_axis(0);
    const double e_mean0[] = {166.5, 456.75, 323.25, 272.75, 163.75}, e_var0[] = {20594.75, 132550.6875, 93385.6875, 131230.6875, 28830.1875};
    for (int i = 0; i < 5; i++) EXPECT(std::fabs(mv0.first[i] - e_mean0[i]) < 1e-7 && std::fabs(mv0.second[i] - e_var0[i]) < 1e-7);
    auto mv1 = mtx.mean_var_axis(1);
    const double e_mean1[] = {267.2, 326.4, 359.6, 153.2}, e_var1[] = {121461.76, 55445.84, 152550.64, 18732.16};
    for (int i = 0; i < 4; i++) EXPECT(std::fabs(mv1.first[i] - e_mean1[i]) < 1e-7 && std::fabs(mv1.second[i] - e_var1[i]) < 1e-7);
    auto dual = mtx.sum_rows_dual({1, 2, 3}, {2, 3, 4});
    const uint64_t e1[] = {936, 1107, 1798, 370}, e2[] = {264, 816, 1665, 294};
    for (int i = 0; i < 4; i++) EXPECT(dual.first[i] == e1[i] && dual.second[i] == e2[i]);
    auto sf = mtx.size_factors();  // counts / median (diff_exp.rs:314-334): totals 666 1827 1293 1091 655 -> median 1091
    EXPECT(std::fabs(sf[0] - 666.0 / 1091.0) < 1e-15 && std::fabs(sf[3] - 1.0) < 1e-15);
}

static void test_irlba(Context &ctx) {  // dim_red/test.rs:133-139: Irlba { tol: 1e-5 }, A^T u = s v to the tolerance
    const uint32_t m = 200;
    const uint64_t n = 700;
    std::vector<uint32_t> dense(m * n, 0);
    uint64_t s = 777;
    for (size_t i = 0; i < dense.size(); i++) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        uint32_t r = (uint32_t)(s >> 40);
        dense[i] = (r % 4 == 0) ? 1 + (r >> 8) % 5 + ((i / n) % 3 == 0 ? 3 : 0) : 0;
    }
    auto mtx = sqz::AdaptiveMat::from_dense(ctx, m, n, dense);
    auto a = normalization::normalize(mtx, Normalization::CellRanger);
    dim_red::Irlba irl;
    irl.tol = 1e-5;
    auto res = irl.run_pca(a, 5);
    EXPECT(res.u.rows == m && res.v.rows == n && res.s.size() == 5);
    for (size_t i = 1; i < 5; i++) EXPECT(res.s[i] <= res.s[i - 1]);
    Array2 ut(5, m);
    for (size_t r = 0; r < m; r++)
        for (size_t j = 0; j < 5; j++) ut(j, r) = res.u(r, j);
    Array2 uta = a.dot_left(ut);
    double worst = 0.0;
    for (size_t j = 0; j < 5; j++)
        for (size_t c = 0; c < n; c++) worst = std::fmax(worst, std::fabs(uta(j, c) - res.s[j] * res.v(c, j)));
    EXPECT(worst < 1e-3 * res.s[0]);
}

int main() {
    try {
        Context ctx(0);
        test_cellranger_normalisation(ctx);
        test_packed_constructor(ctx);
        test_bksvd(ctx);
        test_mean_var_and_sum_fns(ctx);
        test_irlba(ctx);
    } catch (const Error &e) {
        std::printf("FAIL: exception %d: %s\n", e.code, e.what());
        return 2;
    }
    std::printf(failures ? "%d FAILURES\n" : "ALL PASSED\n", failures);
    return failures ? 1 : 0;
}
