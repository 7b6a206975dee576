// CPU test of scan_rs_b200/csrc/eig_host.h (the k largest eigenpairs of the small Gram matrix, host side): matrices with a KNOWN
// spectrum G = Q diag(lambda) Q^T -- separated, PCA-like decay over ten orders of magnitude, clustered, repeated, rank-deficient, tiny
// orders -- must come back with the right eigenvalues, small residuals on the original matrix and orthonormal vectors.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../scan_rs_b200/csrc/eig_host.h"

static int failures = 0;
#define EXPECT(c)                                                     \
    do {                                                              \
        if (!(c)) {                                                   \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); \
            failures++;                                               \
        }                                                             \
    } while (0)

static std::vector<double> random_orthogonal(int n, std::mt19937_64 &rng) {
    std::normal_distribution<double> nd;
    std::vector<double> q((size_t)n * n);
    for (auto &x : q) x = nd(rng);
    for (int pass = 0; pass < 2; pass++)  // Gram-Schmidt twice over the columns
        for (int j = 0; j < n; j++) {
            for (int i = 0; i < j; i++) {
                double d = 0;
                for (int r = 0; r < n; r++) d += q[(size_t)r * n + i] * q[(size_t)r * n + j];
                for (int r = 0; r < n; r++) q[(size_t)r * n + j] -= d * q[(size_t)r * n + i];
            }
            double s = 0;
            for (int r = 0; r < n; r++) s += q[(size_t)r * n + j] * q[(size_t)r * n + j];
            s = std::sqrt(s);
            for (int r = 0; r < n; r++) q[(size_t)r * n + j] /= s;
        }
    return q;
}

static void run_case(const char *name, std::vector<double> lam, int k, std::mt19937_64 &rng) {
    const int n = (int)lam.size();
    auto q = random_orthogonal(n, rng);
    std::vector<double> G((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double s = 0;
            for (int t = 0; t < n; t++) s += q[(size_t)i * n + t] * lam[t] * q[(size_t)j * n + t];
            G[(size_t)i * n + j] = s;
        }
    std::vector<double> ev((size_t)k), vec((size_t)n * k);
    const bool ok = eig_host::topk(G.data(), n, k, ev.data(), vec.data());
    EXPECT(ok);
    if (!ok) {
        std::printf("  case %s: solver declined\n", name);
        return;
    }
    std::sort(lam.begin(), lam.end());
    double lmax = std::max(std::fabs(lam.front()), std::fabs(lam.back())), worst_ev = 0, worst_res = 0, worst_orth = 0;
    for (int i = 0; i < k; i++) {
        worst_ev = std::max(worst_ev, std::fabs(ev[i] - lam[n - k + i]));
        for (int r = 0; r < n; r++) {
            double s = 0;
            for (int c = 0; c < n; c++) s += G[(size_t)r * n + c] * vec[(size_t)i * n + c];
            worst_res = std::max(worst_res, std::fabs(s - ev[i] * vec[(size_t)i * n + r]));
        }
        for (int j = 0; j <= i; j++) {
            double d = 0;
            for (int r = 0; r < n; r++) d += vec[(size_t)i * n + r] * vec[(size_t)j * n + r];
            worst_orth = std::max(worst_orth, std::fabs(d - (i == j ? 1.0 : 0.0)));
        }
    }
    std::printf("  case %-28s n=%3d k=%3d  |ev - exact| %.2e  residual %.2e  orthonormality %.2e  (lambda_max %.3g)\n", name, n, k, worst_ev / lmax,
                worst_res / lmax, worst_orth, lmax);
    EXPECT(worst_ev <= 1e-13 * lmax * n);
    EXPECT(worst_res <= 1e-12 * lmax);
    EXPECT(worst_orth <= 1e-12);
}

int main() {
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> ud(0.0, 1.0);
    {
        std::vector<double> lam(100);
        for (int i = 0; i < 100; i++) lam[i] = 1.0 + i;
        run_case("separated", lam, 10, rng);
        run_case("separated, all vectors", lam, 100, rng);
    }
    {
        std::vector<double> lam(100);
        for (int i = 0; i < 100; i++) lam[i] = 4.0e9 * std::pow(10.0, -0.1 * i);  // sigma^2 of a fast-decaying spectrum
        run_case("decay over ten decades", lam, 10, rng);
        run_case("decay, k = 30", lam, 30, rng);
    }
    {
        std::vector<double> lam(100, 1.0);
        for (int i = 0; i < 100; i++) lam[i] = ud(rng);
        lam[99] = lam[98] = lam[97] = 7.5;  // a triple largest eigenvalue
        lam[96] = 7.5 - 1e-13;              // and one more inside rounding distance of it
        lam[95] = 3.0;
        lam[94] = 3.0 + 1e-9;
        run_case("repeated and clustered top", lam, 10, rng);
    }
    {
        std::vector<double> lam(60, 0.0);
        for (int i = 0; i < 12; i++) lam[i] = 5.0 + i;  // rank 12, k reaches into the null space
        run_case("rank-deficient", lam, 20, rng);
    }
    {
        std::vector<double> lam(128);
        for (int i = 0; i < 128; i++) lam[i] = ud(rng) * 1e-3 + (i % 7 == 0 ? 50.0 * ud(rng) : 0.0);
        run_case("order 128", lam, 64, rng);
    }
    run_case("order 1", {3.5}, 1, rng);
    run_case("order 2", {2.0, 9.0}, 2, rng);
    run_case("order 3", {2.0, 9.0, 9.0}, 2, rng);
    {
        std::vector<double> lam(50);
        for (int i = 0; i < 50; i++) lam[i] = -10.0 + 0.5 * i;  // indefinite: the routine is a symmetric solver, not a PSD one
        run_case("indefinite", lam, 5, rng);
    }
    {
        std::vector<double> z(16, 0.0), ev(4), vec(16);
        EXPECT(eig_host::topk(z.data(), 4, 4, ev.data(), vec.data()));  // zero matrix
        z[5] = NAN;
        EXPECT(!eig_host::topk(z.data(), 4, 2, ev.data(), vec.data()));  // non-finite input is declined
    }
    std::printf(failures ? "%d FAILURES\n" : "ALL PASSED\n", failures);
    return failures ? 1 : 0;
}
