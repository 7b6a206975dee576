// eig_host.h -- the k largest eigenpairs of a small symmetric matrix on the host (plain C++, no CUDA: shared by dense.cu and by
// tests/cpp/eig_host_test.cpp).
//
// Where it stands on the path: the SVD of the projected block T (bk_svd.rs:134, `svddc_into`) is taken here through the Gram matrix
// G = T^T T (w x w, w = b.q = 100 at k = 10) and only the first k triplets are kept (bk_svd.rs:136-138).  cuSOLVER's syevd spends
// 2.7 ms on that 100 x 100 problem (~230 launches: the unblocked back-transformation alone is 198 of them), identical on every rank
// of a sharded run -- the largest fixed cost of a step.  Only k of the w eigenvectors are needed, so:
//   1. Householder tridiagonalisation (reflectors kept),                      O(4/3 w^3)
//   2. all eigenvalues of the tridiagonal by implicit QL without vectors,      O(w^2)
//   3. inverse iteration on the tridiagonal for the k largest (pivoted LU of T - lambda I, modified Gram-Schmidt against the vectors
//      already found, three sweeps),                                           O(k w)
//   4. back-transformation of those k vectors through the reflectors,          O(k w^2)
//   5. a check against the ORIGINAL matrix: || G v - lambda v || <= 1e-12 lambda_max and | V^T V - I | <= 1e-12.
// A failed check (or a non-finite entry, or QL not converging) returns false and the caller runs the library eigensolver instead.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace eig_host {

// G: symmetric n x n (full storage, either major).  On success: lam[0..k) the k largest eigenvalues in ASCENDING order and
// vec (column-major n x k) their orthonormal eigenvectors in the same order.
inline bool topk(const double *G, int n, int k, double *lam, double *vec) {
    if (n <= 0 || k <= 0 || k > n) return false;
    for (long i = 0; i < (long)n * n; i++)
        if (!std::isfinite(G[i])) return false;
    std::vector<double> a((size_t)n * n), d((size_t)n), e((size_t)n, 0.0), tau((size_t)n, 0.0), p((size_t)n), wv((size_t)n);
    double gmax = 0.0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            a[(size_t)i * n + j] = 0.5 * (G[(size_t)i * n + j] + G[(size_t)j * n + i]);
            gmax = std::max(gmax, std::fabs(a[(size_t)i * n + j]));
        }
    // entries are scaled to at most 1 (power of two: exact), so that sums of squares inside the QL sweeps cannot overflow
    int gexp = 0;
    if (gmax > 0.0) {
        std::frexp(gmax, &gexp);
        const double dn = std::ldexp(1.0, -gexp);
        for (auto &x : a) x *= dn;
    }
    const std::vector<double> gs(a);  // the (scaled) symmetric matrix itself, for the final check
    auto A = [&](int i, int j) -> double & { return a[(size_t)i * n + j]; };
    // ---- 1. tridiagonalisation: H_j = I - tau_j v_j v_j^T with v_j = [0.., 1 (row j+1), A(j+2.., j)]
    for (int j = 0; j + 2 < n; j++) {
        const int m = n - j - 1;  // length of the column below the diagonal
        double alpha = A(j + 1, j), xn = 0.0;
        for (int i = j + 2; i < n; i++) xn += A(i, j) * A(i, j);
        xn = std::sqrt(xn);
        if (xn == 0.0) {
            tau[j] = 0.0;
            e[j] = alpha;
            continue;
        }
        const double beta = -std::copysign(std::hypot(alpha, xn), alpha);
        tau[j] = (beta - alpha) / beta;
        const double sc = 1.0 / (alpha - beta);
        for (int i = j + 2; i < n; i++) A(i, j) *= sc;
        e[j] = beta;
        // v in wv[0..m): v[0] = 1
        wv[0] = 1.0;
        for (int i = 1; i < m; i++) wv[i] = A(j + 1 + i, j);
        // p = tau * B v over the trailing block B = A(j+1.., j+1..) (kept full and symmetric)
        // (B is symmetric: B v is accumulated as sum_c v[c] . row_c, an axpy the compiler vectorises; a dot product per row is a
        // reduction it must keep scalar)
        for (int r = 0; r < m; r++) p[r] = 0.0;
        for (int c = 0; c < m; c++) {
            const double *row = &a[(size_t)(j + 1 + c) * n + (j + 1)];
            const double vc = wv[c];
            for (int r = 0; r < m; r++) p[r] += vc * row[r];
        }
        double pv = 0.0;
        for (int r = 0; r < m; r++) {
            p[r] *= tau[j];
            pv += p[r] * wv[r];
        }
        const double al = -0.5 * tau[j] * pv;
        for (int r = 0; r < m; r++) p[r] += al * wv[r];  // w = p - (tau/2)(p.v) v
        for (int r = 0; r < m; r++) {                    // B -= v w^T + w v^T
            double *row = &a[(size_t)(j + 1 + r) * n + (j + 1)];
            const double vr = wv[r], wr = p[r];
            for (int c = 0; c < m; c++) row[c] -= vr * p[c] + wr * wv[c];
        }
    }
    for (int i = 0; i < n; i++) d[i] = A(i, i);
    if (n >= 2) e[n - 2] = A(n - 1, n - 2);
    // ---- 2. eigenvalues of the tridiagonal (d, e): implicit QL, no vectors (EISPACK tql1 / Numerical Recipes tqli without z)
    std::vector<double> dd(d), ee(e);
    ee[n - 1] = 0.0;
    for (int l = 0; l < n; l++) {
        int iter = 0, mm;
        do {
            for (mm = l; mm < n - 1; mm++) {
                const double s = std::fabs(dd[mm]) + std::fabs(dd[mm + 1]);
                if (std::fabs(ee[mm]) <= 2.3e-16 * s) break;
            }
            if (mm != l) {
                if (iter++ == 60) return false;
                double g = (dd[l + 1] - dd[l]) / (2.0 * ee[l]);
                double r = std::sqrt(g * g + 1.0);
                g = dd[mm] - dd[l] + ee[l] / (g + std::copysign(r, g));
                double s = 1.0, c = 1.0, pp = 0.0;
                int i;
                for (i = mm - 1; i >= l; i--) {
                    double f = s * ee[i];
                    const double b = c * ee[i];
                    r = std::sqrt(f * f + g * g);
                    ee[i + 1] = r;
                    if (r == 0.0) {
                        dd[i + 1] -= pp;
                        ee[mm] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = dd[i + 1] - pp;
                    r = (dd[i] - g) * s + 2.0 * c * b;
                    pp = s * r;
                    dd[i + 1] = g + pp;
                    g = c * r - b;
                }
                if (r == 0.0 && i >= l) continue;
                dd[l] -= pp;
                ee[l] = g;
                ee[mm] = 0.0;
            }
        } while (mm != l);
    }
    std::sort(dd.begin(), dd.end());  // ascending
    double tnorm = 0.0;              // 1-norm of the tridiagonal
    for (int i = 0; i < n; i++) tnorm = std::max(tnorm, std::fabs(d[i]) + (i ? std::fabs(e[i - 1]) : 0.0) + (i + 1 < n ? std::fabs(e[i]) : 0.0));
    if (!(tnorm > 0.0)) {  // the zero matrix: any orthonormal set
        for (int i = 0; i < k; i++) {
            lam[i] = 0.0;
            for (int r = 0; r < n; r++) vec[(size_t)i * n + r] = r == (n - k + i) ? 1.0 : 0.0;
        }
        return true;
    }
    // ---- 3. inverse iteration for the k largest, from the largest down; zs[i] = vector of dd[n-1-i]
    std::vector<double> zs((size_t)k * n), du((size_t)n), dl((size_t)n), dg((size_t)n), du2((size_t)n), x((size_t)n);
    std::vector<int> piv((size_t)n);
    const double eps = 2.220446049250313e-16, pert = 4.0 * eps * tnorm;
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    double last = 0.0;
    for (int i = 0; i < k; i++) {
        double lm = dd[n - 1 - i];
        // eigenvalues closer than the perturbation get separated shifts (as dstein does), so their iterations differ
        if (i > 0 && last - lm < pert) lm = last - pert;
        last = lm;
        // LU of T - lm I with partial pivoting (dgttrf): dg diagonal, du / du2 the two super-diagonals, dl multipliers
        for (int r = 0; r < n; r++) {
            dg[r] = d[r] - lm;
            du[r] = r + 1 < n ? e[r] : 0.0;
            dl[r] = r + 1 < n ? e[r] : 0.0;
            du2[r] = 0.0;
        }
        for (int r = 0; r + 1 < n; r++) {
            if (std::fabs(dg[r]) >= std::fabs(dl[r])) {
                piv[r] = 0;
                if (dg[r] == 0.0) dg[r] = pert;  // exactly singular pivot
                const double f = dl[r] / dg[r];
                dl[r] = f;
                dg[r + 1] -= f * du[r];
            } else {  // swap rows r and r + 1
                piv[r] = 1;
                const double f = dg[r] / dl[r];
                dg[r] = dl[r];
                dl[r] = f;
                const double t = du[r];
                du[r] = dg[r + 1];
                dg[r + 1] = t - f * du[r];
                if (r + 2 < n) {
                    du2[r] = du[r + 1];
                    du[r + 1] = -f * du2[r];
                }
            }
        }
        // a vanishing pivot (lm is an eigenvalue to working accuracy: that is the point) is replaced by one of size eps ||T||: the
        // perturbation is harmless to inverse iteration and keeps the back substitution finite
        for (int r = 0; r < n; r++)
            if (std::fabs(dg[r]) < eps * tnorm) dg[r] = std::copysign(eps * tnorm, dg[r]);
        double *z = &zs[(size_t)i * n];
        for (int r = 0; r < n; r++) {  // deterministic start vector, different for every i
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            x[r] = 0.5 + (double)(rng >> 11) * (1.0 / 9007199254740992.0);
        }
        for (int sweep = 0; sweep < 4; sweep++) {
            // orthogonalise the right-hand side against the vectors already found, scale
            for (int q = 0; q < i; q++) {
                const double *zq = &zs[(size_t)q * n];
                double dot = 0.0;
                for (int r = 0; r < n; r++) dot += zq[r] * x[r];
                for (int r = 0; r < n; r++) x[r] -= dot * zq[r];
            }
            double nx = 0.0;
            for (int r = 0; r < n; r++) nx = std::max(nx, std::fabs(x[r]));
            if (!(nx > 0.0) || !std::isfinite(nx)) return false;
            for (int r = 0; r < n; r++) x[r] /= nx;
            // forward: apply the row interchanges and multipliers
            for (int r = 0; r + 1 < n; r++) {
                if (piv[r]) {
                    const double t = x[r];
                    x[r] = x[r + 1];
                    x[r + 1] = t - dl[r] * x[r];
                } else {
                    x[r + 1] -= dl[r] * x[r];
                }
            }
            // backward: U x = y
            x[n - 1] /= dg[n - 1];
            if (n >= 2) x[n - 2] = (x[n - 2] - du[n - 2] * x[n - 1]) / dg[n - 2];
            for (int r = n - 3; r >= 0; r--) x[r] = (x[r] - du[r] * x[r + 1] - du2[r] * x[r + 2]) / dg[r];
        }
        for (int q = 0; q < i; q++) {
            const double *zq = &zs[(size_t)q * n];
            double dot = 0.0;
            for (int r = 0; r < n; r++) dot += zq[r] * x[r];
            for (int r = 0; r < n; r++) x[r] -= dot * zq[r];
        }
        double nrm = 0.0, big = 0.0;
        for (int r = 0; r < n; r++) big = std::max(big, std::fabs(x[r]));
        if (!(big > 0.0) || !std::isfinite(big)) return false;
        for (int r = 0; r < n; r++) {
            x[r] /= big;
            nrm += x[r] * x[r];
        }
        nrm = std::sqrt(nrm);
        for (int r = 0; r < n; r++) z[r] = x[r] / nrm;
    }
    // ---- 4. back-transformation v = H_0 H_1 ... H_{n-3} z, and 5. the check against the original matrix
    double lmax = std::max(std::fabs(dd[n - 1]), std::fabs(dd[0]));
    for (int i = 0; i < k; i++) {
        double *z = &zs[(size_t)i * n];
        for (int j = n - 3; j >= 0; j--) {
            if (tau[j] == 0.0) continue;
            double s = z[j + 1];
            for (int r = j + 2; r < n; r++) s += A(r, j) * z[r];
            s *= tau[j];
            z[j + 1] -= s;
            for (int r = j + 2; r < n; r++) z[r] -= s * A(r, j);
        }
    }
    for (int i = 0; i < k; i++) {
        const double *z = &zs[(size_t)i * n];
        // Rayleigh quotient on the original matrix and the residual (G z as a sum of rows: symmetric)
        double rq = 0.0;
        for (int r = 0; r < n; r++) p[r] = 0.0;
        for (int c = 0; c < n; c++) {
            const double *row = &gs[(size_t)c * n];
            const double zc = z[c];
            for (int r = 0; r < n; r++) p[r] += zc * row[r];
        }
        for (int r = 0; r < n; r++) rq += p[r] * z[r];
        double res = 0.0;
        for (int r = 0; r < n; r++) res += (p[r] - rq * z[r]) * (p[r] - rq * z[r]);
        if (!(std::sqrt(res) <= 1.0e-12 * lmax)) return false;
        for (int q = 0; q <= i; q++) {
            const double *zq = &zs[(size_t)q * n];
            double dot = 0.0;
            for (int r = 0; r < n; r++) dot += zq[r] * z[r];
            if (!(std::fabs(dot - (q == i ? 1.0 : 0.0)) <= 1.0e-12)) return false;
        }
        // outputs ascending: position k-1-i
        lam[k - 1 - i] = std::ldexp(rq, gexp);
        for (int r = 0; r < n; r++) vec[(size_t)(k - 1 - i) * n + r] = z[r];
    }
    for (int i = 0; i + 1 < k; i++)
        if (lam[i] > lam[i + 1] + 1.0e-12 * std::ldexp(lmax, gexp)) return false;  // Rayleigh quotients out of order: something went wrong
    return true;
}

}  // namespace eig_host
