// map.cuh -- the per-nonzero value map of the normalized matrix, evaluated in the SpMM load
// (never materialised).  Restates the reference's composed MatrixMap in the same op order:
//   kind 1: x0 = v as f64; x1 = col_scale[c] * x0; x2 = log_b(x1 + 1.0); x3 = row_scale[r] * x2
//           (sqz/src/matrix_map.rs:246-252, :301-303; scan-rs/src/normalization.rs:169-177;
//            sqz/src/mat.rs:966-981)
//   kind 2: binomial deviance residual - zero term   (scan-rs/src/normalization.rs:279-299)
//   kind 3: binomial Pearson residual - zero term    (scan-rs/src/normalization.rs:338-351)
#pragma once
#include "common.cuh"

struct MapDev {
    int kind;
    int log_base;
    const double *col;  // kind 1: col_scale[c]; kinds 2/3: n[c]
    const double *row;  // kind 1: row_scale[r] or nullptr; kinds 2/3: pi[r]
};

// the cell-dependent part of kind 1: log_b(cs * v + 1.0) (or cs * v when there is no log)
__device__ __forceinline__ double map_log_part(int log_base, double cs, u32 v) {
    double x = cs * (double)v;
    switch (log_base) {
    case 2: return log2(x + 1.0);
    case 1: return log(x + 1.0);
    case 10: return log10(x + 1.0);
    default: return x;
    }
}

__device__ __forceinline__ double a_ln_a_over_b(double a, double b) { return a == 0.0 ? 0.0 : a * log(a / b); }

__device__ __forceinline__ double map_binom_dev(double v, double n, double pi) {
    double mu = n * pi;
    double d = v - mu;
    double sign = (d != d) ? d : (signbit(d) ? -1.0 : 1.0);
    double t = 2.0 * (a_ln_a_over_b(v, mu) + a_ln_a_over_b(n - v, n - mu));
    double residual = sign * sqrt(fmax(t, 0.0));
    double zero_term = -(sqrt(2.0 * n * log(1.0 / (1.0 - pi))));
    return residual - zero_term;
}

__device__ __forceinline__ double map_binom_pearson(double v, double n, double pi) {
    double mu = n * pi;
    double residual = (v - mu) / sqrt(mu * (1.0 - pi));
    double zero_term = -sqrt(n * pi / (1.0 - pi));
    return residual - zero_term;
}

// full map value for (gene g, cell parameter cp = col[c]); `with_row` folds row_scale in
__device__ __forceinline__ double map_full(const MapDev &mp, u32 v, u32 g, double cp, bool with_row) {
    if (mp.kind == 1) {
        double x = map_log_part(mp.log_base, cp, v);
        if (with_row && mp.row) x = mp.row[g] * x;
        return x;
    }
    double pi = mp.row[g];
    return mp.kind == 2 ? map_binom_dev((double)v, cp, pi) : map_binom_pearson((double)v, cp, pi);
}
