// map.cuh -- the per-nonzero value map of the normalized matrix, evaluated in the SpMM load
// (never materialised).  Restates the reference's composed MatrixMap in the same op order:
//   kind 1: x0 = v as f64; x1 = col_scale[c] * x0; x2 = log_b(x1 + 1.0); x3 = row_scale[r] * x2
//           (sqz/src/matrix_map.rs:246-252, :301-303; scan-rs/src/normalization.rs:169-177;
//            sqz/src/mat.rs:966-981)
//   kind 2: binomial deviance residual - zero term   (scan-rs/src/normalization.rs:279-299)
//   kind 3: binomial Pearson residual - zero term    (scan-rs/src/normalization.rs:338-351)
//
// The logarithm of kind 1 is the per-nonzero cost that matters (one per nonzero per pass; CUDA's
// f64 log2 is ~100 instructions).  sb_log2_ge1 is a table-driven f64 log2 for arguments >= 1
// (x1 + 1.0 always is): 7-bit table of 11-bit reciprocals, one FMA for the reduced argument, degree-8
// polynomial -- ~20 instructions, error <= ~1.5 ulp (tests/test_gpu_parity.py::test_device_log).
// The reference's libm log2/ln/log10 are themselves only specified to ~1 ulp.
#pragma once
#include "common.cuh"
#include "log_table.h"

struct MapDev {
    int kind;
    int log_base;
    const double *col;  // kind 1: col_scale[c]; kinds 2/3: n[c]
    const double *row;  // kind 1: row_scale[r] or nullptr; kinds 2/3: pi[r]
    const double *l1;      // kind 1: L_c(1) = log_b(col_scale[c] + 1) per cell (the value of a count of 1); else nullptr
    const double *inv_l1;  // kind 1: 1 / L_c(1) (0 where that is not finite); else nullptr
};

struct __align__(16) LogEnt {
    double R;
    double l2c;
};
__device__ const LogEnt sb_log_table[128] = SB_LOG_TABLE_INIT;

// log2(y) for finite y >= 1
__device__ __forceinline__ double sb_log2_ge1(double y, const LogEnt *__restrict__ tab) {
    const int hi = __double2hiint(y), lo = __double2loint(y);
    const int e = (hi >> 20) - 1023;
    const LogEnt t = tab[(hi >> 13) & 127];
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
    const double r = fma(m, t.R, -1.0);
    double p = SB_LOG2_C8;
    p = fma(p, r, SB_LOG2_C7);
    p = fma(p, r, SB_LOG2_C6);
    p = fma(p, r, SB_LOG2_C5);
    p = fma(p, r, SB_LOG2_C4);
    p = fma(p, r, SB_LOG2_C3);
    p = fma(p, r, SB_LOG2_C2);
    p = fma(p, r, SB_LOG2_C1);
    return (double)e + fma(r, p, t.l2c);
}

// the cell-dependent part of kind 1: log_b(cs * v + 1.0) (or cs * v when there is no log)
__device__ __forceinline__ double map_log_part(int log_base, double cs, u32 v, const LogEnt *__restrict__ tab) {
    const double x = cs * (double)v;
    if (log_base == 0) return x;
    const double y = x + 1.0;
    if (!(y >= 1.0 && y < 1.0e300)) {  // inf / nan / negative scale: defer to the library
        return log_base == 2 ? log2(y) : log_base == 1 ? log(y) : log10(y);
    }
    const double l2 = sb_log2_ge1(y, tab);
    switch (log_base) {
    case 2: return l2;
    case 1: return l2 * 0x1.62e42fefa39efp-1;   // ln 2
    default: return l2 * 0x1.34413509f79ffp-2;  // log10 2
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
__device__ __forceinline__ double map_full(const MapDev &mp, u32 v, u32 g, double cp, bool with_row, const LogEnt *__restrict__ tab) {
    if (mp.kind == 1) {
        double x = map_log_part(mp.log_base, cp, v, tab);
        if (with_row && mp.row) x = mp.row[g] * x;
        return x;
    }
    double pi = mp.row[g];
    return mp.kind == 2 ? map_binom_dev((double)v, cp, pi) : map_binom_pearson((double)v, cp, pi);
}
