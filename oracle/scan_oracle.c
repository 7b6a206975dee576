/*
 * scan_oracle.c -- CPU restatement of the scan-rs normalize -> PCA sparse loops.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (scan_rs_b200/, include/)
 * may link, import or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker or as the
 * timed CPU stand-in.  Compile with -ffp-contract=off and without -ffast-math so the
 * arithmetic is the reference's: separate multiply and add, strict left-to-right sums.
 *
 * Layout: the count matrix is genes(rows) x cells(cols), gene-major CSR with u64
 * indptr, u32 column index, u32 count -- the lossless content of the reference's
 * AdaptiveMat (sqz/src/mat.rs:34-42); the AdaptiveVec encodings (sqz/src/vec.rs) are
 * lossless and iterate ascending, so a plain CSR visits identical values in
 * identical order.
 *
 * Every function cites the reference lines it restates (paths relative to the
 * reference checkout).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Per-nonzero value map (sqz/src/matrix_map.rs).  One struct covers the map chains
 * the normalize path can build:
 *   kind 0: v as f64                                   MatrixIntoMap   matrix_map.rs:119-130
 *   kind 1: row_scale[r] * logb(col_scale[c]*v + 1)    ComposedMap of ScaleAxis(Axis(1)) :246-252,
 *           ScalarMap log1p :301-303, ScaleAxis(Axis(0)); built by normalization.rs:169-177
 *           and mat.rs:966-981.  Stages whose pointer is NULL / base 0 are skipped.
 *   kind 2: binomial deviance residual minus its zero term   normalization.rs:279-299
 *   kind 3: binomial Pearson residual minus its zero term    normalization.rs:338-351
 * `square` applies x -> x*x after the log and before the row scale, restating
 * `self.view().apply(|x| x.powi(2))` (mat.rs:995). */
typedef struct {
    int kind;
    int log_base;            /* 0 none, 1 ln, 2 log2, 10 log10 */
    int square;
    const double *col_scale; /* [cols] or NULL */
    const double *row_scale; /* [rows] or NULL */
    const double *bn;        /* binomial: n[c] */
    const double *bpi;       /* binomial: pi[r] */
} orc_map;

static inline double a_ln_a_over_b(double a, double b) { /* normalization.rs:264-270 */
    if (a == 0.0) return 0.0;
    return a * log(a / b);
}

static inline double signum(double x) { /* f64::signum: +-1 incl. for +-0, NaN -> NaN */
    if (isnan(x)) return x;
    return signbit(x) ? -1.0 : 1.0;
}

static inline double orc_apply(const orc_map *mp, uint32_t v, uint64_t r, uint64_t c) {
    double x = (double)v;
    switch (mp->kind) {
    case 0:
        return mp->square ? x * x : x;
    case 1:
        if (mp->col_scale) x = mp->col_scale[c] * x;         /* matrix_map.rs:250 */
        if (mp->log_base == 2) x = log2(x + 1.0);             /* normalization.rs:174 */
        else if (mp->log_base == 1) x = log(x + 1.0);         /* :173 */
        else if (mp->log_base == 10) x = log10(x + 1.0);      /* :175 */
        if (mp->square) x = x * x;                            /* mat.rs:995 */
        if (mp->row_scale) x = mp->row_scale[r] * x;          /* matrix_map.rs:249 */
        return x;
    case 2: { /* normalization.rs:279-299 */
        double n = mp->bn[c], pi = mp->bpi[r], mu = n * pi;
        double sign = signum(x - mu);
        double t = 2.0 * (a_ln_a_over_b(x, mu) + a_ln_a_over_b(n - x, n - mu));
        double residual = sign * sqrt(fmax(t, 0.0));
        double zero_term = -(sqrt(2.0 * n * log(1.0 / (1.0 - pi))));
        double o = residual - zero_term;
        return mp->square ? o * o : o;
    }
    case 3: { /* normalization.rs:338-351 */
        double n = mp->bn[c], pi = mp->bpi[r], mu = n * pi;
        double residual = (x - mu) / sqrt(mu * (1.0 - pi));
        double zero_term = -sqrt(n * pi / (1.0 - pi));
        double o = residual - zero_term;
        return mp->square ? o * o : o;
    }
    }
    return x;
}

/* sum_axis::<u32> (sqz/src/mat.rs:377-406), CSR branch.  axis 0 -> per-column (cell)
 * totals, axis 1 -> per-row (gene) totals.  u32 `+=` wraps in a release build. */
void orc_sum_axis_u32(uint64_t rows, uint64_t cols, const uint64_t *indptr, const uint32_t *idx,
                      const uint32_t *val, int axis, uint32_t *out) {
    uint64_t sz = axis == 0 ? cols : rows;
    memset(out, 0, sz * sizeof(uint32_t));
    for (uint64_t r = 0; r < rows; r++)
        for (uint64_t k = indptr[r]; k < indptr[r + 1]; k++) {
            if (axis == 0) out[idx[k]] += val[k];
            else out[r] += val[k];
        }
}

/* u64 totals: hdf5-io/src/matrix.rs:106-114 (compute_genes_filter row sums) and the
 * builder-defined HVG moments: sum v and sum v*v per row / per column. */
void orc_sum_axis_u64(uint64_t rows, uint64_t cols, const uint64_t *indptr, const uint32_t *idx,
                      const uint32_t *val, int axis, int square, uint64_t *out) {
    uint64_t sz = axis == 0 ? cols : rows;
    memset(out, 0, sz * sizeof(uint64_t));
    for (uint64_t r = 0; r < rows; r++)
        for (uint64_t k = indptr[r]; k < indptr[r + 1]; k++) {
            uint64_t v = val[k];
            if (square) v = v * v;
            if (axis == 0) out[idx[k]] += v;
            else out[r] += v;
        }
}

/* sum_axis::<f64> through the map (sqz/src/mat.rs:377-406): strict storage order. */
void orc_sum_axis_map(uint64_t rows, uint64_t cols, const uint64_t *indptr, const uint32_t *idx,
                      const uint32_t *val, const orc_map *mp, int axis, double *out) {
    uint64_t sz = axis == 0 ? cols : rows;
    for (uint64_t i = 0; i < sz; i++) out[i] = 0.0;
    for (uint64_t r = 0; r < rows; r++)
        for (uint64_t k = indptr[r]; k < indptr[r + 1]; k++) {
            double x = orc_apply(mp, val[k], r, idx[k]);
            if (axis == 0) out[idx[k]] += x;
            else out[r] += x;
        }
}

/* sum_axis_exclude (sqz/src/mat.rs:729-762), CSR branch; masks replace the HashSets. */
void orc_sum_axis_exclude(uint64_t rows, uint64_t cols, const uint64_t *indptr, const uint32_t *idx,
                          const uint32_t *val, const orc_map *mp, int axis,
                          const uint8_t *excl_rows, const uint8_t *excl_cols, double *out) {
    uint64_t sz = axis == 0 ? cols : rows;
    for (uint64_t i = 0; i < sz; i++) out[i] = 0.0;
    for (uint64_t r = 0; r < rows; r++) {
        if (excl_rows[r]) continue;
        for (uint64_t k = indptr[r]; k < indptr[r + 1]; k++) {
            uint64_t c = idx[k];
            if (excl_cols[c]) continue;
            double x = orc_apply(mp, val[k], r, c);
            if (axis == 0) out[c] += x;
            else out[r] += x;
        }
    }
}

/* partition_on_thresholds fixpoint (sqz/src/mat.rs:776-802).  Fills the exclusion
 * masks; returns the number of sweeps.  has_row/has_col mirror the Option arguments. */
int orc_partition_masks(uint64_t rows, uint64_t cols, const uint64_t *indptr, const uint32_t *idx,
                        const uint32_t *val, const orc_map *mp, int has_row, double row_thr,
                        int has_col, double col_thr, uint8_t *excl_rows, uint8_t *excl_cols) {
    double *sum = (double *)malloc(sizeof(double) * (rows > cols ? rows : cols) + 8);
    memset(excl_rows, 0, rows);
    memset(excl_cols, 0, cols);
    int sweeps = 0;
    for (;;) {
        int updated = 0;
        sweeps++;
        if (has_col) {
            orc_sum_axis_exclude(rows, cols, indptr, idx, val, mp, 0, excl_rows, excl_cols, sum);
            for (uint64_t c = 0; c < cols; c++)
                if (sum[c] < col_thr && !excl_cols[c]) { excl_cols[c] = 1; updated = 1; }
        }
        if (has_row) {
            orc_sum_axis_exclude(rows, cols, indptr, idx, val, mp, 1, excl_rows, excl_cols, sum);
            for (uint64_t r = 0; r < rows; r++)
                if (sum[r] < row_thr && !excl_rows[r]) { excl_rows[r] = 1; updated = 1; }
        }
        if (!updated) break;
    }
    free(sum);
    return sweeps;
}

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

/* median_mut (scan-rs/src/stats.rs:13-38): sorts in place; even length -> midpoint in the
 * element type (u32 add then integer divide).  Returns 0 and leaves *out alone if empty. */
int orc_median_u32(uint32_t *xs, uint64_t n, uint32_t *out) {
    if (n == 0) return 0;
    qsort(xs, n, sizeof(uint32_t), cmp_u32);
    if (n % 2 == 0) *out = (uint32_t)(xs[n / 2] + xs[n / 2 - 1]) / 2u;
    else *out = xs[n / 2];
    return 1;
}

/* CSR x dense (gather): csrmat_densemat_mult + vec_mulacc_dense_rowmaj_std
 * (sqz/src/prod.rs:30-51, 124-148).  out (rows x w) must be zeroed by the caller
 * (mat.rs:1085).  `*o = *o + *r * lval`: multiply, then add. */
void orc_spmm_gather(uint64_t rows, uint64_t cols, const uint64_t *indptr, const uint32_t *idx,
                     const uint32_t *val, const orc_map *mp, const double *rhs, uint64_t w,
                     double *out) {
    (void)cols;
    for (uint64_t r = 0; r < rows; r++) {
        double *o = out + r * w;
        for (uint64_t k = indptr[r]; k < indptr[r + 1]; k++) {
            uint64_t c = idx[k];
            double lval = orc_apply(mp, val[k], r, c);
            const double *rr = rhs + c * w;
            for (uint64_t j = 0; j < w; j++) o[j] = o[j] + rr[j] * lval;
        }
    }
}

/* dense x sparse via the transposed view: Dot<AdaptiveMat> for ArrayBase (mat.rs:1114-1133)
 * -> cscmat_densemat_mult + colvec_mulacc_dense_rowmaj (prod.rs:56-81, 154-214).
 * A.t() is (cols x rows) stored CSC whose "columns" are this matrix's rows, so for each
 * gene row g with dense line y = Y[g,:] (= B[:,g] in the caller's b x m block):
 * out[c,:] += y * map(v, g, c).  out is (cols x w), zeroed by the caller (mat.rs:1128).
 * The TransposeMap swaps (r,c) back (matrix_map.rs:62-66) so the map sees (gene, cell). */
void orc_spmm_scatter(uint64_t rows, uint64_t cols, const uint64_t *indptr, const uint32_t *idx,
                      const uint32_t *val, const orc_map *mp, const double *y, uint64_t w,
                      double *out) {
    (void)cols;
    for (uint64_t g = 0; g < rows; g++) {
        const double *yy = y + g * w;
        for (uint64_t k = indptr[g]; k < indptr[g + 1]; k++) {
            uint64_t c = idx[k];
            double lval = orc_apply(mp, val[k], g, c);
            double *o = out + c * w;
            for (uint64_t j = 0; j < w; j++) o[j] = o[j] + yy[j] * lval;
        }
    }
}

/* to_dense of the mapped sparse part (mat.rs:190-204): out is rows x cols, zeroed here. */
void orc_to_dense(uint64_t rows, uint64_t cols, const uint64_t *indptr, const uint32_t *idx,
                  const uint32_t *val, const orc_map *mp, double *out) {
    for (uint64_t i = 0; i < rows * cols; i++) out[i] = 0.0;
    for (uint64_t r = 0; r < rows; r++)
        for (uint64_t k = indptr[r]; k < indptr[r + 1]; k++)
            out[r * cols + idx[k]] = orc_apply(mp, val[k], r, idx[k]);
}

/* ---- multi-threaded variants: NOT the reference's arithmetic order, only used as the
 * "all host cores" CPU timing in bench.py --impl reference (the reference path itself is
 * single-threaded: prod.rs has no rayon, MKL is -seq).  Row-parallel gather is order
 * preserving; the scatter is made cell-parallel through per-thread partial blocks. ---- */
void orc_spmm_gather_omp(uint64_t rows, uint64_t cols, const uint64_t *indptr, const uint32_t *idx,
                         const uint32_t *val, const orc_map *mp, const double *rhs, uint64_t w,
                         double *out) {
    (void)cols;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < (int64_t)rows; r++) {
        double *o = out + (uint64_t)r * w;
        for (uint64_t k = indptr[r]; k < indptr[r + 1]; k++) {
            uint64_t c = idx[k];
            double lval = orc_apply(mp, val[k], (uint64_t)r, c);
            const double *rr = rhs + c * w;
            for (uint64_t j = 0; j < w; j++) o[j] = o[j] + rr[j] * lval;
        }
    }
}

void orc_spmm_scatter_omp(uint64_t rows, uint64_t cols, const uint64_t *indptr, const uint32_t *idx,
                          const uint32_t *val, const orc_map *mp, const double *y, uint64_t w,
                          double *out) {
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    double *part = (double *)calloc((size_t)nt * cols * w, sizeof(double));
#pragma omp parallel
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        double *po = part + (size_t)t * cols * w;
#pragma omp for schedule(dynamic, 64)
        for (int64_t g = 0; g < (int64_t)rows; g++) {
            const double *yy = y + (uint64_t)g * w;
            for (uint64_t k = indptr[g]; k < indptr[g + 1]; k++) {
                uint64_t c = idx[k];
                double lval = orc_apply(mp, val[k], (uint64_t)g, c);
                double *o = po + c * w;
                for (uint64_t j = 0; j < w; j++) o[j] = o[j] + yy[j] * lval;
            }
        }
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)(cols * w); i++) {
            double s = out[i];
            for (int tt = 0; tt < nt; tt++) s += part[(size_t)tt * cols * w + i];
            out[i] = s;
        }
    }
    free(part);
}

void orc_sum_axis_map_omp(uint64_t rows, uint64_t cols, const uint64_t *indptr, const uint32_t *idx,
                          const uint32_t *val, const orc_map *mp, double *out_rows) {
    (void)cols;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < (int64_t)rows; r++) {
        double s = 0.0;
        for (uint64_t k = indptr[r]; k < indptr[r + 1]; k++) s += orc_apply(mp, val[k], (uint64_t)r, idx[k]);
        out_rows[r] = s;
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py --impl reference pins its own thread count (torchrun exports OMP_NUM_THREADS=1) */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
