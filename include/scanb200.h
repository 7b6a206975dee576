/*
 * scanb200.h -- C ABI of the B200-native normalize -> PCA path of scan-rs.
 *
 * This is the drop-in boundary: plain pointers and sizes, opaque handles,
 * int status codes, no C++/torch types.  A Rust crate binds exactly these symbols from
 * its build.rs (INTEGRATION.md shows the `extern "C"` block and the safe wrappers that
 * keep the reference's entry points: `AdaptiveMat` -> upload, `normalize(..)`,
 * `Pca::run_pca_cancellable`).  Each entry point cites the reference interface it
 * replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - Matrix orientation is the reference's: features/genes are ROWS, barcodes/cells are
 *     COLUMNS (scan-rs/src/mtx.rs:10-51).  m = genes, n = cells.
 *   - All dense blocks are row-major ("standard layout" of ndarray), f64.
 *   - Host buffers are caller-allocated and caller-owned; the library never keeps a host
 *     pointer after the call returns.  Handles own device memory.
 *   - One sb_ctx drives one GPU from one host thread.  For cell-sharded multi-GPU runs
 *     every rank (one process per GPU) creates its own context, joins a communicator
 *     with sb_comm_init and then makes the SAME sequence of calls on its own cell shard;
 *     gene-sized results (totals, U, sigma) come back replicated, cell-sized results
 *     (cell totals, V) come back for the local shard only.
 *   - Every function returns SB_OK (0) or an error code; sb_last_error() returns the
 *     thread-local message.  Nothing panics or aborts across the ABI.
 *   - There is NO CPU fallback: without a CUDA device sb_init fails with SB_ERR_CUDA.
 */
#ifndef SCANB200_H
#define SCANB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_VERSION 100

#if defined(__GNUC__)
#define SB_API __attribute__((visibility("default")))
#else
#define SB_API
#endif

/* status codes */
enum {
    SB_OK = 0,
    SB_ERR_INVALID_SHAPE = 1, /* "The input matrix must be at least 2x2."  bk_svd.rs:73-75 */
    SB_ERR_INVALID_K = 2,     /* "invalid k"                                 bk_svd.rs:77-79 */
    SB_ERR_CANCELLED = 3,     /* snoop::CancellationError                    snoop/src/lib.rs:5-18 */
    SB_ERR_CUDA = 4,
    SB_ERR_NCCL = 5,
    SB_ERR_OOM = 6,
    SB_ERR_INVALID_ARG = 7,   /* the reference's assert!/panic! programmer errors */
    SB_ERR_UNSUPPORTED = 8,
    SB_ERR_LINALG = 9         /* cuSOLVER failure (the reference's LAPACK `?` errors) */
};

/* storage order of host CSR/CSC arrays handed to sb_upload (sprs::CSR / sprs::CSC of a
 * genes x cells matrix, sqz/src/mat.rs:45-65) */
enum { SB_GENE_MAJOR = 0, SB_CELL_MAJOR = 1 };

/* scan_rs::normalization::Normalization (scan-rs/src/normalization.rs:11-28) */
enum {
    SB_NORM_CELLRANGER = 0,
    SB_NORM_CELLRANGER8 = 1,
    SB_NORM_SEURATLOG = 2,
    SB_NORM_BINOMIAL_DEVIANCE = 3,
    SB_NORM_BINOMIAL_PEARSON = 4,
    SB_NORM_WITH_SIZE_FACTORS = 5,
    SB_NORM_LOG_TRANSFORM = 6
};

/* scan_rs::normalization::LogBase (normalization.rs:105-112) */
enum { SB_LOG_E = 1, SB_LOG_TWO = 2, SB_LOG_TEN = 10 };

typedef struct sb_ctx sb_ctx;   /* device context (+ optional NCCL communicator) */
typedef struct sb_mat sb_mat;   /* device-resident u32 count matrix: AdaptiveMat<u32> (sqz/src/mat.rs:34-42) */
typedef struct sb_nmat sb_nmat; /* normalized matrix: LowRankOffset<D, impl MatrixMap<u32,f64>> (sqz/src/low_rank_offset.rs:12-16) */

/* progress / cancel token: CancelProgress::set_progress_check (snoop/src/lib.rs:45-57).
 * Called on the calling thread with the reference's milestones (0.8*i/n_iter, 0.82, 0.93,
 * 1.0; bk_svd.rs:96-114).  A non-zero return cancels: the call stops at that milestone and
 * returns SB_ERR_CANCELLED. */
typedef int (*sb_progress_cb)(double fraction, void *user);

/* ---------------------------------------------------------------- context */
SB_API int sb_version(void);
SB_API const char *sb_last_error(void);
SB_API int sb_init(int device, sb_ctx **out);
SB_API void sb_shutdown(sb_ctx *ctx);
/* Multi-GPU: rank 0 calls sb_comm_unique_id, ships the 128 bytes to the other ranks by any
 * means (torch.distributed, MPI, a file), then every rank calls sb_comm_init. */
SB_API int sb_comm_unique_id(char id[128]);
SB_API int sb_comm_init(sb_ctx *ctx, int nranks, int rank, const char id[128]);
SB_API int sb_sync(sb_ctx *ctx);
/* One host thread, every GPU of the box (the reference's caller is a single process, tools/src/bin/cmd.rs:67-81).
 * sb_multi_init creates one context per device (devices == NULL: 0..n-1), joins them to one communicator and starts one worker
 * thread per device.  sb_multi_run executes fn(rank, ctx, user) on every worker concurrently and returns when all are done (the
 * first non-zero status, with "rank r: message" as the error): inside fn each rank makes the ordinary single-context calls on
 * its own cell shard, exactly the sequence one-process-per-GPU callers make.  Handles created inside fn stay valid across
 * sb_multi_run calls and must be freed before sb_multi_shutdown. */
typedef struct sb_multi sb_multi;
typedef int (*sb_rank_fn)(int rank, sb_ctx *ctx, void *user);
SB_API int sb_multi_init(int n, const int *devices, sb_multi **out);
SB_API int sb_multi_size(const sb_multi *mm);
SB_API int sb_multi_ctx(sb_multi *mm, int rank, sb_ctx **out);
SB_API int sb_multi_run(sb_multi *mm, sb_rank_fn fn, void *user);
SB_API void sb_multi_shutdown(sb_multi *mm);
/* Options: "direct_projection" (0/1, default 0): 1 forces the wide projection pass T = Q^T A of
 * bk_svd.rs:102,131 to run as a sparse product; 0 lets the library use Q^T A = R^-T (K^T A) when R is usable.
 * "gather" (default 1): 1 = panelled gather kernels for the sparse halves of both products, 0 = the first-generation
 * cell-major / gene-major kernels (A/B and accuracy reference).  "dense_genes" (default 2048) / "dense_min_density"
 * (default 0.12): size of the dense hot-gene panel of matrices uploaded afterwards (0 disables it).
 * "upload_sync" (default 1): drain the build stream after every stage of a pipelined upload.
 * "overlap" / "overlap_t" (default 0): experimental concurrent sparse + panel kernels.
 * "dense_max_count" (default 15, 1..15): largest count kept in the dense panel of matrices uploaded afterwards.
 * "gather_split" (default 0): EXPERIMENTAL separate 4-byte stream for the sparse entries with a count of 1 (gather_split.cu);
 * written without hardware access, off until validated.
 * "panel_i8" (default 0): EXPERIMENTAL tcgen05 int8 contraction of the T-side panel (needs dense_max_count <= 3, a
 * 2,048-gene panel and width <= 20); written without hardware access, off until validated. */
SB_API int sb_set_option(sb_ctx *ctx, const char *name, double value);

/* Page-locked host memory for inputs / outputs (optional): copies to and from pinned buffers run at PCIe
 * speed and asynchronously; pageable buffers work too, several times slower for the large V block. */
SB_API int sb_host_alloc(size_t bytes, void **out);
SB_API void sb_host_free(void *p);

/* ---------------------------------------------------------------- count matrix
 * sb_upload replaces AdaptiveMat::from_csmat / new (sqz/src/mat.rs:81-124): the Rust side
 * decodes each AdaptiveVec once (vec.rs `foreach` :1230-1273) into these arrays.  `major`
 * says whether indptr runs over genes (idx = cell) or over cells (idx = gene).  n_local is
 * this rank's number of cells.  Indices must be strictly ascending inside each vector and
 * counts non-zero (what AdaptiveVec guarantees); zeros are dropped. */
SB_API int sb_upload(sb_ctx *ctx, int major, uint32_t m, uint64_t n_local, const uint64_t *indptr,
              const uint32_t *idx, const uint32_t *cnt, sb_mat **out);
/* The reference's own storage as upload input: one AdaptiveVec (sqz/src/vec.rs:1029-1053) described by the raw parts of its Rust
 * struct.  variant follows the enum order: 0 D3, 1 D4, 2 D8, 3 D16, 4 V, 5 S3, 6 S4, 7 S8.
 *   D3/D4/D8/D16 : dense = Dense3.data (u64 words, 21 x 3 bits) / Dense4.data (bytes, 2 x 4 bits) / DenseW.data (u8 / u16), one code
 *                  per position; fb_* = the SimpleSparse fallback holding the values the narrow code cannot (vec.rs:660-1023)
 *   V            : fb_* = SimpleSparse.indexes / values (vec.rs:123-127); dense unused
 *   S3/S4/S8     : CompressedIndexSparse (vec.rs:222-227): index_bytes[n_index] (position inside a block of 256), block_starts
 *                  [n_block_starts]; dense / fb_* describe its dense_data (a D3 / D4 / D8 vector of length n_index)
 * len = the vector's logical length.  All arrays stay owned by the caller. */
typedef struct sb_adaptive_vec {
    uint32_t variant;
    uint32_t reserved;
    uint64_t len;
    const void *dense;
    const uint32_t *fb_idx;
    const uint32_t *fb_val;
    uint64_t fb_len;
    const uint8_t *index_bytes;
    uint64_t n_index;
    const uint32_t *block_starts;
    uint64_t n_block_starts;
} sb_adaptive_vec;
/* AdaptiveVec::foreach (vec.rs:1230-1273) for one vector: (index, value) ascending, stored zeros skipped.  idx / val may be NULL
 * to count only. */
SB_API int sb_adaptive_decode_vec(const sb_adaptive_vec *vec, uint32_t *idx, uint32_t *val, uint64_t capacity, uint64_t *nnz);
/* AdaptiveMat (sqz/src/mat.rs:34-42) -> device matrix: decodes the m (SB_GENE_MAJOR, the reference's CSR storage) or n_local
 * (SB_CELL_MAJOR) vectors on `threads` host threads (0 = all) and uploads. */
SB_API int sb_upload_adaptive(sb_ctx *ctx, int major, uint32_t m, uint64_t n_local, const sb_adaptive_vec *vecs, int threads,
                       sb_mat **out);
/* The same constructor for a cell-major matrix in a narrow host form, for m <= 65536: u16 gene index + u8 count per entry
 * (3 bytes over PCIe instead of 8 -- the host-to-device copy is the largest part of an end-to-end call).  Counts >= 255
 * are written as 255 in cnt8 and listed in the side arrays: entry big_pos[i] (ascending stream positions) has count
 * big_cnt[i].  The Rust side fills these from the same AdaptiveVec `foreach` walk (vec.rs:1230-1273) as sb_upload's. */
SB_API int sb_upload_compact(sb_ctx *ctx, uint32_t m, uint64_t n_local, const uint64_t *indptr, const uint16_t *idx16,
                      const uint8_t *cnt8, uint64_t n_big, const uint64_t *big_pos, const uint32_t *big_cnt, sb_mat **out);
/* The same constructor in the packed host form -- about 1.55 bytes per entry over PCIe (the compact form: 3, the plain one: 8),
 * for any m the library accepts.  Per entry, in the stream order of sb_upload's cell-major arrays:
 *   dgene[j]  : gene - previous gene of the same cell (the previous gene of a cell's first entry is -1), if that is in 1..255;
 *               else 0, and (j, gene) is appended to esc_pos / esc_gene (stream positions ascending)
 *   cnt4[j/2] : count in the low (j even) or high (j odd) nibble, if below 15; else 15, and (j, count) is appended to big_pos /
 *               big_cnt (ascending).  cnt4 has (nnz + 1) / 2 bytes.
 * The Rust side fills these from the same AdaptiveVec `foreach` walk (vec.rs:1230-1273); sb_pack_csc_count / sb_pack_csc_fill do it
 * from plain cell-major u32 arrays on `threads` host threads (0 = all): _count returns the two side-list lengths, _fill writes
 * every array (all caller-allocated; page-locked memory from sb_host_alloc makes the upload a straight DMA).  A zero delta without
 * an escape record, or an escaped gene that does not ascend inside its cell, is SB_ERR_INVALID_ARG. */
SB_API int sb_pack_csc_count(uint64_t n, const uint64_t *indptr, const uint32_t *idx, const uint32_t *cnt, int threads, uint64_t *n_esc,
                      uint64_t *n_big);
SB_API int sb_pack_csc_fill(uint64_t n, const uint64_t *indptr, const uint32_t *idx, const uint32_t *cnt, int threads, uint8_t *dgene,
                     uint8_t *cnt4, uint64_t *esc_pos, uint32_t *esc_gene, uint64_t *big_pos, uint32_t *big_cnt);
SB_API int sb_upload_packed(sb_ctx *ctx, uint32_t m, uint64_t n_local, const uint64_t *indptr, const uint8_t *dgene, const uint8_t *cnt4,
                     uint64_t n_esc, const uint64_t *esc_pos, const uint32_t *esc_gene, uint64_t n_big, const uint64_t *big_pos,
                     const uint32_t *big_cnt, sb_mat **out);
/* The loaders' device half.  sb_upload_unsorted: cell-major arrays whose gene indices are in ANY order inside
 * a cell -- the Cell Ranger 3 defect that hdf5-io/src/matrix.rs:63-78 repairs with `new_from_unsorted_csc`; every cell's entries
 * are sorted on the device, a duplicate index inside a cell is SB_ERR_INVALID_ARG.  sb_filter_genes: compute_genes_filter + the
 * row selection of read_adaptive_csr_matrix (matrix.rs:93-192): type_keep[m] (NULL = all) marks the features whose type matches
 * `retain_feature_like`, min_total is `shrink_row`; kept_rows (room for m) receives the survivors in file order. */
SB_API int sb_upload_unsorted(sb_ctx *ctx, uint32_t m, uint64_t n_local, const uint64_t *indptr, const uint32_t *idx,
                       const uint32_t *cnt, sb_mat **out);
SB_API int sb_filter_genes(sb_mat *mat, const uint8_t *type_keep, uint64_t min_total, uint32_t *kept_rows, uint32_t *n_kept,
                    sb_mat **out);
/* rows(), cols(), shape() (mat.rs:160-176) + nnz() (:155-157); n_global == n_local without a communicator */
SB_API int sb_mat_shape(const sb_mat *mat, uint32_t *m, uint64_t *n_local, uint64_t *n_global, uint64_t *nnz_local);
/* to_csmat (mat.rs:207-239): sizes from sb_mat_shape; indptr has m+1 or n_local+1 entries */
SB_API int sb_download(const sb_mat *mat, int major, uint64_t *indptr, uint32_t *idx, uint32_t *cnt);
SB_API void sb_free_mat(sb_mat *mat);

/* sum_axis::<u32>(Axis(0)) (mat.rs:377-406; normalization.rs:159,161): per-cell UMI totals, wrapping u32 */
SB_API int sb_cell_totals(sb_mat *mat, uint32_t *out_n_local);
/* per-gene totals as u64 (hdf5-io/src/matrix.rs:106-114; sum_axis(Axis(1))); all-reduced over ranks.
 * square != 0 gives sum of v*v (HVG moments, builder-defined) */
SB_API int sb_gene_totals(sb_mat *mat, int square, uint64_t *out_m);
/* number of stored non-zeros per gene, all-reduced */
SB_API int sb_gene_nnz(sb_mat *mat, uint64_t *out_m);
/* median_mut over the cell totals of ALL ranks (scan-rs/src/stats.rs:13-38): u32 midpoint.
 * *nonempty = 0 when there are no cells (the caller then uses 1.0, normalization.rs:166) */
SB_API int sb_median_cell_total(sb_mat *mat, uint32_t *median, int *nonempty);

/* partition_on_thresholds (mat.rs:772-889).  `rows_out`/`cols_out` need room for m / n_local entries and receive the
 * selected indices; kept/residual may be NULL.  Sharded contexts: gene sums are all-reduced every round, cell verdicts stay
 * local; cols_out are LOCAL cell indices and the returned matrices hold this rank's shard. */
SB_API int sb_partition(sb_mat *mat, int has_row_thr, double row_thr, int has_col_thr, double col_thr,
                 sb_mat **kept, sb_mat **residual, uint64_t *rows_out, uint64_t *n_rows_out,
                 uint64_t *cols_out, uint64_t *n_cols_out);
/* select_rows / select_cols (mat.rs:1004-1071): new matrix with the given rows/cols in the given order (sharded contexts:
 * every rank passes the same rows; cols are LOCAL cell indices of its own shard) */
SB_API int sb_select_rows(sb_mat *mat, const uint32_t *rows, uint32_t count, sb_mat **out);
SB_API int sb_select_cols(sb_mat *mat, const uint64_t *cols, uint64_t count, sb_mat **out);
/* Highly-variable-gene selection (NOT in the reference; builder-defined): exact u64
 * sums -> dispersion var/mean in f64 -> top n_top, ties to the lower index; out sorted ascending. */
SB_API int sb_hvg_select(sb_mat *mat, uint32_t n_top, uint32_t *out_idx, uint32_t *out_count);

/* ---------------------------------------------------------------- normalization
 * normalize / normalize_with_size_factor (normalization.rs:46-102) plus the two binomial
 * residual constructors the CLI dispatches to (normalization.rs:233-260, 307-323;
 * tools/src/bin/cmd.rs:67-81).  size_factors (u32[n_local]) only for SB_NORM_WITH_SIZE_FACTORS.
 * The result borrows `mat`: free it before the matrix. */
SB_API int sb_normalize(sb_mat *mat, int norm, const uint32_t *size_factors, sb_nmat **out);
/* log_normalize_with_size_factor (+ optional scale_and_center) with every knob exposed
 * (normalization.rs:138-178; mat.rs:986-1001): has_target=0 -> median of cell totals;
 * center_scale: 0 = sparse log-normalized matrix only, 1 = scale_and_center(Axis(1), None),
 * 2 = scale_and_center(Axis(1), Some(sd_override[m])). */
SB_API int sb_log_normalize(sb_mat *mat, int has_target, double target, int log_base, const uint32_t *size_factors,
                     int center_scale, const double *sd_override, sb_nmat **out);
/* log1p_normalize_fixed_point (normalization.rs:191-213) */
SB_API int sb_normalize_fixed_point(sb_mat *mat, int log_base, uint32_t base, uint32_t exponent, sb_nmat **out);
/* The derived tables: col_scale[n_local], row_scale[m] (1/sd), u[m], v[n_local]; any pointer may be NULL */
SB_API int sb_nmat_params(const sb_nmat *a, double *col_scale, double *row_scale, double *u, double *v);
/* LowRankOffset::to_dense (low_rank_offset.rs:55-57): m x n_local row-major; test/debug sizes only */
SB_API int sb_nmat_to_dense(sb_nmat *a, double *out);
/* LowRankOffset . Array2 (low_rank_offset.rs:68-81): out[m x w] = A . x[n_local x w], all-reduced over ranks */
SB_API int sb_nmat_dot(sb_nmat *a, const double *x, uint32_t w, double *out);
/* Array2 . LowRankOffset (low_rank_offset.rs:83-96): out[w x n_local] = b[w x m] . A */
SB_API int sb_nmat_rdot(sb_nmat *a, const double *b, uint32_t w, double *out);
/* Squared Frobenius norm of the normalized matrix (offset included): the denominator of the DERIVED "variance explained"
 * sigma_i^2 / ||A||_F^2.  Not part of the reference's result (PcaResult is (u, d, v), dim_red/mod.rs:47); north_star names it,
 * so it is offered as a labelled convenience.  Log-chain normalizations only (SB_ERR_UNSUPPORTED otherwise). */
SB_API int sb_nmat_frobenius_sq(sb_nmat *a, double *out);
SB_API void sb_free_nmat(sb_nmat *a);

/* ---------------------------------------------------------------- PCA
 * The start block: SmallRng::seed_from_u64(seed) + Uniform::new(-1.0, 1.0), row-major fill
 * (bk_svd.rs:83-84, :90, :118).  Host-side, deterministic. */
SB_API int sb_omega(uint64_t seed, uint64_t rows, uint64_t cols, double *out);

/* svd_bk (scan-rs/src/dim_red/bk_svd.rs:57-146).  b is clamped to min(m, n, b) (:81).  `omega`
 * may be NULL (generated from `seed` as above) or the caller's start block, row-major:
 * (b x m) when n > m, (n_local x b) when m >= n.  Outputs: U[m x k], S[k], V[n_local x k]
 * row-major -- V is already `vt.reversed_axes()` as run_pca returns it (bk_svd.rs:51). */
SB_API int sb_bksvd(sb_nmat *a, uint32_t k, uint32_t b, uint32_t n_iter, uint64_t seed, const double *omega,
             sb_progress_cb cb, void *user, double *U, double *S, double *V);
/* BkSvd::run_pca_cancellable (bk_svd.rs:48-52): b = ceil(k * k_multiplier), seed 0 */
SB_API int sb_bksvd_run_pca(sb_nmat *a, uint32_t k, double k_multiplier, uint32_t n_iter, sb_progress_cb cb,
                     void *user, double *U, double *S, double *V);
/* Diagnostics of the last sb_bksvd on this context: the condition estimate of the Krylov basis' triangular factor that gated the
 * projection identity (0 when the direct pass ran unconditionally), the a-posteriori probe residual in units of sigma_1 (0 when
 * no check ran) and how many fallbacks (Householder QR after a CholeskyQR breakdown, direct pass after a failed check) were taken. */
SB_API int sb_pca_diagnostics(sb_ctx *ctx, double *cond_r, double *probe_resid, int *fallbacks);
/* svd_rand (scan-rs/src/dim_red/rand_svd.rs:54-129); omega: (l x m) when n > m, (n_local x l) when m >= n */
SB_API int sb_randsvd(sb_nmat *a, uint32_t k, uint32_t l, uint32_t n_iter, uint64_t seed, const double *omega,
               double *U, double *S, double *V);
/* RandSvd::run_pca_cancellable (rand_svd.rs:44-49): l = max(k + 4, floor(k * l_multiplier)), seed 0 */
SB_API int sb_randsvd_run_pca(sb_nmat *a, uint32_t k, double l_multiplier, uint32_t n_iter, double *U, double *S,
                       double *V);

/* irlba (scan-rs/src/dim_red/irlba.rs:71-215): implicitly restarted Lanczos bidiagonalisation, nu triplets, tolerance tol
 * (Irlba::new: 1e-4), at most maxit restarts (50); working dimension m_b = min(nu + 20, 3 nu, n) (:87).  v0[n_local] = this rank's
 * slice of the start vector (normalised inside), or NULL for sb_irlba_start(0, n_global).  The reference draws it from rand_distr's
 * Normal on SmallRng seed 0, a third-party stream that cannot be restated offline (parity unpinned, like Omega).  The
 * `resid[i] < tol * smax` test carries no absolute value in the reference (:176-181); it is kept, with sign-canonical singular
 * vectors of the small matrix (largest-magnitude entry of each right vector positive).  Progress milestones it / maxit (:206).
 * Outputs U[m x nu], S[nu], V[n_local x nu] row-major; mprod_out / iters_out (optional): matrix products, restarts taken. */
SB_API int sb_irlba(sb_nmat *a, uint32_t nu, double tol, uint32_t maxit, const double *v0, sb_progress_cb cb, void *user,
             double *U, double *S, double *V, uint32_t *mprod_out, uint32_t *iters_out);
/* The builder-defined default start vector of sb_irlba: Box-Muller on the Xoshiro256++ stream of sb_omega, entry i for global cell i. */
SB_API int sb_irlba_start(uint64_t seed, uint64_t n, double *out);

/* ---------------------------------------------------------------- moment consumers (diff-exp's reads of the count matrix)
 * mean_var_axis (sqz/src/mat.rs:285-329): V[X] = E[X^2] - E[X]^2 along an axis.  axis 1: per gene over all cells (m values,
 * summed over ranks); axis 0: per local cell over the m genes.  cell_div (n_local values or NULL): the SizeNormalized view
 * `v as f64 / size_factor[c]` (diff-exp/src/diff_exp.rs:340-358; NaN factors become 0).  Without cell_div the sums are exact. */
SB_API int sb_mean_var_axis(sb_mat *mat, int axis, const double *cell_div, double *mean, double *var);
/* mean_var_rows (mat.rs:332-374): per gene over the listed cells (LOCAL indices of this rank's shard; a cell listed twice counts
 * twice, as the reference's CSC branch); the divisor is the number of listed cells over all ranks. */
SB_API int sb_mean_var_rows(sb_mat *mat, const uint64_t *cols, uint64_t n_cols, const double *cell_div, double *mean, double *var);
/* sum_rows / sum_cols / sum_rows_dual (mat.rs:414-476, 484-583) as exact u64 sums: per gene over the listed cells, per listed cell
 * over the genes, per gene over two lists at once (a cell in both lists adds to both). */
SB_API int sb_sum_rows(sb_mat *mat, const uint64_t *cols, uint64_t n_cols, uint64_t *out_m);
SB_API int sb_sum_cols(sb_mat *mat, const uint64_t *cols, uint64_t n_cols, uint64_t *out);
SB_API int sb_sum_rows_dual(sb_mat *mat, const uint64_t *cols1, uint64_t n1, const uint64_t *cols2, uint64_t n2, uint64_t *sum1,
                     uint64_t *sum2);
/* size_factors (diff_exp.rs:314-334): counts per cell / their median (50th percentile with linear interpolation,
 * diff-exp/src/stat.rs:107-163), over the listed local cells (NULL: all) of all ranks; umi_counts (optional) replaces the totals,
 * one value per listed cell.  out[n_local]: 0 at unlisted cells. */
SB_API int sb_size_factors(sb_mat *mat, const uint64_t *cells, uint64_t n_cells, const double *umi_counts, double *out);

/* ---------------------------------------------------------------- kNN on the scores (next step of the pipeline)
 * knn / find_nn (scan-rs/src/nn.rs:38-83): for every query row (n_queries x dim, row-major) the indices of its k nearest points
 * (n_points x dim) in Euclidean distance, nearest first, into out[n_queries x k]; rows with fewer than k candidates are padded
 * with 0xFFFFFFFF (the reference pads with T::max_value(), nn.rs:67).  include_self = 0 skips the point with index
 * self_offset + row (knn() queries the tree with its own points; self_offset serves cell-sharded callers).  Among exactly
 * equidistant points the lower index comes first. */
SB_API int sb_knn(sb_ctx *ctx, const double *points, uint64_t n_points, uint32_t dim, const double *queries, uint64_t n_queries,
           uint32_t k, int include_self, uint64_t self_offset, uint32_t *out);

/* ---------------------------------------------------------------- measurement
 * Device-side timing on the context's own stream (CUDA events) and per-kernel accounting;
 * bench.py builds its roofline object from these.  Not part of the reference interface. */
typedef struct {
    double spmm_t_ms;     /* K7: out[cells x w]  = A^T . Y  (cell-major gather) */
    double spmm_n_ms;     /* K8: out[genes x w]  = A . X    (gene-major panel gather) */
    double moments_ms;    /* K5 */
    double reduce_ms;     /* integer reductions K1/K2 */
    double dense_ms;      /* QR / Gram / eigh / small GEMMs */
    double comm_ms;       /* NCCL collectives */
    double spmm_t_bytes;  /* algorithmic bytes summed over launches (SURVEY 8d formula) */
    double spmm_n_bytes;
    double spmm_t_flops;
    double spmm_n_flops;
    uint64_t spmm_t_launches;
    uint64_t spmm_n_launches;
    uint64_t kernel_launches; /* every kernel this library launched (own + library calls) */
    uint64_t own_kernel_launches;
    double upload_ms;     /* sb_upload: host -> device copies */
    double build_ms;      /* sb_upload: device-side layout build (sorts, dense panel) */
    double output_ms;     /* device -> host copies of U, sigma, V */
} sb_profile;
SB_API int sb_profile_enable(sb_ctx *ctx, int on); /* on: record a CUDA-event pair around every phase */
SB_API int sb_profile_reset(sb_ctx *ctx);
SB_API int sb_profile_get(sb_ctx *ctx, sb_profile *out);
SB_API int sb_timer_begin(sb_ctx *ctx);            /* records an event on the context stream */
SB_API int sb_timer_end(sb_ctx *ctx, float *ms);   /* records, synchronises, returns elapsed ms */
/* write >= 2x L2 worth of bytes so the next timed step starts cold */
SB_API int sb_flush_l2(sb_ctx *ctx);

/* ---------------------------------------------------------------- synthetic workload
 * Test/bench utility, not part of the reference interface: negative-binomial count matrix
 * from a counter-based hash keyed by (seed, gene, global cell), identical bit for bit to
 * the CPU generator in synth/ (see synth_nb.h).  pf[n_clusters x m] are per-cluster gene
 * abundances, depth[n_local] per-cell depths, cluster[n_local] cluster ids. */
SB_API int sb_synth_generate(sb_ctx *ctx, uint32_t m, uint64_t n_local, uint64_t cell_offset, uint64_t seed,
                      uint32_t n_clusters, const double *pf, const double *depth, const uint8_t *cluster,
                      uint32_t r_dispersion, sb_mat **out);

#ifdef __cplusplus
}
#endif
#endif /* SCANB200_H */
