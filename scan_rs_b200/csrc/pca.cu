// pca.cu -- drivers of the truncated SVD: block Krylov (scan-rs/src/dim_red/bk_svd.rs:57-146)
// and randomized (scan-rs/src/dim_red/rand_svd.rs:54-129), both branches (m >= n and n > m),
// on the device-resident normalized matrix.  The sparse products run in spmm.cu; QR / Gram /
// eigh in dense.cu.  With a communicator the cells are sharded over ranks: A^T.Y stays
// shard-local, A.X and the Gram matrix are all-reduced, the gene-side QR is replicated.
#include <cmath>

#include "common.cuh"

int spmm_t(sb_nmat *a, const double *Y, u32 ldy, u32 w, double *out, u32 ldo, double *uy_scratch);
int spmm_n(sb_nmat *a, const double *X, u32 ldx, u32 w, double *P, u32 ldp);
int qr_tall(sb_ctx *ctx, double *A, u64 rows, u32 w, u32 ld, u32 *w_out, double *R_out = nullptr);
int trsm_right_upper(sb_ctx *ctx, double *A, u64 rows, u32 w, u32 ld, const double *R);
int gram(sb_ctx *ctx, const double *A, u64 rows, u32 w, u32 ld, double *G, bool reduce);
int eigh_topk(sb_ctx *ctx, double *G, u32 w, u32 k, double *evals_dev, int *info_dev);
int gemm_tall_small(sb_ctx *ctx, const double *A, u64 rows, u32 w, u32 lda, const double *S, u32 k, u32 lds, double *Out, u32 ldo);
// dense_own.cu: the repo's own tall-skinny kernels (FP64 mma.sync) and the CholeskyQR3 factorisation
int syrk_tall(sb_ctx *ctx, const double *A, u64 rows, u32 w, u32 ld, double *G);
int gemm_tall(sb_ctx *ctx, const double *A, u64 rows, u32 w, u32 lda, const double *S, u32 k, u32 lds, double *Out, u32 ldo, bool s_upper = false);
int qr_chol(sb_ctx *ctx, double *A, double *tmp, u64 rows, u32 w, u32 ld, double *Rinv_out, int *flag, u64 rows_global = 0);
int topk_select(sb_ctx *ctx, const double *W, const double *ev, u32 wq, u32 k, double *Wsel, double *Wsc, double *S);
int comm_allreduce_f64(sb_ctx *ctx, double *buf, size_t count);

// ---------------------------------------------------------------- start block
// rand 0.10.1 SmallRng on 64-bit targets = Xoshiro256++, seeded from a u64 through SplitMix64;
// Uniform::<f64>::new(-1.0, 1.0): value1_2 = f64::from_bits((u64 >> 12) | 0x3FF0..), then
// (value1_2 - 1.0) * 2.0 + (-1.0).  Known answers are in tests/test_host_abi.py.  (The crate
// sources are not available offline, so bit-equality with a Rust build is unpinned)
struct Xoshiro256pp {
    u64 s[4];
    explicit Xoshiro256pp(u64 seed) {
        u64 z = seed;
        for (int i = 0; i < 4; i++) {
            z += 0x9E3779B97F4A7C15ULL;
            u64 x = z;
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
            x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
            s[i] = x ^ (x >> 31);
        }
    }
    static inline u64 rotl(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
    u64 next() {
        u64 r = rotl(s[0] + s[3], 23) + s[0];
        u64 t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
};

// the start block is a pure function of (seed, rows, cols): keep the last one (BkSvd always asks for seed 0)
static const double *omega_cached(sb_ctx *ctx, u64 seed, u64 rows, u64 cols) {
    if (ctx->omega_seed != seed || ctx->omega_rows != rows || ctx->omega_cols != cols || ctx->omega_cache.size() != rows * cols) {
        ctx->omega_cache.resize(rows * cols);
        sb_omega(seed, rows, cols, ctx->omega_cache.data());
        ctx->omega_seed = seed;
        ctx->omega_rows = rows;
        ctx->omega_cols = cols;
    }
    return ctx->omega_cache.data();
}

extern "C" int sb_omega(uint64_t seed, uint64_t rows, uint64_t cols, double *out) {
    if (!out && rows * cols) return sb_fail(SB_ERR_INVALID_ARG, "sb_omega: out is NULL");
    Xoshiro256pp rng(seed);
    for (u64 i = 0; i < rows * cols; i++) {
        u64 bits = (rng.next() >> 12) | 0x3FF0000000000000ULL;
        double v12;
        memcpy(&v12, &bits, 8);
        out[i] = (v12 - 1.0) * 2.0 + (-1.0);
    }
    return SB_OK;
}

// ---------------------------------------------------------------- helpers
static inline u32 even_up(u32 w) { return (w + 1) & ~1u; }

struct Tall {  // row-major rows x w block, even ld, zero padded
    DevBuf<double> buf;
    u64 rows = 0;
    u32 w = 0, ld = 0;
    int init(sb_ctx *ctx, u64 r, u32 width, u64 extra_rows = 0) {
        rows = r;
        w = width;
        ld = even_up(width);
        SB_TRY(buf.alloc((rows + extra_rows) * (size_t)ld));
        SB_CUDA(cudaMemsetAsync(buf.p, 0, std::max<size_t>(1, (rows + extra_rows) * (size_t)ld) * sizeof(double), ctx->stream));
        return SB_OK;
    }
};

// copies src (rows x w, ld lds) into columns [c0, c0+w) of dst (ld ldd)
static int copy_block(sb_ctx *ctx, double *dst, u32 ldd, u32 c0, const double *src, u32 lds, u64 rows, u32 w) {
    if (rows == 0 || w == 0) return SB_OK;
    SB_CUDA(cudaMemcpy2DAsync(dst + c0, ldd * sizeof(double), src, lds * sizeof(double), w * sizeof(double), rows, cudaMemcpyDeviceToDevice,
                              ctx->stream));
    return SB_OK;
}

static int progress(sb_ctx *ctx, sb_progress_cb cb, void *user, double frac) {
    // Without a callback nothing can cancel: no collective, no host synchronisation (every rank must pass a callback or none).
    if (!cb) return SB_OK;
    int cancel = cb(frac, user) != 0;
    if (ctx->nranks > 1) SB_TRY(comm_allreduce_max_i32(ctx, &cancel));
    if (cancel) return sb_fail(SB_ERR_CANCELLED, "cancelled at progress %.3f", frac);
    return SB_OK;
}

// From the projected tall block Tt (rows_t x wq, the cell side when n > m, the gene side when
// m >= n) and the orthonormal basis Q (rows_q x wq): Gram -> eigh -> top-k triplets.
//   left  (rows_t x k) = Tt . W_k . diag(1/sigma)      right (rows_q x k) = Q . W_k
// `reduce_gram` all-reduces the Gram matrix (cell-sharded Tt).  Nothing returns to the host here: S_dev (k values) and the
// eigensolver's status (info_dev) are read by the caller together with the outputs.
static int finish_svd(sb_ctx *ctx, const Tall &Tt, const Tall &Q, u32 wq, u32 k, bool reduce_gram, double *S_dev, int *info_dev, Tall &from_t, Tall &from_q) {
    DevBuf<double> G, ev, Wk;
    SB_TRY(G.alloc((size_t)wq * wq));
    SB_TRY(ev.alloc(wq));
    SB_TRY(Wk.alloc((size_t)wq * k * 2));
    if (ctx->own_dense) {
        ProfScope ps(ctx, PH_DENSE);
        SB_TRY(syrk_tall(ctx, Tt.buf.p, Tt.rows, wq, Tt.ld, G.p));
        if (reduce_gram) SB_TRY(comm_allreduce_f64(ctx, G.p, (size_t)wq * wq));
    } else {
        SB_TRY(gram(ctx, Tt.buf.p, Tt.rows, wq, Tt.ld, G.p, reduce_gram));
    }
    // the w x w Gram matrix (w = 100 at k = 10): only its k largest eigenpairs are used below.  Orders up to 128 are solved on the
    // host (dense.cu: eigh_topk, eig_host.h), larger ones by cuSOLVER's syevd.  (A one-CTA parallel Jacobi solver written for this
    // step was measured at 8 ms against syevd's 2.7: three barriers and 30,000 element updates per tournament round, ~800 rounds.)
    SB_TRY(eigh_topk(ctx, G.p, wq, k, ev.p, info_dev));
    double *dWsel = Wk.p, *dWsc = Wk.p + (size_t)wq * k;
    SB_TRY(topk_select(ctx, G.p, ev.p, wq, k, dWsel, dWsc, S_dev));
    SB_TRY(from_t.init(ctx, Tt.rows, k));
    SB_TRY(from_q.init(ctx, Q.rows, k));
    if (ctx->own_dense) {
        ProfScope ps(ctx, PH_DENSE);
        SB_TRY(gemm_tall(ctx, Tt.buf.p, Tt.rows, wq, Tt.ld, dWsc, k, wq, from_t.buf.p, from_t.ld));
        SB_TRY(gemm_tall(ctx, Q.buf.p, Q.rows, wq, Q.ld, dWsel, k, wq, from_q.buf.p, from_q.ld));
    } else {
        SB_TRY(gemm_tall_small(ctx, Tt.buf.p, Tt.rows, wq, Tt.ld, dWsc, k, wq, from_t.buf.p, from_t.ld));
        SB_TRY(gemm_tall_small(ctx, Q.buf.p, Q.rows, wq, Q.ld, dWsel, k, wq, from_q.buf.p, from_q.ld));
    }
    return SB_OK;
}

static int download_tall(sb_ctx *ctx, const Tall &t, double *host) {
    if (t.rows == 0 || t.w == 0) return SB_OK;
    ProfScope ps(ctx, PH_OUTPUT);
    if (t.ld == t.w)  // contiguous block: one linear copy (the pitched copy of a 1.3M x 10 block runs at under half the PCIe rate)
        SB_CUDA(cudaMemcpyAsync(host, t.buf.p, t.rows * (size_t)t.w * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    else
        SB_CUDA(cudaMemcpy2DAsync(host, t.w * sizeof(double), t.buf.p, t.ld * sizeof(double), t.w * sizeof(double), t.rows, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

// uploads a row-major host block (rows x w) into a Tall; `transpose_in`: host is (w x rows)
static int upload_tall(sb_ctx *ctx, Tall &t, const double *host, bool transpose_in) {
    if (t.rows == 0 || t.w == 0) return SB_OK;
    if (!transpose_in) {
        SB_CUDA(cudaMemcpy2DAsync(t.buf.p, t.ld * sizeof(double), host, t.w * sizeof(double), t.w * sizeof(double), t.rows, cudaMemcpyHostToDevice,
                                  ctx->stream));
        SB_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        std::vector<double> tmp(t.rows * (size_t)t.ld, 0.0);
        for (u32 j = 0; j < t.w; j++)
            for (u64 r = 0; r < t.rows; r++) tmp[r * t.ld + j] = host[(size_t)j * t.rows + r];
        // same stream as the allocation and the zero-fill: a default-stream copy would race with them
        SB_CUDA(cudaMemcpyAsync(t.buf.p, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        SB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return SB_OK;
}

// The projection T = Q^T A (bk_svd.rs:102,131) costs a sparse pass of width b*n_iter.  With K = Q R,
// Q^T A = R^-T (K^T A), and the blocks of K^T A are the products the Krylov loop forms anyway (plus one
// more width-b pass for the last block), so the wide pass becomes a product with R^-1 on the tall block.
// R is ill-conditioned by construction (Krylov blocks are nearly dependent) and the rounding error already
// in K^T A is amplified by cond(R) in the directions R squeezes.  Guard ("Projection"): a 2-norm
// condition estimate of the triangular factor (power + inverse iteration on the host, O(w^2) each):
//   cond <= SB_COND_TRUST   : the identity is used as is (measured margin >= 30x on the sigma / angle bars at 1e9)
//   cond <= SB_COND_VERIFY  : used, then verified a posteriori: || A^T u_i - sigma_i v_i || <= SB_RESID_TOL sigma_1 on
//                             a few of the k triplets (one narrow sparse pass); on failure the direct wide pass reruns
//   otherwise               : the direct wide pass
#define SB_COND_TRUST 1.0e9
#define SB_COND_VERIFY 1.0e12
#define SB_RESID_TOL 1.0e-8

// 2-norm condition estimate of an upper-triangular matrix (column-major w x w, host)
static double cond_upper(const std::vector<double> &M, u32 w) {
    if (w == 0) return 1.0;
    for (u32 i = 0; i < w; i++) {
        const double d = M[(size_t)i * w + i];
        if (!(d == d) || d == 0.0 || std::fabs(d) > 1e300) return INFINITY;
    }
    std::vector<double> x(w), y(w), z(w);
    auto normalize = [&](std::vector<double> &v) {
        double s = 0.0;
        for (double t : v) s += t * t;
        s = std::sqrt(s);
        if (s > 0.0 && s == s && s < 1e300)
            for (double &t : v) t /= s;
        return s;
    };
    // sigma_max: power iteration on M^T M
    for (u32 i = 0; i < w; i++) x[i] = 1.0 + 0.37 * std::sin(1.7 * (double)i);
    normalize(x);
    double smax2 = 0.0;
    for (int it = 0; it < 12; it++) {
        for (u32 i = 0; i < w; i++) {  // y = M x
            double t = 0.0;
            for (u32 j = i; j < w; j++) t += M[(size_t)j * w + i] * x[j];
            y[i] = t;
        }
        for (u32 j = 0; j < w; j++) {  // z = M^T y
            double t = 0.0;
            for (u32 i = 0; i <= j; i++) t += M[(size_t)j * w + i] * y[i];
            z[j] = t;
        }
        smax2 = normalize(z);
        x = z;
    }
    // sigma_min: inverse iteration, M^T y = x (forward), M z = y (backward)
    for (u32 i = 0; i < w; i++) x[i] = 1.0 + 0.41 * std::cos(2.3 * (double)i);
    normalize(x);
    double inv_smin2 = 0.0;
    for (int it = 0; it < 12; it++) {
        for (u32 j = 0; j < w; j++) {
            double t = x[j];
            for (u32 i = 0; i < j; i++) t -= M[(size_t)j * w + i] * y[i];
            y[j] = t / M[(size_t)j * w + j];
        }
        for (u32 ii = w; ii-- > 0;) {
            double t = y[ii];
            for (u32 j = ii + 1; j < w; j++) t -= M[(size_t)j * w + ii] * z[j];
            z[ii] = t / M[(size_t)ii * w + ii];
        }
        inv_smin2 = normalize(z);
        if (!(inv_smin2 == inv_smin2) || inv_smin2 > 1e300) return INFINITY;
        x = z;
    }
    const double c = std::sqrt(smax2 * inv_smin2);
    return c == c ? c : INFINITY;
}

// 0: direct pass, 1: identity trusted, 2: identity + a-posteriori verification
static int projection_mode(sb_ctx *ctx, const double *tri_dev, u32 w, double *cond_out) {
    std::vector<double> h((size_t)w * w);
    if (cudaMemcpyAsync(h.data(), tri_dev, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        *cond_out = INFINITY;
        return 0;
    }
    const double c = cond_upper(h, w);  // cond(R) = cond(R^-1): either factor serves
    *cond_out = c;
    ctx->last_cond_r = c;
    if (TraceScope::on()) fprintf(stderr, "[scanb200] projection: cond(R) ~ %.3e\n", c);
    if (ctx->verify_projection && c <= SB_COND_VERIFY) return 2;
    if (c <= SB_COND_TRUST) return 1;
    if (c <= SB_COND_VERIFY) return 2;
    return 0;
}

// sum over cells of (T2[c, j] - sigma_sel[j] * V[c, sel[j]])^2 for up to 4 probe columns
__global__ void k_probe_resid(const double *__restrict__ T2, u32 ld2, const double *__restrict__ V, u32 ldv, u64 n, const double *__restrict__ S,
                              int4 sel, u32 nsel, double *__restrict__ out) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int se[4] = {sel.x, sel.y, sel.z, sel.w};
    for (u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (u64)gridDim.x * blockDim.x)
        for (u32 j = 0; j < nsel; j++) {
            const double d = T2[c * (size_t)ld2 + j] - S[se[j]] * V[c * (size_t)ldv + se[j]];
            acc[j] = fma(d, d, acc[j]);
        }
    for (u32 j = 0; j < nsel; j++) {
        double v = acc[j];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(out + j, v);
    }
}

// copies columns sel[] of src (rows x *, ld lds) into the leading columns of dst (ld ldd)
__global__ void k_pick_cols(const double *__restrict__ src, u32 lds, u64 rows, int4 sel, u32 nsel, double *__restrict__ dst, u32 ldd) {
    const int se[4] = {sel.x, sel.y, sel.z, sel.w};
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < rows * nsel; i += (u64)gridDim.x * blockDim.x) {
        const u64 r = i / nsel;
        const u32 j = (u32)(i - r * nsel);
        dst[r * ldd + j] = src[r * (size_t)lds + se[j]];
    }
}

static int check_shape(const sb_nmat *a, u32 k) {
    const sb_mat *mt = a->mat;
    if (mt->m < 2 || mt->n_global < 2) return sb_fail(SB_ERR_INVALID_SHAPE, "The input matrix must be at least 2x2.");
    if ((u64)k > std::min<u64>(mt->m, mt->n_global)) return sb_fail(SB_ERR_INVALID_K, "invalid k");
    if (k == 0) return sb_fail(SB_ERR_INVALID_K, "invalid k");
    return SB_OK;
}

// ---------------------------------------------------------------- block Krylov SVD
// Thin QR of a tall block in place.  own: shifted CholeskyQR3 on the repo's kernels (dense_own.cu), no host synchronisation,
// Rinv_out (optional) = R^-1; otherwise LAPACK-style Householder through cuSOLVER (dense.cu), R_out (optional) = R.
// rows_global != 0: A is a cell-side block row-sharded over the ranks; only the CholeskyQR path can factor it (one all-reduce of
// the Gram matrix per round), so `own` must hold.
static int qr_block(sb_ctx *ctx, bool own, Tall &A, Tall &tmp, u32 *wq, double *tri_out, int *flag_dev, u64 rows_global = 0) {
    if (own) {
        *wq = A.w;
        return qr_chol(ctx, A.buf.p, tmp.buf.p, A.rows, A.w, A.ld, tri_out, flag_dev, rows_global);
    }
    if (rows_global) return sb_fail(SB_ERR_UNSUPPORTED, "the Householder QR of a cell-side block is single-rank (sharded runs need own_dense = 1 and a block with 4 w <= n)");
    return qr_tall(ctx, A.buf.p, A.rows, A.w, A.ld, wq, tri_out);
}

// T <- T . R^-1 (own: product with the explicit inverse on the tensor path; else cuBLAS trsm with R)
static int apply_rinv(sb_ctx *ctx, bool own, Tall &T, Tall &tmp, const double *tri) {
    if (!own) return trsm_right_upper(ctx, T.buf.p, T.rows, T.w, T.ld, tri);
    ProfScope ps(ctx, PH_DENSE);
    if (T.w <= 128) return gemm_tall(ctx, T.buf.p, T.rows, T.w, T.ld, tri, T.w, T.w, T.buf.p, T.ld, true);  // one column block per CTA: in place is safe
    SB_TRY(tmp.init(ctx, T.rows, T.w));
    SB_TRY(gemm_tall(ctx, T.buf.p, T.rows, T.w, T.ld, tri, T.w, T.w, tmp.buf.p, tmp.ld, true));
    T.buf.swap(tmp.buf);
    tmp.buf.release();
    return SB_OK;
}

struct PcaStatus {  // read back once, with the outputs
    int chol_flag;   // CholeskyQR breakdown (rank-deficient block)
    int eig_info;    // syevd status
    int pad[2];
};

static int bksvd_impl(sb_nmat *a, uint32_t k, uint32_t b, uint32_t n_iter, uint64_t seed, const double *omega, sb_progress_cb cb, void *user,
                      double *U, double *S, double *V, bool own_qr, bool force_direct, bool *retry_householder, bool *retry_direct) {
    sb_mat *mt = a->mat;
    sb_ctx *ctx = mt->ctx;
    const u32 m = mt->m;
    const u64 n = mt->n, ng = mt->n_global;
    DevBuf<double> uy, S_dev, resid;
    DevBuf<int> status;
    SB_TRY(uy.alloc(even_up(b * n_iter) + 2));
    SB_TRY(S_dev.alloc(k));
    SB_TRY(resid.alloc(4));
    SB_TRY(status.alloc(4));
    SB_CUDA(cudaMemsetAsync(status.p, 0, 4 * sizeof(int), ctx->stream));
    int *chol_flag = status.p, *eig_info = status.p + 1;
    const u32 bq = b * n_iter;
    PcaStatus hst = {0, 0, {0, 0}};
    double cond = 0.0;

    if ((u64)m >= ng) {
        // ---- m >= n (bk_svd.rs:89-115): block on the cell side.  Sharded: B and K are row-sharded like the cells, their QRs are
        // distributed CholeskyQR (Gram matrices all-reduced), A.B is all-reduced by spmm_n, T = A.Q and its SVD are replicated.
        const u64 sh = ctx->nranks > 1 ? ng : 0;  // rows_global of the cell-side blocks
        Tall B, Kc, W, WK, tmpB, tmpK;
        const bool fast = !ctx->direct_projection && !force_direct && (b % 2 == 0) && (u64)bq <= ng;
        const bool own_b = own_qr && ng >= 4ull * b, own_k = own_qr && ng >= 4ull * bq;
        SB_TRY(B.init(ctx, n, b));
        if (!omega) omega = omega_cached(ctx, seed, ng, b) + (size_t)mt->cell_offset * b;  // :90 (n x b, row-major fill): this rank's rows
        SB_TRY(upload_tall(ctx, B, omega, false));
        SB_TRY(Kc.init(ctx, n, bq));
        SB_TRY(W.init(ctx, m, b, 1));
        if (own_b) SB_TRY(tmpB.init(ctx, n, b));
        if (fast) SB_TRY(WK.init(ctx, m, bq, 1));  // column block i-1 keeps A . B_i (B_i = block i of K)
        // sharded: spmm_n all-reduces its output in place and needs a contiguous block, so the product lands in W and the column
        // block of WK is a copy of it
        const bool via_w = sh != 0;
        for (u32 i = 0; i < n_iter; i++) {
            double *Wp = (fast && i > 0 && !via_w) ? WK.buf.p + (size_t)(i - 1) * b : W.buf.p;
            u32 Wld = (fast && i > 0 && !via_w) ? WK.ld : W.ld;
            SB_TRY(spmm_n(a, B.buf.p, B.ld, b, Wp, Wld));                   // A.dot(&B)
            if (fast && i > 0 && via_w) SB_TRY(copy_block(ctx, WK.buf.p, WK.ld, (i - 1) * b, W.buf.p, W.ld, m, b));
            SB_TRY(spmm_t(a, Wp, Wld, b, B.buf.p, B.ld, uy.p));              // (.)^T.dot(A) ^T
            u32 wq = 0;
            SB_TRY(qr_block(ctx, own_b, B, tmpB, &wq, nullptr, chol_flag, sh));  // .qr()?.0   :94
            SB_TRY(copy_block(ctx, Kc.buf.p, Kc.ld, i * b, B.buf.p, B.ld, n, b));  // :95
            SB_TRY(progress(ctx, cb, user, (double)i / (double)n_iter * 0.8));     // :96
        }
        u32 wq = 0;
        DevBuf<double> tri;
        SB_TRY(tri.alloc((size_t)bq * bq));
        if (fast && !via_w) SB_TRY(spmm_n(a, B.buf.p, B.ld, b, WK.buf.p + (size_t)(n_iter - 1) * b, WK.ld));  // A . B_q
        if (fast && via_w) {
            SB_TRY(spmm_n(a, B.buf.p, B.ld, b, W.buf.p, W.ld));
            SB_TRY(copy_block(ctx, WK.buf.p, WK.ld, (n_iter - 1) * b, W.buf.p, W.ld, m, b));
        }
        if (own_k) SB_TRY(tmpK.init(ctx, n, bq));
        SB_TRY(qr_block(ctx, own_k, Kc, tmpK, &wq, tri.p, chol_flag, sh));   // :98
        tmpK.buf.release();
        SB_TRY(progress(ctx, cb, user, 0.82));
        int mode = 0;
        if (fast && wq == bq) mode = projection_mode(ctx, tri.p, bq, &cond);
        if (mode == 2) mode = 0;  // this branch has no a-posteriori check: past the trusted range the direct pass runs
        Tall Tdirect;
        Tall *Tp = &WK;
        if (mode == 1) {
            Tall tmpT;
            WK.rows = m;
            SB_TRY(apply_rinv(ctx, own_k, WK, tmpT, tri.p));                 // T = (A K) R^-1 = A Q
        } else {
            SB_TRY(Tdirect.init(ctx, m, wq, 1));
            SB_TRY(spmm_n(a, Kc.buf.p, Kc.ld, wq, Tdirect.buf.p, Tdirect.ld));  // T = A.dot(&Q)  :102
            Tp = &Tdirect;
        }
        Tall &T = *Tp;
        SB_TRY(progress(ctx, cb, user, 0.93));
        if (k > wq) return sb_fail(SB_ERR_INVALID_K, "invalid k");
        Tall Uo, Vo;
        Kc.w = wq;
        SB_TRY(finish_svd(ctx, T, Kc, wq, k, false, S_dev.p, eig_info, Uo, Vo));  // :105-113
        SB_CUDA(cudaMemcpyAsync(S, S_dev.p, k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(cudaMemcpyAsync(&hst, status.p, sizeof(hst), cudaMemcpyDeviceToHost, ctx->stream));
        SB_TRY(download_tall(ctx, Uo, U));
        SB_TRY(download_tall(ctx, Vo, V));
        if (hst.chol_flag && sh) return sb_fail(SB_ERR_LINALG, "sb_bksvd: CholeskyQR broke down on a sharded cell-side block (rank-deficient Krylov block)");
        if (hst.chol_flag) { *retry_householder = true; return SB_OK; }
        if (hst.eig_info != 0) return sb_fail(SB_ERR_LINALG, "eigendecomposition failed: info = %d", hst.eig_info);
        SB_TRY(progress(ctx, cb, user, 1.0));
        return SB_OK;
    }

    // ---- n > m (bk_svd.rs:116-145): block on the gene side, replicated over ranks
    TraceScope tr_all(ctx, "bksvd: total (n > m)");
    Tall Y, Kt, T, P, TK, tmpP, tmpK;
    const bool fast = !ctx->direct_projection && !force_direct && (b % 2 == 0) && bq <= m;
    const bool own_b = own_qr && m >= 4u * b, own_k = own_qr && m >= 4u * bq;
    {
        TraceScope t0(ctx, "bksvd: omega + buffers");
        SB_TRY(Y.init(ctx, m, b));
        if (!omega) {
            // :118 (b x m, row-major fill).  BkSvd::run_pca always asks for seed 0: the block is generated, transposed and uploaded
            // once per (seed, shape) and then copied on the device (the host transpose + pageable copy were ~2 ms of every call)
            if (ctx->omega_dev_seed != seed || ctx->omega_dev_rows != b || ctx->omega_dev_cols != m || !ctx->omega_dev.p) {
                SB_TRY(upload_tall(ctx, Y, omega_cached(ctx, seed, b, m), true));  // Y[g, j] = B[j, g]
                SB_TRY(ctx->omega_dev.alloc((size_t)m * Y.ld));
                SB_CUDA(cudaMemcpyAsync(ctx->omega_dev.p, Y.buf.p, (size_t)m * Y.ld * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
                ctx->omega_dev_seed = seed;
                ctx->omega_dev_rows = b;
                ctx->omega_dev_cols = m;
            } else {
                SB_CUDA(cudaMemcpyAsync(Y.buf.p, ctx->omega_dev.p, (size_t)m * Y.ld * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            }
        } else {
            SB_TRY(upload_tall(ctx, Y, omega, true));  // Y[g, j] = B[j, g]
        }
        SB_TRY(Kt.init(ctx, m, bq));
        SB_TRY(T.init(ctx, n, b));
        SB_TRY(P.init(ctx, m, b, 1));
        if (own_b) SB_TRY(tmpP.init(ctx, m, b));
        if (fast) SB_TRY(TK.init(ctx, n, bq));  // column block i-1 keeps A^T . Y_i (Y_i = block i of K^T)
    }
    for (u32 i = 0; i < n_iter; i++) {
        double *Tp = (fast && i > 0) ? TK.buf.p + (size_t)(i - 1) * b : T.buf.p;
        u32 Tld = (fast && i > 0) ? TK.ld : T.ld;
        {
            TraceScope t1(ctx, "bksvd: spmm_t");
            SB_TRY(spmm_t(a, Y.buf.p, Y.ld, b, Tp, Tld, uy.p));              // T = B.dot(A)^T        :122
        }
        {
            TraceScope t1(ctx, "bksvd: spmm_n");
            SB_TRY(spmm_n(a, Tp, Tld, b, P.buf.p, P.ld));                    // A.dot(&T)             :123
        }
        TraceScope t2(ctx, "bksvd: qr + copies + progress");
        u32 wq = 0;
        SB_TRY(qr_block(ctx, own_b, P, tmpP, &wq, nullptr, chol_flag));      // .qr()?.0
        SB_TRY(copy_block(ctx, Y.buf.p, Y.ld, 0, P.buf.p, P.ld, m, b));
        SB_TRY(copy_block(ctx, Kt.buf.p, Kt.ld, i * b, P.buf.p, P.ld, m, b));  // K rows i*b..   :124
        SB_TRY(progress(ctx, cb, user, (double)i / (double)n_iter * 0.8));   // :125
    }
    u32 wq = 0;
    DevBuf<double> tri;
    SB_TRY(tri.alloc((size_t)bq * bq));
    if (fast) SB_TRY(spmm_t(a, Y.buf.p, Y.ld, b, TK.buf.p + (size_t)(n_iter - 1) * b, TK.ld, uy.p));  // A^T . Y_q
    if (own_k) SB_TRY(tmpK.init(ctx, m, bq));
    SB_TRY(qr_block(ctx, own_k, Kt, tmpK, &wq, tri.p, chol_flag));           // Q = K.t().qr()?.0     :127
    tmpK.buf.release();
    SB_TRY(progress(ctx, cb, user, 0.82));
    int mode = 0;
    if (fast && wq == bq) mode = projection_mode(ctx, tri.p, bq, &cond);
    Tall Tdirect;
    Tall *Ttp = &TK;
    if (mode >= 1) {
        Tall tmpT;
        SB_TRY(apply_rinv(ctx, own_k, TK, tmpT, tri.p));                     // T^T = (A^T K^T) R^-1 = A^T Q
    } else {
        SB_TRY(Tdirect.init(ctx, n, wq));
        SB_TRY(spmm_t(a, Kt.buf.p, Kt.ld, wq, Tdirect.buf.p, Tdirect.ld, uy.p));  // T = Q.t().dot(A)      :131
        Ttp = &Tdirect;
    }
    Tall &Tt = *Ttp;
    SB_TRY(progress(ctx, cb, user, 0.93));
    if (k > wq) return sb_fail(SB_ERR_INVALID_K, "invalid k");
    Tall Vo, Uo;
    Kt.w = wq;
    TraceScope t9(ctx, "bksvd: gram + eigh + outputs");
    SB_TRY(finish_svd(ctx, Tt, Kt, wq, k, true, S_dev.p, eig_info, Vo, Uo)); // :134-142
    u32 nsel = 0;
    if (mode == 2) {
        // a-posteriori check of the identity on a few triplets: A^T u_i against sigma_i v_i (shard-local rows, summed over ranks)
        nsel = std::min<u32>(k, 4);
        int4 sel = make_int4(0, (int)(k - 1), (int)(k / 2), (int)(k / 4));
        if (nsel < 4) sel = make_int4(0, nsel > 1 ? 1 : 0, nsel > 2 ? 2 : 0, 0);
        Tall Us, Ts;
        SB_TRY(Us.init(ctx, m, nsel));
        SB_TRY(Ts.init(ctx, n, nsel));
        SB_CUDA(cudaMemsetAsync(resid.p, 0, 4 * sizeof(double), ctx->stream));
        k_pick_cols<<<std::max(1u, std::min<u32>(cdiv((u64)m * nsel, 256), 1024u)), 256, 0, ctx->stream>>>(Uo.buf.p, Uo.ld, m, sel, nsel, Us.buf.p, Us.ld);
        count_launch(ctx);
        SB_TRY(spmm_t(a, Us.buf.p, Us.ld, nsel, Ts.buf.p, Ts.ld, uy.p));
        if (n) {
            k_probe_resid<<<(unsigned)std::max<u64>(1, std::min<u64>((n + 255) / 256, (u64)ctx->sm_count * 8)), 256, 0, ctx->stream>>>(
                Ts.buf.p, Ts.ld, Vo.buf.p, Vo.ld, n, S_dev.p, sel, nsel, resid.p);
            count_launch(ctx);
        }
        SB_TRY(comm_allreduce_f64(ctx, resid.p, 4));
    }
    double h_res[4] = {0, 0, 0, 0};
    SB_CUDA(cudaMemcpyAsync(S, S_dev.p, k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(cudaMemcpyAsync(&hst, status.p, sizeof(hst), cudaMemcpyDeviceToHost, ctx->stream));
    if (mode == 2) SB_CUDA(cudaMemcpyAsync(h_res, resid.p, sizeof(h_res), cudaMemcpyDeviceToHost, ctx->stream));
    SB_TRY(download_tall(ctx, Uo, U));
    SB_TRY(download_tall(ctx, Vo, V));  // synchronises
    if (hst.chol_flag) { *retry_householder = true; return SB_OK; }
    if (hst.eig_info != 0) return sb_fail(SB_ERR_LINALG, "eigendecomposition failed: info = %d", hst.eig_info);
    if (mode == 2) {
        double worst = 0.0;
        for (u32 j = 0; j < nsel; j++) worst = std::max(worst, std::sqrt(std::max(h_res[j], 0.0)));
        const double rel = S[0] > 0.0 ? worst / S[0] : (worst > 0.0 ? INFINITY : 0.0);
        ctx->last_probe_resid = rel;
        if (TraceScope::on()) fprintf(stderr, "[scanb200] projection: probe residual %.3e sigma_1 (cond %.3e)\n", rel, cond);
        if (!(rel <= SB_RESID_TOL)) { *retry_direct = true; return SB_OK; }
    }
    SB_TRY(progress(ctx, cb, user, 1.0));
    return SB_OK;
}

extern "C" int sb_bksvd(sb_nmat *a, uint32_t k, uint32_t b, uint32_t n_iter, uint64_t seed, const double *omega, sb_progress_cb cb, void *user,
                        double *U, double *S, double *V) {
    if (!a || !U || !S || !V) return sb_fail(SB_ERR_INVALID_ARG, "sb_bksvd: NULL argument");
    sb_mat *mt = a->mat;
    sb_ctx *ctx = mt->ctx;
    SB_ENTER(ctx);
    SB_TRY(check_shape(a, k));
    if (n_iter == 0) return sb_fail(SB_ERR_INVALID_ARG, "sb_bksvd: n_iter must be >= 1");
    b = (u32)std::min<u64>(std::min<u64>(mt->m, mt->n_global), b);  // bk_svd.rs:81
    if (b < k) return sb_fail(SB_ERR_INVALID_K, "invalid k");
    bool own_qr = ctx->own_dense, force_direct = false;
    ctx->last_cond_r = 0.0;
    ctx->last_probe_resid = 0.0;
    ctx->last_fallbacks = 0;
    for (int attempt = 0; attempt < 3; attempt++) {
        bool retry_h = false, retry_d = false;
        SB_TRY(bksvd_impl(a, k, b, n_iter, seed, omega, cb, user, U, S, V, own_qr, force_direct, &retry_h, &retry_d));
        if (!retry_h && !retry_d) return SB_OK;
        // Both decisions derive from all-reduced (Gram, residual) or replicated (gene-side blocks) data: every rank takes the same turn.
        if (retry_h) own_qr = false;       // CholeskyQR broke down (rank-deficient block): Householder handles it
        if (retry_d) force_direct = true;  // the projection identity failed its check: the reference's wide pass
        ctx->last_fallbacks += 1;
    }
    return sb_fail(SB_ERR_LINALG, "sb_bksvd: no stable path");
}

extern "C" int sb_pca_diagnostics(sb_ctx *ctx, double *cond_r, double *probe_resid, int *fallbacks) {
    if (!ctx) return sb_fail(SB_ERR_INVALID_ARG, "sb_pca_diagnostics: ctx is NULL");
    if (cond_r) *cond_r = ctx->last_cond_r;
    if (probe_resid) *probe_resid = ctx->last_probe_resid;
    if (fallbacks) *fallbacks = ctx->last_fallbacks;
    return SB_OK;
}

extern "C" int sb_bksvd_run_pca(sb_nmat *a, uint32_t k, double k_multiplier, uint32_t n_iter, sb_progress_cb cb, void *user, double *U, double *S,
                                double *V) {
    u32 bsize = (u32)std::ceil((double)k * k_multiplier);  // bk_svd.rs:49
    return sb_bksvd(a, k, bsize, n_iter, 0, nullptr, cb, user, U, S, V);
}

// ---------------------------------------------------------------- randomized SVD
extern "C" int sb_randsvd(sb_nmat *a, uint32_t k, uint32_t l, uint32_t n_iter, uint64_t seed, const double *omega, double *U, double *S, double *V) {
    if (!a || !U || !S || !V) return sb_fail(SB_ERR_INVALID_ARG, "sb_randsvd: NULL argument");
    sb_mat *mt = a->mat;
    sb_ctx *ctx = mt->ctx;
    SB_ENTER(ctx);
    SB_TRY(check_shape(a, k));
    const u32 m = mt->m;
    const u64 n = mt->n, ng = mt->n_global;
    // Sharded: the gene-side blocks are replicated (any QR), the cell-side blocks are row-sharded (distributed CholeskyQR).
    const u64 sh = ctx->nranks > 1 ? ng : 0;
    if ((u64)l > std::min<u64>(m, ng)) return sb_fail(SB_ERR_UNSUPPORTED, "sb_randsvd: l = %u exceeds min(m, n)", l);
    if (sh && !(ctx->own_dense && ng >= 4ull * l)) return sb_fail(SB_ERR_UNSUPPORTED, "sb_randsvd: sharded runs need own_dense = 1 and 4 l <= n");
    if (l < k) return sb_fail(SB_ERR_INVALID_K, "invalid k");
    DevBuf<double> uy;
    SB_TRY(uy.alloc(even_up(l) + 2));
    std::vector<double> h_om;
    DevBuf<double> S_dev;
    DevBuf<int> einfo;
    int h_info = 0;
    SB_TRY(S_dev.alloc(k));
    SB_TRY(einfo.alloc(1));
    SB_CUDA(cudaMemsetAsync(einfo.p, 0, sizeof(int), ctx->stream));
    Tall Qm, Qn, tmpN;  // gene-side (m x l, +1 row for spmm_n) and cell-side (n x l) blocks
    SB_TRY(Qm.init(ctx, m, l, 1));
    SB_TRY(Qn.init(ctx, n, l));
    DevBuf<int> cflag;
    SB_TRY(cflag.alloc(1));
    SB_CUDA(cudaMemsetAsync(cflag.p, 0, sizeof(int), ctx->stream));
    if (sh) SB_TRY(tmpN.init(ctx, n, l));
    u32 wq = 0;
    // the cell-side QR: Householder on one rank (as the reference), distributed CholeskyQR3 on a sharded context
    auto qr_cells = [&]() -> int {
        if (!sh) return qr_tall(ctx, Qn.buf.p, n, l, Qn.ld, &wq);
        wq = l;
        return qr_chol(ctx, Qn.buf.p, tmpN.buf.p, n, l, Qn.ld, nullptr, cflag.p, sh);
    };
    if ((u64)m >= ng) {  // rand_svd.rs:85-105
        if (!omega) {
            h_om.resize((size_t)ng * l);
            SB_TRY(sb_omega(seed, ng, l, h_om.data()));
            omega = h_om.data() + (size_t)mt->cell_offset * l;  // this rank's rows of the n x l start block
        }
        SB_TRY(upload_tall(ctx, Qn, omega, false));
        SB_TRY(spmm_n(a, Qn.buf.p, Qn.ld, l, Qm.buf.p, Qm.ld));              // Q = A.dot(&omega).qr()  :87
        SB_TRY(qr_tall(ctx, Qm.buf.p, m, l, Qm.ld, &wq));
        for (u32 i = 0; i < n_iter; i++) {
            SB_TRY(spmm_t(a, Qm.buf.p, Qm.ld, l, Qn.buf.p, Qn.ld, uy.p));     // Q.t().dot(A)^T .qr()    :90
            SB_TRY(qr_cells());
            SB_TRY(spmm_n(a, Qn.buf.p, Qn.ld, l, Qm.buf.p, Qm.ld));          // A.dot(&Q).qr()          :91
            SB_TRY(qr_tall(ctx, Qm.buf.p, m, l, Qm.ld, &wq));
        }
        SB_TRY(spmm_t(a, Qm.buf.p, Qm.ld, l, Qn.buf.p, Qn.ld, uy.p));         // B = Q.t().dot(A)        :96  (stored as B^T)
        Tall Vo, Uo;
        SB_TRY(finish_svd(ctx, Qn, Qm, l, k, sh != 0, S_dev.p, einfo.p, Vo, Uo)); // U = Q.U_B             :104 (Gram of the sharded B^T: all-reduced)
        SB_CUDA(cudaMemcpyAsync(S, S_dev.p, k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(cudaMemcpyAsync(&h_info, einfo.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SB_TRY(download_tall(ctx, Uo, U));
        SB_TRY(download_tall(ctx, Vo, V));
    } else {  // rand_svd.rs:106-128
        if (!omega) {
            h_om.resize((size_t)l * m);
            SB_TRY(sb_omega(seed, l, m, h_om.data()));
            omega = h_om.data();
        }
        SB_TRY(upload_tall(ctx, Qm, omega, true));
        SB_TRY(spmm_t(a, Qm.buf.p, Qm.ld, l, Qn.buf.p, Qn.ld, uy.p));         // omega.dot(A)^T .qr()    :109
        SB_TRY(qr_cells());
        for (u32 i = 0; i < n_iter; i++) {
            SB_TRY(spmm_n(a, Qn.buf.p, Qn.ld, l, Qm.buf.p, Qm.ld));          // A.dot(&Q).qr()          :112
            SB_TRY(qr_tall(ctx, Qm.buf.p, m, l, Qm.ld, &wq));
            SB_TRY(spmm_t(a, Qm.buf.p, Qm.ld, l, Qn.buf.p, Qn.ld, uy.p));     // Q.t().dot(A)^T .qr()    :113
            SB_TRY(qr_cells());
        }
        SB_TRY(spmm_n(a, Qn.buf.p, Qn.ld, l, Qm.buf.p, Qm.ld));              // B = A.dot(&Q)           :118
        Tall Uo, Vo;
        SB_TRY(finish_svd(ctx, Qm, Qn, l, k, false, S_dev.p, einfo.p, Uo, Vo)); // Va = Vt[:k].Q^T         :126
        SB_CUDA(cudaMemcpyAsync(S, S_dev.p, k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(cudaMemcpyAsync(&h_info, einfo.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SB_TRY(download_tall(ctx, Uo, U));
        SB_TRY(download_tall(ctx, Vo, V));
    }
    if (h_info != 0) return sb_fail(SB_ERR_LINALG, "eigendecomposition failed: info = %d", h_info);
    if (sh) {
        int h_flag = 0;
        SB_CUDA(cudaMemcpyAsync(&h_flag, cflag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (h_flag) return sb_fail(SB_ERR_LINALG, "sb_randsvd: CholeskyQR broke down on a sharded cell-side block");
    }
    return SB_OK;
}

extern "C" int sb_randsvd_run_pca(sb_nmat *a, uint32_t k, double l_multiplier, uint32_t n_iter, double *U, double *S, double *V) {
    u32 l = std::max<u32>(k + 4, (u32)((double)k * l_multiplier));  // rand_svd.rs:46
    return sb_randsvd(a, k, l, n_iter, 0, nullptr, U, S, V);
}
