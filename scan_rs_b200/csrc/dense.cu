// dense.cu -- the small dense steps around the sparse products: thin QR of tall blocks
// (bk_svd.rs:94,98,123,127 -> LAPACK dgeqrf/dorgqr in the reference), the Gram matrix of the
// projected block, its symmetric eigendecomposition (standing in for dgesdd on the wide
// matrix, bk_svd.rs:105,134) and tall x small products.  All tall blocks are row-major
// [rows x w] with an even leading dimension; cuSOLVER/cuBLAS see them as column-major
// [w x rows].
#include "common.cuh"
#include "eig_host.h"

// out (col-major rows x w, ld = rows)  <-  in (row-major rows x w, ld = ldi)
__global__ void k_rm_to_cm(const double *__restrict__ in, u64 rows, u32 w, u32 ldi, double *__restrict__ out) {
    __shared__ double tile[32][33];
    u64 r0 = (u64)blockIdx.x * 32;
    u32 c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        u64 r = r0 + i;
        u32 c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < w) ? in[r * ldi + c] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        u32 c = c0 + i;
        u64 r = r0 + threadIdx.x;
        if (r < rows && c < w) out[(u64)c * rows + r] = tile[threadIdx.x][i];
    }
}

// out (row-major rows x w, ld = ldo)  <-  in (col-major rows x w, ld = rows); pad columns zeroed
__global__ void k_cm_to_rm(const double *__restrict__ in, u64 rows, u32 w, double *__restrict__ out, u32 ldo) {
    __shared__ double tile[32][33];
    u64 r0 = (u64)blockIdx.x * 32;
    u32 c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        u32 c = c0 + i;
        u64 r = r0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < w) ? in[(u64)c * rows + r] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        u64 r = r0 + i;
        u32 c = c0 + threadIdx.x;
        if (r < rows && c < ldo) out[r * ldo + c] = c < w ? tile[threadIdx.x][i] : 0.0;
    }
}

// copies the upper triangle of the leading w x w block of a column-major (ld = rows) factor into R (col-major w x w)
__global__ void k_extract_r(const double *__restrict__ cm, u64 rows, u32 w, double *__restrict__ R) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w * w) {
        u32 c = i / w, r = i % w;
        R[i] = (r <= c && r < rows) ? cm[(u64)c * rows + r] : 0.0;
    }
}

// thin QR: A (row-major rows x w, ld) is replaced by Q (rows x min(rows, w)); returns the new width.
// R_out (device, col-major w x w) optionally receives the triangular factor (needs rows >= w).
int qr_tall(sb_ctx *ctx, double *A, u64 rows, u32 w, u32 ld, u32 *w_out, double *R_out) {
    TraceScope trc(ctx, "dense: qr_tall");
    ProfScope ps(ctx, PH_DENSE);
    u32 kq = (u32)std::min<u64>(rows, w);
    *w_out = kq;
    if (rows == 0 || w == 0) return SB_OK;
    if (rows > 0x7FFFFFFFull) return sb_fail(SB_ERR_UNSUPPORTED, "qr_tall: more than 2^31 rows");
    DevBuf<double> cm, tau, work;
    DevBuf<int> info;
    SB_TRY(cm.alloc(rows * (size_t)w));
    SB_TRY(tau.alloc(kq));
    SB_TRY(info.alloc(1));
    dim3 blk(32, 8), grd(cdiv(rows, 32), cdiv(w, 32));
    k_rm_to_cm<<<grd, blk, 0, ctx->stream>>>(A, rows, w, ld, cm.p);
    count_launch(ctx);
    int lwork1 = 0, lwork2 = 0;
    SB_CUSOLVER(cusolverDnDgeqrf_bufferSize(ctx->cusolver, (int)rows, (int)w, cm.p, (int)rows, &lwork1));
    SB_CUSOLVER(cusolverDnDorgqr_bufferSize(ctx->cusolver, (int)rows, (int)kq, (int)kq, cm.p, (int)rows, tau.p, &lwork2));
    int lwork = std::max(lwork1, lwork2);
    SB_TRY(work.alloc(lwork));
    SB_CUSOLVER(cusolverDnDgeqrf(ctx->cusolver, (int)rows, (int)w, cm.p, (int)rows, tau.p, work.p, lwork, info.p));
    if (R_out && rows >= w) {
        k_extract_r<<<cdiv((u64)w * w, 256), 256, 0, ctx->stream>>>(cm.p, rows, w, R_out);
        count_launch(ctx);
    }
    SB_CUSOLVER(cusolverDnDorgqr(ctx->cusolver, (int)rows, (int)kq, (int)kq, cm.p, (int)rows, tau.p, work.p, lwork, info.p));
    count_launch(ctx, false); count_launch(ctx, false);
    int h_info = 0;
    SB_CUDA(cudaMemcpyAsync(&h_info, info.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    dim3 grd2(cdiv(rows, 32), cdiv(std::max(ld, w), 32));
    k_cm_to_rm<<<grd2, blk, 0, ctx->stream>>>(cm.p, rows, kq, A, ld);
    count_launch(ctx);
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_info != 0) return sb_fail(SB_ERR_LINALG, "QR failed: info = %d", h_info);
    return SB_OK;
}

// G (col-major w x w) = A^T A for A row-major rows x w (ld); all-reduced over ranks when `reduce`
int gram(sb_ctx *ctx, const double *A, u64 rows, u32 w, u32 ld, double *G, bool reduce) {
    TraceScope trc(ctx, "dense: gram");
    ProfScope ps(ctx, PH_DENSE);
    if (rows > 0x7FFFFFFFull) return sb_fail(SB_ERR_UNSUPPORTED, "gram: more than 2^31 rows");
    const double one = 1.0, zero = 0.0;
    if (rows == 0) {
        SB_CUDA(cudaMemsetAsync(G, 0, (size_t)w * w * sizeof(double), ctx->stream));
    } else {
        // A_c = A^T is (w x rows) column-major with ld; G = A_c . A_c^T
        SB_CUBLAS(cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_T, (int)w, (int)w, (int)rows, &one, A, (int)ld, A, (int)ld, &zero, G, (int)w));
        count_launch(ctx, false);
    }
    if (reduce) SB_TRY(comm_allreduce_f64(ctx, G, (size_t)w * w));
    return SB_OK;
}

// symmetric eigendecomposition of G (col-major w x w, overwritten by eigenvectors); evals ascending.  The solver's status
// goes to info_dev (device) and is read by the caller with its outputs: no host synchronisation here.
int eigh(sb_ctx *ctx, double *G, u32 w, double *evals_dev, int *info_dev) {
    TraceScope trc(ctx, "dense: eigh");
    ProfScope ps(ctx, PH_DENSE);
    if (w <= 192 && ctx->eig_jacobi) {
        // small Gram matrices (w = 100 at k = 10): cuSOLVER's Jacobi solver -- a handful of launches instead of the ~230 tiny kernels
        // of syevd's tridiagonalisation + divide and conquer (1.8 ms at w = 100, the largest fixed cost of a sharded step)
        syevjInfo_t par = nullptr;
        SB_CUSOLVER(cusolverDnCreateSyevjInfo(&par));
        cusolverDnXsyevjSetTolerance(par, 1.0e-15);
        cusolverDnXsyevjSetMaxSweeps(par, 100);
        cusolverDnXsyevjSetSortEig(par, 1);
        int lw = 0;
        cusolverStatus_t st = cusolverDnDsyevj_bufferSize(ctx->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)w, G, (int)w, evals_dev, &lw, par);
        DevBuf<double> wk;
        int rc = SB_OK;
        if (st == CUSOLVER_STATUS_SUCCESS) rc = wk.alloc(lw);
        if (st == CUSOLVER_STATUS_SUCCESS && rc == SB_OK)
            st = cusolverDnDsyevj(ctx->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)w, G, (int)w, evals_dev, wk.p, lw, info_dev, par);
        cusolverDnDestroySyevjInfo(par);
        SB_TRY(rc);
        if (st != CUSOLVER_STATUS_SUCCESS) return sb_fail(SB_ERR_LINALG, "cuSOLVER syevj error %d", (int)st);
        count_launch(ctx, false);
        return SB_OK;
    }
    int lwork = 0;
    SB_CUSOLVER(cusolverDnDsyevd_bufferSize(ctx->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)w, G, (int)w, evals_dev, &lwork));
    DevBuf<double> work;
    SB_TRY(work.alloc(lwork));
    SB_CUSOLVER(cusolverDnDsyevd(ctx->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)w, G, (int)w, evals_dev, work.p, lwork, info_dev));
    count_launch(ctx, false);
    return SB_OK;
}

// The k largest eigenpairs of G, in the layout eigh leaves them in (eigenvalues ascending in evals_dev[w-k .. w), their vectors in
// the last k columns of G; the rest of both arrays is unspecified).  Small orders go to the host (eig_host.h: tridiagonalisation,
// QL for the eigenvalues, inverse iteration for the k vectors, a residual / orthonormality check on the original matrix): 0.75 ms
// at w = 100, k = 10 against 2.7 ms of syevd, at the price of one host round trip of 80 KB.  Anything the host path declines
// (non-finite input, a failed check) and every larger order runs through eigh().
int eigh_topk(sb_ctx *ctx, double *G, u32 w, u32 k, double *evals_dev, int *info_dev) {
    if (ctx->eig_host && w <= 128 && k >= 1 && k <= w) {
        TraceScope trc(ctx, "dense: eigh (host, top k)");
        ProfScope ps(ctx, PH_DENSE);
        const size_t need = ((size_t)w * w + (size_t)w * k + k) * sizeof(double);
        if (ctx->eig_pinned_bytes < need) {
            if (ctx->eig_pinned) cudaFreeHost(ctx->eig_pinned);
            ctx->eig_pinned = nullptr;
            ctx->eig_pinned_bytes = 0;
            SB_CUDA(cudaMallocHost(&ctx->eig_pinned, need));
            ctx->eig_pinned_bytes = need;
        }
        double *hG = static_cast<double *>(ctx->eig_pinned), *hvec = hG + (size_t)w * w, *hlam = hvec + (size_t)w * k;
        SB_CUDA(cudaMemcpyAsync(hG, G, (size_t)w * w * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (eig_host::topk(hG, (int)w, (int)k, hlam, hvec)) {
            SB_CUDA(cudaMemcpyAsync(G + (size_t)(w - k) * w, hvec, (size_t)w * k * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            SB_CUDA(cudaMemcpyAsync(evals_dev + (w - k), hlam, (size_t)k * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            SB_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), ctx->stream));
            return SB_OK;
        }
        ctx->eig_host_declined++;
    }
    return eigh(ctx, G, w, evals_dev, info_dev);
}

// Out (row-major rows x k, ldo) = A (row-major rows x w, lda) . S (col-major w x k, lds)
int gemm_tall_small(sb_ctx *ctx, const double *A, u64 rows, u32 w, u32 lda, const double *S, u32 k, u32 lds, double *Out, u32 ldo) {
    TraceScope trc(ctx, "dense: gemm_tall_small");
    ProfScope ps(ctx, PH_DENSE);
    if (rows == 0 || k == 0) return SB_OK;
    if (rows > 0x7FFFFFFFull) return sb_fail(SB_ERR_UNSUPPORTED, "gemm: more than 2^31 rows");
    const double one = 1.0, zero = 0.0;
    SB_CUBLAS(cublasDgemm(ctx->cublas, CUBLAS_OP_T, CUBLAS_OP_N, (int)k, (int)rows, (int)w, &one, S, (int)lds, A, (int)lda, &zero, Out, (int)ldo));
    count_launch(ctx, false);
    return SB_OK;
}

// A (row-major rows x w, ld) <- A . R^-1 for upper-triangular R (col-major w x w): the column-major
// view A_c = A^T (w x rows) gets R^-T applied from the left.
int trsm_right_upper(sb_ctx *ctx, double *A, u64 rows, u32 w, u32 ld, const double *R) {
    TraceScope trc(ctx, "dense: trsm");
    ProfScope ps(ctx, PH_DENSE);
    if (rows == 0 || w == 0) return SB_OK;
    if (rows > 0x7FFFFFFFull) return sb_fail(SB_ERR_UNSUPPORTED, "trsm: more than 2^31 rows");
    const double one = 1.0;
    SB_CUBLAS(cublasDtrsm(ctx->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, (int)w, (int)rows, &one, R, (int)w,
                          A, (int)ld));
    count_launch(ctx, false);
    return SB_OK;
}
