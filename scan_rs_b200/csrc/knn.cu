// knn.cu -- k nearest neighbours on the PCA scores (scan-rs/src/nn.rs:38-83: ball-tree kNN with rayon queries).
// The scores are a small dense block (n x k_pcs) that is already on the GPU side of the pipeline, so the B200 form is a tiled
// brute force: a CTA owns 128 queries (one per thread, coordinates in registers / local memory), streams the points through
// shared memory in tiles (every lane of a warp reads the same point: broadcast loads), and every thread keeps its k best
// in a sorted list.  The distance is evaluated exactly as the reference's `Pt::distance` does (nn.rs:12-20): sum over the
// coordinates of (p1 - p0)^2 in index order, no FMA contraction, then sqrt -- so the neighbour order agrees with the reference
// wherever distances differ; among exactly equidistant points the lower index comes first (the reference's order there is
// the ball tree's traversal order, nn.rs:198-211).
#include "common.cuh"

#define KNN_THREADS 128
#define KNN_TILE 256
#define KNN_MAX_DIM 128
#define KNN_MAX_K 128

template <int DREG>
__global__ void __launch_bounds__(KNN_THREADS) k_knn(const double *__restrict__ P, u64 np, u32 dim, const double *__restrict__ Q, u64 nq, u32 k, int include_self,
                                                     u64 self_offset, u32 *__restrict__ out, double *__restrict__ scratch_d) {
    extern __shared__ double tile[];  // KNN_TILE x dim
    const u64 qi = (u64)blockIdx.x * KNN_THREADS + threadIdx.x;
    const bool live = qi < nq;
    double q[DREG];
#pragma unroll
    for (int j = 0; j < DREG; j++) q[j] = (live && j < (int)dim) ? Q[qi * dim + j] : 0.0;
    // this thread's sorted list lives in global scratch (k doubles) + the output row (k indices): both private to the thread
    double *bd = scratch_d + (live ? qi : 0) * k;
    u32 *bi = out + (live ? qi : 0) * k;
    if (live)
        for (u32 j = 0; j < k; j++) {
            bd[j] = INFINITY;
            bi[j] = 0xFFFFFFFFu;
        }
    double worst = INFINITY;
    const u64 self = self_offset + qi;
    for (u64 p0 = 0; p0 < np; p0 += KNN_TILE) {
        const u32 cnt = (u32)min((u64)KNN_TILE, np - p0);
        __syncthreads();
        for (u32 i = threadIdx.x; i < cnt * dim; i += KNN_THREADS) tile[i] = P[p0 * dim + i];
        __syncthreads();
        if (!live) continue;
        for (u32 t = 0; t < cnt; t++) {
            const double *pt = tile + (size_t)t * dim;
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < DREG; j++)
                if (j < (int)dim) {
                    const double df = __dsub_rn(pt[j], q[j]);
                    s = __dadd_rn(s, __dmul_rn(df, df));
                }
            for (u32 j = DREG; j < dim; j++) {  // coordinates beyond the register block
                const double df = __dsub_rn(pt[j], Q[qi * dim + j]);
                s = __dadd_rn(s, __dmul_rn(df, df));
            }
            const double d = sqrt(s);
            if (d < worst && (include_self || p0 + t != self)) {
                u32 pos = k - 1;  // insertion into the sorted list (strict <: an equidistant later point stays behind)
                while (pos > 0 && bd[pos - 1] > d) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    pos--;
                }
                bd[pos] = d;
                bi[pos] = (u32)(p0 + t);
                worst = bd[k - 1];
            }
        }
    }
}

extern "C" int sb_knn(sb_ctx *ctx, const double *points, uint64_t n_points, uint32_t dim, const double *queries, uint64_t n_queries, uint32_t k,
                      int include_self, uint64_t self_offset, uint32_t *out) {
    if (!ctx || (n_queries && k && !out)) return sb_fail(SB_ERR_INVALID_ARG, "sb_knn: NULL argument");
    if ((n_points && !points) || (n_queries && !queries)) return sb_fail(SB_ERR_INVALID_ARG, "sb_knn: NULL argument");
    if (dim == 0 || dim > KNN_MAX_DIM) return sb_fail(SB_ERR_UNSUPPORTED, "sb_knn: dimension must be 1..%d", KNN_MAX_DIM);
    if (k > KNN_MAX_K) return sb_fail(SB_ERR_UNSUPPORTED, "sb_knn: k must be <= %d", KNN_MAX_K);
    if (n_points > 0xFFFFFFFEull) return sb_fail(SB_ERR_UNSUPPORTED, "sb_knn: more than 2^32 - 2 points");
    if (n_queries == 0 || k == 0) return SB_OK;
    SB_ENTER(ctx);
    DevBuf<double> dP, dQ, dD;
    DevBuf<u32> dO;
    SB_TRY(dP.alloc(n_points * dim));
    SB_TRY(dQ.alloc(n_queries * dim));
    SB_TRY(dD.alloc(n_queries * k));
    SB_TRY(dO.alloc(n_queries * k));
    if (n_points) SB_CUDA(cudaMemcpyAsync(dP.p, points, n_points * dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(cudaMemcpyAsync(dQ.p, queries, n_queries * dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const unsigned blocks = cdiv(n_queries, KNN_THREADS);
    const size_t smem = (size_t)KNN_TILE * dim * sizeof(double);
    if (dim <= 16) {
        SB_CUDA(cudaFuncSetAttribute(k_knn<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_knn<16><<<blocks, KNN_THREADS, smem, ctx->stream>>>(dP.p, n_points, dim, dQ.p, n_queries, k, include_self, self_offset, dO.p, dD.p);
    } else {
        SB_CUDA(cudaFuncSetAttribute(k_knn<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_knn<32><<<blocks, KNN_THREADS, smem, ctx->stream>>>(dP.p, n_points, dim, dQ.p, n_queries, k, include_self, self_offset, dO.p, dD.p);
    }
    count_launch(ctx);
    SB_CUDA(cudaGetLastError());
    SB_CUDA(cudaMemcpyAsync(out, dO.p, n_queries * k * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
