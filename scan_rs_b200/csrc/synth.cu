// synth.cu -- device-side synthetic count matrix (test / bench utility, see synth_nb.h).
// Compiled with -fmad=false so that the sampler's f64 arithmetic matches the CPU generator bit
// for bit.  One warp per cell, lanes over genes, two passes (count, then fill in gene order).
#include "common.cuh"
#include "synth_nb.h"

int mat_from_device_cm(sb_ctx *ctx, u32 m, u64 n, DevBuf<u64> &cm_ptr, DevBuf<uint2> &cm, u64 nnz, sb_mat **out);

__global__ void k_synth(u32 m, u64 n, u64 cell_offset, u64 seed, const double *__restrict__ pf, const double *__restrict__ depth,
                        const unsigned char *__restrict__ cluster, u32 r, const u64 *__restrict__ ptr, u32 *__restrict__ counts,
                        uint2 *__restrict__ out) {
    u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    int lane = threadIdx.x & 31;
    for (u64 c = warp; c < n; c += nwarps) {
        const double d = depth[c];
        const double *p = pf + (size_t)cluster[c] * m;
        u64 base = ptr ? ptr[c] : 0;
        u32 total = 0;
        for (u32 g0 = 0; g0 < m; g0 += 32) {
            u32 g = g0 + lane;
            u32 cnt = 0;
            if (g < m) cnt = synth_nb_count(seed, g, cell_offset + c, d * p[g], r);
            unsigned mask = __ballot_sync(0xffffffffu, cnt != 0);
            if (out && cnt) out[base + total + __popc(mask & ((1u << lane) - 1u))] = make_uint2(g, cnt);
            total += __popc(mask);
        }
        if (counts && lane == 0) counts[c] = total;
    }
}

__global__ void k_widen(const u32 *__restrict__ in, u64 *__restrict__ out, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) out[i] = in[i];
}

// serial-ish exclusive scan for the per-cell counts (n <= a few million): block-wise two level
__global__ void k_scan_blocks(const u32 *__restrict__ counts, u64 n, u64 *__restrict__ block_sums) {
    // each block sums 4096 entries
    __shared__ unsigned long long sh[32];
    u64 b0 = (u64)blockIdx.x * 4096;
    unsigned long long acc = 0;
    for (u64 i = b0 + threadIdx.x; i < min(n, b0 + 4096); i += blockDim.x) acc += counts[i];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += sh[i];
        block_sums[blockIdx.x] = s;
    }
}

__global__ void k_scan_finish(const u32 *__restrict__ counts, u64 n, const u64 *__restrict__ block_offsets, u64 *__restrict__ ptr) {
    // one thread per block of 4096: sequential inside (cheap: n/4096 threads)
    u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 b0 = b * 4096;
    if (b0 >= n) return;
    u64 run = block_offsets[b];
    for (u64 i = b0; i < min(n, b0 + 4096); i++) {
        ptr[i] = run;
        run += counts[i];
    }
    if (min(n, b0 + 4096) == n) ptr[n] = run;
}

extern "C" int sb_synth_generate(sb_ctx *ctx, uint32_t m, uint64_t n_local, uint64_t cell_offset, uint64_t seed, uint32_t n_clusters,
                                 const double *pf, const double *depth, const uint8_t *cluster, uint32_t r_dispersion, sb_mat **out) {
    if (!ctx || !out || !pf || (n_local && (!depth || !cluster)) || r_dispersion == 0 || n_clusters == 0)
        return sb_fail(SB_ERR_INVALID_ARG, "sb_synth_generate: bad argument");
    if (m > SB_GENE_MASK) return sb_fail(SB_ERR_UNSUPPORTED, "too many genes");
    SB_ENTER(ctx);
    *out = nullptr;
    const u64 n = n_local;
    DevBuf<double> d_pf, d_depth;
    DevBuf<unsigned char> d_cl;
    DevBuf<u32> counts;
    DevBuf<u64> ptr, bsum;
    SB_TRY(d_pf.alloc((size_t)n_clusters * m));
    SB_TRY(d_depth.alloc(n));
    SB_TRY(d_cl.alloc(n));
    SB_TRY(counts.alloc(n));
    SB_TRY(ptr.alloc(n + 1));
    SB_CUDA(cudaMemcpyAsync(d_pf.p, pf, (size_t)n_clusters * m * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (n) {
        SB_CUDA(cudaMemcpyAsync(d_depth.p, depth, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        SB_CUDA(cudaMemcpyAsync(d_cl.p, cluster, n, cudaMemcpyHostToDevice, ctx->stream));
    }
    int blocks = ctx->sm_count * 8;
    if (n) {
        k_synth<<<blocks, 256, 0, ctx->stream>>>(m, n, cell_offset, seed, d_pf.p, d_depth.p, d_cl.p, r_dispersion, nullptr, counts.p, nullptr);
        count_launch(ctx);
    }
    // exclusive scan of counts -> ptr
    u64 nb = (n + 4095) / 4096;
    SB_TRY(bsum.alloc(nb + 1));
    std::vector<u64> h_bsum(nb + 1, 0);
    if (n) {
        k_scan_blocks<<<(unsigned)nb, 256, 0, ctx->stream>>>(counts.p, n, bsum.p);
        count_launch(ctx);
        SB_CUDA(cudaMemcpyAsync(h_bsum.data(), bsum.p, nb * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(cudaStreamSynchronize(ctx->stream));
        u64 run = 0;
        for (u64 b = 0; b < nb; b++) {
            u64 t = h_bsum[b];
            h_bsum[b] = run;
            run += t;
        }
        h_bsum[nb] = run;
        SB_CUDA(cudaMemcpyAsync(bsum.p, h_bsum.data(), (nb + 1) * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
        k_scan_finish<<<cdiv(nb, 128), 128, 0, ctx->stream>>>(counts.p, n, bsum.p, ptr.p);
        count_launch(ctx);
    } else {
        SB_CUDA(cudaMemsetAsync(ptr.p, 0, sizeof(u64), ctx->stream));
    }
    u64 nnz = h_bsum[nb];
    DevBuf<uint2> cm;
    SB_TRY(cm.alloc(nnz));
    if (n && nnz) {
        k_synth<<<blocks, 256, 0, ctx->stream>>>(m, n, cell_offset, seed, d_pf.p, d_depth.p, d_cl.p, r_dispersion, ptr.p, nullptr, cm.p);
        count_launch(ctx);
    }
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    SB_CUDA(cudaGetLastError());
    return mat_from_device_cm(ctx, m, n, ptr, cm, nnz, out);
}
