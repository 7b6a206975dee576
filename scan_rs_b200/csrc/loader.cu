// loader.cu -- the device half of the reference's matrix loaders: what happens between "the arrays of the
// file are in memory" and "the filtered count matrix exists".
//   hdf5-io/src/matrix.rs:56-89    read_csc_matrix: CSC (cell-major) arrays of the `matrix` group; Cell Ranger 3 files may carry
//                                  UNSORTED gene indices inside a cell -> `new_from_unsorted_csc` sorts every cell's entries
//   hdf5-io/src/matrix.rs:93-117   compute_genes_filter: drop the features whose type does not match `retain_feature_like` and
//                                  those whose total count (u64) is below `shrink_row`
//   hdf5-io/src/matrix.rs:119-192  read_adaptive_csr_matrix: the surviving features, in file order, become the matrix rows
// The file parsing itself (HDF5 container, gz MatrixMarket text) is host work: scan_rs_b200/h5.py and mtx.py.
#include <cub/cub.cuh>

#include "common.cuh"

int mat_from_device_cm(sb_ctx *ctx, u32 m, u64 n, DevBuf<u64> &cm_ptr, DevBuf<uint2> &cm, u64 nnz, sb_mat **out);
int mat_gene_sums_dev(sb_mat *mat, int mode, const unsigned char *excl_cells, const unsigned char *excl_genes, u64 *d_out, bool allreduce);

__global__ void k_ld_zip(const u32 *__restrict__ idx, const u32 *__restrict__ cnt, u64 nnz, uint2 *__restrict__ out) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (u64)gridDim.x * blockDim.x) out[i] = make_uint2(idx[i], cnt[i]);
}

// flags: [0] pointer not monotone, [1] index out of range, [2] duplicate index inside a cell (after the sort)
__global__ void k_ld_check(const u64 *__restrict__ ptr, const u32 *__restrict__ idx, u64 n, u32 m, int *__restrict__ flags) {
    u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (u64 c = warp; c < n; c += nwarps) {
        const u64 s = ptr[c], e = ptr[c + 1];
        if (e < s) {
            if (lane == 0) flags[0] = 1;
            continue;
        }
        for (u64 k = s + lane; k < e; k += 32) {
            const u32 g = idx[k];
            if (g >= m) flags[1] = 1;
            if (k > s && idx[k - 1] == g) flags[2] = 1;
        }
    }
}

// Cell-major arrays with the gene indices of a cell in ANY order (the Cell Ranger 3 defect): every cell's entries are sorted by
// gene on the device (one segmented radix sort), duplicates inside a cell are an error (sprs rejects them too), zeros dropped
// like every other constructor.
extern "C" int sb_upload_unsorted(sb_ctx *ctx, uint32_t m, uint64_t n_local, const uint64_t *indptr, const uint32_t *idx, const uint32_t *cnt,
                                  sb_mat **out) {
    if (!ctx || !out || !indptr) return sb_fail(SB_ERR_INVALID_ARG, "sb_upload_unsorted: NULL argument");
    *out = nullptr;
    SB_ENTER(ctx);
    if (m > SB_GENE_MASK) return sb_fail(SB_ERR_UNSUPPORTED, "sb_upload_unsorted: more than %u genes", SB_GENE_MASK);
    if (indptr[0] != 0) return sb_fail(SB_ERR_INVALID_ARG, "sb_upload_unsorted: indptr is not monotone from 0");
    for (u64 i = 0; i < n_local; i++)
        if (indptr[i + 1] < indptr[i]) return sb_fail(SB_ERR_INVALID_ARG, "sb_upload_unsorted: indptr is not monotone from 0 (at %llu)", (unsigned long long)i);
    const u64 nnz = indptr[n_local];
    if (nnz && (!idx || !cnt)) return sb_fail(SB_ERR_INVALID_ARG, "sb_upload_unsorted: NULL idx/cnt");
    if (nnz > 0x7FFFFFFFull) return sb_fail(SB_ERR_UNSUPPORTED, "sb_upload_unsorted: more than 2^31 entries per call (shard the cells)");
    DevBuf<u64> d_ptr;
    DevBuf<u32> k_in, v_in, k_out, v_out;
    DevBuf<uint2> cm;
    DevBuf<int> flags;
    SB_TRY(d_ptr.alloc(n_local + 1));
    SB_TRY(k_in.alloc(nnz));
    SB_TRY(v_in.alloc(nnz));
    SB_TRY(k_out.alloc(nnz));
    SB_TRY(v_out.alloc(nnz));
    SB_TRY(cm.alloc(nnz));
    SB_TRY(flags.alloc(4));
    SB_CUDA(cudaMemsetAsync(flags.p, 0, 4 * sizeof(int), ctx->stream));
    {
        ProfScope ps(ctx, PH_UPLOAD);
        SB_CUDA(cudaMemcpyAsync(d_ptr.p, indptr, (n_local + 1) * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
        if (nnz) {
            SB_CUDA(cudaMemcpyAsync(k_in.p, idx, nnz * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
            SB_CUDA(cudaMemcpyAsync(v_in.p, cnt, nnz * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    ProfScope pb(ctx, PH_BUILD);
    if (nnz) {
        int end_bit = 1;
        while (end_bit < 32 && ((m ? m - 1 : 0) >> end_bit)) end_bit++;
        size_t tmp_bytes = 0;
        SB_CUDA(cub::DeviceSegmentedRadixSort::SortPairs(nullptr, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, (int)nnz, (int)n_local, d_ptr.p, d_ptr.p + 1, 0,
                                                         end_bit, ctx->stream));
        DevBuf<char> tmp;
        SB_TRY(tmp.alloc(tmp_bytes));
        SB_CUDA(cub::DeviceSegmentedRadixSort::SortPairs(tmp.p, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, (int)nnz, (int)n_local, d_ptr.p, d_ptr.p + 1, 0,
                                                         end_bit, ctx->stream));
        count_launch(ctx, false);
        const int grid = (int)std::max<u64>(1, std::min<u64>((n_local * 32 + 255) / 256, (u64)ctx->sm_count * 16));
        k_ld_check<<<grid, 256, 0, ctx->stream>>>(d_ptr.p, k_out.p, n_local, m, flags.p);
        k_ld_zip<<<(int)std::max<u64>(1, std::min<u64>((nnz + 255) / 256, (u64)ctx->sm_count * 16)), 256, 0, ctx->stream>>>(k_out.p, v_out.p, nnz, cm.p);
        count_launch(ctx); count_launch(ctx);
    }
    int h[4] = {0, 0, 0, 0};
    SB_CUDA(cudaMemcpyAsync(h, flags.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h[1]) return sb_fail(SB_ERR_INVALID_ARG, "sb_upload_unsorted: gene index out of range %u", m);
    if (h[2]) return sb_fail(SB_ERR_INVALID_ARG, "sb_upload_unsorted: duplicate gene index inside a cell");
    return mat_from_device_cm(ctx, m, n_local, d_ptr, cm, nnz, out);
}

// compute_genes_filter + the row selection of read_adaptive_csr_matrix (hdf5-io/src/matrix.rs:93-192).  type_keep[m] (may be
// NULL = keep all): 1 for the features whose type matches `retain_feature_like` (a host string test); min_total: shrink_row
// (0 = None).  A feature survives if its type is kept AND its total count (u64, summed over all ranks) is >= min_total.
// kept_rows (room for m) receives the surviving feature indices in file order; *out the matrix of those rows.
extern "C" int sb_filter_genes(sb_mat *mat, const uint8_t *type_keep, uint64_t min_total, uint32_t *kept_rows, uint32_t *n_kept, sb_mat **out) {
    if (!mat || !n_kept) return sb_fail(SB_ERR_INVALID_ARG, "sb_filter_genes: NULL argument");
    sb_ctx *ctx = mat->ctx;
    SB_ENTER(ctx);
    if (out) *out = nullptr;
    const u32 m = mat->m;
    DevBuf<u64> tot;
    SB_TRY(tot.alloc(std::max<u32>(m, 1)));
    SB_TRY(mat_gene_sums_dev(mat, 0, nullptr, nullptr, tot.p, ctx->nranks > 1));
    std::vector<u64> h(m);
    if (m) SB_CUDA(cudaMemcpyAsync(h.data(), tot.p, (size_t)m * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<u32> keep;
    for (u32 g = 0; g < m; g++) {
        if (type_keep && !type_keep[g]) continue;  // remove_unlike (:101-104)
        if (h[g] < min_total) continue;            // :107-114
        keep.push_back(g);
    }
    *n_kept = (u32)keep.size();
    if (kept_rows)
        for (size_t i = 0; i < keep.size(); i++) kept_rows[i] = keep[i];
    if (out) return sb_select_rows(mat, keep.data(), (u32)keep.size(), out);
    return SB_OK;
}
