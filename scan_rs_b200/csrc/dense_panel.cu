// dense_panel.cu -- the dense half of the hybrid SpMM on the FP64 tensor path.
//
// The `gd` most expressed genes live in a u8 panel D[n x gd] (counts 1..15, 0 elsewhere; ~1/3 non-zero).
// The sparse gather spends a 160-byte operand row per nonzero and is pinned to the L1/shared pipe
// (DESIGN.md 3); a dense contraction reuses operands from registers.  Register-tiled DFMA still needed
// ~22 shared-memory wavefronts per 1280 FMAs (measured 30-43 % of FP64 peak); mma.sync.m8n8k4.f64 needs
// ~6, and the microbenchmark in profiles/microbench shows DMMA at the full FP64 rate (37 TFLOP/s).
//
//   dense_t : T[c, :] += sum_j L_c(D[c,j]) * rs[hot_j] * Y[hot_j, :]      M = cells, N = cols, K = genes
//   dense_n : P[hot_j, :] += sum_c L_c(D[c,j]) * X[c, :]                  M = genes, N = cols, K = cells
//   dense_moments : S1, S2[hot_j] += sum_c L, L^2                         (scalar, bandwidth-bound)
// L_c(v) = log_b(cs_c * v + 1), a 16-entry per-cell table (L_c(0) = 0): the same function of the same
// inputs as the sparse kernels' map, so both halves see identical values.
//
// m8n8k4 fragments (PTX ISA): A[row = lane/4][col = lane%4], B[row = lane%4][col = lane/4],
// C[row = lane/4][col = 2*(lane%4) + {0,1}].
#include "common.cuh"
#include "map.cuh"

#define LUT_STRIDE 17  // doubles per table row: spreads rows over the banks

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------- dense_t
#define DT_MC 4  // m-tiles (8 cells) per warp: a CTA of T threads works on tiles of T cells

template <int NT, int DT_THREADS>
__global__ void __launch_bounds__(DT_THREADS, 512 / DT_THREADS)  // <= 128 registers in both configurations
k_dense_t(const unsigned char *__restrict__ D, u32 gd, u64 n, const double *__restrict__ cs, int log_base, const u32 *__restrict__ hot_idx,
          const double *__restrict__ row_scale, const double *__restrict__ Y, u32 ldy, u32 col0, u32 w, u32 gchunk, double *__restrict__ out,
          u32 ldo, int atomic_out) {
    constexpr int DT_CELLS = DT_THREADS / 32 * 8 * DT_MC;
    constexpr int YS = NT * 8 + 1;  // row stride of the staged Y chunk (== 1 mod 4: conflict-free B fragments)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *Ys = reinterpret_cast<double *>(smem_raw);            // gchunk x YS
    double *lut = Ys + (size_t)gchunk * YS;                        // DT_CELLS x LUT_STRIDE
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int r = lane >> 2, k = lane & 3;
    const u64 ntiles = (n + DT_CELLS - 1) / DT_CELLS;
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const u64 tile0 = tile * DT_CELLS;
        __syncthreads();  // the previous tile's readers of lut / Ys are done
        {
            const u64 c = tile0 + t;
            const bool ok = c < n;
            const double s = ok ? cs[c] : 0.0;
            lut[t * LUT_STRIDE] = 0.0;
#pragma unroll 1
            for (u32 v = 1; v < SB_DENSE_LUT; v++) lut[t * LUT_STRIDE + v] = ok ? map_log_part(log_base, s, v, sb_log_table) : 0.0;
        }
        double c[DT_MC][NT][2];
#pragma unroll
        for (int mt = 0; mt < DT_MC; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++) c[mt][nt][0] = c[mt][nt][1] = 0.0;
        // this lane's cells: one per m-tile
        const unsigned char *drow[DT_MC];
        bool cok[DT_MC];
        int lrow[DT_MC];
#pragma unroll
        for (int mt = 0; mt < DT_MC; mt++) {
            const int cl = warp * (8 * DT_MC) + mt * 8 + r;
            const u64 cell = tile0 + cl;
            cok[mt] = cell < n;
            drow[mt] = D + (cok[mt] ? cell : 0) * (u64)gd + 4 * k;
            lrow[mt] = cl * LUT_STRIDE;
        }
        for (u32 g0 = 0; g0 < gd; g0 += gchunk) {
            const u32 gc = min(gchunk, gd - g0);
            __syncthreads();
            for (u32 i = t; i < gc * (NT * 8); i += DT_THREADS) {
                const u32 j = i / (NT * 8), cc = i - j * (NT * 8), col = col0 + cc;
                const u32 g = hot_idx[g0 + j];
                double y = col < w ? Y[(size_t)g * ldy + col] : 0.0;
                if (row_scale) y *= row_scale[g];
                Ys[j * YS + cc] = y;
            }
            __syncthreads();
            for (u32 g16 = 0; g16 < gc; g16 += 16) {
                // lane (r, k) holds the counts of genes g16 + 4k .. 4k+3 of its cells; k-step s uses gene g16 + 4k + s
                u32 dw[DT_MC];
#pragma unroll
                for (int mt = 0; mt < DT_MC; mt++) dw[mt] = cok[mt] ? *reinterpret_cast<const u32 *>(drow[mt] + g0 + g16) : 0u;
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    double b[NT];
                    const double *yr = Ys + (size_t)(g16 + 4 * k + s) * YS + r;
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) b[nt] = yr[nt * 8];
#pragma unroll
                    for (int mt = 0; mt < DT_MC; mt++) {
                        const double a = lut[lrow[mt] + ((dw[mt] >> (8 * s)) & 0xFFu)];
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) dmma(c[mt][nt][0], c[mt][nt][1], a, b[nt]);
                    }
                }
            }
        }
#pragma unroll
        for (int mt = 0; mt < DT_MC; mt++) {
            if (!cok[mt]) continue;
            const u64 cell = tile0 + warp * (8 * DT_MC) + mt * 8 + r;
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                const u32 col = col0 + nt * 8 + 2 * k;
                double *o = out + cell * (size_t)ldo + col;
                if (atomic_out) {  // overlapped with the sparse kernel: both add into the zeroed block
                    if (col < w) atomicAdd(o, c[mt][nt][0]);
                    if (col + 1 < w) atomicAdd(o + 1, c[mt][nt][1]);
                } else if (col + 1 < w) {
                    double2 cur = *reinterpret_cast<double2 *>(o);
                    cur.x += c[mt][nt][0];
                    cur.y += c[mt][nt][1];
                    *reinterpret_cast<double2 *>(o) = cur;
                } else if (col < w) {
                    o[0] += c[mt][nt][0];
                }
            }
        }
    }
}

// ---------------------------------------------------------------- dense_n
#define DN_MG 4  // m-tiles (8 genes) per warp: a CTA of T threads covers T panel columns and stages T/16 cells per sync

template <int NT, int DN_THREADS>
__global__ void __launch_bounds__(DN_THREADS, 512 / DN_THREADS)  // <= 128 registers in both configurations
k_dense_n(const unsigned char *__restrict__ D, u32 gd, u64 n, const double *__restrict__ cs, int log_base, const u32 *__restrict__ hot_idx,
          const double *__restrict__ X, u32 ldx, u32 col0, u32 w, double *__restrict__ P, u32 ldp) {
    constexpr int DN_GENES = DN_THREADS / 32 * 8 * DN_MG;
    constexpr int DN_GROUP = DN_THREADS / SB_DENSE_LUT;  // one table entry per thread per staged group
    constexpr int XS = NT * 8 + 4;  // row stride of the staged X rows (== 4 mod 16: conflict-free B fragments)
    __shared__ __align__(16) double Xs[2][DN_GROUP][XS];
    __shared__ __align__(16) double lut[2][DN_GROUP][LUT_STRIDE];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int g = lane >> 2, k = lane & 3;
    const u32 gene0 = blockIdx.x * DN_GENES + warp * (8 * DN_MG) + g * DN_MG;  // this lane's DN_MG consecutive panel columns
    const bool gok = gene0 < gd;                                               // gd is a multiple of 64: whole 4-gene groups
    const u64 per = (((n + gridDim.y - 1) / gridDim.y) + DN_GROUP - 1) / DN_GROUP * DN_GROUP;
    const u64 c_lo = min(n, (u64)blockIdx.y * per), c_hi = min(n, c_lo + per);
    double c[DN_MG][NT][2];
#pragma unroll
    for (int mt = 0; mt < DN_MG; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) c[mt][nt][0] = c[mt][nt][1] = 0.0;

    auto stage = [&](int buf, u64 cbase) {
        for (int i = t; i < DN_GROUP * (NT * 8); i += DN_THREADS) {
            const int cc = i / (NT * 8), j = i - cc * (NT * 8);
            const u64 cell = cbase + cc;
            Xs[buf][cc][j] = (cell < c_hi && col0 + j < w) ? X[cell * (size_t)ldx + col0 + j] : 0.0;
        }
        {
            const int cc = t / SB_DENSE_LUT, v = t % SB_DENSE_LUT;  // DN_GROUP * 16 == DN_THREADS
            const u64 cell = cbase + cc;
            lut[buf][cc][v] = (cell < c_hi && v > 0) ? map_log_part(log_base, cs[cell], (u32)v, sb_log_table) : 0.0;
        }
    };

    if (c_lo < c_hi) stage(0, c_lo);
    __syncthreads();
    int buf = 0;
    const unsigned char *dcol = D + gene0;
    for (u64 cb = c_lo; cb < c_hi; cb += DN_GROUP, buf ^= 1) {
        if (cb + DN_GROUP < c_hi) stage(buf ^ 1, cb + DN_GROUP);
        u32 dw[DN_GROUP / 4];
#pragma unroll
        for (int s = 0; s < DN_GROUP / 4; s++) {
            const u64 cell = cb + 4 * s + k;
            dw[s] = (gok && cell < c_hi) ? *reinterpret_cast<const u32 *>(dcol + cell * (u64)gd) : 0u;
        }
#pragma unroll
        for (int s = 0; s < DN_GROUP / 4; s++) {
            double b[NT];
#pragma unroll
            for (int nt = 0; nt < NT; nt++) b[nt] = Xs[buf][4 * s + k][nt * 8 + g];
            const double *lrow = lut[buf][4 * s + k];
#pragma unroll
            for (int mt = 0; mt < DN_MG; mt++) {
                const double a = lrow[(dw[s] >> (8 * mt)) & 0xFFu];
#pragma unroll
                for (int nt = 0; nt < NT; nt++) dmma(c[mt][nt][0], c[mt][nt][1], a, b[nt]);
            }
        }
        __syncthreads();
    }
    if (gok) {
#pragma unroll
        for (int mt = 0; mt < DN_MG; mt++) {
            double *o = P + (size_t)hot_idx[gene0 + mt] * ldp;
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                const u32 col = col0 + nt * 8 + 2 * k;
                if (col < w && c[mt][nt][0] != 0.0) atomicAdd(o + col, c[mt][nt][0]);
                if (col + 1 < w && c[mt][nt][1] != 0.0) atomicAdd(o + col + 1, c[mt][nt][1]);
            }
        }
    }
}

// ---------------------------------------------------------------- dense_moments (thread = 2 panel columns)
#define DM_THREADS 512
#define DM_GENES 1024
#define DM_GROUP 4
__global__ void __launch_bounds__(DM_THREADS, 1)
k_dense_moments(const unsigned char *__restrict__ D, u32 gd, u64 n, const double *__restrict__ cs, int log_base, const u32 *__restrict__ hot_idx,
                double *__restrict__ S1, double *__restrict__ S2) {
    __shared__ double lut[2][DM_GROUP][SB_DENSE_LUT];
    const int t = threadIdx.x;
    const u32 j0 = blockIdx.x * DM_GENES + t, j1 = j0 + DM_THREADS;
    const bool ok0 = j0 < gd, ok1 = j1 < gd;
    const u64 per = (n + gridDim.y - 1) / gridDim.y;
    const u64 c_lo = min(n, (u64)blockIdx.y * per), c_hi = min(n, c_lo + per);
    double s10 = 0.0, s20 = 0.0, s11 = 0.0, s21 = 0.0;
    auto stage = [&](int buf, u64 cbase) {
        if (t < DM_GROUP * SB_DENSE_LUT) {
            const int cc = t / SB_DENSE_LUT, v = t % SB_DENSE_LUT;
            const u64 cell = cbase + cc;
            lut[buf][cc][v] = (cell < c_hi && v > 0) ? map_log_part(log_base, cs[cell], (u32)v, sb_log_table) : 0.0;
        }
    };
    if (c_lo < c_hi) stage(0, c_lo);
    __syncthreads();
    int buf = 0;
    for (u64 cb = c_lo; cb < c_hi; cb += DM_GROUP, buf ^= 1) {
        if (cb + DM_GROUP < c_hi) stage(buf ^ 1, cb + DM_GROUP);
#pragma unroll
        for (int cc = 0; cc < DM_GROUP; cc++) {
            const u64 cell = cb + cc;
            const u32 v0 = (ok0 && cell < c_hi) ? D[cell * (u64)gd + j0] : 0u, v1 = (ok1 && cell < c_hi) ? D[cell * (u64)gd + j1] : 0u;
            const double l0 = lut[buf][cc][v0], l1 = lut[buf][cc][v1];
            s10 += l0;
            s20 = fma(l0, l0, s20);
            s11 += l1;
            s21 = fma(l1, l1, s21);
        }
        __syncthreads();
    }
    if (ok0) {
        atomicAdd(S1 + hot_idx[j0], s10);
        atomicAdd(S2 + hot_idx[j0], s20);
    }
    if (ok1) {
        atomicAdd(S1 + hot_idx[j1], s11);
        atomicAdd(S2 + hot_idx[j1], s21);
    }
}

// ---------------------------------------------------------------- launchers
// columns are processed in passes of at most 24 (three n-tiles); w = 20 is one pass with 4 padded columns
static inline u32 pass_width(u32 remaining) { return remaining >= 24 ? 24u : remaining; }

// `overlap`: half-size CTAs (256 threads, 32 K registers, < 190 KB shared) so that two CTAs of the sparse kernel stay
// resident beside each of them, atomic epilogue, launched on `stream` (the auxiliary stream)
// panel_i8.cu (experimental, default off)
bool dense_t_i8_usable(const sb_nmat *a, u32 w);
int dense_t_i8(sb_nmat *a, const double *Y, u32 ldy, u32 w, double *out, u32 ldo);

// planes.cu: the same three products over the bit planes (panel_mode 2)
int planes_t(sb_nmat *a, const double *Y, u32 ldy, u32 w, double *out, u32 ldo, const GatherLayout *gl = nullptr, const MapDev *mp = nullptr,
             long long *cycles = nullptr);
int planes_n(sb_nmat *a, const double *X, u32 ldx, u32 w, double *P, u32 ldp);
int planes_moments(sb_nmat *a, double *S1, double *S2);

int dense_t(sb_nmat *a, const double *Y, u32 ldy, u32 w, double *out, u32 ldo, cudaStream_t stream, bool overlap) {
    sb_mat *mt = a->mat;
    sb_ctx *ctx = mt->ctx;
    if (mt->gd == 0 || mt->n == 0 || w == 0) return SB_OK;
    if (mt->pl.active) {
        if (stream != ctx->stream) return sb_fail(SB_ERR_UNSUPPORTED, "dense_t: the plane kernels run on the library stream only");
        return planes_t(a, Y, ldy, w, out, ldo);
    }
    if (!overlap && stream == ctx->stream && dense_t_i8_usable(a, w)) return dense_t_i8(a, Y, ldy, w, out, ldo);
    const int threads = overlap ? 256 : 512;
    const size_t lut_bytes = (size_t)threads * LUT_STRIDE * sizeof(double);
    u64 ntiles = (mt->n + threads - 1) / threads;
    int blocks = (int)std::min<u64>(ntiles, (u64)ctx->sm_count);
    const double *rs = a->has_row_scale ? a->row_scale.p : nullptr;
    for (u32 col0 = 0; col0 < w;) {
        const u32 pw = pass_width(w - col0);
        const int nt = (int)((pw + 7) / 8);
        const u32 ys = nt * 8 + 1;
        const size_t budget = overlap ? 184 * 1024 : 220 * 1024;
        u32 gchunk = (u32)((budget - lut_bytes) / (ys * sizeof(double)));
        gchunk &= ~15u;
        if (gchunk > mt->gd) gchunk = mt->gd;
        const size_t smem = (size_t)gchunk * ys * sizeof(double) + lut_bytes;
        cudaError_t e = cudaSuccess;
#define LAUNCH_DT(NTV, THR)                                                                                                                    \
    e = cudaFuncSetAttribute(k_dense_t<NTV, THR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                     \
    if (e == cudaSuccess)                                                                                                                      \
        k_dense_t<NTV, THR><<<blocks, THR, smem, stream>>>(mt->D.p, mt->gd, mt->n, a->col_scale.p, a->log_base, mt->hot_idx.p, rs, Y, ldy, col0, w,   \
                                                            gchunk, out, ldo, overlap ? 1 : 0);
        if (overlap) {
            if (nt == 3) { LAUNCH_DT(3, 256) } else if (nt == 2) { LAUNCH_DT(2, 256) } else { LAUNCH_DT(1, 256) }
        } else {
            if (nt == 3) { LAUNCH_DT(3, 512) } else if (nt == 2) { LAUNCH_DT(2, 512) } else { LAUNCH_DT(1, 512) }
        }
#undef LAUNCH_DT
        if (e != cudaSuccess) return sb_fail(SB_ERR_CUDA, "dense_t: %s", cudaGetErrorString(e));
        count_launch(ctx);
        col0 += nt * 8;
    }
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

// `overlap`: 256-thread CTAs (28 K registers, 12 KB shared) that fit beside a 512-thread CTA of the sparse kernel
int dense_n(sb_nmat *a, const double *X, u32 ldx, u32 w, double *P, u32 ldp, cudaStream_t stream, bool overlap) {
    sb_mat *mt = a->mat;
    sb_ctx *ctx = mt->ctx;
    if (mt->gd == 0 || mt->n == 0 || w == 0) return SB_OK;
    if (mt->pl.active) {
        if (stream != ctx->stream) return sb_fail(SB_ERR_UNSUPPORTED, "dense_n: the plane kernels run on the library stream only");
        return planes_n(a, X, ldx, w, P, ldp);
    }
    const u32 threads = overlap ? 256 : 512;
    const u32 genes_per_cta = threads;  // threads / 32 warps x 32 genes
    const u32 group = threads / SB_DENSE_LUT;
    u32 gblocks = (mt->gd + genes_per_cta - 1) / genes_per_cta;
    u32 ranges = std::max<u32>(1, (u32)ctx->sm_count / gblocks);
    u64 max_ranges = (mt->n + group - 1) / group;
    if (ranges > max_ranges) ranges = (u32)max_ranges;
    dim3 grid(gblocks, ranges);
    for (u32 col0 = 0; col0 < w;) {
        const u32 pw = pass_width(w - col0);
        const int nt = (int)((pw + 7) / 8);
#define LAUNCH_DN(NTV, THR) \
    k_dense_n<NTV, THR><<<grid, THR, 0, stream>>>(mt->D.p, mt->gd, mt->n, a->col_scale.p, a->log_base, mt->hot_idx.p, X, ldx, col0, w, P, ldp);
        if (overlap) {
            if (nt == 3) { LAUNCH_DN(3, 256) } else if (nt == 2) { LAUNCH_DN(2, 256) } else { LAUNCH_DN(1, 256) }
        } else {
            if (nt == 3) { LAUNCH_DN(3, 512) } else if (nt == 2) { LAUNCH_DN(2, 512) } else { LAUNCH_DN(1, 512) }
        }
#undef LAUNCH_DN
        count_launch(ctx);
        col0 += nt * 8;
    }
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

int dense_moments(sb_nmat *a, double *S1, double *S2) {
    sb_mat *mt = a->mat;
    sb_ctx *ctx = mt->ctx;
    if (mt->gd == 0 || mt->n == 0) return SB_OK;
    if (mt->pl.active) return planes_moments(a, S1, S2);
    u32 gblocks = (mt->gd + DM_GENES - 1) / DM_GENES;
    u32 ranges = std::max<u32>(1, (u32)ctx->sm_count * 2 / gblocks);
    u64 max_ranges = (mt->n + DM_GROUP - 1) / DM_GROUP;
    if (ranges > max_ranges) ranges = (u32)max_ranges;
    k_dense_moments<<<dim3(gblocks, ranges), DM_THREADS, 0, ctx->stream>>>(mt->D.p, mt->gd, mt->n, a->col_scale.p, a->log_base, mt->hot_idx.p, S1, S2);
    count_launch(ctx);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}
