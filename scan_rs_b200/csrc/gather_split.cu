// gather_split.cu -- EXPERIMENTAL, default off (sb_set_option("gather_split", 1)), written at the end of round 1 without a
// GPU to run it on: it compiles for sm_100a but has not executed on hardware.  The product path is gather.cu.
//
// 94 % of the cold entries have a count of 1, and under the log chain such an entry needs nothing but its row offset:
// on the N side the staged row already is L_c(1) . X[c,:], on the T side the run is scaled by L_c(1) at the flush.  The
// combined stream still makes every warp run the general-entry prologue (log evaluation, factor staging, predicated
// factor loads: ~2 of the ~10 instructions per entry).  Here each side's stream is split once, lazily:
//   ones stream    : u32 {run key | row_local << 22}, 4 B per entry, walked by k_gather_ones (no value path at all)
//   general stream : the remaining uint2 entries, walked by the ordinary k_gather (gather.cu)
// Both keep the order of the combined stream, so panels / (block, panel) segments map through a prefix sum of the
// "count == 1" flags.
#include <cub/cub.cuh>

#include "common.cuh"
#include "map.cuh"

#define FULLMASK 0xffffffffu
#define GS_THREADS 1024
#define GS_NONE 0xFFFFFFFFu
#define GS_TILE 20u

int gather_run(sb_ctx *ctx, const GatherLayout &L, int mode, const MapDev &mp, u64 n_cells, const double *B, u32 ldb, u32 w, double *out, u32 ldo);

// ---------------------------------------------------------------- kernel over a ones stream
// Staging buffer of one 8-lane group: offs[8] u16 | keys[8] u32 | (T side) l1s[8] f64
template <int MODE>
struct GsStage {
    static constexpr u32 KEYS = 16u, L1 = 48u, STRIDE = MODE == 1 ? 128u : 64u;
};

template <int MODE, int TAIL>
__global__ void __launch_bounds__(GS_THREADS, 1)
k_gather_ones(const u32 *__restrict__ ent, const GUnit *__restrict__ units, const u32 *__restrict__ cta_first, u32 n_units, u32 rows, u64 n_cells,
              MapDev mp, const double *__restrict__ B, u32 ldb, u32 col0, u32 wt, u32 w, const u32 *__restrict__ slot_gene, double *__restrict__ out,
              u32 ldo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *XA = reinterpret_cast<double *>(smem_raw);                   // (rows + 1) x 16; row `rows` is all zero
    double *XB = XA + (size_t)(rows + 1) * 16;                            // (rows + 1) x 4 (TAIL only)
    unsigned char *stage_all = reinterpret_cast<unsigned char *>(XB + (TAIL ? (size_t)(rows + 1) * 4 : 0));

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int grp = lane >> 3, lig = lane & 7;
    const u32 xa_sa = (u32)__cvta_generic_to_shared(XA) + (u32)lig * 16u;
    const u32 xb_sa = (u32)__cvta_generic_to_shared(XB) + (u32)(lig & 3) * 8u;
    const u32 stage_grp = (u32)__cvta_generic_to_shared(stage_all) + (u32)(wib * 4 + grp) * GsStage<MODE>::STRIDE;
    const u32 zero_off16 = rows * 8u;
    const u32 cA = 2u * lig, cB = 16u + (lig & 3);
    const u32 ok0m = cA < wt && col0 + cA < w, ok1m = cA + 1 < wt && col0 + cA + 1 < w;
    const u32 okbm = TAIL && cB < wt && col0 + cB < w;
    const u32 wt_even = (wt + 1) & ~1u;
    double *const oA = out + col0 + cA;
    const long long tdelta = (long long)cB - (long long)cA;
    const u64 ldo64 = (u64)ldo;

    const u32 u_begin = cta_first[blockIdx.x], u_end = min(n_units, cta_first[blockIdx.x + 1]);
    u32 staged = GS_NONE;

    for (u32 ui = u_begin; ui < u_end; ui++) {
        const GUnit un = units[ui];
        if (un.panel != staged) {
            __syncthreads();
            constexpr u32 PAIRS = TAIL ? 10 : 8;
            for (u32 i = threadIdx.x; i < (rows + 1) * PAIRS; i += blockDim.x) {
                const u32 r = i / PAIRS, j2 = (i - r * PAIRS) * 2, col = col0 + j2;
                double2 val = make_double2(0.0, 0.0);
                if (r < rows && j2 < wt_even && col < w) {
                    if (MODE == 0) {  // the row of cell c is staged as L_c(1) . X[c,:]
                        const u64 cell = (u64)un.panel * rows + r;
                        if (cell < n_cells) {
                            val = *reinterpret_cast<const double2 *>(B + cell * (size_t)ldb + col);
                            const double l1 = mp.l1[cell];
                            val.x *= l1;
                            val.y *= l1;
                        }
                    } else {
                        const u32 g = slot_gene[(size_t)un.panel * rows + r];
                        if (g != GS_NONE) {
                            val = *reinterpret_cast<const double2 *>(B + (size_t)g * ldb + col);
                            if (mp.row) {
                                const double rs = mp.row[g];
                                val.x *= rs;
                                val.y *= rs;
                            }
                        }
                    }
                }
                if (j2 < 16) *reinterpret_cast<double2 *>(XA + (size_t)r * 16 + j2) = val;
                else *reinterpret_cast<double2 *>(XB + (size_t)r * 4 + (j2 - 16)) = val;
            }
            staged = un.panel;
            __syncthreads();
        }
        u64 span = (un.end - un.begin + nw - 1) / nw;
        span = (span + 31) & ~(u64)31;
        const u64 wb = min(un.end, un.begin + (u64)wib * span), we = min(un.end, wb + span);
        const u64 qspan = span >> 2;
        const u64 gb = min(we, wb + (u64)grp * qspan), ge = min(we, gb + qspan);
        const u32 nchunks = (u32)((min(we, wb + qspan) - wb + 7) >> 3);

        u32 cur = GS_NONE, carry = GS_NONE;
        double curL1 = 0.0;
        double a0 = 0.0, a1 = 0.0, bt = 0.0;

        auto flush = [&]() {
            if (cur == GS_NONE) return;
            if (MODE == 1) {
                a0 *= curL1;
                a1 *= curL1;
                bt *= curL1;
            }
            double *o = oA + (u64)cur * ldo64;
            asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p red.global.add.f64 [%0], %1; }" ::"l"(o), "d"(a0), "r"(ok0m) : "memory");
            asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p red.global.add.f64 [%0 + 8], %1; }" ::"l"(o), "d"(a1), "r"(ok1m) : "memory");
            if (TAIL) asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p red.global.add.f64 [%0], %1; }" ::"l"(o + tdelta), "d"(bt), "r"(okbm) : "memory");
            a0 = 0.0;
            a1 = 0.0;
            bt = 0.0;
        };

        u32 znext = (gb + lig < ge) ? __ldcs(ent + gb + lig) : GS_NONE;
        for (u32 ch = 0; ch < nchunks; ch++) {
            const u32 z = znext;
            const u64 kn = gb + (u64)(ch + 1) * 8 + lig;
            znext = (kn < ge) ? __ldcs(ent + kn) : GS_NONE;
            const bool valid = z != GS_NONE;
            const u32 key = z & SB_GENE_MASK, local = z >> SB_GENE_BITS;
            double l1v = 0.0;
            if (MODE == 1 && valid) l1v = mp.l1[key];
            u32 prev = __shfl_up_sync(FULLMASK, key, 1, 8);
            if (lig == 0) prev = carry;
            const bool head = valid && key != prev;
            carry = __shfl_sync(FULLMASK, key, 7, 8);
            const u32 myh = (__ballot_sync(FULLMASK, head) >> (8 * grp)) & 0xFFu;
            __syncwarp();
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(stage_grp + (u32)lig * 2u), "h"((unsigned short)(valid ? local * 8u : zero_off16)) : "memory");
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(stage_grp + GsStage<MODE>::KEYS + (u32)lig * 4u), "r"(key) : "memory");
            if (MODE == 1) asm volatile("st.shared.f64 [%0], %1;" ::"r"(stage_grp + GsStage<MODE>::L1 + (u32)lig * 8u), "d"(l1v) : "memory");
            __syncwarp();
            u32 ov[4];
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ov[0]), "=r"(ov[1]), "=r"(ov[2]), "=r"(ov[3]) : "r"(stage_grp));
#pragma unroll
            for (int s = 0; s < 8; s += 2) {
                const u32 off0 = (ov[s >> 1] & 0xFFFFu) << 4, off1 = (ov[s >> 1] >> 16) << 4;
                double tl = 0.0;
                if (TAIL) {
                    const u32 toff = (lig < 4 ? off0 : off1) >> 2;
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(tl) : "r"(xb_sa + toff));
                }
                double x0, x1;
                if (myh & (1u << s)) {
                    flush();
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cur) : "r"(stage_grp + GsStage<MODE>::KEYS + (u32)s * 4u));
                    if (MODE == 1) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(curL1) : "r"(stage_grp + GsStage<MODE>::L1 + (u32)s * 8u));
                }
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(xa_sa + off0));
                a0 += x0;
                a1 += x1;
                if (TAIL && lig < 4) bt += tl;
                if (myh & (2u << s)) {
                    flush();
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cur) : "r"(stage_grp + GsStage<MODE>::KEYS + (u32)(s + 1) * 4u));
                    if (MODE == 1) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(curL1) : "r"(stage_grp + GsStage<MODE>::L1 + (u32)(s + 1) * 8u));
                }
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(xa_sa + off1));
                a0 += x0;
                a1 += x1;
                if (TAIL && lig >= 4) bt += tl;
            }
        }
        flush();
    }
}

static size_t gs_smem(u32 rows, bool tail, int mode) {
    return (size_t)(rows + 1) * (tail ? 20 : 16) * 8 + (size_t)(GS_THREADS / 32) * 4 * (mode == 1 ? 128 : 64);
}

static int gather_ones_run(sb_ctx *ctx, const GatherLayout &L, const u32 *ones, int mode, const MapDev &mp, u64 n_cells, const double *B, u32 ldb,
                           u32 w, double *out, u32 ldo) {
    if (L.nnz == 0 || L.n_units == 0 || w == 0) return SB_OK;
    for (u32 col0 = 0; col0 < w; col0 += GS_TILE) {
        const u32 wt = std::min(GS_TILE, w - col0);
        const bool tail = wt > 16;
        const size_t smem = gs_smem(L.rows, tail, mode);
        cudaError_t e;
#define GS_LAUNCH(M, T)                                                                                                                     \
    e = cudaFuncSetAttribute(k_gather_ones<M, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                  \
    if (e == cudaSuccess)                                                                                                                   \
        k_gather_ones<M, T><<<L.grid, GS_THREADS, smem, ctx->stream>>>(ones, L.units.p, L.cta_first.p, L.n_units, L.rows, n_cells, mp, B, ldb, col0, \
                                                                       wt, w, L.slot_gene.p, out, ldo);
        if (mode == 0) {
            if (tail) { GS_LAUNCH(0, 1) } else { GS_LAUNCH(0, 0) }
        } else {
            if (tail) { GS_LAUNCH(1, 1) } else { GS_LAUNCH(1, 0) }
        }
#undef GS_LAUNCH
        if (e != cudaSuccess) return sb_fail(SB_ERR_CUDA, "gather_ones_run: %zu B of shared memory: %s", smem, cudaGetErrorString(e));
        count_launch(ctx);
    }
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

// ---------------------------------------------------------------- splitting a combined stream
__global__ void k_gs_flags(const uint2 *__restrict__ ent, u64 nnz, u32 *__restrict__ flag) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (u64)gridDim.x * blockDim.x) flag[i] = ent[i].y == 1u ? 1u : 0u;
}

__global__ void k_gs_scatter(const uint2 *__restrict__ ent, u64 nnz, const u32 *__restrict__ pos1, u32 *__restrict__ ones, uint2 *__restrict__ gen) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (u64)gridDim.x * blockDim.x) {
        const uint2 z = ent[i];
        const u32 p = pos1[i];
        if (z.y == 1u) ones[p] = z.x;
        else gen[i - p] = z;
    }
}

// ones-position of every boundary of the combined stream (a boundary at nnz maps to the number of ones)
__global__ void k_gs_bounds(const u64 *__restrict__ bnd, u32 nb, const u32 *__restrict__ pos1, u64 nnz, u32 n_ones, u64 *__restrict__ out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb) out[i] = bnd[i] < nnz ? (u64)pos1[bnd[i]] : (u64)n_ones;
}

static inline int gs_grid(u64 items, sb_ctx *ctx) {
    u64 blocks = (items + 255) / 256, cap = (u64)ctx->sm_count * 16;
    return (int)std::max<u64>(1, std::min(blocks, cap));
}

// Splits `ent` (order kept) and maps the boundaries `bnd` (ascending, last == nnz) into both streams.
static int split_stream(sb_ctx *ctx, const uint2 *ent, u64 nnz, const std::vector<u64> &bnd, DevBuf<u32> &ones, DevBuf<uint2> &gen,
                        std::vector<u64> &bnd_ones, std::vector<u64> &bnd_gen) {
    if (nnz > 0xFFFFFFF0ull) return sb_fail(SB_ERR_UNSUPPORTED, "gather_split: more than 2^32 entries in one stream");
    const u32 nb = (u32)bnd.size();
    bnd_ones.assign(nb, 0);
    bnd_gen.assign(nb, 0);
    DevBuf<u32> flag, pos1;
    SB_TRY(flag.alloc(nnz + 1));
    SB_TRY(pos1.alloc(nnz + 1));
    SB_CUDA(cudaMemsetAsync(flag.p, 0, (nnz + 1) * sizeof(u32), ctx->stream));
    if (nnz) {
        k_gs_flags<<<gs_grid(nnz, ctx), 256, 0, ctx->stream>>>(ent, nnz, flag.p);
        count_launch(ctx);
    }
    size_t tmp_bytes = 0;
    SB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flag.p, pos1.p, (unsigned long long)(nnz + 1), ctx->stream));
    DevBuf<char> tmp;
    SB_TRY(tmp.alloc(tmp_bytes));
    SB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, flag.p, pos1.p, (unsigned long long)(nnz + 1), ctx->stream));
    count_launch(ctx, false);
    u32 n_ones = 0;
    SB_CUDA(cudaMemcpyAsync(&n_ones, pos1.p + nnz, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    SB_TRY(ones.alloc(n_ones));
    SB_TRY(gen.alloc(nnz - n_ones));
    if (nnz) {
        k_gs_scatter<<<gs_grid(nnz, ctx), 256, 0, ctx->stream>>>(ent, nnz, pos1.p, ones.p, gen.p);
        count_launch(ctx);
    }
    DevBuf<u64> d_bnd, d_out;
    SB_TRY(d_bnd.alloc(nb));
    SB_TRY(d_out.alloc(nb));
    if (nb) {
        SB_CUDA(cudaMemcpyAsync(d_bnd.p, bnd.data(), nb * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
        k_gs_bounds<<<cdiv(nb, 256), 256, 0, ctx->stream>>>(d_bnd.p, nb, pos1.p, nnz, n_ones, d_out.p);
        count_launch(ctx);
        SB_CUDA(cudaMemcpyAsync(bnd_ones.data(), d_out.p, nb * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    }
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (u32 i = 0; i < nb; i++) bnd_gen[i] = bnd[i] - bnd_ones[i];
    return SB_OK;
}

static int upload_units(sb_ctx *ctx, GatherLayout &L, const std::vector<GUnit> &units, const std::vector<u32> &first) {
    L.n_units = (u32)units.size();
    L.grid = (u32)first.size() - 1;
    SB_TRY(L.units.alloc(units.size()));
    SB_TRY(L.cta_first.alloc(first.size()));
    if (!units.empty()) SB_CUDA(cudaMemcpyAsync(L.units.p, units.data(), units.size() * sizeof(GUnit), cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(cudaMemcpyAsync(L.cta_first.p, first.data(), first.size() * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    L.ready = true;
    return SB_OK;
}

// Builds the split streams of both sides from the combined gather layouts (lazily, at the first product that wants them).
int gather_split_build(sb_mat *mt) {
    sb_ctx *ctx = mt->ctx;
    GatherSplit &S = mt->split;
    S.tried = true;
    S.ready = false;
    if (!mt->gn.ready || !mt->gt.ready) return SB_OK;
    const u32 G = (u32)ctx->sm_count;
    const bool cold = mt->gd > 0;
    std::vector<GUnit> units;
    std::vector<u32> first;
    // ---- N side: boundaries = panel bases
    {
        std::vector<u64> base((size_t)mt->np + 1, 0), b1, bg;
        const u64 *d_base = cold ? mt->cold_gm_base.p : mt->gm_base.p;
        if (mt->np) SB_CUDA(cudaMemcpyAsync(base.data(), d_base, base.size() * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(cudaStreamSynchronize(ctx->stream));
        SB_TRY(split_stream(ctx, mt->gn.ent, mt->gn.nnz, base, S.n_ones, S.n_gen, b1, bg));
        S.n1.rows = S.ng.rows = mt->gn.rows;
        S.n1.npanels = S.ng.npanels = mt->gn.npanels;
        S.n1.nnz = b1.back();
        S.ng.nnz = bg.back();
        S.ng.ent = S.n_gen.p;
        gather_units_n(b1, S.n1.nnz, G, units, first);
        SB_TRY(upload_units(ctx, S.n1, units, first));
        gather_units_n(bg, S.ng.nnz, G, units, first);
        SB_TRY(upload_units(ctx, S.ng, units, first));
    }
    // ---- T side: boundaries = (block, panel) segment starts
    {
        const size_t nseg = mt->t_seg_len.size();
        std::vector<u64> pos(nseg + 1, 0), b1, bg;
        for (size_t k = 0; k < nseg; k++) pos[k + 1] = pos[k] + mt->t_seg_len[k];
        if (pos.back() != mt->gt.nnz) return sb_fail(SB_ERR_UNSUPPORTED, "gather_split: T-side segment table does not match the stream");
        SB_TRY(split_stream(ctx, mt->gt.ent, mt->gt.nnz, pos, S.t_ones, S.t_gen, b1, bg));
        std::vector<u64> len1(nseg), leng(nseg), run1(nseg), rung(nseg);
        for (size_t k = 0; k < nseg; k++) {
            len1[k] = b1[k + 1] - b1[k];
            leng[k] = bg[k + 1] - bg[k];
            run1[k] = std::min(mt->t_seg_runs[k], len1[k]);  // run counts per stream: bounded by the combined stream's
            rung[k] = std::min(mt->t_seg_runs[k], leng[k]);
        }
        S.t1.rows = S.tg.rows = mt->gt.rows;
        S.t1.npanels = S.tg.npanels = mt->gt.npanels;
        S.t1.nnz = b1.back();
        S.tg.nnz = bg.back();
        S.tg.ent = S.t_gen.p;
        const size_t nslots = (size_t)mt->gt.npanels * mt->gt.rows;
        SB_TRY(S.t1.slot_gene.alloc(nslots));
        SB_TRY(S.tg.slot_gene.alloc(nslots));
        SB_CUDA(cudaMemcpyAsync(S.t1.slot_gene.p, mt->gt.slot_gene.p, nslots * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
        SB_CUDA(cudaMemcpyAsync(S.tg.slot_gene.p, mt->gt.slot_gene.p, nslots * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
        gather_units_t(len1, run1, mt->gt.npanels, G, 5.0, units, first);
        SB_TRY(upload_units(ctx, S.t1, units, first));
        gather_units_t(leng, rung, mt->gt.npanels, G, 5.0, units, first);
        SB_TRY(upload_units(ctx, S.tg, units, first));
    }
    S.ready = true;
    return SB_OK;
}

// one product over the split streams of a side (mode 0: N, 1: T); log-chain maps only
int gather_split_run(sb_mat *mt, int mode, const MapDev &mp, const double *B, u32 ldb, u32 w, double *out, u32 ldo) {
    sb_ctx *ctx = mt->ctx;
    GatherSplit &S = mt->split;
    if (mode == 0) {
        SB_TRY(gather_ones_run(ctx, S.n1, S.n_ones.p, 0, mp, mt->n, B, ldb, w, out, ldo));
        return gather_run(ctx, S.ng, 0, mp, mt->n, B, ldb, w, out, ldo);
    }
    SB_TRY(gather_ones_run(ctx, S.t1, S.t_ones.p, 1, mp, mt->n, B, ldb, w, out, ldo));
    return gather_run(ctx, S.tg, 1, mp, mt->n, B, ldb, w, out, ldo);
}
