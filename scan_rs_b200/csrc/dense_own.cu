// dense_own.cu -- the repo's own kernels for the tall-skinny dense steps of the Krylov loop, replacing
// the cuSOLVER / cuBLAS calls of dense.cu on the common path:
//   syrk_tall      G = A^T A for a tall row-major block (FP64 mma.sync m8n8k4, shared-memory staged, symmetric blocks once)
//   gemm_tall      Out = A . S for a tall block and a small matrix (same tensor path)
//   chol_inv       Cholesky of a small SPD matrix (optionally shifted) and the inverse of its factor, one CTA
//   qr_chol        thin QR of a tall block by shifted CholeskyQR3 (Fukaya et al. 2020): three rounds of Gram -> Cholesky -> A R^-1.
//                  Replaces LAPACK dgeqrf + dorgqr at bk_svd.rs:94,98,123,127.  The Krylov basis K is ill-conditioned by
//                  construction (cond 1e8..1e9): plain CholeskyQR fails there, the shifted first round brings the
//                  condition number down to ~1e5 and two more rounds restore orthogonality to rounding.  Everything is a GEMM-shaped
//                  kernel or a one-CTA kernel on a b x b matrix: no host synchronisation, no Householder panel latency, and nothing
//                  that has to be replicated at length over ranks.  Breakdown (a non-positive pivot: exactly rank-deficient input)
//                  raises a device flag that the PCA driver reads with its outputs; it then reruns with the Householder path.
#include "common.cuh"

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------- G += A^T A
#define SY_THREADS 512
#define SY_BLK 128
#define SY_RC 32
#define SY_STRIDE 132  // == 4 mod 16 doubles: the four k rows of a fragment (8-byte words, 16 per 128-byte line) land in disjoint banks
#define SY_STAGES 3    // shared-memory ring of row stages filled by cp.async

// 16-byte asynchronous copy global -> shared; bytes past `src_bytes` (0 .. 16) are zero-filled
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, u32 src_bytes) {
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc, u32 src_bytes) {
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A warp owns a 32 x 32 region (4 x 4 MMA tiles of 8 x 8) of the 128 x 128 block.  Tiles past the matrix order, and in a diagonal
// block the tiles strictly below the tile diagonal (the mirror image is stored from above), are skipped: at w = 100 that leaves
// 91 of 256 tiles.  Warps are tied to a scheduler by warp id modulo 4, so `wmap` (host: greedy longest-first over the four
// schedulers) says which region each warp takes: 26 tiles on the busiest scheduler instead of 64.
struct SyWarpMap { unsigned char region[16]; };  // region = wr << 2 | wc

// grid (row chunks, block pairs bi <= bj).  G: column-major w x w per row chunk; only blocks with bi <= bj are written.
// Row stages of 16 rows travel global -> shared by cp.async through a ring of SY_STAGES slots (the first version staged through
// registers one stage ahead and spent its time waiting: 1.39 ms for a 1.3M x 100 block at 37 % issue activity, one barrier and one
// exposed global round trip per 16 rows; profiles/exp_dense_tileskip_r02.log).
__global__ void __launch_bounds__(SY_THREADS, 1)
k_syrk_tall(const double *__restrict__ A, u64 rows, u32 w, u32 ld, double *__restrict__ G, u32 nb, SyWarpMap wmap) {
    extern __shared__ __align__(16) double sy_smem[];
    double(*sI)[SY_RC][SY_STRIDE] = reinterpret_cast<double(*)[SY_RC][SY_STRIDE]>(sy_smem);
    double(*sJ)[SY_RC][SY_STRIDE] = reinterpret_cast<double(*)[SY_RC][SY_STRIDE]>(sy_smem + SY_STAGES * SY_RC * SY_STRIDE);
    // decode the block pair
    u32 bi = 0, bj = 0;
    {
        u32 p = blockIdx.y;
        for (bi = 0; bi < nb; bi++) {
            const u32 cnt = nb - bi;
            if (p < cnt) {
                bj = bi + p;
                break;
            }
            p -= cnt;
        }
    }
    const bool diag = bi == bj;
    const u32 I0 = bi * SY_BLK, J0 = bj * SY_BLK;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int wr = diag ? (wmap.region[warp] >> 2) : (warp >> 2), wc = diag ? (wmap.region[warp] & 3) : (warp & 3);
    const int fr = lane >> 2, fk = lane & 3;
    u32 tmask = 0;  // bit 4 mt + nt: this warp computes tile (mt, nt)
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            const u32 ti = wr * 4 + mt, tj = wc * 4 + nt;
            if (I0 + ti * 8 < w && J0 + tj * 8 < w && (!diag || ti <= tj)) tmask |= 1u << (4 * mt + nt);
        }
    u64 per = (rows + gridDim.x - 1) / gridDim.x;
    per = (per + SY_RC - 1) / SY_RC * SY_RC;
    const u64 r_lo = min(rows, (u64)blockIdx.x * per), r_hi = min(rows, r_lo + per);
    if (r_lo >= r_hi) return;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

    // each thread moves 4 doubles (two 16-byte copies) per operand per stage: row = t / 32, columns 4 (t % 32) .. + 3
    const int lrow = t >> 5, lcol = (t & 31) * 4;
    const bool aligned16 = (ld & 1u) == 0 && ((size_t)A & 15u) == 0;  // row starts on 16 bytes (else: 8-byte copies)
    auto issue = [&](u64 r0, int slot) {
#pragma unroll
        for (int rr = 0; rr < SY_RC; rr += SY_THREADS / 32) {
            const int sr = lrow + rr;
            const u64 r = r0 + sr;
            const bool rok = r < r_hi;
            const double *rowp = A + (rok ? r : r_lo) * (size_t)ld;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const u32 ci = I0 + lcol + 2 * h, cj = J0 + lcol + 2 * h;
                const u32 bi_ = !rok || ci >= w ? 0u : min(16u, (w - ci) * 8u), bj_ = !rok || cj >= w ? 0u : min(16u, (w - cj) * 8u);
                if (aligned16) {
                    cp_async16(&sI[slot][sr][lcol + 2 * h], rowp + (bi_ ? ci : 0), bi_);
                    if (!diag) cp_async16(&sJ[slot][sr][lcol + 2 * h], rowp + (bj_ ? cj : 0), bj_);
                } else {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const u32 b1 = bi_ > 8u * e ? 8u : 0u, b2 = bj_ > 8u * e ? 8u : 0u;
                        cp_async8(&sI[slot][sr][lcol + 2 * h + e], rowp + (b1 ? ci + e : 0), b1);
                        if (!diag) cp_async8(&sJ[slot][sr][lcol + 2 * h + e], rowp + (b2 ? cj + e : 0), b2);
                    }
                }
            }
        }
    };
    const u64 nst = (r_hi - r_lo + SY_RC - 1) / SY_RC;
    for (int p = 0; p < SY_STAGES - 1; p++) {
        if ((u64)p < nst) issue(r_lo + (u64)p * SY_RC, p);
        cp_async_commit();
    }
    for (u64 st = 0; st < nst; st++) {
        cp_async_wait<SY_STAGES - 2>();  // stage st has landed (for this thread's copies; the barrier publishes everybody's)
        __syncthreads();                 // ... and every warp is done with stage st - 1, whose slot is refilled next
        if (st + SY_STAGES - 1 < nst) issue(r_lo + (st + SY_STAGES - 1) * SY_RC, (int)((st + SY_STAGES - 1) % SY_STAGES));
        cp_async_commit();
        const int buf = (int)(st % SY_STAGES);
        double(*sB)[SY_STRIDE] = diag ? sI[buf] : sJ[buf];
#pragma unroll
        for (int ks = 0; ks < SY_RC / 4; ks++) {
            double a[4], b[4];
#pragma unroll
            for (int mt = 0; mt < 4; mt++) a[mt] = sI[buf][ks * 4 + fk][(wr * 4 + mt) * 8 + fr];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) b[nt] = sB[ks * 4 + fk][(wc * 4 + nt) * 8 + fr];
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++)
                    if (tmask & (1u << (4 * mt + nt))) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
    }
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            if (!(tmask & (1u << (4 * mt + nt)))) continue;
            const u32 i = I0 + (wr * 4 + mt) * 8 + fr;
            const u32 j = J0 + (wc * 4 + nt) * 8 + 2 * fk;
            // plain stores into this row chunk's own copy of G: k_syrk_reduce adds the chunks up in a fixed order, so the result is
            // bit-reproducible -- the gene-side blocks are factored redundantly on every rank and must come out identical there
            double *Gc = G + (size_t)blockIdx.x * w * w;
            if (i < w && j < w) Gc[(size_t)j * w + i] = acc[mt][nt][0];
            if (i < w && j + 1 < w) Gc[(size_t)(j + 1) * w + i] = acc[mt][nt][1];
            if (diag && wr * 4 + mt < wc * 4 + nt) {  // the skipped tile below the diagonal is this one's transpose
                if (i < w && j < w) Gc[(size_t)i * w + j] = acc[mt][nt][0];
                if (i < w && j + 1 < w) Gc[(size_t)i * w + j + 1] = acc[mt][nt][1];
            }
        }
}

// which 32 x 32 region of a diagonal block every warp takes: regions sorted by their number of live tiles, each handed to the
// scheduler (warp id modulo 4) with the least work so far that still has a free warp
static SyWarpMap syrk_warp_map(u32 w_in_block) {
    int cnt[16], order[16];
    for (int wr = 0; wr < 4; wr++)
        for (int wc = 0; wc < 4; wc++) {
            int c = 0;
            for (int mt = 0; mt < 4; mt++)
                for (int nt = 0; nt < 4; nt++) {
                    const u32 ti = wr * 4 + mt, tj = wc * 4 + nt;
                    if (ti * 8 < w_in_block && tj * 8 < w_in_block && ti <= tj) c++;
                }
            cnt[wr * 4 + wc] = c;
            order[wr * 4 + wc] = wr * 4 + wc;
        }
    std::stable_sort(order, order + 16, [&](int x, int y) { return cnt[x] > cnt[y]; });
    int load[4] = {0, 0, 0, 0}, used[4] = {0, 0, 0, 0};
    SyWarpMap m;
    for (int k = 0; k < 16; k++) {
        int best = -1;
        for (int sch = 0; sch < 4; sch++)
            if (used[sch] < 4 && (best < 0 || load[sch] < load[best])) best = sch;
        m.region[best + 4 * used[best]] = (unsigned char)order[k];
        load[best] += cnt[order[k]];
        used[best]++;
    }
    return m;
}

// ---------------------------------------------------------------- G = A^T A for w <= 128 (one diagonal block): live tiles only
// The general kernel above gives every warp a 32 x 32 region and masks the MMA tiles it does not need -- but a predicated-off
// DMMA still occupies the FP64 tensor pipe (ncu: sm__inst_executed_pipe_tensor_subpipe_dmma 86 % with 91 of 256 tiles live, and
// the same duration as without the mask).  Here the T (T + 1) / 2 tiles on or above the tile diagonal (T = ceil(w / 8); 91 at
// w = 100) are dealt out to the 16 warps as a list, NQ tiles per warp, so only live tiles are ever issued; a slot past the end of
// the list repeats tile 0 and is not stored.  Same row-stage ring, same per-chunk copies of G, same fixed-order reduction.
struct SyTileList { unsigned char ti[16][9], tj[16][9], live[16]; };

template <int NQ>
__global__ void __launch_bounds__(SY_THREADS, 1)
k_syrk_diag(const double *__restrict__ A, u64 rows, u32 w, u32 ld, double *__restrict__ G, SyTileList tl) {
    extern __shared__ __align__(16) double sy_smem[];
    double(*sI)[SY_RC][SY_STRIDE] = reinterpret_cast<double(*)[SY_RC][SY_STRIDE]>(sy_smem);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    u64 per = (rows + gridDim.x - 1) / gridDim.x;
    per = (per + SY_RC - 1) / SY_RC * SY_RC;
    const u64 r_lo = min(rows, (u64)blockIdx.x * per), r_hi = min(rows, r_lo + per);
    if (r_lo >= r_hi) return;
    u32 ca[NQ], cb[NQ];  // column (in doubles) of this lane's fragment element in the tiles' row / column operand
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        ca[q] = tl.ti[warp][q] * 8u + fr;
        cb[q] = tl.tj[warp][q] * 8u + fr;
    }
    const u32 nlive = tl.live[warp];
    double acc[NQ][2];
#pragma unroll
    for (int q = 0; q < NQ; q++) acc[q][0] = acc[q][1] = 0.0;
    const int lrow = t >> 5, lcol = (t & 31) * 4;
    const bool aligned16 = (ld & 1u) == 0 && ((size_t)A & 15u) == 0;
    auto issue = [&](u64 r0, int slot) {
#pragma unroll
        for (int rr = 0; rr < SY_RC; rr += SY_THREADS / 32) {
            const int sr = lrow + rr;
            const u64 r = r0 + sr;
            const bool rok = r < r_hi;
            const double *rowp = A + (rok ? r : r_lo) * (size_t)ld;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const u32 ci = lcol + 2 * h;
                const u32 bi_ = !rok || ci >= w ? 0u : min(16u, (w - ci) * 8u);
                if (aligned16) {
                    cp_async16(&sI[slot][sr][lcol + 2 * h], rowp + (bi_ ? ci : 0), bi_);
                } else {
                    cp_async8(&sI[slot][sr][lcol + 2 * h], rowp + (bi_ ? ci : 0), bi_ ? 8u : 0u);
                    cp_async8(&sI[slot][sr][lcol + 2 * h + 1], rowp + (bi_ > 8u ? ci + 1 : 0), bi_ > 8u ? 8u : 0u);
                }
            }
        }
    };
    const u64 nst = (r_hi - r_lo + SY_RC - 1) / SY_RC;
    for (int p = 0; p < SY_STAGES - 1; p++) {
        if ((u64)p < nst) issue(r_lo + (u64)p * SY_RC, p);
        cp_async_commit();
    }
    for (u64 st = 0; st < nst; st++) {
        cp_async_wait<SY_STAGES - 2>();
        __syncthreads();
        if (st + SY_STAGES - 1 < nst) issue(r_lo + (st + SY_STAGES - 1) * SY_RC, (int)((st + SY_STAGES - 1) % SY_STAGES));
        cp_async_commit();
        const int buf = (int)(st % SY_STAGES);
#pragma unroll
        for (int ks = 0; ks < SY_RC / 4; ks++) {
            const double *krow = sI[buf][ks * 4 + fk];
#pragma unroll
            for (int q = 0; q < NQ; q++) dmma884(acc[q][0], acc[q][1], krow[ca[q]], krow[cb[q]]);
        }
    }
    double *Gc = G + (size_t)blockIdx.x * w * w;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        if ((u32)q >= nlive) continue;
        const u32 ti = tl.ti[warp][q], tj = tl.tj[warp][q];
        const u32 i = ti * 8 + fr, j = tj * 8 + 2 * fk;
        if (i < w && j < w) Gc[(size_t)j * w + i] = acc[q][0];
        if (i < w && j + 1 < w) Gc[(size_t)(j + 1) * w + i] = acc[q][1];
        if (ti < tj) {  // mirror image below the tile diagonal
            if (i < w && j < w) Gc[(size_t)i * w + j] = acc[q][0];
            if (i < w && j + 1 < w) Gc[(size_t)i * w + j + 1] = acc[q][1];
        }
    }
}

// the live tiles of a w x w Gram matrix, dealt round-robin to the 16 warps (warp id modulo 4 is the scheduler: round-robin keeps
// the four of them level); returns the tiles per warp
static int syrk_tile_list(u32 w, SyTileList &tl) {
    const int T = (int)((w + 7) / 8);
    memset(&tl, 0, sizeof(tl));
    int n = 0;
    for (int ti = 0; ti < T; ti++)
        for (int tj = ti; tj < T; tj++, n++) {
            const int wp = n % 16, q = n / 16;
            tl.ti[wp][q] = (unsigned char)ti;
            tl.tj[wp][q] = (unsigned char)tj;
            tl.live[wp] = (unsigned char)(q + 1);
        }
    return (n + 15) / 16;
}

// G[e] = sum over the row chunks of the blocks on or above the block diagonal.  One warp per element: lane l adds chunks l, l + 32,
// ... in order, then a fixed shuffle tree -- the same association on every rank and every run.
__global__ void k_syrk_reduce(const double *__restrict__ parts, u32 chunks, u32 w, double *__restrict__ G) {
    const u64 idx = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    if (idx >= (u64)w * w) return;
    const u32 j = (u32)(idx / w), i = (u32)(idx - (u64)j * w);
    double s = 0.0;
    if (i / SY_BLK <= j / SY_BLK) {
        for (u32 c = lane; c < chunks; c += 32) s += parts[(size_t)c * w * w + idx];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    if (lane == 0) G[idx] = s;
}

// fills the blocks below the block diagonal from their transposes
__global__ void k_symmetrize_blocks(double *__restrict__ G, u32 w) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (u64)w * w) return;
    const u32 j = (u32)(idx / w), i = (u32)(idx - (u64)j * w);  // element (i, j)
    if (i / SY_BLK > j / SY_BLK) G[idx] = G[(size_t)i * w + j];
}

int syrk_tall(sb_ctx *ctx, const double *A, u64 rows, u32 w, u32 ld, double *G) {
    SB_CUDA(cudaMemsetAsync(G, 0, (size_t)w * w * sizeof(double), ctx->stream));
    if (rows == 0 || w == 0) return SB_OK;
    const u32 nb = (w + SY_BLK - 1) / SY_BLK;
    const u32 pairs = nb * (nb + 1) / 2;
    u32 chunks = std::max<u32>(1, (u32)std::min<u64>((rows + 4 * SY_RC - 1) / (4 * SY_RC), std::max<u32>(1, (u32)ctx->sm_count * 2 / pairs)));
    const size_t smem = (size_t)2 * SY_STAGES * SY_RC * SY_STRIDE * sizeof(double);
    SB_CUDA(cudaFuncSetAttribute(k_syrk_tall, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // row chunks that get no rows return early: only the first `valid` copies are written (the kernel's own split, restated)
    u64 per = (rows + chunks - 1) / chunks;
    per = (per + SY_RC - 1) / SY_RC * SY_RC;
    const u32 valid = (u32)((rows + per - 1) / per);
    // the per-chunk copies live in a grow-only buffer of the context: a fresh stream-ordered allocation per call (a few hundred KB to
    // 24 MB, a dozen times per PCA) fragmented the pool and the large per-pass buffers of the products paid for its remapping
    // (measured: 18.7 -> 55 ms per step at 162k cells)
    SB_TRY(ctx->syrk_parts.ensure((size_t)valid * w * w));
    double *parts = ctx->syrk_parts.p;
    if (nb == 1) {  // one diagonal block: only its live tiles are issued
        SyTileList tl;
        const int nq = syrk_tile_list(w, tl);
        const size_t smem1 = (size_t)SY_STAGES * SY_RC * SY_STRIDE * sizeof(double);
        if (nq <= 3) {
            SB_CUDA(cudaFuncSetAttribute(k_syrk_diag<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
            k_syrk_diag<3><<<chunks, SY_THREADS, smem1, ctx->stream>>>(A, rows, w, ld, parts, tl);
        } else if (nq <= 6) {
            SB_CUDA(cudaFuncSetAttribute(k_syrk_diag<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
            k_syrk_diag<6><<<chunks, SY_THREADS, smem1, ctx->stream>>>(A, rows, w, ld, parts, tl);
        } else {
            SB_CUDA(cudaFuncSetAttribute(k_syrk_diag<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
            k_syrk_diag<9><<<chunks, SY_THREADS, smem1, ctx->stream>>>(A, rows, w, ld, parts, tl);
        }
    } else {
        // (several blocks per side: the map of a full diagonal block; a partial last block is then only balanced approximately)
        const SyWarpMap wmap = syrk_warp_map(SY_BLK);
        k_syrk_tall<<<dim3(chunks, pairs), SY_THREADS, smem, ctx->stream>>>(A, rows, w, ld, parts, nb, wmap);
    }
    k_syrk_reduce<<<cdiv((u64)w * w * 32, 256), 256, 0, ctx->stream>>>(parts, valid, w, G);
    count_launch(ctx); count_launch(ctx);
    if (nb > 1) {
        k_symmetrize_blocks<<<cdiv((u64)w * w, 256), 256, 0, ctx->stream>>>(G, w);
        count_launch(ctx);
    }
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

// ---------------------------------------------------------------- Out = A . S
#define GM_THREADS 512
#define GM_KC 16
#define GM_ASTRIDE 20   // == 4 mod 32 doubles: eight rows x four k land in 32 distinct doubles
#define GM_SSTRIDE 132  // == 4 mod 16 doubles: the four k rows of a B fragment land in disjoint banks

// A row-major rows x w (lda); S column-major w x k (lds); Out row-major rows x k (ldo; pad columns k..ldo-1 are zeroed).
// s_upper: S is upper triangular (the R^-1 of the CholeskyQR rounds and of the projection): the k steps below a column tile's
// diagonal multiply zeros and are skipped, as are column tiles past k.  Warps are laid out so that every scheduler (warp id modulo
// 4) gets one warp of each column group -- the work per column group grows with its index once the triangle is skipped.
// Operand chunks travel global -> shared by cp.async through a ring of GM_STAGES slots (see k_syrk_tall).
#define GM_STAGES 3
__global__ void __launch_bounds__(GM_THREADS, 1)
k_gemm_tall(const double *__restrict__ A, u64 rows, u32 w, u32 lda, const double *__restrict__ S, u32 k, u32 lds, double *__restrict__ Out, u32 ldo,
            int s_upper) {
    extern __shared__ __align__(16) double gm_smem[];
    double(*sA)[128][GM_ASTRIDE] = reinterpret_cast<double(*)[128][GM_ASTRIDE]>(gm_smem);
    double(*sS)[GM_KC][GM_SSTRIDE] = reinterpret_cast<double(*)[GM_KC][GM_SSTRIDE]>(gm_smem + GM_STAGES * 128 * GM_ASTRIDE);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int wr = warp & 3, wc = warp >> 2;
    const int fr = lane >> 2, fk = lane & 3;
    const u64 r0 = (u64)blockIdx.x * 128;
    const u32 c0 = blockIdx.y * 128;
    u32 klim[4];  // k steps [0, klim) contribute to column tile nt of this warp (0: the tile lies past k)
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
        const u32 c8 = c0 + (wc * 4 + nt) * 8;
        klim[nt] = c8 >= k ? 0u : (s_upper ? min(w, c8 + 8u) : w);
    }
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    // staging: A chunk 128 rows x 16 k: thread -> row t / 4, k 4 (t % 4) .. + 3; S chunk 16 k x 128 cols: thread -> col t / 4, k 4 (t % 4) .. + 3
    const int arow = t >> 2, ak = (t & 3) * 4;
    const bool aligned16 = (lda & 1u) == 0 && ((size_t)A & 15u) == 0;
    auto issue = [&](u32 k0, int slot) {
        const u64 r = r0 + arow;
        const u32 c = c0 + arow;
        const bool rok = r < rows, cok = c < k;
        const double *arowp = A + (rok ? r : 0) * (size_t)lda;
        const double *scolp = S + (size_t)(cok ? c : 0) * lds;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const u32 kk = k0 + ak + 2 * h;
            const u32 ab = !rok || kk >= w ? 0u : min(16u, (w - kk) * 8u);
            if (aligned16) {
                cp_async16(&sA[slot][arow][ak + 2 * h], arowp + (ab ? kk : 0), ab);
            } else {
                cp_async8(&sA[slot][arow][ak + 2 * h], arowp + (ab ? kk : 0), ab ? 8u : 0u);
                cp_async8(&sA[slot][arow][ak + 2 * h + 1], arowp + (ab > 8u ? kk + 1 : 0), ab > 8u ? 8u : 0u);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const u32 kk = k0 + ak + i;
            const bool ok = cok && kk < w;
            cp_async8(&sS[slot][ak + i][arow], scolp + (ok ? kk : 0), ok ? 8u : 0u);
        }
    };
    const u32 nst = (w + GM_KC - 1) / GM_KC;
    for (int p = 0; p < GM_STAGES - 1; p++) {
        if ((u32)p < nst) issue((u32)p * GM_KC, p);
        cp_async_commit();
    }
    for (u32 st = 0; st < nst; st++) {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();
        if (st + GM_STAGES - 1 < nst) issue((st + GM_STAGES - 1) * GM_KC, (int)((st + GM_STAGES - 1) % GM_STAGES));
        cp_async_commit();
        const int buf = (int)(st % GM_STAGES);
        const u32 k0 = st * GM_KC;
#pragma unroll
        for (int ks = 0; ks < GM_KC / 4; ks++) {
            double a[4], b[4];
#pragma unroll
            for (int mt = 0; mt < 4; mt++) a[mt] = sA[buf][(wr * 4 + mt) * 8 + fr][ks * 4 + fk];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) b[nt] = sS[buf][ks * 4 + fk][(wc * 4 + nt) * 8 + fr];
#pragma unroll
            for (int nt = 0; nt < 4; nt++)
                if (k0 + ks * 4 < klim[nt]) {
#pragma unroll
                    for (int mt = 0; mt < 4; mt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
                }
        }
    }
    // (in place use, Out == A: every copy of this CTA's rows has landed -- the last stage was waited for above)
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
        const u64 r = r0 + (wr * 4 + mt) * 8 + fr;
        if (r >= rows) continue;
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            const u32 c = c0 + (wc * 4 + nt) * 8 + 2 * fk;  // even; ldo is even
            if (c + 1 < ldo) *reinterpret_cast<double2 *>(Out + r * (size_t)ldo + c) = make_double2(c < k ? acc[mt][nt][0] : 0.0, c + 1 < k ? acc[mt][nt][1] : 0.0);
        }
    }
}

// ---------------------------------------------------------------- Out = A . S for w <= 128, k <= 128: live tiles only
// One warp per 8-row tile of the CTA's 128 rows, CT column tiles of 8 each; the k loop is unrolled completely (at most 8 stages of
// 16), so with an upper-triangular S the (k step, column tile) pairs that multiply zeros are left out at COMPILE time: a
// predicated-off DMMA still occupies the FP64 tensor pipe (see k_syrk_diag).  At w = k = 100: 181 of 400 tile steps per warp.
template <int CT, bool UPPER>
__global__ void __launch_bounds__(GM_THREADS, 1)
k_gemm_rows(const double *__restrict__ A, u64 rows, u32 w, u32 lda, const double *__restrict__ S, u32 k, u32 lds, double *__restrict__ Out, u32 ldo) {
    extern __shared__ __align__(16) double gm_smem[];
    double(*sA)[128][GM_ASTRIDE] = reinterpret_cast<double(*)[128][GM_ASTRIDE]>(gm_smem);
    double(*sS)[GM_KC][GM_SSTRIDE] = reinterpret_cast<double(*)[GM_KC][GM_SSTRIDE]>(gm_smem + GM_STAGES * 128 * GM_ASTRIDE);
    const int t = threadIdx.x, lane = t & 31, rt = t >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    const u64 r0 = (u64)blockIdx.x * 128;
    double acc[CT][2];
#pragma unroll
    for (int c = 0; c < CT; c++) acc[c][0] = acc[c][1] = 0.0;
    const int arow = t >> 2, ak = (t & 3) * 4;
    const bool aligned16 = (lda & 1u) == 0 && ((size_t)A & 15u) == 0;
    auto issue = [&](u32 k0, int slot) {
        const u64 r = r0 + arow;
        const u32 c = arow;
        const bool rok = r < rows, cok = c < k;
        const double *arowp = A + (rok ? r : 0) * (size_t)lda;
        const double *scolp = S + (size_t)(cok ? c : 0) * lds;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const u32 kk = k0 + ak + 2 * h;
            const u32 ab = !rok || kk >= w ? 0u : min(16u, (w - kk) * 8u);
            if (aligned16) {
                cp_async16(&sA[slot][arow][ak + 2 * h], arowp + (ab ? kk : 0), ab);
            } else {
                cp_async8(&sA[slot][arow][ak + 2 * h], arowp + (ab ? kk : 0), ab ? 8u : 0u);
                cp_async8(&sA[slot][arow][ak + 2 * h + 1], arowp + (ab > 8u ? kk + 1 : 0), ab > 8u ? 8u : 0u);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const u32 kk = k0 + ak + i;
            const bool ok = cok && kk < w;
            cp_async8(&sS[slot][ak + i][arow], scolp + (ok ? kk : 0), ok ? 8u : 0u);
        }
    };
    const u32 nst = (w + GM_KC - 1) / GM_KC;  // <= 8
    for (int p = 0; p < GM_STAGES - 1; p++) {
        if ((u32)p < nst) issue((u32)p * GM_KC, p);
        cp_async_commit();
    }
#pragma unroll
    for (int st = 0; st < 128 / GM_KC; st++) {
        if ((u32)st < nst) {
            cp_async_wait<GM_STAGES - 2>();
            __syncthreads();
            if ((u32)st + GM_STAGES - 1 < nst) issue(((u32)st + GM_STAGES - 1) * GM_KC, (st + GM_STAGES - 1) % GM_STAGES);
            cp_async_commit();
            const int buf = st % GM_STAGES;
#pragma unroll
            for (int ks = 0; ks < GM_KC / 4; ks++) {
                const int kk = st * GM_KC + ks * 4;  // compile time
                const double a = sA[buf][rt * 8 + fr][ks * 4 + fk];
                const double *brow = sS[buf][ks * 4 + fk];
#pragma unroll
                for (int c = 0; c < CT; c++)
                    if (!UPPER || c * 8 + 8 > kk) dmma884(acc[c][0], acc[c][1], a, brow[c * 8 + fr]);
            }
        }
    }
    const u64 r = r0 + rt * 8 + fr;
    if (r < rows) {
#pragma unroll
        for (int c = 0; c < 16; c++) {  // also zero the pad columns up to ldo
            const u32 col = c * 8 + 2 * fk;
            if (col + 1 < ldo) {
                const double v0 = c < CT && col < k ? acc[c < CT ? c : 0][0] : 0.0, v1 = c < CT && col + 1 < k ? acc[c < CT ? c : 0][1] : 0.0;
                *reinterpret_cast<double2 *>(Out + r * (size_t)ldo + col) = make_double2(v0, v1);
            }
        }
    }
}

// The same product for a narrow S (k <= 16 output columns, w <= 128): the projection of the tall block on the k kept singular
// vectors (n x 100 by 100 x 10 at C3).  k_gemm_tall computes 128-column tiles whatever k is (1.37 ms there: 12x the flops); this one is
// a plain row-per-thread kernel bound by reading A once: S sits in shared memory (broadcast reads), A is staged 16 columns at a time
// through a padded tile so that the global reads are coalesced and the per-row reads conflict-free.
#define GS_ROWS 128
#define GS_KC 16
__global__ void __launch_bounds__(GS_ROWS)
k_gemm_skinny(const double *__restrict__ A, u64 rows, u32 w, u32 lda, const double *__restrict__ S, u32 k, u32 lds, double *__restrict__ Out, u32 ldo) {
    __shared__ double sS[128][16];                // S[kk][j], zero beyond (w, k)
    __shared__ double sA[GS_ROWS][GS_KC + 1];
    const u32 t = threadIdx.x;
    for (u32 i = t; i < 128 * 16; i += GS_ROWS) {
        const u32 kk = i >> 4, j = i & 15;
        sS[kk][j] = (kk < w && j < k) ? S[(size_t)j * lds + kk] : 0.0;
    }
    const u64 r0 = (u64)blockIdx.x * GS_ROWS;
    double acc[16];
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0.0;
    for (u32 k0 = 0; k0 < w; k0 += GS_KC) {
        __syncthreads();  // the previous chunk has been consumed (first pass: sS is complete)
        for (u32 i = t; i < GS_ROWS * GS_KC; i += GS_ROWS) {
            const u32 rr = i / GS_KC, cc = i - rr * GS_KC;
            const u64 r = r0 + rr;
            sA[rr][cc] = (r < rows && k0 + cc < w) ? A[r * (size_t)lda + k0 + cc] : 0.0;
        }
        __syncthreads();
        const u32 kn = min((u32)GS_KC, w - k0);
        for (u32 cc = 0; cc < kn; cc++) {
            const double a = sA[t][cc];
            const double2 *srow = reinterpret_cast<const double2 *>(sS[k0 + cc]);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double2 sv = srow[j];
                acc[2 * j] = fma(a, sv.x, acc[2 * j]);
                acc[2 * j + 1] = fma(a, sv.y, acc[2 * j + 1]);
            }
        }
    }
    const u64 r = r0 + t;
    if (r < rows) {
#pragma unroll
        for (int j = 0; j < 16; j += 2)
            if ((u32)j + 1 < ldo) *reinterpret_cast<double2 *>(Out + r * (size_t)ldo + j) = make_double2((u32)j < k ? acc[j] : 0.0, (u32)j + 1 < k ? acc[j + 1] : 0.0);
    }
}

int gemm_tall(sb_ctx *ctx, const double *A, u64 rows, u32 w, u32 lda, const double *S, u32 k, u32 lds, double *Out, u32 ldo, bool s_upper) {
    if (rows == 0 || k == 0) return SB_OK;
    if (ldo & 1) return sb_fail(SB_ERR_INVALID_ARG, "gemm_tall: odd leading dimension");
    if (k <= 16 && ldo <= 16 && w <= 128 && ctx->gemm_skinny) {
        k_gemm_skinny<<<(unsigned)cdiv(rows, GS_ROWS), GS_ROWS, 0, ctx->stream>>>(A, rows, w, lda, S, k, lds, Out, ldo);
        count_launch(ctx);
        SB_CUDA(cudaGetLastError());
        return SB_OK;
    }
    if (w <= 128 && k <= 128 && ldo <= 128) {  // one warp per row tile, live tile steps only
        const size_t smem1 = (size_t)(GM_STAGES * 128 * GM_ASTRIDE + GM_STAGES * GM_KC * GM_SSTRIDE) * sizeof(double);
        const unsigned grid = (unsigned)cdiv(rows, 128);
#define GR_LAUNCH(CTV, UP)                                                                                                     \
    do {                                                                                                                       \
        SB_CUDA(cudaFuncSetAttribute(k_gemm_rows<CTV, UP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));          \
        k_gemm_rows<CTV, UP><<<grid, GM_THREADS, smem1, ctx->stream>>>(A, rows, w, lda, S, k, lds, Out, ldo);                  \
    } while (0)
        const u32 ct = (k + 7) / 8;
        if (s_upper) {
            if (ct <= 3) GR_LAUNCH(3, true); else if (ct <= 8) GR_LAUNCH(8, true); else if (ct <= 13) GR_LAUNCH(13, true); else GR_LAUNCH(16, true);
        } else {
            if (ct <= 3) GR_LAUNCH(3, false); else if (ct <= 8) GR_LAUNCH(8, false); else if (ct <= 13) GR_LAUNCH(13, false); else GR_LAUNCH(16, false);
        }
#undef GR_LAUNCH
        count_launch(ctx);
        SB_CUDA(cudaGetLastError());
        return SB_OK;
    }
    const u32 cols = (std::max(k, ldo) + 127) / 128;  // also zero the pad columns
    const size_t smem = (size_t)(GM_STAGES * 128 * GM_ASTRIDE + GM_STAGES * GM_KC * GM_SSTRIDE) * sizeof(double);
    SB_CUDA(cudaFuncSetAttribute(k_gemm_tall, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gemm_tall<<<dim3(cdiv(rows, 128), cols), GM_THREADS, smem, ctx->stream>>>(A, rows, w, lda, S, k, lds, Out, ldo, s_upper ? 1 : 0);
    count_launch(ctx);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

// ---------------------------------------------------------------- small SPD: blocked Cholesky + inverse of the factor
// G: column-major w x w symmetric (upper triangle read and overwritten by R, G + shift I = R^T R); Rinv (column-major, upper, zero
// below) = R^-1.  shift = shift_coef * trace(G).  flag |= 1 on a non-positive pivot.  Blocks of 64: per block step one CTA
// factors the diagonal block in shared memory (and inverts it), a row of CTAs solves the block row, a triangle of CTAs updates
// the trailing matrix; the inverse is assembled block column by block column (one CTA each, back substitution over block rows).
// A first version did all of it in one CTA straight from global memory: 100 ms for w = 1000 (b = 200 Krylov blocks, k = 100).
#define CB 64
#define CB_LD 65
#define CH_THREADS 1024

__device__ __forceinline__ double *ch_elem(double *G, u32 w, u32 r, u32 c) { return G + (size_t)c * w + r; }

__global__ void k_chol_shift(double *__restrict__ G, u32 w, double shift_coef) {
    __shared__ double s_red[32];
    const u32 t = threadIdx.x;
    double tr = 0.0;
    for (u32 i = t; i < w; i += blockDim.x) tr += G[(size_t)i * w + i];
    for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
    if ((t & 31) == 0) s_red[t >> 5] = tr;
    __syncthreads();
    if (t < 32) {
        double v = t < (blockDim.x >> 5) ? s_red[t] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (t == 0) s_red[0] = v;
    }
    __syncthreads();
    const double shift = shift_coef * s_red[0];
    for (u32 i = t; i < w; i += blockDim.x) G[(size_t)i * w + i] += shift;
}

// factors the diagonal block at j0 (nb <= 64 rows) in shared memory, writes R_jj back and its inverse into Dinv (64 x 64, col-major)
// (shift_coef != 0 only when the block is the whole matrix: the trace of the block is then the trace of G; Rinv_small != NULL
// writes the inverse as the w x w result directly -- one launch does the whole factorisation of a matrix of at most 64 columns)
__global__ void __launch_bounds__(CH_THREADS, 1) k_chol_diag(double *__restrict__ G, u32 w, u32 j0, u32 nb, double *__restrict__ Dinv, int *__restrict__ flag,
                                                             double shift_coef, double *__restrict__ Rinv_small) {
    extern __shared__ __align__(16) double ch_dsm[];
    double(*A)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(ch_dsm);
    double(*X)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(ch_dsm + CB * CB_LD);
    __shared__ double s_piv;
    const u32 t = threadIdx.x;
    for (u32 i = t; i < CB * CB; i += CH_THREADS) {
        const u32 r = i % CB, c = i / CB;
        A[r][c] = (r <= c && c < nb) ? *ch_elem(G, w, j0 + r, j0 + c) : 0.0;
        X[r][c] = 0.0;
    }
    __syncthreads();
    if (shift_coef != 0.0) {
        if (t == 0) {
            double tr = 0.0;
            for (u32 i = 0; i < nb; i++) tr += A[i][i];
            s_piv = shift_coef * tr;
        }
        __syncthreads();
        if (t < nb) A[t][t] += s_piv;
        __syncthreads();
    }
    for (u32 k = 0; k < nb; k++) {
        if (t == 0) {
            double d = A[k][k];
            if (!(d > 0.0)) {
                atomicOr(flag, 1);
                d = 1.0;
            }
            s_piv = sqrt(d);
            A[k][k] = s_piv;
        }
        __syncthreads();
        const double inv = 1.0 / s_piv;
        if (t > k && t < nb) A[k][t] *= inv;
        __syncthreads();
        for (u32 e = t; e < nb * nb; e += CH_THREADS) {  // the live nb x nb corner only (nb = 20 for the Krylov blocks: 400 of 4,096 entries)
            const u32 r = e % nb, c = e / nb;
            if (r > k && r <= c) A[r][c] -= A[k][r] * A[k][c];
        }
        __syncthreads();
    }
    if (t < nb) {  // inverse of the upper factor, one column per thread
        const u32 c = t;
        X[c][c] = 1.0 / A[c][c];
        for (u32 ii = c; ii-- > 0;) {
            double sum = 0.0;
            for (u32 kk = ii + 1; kk <= c; kk++) sum += A[ii][kk] * X[kk][c];
            X[ii][c] = -sum / A[ii][ii];
        }
    }
    __syncthreads();
    for (u32 i = t; i < CB * CB; i += CH_THREADS) {
        const u32 r = i % CB, c = i / CB;
        if (r <= c && c < nb) *ch_elem(G, w, j0 + r, j0 + c) = A[r][c];
        if (Dinv) Dinv[(size_t)c * CB + r] = X[r][c];
        if (Rinv_small && r < nb && c < nb) Rinv_small[(size_t)c * w + r] = X[r][c];
    }
}

// block row j: R[j, c] = R_jj^-T G[j, c] for the block columns c right of j (one CTA each)
__global__ void __launch_bounds__(CH_THREADS, 1) k_chol_panel(double *__restrict__ G, u32 w, u32 j0, u32 nb, const double *__restrict__ Dinv) {
    extern __shared__ __align__(16) double ch_dsm[];
    double(*D)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(ch_dsm);
    double(*B)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(ch_dsm + CB * CB_LD);
    const u32 t = threadIdx.x;
    const u32 c0 = j0 + nb + blockIdx.x * CB;
    const u32 nc = min((u32)CB, w - c0);
    for (u32 i = t; i < CB * CB; i += CH_THREADS) {
        const u32 r = i % CB, c = i / CB;
        D[r][c] = Dinv[(size_t)c * CB + r];
        B[r][c] = (r < nb && c < nc) ? *ch_elem(G, w, j0 + r, c0 + c) : 0.0;
    }
    __syncthreads();
    for (u32 e = t; e < CB * CB; e += CH_THREADS) {
        const u32 r = e % CB, c = e / CB;
        if (r < nb && c < nc) {
            double sum = 0.0;
            for (u32 k = 0; k <= r; k++) sum += D[k][r] * B[k][c];  // (D^T B)[r][c], D upper
            *ch_elem(G, w, j0 + r, c0 + c) = sum;
        }
    }
}

// trailing update: G[r, c] -= R[j, r]^T R[j, c] for block pairs j < r <= c (blockIdx.x enumerates the pairs)
__global__ void __launch_bounds__(CH_THREADS, 1) k_chol_update(double *__restrict__ G, u32 w, u32 j0, u32 nb, u32 nrem) {
    extern __shared__ __align__(16) double ch_dsm[];
    double(*Rr)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(ch_dsm);
    double(*Rc)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(ch_dsm + CB * CB_LD);
    u32 br = 0, bc = 0;
    {
        u32 p = blockIdx.x;
        for (br = 0; br < nrem; br++) {
            const u32 cnt = nrem - br;
            if (p < cnt) {
                bc = br + p;
                break;
            }
            p -= cnt;
        }
    }
    const u32 t = threadIdx.x;
    const u32 r0 = j0 + nb + br * CB, c0 = j0 + nb + bc * CB;
    const u32 nr = min((u32)CB, w - r0), nc = min((u32)CB, w - c0);
    for (u32 i = t; i < CB * CB; i += CH_THREADS) {
        const u32 k = i % CB, c = i / CB;
        Rr[k][c] = (k < nb && c < nr) ? *ch_elem(G, w, j0 + k, r0 + c) : 0.0;
        Rc[k][c] = (k < nb && c < nc) ? *ch_elem(G, w, j0 + k, c0 + c) : 0.0;
    }
    __syncthreads();
    for (u32 e = t; e < CB * CB; e += CH_THREADS) {
        const u32 r = e % CB, c = e / CB;
        if (r < nr && c < nc && r0 + r <= c0 + c) {
            double sum = 0.0;
            for (u32 k = 0; k < nb; k++) sum += Rr[k][r] * Rc[k][c];
            *ch_elem(G, w, r0 + r, c0 + c) -= sum;
        }
    }
}

// Rinv block column bj (one CTA): X_jj = D_j; X_ij = -D_i (sum_{i < k <= j} R_ik X_kj), block rows i descending.  Dall: the
// inverted diagonal blocks, 64 x 64 each.
__global__ void __launch_bounds__(CH_THREADS, 1) k_tri_inv_blocks(const double *__restrict__ G, u32 w, const double *__restrict__ Dall, double *__restrict__ Rinv) {
    extern __shared__ __align__(16) double ch_dsm[];
    double(*Ablk)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(ch_dsm);
    double(*Xblk)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(ch_dsm + CB * CB_LD);
    double(*Acc)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(ch_dsm + 2 * CB * CB_LD);
    const u32 t = threadIdx.x;
    const u32 bj = blockIdx.x, c0 = bj * CB, nc = min((u32)CB, w - c0);
    // zero the block column (also below the diagonal), then the diagonal block
    for (u32 i = t; i < w * nc; i += CH_THREADS) Rinv[(size_t)(c0 + i / w) * w + i % w] = 0.0;
    __syncthreads();
    for (u32 i = t; i < CB * CB; i += CH_THREADS) {
        const u32 r = i % CB, c = i / CB;
        if (r < nc && c < nc) Rinv[(size_t)(c0 + c) * w + c0 + r] = Dall[(size_t)bj * CB * CB + (size_t)c * CB + r];
    }
    __syncthreads();
    for (u32 bi = bj; bi-- > 0;) {
        const u32 r0 = bi * CB;  // full block of 64 rows (only the last block can be short, and bi < bj)
        for (u32 i = t; i < CB * CB; i += CH_THREADS) Acc[i % CB][i / CB] = 0.0;
        __syncthreads();
        for (u32 bk = bi + 1; bk <= bj; bk++) {
            const u32 k0 = bk * CB, nk = min((u32)CB, w - k0);
            for (u32 i = t; i < CB * CB; i += CH_THREADS) {
                const u32 r = i % CB, c = i / CB;
                Ablk[r][c] = c < nk ? G[(size_t)(k0 + c) * w + r0 + r] : 0.0;                      // R[bi, bk] (row r, col c)
                Xblk[r][c] = (r < nk && c < nc) ? Rinv[(size_t)(c0 + c) * w + k0 + r] : 0.0;       // X[bk, bj]
            }
            __syncthreads();
            for (u32 e = t; e < CB * CB; e += CH_THREADS) {
                const u32 r = e % CB, c = e / CB;
                double sum = 0.0;
                for (u32 k = 0; k < CB; k++) sum += Ablk[r][k] * Xblk[k][c];
                Acc[r][c] += sum;
            }
            __syncthreads();
        }
        // X[bi, bj] = -D_bi . Acc
        for (u32 i = t; i < CB * CB; i += CH_THREADS) {
            const u32 r = i % CB, c = i / CB;
            Ablk[r][c] = Dall[(size_t)bi * CB * CB + (size_t)c * CB + r];
        }
        __syncthreads();
        for (u32 e = t; e < CB * CB; e += CH_THREADS) {
            const u32 r = e % CB, c = e / CB;
            if (c < nc) {
                double sum = 0.0;
                for (u32 k = r; k < CB; k++) sum += Ablk[r][k] * Acc[k][c];  // D upper
                Rinv[(size_t)(c0 + c) * w + r0 + r] = -sum;
            }
        }
        __syncthreads();
    }
}

int chol_inv(sb_ctx *ctx, double *G, u32 w, double shift_coef, double *Rinv, int *flag) {
    if (w == 0) return SB_OK;
    const u32 nblk = (w + CB - 1) / CB;
    DevBuf<double> Dall;
    SB_TRY(Dall.alloc((size_t)nblk * CB * CB));
    const size_t sm2 = (size_t)2 * CB * CB_LD * sizeof(double), sm3 = (size_t)3 * CB * CB_LD * sizeof(double);
    SB_CUDA(cudaFuncSetAttribute(k_chol_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));  // per device: cheap, so every call
    SB_CUDA(cudaFuncSetAttribute(k_chol_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
    SB_CUDA(cudaFuncSetAttribute(k_chol_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
    SB_CUDA(cudaFuncSetAttribute(k_tri_inv_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
    if (nblk == 1) {  // the whole matrix is one block: shift, factor and invert in one launch
        k_chol_diag<<<1, CH_THREADS, sm2, ctx->stream>>>(G, w, 0, w, nullptr, flag, shift_coef, Rinv);
        count_launch(ctx);
        SB_CUDA(cudaGetLastError());
        return SB_OK;
    }
    if (shift_coef != 0.0) {
        k_chol_shift<<<1, 1024, 0, ctx->stream>>>(G, w, shift_coef);
        count_launch(ctx);
    }
    for (u32 b = 0; b < nblk; b++) {
        const u32 j0 = b * CB, nb = std::min<u32>(CB, w - j0);
        k_chol_diag<<<1, CH_THREADS, sm2, ctx->stream>>>(G, w, j0, nb, Dall.p + (size_t)b * CB * CB, flag, 0.0, nullptr);
        count_launch(ctx);
        const u32 nrem = nblk - b - 1;
        if (nrem) {
            k_chol_panel<<<nrem, CH_THREADS, sm2, ctx->stream>>>(G, w, j0, nb, Dall.p + (size_t)b * CB * CB);
            k_chol_update<<<nrem * (nrem + 1) / 2, CH_THREADS, sm2, ctx->stream>>>(G, w, j0, nb, nrem);
            count_launch(ctx); count_launch(ctx);
        }
    }
    k_tri_inv_blocks<<<nblk, CH_THREADS, sm3, ctx->stream>>>(G, w, Dall.p, Rinv);
    count_launch(ctx);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

// C (w x w, column-major) = A . B for upper-triangular A, B (small; one thread per output element)
__global__ void k_tri_mul(const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C, u32 w) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (u64)w * w) return;
    const u32 j = (u32)(idx / w), i = (u32)(idx - (u64)j * w);
    double s = 0.0;
    if (i <= j)
        for (u32 kk = i; kk <= j; kk++) s += A[(size_t)kk * w + i] * B[(size_t)j * w + kk];
    C[idx] = s;
}

// ---------------------------------------------------------------- thin QR by shifted CholeskyQR3
// A (row-major rows x w, ld) is replaced by Q; Rinv_out (device, column-major w x w, may be NULL) receives R^-1 with A_in = Q R.
// tmp: a scratch block of the same shape as A.  Needs rows >= w.  *flag is raised on breakdown (the caller falls back).
// rows_global != 0: A is row-sharded over the ranks (a cell-side block) -- the w x w Gram matrices are all-reduced, everything
// else is local, and every rank ends with its rows of the same Q and the same R^-1 (distributed CholeskyQR).
int qr_chol(sb_ctx *ctx, double *A, double *tmp, u64 rows, u32 w, u32 ld, double *Rinv_out, int *flag, u64 rows_global) {
    ProfScope ps(ctx, PH_DENSE);
    DevBuf<double> G, Ri, Racc, Rtmp;
    SB_TRY(G.alloc((size_t)w * w));
    SB_TRY(Ri.alloc((size_t)w * w));
    if (Rinv_out) SB_TRY(Rtmp.alloc((size_t)w * w));
    // shift of the first round: 11 (m n + n (n + 1)) u ||A||_2^2, with the trace as the bound on ||A||_2^2
    const double u = 1.1102230246251565e-16;
    const double coef0 = 11.0 * ((double)(rows_global ? rows_global : rows) * w + (double)w * (w + 1)) * u;
    double *src = A, *dst = tmp;
    for (int pass = 0; pass < 3; pass++) {
        SB_TRY(syrk_tall(ctx, src, rows, w, ld, G.p));
        if (rows_global) SB_TRY(comm_allreduce_f64(ctx, G.p, (size_t)w * w));
        SB_TRY(chol_inv(ctx, G.p, w, pass == 0 ? coef0 : 0.0, Ri.p, flag));
        SB_TRY(gemm_tall(ctx, src, rows, w, ld, Ri.p, w, w, dst, ld, true));  // R^-1 is upper triangular
        if (Rinv_out) {
            if (pass == 0) {
                SB_CUDA(cudaMemcpyAsync(Rinv_out, Ri.p, (size_t)w * w * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            } else {
                k_tri_mul<<<cdiv((u64)w * w, 256), 256, 0, ctx->stream>>>(Rinv_out, Ri.p, Rtmp.p, w);
                count_launch(ctx);
                SB_CUDA(cudaMemcpyAsync(Rinv_out, Rtmp.p, (size_t)w * w * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            }
        }
        std::swap(src, dst);
    }
    // three swaps: the result is in `tmp`
    SB_CUDA(cudaMemcpyAsync(A, tmp, rows * (size_t)ld * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return SB_OK;
}

// ---------------------------------------------------------------- top-k eigenpairs -> projection matrices (device resident)
// W: eigenvectors, column-major wq x wq, eigenvalues ascending.  Wsel[i] = eigenvector of the i-th largest eigenvalue,
// Wsc[i] = Wsel[i] / sigma_i, S[i] = sigma_i = sqrt(max(lambda, 0)).
__global__ void k_topk(const double *__restrict__ W, const double *__restrict__ ev, u32 wq, u32 k, double *__restrict__ Wsel, double *__restrict__ Wsc,
                       double *__restrict__ S) {
    const u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= wq * k) return;
    const u32 i = idx / wq, r = idx - i * wq;
    const u32 src = wq - 1 - i;
    const double lam = ev[src];
    const double sig = lam > 0.0 ? sqrt(lam) : 0.0;
    const double x = W[(size_t)src * wq + r];
    Wsel[idx] = x;
    Wsc[idx] = sig > 0.0 ? x * (1.0 / sig) : 0.0;
    if (r == 0) S[i] = sig;
}

int topk_select(sb_ctx *ctx, const double *W, const double *ev, u32 wq, u32 k, double *Wsel, double *Wsc, double *S) {
    if (wq == 0 || k == 0) return SB_OK;
    k_topk<<<cdiv((u64)wq * k, 256), 256, 0, ctx->stream>>>(W, ev, wq, k, Wsel, Wsc, S);
    count_launch(ctx);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}
