// panel_i8.cu -- EXPERIMENTAL, default off (sb_set_option("panel_i8", 1)), written at the end of round 1 WITHOUT a GPU
// to run it on: it compiles for sm_100a and its error model is checked on the CPU (scripts/proto_int8_panel.py), but it
// has not executed on hardware yet.  Nothing calls it unless the option is set; the product path is the FP64 mma.sync
// kernel k_dense_t (dense_panel.cu).  DESIGN.md 7(1) has the plan and the sizing.
//
// The T side of the dense hot-gene panel on the INTEGER tensor cores:
//     T[c,:] += sum_j L_c(D[c,j]) * rs_j * Y[hot_j,:]  =  sum_{k=1..3} L_c(k) * ( [D == k] . Ys )[c,:]
// [D == k] is a 0/1 matrix, so Ys (2,048 x 20 f64, row-scaled) is cut per column into 8 balanced base-128 digits of a
// 54-bit fixed-point number and every [D == k] . digit plane is an exact int8 x int8 -> int32 product
// (tcgen05.mma.kind::i8).  The 8 planes ride in one MMA as N = 8 x 20 = 160 columns; the three count levels accumulate
// into three TMEM accumulators (3 x 160 = 480 of 512 columns).  The panel must have been built with counts 1..3 only
// (sb_set_option("dense_max_count", 3) before the upload): larger counts stay on the sparse side.
//
// One CTA owns 1,024 of the panel's genes and keeps their digit planes resident in shared memory as the B operand
// (K-major, no swizzle: core matrices of 8 rows x 16 bytes; 160 KB); it walks the cell tiles (128 cells), turning D tiles
// of 128 cells x 64 genes into the three 0/1 A tiles in the same core-matrix layout (2 stages x 24 KB).
//   all 8 warps : producers (D -> A tiles) ; thread 0 issues the MMAs ; warps 0-3 : epilogue (tcgen05.ld, recombination)
#include "common.cuh"
#include "map.cuh"

#define PI_CELLS 128u
#define PI_GENES 1024u
#define PI_KB 64u
#define PI_COLS 20u
#define PI_DIGITS 8u
#define PI_N (PI_DIGITS * PI_COLS)  // 160
#define PI_LEVELS 3u
#define PI_THREADS 256
#define PI_B_BYTES (PI_GENES * PI_N)                 // 163,840
#define PI_A_LEVEL_BYTES (PI_CELLS * PI_KB)           // 8,192
#define PI_A_STAGE_BYTES (PI_LEVELS * PI_A_LEVEL_BYTES)
#define PI_FIXED_BITS 54

// ---------------------------------------------------------------- digit planes of Ys
__global__ void k_pi_colmax(const double *__restrict__ Y, u32 ldy, const u32 *__restrict__ hot_idx, const double *__restrict__ rs, u32 gd, u32 col0,
                            u32 w, unsigned long long *__restrict__ colmax_bits) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= gd * PI_COLS) return;
    const u32 j = i / PI_COLS, c = i - j * PI_COLS;
    const u32 g = hot_idx[j];
    double v = (col0 + c < w) ? Y[(size_t)g * ldy + col0 + c] : 0.0;
    if (rs) v *= rs[g];
    atomicMax(&colmax_bits[c], (unsigned long long)__double_as_longlong(fabs(v)));  // non-negative doubles order like their bits
}

// Bd[half][k / 16][n / 8][n % 8][k % 16] = digit (n / 20) of column (n % 20) of gene k of that half; scale2[c] = 2^(e_c - 54)
__global__ void k_pi_digits(const double *__restrict__ Y, u32 ldy, const u32 *__restrict__ hot_idx, const double *__restrict__ rs, u32 gd, u32 col0,
                            u32 w, const unsigned long long *__restrict__ colmax_bits, signed char *__restrict__ Bd, double *__restrict__ scale2) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= gd * PI_COLS) return;
    const u32 j = i / PI_COLS, c = i - j * PI_COLS;
    const double mx = __longlong_as_double((long long)colmax_bits[c]);
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);  // mx = f * 2^e, f in [0.5, 1): |x| <= mx < 2^e
    if (j == 0) scale2[c] = ldexp(1.0, e - PI_FIXED_BITS);
    const u32 g = hot_idx[j];
    double v = (col0 + c < w) ? Y[(size_t)g * ldy + col0 + c] : 0.0;
    if (rs) v *= rs[g];
    long long q = __double2ll_rn(ldexp(v, PI_FIXED_BITS - e));  // |q| <= 2^54, inside the 8-digit balanced range (0.99 * 2^55)
    const u32 half = j / PI_GENES, k = j - half * PI_GENES;
    signed char *base = Bd + (size_t)half * PI_B_BYTES + (size_t)(k >> 4) * (PI_N / 8 * 128) + (k & 15);
#pragma unroll
    for (u32 s = 0; s < PI_DIGITS; s++) {
        const long long d = ((q + 64) & 127) - 64;  // balanced digit in [-64, 63]
        q = (q - d) >> 7;
        const u32 n = s * PI_COLS + c;
        base[(size_t)(n >> 3) * 128 + (n & 7) * 16] = (signed char)d;
    }
}

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): core matrices of 8 rows x 16 bytes;
// lbo = bytes between core matrices adjacent in K, sbo = bytes between core matrices adjacent in M / N
__device__ __forceinline__ unsigned long long pi_smem_desc(u32 saddr, u32 lbo, u32 sbo) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3FFFu);
    d |= (unsigned long long)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;  // descriptor version 1 (Blackwell); base offset 0, layout type 0 = no swizzle
    return d;
}

// cute::UMMA::InstrDescriptor for kind::i8: D = S32, A = B = signed 8-bit, both K-major, N = 160, M = 128
__device__ __forceinline__ u32 pi_instr_desc() {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((PI_N >> 3) << 17) | ((PI_CELLS >> 4) << 24);
}

__device__ __forceinline__ void pi_mma(u32 tmem_c, unsigned long long da, unsigned long long db, u32 idesc, u32 accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_c),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void pi_commit(u32 mbar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_saddr) : "memory");
}

__device__ __forceinline__ void pi_mbar_wait(u32 mbar_saddr, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "PI_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PI_DONE_%=;\n\t"
        "bra PI_WAIT_%=;\n\t"
        "PI_DONE_%=:\n\t}\n" ::"r"(mbar_saddr),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------- the kernel
// grid (ctas per half, 2 halves of the panel's 2,048 genes); dynamic shared memory:
//   [ B planes 163,840 | A stages 2 x 24,576 | L table 128 x 3 f64 | scale2 20 f64 | mbarriers 3 x 8 | tmem slot ]
__global__ void __launch_bounds__(PI_THREADS, 1)
k_panel_t_i8(const unsigned char *__restrict__ D, u32 gd, u64 n, const double *__restrict__ cs, int log_base, const signed char *__restrict__ Bd,
             const double *__restrict__ scale2_g, u32 col0, u32 w, double *__restrict__ out, u32 ldo) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sB = smem;
    unsigned char *sA = sB + PI_B_BYTES;
    double *sL = reinterpret_cast<double *>(sA + 2 * PI_A_STAGE_BYTES);
    double *sScale = sL + PI_CELLS * PI_LEVELS;
    unsigned long long *sBar = reinterpret_cast<unsigned long long *>(sScale + PI_COLS);  // [0,1] stage free, [2] accumulators full
    u32 *sTmem = reinterpret_cast<u32 *>(sBar + 3);

    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u32 half = blockIdx.y;
    const u32 bar0 = smem_u32(sBar);

    if (tid == 0) {
        for (u32 i = 0; i < 3; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // 512 TMEM columns: 3 accumulators of 160
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(sTmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // resident B operand: this half's digit planes
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(Bd + (size_t)half * PI_B_BYTES);
        uint4 *dst = reinterpret_cast<uint4 *>(sB);
        for (u32 i = tid; i < PI_B_BYTES / 16; i += PI_THREADS) dst[i] = src[i];
    }
    if (tid < PI_COLS) sScale[tid] = scale2_g[tid];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core's reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem = *sTmem;
    const u32 idesc = pi_instr_desc();
    const u32 sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);

    u32 uses[2] = {0u, 0u};  // how often each A stage has been handed to the tensor core
    u32 tiles_done = 0;
    const u64 ntiles = (n + PI_CELLS - 1) / PI_CELLS;
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tiles_done++) {
        const u64 c0 = tile * PI_CELLS;
        if (tid < PI_CELLS) {
            const u64 c = c0 + tid;
            const double s = c < n ? cs[c] : 0.0;
#pragma unroll
            for (u32 k = 0; k < PI_LEVELS; k++) sL[tid * PI_LEVELS + k] = c < n ? map_log_part(log_base, s, k + 1, sb_log_table) : 0.0;
        }
        for (u32 kb = 0; kb < PI_GENES / PI_KB; kb++) {
            const u32 st = kb & 1u;
            // the tensor core is done reading this stage's previous contents?
            if (uses[st] > 0) pi_mbar_wait(bar0 + 8 * st, (uses[st] - 1) & 1u);
            uses[st]++;
            unsigned char *stage = sA + st * PI_A_STAGE_BYTES;
            for (u32 id = tid; id < PI_CELLS * (PI_KB / 16); id += PI_THREADS) {
                const u32 m = id & (PI_CELLS - 1), kc = id / PI_CELLS;
                uint4 d4 = make_uint4(0u, 0u, 0u, 0u);
                if (c0 + m < n) d4 = *reinterpret_cast<const uint4 *>(D + (c0 + m) * (u64)gd + half * PI_GENES + kb * PI_KB + kc * 16);
                const u32 off = kc * (PI_CELLS / 8 * 128) + (m >> 3) * 128 + (m & 7) * 16;
#pragma unroll
                for (u32 lv = 0; lv < PI_LEVELS; lv++) {
                    const u32 pat = (lv + 1) * 0x01010101u;
                    uint4 a;
                    a.x = __vcmpeq4(d4.x, pat) & 0x01010101u;
                    a.y = __vcmpeq4(d4.y, pat) & 0x01010101u;
                    a.z = __vcmpeq4(d4.z, pat) & 0x01010101u;
                    a.w = __vcmpeq4(d4.w, pat) & 0x01010101u;
                    *reinterpret_cast<uint4 *>(stage + lv * PI_A_LEVEL_BYTES + off) = a;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (u32 lv = 0; lv < PI_LEVELS; lv++)
#pragma unroll
                    for (u32 kk = 0; kk < PI_KB / 32; kk++) {  // K = 32 bytes per MMA = two 16-byte core-matrix columns
                        const unsigned long long da =
                            pi_smem_desc(sA_addr + st * PI_A_STAGE_BYTES + lv * PI_A_LEVEL_BYTES + kk * 2 * (PI_CELLS / 8 * 128), PI_CELLS / 8 * 128, 128);
                        const unsigned long long db = pi_smem_desc(sB_addr + (kb * (PI_KB / 16) + kk * 2) * (PI_N / 8 * 128), PI_N / 8 * 128, 128);
                        pi_mma(tmem + lv * PI_N, da, db, idesc, (kb | kk) ? 1u : 0u);
                    }
                pi_commit(bar0 + 8 * st);                                  // stage reusable once these MMAs have read it
                if (kb + 1 == PI_GENES / PI_KB) pi_commit(bar0 + 16);      // accumulators complete
            }
        }
        // epilogue: one thread per cell row (warps 0-3 own TMEM lanes 32 w .. 32 w + 31)
        pi_mbar_wait(bar0 + 16, tiles_done & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp < 4) {
            const u32 row = warp * 32 + lane;
            const u64 cell = c0 + row;
            double res[PI_COLS];
#pragma unroll
            for (u32 j = 0; j < PI_COLS; j++) res[j] = 0.0;
#pragma unroll 1
            for (u32 lv = 0; lv < PI_LEVELS; lv++) {
                double tmp[PI_COLS];
#pragma unroll
                for (u32 j = 0; j < PI_COLS; j++) tmp[j] = 0.0;
#pragma unroll
                for (u32 cb = 0; cb < PI_N; cb += 16) {
                    u32 r[16];
                    const u32 taddr = tmem + ((warp * 32u) << 16) + lv * PI_N + cb;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (u32 i = 0; i < 16; i++) {
                        const u32 nn = cb + i, s = nn / PI_COLS, j = nn - s * PI_COLS;  // compile-time after unrolling
                        tmp[j] = fma((double)(int)r[i], (double)(1ull << (7 * s)), tmp[j]);  // exact product, one rounding per add
                    }
                }
                const double lk = sL[row * PI_LEVELS + lv];
#pragma unroll
                for (u32 j = 0; j < PI_COLS; j++) res[j] = fma(lk, tmp[j], res[j]);
            }
            if (cell < n) {
                double *o = out + cell * (size_t)ldo + col0;
#pragma unroll
                for (u32 j = 0; j < PI_COLS; j++)
                    if (col0 + j < w) atomicAdd(o + j, res[j] * sScale[j]);  // the other half of the genes adds into the same row
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // accumulators and the L table are free for the next tile
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------- launcher
// usable: log-chain map, panel of exactly 2,048 genes built with counts 1..3, one column pass of <= 20 columns
bool dense_t_i8_usable(const sb_nmat *a, u32 w) {
    const sb_mat *mt = a->mat;
    return mt->ctx->panel_i8 && a->kind == 1 && mt->gd == 2 * PI_GENES && mt->dense_max_count <= PI_LEVELS && w <= PI_COLS;
}

int dense_t_i8(sb_nmat *a, const double *Y, u32 ldy, u32 w, double *out, u32 ldo) {
    sb_mat *mt = a->mat;
    sb_ctx *ctx = mt->ctx;
    if (mt->n == 0 || w == 0) return SB_OK;
    DevBuf<unsigned long long> colmax;
    DevBuf<signed char> Bd;
    DevBuf<double> scale2;
    SB_TRY(colmax.alloc(PI_COLS));
    SB_TRY(Bd.alloc((size_t)2 * PI_B_BYTES));
    SB_TRY(scale2.alloc(PI_COLS));
    SB_CUDA(cudaMemsetAsync(colmax.p, 0, PI_COLS * sizeof(unsigned long long), ctx->stream));
    const double *rs = a->has_row_scale ? a->row_scale.p : nullptr;
    const u32 items = mt->gd * PI_COLS;
    k_pi_colmax<<<cdiv(items, 256), 256, 0, ctx->stream>>>(Y, ldy, mt->hot_idx.p, rs, mt->gd, 0, w, colmax.p);
    k_pi_digits<<<cdiv(items, 256), 256, 0, ctx->stream>>>(Y, ldy, mt->hot_idx.p, rs, mt->gd, 0, w, colmax.p, Bd.p, scale2.p);
    const size_t smem = (size_t)PI_B_BYTES + 2 * PI_A_STAGE_BYTES + (PI_CELLS * PI_LEVELS + PI_COLS) * sizeof(double) + 3 * 8 + 16;
    cudaError_t e = cudaFuncSetAttribute(k_panel_t_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return sb_fail(SB_ERR_CUDA, "dense_t_i8: %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    const u64 ntiles = (mt->n + PI_CELLS - 1) / PI_CELLS;
    const u32 per_half = (u32)std::max<u64>(1, std::min<u64>(ntiles, (u64)ctx->sm_count / 2));
    k_panel_t_i8<<<dim3(per_half, 2), PI_THREADS, smem, ctx->stream>>>(mt->D.p, mt->gd, mt->n, a->col_scale.p, a->log_base, Bd.p, scale2.p, 0, w, out, ldo);
    count_launch(ctx); count_launch(ctx); count_launch(ctx);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}
