// umma_i8_rate.cu -- cycles per tcgen05.mma.kind::i8 (M = 128, K = 32, cta_group::1, SWIZZLE_NONE operands in shared memory) as a
// function of N, of the operand major-ness and of the accumulator pattern.  One CTA per SM, one issuing thread; the other
// warps idle.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_rate umma_i8_rate.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../scan_rs_b200/csrc/tc05.cuh"
using namespace tc05;

__global__ void __launch_bounds__(128, 1) k_rate(uint32_t N, int mn_major, int n_acc, int iters, int per_commit, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long bar;
    __shared__ uint32_t tslot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc_512(smem_u32(&tslot));
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tslot;
    long long t0 = 0, t1 = 0;
    if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = instr_desc_i8(128, N, mn_major, mn_major);
            const uint32_t a = smem_u32(smem), b = a + 16384;
            // K-major: lbo = next 16 K bytes (M/8 core matrices away), sbo = 128; MN-major: lbo = 8 K rows, sbo = 16 MN elements
            const uint64_t da = mn_major ? smem_desc(a, 128, 512) : smem_desc(a, 2048, 128);
            const uint64_t db = mn_major ? smem_desc(b, N / 16 * 128, 128) : smem_desc(b, N / 8 * 128, 128);
            uint32_t phase = 0;
            t0 = clock64();
            for (int i = 0; i < iters; i += per_commit) {
                for (int j = 0; j < per_commit; j++) mma_i8(tmem + ((i + j) % n_acc) * 160, da, db, idesc, 1u);
                commit(smem_u32(&bar));
                if (i + per_commit >= iters) { mbar_wait(smem_u32(&bar), phase); }
                else if ((i / per_commit) % 8 == 7) { mbar_wait(smem_u32(&bar), phase); }  // keep at most 8 commits in flight
                if (i + per_commit >= iters || (i / per_commit) % 8 == 7) phase ^= 0;     // phases tracked below
            }
            t1 = clock64();
        }
        __syncwarp();
    }
    __syncthreads();
    if (tid == 32 && blockIdx.x == 0) out[0] = t1 - t0;
    if (warp == 0) tmem_dealloc_512(tmem);
}

// simpler and exact: issue all MMAs, one commit at the end, wait
__global__ void __launch_bounds__(128, 1) k_rate2(uint32_t N, int mn_major, int n_acc, int iters, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long bar;
    __shared__ uint32_t tslot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc_512(smem_u32(&tslot));
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tslot;
    long long t0 = 0, t1 = 0;
    if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = instr_desc_i8(128, N, mn_major, mn_major);
            const uint32_t a = smem_u32(smem), b = a + 16384;
            const uint64_t da = mn_major ? smem_desc(a, 128, 512) : smem_desc(a, 2048, 128);
            const uint64_t db = mn_major ? smem_desc(b, N / 16 * 128, 128) : smem_desc(b, N / 8 * 128, 128);
            t0 = clock64();
            for (int i = 0; i < iters; i++) mma_i8(tmem + (i % n_acc) * 160, da, db, idesc, 1u);
            commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            t1 = clock64();
        }
        __syncwarp();
    }
    __syncthreads();
    if (tid == 32 && blockIdx.x == 0) out[0] = t1 - t0;
    if (warp == 0) tmem_dealloc_512(tmem);
}

int main() {
    long long *d, h;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k_rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 4096;
    for (int mn = 0; mn < 2; mn++)
        for (uint32_t N : {64u, 128u, 144u, 160u, 192u, 256u})
            for (int n_acc : {1, 3}) {
                if ((uint32_t)n_acc * 160 > 512 || (n_acc > 1 && N > 160)) continue;
                for (int grid : {1, 148}) {
                    k_rate2<<<grid, 128, 64 * 1024>>>(N, mn, n_acc, iters, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                    printf("%s N=%3u accumulators=%d grid=%3d : %7.1f cycles per MMA (%s)\n", mn ? "MN-major" : "K-major ", N, n_acc, grid, (double)h / iters,
                           cudaGetErrorString(e));
                }
            }
    return 0;
}
