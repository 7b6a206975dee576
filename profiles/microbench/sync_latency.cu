// sync_latency.cu -- latencies that bound a shared-memory stage ring feeding tcgen05.mma (B200):
//   (a) issue n MMAs (M = 128, N = 144, K = 32, kind::i8) + tcgen05.commit -> the issuing thread sees the mbarrier phase complete
//   (b) mbarrier ping-pong between two warps: try_wait (may suspend) vs test_wait spinning
//   (c) the same ping-pong with fence.proxy.async.shared::cta before every arrive (what a generic-proxy producer needs)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../scan_rs_b200/csrc/tc05.cuh"
using namespace tc05;

__device__ __forceinline__ void mbar_spin(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128, 1) k_lat(long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long bar[4];
    __shared__ uint32_t tslot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (uint32_t i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (tid == 0) { for (int i = 0; i < 4; i++) mbar_init(smem_u32(&bar[i]), 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc_512(smem_u32(&tslot));
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tslot;
    const uint32_t b0 = smem_u32(&bar[0]), b1 = smem_u32(&bar[1]), b2 = smem_u32(&bar[2]);
    // (a) MMA + commit latency
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = instr_desc_i8(128, 144, false, false);
        const uint32_t a = smem_u32(smem), b = a + 16384;
        const uint64_t da = smem_desc(a, 2048, 128), db = smem_desc(b, 144 / 8 * 128, 128);
        uint32_t phase = 0;
        int slot = 0;
        for (int n : {1, 2, 4, 8, 16}) {
            long long best = 1ll << 60;
            for (int rep = 0; rep < 20; rep++) {
                const long long t0 = clock64();
                for (int i = 0; i < n; i++) mma_i8(tmem, da, db, idesc, 1u);
                commit(b0);
                mbar_spin(b0, phase);
                const long long t1 = clock64();
                phase ^= 1;
                best = min(best, t1 - t0);
            }
            out[slot++] = best;
        }
    }
    __syncthreads();
    // (b), (c) ping-pong between warp 2 and warp 3 (lane 0 each)
    for (int mode = 0; mode < 4; mode++) {
        const bool spin = mode & 1, fence = mode & 2;
        const int iters = 200;
        __syncthreads();
        if (warp == 2 && lane == 0) {
            const long long t0 = clock64();
            for (int i = 0; i < iters; i++) {
                if (fence) fence_async_smem();
                mbar_arrive(b1);
                if (spin) mbar_spin(b2, i & 1); else mbar_wait(b2, i & 1);
            }
            out[8 + mode] = (clock64() - t0) / iters;  // round trip
        }
        if (warp == 3 && lane == 0) {
            for (int i = 0; i < iters; i++) {
                if (spin) mbar_spin(b1, i & 1); else mbar_wait(b1, i & 1);
                if (fence) fence_async_smem();
                mbar_arrive(b2);
            }
        }
        __syncthreads();
        // both barriers completed `iters` phases: an even count keeps parity aligned for the next mode
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc_512(tmem);
}

int main() {
    long long *d, h[16] = {0};
    cudaMalloc(&d, sizeof(h));
    cudaMemset(d, 0, sizeof(h));
    cudaFuncSetAttribute(k_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k_lat<<<1, 128, 64 * 1024>>>(d);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int ns[5] = {1, 2, 4, 8, 16};
    for (int i = 0; i < 5; i++) printf("issue %2d MMA(s) + commit -> phase observed: %lld cycles\n", ns[i], h[i]);
    const char *names[4] = {"try_wait", "test_wait spin", "try_wait + fence.proxy.async", "test_wait spin + fence.proxy.async"};
    for (int m = 0; m < 4; m++) printf("mbarrier ping-pong round trip, %-36s: %lld cycles\n", names[m], h[8 + m]);
    return 0;
}
