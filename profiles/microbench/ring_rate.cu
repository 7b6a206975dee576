// ring_rate.cu -- cycles per MMA of a producer / MMA-issuer ring with trivial producers (no data movement): the synchronisation
// ceiling of the planes kernels.  Parameters: NSTG stages in the ring, MPS MMAs per stage, PW producer warps per stage.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../scan_rs_b200/csrc/tc05.cuh"
using namespace tc05;

template <int NSTG, int MPS, int PW, bool SPIN_PROD>
__global__ void __launch_bounds__(32 * (2 + NSTG * PW), 1) k_ring(int stages, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long full[NSTG], empty[NSTG], done;
    __shared__ uint32_t tslot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (uint32_t i = tid; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (tid == 0) {
        for (int i = 0; i < NSTG; i++) { mbar_init(smem_u32(&full[i]), PW); mbar_init(smem_u32(&empty[i]), 1); }
        mbar_init(smem_u32(&done), 1);
        mbar_init_fence();
    }
    if (warp == 0) tmem_alloc_512(smem_u32(&tslot));
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tslot;
    const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
    if (warp == 1) {
        const uint32_t idesc = instr_desc_i8(128, 144, false, false);
        const uint32_t a16 = smem_u32(smem) >> 4, b16 = a16 + (65536 >> 4);
        const uint64_t da_hi = smem_desc(0, 2048, 128), db_hi = smem_desc(0, 144 / 8 * 128, 128);
        const long long t0 = clock64();
        for (int s = 0; s < stages; s++) {
            const uint32_t ss = s % NSTG;
            mbar_spin(full0 + 8 * ss, (s / NSTG) & 1u);
            fence_after_sync();
            if (elect_one()) {
#pragma unroll
                for (int t = 0; t < MPS; t++) mma_i8(tmem, da_hi | (a16 + ss * 256 + (t & 3) * 64), db_hi | (b16 + (t & 7) * 288), idesc, 1u);
                commit(empty0 + 8 * ss);
            }
            __syncwarp();
        }
        if (elect_one()) commit(smem_u32(&done));
        __syncwarp();
        mbar_wait(smem_u32(&done), 0);
        if (lane == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
    } else if (warp >= 2) {
        const int p = warp - 2, ss = p / PW;
        int fills = 0;
        for (int s = ss; s < stages; s += NSTG, fills++) {
            if (fills > 0) { if (SPIN_PROD) mbar_spin(empty0 + 8 * ss, (fills - 1) & 1u); else mbar_wait(empty0 + 8 * ss, (fills - 1) & 1u); }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * ss);
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc_512(tmem);
}

template <int NSTG, int MPS, int PW, bool SPIN>
void run(long long *d) {
    const int stages = 4096;
    auto k = k_ring<NSTG, MPS, PW, SPIN>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    k<<<148, 32 * (2 + NSTG * PW), 96 * 1024>>>(stages, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("stages in ring %2d, MMAs per stage %2d, producer warps per stage %d, producers %s: %6.1f cycles per MMA (floor 72.1)  %s\n", NSTG, MPS, PW,
           SPIN ? "spin    " : "try_wait", (double)h / stages / MPS, cudaGetErrorString(e));
}

int main() {
    long long *d;
    cudaMalloc(&d, 8);
    run<4, 4, 2, false>(d);
    run<4, 4, 2, true>(d);
    run<4, 4, 1, false>(d);
    run<8, 2, 1, false>(d);
    run<8, 4, 1, false>(d);
    run<8, 4, 2, false>(d);
    run<4, 8, 2, false>(d);
    run<2, 8, 4, false>(d);
    run<3, 12, 4, false>(d);
    run<3, 12, 4, true>(d);
    run<2, 16, 4, false>(d);
    return 0;
}
