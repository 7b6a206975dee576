// fp64_rate.cu -- measured FP64 throughput of DFMA and of mma.sync.m8n8k4.f64 (DMMA) on this GPU.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_rate.cu -o fp64_rate && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters) {
    double a[8];
    double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.9999999;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], x, y);
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double *out, int iters) {
    double c[8][2];
    double a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 4 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int threads : {256, 512, 1024}) {
        for (int which = 0; which < 2; which++) {
            int iters = 20000;
            dim3 grid(sms * (1024 / threads) * 1);
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (which == 0) k_dfma<<<grid, threads>>>(out, iters);
                else k_dmma<<<grid, threads>>>(out, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double fma = which == 0 ? (double)grid.x * threads * iters * 8 : (double)grid.x * (threads / 32) * iters * 8 * 256.0;
            printf("%s threads/CTA=%d CTAs=%d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at %d MHz)\n", which ? "DMMA m8n8k4" : "DFMA", threads, grid.x, ms,
                   2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / (p.clockRate * 1e3), p.clockRate / 1000);
        }
    }
    return 0;
}
