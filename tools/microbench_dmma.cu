// DMMA (mma.sync.aligned.m8n8k4.f64) issue-rate probe on sm_100a (dev tool):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_dmma tools/microbench_dmma.cu
// Prints the FP64 tensor-core throughput next to the DFMA throughput measured the same way.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b)
{
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static double time_ms(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    const double warps = (double)blocks * threads / 32;
    {
        double ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, iters, 0.999999, 1e-9); });
        printf("DFMA            : %.2f TFLOP/s\n", 2.0 * 32 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
    }
#define RUN(N)                                                                                                   \
    {                                                                                                            \
        double ms = time_ms([&] { k_dmma<N><<<blocks, threads>>>(out, iters, 0.999999, 1e-9); });                  \
        printf("DMMA m8n8k4 x%-2d : %.2f TFLOP/s (%.2f cycles per warp-instruction per SM sub-partition at 1.965 GHz)\n", N,       \
               2.0 * 256 * N * iters * warps / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.965e9 / (N * (double)iters * warps / (p.multiProcessorCount * 4)));  \
    }
    RUN(1) RUN(2) RUN(4) RUN(8)
    return 0;
}
