// B200 FP64 micro-characterisation (dev tool): dependent-DFMA latency, throughput vs warps x ILP,
// MUFU.RCP64H / F2F / LDC latencies.  nvcc -arch=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kc[64];
template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void rcp_chain(double* out, long long* cyc, int iters)
{
    double x = 1.0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void cvt_chain(double* out, long long* cyc, int iters)
{
    double x = 1.0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) { float f = (float)x; x = (double)f + 1e-9; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void ldc_chain(double* out, long long* cyc, int iters)
{
    double x = 0;
    int idx = threadIdx.x & 1;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) { x = kc[(idx + (int)x) & 63]; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
double run_tput(int blocks_per_sm, int threads, int sms)
{
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 1024 * 1024 * 8); cudaMalloc(&cyc, 8);
    int iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_chain<ILP><<<blocks_per_sm * sms, threads>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaEventRecord(e0);
    dfma_chain<ILP><<<blocks_per_sm * sms, threads>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 16.0 * ILP * iters * (double)blocks_per_sm * sms * threads;
    cudaFree(out); cudaFree(cyc);
    return fl / (ms * 1e-3) / 1e12;
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    int iters = 512;
    dfma_chain<1><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.2f cycles\n", (double)h / (iters * 16));
    dfma_chain<2><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ILP2 single warp: %.2f cycles per DFMA\n", (double)h / (iters * 32));
    dfma_chain<4><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ILP4 single warp: %.2f cycles per DFMA\n", (double)h / (iters * 64));
    dfma_chain<8><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ILP8 single warp: %.2f cycles per DFMA\n", (double)h / (iters * 128));
    dfma_chain<1><<<1, 128>>>(out, cyc, iters, 0.999999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("4 warps (1/SMSP) ILP1: %.2f cycles per DFMA per warp\n", (double)h / (iters * 16));
    dfma_chain<1><<<1, 256>>>(out, cyc, iters, 0.999999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("8 warps (2/SMSP) ILP1: %.2f cycles per DFMA per warp\n", (double)h / (iters * 16));
    dfma_chain<1><<<1, 512>>>(out, cyc, iters, 0.999999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("16 warps (4/SMSP) ILP1: %.2f cycles per DFMA per warp\n", (double)h / (iters * 16));
    dfma_chain<1><<<1, 1024>>>(out, cyc, iters, 0.999999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("32 warps (8/SMSP) ILP1: %.2f cycles per DFMA per warp\n", (double)h / (iters * 16));
    rcp_chain<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent MUFU.RCP64H(+mov) latency: %.2f cycles\n", (double)h / (iters * 16));
    cvt_chain<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("F2F.F32.F64 + F2F.F64.F32 + DADD round trip: %.2f cycles\n", (double)h / (iters * 16));
    ldc_chain<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent LDC.64 (indexed) + F2I round trip: %.2f cycles\n", (double)h / (iters * 16));
    printf("throughput TFLOP/s: ILP1 8x256: %.2f  ILP2 8x256: %.2f  ILP8 8x256: %.2f  ILP8 2x256: %.2f  ILP1 2x256: %.2f ILP2 3x256: %.2f ILP1 3x256: %.2f\n",
           run_tput<1>(8, 256, sms), run_tput<2>(8, 256, sms), run_tput<8>(8, 256, sms), run_tput<8>(2, 256, sms), run_tput<1>(2, 256, sms),
           run_tput<2>(3, 256, sms), run_tput<1>(3, 256, sms));
    return 0;
}
