// Does a half-rate DFMA occupy the issue port for two cycles on B200?  (dev tool)
// Per loop trip: 8 independent DFMA chains plus NF independent FFMA (or IMAD) chains, interleaved.
#include <cstdio>
#include <cuda_runtime.h>
template <int NF32, int NI32>
__global__ void mix(double* out, long long* cyc, int iters, double a, double b, float fa, float fb, int ia)
{
    double x[8];
    float y[NF32 > 0 ? NF32 : 1];
    int z[NI32 > 0 ? NI32 : 1];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < NF32; i++) y[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < NI32; i++) z[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                x[i] = fma(x[i], a, b);
                if (i < NF32) y[i] = fmaf(y[i], fa, fb);
                if (i + 8 < NF32) y[i + 8] = fmaf(y[i + 8], fa, fb);
                if (i < NI32) z[i] = z[i] * ia + 7;
                if (i + 8 < NI32) z[i + 8] = z[i + 8] * ia + 7;
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
#pragma unroll
    for (int i = 0; i < NF32; i++) s += y[i];
#pragma unroll
    for (int i = 0; i < NI32; i++) s += z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NF32, int NI32>
void run(const char* name, int threads)
{
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    int iters = 2000;
    mix<NF32, NI32><<<1, threads>>>(out, cyc, iters, 0.999999, 1e-9, 0.9999f, 1e-3f, 3);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s threads %4d: %.2f cycles per group of 8 DFMA (+%d FFMA +%d IMAD)\n", name, threads, (double)h / (iters * 4), NF32, NI32);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int threads : {128, 512}) {
        run<0, 0>("8 DFMA", threads);
        run<8, 0>("8 DFMA + 8 FFMA", threads);
        run<16, 0>("8 DFMA + 16 FFMA", threads);
        run<0, 8>("8 DFMA + 8 IMAD", threads);
        run<0, 16>("8 DFMA + 16 IMAD", threads);
        run<8, 8>("8 DFMA + 8 FFMA + 8 IMAD", threads);
    }
    return 0;
}
