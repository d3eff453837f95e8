// Per-instruction issue cost table on B200 at 4 warps / SM sub-partition (dev tool): 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kc[64];
enum { OP_DFMA, OP_DMUL, OP_DADD, OP_DSETP_SEL, OP_MUFU_RCP, OP_MUFU_RSQ, OP_F2F, OP_DFMA_LDC, OP_FFMA, OP_DMNMX, OP_F2I };
template <int OP>
__global__ void k(double* out, long long* cyc, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == OP_DFMA) x[i] = fma(x[i], a, b);
                if (OP == OP_DMUL) x[i] = x[i] * a;
                if (OP == OP_DADD) x[i] = x[i] + b;
                if (OP == OP_DSETP_SEL) x[i] = (x[i] > a) ? b : x[i] + 1e-300;     // DSETP + DADD + SEL
                if (OP == OP_MUFU_RCP) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i])); x[i] = r; }
                if (OP == OP_MUFU_RSQ) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i])); x[i] = r; }
                if (OP == OP_F2F) { float f = (float)x[i]; x[i] = (double)f; }
                if (OP == OP_DFMA_LDC) x[i] = fma(x[i], kc[(u * 8 + i) & 63], kc[(u * 8 + i + 32) & 63]);
                if (OP == OP_FFMA) { float f = __double2float_rn(x[i]); f = fmaf(f, 0.999f, 1e-3f); x[i] = f; }
                if (OP == OP_DMNMX) x[i] = fmax(x[i], a);
                if (OP == OP_F2I) { int q = __double2int_rn(x[i]); x[i] = (double)q; }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name)
{
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 1000, threads = 512;
    k<OP><<<1, threads>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s %.2f cycles per source-level op per warp (4 warps / sub-partition)\n", name, (double)h / (iters * 32.0 * 4));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    double h[64]; for (int i = 0; i < 64; i++) h[i] = 0.999 + 1e-6 * i;
    cudaMemcpyToSymbol(kc, h, sizeof(h));
    run<OP_DFMA>("DFMA");
    run<OP_DMUL>("DMUL");
    run<OP_DADD>("DADD");
    run<OP_DSETP_SEL>("DSETP + DADD + SEL");
    run<OP_DMNMX>("fmax (DSETP+SEL / DMNMX)");
    run<OP_MUFU_RCP>("MUFU.RCP64H (+mov)");
    run<OP_MUFU_RSQ>("MUFU.RSQ64H (+mov)");
    run<OP_F2F>("F2F f64->f32 + f32->f64");
    run<OP_F2I>("F2I f64->s32 + I2F s32->f64");
    run<OP_DFMA_LDC>("DFMA with two constant operands");
    run<OP_FFMA>("F2F + FFMA + F2F");
    return 0;
}
