// Cost of feeding a per-use constant addend to DFMA on B200 (dev tool): constant bank vs shared memory vs registers.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kc[64];
enum { SRC_REG, SRC_LDC, SRC_LDS64, SRC_LDS128, SRC_LDG };
template <int SRC>
__global__ void k(double* out, long long* cyc, int iters, double a, const double* gk)
{
    __shared__ __align__(16) double sk[64];
    if (threadIdx.x < 64) sk[threadIdx.x] = 0.999 + 1e-6 * threadIdx.x;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    __syncthreads();
    unsigned sbase = (unsigned)__cvta_generic_to_shared(sk);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                double c0, c1;
                const int idx = u * 8 + i;
                if (SRC == SRC_REG) { c0 = a; c1 = a; }
                if (SRC == SRC_LDC) {
                    asm volatile("ld.const.f64 %0, [kc+%1];" : "=d"(c0) : "n"(0), "r"(it) : "memory");
                    asm volatile("ld.const.f64 %0, [kc+%1];" : "=d"(c1) : "n"(8), "r"(it) : "memory");
                }
                if (SRC == SRC_LDS64) {
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(c0) : "r"(sbase + idx * 8));
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(c1) : "r"(sbase + idx * 8 + 8));
                }
                if (SRC == SRC_LDS128) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c0), "=d"(c1) : "r"(sbase + idx * 8));
                if (SRC == SRC_LDG) {
                    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(c0) : "l"(gk + idx));
                    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(c1) : "l"(gk + idx + 1));
                }
                x[i] = fma(x[i], a, c0);
                x[i + 1] = fma(x[i + 1], a, c1);
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
template <int SRC>
void run(const char* name, const double* gk)
{
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    for (int threads : {128, 512, 768}) {
        const int iters = 1000;
        k<SRC><<<1, threads>>>(out, cyc, iters, 0.999999, gk);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-30s %d warps/sub-partition: %.2f cycles per DFMA per warp\n", name, threads / 128, (double)h / (iters * 32.0 * (threads / 128)));
    }
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    double h[64]; for (int i = 0; i < 64; i++) h[i] = 0.999 + 1e-6 * i;
    cudaMemcpyToSymbol(kc, h, sizeof(h));
    double* gk; cudaMalloc(&gk, sizeof(h)); cudaMemcpy(gk, h, sizeof(h), cudaMemcpyHostToDevice);
    run<SRC_REG>("addend in a register", gk);
    run<SRC_LDC>("addend by ld.const.f64", gk);
    run<SRC_LDS64>("addend by ld.shared.f64", gk);
    run<SRC_LDS128>("two addends by ld.shared.v2", gk);
    run<SRC_LDG>("addend by ld.global.nc.f64", gk);
    return 0;
}
