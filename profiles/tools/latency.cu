// latency.cu — dependent-issue latencies (SM cycles) of the instructions on the kernel's critical path, one warp.
#include <cuda_runtime.h>
#include <cstdio>
#define N 2048
template <class F> __device__ long long chain(F f, double& x) {
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = f(x);
    return clock64() - t0;
}
__global__ void k(double* out, long long* cyc, double a, double b, float fa) {
    double x = a + threadIdx.x * 1e-3;
    cyc[0] = chain([=](double v) { return fma(v, a, b); }, x);
    cyc[1] = chain([=](double v) { return v * a; }, x);
    cyc[2] = chain([=](double v) { return v + b; }, x);
    cyc[3] = chain([=](double v) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v)); return y + b; }, x);   // MUFU.RCP64H + DADD
    cyc[4] = chain([=](double v) { return (double)((float)v) + b; }, x);                                                       // F2F.F32.F64 + F2F.F64.F32 + DADD
    float f = fa + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(f)); f = y + fa; }          // MUFU.LG2 + FADD
    cyc[5] = clock64() - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = fmaf(f, fa, fa);
    cyc[6] = clock64() - t0;
    __shared__ double sm[64];
    sm[threadIdx.x] = x; __syncwarp();
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = (int)sm[idx & 31] & 31;
    cyc[7] = clock64() - t0;
    out[threadIdx.x] = x + f + idx;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 256); cudaMalloc(&cyc, 64);
    k<<<1, 32>>>(out, cyc, 0.999, 1e-3, 1.5f); k<<<1, 32>>>(out, cyc, 0.999, 1e-3, 1.5f);
    long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    const char* nm[8] = {"DFMA", "DMUL", "DADD", "MUFU.RCP64H+DADD", "F2F.F32.F64+F2F.F64.F32+DADD", "MUFU.LG2+FADD", "FFMA", "LDS+F2I+LOP (dependent)"};
    for (int i = 0; i < 8; ++i) printf("%-32s %.1f cycles per link\n", nm[i], (double)h[i] / N);
    return 0;
}
