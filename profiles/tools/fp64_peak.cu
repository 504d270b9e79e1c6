// fp64_peak.cu — measures the FP64 roofline denominator the driver's MEASURED_PEAKS.json lacks:
// sustained DFMA throughput of the whole chip (8 independent FMA chains per thread, 1024 threads/SM x 2 CTAs).
// Built by __graft_entry__.build() into profiles/tools/libfp64_peak.so; called by bench.py (ctypes).
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(512) dfma_chain(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" double fp64_peak_tflops(int device, int iters, int reps, double* best_ms) {
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, device);
    const int blocks = p.multiProcessorCount * 4, threads = 512;
    double* out; if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_chain<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);   // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        dfma_chain<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    if (best_ms) *best_ms = best;
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * blocks * threads;
    return flops / (best * 1e-3) / 1e12;
}
