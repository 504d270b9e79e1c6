// e2e_floor.cu — host-clock cost of the pieces of one synchronous C-ABI call (C2-sized: 74 KB in, 82 KB out):
// H2D copy launch, kernel launch, completion wait (cudaStreamSynchronize vs. a flag in pinned memory), zero-copy reads.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>
#include <algorithm>
using clk = std::chrono::steady_clock;
__device__ __forceinline__ void busy(long long cyc) { long long t0 = clock64(); while (clock64() - t0 < cyc) {} }
// mimics K1: every CTA reads its 32-chain slice of `in`, works `cyc` cycles, CTA 0 of each group writes outputs; the
// last CTA of the grid raises `flag`
__global__ void k_work(const double* __restrict__ in, double* out, int n, int ncol, long long cyc,
                       unsigned* counter, volatile unsigned* flag, unsigned seq) {
    double acc = 0;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    if (threadIdx.x < 32) for (int k = 0; k < ncol; ++k) acc += in[c + (size_t)k * n];
    busy(cyc);
    if (blockIdx.y == 0 && threadIdx.x < 32) for (int k = 0; k <= ncol; ++k) out[c + (size_t)k * n] = acc + k;
    if (flag) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            unsigned t = atomicAdd(counter, 1u);
            if (t == gridDim.x * gridDim.y - 1) { *counter = 0; __threadfence_system(); *flag = seq; }
        }
    }
}
template <class F> double med(F f, int reps = 400) {
    std::vector<double> v;
    for (int i = 0; i < reps + 50; ++i) { auto a = clk::now(); f(i); auto b = clk::now(); if (i >= 50) v.push_back(std::chrono::duration<double, std::micro>(b - a).count()); }
    std::sort(v.begin(), v.end()); return v[v.size() / 2];
}
int main() {
    const int n = 1024, ncol = 9; const size_t ib = (size_t)n * ncol * 8, ob = (size_t)n * (ncol + 1) * 8;
    double *h_in, *h_out, *d_in, *d_out; unsigned *counter; volatile unsigned* flag;
    cudaMallocHost(&h_in, ib); cudaMallocHost(&h_out, ob); cudaMallocHost((void**)&flag, 64); *flag = 0;
    cudaMalloc(&d_in, ib); cudaMalloc(&d_out, ob); cudaMalloc(&counter, 4); cudaMemset(counter, 0, 4);
    for (size_t i = 0; i < (size_t)n * ncol; ++i) h_in[i] = 1.0;
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    dim3 grid(32, 9), blk(256);
    for (long long cyc : {0LL, 20000LL}) {   // 20000 cycles ~ 10.5 us: C2's kernel body
        printf("--- kernel body %lld cycles\n", cyc);
        printf("kernel(dev in, dev out) + sync                 : %6.2f us\n", med([&](int) { k_work<<<grid, blk, 0, st>>>(d_in, d_out, n, ncol, cyc, counter, nullptr, 0); cudaStreamSynchronize(st); }));
        printf("kernel(dev in, host out) + sync                : %6.2f us\n", med([&](int) { k_work<<<grid, blk, 0, st>>>(d_in, h_out, n, ncol, cyc, counter, nullptr, 0); cudaStreamSynchronize(st); }));
        printf("H2D + kernel(dev in, host out) + sync          : %6.2f us\n", med([&](int) { cudaMemcpyAsync(d_in, h_in, ib, cudaMemcpyHostToDevice, st); k_work<<<grid, blk, 0, st>>>(d_in, h_out, n, ncol, cyc, counter, nullptr, 0); cudaStreamSynchronize(st); }));
        printf("H2D + kernel(dev in, host out) + flag spin     : %6.2f us\n", med([&](int i) { cudaMemcpyAsync(d_in, h_in, ib, cudaMemcpyHostToDevice, st); k_work<<<grid, blk, 0, st>>>(d_in, h_out, n, ncol, cyc, counter, flag, (unsigned)i + 1); while (*flag != (unsigned)i + 1) {} }));
        printf("kernel(host in zero-copy, host out) + sync     : %6.2f us\n", med([&](int) { k_work<<<grid, blk, 0, st>>>(h_in, h_out, n, ncol, cyc, counter, nullptr, 0); cudaStreamSynchronize(st); }));
        printf("kernel(host in zero-copy, host out) + flag spin: %6.2f us\n", med([&](int i) { k_work<<<grid, blk, 0, st>>>(h_in, h_out, n, ncol, cyc, counter, flag, (unsigned)i + 1); while (*flag != (unsigned)i + 1) {} }));
        printf("kernel(dev in, host out) + flag spin           : %6.2f us\n", med([&](int i) { k_work<<<grid, blk, 0, st>>>(d_in, h_out, n, ncol, cyc, counter, flag, (unsigned)i + 1); while (*flag != (unsigned)i + 1) {} }));
        cudaStreamSynchronize(st);
    }
    // the pieces on the host side alone
    printf("cudaMemcpyAsync call (host time)               : %6.2f us\n", med([&](int) { cudaMemcpyAsync(d_in, h_in, ib, cudaMemcpyHostToDevice, st); }, 200)); cudaStreamSynchronize(st);
    printf("kernel launch call (host time)                 : %6.2f us\n", med([&](int) { k_work<<<grid, blk, 0, st>>>(d_in, d_out, n, ncol, 0, counter, nullptr, 0); }, 200)); cudaStreamSynchronize(st);
    printf("cudaStreamSynchronize on idle stream           : %6.2f us\n", med([&](int) { cudaStreamSynchronize(st); }, 200));
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
