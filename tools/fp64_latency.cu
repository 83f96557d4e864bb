// Microbenchmark: dependent-issue latency and throughput of FP64 instructions on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void chain_kernel(double* out, int iters, double a, double b, long long* cycles) {
    double v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = threadIdx.x + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) v[c] = fma(v[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void mufu_chain(double* out, int iters, long long* cycles) {
    double v = 1.5 + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        double y;
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
        v = y + 2.0;
    }
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    double* d; long long* c; long long h;
    cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 8);
    const int iters = 4096;
    // one warp on one SM: pure dependent latency
#define RUN(CH, WARPS) do { chain_kernel<CH><<<1, 32 * WARPS>>>(d, iters, 1.0000001, 1e-9, c); cudaDeviceSynchronize(); \
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
        printf("chains/thread %d warps/CTA %2d : %.2f cycles per DFMA issue slot per warp (%.2f cycles per loop iter)\n", CH, WARPS, (double)h / iters / CH, (double)h / iters); } while (0)
    RUN(1, 1); RUN(2, 1); RUN(4, 1); RUN(8, 1);
    RUN(1, 4); RUN(1, 8); RUN(1, 16); RUN(1, 32);
    RUN(2, 16); RUN(4, 8); RUN(8, 4);
    mufu_chain<<<1, 32>>>(d, iters, c); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("MUFU.RSQ64H + DADD dependent pair: %.2f cycles\n", (double)h / iters);
    return 0;
}
