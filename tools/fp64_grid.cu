// Microbenchmark: FP64 issue rate per SM sub-partition as a function of (independent chains per warp, warps per
// sub-partition).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_grid fp64_grid.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void chain_kernel(double* out, int iters, double a, double b, long long* cycles) {
    double v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = threadIdx.x + c;
    __syncthreads();
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

template <int CH>
void row(double* d, long long* c) {
    const int iters = 4096;
    printf("chains %d:", CH);
    for (int w : {1, 2, 3, 4, 5, 6, 8}) {
        long long h;
        chain_kernel<CH><<<1, 128 * w>>>(d, iters, 1.0000001, 1e-9, c);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("  w%d %.2f", w, (double)h / iters / (CH * w));
    }
    printf("   (cycles per DFMA warp-instruction per sub-partition; 2.00 = pipe peak)\n");
}

int main() {
    double* d; long long* c;
    cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 8);
    row<1>(d, c); row<2>(d, c); row<3>(d, c); row<4>(d, c); row<6>(d, c);
    return 0;
}
