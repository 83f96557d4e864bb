// Does prefetch.global.L1 (CCTL.E.PF1) really fill L1 on sm_100a?  Load latency after a prefetch.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o prefetch_test prefetch_test.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0 none, 1 prefetch.L2, 2 prefetch.L1, 3 real load first (L1 warm)
__global__ void k(const double* data, size_t stride, long long* out, int reps) {
    long long total = 0;
    double acc = 0;
    for (int r = 0; r < reps; ++r) {
        const double* p = data + (size_t)(r * 32 + threadIdx.x) * stride;
        if (MODE == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        if (MODE == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
        if (MODE == 3) acc += *(volatile const double*)p;
        long long w = clock64();
        while (clock64() - w < 4000) {}
        long long t0 = clock64();
        double v;
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
        acc += v;
        long long t1 = clock64();
        if (acc == 12345.678) t1 += 1;   // use the value before the second clock is trusted
        total += t1 - t0;
    }
    if (threadIdx.x == 0) out[0] = total / reps;
    if (acc == 1.2345) out[1] = 1;
}

int main() {
    const size_t n = (size_t)1 << 28;   // 2 GiB of doubles: every access a fresh line
    double* d; long long* o; long long h;
    cudaMalloc(&d, n * 8); cudaMemset(d, 0, n * 8); cudaMalloc(&o, 16);
    const size_t stride = 4099 * 16;    // doubles: far apart, different lines and pages
    const char* names[] = {"no prefetch", "prefetch.global.L2", "prefetch.global.L1", "loaded before (L1 warm)"};
    for (int m = 0; m < 4; ++m) {
        if (m == 0) k<0><<<1, 32>>>(d, stride, o, 64);
        if (m == 1) k<1><<<1, 32>>>(d, stride, o, 64);
        if (m == 2) k<2><<<1, 32>>>(d, stride, o, 64);
        if (m == 3) k<3><<<1, 32>>>(d, stride, o, 64);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, o, 8, cudaMemcpyDeviceToHost);
        printf("%-28s : %lld cycles per dependent load (32 scattered lanes)\n", names[m], h);
    }
    return 0;
}
