// jme_peer.cuh - owner-mode exchange of a multi-GPU job over NVLink peer memory.
//
// One process per GPU.  Every rank owns 1 / world of the atoms (by original index), supplies their coordinates and
// wants their gradient rows plus the job-wide energies.  Around the local evaluation that takes
//     an all-gather of the coordinates, a reduce-scatter of the gradient and an all-reduce of the 10-double header.
// With NCCL those are three collectives of 60 - 140 us each at N = 8 (the step itself is 0.6 ms there); here they are
// two small kernels over peer-mapped memory:
//     push   every rank stores its coordinate slice into the full-coordinate buffer of EVERY rank (NVLink stores),
//            fences, and raises its push flag on every rank;
//     pull   after the local evaluation every rank raises its done flag on every rank, waits for the others' and reads
//            the rows of ITS atoms from everybody's partial gradient (NVLink loads), adding them in rank order - a fixed
//            order, so the result does not depend on timing - and likewise the headers.
// Each rank exports ONE allocation (cudaIpcGetMemHandle): control words, the full-coordinate buffer, the result buffer.
// Flags carry the step number, and that is all the flow control there is.  A rank's evaluation of step s + 1 waits
// for everybody's push of s + 1, which follows their pull of s in stream order - so nobody still reads the result
// buffer it is about to overwrite; and a peer's push of s + 1 follows its pull of s, which waited for this rank's
// done flag of s - so the coordinate buffer is not overwritten while this rank's marshalling kernel of step s reads it.
// Waiting threads give up after kPeerTimeoutNs and poison the result (NaN total) instead of hanging the GPU.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "jme_kernels.cuh"

namespace jme {

constexpr int kPeerMaxWorld = 16;
constexpr unsigned long long kPeerTimeoutNs = 4000000000ull;      // 4 s
constexpr size_t kPeerCtrlBytes = 4096;
// control words of a rank's block (unsigned long long each)
constexpr int kPeerPushSeq = 0;                    // [world] step of the last coordinate slice rank p pushed here
constexpr int kPeerDoneSeq = kPeerMaxWorld;        // [world] step whose local evaluation rank p has finished
constexpr int kPeerError = 2 * kPeerMaxWorld;      // != 0: a wait on this rank timed out
constexpr int kPeerBlockCtr = 2 * kPeerMaxWorld + 1;    // [world] "last block" counters of the push kernel

struct PeerPtrs { unsigned char* base[kPeerMaxWorld]; };

__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double2 ld_sys_f64x2(const double* p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long peer_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= seq; false (and the rank's error word set) after the timeout
__device__ __forceinline__ bool peer_spin(const unsigned long long* flag, unsigned long long seq, unsigned long long* err) {
    const unsigned long long t0 = peer_now_ns();
    while (ld_sys_u64(flag) < seq) {
        if (peer_now_ns() - t0 > kPeerTimeoutNs) { st_sys_u64(err, 1ull); return false; }
        __nanosleep(200);
    }
    __threadfence_system();
    return true;
}

// grid (blocks, world): column p stores this rank's slice into rank p's coordinate buffer; the last block of a column
// to finish raises this rank's push flag there
__global__ void __launch_bounds__(256)
peer_push_kernel(PeerPtrs P, int rank, size_t xyz_off, long long slice, const double* __restrict__ src, unsigned long long seq) {
    const int p = blockIdx.y;
    double2* dst = reinterpret_cast<double2*>(P.base[p] + xyz_off) + (slice / 2) * rank;
    const double2* s2 = reinterpret_cast<const double2*>(src);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < slice / 2; i += (long long)gridDim.x * blockDim.x) dst[i] = s2[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long* ctrl = reinterpret_cast<unsigned long long*>(P.base[rank]);
        const unsigned long long t = atomicAdd(ctrl + kPeerBlockCtr + p, 1ull);
        if (t == gridDim.x - 1) {
            ctrl[kPeerBlockCtr + p] = 0;
            __threadfence_system();
            st_sys_u64(reinterpret_cast<unsigned long long*>(P.base[p]) + kPeerPushSeq + rank, seq);
        }
    }
}

// one block: thread p waits for flag p of this rank's block (push or done flags)
__global__ void peer_wait_kernel(unsigned char* base, int first_flag, int world, unsigned long long seq) {
    unsigned long long* ctrl = reinterpret_cast<unsigned long long*>(base);
    if ((int)threadIdx.x < world) peer_spin(ctrl + first_flag + threadIdx.x, seq, ctrl + kPeerError);
}

// one block: the local evaluation of step seq is complete (stream order); tell every rank
__global__ void peer_done_kernel(PeerPtrs P, int rank, int world, unsigned long long seq) {
    __threadfence_system();
    if ((int)threadIdx.x < world)
        st_sys_u64(reinterpret_cast<unsigned long long*>(P.base[threadIdx.x]) + kPeerDoneSeq + rank, seq);
}

// every block waits for all done flags, then sums the rows of this rank's atoms over the ranks' partial results in
// rank order; block 0 also sums the headers.  A timeout anywhere poisons the total.
__global__ void __launch_bounds__(256)
peer_pull_kernel(PeerPtrs P, int rank, int world, size_t out_off, long long slice, double* __restrict__ grad_own,
                 double* __restrict__ header, unsigned long long seq) {
    unsigned long long* ctrl = reinterpret_cast<unsigned long long*>(P.base[rank]);
    if ((int)threadIdx.x < world) peer_spin(ctrl + kPeerDoneSeq + threadIdx.x, seq, ctrl + kPeerError);
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < slice / 2; i += (long long)gridDim.x * blockDim.x) {
        double2 v[kPeerMaxWorld];
#pragma unroll
        for (int p = 0; p < kPeerMaxWorld; ++p) {                 // all loads in flight, then the sum in rank order
            v[p] = make_double2(0.0, 0.0);
            if (p < world)
                v[p] = ld_sys_f64x2(reinterpret_cast<const double*>(P.base[p] + out_off) + kOutHeader + slice * rank + 2 * i);
        }
        double2 acc = v[0];
#pragma unroll
        for (int p = 1; p < kPeerMaxWorld; ++p)
            if (p < world) { acc.x += v[p].x; acc.y += v[p].y; }
        reinterpret_cast<double2*>(grad_own)[i] = acc;
    }
    if (blockIdx.x == 0 && threadIdx.x < kOutHeader) {
        double acc = 0.0;
        for (int p = 0; p < world; ++p) {
            const unsigned long long bits = ld_sys_u64(reinterpret_cast<const unsigned long long*>(P.base[p] + out_off) + threadIdx.x);
            acc += __longlong_as_double((long long)bits);
        }
        if (threadIdx.x == 0 && ld_sys_u64(ctrl + kPeerError) != 0) acc = __longlong_as_double(0x7ff8000000000000ll);
        header[threadIdx.x] = acc;
    }
}

}  // namespace jme
