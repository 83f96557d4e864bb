// jme_cells.cuh - cell-list path for cutoff > 0 (the reference has no neighbour list at all: it tests
// every non-excluded pair against the cutoff, MMFF94VdwComponent.java:136-139; this path selects the
// SAME pair set with O(N) work).
//
// Per evaluation, all on the device, no host round trip:
//   1. bounding box (two-stage min/max reduction)              -> grid origin / edge / dims in device memory
//   2. cell index per atom, per-cell counts (integer atomics)   -> exclusive scan -> cell start offsets
//   3. scatter to cells, then rank inside each cell by original atom index, so the sorted order (and
//      with it every later summation order) is deterministic; SoA gather into sorted arrays (FP64
//      positions/charges, type, and an FP32 copy of the positions relative to the grid origin)
//   4. cells_pair_kernel: one warp per atom i.  Lanes scan the candidate atoms of the neighbouring cell
//      rows with a conservative FP32 distance test (superset of the true pair set), compact the
//      survivors into a per-warp shared-memory queue, and every time 32 are queued the warp evaluates
//      them with all lanes busy: exact FP64 predicate (bit-identical to the reference's
//      sqrt(dx*dx+dy*dy+dz*dz) > cutoff via the r^2 threshold), exclusion / 1-4 lookup, vdW + Coulomb.
//      Each pair is evaluated from both of its atoms (no Newton's third law): no scatter, no atomics,
//      deterministic; energies are halved.  The row ranges of an atom are computed lane-parallel (one
//      lane per neighbour cell row, culled against the cutoff sphere in FP32 with padded margins).
#pragma once

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "jme_host.hpp"
#include "jme_kernels.cuh"

namespace jme {

#ifndef JME_CELL_WARPS
#define JME_CELL_WARPS 8
#endif
#ifndef JME_CELL_MIN_BLOCKS
#define JME_CELL_MIN_BLOCKS 3
#endif
constexpr int kCellWarps = JME_CELL_WARPS;
constexpr int kCellThreads = kCellWarps * 32;
constexpr int kScanBlock = 1024;          // elements per scan block

struct GridParams {
    double ox, oy, oz;       // origin (bounding-box minimum)
    double edge, inv_edge;
    int nx, ny, nz, ncell;
    int reach;               // neighbour cells to visit per direction: ceil(cutoff / edge)
    float r2f_max;           // conservative FP32 candidate threshold on r^2
    double cutoff_pad;       // cutoff + margin used for row culling
    float edge_f, inv_edge_f, cut2_pad_f;   // FP32 copies for the per-row culling (margins included)
};

struct CellArrays {
    GridParams* grid;
    int n, n_pad, max_cells;
    int* cell_of;            // [n] cell index per original atom
    int* count;              // [max_cells + 1]
    int* start;              // [max_cells + 1]
    int* cursor;             // [max_cells]
    int* block_sums;         // scan scratch
    int* order_tmp;          // [n]
    int* orig;               // [n] sorted position -> original atom
    int* scell;              // [n] cell of sorted position
    double2* sxy;            // [n] sorted x, y   } two 16-byte records per atom: a gather of either touches
    double2* szq;            // [n] sorted z, q   } half the cache lines a 32-byte record would
    float4* fxyz;            // [n] sorted positions relative to the grid origin, FP32 (w unused)
    int* stype;
    const long long* excl_ptr;    // CSR over original atom index
    const unsigned int* excl_ent;
    double* bbox_part;       // [blocks][6]
    int bbox_blocks;
};

// ---- 1. bounding box ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bbox_kernel(DeviceSystem S, double* __restrict__ part) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < S.n; a += gridDim.x * blockDim.x) {
        const double v[3] = {S.x[a], S.y[a], S.z[a]};
#pragma unroll
        for (int c = 0; c < 3; ++c) { lo[c] = fmin(lo[c], v[c]); hi[c] = fmax(hi[c], v[c]); }
    }
    __shared__ double s[8][6];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if (lane == 0) { s[warp][c] = lo[c]; s[warp][3 + c] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = s[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fmin(v, s[w][threadIdx.x]) : fmax(v, s[w][threadIdx.x]);
        part[6 * blockIdx.x + threadIdx.x] = v;
    }
}

__global__ void grid_setup_kernel(const double* __restrict__ part, int n_part, double cutoff, int max_cells, GridParams* g) {
    // one warp: lanes stride over the per-block bounding boxes, shuffle min/max, lane 0 derives the grid
    const int lane = threadIdx.x;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int b = lane; b < n_part; b += 32)
        for (int c = 0; c < 3; ++c) { lo[c] = fmin(lo[c], part[6 * b + c]); hi[c] = fmax(hi[c], part[6 * b + 3 + c]); }
    for (int c = 0; c < 3; ++c)
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    if (lane != 0) return;
    double ext[3], extent = 0;
    for (int c = 0; c < 3; ++c) { ext[c] = fmax(hi[c] - lo[c], 0.0); extent = fmax(extent, ext[c]); }
    // cell edge: half the cutoff (5x5x5 stencil, about 27 % of it inside the sphere before row culling),
    // grown if the box would need more cells than were allocated
    double edge = 0.5 * cutoff;
    for (;;) {
        const double nx = floor(ext[0] / edge) + 1, ny = floor(ext[1] / edge) + 1, nz = floor(ext[2] / edge) + 1;
        if (nx * ny * nz <= (double)max_cells) {
            g->nx = (int)nx; g->ny = (int)ny; g->nz = (int)nz;
            break;
        }
        edge *= 1.25;
    }
    g->ox = lo[0]; g->oy = lo[1]; g->oz = lo[2];
    g->edge = edge;
    g->inv_edge = 1.0 / edge;
    g->ncell = g->nx * g->ny * g->nz;
    g->reach = (int)ceil(cutoff / edge);
    // FP32 candidate test on positions relative to the origin: coordinate error <= extent * 2^-24 each,
    // so |r_f - r| <= 4 * extent * 2^-23 with room to spare; r2 margin = 2 r dr + float product rounding
    const double dr = 8.0 * (extent + cutoff) * 1.1920929e-7 + 1e-6;
    const double rpad = cutoff + dr;
    g->r2f_max = (float)(rpad * rpad * (1.0 + 1e-6)) * (1.0f + 4e-7f);
    g->cutoff_pad = cutoff + 2.0 * dr + 1e-6;
    // row culling runs in FP32 on origin-relative coordinates: pad by the FP32 coordinate / cell-boundary
    // error once more, and round every FP32 constant in the conservative direction
    const double cpad = g->cutoff_pad + 4.0 * dr;
    g->edge_f = (float)edge;
    g->inv_edge_f = (float)(1.0 / edge);
    g->cut2_pad_f = (float)(cpad * cpad) * (1.0f + 4e-7f);
}

// ---- 2. binning --------------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(double v, double o, double inv_edge, int n) {
    int c = (int)floor((v - o) * inv_edge);
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

__global__ void zero_counts_kernel(const GridParams* __restrict__ g, int* __restrict__ count, int* __restrict__ cursor) {
    const int nc = g->ncell;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= nc; c += gridDim.x * blockDim.x) {
        count[c] = 0;
        if (c < nc) cursor[c] = 0;
    }
}

__global__ void cell_assign_kernel(DeviceSystem S, const GridParams* __restrict__ g, int* __restrict__ cell_of, int* __restrict__ count) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= S.n) return;
    const int cx = cell_coord(S.x[a], g->ox, g->inv_edge, g->nx);
    const int cy = cell_coord(S.y[a], g->oy, g->inv_edge, g->ny);
    const int cz = cell_coord(S.z[a], g->oz, g->inv_edge, g->nz);
    const int c = (cz * g->ny + cy) * g->nx + cx;
    cell_of[a] = c;
    atomicAdd(&count[c], 1);
}

// exclusive scan of count[0..ncell] -> start[0..ncell] in three small launches
__global__ void __launch_bounds__(kScanBlock) scan_local_kernel(const GridParams* __restrict__ g, const int* __restrict__ count,
                                                                int* __restrict__ start, int* __restrict__ block_sums) {
    const int total = g->ncell + 1;
    const int base = blockIdx.x * kScanBlock;
    if (base >= total) return;
    __shared__ int s[kScanBlock];
    const int i = base + threadIdx.x;
    const int v = i < total ? count[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < kScanBlock; o <<= 1) {
        const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < total) start[i] = s[threadIdx.x] - v;
    if (threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = s[threadIdx.x];
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(const GridParams* __restrict__ g, int* __restrict__ block_sums) {
    // serial over chunks of 1024 block sums (at most max_cells / 1024 entries)
    const int nb = (g->ncell + 1 + kScanBlock - 1) / kScanBlock;
    __shared__ int s[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? block_sums[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nb) block_sums[i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanBlock) scan_add_kernel(const GridParams* __restrict__ g, int* __restrict__ start,
                                                              const int* __restrict__ block_sums) {
    const int total = g->ncell + 1;
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    if (i < total) start[i] += block_sums[blockIdx.x];
}

// ---- 3. sort -----------------------------------------------------------------------------------
__global__ void cell_scatter_kernel(int n, const int* __restrict__ cell_of, const int* __restrict__ start,
                                    int* __restrict__ cursor, int* __restrict__ order_tmp) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int c = cell_of[a];
    order_tmp[start[c] + atomicAdd(&cursor[c], 1)] = a;
}

// position p of the atomics-ordered list -> its rank among the atoms of its cell by original index
__global__ void cell_rank_gather_kernel(DeviceSystem S, CellArrays C) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= S.n) return;
    const int a = C.order_tmp[p];
    const int c = C.cell_of[a];
    const int b = C.start[c], e = C.start[c + 1];
    int rank = 0;
    for (int q = b; q < e; ++q) rank += C.order_tmp[q] < a;
    const int d = b + rank;
    const GridParams* g = C.grid;
    C.orig[d] = a;
    C.scell[d] = c;
    const double x = S.x[a], y = S.y[a], z = S.z[a];
    C.sxy[d] = make_double2(x, y);
    C.szq[d] = make_double2(z, S.q[a]);
    C.stype[d] = S.type[a];
    C.fxyz[d] = make_float4((float)(x - g->ox), (float)(y - g->oy), (float)(z - g->oz), 0.0f);
}

// ---- 4. pairs ----------------------------------------------------------------------------------
struct CellPairArgs {
    int i_begin, i_end;            // sorted positions handled by this shard
    double* grad;                  // [3][n_pad] by original atom index
    double* epart;                 // [warps][2]
    unsigned long long* cpart;     // [warps][2]
    // pair-list emission (EMIT only)
    int* out_i; int* out_j; unsigned char* out_14; long long capacity; unsigned long long* cursor;
};

constexpr int kQueue = 128;          // per-warp candidate ring (power of two)

template <bool GRAD, bool EMIT>
__global__ void __launch_bounds__(kCellThreads, JME_CELL_MIN_BLOCKS)
cells_pair_kernel(DeviceSystem S, CellArrays C, CellPairArgs P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PairConst* s_table = reinterpret_cast<PairConst*>(smem_raw);
    const int TT = S.n_types * S.n_types;
    unsigned int* s_queue = reinterpret_cast<unsigned int*>(s_table + TT);
    int2* s_rows = reinterpret_cast<int2*>(s_queue + kCellWarps * kQueue);
    unsigned int* s_excl = reinterpret_cast<unsigned int*>(s_rows + kCellWarps * 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned int* queue = s_queue + warp * kQueue;
    int2* rows = s_rows + warp * 32;
    unsigned int* excl = s_excl + warp * 32;
    for (int t = threadIdx.x; t < TT; t += kCellThreads) s_table[t] = S.table[t];
    __syncthreads();

    const GridParams g = *C.grid;
    const int gw = blockIdx.x * kCellWarps + warp;
    const int n_warps = gridDim.x * kCellWarps;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int width = 2 * g.reach + 1;
    const int n_rows = width * width;          // <= 25: reach <= 2 by construction (edge >= cutoff / 2)
    const double2* __restrict__ sxy = C.sxy;
    const double2* __restrict__ szq = C.szq;
    const float4* __restrict__ fxyz = C.fxyz;
    const int* __restrict__ stype = C.stype;
    const int* __restrict__ orig = C.orig;

    double ev_w = 0, ee_w = 0;
    unsigned long long n_eval_w = 0, n_cand_w = 0;

    for (int i = P.i_begin + gw; i < P.i_end; i += n_warps) {
        const double2 pi_xy = sxy[i], pi_zq = szq[i];
        const float4 fi = fxyz[i];
        const double qi2 = 2.0 * S.ele_scale * pi_zq.y;      // doubled energies, see pair_terms
        const PairConst* trow = s_table + stype[i] * S.n_types;
        const int oi = orig[i];
        const long long eb = C.excl_ptr[oi];
        const int n_excl = (int)(C.excl_ptr[oi + 1] - eb);
        const int ci = C.scell[i];
        const int cx = ci % g.nx, cy = (ci / g.nx) % g.ny, cz = ci / (g.nx * g.ny);

        // ---- lane-parallel: the candidate range of neighbour cell row `lane`, culled against the sphere
        {
            int jb = 0, je = 0;
            if (lane < n_rows) {
                const int ddy = lane % width - g.reach, ddz = lane / width - g.reach;
                const int y2 = cy + ddy, z2 = cz + ddz;
                if (y2 >= 0 && y2 < g.ny && z2 >= 0 && z2 < g.nz) {
                    // distance from the atom to the cell slabs y2 / z2 (0 inside), origin-relative FP32
                    float my = 0.f, mz = 0.f;
                    if (ddy > 0) my = y2 * g.edge_f - fi.y; else if (ddy < 0) my = fi.y - (y2 + 1) * g.edge_f;
                    if (ddz > 0) mz = z2 * g.edge_f - fi.z; else if (ddz < 0) mz = fi.z - (z2 + 1) * g.edge_f;
                    my = fmaxf(my, 0.f);
                    mz = fmaxf(mz, 0.f);
                    const float rem = g.cut2_pad_f - (my * my + mz * mz);
                    if (rem >= 0.f) {
                        const float xr = sqrtf(rem) * (1.0f + 1e-6f);
                        int x0 = (int)floorf((fi.x - xr) * g.inv_edge_f - 1e-4f);
                        int x1 = (int)floorf((fi.x + xr) * g.inv_edge_f + 1e-4f);
                        x0 = max(max(x0, cx - g.reach), 0);
                        x1 = min(min(x1, cx + g.reach), g.nx - 1);
                        if (x0 <= x1) {
                            const int row = (z2 * g.ny + y2) * g.nx;
                            jb = C.start[row + x0];
                            je = C.start[row + x1 + 1];
                        }
                    }
                }
            }
            __syncwarp();
            rows[lane] = make_int2(jb, je);
            if (n_excl > 0) excl[lane] = lane < n_excl ? C.excl_ent[eb + lane] : 0xffffffffu;
            __syncwarp();
        }

        double gix = 0, giy = 0, giz = 0, ev = 0, ee = 0;
        unsigned int n_eval = 0, n_cand = 0;
        unsigned int q_head = 0, q_tail = 0;       // ring indices (monotonic; entries at index & (kQueue-1))

        // ---- heavy stage: the 32 (or, at the end, fewer) oldest queue entries, lane l taking entry l
        auto process = [&](int m) {
            const bool active = lane < m;
            const int j = active ? (int)queue[(q_head + lane) & (kQueue - 1)] : i;
            q_head += m;
            const double2 pj_xy = sxy[j], pj_zq = szq[j];
            const int tj = stype[j];
            const double dx = pi_xy.x - pj_xy.x, dy = pi_xy.y - pj_xy.y, dz = pi_zq.x - pj_zq.x;
            // strict double, left-to-right, no FMA: Point3d.distance squared
            const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            bool on = active && !(r2 > S.r2_max);
            bool f14 = false;
            bool counted = on && j > i;            // each unordered pair is counted from its lower sorted index
            int oj = 0;
            if (EMIT) oj = orig[j];
            if (n_excl > 0) {                       // warp-uniform: the list belongs to atom i
                if (!EMIT) oj = orig[j];
                const int n_sm = min(n_excl, 32);
                for (int e = 0; e < n_sm; ++e) {
                    const unsigned int x = excl[e];
                    if ((int)(x & ~kFlag14) == oj) {
                        if (x & kFlag14) f14 = true; else on = false;
                    }
                }
                for (int e = 32; e < n_excl; ++e) {
                    const unsigned int x = C.excl_ent[eb + e];
                    if ((int)(x & ~kFlag14) == oj) {
                        if (x & kFlag14) f14 = true; else on = false;
                    }
                }
                counted = on && j > i;
            }
            n_eval += __popc(__ballot_sync(0xffffffffu, counted));
            if (EMIT) {
                if (counted) {
                    const unsigned long long p = atomicAdd(P.cursor, 1ull);
                    if ((long long)p < P.capacity) { P.out_i[p] = min(oi, oj); P.out_j[p] = max(oi, oj); P.out_14[p] = f14; }
                }
            } else {
                double qq2 = qi2 * pj_zq.y;
                if (f14) qq2 *= kEle14;
                // coincident atoms: keep the (finite) buffered energy, no gradient direction exists
                double e1, e2, f;
                pair_terms<GRAD>(on ? clamp_r2(r2) : 1.0, trow[tj], qq2, e1, e2, f);
                ev += on ? e1 : 0.0;
                ee += on ? e2 : 0.0;
                if (GRAD) {
                    f = on ? f : 0.0;
                    gix = fma(f, dx, gix);
                    giy = fma(f, dy, giy);
                    giz = fma(f, dz, giz);
                }
            }
        };

        // ---- candidate scan: trips of 64 candidates (two loads in flight) over the culled row ranges
        for (int r = 0; r < n_rows; ++r) {
            const int2 rg = rows[r];
            for (int base = rg.x; base < rg.y; base += 64) {
                const int j0 = base + lane, j1 = j0 + 32;
                const float4 a0 = fxyz[min(j0, rg.y - 1)];
                const float4 a1 = fxyz[min(j1, rg.y - 1)];
                const float x0 = fi.x - a0.x, y0 = fi.y - a0.y, z0 = fi.z - a0.z;
                const float x1 = fi.x - a1.x, y1 = fi.y - a1.y, z1 = fi.z - a1.z;
                const bool h0 = (x0 * x0 + y0 * y0 + z0 * z0) <= g.r2f_max && j0 < rg.y && j0 != i;
                const bool h1 = (x1 * x1 + y1 * y1 + z1 * z1) <= g.r2f_max && j1 < rg.y && j1 != i;
                const unsigned int m0 = __ballot_sync(0xffffffffu, h0);
                const unsigned int m1 = __ballot_sync(0xffffffffu, h1);
                if ((m0 | m1) == 0) continue;
                const int c0n = __popc(m0);
                if (h0) queue[(q_tail + __popc(m0 & lt_mask)) & (kQueue - 1)] = (unsigned int)j0;
                if (h1) queue[(q_tail + c0n + __popc(m1 & lt_mask)) & (kQueue - 1)] = (unsigned int)j1;
                q_tail += c0n + __popc(m1);
                n_cand += c0n + __popc(m1);
                __syncwarp();
                while (q_tail - q_head >= 32) process(32);
            }
        }
        if (q_tail != q_head) process((int)(q_tail - q_head));
        __syncwarp();

        if (GRAD && !EMIT) {
            gix = warp_sum(gix);
            giy = warp_sum(giy);
            giz = warp_sum(giz);
            if (lane == 0) {
                P.grad[oi] = gix;
                P.grad[S.n_pad + oi] = giy;
                P.grad[2 * (size_t)S.n_pad + oi] = giz;
            }
        }
        ev_w += ev;
        ee_w += ee;
        n_eval_w += n_eval;
        n_cand_w += n_cand;
    }
    if (!EMIT) {
        // energies were accumulated doubled (pair_terms) and every pair was seen from both atoms: x 1/4
        ev_w = 0.25 * warp_sum(ev_w);
        ee_w = 0.25 * warp_sum(ee_w);
        if (lane == 0) {
            P.epart[2 * gw] = ev_w;
            P.epart[2 * gw + 1] = ee_w;
            P.cpart[2 * gw] = n_eval_w;
            P.cpart[2 * gw + 1] = n_cand_w;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------
struct CellPlan {
    CellArrays arr{};
    double* d_grad = nullptr;
    double* d_epart = nullptr;
    unsigned long long* d_cpart = nullptr;
    int n_epart = 0;
    int grid_pairs = 0;
    int i_begin = 0, i_end = 0, world = 1;
    double cutoff = 0;
    bool built = false;
    bool sorted_valid = false;
    std::vector<void*> new_allocs;

    std::vector<void*> take_new_allocs() { std::vector<void*> v; v.swap(new_allocs); return v; }
    void coords_changed() { sorted_valid = false; }
};

template <class T>
inline bool cell_alloc(CellPlan& p, T** out, size_t count, std::string& err) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) { err = std::string("cudaMalloc (cell path): ") + cudaGetErrorString(e); return false; }
    p.new_allocs.push_back(q);
    *out = static_cast<T*>(q);
    return true;
}

inline int build_cell_plan(CellPlan& p, const HostSystem& hs, const DeviceSystem& ds, double cutoff, int rank, int world,
                           std::string& err) {
    const int n = ds.n;
    if (!p.built) {
        CellArrays& a = p.arr;
        a.n = n; a.n_pad = ds.n_pad;
        a.max_cells = std::max(4096, 4 * n);
        a.bbox_blocks = std::max(1, std::min(592, (n + 255) / 256));
        const int scan_blocks = (a.max_cells + 1 + kScanBlock - 1) / kScanBlock;
        bool ok = cell_alloc(p, &a.grid, 1, err) && cell_alloc(p, &a.cell_of, n, err) && cell_alloc(p, &a.count, a.max_cells + 1, err) &&
                  cell_alloc(p, &a.start, a.max_cells + 1, err) && cell_alloc(p, &a.cursor, a.max_cells, err) &&
                  cell_alloc(p, &a.block_sums, scan_blocks, err) && cell_alloc(p, &a.order_tmp, n, err) &&
                  cell_alloc(p, &a.orig, n, err) && cell_alloc(p, &a.scell, n, err) && cell_alloc(p, &a.sxy, n, err) &&
                  cell_alloc(p, &a.szq, n, err) && cell_alloc(p, &a.fxyz, n, err) &&
                  cell_alloc(p, &a.stype, n, err) && cell_alloc(p, &a.bbox_part, 6 * (size_t)a.bbox_blocks, err) &&
                  cell_alloc(p, &p.d_grad, 3 * (size_t)ds.n_pad, err);
        if (!ok) return JME_ERR_NOMEM;
        long long* d_ptr = nullptr;
        unsigned int* d_ent = nullptr;
        if (!cell_alloc(p, &d_ptr, hs.excl_ptr.size(), err) || !cell_alloc(p, &d_ent, hs.excl_ent.size(), err)) return JME_ERR_NOMEM;
        std::vector<long long> ptr(hs.excl_ptr.begin(), hs.excl_ptr.end());
        if (ptr.empty()) ptr.push_back(0);
        if (cudaMemcpy(d_ptr, ptr.data(), ptr.size() * sizeof(long long), cudaMemcpyHostToDevice) != cudaSuccess ||
            (!hs.excl_ent.empty() && cudaMemcpy(d_ent, hs.excl_ent.data(), hs.excl_ent.size() * sizeof(unsigned int), cudaMemcpyHostToDevice) != cudaSuccess)) {
            err = "upload of the exclusion list failed";
            return JME_ERR_CUDA;
        }
        a.excl_ptr = d_ptr; a.excl_ent = d_ent;
        // persistent-style launch: enough warps to fill the chip, grid-stride over atoms
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        p.grid_pairs = std::max(1, std::min((n + kCellWarps - 1) / kCellWarps, sms * 12));
        p.n_epart = p.grid_pairs * kCellWarps;
        if (!cell_alloc(p, &p.d_epart, 2 * (size_t)p.n_epart, err) || !cell_alloc(p, &p.d_cpart, 2 * (size_t)p.n_epart, err)) return JME_ERR_NOMEM;
        p.built = true;
    }
    cudaMemset(p.d_grad, 0, 3 * (size_t)ds.n_pad * sizeof(double));
    cudaMemset(p.d_epart, 0, 2 * (size_t)p.n_epart * sizeof(double));
    cudaMemset(p.d_cpart, 0, 2 * (size_t)p.n_epart * sizeof(unsigned long long));
    p.cutoff = cutoff;
    p.world = world;
    p.i_begin = (int)((long long)n * rank / world);
    p.i_end = (int)((long long)n * (rank + 1) / world);
    p.sorted_valid = false;
    return JME_OK;
}

inline size_t cells_smem_bytes(const DeviceSystem& ds) {
    return (size_t)ds.n_types * ds.n_types * sizeof(PairConst) +
           kCellWarps * (kQueue * sizeof(unsigned int) + 32 * sizeof(int2) + 32 * sizeof(unsigned int));
}

#define JME_CELL_CHECK()                                                                           \
    do {                                                                                           \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess) { err = std::string("cell path launch: ") + cudaGetErrorString(e_); return JME_ERR_CUDA; } \
    } while (0)

// cell assignment + deterministic sort of the current coordinates
inline int run_cell_sort(CellPlan& p, const DeviceSystem& ds, cudaStream_t st, uint64_t& launches, std::string& err) {
    if (p.sorted_valid) return JME_OK;
    CellArrays& a = p.arr;
    const int n = ds.n;
    if (n == 0) { p.sorted_valid = true; return JME_OK; }
    const int nb = (n + 255) / 256;
    const int scan_blocks = (a.max_cells + 1 + kScanBlock - 1) / kScanBlock;
    bbox_kernel<<<a.bbox_blocks, 256, 0, st>>>(ds, a.bbox_part);
    grid_setup_kernel<<<1, 32, 0, st>>>(a.bbox_part, a.bbox_blocks, p.cutoff, a.max_cells, a.grid);
    zero_counts_kernel<<<std::min(592, (a.max_cells + 256) / 256), 256, 0, st>>>(a.grid, a.count, a.cursor);
    cell_assign_kernel<<<nb, 256, 0, st>>>(ds, a.grid, a.cell_of, a.count);
    scan_local_kernel<<<scan_blocks, kScanBlock, 0, st>>>(a.grid, a.count, a.start, a.block_sums);
    scan_sums_kernel<<<1, 1024, 0, st>>>(a.grid, a.block_sums);
    scan_add_kernel<<<scan_blocks, kScanBlock, 0, st>>>(a.grid, a.start, a.block_sums);
    cell_scatter_kernel<<<nb, 256, 0, st>>>(n, a.cell_of, a.start, a.cursor, a.order_tmp);
    cell_rank_gather_kernel<<<nb, 256, 0, st>>>(ds, a);
    launches += 9;
    JME_CELL_CHECK();
    p.sorted_valid = true;
    return JME_OK;
}

inline int run_cell_path(CellPlan& p, const DeviceSystem& ds, bool grad, cudaStream_t st, uint64_t& launches, std::string& err) {
    int rc = run_cell_sort(p, ds, st, launches, err);
    if (rc) return rc;
    if (ds.n == 0) return JME_OK;
    if (grad && p.world > 1) cudaMemsetAsync(p.d_grad, 0, 3 * (size_t)ds.n_pad * sizeof(double), st);
    CellPairArgs P{};
    P.i_begin = p.i_begin; P.i_end = p.i_end;
    P.grad = p.d_grad; P.epart = p.d_epart; P.cpart = p.d_cpart;
    const size_t smem = cells_smem_bytes(ds);
    if (grad) {
        auto k = cells_pair_kernel<true, false>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<p.grid_pairs, kCellThreads, smem, st>>>(ds, p.arr, P);
    } else {
        auto k = cells_pair_kernel<false, false>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<p.grid_pairs, kCellThreads, smem, st>>>(ds, p.arr, P);
    }
    ++launches;
    JME_CELL_CHECK();
    return JME_OK;
}

inline int run_cell_pairlist(CellPlan& p, const DeviceSystem& ds, int* d_i, int* d_j, unsigned char* d_f, long long capacity,
                             unsigned long long* d_cursor, cudaStream_t st, uint64_t& launches, std::string& err) {
    int rc = run_cell_sort(p, ds, st, launches, err);
    if (rc) return rc;
    if (ds.n == 0) return JME_OK;
    CellPairArgs P{};
    P.i_begin = p.i_begin; P.i_end = p.i_end;
    P.grad = p.d_grad; P.epart = p.d_epart; P.cpart = p.d_cpart;
    P.out_i = d_i; P.out_j = d_j; P.out_14 = d_f; P.capacity = capacity; P.cursor = d_cursor;
    const size_t smem = cells_smem_bytes(ds);
    auto k = cells_pair_kernel<false, true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<p.grid_pairs, kCellThreads, smem, st>>>(ds, p.arr, P);
    ++launches;
    JME_CELL_CHECK();
    return JME_OK;
}

}  // namespace jme
