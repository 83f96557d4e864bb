// jme_cells.cuh - cell-list path for cutoff > 0 (the reference has no neighbour list at all: it tests
// every non-excluded pair against the cutoff, MMFF94VdwComponent.java:136-139; this path selects the
// SAME pair set with O(N) work).
//
// Per evaluation, all on the device, no host round trip:
//   1. bounding box (two-stage min/max reduction)              -> grid origin / edge / dims in device memory
//   2. cell index per atom, per-cell counts (integer atomics)   -> exclusive scan -> cell start offsets
//   3. scatter to cells, then rank inside each cell by original atom index, so the sorted order (and
//      with it every later summation order) is deterministic; gather into sorted arrays (one 32-byte
//      {x, y, z, q} record per atom, type, an FP32 copy of the positions relative to the grid origin that also
//      carries the type) and rename the exclusion rows to sorted positions
//   4a. cells_list_kernel (the neighbour-list build, integer / FP32 only): one warp per CELL scans the 25
//      neighbour cell rows once for all atoms of the cell - the row ranges are flattened into one virtual index
//      space so that every trip tests 128 candidates (4 per lane, coalesced FP32 positions) against each atom
//      of the cell with a conservative distance test (padded margins: a superset of the true pair set) and
//      appends the survivors (sorted position + type, 4 bytes) to that atom's survivor list (fixed capacity,
//      streaming stores).
//   4b. cells_pair_kernel (FP64): a warp visits up to 32 consecutive atoms and walks their lists as one stream
//      of 64-entry trips with all lanes busy: exact predicate (bit-identical to the reference's
//      sqrt(dx*dx+dy*dy+dz*dz) > cutoff via the r^2 threshold), exclusion / 1-4 lookup, vdW + Coulomb.  See the
//      comment on the kernel for the software pipeline.  Each pair is evaluated from both of its atoms (no
//      Newton's third law): no scatter, no atomics, deterministic; energies are halved.
#pragma once

#include <cuda_runtime.h>

#include <string>
#include <type_traits>
#include <vector>

#include "jme_host.hpp"
#include "jme_kernels.cuh"

namespace jme {

#ifndef JME_CELL_WARPS
#define JME_CELL_WARPS 8
#endif
#ifndef JME_CELL_MIN_BLOCKS
#define JME_CELL_MIN_BLOCKS 2
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
    // cluster path (jme_clusters.cuh): half shell of neighbour rows and the segment decomposition of the cell rows
    int n_half;              // rows in the half shell: reach * (2 reach + 1) + reach + 1  (13 for reach 2)
    int seg_len;             // cells per segment along x (>= 2 reach)
    int nsx;                 // segments per cell row
    int n_seg;               // ny * nz * nsx
    int seg_lo, seg_hi;      // the segments of THIS shard (grid_setup: equal estimated pair work per rank)
};

// Cells whose atoms a shard of the cluster path touches: the cells of its segments plus the cells its force windows
// reach forward (reach cell layers + reach cell rows + the rest of the last row).  Everything before and after
// needs no sorted record on this rank.  world == 1 or the per-atom-list path (n_seg == 0): all cells.
__device__ __forceinline__ void shard_cell_range(const GridParams& g, int rank, int world, int& c_lo, int& c_hi) {
    c_lo = 0; c_hi = g.ncell;
    if (world <= 1 || g.n_seg <= 0) return;
    (void)rank;
    const long long s0 = g.seg_lo, s1 = g.seg_hi;
    auto first_cell = [&](long long s) { return s >= g.n_seg ? g.ncell : (int)(s / g.nsx) * g.nx + (int)(s % g.nsx) * g.seg_len; };
    c_lo = first_cell(s0);
    const long long last_row = (s1 + g.nsx - 1) / g.nsx;                       // one past the last row with a segment of the shard
    const long long hi_row = last_row + (long long)g.reach * g.ny + g.reach;   // rows its windows can reach
    c_hi = (int)(hi_row * g.nx < (long long)g.ncell ? hi_row * g.nx : (long long)g.ncell);
    if (s1 <= s0) c_hi = c_lo;
}

struct alignas(32) PosQ { double x, y, z, q; };
constexpr int kPadType = 0x80;          // stype of a padding record (cluster path): type 0, flag bit 7

// entries of the survivor lists: sorted position of the partner in the low 24 bits, its type index above
constexpr int kEntryBits = 24;
constexpr unsigned int kEntryMask = (1u << kEntryBits) - 1u;

__device__ __forceinline__ PosQ load_posq(const PosQ* p) {
    PosQ r;     // LDG.E.256: one request, one sector
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.q) : "l"(p));
    return r;
}

struct CellArrays {
    GridParams* grid;
    int n, n_pad, max_cells;
    int pad;                 // every cell is padded to a multiple of `pad` sorted positions (1: no padding)
    int n_spad;              // capacity of the arrays indexed by sorted position: n + (pad - 1) * min(n, max_cells)
    int seg_target;          // wanted number of segments (cluster path), 0 otherwise
    int ring;                // cluster path: slots of a warp's force window (shared by the rows, cl_layout_kernel)
    int* flags;              // [4] device flags: [0] list overflow, [1] non-finite coordinate, [2] window fallback used
    int* cell_of;            // [n] cell index per original atom
    int* count;              // [max_cells + 1]
    int* start;              // [max_cells + 1]
    int* cursor;             // [max_cells]
    int* block_sums;         // scan scratch
    int* order_tmp;          // [n_spad]
    int* orig;               // [n_spad] sorted position -> original atom (-1: padding)
    int* scell;              // [n_spad] cell of sorted position
    PosQ* spq;               // [n_spad] sorted x, y, z, q: one 32-byte record = one sector, one 256-bit load per gather
    float4* fxyz;            // [n_spad] sorted positions relative to the grid origin, FP32; w = type index (int bits)
    int* stype;              // [n_spad] type index by sorted position
    int* sorted_of;          // [n] original atom -> sorted position
    const long long* excl_ptr;    // CSR over original atom index
    const unsigned int* excl_ent; // partners by original index (bit 31 = 1-4 pair)
    unsigned int* sexcl_ent;      // the same rows with the partners as SORTED positions (rebuilt after every sort)
    long long n_excl_ent;
    double* bbox_part;       // [blocks][6]
    int bbox_blocks;
    int rank, world;         // shard of the cluster path (only the cells it touches are sorted)
    int* work_ctr;           // [2] dynamic work distribution: next (cell, part) item / next atom group (or segment);
                             // work_ctr and flags are one 32-byte control block
};

// ---- 1. bounding box ---------------------------------------------------------------------------
__device__ void grid_setup(const double* part, int n_part, double cutoff, int max_cells, int seg_target, int rank, int world,
                           GridParams* g);

// per-block bounding boxes; the last block to finish (counter) reduces them and sets up the grid - one launch less
// in a chain that is launch-latency bound for small systems
__global__ void __launch_bounds__(256) bbox_kernel(DeviceSystem S, double* part, int* __restrict__ flags, int* __restrict__ done_ctr,
                                                   double cutoff, int max_cells, int seg_target, int rank, int world, GridParams* g) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    bool bad = false;
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < S.n; a += gridDim.x * blockDim.x) {
        const double v[3] = {S.x[a], S.y[a], S.z[a]};
        // a non-finite coordinate has no cell: it stays out of the box and poisons the result (NaN total)
        if (!(fabs(v[0]) < 1e300 && fabs(v[1]) < 1e300 && fabs(v[2]) < 1e300)) { bad = true; continue; }
#pragma unroll
        for (int c = 0; c < 3; ++c) { lo[c] = fmin(lo[c], v[c]); hi[c] = fmax(hi[c], v[c]); }
    }
    if (bad) flags[1] = 1;
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
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done_ctr, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return;
    __threadfence();
    grid_setup(part, (int)gridDim.x, cutoff, max_cells, seg_target, rank, world, g);
}

// one warp (the first warp of the last bbox block to finish): lanes stride over the per-block bounding boxes, shuffle
// min / max, lane 0 derives the grid
__device__ void grid_setup(const double* part, int n_part, double cutoff, int max_cells, int seg_target,
                           int rank, int world, GridParams* g) {
    const int lane = threadIdx.x & 31;
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
    for (int c = 0; c < 3; ++c) {
        if (!(lo[c] <= hi[c])) { lo[c] = 0; hi[c] = 0; }          // no finite atom at all
        ext[c] = fmax(hi[c] - lo[c], 0.0);
        extent = fmax(extent, ext[c]);
    }
    // cell edge: half the cutoff plus a hair (5x5x5 stencil; the hair keeps two atoms at exactly r == cutoff within two
    // cells of each other whatever floor() does at the cell boundaries), grown if the box would need more cells than
    // were allocated.  The growth is bounded: 400 steps of x1.25 cover any finite extent.
    double edge = 0.5 * cutoff * (1.0 + 1e-9);
    g->nx = g->ny = g->nz = 1;
    for (int it = 0; it < 400; ++it) {
        const double nx = floor(ext[0] / edge) + 1, ny = floor(ext[1] / edge) + 1, nz = floor(ext[2] / edge) + 1;
        if (nx * ny * nz <= (double)max_cells) {
            g->nx = (int)nx; g->ny = (int)ny; g->nz = (int)nz;
            break;
        }
        edge *= 1.25;
        if (it == 399) edge = fmax(extent, cutoff) * 2.0;          // unreachable for finite input: one cell
    }
    g->ox = lo[0]; g->oy = lo[1]; g->oz = lo[2];
    g->edge = edge;
    g->inv_edge = 1.0 / edge;
    g->ncell = g->nx * g->ny * g->nz;
    g->reach = (int)ceil(cutoff / edge);
    if (g->reach < 1) g->reach = 1;
    if (g->reach > 2) g->reach = 2;                               // edge >= cutoff / 2 by construction
    {   // segments of the cell rows for the cluster path: about seg_target of them, at least 2 reach + 1 cells long
        const int rows = g->ny * g->nz;
        // (a window spans seg_len + 2 reach cells: with seg_len >= 2 reach at most two segments of a row cover a cell,
        // which is what the two slot colours of the j-side partials and cl_cellmask_kernel assume)
        const int lmin = 2 * g->reach;
        int nsx = seg_target > 0 ? (seg_target + rows - 1) / rows : 1;
        int len = g->nx / (nsx > 0 ? nsx : 1);
        if (len < lmin) len = lmin;
        if (len > 12) len = 12;                                   // kClMaxSegLen: per-warp window tables are this long
        g->seg_len = len;
        g->nsx = (g->nx + len - 1) / len;
        g->n_seg = rows * g->nsx;
        g->n_half = g->reach * (2 * g->reach + 1) + g->reach + 1;
        // The shard of this rank: segments [seg_lo, seg_hi).  The pair work of a segment is estimated by the rows of the
        // half shell that exist for its cell row, weighted with their share of the pairs at uniform density (the rows
        // at the far end of the box have no forward neighbours: equal segment COUNTS gave the last rank 15 % less work
        // than the first).  Every rank computes the same boundaries.
        g->seg_lo = (int)((long long)g->n_seg * rank / world);
        g->seg_hi = (int)((long long)g->n_seg * (rank + 1) / world);
        if (world > 1 && g->reach == 2 && g->ny >= 5 && g->nz >= 3) {
            const int wt[13] = {122, 198, 54, 25, 135, 198, 135, 25, 1, 25, 54, 25, 1};    // tools/fill_experiment.py, per mille
            const int ny = g->ny, nz = g->nz;
            auto wrow = [&](int y, int z) {
                int w = 0;
                for (int o = 0; o < 13; ++o) {
                    const int dy = o <= 2 ? o : (o - 3) % 5 - 2, dz = o <= 2 ? 0 : 1 + (o - 3) / 5;
                    if (y + dy >= 0 && y + dy < ny && z + dz < nz) w += wt[o];
                }
                return w;
            };
            // y classes: 0, 1, interior, ny - 2, ny - 1;  z classes: interior, nz - 2, nz - 1
            auto ycls = [&](int y) { return y < 2 ? y : (y >= ny - 2 ? 3 + (y - (ny - 2)) : 2); };
            auto zcls = [&](int z) { return z >= nz - 2 ? 1 + (z - (nz - 2)) : 0; };
            const int yrep[5] = {0, 1, 2, ny - 2, ny - 1}, zrep[3] = {0, nz - 2, nz - 1};
            long long w[3][5], layer[3];
            for (int a = 0; a < 3; ++a) {
                for (int b = 0; b < 5; ++b) w[a][b] = wrow(yrep[b], zrep[a]);
                layer[a] = w[a][0] + w[a][1] + (long long)(ny - 4) * w[a][2] + w[a][3] + w[a][4];
            }
            long long total = 0;
            for (int z = 0; z < nz; ++z) total += layer[zcls(z)];
            total *= g->nx;                                        // weight of a segment = row weight x its cells
            auto boundary = [&](long long target) -> int {         // first segment at which the running weight reaches target
                long long acc = 0;
                for (int z = 0; z < nz; ++z) {
                    const long long lw = layer[zcls(z)] * g->nx;
                    if (acc + lw <= target && z + 1 < nz) { acc += lw; continue; }
                    for (int y = 0; y < ny; ++y) {
                        const long long rw = w[zcls(z)][ycls(y)];
                        if (acc + rw * g->nx <= target && !(z + 1 == nz && y + 1 == ny)) { acc += rw * g->nx; continue; }
                        for (int sx = 0; sx < g->nsx; ++sx) {
                            const int cells = (sx + 1 == g->nsx) ? g->nx - sx * g->seg_len : g->seg_len;
                            if (acc + rw * cells > target) return (z * ny + y) * g->nsx + sx;
                            acc += rw * cells;
                        }
                        return (z * ny + y + 1) * g->nsx;
                    }
                }
                return g->n_seg;
            };
            g->seg_lo = rank == 0 ? 0 : boundary(total * rank / world);
            g->seg_hi = rank + 1 == world ? g->n_seg : boundary(total * (rank + 1) / world);
        }
    }
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
                                                                int* __restrict__ start, int* __restrict__ block_sums, int pad) {
    const int total = g->ncell + 1;
    __shared__ int s[kScanBlock];
    // (a fixed modest grid striding over the chunks: the cell count is known on the device only, and launching a
    // block per chunk of the allocated maximum cost 7 us of empty blocks)
    for (int base = blockIdx.x * kScanBlock; base < total; base += gridDim.x * kScanBlock) {
        const int i = base + threadIdx.x;
        // cells are padded to a multiple of `pad` sorted positions (cluster path): start[] runs over the padded sizes
        const int v = i < total ? (count[i] + pad - 1) / pad * pad : 0;
        __syncthreads();
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < kScanBlock; o <<= 1) {
            const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < total) start[i] = s[threadIdx.x] - v;
        if (threadIdx.x == kScanBlock - 1) block_sums[base / kScanBlock] = s[threadIdx.x];
    }
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
    for (int i = blockIdx.x * kScanBlock + threadIdx.x; i < total; i += gridDim.x * kScanBlock) start[i] += block_sums[i / kScanBlock];
}

// ---- 3. sort -----------------------------------------------------------------------------------
__global__ void cell_scatter_kernel(int n, const int* __restrict__ cell_of, const int* __restrict__ start,
                                    int* __restrict__ cursor, int* __restrict__ order_tmp, const GridParams* __restrict__ g,
                                    int rank, int world) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int c = cell_of[a];
    int c_lo, c_hi;
    shard_cell_range(*g, rank, world, c_lo, c_hi);
    if (c < c_lo || c >= c_hi) return;
    order_tmp[start[c] + atomicAdd(&cursor[c], 1)] = a;
}

// atom a -> its rank among the atoms of its cell by original index (the atomics-ordered scatter only groups the
// atoms by cell), then the gather into the sorted arrays.  With padded cells (cluster path) the last atom of a cell
// also writes the cell's padding records: far away, never selected by the FP32 test (NaN) nor by the exact one
__global__ void cell_rank_gather_kernel(DeviceSystem S, CellArrays C) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= S.n) return;
    const int c = C.cell_of[a];
    {
        int c_lo, c_hi;
        shard_cell_range(*C.grid, C.rank, C.world, c_lo, c_hi);
        if (c < c_lo || c >= c_hi) { C.sorted_of[a] = 0x7fffffff; return; }    // not sorted on this rank: matches nothing
    }
    const int b = C.start[c], m = C.count[c];
    int rank = 0;
    for (int q = b; q < b + m; ++q) rank += C.order_tmp[q] < a;
    const int d = b + rank;
    const GridParams* g = C.grid;
    C.orig[d] = a;
    C.scell[d] = c;
    const double x = S.x[a], y = S.y[a], z = S.z[a];
    PosQ r; r.x = x; r.y = y; r.z = z; r.q = S.q[a];
    C.spq[d] = r;
    const int t = S.type[a];
    C.stype[d] = t;
    C.sorted_of[a] = d;
    C.fxyz[d] = make_float4((float)(x - g->ox), (float)(y - g->oy), (float)(z - g->oz), __int_as_float(t));
    if (C.pad > 1 && rank == m - 1) {
        const int e = C.start[c + 1];
        for (int q = b + m; q < e; ++q) {
            C.orig[q] = -1;
            C.scell[q] = c;
            // padding: never listed as a partner (NaN in the FP32 copy); as a member of a cluster it is switched off
            // by the flag bit of its type, and it sits on a real atom of the cell so that nothing overflows
            PosQ f; f.x = x; f.y = y; f.z = z; f.q = 0.0;
            C.spq[q] = f;
            C.stype[q] = kPadType;
            C.fxyz[q] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, __int_as_float(0));
        }
    }
}

// exclusion / 1-4 rows with the partners renamed to sorted positions: the evaluation kernel then compares list
// entries directly, without a gather of the partner's original index
__global__ void cell_excl_map_kernel(CellArrays C) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= C.n_excl_ent) return;
    const unsigned int x = C.excl_ent[k];
    C.sexcl_ent[k] = (unsigned int)C.sorted_of[x & ~kFlag14] | (x & kFlag14);
}

// ---- 4. pairs ----------------------------------------------------------------------------------
struct CellPairArgs {
    int i_begin, i_end;            // sorted positions handled by this shard
    const unsigned int* list;      // [n][cap] survivor lists (cells_list_kernel)
    const int* count;              // [n]
    int cap;
    int group;                     // atoms a warp takes per visit (1..32)
    double* grad;                  // [3][n_pad] by original atom index
    double* epart;                 // [warps][2]
    unsigned long long* cpart;     // [warps][2]
    // pair-list emission (EMIT only)
    int* out_i; int* out_j; unsigned char* out_14; long long capacity; unsigned long long* cursor;
};

// ---- 4a. survivor lists --------------------------------------------------------------------------
// One warp per CELL.  The 25 neighbour cell rows are scanned once for all atoms of the cell: a trip loads 64
// candidate positions (FP32, coalesced) and tests them against every atom i of the cell (i broadcast from
// shared memory) with the conservative FP32 distance test; survivors are appended to the survivor list of i
// (fixed capacity per atom, L2-resident while hot).  Integer / FP32 only, few registers, high occupancy.
#ifndef JME_LIST_WARPS
#define JME_LIST_WARPS 8
#endif
#ifndef JME_LIST_MIN_BLOCKS
#define JME_LIST_MIN_BLOCKS 4
#endif
#ifndef JME_LIST_K
#define JME_LIST_K 4
#endif
constexpr int kListK = JME_LIST_K;          // candidates per lane per trip (a trip tests 32 * kListK candidates)
static_assert(kListK >= 2 && kListK % 2 == 0, "the packed FP32 test takes the candidates two by two");
constexpr int kListWarps = JME_LIST_WARPS;
constexpr int kListThreads = kListWarps * 32;

struct CellListArgs {
    int i_begin, i_end;            // sorted positions handled by this shard
    unsigned int* list;            // [n][cap] survivor lists by sorted position
    int* count;                    // [n] list lengths (clamped to cap)
    int cap;
    int parts;                     // warps sharing one cell (1 for large systems)
    int* overflow;                 // set to 1 if any list had to be truncated
};

// if (pred) base[index] = v, as a predicated streaming store with a single-instruction address
__device__ __forceinline__ void st_cs_if(bool pred, unsigned long long base, unsigned int index, unsigned int v) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 a;\n\tsetp.ne.u32 p, %0, 0;\n\tmad.wide.u32 a, %1, 4, %2;\n\t"
                 "@p st.global.cs.u32 [a], %3;\n\t}"
                 ::"r"((unsigned int)pred), "r"(index), "l"(base), "r"(v) : "memory");
}

// squared distance of two candidates (a.x / a.y of each argument) from one atom, packed FP32 (sm_100 f32x2)
__device__ __forceinline__ unsigned long long f2_bits(float2 v) {
    return ((unsigned long long)__float_as_uint(v.y) << 32) | __float_as_uint(v.x);
}
__device__ __forceinline__ float2 dist2_x2(unsigned long long xi, unsigned long long yi, unsigned long long zi,
                                           unsigned long long xa, unsigned long long ya, unsigned long long za) {
    unsigned long long dx, dy, dz, d2;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(xi), "l"(xa));
    asm("sub.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(yi), "l"(ya));
    asm("sub.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(zi), "l"(za));
    asm("mul.f32x2 %0, %1, %2;" : "=l"(d2) : "l"(dx), "l"(dx));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d2) : "l"(dy), "l"(dy), "l"(d2));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d2) : "l"(dz), "l"(dz), "l"(d2));
    return make_float2(__uint_as_float((unsigned int)d2), __uint_as_float((unsigned int)(d2 >> 32)));
}

__global__ void __launch_bounds__(kListThreads, JME_LIST_MIN_BLOCKS)
cells_list_kernel(CellArrays C, CellListArgs P) {
    __shared__ float2 s_fi[kListWarps][32][4];  // atom il as (x,x) (y,y) (z,z) -: operands of the packed FP32 test
    __shared__ int s_jb[kListWarps][32];        // first candidate of neighbour row r
    __shared__ int s_off[kListWarps][33];       // exclusive prefix of the row lengths (virtual candidate index)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 (*fis)[4] = s_fi[warp];
    int* jbs = s_jb[warp];
    int* off = s_off[warp];
    const GridParams g = *C.grid;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int width = 2 * g.reach + 1;
    const int n_rows = width * width;          // <= 25: reach <= 2 by construction (edge >= cutoff / 2)
    const float4* __restrict__ fxyz = C.fxyz;

    // work item = (cell, part): a cell's atoms are split among P.parts warps when there are few cells
    // (cells differ in work: the items are handed out dynamically; a list does not depend on who builds it)
    const int n_items = g.ncell * P.parts;
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(C.work_ctr, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_items) break;
        const int c = w / P.parts, part = w % P.parts;
        const int c0 = C.start[c], c1 = C.start[c + 1];
        const int per = (c1 - c0 + P.parts - 1) / P.parts;
        const int a0 = c0 + part * per;
        const int s0 = max(a0, P.i_begin), s1 = min(min(c1, a0 + per), P.i_end);
        if (s0 >= s1) continue;
        const int cx = c % g.nx, cy = (c / g.nx) % g.ny, cz = c / (g.nx * g.ny);
        int total;
        {   // candidate range of neighbour row `lane` (shared by all atoms of the cell), and the prefix sums that
            // flatten the <= 25 ranges into one virtual index space so that every trip has all 64 slots filled
            int jb = 0, len = 0;
            if (lane < n_rows) {
                const int y2 = cy + lane % width - g.reach, z2 = cz + lane / width - g.reach;
                if (y2 >= 0 && y2 < g.ny && z2 >= 0 && z2 < g.nz) {
                    const int x0 = max(cx - g.reach, 0), x1 = min(cx + g.reach, g.nx - 1);
                    const int row = (z2 * g.ny + y2) * g.nx;
                    jb = C.start[row + x0];
                    len = C.start[row + x1 + 1] - jb;
                }
            }
            int incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            total = __shfl_sync(0xffffffffu, incl, 31);
            __syncwarp();
            jbs[lane] = jb;
            off[lane] = incl - len;
            if (lane == 31) off[32] = incl;
        }
        for (int ch = s0; ch < s1; ch += 32) {              // at most 32 atoms of the cell at a time
            const int nc = min(32, s1 - ch);
            __syncwarp();
            {
                const float4 f = fxyz[min(ch + lane, s1 - 1)];
                fis[lane][0] = make_float2(f.x, f.x);
                fis[lane][1] = make_float2(f.y, f.y);
                fis[lane][2] = make_float2(f.z, f.z);
            }
            int my_cnt = 0;                                   // lane il holds the list length of atom ch + il
            __syncwarp();
            unsigned long long rows = reinterpret_cast<unsigned long long>(P.list + (size_t)ch * P.cap);
            asm volatile("" : "+l"(rows));                    // keep the base in registers: addresses below are one IMAD.WIDE
            int r[kListK];                                    // row of this lane's k-th virtual index (monotone)
#pragma unroll
            for (int k = 0; k < kListK; ++k) r[k] = 0;
            for (int vb = 0; vb < total; vb += 32 * kListK) {
                // candidates two by two: the distance test runs on packed FP32 pairs (FADD2 / FMUL2 / FFMA2 of
                // sm_100: per element the same IEEE operations as the scalar sequence mul, fma, fma)
                float ax[kListK], ay[kListK], az[kListK];
                unsigned int en[kListK];
#pragma unroll
                for (int k = 0; k < kListK; ++k) {
                    const int vi = vb + 32 * k + lane;
                    const bool valid = vi < total;
                    while (r[k] < 31 && vi >= off[r[k] + 1]) ++r[k];
                    const int j = valid ? jbs[r[k]] + (vi - off[r[k]]) : 0;
                    const float4 a = fxyz[j];
                    ax[k] = valid ? a.x : 1e18f;              // never passes the distance test
                    ay[k] = a.y; az[k] = a.z;
                    en[k] = (unsigned int)j | ((unsigned int)__float_as_int(a.w) << kEntryBits);
                }
                // packed once per trip, so that the pairs stay in aligned register pairs through the atom loop
                unsigned long long px[kListK / 2], py[kListK / 2], pz[kListK / 2];
#pragma unroll
                for (int k = 0; k < kListK; k += 2) {
                    px[k / 2] = f2_bits(make_float2(ax[k], ax[k + 1]));
                    py[k / 2] = f2_bits(make_float2(ay[k], ay[k + 1]));
                    pz[k / 2] = f2_bits(make_float2(az[k], az[k + 1]));
                    asm volatile("" : "+l"(px[k / 2]), "+l"(py[k / 2]), "+l"(pz[k / 2]));
                }
                for (int il = 0; il < nc; ++il) {
                    const float4 fxy = *reinterpret_cast<const float4*>(&fis[il][0]);   // (x, x, y, y)
                    const float2 fzz = fis[il][2];
                    bool h[kListK];
                    unsigned int m[kListK], any = 0;
#pragma unroll
                    for (int k = 0; k < kListK; k += 2) {
                        const float2 d2 = dist2_x2(f2_bits(make_float2(fxy.x, fxy.y)), f2_bits(make_float2(fxy.z, fxy.w)),
                                                   f2_bits(fzz), px[k / 2], py[k / 2], pz[k / 2]);
                        // the atom itself (distance 0) passes too; the evaluation kernel drops it
                        h[k] = d2.x <= g.r2f_max;
                        h[k + 1] = d2.y <= g.r2f_max;
                        m[k] = __ballot_sync(0xffffffffu, h[k]);
                        m[k + 1] = __ballot_sync(0xffffffffu, h[k + 1]);
                        any |= m[k] | m[k + 1];
                    }
                    if (any == 0) continue;
                    const int cnt = __shfl_sync(0xffffffffu, my_cnt, il);
                    int tot = 0;
#pragma unroll
                    for (int k = 0; k < kListK; ++k) tot += __popc(m[k]);
                    if (cnt + tot <= P.cap) {
                        unsigned int o = (unsigned int)(il * P.cap + cnt);
#pragma unroll
                        for (int k = 0; k < kListK; ++k) {
                            // predicated streaming store: no divergent branch around every store
                            st_cs_if(h[k], rows, o + __popc(m[k] & lt_mask), en[k]);
                            o += __popc(m[k]);
                        }
                        if (lane == il) my_cnt += tot;
                    } else if (lane == 0) {
                        *P.overflow = 1;
                    }
                }
            }
            if (lane < nc) P.count[ch + lane] = my_cnt;
        }
    }
}

// ---- 4b. evaluation --------------------------------------------------------------------------------

constexpr int kExclSmem = 64;     // exclusion entries of the current atom staged per warp

// atoms with more than kExclSmem excluded / 1-4 partners (rare): the rest of the row is read from global
// memory, out of line so that the hot loop holds no load of its own besides the pipelined ones
__device__ __noinline__ int excl_tail(const unsigned int* row, int n_excl, int oj) {   // oj: sorted position
    int hit = 0;                                  // bit 0: excluded, bit 1: 1-4 pair
    for (int e = kExclSmem; e < n_excl; ++e) {
        const unsigned int x = row[e];
        if ((int)(x & ~kFlag14) == oj) hit |= (x & kFlag14) ? 2 : 1;
    }
    return hit;
}

struct alignas(32) GroupAtom { PosQ p; long long eb; int hits, oi, nex, t; int pad[2]; };   // 64 bytes
#ifndef JME_TRIP_K
#define JME_TRIP_K 2
#endif
constexpr int kTripK = JME_TRIP_K;          // list entries per lane per trip (independent FP64 chains per lane)
constexpr int kTripSpan = 32 * kTripK;
constexpr int kChunk = 128;       // list entries per asynchronous copy (one 16-byte piece per lane)
constexpr int kRingSlots = 3;     // chunk being read + two in flight
static_assert(kChunk % kTripSpan == 0, "a trip must not straddle two chunks");

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned int s = (unsigned int)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A warp visits G <= 32 consecutive (cell-sorted) atoms.  Their survivor lists are walked as ONE stream of
// 32-entry trips (the atom changes inside a trip, warp-uniformly), so the software pipeline never drains:
//   * list entries travel global -> shared memory by cp.async, two 128-entry chunks ahead of their use and outside
//     the register scoreboards (a scoreboard wait covers every load outstanding on it, so register loads can only
//     ever be one trip ahead);
//   * the partner records (x, y, z, q; one 256-bit load) are gathered one trip ahead into the other of two register
//     sets (loop unrolled by two: no end-of-trip copies that would drag the wait into the issuing trip), and a trip
//     first consumes its own operands, then issues the next trip's gather, then does the arithmetic;
//   * nothing else in a trip takes a long-latency scoreboard: the partner's type comes with the list entry,
//     exclusions are compared as sorted positions, the next atom's exclusion row is fetched a whole atom ahead.
// Every pair is evaluated from both of its atoms: the gradient of i stays in registers, no scatter, no atomics.
template <bool GRAD, bool EMIT, bool TSM>
__global__ void __launch_bounds__(kCellThreads, JME_CELL_MIN_BLOCKS)
cells_pair_kernel(DeviceSystem S, CellArrays C, CellPairArgs P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PairConst* s_table = reinterpret_cast<PairConst*>(smem_raw);
    const int TT = TSM ? S.n_types * S.n_types : 0;       // a table too large for shared memory is read through the caches
    GroupAtom* s_atoms = reinterpret_cast<GroupAtom*>(s_table + TT);               // 32-byte aligned
    unsigned int* s_ring = reinterpret_cast<unsigned int*>(s_atoms + kCellWarps * 32);
    unsigned int* s_excl = s_ring + kCellWarps * kRingSlots * kChunk;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    GroupAtom* atoms = s_atoms + warp * 32;
    unsigned int* ring = s_ring + warp * kRingSlots * kChunk;
    unsigned int* excl = s_excl + warp * kExclSmem;
    for (int t = threadIdx.x; t < TT; t += kCellThreads) s_table[t] = S.table[t];
    __syncthreads();

    const unsigned int FULL = 0xffffffffu;
    const PosQ* __restrict__ spq = C.spq;

    // groups of G atoms (<= 32; small systems use small groups) are handed out dynamically; everything a group
    // produces (its atoms' gradients, its energy / count partials) is written per group, so the result does not
    // depend on which warp took it
    const int G = P.group;
    for (;;) {
        int grp = 0;
        if (lane == 0) grp = atomicAdd(C.work_ctr + 1, 1);
        grp = __shfl_sync(FULL, grp, 0);
        const int base = P.i_begin + G * grp;
        if (base >= P.i_end) break;
        double ev = 0, ee = 0;                   // per-lane partial sums of this group
        unsigned long long n_cand_w = 0;
        // lane l fetches the per-atom data of atom base + l in one coalesced round and parks it in shared memory
        const int n_here = min(G, P.i_end - base);
        __syncwarp();                            // the previous group's reads of the ring and of atoms[] are over
        bool group_excl;
        {
            const int im = min(base + lane, P.i_end - 1);
            GroupAtom a;
            a.p = spq[im];
            a.hits = lane < n_here ? P.count[im] : 0;
            a.oi = C.orig[im];
            a.eb = C.excl_ptr[a.oi];
            a.nex = (int)(C.excl_ptr[a.oi + 1] - a.eb);
            a.t = C.stype[im];
            atoms[lane] = a;
            group_excl = __any_sync(FULL, lane < n_here && a.nex > 0);
        }
        __syncwarp();

        // chunk requests run over the same (atom, chunk) sequence as the trips, two chunks ahead
        int q_il = 0, q_c = 0, q_slot = 0;
        auto request = [&]() {
            if (q_il < n_here) {
                const unsigned int* src = P.list + ((size_t)(base + q_il) * P.cap + kChunk * q_c + 4 * lane);
                cp_async16(ring + q_slot * kChunk + 4 * lane, src);
                q_slot = q_slot == kRingSlots - 1 ? 0 : q_slot + 1;
                if (kChunk * (q_c + 1) < atoms[q_il].hits) ++q_c; else { ++q_il; q_c = 0; }
            }
            cp_async_commit();                   // an empty group keeps the group count uniform
        };
        request(); request(); request();
        cp_async_wait<2>();                      // chunk 0 has landed (the one exposed latency per group)
        __syncwarp();

        // current position in the stream and the per-atom state that goes with it
        int il = 0, b = 0, slot = 0;
        int n_hits = 0, n_excl = 0, oi = 0, i = base;
        long long eb = 0;
        double pix = 0, piy = 0, piz = 0, qi2 = 0;
        const PairConst* tbase = TSM ? s_table : S.table;
        const PairConst* trow = tbase;
        double gix = 0, giy = 0, giz = 0;
        unsigned int n_eval = 0;
        // exclusion row of the NEXT atom, fetched one atom ahead (first atom: now)
        unsigned int xe0 = 0xffffffffu, xe1 = 0xffffffffu;
        auto fetch_excl = [&](int a) {
            xe0 = xe1 = 0xffffffffu;
            if (a < n_here) {
                const long long e0 = atoms[a].eb;
                const int ne = atoms[a].nex;
                if (lane < ne) xe0 = C.sexcl_ent[e0 + lane];
                if (32 + lane < ne) xe1 = C.sexcl_ent[e0 + 32 + lane];
            }
        };
        if (group_excl) fetch_excl(0);

        struct Nb { unsigned int e[kTripK]; PosQ p[kTripK]; };
        auto gather = [&](Nb& n, int k, unsigned int e) {
            n.e[k] = e;
            unsigned long long addr;              // spq + 32 * index in one IMAD.WIDE.U32
            asm("mad.wide.u32 %0, %1, 32, %2;" : "=l"(addr) : "r"(e & kEntryMask), "l"(spq));
            n.p[k] = load_posq(reinterpret_cast<const PosQ*>(addr));
        };
        Nb A, B = {};
#pragma unroll
        for (int k = 0; k < kTripK; ++k)
            gather(A, k, 32 * k + lane < atoms[0].hits ? ring[32 * k + lane] : (unsigned int)base);

        // the trip body exists twice: groups without any exclusion row (ions, noble gases) run an instance with the
        // exclusion handling compiled out
        auto trip = [&](auto has_excl, const Nb& p, Nb& nx) -> bool {
            constexpr bool kHasExcl = decltype(has_excl)::value;
            if (b == 0) {                        // a new atom starts with this trip (warp-uniform)
                i = base + il;
                const GroupAtom& a = atoms[il];
                n_hits = a.hits; oi = a.oi; eb = a.eb; n_excl = a.nex;
                if (!kHasExcl) n_excl = 0;
                pix = a.p.x; piy = a.p.y; piz = a.p.z;
                qi2 = 2.0 * S.ele_scale * a.p.q;                            // doubled energies, see pair_terms
                trow = tbase + a.t * S.n_types;
                gix = giy = giz = 0;
                if (kHasExcl) {
                    if (n_excl > 0) {
                        __syncwarp();
                        excl[lane] = xe0; excl[32 + lane] = xe1;
                        __syncwarp();
                    }
                    fetch_excl(il + 1);
                }
                n_cand_w += n_hits;
            }
            double dx[kTripK], dy[kTripK], dz[kTripK], r2[kTripK];
#pragma unroll
            for (int k = 0; k < kTripK; ++k) {
                dx[k] = pix - p.p[k].x; dy[k] = piy - p.p[k].y; dz[k] = piz - p.p[k].z;
                // strict double, left-to-right, no FMA: Point3d.distance squared
                r2[k] = __dadd_rn(__dadd_rn(__dmul_rn(dx[k], dx[k]), __dmul_rn(dy[k], dy[k])), __dmul_rn(dz[k], dz[k]));
            }

            // next position of the stream; entering a chunk = its copy has landed, the slot two behind is reused
            const bool last = b + kTripSpan >= n_hits;
            int il2 = il, b2 = b + kTripSpan, slot2 = slot;
            if (last) { il2 = il + 1; b2 = 0; }
            if (last || (b2 & (kChunk - 1)) == 0) {
                slot2 = slot == kRingSlots - 1 ? 0 : slot + 1;
                __syncwarp();                    // every lane is done with the slot about to be overwritten
                request();
                cp_async_wait<2>();
                __syncwarp();
            }
            if (il2 < n_here) {
                const int hits2 = last ? atoms[il2].hits : n_hits;
                const unsigned int* src = ring + slot2 * kChunk + (b2 & (kChunk - 1)) + lane;
#pragma unroll
                for (int k = 0; k < kTripK; ++k) {
                    unsigned int e = b2 + 32 * k + lane < hits2 ? src[32 * k] : (unsigned int)base;
                    // the address carries a (never taken) dependence on r2, so that this load is issued after the
                    // wait for this trip's operands and not before it
                    // (a NaN r2 may carry the sign bit: the mask keeps the index inside the arrays)
                    e = max(e, ((unsigned int)__double2hiint(r2[k]) >> 31) & (e >> 31));
                    gather(nx, k, e);
                }
            }

            // phase 1: exclusion / 1-4 flags of this lane's partners (bit 0: excluded, bit 1: 1-4 pair)
            int xf[kTripK];
#pragma unroll
            for (int k = 0; k < kTripK; ++k) xf[k] = 0;
            if (kHasExcl && n_excl > 0) {          // warp-uniform: the list belongs to atom i
                const int n_sm = min(n_excl, kExclSmem);
                for (int e = 0; e < n_sm; ++e) {
                    const unsigned int x = excl[e];
                    const int xj = (int)(x & ~kFlag14);
#pragma unroll
                    for (int k = 0; k < kTripK; ++k)
                        if (xj == (int)(p.e[k] & kEntryMask)) xf[k] |= (x & kFlag14) ? 2 : 1;
                }
                if (n_excl > kExclSmem) {
#pragma unroll
                    for (int k = 0; k < kTripK; ++k) xf[k] |= excl_tail(C.sexcl_ent + eb, n_excl, (int)(p.e[k] & kEntryMask));
                }
            }
            // phase 2: branch-free arithmetic, kTripK independent chains for the scheduler to interleave.  A pair
            // that does not count (beyond the cutoff, excluded, the atom itself, a padding lane) is evaluated like
            // the others with its two prefactors multiplied by zero: r2 is clamped away from 0, so every
            // intermediate is finite and the products are exact zeros - no selects on the results
#pragma unroll
            for (int k = 0; k < kTripK; ++k) {
                const int jc = (int)(p.e[k] & kEntryMask), tj = (int)(p.e[k] >> kEntryBits);
                // exact predicate: the list holds the atom itself and pairs the FP32 test let through
                const bool on = b + 32 * k + lane < n_hits && jc != i && !(r2[k] > S.r2_max) && !(xf[k] & 1);
                if (on && jc > i) ++n_eval;           // each unordered pair is counted from its lower sorted index
                if (EMIT) {
                    if (on && jc > i) {
                        const int oj = C.orig[jc];
                        const unsigned long long q = atomicAdd(P.cursor, 1ull);
                        if ((long long)q < P.capacity) { P.out_i[q] = min(oi, oj); P.out_j[q] = max(oi, oj); P.out_14[q] = (xf[k] & 2) != 0; }
                    }
                } else {
                    // 0, 1 or 0.75 (1-4 pairs) as the high word of a double whose low word is zero
                    const int hv = on ? 0x3ff00000 : 0;
                    const int he = on ? ((xf[k] & 2) ? 0x3fe80000 : 0x3ff00000) : 0;
                    PairConst pc = trow[tj];
                    pc.k1 *= __hiloint2double(hv, 0);
                    const double qq2 = (qi2 * p.p[k].q) * __hiloint2double(he, 0);
                    double e1, e2v, f;
                    pair_terms<GRAD>(clamp_r2(r2[k]), pc, qq2, e1, e2v, f);
                    ev += e1;
                    ee += e2v;
                    if (GRAD) {
                        gix = fma(f, dx[k], gix);
                        giy = fma(f, dy[k], giy);
                        giz = fma(f, dz[k], giz);
                    }
                }
            }
            if (last && GRAD && !EMIT) {           // the atom is complete
                const double sx = warp_sum(gix), sy = warp_sum(giy), sz = warp_sum(giz);
                if (lane == 0) {
                    P.grad[oi] = sx;
                    P.grad[S.n_pad + oi] = sy;
                    P.grad[2 * (size_t)S.n_pad + oi] = sz;
                }
            }
            il = il2; b = b2; slot = slot2;
            return il < n_here;
        };
        if (group_excl) {
            while (true) {
                if (!trip(std::true_type{}, A, B)) break;
                if (!trip(std::true_type{}, B, A)) break;
            }
        } else {
            while (true) {
                if (!trip(std::false_type{}, A, B)) break;
                if (!trip(std::false_type{}, B, A)) break;
            }
        }
        cp_async_wait<0>();                      // only empty groups can be pending here
        if (!EMIT) {
            // energies were accumulated doubled (pair_terms) and every pair was seen from both atoms: x 1/4
            ev = 0.25 * warp_sum(ev);
            ee = 0.25 * warp_sum(ee);
            for (int o = 16; o; o >>= 1) n_eval += __shfl_xor_sync(FULL, n_eval, o);
            if (lane == 0) {
                P.epart[2 * (size_t)grp] = ev;
                P.epart[2 * (size_t)grp + 1] = ee;
                P.cpart[2 * (size_t)grp] = n_eval;
                P.cpart[2 * (size_t)grp + 1] = n_cand_w;
            }
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------
struct CellPlan {
    CellArrays arr{};
    double* d_grad = nullptr;
    double* d_epart = nullptr;
    unsigned long long* d_cpart = nullptr;
    int n_epart = 0, epart_cap = 0;
    int grid_pairs = 0;
    int i_begin = 0, i_end = 0, world = 1, rank = 0;
    unsigned int* d_list = nullptr;
    int* d_count = nullptr;
    int* d_overflow = nullptr;     // = arr.flags
    int list_cap = 0;
    int grid_list = 0;
    double cutoff = 0;
    bool built = false;
    bool sorted_valid = false;
    // cluster path (jme_clusters.cuh); cluster == 0: the round-1 per-atom survivor lists
    int cluster = 0;
    int n_spad = 0;
    int sms = 148;
    int cl_warps = 0, cl_table_in_smem = 1;
    size_t cl_smem_grad = 0, cl_smem_energy = 0;
    uint2* d_clist = nullptr;      // cluster lists: [cluster_cap][list_cap] 64-bit entries
    int cluster_cap = 0;           // clusters the lists hold (grown with the lists when the sort produces more)
    double* d_gi = nullptr;
    double* d_gj = nullptr;
    double* d_gj_ovf = nullptr;
    int* d_excl_src = nullptr;
    uint2* d_cellmask = nullptr;   // [max_cells] which force slots were written for a cell (cl_cellmask_kernel)
    unsigned int* d_layout = nullptr;   // [max_cells][16] window layout per segment (cl_layout_kernel)
    std::vector<void*> new_allocs;
    std::vector<void*> all_allocs;         // everything the plan allocated (released when the plan changes its kind)

    std::vector<void*> take_new_allocs() { std::vector<void*> v; v.swap(new_allocs); return v; }
    void coords_changed() { sorted_valid = false; }
};

template <class T>
inline bool cell_alloc(CellPlan& p, T** out, size_t count, std::string& err) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) { err = std::string("cudaMalloc (cell path): ") + cudaGetErrorString(e); return false; }
    p.new_allocs.push_back(q);
    p.all_allocs.push_back(q);
    *out = static_cast<T*>(q);
    return true;
}

// `cluster` = 0: round-1 per-atom survivor lists; 2 or 4: cluster path (jme_clusters.cuh), cells padded to that size
inline int build_cell_plan(CellPlan& p, const HostSystem& hs, const DeviceSystem& ds, double cutoff, int rank, int world,
                           int cluster, std::string& err) {
    const int n = ds.n;
    if (cluster == 0 && ((long long)n > (1ll << kEntryBits) || ds.n_types > 256)) {
        err = "cell path (per-atom lists): at most 2^24 atoms and 256 atom types (survivor-list entries are 24 + 8 bits)";
        return JME_ERR_UNSUPPORTED;
    }
    if (cluster != 0 && ds.n_types > 128) {
        err = "cell path: at most 128 atom types (list entries carry 7 type bits)";
        return JME_ERR_UNSUPPORTED;
    }
    if ((long long)n * std::max(1, cluster) > 0x7fffff00ll) {
        err = "cell path: too many atoms for 32-bit sorted positions";
        return JME_ERR_UNSUPPORTED;
    }
    if (!p.built) {
        CellArrays& a = p.arr;
        p.cluster = cluster;
        a.n = n; a.n_pad = ds.n_pad;
        a.max_cells = std::max(4096, 4 * n);
        a.pad = std::max(1, cluster);
        // padded cells: at most (pad - 1) padding records per non-empty cell
        a.n_spad = p.n_spad = n + (a.pad - 1) * std::min(n, a.max_cells);
        a.bbox_blocks = std::max(1, std::min(592, (n + 255) / 256));
        const int scan_blocks = (a.max_cells + 1 + kScanBlock - 1) / kScanBlock;
        const size_t ns = (size_t)p.n_spad + 64;         // spare elements: the gather of cl_pair_kernel, cl_pitch
        bool ok = cell_alloc(p, &a.grid, 1, err) && cell_alloc(p, &a.cell_of, n, err) && cell_alloc(p, &a.count, a.max_cells + 1, err) &&
                  cell_alloc(p, &a.start, a.max_cells + 1, err) && cell_alloc(p, &a.cursor, a.max_cells, err) &&
                  cell_alloc(p, &a.block_sums, scan_blocks, err) && cell_alloc(p, &a.order_tmp, ns, err) &&
                  cell_alloc(p, &a.orig, ns, err) && cell_alloc(p, &a.scell, ns, err) && cell_alloc(p, &a.spq, ns, err) &&
                  cell_alloc(p, &a.fxyz, ns, err) && cell_alloc(p, &a.stype, ns, err) && cell_alloc(p, &a.sorted_of, n, err) &&
                  cell_alloc(p, &a.bbox_part, 6 * (size_t)a.bbox_blocks, err) && cell_alloc(p, &a.work_ctr, 8, err);
        if (!ok) return JME_ERR_NOMEM;
        a.flags = a.work_ctr + 4;
        p.d_overflow = a.flags;
        cudaMemset(a.work_ctr, 0, 8 * sizeof(int));
        cudaMemset(a.spq, 0, ns * sizeof(PosQ));
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
        a.n_excl_ent = (long long)hs.excl_ent.size();
        if (!cell_alloc(p, &a.sexcl_ent, hs.excl_ent.size(), err)) return JME_ERR_NOMEM;
        cudaDeviceGetAttribute(&p.sms, cudaDevAttrMultiProcessorCount, 0);
        p.grid_list = std::max(1, p.sms * JME_LIST_MIN_BLOCKS);
        if (cluster == 0) {
            if (!cell_alloc(p, &p.d_grad, 3 * (size_t)ds.n_pad, err)) return JME_ERR_NOMEM;
            // persistent-style launch: enough warps to fill the chip, grid-stride over atoms
            p.grid_pairs = std::max(1, std::min((n + kCellWarps - 1) / kCellWarps, p.sms * JME_CELL_MIN_BLOCKS));
            p.n_epart = std::max(1, n);           // one energy / count partial per atom group; groups hold >= 1 atom
            if (!cell_alloc(p, &p.d_epart, 2 * (size_t)p.n_epart, err) || !cell_alloc(p, &p.d_cpart, 2 * (size_t)p.n_epart, err)) return JME_ERR_NOMEM;
            p.epart_cap = p.n_epart;
            // survivor lists: fixed capacity per atom (1024 entries; less only if that would exceed 16 GiB)
            p.list_cap = 1024;
            while (p.list_cap > 128 && (size_t)n * p.list_cap * sizeof(unsigned int) > ((size_t)16 << 30)) p.list_cap /= 2;
            if (!cell_alloc(p, &p.d_list, (size_t)std::max(1, n) * p.list_cap, err) || !cell_alloc(p, &p.d_count, n, err)) return JME_ERR_NOMEM;
            a.seg_target = 0;
        } else {
            // one energy / count partial per segment; segments <= cells
            p.n_epart = p.epart_cap = a.max_cells;
            if (!cell_alloc(p, &p.d_epart, 2 * (size_t)p.n_epart, err) || !cell_alloc(p, &p.d_cpart, 2 * (size_t)p.n_epart, err)) return JME_ERR_NOMEM;
            // cluster lists: one per cluster (every cluster holds at least one atom: <= n clusters)
            // about 290 entries per cluster of 4 at liquid density and 10 A; every cluster holds at least one atom, at
            // liquid density about 3.6: room for n / 2 clusters to start with (both grow on overflow, jme_api.cu)
            p.list_cap = 512;
            p.cluster_cap = std::min(std::max(1, n), std::max(4096, n / 2));
            if (!cell_alloc(p, &p.d_clist, (size_t)p.cluster_cap * p.list_cap, err) || !cell_alloc(p, &p.d_count, p.cluster_cap, err) ||
                !cell_alloc(p, &p.d_gi, 3 * ns, err) || !cell_alloc(p, &p.d_gj, (size_t)2 * 13 * 3 * ns, err) ||
                !cell_alloc(p, &p.d_gj_ovf, 3 * ns, err) || !cell_alloc(p, &p.d_excl_src, hs.excl_ent.size(), err) ||
                !cell_alloc(p, &p.d_cellmask, (size_t)a.max_cells, err) || !cell_alloc(p, &p.d_layout, (size_t)16 * a.max_cells, err))
                return JME_ERR_NOMEM;
            cudaMemset(p.d_gj_ovf, 0, 3 * ns * sizeof(double));      // the reduction leaves it zero again after every use
            std::vector<int> src(hs.excl_ent.size());
            for (int64_t i = 0; i < hs.n; ++i)
                for (int64_t e = hs.excl_ptr[i]; e < hs.excl_ptr[i + 1]; ++e) src[(size_t)e] = (int)i;
            if (!src.empty() && cudaMemcpy(p.d_excl_src, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
                err = "upload of the exclusion sources failed";
                return JME_ERR_CUDA;
            }
        }
        p.built = true;
    }
    if (p.d_grad) cudaMemset(p.d_grad, 0, 3 * (size_t)ds.n_pad * sizeof(double));
    cudaMemset(p.d_epart, 0, 2 * (size_t)p.epart_cap * sizeof(double));
    cudaMemset(p.d_cpart, 0, 2 * (size_t)p.epart_cap * sizeof(unsigned long long));
    p.cutoff = cutoff;
    p.world = world;
    p.rank = rank;
    p.arr.rank = cluster != 0 ? rank : 0;        // the per-atom-list path shards by atom index and sorts everything
    p.arr.world = cluster != 0 ? world : 1;
    p.i_begin = (int)((long long)n * rank / world);
    p.i_end = (int)((long long)n * (rank + 1) / world);
    p.sorted_valid = false;
    return JME_OK;
}

inline bool cells_table_in_smem(const DeviceSystem& ds) { return (size_t)ds.n_types * ds.n_types * sizeof(PairConst) <= 48 * 1024; }
inline size_t cells_smem_bytes(const DeviceSystem& ds) {
    return (cells_table_in_smem(ds) ? (size_t)ds.n_types * ds.n_types * sizeof(PairConst) : 0) +
           kCellWarps * (32 * sizeof(GroupAtom) + (kExclSmem + kRingSlots * kChunk) * sizeof(unsigned int));
}

template <bool GRAD, bool EMIT>
inline void launch_cells_pair(CellPlan& p, const DeviceSystem& ds, const CellPairArgs& P, cudaStream_t st) {
    const size_t smem = cells_smem_bytes(ds);
    if (cells_table_in_smem(ds)) {
        auto k = cells_pair_kernel<GRAD, EMIT, true>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<p.grid_pairs, kCellThreads, smem, st>>>(ds, p.arr, P);
    } else {
        auto k = cells_pair_kernel<GRAD, EMIT, false>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<p.grid_pairs, kCellThreads, smem, st>>>(ds, p.arr, P);
    }
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
    cudaMemsetAsync(a.work_ctr, 0, 8 * sizeof(int), st);          // work counters and flags
    bbox_kernel<<<a.bbox_blocks, 256, 0, st>>>(ds, a.bbox_part, a.flags, a.work_ctr + 2, p.cutoff, a.max_cells, a.seg_target, a.rank,
                                               a.world, a.grid);
    zero_counts_kernel<<<std::min(592, (a.max_cells + 256) / 256), 256, 0, st>>>(a.grid, a.count, a.cursor);
    cell_assign_kernel<<<nb, 256, 0, st>>>(ds, a.grid, a.cell_of, a.count);
    const int scan_grid = std::min(scan_blocks, 2 * std::max(1, p.sms));
    scan_local_kernel<<<scan_grid, kScanBlock, 0, st>>>(a.grid, a.count, a.start, a.block_sums, a.pad);
    scan_sums_kernel<<<1, 1024, 0, st>>>(a.grid, a.block_sums);
    scan_add_kernel<<<scan_grid, kScanBlock, 0, st>>>(a.grid, a.start, a.block_sums);
    cell_scatter_kernel<<<nb, 256, 0, st>>>(n, a.cell_of, a.start, a.cursor, a.order_tmp, a.grid, a.rank, a.world);
    cell_rank_gather_kernel<<<nb, 256, 0, st>>>(ds, a);
    launches += 8;
    if (a.n_excl_ent > 0) {
        cell_excl_map_kernel<<<(unsigned int)((a.n_excl_ent + 255) / 256), 256, 0, st>>>(a);
        ++launches;
    }
    JME_CELL_CHECK();
    p.sorted_valid = true;
    return JME_OK;
}

// survivor lists for this shard's atoms (after the sort)
inline int run_cell_lists(CellPlan& p, const DeviceSystem& ds, cudaStream_t st, uint64_t& launches, std::string& err) {
    if (ds.n == 0) return JME_OK;
    CellListArgs A{};
    A.i_begin = p.i_begin; A.i_end = p.i_end;
    A.list = p.d_list; A.count = p.d_count; A.cap = p.list_cap; A.overflow = p.d_overflow;
    // about 12 atoms per cell at liquid density: split cells among warps when that leaves resident warps idle
    const long long resident = (long long)p.grid_list * kListWarps;
    A.parts = (int)std::max(1ll, std::min(8ll, 2 * resident * 12 / std::max(1, ds.n)));
    cudaMemsetAsync(p.arr.work_ctr, 0, sizeof(int), st);
    cells_list_kernel<<<p.grid_list, kListThreads, 0, st>>>(p.arr, A);
    ++launches;
    JME_CELL_CHECK();
    return JME_OK;
}

inline void fill_pair_args(CellPlan& p, CellPairArgs& P) {
    P.i_begin = p.i_begin; P.i_end = p.i_end;
    P.list = p.d_list; P.count = p.d_count; P.cap = p.list_cap;
    // enough warp visits to keep every resident warp busy a few times over
    const long long resident = (long long)p.grid_pairs * kCellWarps;
#ifndef JME_GROUP_MAX
#define JME_GROUP_MAX 32
#endif
    P.group = (int)std::max(1ll, std::min((long long)JME_GROUP_MAX, (long long)(p.i_end - p.i_begin) / (resident * 2)));
    P.grad = p.d_grad; P.epart = p.d_epart; P.cpart = p.d_cpart;
    p.n_epart = std::max(1, (p.i_end - p.i_begin + P.group - 1) / P.group);    // partials written: one per group
}

inline int run_cell_path(CellPlan& p, const DeviceSystem& ds, bool grad, cudaStream_t st, uint64_t& launches, std::string& err) {
    if (ds.n == 0) return JME_OK;
    if (grad && p.world > 1) cudaMemsetAsync(p.d_grad, 0, 3 * (size_t)ds.n_pad * sizeof(double), st);
    CellPairArgs P{};
    fill_pair_args(p, P);
    cudaMemsetAsync(p.arr.work_ctr + 1, 0, sizeof(int), st);
    if (grad) launch_cells_pair<true, false>(p, ds, P, st);
    else launch_cells_pair<false, false>(p, ds, P, st);
    ++launches;
    JME_CELL_CHECK();
    return JME_OK;
}

inline int run_cell_pairlist(CellPlan& p, const DeviceSystem& ds, int* d_i, int* d_j, unsigned char* d_f, long long capacity,
                             unsigned long long* d_cursor, cudaStream_t st, uint64_t& launches, std::string& err) {
    int rc = run_cell_sort(p, ds, st, launches, err);
    if (rc) return rc;
    if ((rc = run_cell_lists(p, ds, st, launches, err))) return rc;
    if (ds.n == 0) return JME_OK;
    CellPairArgs P{};
    fill_pair_args(p, P);
    P.out_i = d_i; P.out_j = d_j; P.out_14 = d_f; P.capacity = capacity; P.cursor = d_cursor;
    cudaMemsetAsync(p.arr.work_ctr + 1, 0, sizeof(int), st);
    launch_cells_pair<false, true>(p, ds, P, st);
    ++launches;
    JME_CELL_CHECK();
    return JME_OK;
}

// which list storage overflowed since the flags were last cleared (results are then incomplete):
// bit 0 = the entries of a list, bit 1 = the number of clusters
inline int cell_overflowed(CellPlan& p) {
    int v[4] = {0, 0, 0, 0};
    if (p.arr.flags) cudaMemcpy(v, p.arr.flags, sizeof(v), cudaMemcpyDeviceToHost);
    return (v[0] ? 1 : 0) | (v[3] ? 2 : 0);
}

}  // namespace jme
