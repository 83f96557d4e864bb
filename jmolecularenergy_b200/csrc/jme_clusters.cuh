// jme_clusters.cuh - the cutoff path: cluster lists, Newton's third law, sliding force windows.
//
// The reference tests every non-excluded pair against the cutoff (MMFF94VdwComponent.java:136-139,
// MMFF94ElectrostaticComponent.java:139-141); this path selects the SAME pair set with O(N) work and evaluates
// every unordered pair exactly ONCE.  It replaces the round-1 per-atom survivor lists (each pair evaluated from
// both of its atoms), see DESIGN.md section 4.
//
//   sort        (jme_cells.cuh) cell edge = cutoff / 2, atoms sorted by cell (z, y, x), rank inside a cell by original
//               index; every cell is padded to a multiple of C sorted positions, so CLUSTERS of C consecutive
//               positions never straddle a cell.  Padding records are never listed and are switched off as members.
//   cl_list     FP32, conservative: per cluster the list of partner ATOMS j later in the sorted order (the half
//               shell: the rest of the own cell row, then the rows dy > 0 and dz > 0) that are inside the padded
//               cutoff of at least one atom of the cluster.  One 64-bit entry per (cluster, j), everything the FP64
//               kernel needs precomputed: { sorted position of j | slot of j in the force window, type of j, C
//               "pair excluded" bits, C "1-4 pair" bits }.
//   cl_mark     one thread per exclusion entry: finds the pair's list entry (binary search, entries are sorted by
//               position) and sets the excluded / 1-4 bit of the atom, so the FP64 kernel needs no exclusion lookup.
//   cl_pair     FP64.  A warp sweeps a SEGMENT (<= 12 consecutive cells of one cell row) cluster by cluster; a lane
//               takes 4 / C list entries per trip and evaluates them against the C atoms of the cluster = 4
//               independent FP64 chains per lane (what the FP64 pipe needs).  The exact
//               predicate is the reference's (strict double r^2 against the threshold that is bit-equivalent to
//               sqrt(..) > cutoff).  i-side forces stay in registers for the whole list and are reduced across the
//               warp once per cluster; j-side forces are accumulated in a per-warp SHARED-MEMORY WINDOW that holds
//               one FP64 (x, y, z) accumulator for every atom of the 13 neighbour rows within reach of the current
//               cell and slides along x with the sweep: a cell that leaves the window is flushed to its slot array
//               (coalesced), so a j-atom receives exactly one partial per (row of the half shell, segment) - no
//               atomics, every partial has one writer and the final sum a fixed order => bit-reproducible.
//               Entries and partner records travel through registers: entries two trips ahead (their cache lines
//               are prefetched to L2 a whole cluster ahead), records one trip ahead.  Lists are padded to whole trips
//               and the trips of a batch of clusters are numbered through, so the trip loop is one basic block.
//   cl_reduce   per cell: which of the <= 26 j-side slots were written for it follows from the grid (nothing is
//               stored); per atom: i-side sum + those slots + bonded slots; unit conversion; scatter to the
//               caller's atom order.
#pragma once

#include "jme_cells.cuh"

namespace jme {

constexpr int kClMaxRows = 13;        // rows of the half shell at reach 2
constexpr int kClMaxSegLen = 12;      // cells per segment (GridParams::seg_len is clamped to this in grid_setup)
constexpr int kClRowPitch = 16;       // row-table entries per cell
constexpr int kClBatch = 32;          // clusters staged per batch
#ifndef JME_CL_WARPS
#define JME_CL_WARPS 8
#endif
// warps per CTA, one CTA per SM, 230 registers per thread.  A third warp on any scheduler caps the registers at 168
// (16 K registers per scheduler): measured on 10^6 ions 9 / 10 warps 1.91 / 1.81 ms against 1.68 ms with 8, and the
// smaller windows of 10 warps overflow into the atomics
constexpr int kClWarps = JME_CL_WARPS;
constexpr int kClMaxRing = 128;       // upper bound of the window slots per neighbour row (11 slot bits: 13 * 128 + 1)

// 64-bit list entry: .x = sorted position of the partner, .y = meta
constexpr int kMetaSlotBits = 11;                               // window slot: row * ring + ring index; 13 * ring = dummy slot
constexpr unsigned int kMetaSlotMask = (1u << kMetaSlotBits) - 1u;
constexpr int kMetaTypeShift = 11;                              // 7 bits
constexpr unsigned int kMetaOvf = 1u << 18;                     // the partner does not fit the ring of its row here: atomics
constexpr int kMetaExclShift = 24;                              // C bits: pair (atom a of the cluster, partner) does not count
constexpr int kMetaF14Shift = 28;                               // C bits: ... is a 1-4 pair

template <int C>
struct ClFmt {
    static_assert(C == 2 || C == 4, "cluster size");
    static constexpr int kK = 4 / C;                               // list entries per lane per trip: C * kK = 4 chains
    static constexpr int kSpan = 32 * kK;                          // entries per trip
    static constexpr unsigned int kAllExcl = ((1u << C) - 1u) << kMetaExclShift;
};

// half-shell row o -> (dy, dz):  o = 0 own row; 1..reach: dz = 0, dy = o; then dz = 1..reach, dy = -reach..reach
__device__ __forceinline__ void cl_row_offset(int o, int reach, int& dy, int& dz) {
    if (o <= reach) { dy = o; dz = 0; return; }
    const int W = 2 * reach + 1, t = o - reach - 1;
    dz = 1 + t / W;
    dy = t % W - reach;
}
__device__ __forceinline__ int cl_row_index(int dy, int dz, int reach) {       // -1: not in the half shell
    if (dz == 0) return (dy >= 0 && dy <= reach) ? dy : -1;
    if (dz < 0 || dz > reach || dy < -reach || dy > reach) return -1;
    return reach + 1 + (dz - 1) * (2 * reach + 1) + (dy + reach);
}
// cells of the segments [s0, s1) of this shard: contiguous because segments tile the cell rows in order
__device__ __forceinline__ int cl_seg_first_cell(const GridParams& g, long long s) {
    return s >= g.n_seg ? g.ncell : (int)(s / g.nsx) * g.nx + (int)(s % g.nsx) * g.seg_len;
}
// pitch (in positions) of the per-position force arrays: the number of sorted positions of THIS evaluation (known on
// the device only), not the worst case they are allocated for - the slot arrays then lie a few MB apart instead of
// 16 large pages apart, which spreads them over the TLB sets
__device__ __forceinline__ size_t cl_pitch(int n_sorted) { return ((size_t)n_sorted + 63) / 32 * 32; }
// t mod ring for 0 <= t < 2^31 without a division: magic = floor(2^32 / ring) underestimates the quotient by at most 1
__device__ __forceinline__ int cl_ring_mod(int t, int ring, unsigned int magic) {
    int r = t - (int)__umulhi((unsigned int)t, magic) * ring;
    return r >= ring ? r - ring : r;
}

// ---- force-window layout ------------------------------------------------------------------------------
// The window of a sweeping warp is one pool of `pool` FP64 slots (x 3 planes) shared by the <= 13 rows of the half
// shell.  Row o of segment s gets a ring of cap(s, o) slots = the largest number of atoms its (2 reach + 1)-cell
// window ever holds while the segment is swept, so no partner ever falls outside its ring; only if the rows together
// need more than the pool are the rings shortened in proportion (partners beyond a ring then go through atomics).
// One warp per segment; layout[seg * 16 + o] = base | cap << 16.
__global__ void __launch_bounds__(256)
cl_layout_kernel(CellArrays A, int pool, unsigned int* __restrict__ layout) {
    const GridParams& g = *A.grid;
    const int lane = threadIdx.x & 31;
    const int n_warps = gridDim.x * (blockDim.x >> 5);
    for (int seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); seg < g.n_seg; seg += n_warps) {
        const int row = seg / g.nsx, sx = seg % g.nsx;
        const int xa = sx * g.seg_len, xb = min(xa + g.seg_len, g.nx);
        const int cy = row % g.ny, cz = row / g.ny;
        int need = 0;
        if (lane < g.n_half) {
            int dy, dz;
            cl_row_offset(lane, g.reach, dy, dz);
            const int y2 = cy + dy, z2 = cz + dz;
            if (y2 >= 0 && y2 < g.ny && z2 < g.nz) {
                const int rb = (z2 * g.ny + y2) * g.nx;
                for (int x = xa; x < xb; ++x)
                    need = max(need, A.start[rb + min(x + g.reach, g.nx - 1) + 1] - A.start[rb + max(x - g.reach, 0)]);
            }
        }
        need = min(need, 0xffff);
        int total = need;
#pragma unroll
        for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        int cap = need;
        if (total > pool) cap = (int)((long long)need * pool / total);     // (a row that needs slots keeps at least ... 0: all atomics)
        int incl = cap;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane < 16) layout[seg * 16 + lane] = (unsigned int)(incl - cap) | ((unsigned int)cap << 16);
    }
}

// ---- cluster lists -----------------------------------------------------------------------------------
struct ClListArgs {
    uint2* list;                   // [clusters][cap]
    int* count;                    // [clusters]
    int cap;
    int ccap;                      // clusters the list storage holds (more: overflow flag, the host grows and retries)
    int parts;                     // warps sharing one cell (small systems)
    int rank, world;
    const unsigned int* layout;    // [segments][16] ring of every half-shell row in the sweeping warp's window (cl_layout_kernel)
    int pool;                      // the dummy window slot (padding entries)
};

// if (pred) base[index] = v (8 bytes), as a predicated streaming store with a single-instruction address
__device__ __forceinline__ void st_cs_if64(bool pred, unsigned long long base, unsigned int index, uint2 v) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 a;\n\tsetp.ne.u32 p, %0, 0;\n\tmad.wide.u32 a, %1, 8, %2;\n\t"
                 "@p st.global.cs.v2.u32 [a], {%3, %4};\n\t}"
                 ::"r"((unsigned int)pred), "r"(index), "l"(base), "r"(v.x), "r"(v.y) : "memory");
}

constexpr int kClListMinBlocks = 4;   // measured: 2 / 3 / 4 resident CTAs (102 / 80 / 64 registers) build the 10^6-ion lists in 0.76 / 0.69 / 0.68 ms, 5 (48 registers, spills) in 0.78

// shared-memory scratch of one list-building warp (1 604 bytes)
struct ClBuildSmem {
    float2 (*fis)[4];    // [32][4] atom il as (x,x) (y,y) (z,z) -: operands of the packed FP32 test
    int4* rings;         // [16] ring of row o: {first slot, slots, floor(2^32 / slots), -}
    int4* rowa;          // [16] row o: {first candidate - its virtual index, ring origin (first position of the row's window
                         //      for the SEGMENT), first position of its window for this cell, -}
    int* off;            // [17] exclusive prefix of the candidate counts (virtual candidate index)
};
constexpr size_t kClBuildSmemBytes = 32 * 4 * sizeof(float2) + 2 * 16 * sizeof(int4) + 17 * sizeof(int);
__device__ __forceinline__ ClBuildSmem cl_build_smem(unsigned char* base) {      // base: 16-byte aligned
    ClBuildSmem S;
    S.fis = reinterpret_cast<float2 (*)[4]>(base);
    S.rings = reinterpret_cast<int4*>(base + 1024);
    S.rowa = reinterpret_cast<int4*>(base + 1280);
    S.off = reinterpret_cast<int*>(base + 1536);
    return S;
}

// The lists of the clusters [k0, k1) of cell c, by one warp (called by cl_list_kernel, one launch for all cells, or by the
// sweeping warp of cl_pair_kernel for the cells of its segment).
template <int C>
__device__ __forceinline__ void cl_build_cell(const CellArrays& A, const GridParams& g, const ClListArgs& P, int c, int k0, int k1,
                                              int lane, const ClBuildSmem& SM) {
    using F = ClFmt<C>;
    float2 (*fis)[4] = SM.fis;
    int4* rowa = SM.rowa;
    int4* rings = SM.rings;
    int* off = SM.off;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int H = g.n_half;
    const float4* __restrict__ fxyz = A.fxyz;
    const float qnan = __int_as_float(0x7fc00000);
    const int c0 = A.start[c];
    const int cx = c % g.nx, cy = (c / g.nx) % g.ny, cz = c / (g.nx * g.ny);
    const int xa = cx / g.seg_len * g.seg_len;                   // first cell of the segment that sweeps this cell
    const int seg = (cz * g.ny + cy) * g.nsx + cx / g.seg_len;
    const int a_begin = c0 + C * k0, a_end = c0 + C * k1;        // sorted positions of this item's clusters
    int total;
    {   // candidate range of half-shell row `lane` and the prefix sums that flatten the <= 13 ranges into one
        // virtual index space, so that every trip has all its slots filled
        int jb = 0, len = 0, cw = 0, p0 = 0;
        if (lane < H) {
            int dy, dz;
            cl_row_offset(lane, g.reach, dy, dz);
            const int y2 = cy + dy, z2 = cz + dz;
            if (y2 >= 0 && y2 < g.ny && z2 < g.nz) {
                const int x0 = max(cx - g.reach, 0), x1 = min(cx + g.reach, g.nx - 1);
                const int row = (z2 * g.ny + y2) * g.nx;
                cw = A.start[row + x0];
                p0 = A.start[row + max(xa - g.reach, 0)];
                // own row: only partners later in the sorted order than the first atom of the first cluster
                jb = lane == 0 ? a_begin + 1 : cw;
                len = A.start[row + x1 + 1] - jb;
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
        if (lane < 16) {
            rowa[lane] = make_int4(jb - (incl - len), p0, cw, 0); off[lane] = incl - len;
            const unsigned int lay = P.layout[seg * 16 + lane];
            const int cap = (int)(lay >> 16);
            rings[lane] = make_int4((int)(lay & 0xffffu), cap, cap > 0 ? (int)(0x100000000ull / (unsigned long long)cap) : 0, 0);
        }
        if (lane == 16) off[16] = incl - len;                     // = total for lane >= H (len == 0)
    }
    for (int ch = a_begin; ch < a_end; ch += 32) {              // at most 32 atoms (32 / C clusters) at a time
        const int nc = min(32, a_end - ch);                       // multiple of C
        const int nq = nc / C;
        if (ch / C + nq > P.ccap) {                               // more clusters than list storage
            if (lane == 0) A.flags[3] = 1;
            break;
        }
        __syncwarp();
        {
            const float4 f = fxyz[min(ch + lane, a_end - 1)];
            fis[lane][0] = make_float2(f.x, f.x);
            fis[lane][1] = make_float2(f.y, f.y);
            fis[lane][2] = make_float2(f.z, f.z);
        }
        int my_cnt = 0;                                           // lane kq holds the list length of cluster ch / C + kq
        bool beyond = false;                                      // a candidate of this lane lies beyond the ring of its row
        __syncwarp();
        unsigned long long rows = reinterpret_cast<unsigned long long>(P.list + (size_t)(ch / C) * P.cap);
        asm volatile("" : "+l"(rows));                            // keep the base in registers
        int r[kListK], nxt[kListK];                               // row of this lane's k-th virtual index (monotone) and where
#pragma unroll                                                    // the next row starts (kept in a register: no table read
        for (int k = 0; k < kListK; ++k) { r[k] = 0; nxt[k] = off[1]; }   // per trip when no boundary is crossed)
        for (int vb = 0; vb < total; vb += 32 * kListK) {
            float ax[kListK], ay[kListK], az[kListK];
            unsigned int meta[kListK];
            int jj[kListK];
#pragma unroll
            for (int k = 0; k < kListK; ++k) {
                const int vi = vb + 32 * k + lane;
                const bool valid = vi < total;
                while (r[k] < 15 && vi >= nxt[k]) { ++r[k]; nxt[k] = off[r[k] + 1]; }
                const int4 ra = rowa[r[k]];
                const int j = valid ? ra.x + vi : 0;
                const float4 a = fxyz[j];
                ax[k] = valid ? a.x : qnan;                       // NaN never passes the distance test
                ay[k] = a.y; az[k] = a.z;
                jj[k] = j;
                // the partner's slot in the force window of the sweeping warp: ring index counted from the ring
                // origin of its row; a partner further from the window start than the ring is long is flagged
                const int4 rg = rings[r[k]];
                const int slot = rg.x + (rg.y > 0 ? cl_ring_mod(valid ? j - ra.y : 0, rg.y, (unsigned int)rg.z) : 0);
                const bool ovf = valid && j - ra.z >= rg.y;
                beyond = beyond || ovf;
                meta[k] = (unsigned int)slot | ((unsigned int)__float_as_int(a.w) << kMetaTypeShift) | (ovf ? kMetaOvf : 0u);
            }
            unsigned long long px[kListK / 2], py[kListK / 2], pz[kListK / 2];
#pragma unroll
            for (int k = 0; k < kListK; k += 2) {
                px[k / 2] = f2_bits(make_float2(ax[k], ax[k + 1]));
                py[k / 2] = f2_bits(make_float2(ay[k], ay[k + 1]));
                pz[k / 2] = f2_bits(make_float2(az[k], az[k + 1]));
                asm volatile("" : "+l"(px[k / 2]), "+l"(py[k / 2]), "+l"(pz[k / 2]));
            }
            for (int kq = 0; kq < nq; ++kq) {
                const int p0 = ch + C * kq;                       // first position of the cluster
                bool h[kListK];
#pragma unroll
                for (int k = 0; k < kListK; ++k) h[k] = false;
#pragma unroll
                for (int a = 0; a < C; ++a) {
                    const float4 fxy = *reinterpret_cast<const float4*>(&fis[C * kq + a][0]);   // (x, x, y, y)
                    const float2 fzz = fis[C * kq + a][2];
#pragma unroll
                    for (int k = 0; k < kListK; k += 2) {
                        const float2 d2 = dist2_x2(f2_bits(make_float2(fxy.x, fxy.y)), f2_bits(make_float2(fxy.z, fxy.w)),
                                                   f2_bits(fzz), px[k / 2], py[k / 2], pz[k / 2]);
                        h[k] = h[k] || d2.x <= g.r2f_max;
                        h[k + 1] = h[k + 1] || d2.y <= g.r2f_max;
                    }
                }
                unsigned int m[kListK], any = 0;
#pragma unroll
                for (int k = 0; k < kListK; ++k) {
                    h[k] = h[k] && jj[k] > p0;                    // half shell: partners later than the cluster's first atom
                    m[k] = __ballot_sync(0xffffffffu, h[k]);
                    any |= m[k];
                }
                if (any == 0) continue;
                const int cnt = __shfl_sync(0xffffffffu, my_cnt, kq);
                int tot = 0;
#pragma unroll
                for (int k = 0; k < kListK; ++k) tot += __popc(m[k]);
                if (cnt + tot <= P.cap) {
                    unsigned int o = (unsigned int)(kq * P.cap + cnt);
#pragma unroll
                    for (int k = 0; k < kListK; ++k) {
                        // a partner inside the own cluster: the atoms a >= j - p0 do not come before it in the sorted
                        // order - their pair bits are set as "excluded", the evaluation needs no order test
                        const int dj = jj[k] - p0;
                        const unsigned int own = dj < C ? ((((1u << C) - 1u) >> dj) << dj) << kMetaExclShift : 0u;
                        st_cs_if64(h[k], rows, o + __popc(m[k] & lt_mask), make_uint2((unsigned int)jj[k], meta[k] | own));
                        o += __popc(m[k]);
                    }
                    if (lane == kq) my_cnt += tot;
                } else if (lane == 0) {
                    A.flags[0] = 1;
                }
            }
        }
        // every list is padded to whole trips (at least one) with dummy entries - the cluster's own first atom,
        // every pair excluded, the dummy window slot - so that the FP64 kernel loads entries unconditionally; bit 30
        // of the count: some candidate of this cell lies beyond its ring, the clusters' trips need the atomics
        const bool cell_ovf = __any_sync(0xffffffffu, beyond);
        for (int kq = 0; kq < nq; ++kq) {
            const int cnt = __shfl_sync(0xffffffffu, my_cnt, kq);
            const int full = min(P.cap, max(F::kSpan, (cnt + F::kSpan - 1) / F::kSpan * F::kSpan));
            for (int i = cnt + lane; i < full; i += 32)
                st_cs_if64(true, rows, (unsigned int)(kq * P.cap + i),
                           make_uint2((unsigned int)(ch + C * kq), (unsigned int)P.pool | F::kAllExcl));
        }
        if (lane < nq) P.count[ch / C + lane] = my_cnt | (cell_ovf ? 0x40000000 : 0);
    }
}

template <int C>
__global__ void __launch_bounds__(kListThreads, kClListMinBlocks)
cl_list_kernel(CellArrays A, ClListArgs P) {
    __shared__ __align__(16) unsigned char s_build[kListWarps][(kClBuildSmemBytes + 15) / 16 * 16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ClBuildSmem SM = cl_build_smem(s_build[warp]);
    const GridParams g = *A.grid;
    const long long s0 = g.seg_lo, s1 = g.seg_hi;
    const int cell0 = cl_seg_first_cell(g, s0), cell1 = cl_seg_first_cell(g, s1);
    const int n_items = (cell1 - cell0) * P.parts;
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(A.work_ctr, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_items) break;
        const int c = cell0 + w / P.parts, part = w % P.parts;
        const int nclu = (A.start[c + 1] - A.start[c]) / C;
        const int per = (nclu + P.parts - 1) / P.parts;
        const int k0 = part * per, k1 = min(nclu, k0 + per);
        if (k0 >= k1) continue;
        cl_build_cell<C>(A, g, P, c, k0, k1, lane, SM);
    }
}

// ---- exclusion / 1-4 bits ------------------------------------------------------------------------------
// One thread per (directed) exclusion entry; the direction whose source atom comes first in the sorted order owns
// the pair.  The partner's entry in the source cluster's list is found by binary search on the position - the list
// kernel emits entries in that order - and the atom's excluded / 1-4 bit is OR-ed in.  Pairs beyond the cutoff have
// no entry: nothing to mark.
template <int C>
__global__ void __launch_bounds__(256)
cl_mark_kernel(CellArrays A, const int* __restrict__ excl_src, uint2* __restrict__ list, const int* __restrict__ count,
               int cap, int ccap, int rank, int world) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.n_excl_ent) return;
    const unsigned int x = A.sexcl_ent[k];
    const unsigned int sj = x & ~kFlag14;
    const int si = A.sorted_of[excl_src[k]];
    if ((unsigned int)si >= sj) return;
    const GridParams& g = *A.grid;
    {
        const int ci = A.scell[si];
        const long long s0 = g.seg_lo, s1 = g.seg_hi;
        if (ci < cl_seg_first_cell(g, s0) || ci >= cl_seg_first_cell(g, s1)) return;      // not this shard's list
    }
    const int q = si / C, a = si % C;
    if (q >= ccap) return;
    uint2* lst = list + (size_t)q * cap;
    const int n = min(count[q] & 0x3fffffff, cap);
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (lst[mid].x < sj) lo = mid + 1; else hi = mid;
    }
    if (lo < n && lst[lo].x == sj)
        atomicOr(&lst[lo].y, 1u << (((x & kFlag14) ? kMetaF14Shift : kMetaExclShift) + a));
}

// ---- evaluation ----------------------------------------------------------------------------------------
struct ClPairArgs {
    const uint2* list;             // [clusters][cap]
    const int* count;              // [clusters]
    int cap;
    int ccap;                      // clusters the list storage holds
    int rank, world;
    int pool;                      // slots of a warp's force window (cl_layout_kernel shares them among the rows)
    const unsigned int* layout;    // [segments][16]
    int n_spad;                    // (capacity of the per-position arrays below)
    double* gi;                    // [position][3] i-side sums by sorted position
    double* gj;                    // [2 * kClMaxRows][cl_pitch][3] j-side window flushes: slot 2 * row + (segment & 1)
    double* gj_ovf;                // [position][3] atomically accumulated contributions of atoms beyond a window's capacity
    double* epart;                 // [n_seg][2] per-segment energies (vdw, ele)
    unsigned long long* cpart;     // [n_seg][2] pairs evaluated, candidate pairs
    // pair-list emission (EMIT only)
    int* out_i; int* out_j; unsigned char* out_14; long long capacity; unsigned long long* cursor;
};

// sum of v[i] over the 32 lanes for N values at the cost of about N + log2(32 / N) shuffles instead of 5 N: in every
// halving step a lane keeps one half of its values and hands the other half to its partner.  On return v[0] of lane
// l holds the warp total of value  l / (32 / N).
template <int N>
__device__ __forceinline__ void warp_multi_sum(double (&v)[N], int lane) {
    static_assert(N == 1 || N == 2 || N == 4 || N == 8 || N == 16 || N == 32, "power of two");
    int m = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1, m >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const double send = up ? v[i] : v[i + n / 2];
            const double keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
    }
    for (; m > 0; m >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], m);
}

// per-warp shared memory: the force window (3 planes of pool + 8 doubles: the extra ones hold the dummy slot of the
// padding entries), the row table (window start of every half-shell row for every cell of the segment + the final
// frontier), the row constants, the batch's cluster data, the trip table and the parked cluster atoms
__host__ __device__ constexpr size_t cl_window_plane(int pool) { return (size_t)pool + 8; }
constexpr int kClNextBytes = 2 * (4 * 32 + 16);   // the current and the next cluster's atoms (C records + C types), ping-pong
constexpr int kClTripTab = 254;                   // trips of a batch; two more descriptors pad the table (the look-ahead of the pipeline)
__host__ __device__ constexpr size_t cl_warp_smem(bool grad, int pool) {
    return (grad ? 3 * cl_window_plane(pool) * sizeof(double) : 0) + (kClMaxSegLen + 1) * kClRowPitch * sizeof(int) +
           kClRowPitch * sizeof(int4) + 2 * kClBatch * sizeof(int) + kClNextBytes + (kClTripTab + 2) * sizeof(unsigned short);
}

// (tried and measured no faster on 10^6 ions: an L2 evict-first policy on these loads, an L2 prefetch of the partner
// records two trips ahead, and entries staged through a cp.async ring in shared memory instead of registers)
__device__ __forceinline__ uint2 ld_entry(const uint2* p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// TSM: the type-pair table sits in shared memory (LDS.128); otherwise it is read through L1 / L2
//
// Structure of a sweep: segment -> batches of <= 32 clusters -> clusters -> trips.  The trips of a batch are numbered
// through (trip table: list offset of every trip), so the software pipeline - entries two trips ahead, partner records
// one trip ahead - runs across cluster boundaries, and the trip loop of a cluster is ONE basic block: every list is
// padded to whole trips by the list kernel (dummy entries: everything excluded, dummy window slot), every load is
// unconditional, the window update of a trip is carried into the next one.  Clusters whose list holds a partner
// beyond its ring (count bit 30) run a second instance of the loop that has the atomics.
template <int C, bool GRAD, bool EMIT, bool TSM>
__global__ void __launch_bounds__(32 * kClWarps, 1)
cl_pair_kernel(DeviceSystem S, CellArrays A, ClPairArgs P) {
    using F = ClFmt<C>;
    constexpr int K = F::kK;
    constexpr unsigned int FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TT = S.n_types * S.n_types;
    const int POOL = P.pool;
    const int plane = (int)cl_window_plane(POOL);                // doubles between the x, y and z planes
    PairConst* s_table = reinterpret_cast<PairConst*>(smem_raw);
    unsigned char* wbase = smem_raw + (TSM ? (((size_t)TT * sizeof(PairConst) + 15) & ~(size_t)15) : 0) +
                           (size_t)warp * cl_warp_smem(GRAD, POOL);
    double* acc = reinterpret_cast<double*>(wbase);                                      // [3][plane] (GRAD)
    int4* rowc = reinterpret_cast<int4*>(wbase + (GRAD ? 3 * (size_t)plane * sizeof(double) : 0));   // row o: {ring origin, first slot, slots, floor(2^32 / slots)}
    int* rowtab = reinterpret_cast<int*>(rowc + kClRowPitch);                            // [cell of the segment (+ the end)][row]: window start
    int* s_cnt = rowtab + (kClMaxSegLen + 1) * kClRowPitch;                              // list length of the batch's clusters (bit 30: needs the atomics)
    int* s_ci = s_cnt + kClBatch;                                                        // their cell (index inside the segment)
    uint4* s_next = reinterpret_cast<uint4*>(s_ci + kClBatch);                           // the next cluster's C records, then its C types
    unsigned short* s_trip = reinterpret_cast<unsigned short*>(reinterpret_cast<unsigned char*>(s_next) + kClNextBytes);
    if (TSM) {
        for (int t = threadIdx.x; t < TT; t += blockDim.x) s_table[t] = S.table[t];
    }
    if (GRAD) {
        for (int t = lane; t < 3 * plane; t += 32) acc[t] = 0.0;                         // flushes leave it zero again
    }
    __syncthreads();

    const GridParams g = *A.grid;
    const int H = g.n_half;
    const PosQ* __restrict__ spq = A.spq;
    const long long s0 = g.seg_lo, s1 = g.seg_hi;
    const size_t np = cl_pitch(A.start[g.ncell]);                // positions per slot array (xyz interleaved)
    const int tpc = P.cap / F::kSpan;                            // trips a list can hold (the cap is a multiple of 64)

    for (;;) {
        long long seg = 0;
        if (lane == 0) seg = s0 + atomicAdd(A.work_ctr + 1, 1);
        seg = __shfl_sync(FULL, seg, 0);
        if (seg >= s1) break;
        const int row = (int)(seg / g.nsx), sx = (int)(seg % g.nsx);
        const int xa = sx * g.seg_len, xb = min(xa + g.seg_len, g.nx);
        const int cy = row % g.ny, cz = row / g.ny;
        const int rowbase = row * g.nx;
        const int q0 = A.start[rowbase + xa] / C;
        const int q1 = min(A.start[rowbase + xb] / C, P.ccap);   // the segment's clusters (beyond the list storage: flags[3], the host retries)
        double ev = 0, ee = 0;
        unsigned long long n_cand_w = 0;
        unsigned int n_eval = 0;
        if (q0 < q1) {
            {   // lane o < H looks after half-shell row o: the ring of the row and where its window starts for every
                // cell of the segment (the last line: where it ends when the sweep is over)
                int rb = -1;
                if (lane < H) {
                    int dy, dz;
                    cl_row_offset(lane, g.reach, dy, dz);
                    const int y2 = cy + dy, z2 = cz + dz;
                    if (y2 >= 0 && y2 < g.ny && z2 < g.nz) rb = (z2 * g.ny + y2) * g.nx;
                }
                __syncwarp();
                if (lane < kClRowPitch) {
                    const unsigned int lay = P.layout[seg * 16 + lane];
                    const int cap = (int)(lay >> 16);
                    rowc[lane] = make_int4(rb >= 0 ? A.start[rb + max(xa - g.reach, 0)] : 0, (int)(lay & 0xffffu), cap,
                                           cap > 0 ? (int)(0x100000000ull / (unsigned long long)cap) : 0);
                }
                const int nc = xb - xa;
                for (int ci = 0; ci <= nc; ++ci) {
                    int v = 0;
                    if (rb >= 0) v = ci < nc ? A.start[rb + max(xa + ci - g.reach, 0)] : A.start[rb + min(xb - 1 + g.reach, g.nx - 1) + 1];
                    if (lane < kClRowPitch) rowtab[ci * kClRowPitch + lane] = v;
                }
                __syncwarp();
            }
            const int slot_par = sx & 1;
            // flush the window rows from the frontier of cell `from` to that of cell `to`: all rows at once, two lanes
            // per row (each takes half of the row's positions); positions beyond a row's ring were never accumulated
            // here (they went through the atomics): zeros
            auto flush = [&](int from, int to) {
                if constexpr (GRAD) {
                    const int o = lane >> 1;
                    if (o < H) {
                        const int lo = rowtab[from * kClRowPitch + o], hi = rowtab[to * kClRowPitch + o];
                        const int n = hi - lo;
                        if (n > 0) {
                            const int4 rc = rowc[o];
                            const int half = (n + 1) >> 1;
                            const int d0 = (lane & 1) ? half : 0, d1 = (lane & 1) ? n : half;
                            double* dst = P.gj + ((size_t)(2 * o + slot_par) * np + lo) * 3;
                            int ri = rc.z > 0 ? cl_ring_mod(lo + d0 - rc.x, rc.z, (unsigned int)rc.w) : 0;
#pragma unroll 2
                            for (int d = d0; d < d1; ++d) {
                                const int idx = rc.y + ri;
                                double vx = 0, vy = 0, vz = 0;
                                if (d < rc.z) {
                                    vx = acc[idx]; vy = acc[plane + idx]; vz = acc[2 * plane + idx];
                                    acc[idx] = 0.0; acc[plane + idx] = 0.0; acc[2 * plane + idx] = 0.0;
                                }
                                dst[3 * d] = vx;
                                dst[3 * d + 1] = vy;
                                dst[3 * d + 2] = vz;
                                if (++ri >= rc.z) ri = 0;
                            }
                        }
                    }
                    __syncwarp();
                }
            };
            int cur_ci = 0;                                       // the windows stand at the segment's first cell

            for (int qb = q0; qb < q1;) {
                __syncwarp();                                     // the previous batch's reads of the tables are over
                int nq, ntrip;
                {   // the batch: up to 32 clusters, fewer if their trips do not fit the table
                    const int q = qb + lane;
                    const int raw = q < q1 ? P.count[q] : 0;
                    const int cnt = min(raw & 0x3fffffff, P.cap);
                    int T = q < q1 ? min(max(1, (cnt + F::kSpan - 1) / F::kSpan), kClTripTab) : 0;
                    int incl = T;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += t;
                    }
                    nq = __popc(__ballot_sync(FULL, q < q1 && incl <= kClTripTab));      // (>= 1: a list has at most kClTripTab trips)
                    ntrip = __shfl_sync(FULL, incl, nq - 1);
                    s_cnt[lane] = cnt | (raw & 0x40000000);
                    s_ci[lane] = q < q1 ? A.scell[C * q] - rowbase - xa : 0;
                    if (lane < nq) {
                        for (int t = 0; t < T; ++t) s_trip[incl - T + t] = (unsigned short)(lane * tpc + t);
                    }
                    if (lane < 2) s_trip[ntrip + lane] = 0;       // look-ahead beyond the batch: its first trip again (never evaluated)
                }
                __syncwarp();
                {   // the batch's first two lists on their way to L2
                    const int n0 = s_cnt[0] & 0x3fffffff, n1 = nq > 1 ? s_cnt[1] & 0x3fffffff : 0;
                    if (16 * lane < n0) prefetch_l2(P.list + (size_t)qb * P.cap + 16 * lane);
                    if (16 * lane < n1) prefetch_l2(P.list + (size_t)(qb + 1) * P.cap + 16 * lane);
                }
                const uint2* lbase = P.list + (size_t)qb * P.cap + lane;
                auto load_entries = [&](int k, uint2 (&E)[K]) {   // trip k of the batch
                    const uint2* src = lbase + (size_t)s_trip[k] * F::kSpan;
#pragma unroll
                    for (int kk = 0; kk < K; ++kk) E[kk] = ld_entry(src + 32 * kk);
                };

                int ipos0 = 0;
                double qi2[C], gix[C], giy[C], giz[C];
                unsigned int trow[C];                             // byte offset of the atom's row of the type-pair table
                unsigned int pad_excl = 0;                        // "excluded" bits of the cluster's padding atoms
                // a cluster's atoms (C records of 32 bytes, C types) are fetched one cluster ahead by the first 2 C + 1
                // lanes, 16 bytes each, and parked in shared memory: no register copies of them in the hot loop
                auto fetch_cluster = [&](int q) -> uint4 {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (lane < 2 * C) v = reinterpret_cast<const uint4*>(spq + (size_t)C * q)[lane];
                    else if (lane < 2 * C + C) v.x = (unsigned int)A.stype[(size_t)C * q + (lane - 2 * C)];
                    return v;
                };
                int cb = 0;                                       // buffer of the cluster being evaluated
                auto park_cluster = [&](uint4 v) {                // into the other buffer
                    uint4* dst = s_next + (cb ^ 1) * 9;
                    if (lane < 2 * C) dst[lane] = v;
                    else if (lane < 2 * C + C) reinterpret_cast<unsigned int*>(dst + 2 * C)[lane - 2 * C] = v.x;
                };
                uint4 nxt = fetch_cluster(qb);
                __syncwarp();
                park_cluster(nxt);
                __syncwarp();
                const double* cp = reinterpret_cast<const double*>(s_next);      // x, y, z, q of the cluster's atoms
                uint2 E0[K], E1[K];
                PosQ G[K];
                load_entries(0, E0);
                load_entries(1, E1);
#pragma unroll
                for (int kk = 0; kk < K; ++kk) G[kk] = load_posq(spq + E0[kk].x);
                // the window update a trip leaves to the next one (slot POOL: the dummy slot)
                double pjx[K], pjy[K], pjz[K];
                int pri[K];
#pragma unroll
                for (int kk = 0; kk < K; ++kk) { pjx[kk] = pjy[kk] = pjz[kk] = 0.0; pri[kk] = POOL; }
                auto window_update = [&]() {
                    if constexpr (GRAD && !EMIT) {
                    double a0[K], a1[K], a2[K];
#pragma unroll
                    for (int kk = 0; kk < K; ++kk) { a0[kk] = acc[pri[kk]]; a1[kk] = acc[plane + pri[kk]]; a2[kk] = acc[2 * plane + pri[kk]]; }
#pragma unroll
                    for (int kk = 0; kk < K; ++kk) {
                        // plain read-modify-write: this warp is the only writer and the partners of one trip are distinct
                        // (K == 2: the two entries of a lane are 32 list places apart - distinct partners, distinct slots;
                        // dummy entries all add 0 to the dummy slot)
                        acc[pri[kk]] = a0[kk] + pjx[kk];
                        acc[plane + pri[kk]] = a1[kk] + pjy[kk];
                        acc[2 * plane + pri[kk]] = a2[kk] + pjz[kk];
                    }
                    }
                };

                int k = 0;                                        // trip of the batch
                // the trips of one cluster: E0 / G are this trip's entries and records, E1 the next trip's entries
                auto run_trips = [&](auto ovf_tag, const int T) {
                    constexpr bool OVF = decltype(ovf_tag)::value;
#pragma unroll 1
                    for (int t = 0; t < T; ++t, ++k) {
                        window_update();                          // of the previous trip
                        double dx[K][C], dy[K][C], dz[K][C], r2[K][C];
#pragma unroll
                        for (int kk = 0; kk < K; ++kk) {
#pragma unroll
                            for (int a = 0; a < C; ++a) {
                                // (the cluster's coordinates stay in shared memory: broadcast loads, 24 registers less)
                                dx[kk][a] = cp[4 * a] - G[kk].x; dy[kk][a] = cp[4 * a + 1] - G[kk].y; dz[kk][a] = cp[4 * a + 2] - G[kk].z;
                                // strict double, left-to-right, no FMA: Point3d.distance squared
                                r2[kk][a] = __dadd_rn(__dadd_rn(__dmul_rn(dx[kk][a], dx[kk][a]), __dmul_rn(dy[kk][a], dy[kk][a])),
                                                      __dmul_rn(dz[kk][a], dz[kk][a]));
                            }
                        }
                        constexpr int NC = K * C;
                        double r2c[NC], qq2[NC], e1[NC], e2v[NC], f[NC];
                        PairConst pc[NC];
                        unsigned int em[K];
                        // a pair that does not count (beyond the cutoff, excluded, not later in the sorted order, padding,
                        // a dummy entry) has its two prefactors set to zero: r2 is clamped away from 0 and bounded by
                        // the list construction, so every intermediate is finite and the products are exact zeros
#pragma unroll
                        for (int kk = 0; kk < K; ++kk) {
                            em[kk] = E0[kk].y | pad_excl;
                            const unsigned int tj = ((em[kk] >> kMetaTypeShift) & 127u) * (unsigned int)sizeof(PairConst);
#pragma unroll
                            for (int a = 0; a < C; ++a) {
                                const int c = kk * C + a;
                                const bool on = !(r2[kk][a] > S.r2_max) && !((em[kk] >> (kMetaExclShift + a)) & 1u);
                                const bool f14 = (em[kk] >> (kMetaF14Shift + a)) & 1u;
                                if (EMIT) {
                                    if (on) {
                                        const int oi = A.orig[ipos0 + a], oj = A.orig[E0[kk].x];
                                        const unsigned long long q = atomicAdd(P.cursor, 1ull);
                                        if ((long long)q < P.capacity) { P.out_i[q] = min(oi, oj); P.out_j[q] = max(oi, oj); P.out_14[q] = f14; }
                                    }
                                }
                                if (TSM) pc[c] = *reinterpret_cast<const PairConst*>(smem_raw + trow[a] + tj);
                                else pc[c] = *reinterpret_cast<const PairConst*>(reinterpret_cast<const unsigned char*>(S.table) + trow[a] + tj);
                                pc[c].k1 = on ? pc[c].k1 : 0.0;
                                // 0, 1 or 0.75 (1-4 pairs) as the high word of a double whose low word is zero
                                const unsigned int sch = on ? (f14 ? 0x3fe80000u : 0x3ff00000u) : 0u;
                                n_eval += sch >> 29;              // 1 for both
                                qq2[c] = (qi2[a] * G[kk].q) * __hiloint2double((int)sch, 0);
                                r2c[c] = clamp_r2(r2[kk][a]);
                            }
                        }
                        // the pipeline: records of the next trip (G is free now), entries of the one after it
#pragma unroll
                        for (int kk = 0; kk < K; ++kk) {
                            // the address carries a (never taken, for non-negative r2) dependence on r2, so that this load
                            // is issued after the wait for this trip's operands and not before it (a NaN r2 with the sign
                            // bit set moves the index by one: the arrays have a spare element, the result is NaN anyway)
                            unsigned long long addr;
                            asm("mad.wide.u32 %0, %1, 32, %2;" : "=l"(addr)
                                : "r"(E1[kk].x + ((unsigned int)__double2hiint(r2c[kk * C]) >> 31)), "l"(spq));
                            G[kk] = load_posq(reinterpret_cast<const PosQ*>(addr));
                        }
                        uint2 E2[K];
                        load_entries(k + 2, E2);
                        if (!EMIT) {
                            pair_terms_n<GRAD, NC>(r2c, pc, qq2, e1, e2v, f);
#pragma unroll
                            for (int c = 0; c < NC; ++c) { ev += e1[c]; ee += e2v[c]; }
                        }
                        if (GRAD && !EMIT) {
#pragma unroll
                            for (int kk = 0; kk < K; ++kk) {
                                double gjx = 0, gjy = 0, gjz = 0;
#pragma unroll
                                for (int a = 0; a < C; ++a) {
                                    const int c = kk * C + a;
                                    gix[a] = fma(f[c], dx[kk][a], gix[a]);
                                    giy[a] = fma(f[c], dy[kk][a], giy[a]);
                                    giz[a] = fma(f[c], dz[kk][a], giz[a]);
                                    gjx = fma(-f[c], dx[kk][a], gjx);
                                    gjy = fma(-f[c], dy[kk][a], gjy);
                                    gjz = fma(-f[c], dz[kk][a], gjz);
                                }
                                int ri = (int)(em[kk] & kMetaSlotMask);
                                if (OVF) {                        // a row window denser than its ring: atomics (order not fixed)
                                    if (em[kk] & kMetaOvf) {
                                        atomicAdd(P.gj_ovf + 3 * (size_t)E0[kk].x, gjx);
                                        atomicAdd(P.gj_ovf + 3 * (size_t)E0[kk].x + 1, gjy);
                                        atomicAdd(P.gj_ovf + 3 * (size_t)E0[kk].x + 2, gjz);
                                        A.flags[2] = 1;
                                        ri = POOL;                // and nothing to the window
                                        gjx = gjy = gjz = 0.0;
                                    }
                                }
                                pjx[kk] = gjx; pjy[kk] = gjy; pjz[kk] = gjz; pri[kk] = ri;
                            }
                        }
#pragma unroll
                        for (int kk = 0; kk < K; ++kk) { E0[kk] = E1[kk]; E1[kk] = E2[kk]; }
                    }
                };

                for (int il = 0; il < nq; ++il) {
                    // ---- a new cluster: the parked atoms become the current ones
                    ipos0 = C * (qb + il);
                    const int cnt_raw = s_cnt[il];
                    const int cnt = cnt_raw & 0x3fffffff;
                    const int ci_new = s_ci[il];
                    pad_excl = 0;
                    cb ^= 1;
                    cp = reinterpret_cast<const double*>(s_next + cb * 9);
                    const int* nt = reinterpret_cast<const int*>(s_next + cb * 9 + 2 * C);
#pragma unroll
                    for (int a = 0; a < C; ++a) {
                        qi2[a] = 2.0 * S.ele_scale * cp[4 * a + 3];              // doubled energies, see pair_terms
                        const int ty = nt[a];
                        trow[a] = (unsigned int)((ty & 0x7f) * S.n_types * (int)sizeof(PairConst));
                        if (ty & kPadType) pad_excl |= 1u << (kMetaExclShift + a);
                        gix[a] = giy[a] = giz[a] = 0.0;
                    }
                    if (il + 1 < nq) nxt = fetch_cluster(qb + il + 1);           // the next cluster's atoms, one cluster ahead
                    if (il + 2 < nq && 16 * lane < (s_cnt[il + 2] & 0x3fffffff)) // and the list after the next one towards L2
                        prefetch_l2(P.list + (size_t)(qb + il + 2) * P.cap + 16 * lane);
                    if (GRAD && ci_new != cur_ci) {               // the sweep moves on: cells that left the windows are flushed
                        window_update();                          // (the last trip's update first)
#pragma unroll
                        for (int kk = 0; kk < K; ++kk) { pjx[kk] = pjy[kk] = pjz[kk] = 0.0; pri[kk] = POOL; }
                        __syncwarp();
                        flush(cur_ci, ci_new);
                        cur_ci = ci_new;
                    }
                    n_cand_w += (unsigned long long)cnt * C;
                    const int T = min(max(1, (cnt + F::kSpan - 1) / F::kSpan), kClTripTab);
                    if (cnt_raw & 0x40000000) run_trips(std::true_type{}, T);
                    else run_trips(std::false_type{}, T);
                    // ---- the cluster is complete
                    if (il + 1 < nq) {                            // park the atoms fetched at its start
                        __syncwarp();
                        park_cluster(nxt);
                        __syncwarp();
                    }
                    if (GRAD && !EMIT) {                          // its i-side sums over the warp
                        constexpr int NV = C == 2 ? 8 : 16;
                        double v[NV];
#pragma unroll
                        for (int i = 0; i < NV; ++i) v[i] = 0.0;
#pragma unroll
                        for (int a = 0; a < C; ++a) { v[3 * a] = gix[a]; v[3 * a + 1] = giy[a]; v[3 * a + 2] = giz[a]; }
                        warp_multi_sum<NV>(v, lane);
                        const int vi = lane / (32 / NV);          // value completed in this lane
                        if ((lane & (32 / NV - 1)) == 0 && vi < 3 * C)
                            P.gi[3 * (size_t)ipos0 + vi] = v[0];
                    }
                }
                window_update();                                  // the batch's last trip
                qb += nq;
            }
            __syncwarp();
            flush(cur_ci, xb - xa);                               // the segment is done: everything still in the windows goes out
        }
        if (!EMIT) {
            ev = 0.5 * warp_sum(ev);                              // energies were accumulated doubled (pair_terms)
            ee = 0.5 * warp_sum(ee);
            for (int o = 16; o; o >>= 1) n_eval += __shfl_xor_sync(FULL, n_eval, o);
            if (lane == 0) {
                P.epart[2 * (size_t)seg] = ev;
                P.epart[2 * (size_t)seg + 1] = ee;
                P.cpart[2 * (size_t)seg] = n_eval;
                P.cpart[2 * (size_t)seg + 1] = n_cand_w;
            }
        }
    }
}

// ---- final reduction of the cluster path ------------------------------------------------------------------
struct ClReduceArgs {
    int n_spad;
    int rank, world;
    const double* gi;
    const double* gj;
    double* gj_ovf;
    const uint2* cellmask;         // [cells] cl_cellmask_kernel
    ReduceArgs R;                  // bonded slots, energy partials, units, outputs (gpart / direct unused)
};

// Which slots were written for a cell depends on the cell only: one thread per cell works that out
// (mask.x: bit o = slot 2 o + (sa & 1) holds a flush of half-shell row o, bit 29 = sa & 1, bit 30 = sb & 1, bit 31 = the
// cell's own segment belongs to this shard; mask.y: the same for the second writer sb), then one thread per sorted
// position sums them.
__global__ void __launch_bounds__(256)
cl_cellmask_kernel(CellArrays A, int rank, int world, uint2* __restrict__ mask) {
    const GridParams& g = *A.grid;
    const long long s0 = g.seg_lo, s1 = g.seg_hi;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < g.ncell; c += gridDim.x * blockDim.x) {
        if (A.count[c] == 0) { mask[c] = make_uint2(0u, 0u); continue; }
        const int cx = c % g.nx, cy = (c / g.nx) % g.ny, cz = c / (g.nx * g.ny);
        // the segments of an i-row whose windows covered this cell - cells [xa - reach, xb - 1 + reach] - are sa and
        // sb (at most two, of different parity); row o of the half shell belongs to the i-row (cy - dy, cz - dz)
        const int sa = max(cx - g.reach, 0) / g.seg_len, sb = min((cx + g.reach) / g.seg_len, g.nsx - 1);
        unsigned int ma = 0, mb = 0;
        for (int o = 0; o < g.n_half; ++o) {
            int dy, dz;
            cl_row_offset(o, g.reach, dy, dz);
            const int yi = cy - dy, zi = cz - dz;
            if (yi < 0 || yi >= g.ny || zi < 0) continue;
            const int irow = zi * g.ny + yi;
            auto wrote = [&](int sx) {       // in this shard and holding at least one atom
                const long long seg = (long long)irow * g.nsx + sx;
                if (seg < s0 || seg >= s1) return false;
                const int xa = sx * g.seg_len, xb = min(xa + g.seg_len, g.nx);
                return A.start[irow * g.nx + xb] != A.start[irow * g.nx + xa];
            };
            if (wrote(sa)) ma |= 1u << o;
            if (sb != sa && wrote(sb)) mb |= 1u << o;
        }
        const long long own_seg = (long long)(cz * g.ny + cy) * g.nsx + cx / g.seg_len;
        if (own_seg >= s0 && own_seg < s1) ma |= 1u << 31;
        ma |= (unsigned int)(sa & 1) << 29 | (unsigned int)(sb & 1) << 30;
        mask[c] = make_uint2(ma, mb);
    }
}

__global__ void __launch_bounds__(kReduceWarps * 32, 3)
cl_reduce_kernel(CellArrays A, ClReduceArgs Q) {
    const GridParams& g = *A.grid;
    if (blockIdx.x == gridDim.x - 1) {
        // last block: energies and pair counts over this shard's segments
        const long long s0 = g.seg_lo, s1 = g.seg_hi;
        ReduceArgs R = Q.R;
        R.epart += 2 * s0; R.cpart += 2 * s0; R.n_epart = (int)(s1 - s0);
        R.overflow = A.flags;
        reduce_energy_block(R);
        return;
    }
    if (!Q.R.want_grad) return;
    const ReduceArgs& R = Q.R;
    const bool use_ovf = A.flags[2] != 0;
    const int n_sorted = A.start[g.ncell];
    const size_t np = cl_pitch(n_sorted);
    // a shard only holds (and only wrote forces for) the positions of the cells it touches; the caller zeroed the rest
    int c_lo, c_hi;
    shard_cell_range(g, Q.rank, Q.world, c_lo, c_hi);
    const int p_lo = A.start[c_lo], p_hi = A.start[c_hi];
    const bool with_bonded = Q.world == 1;                        // sharded: bonded slots are added by cl_bonded_add_kernel
    for (int p = p_lo + blockIdx.x * blockDim.x + threadIdx.x; p < p_hi; p += (gridDim.x - 1) * blockDim.x) {
        const int a = A.orig[p];
        if (a < 0) continue;                                      // padding
        const uint2 m = Q.cellmask[A.scell[p]];
        const unsigned int pa = (m.x >> 29) & 1u, pb = (m.x >> 30) & 1u;
        double gx = 0, gy = 0, gz = 0;
        if (m.x >> 31) { gx = Q.gi[3 * (size_t)p]; gy = Q.gi[3 * (size_t)p + 1]; gz = Q.gi[3 * (size_t)p + 2]; }
        // the slot loads of an atom are independent: issue them in two batches (predicated), sum in row order
#pragma unroll
        for (int o0 = 0; o0 < kClMaxRows; o0 += 7) {
            double vx[7], vy[7], vz[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const int o = o0 + i;
                const double* src = Q.gj + ((size_t)(2 * o + pa) * np + p) * 3;
                const bool u = o < kClMaxRows && ((m.x >> o) & 1u);
                vx[i] = u ? src[0] : 0.0; vy[i] = u ? src[1] : 0.0; vz[i] = u ? src[2] : 0.0;
            }
#pragma unroll
            for (int i = 0; i < 7; ++i) { gx += vx[i]; gy += vy[i]; gz += vz[i]; }
        }
        for (unsigned int v = m.y & 0x1fffu; v; v &= v - 1) {     // cells next to a segment boundary: the second writer
            const double* src = Q.gj + ((size_t)(2 * (__ffs(v) - 1) + pb) * np + p) * 3;
            gx += src[0]; gy += src[1]; gz += src[2];
        }
        if (use_ovf) {       // (and back to zero for the next evaluation)
            double* ov = Q.gj_ovf + 3 * (size_t)p;
            gx += ov[0]; gy += ov[1]; gz += ov[2];
            ov[0] = 0.0; ov[1] = 0.0; ov[2] = 0.0;
        }
        if (with_bonded) {
            for (long long e = R.slot_ptr[a]; e < R.slot_ptr[a + 1]; ++e) {
                const long long sl = R.slot_idx[e];
                if (sl >= R.slot_lo && sl < R.slot_hi) {
                    gx += R.slots[3 * sl];
                    gy += R.slots[3 * sl + 1];
                    gz += R.slots[3 * sl + 2];
                }
            }
        }
        double* o = R.out + kOutHeader + 3 * (size_t)a;
        o[0] = convert_unit(gx, R.unit_num, R.unit_den);
        o[1] = convert_unit(gy, R.unit_num, R.unit_den);
        o[2] = convert_unit(gz, R.unit_num, R.unit_den);
    }
}

// sharded evaluation: the bonded terms are split by term index, not by space - their gradient slots are added to the
// (zeroed, then partly written) output by original atom index
__global__ void __launch_bounds__(256)
cl_bonded_add_kernel(ReduceArgs R) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= R.n) return;
    double gx = 0, gy = 0, gz = 0;
    bool any = false;
    for (long long e = R.slot_ptr[a]; e < R.slot_ptr[a + 1]; ++e) {
        const long long sl = R.slot_idx[e];
        if (sl >= R.slot_lo && sl < R.slot_hi) {
            gx += R.slots[3 * sl];
            gy += R.slots[3 * sl + 1];
            gz += R.slots[3 * sl + 2];
            any = true;
        }
    }
    if (!any) return;
    double* o = R.out + kOutHeader + 3 * (size_t)a;
    o[0] += convert_unit(gx, R.unit_num, R.unit_den);
    o[1] += convert_unit(gy, R.unit_num, R.unit_den);
    o[2] += convert_unit(gz, R.unit_num, R.unit_den);
}

// ---- host side ---------------------------------------------------------------------------------------
// launch geometry of the cluster path: warps per CTA of the gradient kernel from what the windows leave of the
// shared memory; a type-pair table too large to sit next to them is read through L1 / L2 instead
inline int cl_configure(CellPlan& p, const DeviceSystem& ds, std::string& err) {
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t table = (((size_t)ds.n_types * ds.n_types * sizeof(PairConst)) + 15) & ~(size_t)15;
    p.cl_table_in_smem = table <= 48 * 1024 ? 1 : 0;
    const size_t tb = p.cl_table_in_smem ? table : 0;
    // the slot pool of a warp's force window from what kClWarps warps leave of the shared memory
    const long long per_warp = ((long long)max_smem - (long long)tb) / kClWarps;
    long long pool = (per_warp - (long long)cl_warp_smem(true, 0)) / (3 * (long long)sizeof(double));
    pool = std::min<long long>((1 << kMetaSlotBits) - 1, pool / 8 * 8);
    if (pool < 256) { err = "cell path: not enough shared memory for the force windows"; return JME_ERR_UNSUPPORTED; }
    p.arr.ring = (int)pool;
    p.cl_warps = kClWarps;
    p.cl_smem_grad = tb + kClWarps * cl_warp_smem(true, (int)pool);
    p.cl_smem_energy = tb + kClWarps * cl_warp_smem(false, (int)pool);
    // segments per resident warp: eight on one GPU (measured on 10^6 ions: segments of 12 / 11 / 8 / 6 / 5 cells give a
    // pair sweep of 1.74 / 1.72 / 1.68 / 1.68 / 1.69 ms - shorter segments balance the warps better, longer ones flush
    // their windows less often), and six per warp of THIS rank's shard in a multi-GPU job (the shards are 1 / world of
    // the segments each: with the single-GPU segment length a rank at N = 4 had 1.6 segments per warp and its pair
    // kernel ran 45 % over its share; measured at N = 2: 8 / 12 / 16 per warp give steps of 1.503 / 1.483 / 1.498 ms)
    int seg_mult = p.world > 1 ? 6 * p.world : 8;
    if (const char* e = std::getenv("JME_SEG_MULT")) seg_mult = std::max(1, std::atoi(e));      // (development knob)
    p.arr.seg_target = seg_mult * p.sms * kClWarps;
    return JME_OK;
}

template <int C>
inline int cl_run_lists_t(CellPlan& p, const DeviceSystem& ds, cudaStream_t st, uint64_t& launches, std::string& err) {
    ClListArgs A{};
    A.list = p.d_clist; A.count = p.d_count; A.cap = p.list_cap; A.ccap = p.cluster_cap; A.rank = p.rank; A.world = p.world;
    A.layout = p.d_layout; A.pool = p.arr.ring;
    const int grid = p.sms * kClListMinBlocks;
    const long long resident = (long long)grid * kListWarps;
    A.parts = (int)std::max(1ll, std::min(8ll, 2 * resident * 12 / std::max(1, ds.n)));
    cudaMemsetAsync(p.arr.work_ctr, 0, 2 * sizeof(int), st);
    cl_layout_kernel<<<std::max(1, std::min(4 * p.sms, (ds.n + 2047) / 2048)), 256, 0, st>>>(p.arr, p.arr.ring, p.d_layout);
    ++launches;
    cl_list_kernel<C><<<grid, kListThreads, 0, st>>>(p.arr, A);
    ++launches;
    if (p.arr.n_excl_ent > 0) {
        cl_mark_kernel<C><<<(unsigned int)((p.arr.n_excl_ent + 255) / 256), 256, 0, st>>>(p.arr, p.d_excl_src, p.d_clist, p.d_count,
                                                                                        p.list_cap, p.cluster_cap, p.rank, p.world);
        ++launches;
    }
    JME_CELL_CHECK();
    return JME_OK;
}

inline int cl_run_lists(CellPlan& p, const DeviceSystem& ds, cudaStream_t st, uint64_t& launches, std::string& err) {
    if (ds.n == 0) return JME_OK;
    return p.cluster == 2 ? cl_run_lists_t<2>(p, ds, st, launches, err) : cl_run_lists_t<4>(p, ds, st, launches, err);
}

inline void cl_fill_pair_args(CellPlan& p, ClPairArgs& P) {
    P.list = p.d_clist; P.count = p.d_count; P.cap = p.list_cap; P.ccap = p.cluster_cap;
    P.rank = p.rank; P.world = p.world; P.n_spad = p.n_spad + 1; P.pool = p.arr.ring; P.layout = p.d_layout;
    P.gi = p.d_gi; P.gj = p.d_gj; P.gj_ovf = p.d_gj_ovf;
    P.epart = p.d_epart; P.cpart = p.d_cpart;
}

template <int C, bool GRAD, bool EMIT, bool TSM>
inline int cl_launch_pairs_t(CellPlan& p, const DeviceSystem& ds, const ClPairArgs& P, cudaStream_t st, std::string& err) {
    auto k = cl_pair_kernel<C, GRAD, EMIT, TSM>;
    const size_t smem = GRAD ? p.cl_smem_grad : p.cl_smem_energy;
    const int warps = kClWarps;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { err = std::string("cell path: shared memory request refused: ") + cudaGetErrorString(e); return JME_ERR_CUDA; }
    }
    k<<<p.sms, 32 * warps, smem, st>>>(ds, p.arr, P);
    JME_CELL_CHECK();
    return JME_OK;
}

template <int C, bool GRAD, bool EMIT>
inline int cl_launch_pairs(CellPlan& p, const DeviceSystem& ds, const ClPairArgs& P, cudaStream_t st, std::string& err) {
    return p.cl_table_in_smem ? cl_launch_pairs_t<C, GRAD, EMIT, true>(p, ds, P, st, err)
                              : cl_launch_pairs_t<C, GRAD, EMIT, false>(p, ds, P, st, err);
}

inline int cl_run_pairs(CellPlan& p, const DeviceSystem& ds, bool grad, cudaStream_t st, uint64_t& launches, std::string& err) {
    if (ds.n == 0) return JME_OK;
    ClPairArgs P{};
    cl_fill_pair_args(p, P);
    cudaMemsetAsync(p.arr.work_ctr + 1, 0, sizeof(int), st);
    ++launches;
    if (p.cluster == 2) return grad ? cl_launch_pairs<2, true, false>(p, ds, P, st, err) : cl_launch_pairs<2, false, false>(p, ds, P, st, err);
    return grad ? cl_launch_pairs<4, true, false>(p, ds, P, st, err) : cl_launch_pairs<4, false, false>(p, ds, P, st, err);
}

// final reduction (gradient: one thread per sorted position + the energy block; energy only: the energy block)
inline int cl_run_reduce(CellPlan& p, const DeviceSystem& ds, const ReduceArgs& R, cudaStream_t st, uint64_t& launches, std::string& err) {
    ClReduceArgs Q{};
    Q.n_spad = p.n_spad + 1; Q.rank = p.rank; Q.world = p.world;
    Q.gi = p.d_gi; Q.gj = p.d_gj; Q.gj_ovf = p.d_gj_ovf; Q.cellmask = p.d_cellmask;
    Q.R = R;
    const int nt = kReduceWarps * 32;
    if (R.want_grad && p.world > 1)          // a shard writes the atoms it touched only
        cudaMemsetAsync(R.out + kOutHeader, 0, 3 * (size_t)ds.n * sizeof(double), st);
    if (R.want_grad) {
        cl_cellmask_kernel<<<std::max(1, std::min(4 * p.sms, (ds.n + 255) / 256)), 256, 0, st>>>(p.arr, p.rank, p.world, p.d_cellmask);
        ++launches;
    }
    // gradient: a grid-stride loop over the sorted positions (their number is only known on the device)
    const int grid = (R.want_grad ? std::max(1, std::min(16 * p.sms, (ds.n + nt - 1) / nt)) : 0) + 1;
    cl_reduce_kernel<<<grid, nt, 0, st>>>(p.arr, Q);
    ++launches;
    if (R.want_grad && p.world > 1 && R.slot_hi > R.slot_lo) {
        cl_bonded_add_kernel<<<(ds.n + 255) / 256, 256, 0, st>>>(R);
        ++launches;
    }
    JME_CELL_CHECK();
    return JME_OK;
}

inline int cl_run_pairlist(CellPlan& p, const DeviceSystem& ds, int* d_i, int* d_j, unsigned char* d_f, long long capacity,
                           unsigned long long* d_cursor, cudaStream_t st, uint64_t& launches, std::string& err) {
    int rc = run_cell_sort(p, ds, st, launches, err);
    if (rc) return rc;
    if ((rc = cl_run_lists(p, ds, st, launches, err))) return rc;
    if (ds.n == 0) return JME_OK;
    ClPairArgs P{};
    cl_fill_pair_args(p, P);
    P.out_i = d_i; P.out_j = d_j; P.out_14 = d_f; P.capacity = capacity; P.cursor = d_cursor;
    cudaMemsetAsync(p.arr.work_ctr + 1, 0, sizeof(int), st);
    ++launches;
    return p.cluster == 2 ? cl_launch_pairs<2, false, true>(p, ds, P, st, err) : cl_launch_pairs<4, false, true>(p, ds, P, st, err);
}

}  // namespace jme
