// jme_clusters.cuh - the cutoff path: cluster lists, Newton's third law, sliding force windows.
//
// The reference tests every non-excluded pair against the cutoff (MMFF94VdwComponent.java:136-139,
// MMFF94ElectrostaticComponent.java:139-141); this path selects the SAME pair set with O(N) work and evaluates
// every unordered pair exactly ONCE.  It replaces the round-1 per-atom survivor lists (each pair evaluated from
// both of its atoms), see DESIGN.md section 4.
//
//   sort        (jme_cells.cuh) cell edge = cutoff / 2, atoms sorted by cell (z, y, x), rank inside a cell by original
//               index; every cell is padded to a multiple of C sorted positions, so CLUSTERS of C consecutive
//               positions never straddle a cell.  Padding records sit far away and never pass a distance test.
//   cl_list     FP32, conservative: per cluster the list of partner ATOMS j later in the sorted order (the half
//               shell: the rest of the own cell row, then the rows dy > 0 and dz > 0) that are inside the padded
//               cutoff of at least one atom of the cluster.  One 32-bit entry per (cluster, j):
//               [ position of j relative to the row's window | row of the half shell | type of j | C excluded bits
//               | C one-four bits ].
//   cl_mark     one thread per exclusion entry: finds the pair's list entry (binary search, entries are sorted) and
//               sets the excluded / 1-4 bit of the atom, so the FP64 kernel needs no exclusion lookup at all.
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
//   cl_reduce   per atom: i-side sum + the <= 26 j-side slots that were written for its cell (derived from the grid,
//               nothing stored) + bonded slots; unit conversion; scatter to the caller's atom order.
#pragma once

#include "jme_cells.cuh"

namespace jme {

constexpr int kClRing = 128;          // window slots (atoms) per neighbour row and warp; power of two
constexpr int kClMaxRows = 13;        // rows of the half shell at reach 2
constexpr int kClMaxSegLen = 12;      // cells per segment (GridParams::seg_len is clamped to this in grid_setup_kernel)
constexpr int kClRowPitch = 16;       // row-info entries per cell
constexpr int kClBatch = 32;          // clusters staged per batch
constexpr int kClMaxWarps = 5;        // warps per CTA of the gradient kernel (one CTA per SM: the windows fill shared memory)
constexpr int kClEnergyWarps = 8;     // energy-only kernel (no windows)

template <int C>
struct ClFmt {
    static_assert(C == 1 || C == 2 || C == 4, "cluster size");
    static constexpr int kK = 4 / C;                               // list entries per lane per trip: C * kK = 4 chains
    static constexpr int kSpan = 32 * kK;                          // entries per trip
    static constexpr int kRelBits = (32 - 2 * C - 11) < 16 ? (32 - 2 * C - 11) : 16;
    static constexpr unsigned int kRelMask = (1u << kRelBits) - 1u;
    static constexpr int kRowShift = kRelBits;                     // 4 bits: row of the half shell
    static constexpr int kTypeShift = kRelBits + 4;                // 7 bits: type index of j
    static constexpr unsigned int kKeyMask = (1u << (kRelBits + 4)) - 1u;   // (row, rel): entries are sorted by it
    static constexpr int kExclShift = 32 - 2 * C;                  // C bits: pair (atom a of the cluster, j) is excluded
    static constexpr int kF14Shift = 32 - C;                       // C bits: ... is a 1-4 pair
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

// ---- cluster lists -----------------------------------------------------------------------------------
struct ClListArgs {
    unsigned int* list;            // [clusters][cap]
    int* count;                    // [clusters]
    int cap;
    int parts;                     // warps sharing one cell (small systems)
    int rank, world;
};

template <int C>
__global__ void __launch_bounds__(kListThreads, JME_LIST_MIN_BLOCKS)
cl_list_kernel(CellArrays A, ClListArgs P) {
    using F = ClFmt<C>;
    __shared__ float2 s_fi[kListWarps][32][4];  // atom il as (x,x) (y,y) (z,z) -: operands of the packed FP32 test
    __shared__ int s_jb[kListWarps][16];        // first candidate of half-shell row o
    __shared__ int s_cw[kListWarps][16];        // first position of row o's window (entries are relative to it)
    __shared__ int s_off[kListWarps][17];       // exclusive prefix of the candidate counts (virtual candidate index)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 (*fis)[4] = s_fi[warp];
    int* jbs = s_jb[warp];
    int* cws = s_cw[warp];
    int* off = s_off[warp];
    const GridParams g = *A.grid;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int H = g.n_half;
    const float4* __restrict__ fxyz = A.fxyz;
    const float qnan = __int_as_float(0x7fc00000);

    const long long s0 = (long long)g.n_seg * P.rank / P.world, s1 = (long long)g.n_seg * (P.rank + 1) / P.world;
    const int cell0 = cl_seg_first_cell(g, s0), cell1 = cl_seg_first_cell(g, s1);
    const int n_items = (cell1 - cell0) * P.parts;
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(A.work_ctr, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_items) break;
        const int c = cell0 + w / P.parts, part = w % P.parts;
        const int c0 = A.start[c], c1 = A.start[c + 1];
        const int nclu = (c1 - c0) / C;
        const int per = (nclu + P.parts - 1) / P.parts;
        const int k0 = part * per, k1 = min(nclu, k0 + per);
        if (k0 >= k1) continue;
        const int cx = c % g.nx, cy = (c / g.nx) % g.ny, cz = c / (g.nx * g.ny);
        const int a_begin = c0 + C * k0, a_end = c0 + C * k1;        // sorted positions of this item's clusters
        int total;
        {   // candidate range of half-shell row `lane` and the prefix sums that flatten the <= 13 ranges into one
            // virtual index space, so that every trip has all its slots filled
            int jb = 0, len = 0, cw = 0;
            if (lane < H) {
                int dy, dz;
                cl_row_offset(lane, g.reach, dy, dz);
                const int y2 = cy + dy, z2 = cz + dz;
                if (y2 >= 0 && y2 < g.ny && z2 < g.nz) {
                    const int x0 = max(cx - g.reach, 0), x1 = min(cx + g.reach, g.nx - 1);
                    const int row = (z2 * g.ny + y2) * g.nx;
                    cw = A.start[row + x0];
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
            if (lane < 16) { jbs[lane] = jb; cws[lane] = cw; off[lane] = incl - len; }
            if (lane == 16) off[16] = incl - len;                     // = total for lane >= H (len == 0)
        }
        for (int ch = a_begin; ch < a_end; ch += 32) {              // at most 32 atoms (32 / C clusters) at a time
            const int nc = min(32, a_end - ch);                       // multiple of C
            const int nq = nc / C;
            __syncwarp();
            {
                const float4 f = fxyz[min(ch + lane, a_end - 1)];
                fis[lane][0] = make_float2(f.x, f.x);
                fis[lane][1] = make_float2(f.y, f.y);
                fis[lane][2] = make_float2(f.z, f.z);
            }
            int my_cnt = 0;                                           // lane kq holds the list length of cluster ch / C + kq
            __syncwarp();
            unsigned long long rows = reinterpret_cast<unsigned long long>(P.list + (size_t)(ch / C) * P.cap);
            asm volatile("" : "+l"(rows));                            // keep the base in registers
            int r[kListK];                                            // row of this lane's k-th virtual index (monotone)
#pragma unroll
            for (int k = 0; k < kListK; ++k) r[k] = 0;
            for (int vb = 0; vb < total; vb += 32 * kListK) {
                float ax[kListK], ay[kListK], az[kListK];
                unsigned int en[kListK];
                int jj[kListK];
#pragma unroll
                for (int k = 0; k < kListK; ++k) {
                    const int vi = vb + 32 * k + lane;
                    const bool valid = vi < total;
                    while (r[k] < 15 && vi >= off[r[k] + 1]) ++r[k];
                    const int j = valid ? jbs[r[k]] + (vi - off[r[k]]) : 0;
                    const float4 a = fxyz[j];
                    ax[k] = valid ? a.x : qnan;                       // NaN never passes the distance test
                    ay[k] = a.y; az[k] = a.z;
                    jj[k] = j;
                    const unsigned int rel = (unsigned int)(j - cws[r[k]]);
                    if (valid && rel > F::kRelMask) A.flags[3] = 1;   // a row window with more atoms than the entry format holds
                    en[k] = (rel & F::kRelMask) | ((unsigned int)r[k] << F::kRowShift) |
                            ((unsigned int)__float_as_int(a.w) << F::kTypeShift);
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
                            st_cs_if(h[k], rows, o + __popc(m[k] & lt_mask), en[k]);
                            o += __popc(m[k]);
                        }
                        if (lane == kq) my_cnt += tot;
                    } else if (lane == 0) {
                        A.flags[0] = 1;
                    }
                }
            }
            if (lane < nq) P.count[ch / C + lane] = my_cnt;
        }
    }
}

// ---- exclusion / 1-4 bits ------------------------------------------------------------------------------
// One thread per (directed) exclusion entry; the direction whose source atom comes first in the sorted order owns
// the pair.  The partner's entry in the source cluster's list is found by binary search on (row, rel) - the list
// kernel emits entries in that order - and the atom's excluded / 1-4 bit is OR-ed in.  Pairs beyond the cutoff have
// no entry: nothing to mark.
template <int C>
__global__ void __launch_bounds__(256)
cl_mark_kernel(CellArrays A, const int* __restrict__ excl_src, unsigned int* __restrict__ list, const int* __restrict__ count,
               int cap, int rank, int world) {
    using F = ClFmt<C>;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.n_excl_ent) return;
    const unsigned int x = A.sexcl_ent[k];
    const int sj = (int)(x & ~kFlag14);
    const int si = A.sorted_of[excl_src[k]];
    if (si >= sj) return;
    const GridParams& g = *A.grid;
    const int ci = A.scell[si], cj = A.scell[sj];
    {
        const long long s0 = (long long)g.n_seg * rank / world, s1 = (long long)g.n_seg * (rank + 1) / world;
        if (ci < cl_seg_first_cell(g, s0) || ci >= cl_seg_first_cell(g, s1)) return;      // not this shard's list
    }
    const int nxy = g.nx * g.ny;
    const int ix = ci % g.nx, iy = (ci / g.nx) % g.ny, iz = ci / nxy;
    const int jx = cj % g.nx, jy = (cj / g.nx) % g.ny, jz = cj / nxy;
    const int o = cl_row_index(jy - iy, jz - iz, g.reach);
    if (o < 0 || jx - ix > g.reach || ix - jx > g.reach) return;
    const int cw = A.start[(jz * g.ny + jy) * g.nx + max(ix - g.reach, 0)];
    const unsigned int key = ((unsigned int)o << F::kRowShift) | (unsigned int)(sj - cw);
    const int q = si / C, a = si % C;
    unsigned int* lst = list + (size_t)q * cap;
    int lo = 0, hi = min(count[q], cap);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((lst[mid] & F::kKeyMask) < key) lo = mid + 1; else hi = mid;
    }
    if (lo < min(count[q], cap) && (lst[lo] & F::kKeyMask) == key)
        atomicOr(&lst[lo], 1u << (((x & kFlag14) ? F::kF14Shift : F::kExclShift) + a));
}

// ---- evaluation ----------------------------------------------------------------------------------------
struct ClPairArgs {
    const unsigned int* list;      // [clusters][cap]
    const int* count;              // [clusters]
    int cap;
    int rank, world;
    int n_spad;                    // pitch of the per-position arrays below
    double* gi;                    // [3][n_spad] i-side sums by sorted position
    double* gj;                    // [2 * kClMaxRows][3][n_spad] j-side window flushes: slot 2 * row + (segment & 1)
    double* gj_ovf;                // [3][n_spad] atomically accumulated contributions of atoms beyond a window's capacity
    double* epart;                 // [n_seg][2] per-segment energies (vdw, ele)
    unsigned long long* cpart;     // [n_seg][2] pairs evaluated, candidate pairs
    int table_in_smem;
    // pair-list emission (EMIT only)
    int* out_i; int* out_j; unsigned char* out_14; long long capacity; unsigned long long* cursor;
};

// sum of v[i] over the 32 lanes for N values at the cost of about N + log2(32 / N) shuffles instead of 5 N: in every
// halving step a lane keeps one half of its values and hands the other half to its partner.  On return v[0] of lane
// l holds the warp total of value  (l >> (5 - log2 N)) ... i.e. value i is complete in lane  i * (32 / N)  (and in
// the lanes that differ from it in the low log2(32 / N) bits).
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

template <int C>
struct ClNb {                      // what a lane holds of its ClFmt<C>::kK partners of one trip
    unsigned int e[ClFmt<C>::kK];
    PosQ p[ClFmt<C>::kK];
};

__host__ __device__ constexpr size_t cl_warp_smem(bool grad) {
    return (grad ? (size_t)3 * kClMaxRows * kClRing * sizeof(double) : 0) + kRingSlots * kChunk * sizeof(unsigned int) +
           kClMaxSegLen * kClRowPitch * sizeof(int2) + 2 * kClBatch * sizeof(int);
}

template <int C, bool GRAD, bool EMIT>
__global__ void __launch_bounds__(32 * (GRAD ? kClMaxWarps : kClEnergyWarps), 1)
cl_pair_kernel(DeviceSystem S, CellArrays A, ClPairArgs P) {
    using F = ClFmt<C>;
    constexpr int K = F::kK;
    constexpr int R = kClRing;
    constexpr unsigned int FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const int TT = S.n_types * S.n_types;
    PairConst* s_table = reinterpret_cast<PairConst*>(smem_raw);
    unsigned char* wbase = smem_raw + (P.table_in_smem ? (((size_t)TT * sizeof(PairConst) + 15) & ~(size_t)15) : 0) +
                           (size_t)warp * cl_warp_smem(GRAD);
    double* acc = reinterpret_cast<double*>(wbase);                                      // [3][kClMaxRows * R] (GRAD)
    unsigned int* ring = reinterpret_cast<unsigned int*>(wbase + (GRAD ? (size_t)3 * kClMaxRows * R * sizeof(double) : 0));
    int2* rowinfo = reinterpret_cast<int2*>(ring + kRingSlots * kChunk);                 // [cell of the segment][row]: {window start, window start - ring origin}
    int* s_cnt = reinterpret_cast<int*>(rowinfo + kClMaxSegLen * kClRowPitch);           // list length of the batch's clusters
    int* s_ci = s_cnt + kClBatch;                                                        // their cell (index inside the segment)
    if (P.table_in_smem) {
        for (int t = threadIdx.x; t < TT; t += blockDim.x) s_table[t] = S.table[t];
    }
    if (GRAD) {
        for (int t = lane; t < 3 * kClMaxRows * R; t += 32) acc[t] = 0.0;                // flushes leave it zero again
    }
    __syncthreads();
    const PairConst* table = P.table_in_smem ? s_table : S.table;

    const GridParams g = *A.grid;
    const int H = g.n_half;
    const PosQ* __restrict__ spq = A.spq;
    const long long s0 = (long long)g.n_seg * P.rank / P.world, s1 = (long long)g.n_seg * (P.rank + 1) / P.world;
    constexpr int kAccPlane = kClMaxRows * R;                     // doubles between the x, y and z planes

    for (;;) {
        long long seg = 0;
        if (lane == 0) seg = s0 + atomicAdd(A.work_ctr + 1, 1);
        seg = __shfl_sync(FULL, seg, 0);
        if (seg >= s1) break;
        const int row = (int)(seg / g.nsx), sx = (int)(seg % g.nsx);
        const int xa = sx * g.seg_len, xb = min(xa + g.seg_len, g.nx);
        const int cy = row % g.ny, cz = row / g.ny;
        const int rowbase = row * g.nx;
        const int q0 = A.start[rowbase + xa] / C, q1 = A.start[rowbase + xb] / C;      // the segment's clusters
        double ev = 0, ee = 0;
        unsigned long long n_cand_w = 0;
        unsigned int n_eval = 0;
        if (q0 < q1) {
            // lane o < H looks after half-shell row o: its cell row, the ring origin (sorted position of ring slot 0)
            // and the flush frontier (first position not yet written to the slot array)
            int my_rowbase = -1, my_p0 = 0, my_lo = 0;
            if (lane < H) {
                int dy, dz;
                cl_row_offset(lane, g.reach, dy, dz);
                const int y2 = cy + dy, z2 = cz + dz;
                if (y2 >= 0 && y2 < g.ny && z2 < g.nz) {
                    my_rowbase = (z2 * g.ny + y2) * g.nx;
                    my_p0 = my_lo = A.start[my_rowbase + max(xa - g.reach, 0)];
                }
            }
            __syncwarp();
            for (int ci = 0; ci < xb - xa; ++ci) {
                int2 v = make_int2(0, 0);
                if (my_rowbase >= 0) {
                    v.x = A.start[my_rowbase + max(xa + ci - g.reach, 0)];
                    v.y = v.x - my_p0;
                }
                if (lane < kClRowPitch) rowinfo[ci * kClRowPitch + lane] = v;
            }
            __syncwarp();
            const int slot_par = sx & 1;
            // flush the window rows up to the given frontier (per row: `target`, held by lane o)
            auto flush_to = [&](int target) {
                for (int o = 0; o < (GRAD ? H : 0); ++o) {
                    const int rb = __shfl_sync(FULL, my_rowbase, o);
                    const int lo = __shfl_sync(FULL, my_lo, o);
                    const int hi = __shfl_sync(FULL, target, o);
                    const int p0 = __shfl_sync(FULL, my_p0, o);
                    if (rb < 0) continue;
                    double* dst = P.gj + (size_t)(2 * o + slot_par) * 3 * P.n_spad;
                    for (int pos = lo + lane; pos < hi; pos += 32) {
                        const bool in_ring = pos - lo < R;            // positions beyond the capacity were never accumulated here
                        const int idx = o * R + ((pos - p0) & (R - 1));
                        double vx = 0, vy = 0, vz = 0;
                        if (in_ring) {
                            vx = acc[idx]; vy = acc[kAccPlane + idx]; vz = acc[2 * kAccPlane + idx];
                            acc[idx] = 0.0; acc[kAccPlane + idx] = 0.0; acc[2 * kAccPlane + idx] = 0.0;
                        }
                        dst[pos] = vx;
                        dst[P.n_spad + pos] = vy;
                        dst[2 * (size_t)P.n_spad + pos] = vz;
                    }
                    if (lane == o) my_lo = max(my_lo, hi);
                }
                __syncwarp();
            };
            int cur_ci = 0;                                       // the windows stand at the segment's first cell

            for (int qb = q0; qb < q1; qb += kClBatch) {
                const int nq = min(kClBatch, q1 - qb);
                __syncwarp();                                     // the previous batch's reads of the ring / s_cnt are over
                s_cnt[lane] = lane < nq ? min(P.count[qb + lane], P.cap) : 0;
                s_ci[lane] = lane < nq ? A.scell[C * (qb + lane)] - rowbase - xa : 0;
                __syncwarp();

                // chunk requests run over the same (cluster, chunk) sequence as the trips, two chunks ahead
                int q_il = 0, q_c = 0, q_slot = 0;
                auto request = [&]() {
                    if (q_il < nq) {
                        const unsigned int* src = P.list + ((size_t)(qb + q_il) * P.cap + kChunk * q_c + 4 * lane);
                        cp_async16(ring + q_slot * kChunk + 4 * lane, src);
                        q_slot = q_slot == kRingSlots - 1 ? 0 : q_slot + 1;
                        if (kChunk * (q_c + 1) < s_cnt[q_il]) ++q_c; else { ++q_il; q_c = 0; }
                    }
                    cp_async_commit();
                };
                request(); request(); request();
                cp_async_wait<2>();
                __syncwarp();

                int il = 0, b = 0, slot = 0, ci = 0, n_hits = 0, ipos0 = 0;
                double xi[C], yi[C], zi[C], qi2[C], gix[C], giy[C], giz[C];
                const PairConst* trow[C];
                PosQ nxt_p[C];
                int nxt_t[C];
#pragma unroll
                for (int a = 0; a < C; ++a) {
                    nxt_p[a] = load_posq(spq + (size_t)C * qb + a);
                    nxt_t[a] = A.stype[(size_t)C * qb + a];
                    xi[a] = yi[a] = zi[a] = qi2[a] = gix[a] = giy[a] = giz[a] = 0.0;
                    trow[a] = table;
                }

                auto gather = [&](ClNb<C>& n, int k, unsigned int e, int cell) {
                    n.e[k] = e;
                    const int2 info = rowinfo[cell * kClRowPitch + ((e >> F::kRowShift) & 15u)];
                    unsigned long long addr;
                    asm("mad.wide.u32 %0, %1, 32, %2;" : "=l"(addr) : "r"((unsigned int)info.x + (e & F::kRelMask)), "l"(spq));
                    n.p[k] = load_posq(reinterpret_cast<const PosQ*>(addr));
                };
                ClNb<C> NA, NB = {};
#pragma unroll
                for (int k = 0; k < K; ++k) gather(NA, k, 32 * k + lane < s_cnt[0] ? ring[32 * k + lane] : 0u, s_ci[0]);

                auto trip = [&](const ClNb<C>& p, ClNb<C>& nx) -> bool {
                    if (b == 0) {                             // a new cluster starts with this trip (warp-uniform)
                        n_hits = s_cnt[il];
                        ipos0 = C * (qb + il);
                        const int ci_new = s_ci[il];
#pragma unroll
                        for (int a = 0; a < C; ++a) {
                            xi[a] = nxt_p[a].x; yi[a] = nxt_p[a].y; zi[a] = nxt_p[a].z;
                            qi2[a] = 2.0 * S.ele_scale * nxt_p[a].q;             // doubled energies, see pair_terms
                            trow[a] = table + nxt_t[a] * S.n_types;
                            gix[a] = giy[a] = giz[a] = 0.0;
                        }
                        if (GRAD && ci_new != cur_ci) {       // the sweep moves on: cells that left the windows are flushed
                            int target = 0;
                            if (lane < kClRowPitch) target = rowinfo[ci_new * kClRowPitch + lane].x;
                            flush_to(target);
                            cur_ci = ci_new;
                        }
                        ci = ci_new;
                        if (il + 1 < nq) {                    // the next cluster's atoms, one cluster ahead
#pragma unroll
                            for (int a = 0; a < C; ++a) {
                                nxt_p[a] = load_posq(spq + (size_t)C * (qb + il + 1) + a);
                                nxt_t[a] = A.stype[(size_t)C * (qb + il + 1) + a];
                            }
                        }
                        n_cand_w += (unsigned long long)n_hits * C;
                    }
                    int jpos[K], ridx[K], rel[K];
                    double dx[K][C], dy[K][C], dz[K][C], r2[K][C];
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const unsigned int e = p.e[k];
                        const int o = (int)((e >> F::kRowShift) & 15u);
                        const int2 info = rowinfo[ci * kClRowPitch + o];
                        rel[k] = (int)(e & F::kRelMask);
                        jpos[k] = info.x + rel[k];
                        ridx[k] = o * R + ((info.y + rel[k]) & (R - 1));
#pragma unroll
                        for (int a = 0; a < C; ++a) {
                            dx[k][a] = xi[a] - p.p[k].x; dy[k][a] = yi[a] - p.p[k].y; dz[k][a] = zi[a] - p.p[k].z;
                            // strict double, left-to-right, no FMA: Point3d.distance squared
                            r2[k][a] = __dadd_rn(__dadd_rn(__dmul_rn(dx[k][a], dx[k][a]), __dmul_rn(dy[k][a], dy[k][a])),
                                                 __dmul_rn(dz[k][a], dz[k][a]));
                        }
                    }
                    // next position of the stream; entering a chunk = its copy has landed, the slot two behind is reused
                    const bool last = b + F::kSpan >= n_hits;
                    int il2 = il, b2 = b + F::kSpan, slot2 = slot;
                    if (last) { il2 = il + 1; b2 = 0; }
                    if (last || (b2 & (kChunk - 1)) == 0) {
                        slot2 = slot == kRingSlots - 1 ? 0 : slot + 1;
                        __syncwarp();
                        request();
                        cp_async_wait<2>();
                        __syncwarp();
                    }
                    if (il2 < nq) {
                        const int hits2 = last ? s_cnt[il2] : n_hits;
                        const int ci2 = last ? s_ci[il2] : ci;
                        const unsigned int* src = ring + slot2 * kChunk + (b2 & (kChunk - 1)) + lane;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            unsigned int e = b2 + 32 * k + lane < hits2 ? src[32 * k] : 0u;
                            // the address carries a (never taken, for non-negative r2) dependence on r2, so that this load
                            // is issued after the wait for this trip's operands and not before it.  A NaN r2 with the sign
                            // bit set moves the index by one row bit at most: it stays inside the row table, and the
                            // result is NaN anyway
                            e ^= ((unsigned int)__double2hiint(r2[k][0]) >> 31) << F::kRowShift;
                            gather(nx, k, e, ci2);
                        }
                    }
                    // arithmetic: K x C independent chains.  A pair that does not count (beyond the cutoff, excluded, not
                    // later in the sorted order, padding) is evaluated at r2 = 1 with zero prefactors: exact zeros
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const unsigned int e = p.e[k];
                        const bool valid = b + 32 * k + lane < n_hits;
                        const int tj = (int)((e >> F::kTypeShift) & 127u);
                        double gjx = 0, gjy = 0, gjz = 0;
#pragma unroll
                        for (int a = 0; a < C; ++a) {
                            const bool on = valid && jpos[k] > ipos0 + a && !(r2[k][a] > S.r2_max) && !((e >> (F::kExclShift + a)) & 1u);
                            const bool f14 = (e >> (F::kF14Shift + a)) & 1u;
                            if (on) ++n_eval;
                            if (EMIT) {
                                if (on) {
                                    const int oi = A.orig[ipos0 + a], oj = A.orig[jpos[k]];
                                    const unsigned long long q = atomicAdd(P.cursor, 1ull);
                                    if ((long long)q < P.capacity) { P.out_i[q] = min(oi, oj); P.out_j[q] = max(oi, oj); P.out_14[q] = f14; }
                                }
                            } else {
                                PairConst pc = trow[a][tj];
                                pc.k1 = on ? pc.k1 : 0.0;
                                const double qq2 = on ? (qi2[a] * p.p[k].q) * (f14 ? kEle14 : 1.0) : 0.0;
                                double e1, e2v, f;
                                pair_terms<GRAD>(on ? clamp_r2(r2[k][a]) : 1.0, pc, qq2, e1, e2v, f);
                                ev += e1;
                                ee += e2v;
                                if (GRAD) {
                                    gix[a] = fma(f, dx[k][a], gix[a]);
                                    giy[a] = fma(f, dy[k][a], giy[a]);
                                    giz[a] = fma(f, dz[k][a], giz[a]);
                                    gjx = fma(-f, dx[k][a], gjx);
                                    gjy = fma(-f, dy[k][a], gjy);
                                    gjz = fma(-f, dz[k][a], gjz);
                                }
                            }
                        }
                        if (GRAD && !EMIT && valid) {
                            if (rel[k] < R) {                     // inside the window: plain read-modify-write, this warp is the only writer
                                acc[ridx[k]] += gjx;
                                acc[kAccPlane + ridx[k]] += gjy;
                                acc[2 * kAccPlane + ridx[k]] += gjz;
                            } else {                              // a row window denser than its capacity: atomics (order not fixed)
                                atomicAdd(P.gj_ovf + jpos[k], gjx);
                                atomicAdd(P.gj_ovf + P.n_spad + jpos[k], gjy);
                                atomicAdd(P.gj_ovf + 2 * (size_t)P.n_spad + jpos[k], gjz);
                                A.flags[2] = 1;
                            }
                        }
                    }
                    __syncwarp();                                 // window updates of this trip before those of the next
                    if (last && GRAD && !EMIT) {                  // the cluster is complete: reduce its i-side sums over the warp
                        constexpr int NV = C == 1 ? 4 : (C == 2 ? 8 : 16);
                        double v[NV];
#pragma unroll
                        for (int i = 0; i < NV; ++i) v[i] = 0.0;
#pragma unroll
                        for (int a = 0; a < C; ++a) { v[3 * a] = gix[a]; v[3 * a + 1] = giy[a]; v[3 * a + 2] = giz[a]; }
                        warp_multi_sum<NV>(v, lane);
                        const int vi = lane / (32 / NV);          // value completed in this lane
                        if ((lane & (32 / NV - 1)) == 0 && vi < 3 * C)
                            P.gi[(size_t)(vi % 3) * P.n_spad + ipos0 + vi / 3] = v[0];
                    }
                    il = il2; b = b2; slot = slot2;
                    return il < nq;
                };
                while (true) {
                    if (!trip(NA, NB)) break;
                    if (!trip(NB, NA)) break;
                }
                cp_async_wait<0>();
            }
            {   // the segment is done: everything still in the windows goes out
                int target = 0;
                if (my_rowbase >= 0) target = A.start[my_rowbase + min(xb - 1 + g.reach, g.nx - 1) + 1];
                flush_to(target);
            }
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
    (void)n_warps;
}

// ---- final reduction of the cluster path ------------------------------------------------------------------
struct ClReduceArgs {
    int n_spad;
    int rank, world;
    const double* gi;
    const double* gj;
    const double* gj_ovf;
    ReduceArgs R;                  // bonded slots, energy partials, units, outputs (gpart / direct unused)
};

__global__ void __launch_bounds__(kReduceWarps * 32)
cl_reduce_kernel(CellArrays A, ClReduceArgs Q) {
    const GridParams& g = *A.grid;
    if (blockIdx.x == gridDim.x - 1) {
        // last block: energies and pair counts over this shard's segments
        ReduceArgs R = Q.R;
        const long long s0 = (long long)g.n_seg * Q.rank / Q.world, s1 = (long long)g.n_seg * (Q.rank + 1) / Q.world;
        R.epart += 2 * s0; R.cpart += 2 * s0; R.n_epart = (int)(s1 - s0);
        R.overflow = A.flags;
        reduce_energy_block(R);
        return;
    }
    if (!Q.R.want_grad) return;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.start[g.ncell]) return;
    const int a = A.orig[p];
    if (a < 0) return;                                            // padding
    const int c = A.scell[p];
    const int cx = c % g.nx, cy = (c / g.nx) % g.ny, cz = c / (g.nx * g.ny);
    const long long s0 = (long long)g.n_seg * Q.rank / Q.world, s1 = (long long)g.n_seg * (Q.rank + 1) / Q.world;
    double gx = 0, gy = 0, gz = 0;
    {   // i-side: written by the segment that holds the atom's cell
        const long long seg = (long long)(cz * g.ny + cy) * g.nsx + cx / g.seg_len;
        if (seg >= s0 && seg < s1) {
            gx = Q.gi[p]; gy = Q.gi[Q.n_spad + p]; gz = Q.gi[2 * (size_t)Q.n_spad + p];
        }
    }
    // j-side: for every row o of the half shell, the segments of the i-row (cy - dy, cz - dz) whose windows covered
    // this cell - cells [xa - reach, xb - 1 + reach] - wrote slot 2 o + (segment & 1), provided they hold any atom
    for (int o = 0; o < g.n_half; ++o) {
        int dy, dz;
        cl_row_offset(o, g.reach, dy, dz);
        const int yi = cy - dy, zi = cz - dz;
        if (yi < 0 || yi >= g.ny || zi < 0) continue;
        const int irow = zi * g.ny + yi;
        const int sa = max(cx - g.reach, 0) / g.seg_len, sb = min((cx + g.reach) / g.seg_len, g.nsx - 1);
        for (int sx = sa; sx <= sb; ++sx) {
            const long long seg = (long long)irow * g.nsx + sx;
            if (seg < s0 || seg >= s1) continue;
            const int xa = sx * g.seg_len, xb = min(xa + g.seg_len, g.nx);
            if (A.start[irow * g.nx + xb] == A.start[irow * g.nx + xa]) continue;       // empty segment: wrote nothing
            const double* src = Q.gj + (size_t)(2 * o + (sx & 1)) * 3 * Q.n_spad;
            gx += src[p]; gy += src[Q.n_spad + p]; gz += src[2 * (size_t)Q.n_spad + p];
        }
    }
    if (A.flags[2]) {
        gx += Q.gj_ovf[p]; gy += Q.gj_ovf[Q.n_spad + p]; gz += Q.gj_ovf[2 * (size_t)Q.n_spad + p];
    }
    const ReduceArgs& R = Q.R;
    for (long long e = R.slot_ptr[a]; e < R.slot_ptr[a + 1]; ++e) {
        const long long s = R.slot_idx[e];
        if (s >= R.slot_lo && s < R.slot_hi) {
            gx += R.slots[3 * s];
            gy += R.slots[3 * s + 1];
            gz += R.slots[3 * s + 2];
        }
    }
    double* o = R.out + kOutHeader + 3 * (size_t)a;
    o[0] = convert_unit(gx, R.unit_num, R.unit_den);
    o[1] = convert_unit(gy, R.unit_num, R.unit_den);
    o[2] = convert_unit(gz, R.unit_num, R.unit_den);
}

// ---- host side ---------------------------------------------------------------------------------------
// launch geometry of the cluster path: warps per CTA of the gradient kernel from what the windows leave of the
// shared memory; a type-pair table too large to sit next to them is read through L1 / L2 instead
inline int cl_configure(CellPlan& p, const DeviceSystem& ds, std::string& err) {
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t table = (((size_t)ds.n_types * ds.n_types * sizeof(PairConst)) + 15) & ~(size_t)15;
    p.cl_table_in_smem = table <= 64 * 1024 ? 1 : 0;
    const size_t tb = p.cl_table_in_smem ? table : 0;
    p.cl_warps = (int)std::min<size_t>(kClMaxWarps, ((size_t)max_smem - tb) / cl_warp_smem(true));
    if (p.cl_warps < 1) { err = "cell path: not enough shared memory for one force window"; return JME_ERR_UNSUPPORTED; }
    p.cl_smem_grad = tb + p.cl_warps * cl_warp_smem(true);
    p.cl_smem_energy = tb + kClEnergyWarps * cl_warp_smem(false);
    // about four segments per resident warp of the gradient kernel
    p.arr.seg_target = 4 * p.sms * p.cl_warps;
    return JME_OK;
}

template <int C>
inline int cl_run_lists_t(CellPlan& p, const DeviceSystem& ds, cudaStream_t st, uint64_t& launches, std::string& err) {
    ClListArgs A{};
    A.list = p.d_list; A.count = p.d_count; A.cap = p.list_cap; A.rank = p.rank; A.world = p.world;
    const long long resident = (long long)p.grid_list * kListWarps;
    A.parts = (int)std::max(1ll, std::min(8ll, 2 * resident * 12 / std::max(1, ds.n)));
    cudaMemsetAsync(p.arr.work_ctr, 0, 2 * sizeof(int), st);
    cl_list_kernel<C><<<p.grid_list, kListThreads, 0, st>>>(p.arr, A);
    ++launches;
    if (p.arr.n_excl_ent > 0) {
        cl_mark_kernel<C><<<(unsigned int)((p.arr.n_excl_ent + 255) / 256), 256, 0, st>>>(p.arr, p.d_excl_src, p.d_list, p.d_count,
                                                                                        p.list_cap, p.rank, p.world);
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
    P.list = p.d_list; P.count = p.d_count; P.cap = p.list_cap;
    P.rank = p.rank; P.world = p.world; P.n_spad = p.n_spad + 1;
    P.gi = p.d_gi; P.gj = p.d_gj; P.gj_ovf = p.d_gj_ovf;
    P.epart = p.d_epart; P.cpart = p.d_cpart;
    P.table_in_smem = p.cl_table_in_smem;
}

template <int C, bool GRAD, bool EMIT>
inline int cl_launch_pairs(CellPlan& p, const DeviceSystem& ds, const ClPairArgs& P, cudaStream_t st, std::string& err) {
    auto k = cl_pair_kernel<C, GRAD, EMIT>;
    const size_t smem = GRAD ? p.cl_smem_grad : p.cl_smem_energy;
    const int warps = GRAD ? p.cl_warps : kClEnergyWarps;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { err = std::string("cell path: shared memory request refused: ") + cudaGetErrorString(e); return JME_ERR_CUDA; }
    }
    const int grid = GRAD ? p.sms : 2 * p.sms;
    k<<<grid, 32 * warps, smem, st>>>(ds, p.arr, P);
    JME_CELL_CHECK();
    return JME_OK;
}

inline int cl_run_pairs(CellPlan& p, const DeviceSystem& ds, bool grad, cudaStream_t st, uint64_t& launches, std::string& err) {
    if (ds.n == 0) return JME_OK;
    ClPairArgs P{};
    cl_fill_pair_args(p, P);
    cudaMemsetAsync(p.arr.work_ctr + 1, 0, sizeof(int), st);
    if (grad) cudaMemsetAsync(p.d_gj_ovf, 0, 3 * (size_t)(p.n_spad + 1) * sizeof(double), st);
    ++launches;
    if (p.cluster == 2) return grad ? cl_launch_pairs<2, true, false>(p, ds, P, st, err) : cl_launch_pairs<2, false, false>(p, ds, P, st, err);
    return grad ? cl_launch_pairs<4, true, false>(p, ds, P, st, err) : cl_launch_pairs<4, false, false>(p, ds, P, st, err);
}

// final reduction (gradient: one thread per sorted position + the energy block; energy only: the energy block)
inline int cl_run_reduce(CellPlan& p, const DeviceSystem& ds, const ReduceArgs& R, cudaStream_t st, uint64_t& launches, std::string& err) {
    ClReduceArgs Q{};
    Q.n_spad = p.n_spad + 1; Q.rank = p.rank; Q.world = p.world;
    Q.gi = p.d_gi; Q.gj = p.d_gj; Q.gj_ovf = p.d_gj_ovf;
    Q.R = R;
    const int nt = kReduceWarps * 32;
    const int grid = (R.want_grad ? (p.n_spad + nt - 1) / nt : 0) + 1;
    cl_reduce_kernel<<<grid, nt, 0, st>>>(p.arr, Q);
    ++launches;
    JME_CELL_CHECK();
    (void)ds;
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
