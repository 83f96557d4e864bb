// jme_kernels.cuh - CUDA kernels of the MMFF94 energy/gradient path for sm_100a (B200).
//
// Nothing here is a dense contraction, so tensor cores are deliberately unused: the nonbonded
// kernels are bound by the FP64 FMA pipe (about 58 FP64 instructions per pair), everything else by
// HBM/L2 bandwidth or launch latency.  Design notes per kernel are in DESIGN.md.
//
//   ap_kernel        all-pairs tiles: one warp = one 32-atom i-block x a run of 32-atom j-tiles,
//                    j-tile staged in shared memory, Newton's third law inside every tile with the
//                    j-accumulators rotating through the warp, exclusions / 1-4 flags as 32x32 bit
//                    masks, deterministic (every partial gradient has exactly one writer).
//   bonded_kernel    one thread per bonded term, gradients to per-term slots.
//   reduce_kernel    per-atom fixed-order sum of the tile partials and bonded slots (CSR gather),
//                    unit conversion, energy/pair-count tree reduction.
//   cells_*          cell-list path for cutoff > 0 (jme_cells.cuh).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "jme_math.cuh"

namespace jme {

constexpr int kTile = 32;
constexpr int kApWarps = 4;                    // warps per CTA in ap_kernel (2 CTAs/SM: 2 warps per scheduler)
constexpr int kApThreads = kApWarps * 32;
constexpr int kReduceWarps = 8;
constexpr int kBondedThreads = 128;

struct DeviceSystem {
    int n;                 // atoms
    int n_pad;             // padded to a multiple of 128 (kSuper)
    int n_blocks;          // n_pad / 32 (n_pad is a multiple of 128: whole superblocks)
    int n_types;
    const double* x;       // [n_pad] SoA coordinates (padding = 0)
    const double* y;
    const double* z;
    const double* q;       // [n_pad] charges (padding = 0)
    const int* type;       // [n_pad] dense type index (padding = 0)
    const PairConst* table;    // [n_types*n_types]
    double r2_max;         // cutoff threshold on r^2 (jme_host.hpp cutoff_r2_threshold); unused without cutoff
    double ele_scale;      // 332.0716 / dielectric
};

// ---- all-pairs decomposition ---------------------------------------------------------------------
// Atoms are grouped in 32-atom blocks and kNI consecutive blocks form a SUPERBLOCK.  A warp owns one
// superblock S as its i-side: lane l holds the kNI atoms  (S*kNI + a)*32 + l,  a = 0..kNI-1  ("chains").
// It meets 32-atom j-tiles one after the other; every j atom is loaded once per step and paired with all
// kNI i atoms of the lane - kNI independent FP64 dependency chains per lane, which is what the FP64 pipe
// needs (measured on B200, tools/fp64_latency.cu: 8.3 cycles dependent latency; 2.25 cycles per warp
// instruction with >= 4 chains in ONE warp, but 3.1 when the same parallelism comes from 8 warps).
// The j-tile sequence of row S is  q = 0 .. (H(S)+1)*kNI - 1 :  d = q / kNI, b = q % kNI,
// j-block = ((S + d) mod nS)*kNI + b  (circulant over superblocks: every superblock pair exactly once);
// d = 0 is the superblock against itself, where chain a meets tile b only if a <= b and the tile a == b
// keeps only j > i.
constexpr int kNI = 4;
constexpr int kSuper = kNI * kTile;          // atoms per superblock (128)

struct ApItem {
    int S, q0, nq, tile_ofs;                 // tiles q0 .. q0+nq-1 of row S; tile_ofs indexes tile_mask[]
};

struct TileMask {
    uint32_t excl[kNI][32];     // [chain][i-lane], bit = j index in the tile: excluded (1-2, 1-3, padding)
    uint32_t f14[kNI][32];      // 1-4 pair (electrostatics x0.75)
};

__device__ __forceinline__ double shfl_rot1(double v, int src_lane) {
    return __shfl_sync(0xffffffffu, v, src_lane);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// all-pairs tiles
// ---------------------------------------------------------------------------------------------
// gpart layout: [slot][3][n_pad].  Slot S (< n_super) receives the j-side gradients produced by the items of
// superblock row S; slot n_super + c receives the i-side gradients of chunk c of every row.  Each
// (slot, atom) has exactly one writer, so the later fixed-order sum is deterministic.

// r2 with its exponent clamped from below (integer max on the high word): coincident atoms (r2 == 0) keep
// their finite buffered energy and get no gradient (f * dx = 0), without a NaN from 1/sqrt(0)
__device__ __forceinline__ double clamp_r2(double r2) {
    return __hiloint2double(max(__double2hiint(r2), 0x01a00000), __double2loint(r2));
}

struct ApLaneState {            // the kNI i-atoms of a lane
    double xi[kNI], yi[kNI], zi[kNI], qi2[kNI];
    double gix[kNI], giy[kNI], giz[kNI];
    const PairConst* trow[kNI];
};

// One 32-atom j-tile against the kNI i-chains of every lane.  At step k lane l meets j-atom (l + k) & 31;
// the j-side accumulator travels with the j atom from lane l+1 to lane l between steps (warp shuffle), so
// Newton's third law costs no atomics and no scatter.
//   MASKED = false : every pair of the tile interacts with every chain - no predicates, no selects
//   MASKED = true  : per chain: active flag, "self" rule (tile == own block: only j > i), exclusion / 1-4 /
//                    padding bit masks
template <bool GRAD, bool CUTOFF, bool MASKED>
__device__ __forceinline__ void ap_tile(ApLaneState& L, const double* __restrict__ sx, const double* __restrict__ sy,
                                        const double* __restrict__ sz, const double* __restrict__ sq,
                                        const int* __restrict__ st, const int lane, const int k_begin, const int k_end,
                                        const int active_chains,
                                        const int self_chain, const uint32_t (&m_excl)[kNI], const uint32_t (&m_14)[kNI],
                                        const double r2_max, double& gjx, double& gjy, double& gjz,
                                        double& ev, double& ee, unsigned int& n_eval, unsigned int& n_cand) {
    const int src = (lane + 1) & 31;
#pragma unroll 1
    for (int k = k_begin; k < k_end; ++k) {
        if (GRAD && k != k_begin) {
            gjx = __shfl_sync(0xffffffffu, gjx, src);
            gjy = __shfl_sync(0xffffffffu, gjy, src);
            gjz = __shfl_sync(0xffffffffu, gjz, src);
        }
        const int jj = (lane + k) & 31;
        const double xj = sx[jj], yj = sy[jj], zj = sz[jj], qj = sq[jj];
        const int tj = st[jj];
#pragma unroll
        for (int a = 0; a < kNI; ++a) {
            const double dx = L.xi[a] - xj, dy = L.yi[a] - yj, dz = L.zi[a] - zj;
            double r2;
            if (CUTOFF) {
                // strict double, left-to-right, no FMA: the r2 the reference's Point3d.distance squares
                r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            } else {
                r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            }
            double qq2 = L.qi2[a] * qj;
            bool on = true;
            if (MASKED) {
                on = (a < active_chains) && !((m_excl[a] >> jj) & 1u);
                if (a == self_chain) on = on && (jj > lane);
                n_cand += __popc(__ballot_sync(0xffffffffu, on));
                if ((m_14[a] >> jj) & 1u) qq2 *= kEle14;
            }
            if (CUTOFF) {
                on = on && !(r2 > r2_max);
                n_eval += __popc(__ballot_sync(0xffffffffu, on));
            }
            double e1, e2, f;
            pair_terms<GRAD>((MASKED || CUTOFF) ? (on ? clamp_r2(r2) : 1.0) : clamp_r2(r2), L.trow[a][tj], qq2, e1, e2, f);
            if (MASKED || CUTOFF) {
                e1 = on ? e1 : 0.0;
                e2 = on ? e2 : 0.0;
                f = on ? f : 0.0;
            }
            ev += e1;
            ee += e2;
            if (GRAD) {
                L.gix[a] = fma(f, dx, L.gix[a]);
                L.giy[a] = fma(f, dy, L.giy[a]);
                L.giz[a] = fma(f, dz, L.giz[a]);
                gjx = fma(-f, dx, gjx);
                gjy = fma(-f, dy, gjy);
                gjz = fma(-f, dz, gjz);
            }
        }
    }
    if (!MASKED) n_cand += 32 * (k_end - k_begin) * kNI;
}

// KSPLIT = 1: every warp of the CTA works on its own item (large systems: plenty of items).
// KSPLIT = 4: the four warps of the CTA share ONE item and each takes a quarter of the 32 steps of every tile
//             (small systems: 4x more parallelism per tile); the partial sums of the four warps are combined in
//             shared memory in a fixed order, so the outputs are the same buffers as for KSPLIT = 1.
// TSM: the type-pair table is staged in shared memory; a table too large for that (about 40+ atom types) is read through
// the cache hierarchy instead
template <bool GRAD, bool CUTOFF, int KSPLIT, bool TSM>
__global__ void __launch_bounds__(kApThreads, 2)
ap_kernel(DeviceSystem S, const ApItem* __restrict__ items, int item_begin, int item_end,
          const int* __restrict__ tile_mask, const TileMask* __restrict__ masks, int n_super,
          double* __restrict__ gpart, int tiles_per_item,
          double* __restrict__ epart, unsigned long long* __restrict__ cpart) {
    static_assert(KSPLIT == 1 || KSPLIT == kApWarps, "KSPLIT must be 1 or the number of warps per CTA");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // shared: pair table, per-warp j staging (x, y, z, q: 4 x 32 doubles; type: 32 ints), KSPLIT combine area
    PairConst* s_table = reinterpret_cast<PairConst*>(smem_raw);
    const int TT = TSM ? S.n_types * S.n_types : 0;
    double* s_stage = reinterpret_cast<double*>(s_table + TT);
    double* s_comb = s_stage + kApWarps * (4 * 32 + 16);       // [warps][kNI*3 + 3][32] (KSPLIT > 1 only)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sx = s_stage + warp * (4 * 32 + 16);
    double* sy = sx + 32;
    double* sz = sy + 32;
    double* sq = sz + 32;
    int* st = reinterpret_cast<int*>(sq + 32);

    for (int t = threadIdx.x; t < TT; t += kApThreads) s_table[t] = S.table[t];
    __syncthreads();

    const int item_idx = item_begin + (KSPLIT == 1 ? blockIdx.x * kApWarps + warp : (int)blockIdx.x);
    if (item_idx >= item_end) return;           // CTA-uniform when KSPLIT > 1
    const ApItem it = items[item_idx];
    const int k_begin = KSPLIT == 1 ? 0 : warp * (32 / kApWarps);
    const int k_end = KSPLIT == 1 ? 32 : k_begin + 32 / kApWarps;

    ApLaneState L;
#pragma unroll
    for (int a = 0; a < kNI; ++a) {
        const int ia = (it.S * kNI + a) * kTile + lane;
        L.xi[a] = S.x[ia]; L.yi[a] = S.y[ia]; L.zi[a] = S.z[ia];
        L.qi2[a] = 2.0 * S.ele_scale * S.q[ia];             // doubled energies, see pair_terms
        L.trow[a] = (TSM ? s_table : S.table) + S.type[ia] * S.n_types;
        L.gix[a] = L.giy[a] = L.giz[a] = 0.0;
    }
    double ev = 0, ee = 0;
    unsigned int n_eval = 0, n_cand = 0;

    auto j_block = [&](int q) {
        int T = it.S + q / kNI;
        if (T >= n_super) T -= n_super;
        return T * kNI + q % kNI;
    };
    // software pipeline: the next tile's atoms are loaded into registers while this tile computes
    int Jn = j_block(it.q0);
    double px = S.x[Jn * kTile + lane], py = S.y[Jn * kTile + lane], pz = S.z[Jn * kTile + lane];
    double pq = S.q[Jn * kTile + lane];
    int pt = S.type[Jn * kTile + lane];
    int pm = tile_mask[it.tile_ofs];

    for (int t = 0; t < it.nq; ++t) {
        const int q = it.q0 + t;
        const int J = Jn;
        __syncwarp();
        sx[lane] = px;
        sy[lane] = py;
        sz[lane] = pz;
        sq[lane] = pq;
        st[lane] = pt;
        const int mi = pm;
        __syncwarp();
        if (t + 1 < it.nq) {
            Jn = j_block(q + 1);
            const int na = Jn * kTile + lane;
            px = S.x[na]; py = S.y[na]; pz = S.z[na]; pq = S.q[na]; pt = S.type[na];
            pm = tile_mask[it.tile_ofs + t + 1];
        }

        const bool self = q < kNI;                 // d == 0: the superblock against itself
        double gjx = 0, gjy = 0, gjz = 0;
        if (self || mi >= 0) {
            uint32_t m_excl[kNI], m_14[kNI];
#pragma unroll
            for (int a = 0; a < kNI; ++a) {
                m_excl[a] = mi >= 0 ? masks[mi].excl[a][lane] : 0u;
                m_14[a] = mi >= 0 ? masks[mi].f14[a][lane] : 0u;
            }
            // self tiles: tile b = q meets chains a <= b only, chain a == b keeps j > i
            ap_tile<GRAD, CUTOFF, true>(L, sx, sy, sz, sq, st, lane, k_begin, k_end, self ? q + 1 : kNI, self ? q : -1,
                                        m_excl, m_14, S.r2_max, gjx, gjy, gjz, ev, ee, n_eval, n_cand);
        } else {
            const uint32_t zero[kNI] = {};
            ap_tile<GRAD, CUTOFF, false>(L, sx, sy, sz, sq, st, lane, k_begin, k_end, kNI, -1, zero, zero, S.r2_max,
                                         gjx, gjy, gjz, ev, ee, n_eval, n_cand);
        }
        if (GRAD) {
            // after its last step the accumulator in lane l belongs to tile atom (l + k_end - 1) & 31
            const int jl = (lane + k_end - 1) & 31;
            double* gj = gpart + (size_t)it.S * 3 * S.n_pad + J * kTile;
            if (KSPLIT == 1) {
                gj[jl] = gjx;
                gj[S.n_pad + jl] = gjy;
                gj[2 * S.n_pad + jl] = gjz;
            } else {
                double* c = s_comb + warp * ((kNI * 3 + 3) * 32);
                __syncthreads();                    // previous tile's combine has been consumed
                c[jl] = gjx; c[32 + jl] = gjy; c[64 + jl] = gjz;
                __syncthreads();
                if (warp == 0) {
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        double v = 0;
#pragma unroll
                        for (int w = 0; w < kApWarps; ++w) v += s_comb[w * ((kNI * 3 + 3) * 32) + cc * 32 + lane];
                        gj[cc * (size_t)S.n_pad + lane] = v;
                    }
                }
            }
        }
    }
    ev = warp_sum(ev);
    ee = warp_sum(ee);
    if (KSPLIT == 1) {
        if (GRAD) {
            const int chunk = it.q0 / tiles_per_item;
            double* gi = gpart + (size_t)(n_super + chunk) * 3 * S.n_pad;
#pragma unroll
            for (int a = 0; a < kNI; ++a) {
                const int ia = (it.S * kNI + a) * kTile + lane;
                gi[ia] = L.gix[a];
                gi[S.n_pad + ia] = L.giy[a];
                gi[2 * S.n_pad + ia] = L.giz[a];
            }
        }
        if (lane == 0) {
            // energies were accumulated doubled (pair_terms); ballots are warp-uniform: every lane holds the counts
            epart[2 * item_idx] = 0.5 * ev;
            epart[2 * item_idx + 1] = 0.5 * ee;
            cpart[2 * item_idx] = CUTOFF ? n_eval : n_cand;
            cpart[2 * item_idx + 1] = n_cand;
        }
    } else {
        double* c = s_comb + warp * ((kNI * 3 + 3) * 32);
        __syncthreads();
        if (GRAD) {
#pragma unroll
            for (int a = 0; a < kNI; ++a) {
                c[(3 * a) * 32 + lane] = L.gix[a];
                c[(3 * a + 1) * 32 + lane] = L.giy[a];
                c[(3 * a + 2) * 32 + lane] = L.giz[a];
            }
        }
        if (lane == 0) {
            c[kNI * 3 * 32] = ev;
            c[kNI * 3 * 32 + 1] = ee;
            c[kNI * 3 * 32 + 2] = (double)(CUTOFF ? n_eval : n_cand);
            c[kNI * 3 * 32 + 3] = (double)n_cand;
        }
        __syncthreads();
        if (warp == 0) {
            const int stride = (kNI * 3 + 3) * 32;
            if (GRAD) {
                const int chunk = it.q0 / tiles_per_item;
                double* gi = gpart + (size_t)(n_super + chunk) * 3 * S.n_pad;
#pragma unroll
                for (int a = 0; a < kNI; ++a) {
                    const int ia = (it.S * kNI + a) * kTile + lane;
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        double v = 0;
#pragma unroll
                        for (int w = 0; w < kApWarps; ++w) v += s_comb[w * stride + (3 * a + cc) * 32 + lane];
                        gi[cc * (size_t)S.n_pad + ia] = v;
                    }
                }
            }
            if (lane == 0) {
                double e0 = 0, e1 = 0, c0 = 0, c1 = 0;
                for (int w = 0; w < kApWarps; ++w) {
                    e0 += s_comb[w * stride + kNI * 3 * 32];
                    e1 += s_comb[w * stride + kNI * 3 * 32 + 1];
                    c0 += s_comb[w * stride + kNI * 3 * 32 + 2];
                    c1 += s_comb[w * stride + kNI * 3 * 32 + 3];
                }
                epart[2 * item_idx] = 0.5 * e0;
                epart[2 * item_idx + 1] = 0.5 * e1;
                cpart[2 * item_idx] = (unsigned long long)c0;
                cpart[2 * item_idx + 1] = (unsigned long long)c1;
            }
        }
    }
}

// Diagnostic twin of ap_kernel: same traversal and predicate, emits the selected pairs instead of
// evaluating them (jme_pair_list).  Order is arbitrary; callers sort.
template <bool CUTOFF>
__global__ void __launch_bounds__(kApThreads)
ap_pairlist_kernel(DeviceSystem S, const ApItem* __restrict__ items, int item_begin, int item_end,
                   const int* __restrict__ tile_mask, const TileMask* __restrict__ masks, int n_super,
                   int* __restrict__ out_i, int* __restrict__ out_j, unsigned char* __restrict__ out_14,
                   long long capacity, unsigned long long* __restrict__ cursor) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item_idx = item_begin + blockIdx.x * kApWarps + warp;
    if (item_idx >= item_end) return;
    const ApItem it = items[item_idx];
    for (int t = 0; t < it.nq; ++t) {
        const int q = it.q0 + t;
        int T = it.S + q / kNI;
        if (T >= n_super) T -= n_super;
        const int J = T * kNI + q % kNI;
        const int mi = tile_mask[it.tile_ofs + t];
        const bool self = q < kNI;
        for (int a = 0; a < kNI; ++a) {
            if (self && a > q) continue;
            const int ia = (it.S * kNI + a) * kTile + lane;
            const double xi = S.x[ia], yi = S.y[ia], zi = S.z[ia];
            const uint32_t m_excl = mi >= 0 ? masks[mi].excl[a][lane] : 0u;
            const uint32_t m_14 = mi >= 0 ? masks[mi].f14[a][lane] : 0u;
            for (int k = 0; k < 32; ++k) {
                const int jj = (lane + k) & 31;
                const int ja = J * kTile + jj;
                const double dx = xi - S.x[ja], dy = yi - S.y[ja], dz = zi - S.z[ja];
                const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                bool on = !((m_excl >> jj) & 1u);
                if (self && a == q) on = on && (jj > lane);
                if (CUTOFF) on = on && !(r2 > S.r2_max);
                if (on) {
                    const unsigned long long p = atomicAdd(cursor, 1ull);
                    if ((long long)p < capacity) {
                        out_i[p] = min(ia, ja);
                        out_j[p] = max(ia, ja);
                        out_14[p] = (m_14 >> jj) & 1u;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// bonded terms: one thread per term
// ---------------------------------------------------------------------------------------------
struct BondedTerms {
    int n_bond, n_angle, n_oop, n_tor;
    const int* bond_at;       // [n_bond][2]
    const double* bond_par;   // [n_bond][2]
    const int* angle_at;      // [n_angle][3]
    const double* angle_par;  // [n_angle][6]
    const unsigned char* angle_linear;
    const int* oop_at;        // [n_oop][4]
    const double* oop_par;    // [n_oop]
    const int* tor_at;        // [n_tor][4]
    const double* tor_par;    // [n_tor][3]
};

__device__ __forceinline__ Vec3 load_pos(const DeviceSystem& S, int a) { return Vec3{S.x[a], S.y[a], S.z[a]}; }
__device__ __forceinline__ void store_slot(double* slots, long long s, Vec3 g) {
    slots[3 * s] = g.x;
    slots[3 * s + 1] = g.y;
    slots[3 * s + 2] = g.z;
}

// Terms [term_begin, term_end) of the concatenated list bonds | angles | oops | torsions.
// Slot numbering: bond b -> 2b, 2b+1; angle a -> 2nb + 3a + {0,1,2}; oop o -> 2nb + 3na + 4o + {0..3};
// torsion t -> 2nb + 3na + 4no + 4t + {0..3}.  bpart: [gridDim.x][5] block partial energies.
template <bool GRAD>
__global__ void __launch_bounds__(kBondedThreads)
bonded_kernel(DeviceSystem S, BondedTerms B, long long term_begin, long long term_end,
              double* __restrict__ slots, double* __restrict__ bpart) {
    const long long t = term_begin + (long long)blockIdx.x * kBondedThreads + threadIdx.x;
    double e[5] = {0, 0, 0, 0, 0};
    if (t < term_end) {
        long long u = t;
        if (u < B.n_bond) {
            const int i = B.bond_at[2 * u], j = B.bond_at[2 * u + 1];
            Vec3 gi, gj;
            e[0] = bond_term(load_pos(S, i), load_pos(S, j), B.bond_par[2 * u], B.bond_par[2 * u + 1], gi, gj);
            if (GRAD) { store_slot(slots, 2 * u, gi); store_slot(slots, 2 * u + 1, gj); }
        } else if ((u -= B.n_bond) < B.n_angle) {
            const int i = B.angle_at[3 * u], j = B.angle_at[3 * u + 1], k = B.angle_at[3 * u + 2];
            const double* p = B.angle_par + 6 * u;
            Vec3 gi, gj, gk;
            angle_stbn_term(load_pos(S, i), load_pos(S, j), load_pos(S, k), p[0], p[1], B.angle_linear[u] != 0,
                            p[2], p[3], p[4], p[5], e[1], e[2], gi, gj, gk);
            if (GRAD) {
                const long long s = 2ll * B.n_bond + 3 * u;
                store_slot(slots, s, gi); store_slot(slots, s + 1, gj); store_slot(slots, s + 2, gk);
            }
        } else if ((u -= B.n_angle) < B.n_oop) {
            const int* at = B.oop_at + 4 * u;
            Vec3 gi, gj, gk, gl;
            e[3] = oop_term(load_pos(S, at[0]), load_pos(S, at[1]), load_pos(S, at[2]), load_pos(S, at[3]), B.oop_par[u], gi, gj, gk, gl);
            if (GRAD) {
                const long long s = 2ll * B.n_bond + 3ll * B.n_angle + 4 * u;
                store_slot(slots, s, gi); store_slot(slots, s + 1, gj); store_slot(slots, s + 2, gk); store_slot(slots, s + 3, gl);
            }
        } else {
            u -= B.n_oop;
            const int* at = B.tor_at + 4 * u;
            const double* p = B.tor_par + 3 * u;
            Vec3 gi, gj, gk, gl;
            e[4] = torsion_term(load_pos(S, at[0]), load_pos(S, at[1]), load_pos(S, at[2]), load_pos(S, at[3]), p[0], p[1], p[2], gi, gj, gk, gl);
            if (GRAD) {
                const long long s = 2ll * B.n_bond + 3ll * B.n_angle + 4ll * B.n_oop + 4 * u;
                store_slot(slots, s, gi); store_slot(slots, s + 1, gj); store_slot(slots, s + 2, gk); store_slot(slots, s + 3, gl);
            }
        }
    }
    __shared__ double s_e[kBondedThreads / 32][5];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const double v = warp_sum(e[c]);
        if (lane == 0) s_e[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double v = 0;
        for (int w = 0; w < kBondedThreads / 32; ++w) v += s_e[w][threadIdx.x];
        bpart[5 * blockIdx.x + threadIdx.x] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// final reduction
// ---------------------------------------------------------------------------------------------
struct ReduceArgs {
    int n, n_pad;
    const double* gpart;            // [n_slots][3][n_pad]
    int n_slots;
    const double* direct;           // optional [3][n_pad] gradient written directly per atom (cell path), or null
    const double* slots;            // bonded slots
    const long long* slot_ptr;      // [n+1] CSR atom -> bonded slot ids (ascending)
    const long long* slot_idx;
    long long slot_lo, slot_hi;     // only slots in [slot_lo, slot_hi) were written by this shard
    const double* epart; int n_epart;               // [n_epart][2] (vdw, ele)
    const unsigned long long* cpart;                // [n_epart][2] (evaluated, candidates)
    const double* bpart; int n_bpart;               // [n_bpart][5]
    double unit_num, unit_den;      // (v * num) / den, ForceField.java:157; both 1 for kcal/mol
    int want_grad;
    double* out;                    // [0] total [1..7] comps [8] pairs evaluated [9] candidates [10..] gradient AoS
    double* out_host;               // optional: mapped pinned host copy of the header (zero-copy result), kOutHeader + 1
                                    // doubles: [kOutHeader] = overflow flag
    const int* overflow;            // optional: the cell paths' device flags [4]: [0] a list was truncated, [1] non-finite
                                    // coordinate, [2] window fallback used (informational), [3] more clusters than list storage
};
constexpr int kOutHeader = 10;

__device__ __forceinline__ double convert_unit(double v, double num, double den) {
    return (num == den) ? v : (v * num) / den;
}

// energies and pair counts: fixed-order strided sums + shared-memory tree (one block of kReduceWarps * 32 threads)
__device__ __forceinline__ void reduce_energy_block(const ReduceArgs& A) {
    constexpr int NT = kReduceWarps * 32;
    __shared__ double s_e[7][NT];
    __shared__ unsigned long long s_c[2][NT];
    double e[7] = {0, 0, 0, 0, 0, 0, 0};
    unsigned long long c0 = 0, c1 = 0;
    for (int i = threadIdx.x; i < A.n_bpart; i += NT)
        for (int c = 0; c < 5; ++c) e[c] += A.bpart[5 * (size_t)i + c];
    for (int i = threadIdx.x; i < A.n_epart; i += NT) {
        e[5] += A.epart[2 * (size_t)i];
        e[6] += A.epart[2 * (size_t)i + 1];
        c0 += A.cpart[2 * (size_t)i];
        c1 += A.cpart[2 * (size_t)i + 1];
    }
    for (int c = 0; c < 7; ++c) s_e[c][threadIdx.x] = e[c];
    s_c[0][threadIdx.x] = c0;
    s_c[1][threadIdx.x] = c1;
    __syncthreads();
    for (int o = NT / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            for (int c = 0; c < 7; ++c) s_e[c][threadIdx.x] += s_e[c][threadIdx.x + o];
            s_c[0][threadIdx.x] += s_c[0][threadIdx.x + o];
            s_c[1][threadIdx.x] += s_c[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double total = 0;
        for (int c = 0; c < 7; ++c) {      // MMFF94.java:390-392: sum in component order
            const double v = convert_unit(s_e[c][0], A.unit_num, A.unit_den);
            A.out[1 + c] = v;
            total += v;
        }
        // a truncated survivor list means missing pairs: the total is poisoned (NaN) so that no caller - host or
        // device side, before or after an all-reduce - can mistake the result for a valid energy
        // bit 0: retry with longer lists, bit 1: retry with storage for more clusters, bit 2: non-finite coordinate
        const int ovf = A.overflow ? ((A.overflow[0] ? 1 : 0) | (A.overflow[3] ? 2 : 0) | (A.overflow[1] ? 4 : 0)) : 0;
        A.out[0] = ovf ? __longlong_as_double(0x7ff8000000000000ll) : total;
        A.out[8] = (double)s_c[0][0];
        A.out[9] = (double)s_c[1][0];
        if (A.out_host) {
            for (int c = 0; c < kOutHeader; ++c) A.out_host[c] = A.out[c];
            A.out_host[kOutHeader] = (double)ovf;
            __threadfence_system();
        }
    }
}

__global__ void __launch_bounds__(kReduceWarps * 32)
reduce_kernel(ReduceArgs A) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // no partial-gradient slots (cell path): one atom per THREAD; otherwise 32 atoms per block, warps over slots
    const int atoms_per_block = A.n_slots == 0 ? kReduceWarps * 32 : 32;
    const int n_atom_blocks = (A.n + atoms_per_block - 1) / atoms_per_block;
    __shared__ double s_g[kReduceWarps][3][32];
    // energy-only launches have a single block: the energy block
    const int bid = blockIdx.x + (A.want_grad ? 0 : n_atom_blocks);
    if (bid < n_atom_blocks && A.n_slots == 0) {
        const int a = bid * atoms_per_block + (int)threadIdx.x;
        if (a < A.n) {
            double gx = 0, gy = 0, gz = 0;
            if (A.direct) {
                gx = A.direct[a];
                gy = A.direct[A.n_pad + a];
                gz = A.direct[2 * (size_t)A.n_pad + a];
            }
            for (long long e = A.slot_ptr[a]; e < A.slot_ptr[a + 1]; ++e) {
                const long long s = A.slot_idx[e];
                if (s >= A.slot_lo && s < A.slot_hi) {
                    gx += A.slots[3 * s];
                    gy += A.slots[3 * s + 1];
                    gz += A.slots[3 * s + 2];
                }
            }
            double* o = A.out + kOutHeader + 3 * (size_t)a;
            o[0] = convert_unit(gx, A.unit_num, A.unit_den);
            o[1] = convert_unit(gy, A.unit_num, A.unit_den);
            o[2] = convert_unit(gz, A.unit_num, A.unit_den);
        }
        return;
    }
    if (bid < n_atom_blocks) {
        const int a = bid * 32 + lane;
        double gx = 0, gy = 0, gz = 0;
        for (int s = warp; s < A.n_slots; s += kReduceWarps) {
            const double* p = A.gpart + (size_t)s * 3 * A.n_pad + a;
            gx += p[0];
            gy += p[A.n_pad];
            gz += p[2 * (size_t)A.n_pad];
        }
        s_g[warp][0][lane] = gx;
        s_g[warp][1][lane] = gy;
        s_g[warp][2][lane] = gz;
        __syncthreads();
        if (warp == 0 && a < A.n) {
            gx = gy = gz = 0;
#pragma unroll
            for (int w = 0; w < kReduceWarps; ++w) {
                gx += s_g[w][0][lane];
                gy += s_g[w][1][lane];
                gz += s_g[w][2][lane];
            }
            if (A.direct) {
                gx += A.direct[a];
                gy += A.direct[A.n_pad + a];
                gz += A.direct[2 * (size_t)A.n_pad + a];
            }
            for (long long e = A.slot_ptr[a]; e < A.slot_ptr[a + 1]; ++e) {
                const long long s = A.slot_idx[e];
                if (s >= A.slot_lo && s < A.slot_hi) {
                    gx += A.slots[3 * s];
                    gy += A.slots[3 * s + 1];
                    gz += A.slots[3 * s + 2];
                }
            }
            double* o = A.out + kOutHeader + 3 * (size_t)a;
            o[0] = convert_unit(gx, A.unit_num, A.unit_den);
            o[1] = convert_unit(gy, A.unit_num, A.unit_den);
            o[2] = convert_unit(gz, A.unit_num, A.unit_den);
        }
        return;
    }
    reduce_energy_block(A);
}

// ---------------------------------------------------------------------------------------------
// marshalling helpers
// ---------------------------------------------------------------------------------------------
__global__ void aos_to_soa_kernel(const double* __restrict__ xyz, int n, double* __restrict__ x,
                                  double* __restrict__ y, double* __restrict__ z) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n) {
        x[a] = xyz[3 * (size_t)a];
        y[a] = xyz[3 * (size_t)a + 1];
        z[a] = xyz[3 * (size_t)a + 2];
    }
}

__global__ void soa_to_aos_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                  const double* __restrict__ z, int n, double* __restrict__ xyz) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n) {
        xyz[3 * (size_t)a] = x[a];
        xyz[3 * (size_t)a + 1] = y[a];
        xyz[3 * (size_t)a + 2] = z[a];
    }
}

__global__ void set_coord1_kernel(double* x, double* y, double* z, int atom, int axis, double value) {
    (axis == 0 ? x : axis == 1 ? y : z)[atom] = value;
}

// FP64 FMA-pipe peak: 8 independent dependent-chains of DFMA per thread.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
    for (int i = 0; i < iters; ++i) {
        v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
        v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}

}  // namespace jme
