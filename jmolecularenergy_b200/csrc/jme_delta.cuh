// jme_delta.cuh - the energy change of a single-coordinate move (the minimizers' inner step).
//
// The reference's minimizers (minimizer/GradientDescentMinimizer.java:75-93, AdamMinimizer.java:89-101) move ONE
// coordinate of ONE atom, evaluate the whole energy again and look at the difference.  Only the terms that contain the
// moved atom change: its N - 1 nonbonded pairs and the few bonded terms it takes part in.  delta_kernel evaluates exactly
// those, at the old and at the new position, with the same term arithmetic as the full evaluation (pair_terms, bond_term,
// angle_stbn_term, oop_term, torsion_term), and returns the difference per energy component - O(N) work instead of
// O(N^2) (or O(N x neighbours)), and a difference that does not carry the rounding of two sums over all pairs.
// One launch: every block takes a stride of partner atoms; the last block to finish adds the partials in block order,
// the bonded terms of the atom (its rows of the atom -> slot table the gradient gather uses), converts the units, writes
// the seven component changes and their sum to mapped host memory and only then stores the new coordinate.
// Opt-in (jme_energy_delta, jme_set_incremental): the default minimizer loops evaluate the full energy like the reference.
#pragma once

#include "jme_kernels.cuh"

namespace jme {

constexpr int kDeltaThreads = 256;
constexpr int kDeltaMaxBlocks = 64;
constexpr int kDeltaExclSmem = 256;        // exclusion-row entries staged in shared memory (longer rows: read from global)

struct DeltaArgs {
    DeviceSystem S;
    double* x; double* y; double* z;       // the same coordinate arrays, writable
    int atom, axis;
    double value;                          // the coordinate after the move
    int r_atom, r_axis;                    // a pending undo of the previous (rejected) move, applied first: one coordinate
    double r_value;                        // back to r_value (r_atom < 0: none) - saves the minimizers a launch per rejection
    int cutoff_on;
    const long long* excl_ptr;             // exclusion rows by atom: partner | kFlag14
    const unsigned int* excl_ent;
    BondedTerms B;
    const long long* slot_ptr;             // atom -> bonded gradient slots (= the bonded terms it takes part in)
    const long long* slot_idx;
    double* part;                          // [blocks][2] (vdw, ele) partial differences
    unsigned int* done_ctr;
    double unit_num, unit_den;
    double* out_host;                      // mapped: [0] total, [1..7] components, [8] sequence number (written last)
    double seq;
};

__device__ __forceinline__ double delta_r2(const Vec3& a, const Vec3& b) {
    const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    // strict double, left to right, no FMA: the value the cutoff predicate of the full evaluation sees
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__global__ void __launch_bounds__(kDeltaThreads)
delta_kernel(DeltaArgs A) {
    const DeviceSystem& S = A.S;
    __shared__ unsigned int s_ex[kDeltaExclSmem];
    __shared__ double s_red[kDeltaThreads / 32][7];
    __shared__ int s_last;
    const int a = A.atom;
    auto load = [&](int k) {                                      // coordinates as they are once the pending undo is applied
        Vec3 p = load_pos(S, k);
        if (k == A.r_atom) { if (A.r_axis == 0) p.x = A.r_value; else if (A.r_axis == 1) p.y = A.r_value; else p.z = A.r_value; }
        return p;
    };
    const Vec3 p_old = load(a);
    Vec3 p_new = p_old;
    if (A.axis == 0) p_new.x = A.value; else if (A.axis == 1) p_new.y = A.value; else p_new.z = A.value;
    const long long e0 = A.excl_ptr[a], e1 = A.excl_ptr[a + 1];
    const int nex = (int)(e1 - e0);
    for (int t = threadIdx.x; t < min(nex, kDeltaExclSmem); t += blockDim.x) s_ex[t] = A.excl_ent[e0 + t];
    __syncthreads();
    const double qa2 = 2.0 * S.ele_scale * S.q[a];                 // doubled energies, see pair_terms
    const PairConst* row = S.table + (size_t)S.type[a] * S.n_types;
    double dv = 0, de = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < S.n; j += gridDim.x * blockDim.x) {
        if (j == a) continue;
        int kind = 0;                                             // 0 counts, 1 excluded (1-2, 1-3), 2 a 1-4 pair
        for (int t = 0; t < nex; ++t) {
            const unsigned int x = t < kDeltaExclSmem ? s_ex[t] : A.excl_ent[e0 + t];
            if ((x & ~kFlag14) == (unsigned int)j) { kind = (x & kFlag14) ? 2 : 1; break; }
        }
        if (kind == 1) continue;
        const Vec3 pj = load(j);
        const PairConst pc = row[S.type[j]];
        const double qq2 = (qa2 * S.q[j]) * (kind == 2 ? 0.75 : 1.0);
        double ev, ee, f;
        const double r2o = delta_r2(p_old, pj);
        if (!A.cutoff_on || !(r2o > S.r2_max)) {
            pair_terms<false>(clamp_r2(r2o), pc, qq2, ev, ee, f);
            dv -= ev; de -= ee;
        }
        const double r2n = delta_r2(p_new, pj);
        if (!A.cutoff_on || !(r2n > S.r2_max)) {
            pair_terms<false>(clamp_r2(r2n), pc, qq2, ev, ee, f);
            dv += ev; de += ee;
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    dv = warp_sum(dv); de = warp_sum(de);
    if (lane == 0) { s_red[warp][0] = dv; s_red[warp][1] = de; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0, e = 0;
        for (int w = 0; w < kDeltaThreads / 32; ++w) { v += s_red[w][0]; e += s_red[w][1]; }
        A.part[2 * blockIdx.x] = 0.5 * v;
        A.part[2 * blockIdx.x + 1] = 0.5 * e;
        __threadfence();
        s_last = atomicAdd(A.done_ctr, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- last block: the bonded terms of the atom, old and new
    auto pos = [&](int k, bool moved) { return k == a ? (moved ? p_new : p_old) : load(k); };
    double db[5] = {0, 0, 0, 0, 0};                                // bond, angle, stretch-bend, out-of-plane, torsion
    const BondedTerms& B = A.B;
    for (long long s = A.slot_ptr[a] + threadIdx.x; s < A.slot_ptr[a + 1]; s += blockDim.x) {
        long long u = A.slot_idx[s];
        Vec3 g0, g1, g2, g3;
        for (int moved = 0; moved < 2; ++moved) {
            const double sign = moved ? 1.0 : -1.0;
            long long v = u;
            if (v < 2ll * B.n_bond) {
                const long long t = v / 2;
                db[0] += sign * bond_term(pos(B.bond_at[2 * t], moved), pos(B.bond_at[2 * t + 1], moved), B.bond_par[2 * t],
                                          B.bond_par[2 * t + 1], g0, g1);
            } else if ((v -= 2ll * B.n_bond) < 3ll * B.n_angle) {
                const long long t = v / 3;
                const double* p = B.angle_par + 6 * t;
                double ea = 0, es = 0;
                angle_stbn_term(pos(B.angle_at[3 * t], moved), pos(B.angle_at[3 * t + 1], moved), pos(B.angle_at[3 * t + 2], moved),
                                p[0], p[1], B.angle_linear[t] != 0, p[2], p[3], p[4], p[5], ea, es, g0, g1, g2);
                db[1] += sign * ea;
                db[2] += sign * es;
            } else if ((v -= 3ll * B.n_angle) < 4ll * B.n_oop) {
                const long long t = v / 4;
                const int* at = B.oop_at + 4 * t;
                db[3] += sign * oop_term(pos(at[0], moved), pos(at[1], moved), pos(at[2], moved), pos(at[3], moved), B.oop_par[t],
                                         g0, g1, g2, g3);
            } else {
                v -= 4ll * B.n_oop;
                const long long t = v / 4;
                const int* at = B.tor_at + 4 * t;
                const double* p = B.tor_par + 3 * t;
                db[4] += sign * torsion_term(pos(at[0], moved), pos(at[1], moved), pos(at[2], moved), pos(at[3], moved), p[0], p[1],
                                             p[2], g0, g1, g2, g3);
            }
        }
    }
    __syncthreads();                                              // (s_red is reused)
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const double v = warp_sum(db[c]);
        if (lane == 0) s_red[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double comp[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 5; ++c)
            for (int w = 0; w < kDeltaThreads / 32; ++w) comp[c] += s_red[w][c];
        for (unsigned int b = 0; b < gridDim.x; ++b) { comp[5] += A.part[2 * b]; comp[6] += A.part[2 * b + 1]; }
        double total = 0;
        for (int c = 0; c < 7; ++c) {
            const double v = convert_unit(comp[c], A.unit_num, A.unit_den);
            A.out_host[1 + c] = v;
            total += v;
        }
        A.out_host[0] = total;
        // every block has read the old positions by now: the pending undo, then the move
        if (A.r_atom >= 0) (A.r_axis == 0 ? A.x : (A.r_axis == 1 ? A.y : A.z))[A.r_atom] = A.r_value;
        (A.axis == 0 ? A.x : (A.axis == 1 ? A.y : A.z))[a] = A.value;
        *A.done_ctr = 0;
        __threadfence_system();
        A.out_host[8] = A.seq;
        __threadfence_system();
    }
}

}  // namespace jme
