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

#include <cooperative_groups.h>

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

// ---- the whole minimizer loop on the device ------------------------------------------------------------------
// minimize_kernel runs the reference's coordinate-wise loops (GradientDescentMinimizer.java:49-105, AdamMinimizer.java:
// 62-120) entirely inside ONE thread-block cluster (kMinimizeCtas CTAs on neighbouring SMs): every CTA keeps the
// coordinates in its shared memory and the loop's scalars in its thread 0, every trial move is answered by the same
// per-atom difference as delta_kernel - the cluster's threads stride over the partners and over the atom's bonded
// terms - each CTA stores its seven partial sums into every CTA's shared memory (distributed shared memory), one
// cluster barrier, and every CTA adds them in rank order and takes the same decision.  No launch, no host round trip
// per trial (measured at 5 001 atoms: 11.5 / 7.7 / 5.8 / 5.3 us per trial with 1 / 2 / 4 / 8 CTAs).
// jme_set_incremental(h, 2) selects it for systems whose coordinates fit (n <= kMinimizeMaxAtoms).
constexpr int kMinimizeThreads = 512;
#ifndef JME_MINIMIZE_CTAS
#define JME_MINIMIZE_CTAS 8
#endif
constexpr int kMinimizeCtas = JME_MINIMIZE_CTAS;   // 8 = the portable cluster size
constexpr int kMinimizeMaxAtoms = 8192;            // 3 x 8 bytes x 8192 = 192 KB of shared memory

struct MinimizeArgs {
    DeviceSystem S;
    double* x; double* y; double* z;               // coordinates in, final coordinates out
    int cutoff_on;
    const long long* excl_ptr;
    const unsigned int* excl_ent;
    BondedTerms B;
    const long long* slot_ptr;
    const long long* slot_idx;
    double unit_num, unit_den;
    int adam;                                      // 0: GradientDescentMinimizer, 1: AdamMinimizer
    int max_steps;
    double step_size, threshold, beta1, beta2, epsilon;
    double e0;                                     // energy of the initial coordinates (one full evaluation)
    double* grads;                                 // [CTAs][3n] work: the per-coordinate slopes
    double* adam_m; double* adam_v; int* adam_t;   // [CTAs][3n] work (Adam)
    double* log; int capacity;                     // log[k] = energy after step k, k >= 1 (log[0] is the caller's e0)
    long long* out_calls; int* out_steps; double* out_energy;
};

__global__ void __launch_bounds__(kMinimizeThreads, 1)
minimize_kernel(MinimizeArgs A) {
    extern __shared__ __align__(16) unsigned char min_smem[];
    const DeviceSystem& S = A.S;
    const int n = S.n;
    double* sx = reinterpret_cast<double*>(min_smem);
    double* sy = sx + n;
    double* sz = sy + n;
    __shared__ double s_red[kMinimizeThreads / 32][7];
    __shared__ double s_part[2][kMinimizeCtas][8];  // [trial parity][CTA][component]: every CTA's partial sums, in every CTA
    __shared__ double s_bc[4];                     // broadcast: [0] the move, [1] go on (1) / skip (0), [2] the accepted difference
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int n_cta = (int)cluster.num_blocks(), cta = (int)cluster.block_rank();
    unsigned int trial = 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = tid; k < n; k += blockDim.x) { sx[k] = A.x[k]; sy[k] = A.y[k]; sz[k] = A.z[k]; }
    // the loop's per-coordinate state, one private copy per CTA (every CTA runs the same scalar bookkeeping on the same
    // inputs and so takes the same decisions: nothing is shared, nothing is broadcast)
    double* grads = A.grads + (size_t)cta * 3 * n;
    double* adam_m = A.adam ? A.adam_m + (size_t)cta * 3 * n : nullptr;
    double* adam_v = A.adam ? A.adam_v + (size_t)cta * 3 * n : nullptr;
    int* adam_t = A.adam ? A.adam_t + (size_t)cta * 3 * n : nullptr;
    for (int k = tid; k < 3 * n; k += blockDim.x) {
        grads[k] = A.step_size;
        if (A.adam) { adam_m[k] = 0.0; adam_v[k] = 0.0; adam_t[k] = 1; }
    }
    __syncthreads();
    auto spos = [&](int k) { return Vec3{sx[k], sy[k], sz[k]}; };
    const BondedTerms& B = A.B;

    // the energy change of moving coordinate (a, axis) to `value` (block- and cluster-wide work; the result on thread 0)
    auto delta_of = [&](int a, int axis, double value) -> double {
        const Vec3 p_old = spos(a);
        Vec3 p_new = p_old;
        if (axis == 0) p_new.x = value; else if (axis == 1) p_new.y = value; else p_new.z = value;
        const long long e0 = A.excl_ptr[a], e1 = A.excl_ptr[a + 1];
        const double qa2 = 2.0 * S.ele_scale * S.q[a];
        const PairConst* row = S.table + (size_t)S.type[a] * S.n_types;
        double c[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int j = cta * blockDim.x + tid; j < n; j += n_cta * blockDim.x) {
            if (j == a) continue;
            int kind = 0;
            for (long long t = e0; t < e1; ++t) {
                const unsigned int xe = A.excl_ent[t];
                if ((xe & ~kFlag14) == (unsigned int)j) { kind = (xe & kFlag14) ? 2 : 1; break; }
            }
            if (kind == 1) continue;
            const Vec3 pj = spos(j);
            const PairConst pc = row[S.type[j]];
            const double qq2 = (qa2 * S.q[j]) * (kind == 2 ? 0.75 : 1.0);
            double ev, ee, f;
            const double r2o = delta_r2(p_old, pj);
            if (!A.cutoff_on || !(r2o > S.r2_max)) { pair_terms<false>(clamp_r2(r2o), pc, qq2, ev, ee, f); c[5] -= ev; c[6] -= ee; }
            const double r2n = delta_r2(p_new, pj);
            if (!A.cutoff_on || !(r2n > S.r2_max)) { pair_terms<false>(clamp_r2(r2n), pc, qq2, ev, ee, f); c[5] += ev; c[6] += ee; }
        }
        c[5] *= 0.5; c[6] *= 0.5;                   // (pair_terms returns doubled energies)
        auto pos = [&](int k, bool moved) { return k == a ? (moved ? p_new : p_old) : spos(k); };
        // the atom's bonded terms, spread over the cluster: the last threads of every CTA (the first ones have the most partners)
        for (long long s = A.slot_ptr[a] + (long long)(blockDim.x - 1 - tid) * n_cta + cta; s < A.slot_ptr[a + 1]; s += (long long)blockDim.x * n_cta) {
            const long long u = A.slot_idx[s];
            Vec3 g0, g1, g2, g3;
            for (int moved = 0; moved < 2; ++moved) {
                const double sign = moved ? 1.0 : -1.0;
                long long v = u;
                if (v < 2ll * B.n_bond) {
                    const long long t = v / 2;
                    c[0] += sign * bond_term(pos(B.bond_at[2 * t], moved), pos(B.bond_at[2 * t + 1], moved), B.bond_par[2 * t],
                                             B.bond_par[2 * t + 1], g0, g1);
                } else if ((v -= 2ll * B.n_bond) < 3ll * B.n_angle) {
                    const long long t = v / 3;
                    const double* p = B.angle_par + 6 * t;
                    double ea = 0, es = 0;
                    angle_stbn_term(pos(B.angle_at[3 * t], moved), pos(B.angle_at[3 * t + 1], moved), pos(B.angle_at[3 * t + 2], moved),
                                    p[0], p[1], B.angle_linear[t] != 0, p[2], p[3], p[4], p[5], ea, es, g0, g1, g2);
                    c[1] += sign * ea;
                    c[2] += sign * es;
                } else if ((v -= 3ll * B.n_angle) < 4ll * B.n_oop) {
                    const long long t = v / 4;
                    const int* at = B.oop_at + 4 * t;
                    c[3] += sign * oop_term(pos(at[0], moved), pos(at[1], moved), pos(at[2], moved), pos(at[3], moved), B.oop_par[t],
                                            g0, g1, g2, g3);
                } else {
                    v -= 4ll * B.n_oop;
                    const long long t = v / 4;
                    const int* at = B.tor_at + 4 * t;
                    const double* p = B.tor_par + 3 * t;
                    c[4] += sign * torsion_term(pos(at[0], moved), pos(at[1], moved), pos(at[2], moved), pos(at[3], moved), p[0],
                                                p[1], p[2], g0, g1, g2, g3);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const double v = warp_sum(c[k]);
            if (lane == 0) s_red[warp][k] = v;
        }
        __syncthreads();
        const unsigned int par = trial & 1u;
        ++trial;
        if (warp == 0) {
            // this CTA's seven sums, stored into every CTA's copy of the partial table (lane r -> CTA r)
            double mine[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) mine[k] = warp_sum(lane < kMinimizeThreads / 32 ? s_red[lane][k] : 0.0);
            if (lane < n_cta) {
                double* dst = cluster.map_shared_rank(&s_part[par][cta][0], lane);
#pragma unroll
                for (int k = 0; k < 7; ++k) dst[k] = mine[k];
            }
        }
        cluster.sync();                                            // (also a barrier of this CTA; the table alternates with the
        double total = 0;                                          // trial, so one barrier per trial is enough)
        if (warp == 0) {                                           // only thread 0 needs the value: it takes the decision
            double v = 0;                                          // lane k: component k over the CTAs in rank order, converted
            if (lane < 7) {
                for (int r = 0; r < n_cta; ++r) v += s_part[par][r][lane];
                v = convert_unit(v, A.unit_num, A.unit_den);
            }
#pragma unroll
            for (int k = 0; k < 7; ++k) total += __shfl_sync(0xffffffffu, v, k);    // components in the reference's order
        }
        return total;
    };
    auto set_coord = [&](int a, int axis, double value) { (axis == 0 ? sx : (axis == 1 ? sy : sz))[a] = value; };
    auto get_coord = [&](int a, int axis) { return (axis == 0 ? sx : (axis == 1 ? sy : sz))[a]; };

    long long calls = 1;                                          // the caller's evaluation of e0
    double g_next = tid == 0 ? grads[0] : 0.0;                    // thread 0 reads the slope of the NEXT trial during this one
    int step = 0;
    bool initialized = false;
    double last_energy = A.e0, previous = 0;
    while (step < A.max_steps) {
        if (step != 0 && fabs(previous - last_energy) <= A.threshold) break;
        previous = last_energy;
        for (int a = 0; a < n; ++a)
            for (int i = 0; i < 3; ++i) {
                // thread 0 decides the trial (every thread could: the inputs are shared - but pow() once is enough)
                if (tid == 0) {
                    const double g = g_next;
                    const int nxt = 3 * a + i + 1;
                    g_next = grads[nxt < 3 * n ? nxt : 0];        // (index 0 again after the last one: the next sweep starts there)
                    double distance = 0;
                    double go = 1;
                    if (g == 0) {
                        go = 0;
                    } else if (!A.adam) {
                        const double mag = fmin(fabs(A.step_size * g), .3);
                        distance = mag * ((g > 0) - (g < 0));
                    } else {
                        const double x = get_coord(a, i);
                        const int t = ++adam_t[3 * a + i];
                        const double m = adam_m[3 * a + i] = A.beta1 * adam_m[3 * a + i] + (1 - A.beta1) * g;
                        const double v = adam_v[3 * a + i] = A.beta2 * adam_v[3 * a + i] + (1 - A.beta2) * g * g;
                        const double m_hat = m / (1 - pow(A.beta1, (double)t));
                        const double v_hat = v / (1 - pow(A.beta2, (double)t));
                        distance = (x - A.step_size * m_hat / (sqrt(v_hat) + A.epsilon)) - x;
                    }
                    s_bc[0] = distance;
                    s_bc[1] = go;
                }
                __syncthreads();
                const double distance = s_bc[0];
                const bool go = s_bc[1] != 0;
                if (!go) { __syncthreads(); continue; }
                const double x_old = get_coord(a, i);
                const double x_new = x_old + distance;            // movePoint: p += d
                calls += A.adam ? 2 : 1;                           // (Adam asks for the energy before the move as well)
                const double difference = delta_of(a, i, x_new);
                if (tid == 0) {
                    double g;
                    bool keep;
                    if (!A.adam) {
                        g = -difference / (distance == 0 ? 1 : distance);
                        keep = initialized && !(difference > 0);
                        if (initialized) g = g * (difference > 0 ? 0.9 : 1.05);
                    } else {
                        g = -difference / fabs(distance);
                        keep = initialized && !(difference > 0);
                    }
                    grads[3 * a + i] = g;
                    set_coord(a, i, keep ? x_new : x_new + (-distance));     // the undo is a second movePoint, as in the reference
                    s_bc[2] = keep ? difference : 0.0;
                }
                __syncthreads();
                last_energy += s_bc[2];
            }
        if (!initialized) {
            initialized = true;
        } else {
            ++step;
            if (cta == 0 && tid == 0 && step < A.capacity) A.log[step] = last_energy;
        }
    }
    cluster.sync();                                                // nobody leaves while a peer may still write into its table
    if (cta == 0) {
        for (int k = tid; k < n; k += blockDim.x) { A.x[k] = sx[k]; A.y[k] = sy[k]; A.z[k] = sz[k]; }
        if (tid == 0) { *A.out_calls = calls; *A.out_steps = step; *A.out_energy = last_energy; }
    }
}

}  // namespace jme
