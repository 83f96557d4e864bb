// jme_host.hpp - host-side marshalling shared by the C ABI (jme_api.cu) and tests/hostcheck:
// validation of a jme_system_t, dense type indices, the type-pair vdW table, the topological
// pair predicate (exclusions / 1-4 flags) and term enumeration.  Pure C++17, no CUDA.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/jme_b200.h"
#include "jme_math.cuh"

namespace jme {

struct HostSystem {
    int64_t n = 0;
    int n_types = 0;                       // types that occur, dense 0..n_types-1
    std::vector<int32_t> type_idx;         // [n] dense index
    std::vector<int32_t> type_of_idx;      // [n_types] MMFF numeric type
    std::vector<double> charge;            // [n]
    std::vector<PairConst> pair_table;     // [n_types*n_types], symmetric
    std::vector<double> pair_rstar, pair_eps;   // same shape, for diagnostics/tests

    // exclusion CSR over original atom indices: partners at topological distance 1..3, sorted;
    // bit 31 of an entry = "exactly 3 bonds" (1-4 pair: interacts, electrostatics x0.75)
    std::vector<int64_t> excl_ptr;         // [n+1]
    std::vector<uint32_t> excl_ent;

    // bonded terms, parameters widened to double, atoms in the reference's evaluation order
    std::vector<int32_t> bond_at;          // [nb][2]
    std::vector<double> bond_par;          // [nb][2] kb, r0
    std::vector<int32_t> angle_at;         // [na][3]
    std::vector<double> angle_par;         // [na][6] ka, theta0, kba_ijk, kba_kji, r0_ij, r0_jk
    std::vector<uint8_t> angle_linear;     // [na]
    std::vector<int32_t> oop_at;           // [no][4]
    std::vector<double> oop_par;           // [no] koop
    std::vector<int32_t> tor_at;           // [nt][4] (swapped like MMFF94TorsionComponent.java:117-124)
    std::vector<double> tor_par;           // [nt][3]
};

constexpr uint32_t kFlag14 = 0x80000000u;

// Returns "" on success, else the error text.
std::string build_host_system(const jme_system_t* sys, HostSystem& out);

// R*ij / eps_ij of MMFF94VdwComponent.java:106-135 for one pair of vdW rows.
void vdw_pair_constants(float alpha_i, float N_i, float A_i, float G_i, int DA_i,
                        float alpha_j, float N_j, float A_j, float G_j, int DA_j,
                        double& rstar, double& eps);

// CSR of partners at bond distance 1..3 (flag bit for exactly 3).  MMFF94.java:162-170,1106-1140.
void build_exclusions(int64_t n, int64_t nb, const int32_t* bi, const int32_t* bj,
                      std::vector<int64_t>& ptr, std::vector<uint32_t>& ent);

// Largest double T with sqrt(T) <= cutoff:  (sqrt(r2) > cutoff) == (r2 > T) for all doubles r2, which
// turns the reference's test on the rounded square root (MMFF94VdwComponent.java:136-137) into a test
// on r2 with identical outcome.
double cutoff_r2_threshold(double cutoff);

// MMFF94.java:158-212 on index arrays.
void enumerate_terms(int64_t n, const int32_t* type, const uint8_t* linear, int64_t nb,
                     const int32_t* bi, const int32_t* bj, std::vector<int32_t>& angles,
                     std::vector<int32_t>& oops, std::vector<int32_t>& torsions);

// 0 = interacts, 1 = excluded, 2 = interacts as a 1-4 pair
inline int pair_class(const HostSystem& h, int64_t i, int64_t j) {
    for (int64_t e = h.excl_ptr[i]; e < h.excl_ptr[i + 1]; ++e) {
        uint32_t v = h.excl_ent[e];
        if ((int64_t)(v & ~kFlag14) == j) return (v & kFlag14) ? 2 : 1;
    }
    return 0;
}

}  // namespace jme
