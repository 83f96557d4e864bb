// TEST HARNESS ONLY: compiles the product's per-term math (csrc/jme_math.cuh) and host marshalling
// (csrc/jme_host.cpp) for the CPU so that the analytic gradients, the type-pair table, the exclusion
// CSR and the r^2 cutoff threshold can be checked against the oracle without a GPU.  It is not a CPU
// fallback: nothing in the package loads it.
#include <cstring>
#include <vector>

#include "../../jmolecularenergy_b200/csrc/jme_host.hpp"

using namespace jme;

static Vec3 P(const double* xyz, int a) { return Vec3{xyz[3 * a], xyz[3 * a + 1], xyz[3 * a + 2]}; }
static void add(double* g, int a, Vec3 v) { g[3 * a] += v.x; g[3 * a + 1] += v.y; g[3 * a + 2] += v.z; }

extern "C" int hostcheck_evaluate(const jme_system_t* s, const double* xyz, double cutoff, double dielectric,
                                  double comps[7], double* grad, uint64_t* n_pairs) {
    HostSystem h;
    std::string err = build_host_system(s, h);
    if (!err.empty()) return JME_ERR_INVALID;
    const int64_t n = h.n;
    for (int c = 0; c < 7; ++c) comps[c] = 0;
    std::memset(grad, 0, sizeof(double) * 3 * n);
    for (size_t b = 0; b < h.bond_par.size() / 2; ++b) {
        Vec3 gi, gj;
        int i = h.bond_at[2 * b], j = h.bond_at[2 * b + 1];
        comps[JME_BOND] += bond_term(P(xyz, i), P(xyz, j), h.bond_par[2 * b], h.bond_par[2 * b + 1], gi, gj);
        add(grad, i, gi); add(grad, j, gj);
    }
    for (size_t a = 0; a < h.angle_linear.size(); ++a) {
        Vec3 gi, gj, gk;
        double ea, es;
        int i = h.angle_at[3 * a], j = h.angle_at[3 * a + 1], k = h.angle_at[3 * a + 2];
        const double* p = &h.angle_par[6 * a];
        angle_stbn_term(P(xyz, i), P(xyz, j), P(xyz, k), p[0], p[1], h.angle_linear[a] != 0, p[2], p[3], p[4], p[5], ea, es, gi, gj, gk);
        comps[JME_ANGLE] += ea; comps[JME_STRETCH_BEND] += es;
        add(grad, i, gi); add(grad, j, gj); add(grad, k, gk);
    }
    for (size_t o = 0; o < h.oop_par.size(); ++o) {
        Vec3 gi, gj, gk, gl;
        const int32_t* at = &h.oop_at[4 * o];
        comps[JME_OOP] += oop_term(P(xyz, at[0]), P(xyz, at[1]), P(xyz, at[2]), P(xyz, at[3]), h.oop_par[o], gi, gj, gk, gl);
        add(grad, at[0], gi); add(grad, at[1], gj); add(grad, at[2], gk); add(grad, at[3], gl);
    }
    for (size_t t = 0; t < h.tor_par.size() / 3; ++t) {
        Vec3 gi, gj, gk, gl;
        const int32_t* at = &h.tor_at[4 * t];
        comps[JME_TORSION] += torsion_term(P(xyz, at[0]), P(xyz, at[1]), P(xyz, at[2]), P(xyz, at[3]),
                                           h.tor_par[3 * t], h.tor_par[3 * t + 1], h.tor_par[3 * t + 2], gi, gj, gk, gl);
        add(grad, at[0], gi); add(grad, at[1], gj); add(grad, at[2], gk); add(grad, at[3], gl);
    }
    const double T2 = cutoff > 0 ? cutoff_r2_threshold(cutoff) : 0;
    uint64_t pairs = 0;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i + 1; j < n; ++j) {
            int cls = pair_class(h, i, j);
            if (cls == 1) continue;
            double dx = xyz[3 * i] - xyz[3 * j], dy = xyz[3 * i + 1] - xyz[3 * j + 1], dz = xyz[3 * i + 2] - xyz[3 * j + 2];
            double r2 = dx * dx + dy * dy + dz * dz;
            if (cutoff > 0 && r2 > T2) continue;
            ++pairs;
            double qq2 = (2.0 * kEleC * h.charge[i] / dielectric) * h.charge[j];
            if (cls == 2) qq2 *= kEle14;
            double ev2, ee2, f;
            pair_terms<true>(r2, h.pair_table[(size_t)h.type_idx[i] * h.n_types + h.type_idx[j]], qq2, ev2, ee2, f);
            comps[JME_VDW] += 0.5 * ev2; comps[JME_ELE] += 0.5 * ee2;
            grad[3 * i] += f * dx; grad[3 * i + 1] += f * dy; grad[3 * i + 2] += f * dz;
            grad[3 * j] -= f * dx; grad[3 * j + 1] -= f * dy; grad[3 * j + 2] -= f * dz;
        }
    *n_pairs = pairs;
    return JME_OK;
}
