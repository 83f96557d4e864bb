// TEST INFRASTRUCTURE ONLY - CPU restatement ("oracle") of JMolecularEnergy's MMFF94 energy path.
//
// Nothing in the product (jmolecularenergy_b200/, libjme_b200.so) links, loads or calls this file.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and
// only as the checker / the CPU baseline.
//
// Why a restatement: the reference is Java on CDK 2.10 and neither a JVM nor the CDK jars exist in
// this image, so it cannot be executed here.  Parity status: PINNED - the
// restatement reproduces the reference's own golden energies (MMFF94.energies OPTIMOL column, the
// values MMFF94Test.java:79-81 asserts) for the hand-typable molecules under tests/golden/ to
// 1e-5 kcal/mol (tests/test_oracle_golden.py).  Gradients: the reference has none; the gradient
// here is forward-mode automatic differentiation of the restated energy, cross-checked against
// __float128 central differences of the same expression.
//
// Every function cites the reference lines it follows; paths are relative to
// /root/reference/src/main/java/org/jme/forcefield .  Operation order is kept (compile with
// -ffp-contract=off: Java never contracts a*b+c).
//
// Build: oracle/Makefile -> oracle/libjme_oracle.so

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <queue>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include <quadmath.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/jme_b200.h"

namespace {

// ---------------------------------------------------------------------------------------------
// scalar types: double (value path), Dual<N> (value path + exact partials), __float128 (FD check)
// ---------------------------------------------------------------------------------------------
template <int N>
struct Dual {
    double v;
    double d[N];
    Dual() : v(0) { for (int i = 0; i < N; ++i) d[i] = 0; }
    Dual(double x) : v(x) { for (int i = 0; i < N; ++i) d[i] = 0; }   // NOLINT: implicit on purpose
};
template <int N> Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v / b.v; for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }
template <int N> Dual<N> operator+(double a, const Dual<N>& b) { return Dual<N>(a) + b; }
template <int N> Dual<N> operator+(const Dual<N>& a, double b) { return a + Dual<N>(b); }
template <int N> Dual<N> operator-(double a, const Dual<N>& b) { return Dual<N>(a) - b; }
template <int N> Dual<N> operator-(const Dual<N>& a, double b) { return a - Dual<N>(b); }
template <int N> Dual<N> operator*(double a, const Dual<N>& b) { Dual<N> r; r.v = a * b.v; for (int i = 0; i < N; ++i) r.d[i] = a * b.d[i]; return r; }
template <int N> Dual<N> operator*(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r; }
template <int N> Dual<N> operator/(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v / b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b; return r; }
template <int N> Dual<N> operator/(double a, const Dual<N>& b) { return Dual<N>(a) / b; }
template <int N> bool operator>(const Dual<N>& a, double b) { return a.v > b; }
template <int N> bool operator<(const Dual<N>& a, double b) { return a.v < b; }

template <int N> Dual<N> chain(const Dual<N>& a, double f, double df) { Dual<N> r; r.v = f; for (int i = 0; i < N; ++i) r.d[i] = df * a.d[i]; return r; }

inline double m_sqrt(double x) { return std::sqrt(x); }
inline double m_exp(double x) { return std::exp(x); }
inline double m_pow(double x, double p) { return std::pow(x, p); }
inline double m_acos(double x) { return std::acos(x); }
inline double m_asin(double x) { return std::asin(x); }
inline double m_sin(double x) { return std::sin(x); }
inline double m_cos(double x) { return std::cos(x); }
inline double value_of(double x) { return x; }

inline __float128 m_sqrt(__float128 x) { return sqrtq(x); }
inline __float128 m_exp(__float128 x) { return expq(x); }
inline __float128 m_pow(__float128 x, double p) { return powq(x, (__float128)p); }
inline __float128 m_acos(__float128 x) { return acosq(x); }
inline __float128 m_asin(__float128 x) { return asinq(x); }
inline __float128 m_sin(__float128 x) { return sinq(x); }
inline __float128 m_cos(__float128 x) { return cosq(x); }
inline double value_of(__float128 x) { return (double)x; }

template <int N> Dual<N> m_sqrt(const Dual<N>& x) { double s = std::sqrt(x.v); return chain(x, s, 0.5 / s); }
template <int N> Dual<N> m_exp(const Dual<N>& x) { double e = std::exp(x.v); return chain(x, e, e); }
template <int N> Dual<N> m_pow(const Dual<N>& x, double p) { double f = std::pow(x.v, p); return chain(x, f, p * std::pow(x.v, p - 1.0)); }
// d/dx acos = -1/sqrt((1-x)(1+x)): the factored form keeps full relative accuracy next to |x| = 1.
// At |x| == 1 exactly (clamped or exactly collinear/planar input) the derivative is taken as 0: the
// energies built on these angles are even functions of the angle there, so that is their true gradient
// except for the ideally-linear angle form, which the tests never evaluate at exactly 180 degrees.
inline double inv_sqrt_1mx2(double x) { double t = (1.0 - x) * (1.0 + x); return t > 0 ? 1.0 / std::sqrt(t) : 0.0; }
template <int N> Dual<N> m_acos(const Dual<N>& x) { return chain(x, std::acos(x.v), -inv_sqrt_1mx2(x.v)); }
template <int N> Dual<N> m_asin(const Dual<N>& x) { return chain(x, std::asin(x.v), inv_sqrt_1mx2(x.v)); }
template <int N> Dual<N> m_sin(const Dual<N>& x) { return chain(x, std::sin(x.v), std::cos(x.v)); }
template <int N> Dual<N> m_cos(const Dual<N>& x) { return chain(x, std::cos(x.v), -std::sin(x.v)); }
template <int N> double value_of(const Dual<N>& x) { return x.v; }

template <class T> struct P3 { T x, y, z; };

const double PI = 3.14159265358979323846;   // java.lang.Math.PI

// Math.toRadians (Java 8, pom.xml:31-32): angdeg / 180.0 * PI
template <class T> T to_radians(const T& deg) { return deg / 180.0 * PI; }

// ForceField.EnergyUnit.convertTo from KCAL_PER_MOL (ForceField.java:153-158)
const double UNIT_TO_J[3] = {4184.0, 1000.0, 1.0};
template <class T> T convert(const T& v, int unit) {
    if (unit == JME_KCAL_PER_MOL) return v;
    return (v * UNIT_TO_J[JME_KCAL_PER_MOL]) / UNIT_TO_J[unit];
}

// javax.vecmath.Point3d.distance: sqrt(dx*dx + dy*dy + dz*dz), dx = this.x - p.x
template <class T> T distance(const P3<T>& a, const P3<T>& b) {
    T dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return m_sqrt(dx * dx + dy * dy + dz * dz);
}
template <class T> T clamp_to_one(const T& n) {   // GeometryUtils.java:73-75
    if (n > 1.0) return T(1.0);
    if (n < -1.0) return T(-1.0);
    return n;
}
// GeometryUtils.calculateAngle, GeometryUtils.java:43-51 (radians)
template <class T> T calc_angle(const P3<T>& i, const P3<T>& j, const P3<T>& k) {
    T ux = i.x - j.x, uy = i.y - j.y, uz = i.z - j.z;
    T vx = k.x - j.x, vy = k.y - j.y, vz = k.z - j.z;
    return m_acos(clamp_to_one((ux * vx + uy * vy + uz * vz) / (distance(i, j) * distance(k, j))));
}
template <class T> P3<T> cross(const P3<T>& a, const P3<T>& b) {   // vecmath Vector3d.cross
    return P3<T>{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class T> T dot(const P3<T>& a, const P3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> T length(const P3<T>& a) { return m_sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
template <class T> P3<T> normalize(const P3<T>& a) {               // vecmath: scale by 1.0/sqrt(...)
    T n = 1.0 / m_sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    return P3<T>{a.x * n, a.y * n, a.z * n};
}
// GeometryUtils.calculateTorsionAngle, GeometryUtils.java:61-71 (radians, unsigned)
template <class T> T calc_torsion(const P3<T>& i, const P3<T>& j, const P3<T>& k, const P3<T>& l) {
    P3<T> ij{i.x - j.x, i.y - j.y, i.z - j.z};
    P3<T> kj{k.x - j.x, k.y - j.y, k.z - j.z};
    P3<T> jk{-kj.x, -kj.y, -kj.z};
    P3<T> lk{l.x - k.x, l.y - k.y, l.z - k.z};
    P3<T> n1 = cross(ij, kj), n2 = cross(jk, lk);
    return m_acos(clamp_to_one(dot(n1, n2) / (length(n1) * length(n2))));
}
// GeometryUtils.calculateOutOfPlaneAngle, GeometryUtils.java:85-95 (radians, asin NOT clamped)
template <class T> T calc_oop(const P3<T>& i, const P3<T>& j, const P3<T>& k, const P3<T>& l) {
    P3<T> ij = normalize(P3<T>{i.x - j.x, i.y - j.y, i.z - j.z});
    P3<T> kj = normalize(P3<T>{k.x - j.x, k.y - j.y, k.z - j.z});
    P3<T> perp = cross(ij, kj);
    P3<T> lj = normalize(P3<T>{l.x - j.x, l.y - j.y, l.z - j.z});
    return m_asin(dot(perp, lj) / m_sin(calc_angle(i, j, k)));
}

// ---- per-term energies (kcal/mol before convert) --------------------------------------------
// MMFF94BondStretchingComponent.java:99-111
template <class T> T e_bond(const P3<T>& i, const P3<T>& j, float kb, float r0, int unit) {
    T length = distance(i, j);
    T dr = length - (double)r0;
    T e = .5 * 143.9325 * (double)kb * dr * dr * (1.0 - 2.0 * dr + (7.0 / 12.0) * 4.0 * dr * dr);
    return convert(e, unit);
}
// MMFF94AngleBendingComponent.java:116-131
template <class T> T e_angle(const P3<T>& i, const P3<T>& j, const P3<T>& k, float ka, float theta0, bool linear, int unit) {
    T angle = calc_angle(i, j, k) * 180.0 / PI;
    T dth = angle - (double)theta0;
    T e;
    if (linear) {
        e = 143.9325 * (double)ka * (1.0 + m_cos(to_radians(angle)));
    } else {
        const double c = to_radians(to_radians(143.9325));
        e = c * (double)ka * .5 * dth * dth * (1.0 - 0.006981317 * dth);
    }
    return convert(e, unit);
}
// MMFF94StretchBendComponent.java:123-140
template <class T> T e_stbn(const P3<T>& i, const P3<T>& j, const P3<T>& k, float theta0, float kba_ijk, float kba_kji,
                            float r0_ij, float r0_jk, bool linear, int unit) {
    if (linear) return T(0.0);
    T angle = calc_angle(i, j, k) * 180.0 / PI;
    T dth = angle - (double)theta0;
    T drij = distance(i, j) - (double)r0_ij;
    T drjk = distance(j, k) - (double)r0_jk;
    T e = 2.51210 * ((double)kba_ijk * drij + (double)kba_kji * drjk) * dth;
    return convert(e, unit);
}
// MMFF94OutOfPlaneComponent.java:122-134
template <class T> T e_oop(const P3<T>& i, const P3<T>& j, const P3<T>& k, const P3<T>& l, float koop, int unit) {
    T angle = calc_oop(i, j, k, l) * 180.0 / PI;
    const double c = to_radians(to_radians(143.9325));
    T e = c * .5 * (double)koop * angle * angle;
    return convert(e, unit);
}
// MMFF94TorsionComponent.java:111-132 (the i<->l / j<->k swap at :117-124 leaves the angle unchanged
// in exact arithmetic but changes rounding, so it is restated too)
template <class T> T e_torsion(P3<T> i, P3<T> j, P3<T> k, P3<T> l, int ti, int tj, int tk, int tl, int idx_j, int idx_k,
                               float v1, float v2, float v3, int unit) {
    if (tk < tj || (tj == tk && ti > tl) || (tj == tk && ti == tl && idx_j > idx_k)) {
        std::swap(j, k);
        std::swap(i, l);
    }
    T phi = calc_torsion(i, j, k, l);
    T e = 0.5 * ((double)v1 * (1.0 + m_cos(phi)) + (double)v2 * (1.0 - m_cos(2.0 * phi)) + (double)v3 * (1.0 + m_cos(3.0 * phi)));
    return convert(e, unit);
}

struct VdwRow { float alpha, N, A, G; int DA; };

// R*ij and eps_ij exactly as MMFF94VdwComponent.java:106-135 computes them for every pair
inline void vdw_constants(const VdwRow& a, const VdwRow& b, double& rstar, double& eps) {
    double rii = a.A * std::pow((double)a.alpha, 0.25);
    double rjj = b.A * std::pow((double)b.alpha, 0.25);
    double B = (a.DA == 1 || b.DA == 1) ? 0 : 0.2;
    double beta = 12;
    double gamma = (rii - rjj) / (rii + rjj);
    rstar = 0.5 * (rii + rjj) * (1 + B * (1 - std::exp(-beta * std::pow(gamma, 2))));
    eps = 181.16 * a.G * b.G * a.alpha * b.alpha;
    float ai = a.alpha / a.N, aj = b.alpha / b.N;             // float / float, :127
    eps /= (std::sqrt((double)ai) + std::sqrt((double)aj));
    eps /= std::pow(rstar, 6);
    if (a.DA + b.DA == 3) {
        rstar *= 0.8;
        eps *= 0.5;
    }
}
// MMFF94VdwComponent.java:136-142, given the distance
template <class T> T e_vdw(const T& r, double rstar, double eps, int unit) {
    T e = eps * m_pow((1.07 * rstar) / (r + 0.07 * rstar), 7);
    e = e * (((1.12 * std::pow(rstar, 7)) / (m_pow(r, 7) + 0.12 * std::pow(rstar, 7))) - 2.0);
    return convert(e, unit);
}
// MMFF94ElectrostaticComponent.java:134-151
template <class T> T e_ele(const T& r, double qi, double qj, double dielectric, bool is14, int unit) {
    T e = 332.0716 * qi * qj / (dielectric * (r + 0.05));
    if (is14) e = e * 0.75;
    return convert(e, unit);
}

// ---- topology (MMFF94.java:158-212, 1106-1140) ----------------------------------------------
struct Topology {
    std::vector<std::vector<int>> adj;          // neighbours in bond-list order
    std::vector<std::vector<int>> excl;         // per atom: sorted partners at distance 1..2
    std::vector<std::vector<int>> p14;          // per atom: sorted partners at distance exactly 3
};

Topology build_topology(int64_t n, int64_t nb, const int32_t* bi, const int32_t* bj) {
    Topology t;
    t.adj.assign(n, {});
    for (int64_t b = 0; b < nb; ++b) {
        t.adj[bi[b]].push_back(bj[b]);
        t.adj[bj[b]].push_back(bi[b]);
    }
    t.excl.assign(n, {});
    t.p14.assign(n, {});
    std::vector<int> depth(n, -1);
    for (int64_t a = 0; a < n; ++a) {
        if (t.adj[a].empty()) continue;
        // breadth-first to depth 3 = isSeparatedByNOrMoreBonds(a, *, 3): first hit decides
        std::vector<int> seen{(int)a};
        depth[a] = 0;
        std::vector<int> frontier{(int)a};
        for (int d = 1; d <= 3; ++d) {
            std::vector<int> next;
            for (int u : frontier)
                for (int v : t.adj[u])
                    if (depth[v] < 0) {
                        depth[v] = d;
                        seen.push_back(v);
                        next.push_back(v);
                    }
            frontier.swap(next);
        }
        for (int v : seen) {
            if (depth[v] == 1 || depth[v] == 2) t.excl[a].push_back(v);
            else if (depth[v] == 3) t.p14[a].push_back(v);
            depth[v] = -1;
        }
        std::sort(t.excl[a].begin(), t.excl[a].end());
        std::sort(t.p14[a].begin(), t.p14[a].end());
    }
    return t;
}

struct Sys {
    const jme_system_t* s;
    std::vector<VdwRow> row_of_atom;
    std::map<std::pair<int, int>, float> r0_of_bond;
    Topology topo;
    std::string error;
};

bool prepare(const jme_system_t* s, Sys& out) {
    out.s = s;
    std::unordered_map<int, VdwRow> rows;
    for (int t = 0; t < s->n_vdw_types; ++t)
        rows[s->vdw_type[t]] = VdwRow{s->vdw_alpha[t], s->vdw_N[t], s->vdw_A[t], s->vdw_G[t], (int)s->vdw_DA[t]};
    out.row_of_atom.resize(s->n_atoms);
    for (int64_t a = 0; a < s->n_atoms; ++a) {
        auto it = rows.find(s->atom_type[a]);
        if (it == rows.end()) { out.error = "atom type without vdW row"; return false; }
        out.row_of_atom[a] = it->second;
    }
    for (int64_t b = 0; b < s->n_bonds; ++b) {
        out.r0_of_bond[{s->bond_i[b], s->bond_j[b]}] = s->bond_r0[b];
        out.r0_of_bond[{s->bond_j[b], s->bond_i[b]}] = s->bond_r0[b];
    }
    out.topo = build_topology(s->n_atoms, s->n_bonds, s->bond_i, s->bond_j);
    return true;
}

template <class T> P3<T> pt(const double* xyz, int64_t a) { return P3<T>{T(xyz[3 * a]), T(xyz[3 * a + 1]), T(xyz[3 * a + 2])}; }

// seeds partial derivative slots: atom `slot` of the term owns d[3*slot .. 3*slot+2]
template <int N> P3<Dual<N>> pt_seed(const double* xyz, int64_t a, int slot) {
    P3<Dual<N>> p{Dual<N>(xyz[3 * a]), Dual<N>(xyz[3 * a + 1]), Dual<N>(xyz[3 * a + 2])};
    p.x.d[3 * slot] = 1; p.y.d[3 * slot + 1] = 1; p.z.d[3 * slot + 2] = 1;
    return p;
}
template <int N> void scatter(const Dual<N>& e, const int64_t* atoms, int n_atoms_in_term, double* grad) {
    for (int a = 0; a < n_atoms_in_term; ++a)
        for (int c = 0; c < 3; ++c) grad[3 * atoms[a] + c] += e.d[3 * a + c];
}

inline int pair_class(const Topology& t, int i, int j) {   // 0 = interacts, 1 = excluded, 2 = 1-4
    if (std::binary_search(t.excl[i].begin(), t.excl[i].end(), j)) return 1;
    if (std::binary_search(t.p14[i].begin(), t.p14[i].end(), j)) return 2;
    return 0;
}

// Full evaluation in double; when grad != nullptr also the AD gradient.  Sums run in the order the
// arrays are given (the reference iterates HashMaps in identity-hash order - SURVEY.md appendix A.4 -
// so no order is "the" order; this one is fixed and documented).  vdW and electrostatics are
// accumulated separately like the two component passes (MMFF94.java:390-392).
int evaluate(const Sys& S, const double* xyz, double cutoff, double dielectric, int unit,
             double comps[7], double* total, double* grad, uint64_t* n_pairs, uint64_t* n_candidates) {
    const jme_system_t* s = S.s;
    for (int c = 0; c < 7; ++c) comps[c] = 0;
    if (grad) std::memset(grad, 0, sizeof(double) * 3 * s->n_atoms);
    for (int64_t b = 0; b < s->n_bonds; ++b) {
        int64_t at[2] = {s->bond_i[b], s->bond_j[b]};
        if (grad) {
            Dual<6> e = e_bond(pt_seed<6>(xyz, at[0], 0), pt_seed<6>(xyz, at[1], 1), s->bond_kb[b], s->bond_r0[b], unit);
            comps[JME_BOND] += e.v; scatter(e, at, 2, grad);
        } else {
            comps[JME_BOND] += e_bond(pt<double>(xyz, at[0]), pt<double>(xyz, at[1]), s->bond_kb[b], s->bond_r0[b], unit);
        }
    }
    for (int64_t a = 0; a < s->n_angles; ++a) {
        int64_t at[3] = {s->angle_i[a], s->angle_j[a], s->angle_k[a]};
        bool lin = s->linear[at[1]] != 0;
        auto r1 = S.r0_of_bond.find({(int)at[0], (int)at[1]});
        auto r2 = S.r0_of_bond.find({(int)at[1], (int)at[2]});
        if (r1 == S.r0_of_bond.end() || r2 == S.r0_of_bond.end()) return JME_ERR_INVALID;
        if (grad) {
            auto pi = pt_seed<9>(xyz, at[0], 0); auto pj = pt_seed<9>(xyz, at[1], 1); auto pk = pt_seed<9>(xyz, at[2], 2);
            Dual<9> e = e_angle(pi, pj, pk, s->angle_ka[a], s->angle_theta0[a], lin, unit);
            comps[JME_ANGLE] += e.v; scatter(e, at, 3, grad);
            Dual<9> f = e_stbn(pi, pj, pk, s->angle_theta0[a], s->stbn_kba_ijk[a], s->stbn_kba_kji[a], r1->second, r2->second, lin, unit);
            comps[JME_STRETCH_BEND] += f.v; scatter(f, at, 3, grad);
        } else {
            auto pi = pt<double>(xyz, at[0]); auto pj = pt<double>(xyz, at[1]); auto pk = pt<double>(xyz, at[2]);
            comps[JME_ANGLE] += e_angle(pi, pj, pk, s->angle_ka[a], s->angle_theta0[a], lin, unit);
            comps[JME_STRETCH_BEND] += e_stbn(pi, pj, pk, s->angle_theta0[a], s->stbn_kba_ijk[a], s->stbn_kba_kji[a], r1->second, r2->second, lin, unit);
        }
    }
    for (int64_t o = 0; o < s->n_oops; ++o) {
        int64_t at[4] = {s->oop_i[o], s->oop_j[o], s->oop_k[o], s->oop_l[o]};
        if (grad) {
            Dual<12> e = e_oop(pt_seed<12>(xyz, at[0], 0), pt_seed<12>(xyz, at[1], 1), pt_seed<12>(xyz, at[2], 2), pt_seed<12>(xyz, at[3], 3), s->oop_koop[o], unit);
            comps[JME_OOP] += e.v; scatter(e, at, 4, grad);
        } else {
            comps[JME_OOP] += e_oop(pt<double>(xyz, at[0]), pt<double>(xyz, at[1]), pt<double>(xyz, at[2]), pt<double>(xyz, at[3]), s->oop_koop[o], unit);
        }
    }
    for (int64_t t = 0; t < s->n_torsions; ++t) {
        int64_t at[4] = {s->tor_i[t], s->tor_j[t], s->tor_k[t], s->tor_l[t]};
        int ti = s->atom_type[at[0]], tj = s->atom_type[at[1]], tk = s->atom_type[at[2]], tl = s->atom_type[at[3]];
        if (grad) {
            Dual<12> e = e_torsion(pt_seed<12>(xyz, at[0], 0), pt_seed<12>(xyz, at[1], 1), pt_seed<12>(xyz, at[2], 2), pt_seed<12>(xyz, at[3], 3),
                                   ti, tj, tk, tl, (int)at[1], (int)at[2], s->tor_v1[t], s->tor_v2[t], s->tor_v3[t], unit);
            comps[JME_TORSION] += e.v; scatter(e, at, 4, grad);
        } else {
            comps[JME_TORSION] += e_torsion(pt<double>(xyz, at[0]), pt<double>(xyz, at[1]), pt<double>(xyz, at[2]), pt<double>(xyz, at[3]),
                                            ti, tj, tk, tl, (int)at[1], (int)at[2], s->tor_v1[t], s->tor_v2[t], s->tor_v3[t], unit);
        }
    }
    uint64_t pairs = 0, cand = 0;
    for (int64_t i = 0; i < s->n_atoms; ++i) {
        for (int64_t j = i + 1; j < s->n_atoms; ++j) {
            int cls = pair_class(S.topo, (int)i, (int)j);
            if (cls == 1) continue;
            ++cand;
            // the reference evaluates R*, eps before the cutoff test (MMFF94VdwComponent.java:106-139);
            // they do not depend on the distance, so testing first changes no result, only the cost
            if (cutoff > 0 && distance(pt<double>(xyz, i), pt<double>(xyz, j)) > cutoff) continue;
            double rstar, eps;
            vdw_constants(S.row_of_atom[i], S.row_of_atom[j], rstar, eps);
            int64_t at[2] = {i, j};
            if (grad) {
                Dual<6> r = distance(pt_seed<6>(xyz, i, 0), pt_seed<6>(xyz, j, 1));
                ++pairs;
                Dual<6> ev = e_vdw(r, rstar, eps, unit);
                Dual<6> ee = e_ele(r, s->charge[i], s->charge[j], dielectric, cls == 2, unit);
                comps[JME_VDW] += ev.v; comps[JME_ELE] += ee.v;
                scatter(ev, at, 2, grad); scatter(ee, at, 2, grad);
            } else {
                double r = distance(pt<double>(xyz, i), pt<double>(xyz, j));
                ++pairs;
                comps[JME_VDW] += e_vdw(r, rstar, eps, unit);
                comps[JME_ELE] += e_ele(r, s->charge[i], s->charge[j], dielectric, cls == 2, unit);
            }
        }
    }
    double e = 0;
    for (int c = 0; c < 7; ++c) e += comps[c];   // MMFF94.java:390-392
    *total = e;
    if (n_pairs) *n_pairs = pairs;
    if (n_candidates) *n_candidates = cand;
    return JME_OK;
}

// Sum of every term that involves atom `a`, in scalar type T - the only part of E that moves when a
// coordinate of `a` moves; used for finite differences.
template <class T>
T energy_involving(const Sys& S, const std::vector<T>& x, int64_t a, double cutoff, double dielectric, int unit) {
    const jme_system_t* s = S.s;
    auto P = [&](int64_t n) { return P3<T>{x[3 * n], x[3 * n + 1], x[3 * n + 2]}; };
    T e = T(0.0);
    for (int64_t b = 0; b < s->n_bonds; ++b)
        if (s->bond_i[b] == a || s->bond_j[b] == a) e = e + e_bond(P(s->bond_i[b]), P(s->bond_j[b]), s->bond_kb[b], s->bond_r0[b], unit);
    for (int64_t n = 0; n < s->n_angles; ++n) {
        int64_t i = s->angle_i[n], j = s->angle_j[n], k = s->angle_k[n];
        if (i != a && j != a && k != a) continue;
        bool lin = s->linear[j] != 0;
        e = e + e_angle(P(i), P(j), P(k), s->angle_ka[n], s->angle_theta0[n], lin, unit);
        e = e + e_stbn(P(i), P(j), P(k), s->angle_theta0[n], s->stbn_kba_ijk[n], s->stbn_kba_kji[n],
                       S.r0_of_bond.at({(int)i, (int)j}), S.r0_of_bond.at({(int)j, (int)k}), lin, unit);
    }
    for (int64_t n = 0; n < s->n_oops; ++n) {
        int64_t i = s->oop_i[n], j = s->oop_j[n], k = s->oop_k[n], l = s->oop_l[n];
        if (i != a && j != a && k != a && l != a) continue;
        e = e + e_oop(P(i), P(j), P(k), P(l), s->oop_koop[n], unit);
    }
    for (int64_t n = 0; n < s->n_torsions; ++n) {
        int64_t i = s->tor_i[n], j = s->tor_j[n], k = s->tor_k[n], l = s->tor_l[n];
        if (i != a && j != a && k != a && l != a) continue;
        e = e + e_torsion(P(i), P(j), P(k), P(l), s->atom_type[i], s->atom_type[j], s->atom_type[k], s->atom_type[l],
                          (int)j, (int)k, s->tor_v1[n], s->tor_v2[n], s->tor_v3[n], unit);
    }
    for (int64_t b = 0; b < s->n_atoms; ++b) {
        if (b == a) continue;
        int cls = pair_class(S.topo, (int)a, (int)b);
        if (cls == 1) continue;
        double rstar, eps;
        int64_t lo = std::min(a, b), hi = std::max(a, b);
        vdw_constants(S.row_of_atom[lo], S.row_of_atom[hi], rstar, eps);
        T r = distance(P(lo), P(hi));
        if (cutoff > 0 && value_of(r) > cutoff) continue;
        e = e + e_vdw(r, rstar, eps, unit) + e_ele(r, s->charge[lo], s->charge[hi], dielectric, cls == 2, unit);
    }
    return e;
}

}  // namespace

// =============================================================================================
// C interface (ctypes) - names prefixed oracle_
// =============================================================================================
extern "C" {

// Energy (and AD gradient when grad != NULL) with the reference-faithful per-pair arithmetic.
int oracle_evaluate(const jme_system_t* s, const double* xyz, double cutoff, double dielectric, int unit,
                    double comps[7], double* total, double* grad, uint64_t* n_pairs, uint64_t* n_candidates) {
    Sys S;
    if (!prepare(s, S)) return JME_ERR_INVALID;
    return evaluate(S, xyz, cutoff, dielectric, unit, comps, total, grad, n_pairs, n_candidates);
}

// __float128 central differences of the restated energy for the listed atoms: out[n][3].
int oracle_fd_gradient(const jme_system_t* s, const double* xyz, double cutoff, double dielectric, int unit,
                       const int64_t* atoms, int64_t n_sel, double h, double* out) {
    Sys S;
    if (!prepare(s, S)) return JME_ERR_INVALID;
    std::vector<__float128> x(3 * s->n_atoms);
    for (int64_t n = 0; n < 3 * s->n_atoms; ++n) x[n] = xyz[n];
    for (int64_t n = 0; n < n_sel; ++n) {
        int64_t a = atoms[n];
        for (int c = 0; c < 3; ++c) {
            __float128 keep = x[3 * a + c];
            x[3 * a + c] = keep + (__float128)h;
            __float128 ep = energy_involving<__float128>(S, x, a, cutoff, dielectric, unit);
            x[3 * a + c] = keep - (__float128)h;
            __float128 em = energy_involving<__float128>(S, x, a, cutoff, dielectric, unit);
            x[3 * a + c] = keep;
            out[3 * n + c] = (double)((ep - em) / (2 * (__float128)h));
        }
    }
    return JME_OK;
}

// Selected pair set (i<j): not excluded and not beyond the cutoff (MMFF94.java:162-170 +
// MMFF94VdwComponent.java:136-139).  Returns the count; fills up to `capacity`.
int64_t oracle_pair_list(const jme_system_t* s, const double* xyz, double cutoff,
                         int32_t* pi, int32_t* pj, uint8_t* is14, int64_t capacity) {
    Topology t = build_topology(s->n_atoms, s->n_bonds, s->bond_i, s->bond_j);
    int64_t n = 0;
    for (int64_t i = 0; i < s->n_atoms; ++i)
        for (int64_t j = i + 1; j < s->n_atoms; ++j) {
            int cls = pair_class(t, (int)i, (int)j);
            if (cls == 1) continue;
            if (cutoff > 0) {
                double r = distance(pt<double>(xyz, i), pt<double>(xyz, j));
                if (r > cutoff) continue;
            }
            if (n < capacity) { pi[n] = (int32_t)i; pj[n] = (int32_t)j; is14[n] = cls == 2; }
            ++n;
        }
    return n;
}

// Term enumeration of MMFF94.java:158-212 on index arrays (first-insertion order, HashMap.put
// de-duplication).  Outputs may be NULL to query counts.
int oracle_enumerate_terms(int64_t n_atoms, const int32_t* type, const uint8_t* linear, int64_t n_bonds,
                           const int32_t* bi, const int32_t* bj, int32_t* angles, int64_t* n_angles,
                           int32_t* oops, int64_t* n_oops, int32_t* torsions, int64_t* n_torsions) {
    Topology t = build_topology(n_atoms, n_bonds, bi, bj);
    std::vector<std::array<int, 3>> A;
    std::vector<std::array<int, 4>> O, Tn;
    std::set<std::array<int, 3>> sa;
    std::set<std::array<int, 4>> so, st;
    for (int atom = 0; atom < n_atoms; ++atom) {
        const auto& nb = t.adj[atom];
        if (nb.size() < 2) continue;
        int tj = type[atom];
        for (int atom2 : nb)
            for (int atom3 : nb) {
                if (atom3 == atom2) continue;
                int ti = type[atom2], tk = type[atom3];
                bool first = ti < tk || (ti == tk && atom2 < atom3);
                if (first && sa.insert({atom2, atom, atom3}).second) A.push_back({atom2, atom, atom3});
                if (nb.size() >= 3 && first)
                    for (int atom4 : nb) {
                        if (atom4 == atom2 || atom4 == atom3) continue;
                        if (so.insert({atom2, atom, atom3, atom4}).second) O.push_back({atom2, atom, atom3, atom4});
                    }
                if (!linear[atom] && !linear[atom3] && t.adj[atom3].size() >= 2)
                    for (int atom4 : t.adj[atom3]) {
                        if (atom4 == atom || atom4 == atom2) continue;
                        int tl = type[atom4];
                        if (tj < tk || (tj == tk && ti < tl) || (tj == tk && ti == tl && atom < atom3))
                            if (st.insert({atom2, atom, atom3, atom4}).second) Tn.push_back({atom2, atom, atom3, atom4});
                    }
            }
    }
    if (angles) for (size_t n = 0; n < A.size(); ++n) for (int c = 0; c < 3; ++c) angles[3 * n + c] = A[n][c];
    if (oops) for (size_t n = 0; n < O.size(); ++n) for (int c = 0; c < 4; ++c) oops[4 * n + c] = O[n][c];
    if (torsions) for (size_t n = 0; n < Tn.size(); ++n) for (int c = 0; c < 4; ++c) torsions[4 * n + c] = Tn[n][c];
    *n_angles = (int64_t)A.size(); *n_oops = (int64_t)O.size(); *n_torsions = (int64_t)Tn.size();
    return JME_OK;
}

// The strict-double cutoff predicate as a threshold on r^2: largest double T with sqrt(T) <= cutoff,
// so that (sqrt(r2) > cutoff) == (r2 > T) for every double r2.  Restated here independently of the
// product's own computation so that tests can compare the two.
double oracle_cutoff_r2_threshold(double cutoff) {
    double t = cutoff * cutoff;
    while (std::sqrt(t) > cutoff) t = std::nextafter(t, 0.0);
    while (std::sqrt(std::nextafter(t, INFINITY)) <= cutoff) t = std::nextafter(t, INFINITY);
    return t;
}


int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---------------------------------------------------------------------------------------------
// Multi-threaded nonbonded evaluation for large systems: the timed CPU baseline and the checker
// for systems where the O(N^2) scan above is too slow.  Same pair predicate, same energy
// expressions; R*, eps come from a type-pair table (or are recomputed per pair when
// faithful_constants != 0, which is what the reference does, MMFF94VdwComponent.java:106-135);
// x^7 by multiplication and a hand-derived radial derivative instead of AD; Newton's third law
// with one gradient buffer per thread; a cell list (edge = cutoff) when cutoff > 0.  Its results are
// checked against oracle_evaluate in tests/test_oracle_fast.py.
// Bonded terms are added serially with the AD path (they are << 1 % of the work).
// ---------------------------------------------------------------------------------------------
int oracle_fast_evaluate(const jme_system_t* s, const double* xyz, double cutoff, double dielectric, int unit,
                         int threads, int faithful_constants, double comps[7], double* total, double* grad,
                         uint64_t* n_pairs) {
    Sys S;
    if (!prepare(s, S)) return JME_ERR_INVALID;
    const int64_t n = s->n_atoms;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    // bonded part (and zeroing) through the AD evaluator restricted to bonded terms
    {
        jme_system_t b = *s;   // same system, no atoms' pair loop: evaluate bonded by zeroing n_atoms' pair scan
        // evaluate() always scans pairs; call the term loops directly instead
        for (int c = 0; c < 7; ++c) comps[c] = 0;
        if (grad) std::memset(grad, 0, sizeof(double) * 3 * n);
        (void)b;
        for (int64_t q = 0; q < s->n_bonds; ++q) {
            int64_t at[2] = {s->bond_i[q], s->bond_j[q]};
            Dual<6> e = e_bond(pt_seed<6>(xyz, at[0], 0), pt_seed<6>(xyz, at[1], 1), s->bond_kb[q], s->bond_r0[q], unit);
            comps[JME_BOND] += e.v; if (grad) scatter(e, at, 2, grad);
        }
        for (int64_t a = 0; a < s->n_angles; ++a) {
            int64_t at[3] = {s->angle_i[a], s->angle_j[a], s->angle_k[a]};
            bool lin = s->linear[at[1]] != 0;
            auto r1 = S.r0_of_bond.find({(int)at[0], (int)at[1]});
            auto r2 = S.r0_of_bond.find({(int)at[1], (int)at[2]});
            if (r1 == S.r0_of_bond.end() || r2 == S.r0_of_bond.end()) return JME_ERR_INVALID;
            auto pi = pt_seed<9>(xyz, at[0], 0); auto pj = pt_seed<9>(xyz, at[1], 1); auto pk = pt_seed<9>(xyz, at[2], 2);
            Dual<9> e = e_angle(pi, pj, pk, s->angle_ka[a], s->angle_theta0[a], lin, unit);
            comps[JME_ANGLE] += e.v; if (grad) scatter(e, at, 3, grad);
            Dual<9> f = e_stbn(pi, pj, pk, s->angle_theta0[a], s->stbn_kba_ijk[a], s->stbn_kba_kji[a], r1->second, r2->second, lin, unit);
            comps[JME_STRETCH_BEND] += f.v; if (grad) scatter(f, at, 3, grad);
        }
        for (int64_t o = 0; o < s->n_oops; ++o) {
            int64_t at[4] = {s->oop_i[o], s->oop_j[o], s->oop_k[o], s->oop_l[o]};
            Dual<12> e = e_oop(pt_seed<12>(xyz, at[0], 0), pt_seed<12>(xyz, at[1], 1), pt_seed<12>(xyz, at[2], 2), pt_seed<12>(xyz, at[3], 3), s->oop_koop[o], unit);
            comps[JME_OOP] += e.v; if (grad) scatter(e, at, 4, grad);
        }
        for (int64_t t = 0; t < s->n_torsions; ++t) {
            int64_t at[4] = {s->tor_i[t], s->tor_j[t], s->tor_k[t], s->tor_l[t]};
            Dual<12> e = e_torsion(pt_seed<12>(xyz, at[0], 0), pt_seed<12>(xyz, at[1], 1), pt_seed<12>(xyz, at[2], 2), pt_seed<12>(xyz, at[3], 3),
                                   s->atom_type[at[0]], s->atom_type[at[1]], s->atom_type[at[2]], s->atom_type[at[3]],
                                   (int)at[1], (int)at[2], s->tor_v1[t], s->tor_v2[t], s->tor_v3[t], unit);
            comps[JME_TORSION] += e.v; if (grad) scatter(e, at, 4, grad);
        }
    }
    // dense type index + type-pair table
    std::vector<int> tindex(n);
    std::vector<int> types;
    {
        std::map<int, int> idx;
        for (int64_t a = 0; a < n; ++a) {
            auto it = idx.find(s->atom_type[a]);
            if (it == idx.end()) { it = idx.emplace(s->atom_type[a], (int)types.size()).first; types.push_back((int)a); }
            tindex[a] = it->second;
        }
    }
    const int T = (int)types.size();
    std::vector<double> tab_r(T * T), tab_e(T * T);
    for (int a = 0; a < T; ++a)
        for (int b = 0; b < T; ++b) {
            // the reference orders the pair by atom index (i<j); R*, eps are symmetric in exact arithmetic
            // and the table is filled with (min,max) ordering of the dense index to make it symmetric
            int lo = std::min(a, b), hi = std::max(a, b);
            vdw_constants(S.row_of_atom[types[lo]], S.row_of_atom[types[hi]], tab_r[a * T + b], tab_e[a * T + b]);
        }
    const double unit_scale_num = (unit == JME_KCAL_PER_MOL) ? 1.0 : UNIT_TO_J[JME_KCAL_PER_MOL];
    const double unit_scale_den = (unit == JME_KCAL_PER_MOL) ? 1.0 : UNIT_TO_J[unit];

    std::vector<std::vector<double>> tg(threads);
    std::vector<double> t_vdw(threads, 0.0), t_ele(threads, 0.0);
    std::vector<uint64_t> t_pairs(threads, 0);

    auto pair_kernel = [&](int64_t i, int64_t j, double* g, double& ev_acc, double& ee_acc, uint64_t& cnt) {
        // i < j required by callers
        int cls = 0;
        if (!S.topo.excl[i].empty() || !S.topo.p14[i].empty()) cls = pair_class(S.topo, (int)i, (int)j);
        if (cls == 1) return;
        double dx = xyz[3 * i] - xyz[3 * j], dy = xyz[3 * i + 1] - xyz[3 * j + 1], dz = xyz[3 * i + 2] - xyz[3 * j + 2];
        double r = std::sqrt(dx * dx + dy * dy + dz * dz);
        if (cutoff > 0 && r > cutoff) return;
        ++cnt;
        double R, eps;
        if (faithful_constants) vdw_constants(S.row_of_atom[i], S.row_of_atom[j], R, eps);
        else { R = tab_r[tindex[i] * T + tindex[j]]; eps = tab_e[tindex[i] * T + tindex[j]]; }
        double sden = r + 0.07 * R;
        double a = (1.07 * R) / sden;
        double a2 = a * a, a4 = a2 * a2, a7 = a4 * a2 * a;
        double R2 = R * R, R4 = R2 * R2, R7 = R4 * R2 * R;
        double r2 = r * r, r4 = r2 * r2, r6 = r4 * r2, r7 = r6 * r;
        double tden = r7 + 0.12 * R7;
        double b = (1.12 * R7) / tden;
        double ev = eps * a7 * (b - 2.0);
        double dev = eps * a7 * (-7.0 * (b - 2.0) / sden - 7.0 * b * r6 / tden);
        double u = r + 0.05;
        double ee = 332.0716 * s->charge[i] * s->charge[j] / (dielectric * u);
        if (cls == 2) ee *= 0.75;
        double dee = -ee / u;
        ev = (ev * unit_scale_num) / unit_scale_den;
        ee = (ee * unit_scale_num) / unit_scale_den;
        ev_acc += ev; ee_acc += ee;
        if (g) {
            double de = ((dev + dee) * unit_scale_num) / unit_scale_den;
            double f = r > 0 ? de / r : 0.0;
            g[3 * i] += f * dx; g[3 * i + 1] += f * dy; g[3 * i + 2] += f * dz;
            g[3 * j] -= f * dx; g[3 * j + 1] -= f * dy; g[3 * j + 2] -= f * dz;
        }
    };

    if (cutoff <= 0) {
#pragma omp parallel num_threads(threads)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            if (grad) tg[tid].assign(3 * n, 0.0);
            double ev = 0, ee = 0; uint64_t cnt = 0;
            double* g = grad ? tg[tid].data() : nullptr;
#pragma omp for schedule(dynamic, 16)
            for (int64_t i = 0; i < n; ++i)
                for (int64_t j = i + 1; j < n; ++j) pair_kernel(i, j, g, ev, ee, cnt);
            t_vdw[tid] = ev; t_ele[tid] = ee; t_pairs[tid] = cnt;
        }
    } else {
        // cell list, edge = cutoff, half shell
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int64_t a = 0; a < n; ++a)
            for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], xyz[3 * a + c]); hi[c] = std::max(hi[c], xyz[3 * a + c]); }
        int64_t dim[3];
        for (int c = 0; c < 3; ++c) dim[c] = std::max<int64_t>(1, (int64_t)std::floor((hi[c] - lo[c]) / cutoff) + 1);
        int64_t ncell = dim[0] * dim[1] * dim[2];
        std::vector<int64_t> cell_of(n), start(ncell + 1, 0);
        for (int64_t a = 0; a < n; ++a) {
            int64_t c3[3];
            for (int c = 0; c < 3; ++c) c3[c] = std::min<int64_t>(dim[c] - 1, (int64_t)std::floor((xyz[3 * a + c] - lo[c]) / cutoff));
            cell_of[a] = (c3[2] * dim[1] + c3[1]) * dim[0] + c3[0];
            ++start[cell_of[a] + 1];
        }
        for (int64_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
        std::vector<int64_t> fill(start.begin(), start.end() - 1), order(n);
        for (int64_t a = 0; a < n; ++a) order[fill[cell_of[a]]++] = a;
#pragma omp parallel num_threads(threads)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            if (grad) tg[tid].assign(3 * n, 0.0);
            double ev = 0, ee = 0; uint64_t cnt = 0;
            double* g = grad ? tg[tid].data() : nullptr;
#pragma omp for schedule(dynamic, 4)
            for (int64_t c = 0; c < ncell; ++c) {
                int64_t cx = c % dim[0], cy = (c / dim[0]) % dim[1], cz = c / (dim[0] * dim[1]);
                for (int64_t p = start[c]; p < start[c + 1]; ++p) {
                    int64_t i = order[p];
                    for (int64_t q = p + 1; q < start[c + 1]; ++q) {
                        int64_t j = order[q];
                        if (i < j) pair_kernel(i, j, g, ev, ee, cnt); else pair_kernel(j, i, g, ev, ee, cnt);
                    }
                }
                // 13 forward neighbour cells
                for (int dz = 0; dz <= 1; ++dz)
                    for (int dy = (dz == 0 ? 0 : -1); dy <= 1; ++dy)
                        for (int dx = ((dz == 0 && dy == 0) ? 1 : -1); dx <= 1; ++dx) {
                            int64_t nx = cx + dx, ny = cy + dy, nz = cz + dz;
                            if (nx < 0 || ny < 0 || nz < 0 || nx >= dim[0] || ny >= dim[1] || nz >= dim[2]) continue;
                            int64_t d = (nz * dim[1] + ny) * dim[0] + nx;
                            for (int64_t p = start[c]; p < start[c + 1]; ++p) {
                                int64_t i = order[p];
                                for (int64_t q = start[d]; q < start[d + 1]; ++q) {
                                    int64_t j = order[q];
                                    if (i < j) pair_kernel(i, j, g, ev, ee, cnt); else pair_kernel(j, i, g, ev, ee, cnt);
                                }
                            }
                        }
            }
            t_vdw[tid] = ev; t_ele[tid] = ee; t_pairs[tid] = cnt;
        }
    }
    uint64_t pairs = 0;
    for (int t = 0; t < threads; ++t) { comps[JME_VDW] += t_vdw[t]; comps[JME_ELE] += t_ele[t]; pairs += t_pairs[t]; }
    if (grad)
        for (int t = 0; t < threads; ++t)
            if (!tg[t].empty())
                for (int64_t q = 0; q < 3 * n; ++q) grad[q] += tg[t][q];
    double e = 0;
    for (int c = 0; c < 7; ++c) e += comps[c];
    *total = e;
    if (n_pairs) *n_pairs = pairs;
    return JME_OK;
}

}  // extern "C"
