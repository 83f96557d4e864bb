// jme_math.cuh - per-term MMFF94 energy + analytic Cartesian gradient, shared by every kernel.
//
// All functions are __host__ __device__ so that tests/hostcheck can compile the very same
// expressions with g++ and compare them with the oracle's automatic differentiation before
// any GPU time is spent.  Values follow the reference's expressions (file:line cited per
// function, paths under /root/reference/src/main/java/org/jme/forcefield); the gradients
// have no counterpart in the reference (it has none) and are derived here.
//
// Everything is FP64.  Energies are in kcal/mol; unit conversion (ForceField.java:153-158) is
// applied to the component sums by the final reduction kernel.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define JME_HD __host__ __device__ __forceinline__
#else
#define JME_HD inline
#endif

namespace jme {

constexpr double kPi = 3.14159265358979323846;              // java.lang.Math.PI
constexpr double kRadToDeg = 180.0 / kPi;
constexpr double kBondC = .5 * 143.9325;                      // MMFF94BondStretchingComponent.java:105
constexpr double kCubicStretch4 = (7.0 / 12.0) * 4;           // same line
constexpr double kAngleC = 143.9325 / 180.0 * kPi / 180.0 * kPi;  // Math.toRadians(Math.toRadians(143.9325)), AngleBending :124
constexpr double kAngleCb = 0.006981317;                      // AngleBending :124
constexpr double kAngleLin = 143.9325;                        // AngleBending :122
constexpr double kStbnC = 2.51210;                            // StretchBend :134
constexpr double kEleC = 332.0716;                            // Electrostatic :142
constexpr double kEleBuf = 0.05;                              // Electrostatic :142
constexpr double kEle14 = 0.75;                               // Electrostatic :144

struct Vec3 {
    double x, y, z;
};
JME_HD Vec3 operator+(Vec3 a, Vec3 b) { return Vec3{a.x + b.x, a.y + b.y, a.z + b.z}; }
JME_HD Vec3 operator-(Vec3 a, Vec3 b) { return Vec3{a.x - b.x, a.y - b.y, a.z - b.z}; }
JME_HD Vec3 operator*(double s, Vec3 a) { return Vec3{s * a.x, s * a.y, s * a.z}; }
JME_HD Vec3 operator-(Vec3 a) { return Vec3{-a.x, -a.y, -a.z}; }
JME_HD double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
JME_HD Vec3 cross(Vec3 a, Vec3 b) { return Vec3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
JME_HD double norm(Vec3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
JME_HD double clamp1(double n) { return n > 1 ? 1 : (n < -1 ? -1 : n); }     // GeometryUtils.java:73-75

// ---------------------------------------------------------------------------------------------
// Nonbonded pair: buffered 14-7 vdW (MMFF94VdwComponent.java:140-141) + buffered Coulomb
// (MMFF94ElectrostaticComponent.java:142-145).
//
// Type-pair constants (precomputed on the host with the reference's operation order, see
// jme_api.cu build_pair_table):   c2 = 0.07 R*   c3 = 1.12 R*^7   c4 = 0.12 R*^7
//                                 k1 = 2 eps * (1.07 R*)^7   (doubled, see pair_terms)
// so that  2 E_vdw = 2 eps (1.07R/(r+0.07R))^7 (1.12R^7/(r^7+0.12R^7) - 2) = k1 s^-7 (c3/t - 2)
// with s = r + c2, t = r^7 + c4.
// ---------------------------------------------------------------------------------------------
struct alignas(16) PairConst {      // two 128-bit loads
    double k1, c2, c3, c4;
};

// r and 1/(2r) from r2 (> 0).  Device: MUFU.RSQ64H seed (about 20 bits) + ONE third-order (Halley) step:
// with e = 1 - r2 y0^2,  1/sqrt(r2) = y0 (1 - e)^(-1/2) = y0 (1 + e/2 + 3e^2/8 + O(e^3)); e ~ 2^-20 leaves a
// relative error ~ 2^-58, i.e. <= 1 ulp after the final roundings, with no slow path.  The host build of this
// header (tests/hostcheck) uses libm.
JME_HD void sqrt_and_half_rsqrt(double r2, double& r, double& half_rinv) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));
    const double t = r2 * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double q = e * p;
    y = fma(y, q, y);
    r = r2 * y;
    half_rinv = 0.5 * y;
#else
    r = sqrt(r2);
    half_rinv = 0.5 / r;
#endif
}

// 1/x for finite x well inside the normal range.  Device: MUFU.RCP64H seed + one third-order step:
// e = 1 - x y0,  1/x = y0 (1 + e + e^2 + O(e^3)).
JME_HD double fast_rcp(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double q = fma(e, e, e);
    return fma(y, q, y);
#else
    return 1.0 / x;
#endif
}

// One pair with r2 > 0.  To save one multiply per pair the kernels work with DOUBLED energies:
//   pc.k1 = 2 eps (1.07 R*)^7   and   qq2 = 2 * 332.0716 qi qj / D  (x0.75 for a 1-4 pair)
// so e2_vdw = 2 E_vdw, e2_ele = 2 E_ele, and f_over_r = (dE/dr)/r exactly (the factor 2 cancels against
// the 1/(2r) the square-root step returns):  dE/dx_i = f_over_r * (x_i - x_j).  Callers halve the energy
// sums (exact: a power of two).
template <bool GRAD>
JME_HD void pair_terms(double r2, const PairConst& pc, double qq2, double& e2_vdw, double& e2_ele, double& f_over_r) {
    double r, hrinv;
    sqrt_and_half_rsqrt(r2, r, hrinv);
    const double s = r + pc.c2;
    const double u = r + kEleBuf;
    const double r4 = r2 * r2;
    const double r6 = r4 * r2;
    const double t = fma(r6, r, pc.c4);
    // one reciprocal for 1/s, 1/t, 1/u
    const double st = s * t;
    const double iq = fast_rcp(st * u);
    const double iu = iq * st;
    const double ist = iq * u;
    const double is = ist * t;
    const double it = ist * s;
    const double is2 = is * is;
    const double is4 = is2 * is2;
    const double is7 = is4 * is2 * is;
    const double w = pc.k1 * is7;
    const double b = pc.c3 * it;
    e2_vdw = w * (b - 2.0);
    e2_ele = qq2 * iu;
    if (GRAD) {
        // dE_vdw/dr = -7 [ E_vdw/s + w b r^6 / t ]      dE_ele/dr = -E_ele/u
        const double d = fma(w, (b * r6) * it, e2_vdw * is);
        const double g = fma(7.0, d, e2_ele * iu);
        f_over_r = -(g * hrinv);
    } else {
        f_over_r = 0.0;
    }
}

// The same arithmetic for N independent pairs, written statement by statement across the pairs: the N dependency
// chains reach the assembler already interleaved, which is what the FP64 pipe wants (one warp instruction per two
// cycles, eight cycles of dependent latency) when few warps are resident.  Bit-identical to N calls of pair_terms.
template <bool GRAD, int N>
JME_HD void pair_terms_n(const double (&r2)[N], const PairConst (&pc)[N], const double (&qq2)[N],
                         double (&e2_vdw)[N], double (&e2_ele)[N], double (&f_over_r)[N]) {
    double r[N], hrinv[N];
#if defined(__CUDA_ARCH__)
    double y[N], t0[N], e0[N], p0[N], q0[N];
#pragma unroll
    for (int i = 0; i < N; ++i) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(r2[i]));
#pragma unroll
    for (int i = 0; i < N; ++i) t0[i] = r2[i] * y[i];
#pragma unroll
    for (int i = 0; i < N; ++i) e0[i] = fma(-t0[i], y[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; ++i) p0[i] = fma(0.375, e0[i], 0.5);
#pragma unroll
    for (int i = 0; i < N; ++i) q0[i] = e0[i] * p0[i];
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = fma(y[i], q0[i], y[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = r2[i] * y[i];
#pragma unroll
    for (int i = 0; i < N; ++i) hrinv[i] = 0.5 * y[i];
#else
    for (int i = 0; i < N; ++i) { r[i] = sqrt(r2[i]); hrinv[i] = 0.5 / r[i]; }
#endif
    double s[N], u[N], r4[N], r6[N], t[N], st[N], x[N], iq[N];
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = r[i] + pc[i].c2;
#pragma unroll
    for (int i = 0; i < N; ++i) u[i] = r[i] + kEleBuf;
#pragma unroll
    for (int i = 0; i < N; ++i) r4[i] = r2[i] * r2[i];
#pragma unroll
    for (int i = 0; i < N; ++i) r6[i] = r4[i] * r2[i];
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = fma(r6[i], r[i], pc[i].c4);
#pragma unroll
    for (int i = 0; i < N; ++i) st[i] = s[i] * t[i];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = st[i] * u[i];
#if defined(__CUDA_ARCH__)
    {
        double e1[N], q1[N];
#pragma unroll
        for (int i = 0; i < N; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(iq[i]) : "d"(x[i]));
#pragma unroll
        for (int i = 0; i < N; ++i) e1[i] = fma(-x[i], iq[i], 1.0);
#pragma unroll
        for (int i = 0; i < N; ++i) q1[i] = fma(e1[i], e1[i], e1[i]);
#pragma unroll
        for (int i = 0; i < N; ++i) iq[i] = fma(iq[i], q1[i], iq[i]);
    }
#else
    for (int i = 0; i < N; ++i) iq[i] = 1.0 / x[i];
#endif
    double iu[N], ist[N], is[N], it[N], is2[N], is4[N], is7[N], w[N], b[N];
#pragma unroll
    for (int i = 0; i < N; ++i) iu[i] = iq[i] * st[i];
#pragma unroll
    for (int i = 0; i < N; ++i) ist[i] = iq[i] * u[i];
#pragma unroll
    for (int i = 0; i < N; ++i) is[i] = ist[i] * t[i];
#pragma unroll
    for (int i = 0; i < N; ++i) it[i] = ist[i] * s[i];
#pragma unroll
    for (int i = 0; i < N; ++i) is2[i] = is[i] * is[i];
#pragma unroll
    for (int i = 0; i < N; ++i) is4[i] = is2[i] * is2[i];
#pragma unroll
    for (int i = 0; i < N; ++i) is7[i] = is4[i] * is2[i] * is[i];
#pragma unroll
    for (int i = 0; i < N; ++i) w[i] = pc[i].k1 * is7[i];
#pragma unroll
    for (int i = 0; i < N; ++i) b[i] = pc[i].c3 * it[i];
#pragma unroll
    for (int i = 0; i < N; ++i) e2_vdw[i] = w[i] * (b[i] - 2.0);
#pragma unroll
    for (int i = 0; i < N; ++i) e2_ele[i] = qq2[i] * iu[i];
    if (GRAD) {
        double d[N], g[N];
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = fma(w[i], (b[i] * r6[i]) * it[i], e2_vdw[i] * is[i]);
#pragma unroll
        for (int i = 0; i < N; ++i) g[i] = fma(7.0, d[i], e2_ele[i] * iu[i]);
#pragma unroll
        for (int i = 0; i < N; ++i) f_over_r[i] = -(g[i] * hrinv[i]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) f_over_r[i] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// Bonded terms.  Each returns the energy and writes dE/d(position) of its atoms.
// ---------------------------------------------------------------------------------------------

// MMFF94BondStretchingComponent.java:99-111
JME_HD double bond_term(Vec3 pi, Vec3 pj, double kb, double r0, Vec3& gi, Vec3& gj) {
    const Vec3 d = pi - pj;
    const double len = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
    const double dr = len - r0;
    const double poly = 1 - 2 * dr + kCubicStretch4 * dr * dr;
    const double e = kBondC * kb * dr * dr * poly;
    const double dpoly = -2 + 2 * kCubicStretch4 * dr;
    const double de = kBondC * kb * (2 * dr * poly + dr * dr * dpoly);
    const double f = len > 0 ? de / len : 0.0;
    gi = f * d;
    gj = -gi;
    return e;
}

// Angle bend (MMFF94AngleBendingComponent.java:116-131) and stretch-bend
// (MMFF94StretchBendComponent.java:123-140) share the key (i,j,k) and the angle, so they are one term
// here.  theta from GeometryUtils.calculateAngle (GeometryUtils.java:43-51).
JME_HD void angle_stbn_term(Vec3 pi, Vec3 pj, Vec3 pk, double ka, double theta0, bool linear,
                            double kba_ijk, double kba_kji, double r0_ij, double r0_jk,
                            double& e_angle, double& e_stbn, Vec3& gi, Vec3& gj, Vec3& gk) {
    const Vec3 u = pi - pj, v = pk - pj;
    const double lu = sqrt(u.x * u.x + u.y * u.y + u.z * u.z);
    const double lv = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    const double craw = (u.x * v.x + u.y * v.y + u.z * v.z) / (lu * lv);
    const double c = clamp1(craw);
    const double theta = acos(c) * 180 / kPi;
    const double dth = theta - theta0;
    // dc/dr_i, dc/dr_k  (zero where the clamp is active)
    const double on = (craw > 1 || craw < -1) ? 0.0 : 1.0;
    const Vec3 uh = (1.0 / lu) * u, vh = (1.0 / lv) * v;
    const Vec3 dci = (on / lu) * (vh - c * uh);
    const Vec3 dck = (on / lv) * (uh - c * vh);
    double de_dc;          // dE/dc summed over both terms
    Vec3 gi_len{0, 0, 0}, gk_len{0, 0, 0};   // stretch part of the stretch-bend
    if (linear) {
        // 143.9325 ka (1 + cos(toRadians(theta)))  and  cos(theta) == c
        e_angle = kAngleLin * ka * (1 + cos(theta / 180.0 * kPi));
        e_stbn = 0.0;                         // StretchBend :125-127
        de_dc = kAngleLin * ka;
    } else {
        e_angle = kAngleC * ka * .5 * dth * dth * (1 - kAngleCb * dth);
        const double drij = lu - r0_ij, drjk = lv - r0_jk;
        const double sb = kba_ijk * drij + kba_kji * drjk;
        e_stbn = kStbnC * sb * dth;
        const double sin2 = (1.0 - c) * (1.0 + c);
        const double dth_dc = sin2 > 0 ? -kRadToDeg / sqrt(sin2) : 0.0;
        const double de_dth = kAngleC * ka * .5 * dth * (2 - 3 * kAngleCb * dth) + kStbnC * sb;
        de_dc = de_dth * dth_dc;
        gi_len = (kStbnC * dth * kba_ijk) * uh;
        gk_len = (kStbnC * dth * kba_kji) * vh;
    }
    gi = de_dc * dci + gi_len;
    gk = de_dc * dck + gk_len;
    gj = -(gi + gk);
}

// Wilson out-of-plane angle, MMFF94OutOfPlaneComponent.java:122-134 + GeometryUtils.java:85-95
// (asin argument NOT clamped, like the reference).
JME_HD double oop_term(Vec3 pi, Vec3 pj, Vec3 pk, Vec3 pl, double koop, Vec3& gi, Vec3& gj, Vec3& gk, Vec3& gl) {
    const Vec3 p = pi - pj, q = pk - pj, w = pl - pj;
    const double lp = sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
    const double lq = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    const double lw = sqrt(w.x * w.x + w.y * w.y + w.z * w.z);
    const double np = 1.0 / lp, nq = 1.0 / lq, nw = 1.0 / lw;      // vecmath normalize: scale by 1/len
    const Vec3 a = np * p, b = nq * q, d = nw * w;
    const Vec3 perp = cross(a, b);
    const double T = dot(perp, d);
    // sin(calculateAngle(i,j,k)): the reference recomputes the angle from the raw vectors
    const double craw = (p.x * q.x + p.y * q.y + p.z * q.z) / (lp * lq);
    const double c = clamp1(craw);
    const double sin_t = sin(acos(c));
    const double S = T / sin_t;
    const double chi = asin(S) * 180 / kPi;
    const double e = kAngleC * .5 * koop * chi * chi;
    // dE/dS
    const double cos2 = (1.0 - S) * (1.0 + S);
    const double de_dS = cos2 > 0 ? kAngleC * koop * chi * kRadToDeg / sqrt(cos2) : 0.0;
    // dS/d(unit vectors)
    const double is = 1.0 / sin_t;
    const double k = (craw > 1 || craw < -1) ? 0.0 : T * c * is * is * is;
    const Vec3 dSa = is * cross(b, d) + k * b;
    const Vec3 dSb = is * cross(d, a) + k * a;
    const Vec3 dSd = is * perp;
    // through the normalisations: (I - a a^T)/|p|
    const Vec3 ga = np * (dSa - dot(dSa, a) * a);
    const Vec3 gb = nq * (dSb - dot(dSb, b) * b);
    const Vec3 gd = nw * (dSd - dot(dSd, d) * d);
    gi = de_dS * ga;
    gk = de_dS * gb;
    gl = de_dS * gd;
    gj = -(gi + gk + gl);
    return e;
}

// MMFF94TorsionComponent.java:111-132 + GeometryUtils.java:61-71.  The caller passes the atoms in the
// order the reference evaluates them (after the swap of Torsion :117-124).
JME_HD double torsion_term(Vec3 pi, Vec3 pj, Vec3 pk, Vec3 pl, double v1, double v2, double v3,
                           Vec3& gi, Vec3& gj, Vec3& gk, Vec3& gl) {
    const Vec3 ij = pi - pj, kj = pk - pj, lk = pl - pk;
    const Vec3 jk = -kj;
    const Vec3 n1 = cross(ij, kj), n2 = cross(jk, lk);
    const double l1 = sqrt(n1.x * n1.x + n1.y * n1.y + n1.z * n1.z);
    const double l2 = sqrt(n2.x * n2.x + n2.y * n2.y + n2.z * n2.z);
    const double craw = (n1.x * n2.x + n1.y * n2.y + n1.z * n2.z) / (l1 * l2);
    const double c = clamp1(craw);
    const double phi = acos(c);
    const double e = 0.5 * (v1 * (1 + cos(phi)) + v2 * (1 - cos(2 * phi)) + v3 * (1 + cos(3 * phi)));
    // dE/dc with cos 2phi = 2c^2 - 1, cos 3phi = 4c^3 - 3c: no 1/sin(phi)
    const double on = (craw > 1 || craw < -1) ? 0.0 : 1.0;
    const double de_dc = on * 0.5 * (v1 - 4 * v2 * c + v3 * (12 * c * c - 3));
    const Vec3 h1 = (1.0 / l1) * n1, h2 = (1.0 / l2) * n2;
    const Vec3 A = (de_dc / l1) * (h2 - c * h1);     // dE/dn1
    const Vec3 B = (de_dc / l2) * (h1 - c * h2);     // dE/dn2
    const Vec3 d_ij = cross(kj, A);
    const Vec3 d_kj = cross(A, ij);
    const Vec3 d_jk = cross(lk, B);
    const Vec3 d_lk = cross(B, jk);
    gi = d_ij;
    gl = d_lk;
    gj = d_jk - d_ij - d_kj;
    gk = d_kj - d_jk - d_lk;
    return e;
}

}  // namespace jme
