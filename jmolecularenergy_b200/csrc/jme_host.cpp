// jme_host.cpp - see jme_host.hpp.  Compile with -ffp-contract=off: the type-pair constants must be
// rounded like the reference's Java expressions.
#include "jme_host.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <map>
#include <set>
#include <unordered_map>

namespace jme {

void vdw_pair_constants(float alpha_i, float N_i, float A_i, float G_i, int DA_i,
                        float alpha_j, float N_j, float A_j, float G_j, int DA_j,
                        double& rstar, double& eps) {
    // MMFF94VdwComponent.java:112-123 - the "like pair" branch (:112-114) equals the unlike formula
    // bit for bit when both rows coincide (gamma = 0), so one formula serves a numeric-type table.
    const double rii = A_i * std::pow((double)alpha_i, 0.25);
    const double rjj = A_j * std::pow((double)alpha_j, 0.25);
    const double B = (DA_i == 1 || DA_j == 1) ? 0 : 0.2;
    const double beta = 12;
    const double gamma = (rii - rjj) / (rii + rjj);
    rstar = 0.5 * (rii + rjj) * (1 + B * (1 - std::exp(-beta * std::pow(gamma, 2))));
    // :126-128, alpha / N is float / float
    eps = 181.16 * G_i * G_j * alpha_i * alpha_j;
    const float ai = alpha_i / N_i, aj = alpha_j / N_j;
    eps /= (std::sqrt((double)ai) + std::sqrt((double)aj));
    eps /= std::pow(rstar, 6);
    if (DA_i + DA_j == 3) {   // :130-135 donor + acceptor, after eps used the unscaled R*
        rstar *= 0.8;
        eps *= 0.5;
    }
}

void build_exclusions(int64_t n, int64_t nb, const int32_t* bi, const int32_t* bj,
                      std::vector<int64_t>& ptr, std::vector<uint32_t>& ent) {
    std::vector<int64_t> deg(n + 1, 0);
    for (int64_t b = 0; b < nb; ++b) { ++deg[bi[b] + 1]; ++deg[bj[b] + 1]; }
    for (int64_t a = 0; a < n; ++a) deg[a + 1] += deg[a];
    std::vector<int32_t> adj(2 * nb);
    {
        std::vector<int64_t> fill(deg.begin(), deg.end() - 1);
        for (int64_t b = 0; b < nb; ++b) { adj[fill[bi[b]]++] = bj[b]; adj[fill[bj[b]]++] = bi[b]; }
    }
    ptr.assign(n + 1, 0);
    ent.clear();
    std::vector<int8_t> depth(n, -1);
    std::vector<int32_t> seen, frontier, next;
    std::vector<uint32_t> row;
    for (int64_t a = 0; a < n; ++a) {
        ptr[a] = (int64_t)ent.size();
        if (deg[a + 1] == deg[a]) continue;
        // breadth-first search to depth 3: the first depth an atom is met at is its bond distance,
        // which is what isSeparatedByNOrMoreBonds(a, x, 3) reports (MMFF94.java:1106-1140)
        seen.assign(1, (int32_t)a);
        frontier.assign(1, (int32_t)a);
        depth[a] = 0;
        for (int d = 1; d <= 3; ++d) {
            next.clear();
            for (int32_t u : frontier)
                for (int64_t e = deg[u]; e < deg[u + 1]; ++e) {
                    int32_t v = adj[e];
                    if (depth[v] < 0) { depth[v] = (int8_t)d; seen.push_back(v); next.push_back(v); }
                }
            frontier.swap(next);
        }
        row.clear();
        for (int32_t v : seen) {
            if (depth[v] >= 1) row.push_back((uint32_t)v | (depth[v] == 3 ? kFlag14 : 0u));
            depth[v] = -1;
        }
        std::sort(row.begin(), row.end(), [](uint32_t x, uint32_t y) { return (x & ~kFlag14) < (y & ~kFlag14); });
        ent.insert(ent.end(), row.begin(), row.end());
    }
    ptr[n] = (int64_t)ent.size();
}

double cutoff_r2_threshold(double cutoff) {
    double t = cutoff * cutoff;
    while (std::sqrt(t) > cutoff) t = std::nextafter(t, 0.0);
    while (std::sqrt(std::nextafter(t, INFINITY)) <= cutoff) t = std::nextafter(t, INFINITY);
    return t;
}

void enumerate_terms(int64_t n, const int32_t* type, const uint8_t* linear, int64_t nb,
                     const int32_t* bi, const int32_t* bj, std::vector<int32_t>& angles,
                     std::vector<int32_t>& oops, std::vector<int32_t>& torsions) {
    std::vector<std::vector<int32_t>> adj(n);
    for (int64_t b = 0; b < nb; ++b) { adj[bi[b]].push_back(bj[b]); adj[bj[b]].push_back(bi[b]); }
    std::set<std::array<int32_t, 3>> sa;
    std::set<std::array<int32_t, 4>> so, st;
    angles.clear(); oops.clear(); torsions.clear();
    for (int32_t atom = 0; atom < n; ++atom) {
        const auto& nbr = adj[atom];
        if (nbr.size() < 2) continue;                       // MMFF94.java:172
        const int tj = type[atom];
        for (int32_t a2 : nbr)
            for (int32_t a3 : nbr) {
                if (a3 == a2) continue;
                const int ti = type[a2], tk = type[a3];
                const bool first = ti < tk || (ti == tk && a2 < a3);              // :180
                if (first && sa.insert({a2, atom, a3}).second) angles.insert(angles.end(), {a2, atom, a3});
                if (nbr.size() >= 3 && first)                                     // :184-193
                    for (int32_t a4 : nbr) {
                        if (a4 == a2 || a4 == a3) continue;
                        if (so.insert({a2, atom, a3, a4}).second) oops.insert(oops.end(), {a2, atom, a3, a4});
                    }
                if (!linear[atom] && !linear[a3] && adj[a3].size() >= 2)          // :196-210
                    for (int32_t a4 : adj[a3]) {
                        if (a4 == atom || a4 == a2) continue;
                        const int tl = type[a4];
                        if (tj < tk || (tj == tk && ti < tl) || (tj == tk && ti == tl && atom < a3))
                            if (st.insert({a2, atom, a3, a4}).second) torsions.insert(torsions.end(), {a2, atom, a3, a4});
                    }
            }
    }
}

std::string build_host_system(const jme_system_t* s, HostSystem& h) {
    if (!s) return "null system";
    if (s->n_atoms < 0 || s->n_atoms > 0x3fffffff) return "n_atoms out of range";
    const int64_t n = s->n_atoms;
    if (n > 0 && (!s->atom_type || !s->charge || !s->linear)) return "null per-atom array";
    if (s->n_vdw_types < 0 || (s->n_vdw_types > 0 && (!s->vdw_type || !s->vdw_alpha || !s->vdw_N || !s->vdw_A || !s->vdw_G || !s->vdw_DA)))
        return "null vdW table";
    if (s->n_bonds < 0 || s->n_angles < 0 || s->n_oops < 0 || s->n_torsions < 0) return "negative term count";
    if (s->n_bonds && (!s->bond_i || !s->bond_j || !s->bond_kb || !s->bond_r0)) return "null bond array";
    if (s->n_angles && (!s->angle_i || !s->angle_j || !s->angle_k || !s->angle_ka || !s->angle_theta0 || !s->stbn_kba_ijk || !s->stbn_kba_kji))
        return "null angle array";
    if (s->n_oops && (!s->oop_i || !s->oop_j || !s->oop_k || !s->oop_l || !s->oop_koop)) return "null out-of-plane array";
    if (s->n_torsions && (!s->tor_i || !s->tor_j || !s->tor_k || !s->tor_l || !s->tor_v1 || !s->tor_v2 || !s->tor_v3))
        return "null torsion array";
    auto bad = [n](int32_t a) { return a < 0 || a >= n; };

    h = HostSystem();
    h.n = n;
    // dense type indices in order of first appearance in the vdW table
    std::unordered_map<int, int> row_of_type;
    for (int t = 0; t < s->n_vdw_types; ++t) {
        const int ty = s->vdw_type[t];
        if (!((ty >= 1 && ty <= 82) || (ty >= 87 && ty <= 99))) return "vdw_type is not an MMFF94 numeric type";
        row_of_type.emplace(ty, t);
    }
    std::map<int, int> dense;   // type -> dense idx, only the types that occur
    h.type_idx.resize(n);
    for (int64_t a = 0; a < n; ++a) {
        const int ty = s->atom_type[a];
        if (!row_of_type.count(ty)) return "MMFF94 parameters need to be assigned first (atom type " + std::to_string(ty) + " has no vdW row)";
        dense.emplace(ty, 0);
    }
    for (auto& kv : dense) { kv.second = (int)h.type_of_idx.size(); h.type_of_idx.push_back(kv.first); }
    for (int64_t a = 0; a < n; ++a) h.type_idx[a] = dense[s->atom_type[a]];
    h.n_types = (int)h.type_of_idx.size();
    h.charge.assign(s->charge, s->charge + n);

    const int T = h.n_types;
    h.pair_table.resize((size_t)T * T);
    h.pair_rstar.resize((size_t)T * T);
    h.pair_eps.resize((size_t)T * T);
    for (int a = 0; a < T; ++a)
        for (int b = a; b < T; ++b) {
            const int ra = row_of_type[h.type_of_idx[a]], rb = row_of_type[h.type_of_idx[b]];
            double R, eps;
            vdw_pair_constants(s->vdw_alpha[ra], s->vdw_N[ra], s->vdw_A[ra], s->vdw_G[ra], s->vdw_DA[ra],
                               s->vdw_alpha[rb], s->vdw_N[rb], s->vdw_A[rb], s->vdw_G[rb], s->vdw_DA[rb], R, eps);
            PairConst pc;
            const double R7 = std::pow(R, 7);                 // Math.pow(RStarIJ, 7), Vdw :141
            pc.c2 = 0.07 * R;                                 // :140
            pc.c3 = 1.12 * R7;                                // :141
            pc.c4 = 0.12 * R7;                                // :141
            pc.k1 = 2.0 * (eps * std::pow(1.07 * R, 7));      // 2 eps (1.07 R*)^7: folded numerator of :140, doubled (pair_terms)
            h.pair_table[(size_t)a * T + b] = h.pair_table[(size_t)b * T + a] = pc;
            h.pair_rstar[(size_t)a * T + b] = h.pair_rstar[(size_t)b * T + a] = R;
            h.pair_eps[(size_t)a * T + b] = h.pair_eps[(size_t)b * T + a] = eps;
        }

    for (int64_t b = 0; b < s->n_bonds; ++b)
        if (bad(s->bond_i[b]) || bad(s->bond_j[b]) || s->bond_i[b] == s->bond_j[b]) return "bond atom index out of range";
    build_exclusions(n, s->n_bonds, s->bond_i, s->bond_j, h.excl_ptr, h.excl_ent);

    std::map<std::pair<int32_t, int32_t>, float> r0_of;
    h.bond_at.resize(2 * s->n_bonds);
    h.bond_par.resize(2 * s->n_bonds);
    for (int64_t b = 0; b < s->n_bonds; ++b) {
        h.bond_at[2 * b] = s->bond_i[b];
        h.bond_at[2 * b + 1] = s->bond_j[b];
        h.bond_par[2 * b] = s->bond_kb[b];
        h.bond_par[2 * b + 1] = s->bond_r0[b];
        r0_of[{s->bond_i[b], s->bond_j[b]}] = s->bond_r0[b];
        r0_of[{s->bond_j[b], s->bond_i[b]}] = s->bond_r0[b];
    }
    h.angle_at.resize(3 * s->n_angles);
    h.angle_par.resize(6 * s->n_angles);
    h.angle_linear.resize(s->n_angles);
    for (int64_t a = 0; a < s->n_angles; ++a) {
        const int32_t i = s->angle_i[a], j = s->angle_j[a], k = s->angle_k[a];
        if (bad(i) || bad(j) || bad(k)) return "angle atom index out of range";
        auto r1 = r0_of.find({i, j}), r2 = r0_of.find({j, k});
        if (r1 == r0_of.end() || r2 == r0_of.end()) return "the atoms must be bonded in the order i-j-k";   // StretchBend :104
        h.angle_at[3 * a] = i; h.angle_at[3 * a + 1] = j; h.angle_at[3 * a + 2] = k;
        double* p = &h.angle_par[6 * a];
        p[0] = s->angle_ka[a]; p[1] = s->angle_theta0[a]; p[2] = s->stbn_kba_ijk[a]; p[3] = s->stbn_kba_kji[a];
        p[4] = r1->second; p[5] = r2->second;
        h.angle_linear[a] = s->linear[j] ? 1 : 0;
    }
    h.oop_at.resize(4 * s->n_oops);
    h.oop_par.resize(s->n_oops);
    for (int64_t o = 0; o < s->n_oops; ++o) {
        const int32_t at[4] = {s->oop_i[o], s->oop_j[o], s->oop_k[o], s->oop_l[o]};
        for (int c = 0; c < 4; ++c) { if (bad(at[c])) return "out-of-plane atom index out of range"; h.oop_at[4 * o + c] = at[c]; }
        h.oop_par[o] = s->oop_koop[o];
    }
    h.tor_at.resize(4 * s->n_torsions);
    h.tor_par.resize(3 * s->n_torsions);
    for (int64_t t = 0; t < s->n_torsions; ++t) {
        int32_t i = s->tor_i[t], j = s->tor_j[t], k = s->tor_k[t], l = s->tor_l[t];
        if (bad(i) || bad(j) || bad(k) || bad(l)) return "torsion atom index out of range";
        const int ti = s->atom_type[i], tj = s->atom_type[j], tk = s->atom_type[k], tl = s->atom_type[l];
        if (tk < tj || (tj == tk && ti > tl) || (tj == tk && ti == tl && j > k)) {   // Torsion :117-124
            std::swap(j, k);
            std::swap(i, l);
        }
        h.tor_at[4 * t] = i; h.tor_at[4 * t + 1] = j; h.tor_at[4 * t + 2] = k; h.tor_at[4 * t + 3] = l;
        h.tor_par[3 * t] = s->tor_v1[t]; h.tor_par[3 * t + 1] = s->tor_v2[t]; h.tor_par[3 * t + 2] = s->tor_v3[t];
    }
    return "";
}

}  // namespace jme
