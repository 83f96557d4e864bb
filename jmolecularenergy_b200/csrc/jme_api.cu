// jme_api.cu - the C ABI of include/jme_b200.h: handle, marshalling to device SoA buffers, launch
// plans and the evaluation sequence.  No CPU fallback: every entry point that computes needs a CUDA
// device and fails with JME_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/jme_b200.h"
#include "jme_cells.cuh"
#include "jme_clusters.cuh"
#include "jme_peer.cuh"
#include "jme_delta.cuh"
#include "jme_host.hpp"
#include "jme_kernels.cuh"

using namespace jme;

namespace {

thread_local std::string g_create_error;

const double kUnitToJoule[3] = {4184.0, 1000.0, 1.0};   // ForceField.java:135-137

struct DeviceAllocs {
    std::vector<void*> ptrs;
    ~DeviceAllocs() { for (void* p : ptrs) cudaFree(p); }
};

}  // namespace

struct jme_handle {
    HostSystem hs;
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;

    double cutoff = 0, dielectric = 1;
    int unit = JME_KCAL_PER_MOL;
    int path = JME_PATH_AUTO;
    int rank = 0, world = 1;
    bool coords_set = false;
    bool plan_dirty = true;
    int active_path = JME_PATH_ALLPAIRS;

    DeviceAllocs mem;
    DeviceSystem ds{};
    BondedTerms bt{};
    double *d_x = nullptr, *d_y = nullptr, *d_z = nullptr, *d_q = nullptr;
    int* d_type = nullptr;
    PairConst* d_table = nullptr;
    double* d_xyz_stage = nullptr;

    // all-pairs plan
    std::vector<ApItem> items;
    ApItem* d_items = nullptr;
    int* d_tile_mask = nullptr;
    TileMask* d_masks = nullptr;
    int tiles_per_item = 1, n_chunks = 1, n_slots = 0;
    int item_begin = 0, item_end = 0;
    double* d_gpart = nullptr;
    size_t gpart_bytes = 0;
    double* d_epart = nullptr;
    unsigned long long* d_cpart = nullptr;
    int n_epart = 0;

    // bonded plan
    double* d_slots = nullptr;
    long long *d_slot_ptr = nullptr, *d_slot_idx = nullptr;
    double* d_bpart = nullptr;
    int n_bpart = 0;
    long long term_begin = 0, term_end = 0, slot_lo = 0, slot_hi = 0;

    // cell path
    CellPlan cells;

    double* d_out = nullptr;
    double* h_out = nullptr;   // pinned, mapped header: reduce_kernel writes the results straight into it
    double* h_out_dev = nullptr;   // device alias of h_out
    uint64_t last_pairs = 0, last_cand = 0;

    // single-coordinate energy differences (jme_delta.cuh)
    long long* d_dx_excl_ptr = nullptr;
    unsigned int* d_dx_excl_ent = nullptr;
    double* d_delta_part = nullptr;
    unsigned int* d_delta_ctr = nullptr;
    double* h_delta = nullptr;     // pinned, mapped: [0] total, [1..7] components, [8] sequence number
    double* h_delta_dev = nullptr;
    uint64_t delta_seq = 0;
    int incremental = 0;           // native minimizer loops: 0 full evaluations, 1 jme_energy_delta per trial, 2 the whole loop on the device

    // CUDA graphs of the all-pairs evaluation sequence {nonbonded | bonded} -> reduce (launch-bound small
    // systems): one per (gradient?, output buffer); rebuilt whenever anything baked into the kernel arguments changes
    struct EvalGraph { cudaGraphExec_t exec = nullptr; uint64_t epoch = 0; double* out = nullptr; cudaStream_t stream = nullptr; uint64_t kernels = 0; };
    EvalGraph graphs[2];
    uint64_t graph_epoch = 1;
    cudaStream_t side_stream = nullptr;
    cudaEvent_t fork_event = nullptr, join_event = nullptr;
    bool use_graphs = true;

    // profiling ring: 3 events per evaluation (before build, before pairs, after pairs)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    size_t prof_used = 0;     // evaluations recorded since the last read
    ~jme_handle() {
        for (cudaEvent_t e : prof_events) cudaEventDestroy(e);
        for (auto& g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
        if (fork_event) cudaEventDestroy(fork_event);
        if (join_event) cudaEventDestroy(join_event);
        if (side_stream) cudaStreamDestroy(side_stream);
    }
};

namespace {
constexpr int kListOverflow = 100;     // internal: finish_host saw an overflow flag of the cell path (+ 1: entries, + 2: clusters)

#define JME_CUDA(h, call)                                                                        \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                       \
            return JME_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

template <class T>
int dev_alloc(jme_handle* h, T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) {
        h->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? JME_ERR_NOMEM : JME_ERR_CUDA;
    }
    h->mem.ptrs.push_back(q);
    *p = static_cast<T*>(q);
    return JME_OK;
}

template <class T>
int dev_upload(jme_handle* h, T** p, const std::vector<T>& v) {
    int rc = dev_alloc(h, p, v.size());
    if (rc) return rc;
    if (!v.empty()) JME_CUDA(h, cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return JME_OK;
}

void dev_free(jme_handle* h, void* p) {
    if (!p) return;
    auto& v = h->mem.ptrs;
    auto it = std::find(v.begin(), v.end(), p);
    if (it != v.end()) v.erase(it);
    cudaFree(p);
}

int fail(jme_handle* h, int code, const std::string& msg) {
    h->err = msg;
    return code;
}

// ---- all-pairs plan -------------------------------------------------------------------------
// Circulant decomposition over SUPERBLOCKS (kNI 32-atom blocks): row S owns the superblock pairs
// (S, (S+d) mod nS) for d = 0 .. H(S); every unordered superblock pair appears exactly once and every row has
// the same length (+-1), so rows - and ranks owning contiguous item ranges - are balanced.  Row S visits the
// j-tiles q = d*kNI + b (block ((S+d) mod nS)*kNI + b).
inline int half_ring(int nS, int S) {
    const int full = (nS - 1) / 2;
    return full + ((nS % 2 == 0 && S < nS / 2) ? 1 : 0);
}

int build_allpairs_plan(jme_handle* h) {
    const HostSystem& hs = h->hs;
    const int nS = h->ds.n_pad / kSuper;
    const int n = (int)hs.n;
    long long total_tiles = 0;
    for (int S = 0; S < nS; ++S) total_tiles += (long long)(half_ring(nS, S) + 1) * kNI;
    if (total_tiles > (1ll << 30)) return fail(h, JME_ERR_UNSUPPORTED, "system too large for the all-pairs path; set a cutoff to use the cell path");
    // enough items to fill 148 SMs x 8 resident warps a few times over, but >= 1 tile each
    long long tpi = total_tiles / (148ll * 8 * 4);
    tpi = std::max(1ll, std::min(32ll, tpi));
    h->tiles_per_item = (int)tpi;
    const int max_row = ((nS - 1) / 2 + 1 + 1) * kNI;
    h->n_chunks = (max_row + (int)tpi - 1) / (int)tpi;
    h->n_slots = nS + h->n_chunks;
    const size_t gpart_bytes = (size_t)h->n_slots * 3 * h->ds.n_pad * sizeof(double);
    if (gpart_bytes > (size_t)96 << 30) return fail(h, JME_ERR_UNSUPPORTED, "all-pairs partial-gradient buffer would exceed 96 GiB; set a cutoff to use the cell path");

    h->items.clear();
    std::vector<int> row_first_item(nS + 1, 0);
    int tile_ofs = 0;
    for (int S = 0; S < nS; ++S) {
        row_first_item[S] = (int)h->items.size();
        const int len = (half_ring(nS, S) + 1) * kNI;
        for (int q0 = 0; q0 < len; q0 += (int)tpi) {
            ApItem it{S, q0, std::min((int)tpi, len - q0), tile_ofs};
            tile_ofs += it.nq;
            h->items.push_back(it);
        }
    }
    row_first_item[nS] = (int)h->items.size();

    // tile masks
    std::vector<int> tile_mask((size_t)total_tiles, -1);
    std::vector<TileMask> masks;
    auto tile_index = [&](int S, int q) {
        const ApItem& it = h->items[row_first_item[S] + q / (int)tpi];
        return it.tile_ofs + (q - it.q0);
    };
    auto mask_of = [&](int S, int q) -> TileMask& {
        int& m = tile_mask[tile_index(S, q)];
        if (m < 0) {
            m = (int)masks.size();
            masks.emplace_back();
            std::memset(&masks.back(), 0, sizeof(TileMask));
        }
        return masks[m];
    };
    // which row owns the superblock pair {A, B}, and at which offset
    auto owner = [&](int A, int B, int& S, int& d) {
        const int d1 = (B - A + nS) % nS, d2 = nS - d1;
        if (d1 < d2 || (d1 == d2 && A < B)) { S = A; d = d1; } else { S = B; d = d2; }
    };
    for (int64_t i = 0; i < n; ++i)
        for (int64_t e = hs.excl_ptr[i]; e < hs.excl_ptr[i + 1]; ++e) {
            const uint32_t v = hs.excl_ent[e];
            const int64_t j = v & ~kFlag14;
            if (j <= i) continue;
            const bool is14 = (v & kFlag14) != 0;
            const int A = (int)(i / kTile), B = (int)(j / kTile), li = (int)(i % kTile), lj = (int)(j % kTile);
            const int SA = A / kNI, SB = B / kNI;
            int S, q, chain, row, bit;
            if (SA == SB) {
                // inside one superblock the pair is met by chain A%kNI (the lower block) on tile B%kNI;
                // on the diagonal tile (A == B) the kernel keeps j > i, i.e. lane li meets bit lj
                S = SA; q = B % kNI; chain = A % kNI; row = li; bit = lj;
            } else {
                int d;
                owner(SA, SB, S, d);
                if (S == SA) { q = d * kNI + B % kNI; chain = A % kNI; row = li; bit = lj; }
                else { q = d * kNI + A % kNI; chain = B % kNI; row = lj; bit = li; }
            }
            TileMask& m = mask_of(S, q);
            if (is14) m.f14[chain][row] |= 1u << bit; else m.excl[chain][row] |= 1u << bit;
        }
    if (h->ds.n_pad > n) {
        // padding atoms (index >= n) never interact
        auto valid_in_block = [&](int B) { return std::max(0, std::min(kTile, n - B * kTile)); };
        const int first_pad_block = n / kTile;
        for (int S = 0; S < nS; ++S) {
            const int len = (half_ring(nS, S) + 1) * kNI;
            const bool i_has_pad = (S * kNI + kNI - 1) >= first_pad_block;
            for (int q = 0; q < len; ++q) {
                const int J = ((S + q / kNI) % nS) * kNI + q % kNI;
                if (!i_has_pad && J < first_pad_block) continue;
                TileMask& m = mask_of(S, q);
                const int vj = valid_in_block(J);
                const uint32_t pad_bits = vj >= 32 ? 0u : (0xffffffffu << vj);
                for (int a = 0; a < kNI; ++a) {
                    const int vi = valid_in_block(S * kNI + a);
                    for (int r = 0; r < 32; ++r) {
                        if (r >= vi) m.excl[a][r] = 0xffffffffu;
                        m.excl[a][r] |= pad_bits;
                    }
                }
            }
        }
    }

    dev_free(h, h->d_items); dev_free(h, h->d_tile_mask); dev_free(h, h->d_masks);
    dev_free(h, h->d_gpart); dev_free(h, h->d_epart); dev_free(h, h->d_cpart);
    int rc;
    if ((rc = dev_upload(h, &h->d_items, h->items))) return rc;
    if ((rc = dev_upload(h, &h->d_tile_mask, tile_mask))) return rc;
    if ((rc = dev_upload(h, &h->d_masks, masks))) return rc;
    if ((rc = dev_alloc(h, &h->d_gpart, gpart_bytes / sizeof(double)))) return rc;
    h->gpart_bytes = gpart_bytes;
    JME_CUDA(h, cudaMemsetAsync(h->d_gpart, 0, gpart_bytes, h->stream));
    h->n_epart = (int)h->items.size();
    if ((rc = dev_alloc(h, &h->d_epart, 2 * (size_t)h->n_epart))) return rc;
    if ((rc = dev_alloc(h, &h->d_cpart, 2 * (size_t)h->n_epart))) return rc;
    JME_CUDA(h, cudaMemsetAsync(h->d_epart, 0, 2 * (size_t)h->n_epart * sizeof(double), h->stream));
    JME_CUDA(h, cudaMemsetAsync(h->d_cpart, 0, 2 * (size_t)h->n_epart * sizeof(unsigned long long), h->stream));
    const long long ni = (long long)h->items.size();
    h->item_begin = (int)(ni * h->rank / h->world);
    h->item_end = (int)(ni * (h->rank + 1) / h->world);
    return JME_OK;
}

// ---- bonded plan ----------------------------------------------------------------------------
int build_bonded_plan(jme_handle* h) {
    const HostSystem& hs = h->hs;
    const long long nb = h->bt.n_bond, na = h->bt.n_angle, no = h->bt.n_oop, nt = h->bt.n_tor;
    const long long terms = nb + na + no + nt;
    auto slot_of_term = [&](long long t) -> long long {
        if (t <= nb) return 2 * t;
        t -= nb;
        if (t <= na) return 2 * nb + 3 * t;
        t -= na;
        if (t <= no) return 2 * nb + 3 * na + 4 * t;
        t -= no;
        return 2 * nb + 3 * na + 4 * no + 4 * t;
    };
    h->term_begin = terms * h->rank / h->world;
    h->term_end = terms * (h->rank + 1) / h->world;
    h->slot_lo = slot_of_term(h->term_begin);
    h->slot_hi = slot_of_term(h->term_end);
    h->n_bpart = (int)((h->term_end - h->term_begin + kBondedThreads - 1) / kBondedThreads);
    if (h->d_slots == nullptr) {
        const long long n_slots = slot_of_term(terms);
        // CSR atom -> slots, ascending slot id
        std::vector<long long> ptr(hs.n + 1, 0), idx((size_t)n_slots);
        auto for_each_slot = [&](auto&& fn) {
            for (long long b = 0; b < nb; ++b) for (int c = 0; c < 2; ++c) fn(hs.bond_at[2 * b + c], 2 * b + c);
            for (long long a = 0; a < na; ++a) for (int c = 0; c < 3; ++c) fn(hs.angle_at[3 * a + c], 2 * nb + 3 * a + c);
            for (long long o = 0; o < no; ++o) for (int c = 0; c < 4; ++c) fn(hs.oop_at[4 * o + c], 2 * nb + 3 * na + 4 * o + c);
            for (long long t = 0; t < nt; ++t) for (int c = 0; c < 4; ++c) fn(hs.tor_at[4 * t + c], 2 * nb + 3 * na + 4 * no + 4 * t + c);
        };
        for_each_slot([&](int atom, long long) { ++ptr[atom + 1]; });
        for (int64_t a = 0; a < hs.n; ++a) ptr[a + 1] += ptr[a];
        std::vector<long long> fill(ptr.begin(), ptr.end() - 1);
        for_each_slot([&](int atom, long long s) { idx[fill[atom]++] = s; });
        int rc;
        if ((rc = dev_alloc(h, &h->d_slots, 3 * (size_t)n_slots))) return rc;
        if ((rc = dev_upload(h, &h->d_slot_ptr, ptr))) return rc;
        if ((rc = dev_upload(h, &h->d_slot_idx, idx))) return rc;
    }
    dev_free(h, h->d_bpart);
    return dev_alloc(h, &h->d_bpart, 5 * (size_t)std::max(1, h->n_bpart));
}

// atoms from which the cutoff path evaluates cluster lists with Newton's third law instead of per-atom lists (below,
// the sweep of a few hundred segments leaves most of the chip idle; measured crossover)
constexpr int64_t kClusterPathMinAtoms = 90000;

int choose_path(const jme_handle* h) {
    if (h->path == JME_PATH_ALLPAIRS) return JME_PATH_ALLPAIRS;
    if (h->path == JME_PATH_CELLS || h->path == JME_PATH_CELLS_LISTS || h->path == JME_PATH_CELLS_CLUSTERS) return JME_PATH_CELLS;
    // AUTO: the cell path pays off once the cutoff sphere holds a small fraction of the atoms
    if (h->cutoff > 0 && h->hs.n >= 6000) return JME_PATH_CELLS;
    return JME_PATH_ALLPAIRS;
}

int ensure_plan(jme_handle* h) {
    if (!h->plan_dirty) return JME_OK;
    int rc;
    const int p = choose_path(h);
    if (p == JME_PATH_CELLS && !(h->cutoff > 0)) return fail(h, JME_ERR_INVALID, "the cell path needs a cutoff > 0");
    if (p == JME_PATH_ALLPAIRS) {
        if ((rc = build_allpairs_plan(h))) return rc;
    } else {
        // the two evaluation schemes of the cutoff path: cluster lists (4 atoms per cluster) for large systems,
        // per-atom lists below; JME_PATH_CELLS_LISTS / _CLUSTERS force one, JME_CLUSTER (0, 2, 4) is a development knob
        int cluster = h->hs.n / std::max(1, h->world) >= kClusterPathMinAtoms ? 4 : 0;       // atoms per rank
        if (h->path == JME_PATH_CELLS_LISTS) cluster = 0;
        if (h->path == JME_PATH_CELLS_CLUSTERS) cluster = 4;
        if (const char* e = std::getenv("JME_CLUSTER")) {
            const int v = std::atoi(e);
            if (v == 0 || v == 2 || v == 4) cluster = v;
        }
        if (h->cells.built && h->cells.cluster != cluster) {          // the plan changes its kind: release the old one
            for (void* q : h->cells.all_allocs) dev_free(h, q);
            h->cells = CellPlan();
        }
        rc = build_cell_plan(h->cells, h->hs, h->ds, h->cutoff, h->rank, h->world, cluster, h->err);
        for (void* q : h->cells.take_new_allocs()) h->mem.ptrs.push_back(q);
        if (rc) return rc;
        if (cluster != 0 && (rc = cl_configure(h->cells, h->ds, h->err))) return rc;
    }
    if ((rc = build_bonded_plan(h))) return rc;
    h->active_path = p;
    h->ds.r2_max = h->cutoff > 0 ? cutoff_r2_threshold(h->cutoff) : 0.0;
    h->ds.ele_scale = kEleC / h->dielectric;
    h->plan_dirty = false;
    ++h->graph_epoch;
    return JME_OK;
}

constexpr size_t kProfRing = 4096;

// returns the 3 events of the next profiling slot, or nullptr when profiling is off / the ring is full
cudaEvent_t* prof_slot(jme_handle* h) {
    if (!h->profiling || h->prof_used >= kProfRing) return nullptr;
    while (h->prof_events.size() < 3 * (h->prof_used + 1)) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        h->prof_events.push_back(e);
    }
    return &h->prof_events[3 * h->prof_used];
}

constexpr size_t kTableSmemLimit = 48 * 1024;       // larger type-pair tables stay in global memory

size_t ap_smem_bytes(const DeviceSystem& ds, bool tsm) {
    return (tsm ? (size_t)ds.n_types * ds.n_types * sizeof(PairConst) : 0) + kApWarps * (4 * 32 + 16) * sizeof(double) +
           kApWarps * (kNI * 3 + 3) * 32 * sizeof(double);
}

template <bool GRAD, bool CUTOFF, int KSPLIT, bool TSM>
int launch_ap_kt(jme_handle* h) {
    const int n_items = h->item_end - h->item_begin;
    if (n_items <= 0) return JME_OK;
    const size_t smem = ap_smem_bytes(h->ds, TSM);
    auto kern = ap_kernel<GRAD, CUTOFF, KSPLIT, TSM>;
    if (smem > 48 * 1024) JME_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = KSPLIT == 1 ? (n_items + kApWarps - 1) / kApWarps : n_items;
    kern<<<grid, kApThreads, smem, h->stream>>>(h->ds, h->d_items, h->item_begin, h->item_end, h->d_tile_mask, h->d_masks,
                                                h->ds.n_pad / kSuper, h->d_gpart, h->tiles_per_item, h->d_epart, h->d_cpart);
    ++h->launches;
    JME_CUDA(h, cudaGetLastError());
    return JME_OK;
}

template <bool GRAD, bool CUTOFF, int KSPLIT>
int launch_ap_k(jme_handle* h) {
    const bool tsm = (size_t)h->ds.n_types * h->ds.n_types * sizeof(PairConst) <= kTableSmemLimit;
    return tsm ? launch_ap_kt<GRAD, CUTOFF, KSPLIT, true>(h) : launch_ap_kt<GRAD, CUTOFF, KSPLIT, false>(h);
}

// Small systems do not have enough tiles to fill the chip with one warp per item: let the four warps of a CTA
// share an item (a quarter of each tile's steps each).
template <bool GRAD, bool CUTOFF>
int launch_ap(jme_handle* h) {
    const long long n_items = h->item_end - h->item_begin;
    return n_items < 148ll * 8 * 3 ? launch_ap_k<GRAD, CUTOFF, kApWarps>(h) : launch_ap_k<GRAD, CUTOFF, 1>(h);
}

// The launches of one evaluation; results land in h->d_out (or d_out_user).  With `fork` the bonded kernel runs
// on a side stream next to the nonbonded kernel (used while capturing the evaluation into a CUDA graph).
int enqueue_launches(jme_handle* h, bool grad, double* d_out_user, bool fork) {
    int rc;
    const bool cut = h->cutoff > 0;
    ReduceArgs A{};
    A.n = h->ds.n; A.n_pad = h->ds.n_pad;
    const long long n_terms = h->term_end - h->term_begin;
    auto launch_bonded = [&](cudaStream_t st) -> int {
        if (n_terms <= 0) return JME_OK;
        if (grad) bonded_kernel<true><<<h->n_bpart, kBondedThreads, 0, st>>>(h->ds, h->bt, h->term_begin, h->term_end, h->d_slots, h->d_bpart);
        else bonded_kernel<false><<<h->n_bpart, kBondedThreads, 0, st>>>(h->ds, h->bt, h->term_begin, h->term_end, h->d_slots, h->d_bpart);
        ++h->launches;
        JME_CUDA(h, cudaGetLastError());
        return JME_OK;
    };
    if (fork && n_terms > 0) {
        JME_CUDA(h, cudaEventRecord(h->fork_event, h->stream));
        JME_CUDA(h, cudaStreamWaitEvent(h->side_stream, h->fork_event, 0));
        if ((rc = launch_bonded(h->side_stream))) return rc;
        JME_CUDA(h, cudaEventRecord(h->join_event, h->side_stream));
    }
    cudaEvent_t* pe = prof_slot(h);
    if (pe) cudaEventRecord(pe[0], h->stream);
    if (h->active_path == JME_PATH_ALLPAIRS) {
        if (pe) cudaEventRecord(pe[1], h->stream);
        if (grad) rc = cut ? launch_ap<true, true>(h) : launch_ap<true, false>(h);
        else rc = cut ? launch_ap<false, true>(h) : launch_ap<false, false>(h);
        if (rc) return rc;
        A.gpart = h->d_gpart; A.n_slots = h->n_slots; A.direct = nullptr;
        A.epart = h->d_epart; A.cpart = h->d_cpart; A.n_epart = h->n_epart;
    } else if (h->cells.cluster != 0) {
        if ((rc = run_cell_sort(h->cells, h->ds, h->stream, h->launches, h->err))) return rc;
        if ((rc = cl_run_lists(h->cells, h->ds, h->stream, h->launches, h->err))) return rc;
        if (pe) cudaEventRecord(pe[1], h->stream);
        if ((rc = cl_run_pairs(h->cells, h->ds, grad, h->stream, h->launches, h->err))) return rc;
        A.gpart = nullptr; A.n_slots = 0; A.direct = nullptr;
        A.overflow = h->cells.arr.flags;
        A.epart = h->cells.d_epart; A.cpart = h->cells.d_cpart; A.n_epart = 0;      // shard range taken on the device
    } else {
        if ((rc = run_cell_sort(h->cells, h->ds, h->stream, h->launches, h->err))) return rc;
        if ((rc = run_cell_lists(h->cells, h->ds, h->stream, h->launches, h->err))) return rc;
        if (pe) cudaEventRecord(pe[1], h->stream);
        if ((rc = run_cell_path(h->cells, h->ds, grad, h->stream, h->launches, h->err))) return rc;
        A.gpart = nullptr; A.n_slots = 0; A.direct = h->cells.d_grad;
        A.overflow = h->cells.d_overflow;
        A.epart = h->cells.d_epart; A.cpart = h->cells.d_cpart; A.n_epart = h->cells.n_epart;
    }
    if (pe) { cudaEventRecord(pe[2], h->stream); ++h->prof_used; }
    if (fork && n_terms > 0) JME_CUDA(h, cudaStreamWaitEvent(h->stream, h->join_event, 0));
    else if ((rc = launch_bonded(h->stream))) return rc;
    A.slots = h->d_slots; A.slot_ptr = h->d_slot_ptr; A.slot_idx = h->d_slot_idx;
    A.slot_lo = h->slot_lo; A.slot_hi = h->slot_hi;
    A.bpart = h->d_bpart; A.n_bpart = n_terms > 0 ? h->n_bpart : 0;
    A.unit_num = h->unit == JME_KCAL_PER_MOL ? 1.0 : kUnitToJoule[JME_KCAL_PER_MOL];
    A.unit_den = h->unit == JME_KCAL_PER_MOL ? 1.0 : kUnitToJoule[h->unit];
    A.want_grad = grad ? 1 : 0;
    A.out = d_out_user ? d_out_user : h->d_out;
    A.out_host = d_out_user ? nullptr : h->h_out_dev;
    if (h->active_path == JME_PATH_CELLS && h->cells.cluster != 0)
        return cl_run_reduce(h->cells, h->ds, A, h->stream, h->launches, h->err);
    const int apb = A.n_slots == 0 ? kReduceWarps * 32 : 32;      // atoms per block, see reduce_kernel
    const int grid = grad ? (h->ds.n + apb - 1) / apb + 1 : 1;
    reduce_kernel<<<grid, kReduceWarps * 32, 0, h->stream>>>(A);
    ++h->launches;
    JME_CUDA(h, cudaGetLastError());
    return JME_OK;
}

// Enqueue one evaluation.  All-pairs path without profiling: replay a captured CUDA graph.
int enqueue_eval(jme_handle* h, bool grad, double* d_out_user) {
    int rc;
    if (!h->coords_set) return fail(h, JME_ERR_NOT_ASSIGNED, "MMFF94 parameters need to be assigned first (no coordinates set)");
    if ((rc = ensure_plan(h))) return rc;
    // (the legacy default stream cannot be captured)
    if (h->ds.n == 0 || !h->use_graphs || h->profiling || h->stream == nullptr)
        return enqueue_launches(h, grad, d_out_user, false);
    // the captured sequence of the cell path always contains the sort and the list build
    if (h->active_path == JME_PATH_CELLS) h->cells.coords_changed();
    jme_handle::EvalGraph& g = h->graphs[grad ? 1 : 0];
    double* out = d_out_user ? d_out_user : h->d_out;
    if (!g.exec || g.epoch != h->graph_epoch || g.out != out || g.stream != h->stream) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        if (!h->side_stream) {
            JME_CUDA(h, cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
            JME_CUDA(h, cudaEventCreateWithFlags(&h->fork_event, cudaEventDisableTiming));
            JME_CUDA(h, cudaEventCreateWithFlags(&h->join_event, cudaEventDisableTiming));
        }
        const uint64_t launches_before = h->launches;
        cudaGraph_t graph = nullptr;
        JME_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        rc = enqueue_launches(h, grad, d_out_user, true);
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        g.kernels = h->launches - launches_before;
        h->launches = launches_before;
        if (rc || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            h->use_graphs = false;                 // fall back to plain launches for this handle
            h->cells.coords_changed();             // the sort was only captured, never executed
            return rc ? rc : enqueue_launches(h, grad, d_out_user, false);
        }
        ce = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) {
            g.exec = nullptr;
            cudaGetLastError();
            h->use_graphs = false;
            h->cells.coords_changed();
            return enqueue_launches(h, grad, d_out_user, false);
        }
        g.epoch = h->graph_epoch; g.out = out; g.stream = h->stream;
    }
    JME_CUDA(h, cudaGraphLaunch(g.exec, h->stream));
    h->launches += g.kernels;                                            // kernels inside the graph
    return JME_OK;
}

int finish_host(jme_handle* h, bool grad, double* total, double* comps, double* grad_xyz) {
    // the 10-double header arrives through mapped pinned memory (written by reduce_kernel); only the gradient
    // needs a copy
    if (grad && grad_xyz && h->hs.n > 0)
        JME_CUDA(h, cudaMemcpyAsync(grad_xyz, h->d_out + kOutHeader, 3 * (size_t)h->hs.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    JME_CUDA(h, cudaStreamSynchronize(h->stream));
    if (total) *total = h->h_out[0];
    if (comps) for (int c = 0; c < JME_N_COMPONENTS; ++c) comps[c] = h->h_out[1 + c];
    h->last_pairs = (uint64_t)h->h_out[8];
    h->last_cand = (uint64_t)h->h_out[9];
    if (h->active_path == JME_PATH_CELLS && h->h_out[kOutHeader] != 0.0) {
        const int f = (int)h->h_out[kOutHeader];
        if (f & 3) return kListOverflow + (f & 3);            // which storage ran out travels with the code
        // f & 4: a non-finite coordinate - the total is NaN, which is the answer
    }
    return JME_OK;
}

// after an overflow: survivor lists with twice the capacity per atom (the old ones are released), flag cleared
// replace a plan allocation (the plan's own bookkeeping follows)
template <class T>
int realloc_plan(jme_handle* h, T** p, size_t count) {
    auto& all = h->cells.all_allocs;
    all.erase(std::remove(all.begin(), all.end(), static_cast<void*>(*p)), all.end());
    dev_free(h, *p);
    *p = nullptr;
    int rc = dev_alloc(h, p, count);
    if (!rc) all.push_back(*p);
    return rc;
}

// after an overflow: `what` bit 0 = a list was truncated (twice the entries per list), bit 1 = the sort produced more
// clusters than the lists hold (twice the clusters); the old storage is released, the flags are cleared
int grow_cell_lists(jme_handle* h, int what = 1) {
    CellPlan& p = h->cells;
    int cap = p.list_cap, ccap = p.cluster_cap;
    if (what & 1) cap *= 2;
    // (cluster lists: a list is at most kClTripTab trips of the FP64 kernel)
    if (cap > (p.cluster != 0 ? 4096 : (1 << 16))) { h->err = "survivor-list capacity limit"; return JME_ERR_NOMEM; }
    int rc;
    if (p.cluster != 0) {
        if (what & 2) ccap = std::min(std::max(1, h->ds.n), std::max(ccap + 1, ccap * 2));
        rc = realloc_plan(h, &p.d_clist, (size_t)ccap * cap);
        if (!rc && ccap != p.cluster_cap) rc = realloc_plan(h, &p.d_count, (size_t)ccap);
        if (rc) { p.built = false; h->plan_dirty = true; return rc; }
        p.cluster_cap = ccap;
    } else {
        rc = realloc_plan(h, &p.d_list, (size_t)std::max(1, h->ds.n) * cap);
        if (rc) { p.built = false; h->plan_dirty = true; return rc; }
    }
    p.list_cap = cap;
    JME_CUDA(h, cudaMemset(p.arr.flags, 0, 4 * sizeof(int)));
    return JME_OK;
}

// Host-API evaluation on the cell path: when a survivor list overflowed, double the list capacity and evaluate
// again (dense systems / long cutoffs); gives up when the lists would not fit in memory.
int eval_host(jme_handle* h, bool grad, double* total, double* comps, double* grad_xyz) {
    for (;;) {
        int rc = enqueue_eval(h, grad, nullptr);
        if (rc) return rc;
        rc = finish_host(h, grad, total, comps, grad_xyz);
        if (rc <= kListOverflow) return rc;
        const int old_cap = h->cells.list_cap;
        if (int g = grow_cell_lists(h, rc - kListOverflow)) {
            if (g == JME_ERR_NOMEM) h->err = "a cutoff sphere holds more atoms than the survivor-list capacity (" +
                std::to_string(old_cap) + ") and larger lists do not fit in device memory: " + h->err;
            return g;
        }
        ++h->graph_epoch;                  // the list pointers / capacity are baked into the captured graph
    }
}

}  // namespace

// =============================================================================================
extern "C" {

int jme_abi_version(void) { return JME_ABI_VERSION; }
const char* jme_last_create_error(void) { return g_create_error.c_str(); }
const char* jme_last_error(const jme_handle_t* h) { return h ? h->err.c_str() : "null handle"; }
uint64_t jme_launch_count(const jme_handle_t* h) { return h ? h->launches : 0; }

int jme_create(const jme_system_t* sys, jme_handle_t** out) {
    if (!out) { g_create_error = "null output pointer"; return JME_ERR_INVALID; }
    *out = nullptr;
    jme_handle* h = new (std::nothrow) jme_handle();
    if (!h) { g_create_error = "out of host memory"; return JME_ERR_NOMEM; }
    auto bail = [&](int code) { g_create_error = h->err; delete h; return code; };
    std::string e = build_host_system(sys, h->hs);
    if (!e.empty()) { h->err = e; return bail(JME_ERR_INVALID); }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        h->err = std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0") +
                 " (libjme_b200 has no CPU fallback)";
        return bail(JME_ERR_CUDA);
    }
    if (cudaGetDevice(&h->device) != cudaSuccess) { h->err = "cudaGetDevice failed"; return bail(JME_ERR_CUDA); }
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { h->err = "cudaStreamCreate failed"; return bail(JME_ERR_CUDA); }
    h->stream = h->own_stream;

    const HostSystem& hs = h->hs;
    const int n = (int)hs.n;
    const int n_pad = std::max(kSuper, (n + kSuper - 1) / kSuper * kSuper);
    int rc;
    auto alloc_zero = [&](double** p, size_t cnt) -> int {
        int r = dev_alloc(h, p, cnt);
        if (r) return r;
        return cudaMemset(*p, 0, cnt * sizeof(double)) == cudaSuccess ? (int)JME_OK : (int)JME_ERR_CUDA;
    };
    if ((rc = alloc_zero(&h->d_x, n_pad)) || (rc = alloc_zero(&h->d_y, n_pad)) || (rc = alloc_zero(&h->d_z, n_pad)) ||
        (rc = alloc_zero(&h->d_q, n_pad)))
        return bail(rc);
    if ((rc = dev_alloc(h, &h->d_type, n_pad))) return bail(rc);
    if (cudaMemset(h->d_type, 0, n_pad * sizeof(int)) != cudaSuccess) { h->err = "cudaMemset failed"; return bail(JME_ERR_CUDA); }
    if (n > 0) {
        if (cudaMemcpy(h->d_q, hs.charge.data(), n * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(h->d_type, hs.type_idx.data(), n * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
            h->err = "upload of per-atom data failed";
            return bail(JME_ERR_CUDA);
        }
    }
    // a table with at least one entry even for an empty system
    std::vector<PairConst> table = hs.pair_table;
    if (table.empty()) table.push_back(PairConst{0, 1, 0, 1});
    if ((rc = dev_upload(h, &h->d_table, table))) return bail(rc);
    if ((rc = dev_alloc(h, &h->d_xyz_stage, 3 * (size_t)std::max(1, n)))) return bail(rc);
    if ((rc = dev_alloc(h, &h->d_out, kOutHeader + 3 * (size_t)std::max(1, n)))) return bail(rc);
    if (cudaHostAlloc(&h->h_out, (kOutHeader + 1) * sizeof(double), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&h->h_out_dev, h->h_out, 0) != cudaSuccess) { h->err = "cudaHostAlloc (mapped) failed"; return bail(JME_ERR_CUDA); }

    h->ds.n = n; h->ds.n_pad = n_pad; h->ds.n_blocks = n_pad / kTile; h->ds.n_types = std::max(1, hs.n_types);
    h->ds.x = h->d_x; h->ds.y = h->d_y; h->ds.z = h->d_z; h->ds.q = h->d_q; h->ds.type = h->d_type; h->ds.table = h->d_table;
    h->ds.r2_max = 0; h->ds.ele_scale = kEleC;

    // bonded arrays
    int* p_i; double* p_d; unsigned char* p_u;
    h->bt.n_bond = (int)(hs.bond_par.size() / 2); h->bt.n_angle = (int)hs.angle_linear.size();
    h->bt.n_oop = (int)hs.oop_par.size(); h->bt.n_tor = (int)(hs.tor_par.size() / 3);
    if ((rc = dev_upload(h, &p_i, hs.bond_at))) return bail(rc); h->bt.bond_at = p_i;
    if ((rc = dev_upload(h, &p_d, hs.bond_par))) return bail(rc); h->bt.bond_par = p_d;
    if ((rc = dev_upload(h, &p_i, hs.angle_at))) return bail(rc); h->bt.angle_at = p_i;
    if ((rc = dev_upload(h, &p_d, hs.angle_par))) return bail(rc); h->bt.angle_par = p_d;
    if ((rc = dev_upload(h, &p_u, hs.angle_linear))) return bail(rc); h->bt.angle_linear = p_u;
    if ((rc = dev_upload(h, &p_i, hs.oop_at))) return bail(rc); h->bt.oop_at = p_i;
    if ((rc = dev_upload(h, &p_d, hs.oop_par))) return bail(rc); h->bt.oop_par = p_d;
    if ((rc = dev_upload(h, &p_i, hs.tor_at))) return bail(rc); h->bt.tor_at = p_i;
    if ((rc = dev_upload(h, &p_d, hs.tor_par))) return bail(rc); h->bt.tor_par = p_d;
    *out = h;
    return JME_OK;
}

void jme_destroy(jme_handle_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_delta) cudaFreeHost(h->h_delta);
    delete h;
}

int jme_set_options(jme_handle_t* h, double cutoff, double dielectric, int energy_unit) {
    if (!h) return JME_ERR_INVALID;
    if (!(dielectric > 0) || std::isnan(cutoff) || std::isinf(cutoff)) return fail(h, JME_ERR_INVALID, "dielectric must be > 0 and cutoff finite");
    if (energy_unit < 0 || energy_unit > 2) return fail(h, JME_ERR_INVALID, "unknown energy unit");
    if (cutoff < 0) cutoff = 0;     // ForceField.java:175: only cutoff > 0 limits the interaction
    if (cutoff != h->cutoff) h->plan_dirty = true;
    h->cutoff = cutoff;
    h->dielectric = dielectric;
    h->unit = energy_unit;
    h->ds.r2_max = cutoff > 0 ? cutoff_r2_threshold(cutoff) : 0.0;
    h->ds.ele_scale = kEleC / dielectric;
    ++h->graph_epoch;          // cutoff threshold, dielectric scale and unit factors are kernel arguments
    return JME_OK;
}

int jme_set_path(jme_handle_t* h, int path) {
    if (!h || path < 0 || path > JME_PATH_CELLS_CLUSTERS) return JME_ERR_INVALID;
    if (path != h->path) h->plan_dirty = true;
    h->path = path;
    return JME_OK;
}

int jme_active_path(const jme_handle_t* h) {
    if (!h) return JME_ERR_INVALID;
    if (h->plan_dirty) return JME_PATH_AUTO;
    if (h->active_path == JME_PATH_ALLPAIRS) return JME_PATH_ALLPAIRS;
    return h->cells.cluster != 0 ? JME_PATH_CELLS_CLUSTERS : JME_PATH_CELLS_LISTS;
}

int jme_set_shard(jme_handle_t* h, int rank, int world) {
    if (!h) return JME_ERR_INVALID;
    if (world < 1 || rank < 0 || rank >= world) return fail(h, JME_ERR_INVALID, "need 0 <= rank < world");
    if (rank != h->rank || world != h->world) h->plan_dirty = true;
    h->rank = rank;
    h->world = world;
    return JME_OK;
}

int jme_set_stream(jme_handle_t* h, void* stream) {
    if (!h) return JME_ERR_INVALID;
    // NULL is a real stream (the legacy default stream, which is what torch's default stream is)
    h->stream = (stream == JME_STREAM_OWN) ? h->own_stream : static_cast<cudaStream_t>(stream);
    return JME_OK;
}

static int coords_from_device_aos(jme_handle* h, const double* d_xyz) {
    const int n = h->ds.n;
    if (n > 0) {
        aos_to_soa_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(d_xyz, n, h->d_x, h->d_y, h->d_z);
        ++h->launches;
        JME_CUDA(h, cudaGetLastError());
    }
    h->coords_set = true;
    h->cells.coords_changed();
    return JME_OK;
}

int jme_set_coords(jme_handle_t* h, const double* xyz, int64_t n_atoms) {
    if (!h) return JME_ERR_INVALID;
    if (n_atoms != h->hs.n) return fail(h, JME_ERR_INVALID, "atom count differs from the parameterised system");
    if (n_atoms > 0 && !xyz) return fail(h, JME_ERR_INVALID, "null coordinates");
    JME_CUDA(h, cudaSetDevice(h->device));
    if (n_atoms > 0)
        JME_CUDA(h, cudaMemcpyAsync(h->d_xyz_stage, xyz, 3 * (size_t)n_atoms * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return coords_from_device_aos(h, h->d_xyz_stage);
}

int jme_set_coords_device(jme_handle_t* h, const double* d_xyz, int64_t n_atoms) {
    if (!h) return JME_ERR_INVALID;
    if (n_atoms != h->hs.n) return fail(h, JME_ERR_INVALID, "atom count differs from the parameterised system");
    if (n_atoms > 0 && !d_xyz) return fail(h, JME_ERR_INVALID, "null coordinates");
    return coords_from_device_aos(h, d_xyz);
}

int jme_set_coord1(jme_handle_t* h, int64_t atom, int axis, double value) {
    if (!h) return JME_ERR_INVALID;
    if (atom < 0 || atom >= h->hs.n || axis < 0 || axis > 2) return fail(h, JME_ERR_INVALID, "atom or axis out of range");
    if (!h->coords_set) return fail(h, JME_ERR_NOT_ASSIGNED, "set all coordinates once before moving single ones");
    set_coord1_kernel<<<1, 1, 0, h->stream>>>(h->d_x, h->d_y, h->d_z, (int)atom, axis, value);
    ++h->launches;
    JME_CUDA(h, cudaGetLastError());
    h->cells.coords_changed();
    return JME_OK;
}

int jme_get_coords(jme_handle_t* h, double* xyz, int64_t n_atoms) {
    if (!h || !xyz) return JME_ERR_INVALID;
    if (n_atoms != h->hs.n) return fail(h, JME_ERR_INVALID, "atom count differs");
    if (!h->coords_set) return fail(h, JME_ERR_NOT_ASSIGNED, "no coordinates set");
    const int n = h->ds.n;
    if (n == 0) return JME_OK;
    soa_to_aos_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_x, h->d_y, h->d_z, n, h->d_xyz_stage);
    ++h->launches;
    JME_CUDA(h, cudaGetLastError());
    JME_CUDA(h, cudaMemcpyAsync(xyz, h->d_xyz_stage, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    JME_CUDA(h, cudaStreamSynchronize(h->stream));
    return JME_OK;
}

int jme_energy(jme_handle_t* h, double* total, double comps[JME_N_COMPONENTS]) {
    if (!h) return JME_ERR_INVALID;
    JME_CUDA(h, cudaSetDevice(h->device));
    return eval_host(h, false, total, comps, nullptr);
}

int jme_energy_gradient(jme_handle_t* h, double* total, double comps[JME_N_COMPONENTS], double* grad_xyz) {
    if (!h) return JME_ERR_INVALID;
    if (!grad_xyz && h->hs.n > 0) return fail(h, JME_ERR_INVALID, "null gradient buffer");
    JME_CUDA(h, cudaSetDevice(h->device));
    return eval_host(h, true, total, comps, grad_xyz);
}

// The change of the seven components ([1..7]) and of the total ([0]) that moving ONE coordinate causes, and the move
// itself (jme_delta.cuh): O(N) work.  delta: 8 doubles in the handle's energy unit.
namespace {
// device copies of the exclusion rows, partial / control words and the mapped result buffer of the delta path
int ensure_delta_data(jme_handle_t* h) {
    if (h->h_delta) return JME_OK;
    int rc;
    std::vector<long long> ptr(h->hs.excl_ptr.begin(), h->hs.excl_ptr.end());
    std::vector<unsigned int> ent(h->hs.excl_ent.begin(), h->hs.excl_ent.end());
    if ((rc = dev_upload(h, &h->d_dx_excl_ptr, ptr))) return rc;
    if ((rc = dev_upload(h, &h->d_dx_excl_ent, ent))) return rc;
    if ((rc = dev_alloc(h, &h->d_delta_part, 2 * (size_t)kDeltaMaxBlocks))) return rc;
    if ((rc = dev_alloc(h, &h->d_delta_ctr, 1))) return rc;
    JME_CUDA(h, cudaMemset(h->d_delta_ctr, 0, sizeof(unsigned int)));
    if (cudaHostAlloc(&h->h_delta, 16 * sizeof(double), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&h->h_delta_dev, h->h_delta, 0) != cudaSuccess)
        return fail(h, JME_ERR_CUDA, "cudaHostAlloc (mapped) failed");
    h->h_delta[8] = 0.0;
    return JME_OK;
}

// The whole minimizer loop in one kernel (jme_delta.cuh: minimize_kernel).  The coordinates are on the device already and
// e0 is their energy; on return xyz holds the final coordinates.
int minimize_on_device(jme_handle_t* h, bool adam, int max_steps, double step_size, double threshold, double beta1, double beta2,
                       double epsilon, double e0, double* xyz, int64_t* energy_calls, int* steps_done, double* step_energies,
                       int capacity) {
    int rc = ensure_plan(h);
    if (rc) return rc;
    if ((rc = ensure_delta_data(h))) return rc;
    const size_t n = (size_t)h->hs.n;
    const int cap = std::max(1, std::min(capacity, max_steps + 2));
    double *d_grads = nullptr, *d_m = nullptr, *d_v = nullptr, *d_log = nullptr, *d_energy = nullptr;
    int *d_t = nullptr, *d_steps = nullptr;
    long long* d_calls = nullptr;
    auto cleanup = [&]() {
        for (void* q : {(void*)d_grads, (void*)d_m, (void*)d_v, (void*)d_log, (void*)d_energy, (void*)d_t, (void*)d_steps, (void*)d_calls})
            if (q) cudaFree(q);
    };
    auto bail = [&](cudaError_t e, const char* what) { cleanup(); return fail(h, JME_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); };
    cudaError_t ce;
    const size_t nw = (size_t)kMinimizeCtas * 3 * n;          // per-CTA copies of the loop state
    if ((ce = cudaMalloc(&d_grads, nw * sizeof(double))) != cudaSuccess) return bail(ce, "cudaMalloc");
    if (adam) {
        if ((ce = cudaMalloc(&d_m, nw * sizeof(double))) != cudaSuccess) return bail(ce, "cudaMalloc");
        if ((ce = cudaMalloc(&d_v, nw * sizeof(double))) != cudaSuccess) return bail(ce, "cudaMalloc");
        if ((ce = cudaMalloc(&d_t, nw * sizeof(int))) != cudaSuccess) return bail(ce, "cudaMalloc");
    }
    if ((ce = cudaMalloc(&d_log, (size_t)cap * sizeof(double))) != cudaSuccess) return bail(ce, "cudaMalloc");
    if ((ce = cudaMalloc(&d_energy, sizeof(double))) != cudaSuccess) return bail(ce, "cudaMalloc");
    if ((ce = cudaMalloc(&d_steps, sizeof(int))) != cudaSuccess) return bail(ce, "cudaMalloc");
    if ((ce = cudaMalloc(&d_calls, sizeof(long long))) != cudaSuccess) return bail(ce, "cudaMalloc");
    cudaMemsetAsync(d_log, 0, (size_t)cap * sizeof(double), h->stream);
    MinimizeArgs A{};
    A.S = h->ds;
    A.x = h->d_x; A.y = h->d_y; A.z = h->d_z;
    A.cutoff_on = h->cutoff > 0 ? 1 : 0;
    A.excl_ptr = h->d_dx_excl_ptr; A.excl_ent = h->d_dx_excl_ent;
    A.B = h->bt; A.slot_ptr = h->d_slot_ptr; A.slot_idx = h->d_slot_idx;
    A.unit_num = h->unit == JME_KCAL_PER_MOL ? 1.0 : kUnitToJoule[JME_KCAL_PER_MOL];
    A.unit_den = h->unit == JME_KCAL_PER_MOL ? 1.0 : kUnitToJoule[h->unit];
    A.adam = adam ? 1 : 0;
    A.max_steps = max_steps; A.step_size = step_size; A.threshold = threshold;
    A.beta1 = beta1; A.beta2 = beta2; A.epsilon = epsilon;
    A.e0 = e0;
    A.grads = d_grads; A.adam_m = d_m; A.adam_v = d_v; A.adam_t = d_t;
    A.log = d_log; A.capacity = cap;
    A.out_calls = d_calls; A.out_steps = d_steps; A.out_energy = d_energy;
    const size_t smem = 3 * n * sizeof(double);
    if ((ce = cudaFuncSetAttribute(minimize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
        return bail(ce, "minimize_kernel shared memory");
    {   // one cluster of kMinimizeCtas CTAs
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(kMinimizeCtas, 1, 1);
        cfg.blockDim = dim3(kMinimizeThreads, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = h->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kMinimizeCtas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if ((ce = cudaLaunchKernelEx(&cfg, minimize_kernel, A)) != cudaSuccess) return bail(ce, "minimize_kernel launch");
    }
    ++h->launches;
    if ((ce = cudaGetLastError()) != cudaSuccess) return bail(ce, "minimize_kernel launch");
    h->cells.coords_changed();
    std::vector<double> log((size_t)cap);
    long long calls = 0;
    int steps = 0;
    cudaMemcpyAsync(log.data(), d_log, (size_t)cap * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(&calls, d_calls, sizeof(calls), cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(&steps, d_steps, sizeof(steps), cudaMemcpyDeviceToHost, h->stream);
    if ((ce = cudaStreamSynchronize(h->stream)) != cudaSuccess) return bail(ce, "minimize_kernel");
    cleanup();
    if ((rc = jme_get_coords(h, xyz, (int64_t)n))) return rc;
    if (step_energies) for (int k = 1; k <= steps && k < capacity && k < cap; ++k) step_energies[k] = log[(size_t)k];
    if (energy_calls) *energy_calls = calls;
    if (steps_done) *steps_done = steps;
    return JME_OK;
}
}  // namespace

namespace {
struct PendingUndo { int64_t atom = -1; int axis = 0; double value = 0; };
int energy_delta_impl(jme_handle_t* h, int64_t atom, int axis, double value, double* delta, const PendingUndo* undo);
}  // namespace

int jme_energy_delta(jme_handle_t* h, int64_t atom, int axis, double value, double* delta) {
    return energy_delta_impl(h, atom, axis, value, delta, nullptr);
}

namespace {
int energy_delta_impl(jme_handle_t* h, int64_t atom, int axis, double value, double* delta, const PendingUndo* undo) {
    if (!h || !delta) return JME_ERR_INVALID;
    if (atom < 0 || atom >= h->hs.n || axis < 0 || axis > 2) return fail(h, JME_ERR_INVALID, "atom or axis out of range");
    if (!h->coords_set) return fail(h, JME_ERR_NOT_ASSIGNED, "set all coordinates once before moving single ones");
    if (h->world != 1) return fail(h, JME_ERR_UNSUPPORTED, "jme_energy_delta on a sharded handle");
    JME_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_plan(h);                       // cutoff threshold, dielectric scale
    if (rc) return rc;
    if ((rc = ensure_delta_data(h))) return rc;
    DeltaArgs A{};
    A.S = h->ds;
    A.x = h->d_x; A.y = h->d_y; A.z = h->d_z;
    A.atom = (int)atom; A.axis = axis; A.value = value;
    A.r_atom = undo && undo->atom >= 0 ? (int)undo->atom : -1;
    A.r_axis = undo ? undo->axis : 0;
    A.r_value = undo ? undo->value : 0.0;
    A.cutoff_on = h->cutoff > 0 ? 1 : 0;
    A.excl_ptr = h->d_dx_excl_ptr; A.excl_ent = h->d_dx_excl_ent;
    A.B = h->bt; A.slot_ptr = h->d_slot_ptr; A.slot_idx = h->d_slot_idx;
    A.part = h->d_delta_part; A.done_ctr = h->d_delta_ctr;
    A.unit_num = h->unit == JME_KCAL_PER_MOL ? 1.0 : kUnitToJoule[JME_KCAL_PER_MOL];
    A.unit_den = h->unit == JME_KCAL_PER_MOL ? 1.0 : kUnitToJoule[h->unit];
    A.out_host = h->h_delta_dev;
    A.seq = (double)(++h->delta_seq);
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(kDeltaMaxBlocks, (h->hs.n + kDeltaThreads - 1) / kDeltaThreads));
    delta_kernel<<<blocks, kDeltaThreads, 0, h->stream>>>(A);
    ++h->launches;
    JME_CUDA(h, cudaGetLastError());
    h->cells.coords_changed();
    // the result arrives in mapped host memory, its sequence number last: poll it (a stream synchronisation costs more
    // than the kernel); give up after two seconds and let the synchronisation report what happened
    volatile double* flag = h->h_delta + 8;
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spin = 0; *flag != A.seq; ++spin) {
        if ((spin & 0xfff) == 0xfff && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2)) {
            JME_CUDA(h, cudaStreamSynchronize(h->stream));
            break;
        }
    }
    if (*flag != A.seq) return fail(h, JME_ERR_CUDA, "jme_energy_delta: result not written");
    for (int c = 0; c < 8; ++c) delta[c] = h->h_delta[c];
    return JME_OK;
}
}  // namespace

int jme_set_incremental(jme_handle_t* h, int mode) {
    if (!h) return JME_ERR_INVALID;
    if (mode < 0 || mode > 2) return fail(h, JME_ERR_INVALID, "incremental mode: 0, 1 or 2");
    h->incremental = mode;
    return JME_OK;
}

int jme_gd_minimize(jme_handle_t* h, int max_steps, double step_size, double threshold, double* xyz,
                    int64_t* energy_calls, int* steps_done, double* step_energies, int capacity) {
    if (!h || !xyz) return JME_ERR_INVALID;
    JME_CUDA(h, cudaSetDevice(h->device));
    const int64_t n = h->hs.n;
    int rc = jme_set_coords(h, xyz, n);
    if (rc) return rc;
    int64_t calls = 0;
    auto energy = [&](double& e) -> int {
        ++calls;
        return eval_host(h, false, &e, nullptr, nullptr);
    };
    PendingUndo undo;                                              // incremental mode: the undo of a rejected move rides on the next trial
    auto move = [&](int64_t a, int i, double d) -> int {          // movePoint, GradientDescentMinimizer.java:107-115
        xyz[3 * a + i] += d;
        if (h->incremental) { undo.atom = a; undo.axis = i; undo.value = xyz[3 * a + i]; return JME_OK; }
        return jme_set_coord1(h, a, i, xyz[3 * a + i]);
    };
    auto flush_undo = [&]() -> int {
        if (undo.atom < 0) return JME_OK;
        const int r = jme_set_coord1(h, undo.atom, undo.axis, undo.value);
        undo.atom = -1;
        return r;
    };
    int step = 0;
    std::vector<double> grads(3 * (size_t)n, step_size);           // :51-54
    bool initialized = false;
    double last_energy = 0, previous_step_energy = 0;
    if ((rc = energy(last_energy))) return rc;                     // :56
    int n_log = 0;
    if (step_energies && n_log < capacity) step_energies[n_log++] = last_energy;   // onStepConsumer(0, E)
    if (h->incremental == 2 && h->world == 1 && n > 0 && n <= kMinimizeMaxAtoms)
        return minimize_on_device(h, false, max_steps, step_size, threshold, 0, 0, 0, last_energy, xyz, energy_calls, steps_done,
                                  step_energies, capacity);
    while (step < max_steps) {
        if (step != 0 && std::fabs(previous_step_energy - last_energy) <= threshold) break;   // :62-64
        previous_step_energy = last_energy;
        for (int64_t a = 0; a < n; ++a)
            for (int i = 0; i < 3; ++i) {
                double& g = grads[3 * a + i];
                if (g == 0) continue;
                const double mag = std::min(std::fabs(step_size * g), .3);
                const double distance = mag * ((g > 0) - (g < 0));             // Math.signum
                double difference;
                if (h->incremental) {                                            // the move and its energy change in one call
                    double d8[8];
                    xyz[3 * a + i] += distance;
                    ++calls;
                    if ((rc = energy_delta_impl(h, a, i, xyz[3 * a + i], d8, &undo))) return rc;
                    undo.atom = -1;
                    difference = d8[0];
                } else {
                    if ((rc = move(a, i, distance))) return rc;
                    double e;
                    if ((rc = energy(e))) return rc;
                    difference = e - last_energy;                                // :78
                }
                g = -difference / (distance == 0 ? 1 : distance);               // :79
                if (!initialized) {
                    if ((rc = move(a, i, -distance))) return rc;
                } else if (difference > 0) {
                    if ((rc = move(a, i, -distance))) return rc;
                    g = g * 0.9;
                } else {
                    last_energy += difference;
                    g = g * 1.05;
                }
            }
        if (!initialized) {
            initialized = true;
        } else {
            ++step;
            if (step_energies && n_log < capacity) step_energies[n_log++] = last_energy;
        }
    }
    if ((rc = flush_undo())) return rc;
    if (energy_calls) *energy_calls = calls;
    if (steps_done) *steps_done = step;
    return JME_OK;
}

int jme_adam_minimize(jme_handle_t* h, int max_steps, double step_size, double threshold, double beta1, double beta2,
                      double epsilon, double* xyz, int64_t* energy_calls, int* steps_done, double* step_energies, int capacity) {
    if (!h || !xyz) return JME_ERR_INVALID;
    JME_CUDA(h, cudaSetDevice(h->device));
    const int64_t n = h->hs.n;
    int rc = jme_set_coords(h, xyz, n);
    if (rc) return rc;
    int64_t calls = 0;
    auto energy = [&](double& e) -> int {
        ++calls;
        return eval_host(h, false, &e, nullptr, nullptr);
    };
    PendingUndo undo;                                              // incremental mode: see jme_gd_minimize
    auto move = [&](int64_t a, int i, double d) -> int {          // movePoint, AdamMinimizer.java:122-130
        xyz[3 * a + i] += d;
        if (h->incremental) { undo.atom = a; undo.axis = i; undo.value = xyz[3 * a + i]; return JME_OK; }
        return jme_set_coord1(h, a, i, xyz[3 * a + i]);
    };
    auto flush_undo = [&]() -> int {
        if (undo.atom < 0) return JME_OK;
        const int r = jme_set_coord1(h, undo.atom, undo.axis, undo.value);
        undo.atom = -1;
        return r;
    };
    struct Adam { double m = 0, v = 0; int t = 1; };              // AdamOptimizer, :132-161
    std::vector<Adam> adams(3 * (size_t)n);
    std::vector<double> grads(3 * (size_t)n, step_size);           // :63-70
    int step = 0;
    bool initialized = false;
    double last_energy = 0, energy_per_step = 0;
    if ((rc = energy(last_energy))) return rc;                     // :73
    int n_log = 0;
    if (step_energies && n_log < capacity) step_energies[n_log++] = last_energy;
    if (h->incremental == 2 && h->world == 1 && n > 0 && n <= kMinimizeMaxAtoms)
        return minimize_on_device(h, true, max_steps, step_size, threshold, beta1, beta2, epsilon, last_energy, xyz, energy_calls,
                                  steps_done, step_energies, capacity);
    while (step < max_steps) {
        if (step != 0 && std::fabs(energy_per_step - last_energy) <= threshold) break;    // :78-80
        energy_per_step = last_energy;
        for (int64_t a = 0; a < n; ++a)
            for (int i = 0; i < 3; ++i) {
                double& g = grads[3 * a + i];
                if (g == 0) continue;                              // :86-88
                double initial = 0;
                if (h->incremental) ++calls;                       // (the energy before the move is the running one)
                else if ((rc = energy(initial))) return rc;        // :89
                Adam& o = adams[3 * a + i];
                const double x = xyz[3 * a + i];
                ++o.t;                                             // optimize(), :153-160
                o.m = beta1 * o.m + (1 - beta1) * g;
                o.v = beta2 * o.v + (1 - beta2) * g * g;
                const double m_hat = o.m / (1 - std::pow(beta1, o.t));
                const double v_hat = o.v / (1 - std::pow(beta2, o.t));
                const double distance = (x - step_size * m_hat / (std::sqrt(v_hat) + epsilon)) - x;   // :90
                double difference;
                if (h->incremental) {
                    double d8[8];
                    xyz[3 * a + i] += distance;
                    ++calls;
                    if ((rc = energy_delta_impl(h, a, i, xyz[3 * a + i], d8, &undo))) return rc;
                    undo.atom = -1;
                    difference = d8[0];
                } else {
                    if ((rc = move(a, i, distance))) return rc;
                    double e;
                    if ((rc = energy(e))) return rc;
                    difference = e - initial;                      // :93
                }
                g = -difference / std::fabs(distance);             // :94 (a zero step gives +-inf / NaN, like Java)
                if (!initialized || difference > 0) {
                    if ((rc = move(a, i, -distance))) return rc;   // :95-99
                } else {
                    last_energy += difference;                     // :101
                }
            }
        if (!initialized) {
            initialized = true;
        } else {
            ++step;
            if (step_energies && n_log < capacity) step_energies[n_log++] = last_energy;
        }
    }
    if ((rc = flush_undo())) return rc;
    if (energy_calls) *energy_calls = calls;
    if (steps_done) *steps_done = step;
    return JME_OK;
}

int64_t jme_eval_out_size(const jme_handle_t* h, int want_gradient) {
    if (!h) return 0;
    return kOutHeader + (want_gradient ? 3 * h->hs.n : 0);
}

int jme_eval_device(jme_handle_t* h, int want_gradient, double* d_out) {
    if (!h || !d_out) return JME_ERR_INVALID;
    return enqueue_eval(h, want_gradient != 0, d_out);
}

int jme_pair_stats(jme_handle_t* h, uint64_t* pairs_evaluated, uint64_t* pairs_candidate) {
    if (!h) return JME_ERR_INVALID;
    if (pairs_evaluated) *pairs_evaluated = h->last_pairs;
    if (pairs_candidate) *pairs_candidate = h->last_cand;
    return JME_OK;
}

int jme_pair_list(jme_handle_t* h, int32_t* pair_i, int32_t* pair_j, uint8_t* is14, int64_t capacity, int64_t* n_pairs) {
    if (!h || !n_pairs || capacity < 0) return JME_ERR_INVALID;
    if (!h->coords_set) return fail(h, JME_ERR_NOT_ASSIGNED, "no coordinates set");
    JME_CUDA(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = ensure_plan(h))) return rc;
    int *d_i = nullptr, *d_j = nullptr;
    unsigned char* d_f = nullptr;
    unsigned long long* d_cur = nullptr;
    const size_t cap = (size_t)std::max<int64_t>(1, capacity);
    JME_CUDA(h, cudaMalloc(&d_i, cap * sizeof(int)));
    JME_CUDA(h, cudaMalloc(&d_j, cap * sizeof(int)));
    JME_CUDA(h, cudaMalloc(&d_f, cap));
    JME_CUDA(h, cudaMalloc(&d_cur, sizeof(unsigned long long)));
    JME_CUDA(h, cudaMemsetAsync(d_cur, 0, sizeof(unsigned long long), h->stream));
    rc = JME_OK;
    if (h->active_path == JME_PATH_ALLPAIRS) {
        const int n_items = h->item_end - h->item_begin;
        if (n_items > 0) {
            const int grid = (n_items + kApWarps - 1) / kApWarps;
            if (h->cutoff > 0)
                ap_pairlist_kernel<true><<<grid, kApThreads, 0, h->stream>>>(h->ds, h->d_items, h->item_begin, h->item_end, h->d_tile_mask, h->d_masks, h->ds.n_pad / kSuper, d_i, d_j, d_f, capacity, d_cur);
            else
                ap_pairlist_kernel<false><<<grid, kApThreads, 0, h->stream>>>(h->ds, h->d_items, h->item_begin, h->item_end, h->d_tile_mask, h->d_masks, h->ds.n_pad / kSuper, d_i, d_j, d_f, capacity, d_cur);
            ++h->launches;
        }
    } else {
        // a truncated survivor list would silently drop pairs: grow the lists and redo, like eval_host
        for (;;) {
            h->cells.coords_changed();
            rc = h->cells.cluster != 0
                     ? cl_run_pairlist(h->cells, h->ds, d_i, d_j, d_f, capacity, d_cur, h->stream, h->launches, h->err)
                     : run_cell_pairlist(h->cells, h->ds, d_i, d_j, d_f, capacity, d_cur, h->stream, h->launches, h->err);
            int what = 0;
            if (rc || cudaStreamSynchronize(h->stream) != cudaSuccess || !(what = cell_overflowed(h->cells))) break;
            if ((rc = grow_cell_lists(h, what))) break;
            ++h->graph_epoch;
            cudaMemsetAsync(d_cur, 0, sizeof(unsigned long long), h->stream);
        }
    }
    unsigned long long cnt = 0;
    cudaError_t ce = cudaMemcpyAsync(&cnt, d_cur, sizeof(cnt), cudaMemcpyDeviceToHost, h->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
    const size_t m = (size_t)std::min<unsigned long long>(cnt, (unsigned long long)capacity);
    if (ce == cudaSuccess && m > 0 && pair_i && pair_j && is14) {
        cudaMemcpy(pair_i, d_i, m * sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(pair_j, d_j, m * sizeof(int), cudaMemcpyDeviceToHost);
        ce = cudaMemcpy(is14, d_f, m, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_i); cudaFree(d_j); cudaFree(d_f); cudaFree(d_cur);
    if (ce != cudaSuccess) return fail(h, JME_ERR_CUDA, std::string("pair list: ") + cudaGetErrorString(ce));
    *n_pairs = (int64_t)cnt;
    return rc;
}

int jme_set_profiling(jme_handle_t* h, int on) {
    if (!h) return JME_ERR_INVALID;
    h->profiling = on != 0;
    h->prof_used = 0;
    return JME_OK;
}

int jme_profile_read(jme_handle_t* h, float* nonbonded_ms, float* build_ms, int capacity, int* n_read) {
    if (!h || !n_read || capacity < 0) return JME_ERR_INVALID;
    const size_t m = std::min<size_t>(h->prof_used, (size_t)capacity);
    for (size_t k = 0; k < m; ++k) {
        cudaEvent_t* e = &h->prof_events[3 * k];
        float a = 0, b = 0;
        JME_CUDA(h, cudaEventSynchronize(e[2]));
        JME_CUDA(h, cudaEventElapsedTime(&a, e[0], e[1]));
        JME_CUDA(h, cudaEventElapsedTime(&b, e[1], e[2]));
        if (build_ms) build_ms[k] = a;
        if (nonbonded_ms) nonbonded_ms[k] = b;
    }
    *n_read = (int)m;
    h->prof_used = 0;
    return JME_OK;
}

int jme_measure_fp64_peak(double* tflops, double* sm_clock_mhz_estimate) {
    if (!tflops) return JME_ERR_INVALID;
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return JME_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double* d = nullptr;
    if (cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return JME_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return JME_ERR_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    const double fma_count = (double)blocks * threads * (double)iters * 8.0;
    *tflops = 2.0 * fma_count / (best * 1e-3) / 1e12;
    if (sm_clock_mhz_estimate)   // 64 FP64 FMA lanes per SM per clock on sm_100
        *sm_clock_mhz_estimate = fma_count / (best * 1e-3) / (64.0 * prop.multiProcessorCount) / 1e6;
    return JME_OK;
}

int jme_enumerate_terms(int64_t n_atoms, const int32_t* atom_type, const uint8_t* linear, int64_t n_bonds,
                        const int32_t* bond_i, const int32_t* bond_j, int32_t* angles, int64_t* n_angles,
                        int32_t* oops, int64_t* n_oops, int32_t* torsions, int64_t* n_torsions) {
    if (n_atoms < 0 || n_bonds < 0 || !n_angles || !n_oops || !n_torsions) return JME_ERR_INVALID;
    if (n_atoms > 0 && (!atom_type || !linear)) return JME_ERR_INVALID;
    if (n_bonds > 0 && (!bond_i || !bond_j)) return JME_ERR_INVALID;
    for (int64_t b = 0; b < n_bonds; ++b)
        if (bond_i[b] < 0 || bond_i[b] >= n_atoms || bond_j[b] < 0 || bond_j[b] >= n_atoms) return JME_ERR_INVALID;
    std::vector<int32_t> A, O, T;
    enumerate_terms(n_atoms, atom_type, linear, n_bonds, bond_i, bond_j, A, O, T);
    if (angles) std::copy(A.begin(), A.end(), angles);
    if (oops) std::copy(O.begin(), O.end(), oops);
    if (torsions) std::copy(T.begin(), T.end(), torsions);
    *n_angles = (int64_t)A.size() / 3;
    *n_oops = (int64_t)O.size() / 4;
    *n_torsions = (int64_t)T.size() / 4;
    return JME_OK;
}

int jme_exclusions(int64_t n_atoms, int64_t n_bonds, const int32_t* bond_i, const int32_t* bond_j,
                   int64_t* row_ptr, int32_t* partner, uint8_t* is14, int64_t capacity, int64_t* n_entries) {
    if (n_atoms < 0 || n_bonds < 0 || !n_entries) return JME_ERR_INVALID;
    for (int64_t b = 0; b < n_bonds; ++b)
        if (bond_i[b] < 0 || bond_i[b] >= n_atoms || bond_j[b] < 0 || bond_j[b] >= n_atoms) return JME_ERR_INVALID;
    std::vector<int64_t> ptr;
    std::vector<uint32_t> ent;
    build_exclusions(n_atoms, n_bonds, bond_i, bond_j, ptr, ent);
    *n_entries = (int64_t)ent.size();
    if (row_ptr) std::copy(ptr.begin(), ptr.end(), row_ptr);
    if (partner && is14)
        for (int64_t e = 0; e < std::min<int64_t>(capacity, (int64_t)ent.size()); ++e) {
            partner[e] = (int32_t)(ent[e] & ~kFlag14);
            is14[e] = (ent[e] & kFlag14) ? 1 : 0;
        }
    return JME_OK;
}

// ---- owner-mode exchange over NVLink peer memory (jme_peer.cuh) ----------------------------------------------
struct jme_peer_comm {
    int rank = 0, world = 1, device = 0;
    int64_t n = 0, chunk = 0;              // atoms, atoms per rank (even)
    size_t xyz_off = 0, out_off = 0, total = 0;
    unsigned char* base = nullptr;         // this rank's exported block
    unsigned char* peer[kPeerMaxWorld] = {};
    bool opened = false;
    unsigned long long seq = 0;
    int sms = 148;
    std::string err;
};

int jme_peer_create(int rank, int world, int64_t n_atoms, jme_peer_comm_t** out) {
    if (!out || world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world || n_atoms < 0) return JME_ERR_INVALID;
    *out = nullptr;
    jme_peer_comm* c = new (std::nothrow) jme_peer_comm();
    if (!c) return JME_ERR_NOMEM;
    c->rank = rank; c->world = world; c->n = n_atoms;
    c->chunk = ((n_atoms + world - 1) / world + 1) / 2 * 2;
    const size_t xyz_bytes = ((size_t)3 * c->chunk * world * sizeof(double) + 255) / 256 * 256;
    const size_t out_bytes = (((size_t)kOutHeader + (size_t)3 * c->chunk * world) * sizeof(double) + 255) / 256 * 256;
    c->xyz_off = kPeerCtrlBytes;
    c->out_off = c->xyz_off + xyz_bytes;
    c->total = c->out_off + out_bytes;
    cudaError_t e = cudaGetDevice(&c->device);
    if (e == cudaSuccess) e = cudaMalloc(&c->base, c->total);
    if (e == cudaSuccess) e = cudaMemset(c->base, 0, c->total);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, c->device);
    if (e != cudaSuccess) { if (c->base) cudaFree(c->base); delete c; return e == cudaErrorMemoryAllocation ? JME_ERR_NOMEM : JME_ERR_CUDA; }
    c->peer[rank] = c->base;
    *out = c;
    return JME_OK;
}

int64_t jme_peer_chunk(const jme_peer_comm_t* c) { return c ? c->chunk : 0; }
const char* jme_peer_last_error(const jme_peer_comm_t* c) { return c ? c->err.c_str() : "null communicator"; }

int jme_peer_handle(const jme_peer_comm_t* c, void* handle64) {
    if (!c || !handle64) return JME_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hd;
    if (cudaIpcGetMemHandle(&hd, c->base) != cudaSuccess) { cudaGetLastError(); return JME_ERR_CUDA; }
    std::memcpy(handle64, &hd, 64);
    return JME_OK;
}

int jme_peer_open(jme_peer_comm_t* c, const void* handles) {
    if (!c || !handles) return JME_ERR_INVALID;
    if (c->opened) return JME_OK;
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) continue;
        cudaIpcMemHandle_t hd;
        std::memcpy(&hd, static_cast<const unsigned char*>(handles) + 64 * (size_t)p, 64);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            c->err = std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(p) + "): " + cudaGetErrorString(e);
            cudaGetLastError();
            return JME_ERR_CUDA;
        }
        c->peer[p] = static_cast<unsigned char*>(ptr);
    }
    c->opened = true;
    return JME_OK;
}

// One owner-mode step, enqueued on the evaluation handle's stream: d_xyz_own = the 3 * chunk coordinates of this
// rank's atoms, d_grad_own receives their 3 * chunk gradient doubles, d_header the job-wide 10-double header.
int jme_peer_step(jme_peer_comm_t* c, jme_handle_t* h, const double* d_xyz_own, double* d_grad_own, double* d_header) {
    if (!c || !h || !d_xyz_own || !d_grad_own || !d_header) return JME_ERR_INVALID;
    if (!c->opened && c->world > 1) return fail(h, JME_ERR_INVALID, "peer communicator: jme_peer_open first");
    if (h->hs.n != c->n || h->world != c->world || h->rank != c->rank) return fail(h, JME_ERR_INVALID, "peer communicator and handle disagree (atoms / shard)");
    const unsigned long long seq = ++c->seq;
    PeerPtrs P{};
    for (int p = 0; p < c->world; ++p) P.base[p] = c->peer[p];
    const long long slice = 3 * c->chunk;
    cudaStream_t st = h->stream;
    const int bx = (int)std::max<long long>(1, std::min<long long>(2 * c->sms / c->world + 1, (slice / 2 + 255) / 256));
    peer_push_kernel<<<dim3(bx, c->world), 256, 0, st>>>(P, c->rank, c->xyz_off, slice, d_xyz_own, seq);
    peer_wait_kernel<<<1, 32, 0, st>>>(c->base, kPeerPushSeq, c->world, seq);
    h->launches += 2;
    int rc = jme_set_coords_device(h, reinterpret_cast<const double*>(c->base + c->xyz_off), c->n);
    if (rc) return rc;
    if ((rc = jme_eval_device(h, 1, reinterpret_cast<double*>(c->base + c->out_off)))) return rc;
    peer_done_kernel<<<1, 32, 0, st>>>(P, c->rank, c->world, seq);
    const int gx = (int)std::max<long long>(1, std::min<long long>(2 * c->sms, (slice / 2 + 255) / 256));
    peer_pull_kernel<<<gx, 256, 0, st>>>(P, c->rank, c->world, c->out_off, slice, d_grad_own, d_header, seq);
    h->launches += 2;
    JME_CUDA(h, cudaGetLastError());
    return JME_OK;
}

void jme_peer_destroy(jme_peer_comm_t* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < c->world; ++p)
        if (p != c->rank && c->peer[p]) cudaIpcCloseMemHandle(c->peer[p]);
    if (c->base) cudaFree(c->base);
    delete c;
}

}  // extern "C"
