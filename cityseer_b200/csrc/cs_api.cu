// C ABI + host-side graph builder for the B200 centrality library (see include/cityseer_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "cs_shortest.cuh"
#include "cs_shortest3.cuh"
#include "cs_segment.cuh"
#include "cs_segment3.cuh"
#include "cs_simplest.cuh"
#include "cs_tree.cuh"

// ------------------------------------------------------------------------------------------------ error handling
static thread_local std::string g_err;
static int cs_fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CS_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) return cs_fail("CUDA error %s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define CS_CUDA_NULL(call)                                                                              \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            cs_fail("CUDA error %s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__);              \
            return nullptr;                                                                             \
        }                                                                                               \
    } while (0)

// ------------------------------------------------------------------------------------------------ graph handle
struct cs_graph {
    int device = 0;
    int sm_count = 0;
    uint32_t n = 0;
    uint64_t E = 0;  // existing directed edges
    bool is_dual = false;
    int dual_status = 0;  // 0 ok, 1 missing key, 2 more than two endpoints (reported when simplest is called)
    bool twin_missing = false;
    // device graph
    uint32_t *d_in_off = nullptr, *d_out_off = nullptr;
    CsEdge *d_in_rec = nullptr, *d_out_rec = nullptr, *d_ang_rec = nullptr;
    float *d_in_num = nullptr, *d_out_num = nullptr, *d_ang_num = nullptr, *d_in_imp = nullptr, *d_weight = nullptr;
    uint8_t* d_live = nullptr;
    float cached_speed = -1.f, cached_ang_speed = -1.f;
    // per-call arrays
    uint32_t* d_sources = nullptr;
    float* d_src_wt = nullptr;
    uint8_t* d_eligible = nullptr;
    uint64_t sources_cap = 0, n_resident_sources = 0;
    unsigned long long* d_counters = nullptr;
    int* d_error = nullptr;
    uint32_t* d_redo = nullptr;  // segment: sources set aside for the heap-order replay
    uint64_t redo_cap = 0;
    double* d_out = nullptr;
    size_t out_cap = 0;
    double* d_acc = nullptr;  // node-interleaved accumulators [n][cw] + [n][bw] (shortest)
    size_t acc_cap = 0;
    // arena
    uint8_t* d_arena = nullptr;
    size_t arena_bytes = 0;
    CsArenaLayout lay{};
    CsAngLayout ang_lay{};
    int arena_D = 0;
    int arena_kind = -1;
    uint32_t workers = 0, cfg_workers = 0, cfg_rcap = 0;
    bool rcap_user = false;  // reach_capacity pinned by cs_graph_configure: overflow is an error, not a reason to grow
    float cfg_delta = 0.f;
    float mean_edge_len = 0.f;
    cudaStream_t stream = nullptr, side_stream = nullptr, own_stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long* h_progress = nullptr;  // pinned
    std::mutex side_mu;
    // chain-contracted kernel (cs_shortest3.cuh)
    bool v3_ok = false;
    uint32_t v3_J = 0, v3_I = 0;
    size_t v3_ncsec = 0;
    uint2* d3_jinfo = nullptr;
    uint4 *d3_links = nullptr, *d3_ctab = nullptr;
    float *d3_cnum = nullptr, *d3_csec = nullptr, *d3_weight = nullptr, *d3_clen = nullptr, *d3_cimp = nullptr;
    bool v3_loops = false;  // the graph has self-loops (not part of the contracted copy)
    // small node-level arena of the segment heap-order replay when the chain kernel serves the call
    uint8_t* d_arena2 = nullptr;
    CsArenaLayout lay2{};
    uint32_t workers2 = 0;
    int arena2_D = 0;
    uint32_t *d3_int_chain = nullptr, *d3_orig_of_new = nullptr, *d3_new_of_orig = nullptr;
    uint8_t* d3_eligible = nullptr;
    float cached_speed3 = -1.f;
    // OD betweenness: per-call destination lists on the device
    unsigned long long* d_od_off = nullptr;
    uint32_t* d_od_dst = nullptr;
    float* d_od_w = nullptr;
    size_t od_off_cap = 0, od_pairs_cap = 0;
    // the same lists for the chain-contracted kernel: destinations in its node numbering, ascending per origin
    uint32_t* d3_od_dst = nullptr;
    float* d3_od_w = nullptr;
    size_t od3_pairs_cap = 0;
    std::vector<uint32_t> h3_new_of_orig;
    int last_kernel = 1;
    int last_herr = 0;   // device error code of the last call (finish_call)
    int opt_kernel = 0;  // 0 auto (chain-contracted kernel when the graph qualifies, else the global-arena kernel),
                         // 1 global-arena kernel, 3 chain-contracted kernel (required)
    float opt_delta_factor = 12.0f;
    bool seg_optin = false;  // segment kernel: dynamic shared memory opt-in done on this device
    std::vector<uint32_t> in_edge_at;  // container edge id stored at each in-CSR position (tree dumps report edge ids)
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" const char* cs_last_error(void) { return g_err.c_str(); }

extern "C" int cs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------ pinned result buffers
static std::mutex g_pool_mu;
static std::vector<std::pair<void*, uint64_t>> g_pool_free;  // (ptr, bytes), at most a handful
static std::vector<std::pair<void*, uint64_t>> g_pool_live;

extern "C" void* cs_host_alloc(uint64_t bytes) {
    if (bytes == 0) bytes = 8;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_pool_free.size(); ++i) {
        if (g_pool_free[i].second == bytes) {
            void* p = g_pool_free[i].first;
            g_pool_live.push_back(g_pool_free[i]);
            g_pool_free.erase(g_pool_free.begin() + i);
            return p;
        }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        // make room: drop the cached buffers and retry once
        for (auto& f : g_pool_free) cudaFreeHost(f.first);
        g_pool_free.clear();
        if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            cs_fail("cudaHostAlloc of %llu bytes failed", (unsigned long long)bytes);
            return nullptr;
        }
    }
    g_pool_live.push_back({p, bytes});
    return p;
}

extern "C" void cs_host_free(void* ptr) {
    if (!ptr) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_pool_live.size(); ++i) {
        if (g_pool_live[i].first == ptr) {
            auto e = g_pool_live[i];
            g_pool_live.erase(g_pool_live.begin() + i);
            if (g_pool_free.size() < 4) {
                g_pool_free.push_back(e);
            } else {
                cudaFreeHost(e.first);
            }
            return;
        }
    }
}

// Tobler slope penalty and the numerator of edge_travel_seconds, evaluated in f32 left to right exactly as
// centrality.rs:969-1007: (length * imp * slope_pen) / speed.  The division by the per-call speed happens on device.
static inline float slope_penalty(const double* z, uint32_t from, uint32_t to, float length_2d) {
    if (length_2d <= 0.0f) return 1.0f;
    double zf = z[from], zt = z[to];
    if (std::isnan(zf) || std::isnan(zt)) return 1.0f;
    float slope = (float)(zt - zf) / length_2d;
    const float FLAT_FACTOR = 0.839457f;
    float slope_factor = std::exp(-3.5f * std::fabs(slope + 0.05f));
    return FLAT_FACTOR / slope_factor;
}

template <class T>
static int upload(T** dptr, const std::vector<T>& h) {
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    CS_CUDA(cudaMalloc(dptr, bytes));
    if (!h.empty()) CS_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

static int build_v3_graph(cs_graph* g, uint32_t n, const uint8_t* node_exists, const double* xs, const double* ys,
                          const std::vector<uint32_t>& in_off, const std::vector<uint32_t>& out_off,
                          const std::vector<CsEdge>& in_rec, const std::vector<CsEdge>& out_rec,
                          const std::vector<float>& in_num, const std::vector<float>& out_num,
                          const std::vector<float>& in_imp, const std::vector<float>& weight);

extern "C" cs_graph* cs_graph_create(uint32_t node_bound, const uint8_t* node_exists, const uint8_t* live,
                                     const float* weight, const double* xs, const double* ys, const double* z,
                                     uint64_t edge_bound,
                                     const uint8_t* edge_exists, const uint32_t* src, const uint32_t* dst,
                                     const uint32_t* edge_idx, const float* length, const float* angle_sum,
                                     const float* imp_factor, const float* seconds, const int32_t* shared_key,
                                     const uint64_t* stamp, int is_dual, int device) {
    int ndev = cs_device_count();
    if (ndev == 0) {
        cs_fail("no CUDA device available: cityseer_b200 has no CPU fallback");
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        cs_fail("device %d out of range (%d devices)", device, ndev);
        return nullptr;
    }
    if (node_bound == 0) {
        cs_fail("NetworkStructure contains no nodes.");
        return nullptr;
    }
    if (node_bound > CS_NODE_MASK) {
        cs_fail("node_bound %u exceeds the supported maximum of %u nodes", node_bound, CS_NODE_MASK);
        return nullptr;
    }
    CS_CUDA_NULL(cudaSetDevice(device));
    const uint32_t n = node_bound;
    // existing edges in petgraph adjacency order: newest first per node and direction
    std::vector<uint32_t> order;
    order.reserve(edge_bound);
    for (uint64_t e = 0; e < edge_bound; ++e) {
        if (!edge_exists[e]) continue;
        if (src[e] >= n || dst[e] >= n || !node_exists[src[e]] || !node_exists[dst[e]]) {
            cs_fail("edge %llu references a node that does not exist", (unsigned long long)e);
            return nullptr;
        }
        if (!std::isnan(seconds[e])) {
            // transport edge (graph.rs:948-985): travel time given, length NaN; edge_travel_seconds returns it as is
            if (!std::isfinite(seconds[e]) || seconds[e] < 0.0f) {
                cs_fail("Invalid transport edge payload : seconds must be finite and non-negative (edge %llu)", (unsigned long long)e);
                return nullptr;
            }
        } else if (!(imp_factor[e] > 0.0f) || !std::isfinite(imp_factor[e]) || !std::isfinite(length[e])) {
            cs_fail("Invalid edge payload : imp_factor must be finite and positive (> 0.0) and length finite (edge %llu)",
                    (unsigned long long)e);
            return nullptr;
        }
        order.push_back((uint32_t)e);
    }
    const uint64_t E = order.size();
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return stamp[a] > stamp[b]; });
    std::vector<uint32_t> in_off(n + 1, 0), out_off(n + 1, 0);
    for (uint32_t e : order) {
        in_off[dst[e] + 1]++;
        out_off[src[e] + 1]++;
    }
    uint32_t max_deg = 0;
    for (uint32_t i = 0; i < n; ++i) {
        max_deg = std::max(max_deg, std::max(in_off[i + 1], out_off[i + 1]));
        in_off[i + 1] += in_off[i];
        out_off[i + 1] += out_off[i];
    }
    if (max_deg > CS_MAX_DEGREE) {
        cs_fail("node degree %u exceeds the supported maximum of %d", max_deg, CS_MAX_DEGREE);
        return nullptr;
    }
    std::vector<uint32_t> in_fill(in_off.begin(), in_off.end() - 1), out_fill(out_off.begin(), out_off.end() - 1);
    std::vector<uint32_t> in_slot(edge_bound, 0), out_slot(edge_bound, 0);  // CSR position of each edge id
    for (uint32_t e : order) {
        in_slot[e] = in_fill[dst[e]]++;
        out_slot[e] = out_fill[src[e]]++;
    }
    std::vector<uint32_t> out_edge_at(E);  // edge id stored at each out-CSR position
    for (uint32_t e : order) out_edge_at[out_slot[e]] = e;

    // canonical representative per (min, max, edge_idx) group, self-loops excluded (centrality.rs:494-507)
    std::vector<uint8_t> canonical(edge_bound, 0);
    {
        std::vector<uint32_t> idx;
        idx.reserve(E);
        for (uint32_t e : order)
            if (src[e] != dst[e]) idx.push_back(e);
        auto key_lt = [&](uint32_t a, uint32_t b) {
            uint32_t alo = std::min(src[a], dst[a]), ahi = std::max(src[a], dst[a]);
            uint32_t blo = std::min(src[b], dst[b]), bhi = std::max(src[b], dst[b]);
            if (alo != blo) return alo < blo;
            if (ahi != bhi) return ahi < bhi;
            if (edge_idx[a] != edge_idx[b]) return edge_idx[a] < edge_idx[b];
            return a < b;
        };
        std::sort(idx.begin(), idx.end(), key_lt);
        for (size_t i = 0; i < idx.size(); ++i) {
            bool first = i == 0;
            if (!first) {
                uint32_t a = idx[i - 1], b = idx[i];
                first = !(std::min(src[a], dst[a]) == std::min(src[b], dst[b]) &&
                          std::max(src[a], dst[a]) == std::max(src[b], dst[b]) && edge_idx[a] == edge_idx[b]);
            }
            if (first) canonical[idx[i]] = 1;
        }
    }

    // dual endpoint slots (centrality.rs:533-566): first-appearance order over edge references (edge-index order)
    std::vector<int32_t> slot0, slot1;
    int dual_status = 0;
    if (is_dual) {
        slot0.assign(n, -1);
        slot1.assign(n, -1);
        for (uint64_t e = 0; e < edge_bound && dual_status == 0; ++e) {
            if (!edge_exists[e]) continue;
            int32_t k = shared_key ? shared_key[e] : -1;
            if (k < 0) {
                dual_status = 1;
                break;
            }
            for (uint32_t nd : {src[e], dst[e]}) {
                if (slot0[nd] == k || slot1[nd] == k) continue;
                if (slot0[nd] < 0)
                    slot0[nd] = k;
                else if (slot1[nd] < 0)
                    slot1[nd] = k;
                else {
                    dual_status = 2;
                    break;
                }
            }
        }
    }

    std::vector<CsEdge> in_rec(E), out_rec(E), ang_rec;
    std::vector<float> in_num(E), out_num(E), in_imp(E), ang_num;
    if (is_dual && dual_status == 0) {
        ang_rec.resize(E);
        ang_num.resize(E);
    }
    bool twin_missing = false;
    double len_sum = 0.0;
    uint64_t len_cnt = 0;
    for (uint32_t e : order) {
        const uint32_t s = src[e], d = dst[e];
        const float sp = slope_penalty(z, s, d, length[e]);
        // explicit seconds (transport edges, centrality.rs:988-990) travel through the numerator arrays with the sign
        // bit set; the prep kernels pass them on undivided
        const bool fixed_sec = !std::isnan(seconds[e]);
        const float num = fixed_sec ? -seconds[e] : (length[e] * imp_factor[e]) * sp;  // use_impedance = true
        const bool self_loop = s == d;
        if (!fixed_sec) {
            len_sum += length[e];
            len_cnt += 1;
        }
        // twin d->s with the same payload edge_idx: first match in d's out-list (graph.rs:1278-1292)
        int64_t twin = -1;
        for (uint32_t k = out_off[d]; k < out_off[d + 1]; ++k) {
            uint32_t t = out_edge_at[k];
            if (dst[t] == s && edge_idx[t] == edge_idx[e]) {
                twin = t;
                break;
            }
        }
        if (twin < 0) twin_missing = true;
        CsEdge& ir = in_rec[in_slot[e]];
        ir.nbr = s;
        ir.sec = 0.f;
        ir.aux = twin >= 0 ? length[twin] : 0.f;
        // meta[21:16] = 1 + position of the twin (d->s) inside s's in-list: the edge a search item for s skips
        const uint32_t back = (twin >= 0 && !self_loop) ? (in_slot[twin] - in_off[s] + 1u) : 0u;
        ir.meta = (out_slot[e] - out_off[s]) | (twin >= 0 ? 0x100u : 0u) | (self_loop ? 0x200u : 0u) | (back << 16);
        in_num[in_slot[e]] = num;
        in_imp[in_slot[e]] = twin >= 0 ? imp_factor[twin] : 1.0f;
        CsEdge& orc = out_rec[out_slot[e]];
        orc.nbr = d;
        orc.sec = 0.f;
        orc.aux = twin >= 0 ? length[twin] : 0.f;  // length of the twin d->s (segment: origin / last segment lengths)
        orc.meta = (in_slot[e] - in_off[d]) | (canonical[e] ? 0x100u : 0u) | (self_loop ? 0x200u : 0u) |
                   (twin >= 0 ? 0x400u : 0u);
        out_num[out_slot[e]] = num;
        if (!ang_rec.empty()) {
            // angular record: nbr | exit slot at s (bit 30) | entry slot at d (bit 31); sec uses use_impedance = false
            const int32_t k = shared_key[e];
            const uint32_t cslot = slot0[s] == k ? 0u : 1u;
            const uint32_t nslot = slot0[d] == k ? 0u : 1u;
            CsEdge& ar = ang_rec[out_slot[e]];
            ar.nbr = d | (cslot << 30) | (nslot << 31);
            ar.sec = 0.f;
            ar.aux = angle_sum[e];
            float tb = 1e-6f * length[e];  // ANGULAR_ROUTE_TIE_BREAK_FACTOR * length (centrality.rs:667)
            std::memcpy(&ar.meta, &tb, 4);
            ang_num[out_slot[e]] = fixed_sec ? -seconds[e] : (length[e] * 1.0f) * sp;
        }
    }
    if (is_dual && n >= (1u << 30)) {
        cs_fail("dual graphs support at most 2^30 nodes");
        return nullptr;
    }

    cs_graph* g = new cs_graph();
    g->device = device;
    g->n = n;
    g->E = E;
    g->is_dual = is_dual != 0;
    g->dual_status = dual_status;
    g->twin_missing = twin_missing;
    g->mean_edge_len = len_cnt ? (float)(len_sum / (double)len_cnt) : 1.0f;
    g->in_edge_at.resize(E);
    for (uint32_t e : order) g->in_edge_at[in_slot[e]] = e;
    cudaDeviceProp prop;
    CS_CUDA_NULL(cudaGetDeviceProperties(&prop, device));
    g->sm_count = prop.multiProcessorCount;
    std::vector<float> w(weight, weight + n);
    std::vector<uint8_t> lv(n);
    for (uint32_t i = 0; i < n; ++i) lv[i] = (node_exists[i] && live[i]) ? 1 : 0;
    int rc = 0;
    rc |= upload(&g->d_in_off, in_off);
    rc |= upload(&g->d_out_off, out_off);
    rc |= upload(&g->d_in_rec, in_rec);
    rc |= upload(&g->d_out_rec, out_rec);
    rc |= upload(&g->d_in_num, in_num);
    rc |= upload(&g->d_out_num, out_num);
    rc |= upload(&g->d_in_imp, in_imp);
    rc |= upload(&g->d_weight, w);
    rc |= upload(&g->d_live, lv);
    if (!ang_rec.empty()) {
        rc |= upload(&g->d_ang_rec, ang_rec);
        rc |= upload(&g->d_ang_num, ang_num);
    }
    if (!rc) rc = build_v3_graph(g, n, node_exists, xs, ys, in_off, out_off, in_rec, out_rec, in_num, out_num, in_imp, w);
    if (rc) {
        delete g;
        return nullptr;
    }
    CS_CUDA_NULL(cudaMalloc(&g->d_counters, CS_NCOUNTERS * sizeof(unsigned long long)));
    CS_CUDA_NULL(cudaMemset(g->d_counters, 0, CS_NCOUNTERS * sizeof(unsigned long long)));
    CS_CUDA_NULL(cudaMalloc(&g->d_error, sizeof(int)));
    CS_CUDA_NULL(cudaMalloc(&g->d_eligible, n));
    CS_CUDA_NULL(cudaStreamCreateWithFlags(&g->own_stream, cudaStreamNonBlocking));
    g->stream = g->own_stream;
    CS_CUDA_NULL(cudaStreamCreateWithFlags(&g->side_stream, cudaStreamNonBlocking));
    for (auto& e : g->ev) CS_CUDA_NULL(cudaEventCreate(&e));
    CS_CUDA_NULL(cudaHostAlloc(&g->h_progress, sizeof(unsigned long long), cudaHostAllocDefault));
    *g->h_progress = 0;
    return g;
}

extern "C" void cs_graph_destroy(cs_graph* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    cudaDeviceSynchronize();
    for (void* p : {(void*)g->d_in_off, (void*)g->d_out_off, (void*)g->d_in_rec, (void*)g->d_out_rec, (void*)g->d_ang_rec,
                    (void*)g->d_in_num, (void*)g->d_out_num, (void*)g->d_ang_num, (void*)g->d_in_imp, (void*)g->d_weight,
                    (void*)g->d_live, (void*)g->d_sources, (void*)g->d_src_wt, (void*)g->d_eligible, (void*)g->d_counters,
                    (void*)g->d_error, (void*)g->d_out, (void*)g->d_arena, (void*)g->d_acc, (void*)g->d3_jinfo,
                    (void*)g->d3_links, (void*)g->d3_ctab, (void*)g->d3_cnum, (void*)g->d3_csec, (void*)g->d3_weight,
                    (void*)g->d3_int_chain, (void*)g->d3_orig_of_new, (void*)g->d3_new_of_orig, (void*)g->d3_eligible,
                    (void*)g->d_od_off, (void*)g->d_od_dst, (void*)g->d_od_w, (void*)g->d3_od_dst, (void*)g->d3_od_w,
                    (void*)g->d_redo, (void*)g->d3_clen,
                    (void*)g->d3_cimp, (void*)g->d_arena2})
        if (p) cudaFree(p);
    if (g->h_progress) cudaFreeHost(g->h_progress);
    for (auto& e : g->ev)
        if (e) cudaEventDestroy(e);
    if (g->own_stream) cudaStreamDestroy(g->own_stream);
    if (g->side_stream) cudaStreamDestroy(g->side_stream);
    delete g;
}

extern "C" int cs_graph_configure(cs_graph* g, uint32_t reach_capacity, float delta_seconds, uint32_t workers) {
    if (!g) return cs_fail("null graph");
    if (reach_capacity) {
        g->cfg_rcap = reach_capacity;
        g->rcap_user = true;
    }
    if (delta_seconds > 0.f) g->cfg_delta = delta_seconds;
    if (workers) g->cfg_workers = workers;
    if (reach_capacity || workers) g->arena_kind = -1;  // force re-allocation
    return 0;
}

extern "C" int cs_graph_set_option(cs_graph* g, const char* name, double value) {
    if (!g || !name) return cs_fail("null graph or option name");
    const std::string k(name);
    if (k == "kernel") {
        if (value != 0 && value != 1 && value != 3) return cs_fail("option kernel must be 0 (auto), 1 (arena) or 3 (chain-contracted)");
        g->opt_kernel = (int)value;
    } else if (k == "delta_factor") {
        if (!(value > 0)) return cs_fail("option delta_factor must be positive");
        g->opt_delta_factor = (float)value;
    } else {
        return cs_fail("unknown option %s", name);
    }
    return 0;
}

extern "C" int cs_graph_set_stream(cs_graph* g, void* cuda_stream) {
    if (!g) return cs_fail("null graph");
    g->stream = cuda_stream ? (cudaStream_t)cuda_stream : g->own_stream;
    return 0;
}

extern "C" uint64_t cs_progress(cs_graph* g) {
    if (!g) return 0;
    std::lock_guard<std::mutex> lk(g->side_mu);
    cudaSetDevice(g->device);
    if (cudaMemcpyAsync(g->h_progress, g->d_counters + CS_C_PROGRESS, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                        g->side_stream) != cudaSuccess)
        return 0;
    cudaStreamSynchronize(g->side_stream);
    return *g->h_progress;
}

// ------------------------------------------------------------------------------------------------ arena
// Default capacity of a search in reached nodes (junctions for the chain kernel, states for the angular one).  A call
// that overflows it is repeated with four times the capacity (grow_after_overflow) unless the caller pinned the value
// with cs_graph_configure: the arena then only grows for graphs and distances that need it (the reference's own 20 km
// Greater-London runs reach 69 k nodes).
#define CS_DEFAULT_RCAP (1u << 14)

// After a failed call: true when the failure was an arena overflow and the capacity could be raised (the caller repeats
// the call; nothing of the failed attempt has reached a caller-owned buffer unless it asked to accumulate).
static bool grow_after_overflow(cs_graph* g, size_t nstates, uint32_t current) {
    if (!g || g->rcap_user) return false;
    if (g->last_herr != CS_ERR_REACH_OVERFLOW && g->last_herr != CS_ERR_QUEUE_OVERFLOW) return false;
    g->last_herr = 0;
    if ((size_t)current >= nstates) return false;
    g->cfg_rcap = (uint32_t)std::min<size_t>(nstates, (size_t)current * 4);
    g->arena_kind = -1;
    return true;
}

// kind 0 = shortest, 1 = segment, 2 = simplest (two states per node), 3 = chain-contracted shortest (junction states)
static int ensure_arena(cs_graph* g, int kind, int D) {
    if (g->d_arena && g->arena_kind == kind && g->arena_D >= D) return 0;
    if (g->d_arena) {
        CS_CUDA(cudaFree(g->d_arena));
        g->d_arena = nullptr;
    }
    const size_t nstates = kind == 2 ? (size_t)g->n * 2 : kind == 3 ? (size_t)g->v3_J + 1 : g->n;
    uint32_t rcap = g->cfg_rcap ? g->cfg_rcap : CS_DEFAULT_RCAP;
    rcap = (uint32_t)std::min<size_t>(rcap, nstates);
    rcap = std::max(rcap, 32u);
    const uint32_t qcap = rcap * 2 + 64;
    uint32_t workers = g->cfg_workers ? g->cfg_workers
                                      : (uint32_t)g->sm_count * (kind == 3   ? std::max<uint32_t>(CS3_WORKERS_PER_SM, CS3S_WARPS)
                                                                 : kind == 1 ? CS_SEG_MIN_BLOCKS * CS_SEG_WARPS
                                                                             : CS_MIN_BLOCKS * CS_WARPS_PER_CTA);
    // warps of the widest CTA that uses the arena (kind 3 serves both chain kernels)
    const uint32_t gran = kind == 3 ? std::max<uint32_t>(CS3_WORKERS_PER_SM, CS3S_WARPS) : kind == 1 ? CS_SEG_WARPS : CS_WARPS_PER_CTA;
    workers = std::max<uint32_t>(gran, workers / gran * gran);
    CsArenaLayout L{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    L.ds = take(nstates * sizeof(uint2));
    L.node_list = take((size_t)rcap * 4);
    L.qa = take((size_t)qcap * 8);
    L.qb = take((size_t)qcap * 8);
    L.far = take((size_t)qcap * 8);
    L.s_node = take((size_t)rcap * 4);
    L.s_agg = take((size_t)rcap * 4);
    L.predmask = take((size_t)rcap * 8);  // shortest: u32 mask; segment/simplest reuse as 8-byte per-rank scratch
    L.sigma = take((size_t)rcap * 8);
    L.dep = take((size_t)rcap * 2 * D * 8);
    L.bdone = take((size_t)rcap * 8);  // shortest: u8 flags; other kernels: 8-byte per-rank scratch
    if (kind != 3) L.erank = take((size_t)rcap * 16);  // node-level kernels: CSR rows by settle rank
    if (kind == 3) {
        L.frank = take((size_t)rcap * CS3_MAX_LINKS * 16);  // chain kernel: per link {candidate, neighbour, id, far rank}
        L.needm = take((size_t)rcap * 4);                  // chain kernel: links whose far junction continues the path
        L.jrank = take((size_t)rcap * 8);                  // chain kernel: junction record by settle rank
    }
    L.stride = align_up(off, 4096);
    L.rcap = rcap;
    L.qcap = qcap;
    size_t free_b = 0, total_b = 0;
    CS_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t budget = (size_t)((double)free_b * 0.80);
    while (workers > gran && (size_t)workers * L.stride > budget) workers -= gran;
    if ((size_t)workers * L.stride > budget)
        return cs_fail("not enough device memory for the search arena (%zu bytes per worker)", L.stride);
    g->arena_bytes = (size_t)workers * L.stride;
    CS_CUDA(cudaMalloc(&g->d_arena, g->arena_bytes));
    // dense maps start at {inf, none}; each search resets exactly what it touched
    {
        dim3 grid((unsigned)std::min<size_t>((nstates + 255) / 256, 64), workers);
        cs_k_init_ds<<<grid, 256, 0, g->stream>>>(g->d_arena, L.stride, L.ds, nstates);
        CS_CUDA(cudaGetLastError());
        CS_CUDA(cudaStreamSynchronize(g->stream));
    }
    g->lay = L;
    g->workers = workers;
    g->arena_D = D;
    g->arena_kind = kind;
    return 0;
}

// angular arena (two states per node, explicit predecessor lists, lane-0 heap)
static int ensure_arena_angular(cs_graph* g, int D) {
    if (g->d_arena && g->arena_kind == 2 && g->arena_D >= D) return 0;
    if (g->d_arena) {
        CS_CUDA(cudaFree(g->d_arena));
        g->d_arena = nullptr;
    }
    const size_t nstates = (size_t)g->n * 2;
    uint32_t rcap = g->cfg_rcap ? g->cfg_rcap : CS_DEFAULT_RCAP;
    rcap = (uint32_t)std::min<size_t>(std::max<size_t>(rcap, 64), nstates);
    const uint32_t hcap = rcap * 4 + 64;
    uint32_t workers = g->cfg_workers ? g->cfg_workers : (uint32_t)g->sm_count * CS_ANG_MIN_BLOCKS * CS_WARPS_PER_CTA;
    workers = std::max<uint32_t>(CS_WARPS_PER_CTA, workers / CS_WARPS_PER_CTA * CS_WARPS_PER_CTA);
    CsAngLayout L{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    L.ds = take(nstates * sizeof(uint2));
    L.dn = take((size_t)g->n * sizeof(uint2));
    L.st_state = take((size_t)rcap * 4);
    L.st_secs = take((size_t)rcap * 4);
    L.st_cost = take((size_t)rcap * 4);
    L.st_sigma = take((size_t)rcap * 8);
    L.st_np = take((size_t)rcap * 4);
    L.st_preds = take((size_t)rcap * 4 * CS_ANG_MAXPRED);
    L.order = take((size_t)rcap * 4);
    L.pos = take((size_t)rcap * 4);
    L.delta = take((size_t)rcap * 8 * D);
    L.pending = take((size_t)rcap * 4);
    L.heap = take((size_t)hcap * 8);
    L.stride = align_up(off, 4096);
    L.rcap = rcap;
    L.hcap = hcap;
    size_t free_b = 0, total_b = 0;
    CS_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t budget = (size_t)((double)free_b * 0.80);
    while (workers > CS_WARPS_PER_CTA && (size_t)workers * L.stride > budget) workers -= CS_WARPS_PER_CTA;
    if ((size_t)workers * L.stride > budget)
        return cs_fail("not enough device memory for the search arena (%zu bytes per worker)", L.stride);
    g->arena_bytes = (size_t)workers * L.stride;
    CS_CUDA(cudaMalloc(&g->d_arena, g->arena_bytes));
    dim3 grid((unsigned)std::min<size_t>((nstates + 255) / 256, 64), workers);
    cs_k_init_ang<<<grid, 256, 0, g->stream>>>(g->d_arena, L.stride, L.ds, nstates, L.dn, g->n);
    CS_CUDA(cudaGetLastError());
    CS_CUDA(cudaStreamSynchronize(g->stream));
    g->ang_lay = L;
    g->workers = workers;
    g->arena_D = D;
    g->arena_kind = 2;
    return 0;
}

// edge_travel_seconds (centrality.rs:988-1006): numerator / speed, or the edge's explicit seconds (sign bit set)
__device__ __forceinline__ float cs_edge_seconds(float num, float speed) {
    return (__float_as_uint(num) & 0x80000000u) ? fabsf(num) : __fdiv_rn(num, speed);
}
__global__ void cs_k_prep_csec(float* sec, const float* num, size_t m, float speed) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) sec[i] = cs_edge_seconds(num[i], speed);
}

__global__ void cs_k_prep_seconds(CsEdge* rec, const float* num, uint64_t E, float speed) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E) rec[i].sec = cs_edge_seconds(num[i], speed);
}

static int prep_seconds(cs_graph* g, float speed, bool angular, uint32_t* launches) {
    if (g->E == 0) return 0;
    int blocks = (int)((g->E + 255) / 256);
    if (!angular) {
        if (g->cached_speed == speed) return 0;
        cs_k_prep_seconds<<<blocks, 256, 0, g->stream>>>(g->d_in_rec, g->d_in_num, g->E, speed);
        cs_k_prep_seconds<<<blocks, 256, 0, g->stream>>>(g->d_out_rec, g->d_out_num, g->E, speed);
        *launches += 2;
        g->cached_speed = speed;
    } else {
        if (g->cached_ang_speed == speed) return 0;
        cs_k_prep_seconds<<<blocks, 256, 0, g->stream>>>(g->d_ang_rec, g->d_ang_num, g->E, speed);
        *launches += 1;
        g->cached_ang_speed = speed;
    }
    CS_CUDA(cudaGetLastError());
    return 0;
}

static int stage_sources(cs_graph* g, uint64_t n_sources, const uint32_t* sources, const float* source_wt,
                         const uint8_t* eligible) {
    if (sources == nullptr) {
        if (n_sources != g->n_resident_sources)
            return cs_fail("sources == NULL but %llu resident sources were staged, %llu requested",
                           (unsigned long long)g->n_resident_sources, (unsigned long long)n_sources);
        return 0;
    }
    if (n_sources > g->sources_cap) {
        if (g->d_sources) cudaFree(g->d_sources);
        if (g->d_src_wt) cudaFree(g->d_src_wt);
        g->d_sources = nullptr;
        g->d_src_wt = nullptr;
        CS_CUDA(cudaMalloc(&g->d_sources, std::max<uint64_t>(n_sources, 1) * 4));
        CS_CUDA(cudaMalloc(&g->d_src_wt, std::max<uint64_t>(n_sources, 1) * 4));
        g->sources_cap = n_sources;
    }
    for (uint64_t i = 0; i < n_sources; ++i)
        if (sources[i] >= g->n) return cs_fail("node index %u does not exist in the graph", sources[i]);
    if (n_sources) {
        CS_CUDA(cudaMemcpyAsync(g->d_sources, sources, n_sources * 4, cudaMemcpyHostToDevice, g->stream));
        if (source_wt)
            CS_CUDA(cudaMemcpyAsync(g->d_src_wt, source_wt, n_sources * 4, cudaMemcpyHostToDevice, g->stream));
    }
    if (eligible)
        CS_CUDA(cudaMemcpyAsync(g->d_eligible, eligible, g->n, cudaMemcpyHostToDevice, g->stream));
    else if (source_wt)
        CS_CUDA(cudaMemcpyAsync(g->d_eligible, g->d_live, g->n, cudaMemcpyDeviceToDevice, g->stream));
    g->n_resident_sources = n_sources;
    return 0;
}

extern "C" int cs_stage_sources(cs_graph* g, uint64_t n_sources, const uint32_t* sources, const float* source_wt,
                                const uint8_t* eligible) {
    if (!g) return cs_fail("null graph");
    if (!sources || !source_wt) return cs_fail("null source plan");
    CS_CUDA(cudaSetDevice(g->device));
    if (stage_sources(g, n_sources, sources, source_wt, eligible)) return 1;
    CS_CUDA(cudaStreamSynchronize(g->stream));
    return 0;
}

static int acquire_out(cs_graph* g, double* out, int out_on_device, int accumulate, size_t elems, double** d_out,
                       bool zero = true) {
    if (out_on_device) {
        *d_out = out;
    } else {
        if (accumulate) return cs_fail("accumulate requires a device output pointer");
        if (g->out_cap < elems) {
            if (g->d_out) cudaFree(g->d_out);
            g->d_out = nullptr;
            CS_CUDA(cudaMalloc(&g->d_out, elems * sizeof(double)));
            g->out_cap = elems;
        }
        *d_out = g->d_out;
    }
    if (!accumulate && zero) CS_CUDA(cudaMemsetAsync(*d_out, 0, elems * sizeof(double), g->stream));
    return 0;
}

static int finish_call(cs_graph* g, double* out, int out_on_device, size_t elems, double* d_out, cs_stats* stats,
                       uint32_t launches) {
    if (!out_on_device) CS_CUDA(cudaMemcpyAsync(out, d_out, elems * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
    CS_CUDA(cudaEventRecord(g->ev[3], g->stream));
    cudaError_t e = cudaStreamSynchronize(g->stream);
    if (e != cudaSuccess) return cs_fail("CUDA error during centrality kernel: %s", cudaGetErrorString(e));
    unsigned long long h[CS_NCOUNTERS];
    int herr = 0;
    CS_CUDA(cudaMemcpy(h, g->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(&herr, g->d_error, sizeof(int), cudaMemcpyDeviceToHost));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->sources = h[CS_C_SOURCES];
        stats->settled = h[CS_C_SETTLED];
        stats->edge_iters = h[CS_C_EDGE_ITERS];
        stats->sum_ri = h[CS_C_SUM_RI];
        stats->sum_ci = h[CS_C_SUM_CI];
        stats->relaxations = h[CS_C_RELAX];
        stats->fallback_sources = h[CS_C_FALLBACK];
        for (int i = 0; i < CS_MAX_THRESHOLDS; ++i) stats->reach_totals[i] = h[CS_C_REACH0 + i];
        for (int i = 0; i < 8; ++i) stats->phase_cycles[i] = h[CS_C_PHASE0 + i];
        stats->reach_capacity = g->lay.rcap;
        stats->kernel_used = (uint32_t)g->last_kernel;
        cudaEventElapsedTime(&stats->kernel_ms, g->ev[1], g->ev[2]);
        cudaEventElapsedTime(&stats->total_ms, g->ev[0], g->ev[3]);
        stats->gpu_launches = launches;
        stats->workers = g->workers;
    }
    g->last_herr = herr;
    // a warp that stopped on an error left its dense-map entries behind: never reuse the arena as it is
    if (herr) g->arena_kind = -1;
    if (herr == CS_ERR_REACH_OVERFLOW)
        return cs_fail("search arena overflow: a source reached more than %u nodes; raise reach_capacity via cs_graph_configure",
                       g->lay.rcap);
    if (herr == CS_ERR_QUEUE_OVERFLOW)
        return cs_fail("search queue overflow (capacity %u); raise reach_capacity via cs_graph_configure", g->lay.qcap);
    if (herr == CS_ERR_ZERO_TIE)
        return cs_fail("zero-second edge between nodes that tie on (seconds, index): the reference settles such pairs in heap "
                       "order; remove zero-length edges (tools.graphs.nx_simple_geoms does) / give transport edges a "
                       "positive travel time");
    if (herr == CS_ERR_PRED_OVERFLOW)
        return cs_fail("a dual state acquired more than %d tied predecessors (unsupported)", CS_ANG_MAXPRED);
    if (herr) return cs_fail("device error %d", herr);
    return 0;
}

static CsGraphDev graph_dev(const cs_graph* g) {
    CsGraphDev d;
    d.n = g->n;
    d.in_off = g->d_in_off;
    d.in_rec = g->d_in_rec;
    d.out_off = g->d_out_off;
    d.out_rec = g->d_out_rec;
    d.in_imp = g->d_in_imp;
    d.weight = g->d_weight;
    d.live = g->d_live;
    return d;
}

static int check_thresholds(int D, const uint32_t* seconds) {
    if (D < 1 || D > CS_MAX_THRESHOLDS) return cs_fail("number of thresholds must be in [1, %d], got %d", CS_MAX_THRESHOLDS, D);
    if (!seconds) return cs_fail("null thresholds");
    return 0;
}

static float default_delta(const cs_graph* g, float speed) {
    if (g->cfg_delta > 0.f) return g->cfg_delta;
    return std::max(1e-3f, g->opt_delta_factor * g->mean_edge_len / speed);
}

#include "cs_api_v3.inl"

template <int DT, bool OD>
static cudaError_t v3_launch_t(const CsShortest3Params& t, uint32_t workers, cudaStream_t st) {
    constexpr uint32_t smem = cs3_smem_bytes<DT>();
    constexpr uint32_t W = cs3_warps<DT>();
    cudaError_t e = cudaFuncSetAttribute(cs_k_shortest3<DT, OD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(workers / W, (t.n_sources + W - 1) / W);
    if (grid == 0) return cudaSuccess;
    cs_k_shortest3<DT, OD><<<grid, W * 32, smem, st>>>(t);
    return cudaGetLastError();
}
template <bool OD>
static cudaError_t v3_launch_od(const CsShortest3Params& t, uint32_t workers, cudaStream_t st) {
    switch (cs_shortest_dt(t.D)) {
        case 1: return v3_launch_t<1, OD>(t, workers, st);
        case 2: return v3_launch_t<2, OD>(t, workers, st);
        case 3: return v3_launch_t<3, OD>(t, workers, st);
        case 4: return v3_launch_t<4, OD>(t, workers, st);
        case 8: return v3_launch_t<8, OD>(t, workers, st);
        default: return v3_launch_t<CS_MAX_THRESHOLDS, OD>(t, workers, st);
    }
}
static cudaError_t v3_launch(const CsShortest3Params& t, uint32_t workers, cudaStream_t st) {
    return t.od_off ? v3_launch_od<true>(t, workers, st) : v3_launch_od<false>(t, workers, st);
}

// ------------------------------------------------------------------------------------------------ shortest
static int run_shortest(cs_graph* g, int D, const uint32_t* distances, const float* betas, const uint32_t* seconds,
                        float speed, float tol, int closeness, int betweenness, uint64_t n_sources,
                        const uint32_t* sources, const float* source_wt, const uint8_t* eligible, double* out,
                        int out_on_device, int accumulate, cs_stats* stats, float* dump_agg, double* dump_sigma,
                        uint32_t* dump_npred, const uint64_t* od_off = nullptr, const uint32_t* od_dst = nullptr,
                        const float* od_w = nullptr) {
    if (!g) return cs_fail("null graph");
    if (check_thresholds(D, seconds)) return 1;
    if (!closeness && !betweenness)
        return cs_fail("Either or both closeness and betweenness flags is required, but both parameters are False.");
    if (!(speed > 0.f) || !std::isfinite(speed)) return cs_fail("speed_m_s must be finite and positive, got %f", speed);
    if (!(tol >= CS_TIE_EPS)) return cs_fail("Tolerance must be >= TIE_EPSILON to avoid float-comparison bugs");
    CS_CUDA(cudaSetDevice(g->device));
    uint32_t launches = 0;
    CS_CUDA(cudaEventRecord(g->ev[0], g->stream));
    if (stage_sources(g, n_sources, sources, source_wt, eligible)) return 1;
    CS_CUDA(cudaMemsetAsync(g->d_error, 0, sizeof(int), g->stream));
    const size_t elems = (size_t)7 * D * g->n;
    double* d_out = nullptr;
    if (acquire_out(g, out, out_on_device, accumulate, elems, &d_out, false)) return 1;
    const int cw = cs_shortest_cw(D), bw = cs_shortest_bw(D);
    const size_t acc_elems = (size_t)g->n * (cw + bw);
    if (g->acc_cap < acc_elems) {
        if (g->d_acc) cudaFree(g->d_acc);
        g->d_acc = nullptr;
        g->acc_cap = 0;
        CS_CUDA(cudaMalloc(&g->d_acc, acc_elems * sizeof(double)));
        g->acc_cap = acc_elems;
    }
    uint32_t max_sec = 0;
    for (int i = 0; i < D; ++i) max_sec = std::max(max_sec, seconds[i]);

    CS_CUDA(cudaMemsetAsync(g->d_counters, 0, CS_NCOUNTERS * sizeof(unsigned long long), g->stream));
    CS_CUDA(cudaMemsetAsync(g->d_acc, 0, acc_elems * sizeof(double), g->stream));

    CsShortestParams p{};
    p.g = graph_dev(g);
    p.D = D;
    p.closeness = closeness;
    p.betweenness = betweenness;
    p.phase2 = tol > CS_TIE_EPS ? 1 : 0;
    for (int i = 0; i < D; ++i) {
        p.dist_f[i] = (float)distances[i];
        p.beta_f[i] = betas[i];
        p.beta_d[i] = (double)betas[i];
    }
    p.max_seconds = (float)max_sec;
    p.speed = speed;
    p.tol = tol;
    p.sources = g->d_sources;
    p.src_wt = g->d_src_wt;
    p.n_sources = n_sources;
    p.eligible = g->d_eligible;
    p.out = d_out;
    p.acc_c = g->d_acc;
    p.acc_b = g->d_acc + (size_t)g->n * cw;
    p.cw = cw;
    p.bw = bw;
    p.counters = g->d_counters;
    p.error = g->d_error;
    p.delta = default_delta(g, speed);
    p.bin_scale = (float)CS_NBINS / (((float)max_sec + 1.0f) * ((float)max_sec + 1.0f));  // cs_bin is quadratic
    p.dump_agg = dump_agg;
    p.dump_sigma = dump_sigma;
    p.dump_npred = dump_npred;
    if (od_off) {
        const size_t n_pairs = (size_t)od_off[n_sources];
        if (g->od_off_cap < n_sources + 1) {
            if (g->d_od_off) cudaFree(g->d_od_off);
            g->d_od_off = nullptr;
            CS_CUDA(cudaMalloc(&g->d_od_off, (n_sources + 1) * sizeof(unsigned long long)));
            g->od_off_cap = n_sources + 1;
        }
        if (g->od_pairs_cap < std::max<size_t>(n_pairs, 1)) {
            if (g->d_od_dst) cudaFree(g->d_od_dst);
            if (g->d_od_w) cudaFree(g->d_od_w);
            g->d_od_dst = nullptr;
            g->d_od_w = nullptr;
            CS_CUDA(cudaMalloc(&g->d_od_dst, std::max<size_t>(n_pairs, 1) * 4));
            CS_CUDA(cudaMalloc(&g->d_od_w, std::max<size_t>(n_pairs, 1) * 4));
            g->od_pairs_cap = std::max<size_t>(n_pairs, 1);
        }
        for (size_t j = 0; j < n_pairs; ++j)
            if (od_dst[j] >= g->n) return cs_fail("OD destination %u is out of range for node_bound %u", od_dst[j], g->n);
        CS_CUDA(cudaMemcpyAsync(g->d_od_off, od_off, (n_sources + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, g->stream));
        if (n_pairs) {
            CS_CUDA(cudaMemcpyAsync(g->d_od_dst, od_dst, n_pairs * 4, cudaMemcpyHostToDevice, g->stream));
            CS_CUDA(cudaMemcpyAsync(g->d_od_w, od_w, n_pairs * 4, cudaMemcpyHostToDevice, g->stream));
        }
        p.od_off = g->d_od_off;
        p.od_dst = g->d_od_dst;
        p.od_w = g->d_od_w;
    }
    auto launch_v1 = [&](uint64_t m) -> int {
        if (ensure_arena(g, 0, D)) return 1;
        if (prep_seconds(g, speed, false, &launches)) return 1;
        p.arena = g->d_arena;
        p.lay = g->lay;
        const uint32_t grid = (uint32_t)std::min<uint64_t>(g->workers / CS_WARPS_PER_CTA, (m + CS_WARPS_PER_CTA - 1) / CS_WARPS_PER_CTA);
        if (grid == 0) return 0;
        const int threads = CS_WARPS_PER_CTA * 32;
        if (D == 1) cs_k_shortest<1><<<grid, threads, 0, g->stream>>>(p);
        else if (D == 2) cs_k_shortest<2><<<grid, threads, 0, g->stream>>>(p);
        else if (D == 3) cs_k_shortest<3><<<grid, threads, 0, g->stream>>>(p);
        else if (D == 4) cs_k_shortest<4><<<grid, threads, 0, g->stream>>>(p);
        else if (D <= 8) cs_k_shortest<8><<<grid, threads, 0, g->stream>>>(p);
        else cs_k_shortest<CS_MAX_THRESHOLDS><<<grid, threads, 0, g->stream>>>(p);
        launches += 1;
        CS_CUDA(cudaGetLastError());
        return 0;
    };
    const int nblk = (int)((g->n + 255) / 256);


    // ---- chain-contracted kernel (cs_shortest3.cuh): junction-level search, chains walked in place
    const bool dumping = dump_agg || dump_sigma || dump_npred;  // served by the arena kernel
    // auto: the chain-contracted kernel pays off when most nodes are chain interiors (decomposed / OSM-like graphs:
    // cfg #4 1.07 M vs 0.47 M sources/s); on a graph of junctions only the arena kernel is the faster one (cfg #2:
    // 3.7 M vs 2.1 M sources/s)
    bool use_v3 = !dumping && g->v3_ok && n_sources > 0 &&
                  (g->opt_kernel == 3 || (g->opt_kernel == 0 && g->v3_I >= g->v3_J));
    if (g->opt_kernel == 3 && !use_v3 && !dumping && n_sources > 0)
        return cs_fail("the chain-contracted kernel cannot serve this graph (an edge without a mutual twin, or a junction "
                       "with more than %d links)", CS3_MAX_LINKS);
    g->last_kernel = 1;
    if (use_v3) {
        std::vector<uint32_t> dst3;  // OD lists in the contracted numbering: alive until the stream is synchronised below
        std::vector<float> w3;
        if (ensure_arena(g, 3, D)) return 1;
        if (g->cached_speed3 != speed && g->v3_ncsec) {
            cs_k_prep_csec<<<(int)((g->v3_ncsec + 255) / 256), 256, 0, g->stream>>>(g->d3_csec, g->d3_cnum, g->v3_ncsec, speed);
            launches += 1;
            g->cached_speed3 = speed;
        }
        cs_k_permute_u8<<<(g->n + 255) / 256, 256, 0, g->stream>>>(g->d_eligible, g->d3_orig_of_new, g->d3_eligible, g->n);
        launches += 1;
        CsShortest3Params t{};
        t.g.J = g->v3_J;
        t.g.I = g->v3_I;
        t.g.n = g->n;
        t.g.jinfo = g->d3_jinfo;
        t.g.links = g->d3_links;
        t.g.csec = g->d3_csec;
        t.g.ctab = g->d3_ctab;
        t.g.int_chain = g->d3_int_chain;
        t.g.orig_of_new = g->d3_orig_of_new;
        t.g.new_of_orig = g->d3_new_of_orig;
        t.g.weight = g->d3_weight;
        t.D = D;
        t.closeness = closeness;
        t.betweenness = betweenness;
        t.phase2 = p.phase2;
        for (int i = 0; i < D; ++i) {
            t.dist_f[i] = p.dist_f[i];
            t.beta_f[i] = p.beta_f[i];
            t.beta_d[i] = p.beta_d[i];
        }
        t.beta_chain = D > 1;
        for (int i = 0; i + 1 < D; ++i) t.beta_chain = t.beta_chain && (t.beta_d[i] == 2.0 * t.beta_d[i + 1]) && t.beta_d[i] > 0.0;
        t.max_seconds = p.max_seconds;
        t.speed = speed;
        t.tol = tol;
        t.sources = g->d_sources;
        t.src_wt = g->d_src_wt;
        t.n_sources = n_sources;
        t.eligible = g->d3_eligible;
        if (od_off) {
            // the kernel looks a reached node up in its origin's destinations by binary search: translate them to the
            // contracted copy's numbering and sort each origin's slice (destinations are unique per origin)
            const size_t n_pairs = (size_t)od_off[n_sources];
            if (g->od3_pairs_cap < std::max<size_t>(n_pairs, 1)) {
                if (g->d3_od_dst) cudaFree(g->d3_od_dst);
                if (g->d3_od_w) cudaFree(g->d3_od_w);
                g->d3_od_dst = nullptr;
                g->d3_od_w = nullptr;
                g->od3_pairs_cap = 0;
                CS_CUDA(cudaMalloc(&g->d3_od_dst, std::max<size_t>(n_pairs, 1) * 4));
                CS_CUDA(cudaMalloc(&g->d3_od_w, std::max<size_t>(n_pairs, 1) * 4));
                g->od3_pairs_cap = std::max<size_t>(n_pairs, 1);
            }
            std::vector<std::pair<uint32_t, float>> slice;
            dst3.resize(n_pairs);
            w3.resize(n_pairs);
            for (uint64_t k = 0; k < n_sources; ++k) {
                const size_t a = (size_t)od_off[k], b = (size_t)od_off[k + 1];
                slice.clear();
                for (size_t j = a; j < b; ++j) slice.emplace_back(g->h3_new_of_orig[od_dst[j]], od_w[j]);
                std::sort(slice.begin(), slice.end(),
                          [](const std::pair<uint32_t, float>& x, const std::pair<uint32_t, float>& y) { return x.first < y.first; });
                for (size_t j = a; j < b; ++j) {
                    dst3[j] = slice[j - a].first;
                    w3[j] = slice[j - a].second;
                }
            }
            if (n_pairs) {
                CS_CUDA(cudaMemcpyAsync(g->d3_od_dst, dst3.data(), n_pairs * 4, cudaMemcpyHostToDevice, g->stream));
                CS_CUDA(cudaMemcpyAsync(g->d3_od_w, w3.data(), n_pairs * 4, cudaMemcpyHostToDevice, g->stream));
            }
            t.od_off = g->d_od_off;
            t.od_dst = g->d3_od_dst;
            t.od_w = g->d3_od_w;
        }
        t.acc_c = g->d_acc;
        t.acc_b = g->d_acc + (size_t)g->n * 5 * D;
        t.counters = g->d_counters;
        t.error = g->d_error;
        t.arena = g->d_arena;
        t.lay = g->lay;
        t.delta = default_delta(g, speed);
        t.bin_scale = p.bin_scale * ((float)CS3_NBINS / (float)CS_NBINS);
        CS_CUDA(cudaEventRecord(g->ev[1], g->stream));
        CS_CUDA(v3_launch(t, g->workers, g->stream));
        launches += 1;
        CS_CUDA(cudaGetLastError());
        int herr3 = 0;
        if (g->opt_kernel == 0) {
            // a walk that stopped increasing (a piece far below f32 resolution) is outside this kernel's contract:
            // serve the call with the arena kernel instead
            CS_CUDA(cudaMemcpyAsync(&herr3, g->d_error, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
            CS_CUDA(cudaStreamSynchronize(g->stream));
        }
        if (herr3 == CS_ERR_ZERO_TIE) {
            CS_CUDA(cudaMemsetAsync(g->d_error, 0, sizeof(int), g->stream));
            CS_CUDA(cudaMemsetAsync(g->d_counters, 0, CS_NCOUNTERS * sizeof(unsigned long long), g->stream));
            CS_CUDA(cudaMemsetAsync(g->d_acc, 0, acc_elems * sizeof(double), g->stream));
            use_v3 = false;
        } else {
            cs_k_epilogue_shortest3<<<nblk, 256, 0, g->stream>>>(t.acc_c, t.acc_b, d_out, g->d3_orig_of_new, g->n, D, closeness,
                                                                 betweenness, accumulate);
            launches += 1;
            CS_CUDA(cudaGetLastError());
            CS_CUDA(cudaEventRecord(g->ev[2], g->stream));
            g->last_kernel = 3;
            return finish_call(g, out, out_on_device, elems, d_out, stats, launches);
        }
    }
    CS_CUDA(cudaEventRecord(g->ev[1], g->stream));
    if (launch_v1(n_sources)) return 1;
    cs_k_epilogue_shortest<<<nblk, 256, 0, g->stream>>>(p.acc_c, p.acc_b, d_out, g->n, D, cw, bw, closeness, betweenness,
                                                        accumulate);
    launches += 1;
    CS_CUDA(cudaGetLastError());
    CS_CUDA(cudaEventRecord(g->ev[2], g->stream));
    return finish_call(g, out, out_on_device, elems, d_out, stats, launches);
}

extern "C" int cs_centrality_shortest(cs_graph* g, int D, const uint32_t* distances, const float* betas,
                                      const uint32_t* seconds, float speed_m_s, float tolerance, int compute_closeness,
                                      int compute_betweenness, uint64_t n_sources, const uint32_t* sources,
                                      const float* source_wt, const uint8_t* eligible, double* out, int out_on_device,
                                      int accumulate, cs_stats* stats) {
    if (!out) return cs_fail("null output");
    for (;;) {
        const int rc = run_shortest(g, D, distances, betas, seconds, speed_m_s, tolerance, compute_closeness,
                                    compute_betweenness, n_sources, sources, source_wt, eligible, out, out_on_device,
                                    accumulate, stats, nullptr, nullptr, nullptr);
        if (!rc || accumulate || !g) return rc;
        const size_t nstates = g->last_kernel == 3 ? (size_t)g->v3_J + 1 : g->n;
        if (!grow_after_overflow(g, nstates, g->lay.rcap)) return rc;
    }
}

// betweenness_od_shortest (centrality.rs:2419-2540).  `out` is the [7][D][node_bound] layout of centrality_shortest with
// only rows 5 (betweenness) and 6 (betweenness_beta) populated.
extern "C" int cs_betweenness_od_shortest(cs_graph* g, int D, const uint32_t* distances, const float* betas,
                                          const uint32_t* seconds, float speed_m_s, float tolerance, uint64_t n_sources,
                                          const uint32_t* sources, const uint64_t* od_off, const uint32_t* od_dst,
                                          const float* od_w, double* out, int out_on_device, cs_stats* stats) {
    if (!out) return cs_fail("null output");
    if (!sources || !od_off || (od_off[n_sources] && (!od_dst || !od_w))) return cs_fail("null OD arrays");
    std::vector<float> ones(std::max<uint64_t>(n_sources, 1), 1.0f);
    for (;;) {
        const int rc = run_shortest(g, D, distances, betas, seconds, speed_m_s, tolerance, 0, 1, n_sources, sources,
                                    ones.data(), nullptr, out, out_on_device, 0, stats, nullptr, nullptr, nullptr, od_off,
                                    od_dst, od_w);
        if (!rc || !g) return rc;
        const size_t nstates = g->last_kernel == 3 ? (size_t)g->v3_J + 1 : g->n;
        if (!grow_after_overflow(g, nstates, g->lay.rcap)) return rc;
    }
}

extern "C" int cs_shortest_search(cs_graph* g, uint32_t src, uint32_t max_seconds, float speed_m_s, float tolerance,
                                  float* agg_seconds, double* sigma, uint32_t* pred_count) {
    if (!g) return cs_fail("null graph");
    if (src >= g->n) return cs_fail("src_idx %u out of range for network with node_bound %u", src, g->n);
    CS_CUDA(cudaSetDevice(g->device));
    float* d_agg = nullptr;
    double* d_sigma = nullptr;
    uint32_t* d_np = nullptr;
    double* d_dummy = nullptr;
    const size_t n = g->n;
    CS_CUDA(cudaMalloc(&d_agg, n * 4));
    CS_CUDA(cudaMalloc(&d_sigma, n * 8));
    CS_CUDA(cudaMalloc(&d_np, n * 4));
    CS_CUDA(cudaMalloc(&d_dummy, n * 7 * 8));
    std::vector<float> inf(n, INFINITY);
    CS_CUDA(cudaMemcpy(d_agg, inf.data(), n * 4, cudaMemcpyHostToDevice));
    CS_CUDA(cudaMemset(d_sigma, 0, n * 8));
    CS_CUDA(cudaMemset(d_np, 0, n * 4));
    uint32_t dist = 0xffffffffu, sec = max_seconds;  // thresholds are irrelevant for the dump; max_seconds is the cutoff
    float beta = 0.f, wt = 1.f;
    // distance threshold: huge, so nothing is filtered; closeness only (cheap), results discarded
    dist = 4000000000u;
    int rc = run_shortest(g, 1, &dist, &beta, &sec, speed_m_s, tolerance, 1, 0, 1, &src, &wt, nullptr, d_dummy, 1, 0,
                          nullptr, d_agg, d_sigma, d_np);
    if (!rc) {
        cudaMemcpy(agg_seconds, d_agg, n * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(sigma, d_sigma, n * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(pred_count, d_np, n * 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_agg);
    cudaFree(d_sigma);
    cudaFree(d_np);
    cudaFree(d_dummy);
    return rc;
}

// ------------------------------------------------------------------------------------------------ segment / simplest
#include "cs_api_more.inl"
#include "cs_api_seg3.inl"
#include "cs_api_tree.inl"
