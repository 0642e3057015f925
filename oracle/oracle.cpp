// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17, C ABI) of cityseer's localized-centrality hot path, written from the reference's
// semantics (NOT a copy of its code) so that the CUDA product path can be checked against it.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this library.
// The product path (cityseer_b200/) never imports, links or calls it.
//
// Parity pin: the Rust reference cannot be compiled in this environment (no cargo/rustc, crates not vendored), so
// this restatement is pinned against the reference's OWN known-answer tests instead (tests/test_oracle_golden.py):
// diamond-graph constants for all three functions (/root/reference/tests/rustalgos/test_centrality.py:451-598),
// NetworkX betweenness / harmonic closeness on mock_graph (:247-341, :652-674), dual routes (:171-234),
// plateau ratio 1.8 (:872-890), tolerance drift (:893-930), threshold pairing tables (tests/rustalgos/test_common.py).
//
// Reference anchors (all relative to /root/reference/rust/src/):
//   edge_travel_seconds / slope_penalty ........ centrality.rs:969-1007
//   NodeDistance + std BinaryHeap order ......... centrality.rs:358-386 (+ Rust std::collections::BinaryHeap sift rules)
//   dijkstra_brandes_shortest (phase 1 + 2) ..... centrality.rs:1344-1494
//   circuit_ranks_from_traversal ................ centrality.rs:470-526
//   closeness loop (shortest) ................... centrality.rs:1733-1780
//   sorted_brandes_state_indices / backprop ..... centrality.rs:777-791, 823-873, 1783-1868
//   dual_node_endpoint_slots / angular .......... centrality.rs:533-566, 577-775, 793-821, 1986-2126
//   dijkstra_tree_shortest / angular / segment .. centrality.rs:1141-1200, 1202-1332, 1523-1611
//   segment_centrality body ..................... centrality.rs:2189-2403
//   adjacency order (petgraph StableGraph: newest edge first per node & direction) — SURVEY.md Appendix C
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <set>
#include <thread>
#include <tuple>
#include <vector>

namespace {

constexpr float TIE_EPSILON = 1e-4f;                       // centrality.rs:28
constexpr float ANGULAR_ROUTE_TIE_BREAK_FACTOR = 1e-6f;    // centrality.rs:24
constexpr float F_INF = std::numeric_limits<float>::infinity();
constexpr int64_t NONE = -1;

struct Edge {
    uint32_t src, dst, edge_idx;
    float length, angle_sum, imp, seconds;
    int32_t key;  // shared_primal_node_key id (-1 = none)
    uint32_t id;  // petgraph edge index
};

struct Graph {
    uint32_t node_bound = 0;
    std::vector<uint8_t> exists, live;
    std::vector<float> weight;
    std::vector<double> z;  // NaN = None
    std::vector<Edge> edges;          // indexed by petgraph edge index (gaps: exists flag below)
    std::vector<uint8_t> edge_exists;
    // adjacency in petgraph iteration order (newest-added first)
    std::vector<std::vector<uint32_t>> out_adj, in_adj;  // edge ids
    bool is_dual = false;
};

// ---- f32 total order (Rust f32::total_cmp) -------------------------------------------------------------------
inline int32_t total_key(float f) {
    int32_t b;
    std::memcpy(&b, &f, 4);
    b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
    return b;
}
inline bool metric_le(float a, float b) { return total_key(a) <= total_key(b); }  // a <= b in total order

// ---- Rust std BinaryHeap<NodeDistance> restated (max-heap on reversed metric) ---------------------------------
// Ord: x <= y  <=>  x.metric >= y.metric (total order).  centrality.rs:363-370
struct HeapItem {
    uint32_t idx;
    float metric;
};
struct RustHeap {
    std::vector<HeapItem> d;
    static bool le(const HeapItem& a, const HeapItem& b) { return metric_le(b.metric, a.metric); }
    void sift_up(size_t start, size_t pos) {
        HeapItem hole = d[pos];
        while (pos > start) {
            size_t parent = (pos - 1) / 2;
            if (le(hole, d[parent])) break;
            d[pos] = d[parent];
            pos = parent;
        }
        d[pos] = hole;
    }
    void sift_down_to_bottom(size_t pos) {
        size_t end = d.size();
        size_t start = pos;
        HeapItem hole = d[pos];
        size_t child = 2 * pos + 1;
        size_t lim = end >= 2 ? end - 2 : 0;
        while (child <= lim && end >= 2) {
            if (le(d[child], d[child + 1])) child += 1;
            d[pos] = d[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (end >= 1 && child == end - 1) {
            d[pos] = d[child];
            pos = child;
        }
        d[pos] = hole;
        sift_up(start, pos);
    }
    void push(uint32_t idx, float metric) {
        size_t old_len = d.size();
        d.push_back({idx, metric});
        sift_up(0, old_len);
    }
    bool pop(HeapItem& out) {
        if (d.empty()) return false;
        HeapItem item = d.back();
        d.pop_back();
        if (!d.empty()) {
            std::swap(item, d[0]);
            sift_down_to_bottom(0);
        }
        out = item;
        return true;
    }
};

// ---- edge seconds (centrality.rs:969-1007) ---------------------------------------------------------------------
inline float slope_penalty(const Graph& g, uint32_t from, uint32_t to, float length_2d) {
    if (length_2d <= 0.0f) return 1.0f;
    double zf = g.z[from], zt = g.z[to];
    if (std::isnan(zf) || std::isnan(zt)) return 1.0f;
    float slope = (float)(zt - zf) / length_2d;
    const float FLAT_FACTOR = 0.839457f;
    float slope_factor = std::exp(-3.5f * std::fabs(slope + 0.05f));
    return FLAT_FACTOR / slope_factor;
}
inline float edge_travel_seconds(const Graph& g, uint32_t from, uint32_t to, const Edge& e, float speed, bool use_imp) {
    if (!std::isnan(e.seconds)) return e.seconds;
    float sp = slope_penalty(g, from, to, e.length);
    float imp = use_imp ? e.imp : 1.0f;
    return (e.length * imp * sp) / speed;
}

// ---- counters shared with the GPU side -----------------------------------------------------------
struct Counters {
    std::atomic<uint64_t> sources{0}, settled{0}, edge_iters{0}, sum_ri{0}, sum_ci{0}, key_ties{0}, multi_pred{0};
};

struct BState {
    bool visited = false;
    std::vector<uint32_t> preds;
    double sigma = 0.0;
    uint32_t node_idx = 0;
    float route_cost = F_INF;
    float agg_seconds = F_INF;
};
struct Traversal {
    std::vector<uint32_t> visited_state_indices, reached_node_indices;
    std::vector<BState> state;
    std::vector<float> best_route_cost, best_agg_seconds;
};

inline bool contains(const std::vector<uint32_t>& v, uint32_t x) { return std::find(v.begin(), v.end(), x) != v.end(); }

void reached_nodes(Traversal& t) {
    for (uint32_t i = 0; i < t.best_route_cost.size(); ++i)
        if (std::isfinite(t.best_route_cost[i])) t.reached_node_indices.push_back(i);
}

// Workspace of the "optimised CPU" variant: the SAME arithmetic and visiting order as
// the faithful restatement, but the Theta(N) per-source allocations of the reference (centrality.rs:1356-1361, :835-836,
// :1785-1792) are replaced by per-thread arrays that are reset through the list of touched nodes.
struct SparseWork {
    Traversal t;
    std::vector<uint32_t> touched;
    std::vector<size_t> visit_pos;
    std::vector<double> seed, seedb, delta, deltab;
    void ensure(uint32_t n) {
        if (t.state.size() == n) return;
        t.state.assign(n, BState());
        for (uint32_t i = 0; i < n; ++i) t.state[i].node_idx = i;
        t.best_route_cost.assign(n, F_INF);
        t.best_agg_seconds.assign(n, F_INF);
        visit_pos.assign(n, SIZE_MAX);
        seed.assign(n, 0.0);
        seedb.assign(n, 0.0);
        delta.assign(n, 0.0);
        deltab.assign(n, 0.0);
    }
    void reset() {  // back to the state `ensure` leaves, touching only what the last source touched
        for (uint32_t i : touched) {
            BState& b = t.state[i];
            b.visited = false;
            b.preds.clear();
            b.sigma = 0.0;
            b.route_cost = F_INF;
            b.agg_seconds = F_INF;
            t.best_route_cost[i] = F_INF;
            t.best_agg_seconds[i] = F_INF;
        }
        touched.clear();
        t.visited_state_indices.clear();
        t.reached_node_indices.clear();
    }
};

// centrality.rs:1344-1494.  `w` != NULL: sparse-reset variant, `t` is w->t (sized and clean).
void brandes_shortest(const Graph& g, uint32_t src, uint32_t max_seconds, float speed, float tol, Traversal& t, Counters* c,
                      SparseWork* w = nullptr) {
    uint32_t n = g.node_bound;
    if (!w) {
        t.state.assign(n, BState());
        for (uint32_t i = 0; i < n; ++i) t.state[i].node_idx = i;
        t.best_route_cost.assign(n, F_INF);
        t.best_agg_seconds.assign(n, F_INF);
    } else {
        w->touched.push_back(src);
    }
    auto& st = t.state;
    st[src].sigma = 1.0;
    st[src].route_cost = 0.0f;
    st[src].agg_seconds = 0.0f;
    t.best_route_cost[src] = 0.0f;
    t.best_agg_seconds[src] = 0.0f;
    RustHeap heap;
    heap.push(src, 0.0f);
    HeapItem it;
    float last_key = -1.0f;
    uint64_t edge_iters = 0, key_ties = 0;
    while (heap.pop(it)) {
        uint32_t si = it.idx;
        if (st[si].visited) continue;
        st[si].visited = true;
        t.visited_state_indices.push_back(si);
        if (it.metric == last_key) key_ties++;
        last_key = it.metric;
        uint32_t cur = st[si].node_idx;
        for (uint32_t eid : g.in_adj[cur]) {
            edge_iters++;
            const Edge& e = g.edges[eid];
            uint32_t nb = e.src;
            if (nb == cur) continue;
            float es = edge_travel_seconds(g, nb, cur, e, speed, true);
            float cs = st[si].agg_seconds + es;
            if (cs > (float)max_seconds) continue;
            if (st[nb].visited) continue;
            float cr = cs * speed;
            bool improved = cs < st[nb].agg_seconds;
            bool tied = cs <= st[nb].agg_seconds * (1.0f + TIE_EPSILON);
            if (improved) {
                if (w && st[nb].agg_seconds == F_INF) w->touched.push_back(nb);  // first discovery
                if (cs < st[nb].agg_seconds * (1.0f - TIE_EPSILON)) {
                    st[nb].preds.clear();
                    st[nb].sigma = st[si].sigma;
                } else {
                    st[nb].sigma += st[si].sigma;
                }
                st[nb].route_cost = cr;
                st[nb].agg_seconds = cs;
                st[nb].preds.push_back(si);
                heap.push(nb, cr);
            } else if (tied && !contains(st[nb].preds, si)) {
                if (cs < st[nb].agg_seconds) st[nb].agg_seconds = cs;
                st[nb].preds.push_back(si);
                st[nb].sigma += st[si].sigma;
            }
            t.best_route_cost[nb] = st[nb].route_cost;
            t.best_agg_seconds[nb] = st[nb].agg_seconds;
        }
    }
    if (tol > TIE_EPSILON) {
        std::vector<size_t> visit_pos_own;
        if (!w) visit_pos_own.assign(n, SIZE_MAX);
        std::vector<size_t>& visit_pos = w ? w->visit_pos : visit_pos_own;
        for (size_t p = 0; p < t.visited_state_indices.size(); ++p) visit_pos[t.visited_state_indices[p]] = p;
        for (uint32_t idx : t.visited_state_indices) {
            st[idx].preds.clear();
            st[idx].sigma = 0.0;
        }
        st[src].sigma = 1.0;
        for (size_t pos = 0; pos < t.visited_state_indices.size(); ++pos) {
            uint32_t u = t.visited_state_indices[pos];
            for (uint32_t eid : g.in_adj[st[u].node_idx]) {
                const Edge& e = g.edges[eid];
                uint32_t v = e.src;
                if (v == u) continue;
                if (visit_pos[v] <= pos) continue;
                float es = edge_travel_seconds(g, v, st[u].node_idx, e, speed, true);
                float ps = st[u].agg_seconds + es;
                if (ps <= st[v].agg_seconds * (1.0f + tol) && !contains(st[v].preds, u)) {
                    st[v].preds.push_back(u);
                    st[v].sigma += st[u].sigma;
                }
            }
        }
        if (w)
            for (uint32_t idx : t.visited_state_indices) visit_pos[idx] = SIZE_MAX;
    }
    if (!w) {
        reached_nodes(t);
    } else {
        // the reference scans 0..n for finite costs (:1485): the touched nodes in index order are that list
        t.reached_node_indices = w->touched;
        std::sort(t.reached_node_indices.begin(), t.reached_node_indices.end());
    }
    if (c) {
        c->settled += t.visited_state_indices.size();
        c->edge_iters += edge_iters;
        c->key_ties += key_ties;
        uint64_t mp = 0;
        for (uint32_t idx : t.visited_state_indices)
            if (st[idx].preds.size() > 1) mp++;
        c->multi_pred += mp;
    }
}

// centrality.rs:470-526
std::vector<float> circuit_ranks(const Graph& g, const Traversal& t, const std::vector<uint32_t>& distances) {
    size_t dn = distances.size();
    std::vector<size_t> node_counts(dn, 0), edge_counts(dn, 0);
    for (uint32_t ni : t.reached_node_indices) {
        float cost = t.best_route_cost[ni];
        for (size_t i = 0; i < dn; ++i)
            if (cost <= (float)distances[i]) node_counts[i]++;
    }
    std::set<std::tuple<uint32_t, uint32_t, uint32_t>> seen;
    for (uint32_t ni : t.reached_node_indices) {
        for (uint32_t eid : g.out_adj[ni]) {
            const Edge& e = g.edges[eid];
            uint32_t nb = e.dst;
            if (nb == ni) continue;
            if (!std::isfinite(t.best_route_cost[nb])) continue;
            auto key = std::make_tuple(std::min(ni, nb), std::max(ni, nb), e.edge_idx);
            if (!seen.insert(key).second) continue;
            float ec = std::max(t.best_route_cost[ni], t.best_route_cost[nb]);
            for (size_t i = 0; i < dn; ++i)
                if (ec <= (float)distances[i]) edge_counts[i]++;
        }
    }
    std::vector<float> out(dn);
    for (size_t i = 0; i < dn; ++i) {
        if (node_counts[i] == 0)
            out[i] = 0.0f;
        else
            out[i] = (float)std::max<int64_t>((int64_t)edge_counts[i] - (int64_t)node_counts[i] + 1, 0);
    }
    return out;
}

// centrality.rs:777-791
std::vector<uint32_t> sorted_states(const Traversal& t) {
    std::vector<uint32_t> s;
    for (uint32_t idx : t.visited_state_indices)
        if (t.state[idx].sigma > 0.0) s.push_back(idx);
    std::stable_sort(s.begin(), s.end(), [&](uint32_t a, uint32_t b) { return t.state[a].route_cost > t.state[b].route_cost; });
    return s;
}

// centrality.rs:823-873
template <class FInc, class FCred>
void backprop(const Traversal& t, const std::vector<uint32_t>& sorted, uint32_t src_node, const std::vector<double>& seed,
              const std::vector<double>& seed_beta, FInc include, FCred on_credit, SparseWork* w = nullptr) {
    std::vector<double> delta_own, delta_beta_own;
    if (!w) {
        delta_own.assign(t.state.size(), 0.0);
        delta_beta_own.assign(t.state.size(), 0.0);
    }
    std::vector<double>& delta = w ? w->delta : delta_own;  // clean on entry; the caller resets them over `touched`
    std::vector<double>& delta_beta = w ? w->deltab : delta_beta_own;
    for (uint32_t si : sorted) {
        const BState& s = t.state[si];
        if (!include(s)) continue;
        double sigma_w = s.sigma;
        if (sigma_w == 0.0) continue;
        double dep = seed[si] + delta[si];
        double depb = seed_beta[si] + delta_beta[si];
        if (dep == 0.0 && depb == 0.0) continue;
        for (uint32_t p : s.preds) {
            double sigma_v = t.state[p].sigma;
            if (sigma_v == 0.0) continue;
            double f = sigma_v / sigma_w;
            delta[p] += f * dep;
            delta_beta[p] += f * depb;
        }
        if (s.node_idx == src_node) continue;
        double credit = dep - seed[si];
        double creditb = depb - seed_beta[si];
        if (credit > 0.0 || creditb > 0.0) on_credit(s.node_idx, std::max(credit, 0.0), std::max(creditb, 0.0));
    }
}

inline void atomic_add(double* p, double v) {
    auto* a = reinterpret_cast<std::atomic<double>*>(p);
    double old = a->load(std::memory_order_relaxed);
    while (!a->compare_exchange_weak(old, old + v, std::memory_order_relaxed)) {
    }
}

template <class F>
void par_for(uint64_t n, int n_threads, F f) {
    if (n_threads <= 1 || n < 2) {
        for (uint64_t i = 0; i < n; ++i) f(i);
        return;
    }
    std::atomic<uint64_t> next{0};
    std::vector<std::thread> th;
    for (int k = 0; k < n_threads; ++k)
        th.emplace_back([&] {
            for (;;) {
                uint64_t i = next.fetch_add(1);
                if (i >= n) break;
                f(i);
            }
        });
    for (auto& x : th) x.join();
}

// ---- angular -----------------------------------------------------------------------------------------------------
// centrality.rs:533-566: slot index of each shared key at each dual node; returns false on >2 keys or a missing key.
int endpoint_slots(const Graph& g, std::vector<std::array<int32_t, 2>>& slots) {
    slots.assign(g.node_bound, {-1, -1});
    for (size_t eid = 0; eid < g.edges.size(); ++eid) {
        if (!g.edge_exists[eid]) continue;
        const Edge& e = g.edges[eid];
        if (e.key < 0) return 1;  // missing shared_primal_node_key
        for (uint32_t nd : {e.src, e.dst}) {
            auto& s = slots[nd];
            if (s[0] == e.key || s[1] == e.key) continue;
            if (s[0] < 0)
                s[0] = e.key;
            else if (s[1] < 0)
                s[1] = e.key;
            else
                return 2;  // more than two primal endpoints
        }
    }
    return 0;
}
inline int slot_pos(const std::array<int32_t, 2>& s, int32_t key) { return s[0] == key ? 0 : (s[1] == key ? 1 : -1); }

// centrality.rs:577-775
void brandes_angular(const Graph& g, uint32_t src, uint32_t max_seconds, float speed, float tol,
                     const std::vector<std::array<int32_t, 2>>& slots, Traversal& t, Counters* c) {
    uint32_t n = g.node_bound;
    uint32_t sc = n * 2;
    t.state.assign(sc, BState());
    for (uint32_t i = 0; i < sc; ++i) t.state[i].node_idx = i / 2;
    t.best_route_cost.assign(n, F_INF);
    t.best_agg_seconds.assign(n, F_INF);
    auto& st = t.state;
    RustHeap heap;
    t.best_route_cost[src] = 0.0f;
    t.best_agg_seconds[src] = 0.0f;
    for (uint32_t slot = 0; slot < 2; ++slot) {
        uint32_t s = src * 2 + slot;
        st[s].sigma = 1.0;
        st[s].route_cost = 0.0f;
        st[s].agg_seconds = 0.0f;
        heap.push(s, 0.0f);
    }
    HeapItem it;
    uint64_t edge_iters = 0;
    while (heap.pop(it)) {
        uint32_t si = it.idx;
        if (st[si].visited) continue;
        st[si].visited = true;
        t.visited_state_indices.push_back(si);
        uint32_t cur = st[si].node_idx;
        int entry = si % 2;
        for (uint32_t eid : g.out_adj[cur]) {
            edge_iters++;
            const Edge& e = g.edges[eid];
            uint32_t nx = e.dst;
            int cslot = slot_pos(slots[cur], e.key);
            if (cslot != 1 - entry) continue;
            int nslot = slot_pos(slots[nx], e.key);
            uint32_t ns = nx * 2 + (uint32_t)nslot;
            float es = edge_travel_seconds(g, cur, nx, e, speed, false);
            float cs = st[si].agg_seconds + es;
            if (cs > (float)max_seconds) continue;
            if (st[ns].visited) continue;
            float cr = st[si].route_cost + e.angle_sum + (ANGULAR_ROUTE_TIE_BREAK_FACTOR * e.length);
            float cur_cost = st[ns].route_cost;
            bool improved = cr < cur_cost;
            bool tied = cr <= cur_cost * (1.0f + TIE_EPSILON);
            if (improved) {
                if (cr < cur_cost * (1.0f - TIE_EPSILON)) {
                    st[ns].preds.clear();
                    st[ns].sigma = st[si].sigma;
                } else {
                    st[ns].sigma += st[si].sigma;
                }
                st[ns].route_cost = cr;
                st[ns].agg_seconds = cs;
                st[ns].preds.push_back(si);
                heap.push(ns, cr);
            } else if (tied && !contains(st[ns].preds, si)) {
                if (cs < st[ns].agg_seconds) st[ns].agg_seconds = cs;
                st[ns].preds.push_back(si);
                st[ns].sigma += st[si].sigma;
            }
            uint32_t nn = st[ns].node_idx;
            float best = t.best_route_cost[nn];
            if (cr < best * (1.0f - TIE_EPSILON)) {
                t.best_route_cost[nn] = cr;
                t.best_agg_seconds[nn] = cs;
            } else if (cr <= best * (1.0f + TIE_EPSILON)) {
                t.best_agg_seconds[nn] = std::min(t.best_agg_seconds[nn], cs);
            }
        }
    }
    if (tol > TIE_EPSILON) {
        std::vector<size_t> visit_pos(sc, SIZE_MAX);
        for (size_t p = 0; p < t.visited_state_indices.size(); ++p) visit_pos[t.visited_state_indices[p]] = p;
        for (uint32_t idx : t.visited_state_indices) {
            st[idx].preds.clear();
            st[idx].sigma = 0.0;
        }
        for (uint32_t slot = 0; slot < 2; ++slot) st[src * 2 + slot].sigma = 1.0;
        for (size_t pos = 0; pos < t.visited_state_indices.size(); ++pos) {
            uint32_t us = t.visited_state_indices[pos];
            uint32_t un = st[us].node_idx;
            int uentry = us % 2;
            for (uint32_t eid : g.out_adj[un]) {
                const Edge& e = g.edges[eid];
                uint32_t nx = e.dst;
                int uslot = slot_pos(slots[un], e.key);
                if (uslot != 1 - uentry) continue;
                int nslot = slot_pos(slots[nx], e.key);
                uint32_t vs = nx * 2 + (uint32_t)nslot;
                if (visit_pos[vs] <= pos) continue;
                float cr = st[us].route_cost + e.angle_sum + (ANGULAR_ROUTE_TIE_BREAK_FACTOR * e.length);
                if (cr <= st[vs].route_cost * (1.0f + tol) && !contains(st[vs].preds, us)) {
                    st[vs].preds.push_back(us);
                    st[vs].sigma += st[us].sigma;
                }
            }
        }
    }
    reached_nodes(t);
    if (c) {
        c->settled += t.visited_state_indices.size();
        c->edge_iters += edge_iters;
    }
}

// ---- single-predecessor tree searches ------------------------------------------------------------------------------
struct NodeVisit {
    bool visited = false, discovered = false;
    int64_t pred = NONE;
    float short_dist = F_INF, simpl_dist = F_INF;
    int64_t origin_seg = NONE, last_seg = NONE;
    float agg_seconds = F_INF;
};
struct EdgeVisit {
    bool visited = false;
    int64_t start = NONE, end = NONE, edge_idx = NONE;
};

// centrality.rs:1141-1200
void tree_shortest(const Graph& g, uint32_t src, uint32_t max_seconds, float speed, std::vector<uint32_t>& visited,
                   std::vector<NodeVisit>& tm) {
    tm.assign(g.node_bound, NodeVisit());
    tm[src].agg_seconds = 0.0f;
    tm[src].discovered = true;
    tm[src].short_dist = 0.0f;
    RustHeap heap;
    heap.push(src, 0.0f);
    HeapItem it;
    while (heap.pop(it)) {
        uint32_t ni = it.idx;
        if (tm[ni].visited) continue;
        tm[ni].visited = true;
        visited.push_back(ni);
        for (uint32_t eid : g.in_adj[ni]) {
            const Edge& e = g.edges[eid];
            uint32_t nb = e.src;
            if (nb == ni || tm[nb].visited) continue;
            if (tm[ni].pred != NONE && (int64_t)nb == tm[ni].pred) continue;
            float es = edge_travel_seconds(g, nb, ni, e, speed, true);
            float ts = tm[ni].agg_seconds + es;
            if (ts > (float)max_seconds) continue;
            if (ts < tm[nb].agg_seconds) {
                tm[nb].short_dist = ts * speed;
                tm[nb].agg_seconds = ts;
                tm[nb].pred = ni;
                tm[nb].discovered = true;
                heap.push(nb, ts);
            }
        }
    }
}

// centrality.rs:1202-1332
void tree_angular(const Graph& g, uint32_t src, uint32_t max_seconds, float speed,
                  const std::vector<std::array<int32_t, 2>>& slots, std::vector<uint32_t>& visited, std::vector<NodeVisit>& tm) {
    struct AState {
        bool visited = false;
        int64_t pred = NONE;
        float route_metric = F_INF, simpl = F_INF, agg = F_INF;
    };
    uint32_t n = g.node_bound;
    std::vector<AState> st(n * 2);
    tm.assign(n, NodeVisit());
    visited.push_back(src);
    std::vector<uint8_t> reached(n, 0);
    reached[src] = 1;
    tm[src].discovered = true;
    tm[src].visited = true;
    tm[src].simpl_dist = 0.0f;
    tm[src].agg_seconds = 0.0f;
    RustHeap heap;
    for (uint32_t slot = 0; slot < 2; ++slot) {
        uint32_t s = src * 2 + slot;
        st[s].route_metric = 0.0f;
        st[s].simpl = 0.0f;
        st[s].agg = 0.0f;
        heap.push(s, 0.0f);
    }
    HeapItem it;
    while (heap.pop(it)) {
        uint32_t si = it.idx;
        if (st[si].visited) continue;
        st[si].visited = true;
        uint32_t cur = si / 2;
        int entry = si % 2;
        for (uint32_t eid : g.out_adj[cur]) {
            const Edge& e = g.edges[eid];
            uint32_t nx = e.dst;
            int cslot = slot_pos(slots[cur], e.key);
            if (cslot != 1 - entry) continue;
            int nslot = slot_pos(slots[nx], e.key);
            uint32_t ns = nx * 2 + (uint32_t)nslot;
            float es = edge_travel_seconds(g, cur, nx, e, speed, false);
            float cs = st[si].agg + es;
            if (cs > (float)max_seconds) continue;
            float csimpl = st[si].simpl + e.angle_sum;
            float cmetric = csimpl + (ANGULAR_ROUTE_TIE_BREAK_FACTOR * e.length);
            bool improved = cmetric + TIE_EPSILON < st[ns].route_metric;
            bool tied = std::fabs(cmetric - st[ns].route_metric) <= TIE_EPSILON;
            if (improved || (tied && cs < st[ns].agg && si != ns)) {
                st[ns].route_metric = cmetric;
                st[ns].simpl = csimpl;
                st[ns].agg = cs;
                st[ns].pred = si;
                heap.push(ns, cmetric);
                NodeVisit& nv = tm[nx];
                if (!reached[nx]) {
                    reached[nx] = 1;
                    visited.push_back(nx);
                }
                bool node_improved = csimpl + TIE_EPSILON < nv.simpl_dist;
                bool node_tied = std::fabs(csimpl - nv.simpl_dist) <= TIE_EPSILON;
                if (!nv.discovered || node_improved || (node_tied && cs < nv.agg_seconds)) {
                    nv.discovered = true;
                    nv.visited = true;
                    nv.simpl_dist = csimpl;
                    nv.agg_seconds = cs;
                    nv.pred = cur;
                }
            }
        }
    }
}

// centrality.rs:1523-1611
void tree_segment(const Graph& g, uint32_t src, uint32_t max_seconds, float speed, std::vector<uint32_t>& visited_nodes,
                  std::vector<uint32_t>& visited_edges, std::vector<NodeVisit>& tm, std::vector<EdgeVisit>& em, Counters* c) {
    tm.assign(g.node_bound, NodeVisit());
    em.assign(g.edges.size(), EdgeVisit());
    tm[src].short_dist = 0.0f;
    tm[src].agg_seconds = 0.0f;
    tm[src].discovered = true;
    RustHeap heap;
    heap.push(src, 0.0f);
    HeapItem it;
    uint64_t edge_iters = 0;
    while (heap.pop(it)) {
        uint32_t ni = it.idx;
        if (tm[ni].visited) continue;
        tm[ni].visited = true;
        visited_nodes.push_back(ni);
        for (uint32_t eid : g.in_adj[ni]) {
            edge_iters++;
            const Edge& e = g.edges[eid];
            uint32_t nb = e.src;
            if (nb == ni) {
                visited_edges.push_back(eid);
                em[eid] = {true, (int64_t)ni, (int64_t)nb, (int64_t)e.edge_idx};
                continue;
            }
            if (tm[nb].visited) continue;
            visited_edges.push_back(eid);
            em[eid] = {true, (int64_t)ni, (int64_t)nb, (int64_t)e.edge_idx};
            float es = edge_travel_seconds(g, nb, ni, e, speed, true);
            float ts = tm[ni].agg_seconds + es;
            if (ts > (float)max_seconds) continue;
            if (ts < tm[nb].agg_seconds) {
                int64_t origin = (ni == src) ? (int64_t)eid : tm[ni].origin_seg;
                tm[nb].short_dist = ts * speed;
                tm[nb].agg_seconds = ts;
                tm[nb].pred = ni;
                tm[nb].origin_seg = origin;
                tm[nb].last_seg = eid;
                tm[nb].discovered = true;
                heap.push(nb, ts);
            }
        }
    }
    if (c) {
        c->settled += visited_nodes.size();
        c->edge_iters += edge_iters;
    }
}

// graph.rs:1278-1292: first edge start->end (petgraph edges_connecting order = out-list order) with payload edge_idx
const Edge* find_edge(const Graph& g, uint32_t start, uint32_t end, uint32_t edge_idx) {
    for (uint32_t eid : g.out_adj[start]) {
        const Edge& e = g.edges[eid];
        if (e.dst == end && e.edge_idx == edge_idx) return &e;
    }
    return nullptr;
}

}  // namespace

extern "C" {

struct orc_graph {
    Graph g;
};

orc_graph* orc_graph_create(uint32_t node_bound, const uint8_t* node_exists, const uint8_t* live, const float* weight,
                            const double* z, uint64_t edge_bound, const uint8_t* edge_exists, const uint32_t* src,
                            const uint32_t* dst, const uint32_t* edge_idx, const float* length, const float* angle_sum,
                            const float* imp, const float* seconds, const int32_t* shared_key, const uint64_t* stamp,
                            int is_dual) {
    auto* h = new orc_graph();
    Graph& g = h->g;
    g.node_bound = node_bound;
    g.exists.assign(node_exists, node_exists + node_bound);
    g.live.assign(live, live + node_bound);
    g.weight.assign(weight, weight + node_bound);
    g.z.assign(z, z + node_bound);
    g.edges.resize(edge_bound);
    g.edge_exists.assign(edge_exists, edge_exists + edge_bound);
    g.out_adj.resize(node_bound);
    g.in_adj.resize(node_bound);
    g.is_dual = is_dual != 0;
    std::vector<uint32_t> order;
    for (uint64_t i = 0; i < edge_bound; ++i) {
        g.edges[i] = {src[i], dst[i], edge_idx[i], length[i], angle_sum[i], imp[i], seconds[i], shared_key ? shared_key[i] : -1, (uint32_t)i};
        if (edge_exists[i]) order.push_back((uint32_t)i);
    }
    // newest-first adjacency (petgraph add_edge pushes at the head of both lists)
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return stamp[a] > stamp[b]; });
    for (uint32_t eid : order) {
        g.out_adj[g.edges[eid].src].push_back(eid);
        g.in_adj[g.edges[eid].dst].push_back(eid);
    }
    return h;
}
void orc_graph_destroy(orc_graph* h) { delete h; }

// counters_out[7]: sources, settled, edge_iters, sum_ri, sum_ci, key_ties, multi_pred
// out: [7][D][node_bound] = density, farness, cycles, harmonic, beta, betweenness, betweenness_beta
// betweenness_od_shortest (centrality.rs:2419-2540): the Brandes pass of centrality_shortest with seeds only at the OD
// destinations of each origin (weight w, beta seed w * exp(-beta * cost)), credits not scaled by any source weight.
// od_off[k] .. od_off[k + 1] index the destinations / weights of sources[k]; out is [2][D][node_bound].
int orc_betweenness_od(const orc_graph* h, int D, const uint32_t* distances, const float* betas, const uint32_t* seconds,
                       float speed, float tol, uint64_t n_sources, const uint32_t* sources, const uint64_t* od_off,
                       const uint32_t* od_dst, const float* od_w, double* out, int n_threads) {
    const Graph& g = h->g;
    size_t nb = g.node_bound;
    uint32_t max_sec = *std::max_element(seconds, seconds + D);
    auto M = [&](int m, int i, size_t node) -> double* { return out + ((size_t)m * D + i) * nb + node; };
    par_for(n_sources, n_threads, [&](uint64_t k) {
        uint32_t src = sources[k];
        Traversal t;
        brandes_shortest(g, src, max_sec, speed, tol, t, nullptr);
        std::vector<uint32_t> sorted = sorted_states(t);
        std::vector<double> seed(t.state.size()), seedb(t.state.size());
        for (int i = 0; i < D; ++i) {
            float thr = (float)distances[i];
            double beta = (double)betas[i];
            std::fill(seed.begin(), seed.end(), 0.0);
            std::fill(seedb.begin(), seedb.end(), 0.0);
            for (uint64_t j = od_off[k]; j < od_off[k + 1]; ++j) {
                uint32_t dest = od_dst[j];
                if (t.best_route_cost[dest] > thr) continue;
                seed[dest] += (double)od_w[j];
                seedb[dest] += (double)od_w[j] * std::exp(-beta * (double)t.best_route_cost[dest]);
            }
            backprop(
                t, sorted, src, seed, seedb, [&](const BState& s) { return s.route_cost <= thr; },
                [&](uint32_t node, double credit, double creditb) {
                    if (credit > 0.0) atomic_add(M(0, i, node), credit);
                    if (creditb > 0.0) atomic_add(M(1, i, node), creditb);
                });
        }
    });
    return 0;
}

static int centrality_shortest_impl(const orc_graph* h, int D, const uint32_t* distances, const float* betas,
                                    const uint32_t* seconds, float speed, float tol, int closeness, int betweenness,
                                    uint64_t n_sources, const uint32_t* sources, const float* source_wt,
                                    const uint8_t* eligible, double* out, uint64_t* counters_out, uint64_t* reach_totals,
                                    int n_threads, bool sparse) {
    const Graph& g = h->g;
    size_t nb = g.node_bound;
    std::vector<uint32_t> dist(distances, distances + D);
    uint32_t max_sec = *std::max_element(seconds, seconds + D);
    Counters cnt;
    std::vector<std::atomic<uint64_t>> reach(D);
    for (auto& r : reach) r = 0;
    auto M = [&](int m, int i, size_t node) -> double* { return out + ((size_t)m * D + i) * nb + node; };
    par_for(n_sources, n_threads, [&](uint64_t k) {
        uint32_t src = sources[k];
        float wt = source_wt[k];
        static thread_local SparseWork tls_work;
        SparseWork* w = sparse ? &tls_work : nullptr;
        Traversal t_own;
        if (w) w->ensure(g.node_bound);
        Traversal& t = w ? w->t : t_own;
        brandes_shortest(g, src, max_sec, speed, tol, t, &cnt, w);
        cnt.sources++;
        float cycles_wt = wt / g.weight[src];
        if (closeness) {
            std::vector<float> ranks = circuit_ranks(g, t, dist);
            uint64_t sri = 0;
            for (uint32_t to : t.reached_node_indices) {
                if (to == src) continue;
                if (!std::isfinite(t.best_agg_seconds[to])) continue;
                float cost = t.best_route_cost[to];
                for (int i = 0; i < D; ++i) {
                    if (cost <= (float)dist[i]) {
                        sri++;
                        reach[i]++;
                        atomic_add(M(0, i, to), (double)wt);
                        atomic_add(M(1, i, to), (double)(cost * wt));
                        atomic_add(M(2, i, to), (double)(ranks[i] * cycles_wt));
                        atomic_add(M(3, i, to), (double)((1.0f / cost) * wt));
                        atomic_add(M(4, i, to), (double)(std::exp(-betas[i] * cost) * wt));
                    }
                }
            }
            cnt.sum_ri += sri;
        }
        if (betweenness) {
            std::vector<uint32_t> sorted = sorted_states(t);
            std::vector<double> seed_own, seedb_own;
            if (!w) {
                seed_own.resize(t.state.size());
                seedb_own.resize(t.state.size());
            }
            std::vector<double>& seed = w ? w->seed : seed_own;
            std::vector<double>& seedb = w ? w->seedb : seedb_own;
            uint64_t sci = 0;
            for (int i = 0; i < D; ++i) {
                float thr = (float)dist[i];
                double beta = (double)betas[i];
                if (!w) {
                    std::fill(seed.begin(), seed.end(), 0.0);
                    std::fill(seedb.begin(), seedb.end(), 0.0);
                } else {
                    for (uint32_t q : w->touched) seed[q] = seedb[q] = w->delta[q] = w->deltab[q] = 0.0;
                }
                for (uint32_t to : t.reached_node_indices) {
                    if (to == src) continue;
                    if (t.best_route_cost[to] > thr) continue;
                    double pc = eligible[to] ? 0.5 : 1.0;
                    double pb = pc * std::exp(-beta * (double)t.best_route_cost[to]);
                    seed[to] += pc;
                    seedb[to] += pb;
                }
                backprop(
                    t, sorted, src, seed, seedb, [&](const BState& s) { return s.route_cost <= thr; },
                    [&](uint32_t node, double credit, double creditb) {
                        sci++;
                        if (credit > 0.0) atomic_add(M(5, i, node), credit * (double)wt);
                        if (creditb > 0.0) atomic_add(M(6, i, node), creditb * (double)wt);
                    },
                    w);
            }
            if (w)
                for (uint32_t q : w->touched) w->seed[q] = w->seedb[q] = w->delta[q] = w->deltab[q] = 0.0;
            cnt.sum_ci += sci;
        }
        if (w) w->reset();
    });
    if (counters_out) {
        counters_out[0] = cnt.sources;
        counters_out[1] = cnt.settled;
        counters_out[2] = cnt.edge_iters;
        counters_out[3] = cnt.sum_ri;
        counters_out[4] = cnt.sum_ci;
        counters_out[5] = cnt.key_ties;
        counters_out[6] = cnt.multi_pred;
    }
    if (reach_totals)
        for (int i = 0; i < D; ++i) reach_totals[i] = reach[i];
    return 0;
}

int orc_centrality_shortest(const orc_graph* h, int D, const uint32_t* distances, const float* betas, const uint32_t* seconds,
                            float speed, float tol, int closeness, int betweenness, uint64_t n_sources,
                            const uint32_t* sources, const float* source_wt, const uint8_t* eligible, double* out,
                            uint64_t* counters_out, uint64_t* reach_totals, int n_threads) {
    return centrality_shortest_impl(h, D, distances, betas, seconds, speed, tol, closeness, betweenness, n_sources, sources,
                                    source_wt, eligible, out, counters_out, reach_totals, n_threads, false);
}

// "Optimised CPU" variant: identical arithmetic and visiting order, touched-list resets instead of the reference's
// Theta(N) per-source allocations.  Reported beside the faithful restatement so the GPU is not only compared with the
// slow formulation; tests/test_oracle_golden.py asserts it equals the faithful port bit for bit.
int orc_centrality_shortest_opt(const orc_graph* h, int D, const uint32_t* distances, const float* betas,
                                const uint32_t* seconds, float speed, float tol, int closeness, int betweenness,
                                uint64_t n_sources, const uint32_t* sources, const float* source_wt, const uint8_t* eligible,
                                double* out, uint64_t* counters_out, uint64_t* reach_totals, int n_threads) {
    return centrality_shortest_impl(h, D, distances, betas, seconds, speed, tol, closeness, betweenness, n_sources, sources,
                                    source_wt, eligible, out, counters_out, reach_totals, n_threads, true);
}

// out: [4][D][node_bound] = density, farness, harmonic, betweenness.  Returns 1/2 on invalid dual metadata.
int orc_centrality_simplest(const orc_graph* h, int D, const uint32_t* distances, const uint32_t* seconds, float speed,
                            float tol, float angular_scaling_unit, float farness_scaling_offset, int closeness,
                            int betweenness, uint64_t n_sources, const uint32_t* sources, const float* source_wt,
                            const uint8_t* eligible, double* out, uint64_t* counters_out, uint64_t* reach_totals,
                            int n_threads) {
    (void)distances;
    const Graph& g = h->g;
    size_t nb = g.node_bound;
    std::vector<std::array<int32_t, 2>> slots;
    int rc = endpoint_slots(g, slots);
    if (rc) return rc;
    uint32_t max_sec = *std::max_element(seconds, seconds + D);
    Counters cnt;
    std::vector<std::atomic<uint64_t>> reach(D);
    for (auto& r : reach) r = 0;
    auto M = [&](int m, int i, size_t node) -> double* { return out + ((size_t)m * D + i) * nb + node; };
    par_for(n_sources, n_threads, [&](uint64_t k) {
        uint32_t src = sources[k];
        float wt = source_wt[k];
        Traversal t;
        brandes_angular(g, src, max_sec, speed, tol, slots, t, &cnt);
        cnt.sources++;
        if (closeness) {
            uint64_t sri = 0;
            for (uint32_t to : t.reached_node_indices) {
                if (to == src) continue;
                float simpl = t.best_route_cost[to];
                float bsec = t.best_agg_seconds[to];
                if (!std::isfinite(simpl) || !std::isfinite(bsec)) continue;
                for (int i = 0; i < D; ++i) {
                    if (bsec <= (float)seconds[i]) {
                        sri++;
                        reach[i]++;
                        atomic_add(M(0, i, to), (double)wt);
                        float far_ang = farness_scaling_offset + (simpl / angular_scaling_unit);
                        atomic_add(M(1, i, to), (double)(far_ang * wt));
                        float harm_ang = 1.0f + (simpl / angular_scaling_unit);
                        atomic_add(M(2, i, to), (double)((1.0f / harm_ang) * wt));
                    }
                }
            }
            cnt.sum_ri += sri;
        }
        if (betweenness) {
            std::vector<uint32_t> sorted = sorted_states(t);
            std::vector<double> seed(t.state.size()), seedb(t.state.size(), 0.0);
            uint64_t sci = 0;
            for (int i = 0; i < D; ++i) {
                float thr = (float)seconds[i];
                std::fill(seed.begin(), seed.end(), 0.0);
                for (uint32_t to : t.reached_node_indices) {
                    if (to == src) continue;
                    // best_angular_target_states, centrality.rs:793-821
                    float brc = t.best_route_cost[to], bas = t.best_agg_seconds[to];
                    if (!std::isfinite(brc) || !std::isfinite(bas) || bas > thr) continue;
                    uint32_t best[2];
                    int nbest = 0;
                    for (uint32_t slot = 0; slot < 2; ++slot) {
                        uint32_t s = to * 2 + slot;
                        const BState& stt = t.state[s];
                        if (stt.sigma == 0.0 || stt.agg_seconds > thr) continue;
                        if (stt.route_cost <= brc * (1.0f + tol)) best[nbest++] = s;
                    }
                    if (nbest == 0) continue;
                    double pc = eligible[to] ? 0.5 : 1.0;
                    double sigma_total = 0.0;
                    for (int b = 0; b < nbest; ++b) sigma_total += t.state[best[b]].sigma;
                    if (sigma_total == 0.0) continue;
                    for (int b = 0; b < nbest; ++b) seed[best[b]] += pc * (t.state[best[b]].sigma / sigma_total);
                }
                backprop(
                    t, sorted, src, seed, seedb, [&](const BState& s) { return s.agg_seconds <= thr; },
                    [&](uint32_t node, double credit, double) {
                        sci++;
                        if (credit > 0.0) atomic_add(M(3, i, node), credit * (double)wt);
                    });
            }
            cnt.sum_ci += sci;
        }
    });
    if (counters_out) {
        counters_out[0] = cnt.sources;
        counters_out[1] = cnt.settled;
        counters_out[2] = cnt.edge_iters;
        counters_out[3] = cnt.sum_ri;
        counters_out[4] = cnt.sum_ci;
        counters_out[5] = 0;
        counters_out[6] = 0;
    }
    if (reach_totals)
        for (int i = 0; i < D; ++i) reach_totals[i] = reach[i];
    return 0;
}

// out: [4][D][node_bound] = seg density, harmonic, beta, betweenness.  Returns 3 if a twin edge lookup fails
// (the reference panics there, graph.rs:1291).
int orc_segment_centrality(const orc_graph* h, int D, const uint32_t* distances, const float* betas, const uint32_t* seconds,
                           float speed, int closeness, int betweenness, uint64_t n_sources, const uint32_t* sources,
                           double* out, uint64_t* counters_out, int n_threads) {
    const Graph& g = h->g;
    size_t nb = g.node_bound;
    uint32_t max_sec = *std::max_element(seconds, seconds + D);
    Counters cnt;
    std::atomic<int> err{0};
    auto M = [&](int m, int i, size_t node) -> double* { return out + ((size_t)m * D + i) * nb + node; };
    par_for(n_sources, n_threads, [&](uint64_t k) {
        uint32_t src = sources[k];
        std::vector<uint32_t> vn, ve;
        std::vector<NodeVisit> tm;
        std::vector<EdgeVisit> em;
        tree_segment(g, src, max_sec, speed, vn, ve, tm, em, &cnt);
        cnt.sources++;
        for (uint32_t eid : ve) {
            const EdgeVisit& ev = em[eid];
            uint32_t sn = (uint32_t)ev.start, en = (uint32_t)ev.end, ei = (uint32_t)ev.edge_idx;
            const NodeVisit& vn_ = tm[sn];
            const NodeVisit& vm_ = tm[en];
            if (!std::isfinite(vn_.short_dist) && !std::isfinite(vm_.short_dist)) continue;
            if (!closeness) continue;
            bool n_nearer = vn_.short_dist <= vm_.short_dist;
            float a = n_nearer ? vn_.short_dist : vm_.short_dist;
            float b = n_nearer ? vm_.short_dist : vn_.short_dist;
            float a_imp = a, b_imp = b;
            const Edge* pe = find_edge(g, sn, en, ei);
            if (!pe) {
                err = 3;
                return;
            }
            float edge_length = pe->length, imp = pe->imp;
            float c_ = (edge_length + a + b) / 2.0f;
            float d_ = c_;
            float c_imp = a_imp + (c_ - a) * imp;
            float d_imp = c_imp;
            for (int i = D - 1; i >= 0; --i) {
                float df = (float)distances[i];
                float beta = betas[i];
                float neg_beta = -beta;
                float inv_neg_beta = beta != 0.0f ? 1.0f / neg_beta : 0.0f;
                if (a < df) {
                    float cc = c_, cc_imp = c_imp;
                    if (cc > df) {
                        cc = df;
                        cc_imp = a_imp + (df - a) * imp;
                    }
                    atomic_add(M(0, i, src), (double)(cc - a));
                    float seg_harm = a_imp < 1.0f ? std::log(cc_imp)
                                                  : std::log(std::max(cc_imp / a_imp, std::numeric_limits<float>::epsilon()));
                    atomic_add(M(1, i, src), (double)seg_harm);
                    float bet = beta == 0.0f ? cc_imp - a_imp : (std::exp(neg_beta * cc_imp) - std::exp(neg_beta * a_imp)) * inv_neg_beta;
                    atomic_add(M(2, i, src), (double)bet);
                }
                if (b == d_) continue;
                if (b <= df) {
                    float cd = d_, cd_imp = d_imp;
                    if (cd > df) {
                        cd = df;
                        cd_imp = b_imp + (df - b) * imp;
                    }
                    atomic_add(M(0, i, src), (double)(cd - b));
                    float seg_harm = b_imp < 1.0f ? std::log(cd_imp)
                                                  : std::log(std::max(cd_imp / b_imp, std::numeric_limits<float>::epsilon()));
                    atomic_add(M(1, i, src), (double)seg_harm);
                    float bet = beta == 0.0f ? cd_imp - b_imp : (std::exp(neg_beta * cd_imp) - std::exp(neg_beta * b_imp)) * inv_neg_beta;
                    atomic_add(M(2, i, src), (double)bet);
                }
            }
        }
        if (betweenness) {
            uint64_t sci = 0;
            for (uint32_t to : vn) {
                if (to <= src) continue;
                const NodeVisit& tv = tm[to];
                if (!std::isfinite(tv.short_dist)) continue;
                const EdgeVisit& oe = em[(size_t)tv.origin_seg];
                const EdgeVisit& le = em[(size_t)tv.last_seg];
                const Edge* po = find_edge(g, (uint32_t)oe.start, (uint32_t)oe.end, (uint32_t)oe.edge_idx);
                const Edge* pl = find_edge(g, (uint32_t)le.start, (uint32_t)le.end, (uint32_t)le.edge_idx);
                if (!po || !pl) {
                    err = 3;
                    return;
                }
                float o_len = po->length, l_len = pl->length;
                float min_span = tv.short_dist - o_len - l_len;
                float o_1 = min_span, o_2 = min_span + o_len, l_1 = min_span, l_2 = min_span + l_len;
                int64_t cur = tv.pred;
                while (cur != NONE) {
                    if ((uint32_t)cur == src) break;
                    for (int i = D - 1; i >= 0; --i) {
                        float df = (float)distances[i];
                        float beta = betas[i];
                        if (min_span <= df) {
                            float o2s = std::min(o_2, df), l2s = std::min(l_2, df);
                            float auc;
                            if (beta == 0.0f) {
                                auc = (o2s - o_1) + (l2s - l_1);
                            } else {
                                float nbeta = -beta, inb = 1.0f / nbeta;
                                auc = (std::exp(nbeta * o2s) - std::exp(nbeta * o_1)) * inb +
                                      (std::exp(nbeta * l2s) - std::exp(nbeta * l_1)) * inb;
                            }
                            if (std::isfinite(auc) && auc >= 0.0f) {
                                sci++;
                                atomic_add(M(3, i, (size_t)cur), (double)auc);
                            }
                        }
                    }
                    cur = tm[(size_t)cur].pred;
                }
            }
            cnt.sum_ci += sci;
        }
    });
    if (counters_out) {
        counters_out[0] = cnt.sources;
        counters_out[1] = cnt.settled;
        counters_out[2] = cnt.edge_iters;
        counters_out[3] = 0;
        counters_out[4] = cnt.sum_ci;
        counters_out[5] = 0;
        counters_out[6] = 0;
    }
    return err.load();
}

// Tree searches: outputs sized node_bound (pred, short_dist, simpl_dist, agg_seconds, origin_seg, last_seg, flags bit0=visited
// bit1=discovered); visited_* arrays sized node_bound / edge_bound, counts returned through n_* pointers.
static void export_tree(const std::vector<NodeVisit>& tm, int64_t* pred, float* short_dist, float* simpl_dist, float* agg,
                        int64_t* origin_seg, int64_t* last_seg, uint8_t* flags) {
    for (size_t i = 0; i < tm.size(); ++i) {
        pred[i] = tm[i].pred;
        short_dist[i] = tm[i].short_dist;
        simpl_dist[i] = tm[i].simpl_dist;
        agg[i] = tm[i].agg_seconds;
        origin_seg[i] = tm[i].origin_seg;
        last_seg[i] = tm[i].last_seg;
        flags[i] = (uint8_t)((tm[i].visited ? 1 : 0) | (tm[i].discovered ? 2 : 0));
    }
}

int orc_dijkstra_tree_shortest(const orc_graph* h, uint32_t src, uint32_t max_seconds, float speed, uint32_t* visited,
                               uint64_t* n_visited, int64_t* pred, float* short_dist, float* simpl_dist, float* agg,
                               int64_t* origin_seg, int64_t* last_seg, uint8_t* flags) {
    std::vector<uint32_t> v;
    std::vector<NodeVisit> tm;
    tree_shortest(h->g, src, max_seconds, speed, v, tm);
    std::copy(v.begin(), v.end(), visited);
    *n_visited = v.size();
    export_tree(tm, pred, short_dist, simpl_dist, agg, origin_seg, last_seg, flags);
    return 0;
}

int orc_dijkstra_tree_simplest(const orc_graph* h, uint32_t src, uint32_t max_seconds, float speed, uint32_t* visited,
                               uint64_t* n_visited, int64_t* pred, float* short_dist, float* simpl_dist, float* agg,
                               int64_t* origin_seg, int64_t* last_seg, uint8_t* flags) {
    std::vector<std::array<int32_t, 2>> slots;
    int rc = endpoint_slots(h->g, slots);
    if (rc) return rc;
    std::vector<uint32_t> v;
    std::vector<NodeVisit> tm;
    tree_angular(h->g, src, max_seconds, speed, slots, v, tm);
    std::copy(v.begin(), v.end(), visited);
    *n_visited = v.size();
    export_tree(tm, pred, short_dist, simpl_dist, agg, origin_seg, last_seg, flags);
    return 0;
}

int orc_dijkstra_tree_segment(const orc_graph* h, uint32_t src, uint32_t max_seconds, float speed, uint32_t* visited,
                              uint64_t* n_visited, uint32_t* visited_edges, uint64_t* n_visited_edges, int64_t* pred,
                              float* short_dist, float* simpl_dist, float* agg, int64_t* origin_seg, int64_t* last_seg,
                              uint8_t* flags, int64_t* ev_start, int64_t* ev_end, int64_t* ev_edge_idx, uint8_t* ev_visited) {
    std::vector<uint32_t> v, ve;
    std::vector<NodeVisit> tm;
    std::vector<EdgeVisit> em;
    tree_segment(h->g, src, max_seconds, speed, v, ve, tm, em, nullptr);
    std::copy(v.begin(), v.end(), visited);
    *n_visited = v.size();
    std::copy(ve.begin(), ve.end(), visited_edges);
    *n_visited_edges = ve.size();
    export_tree(tm, pred, short_dist, simpl_dist, agg, origin_seg, last_seg, flags);
    for (size_t i = 0; i < em.size(); ++i) {
        ev_start[i] = em[i].start;
        ev_end[i] = em[i].end;
        ev_edge_idx[i] = em[i].edge_idx;
        ev_visited[i] = em[i].visited;
    }
    return 0;
}

// Per-source agg_seconds/route_cost dump for distance-level parity tests: out arrays sized node_bound.
int orc_shortest_distances(const orc_graph* h, uint32_t src, uint32_t max_seconds, float speed, float* agg_seconds,
                           float* route_cost, double* sigma) {
    Traversal t;
    brandes_shortest(h->g, src, max_seconds, speed, TIE_EPSILON, t, nullptr);
    for (uint32_t i = 0; i < h->g.node_bound; ++i) {
        agg_seconds[i] = t.best_agg_seconds[i];
        route_cost[i] = t.best_route_cost[i];
        sigma[i] = t.state[i].sigma;
    }
    return 0;
}

}  // extern "C"
