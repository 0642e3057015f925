// Single-source tree dumps: dijkstra_tree_segment (centrality.rs:1523-1611) and dijkstra_tree_simplest
// (dijkstra_tree_angular, centrality.rs:1202-1332).  Both are inspection entry points of the reference (tests and the
// data-layer callers read the tree of ONE source); their relaxation rules are order dependent (strict `<` against the
// running best / the angular tie rule against the running route metric), so the search is replayed in the reference's
// own order by one lane that owns a binary heap with the Rust sift rules (cs_simplest.cuh) — one launch, one source,
// a few thousand dependent steps.  The many-source paths never come through here.
#pragma once
#include "cs_common.cuh"
#include "cs_simplest.cuh"

struct CsTreeParams {
    uint32_t n;
    const uint32_t* off;  // segment: in-CSR offsets; angular: out-CSR offsets
    const CsEdge* rec;    // segment: in-CSR records; angular: angular out-records
    uint32_t src;
    float max_seconds;
    // per node [n]
    float* agg;        // inf
    float* simpl;      // inf (angular)
    uint32_t* pred;    // CS_NOSLOT
    uint32_t* origin;  // CS_NOSLOT (segment; in-CSR position)
    uint32_t* last;    // CS_NOSLOT (segment; in-CSR position)
    uint8_t* flags;    // bit 0 visited, bit 1 discovered
    // angular states [2n]
    float* st_metric;  // inf
    float* st_simpl;   // inf
    float* st_agg;     // inf
    uint8_t* st_flags; // bit 0 visited; per node bit 0 of reached[] below
    uint8_t* reached;  // [n] angular: node already listed in visited_nodes
    uint32_t* order;   // visited nodes
    uint32_t* eorder;  // visited edges (segment; in-CSR positions)
    uint32_t* counts;  // [0] nodes, [1] edges
    uint2* heap;
    uint32_t heap_cap;
    int* error;
};

__global__ void cs_k_fill_u32(uint32_t* p, size_t n, uint32_t v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// centrality.rs:1523-1611
__global__ void cs_k_tree_segment(CsTreeParams p) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    CsHeap h;
    cs_heap_init(h, nullptr, 0, p.heap);
    uint32_t nv = 0, ne = 0;
    p.agg[p.src] = 0.0f;
    p.flags[p.src] = 2;
    cs_heap_push(h, p.src, 0u);
    while (h.len > 0) {
        const uint32_t cur = cs_heap_pop(h).x;
        if (p.flags[cur] & 1) continue;
        p.flags[cur] |= 1;
        p.order[nv++] = cur;
        const float base = p.agg[cur];
        const uint32_t e1 = p.off[cur + 1];
        for (uint32_t e = p.off[cur]; e < e1; ++e) {
            const CsEdge r = p.rec[e];
            const uint32_t nb = r.nbr;
            if (nb == cur) {
                p.eorder[ne++] = e;
                continue;
            }
            if (p.flags[nb] & 1) continue;
            p.eorder[ne++] = e;
            const float ts = __fadd_rn(base, r.sec);
            if (ts > p.max_seconds) continue;
            if (ts < p.agg[nb]) {
                p.agg[nb] = ts;
                p.pred[nb] = cur;
                p.origin[nb] = cur == p.src ? e : p.origin[cur];
                p.last[nb] = e;
                p.flags[nb] |= 2;
                if (h.len >= p.heap_cap) {
                    *p.error = CS_ERR_QUEUE_OVERFLOW;
                    return;
                }
                cs_heap_push(h, nb, __float_as_uint(ts));
            }
        }
    }
    p.counts[0] = nv;
    p.counts[1] = ne;
}

// centrality.rs:1202-1332
__global__ void cs_k_tree_angular(CsTreeParams p) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    CsHeap h;
    cs_heap_init(h, nullptr, 0, p.heap);
    uint32_t nv = 0;
    p.order[nv++] = p.src;
    p.reached[p.src] = 1;
    p.flags[p.src] = 3;
    p.simpl[p.src] = 0.0f;
    p.agg[p.src] = 0.0f;
    for (uint32_t slot = 0; slot < 2; ++slot) {
        const uint32_t s = p.src * 2 + slot;
        p.st_metric[s] = 0.0f;
        p.st_simpl[s] = 0.0f;
        p.st_agg[s] = 0.0f;
        cs_heap_push(h, s, 0u);
    }
    while (h.len > 0) {
        const uint32_t si = cs_heap_pop(h).x;
        if (p.st_flags[si] & 1) continue;
        p.st_flags[si] |= 1;
        const uint32_t cur = si >> 1, entry = si & 1u;
        const float s_agg = p.st_agg[si], s_simpl = p.st_simpl[si];
        const uint32_t e1 = p.off[cur + 1];
        for (uint32_t e = p.off[cur]; e < e1; ++e) {
            const CsEdge r = p.rec[e];
            const uint32_t cslot = (r.nbr >> 30) & 1u, nslot = r.nbr >> 31, nx = r.nbr & 0x3fffffffu;
            if (cslot != 1u - entry) continue;
            const uint32_t ns = nx * 2 + nslot;
            const float cs = __fadd_rn(s_agg, r.sec);
            if (cs > p.max_seconds) continue;
            const float csimpl = __fadd_rn(s_simpl, r.aux);
            const float cmetric = __fadd_rn(csimpl, __uint_as_float(r.meta));
            const float old = p.st_metric[ns];
            const bool improved = __fadd_rn(cmetric, CS_TIE_EPS) < old;
            const bool tied = fabsf(__fsub_rn(cmetric, old)) <= CS_TIE_EPS;
            if (improved || (tied && cs < p.st_agg[ns] && si != ns)) {
                p.st_metric[ns] = cmetric;
                p.st_simpl[ns] = csimpl;
                p.st_agg[ns] = cs;
                if (h.len >= p.heap_cap) {
                    *p.error = CS_ERR_QUEUE_OVERFLOW;
                    return;
                }
                cs_heap_push(h, ns, __float_as_uint(cmetric));
                if (!p.reached[nx]) {
                    p.reached[nx] = 1;
                    p.order[nv++] = nx;
                }
                const float nsimpl = p.simpl[nx];
                const bool node_improved = __fadd_rn(csimpl, CS_TIE_EPS) < nsimpl;
                const bool node_tied = fabsf(__fsub_rn(csimpl, nsimpl)) <= CS_TIE_EPS;
                if (!(p.flags[nx] & 2) || node_improved || (node_tied && cs < p.agg[nx])) {
                    p.flags[nx] = 3;
                    p.simpl[nx] = csimpl;
                    p.agg[nx] = cs;
                    p.pred[nx] = cur;
                }
            }
        }
    }
    p.counts[0] = nv;
    p.counts[1] = 0;
}
