// Replica of Rust's std::collections::BinaryHeap over (item, key bits) pairs, as the reference uses it with NodeDistance
// (reversed f32 total order, /root/reference/rust/src/centrality.rs:358-386): a max-heap on the reversed order, i.e.
// pop returns the SMALLEST key, and equal keys pop in the deterministic but non-FIFO order that results from
//   push = append + sift_up (stop when the element is <= its parent), and
//   pop  = swap the last element into the root, sift_down_to_bottom (always descend to the larger child, taking the RIGHT
//          child when left <= right), then sift_up from the bottom.
// All keys here are non-negative floats, whose total order equals the unsigned order of their bit patterns.
//
// One lane drives a heap, so every instruction of it is issued at 1/32 lane efficiency: the entries live in the warp's
// shared memory (one LDS.64 / STS.64 per access through a 32-bit shared-space address, no generic-pointer selects).  A
// heap that outgrows its shared-memory image is copied to the arena once and continues there (`spilled`).
#pragma once
#include "cs_common.cuh"

struct CsHeap {
    uint32_t sbase;  // shared-space byte address of entry 0
    uint2* g;        // arena image: entries when spilled (and the only storage when nsm == 0)
    uint32_t nsm;    // entries that fit the shared-memory image
    uint32_t len;
    bool spilled;
};

__device__ __forceinline__ void cs_heap_init(CsHeap& h, void* smem, uint32_t nsm, uint2* arena) {
    h.sbase = smem ? (uint32_t)__cvta_generic_to_shared(smem) : 0u;
    h.g = arena;
    h.nsm = smem ? nsm : 0u;
    h.len = 0;
    h.spilled = h.nsm == 0;
}
__device__ __forceinline__ void cs_heap_clear(CsHeap& h) {
    h.len = 0;
    h.spilled = h.nsm == 0;
}

template <bool SM>
__device__ __forceinline__ uint2 cs_heap_get(const CsHeap& h, uint32_t i) {
    if constexpr (SM) {
        uint2 v;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(h.sbase + i * 8u));
        return v;
    } else {
        return h.g[i];
    }
}
template <bool SM>
__device__ __forceinline__ void cs_heap_set(const CsHeap& h, uint32_t i, uint2 v) {
    if constexpr (SM) {
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(h.sbase + i * 8u), "r"(v.x), "r"(v.y) : "memory");
    } else {
        h.g[i] = v;
    }
}
// Ord of NodeDistance: x <= y  <=>  x.metric >= y.metric
__device__ __forceinline__ bool cs_heap_le(uint2 a, uint2 b) { return a.y >= b.y; }

template <bool SM>
__device__ __forceinline__ void cs_heap_sift_up_t(const CsHeap& h, uint32_t pos) {
    const uint2 hole = cs_heap_get<SM>(h, pos);
    while (pos > 0) {
        const uint32_t parent = (pos - 1) / 2;
        const uint2 pv = cs_heap_get<SM>(h, parent);
        if (cs_heap_le(hole, pv)) break;
        cs_heap_set<SM>(h, pos, pv);
        pos = parent;
    }
    cs_heap_set<SM>(h, pos, hole);
}
template <bool SM>
__device__ __forceinline__ uint2 cs_heap_pop_t(CsHeap& h) {
    uint2 item = cs_heap_get<SM>(h, --h.len);
    if (h.len > 0) {
        const uint2 root = cs_heap_get<SM>(h, 0);
        const uint2 hole = item;  // the former last element travels down from the root
        item = root;
        const uint32_t end = h.len;
        uint32_t pos = 0, child = 1;
        while (end >= 2 && child <= end - 2) {
            const uint2 l = cs_heap_get<SM>(h, child), r = cs_heap_get<SM>(h, child + 1);
            uint2 c = l;
            if (cs_heap_le(l, r)) {
                child += 1;
                c = r;
            }
            cs_heap_set<SM>(h, pos, c);
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            cs_heap_set<SM>(h, pos, cs_heap_get<SM>(h, child));
            pos = child;
        }
        cs_heap_set<SM>(h, pos, hole);
        cs_heap_sift_up_t<SM>(h, pos);
    }
    return item;
}

// push / pop as the search loops call them (one lane)
__device__ __forceinline__ void cs_heap_push(CsHeap& h, uint32_t item, uint32_t key_bits) {
    if (!h.spilled && h.len == h.nsm) {
        for (uint32_t i = 0; i < h.len; ++i) h.g[i] = cs_heap_get<true>(h, i);
        h.spilled = true;
    }
    if (!h.spilled) {
        cs_heap_set<true>(h, h.len, make_uint2(item, key_bits));
        cs_heap_sift_up_t<true>(h, h.len);
    } else {
        cs_heap_set<false>(h, h.len, make_uint2(item, key_bits));
        cs_heap_sift_up_t<false>(h, h.len);
    }
    h.len++;
}
__device__ __forceinline__ uint2 cs_heap_pop(CsHeap& h) {
    return h.spilled ? cs_heap_pop_t<false>(h) : cs_heap_pop_t<true>(h);
}
