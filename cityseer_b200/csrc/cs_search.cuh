// Warp-level building blocks shared by the shortest and segment kernels:
//   cs_p1_search  distance-capped label-correcting search over INCOMING edges (near/far buckets)
//   cs_p2_order   exact settle order of the reached nodes (counting sort + exact rank inside each bin)
// One warp owns one source at a time; all scratch lives in that warp's arena (global memory, L2-resident working set).
#pragma once
#include "cs_common.cuh"

struct CsArenaLayout {
    size_t ds, node_list, qa, qb, far, s_node, s_agg, predmask, sigma, dep, bdone, frank, needm, jrank, erank, stride;
    uint32_t rcap, qcap;
};

struct CsWarpArena {
    uint2* ds;             // dense [n]: {seconds bits, settle rank}; {inf, none} outside a search
    uint32_t* node_list;   // [rcap] reached nodes in discovery order
    uint2* qa;             // [qcap] frontier queues: {node | (skip_pos+1) << 26, seconds bits}
    uint2* qb;
    uint2* far;
    unsigned long long* tmp_key;  // aliases qa after the search: (seconds bits << 32 | node tie key), bin-scattered
    uint32_t* s_node;      // [rcap] node by settle rank
    float* s_agg;          // [rcap] seconds by settle rank
    uint32_t* predmask;    // [rcap] per-rank 32-bit word (shortest: predecessor adjacency mask)
    double* sigma;         // [rcap] per-rank f64
    double* dep;           // [rcap][2D] per-rank f64 vectors
    uint8_t* bdone;        // [rcap * 8] per-rank flags / scratch
    uint4* erank;          // [rcap] node-level kernels: {in_off, in-degree, out_off, out-degree} of the node at each rank
    uint32_t rcap, qcap;
};

#define CS_NODE_BITS 26
#define CS_NODE_MASK ((1u << CS_NODE_BITS) - 1u)

__device__ __forceinline__ CsWarpArena cs_arena(uint8_t* arena, const CsArenaLayout& L, uint32_t worker) {
    uint8_t* base = arena + (size_t)worker * L.stride;
    CsWarpArena A;
    A.ds = reinterpret_cast<uint2*>(base + L.ds);
    A.node_list = reinterpret_cast<uint32_t*>(base + L.node_list);
    A.qa = reinterpret_cast<uint2*>(base + L.qa);
    A.qb = reinterpret_cast<uint2*>(base + L.qb);
    A.far = reinterpret_cast<uint2*>(base + L.far);
    A.tmp_key = reinterpret_cast<unsigned long long*>(base + L.qa);
    A.s_node = reinterpret_cast<uint32_t*>(base + L.s_node);
    A.s_agg = reinterpret_cast<float*>(base + L.s_agg);
    A.predmask = reinterpret_cast<uint32_t*>(base + L.predmask);
    A.sigma = reinterpret_cast<double*>(base + L.sigma);
    A.dep = reinterpret_cast<double*>(base + L.dep);
    A.bdone = base + L.bdone;
    A.erank = reinterpret_cast<uint4*>(base + L.erank);
    A.rcap = L.rcap;
    A.qcap = L.qcap;
    return A;
}

// dense maps of all workers in one launch: blockIdx.y = worker
__global__ void cs_k_init_ds(uint8_t* arena, size_t stride, size_t ds_off, size_t n_states) {
    uint2* ds = reinterpret_cast<uint2*>(arena + (size_t)blockIdx.y * stride + ds_off);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < n_states; i += step) ds[i] = make_uint2(CS_INF_BITS, CS_NOSLOT);
}

// P1. Returns the number of reached nodes R (source included); `fail` != 0 on arena overflow.
// f32 `+` is monotone, so the fixed point of this label-correcting search equals the reference's Dijkstra distances
// bit for bit (centrality.rs:1363-1442 / :1543-1609).  An item remembers which incoming edge leads back to the node
// that produced it: relaxing that edge can never improve (non-negative weights), so it is skipped.
__device__ __forceinline__ uint32_t cs_p1_search(const CsGraphDev& g, const CsWarpArena& A, uint32_t src, float max_seconds,
                                                 float delta, unsigned long long& relax, int& fail) {
    const uint32_t lane = cs_lane();
    const uint32_t ltmask = cs_lanemask_lt();
    uint2* qc = A.qa;
    uint2* qn = A.qb;
    uint2* far = A.far;
    uint32_t nc = 1, nn = 0, nf = 0, count = 1;
    float thr = delta;
    fail = 0;
    if (lane == 0) {
        cs_st(&A.ds[src], make_uint2(0u, CS_NOSLOT));
        cs_st(&A.node_list[0], src);
        cs_st(&qc[0], make_uint2(src, 0u));
    }
    __syncwarp();
    for (;;) {
        while (nc > 0) {
            for (uint32_t b0 = 0; b0 < nc; b0 += 32) {
                const uint32_t idx = b0 + lane;
                bool valid = idx < nc;
                uint32_t v = 0, abits = 0, skip = 0xffffffffu;
                if (valid) {
                    const uint2 it = cs_ld(&qc[idx]);
                    v = it.x & CS_NODE_MASK;
                    skip = (it.x >> CS_NODE_BITS) - 1u;
                    abits = it.y;
                    valid = cs_ld(&A.ds[v].x) == abits;  // stale entries were superseded by a smaller distance
                }
                uint32_t eb = 0, deg = 0;
                if (valid) {
                    eb = __ldg(&g.in_off[v]);
                    deg = __ldg(&g.in_off[v + 1]) - eb;
                }
                const uint32_t maxdeg = __reduce_max_sync(CS_FULL, deg);
                const float a = __uint_as_float(abits);
                // four edges per lane at a time: the edge records, then the atomicMin on the neighbours' distances, are
                // issued for the whole group before their results are consumed (two round trips per group, not per edge)
                for (uint32_t j0 = 0; j0 < maxdeg; j0 += 4) {
                    uint4 raws[4];
                    bool use[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        use[t] = j0 + t < deg && j0 + t != skip;
                        if (use[t]) raws[t] = __ldg(reinterpret_cast<const uint4*>(&g.in_rec[eb + j0 + t]));
                    }
                    uint32_t olds[4], cbs[4];
                    float cands[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        cands[t] = 0.f;
                        cbs[t] = 0;
                        olds[t] = 0;
                        if (use[t]) {
                            cands[t] = __fadd_rn(a, __uint_as_float(raws[t].y));
                            use[t] = raws[t].x != v && !(cands[t] > max_seconds);
                            if (use[t]) {
                                cbs[t] = __float_as_uint(cands[t]);
                                olds[t] = atomicMin(&A.ds[raws[t].x].x, cbs[t]);
                            }
                        }
                    }
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if (j0 + t >= maxdeg) break;  // warp-uniform
                        const uint32_t nb = use[t] ? raws[t].x : 0u;
                        const uint32_t cbits = cbs[t];
                        const float cand = cands[t];
                        const bool improved = use[t] && cbits < olds[t];
                        const bool first = use[t] && olds[t] == CS_INF_BITS;
                        // (position of the twin edge in nb's in-list) + 1, or 0
                        const uint32_t back = use[t] ? (raws[t].w >> 16) & 0x3fu : 0u;
                        uint32_t m = __ballot_sync(CS_FULL, first);
                        if (m) {
                            const uint32_t pos = count + __popc(m & ltmask);
                            if (first && pos < A.rcap) cs_st(&A.node_list[pos], nb);
                            count += __popc(m);
                        }
                        const bool pn = improved && (cand < thr);
                        const bool pf = improved && !pn;
                        const uint2 item = make_uint2(nb | (back << CS_NODE_BITS), cbits);
                        m = __ballot_sync(CS_FULL, pn);
                        if (m) {
                            const uint32_t pos = nn + __popc(m & ltmask);
                            if (pn && pos < A.qcap) cs_st(&qn[pos], item);
                            nn += __popc(m);
                        }
                        m = __ballot_sync(CS_FULL, pf);
                        if (m) {
                            const uint32_t pos = nf + __popc(m & ltmask);
                            if (pf && pos < A.qcap) cs_st(&far[pos], item);
                            nf += __popc(m);
                        }
                        relax += improved ? 1ull : 0ull;
                    }
                }
            }
            if (count > A.rcap || nn > A.qcap || nf > A.qcap) {
                fail = count > A.rcap ? CS_ERR_REACH_OVERFLOW : CS_ERR_QUEUE_OVERFLOW;
                return count;
            }
            uint2* t = qc;
            qc = qn;
            qn = t;
            nc = nn;
            nn = 0;
            __syncwarp();
        }
        if (nf == 0) break;
        // near bucket exhausted: advance the threshold past the smallest live far entry and split the far pile
        float mn = __uint_as_float(CS_INF_BITS);
        for (uint32_t i = lane; i < nf; i += 32) {
            const uint2 it = cs_ld(&far[i]);
            if (cs_ld(&A.ds[it.x & CS_NODE_MASK].x) == it.y) mn = fminf(mn, __uint_as_float(it.y));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(CS_FULL, mn, o));
        if (!(mn < __uint_as_float(CS_INF_BITS))) break;
        thr = mn + delta;
        uint32_t w = 0;
        nc = 0;
        for (uint32_t b0 = 0; b0 < nf; b0 += 32) {
            const uint32_t idx = b0 + lane;
            bool livee = idx < nf;
            uint2 it = make_uint2(0u, 0u);
            if (livee) {
                it = cs_ld(&far[idx]);
                livee = cs_ld(&A.ds[it.x & CS_NODE_MASK].x) == it.y;
            }
            const bool near = livee && (__uint_as_float(it.y) < thr);
            const bool keep = livee && !near;
            __syncwarp();
            uint32_t m = __ballot_sync(CS_FULL, near);
            if (near) cs_st(&qc[nc + __popc(m & ltmask)], it);
            nc += __popc(m);
            m = __ballot_sync(CS_FULL, keep);
            if (keep) cs_st(&far[w + __popc(m & ltmask)], it);
            w += __popc(m);
        }
        nf = w;
        __syncwarp();
    }
    return count;
}

// Monotone bin of a distance: quadratic in the distance, so that bins hold ~equal node counts on a 2-D street network.
__device__ __forceinline__ uint32_t cs_bin(uint32_t ab, float bin_scale) {
    const float a = __uint_as_float(ab);
    return min((uint32_t)(CS_NBINS - 1), (uint32_t)(__fmul_rn(__fmul_rn(a, a), bin_scale)));
}

// P2. Exact settle order by (seconds, node) with the source first: fills s_node / s_agg by rank, writes each node's
// rank into ds[node].y, zeroes sigma / bdone for the ranks in use.  `bins` = CS_NBINS words of shared memory (per warp).
__device__ __forceinline__ void cs_p2_order(const CsGraphDev& g, const CsWarpArena& A, uint32_t* bins, uint32_t src,
                                            uint32_t R, float bin_scale, unsigned long long& edge_iters) {
    const uint32_t lane = cs_lane();
    // node_list is dead once its entries have been scattered into the sort keys: the shortest kernel reuses it as the
    // per-rank "smallest successor" array of its dependency pass, initialised here while the ranks are being written
    uint32_t* minsucc_out = A.node_list;
    for (uint32_t i = lane; i < CS_NBINS; i += 32) bins[i] = 0;
    __syncwarp();
    for (uint32_t i = lane; i < R; i += 32) {
        const uint32_t node = cs_ld(&A.node_list[i]);
        const uint32_t ab = cs_ld(&A.ds[node].x);
        cs_st(reinterpret_cast<uint32_t*>(&A.s_agg[i]), ab);  // discovery-indexed temp, rewritten by rank below
        atomicAdd(&bins[cs_bin(ab, bin_scale)], 1u);
    }
    __syncwarp();
    {
        uint32_t carry = 0;
        for (uint32_t k = 0; k < CS_NBINS / 32; ++k) {
            const uint32_t c = bins[k * 32 + lane];
            uint32_t inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(CS_FULL, inc, o);
                if ((int)lane >= o) inc += t;
            }
            bins[k * 32 + lane] = carry + inc - c;
            carry += __shfl_sync(CS_FULL, inc, 31);
        }
    }
    __syncwarp();
    for (uint32_t i = lane; i < R; i += 32) {
        const uint32_t node = cs_ld(&A.node_list[i]);
        const uint32_t ab = cs_ld(reinterpret_cast<const uint32_t*>(&A.s_agg[i]));
        const uint32_t pos = atomicAdd(&bins[cs_bin(ab, bin_scale)], 1u);
        // the source sorts first among zero-distance nodes (it is always the first settled state)
        cs_st(&A.tmp_key[pos], ((unsigned long long)ab << 32) | (node == src ? 0u : node + 1u));
    }
    __syncwarp();
    // bins[b] now holds the end offset of bin b
    for (uint32_t pos = lane; pos < R; pos += 32) {
        const unsigned long long key = cs_ld(&A.tmp_key[pos]);
        const uint32_t ab = (uint32_t)(key >> 32);
        const uint32_t bin = cs_bin(ab, bin_scale);
        const uint32_t start = bin ? bins[bin - 1] : 0u;
        const uint32_t end = bins[bin];
        uint32_t rank = start;
        for (uint32_t j = start; j < end; ++j) rank += (cs_ld(&A.tmp_key[j]) < key) ? 1u : 0u;
        const uint32_t low = (uint32_t)key;
        const uint32_t node = low ? low - 1u : src;
        cs_st(&A.s_node[rank], node);
        cs_st(&A.s_agg[rank], __uint_as_float(ab));
        cs_st(&A.ds[node].y, rank);
        cs_st(&A.sigma[rank], 0.0);
        cs_st(reinterpret_cast<unsigned long long*>(A.bdone) + rank, 0ull);
        cs_st(&minsucc_out[rank], CS_NOSLOT);
        // the CSR rows of the node, by rank: the later phases read them next to s_node / s_agg (one coalesced round trip)
        // instead of chasing in_off[s_node[r]]
        const uint32_t ib = __ldg(&g.in_off[node]), ie = __ldg(&g.in_off[node + 1]);
        const uint32_t ob = __ldg(&g.out_off[node]), oe = __ldg(&g.out_off[node + 1]);
        cs_st(&A.erank[rank], make_uint4(ib, ie - ib, ob, oe - ob));
        edge_iters += ie - ib;
    }
    __syncwarp();
}

__device__ __forceinline__ void cs_p6_reset(const CsWarpArena& A, uint32_t R) {
    for (uint32_t r = cs_lane(); r < R; r += 32) cs_st(&A.ds[cs_ld(&A.s_node[r])], make_uint2(CS_INF_BITS, CS_NOSLOT));
    __syncwarp();
}
