// segment_centrality on the GPU (reference: /root/reference/rust/src/centrality.rs:1523-1611 tree search,
// :2198-2402 accumulation).  One warp per source, same search / settle-order machinery as cs_shortest.cuh:
//   P1/P2  capped search over incoming edges + exact settle order (cs_search.cuh)
//   S3     forward in settle order: single predecessor = earliest-settled neighbour attaining the final seconds exactly
//          (strict `<` in the reference, :1589), origin / last segment lengths, and the continuous closeness integrals
//          over every visited edge (:2201-2317) accumulated in registers, one f64 add per metric/threshold at [i][src]
//   S5     reverse settle order: subtree sums of the per-target area-under-curve terms (:2319-2402); the reference walks
//          each target's predecessor chain, which adds auc(to) to every strict tree ancestor — i.e. credit(v) =
//          sum over strict descendants `to > src` of auc(to).
#pragma once
#include "cs_search.cuh"
#include "cs_heap.cuh"

struct CsSegmentParams {
    CsGraphDev g;
    int D, closeness, betweenness;
    float dist_f[CS_MAX_THRESHOLDS];
    float beta_f[CS_MAX_THRESHOLDS];
    float max_seconds, speed;
    const uint32_t* sources;
    unsigned long long n_sources;
    double* out;  // [4][D][n]: density, harmonic, beta, betweenness
    unsigned long long* counters;
    int* error;
    uint8_t* arena;
    CsArenaLayout lay;
    float delta, bin_scale;
    // single-source tree dump (dijkstra_tree_shortest, centrality.rs:1141-1200): settle order, predecessor and
    // seconds per node; NULL for centrality runs
    uint32_t* dump_order;  // [n] nodes in settle order
    uint32_t* dump_pred;   // [n] predecessor node, 0xffffffff = none
    float* dump_agg;       // [n] seconds (prefilled with inf by the caller)
    uint32_t* dump_count;  // [1] number of settled nodes
    // batched tree dumps (dijkstra_tree_shortest for many sources in one launch, replay == 1): source slot s owns entries
    // [s * batch_cap, (s + 1) * batch_cap): settled node in pop order, its predecessor (0xffffffff = none), its seconds;
    // b_count[s] = settled nodes (entries beyond batch_cap are dropped, the caller checks the count)
    uint32_t batch_cap;
    uint32_t* b_order;
    uint32_t* b_pred;
    float* b_agg;
    uint32_t* b_count;
    // Equal-key settle order.  The single-predecessor rule (strict `<` in pop order, :1589) makes the tree depend on the
    // order in which the reference's BinaryHeap pops nodes whose seconds are bit-equal.  replay == 0: a source where two
    // tree parents with bit-equal seconds compete for a node is not accumulated but appended to redo_list; replay == 1:
    // the search is replayed by one lane with the Rust heap (cs_seg_replay) and parents are ordered by its pop sequence.
    int replay;
    uint32_t* redo_list;
    // sources a CTA takes per round (1 .. CS_SEG_WARPS; warps beyond it idle at the barriers): a short replay list is
    // spread over all SMs instead of filling a few CTAs with 32 serial heap replays each
    uint32_t src_per_cta;
    // arena slots are indexed blockIdx.x * src_per_cta + warp instead of blockIdx.x * CS_SEG_WARPS + warp: the replay
    // launch behind the chain-contracted kernel runs in a small arena of its own
    int compact_arena;
};

// The heap of the replay (cs_heap.cuh): the first CS_SEG_HEAP_SMEM entries in the warp's shared-memory bins, the arena
// beyond that.
#define CS_SEG_HEAP_SMEM (CS_NBINS / 2)

// Replays dijkstra_tree_segment's heap loop (centrality.rs:1538-1609) over the reached set and records the pop sequence
// number of every settled node in popseq[rank].  Lane 0 only; the other lanes help with the initialisation.
__device__ __forceinline__ void cs_seg_replay(const CsGraphDev& g, const CsWarpArena& A, uint32_t* smem_words, uint32_t R,
                                              float max_seconds, uint32_t* popseq, float* run_agg, int& fail) {
    const uint32_t lane = cs_lane();
    // popseq / run_agg are written and read with ordinary (L1-cached) accesses here: one lane walks them serially, so a
    // hit in L1 instead of an L2 round trip per access halves the replay; the stores write through, and the later phases
    // read them from L2 as usual.  Every entry in use is re-initialised first, so no stale line of an earlier source is read.
    for (uint32_t r = lane; r < R; r += 32) {
        popseq[r] = CS_NOSLOT;
        run_agg[r] = __uint_as_float(CS_INF_BITS);
    }
    __syncwarp();
    if (lane == 0) {
        CsHeap h;
        cs_heap_init(h, smem_words, CS_SEG_HEAP_SMEM, A.qa);  // qa, qb and far are contiguous and dead after the order pass
        const uint32_t cap = 3u * A.qcap;
        uint32_t seq = 0;
        run_agg[0] = 0.0f;
        cs_heap_push(h, 0u, 0u);
        while (h.len > 0) {
            const uint32_t r = cs_heap_pop(h).x;
            if (popseq[r] != CS_NOSLOT) continue;  // lazy deletion (:1545)
            popseq[r] = seq++;
            const uint32_t cur = cs_ld(&A.s_node[r]);
            const float base = run_agg[r];
            const uint32_t eb = __ldg(&g.in_off[cur]);
            const uint32_t e1 = __ldg(&g.in_off[cur + 1]);
            for (uint32_t e = eb; e < e1; ++e) {
                const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&g.in_rec[e]));
                const uint32_t nb = raw.x;
                if (nb == cur) continue;
                const uint2 dnb = cs_ld(&A.ds[nb]);
                if (dnb.x == CS_INF_BITS) continue;  // every candidate of nb exceeds the cutoff
                if (popseq[dnb.y] != CS_NOSLOT) continue;
                const float ts = __fadd_rn(base, __uint_as_float(raw.y));
                if (ts > max_seconds) continue;
                if (ts < run_agg[dnb.y]) {
                    run_agg[dnb.y] = ts;
                    if (h.len >= cap) {
                        fail = CS_ERR_QUEUE_OVERFLOW;
                        break;
                    }
                    cs_heap_push(h, dnb.y, __float_as_uint(ts));
                }
            }
            if (fail) break;
        }
        __threadfence_block();
    }
    fail = __shfl_sync(CS_FULL, fail, 0);
    __syncwarp();
}

// one side of a visited edge: integrals of 1, 1/x, exp(-beta x) from `lo` towards `hi_raw`, clipped at the threshold
__device__ __forceinline__ void cs_seg_terms(float lo, float hi, float hi_imp, float imp, float thr, float beta, double& dens,
                                             double& harm, double& bet) {
    float cur = hi, cur_imp = hi_imp;
    if (cur > thr) {
        cur = thr;
        cur_imp = __fadd_rn(lo, __fmul_rn(__fsub_rn(thr, lo), imp));
    }
    dens += (double)__fsub_rn(cur, lo);
    const float seg_harm = lo < 1.0f ? logf(cur_imp) : logf(fmaxf(__fdiv_rn(cur_imp, lo), 1.1920929e-07f));
    harm += (double)seg_harm;
    float b;
    if (beta == 0.0f) {
        b = __fsub_rn(cur_imp, lo);
    } else {
        const float nb = -beta;
        b = __fmul_rn(__fsub_rn(cs_expf_libm(__fmul_rn(nb, cur_imp)), cs_expf_libm(__fmul_rn(nb, lo))), __fdiv_rn(1.0f, nb));
    }
    bet += (double)b;
}

// CTA shape: ONE CTA per SM whose warps take consecutive sources and move through the phases together (a barrier per
// phase).  The kernel is about 70 KB of SASS against a 32 KB instruction cache next to the SM; with three independent
// 8-warp CTAs the warps of an SM sat in different phases and starved on instruction fetch (same finding as the
// chain-contracted kernel, cs_shortest3.cuh).  Measured on the 1M-node graph (400/800/1600 m): 3 CTAs x 8 warps without
// barriers 796 k sources/s; one CTA of 16 / 24 / 28 / 32 warps 641 k / 847 k / 887 k / 915 k; 2 x 20 / 2 x 24: 867 k / 799 k.
#ifndef CS_SEG_WARPS
#define CS_SEG_WARPS 32
#endif
#ifndef CS_SEG_MIN_BLOCKS
#define CS_SEG_MIN_BLOCKS 1
#endif
#define CS_SEG_SMEM_BYTES (CS_SEG_WARPS * CS_NBINS * 4)
template <int DT>
__global__ void __launch_bounds__(CS_SEG_WARPS * 32, CS_SEG_MIN_BLOCKS) cs_k_segment(const CsSegmentParams p) {
    extern __shared__ __align__(16) uint32_t s_bins_all[];  // [CS_SEG_WARPS][CS_NBINS]
    __shared__ unsigned long long s_base;
    __shared__ int s_err;
    const uint32_t lane = cs_lane();
    const uint32_t wic = cs_warp_in_cta();
    const uint32_t worker = p.compact_arena ? blockIdx.x * p.src_per_cta + (wic < p.src_per_cta ? wic : 0u)
                                            : blockIdx.x * CS_SEG_WARPS + wic;
    uint32_t* bins = s_bins_all + (size_t)wic * CS_NBINS;
    const CsWarpArena A = cs_arena(p.arena, p.lay, worker);
    float2* seglen = reinterpret_cast<float2*>(A.sigma);  // per rank {origin segment length (-1 = pending), last segment length}
    const int D = p.D;
    const size_t n = p.g.n;
    const float f_inf = __uint_as_float(CS_INF_BITS);

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_base = atomicAdd(&p.counters[CS_C_NEXT], (unsigned long long)p.src_per_cta);
            s_err = *reinterpret_cast<volatile int*>(p.error);
        }
        __syncthreads();
        if (s_base >= p.n_sources || s_err != 0) break;
        const unsigned long long si = s_base + wic;
        bool run = wic < p.src_per_cta && si < p.n_sources;  // an idle warp still meets the barriers; its loops are empty
        const uint32_t src = run ? __ldg(&p.sources[si]) : 0u;

        unsigned long long relax = 0, edge_iters = 0, n_ci = 0;
        uint32_t R = 0;
        if (run) {
            int fail = 0;
            R = cs_p1_search(p.g, A, src, p.max_seconds, p.delta, relax, fail);
            if (fail) {
                if (lane == 0) atomicCAS(p.error, 0, fail);
                run = false;
                R = 0;
            }
        }
        __syncthreads();
        if (run) cs_p2_order(p.g, A, bins, src, R, p.bin_scale, edge_iters);
        __syncthreads();
        // pop sequence of the reference's heap (replay mode); A.dep is idle until S5
        uint32_t* popseq = reinterpret_cast<uint32_t*>(A.dep);
        if (run && p.replay) {
            int rfail = 0;
            cs_seg_replay(p.g, A, bins, R, p.max_seconds, popseq, reinterpret_cast<float*>(A.dep) + A.rcap, rfail);
            if (rfail) {
                if (lane == 0) atomicCAS(p.error, 0, rfail);
                cs_p6_reset(A, R);
                run = false;
                R = 0;
            }
        }
        bool ambiguous = false;

        // ------------------------------------------------------------------ S3: tree + closeness, forward
        double dens[DT], harm[DT], bet[DT];
#pragma unroll
        for (int i = 0; i < DT; ++i) dens[i] = harm[i] = bet[i] = 0.0;
        for (uint32_t b0 = 0; b0 < R; b0 += 32) {
            const uint32_t r = b0 + lane;
            const bool valid = r < R;
            uint32_t v = 0, pred_rank = CS_NOSLOT;
            float l_len = 0.f;
            bool pred_is_src = false;
            if (valid) {
                v = cs_ld(&A.s_node[r]);
                const uint32_t av_bits = __float_as_uint(cs_ld(&A.s_agg[r]));
                const float av = __uint_as_float(av_bits);
                // predecessor: earliest-settled u with agg[u] + sec(v->u) == agg[v] exactly (first strict improvement wins)
                uint32_t best_key = 0xffffffffu, best_pos = 0xffffffffu, best_j = 0, best_rank = CS_NOSLOT;
                uint32_t best_bits = 0, best_u = 0;
                if (v != src) {
                    const uint32_t eb = __ldg(&p.g.out_off[v]);
                    const uint32_t deg = __ldg(&p.g.out_off[v + 1]) - eb;
                    for (uint32_t j = 0; j < deg; ++j) {
                        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&p.g.out_rec[eb + j]));
                        const uint32_t u = raw.x;
                        if (u == v) continue;
                        const uint2 du = cs_ld(&A.ds[u]);
                        if (du.x == CS_INF_BITS || du.y >= r) continue;
                        const float c = __fadd_rn(__uint_as_float(du.x), __uint_as_float(raw.y));
                        if (__float_as_uint(c) != av_bits) continue;
                        const uint32_t ipos = raw.w & 0xffu;
                        // two different parents with bit-equal seconds: the winner is decided by the heap's pop order
                        if (best_rank != CS_NOSLOT && du.x == best_bits && u != best_u) ambiguous = true;
                        const uint32_t key = p.replay ? cs_ld(&popseq[du.y]) : du.y;
                        if (key < best_key || (key == best_key && ipos < best_pos)) {
                            best_key = key;
                            best_rank = du.y;
                            best_bits = du.x;
                            best_u = u;
                            best_pos = ipos;
                            best_j = j;
                            l_len = __uint_as_float(raw.z);  // length of the twin u->v = the "last segment" of v
                            pred_is_src = (u == src);
                        }
                    }
                }
                pred_rank = best_rank;
                cs_st(&A.predmask[r], pred_rank == CS_NOSLOT ? 0u : (1u << best_j));
                if (p.b_order) {
                    const uint32_t idx = p.replay ? cs_ld(&popseq[r]) : r;
                    if (idx < p.batch_cap) {
                        const size_t at = (size_t)si * p.batch_cap + idx;
                        p.b_order[at] = v;
                        p.b_pred[at] = pred_rank == CS_NOSLOT ? 0xffffffffu : cs_ld(&A.s_node[pred_rank]);
                        p.b_agg[at] = av;
                    }
                    if (r == 0) p.b_count[si] = R;
                }
                if (p.dump_order) {
                    p.dump_order[p.replay ? cs_ld(&popseq[r]) : r] = v;
                    p.dump_agg[v] = av;
                    p.dump_pred[v] = pred_rank == CS_NOSLOT ? 0xffffffffu : cs_ld(&A.s_node[pred_rank]);
                    if (r == 0) *p.dump_count = R;
                }
                // closeness over the edges visited from v: incoming (m->v) with m not settled before v, or a self-loop
                if (p.closeness) {
                    const float dn = __fmul_rn(av, p.speed);
                    const uint32_t eb = __ldg(&p.g.in_off[v]);
                    const uint32_t deg = __ldg(&p.g.in_off[v + 1]) - eb;
                    for (uint32_t j = 0; j < deg; ++j) {
                        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&p.g.in_rec[eb + j]));
                        const uint32_t m = raw.x;
                        float dm = dn;
                        if (m != v) {
                            const uint2 dmm = cs_ld(&A.ds[m]);
                            if (dmm.x != CS_INF_BITS && dmm.y < r) continue;  // m settled first: it visited this edge pair
                            dm = dmm.x == CS_INF_BITS ? f_inf : __fmul_rn(__uint_as_float(dmm.x), p.speed);
                        }
                        const float len = __uint_as_float(raw.z);  // twin v->m: get_edge_length_unchecked(start, end, idx)
                        const float imp = __ldg(&p.g.in_imp[eb + j]);
                        const bool n_nearer = dn <= dm;
                        const float a = n_nearer ? dn : dm;
                        const float b = n_nearer ? dm : dn;
                        const float c = __fdiv_rn(__fadd_rn(__fadd_rn(len, a), b), 2.0f);
                        const float c_imp = __fadd_rn(a, __fmul_rn(__fsub_rn(c, a), imp));
#pragma unroll
                        for (int i = DT - 1; i >= 0; --i) {
                            if (i >= D) continue;
                            const float thr = p.dist_f[i];
                            if (a < thr) cs_seg_terms(a, c, c_imp, imp, thr, p.beta_f[i], dens[i], harm[i], bet[i]);
                            if (b == c) continue;
                            if (b <= thr) cs_seg_terms(b, c, c_imp, imp, thr, p.beta_f[i], dens[i], harm[i], bet[i]);
                        }
                    }
                }
                // the source and orphans (no exact predecessor: only with zero-length edges) carry no segments
                cs_st(&seglen[r], make_float2((v == src || pred_rank == CS_NOSLOT) ? 0.0f : -1.0f, l_len));
            }
            bool pending = valid && v != src && pred_rank != CS_NOSLOT;
            for (;;) {
                if (pending) {
                    float o_len = l_len;
                    bool ok = true;
                    if (!pred_is_src) {
                        o_len = cs_ld(&seglen[pred_rank]).x;
                        ok = o_len >= 0.0f;
                    }
                    if (ok) {
                        cs_st(&seglen[r], make_float2(o_len, l_len));
                        pending = false;
                    }
                }
                __syncwarp();
                if (!__any_sync(CS_FULL, pending)) break;
            }
        }
        if (run && !p.replay && __any_sync(CS_FULL, ambiguous)) {
            // leave this source to the replay launch: nothing of it has been accumulated yet
            if (lane == 0) cs_st(&p.redo_list[atomicAdd(&p.counters[CS_C_FALLBACK], 1ull)], src);
            cs_p6_reset(A, R);
            run = false;
            R = 0;
        }
        if (run && p.closeness) {
#pragma unroll
            for (int i = 0; i < DT; ++i) {
                if (i < D) {
                    const double d0 = cs_warp_sum(dens[i]), d1 = cs_warp_sum(harm[i]), d2 = cs_warp_sum(bet[i]);
                    if (lane == 0) {
                        cs_red_add(p.out + ((size_t)(0 * D + i)) * n + src, d0);
                        cs_red_add(p.out + ((size_t)(1 * D + i)) * n + src, d1);
                        cs_red_add(p.out + ((size_t)(2 * D + i)) * n + src, d2);
                    }
                }
            }
        }

        __syncthreads();
        // ------------------------------------------------------------------ S5: subtree sums, reverse settle order
        if (run && p.betweenness) {
            for (int b0 = (int)((R - 1) & ~31u); b0 >= 0; b0 -= 32) {
                const uint32_t r = (uint32_t)b0 + lane;
                const bool valid = r < R;
                uint32_t w = 0;
                uint32_t crk[CS_MAX_DEGREE];
                int nchild = 0;
                uint32_t same_chunk = 0;
                double own[DT];
#pragma unroll
                for (int i = 0; i < DT; ++i) own[i] = 0.0;
                if (valid) {
                    w = cs_ld(&A.s_node[r]);
                    const float sd = __fmul_rn(cs_ld(&A.s_agg[r]), p.speed);
                    if (w > src) {
                        // area under the decay curve over the origin and last segments of the src -> w route (:2363-2391)
                        const float2 ol = cs_ld(&seglen[r]);
                        const float ms = __fsub_rn(__fsub_rn(sd, ol.x), ol.y);
                        const float o2 = __fadd_rn(ms, ol.x), l2 = __fadd_rn(ms, ol.y);
#pragma unroll
                        for (int i = 0; i < DT; ++i) {
                            if (i < D && ms <= p.dist_f[i]) {
                                const float thr = p.dist_f[i], beta = p.beta_f[i];
                                const float o2s = fminf(o2, thr), l2s = fminf(l2, thr);
                                float auc;
                                if (beta == 0.0f) {
                                    auc = __fadd_rn(__fsub_rn(o2s, ms), __fsub_rn(l2s, ms));
                                } else {
                                    const float nb = -beta, inb = __fdiv_rn(1.0f, nb);
                                    const float e0 = cs_expf_libm(__fmul_rn(nb, ms));
                                    auc = __fadd_rn(__fmul_rn(__fsub_rn(cs_expf_libm(__fmul_rn(nb, o2s)), e0), inb),
                                                    __fmul_rn(__fsub_rn(cs_expf_libm(__fmul_rn(nb, l2s)), e0), inb));
                                }
                                if (isfinite(auc) && auc >= 0.0f) own[i] = (double)auc;
                            }
                        }
                    }
                    const uint32_t eb = __ldg(&p.g.in_off[w]);
                    const uint32_t deg = __ldg(&p.g.in_off[w + 1]) - eb;
                    for (uint32_t j = 0; j < deg; ++j) {
                        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&p.g.in_rec[eb + j]));
                        const uint32_t x = raw.x;
                        if (x == w) continue;
                        const uint2 dx = cs_ld(&A.ds[x]);
                        if (dx.x == CS_INF_BITS || dx.y <= r) continue;
                        if ((cs_ld(&A.predmask[dx.y]) >> (raw.w & 0xffu)) & 1u) {
                            if (dx.y < (uint32_t)b0 + 32u) same_chunk |= 1u << nchild;
                            crk[nchild++] = dx.y;
                        }
                    }
                }
                bool pending = valid;
                for (;;) {
                    if (pending) {
                        bool ok = true;
                        for (uint32_t mm = same_chunk; mm; mm &= mm - 1) ok = ok && (cs_ld(&A.bdone[crk[__ffs(mm) - 1]]) != 0);
                        if (ok) {
                            double sub[DT];
#pragma unroll
                            for (int i = 0; i < DT; ++i) sub[i] = 0.0;
                            for (int k = 0; k < nchild; ++k) {
                                const double* dc = A.dep + (size_t)crk[k] * D;
#pragma unroll
                                for (int i = 0; i < DT; ++i)
                                    if (i < D) sub[i] += cs_ld(&dc[i]);
                            }
                            double* dr = A.dep + (size_t)r * D;
#pragma unroll
                            for (int i = 0; i < DT; ++i) {
                                if (i < D) {
                                    cs_st(&dr[i], own[i] + sub[i]);
                                    if (w != src && sub[i] > 0.0) {
                                        ++n_ci;
                                        cs_red_add(p.out + ((size_t)(3 * D + i)) * n + w, sub[i]);
                                    }
                                }
                            }
                            cs_st(&A.bdone[r], (uint8_t)1);
                            pending = false;
                        }
                    }
                    __syncwarp();
                    if (!__any_sync(CS_FULL, pending)) break;
                }
            }
        }

        __syncthreads();
        if (!run) continue;
        cs_p6_reset(A, R);
        edge_iters = cs_warp_sum(edge_iters);
        relax = cs_warp_sum(relax);
        n_ci = cs_warp_sum(n_ci);
        if (lane == 0) {
            atomicAdd(&p.counters[CS_C_SOURCES], 1ull);
            atomicAdd(&p.counters[CS_C_SETTLED], (unsigned long long)R);
            atomicAdd(&p.counters[CS_C_EDGE_ITERS], edge_iters);
            atomicAdd(&p.counters[CS_C_RELAX], relax);
            if (n_ci) atomicAdd(&p.counters[CS_C_SUM_CI], n_ci);
            atomicAdd(&p.counters[CS_C_PROGRESS], 1ull);
        }
    }
}
