// centrality_shortest on the GPU: one warp per source, persistent grid, all per-source scratch in a per-warp arena.
//
// Phases per source (reference: /root/reference/rust/src/centrality.rs):
//   P1  distance-capped label-correcting search over INCOMING edges (cs_search.cuh; :1363-1442).
//   P2  exact settle order: counting sort on quantised seconds, exact rank inside each bin by (seconds, node).
//   P3  predecessor sets + sigma by pulling over OUTGOING edges in settle order through the reference's epsilon rule
//       (:1413-1437), or the tolerance rule of phase 2 (:1447-1483); circuit-rank counts (:470-526).
//   P4  closeness terms formed in f32, widened to f64, scattered to the TARGET with red.global.add.f64 (:1733-1780).
//   P5  Brandes dependencies in reverse settle order, all D thresholds at once, pull form (:823-873, :1783-1835).
//   P6  reset the touched entries of the dense per-warp distance map.
#pragma once
#include "cs_search.cuh"

struct CsShortestParams {
    CsGraphDev g;
    int D, closeness, betweenness, phase2;
    float dist_f[CS_MAX_THRESHOLDS];
    float beta_f[CS_MAX_THRESHOLDS];
    double beta_d[CS_MAX_THRESHOLDS];
    float max_seconds, speed, tol;
    const uint32_t* sources;
    const float* src_wt;
    unsigned long long n_sources;
    const uint8_t* eligible;
    double* out;    // [7][D][n] final layout (written by cs_k_epilogue_shortest)
    double* acc_c;  // [n][cw] node-interleaved closeness accumulators: slot q = 5 * i + m
    double* acc_b;  // [n][bw] node-interleaved betweenness accumulators: slot q = i (plain), D + i (beta)
    int cw, bw;
    unsigned long long* counters;
    int* error;
    uint8_t* arena;
    CsArenaLayout lay;
    float delta, bin_scale;
    float* dump_agg;  // optional per-node dumps (single-source debug search)
    double* dump_sigma;
    uint32_t* dump_npred;
    // betweenness_od_shortest (centrality.rs:2419-2540): seeds only at the OD destinations of each source; the
    // destinations / weights of sources[k] are od_dst / od_w [od_off[k], od_off[k + 1]); NULL for plain centrality
    const unsigned long long* od_off;
    const uint32_t* od_dst;
    const float* od_w;
};

#ifndef CS_MIN_BLOCKS
#define CS_MIN_BLOCKS 3
#endif

// first threshold index whose distance admits `cost` (thresholds are strictly increasing), DT if none
template <int DT>
__device__ __forceinline__ int cs_first_threshold(const CsShortestParams& p, float cost) {
    int ti = DT;
#pragma unroll
    for (int i = DT - 1; i >= 0; --i)
        if (i < p.D && cost <= p.dist_f[i]) ti = i;
    return ti;
}

template <int DT>
__global__ void __launch_bounds__(CS_WARPS_PER_CTA * 32, CS_MIN_BLOCKS) cs_k_shortest(const CsShortestParams p) {
    __shared__ uint32_t s_bins_all[CS_WARPS_PER_CTA][CS_NBINS];
    __shared__ uint32_t s_hist_all[CS_WARPS_PER_CTA][2][CS_MAX_THRESHOLDS + 1];
    __shared__ float s_rank_all[CS_WARPS_PER_CTA][CS_MAX_THRESHOLDS];

    const uint32_t lane = cs_lane();
    const uint32_t wic = cs_warp_in_cta();
    const uint32_t worker = blockIdx.x * CS_WARPS_PER_CTA + wic;
    uint32_t* bins = s_bins_all[wic];
    uint32_t* histN = s_hist_all[wic][0];
    uint32_t* histE = s_hist_all[wic][1];
    float* rankf = s_rank_all[wic];
    const CsWarpArena A = cs_arena(p.arena, p.lay, worker);
    uint32_t* minsucc = A.node_list;  // [rcap] after P2 (which leaves it at "none"): smallest successor rank per node
    const int D = p.D;
    const int D2 = 2 * D;
    const float one_minus = 1.0f - CS_TIE_EPS, one_plus = 1.0f + CS_TIE_EPS;
    const float one_plus_tol = 1.0f + p.tol;

    for (;;) {
        unsigned long long si = 0;
        if (lane == 0) si = atomicAdd(&p.counters[CS_C_NEXT], 1ull);
        si = __shfl_sync(CS_FULL, si, 0);
        if (si >= p.n_sources) break;
        if (*reinterpret_cast<volatile int*>(p.error) != 0) break;
        const uint32_t src = __ldg(&p.sources[si]);  // (broadcasting these as cs_uni values costs this kernel 5 %)
        const float wt = __ldg(&p.src_wt[si]);

        // ------------------------------------------------------------------ P1 + P2
        unsigned long long relax = 0, edge_iters = 0;
        int fail = 0;
        long long tc[7];
        tc[0] = clock64();
        const uint32_t R = cs_p1_search(p.g, A, src, p.max_seconds, p.delta, relax, fail);
        tc[1] = clock64();
        if (fail) {
            if (lane == 0) atomicCAS(p.error, 0, fail);
            break;
        }
        if (lane <= CS_MAX_THRESHOLDS) {
            histN[lane] = 0;
            histE[lane] = 0;
        }
        cs_p2_order(p.g, A, bins, src, R, p.bin_scale, edge_iters);
        tc[2] = clock64();

        // ------------------------------------------------------------------ P3: predecessors + sigma (outgoing edges)
        for (uint32_t b0 = 0; b0 < R; b0 += 32) {
            const uint32_t r = b0 + lane;
            const bool valid = r < R;
            uint32_t v = 0;
            // candidate arrays (local memory; typical out-degree on street graphs is 2-4)
            float cc[CS_MAX_DEGREE];
            uint32_t cu[CS_MAX_DEGREE], crk[CS_MAX_DEGREE], cj[CS_MAX_DEGREE];
            int ncand = 0;
            uint32_t pmask_c = 0;
            if (valid) {
                v = cs_ld(&A.s_node[r]);
                const float av = cs_ld(&A.s_agg[r]);
                const float cost_v = __fmul_rn(av, p.speed);
                if (p.closeness) atomicAdd(&histN[cs_first_threshold<DT>(p, cost_v)], 1u);
                const uint4 er = cs_ld(&A.erank[r]);
                const uint32_t eb = er.z, deg = er.w;
                // four out-edges at a time: the edge records, then the neighbours' map entries, are in flight together
                for (uint32_t j0 = 0; j0 < deg; j0 += 4) {
                    uint4 raws[4];
                    uint2 dus[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        raws[t] = j0 + t < deg ? __ldg(reinterpret_cast<const uint4*>(&p.g.out_rec[eb + j0 + t])) : make_uint4(v, 0u, 0u, 0u);
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        dus[t] = raws[t].x != v ? cs_ld(&A.ds[raws[t].x]) : make_uint2(CS_INF_BITS, CS_NOSLOT);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint4 raw = raws[t];
                        const uint32_t j = j0 + t;
                        const uint32_t u = raw.x;
                        if (u == v) continue;  // self-loops and the padding of the last group
                        const uint2 du = dus[t];
                        if (du.x == CS_INF_BITS) continue;
                        const float au = __uint_as_float(du.x);
                        if (p.closeness && (raw.w & 0x100u)) {
                            const float ec = fmaxf(cost_v, __fmul_rn(au, p.speed));
                            atomicAdd(&histE[cs_first_threshold<DT>(p, ec)], 1u);
                        }
                        if (du.y >= r || v == src) continue;  // u must be settled before v; the source has no predecessors
                        const float c = __fadd_rn(au, __uint_as_float(raw.y));
                        if (!p.phase2 && c > p.max_seconds) continue;
                        // insertion by (settle rank of u, position in u's incoming list)
                        const uint32_t skey = du.y;
                        const uint32_t ipos = raw.w & 0xffu;
                        int k = ncand++;
                        while (k > 0 && (crk[k - 1] > skey || (crk[k - 1] == skey && (cj[k - 1] >> 8) > ipos))) {
                            cc[k] = cc[k - 1];
                            cu[k] = cu[k - 1];
                            crk[k] = crk[k - 1];
                            cj[k] = cj[k - 1];
                            --k;
                        }
                        cc[k] = c;
                        cu[k] = u;
                        crk[k] = skey;
                        cj[k] = j | (ipos << 8);
                    }
                }
                if (ncand == 1 && !p.phase2) {
                    pmask_c = 1u;
                } else if (!p.phase2) {
                    // epsilon rule, sequential in settle order (centrality.rs:1413-1437)
                    float old = __uint_as_float(CS_INF_BITS);
                    for (int k = 0; k < ncand; ++k) {
                        const float c = cc[k];
                        if (c < old) {
                            if (c < __fmul_rn(old, one_minus)) pmask_c = 0;
                            pmask_c |= 1u << k;  // no duplicate check on this branch (:1426)
                            old = c;
                        } else if (c <= __fmul_rn(old, one_plus)) {
                            bool dup = false;
                            for (uint32_t mm = pmask_c; mm; mm &= mm - 1) dup |= cu[__ffs(mm) - 1] == cu[k];
                            if (!dup) pmask_c |= 1u << k;
                        }
                    }
                } else {
                    // tolerance rule against final distances (centrality.rs:1457-1482)
                    const float lim = __fmul_rn(av, one_plus_tol);
                    for (int k = 0; k < ncand; ++k) {
                        if (cc[k] <= lim) {
                            bool dup = false;
                            for (uint32_t mm = pmask_c; mm; mm &= mm - 1) dup |= cu[__ffs(mm) - 1] == cu[k];
                            if (!dup) pmask_c |= 1u << k;
                        }
                    }
                }
                // A reached node without an earlier-settled predecessor: its predecessor ties with it on (seconds, index)
                // through a zero-second edge, and the reference settles such pairs in heap order.  Outside this kernel's
                // contract: fail loudly (and never leave sigma at zero, which the nodes behind it would wait on).
                if (v != src && pmask_c == 0) atomicCAS(p.error, 0, CS_ERR_ZERO_TIE);
                uint32_t amask = 0;
                for (uint32_t mm = pmask_c; mm; mm &= mm - 1) amask |= 1u << (cj[__ffs(mm) - 1] & 0xffu);
                cs_st(&A.predmask[r], amask);
                // the dependency pass forms chunks whose nodes do not depend on each other: the smallest rank that has
                // this node as a predecessor bounds the chunk that may contain it
                for (uint32_t mm = pmask_c; mm; mm &= mm - 1) atomicMin(&minsucc[crk[__ffs(mm) - 1]], r);
                if (v == src) cs_st(&A.sigma[r], 1.0);
                if (p.dump_npred) p.dump_npred[v] = __popc(pmask_c);
            }
            // sigma = sum over predecessors in settle order; only predecessors in this same chunk can be pending
            bool pending = valid && v != src;
            for (;;) {
                if (pending) {
                    double s = 0.0;
                    bool ok = true;
                    for (uint32_t mm = pmask_c; mm; mm &= mm - 1) {
                        const double sg = cs_ld(&A.sigma[crk[__ffs(mm) - 1]]);
                        if (sg == 0.0) {
                            ok = false;
                            break;
                        }
                        s += sg;
                    }
                    if (ok) {
                        cs_st(&A.sigma[r], pmask_c ? s : 1.0);
                        pending = false;
                    }
                }
                __syncwarp();
                if (!__any_sync(CS_FULL, pending)) break;
            }
        }
        __syncwarp();
        if (p.dump_agg) {
            for (uint32_t r = lane; r < R; r += 32) {
                const uint32_t node = cs_ld(&A.s_node[r]);
                p.dump_agg[node] = cs_ld(&A.s_agg[r]);
                p.dump_sigma[node] = cs_ld(&A.sigma[r]);
            }
        }

        // ------------------------------------------------------------------ P4: closeness scatter to targets
        tc[3] = clock64();
        unsigned long long n_ri = 0, n_ci = 0;
        if (p.closeness) {
            if (lane < (uint32_t)D) {
                // circuit rank per threshold = max(0, E_i - N_i + 1) over the reached subgraph (centrality.rs:517-525)
                long long ncount = 0, ecount = 0;
                for (int t = 0; t <= (int)lane; ++t) {
                    ncount += histN[t];
                    ecount += histE[t];
                }
                rankf[lane] = ncount == 0 ? 0.0f : (float)max(ecount - ncount + 1ll, 0ll);
                // reachable targets per threshold exclude the source itself
                atomicAdd(&p.counters[CS_C_REACH0 + lane], (unsigned long long)(ncount > 0 ? ncount - 1 : 0));
            }
            __syncwarp();
            const float cycles_wt = __fdiv_rn(wt, __ldg(&p.g.weight[src]));  // centrality.rs:1730
            // Packed scatter: the 5*D accumulators of one target are contiguous (one or two 128-byte lines), so a warp
            // instruction covers 32/LP targets with LP consecutive doubles each instead of 32 unrelated lines
            // (profiles/r01c_red_microbench.txt: 8-9x the red.f64 throughput of the per-metric [M][D][N] layout).
            constexpr int NQ = 5 * DT;
            constexpr int LP = NQ <= 16 ? 16 : 32;
            constexpr int G = 32 / LP;
            const uint32_t ql = lane & (LP - 1);
            const int nq = 5 * D;
            // lane-constant slot description (single round when 5*DT <= LP, else per round below)
            for (uint32_t b0 = 0; b0 < R; b0 += 32) {
                const uint32_t r = b0 + lane;
                uint32_t node = 0;
                float cost = __uint_as_float(CS_INF_BITS);
                if (r < R) {
                    node = cs_ld(&A.s_node[r]);
                    if (node != src) cost = __fmul_rn(cs_ld(&A.s_agg[r]), p.speed);
                }
                // per-target terms formed once by the target's own lane, exactly as centrality.rs:1755-1777 (f32)
                const float far_t = __fmul_rn(cost, wt);
                const float harm_t = __fmul_rn(__fdiv_rn(1.0f, cost), wt);
                float bet_t[DT];
#pragma unroll
                for (int i = 0; i < DT; ++i)
                    bet_t[i] = (i < D && cost <= p.dist_f[i]) ? __fmul_rn(expf(__fmul_rn(-p.beta_f[i], cost)), wt) : 0.0f;
                const uint32_t cnt = min(32u, R - b0);
                for (uint32_t g0 = 0; g0 < cnt; g0 += G) {
                    const int sl = (int)(g0 + lane / LP);
                    const float c = __shfl_sync(CS_FULL, cost, sl);
                    const uint32_t nd = __shfl_sync(CS_FULL, node, sl);
                    const float f1 = __shfl_sync(CS_FULL, far_t, sl);
                    const float f3 = __shfl_sync(CS_FULL, harm_t, sl);
                    float f4[DT];
#pragma unroll
                    for (int i = 0; i < DT; ++i) f4[i] = __shfl_sync(CS_FULL, bet_t[i], sl);
#pragma unroll
                    for (int q0 = 0; q0 < NQ; q0 += LP) {
                        const int q = q0 + (int)ql;
                        if (q < nq) {
                            const int i = q / 5, m = q - 5 * i;
                            if (c <= p.dist_f[i]) {
                                float v = wt;
                                if (m == 0) ++n_ri;
                                if (m == 1) v = f1;
                                if (m == 2) v = __fmul_rn(rankf[i], cycles_wt);
                                if (m == 3) v = f3;
                                if (m == 4) {
#pragma unroll
                                    for (int ii = 0; ii < DT; ++ii)
                                        if (ii == i) v = f4[ii];
                                }
                                cs_red_add(p.acc_c + (size_t)nd * p.cw + q, (double)v);
                            }
                        }
                    }
                }
            }
        }

        // ------------------------------------------------------------------ P5: dependencies, reverse settle order
        tc[4] = clock64();
        double* odw = reinterpret_cast<double*>(A.far);  // per-rank OD weight (the far queue is dead after the search)
        if (p.od_off) {
            for (uint32_t r = lane; r < R; r += 32) cs_st(&odw[r], 0.0);
            __syncwarp();
            for (unsigned long long j = __ldg(&p.od_off[si]) + lane; j < __ldg(&p.od_off[si + 1]); j += 32) {
                const uint2 dd = cs_ld(&A.ds[__ldg(&p.od_dst[j])]);
                if (dd.x != CS_INF_BITS) cs_st(&odw[dd.y], (double)__ldg(&p.od_w[j]));  // destinations are unique per origin
            }
            __syncwarp();
        }
        if (p.betweenness) {
            const double wt_d = (double)wt;
            // reverse settle order in chunks of up to 32 nodes none of which depends on another one of the same chunk
            // (minsucc): every successor of a chunk's node was finished by an earlier chunk, nothing waits
            int hi = (int)R - 1;
            while (hi >= 0) {
                const int rr = hi - (int)lane;
                // the chunk boundary (minsucc) and the per-rank state of the 32 candidates are loaded together: lanes
                // beyond the boundary simply discard what they fetched
                uint32_t ms = 0u, w = 0;
                float agg_w = 0.f;
                double sigma_w = 1.0;
                uint4 er = make_uint4(0u, 0u, 0u, 0u);
                if (rr >= 0) {
                    ms = cs_ld(&minsucc[rr]);
                    w = cs_ld(&A.s_node[rr]);
                    agg_w = cs_ld(&A.s_agg[rr]);
                    sigma_w = cs_ld(&A.sigma[rr]);
                    er = cs_ld(&A.erank[rr]);
                }
                const uint32_t badm = __ballot_sync(CS_FULL, rr < 0 || ms <= (uint32_t)hi);
                const uint32_t cnt = badm ? (uint32_t)__ffs(badm) - 1u : 32u;  // >= 1: minsucc[hi] > hi
                const bool valid = lane < cnt;
                const uint32_t r = (uint32_t)(hi - (int)lane);
                double cr[2 * DT];  // positive credits of this lane's node, slot 2 * i (plain) / 2 * i + 1 (beta-weighted)
#pragma unroll
                for (int q = 0; q < 2 * DT; ++q) cr[q] = 0.0;
                if (valid) {
                    const float cost_w = __fmul_rn(agg_w, p.speed);
                    double acc[DT], accb[DT];
#pragma unroll
                    for (int i = 0; i < DT; ++i) acc[i] = accb[i] = 0.0;
                    const uint32_t eb = er.x, deg = er.y;
                    // four in-edges at a time, each level of the dependent chain (edge record -> neighbour's map entry ->
                    // its predecessor mask -> its sigma) issued for the whole group before anything is consumed: the
                    // round trips overlap instead of adding up edge by edge
                    for (uint32_t j0 = 0; j0 < deg; j0 += 4) {
                        uint32_t xs[4], ps[4], rk[4];
                        bool cand[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            xs[u] = w;
                            ps[u] = 0;
                            if (j0 + u < deg) {
                                const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&p.g.in_rec[eb + j0 + u]));
                                xs[u] = raw.x;
                                ps[u] = raw.w & 0xffu;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            cand[u] = false;
                            rk[u] = 0;
                            if (xs[u] != w) {
                                const uint2 dx = cs_ld(&A.ds[xs[u]]);
                                cand[u] = dx.x != CS_INF_BITS && dx.y > r;
                                rk[u] = dx.y;
                            }
                        }
                        uint32_t pm[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) pm[u] = cand[u] ? cs_ld(&A.predmask[rk[u]]) : 0u;
                        double sx[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            cand[u] = cand[u] && ((pm[u] >> ps[u]) & 1u);  // x continues shortest paths through w (:861-866)
                            sx[u] = cand[u] ? cs_ld(&A.sigma[rk[u]]) : 1.0;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (cand[u]) {
                                const double f = (sx[u] == sigma_w) ? 1.0 : sigma_w / sx[u];
                                const double* dxp = A.dep + (size_t)rk[u] * D2;
#pragma unroll
                                for (int i = 0; i < DT; ++i) {
                                    if (i < D) {
                                        acc[i] += f * cs_ld(&dxp[i]);
                                        accb[i] += f * cs_ld(&dxp[D + i]);
                                    }
                                }
                            }
                        }
                    }
                    const bool is_src = (w == src);
                    const double pc = is_src ? 0.0 : p.od_off ? cs_ld(&odw[r]) : (__ldg(&p.eligible[w]) ? 0.5 : 1.0);
                    double* dr = A.dep + (size_t)r * D2;
#pragma unroll
                    for (int i = 0; i < DT; ++i) {
                        if (i < D) {
                            double seed = 0.0, seedb = 0.0;
                            if (!is_src && cost_w <= p.dist_f[i]) {
                                seed = pc;
                                seedb = pc * exp(-p.beta_d[i] * (double)cost_w);
                            }
                            const double dpn = seed + acc[i], dpb = seedb + accb[i];
                            cs_st(&dr[i], dpn);
                            cs_st(&dr[D + i], dpb);
                            if (!is_src) {
                                const double credit = dpn - seed, creditb = dpb - seedb;
                                if (credit > 0.0 || creditb > 0.0) {
                                    ++n_ci;
                                    if (credit > 0.0) cr[2 * i] = credit * wt_d;
                                    if (creditb > 0.0) cr[2 * i + 1] = creditb * wt_d;
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                // packed credit scatter: 32/LPB nodes per warp instruction, LPB consecutive doubles each
                {
                    constexpr int NQB = 2 * DT;
                    constexpr int LPB = NQB <= 2 ? 2 : NQB <= 4 ? 4 : NQB <= 8 ? 8 : NQB <= 16 ? 16 : 32;
                    constexpr int GB = 32 / LPB;
                    const int q = (int)(lane & (LPB - 1));
                    for (uint32_t g0 = 0; g0 < cnt; g0 += GB) {
                        const int sl = (int)(g0 + lane / LPB);
                        const uint32_t nd = __shfl_sync(CS_FULL, w, sl);
                        double v = 0.0;
#pragma unroll
                        for (int qq = 0; qq < NQB; ++qq) {
                            const double t = __shfl_sync(CS_FULL, cr[qq], sl);
                            if (q == qq) v = t;
                        }
                        if (v > 0.0) cs_red_add(p.acc_b + (size_t)nd * p.bw + q, v);
                    }
                }
                hi -= (int)cnt;
            }
        }

        // ------------------------------------------------------------------ P6: reset the dense map
        tc[5] = clock64();
        cs_p6_reset(A, R);
        tc[6] = clock64();

        edge_iters = cs_warp_sum(edge_iters);
        relax = cs_warp_sum(relax);
        n_ri = cs_warp_sum(n_ri);
        n_ci = cs_warp_sum(n_ci);
        if (lane == 0) {
            atomicAdd(&p.counters[CS_C_SOURCES], 1ull);
            atomicAdd(&p.counters[CS_C_SETTLED], (unsigned long long)R);
            atomicAdd(&p.counters[CS_C_EDGE_ITERS], edge_iters);
            atomicAdd(&p.counters[CS_C_RELAX], relax);
            if (n_ri) atomicAdd(&p.counters[CS_C_SUM_RI], n_ri);
            if (n_ci) atomicAdd(&p.counters[CS_C_SUM_CI], n_ci);
            atomicAdd(&p.counters[CS_C_PROGRESS], 1ull);
#pragma unroll
            for (int k = 0; k < 6; ++k) atomicAdd(&p.counters[CS_C_PHASE0 + k], (unsigned long long)(tc[k + 1] - tc[k]));
        }
    }
}

// Row widths (in doubles) of the node-interleaved accumulators for the kernel instantiation that serves D thresholds.
static inline int cs_shortest_dt(int D) { return D <= 4 ? D : D <= 8 ? 8 : CS_MAX_THRESHOLDS; }
static inline int cs_shortest_cw(int D) {
    const int nq = 5 * cs_shortest_dt(D);
    return nq <= 16 ? 16 : (nq + 31) / 32 * 32;
}
static inline int cs_shortest_bw(int D) {
    const int nq = 2 * cs_shortest_dt(D);
    return nq <= 2 ? 2 : nq <= 4 ? 4 : nq <= 8 ? 8 : nq <= 16 ? 16 : 32;
}

// Epilogue: node-interleaved accumulators -> the reference's [7][D][node_bound] layout (density, farness, cycles,
// harmonic, beta, betweenness, betweenness_beta; centrality.rs:152-215).  `add` != 0 accumulates into `out`.
__global__ void cs_k_epilogue_shortest(const double* __restrict__ acc_c, const double* __restrict__ acc_b, double* out,
                                       uint32_t n, int D, int cw, int bw, int closeness, int betweenness, int add) {
    const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n) return;
    if (closeness) {
        const double* row = acc_c + (size_t)node * cw;
        for (int i = 0; i < D; ++i)
            for (int m = 0; m < 5; ++m) {
                double* o = out + ((size_t)(m * D + i)) * n + node;
                const double v = row[5 * i + m];
                *o = add ? *o + v : v;
            }
    } else if (!add) {
        for (int q = 0; q < 5 * D; ++q) out[(size_t)q * n + node] = 0.0;
    }
    if (betweenness) {
        const double* row = acc_b + (size_t)node * bw;
        for (int i = 0; i < D; ++i)
            for (int b = 0; b < 2; ++b) {
                double* o = out + ((size_t)((5 + b) * D + i)) * n + node;
                const double v = row[2 * i + b];
                *o = add ? *o + v : v;
            }
    } else if (!add) {
        for (int q = 5 * D; q < 7 * D; ++q) out[(size_t)q * n + node] = 0.0;
    }
}
