// centrality_shortest on the GPU: one warp per source, persistent grid, all per-source scratch in a per-warp arena.
//
// Phases per source (reference: /root/reference/rust/src/centrality.rs):
//   P1  distance-capped label-correcting search over INCOMING edges (near/far buckets, 32 frontier nodes per step).
//       f32 `+` is monotone, so the fixed point equals the reference's Dijkstra distances bit for bit (:1363-1442).
//   P2  exact settle order: counting sort on quantised seconds, then exact ranking inside each bin by (seconds, node).
//   P3  predecessor sets + sigma by pulling over OUTGOING edges in settle order through the reference's epsilon rule
//       (:1413-1437), or the tolerance rule of phase 2 (:1447-1483); circuit-rank counts (:470-526).
//   P4  closeness terms formed in f32, widened to f64, scattered to the TARGET with red.global.add.f64 (:1733-1780).
//   P5  Brandes dependencies in reverse settle order, all D thresholds at once, pull form (:823-873, :1783-1835).
//   P6  reset the touched entries of the dense per-warp distance map.
#pragma once
#include "cs_common.cuh"

struct CsArenaLayout {
    size_t ds, node_list, qa, qb, far, s_node, s_agg, predmask, sigma, dep, bdone, stride;
    uint32_t rcap, qcap;
};

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
    double* out;  // [7][D][n]
    unsigned long long* counters;
    int* error;
    uint8_t* arena;
    CsArenaLayout lay;
    float delta, bin_scale;
    float* dump_agg;      // optional per-node dumps (single-source debug search)
    double* dump_sigma;
    uint32_t* dump_npred;
};

__global__ void cs_k_init_ds(uint2* ds, size_t n_total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n_total; i += stride) ds[i] = make_uint2(CS_INF_BITS, CS_NOSLOT);
}

__global__ void __launch_bounds__(CS_WARPS_PER_CTA * 32, 2) cs_k_shortest(const CsShortestParams p) {
    __shared__ uint32_t s_bins_all[CS_WARPS_PER_CTA][CS_NBINS];
    __shared__ uint32_t s_hist_all[CS_WARPS_PER_CTA][2][CS_MAX_THRESHOLDS];
    __shared__ float s_rank_all[CS_WARPS_PER_CTA][CS_MAX_THRESHOLDS];

    const uint32_t lane = cs_lane();
    const uint32_t wic = threadIdx.x >> 5;
    const uint32_t worker = blockIdx.x * CS_WARPS_PER_CTA + wic;
    const uint32_t ltmask = cs_lanemask_lt();
    uint32_t* bins = s_bins_all[wic];
    uint32_t* histN = s_hist_all[wic][0];
    uint32_t* histE = s_hist_all[wic][1];
    float* rankf = s_rank_all[wic];

    uint8_t* base = p.arena + (size_t)worker * p.lay.stride;
    uint2* ds = reinterpret_cast<uint2*>(base + p.lay.ds);
    uint32_t* node_list = reinterpret_cast<uint32_t*>(base + p.lay.node_list);
    uint2* qa = reinterpret_cast<uint2*>(base + p.lay.qa);
    uint2* qb = reinterpret_cast<uint2*>(base + p.lay.qb);
    uint2* far = reinterpret_cast<uint2*>(base + p.lay.far);
    unsigned long long* tmp_key = reinterpret_cast<unsigned long long*>(base + p.lay.qa);  // aliases qa after P1
    uint32_t* s_node = reinterpret_cast<uint32_t*>(base + p.lay.s_node);
    float* s_agg = reinterpret_cast<float*>(base + p.lay.s_agg);
    uint32_t* predmask = reinterpret_cast<uint32_t*>(base + p.lay.predmask);
    double* sigma = reinterpret_cast<double*>(base + p.lay.sigma);
    double* dep = reinterpret_cast<double*>(base + p.lay.dep);
    uint8_t* bdone = base + p.lay.bdone;
    const uint32_t rcap = p.lay.rcap, qcap = p.lay.qcap;
    const int D = p.D;
    const size_t n = p.g.n;
    const float one_minus = 1.0f - CS_TIE_EPS, one_plus = 1.0f + CS_TIE_EPS;
    const float one_plus_tol = 1.0f + p.tol;

    for (;;) {
        unsigned long long si = 0;
        if (lane == 0) si = atomicAdd(&p.counters[CS_C_NEXT], 1ull);
        si = __shfl_sync(CS_FULL, si, 0);
        if (si >= p.n_sources) break;
        if (*reinterpret_cast<volatile int*>(p.error) != 0) break;
        const uint32_t src = __ldg(&p.sources[si]);
        const float wt = __ldg(&p.src_wt[si]);

        // ------------------------------------------------------------------ P1: capped search (incoming edges)
        uint2* qc = qa;
        uint2* qn = qb;
        uint32_t nc = 1, nn = 0, nf = 0, count = 1;
        float thr = p.delta;
        unsigned long long relax = 0;
        bool fail = false;
        if (lane == 0) {
            cs_st(&ds[src], make_uint2(0u, CS_NOSLOT));
            cs_st(&node_list[0], src);
            cs_st(&qc[0], make_uint2(src, 0u));
        }
        __syncwarp();
        for (;;) {
            while (nc > 0) {
                for (uint32_t b0 = 0; b0 < nc; b0 += 32) {
                    const uint32_t idx = b0 + lane;
                    bool valid = idx < nc;
                    uint32_t v = 0, abits = 0;
                    if (valid) {
                        const uint2 it = cs_ld(&qc[idx]);
                        v = it.x;
                        abits = it.y;
                        valid = cs_ld(&ds[v].x) == abits;  // stale entries were superseded by a smaller distance
                    }
                    uint32_t eb = 0, deg = 0;
                    if (valid) {
                        eb = __ldg(&p.g.in_off[v]);
                        deg = __ldg(&p.g.in_off[v + 1]) - eb;
                    }
                    const uint32_t maxdeg = __reduce_max_sync(CS_FULL, deg);
                    const float a = __uint_as_float(abits);
                    for (uint32_t j = 0; j < maxdeg; ++j) {
                        bool improved = false, first = false;
                        uint32_t nb = 0, cbits = 0;
                        float cand = 0.f;
                        if (j < deg) {
                            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&p.g.in_rec[eb + j]));
                            nb = raw.x;
                            cand = __fadd_rn(a, __uint_as_float(raw.y));
                            if (nb != v && !(cand > p.max_seconds)) {
                                cbits = __float_as_uint(cand);
                                const uint32_t old = atomicMin(&ds[nb].x, cbits);
                                improved = cbits < old;
                                first = old == CS_INF_BITS;
                            }
                        }
                        uint32_t m = __ballot_sync(CS_FULL, first);
                        if (m) {
                            const uint32_t pos = count + __popc(m & ltmask);
                            if (first && pos < rcap) cs_st(&node_list[pos], nb);
                            count += __popc(m);
                        }
                        const bool pn = improved && (cand < thr);
                        const bool pf = improved && !pn;
                        m = __ballot_sync(CS_FULL, pn);
                        if (m) {
                            const uint32_t pos = nn + __popc(m & ltmask);
                            if (pn && pos < qcap) cs_st(&qn[pos], make_uint2(nb, cbits));
                            nn += __popc(m);
                        }
                        m = __ballot_sync(CS_FULL, pf);
                        if (m) {
                            const uint32_t pos = nf + __popc(m & ltmask);
                            if (pf && pos < qcap) cs_st(&far[pos], make_uint2(nb, cbits));
                            nf += __popc(m);
                        }
                        relax += improved ? 1ull : 0ull;
                    }
                }
                if (count > rcap || nn > qcap || nf > qcap) {
                    fail = true;
                    break;
                }
                uint2* t = qc;
                qc = qn;
                qn = t;
                nc = nn;
                nn = 0;
                __syncwarp();
            }
            if (fail || nf == 0) break;
            // near bucket exhausted: advance the threshold past the smallest live far entry and split the far pile
            float mn = __uint_as_float(CS_INF_BITS);
            for (uint32_t i = lane; i < nf; i += 32) {
                const uint2 it = cs_ld(&far[i]);
                if (cs_ld(&ds[it.x].x) == it.y) mn = fminf(mn, __uint_as_float(it.y));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(CS_FULL, mn, o));
            if (!(mn < __uint_as_float(CS_INF_BITS))) {
                nf = 0;
                break;
            }
            thr = mn + p.delta;
            uint32_t w = 0;
            nc = 0;
            for (uint32_t b0 = 0; b0 < nf; b0 += 32) {
                const uint32_t idx = b0 + lane;
                bool livee = idx < nf;
                uint2 it = make_uint2(0u, 0u);
                if (livee) {
                    it = cs_ld(&far[idx]);
                    livee = cs_ld(&ds[it.x].x) == it.y;
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
        if (fail) {
            if (lane == 0) atomicCAS(p.error, 0, count > rcap ? CS_ERR_REACH_OVERFLOW : CS_ERR_QUEUE_OVERFLOW);
            break;
        }
        const uint32_t R = count;

        // ------------------------------------------------------------------ P2: exact settle order
        for (uint32_t i = lane; i < CS_NBINS; i += 32) bins[i] = 0;
        if (lane < CS_MAX_THRESHOLDS) {
            histN[lane] = 0;
            histE[lane] = 0;
        }
        __syncwarp();
        for (uint32_t i = lane; i < R; i += 32) {
            const uint32_t node = cs_ld(&node_list[i]);
            const uint32_t ab = cs_ld(&ds[node].x);
            const uint32_t bin = min((uint32_t)(CS_NBINS - 1), (uint32_t)(__uint_as_float(ab) * p.bin_scale));
            atomicAdd(&bins[bin], 1u);
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
            const uint32_t node = cs_ld(&node_list[i]);
            const uint32_t ab = cs_ld(&ds[node].x);
            const uint32_t bin = min((uint32_t)(CS_NBINS - 1), (uint32_t)(__uint_as_float(ab) * p.bin_scale));
            const uint32_t pos = atomicAdd(&bins[bin], 1u);
            // the source sorts first among zero-distance nodes (it is always the first settled state)
            const unsigned long long key = ((unsigned long long)ab << 32) | (node == src ? 0u : node + 1u);
            cs_st(&tmp_key[pos], key);
        }
        __syncwarp();
        unsigned long long edge_iters = 0;
        for (uint32_t pos = lane; pos < R; pos += 32) {
            const unsigned long long key = cs_ld(&tmp_key[pos]);
            const uint32_t ab = (uint32_t)(key >> 32);
            const uint32_t bin = min((uint32_t)(CS_NBINS - 1), (uint32_t)(__uint_as_float(ab) * p.bin_scale));
            const uint32_t start = bin ? bins[bin - 1] : 0u;
            const uint32_t end = bins[bin];
            uint32_t rank = start;
            for (uint32_t j = start; j < end; ++j) rank += (cs_ld(&tmp_key[j]) < key) ? 1u : 0u;
            const uint32_t low = (uint32_t)key;
            const uint32_t node = low ? low - 1u : src;
            cs_st(&s_node[rank], node);
            cs_st(&s_agg[rank], __uint_as_float(ab));
            cs_st(&ds[node].y, rank);
            cs_st(&sigma[rank], 0.0);
            cs_st(&bdone[rank], (uint8_t)0);
            edge_iters += __ldg(&p.g.in_off[node + 1]) - __ldg(&p.g.in_off[node]);
        }
        __syncwarp();

        // ------------------------------------------------------------------ P3: predecessors + sigma (outgoing edges)
        for (uint32_t b0 = 0; b0 < R; b0 += 32) {
            const uint32_t r = b0 + lane;
            const bool valid = r < R;
            uint32_t v = 0;
            float av = 0.f;
            // candidate arrays (local memory; typical out-degree on street graphs is 2-4)
            float cc[CS_MAX_DEGREE];
            uint32_t cu[CS_MAX_DEGREE], crk[CS_MAX_DEGREE], cj[CS_MAX_DEGREE];
            int ncand = 0;
            uint32_t pmask_c = 0;
            if (valid) {
                v = cs_ld(&s_node[r]);
                av = cs_ld(&s_agg[r]);
                const float cost_v = __fmul_rn(av, p.speed);
                if (p.closeness) {
                    int ti = 0;
                    while (ti < D && !(cost_v <= p.dist_f[ti])) ++ti;
                    if (ti < D) atomicAdd(&histN[ti], 1u);
                }
                const uint32_t eb = __ldg(&p.g.out_off[v]);
                const uint32_t deg = __ldg(&p.g.out_off[v + 1]) - eb;
                for (uint32_t j = 0; j < deg; ++j) {
                    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&p.g.out_rec[eb + j]));
                    const uint32_t u = raw.x;
                    if (u == v) continue;
                    const uint2 du = cs_ld(&ds[u]);
                    if (du.x == CS_INF_BITS) continue;
                    const float au = __uint_as_float(du.x);
                    if (p.closeness && (raw.w & 0x100u)) {
                        const float ec = fmaxf(cost_v, __fmul_rn(au, p.speed));
                        int ti = 0;
                        while (ti < D && !(ec <= p.dist_f[ti])) ++ti;
                        if (ti < D) atomicAdd(&histE[ti], 1u);
                    }
                    if (du.y >= r) continue;  // u must be settled before v
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
                if (v == src) ncand = 0;  // the source is settled first and never acquires predecessors
                if (!p.phase2) {
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
                uint32_t amask = 0;
                for (uint32_t mm = pmask_c; mm; mm &= mm - 1) amask |= 1u << (cj[__ffs(mm) - 1] & 0xffu);
                cs_st(&predmask[r], amask);
                if (v == src) cs_st(&sigma[r], 1.0);
                if (p.dump_npred) p.dump_npred[v] = __popc(pmask_c);
            }
            bool pending = valid && v != src;
            for (;;) {
                if (pending) {
                    double s = 0.0;
                    bool ok = true;
                    for (uint32_t mm = pmask_c; mm; mm &= mm - 1) {
                        const double sg = cs_ld(&sigma[crk[__ffs(mm) - 1]]);
                        if (sg == 0.0) {
                            ok = false;
                            break;
                        }
                        s += sg;
                    }
                    if (ok) {
                        cs_st(&sigma[r], s);
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
                const uint32_t node = cs_ld(&s_node[r]);
                p.dump_agg[node] = cs_ld(&s_agg[r]);
                p.dump_sigma[node] = cs_ld(&sigma[r]);
            }
        }

        // ------------------------------------------------------------------ P4: closeness scatter to targets
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
                if (lane == 0 || true) {
                    // reachable targets per threshold exclude the source itself
                    atomicAdd(&p.counters[CS_C_REACH0 + lane], (unsigned long long)(ncount > 0 ? ncount - 1 : 0));
                }
            }
            __syncwarp();
            const float cycles_wt = __fdiv_rn(wt, __ldg(&p.g.weight[src]));  // centrality.rs:1730
            const double wt_d = (double)wt;
            for (uint32_t r = lane; r < R; r += 32) {
                const uint32_t node = cs_ld(&s_node[r]);
                if (node == src) continue;
                const float cost = __fmul_rn(cs_ld(&s_agg[r]), p.speed);
                for (int i = 0; i < D; ++i) {
                    if (cost <= p.dist_f[i]) {
                        double* o = p.out + (size_t)i * n + node;
                        const size_t ms = (size_t)D * n;
                        cs_red_add(o, wt_d);
                        cs_red_add(o + ms, (double)__fmul_rn(cost, wt));
                        cs_red_add(o + 2 * ms, (double)__fmul_rn(rankf[i], cycles_wt));
                        cs_red_add(o + 3 * ms, (double)__fmul_rn(__fdiv_rn(1.0f, cost), wt));
                        cs_red_add(o + 4 * ms, (double)__fmul_rn(expf(__fmul_rn(-p.beta_f[i], cost)), wt));
                        ++n_ri;
                    }
                }
            }
        }

        // ------------------------------------------------------------------ P5: dependencies, reverse settle order
        if (p.betweenness) {
            const double wt_d = (double)wt;
            const int D2 = 2 * D;
            for (int b0 = (int)((R - 1) & ~31u); b0 >= 0; b0 -= 32) {
                const uint32_t r = (uint32_t)b0 + lane;
                const bool valid = r < R;
                uint32_t w = 0;
                uint32_t srk[CS_MAX_DEGREE];
                int nsucc = 0;
                double sigma_w = 1.0;
                float cost_w = 0.f;
                if (valid) {
                    w = cs_ld(&s_node[r]);
                    cost_w = __fmul_rn(cs_ld(&s_agg[r]), p.speed);
                    sigma_w = cs_ld(&sigma[r]);
                    const uint32_t eb = __ldg(&p.g.in_off[w]);
                    const uint32_t deg = __ldg(&p.g.in_off[w + 1]) - eb;
                    for (uint32_t j = 0; j < deg; ++j) {
                        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&p.g.in_rec[eb + j]));
                        const uint32_t x = raw.x;
                        if (x == w) continue;
                        const uint2 dx = cs_ld(&ds[x]);
                        if (dx.x == CS_INF_BITS || dx.y <= r) continue;
                        if ((cs_ld(&predmask[dx.y]) >> (raw.w & 0xffu)) & 1u) srk[nsucc++] = dx.y;
                    }
                }
                bool pending = valid;
                for (;;) {
                    if (pending) {
                        bool ok = true;
                        for (int k = 0; k < nsucc; ++k) ok = ok && (cs_ld(&bdone[srk[k]]) != 0);
                        if (ok) {
                            const double pc = (w == src) ? 0.0 : (__ldg(&p.eligible[w]) ? 0.5 : 1.0);
                            for (int i = 0; i < D; ++i) {
                                double acc = 0.0, accb = 0.0;
                                for (int k = 0; k < nsucc; ++k) {
                                    const double f = sigma_w / cs_ld(&sigma[srk[k]]);
                                    acc += f * cs_ld(&dep[(size_t)srk[k] * D2 + i]);
                                    accb += f * cs_ld(&dep[(size_t)srk[k] * D2 + D + i]);
                                }
                                double seed = 0.0, seedb = 0.0;
                                if (w != src && cost_w <= p.dist_f[i]) {
                                    seed = pc;
                                    seedb = pc * exp(-p.beta_d[i] * (double)cost_w);
                                }
                                const double dpn = seed + acc, dpb = seedb + accb;
                                cs_st(&dep[(size_t)r * D2 + i], dpn);
                                cs_st(&dep[(size_t)r * D2 + D + i], dpb);
                                if (w != src) {
                                    const double credit = dpn - seed, creditb = dpb - seedb;
                                    if (credit > 0.0 || creditb > 0.0) {
                                        ++n_ci;
                                        if (credit > 0.0) cs_red_add(p.out + ((size_t)(5 * D + i)) * n + w, credit * wt_d);
                                        if (creditb > 0.0) cs_red_add(p.out + ((size_t)(6 * D + i)) * n + w, creditb * wt_d);
                                    }
                                }
                            }
                            cs_st(&bdone[r], (uint8_t)1);
                            pending = false;
                        }
                    }
                    __syncwarp();
                    if (!__any_sync(CS_FULL, pending)) break;
                }
            }
        }

        // ------------------------------------------------------------------ P6: reset the dense map
        for (uint32_t r = lane; r < R; r += 32) cs_st(&ds[cs_ld(&s_node[r])], make_uint2(CS_INF_BITS, CS_NOSLOT));
        __syncwarp();

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
        }
    }
}
