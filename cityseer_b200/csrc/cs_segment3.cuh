// segment_centrality, chain-contracted kernel: one warp per source; search, settle order, tree and subtree sums over
// JUNCTIONS only, chain interiors produced by walking the contiguous per-chain arrays of cs_api_v3.inl (same graph copy as
// cs_shortest3.cuh, plus per-piece length / impedance).  Reference: /root/reference/rust/src/centrality.rs:1523-1611 (tree
// search), :2198-2317 (closeness integrals over every visited edge, accumulated at the source), :2319-2402 (per-target
// area-under-curve terms added to every tree ancestor = a subtree sum).  Same f32 arithmetic as cs_segment.cuh:
//   * distances: sequential f32 additions along a chain = the reference's node-by-node additions (:1582), bit-identical;
//   * tree: an interior's predecessor is its neighbour on its own wave's side; a junction's predecessor is the neighbour
//     whose candidate equals its final seconds exactly (strict `<`, :1589).  Whenever two candidates tie exactly - across
//     the meeting piece of a chain or at a junction - the winner depends on pop order: the source is set aside
//     (redo_list) and served by the node-level kernel's heap-order replay (cs_segment.cuh), so this kernel never guesses;
//   * closeness: every chain is evaluated once by its owner end (the earlier-settled junction); the walk emits one record
//     per visited piece {cost of the visiting node, cost of the other end or inf, piece index} into shared memory and the
//     integrals run with one lane per piece, accumulated in registers -> 3 f64 adds per threshold at [m][i][src];
//   * betweenness: reverse settle order over junctions, chunks without internal dependencies (minsucc) as in
//     cs_shortest3.cuh; along a chain the subtree sum is a running sum of the nodes' own terms.
// Self-loops are not part of the contracted copy: graphs that have any are served by cs_segment.cuh.
#pragma once
#include "cs_segment.cuh"
#include "cs_shortest3.cuh"

// CTA shape: one phase-synchronous CTA per SM like the other chain kernel, but this one is bound by the arithmetic of the
// integrals (logf and two libm-exact expf per piece side and threshold), not by dependent loads: resident warps pay even at
// the price of register spills.  Measured on cfg #4 (400/800/1600 m): 16 warps (128 registers) 1.26 M sources/s, 20 (96)
// 1.30 M, 24 (80) 1.36 M, 28 (72) 1.38 M.  Shared memory per warp grows with the threshold count (outflow accumulators).
#ifndef CS3S_WARPS
#define CS3S_WARPS 28  // up to four thresholds; the widest shape decides the arena's worker count
#endif
template <int DT>
__host__ __device__ constexpr uint32_t cs3s_warps() { return DT <= 4 ? CS3S_WARPS : DT <= 8 ? 24u : 20u; }
#define CS3S_NBP 128u  // staged pieces per closeness sub-iteration (12-byte records in the 2 KB region A)

struct CsSegment3Params {
    CsV3Graph g;
    const float* clen;  // per chain piece, block layout of g.csec: length of the edge LEAVING the visiting node
    const float* cimp;  //                                          impedance factor of that edge
    int D, closeness, betweenness;
    float dist_f[CS_MAX_THRESHOLDS];
    float beta_f[CS_MAX_THRESHOLDS];
    float max_seconds, speed;
    const uint32_t* sources;  // ORIGINAL indices
    unsigned long long n_sources;
    double* out;    // [4][D][n] by ORIGINAL index: rows 0..2 (closeness) receive the source's sums directly
    double* acc_b;  // [D][n] by new id: segment betweenness, permuted into out[3] by cs_k_epilogue_segment3
    unsigned long long* counters;
    int* error;
    uint8_t* arena;
    CsArenaLayout lay;  // the chain kernel's layout (kind 3)
    float delta, bin_scale;
    uint32_t* redo_list;  // sources with exactly tied tree parents (ORIGINAL indices), count in counters[CS_C_FALLBACK]
};

template <int DT>
__host__ __device__ constexpr uint32_t cs3s_nb() { return DT <= 3 ? 112u : DT == 4 ? 80u : DT <= 8 ? 40u : 20u; }

// area under the decay curve over the origin and last segments of the src -> node route (centrality.rs:2363-2391)
template <int DT>
__device__ __forceinline__ void cs3s_auc(const CsSegment3Params& p, float sd, float o_len, float l_len, float* own) {
    const float ms = __fsub_rn(__fsub_rn(sd, o_len), l_len);
    const float o2 = __fadd_rn(ms, o_len), l2 = __fadd_rn(ms, l_len);
#pragma unroll
    for (int i = 0; i < DT; ++i) {
        own[i] = 0.0f;
        if (i < p.D && ms <= p.dist_f[i]) {
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
            if (isfinite(auc) && auc >= 0.0f) own[i] = auc;
        }
    }
}

template <int DT>
__global__ void __launch_bounds__(cs3s_warps<DT>() * 32, 1) cs_k_segment3(const CsSegment3Params p) {
    constexpr uint32_t WARPS = cs3s_warps<DT>();
    constexpr uint32_t NB = cs3s_nb<DT>();  // staged nodes per sub-iteration of the subtree pass
    // per-warp shared memory: region A (2 KB): P1 task table | P2 bins | S3 piece records | S5 node ids, costs, lengths;
    // region B (4 KB): S3 chain-block cells | S5 credits (f64) and own terms (f32); region C: link lists, link bytes, outflow
    constexpr uint32_t BYTES_A = CS3_NBINS * 4, BYTES_B = 8 * 32 * 16;
    constexpr uint32_t BYTES_C = DT * 32 * 8 + 512 + 256 + 256;
    constexpr uint32_t BYTES_W = BYTES_A + BYTES_B + BYTES_C;
    static_assert(CS3S_NBP * 12 <= BYTES_A && 4 * NB * 4 <= BYTES_A && NB >= CS3_KMAX, "region A");
    static_assert(DT * NB * 8 + DT * NB * 4 <= BYTES_B, "region B");
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ unsigned long long s_base;
    __shared__ int s_err;

    const uint32_t lane = cs_lane();
    const uint32_t ltmask = cs_lanemask_lt();
    const uint32_t wic = cs_warp_in_cta();
    const uint32_t worker = blockIdx.x * WARPS + wic;
    uint8_t* s_warp = s_dyn + (size_t)wic * BYTES_W;
    uint32_t* bins = reinterpret_cast<uint32_t*>(s_warp);
    uint16_t* ttab = reinterpret_cast<uint16_t*>(bins);
    float* pc_a = reinterpret_cast<float*>(bins);             // S3 piece records: cost of the visiting node
    float* pc_b = pc_a + CS3S_NBP;                            //                   cost of the other end (inf = not reached)
    uint32_t* pc_g = reinterpret_cast<uint32_t*>(pc_b + CS3S_NBP);  //             piece index into clen / cimp
    uint32_t* s_ids = bins;                                   // S5 staged nodes
    float* s_cst = reinterpret_cast<float*>(bins + NB);
    float* s_ll = reinterpret_cast<float*>(bins + 2 * NB);
    float* s_ol = reinterpret_cast<float*>(bins + 3 * NB);
    float* cblk = reinterpret_cast<float*>(s_warp + BYTES_A) + lane * 4;  // the lane's column of 16-byte cells
    double* s_crd = reinterpret_cast<double*>(s_warp + BYTES_A);
    float* s_own = reinterpret_cast<float*>(s_warp + BYTES_A + DT * NB * 8);
    double* s_acc = reinterpret_cast<double*>(s_warp + BYTES_A + BYTES_B);
    uint16_t* s_llist = reinterpret_cast<uint16_t*>(s_warp + BYTES_A + BYTES_B + DT * 32 * 8);
    uint8_t* s_llist2 = s_warp + BYTES_A + BYTES_B + DT * 32 * 8 + 512;
    uint8_t* s_info = s_warp + BYTES_A + BYTES_B + DT * 32 * 8 + 512 + 256;
#define CS3_CB(i) cblk[((i) >> 2) * 128u + ((i) & 3u)]
    const CsWarpArena A = cs_arena(p.arena, p.lay, worker);
    uint8_t* linfo = A.bdone;         // [rcap][8] own-side interiors per link
    uint32_t* minsucc = A.node_list;  // [rcap] after P2: smallest rank whose tree parent hangs off this junction
    // [rcap][8] per link: {candidate seconds bits for the junction (inf = none), first piece length, last piece length,
    // rank of the junction at the other end | its link index << 28}
    uint4* cand = reinterpret_cast<uint4*>(p.arena + (size_t)worker * p.lay.stride + p.lay.frank);
    uint32_t* needm = reinterpret_cast<uint32_t*>(p.arena + (size_t)worker * p.lay.stride + p.lay.needm);
    uint2* jrank = reinterpret_cast<uint2*>(p.arena + (size_t)worker * p.lay.stride + p.lay.jrank);
    float2* seglen = reinterpret_cast<float2*>(A.sigma);  // [rcap] {origin segment length (-1 = pending), last segment length}
    double* dep = A.dep;                                   // [rcap][D] own term + subtree sum
    const CsV3Graph& g = p.g;
    const uint32_t J = g.J;
    const int D = DT <= 4 ? DT : p.D;
    const uint32_t INF = CS_INF_BITS;
    const float f_inf = __uint_as_float(CS_INF_BITS);

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_base = atomicAdd(&p.counters[CS_C_NEXT], (unsigned long long)WARPS);
            s_err = *reinterpret_cast<volatile int*>(p.error);
        }
        __syncthreads();
        if (s_base >= p.n_sources || s_err != 0) break;
        const unsigned long long si = s_base + wic;
        bool run = si < p.n_sources;
        const uint32_t src_orig = run ? __ldg(&p.sources[si]) : 0u;
        CsSrc3 S;
        S.id = run ? __ldg(&g.new_of_orig[src_orig]) : 0u;  // (cs_uni broadcasts of the per-source values: no gain here)
        S.interior = S.id >= J ? 1u : 0u;
        S.slot = S.interior ? J : S.id;
        S.soff = S.ibase = S.k = S.p = S.A = S.B = S.posA = S.posB = 0;
        if (S.interior) {
            const uint32_t c = __ldg(&g.int_chain[S.id - J]);
            const uint4 c0 = __ldg(&g.ctab[2 * c]), c1 = __ldg(&g.ctab[2 * c + 1]);
            S.soff = c0.x;
            S.ibase = c0.y;
            S.k = c0.z;
            S.A = c0.w;
            S.B = c1.x;
            S.posA = c1.y;
            S.posB = c1.z;
            S.p = S.id - J - S.ibase + 1;
        }

        // ------------------------------------------------------------------ P1: junction search (as cs_shortest3.cuh)
        unsigned long long relax = 0, edge_iters = 0, n_interior = 0, n_ci = 0;
        uint32_t R = 1;
        int fail = 0;
        if (run) {
            uint2* qc = A.qa;
            uint2* qn = A.qb;
            uint2* far = A.far;
            uint32_t nc = 1, nn = 0, nf = 0;
            float thr = p.delta;
            if (lane == 0) {
                cs_st(&A.ds[S.slot], make_uint2(0u, CS_NOSLOT));
                cs_st(&A.node_list[0], S.slot);
                cs_st(&qc[0], make_uint2(S.slot, 0u));
            }
            __syncwarp();
            for (;;) {
                while (nc > 0) {
                    for (uint32_t b0 = 0; b0 < nc; b0 += 32) {
                        const uint32_t idx = b0 + lane;
                        bool valid = idx < nc;
                        uint32_t v = 0, abits = 0, skip = 0xffffffffu;
                        uint32_t off = 0, deg = 0;
                        if (valid) {
                            const uint2 it = cs_ld(&qc[idx]);
                            v = it.x & CS_NODE_MASK;
                            skip = (it.x >> CS_NODE_BITS) - 1u;
                            abits = it.y;
                            const uint32_t cur = cs_ld(&A.ds[v].x);
                            uint2 ji = make_uint2(0u, 2u);
                            if (v != J) ji = __ldg(&g.jinfo[v]);
                            valid = cur == abits;
                            off = ji.x;
                            deg = valid ? (ji.y & 0xffu) : 0u;
                        }
                        const uint32_t cnt = deg - (skip < deg ? 1u : 0u);
                        uint32_t incl = cnt;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t up = __shfl_up_sync(CS_FULL, incl, o);
                            if (lane >= (uint32_t)o) incl += up;
                        }
                        const uint32_t total = __shfl_sync(CS_FULL, incl, 31);
                        for (uint32_t q = 0, j = 0; q < cnt; ++q, ++j) {
                            if (j == skip) ++j;
                            ttab[incl - cnt + q] = (uint16_t)(lane | (j << 8));
                        }
                        __syncwarp();
                        for (uint32_t t0 = 0; t0 < total; t0 += 32) {
                            const bool has = t0 + lane < total;
                            const uint32_t te = has ? ttab[t0 + lane] : 0u;
                            const uint32_t tv = __shfl_sync(CS_FULL, v, te & 31u);
                            const uint32_t tab = __shfl_sync(CS_FULL, abits, te & 31u);
                            const uint32_t toff = __shfl_sync(CS_FULL, off, te & 31u);
                            bool improved = false, first = false;
                            uint32_t nb = 0, cbits = 0, back = 0;
                            float cnd = 0.f;
                            if (has) {
                                const CsView V = cs3_view(g, S, tv, toff, te >> 8);
                                cs3_load_block(g, V, cblk, V.sv, V.k + 1);
                                float a = __uint_as_float(tab);
                                bool ok = true;
                                for (uint32_t t = 0; t <= V.k; ++t) {
                                    a = __fadd_rn(a, CS3_CB(V.sv + t));
                                    if (a > p.max_seconds) {
                                        ok = false;
                                        break;
                                    }
                                }
                                if (ok) {
                                    nb = V.far;
                                    cnd = a;
                                    cbits = __float_as_uint(a);
                                    const uint32_t old = atomicMin(&A.ds[nb].x, cbits);
                                    improved = cbits < old;
                                    first = old == INF;
                                    back = V.paf + 1u;
                                }
                            }
                            uint32_t m = __ballot_sync(CS_FULL, first);
                            if (m) {
                                const uint32_t pos = R + __popc(m & ltmask);
                                if (first && pos < A.rcap) cs_st(&A.node_list[pos], nb);
                                R += __popc(m);
                            }
                            const bool pn = improved && (cnd < thr);
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
                        __syncwarp();
                    }
                    if (R > A.rcap || nn > A.qcap || nf > A.qcap) {
                        fail = R > A.rcap ? CS_ERR_REACH_OVERFLOW : CS_ERR_QUEUE_OVERFLOW;
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
                float mn = f_inf;
                for (uint32_t i = lane; i < nf; i += 32) {
                    const uint2 it = cs_ld(&far[i]);
                    if (cs_ld(&A.ds[it.x & CS_NODE_MASK].x) == it.y) mn = fminf(mn, __uint_as_float(it.y));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(CS_FULL, mn, o));
                if (!(mn < f_inf)) break;
                thr = mn + p.delta;
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
        }
        if (fail) {
            if (lane == 0) atomicCAS(p.error, 0, fail);
            run = false;
        }
        if (!run) R = 0;
        __syncthreads();

        // ------------------------------------------------------------------ P2: exact settle order of the junctions
        if (run) {
            for (uint32_t i = lane; i < CS3_NBINS; i += 32) bins[i] = 0;
            __syncwarp();
            for (uint32_t i = lane; i < R; i += 32) {
                const uint32_t node = cs_ld(&A.node_list[i]);
                const uint32_t ab = cs_ld(&A.ds[node].x);
                cs_st(reinterpret_cast<uint32_t*>(&A.s_agg[i]), ab);
                atomicAdd(&bins[cs3_bin(ab, p.bin_scale)], 1u);
            }
            __syncwarp();
            {
                uint32_t carry = 0;
                for (uint32_t k = 0; k < CS3_NBINS / 32; ++k) {
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
                const uint32_t pos = atomicAdd(&bins[cs3_bin(ab, p.bin_scale)], 1u);
                const uint32_t key = node == S.slot ? 0u : __ldg(&g.orig_of_new[node]) + 1u;
                cs_st(&A.tmp_key[pos], ((unsigned long long)ab << 32) | key);
            }
            __syncwarp();
            for (uint32_t pos = lane; pos < R; pos += 32) {
                const unsigned long long key = cs_ld(&A.tmp_key[pos]);
                const uint32_t ab = (uint32_t)(key >> 32);
                const uint32_t bin = cs3_bin(ab, p.bin_scale);
                const uint32_t start = bin ? bins[bin - 1] : 0u;
                const uint32_t end = bins[bin];
                uint32_t rank = start;
                for (uint32_t j = start; j < end; ++j) rank += (cs_ld(&A.tmp_key[j]) < key) ? 1u : 0u;
                const uint32_t low = (uint32_t)key;
                const uint32_t node = low ? __ldg(&g.new_of_orig[low - 1u]) : S.slot;
                cs_st(&A.s_node[rank], node);
                cs_st(&A.s_agg[rank], __uint_as_float(ab));
                cs_st(&A.ds[node].y, rank);
                cs_st(&minsucc[rank], CS_NOSLOT);  // (node_list is dead from here on)
                cs_st(&needm[rank], 0u);
                cs_st(&seglen[rank], make_float2(rank == 0 ? 0.0f : -1.0f, 0.0f));
                const uint2 ji = node == J ? make_uint2(0u, 2u | (2u << 8)) : __ldg(&g.jinfo[node]);
                cs_st(&jrank[rank], ji);
                edge_iters += ji.y >> 8;
            }
            __syncwarp();
        }
        __syncthreads();

        // ------------------------------------------------------------------ S3: chains, tree, closeness integrals (forward)
        double dens[DT], harm[DT], bet[DT];
#pragma unroll
        for (int i = 0; i < DT; ++i) dens[i] = harm[i] = bet[i] = 0.0;
        bool ambiguous = false;
        for (uint32_t b0 = 0; b0 < R; b0 += 32) {
            const uint32_t r = b0 + lane;
            const bool valid = r < R;
            uint32_t v = 0, off = 0, deg = 0, vid = 0, avb = 0;
            if (valid) {
                v = cs_ld(&A.s_node[r]);
                vid = v == J ? S.id : v;
                avb = __float_as_uint(cs_ld(&A.s_agg[r]));
                const uint2 ji = cs_ld(&jrank[r]);
                off = ji.x;
                deg = ji.y & 0xffu;
            }
            uint32_t inc = deg;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(CS_FULL, inc, o);
                if ((int)lane >= o) inc += t;
            }
            const uint32_t totalL = __shfl_sync(CS_FULL, inc, 31);
            __syncwarp();
            for (uint32_t j = 0; j < deg; ++j) s_llist[inc - deg + j] = (uint16_t)(lane | (j << 5));
            __syncwarp();
            // every chain is evaluated once, by its OWNER end: the earlier-settled junction (or the only reached one)
            uint32_t nown = 0;
            for (uint32_t base = 0; base < totalL; base += 32) {
                const uint32_t e = base + lane;
                const bool act = e < totalL;
                const uint32_t code = act ? s_llist[e] : 0u;
                const uint32_t jl = code & 31u, j = code >> 5;
                const uint32_t lv = __shfl_sync(CS_FULL, v, jl), loff = __shfl_sync(CS_FULL, off, jl);
                bool own = false;
                if (act) {
                    uint32_t far, paf;
                    if (lv == J) {
                        far = j == 0 ? S.A : S.B;
                        paf = j == 0 ? S.posA : S.posB;
                    } else {
                        const uint4 L = __ldg(&g.links[loff + j]);
                        far = L.x;
                        paf = (L.w >> 5) & 15u;
                        if (S.interior && (L.w & 15u) > 0 && L.y == S.soff) {
                            far = J;
                            paf = (L.w >> 4) & 1u;
                        }
                    }
                    const uint2 dF = cs_ld(&A.ds[far]);
                    const uint32_t lr = b0 + jl;
                    own = dF.x == INF || lr < dF.y || (lr == dF.y && j < paf);
                }
                const uint32_t m = __ballot_sync(CS_FULL, own);
                if (own) s_llist2[nown + __popc(m & ltmask)] = (uint8_t)code;
                nown += __popc(m);
            }
            __syncwarp();
            for (uint32_t base = 0; base < nown; base += 32) {
                const uint32_t e = base + lane;
                const bool act = e < nown;
                const uint32_t code = act ? s_llist2[e] : 0u;
                const uint32_t jl = code & 31u, j = code >> 5;
                const uint32_t lv = __shfl_sync(CS_FULL, v, jl), loff = __shfl_sync(CS_FULL, off, jl);
                const uint32_t lvid = __shfl_sync(CS_FULL, vid, jl), lavb = __shfl_sync(CS_FULL, avb, jl);
                const uint32_t lr = b0 + jl;
                CsView V;
                V.k = V.sv = V.sF = V.blk = V.nv = V.far = V.paf = V.id1 = V.cnt = 0;
                V.step = 1;
                if (act) V = cs3_view(g, S, lv, loff, j);
                // sub-iterations: as many links as fit the piece staging area (a chain emits at most k + 1 pieces)
                uint32_t remaining = __ballot_sync(CS_FULL, act);
                while (remaining) {
                    const bool mine = (remaining >> lane) & 1u;
                    uint32_t pinc = mine ? V.k + 1u : 0u;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(CS_FULL, pinc, o);
                        if ((int)lane >= o) pinc += t;
                    }
                    const bool go = mine && pinc <= CS3S_NBP;
                    const uint32_t gom = __ballot_sync(CS_FULL, go);  // never empty: k + 1 <= 13
                    uint32_t pos = pinc - (V.k + 1u);                 // this lane's first record
                    const uint32_t pos0 = pos;
                    __syncwarp();
                    if (go) {
                        const float av = __uint_as_float(lavb);
                        const uint2 dF = cs_ld(&A.ds[V.far]);  // issued before the staging wait: the round trips overlap
                        cs3_load_block(g, V, cblk, 0, V.nv);
                        const uint32_t k = V.k;
                        const uint32_t fid = V.far == J ? S.id : V.far;
                        const bool f_reached = dF.x != INF;
                        float a = av;
                        float b = __uint_as_float(dF.x);
                        float a_next = __fadd_rn(a, CS3_CB(V.sv));
                        float b_next = f_reached ? __fadd_rn(b, CS3_CB(V.sF)) : f_inf;
                        uint32_t T = 0, jn = 0;
                        while (T + jn < k) {
                            const bool a_ok = !(a_next > p.max_seconds);
                            const bool b_ok = f_reached && !(b_next > p.max_seconds);
                            if (!a_ok && !b_ok) break;
                            bool take_a = a_ok;
                            if (a_ok && b_ok) {
                                take_a = a_next <= b_next;
                                // the last unsettled node claimed from both sides with bit-equal seconds: its tree parent is
                                // whichever side pops first
                                if (T + jn + 1 == k && a_next == b_next) ambiguous = true;
                            }
                            if (take_a) {
                                if (a_next == a) atomicCAS(p.error, 0, CS_ERR_ZERO_TIE);
                                if (p.closeness) {
                                    pc_a[pos] = __fmul_rn(a, p.speed);
                                    pc_b[pos] = __fmul_rn(a_next, p.speed);
                                    pc_g[pos] = V.blk + V.sv + T;
                                    ++pos;
                                }
                                a = a_next;
                                ++T;
                                a_next = __fadd_rn(a, CS3_CB(V.sv + T));
                            } else {
                                if (b_next == b) atomicCAS(p.error, 0, CS_ERR_ZERO_TIE);
                                if (p.closeness) {
                                    pc_a[pos] = __fmul_rn(b, p.speed);
                                    pc_b[pos] = __fmul_rn(b_next, p.speed);
                                    pc_g[pos] = V.blk + V.sF + jn;
                                    ++pos;
                                }
                                b = b_next;
                                ++jn;
                                b_next = __fadd_rn(b, CS3_CB(V.sF + jn));
                            }
                        }
                        n_interior += T + jn;
                        const uint32_t xid = T == 0 ? lvid : V.id1 + V.step * (int)(T - 1);
                        if (f_reached && T + jn == k) {
                            // the fronts met: X = m_T (v when T == 0) and Y = m_{T+1} (F when T == k) share the last piece,
                            // visited from whichever settles first (the edge record of the reference, :1571-1576)
                            const uint32_t yid = T == k ? fid : V.id1 + V.step * (int)T;
                            const bool x_later = cs3_before(g, S, __float_as_uint(b), yid, __float_as_uint(a), xid);
                            if (p.closeness) {
                                pc_a[pos] = __fmul_rn(x_later ? b : a, p.speed);
                                pc_b[pos] = __fmul_rn(x_later ? a : b, p.speed);
                                pc_g[pos] = x_later ? V.blk + V.sF + jn : V.blk + V.sv + T;
                                ++pos;
                            }
                            // a candidate across the meeting piece that equals an interior's seconds exactly competes
                            // with that interior's own-side parent in pop order
                            if (T >= 1 && __float_as_uint(b_next) == __float_as_uint(a) && !(b_next > p.max_seconds)) ambiguous = true;
                            if (T < k && __float_as_uint(a_next) == __float_as_uint(b) && !(a_next > p.max_seconds)) ambiguous = true;
                        } else if (p.closeness) {
                            // the fronts did not meet: the piece beyond each front was recorded from its settled end
                            pc_a[pos] = __fmul_rn(a, p.speed);
                            pc_b[pos] = f_inf;
                            pc_g[pos] = V.blk + V.sv + T;
                            ++pos;
                            if (f_reached) {
                                pc_a[pos] = __fmul_rn(b, p.speed);
                                pc_b[pos] = f_inf;
                                pc_g[pos] = V.blk + V.sF + jn;
                                ++pos;
                            }
                        }
                        // this end owns T interiors and gets no candidate parent from this link (it settles first)
                        cs_st(&cand[(size_t)lr * 8 + j], make_uint4(INF, 0u, 0u, dF.y));
                        cs_st(&linfo[(size_t)lr * 8 + j], (uint8_t)T);
                        if (f_reached) {
                            // the far end owns jn interiors; when this wave took the whole chain its last node (or this
                            // junction) offers the far junction a parent at a_next seconds
                            uint32_t cc = INF;
                            if (T == k && dF.y != 0 && !(a_next > p.max_seconds)) cc = __float_as_uint(a_next);
                            const float first_len = __ldg(&p.clen[V.blk + V.sv]);
                            const float last_len = __ldg(&p.clen[V.blk + V.sv + k]);
                            cs_st(&cand[(size_t)dF.y * 8 + V.paf],
                                  make_uint4(cc, __float_as_uint(first_len), __float_as_uint(last_len), lr | (j << 28)));
                            cs_st(&linfo[(size_t)dF.y * 8 + V.paf], (uint8_t)jn);
                        }
                        if (p.closeness) {
                            for (; pos < pos0 + k + 1u; ++pos) {  // a cut-off chain emits fewer records than it reserved
                                pc_a[pos] = f_inf;
                                pc_b[pos] = f_inf;
                                pc_g[pos] = V.blk;
                            }
                        }
                    }
                    // every going lane reserved k + 1 records and filled the unused ones with inert entries
                    const uint32_t last = 31u - (uint32_t)__clz(gom);
                    const uint32_t total = __shfl_sync(CS_FULL, pinc, last);
                    __syncwarp();
                    if (p.closeness) {
                        // one lane per piece (centrality.rs:2201-2317, f32 as cs_segment.cuh)
                        for (uint32_t e0 = 0; e0 < total; e0 += 32) {
                            const uint32_t en = e0 + lane;
                            const bool used = en < total;
                            if (used) {
                                const float a = pc_a[en], b = pc_b[en];
                                const uint32_t gi = pc_g[en];
                                const float len = __ldg(&p.clen[gi]);
                                const float imp = __ldg(&p.cimp[gi]);
                                const float lo_c = fminf(a, b), hi_c = fmaxf(a, b);  // a <= b except across a meeting piece
                                const float c = __fdiv_rn(__fadd_rn(__fadd_rn(len, lo_c), hi_c), 2.0f);
                                const float c_imp = __fadd_rn(lo_c, __fmul_rn(__fsub_rn(c, lo_c), imp));
#pragma unroll
                                for (int i = DT - 1; i >= 0; --i) {
                                    if (i >= D) continue;
                                    const float thr = p.dist_f[i];
                                    if (lo_c < thr) cs_seg_terms(lo_c, c, c_imp, imp, thr, p.beta_f[i], dens[i], harm[i], bet[i]);
                                    if (hi_c == c) continue;
                                    if (hi_c <= thr) cs_seg_terms(hi_c, c, c_imp, imp, thr, p.beta_f[i], dens[i], harm[i], bet[i]);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    remaining &= ~gom;
                }
            }
            __syncwarp();
            // S3b, one lane per junction: the parent is the one candidate that equals the junction's seconds exactly
            uint32_t prk = CS_NOSLOT;
            float first_len = 0.f, last_len = 0.f;
            if (valid && r != 0) {
                int nmatch = 0;
                uint32_t pj = 0;
                for (uint32_t j0 = 0; j0 < deg; j0 += 4) {  // four link records per round trip
                    uint4 grp[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        grp[t] = j0 + t < deg ? cs_ld(&cand[(size_t)r * 8 + j0 + t]) : make_uint4(INF, 0u, 0u, 0u);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint4 c4 = grp[t];
                        if (j0 + t >= deg || c4.x != avb) continue;
                        ++nmatch;
                        prk = c4.w & 0x0fffffffu;
                        pj = c4.w >> 28;
                        first_len = __uint_as_float(c4.y);
                        last_len = __uint_as_float(c4.z);
                    }
                }
                if (nmatch == 0) atomicCAS(p.error, 0, CS_ERR_ZERO_TIE);  // reached only through a zero-second tie
                if (nmatch > 1) ambiguous = true;
                if (nmatch >= 1) {
                    atomicMin(&minsucc[prk], r);
                    atomicOr(&needm[prk], 1u << pj);
                } else {
                    prk = CS_NOSLOT;
                }
            }
            bool pending = valid && r != 0 && prk != CS_NOSLOT;
            if (valid && r != 0 && prk == CS_NOSLOT) cs_st(&seglen[r], make_float2(0.0f, 0.0f));
            for (;;) {
                if (pending) {
                    float o_len = first_len;  // the parent junction is the source: the origin segment is this chain's first piece
                    bool ok = true;
                    if (prk != 0) {
                        o_len = cs_ld(&seglen[prk]).x;
                        ok = o_len >= 0.0f;
                    }
                    if (ok) {
                        cs_st(&seglen[r], make_float2(o_len, last_len));
                        pending = false;
                    }
                }
                __syncwarp();
                if (!__any_sync(CS_FULL, pending)) break;
            }
        }
        __syncwarp();
        if (run && __any_sync(CS_FULL, ambiguous)) {
            // exactly tied tree parents: leave the source to the heap-order replay, nothing of it has been accumulated
            if (lane == 0) cs_st(&p.redo_list[atomicAdd(&p.counters[CS_C_FALLBACK], 1ull)], src_orig);
            cs_p6_reset(A, R);
            run = false;
            R = 0;
        }
        if (run && p.closeness) {
            const size_t n = g.n;
#pragma unroll
            for (int i = 0; i < DT; ++i) {
                if (i < D) {
                    const double d0 = cs_warp_sum(dens[i]), d1 = cs_warp_sum(harm[i]), d2 = cs_warp_sum(bet[i]);
                    if (lane == 0) {
                        cs_red_add(p.out + ((size_t)(0 * D + i)) * n + src_orig, d0);
                        cs_red_add(p.out + ((size_t)(1 * D + i)) * n + src_orig, d1);
                        cs_red_add(p.out + ((size_t)(2 * D + i)) * n + src_orig, d2);
                    }
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------------ S5: subtree sums, reverse settle order
        if (run && p.betweenness) {
            int hi = (int)R - 1;
            while (hi >= 0) {
                const int rr = hi - (int)lane;
                // the chunk boundary (minsucc) and the per-rank state of the 32 candidate junctions are loaded together
                uint32_t ms = 0u, w = 0, off = 0, deg = 0, nm = 0;
                float aw = 0.f;
                float2 ol = make_float2(0.f, 0.f);
                unsigned long long info8 = 0ull;
                if (rr >= 0) {
                    ms = cs_ld(&minsucc[rr]);
                    w = cs_ld(&A.s_node[rr]);
                    aw = cs_ld(&A.s_agg[rr]);
                    nm = cs_ld(&needm[rr]);
                    ol = cs_ld(&seglen[rr]);
                    const uint2 ji = cs_ld(&jrank[rr]);
                    off = ji.x;
                    deg = ji.y & 0xffu;
                    info8 = cs_ld(reinterpret_cast<const unsigned long long*>(linfo + (size_t)rr * 8));
                }
                const uint32_t badm = __ballot_sync(CS_FULL, rr < 0 || ms <= (uint32_t)hi);
                const uint32_t cnt = badm ? (uint32_t)__ffs(badm) - 1u : 32u;  // >= 1: minsucc[hi] > hi
                const bool valid = lane < cnt;
                const uint32_t r = (uint32_t)(hi - (int)lane);
                if (!valid) deg = 0;
                uint32_t inc = deg;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(CS_FULL, inc, o);
                    if ((int)lane >= o) inc += t;
                }
                const uint32_t totalL = __shfl_sync(CS_FULL, inc, 31);
                __syncwarp();
                if (valid) {
                    *reinterpret_cast<unsigned long long*>(s_info + lane * 8) = info8;
                    for (uint32_t j = 0; j < deg; ++j) s_llist[inc - deg + j] = (uint16_t)(lane | (j << 5));
                }
#pragma unroll
                for (int i = 0; i < DT; ++i) s_acc[i * 32 + lane] = 0.0;
                __syncwarp();
                for (uint32_t base = 0; base < totalL; base += 32) {
                    const uint32_t e = base + lane;
                    const bool act = e < totalL;
                    const uint32_t code = act ? s_llist[e] : 0u;
                    const uint32_t jl = code & 31u, j = code >> 5;
                    const uint32_t lw = __shfl_sync(CS_FULL, w, jl), loff = __shfl_sync(CS_FULL, off, jl);
                    const float law = __shfl_sync(CS_FULL, aw, jl);
                    const float lol = __shfl_sync(CS_FULL, ol.x, jl);
                    const uint32_t lnm = __shfl_sync(CS_FULL, nm, jl);
                    const uint32_t lr = (uint32_t)(hi - (int)jl);
                    uint32_t T = 0;
                    bool work = false;
                    double dl[DT];  // subtree sum flowing toward the junction along this link
#pragma unroll
                    for (int i = 0; i < DT; ++i) dl[i] = 0.0;
                    CsView V;
                    V.sv = V.id1 = V.blk = 0;
                    V.step = 1;
                    float o_side = lol;  // origin segment of the nodes on this side of the link
                    if (act) {
                        T = s_info[jl * 8 + j] & 15u;
                        const bool needF = (lnm >> j) & 1u;  // the far junction's tree parent hangs on this link
                        work = T > 0 || needF;
                        if (work) V = cs3_view(g, S, lw, loff, j);
                        if (work && lr == 0) o_side = __ldg(&p.clen[V.blk + V.sv]);  // first piece from the source
                        if (needF) {
                            const uint32_t rankF = cs_ld(&cand[(size_t)lr * 8 + j].w) & 0x0fffffffu;
                            const double* dx = dep + (size_t)rankF * D;
#pragma unroll
                            for (int i = 0; i < DT; ++i)
                                if (i < D) dl[i] = cs_ld(&dx[i]);
                        }
                    }
                    uint32_t remaining = __ballot_sync(CS_FULL, act && work);
                    while (remaining) {
                        const bool mine = (remaining >> lane) & 1u;
                        uint32_t tinc = mine ? T : 0u;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t t = __shfl_up_sync(CS_FULL, tinc, o);
                            if ((int)lane >= o) tinc += t;
                        }
                        const bool go = mine && tinc <= NB;
                        const uint32_t gom = __ballot_sync(CS_FULL, go);
                        const uint32_t last = 31u - (uint32_t)__clz(gom);
                        const uint32_t total = __shfl_sync(CS_FULL, tinc, last);
                        const uint32_t offs = tinc - T;
                        __syncwarp();
                        if (go && T) {
                            float a = law;
                            const float* sec = g.csec + V.blk + V.sv;
                            const float* len = p.clen + V.blk + V.sv;
                            for (uint32_t t0 = 0; t0 < T; t0 += 4) {  // the loads of four pieces are in flight together
                                float sx[4], lx[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    sx[u] = t0 + u < T ? __ldg(sec + t0 + u) : 0.f;
                                    lx[u] = t0 + u < T ? __ldg(len + t0 + u) : 0.f;
                                }
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    if (t0 + u < T) {
                                        a = __fadd_rn(a, sx[u]);
                                        s_ids[offs + t0 + u] = V.id1 + V.step * (int)(t0 + u);
                                        s_cst[offs + t0 + u] = __fmul_rn(a, p.speed);
                                        s_ll[offs + t0 + u] = lx[u];
                                        s_ol[offs + t0 + u] = o_side;
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        // own terms, one lane per node: only targets with a larger ORIGINAL index than the source (:2321)
                        for (uint32_t e0 = 0; e0 < total; e0 += 32) {
                            const uint32_t en = e0 + lane;
                            if (en < total) {
                                float own[DT];
                                if (__ldg(&g.orig_of_new[s_ids[en]]) > src_orig) {
                                    cs3s_auc<DT>(p, s_cst[en], s_ol[en], s_ll[en], own);
                                } else {
#pragma unroll
                                    for (int i = 0; i < DT; ++i) own[i] = 0.0f;
                                }
#pragma unroll
                                for (int i = 0; i < DT; ++i) s_own[i * NB + en] = own[i];
                            }
                        }
                        __syncwarp();
                        if (go) {
                            for (uint32_t t = T; t >= 1; --t) {
                                const uint32_t en = offs + t - 1;
#pragma unroll
                                for (int i = 0; i < DT; ++i) {
                                    if (i < D) {
                                        s_crd[i * NB + en] = dl[i];
                                        dl[i] += (double)s_own[i * NB + en];
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        for (uint32_t e0 = 0; e0 < total; e0 += 32) {
                            const uint32_t en = e0 + lane;
                            if (en < total) {
                                double* col = p.acc_b + s_ids[en];
#pragma unroll
                                for (int i = 0; i < DT; ++i) {
                                    if (i < D) {
                                        const double credit = s_crd[i * NB + en];
                                        if (credit > 0.0) {
                                            ++n_ci;
                                            cs_red_add(col + (size_t)i * g.n, credit);
                                        }
                                    }
                                }
                            }
                        }
                        {
                            const uint32_t key = act ? jl : 64u + lane;
                            const uint32_t kprev = __shfl_up_sync(CS_FULL, key, 1);
                            const bool head = act && (lane == 0 || kprev != key);
                            double ov[DT];
#pragma unroll
                            for (int i = 0; i < DT; ++i) ov[i] = go ? dl[i] : 0.0;
#pragma unroll
                            for (int o = 1; o < (int)CS3_MAX_LINKS; o <<= 1) {
                                const bool same = __shfl_down_sync(CS_FULL, key, o) == key && lane + o < 32u;
#pragma unroll
                                for (int q = 0; q < DT; ++q) {
                                    const double t = __shfl_down_sync(CS_FULL, ov[q], o);
                                    if (same) ov[q] += t;
                                }
                            }
                            if (head) {
#pragma unroll
                                for (int i = 0; i < DT; ++i)
                                    if (i < D) s_acc[i * 32 + jl] += ov[i];
                            }
                        }
                        __syncwarp();
                        remaining &= ~gom;
                    }
                }
                __syncwarp();
                if (valid) {
                    const bool is_src = r == 0;
                    const uint32_t wid = w == J ? S.id : w;
                    float own[DT];
#pragma unroll
                    for (int i = 0; i < DT; ++i) own[i] = 0.0f;
                    if (!is_src && __ldg(&g.orig_of_new[wid]) > src_orig) cs3s_auc<DT>(p, __fmul_rn(aw, p.speed), ol.x, ol.y, own);
                    double* dr = dep + (size_t)r * D;
                    double* col = p.acc_b + wid;
#pragma unroll
                    for (int i = 0; i < DT; ++i) {
                        if (i < D) {
                            const double sub = s_acc[i * 32 + lane];
                            cs_st(&dr[i], (double)own[i] + sub);
                            if (!is_src && sub > 0.0) {
                                ++n_ci;
                                cs_red_add(col + (size_t)i * g.n, sub);
                            }
                        }
                    }
                }
                __syncwarp();
                hi -= (int)cnt;
            }
        }
        __syncthreads();

        // ------------------------------------------------------------------ reset the dense map
        cs_p6_reset(A, R);
        edge_iters = cs_warp_sum(edge_iters);
        relax = cs_warp_sum(relax);
        n_ci = cs_warp_sum(n_ci);
        n_interior = cs_warp_sum(n_interior);
        if (lane == 0 && run) {
            atomicAdd(&p.counters[CS_C_SOURCES], 1ull);
            atomicAdd(&p.counters[CS_C_SETTLED], (unsigned long long)R + n_interior);
            atomicAdd(&p.counters[CS_C_EDGE_ITERS], edge_iters + 2ull * n_interior);
            atomicAdd(&p.counters[CS_C_RELAX], relax);
            if (n_ci) atomicAdd(&p.counters[CS_C_SUM_CI], n_ci);
            atomicAdd(&p.counters[CS_C_PROGRESS], 1ull);
        }
    }
#undef CS3_CB
}

// betweenness accumulators by new id -> row 3 of the [4][D][node_bound] result in original index order
__global__ void cs_k_epilogue_segment3(const double* __restrict__ acc_b, double* out, const uint32_t* __restrict__ orig_of_new,
                                       uint32_t n, int D) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t node = orig_of_new[v];
    for (int i = 0; i < D; ++i) out[((size_t)(3 * D + i)) * n + node] += acc_b[(size_t)i * n + v];
}

template <int DT>
static constexpr uint32_t cs3s_smem_bytes() {
    return cs3s_warps<DT>() * (CS3_NBINS * 4 + 8 * 32 * 16 + DT * 32 * 8 + 512 + 256 + 256);
}
