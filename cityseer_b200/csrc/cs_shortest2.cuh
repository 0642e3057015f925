// centrality_shortest, shared-memory kernel: one CTA per source, all latency-critical per-source state in shared memory.
//
// Same phases and the same arithmetic as cs_shortest.cuh (reference: /root/reference/rust/src/centrality.rs), but the
// per-source state no longer streams through HBM:
//   * nodes are renumbered along a Hilbert curve at upload, so the nodes one source reaches occupy few, mostly full
//     pages of 2^pb consecutive ids; a paged map in shared memory (open-addressed page table + one f32 word per slot)
//     replaces the dense per-warp distance map of N entries;
//   * P1 keeps its frontier in shared-memory queues and defers far relaxations with one bit per slot (no far pile);
//     queue items carry the neighbour's CSR offset and degree (stored in the edge record), so the only global access on
//     the search's critical path is the 16-byte edge record itself;
//   * P3 / P5 resolve their in-chunk dependencies (sigma forward, dependencies backward) by spinning on shared-memory
//     rings instead of L2/DRAM round trips;
//   * closeness and betweenness are scattered into node-interleaved accumulators with packed red.global.add.f64.
// A source that exceeds the shared-memory capacities (pages, reached nodes) is appended to a fallback list and handled
// by the global-arena kernel of cs_shortest.cuh afterwards: still the GPU, never a CPU path.
#pragma once
#include "cs_shortest.cuh"

#define CS2_T_MAX 256         // threads per CTA (one source per CTA at a time): 128 or 256, template parameter T
#define CS2_EMPTY 0xffffffffu
#define CS2_MAX_DEG 8         // in/out degree bound of this kernel (4-bit fields, 8-bit predecessor masks)
#ifndef CS2_POLL_NS
#define CS2_POLL_NS 40        // back-off of a warp that found its in-chunk dependencies still pending
#endif
#define CS2_CHAIN_HOPS 12     // degree-2 nodes a search lane walks through before handing back to the queue

struct CsV2Graph {
    uint32_t n;
    const uint4* node2;           // [n] by new id: {in_eb, out_eb, in_deg | out_deg << 8 | live << 16, weight bits}
    const CsEdge* in2;            // in-CSR by new id: {nbr, sec, in_eb of nbr, meta}
                                  //   meta[7:0] position of this edge in nbr's out-list, meta[9] self-loop,
                                  //   meta[19:16] 1 + position of the twin inside nbr's in-list (0 = none), meta[27:24] in-degree of nbr
    const CsEdge* out2;           // out-CSR by new id: {nbr, sec, -, meta}: meta[7:0] position in nbr's in-list,
                                  //   meta[8] canonical representative (circuit rank), meta[9] self-loop
    const uint32_t* orig_of_new;  // [n] original (petgraph) index of each new id: tie-break key of the settle order
};

struct CsV2Smem {
    uint32_t S, TB, pb, rcap, QC, NB, WS, WD, max_pages;
    uint32_t off_dist, off_keys, off_defer, off_rank, off_perm, off_pmask, off_u, off_dep, total;
};

struct CsShortest2Params {
    CsV2Graph g;
    CsV2Smem sm;
    int D, closeness, betweenness, phase2, probe;
    float dist_f[CS_MAX_THRESHOLDS];
    float beta_f[CS_MAX_THRESHOLDS];
    double beta_d[CS_MAX_THRESHOLDS];
    float max_seconds, speed, tol;
    const uint32_t* sources;  // new ids, ascending (spatially adjacent sources run on neighbouring CTAs)
    const float* src_wt;
    unsigned long long n_sources;
    const uint8_t* eligible;  // by new id
    double* acc_c;            // [n][cw] by new id
    double* acc_b;            // [n][bw] by new id
    int cw, bw;
    unsigned long long* counters;
    int* error;
    uint8_t* scratch;         // per CTA: s_agg f32[rcap] | sigma f64[rcap] | dep f64[rcap][2D]
    size_t scratch_stride;
    uint32_t* fallback;       // positions (into sources[]) of the sources this kernel could not hold
    uint32_t* probe_max;      // [2] probe mode: max reached nodes, max pages
    float delta, bin_scale;
    float* dump_agg;          // optional per-node dumps by ORIGINAL index (single-source debug search)
    double* dump_sigma;
    uint32_t* dump_npred;
};

__device__ __forceinline__ uint32_t cs2_hash(uint32_t blk, uint32_t TB) { return __umulhi(blk * 0x9E3779B1u, TB); }

// page of block `blk`, inserting it if absent; on a full table sets *fail and returns page 0 (the source is abandoned)
__device__ __forceinline__ uint32_t cs2_map_insert(volatile uint32_t* keys, uint32_t TB, uint32_t blk, uint32_t max_pages,
                                                   uint32_t* npages, volatile int* fail) {
    uint32_t h = cs2_hash(blk, TB);
    for (uint32_t probe = 0; probe < TB; ++probe) {
        const uint32_t k = keys[h];
        if (k == blk) return h;
        if (k == CS2_EMPTY) {
            const uint32_t old = atomicCAS(const_cast<uint32_t*>(keys) + h, CS2_EMPTY, blk);
            if (old == CS2_EMPTY) {
                if (atomicAdd(npages, 1u) >= max_pages) *fail = CS_ERR_REACH_OVERFLOW;
                return h;
            }
            if (old == blk) return h;
        }
        h = (h + 1 == TB) ? 0 : h + 1;
    }
    *fail = CS_ERR_REACH_OVERFLOW;
    return 0;
}

// page of block `blk` or CS2_EMPTY when the block holds no reached node
__device__ __forceinline__ uint32_t cs2_map_find(const uint32_t* keys, uint32_t TB, uint32_t blk) {
    uint32_t h = cs2_hash(blk, TB);
    for (uint32_t probe = 0; probe < TB; ++probe) {
        const uint32_t k = keys[h];
        if (k == blk) return h;
        if (k == CS2_EMPTY) return CS2_EMPTY;
        h = (h + 1 == TB) ? 0 : h + 1;
    }
    return CS2_EMPTY;
}

template <int DT>
__device__ __forceinline__ int cs2_first_threshold(const CsShortest2Params& p, float cost) {
    int ti = DT;
#pragma unroll
    for (int i = DT - 1; i >= 0; --i)
        if (i < p.D && cost <= p.dist_f[i]) ti = i;
    return ti;
}

__device__ __forceinline__ uint32_t cs2_bin(uint32_t ab, float bin_scale, uint32_t NB) {
    const float a = __uint_as_float(ab);
    return min(NB - 1u, (uint32_t)(__fmul_rn(__fmul_rn(a, a), bin_scale)));
}

template <int DT, int T>
__global__ void __launch_bounds__(T, T >= 512 ? 1 : 512 / T) cs_k_shortest2(const CsShortest2Params p) {
    constexpr uint32_t CS2_T = T;
    constexpr uint32_t CS2_WARPS = T / 32;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t s_cnt[3];       // rotating near-queue counters
    __shared__ uint32_t s_npages, s_min, s_R, s_maxgap;
    __shared__ int s_fail;
    __shared__ unsigned long long s_si;
    __shared__ uint32_t s_histE[CS_MAX_THRESHOLDS + 1];
    __shared__ float s_rankf[CS_MAX_THRESHOLDS];
    __shared__ uint32_t s_scan[CS2_WARPS];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t ltmask = cs_lanemask_lt();
    const CsV2Smem& M = p.sm;
    const uint32_t S = M.S, TB = M.TB, pb = M.pb, pm = (1u << pb) - 1u, QC = M.QC, NB = M.NB, WS = M.WS, WD = M.WD;
    uint32_t* s_dist = reinterpret_cast<uint32_t*>(smem + M.off_dist);
    uint32_t* s_keys = reinterpret_cast<uint32_t*>(smem + M.off_keys);
    uint32_t* s_defer = reinterpret_cast<uint32_t*>(smem + M.off_defer);
    uint16_t* s_rank = reinterpret_cast<uint16_t*>(smem + M.off_rank);
    uint16_t* s_perm = reinterpret_cast<uint16_t*>(smem + M.off_perm);
    uint8_t* s_u = smem + M.off_u;
    uint8_t* s_pmask = smem + M.off_pmask;  // inside the union region, behind the sigma ring (written from P3 on)
    // P1 view of the union region: two queues of three words per item
    uint32_t* qbase = reinterpret_cast<uint32_t*>(s_u);
    // P2 view: bins[NB + 1], tmp u16[rcap]
    uint32_t* s_bins = reinterpret_cast<uint32_t*>(s_u);
    uint16_t* s_tmp = reinterpret_cast<uint16_t*>(s_u + (size_t)(NB + 1) * 4);
    // P3 view: sigma ring f64[WS]
    volatile double* s_sig = reinterpret_cast<volatile double*>(s_u);
    // P5 view: ring of {sigma, dep[2D]} entries + done flags
    const int D = p.D, D2 = 2 * D, ES = D2 + 1;
    volatile double* s_dep = reinterpret_cast<volatile double*>(smem + M.off_dep);

    uint8_t* scr = p.scratch + (size_t)blockIdx.x * p.scratch_stride;
    uint32_t* g_agg = reinterpret_cast<uint32_t*>(scr);
    double* g_sigma = reinterpret_cast<double*>(scr + (((size_t)M.rcap * 4 + 255) & ~(size_t)255));
    double* g_dep = g_sigma + (((size_t)M.rcap + 31) & ~(size_t)31);

    const float one_minus = 1.0f - CS_TIE_EPS, one_plus = 1.0f + CS_TIE_EPS;
    const float one_plus_tol = 1.0f + p.tol;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_si = atomicAdd(&p.counters[CS_C_NEXT], 1ull);
        __syncthreads();
        const unsigned long long si = s_si;
        if (si >= p.n_sources) break;
        if (*reinterpret_cast<volatile int*>(p.error) != 0) break;
        const uint32_t src = __ldg(&p.sources[si]);
        const float wt = __ldg(&p.src_wt[si]);
        long long tc[7];
        tc[0] = clock64();

        // ------------------------------------------------------------------ init
        for (uint32_t i = tid; i < S; i += CS2_T) s_dist[i] = CS_INF_BITS;
        for (uint32_t i = tid; i < TB; i += CS2_T) s_keys[i] = CS2_EMPTY;
        for (uint32_t i = tid; i < S / 32; i += CS2_T) s_defer[i] = 0;
        if (tid <= CS_MAX_THRESHOLDS) {
            s_histE[tid] = 0;
        }
        if (tid == 0) {
            s_npages = 0;
            s_cnt[0] = 1;
            s_cnt[1] = 0;
            s_cnt[2] = 0;
            s_fail = 0;
            s_maxgap = 0;
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t page = cs2_map_insert(s_keys, TB, src >> pb, M.max_pages, &s_npages, &s_fail);
            const uint32_t slot = (page << pb) | (src & pm);
            s_dist[slot] = 0u;
            const uint4 nd = __ldg(&p.g.node2[src]);
            qbase[0] = slot | ((nd.z & 0xffu) << 16);
            qbase[QC] = 0u;
            qbase[2 * QC] = nd.x;
        }
        __syncthreads();

        // ------------------------------------------------------------------ P1: capped label-correcting search
        // (centrality.rs:1363-1442; f32 `+` is monotone, so the fixed point equals the reference's Dijkstra distances)
        unsigned long long relax = 0;
        {
            float thr = p.delta;
            uint32_t it = 0;  // iteration counter selects the rotating queue / counter
            uint32_t n_split = 0;
            const long long t_p1 = clock64();
            long long t_split = 0, t_s0 = 0;
            for (;;) {
                for (;;) {
                    const uint32_t cur = it % 3u, nxt = (it + 1u) % 3u;
                    const uint32_t nc = min(s_cnt[cur], QC);
                    if (tid == 0) s_cnt[(it + 2u) % 3u] = 0;
                    if (nc == 0) break;
                    const uint32_t* qc = qbase + (size_t)(it & 1u) * 3 * QC;
                    uint32_t* qn = qbase + (size_t)((it + 1u) & 1u) * 3 * QC;
                    const uint32_t work = nc * 4u;
                    for (uint32_t base = 0; base < work; base += CS2_T) {
                        const uint32_t idx = base + tid;
                        const uint32_t item = idx >> 2, jj = idx & 3u;
                        uint32_t deg = 0, skip = 0, eb = 0, abits = 0;
                        if (item < nc) {
                            const uint32_t w0 = qc[item];
                            abits = qc[QC + item];
                            eb = qc[2 * QC + item];
                            const uint32_t slot = w0 & 0xffffu;
                            if (s_dist[slot] == abits) {  // else stale: superseded by a smaller distance
                                deg = (w0 >> 16) & 0xfu;
                                skip = (w0 >> 20) & 0xfu;
                                if (jj == 0) {
                                    // this visit also serves a pending deferred visit of the slot: clear the bit first,
                                    // then take the distance as it is now (a concurrent improvement re-sets the bit or
                                    // queues its own item, so nothing is lost)
                                    const uint32_t bit = 1u << (slot & 31u);
                                    if (s_defer[slot >> 5] & bit) {
                                        atomicAnd(&s_defer[slot >> 5], ~bit);
                                        __threadfence_block();
                                        const uint32_t now = *reinterpret_cast<volatile uint32_t*>(&s_dist[slot]);
                                        if (now != abits) {
                                            abits = now;
                                            skip = 0x10u;  // the remembered back edge belongs to the older distance
                                        }
                                    }
                                }
                            }
                        }
                        // the four lanes of an item relax with the same distance
                        abits = __shfl_sync(CS_FULL, abits, lane & ~3u);
                        if (__shfl_sync(CS_FULL, skip, lane & ~3u) == 0x10u) skip = 0;
                        const float a = __uint_as_float(abits);
#pragma unroll
                        for (uint32_t round = 0; round < 2; ++round) {
                            const uint32_t j = jj + 4u * round;
                            if (round && !__any_sync(CS_FULL, j < deg)) break;
                            bool pn = false;
                            uint32_t i0 = 0, i1 = 0, i2 = 0, dslot = 0;
                            if (j < deg && j + 1u != skip) {
                                uint4 raw = __ldg(reinterpret_cast<const uint4*>(&p.g.in2[eb + j]));
                                float from = a;
                                for (int hop = 0;; ++hop) {
                                    const float cand = __fadd_rn(from, __uint_as_float(raw.y));
                                    if ((raw.w & 0x200u) || cand > p.max_seconds) break;
                                    const uint32_t cbits = __float_as_uint(cand);
                                    const uint32_t nb = raw.x;
                                    const uint32_t page = cs2_map_insert(s_keys, TB, nb >> pb, M.max_pages, &s_npages, &s_fail);
                                    const uint32_t nslot = (page << pb) | (nb & pm);
                                    const uint32_t old = atomicMin(&s_dist[nslot], cbits);
                                    if (!(cbits < old)) break;
                                    ++relax;
                                    if (!(cand < thr)) {
                                        atomicOr(&s_defer[nslot >> 5], 1u << (nslot & 31u));
                                        break;
                                    }
                                    const uint32_t ndeg = (raw.w >> 24) & 0xfu, back = (raw.w >> 16) & 0xfu;
                                    if (ndeg == 2u && back != 0u && hop < CS2_CHAIN_HOPS) {
                                        // degree-2 node: its only other incoming edge continues the chain; relax it here
                                        // instead of paying a queue round trip per 20 m segment of a decomposed street
                                        raw = __ldg(reinterpret_cast<const uint4*>(&p.g.in2[raw.z + ((back - 1u) ^ 1u)]));
                                        from = cand;
                                        continue;
                                    }
                                    pn = true;
                                    i0 = nslot | (ndeg << 16) | (back << 20);
                                    i1 = cbits;
                                    i2 = raw.z;
                                    dslot = nslot;
                                    break;
                                }
                            }
                            const uint32_t m = __ballot_sync(CS_FULL, pn);
                            if (m) {
                                uint32_t b = 0;
                                const int leader = __ffs(m) - 1;
                                if ((int)lane == leader) b = atomicAdd(&s_cnt[nxt], (uint32_t)__popc(m));
                                b = __shfl_sync(CS_FULL, b, leader);
                                if (pn) {
                                    const uint32_t pos = b + __popc(m & ltmask);
                                    if (pos < QC) {
                                        qn[pos] = i0;
                                        qn[QC + pos] = i1;
                                        qn[2 * QC + pos] = i2;
                                    } else {
                                        atomicOr(&s_defer[dslot >> 5], 1u << (dslot & 31u));  // queue full: defer
                                    }
                                }
                            }
                        }
                    }
                    __syncthreads();
                    ++it;
                }
                // near bucket exhausted (all threads saw nc == 0 after the same barrier)
                if (s_fail) break;
                t_s0 = clock64();
                // pass 1: smallest deferred distance
                __syncthreads();
                if (tid == 0) s_min = CS_INF_BITS;
                __syncthreads();
                {
                    uint32_t mn = CS_INF_BITS;
                    for (uint32_t w = tid; w < S / 32; w += CS2_T) {
                        for (uint32_t bits = s_defer[w]; bits; bits &= bits - 1) mn = min(mn, s_dist[w * 32 + (__ffs(bits) - 1)]);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(CS_FULL, mn, o));
                    if (lane == 0 && mn != CS_INF_BITS) atomicMin(&s_min, mn);
                }
                __syncthreads();
                const uint32_t mnb = s_min;
                if (mnb == CS_INF_BITS) break;
                thr = __uint_as_float(mnb) + p.delta;
                ++n_split;
                // pass 2: move the deferred slots below the new threshold into the current queue
                {
                    const uint32_t cur = it % 3u;
                    uint32_t* qc = qbase + (size_t)(it & 1u) * 3 * QC;
                    for (uint32_t w = tid; w < S / 32; w += CS2_T) {
                        const uint32_t bits0 = s_defer[w];
                        uint32_t take = 0;
                        for (uint32_t bits = bits0; bits; bits &= bits - 1) {
                            const uint32_t b = __ffs(bits) - 1;
                            if (__uint_as_float(s_dist[w * 32 + b]) < thr) take |= 1u << b;
                        }
                        if (take) {
                            uint32_t pos = atomicAdd(&s_cnt[cur], (uint32_t)__popc(take));
                            uint32_t left = 0;
                            for (uint32_t bits = take; bits; bits &= bits - 1, ++pos) {
                                const uint32_t b = __ffs(bits) - 1;
                                if (pos < QC) {
                                    const uint32_t slot = w * 32 + b;
                                    const uint32_t v = (s_keys[slot >> pb] << pb) | (slot & pm);
                                    const uint4 nd = __ldg(&p.g.node2[v]);
                                    qc[pos] = slot | ((nd.z & 0xffu) << 16);
                                    qc[QC + pos] = s_dist[slot];
                                    qc[2 * QC + pos] = nd.x;
                                } else {
                                    left |= 1u << b;  // no room: stays deferred
                                }
                            }
                            s_defer[w] = (bits0 & ~take) | left;
                        }
                    }
                }
                __syncthreads();
                t_split += clock64() - t_s0;
            }
            if (tid == 0) {
                atomicAdd(&p.counters[CS_C_PHASE0 + 6], (unsigned long long)it);
                atomicAdd(&p.counters[CS_C_PHASE0 + 7], (unsigned long long)n_split);
                atomicAdd(&p.counters[CS_C_PHASE0 + 5], (unsigned long long)t_split);      // debug: split cycles
                atomicAdd(&p.counters[CS_C_REACH0 + 15], (unsigned long long)(t_p1 - tc[0]));  // debug: init cycles
            }
        }
        __syncthreads();
        tc[1] = clock64();
        int fail = s_fail;

        // ------------------------------------------------------------------ P2: exact settle order
        // counting sort into NB bins (quadratic in the distance), exact rank inside each bin on (seconds bits, original
        // node index) with the source first: the reference's pop order whenever keys are distinct.
        uint32_t R = 0;
        if (!fail) {
            for (uint32_t i = tid; i <= NB; i += CS2_T) s_bins[i] = 0;
            __syncthreads();
            for (uint32_t s = tid; s < S; s += CS2_T) {
                const uint32_t d = s_dist[s];
                if (d != CS_INF_BITS) atomicAdd(&s_bins[cs2_bin(d, p.bin_scale, NB)], 1u);
            }
            __syncthreads();
            {
                // exclusive scan: each thread owns NB / CS2_T consecutive bins
                const uint32_t per = NB / CS2_T;
                uint32_t loc = 0;
                for (uint32_t k = 0; k < per; ++k) loc += s_bins[tid * per + k];
                uint32_t inc = loc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(CS_FULL, inc, o);
                    if ((int)lane >= o) inc += t;
                }
                if (lane == 31) s_scan[wid] = inc;
                __syncthreads();
                uint32_t woff = 0;
                for (uint32_t k = 0; k < wid; ++k) woff += s_scan[k];
                uint32_t run = woff + inc - loc;
                for (uint32_t k = 0; k < per; ++k) {
                    const uint32_t c = s_bins[tid * per + k];
                    s_bins[tid * per + k] = run;
                    run += c;
                }
                if (tid == CS2_T - 1) s_R = run;
            }
            __syncthreads();
            R = s_R;
            if (R > M.rcap) fail = CS_ERR_REACH_OVERFLOW;
        }
        if (p.probe) {
            if (tid == 0) {
                atomicMax(&p.probe_max[0], fail ? 0xffffffffu : R);
                atomicMax(&p.probe_max[1], fail ? 0xffffffffu : s_npages);
                atomicAdd(&p.counters[CS_C_SOURCES], 1ull);
            }
            continue;
        }
        if (fail) {
            if (tid == 0) {
                const unsigned long long k = atomicAdd(&p.counters[CS_C_FALLBACK], 1ull);
                p.fallback[k] = (uint32_t)si;
            }
            continue;
        }
        {
            for (uint32_t s = tid; s < S; s += CS2_T) {
                const uint32_t d = s_dist[s];
                if (d != CS_INF_BITS) s_tmp[atomicAdd(&s_bins[cs2_bin(d, p.bin_scale, NB)], 1u)] = (uint16_t)s;
            }
            __syncthreads();
            // s_bins[b] is now the end offset of bin b
            for (uint32_t pos = tid; pos < R; pos += CS2_T) {
                const uint32_t slot = s_tmp[pos];
                const uint32_t d = s_dist[slot];
                const uint32_t b = cs2_bin(d, p.bin_scale, NB);
                const uint32_t start = b ? s_bins[b - 1] : 0u, end = s_bins[b];
                uint32_t rank = start;
                uint32_t tie_self = 0xffffffffu;  // lazily loaded tie key
                for (uint32_t j = start; j < end; ++j) {
                    if (j == pos) continue;
                    const uint32_t os = s_tmp[j];
                    const uint32_t od = s_dist[os];
                    if (od < d) {
                        ++rank;
                    } else if (od == d) {
                        if (tie_self == 0xffffffffu) {
                            const uint32_t v = (s_keys[slot >> pb] << pb) | (slot & pm);
                            tie_self = v == src ? 0u : __ldg(&p.g.orig_of_new[v]) + 1u;
                        }
                        const uint32_t ov = (s_keys[os >> pb] << pb) | (os & pm);
                        const uint32_t tie_o = ov == src ? 0u : __ldg(&p.g.orig_of_new[ov]) + 1u;
                        if (tie_o < tie_self) ++rank;
                    }
                }
                s_rank[slot] = (uint16_t)rank;
            }
            __syncthreads();
            for (uint32_t pos = tid; pos < R; pos += CS2_T) {
                const uint32_t slot = s_tmp[pos];
                const uint32_t r = s_rank[slot];
                s_perm[r] = (uint16_t)slot;
                __stcg(&g_agg[r], s_dist[slot]);
            }
            __syncthreads();
        }
        tc[2] = clock64();

        // ------------------------------------------------------------------ P3: predecessors + sigma (outgoing edges)
        unsigned long long edge_iters = 0;
        for (uint32_t b0 = 0; b0 < R; b0 += CS2_T) {
            const uint32_t r = b0 + tid;
            const bool valid = r < R;
            if (valid) s_sig[r % WS] = 0.0;
            __syncthreads();
            uint32_t v = 0;
            float cc[CS2_MAX_DEG];
            uint32_t cu[CS2_MAX_DEG], crk[CS2_MAX_DEG], cj[CS2_MAX_DEG];
            int ncand = 0;
            uint32_t pmask_c = 0;
            constexpr int NEW = (DT + 1 + 3) / 4;
            unsigned long long ep[NEW];  // circuit-rank edge histogram of this node, 16 bits per threshold bin
#pragma unroll
            for (int k = 0; k < NEW; ++k) ep[k] = 0ull;
            if (valid) {
                const uint32_t slot = s_perm[r];
                v = (s_keys[slot >> pb] << pb) | (slot & pm);
                const float av = __uint_as_float(s_dist[slot]);
                const float cost_v = __fmul_rn(av, p.speed);
                const uint4 nd = __ldg(&p.g.node2[v]);
                const uint32_t eb = nd.y, deg = (nd.z >> 8) & 0xffu;
                edge_iters += nd.z & 0xffu;
                for (uint32_t j = 0; j < deg; ++j) {
                    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&p.g.out2[eb + j]));
                    const uint32_t u = raw.x;
                    if (u == v) continue;
                    const uint32_t upage = cs2_map_find(s_keys, TB, u >> pb);
                    if (upage == CS2_EMPTY) continue;
                    const uint32_t uslot = (upage << pb) | (u & pm);
                    const uint32_t ub = s_dist[uslot];
                    if (ub == CS_INF_BITS) continue;
                    const float au = __uint_as_float(ub);
                    if (p.closeness && (raw.w & 0x100u)) {
                        const float ec = fmaxf(cost_v, __fmul_rn(au, p.speed));
                        const int t = cs2_first_threshold<DT>(p, ec);
#pragma unroll
                        for (int k = 0; k < NEW; ++k)
                            if ((t >> 2) == k) ep[k] += 1ull << ((t & 3) * 16);
                    }
                    const uint32_t urank = s_rank[uslot];
                    if (urank >= r || v == src) continue;  // u must be settled before v; the source has no predecessors
                    const float c = __fadd_rn(au, __uint_as_float(raw.y));
                    if (!p.phase2 && c > p.max_seconds) continue;
                    const uint32_t ipos = raw.w & 0xffu;
                    int k = ncand++;
                    while (k > 0 && (crk[k - 1] > urank || (crk[k - 1] == urank && (cj[k - 1] >> 8) > ipos))) {
                        cc[k] = cc[k - 1];
                        cu[k] = cu[k - 1];
                        crk[k] = crk[k - 1];
                        cj[k] = cj[k - 1];
                        --k;
                    }
                    cc[k] = c;
                    cu[k] = u;
                    crk[k] = urank;
                    cj[k] = j | (ipos << 8);
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
                uint32_t amask = 0, gap = 0;
                for (uint32_t mm = pmask_c; mm; mm &= mm - 1) {
                    const int k = __ffs(mm) - 1;
                    amask |= 1u << (cj[k] & 0xffu);
                    gap = max(gap, r - crk[k]);
                }
                s_pmask[r] = (uint8_t)amask;
                if (gap > s_maxgap) atomicMax(&s_maxgap, gap);
                if (p.dump_npred) p.dump_npred[__ldg(&p.g.orig_of_new[v])] = __popc(pmask_c);
            }
            // circuit-rank edge histogram, aggregated per warp (one shared-memory atomic per warp and non-empty bin)
            if (p.closeness) {
#pragma unroll
                for (int k = 0; k < NEW; ++k) {
                    unsigned long long e = ep[k];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(CS_FULL, e, o);
                    if (lane == 0 && e) {
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const uint32_t c = (uint32_t)(e >> (16 * b)) & 0xffffu;
                            if (c && 4 * k + b <= D) atomicAdd(&s_histE[4 * k + b], c);
                        }
                    }
                }
            }
            // sigma = sum over predecessors in settle order; predecessors of this chunk may still be pending (ring == 0)
            bool pending = valid;
            const uint32_t ring_lo = b0 + 2u * CS2_T >= WS ? b0 + 2u * CS2_T - WS : 0u;  // older ranks: read global
            // predecessor ranks in registers (two cover almost every node; more take the generic path)
            uint32_t pr0 = CS2_EMPTY, pr1 = CS2_EMPTY;
            const bool pmany = __popc(pmask_c) > 2;
            if (pmask_c) {
                pr0 = crk[__ffs(pmask_c) - 1];
                const uint32_t rest = pmask_c & (pmask_c - 1);
                if (rest) pr1 = crk[__ffs(rest) - 1];
            }
            double sfar = 0.0;  // predecessors older than the ring: final, read once from global
            if (pr0 != CS2_EMPTY && pr0 < ring_lo) {
                sfar += __ldcg(&g_sigma[pr0]);
                pr0 = CS2_EMPTY;
            }
            if (pr1 != CS2_EMPTY && pr1 < ring_lo) {
                sfar += __ldcg(&g_sigma[pr1]);
                pr1 = CS2_EMPTY;
            }
            for (;;) {
                if (pending) {
                    double s = 0.0;
                    bool ok = true;
                    if (v == src) {
                        s = 1.0;
                    } else if (!pmany) {
                        // summation order = settle order of the predecessors (pr0 before pr1), far ones first
                        const double s0 = pr0 != CS2_EMPTY ? s_sig[pr0 % WS] : 1.0;
                        const double s1 = pr1 != CS2_EMPTY ? s_sig[pr1 % WS] : 1.0;
                        ok = s0 != 0.0 && s1 != 0.0;
                        s = sfar;
                        if (pr0 != CS2_EMPTY) s += s0;
                        if (pr1 != CS2_EMPTY) s += s1;
                    } else {
                        for (uint32_t mm = pmask_c; mm; mm &= mm - 1) {
                            const uint32_t ur = crk[__ffs(mm) - 1];
                            const double sg = ur >= ring_lo ? s_sig[ur % WS] : __ldcg(&g_sigma[ur]);
                            if (sg == 0.0) {
                                ok = false;
                                break;
                            }
                            s += sg;
                        }
                    }
                    if (ok) {
                        if (s == 0.0) {
                            // a reached node without an earlier-settled predecessor: only zero-length edges whose
                            // endpoints tie on (seconds, index) produce this; fail loudly instead of guessing
                            atomicCAS(p.error, 0, CS_ERR_ZERO_TIE);
                            s = 1.0;
                        }
                        s_sig[r % WS] = s;
                        __stcg(&g_sigma[r], s);
                        pending = false;
                    }
                }
                __syncwarp();
                if (!__any_sync(CS_FULL, pending)) break;
                __nanosleep(CS2_POLL_NS);
            }
        }
        __syncthreads();
        if (p.dump_agg) {
            for (uint32_t r = tid; r < R; r += CS2_T) {
                const uint32_t slot = s_perm[r];
                const uint32_t o = __ldg(&p.g.orig_of_new[(s_keys[slot >> pb] << pb) | (slot & pm)]);
                p.dump_agg[o] = __uint_as_float(s_dist[slot]);
                p.dump_sigma[o] = __ldcg(&g_sigma[r]);
            }
        }
        tc[3] = clock64();

        // ------------------------------------------------------------------ P4: closeness scatter to targets
        unsigned long long n_ri = 0, n_ci = 0;
        if (p.closeness) {
            if (tid < (uint32_t)D) {
                // circuit rank per threshold = max(0, E_i - N_i + 1) over the reached subgraph (centrality.rs:517-525);
                // N_i = reached nodes with cost <= d_i = an upper bound in the settle order (cost is monotone in rank)
                uint32_t lo = 0, hi = R;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    const float c = __fmul_rn(__uint_as_float(s_dist[s_perm[mid]]), p.speed);
                    if (c <= p.dist_f[tid]) lo = mid + 1; else hi = mid;
                }
                const long long ncount = lo;
                long long ecount = 0;
                for (int t = 0; t <= (int)tid; ++t) ecount += s_histE[t];
                s_rankf[tid] = ncount == 0 ? 0.0f : (float)max(ecount - ncount + 1ll, 0ll);
                atomicAdd(&p.counters[CS_C_REACH0 + tid], (unsigned long long)(ncount > 0 ? ncount - 1 : 0));
            }
            __syncthreads();
            const float cycles_wt = __fdiv_rn(wt, __uint_as_float(__ldg(&p.g.node2[src]).w));  // centrality.rs:1730
            constexpr int NQ = 5 * DT;
            constexpr int LP = NQ <= 16 ? 16 : 32;
            constexpr int G = 32 / LP;
            const uint32_t ql = lane & (LP - 1);
            const int nq = 5 * D;
            for (uint32_t b0 = wid * 32u; b0 < R; b0 += CS2_T) {
                const uint32_t r = b0 + lane;
                uint32_t node = 0;
                float cost = __uint_as_float(CS_INF_BITS);
                if (r < R) {
                    const uint32_t slot = s_perm[r];
                    node = (s_keys[slot >> pb] << pb) | (slot & pm);
                    if (node != src) cost = __fmul_rn(__uint_as_float(s_dist[slot]), p.speed);
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
                                float val = wt;
                                if (m == 0) ++n_ri;
                                if (m == 1) val = f1;
                                if (m == 2) val = __fmul_rn(s_rankf[i], cycles_wt);
                                if (m == 3) val = f3;
                                if (m == 4) {
#pragma unroll
                                    for (int ii = 0; ii < DT; ++ii)
                                        if (ii == i) val = f4[ii];
                                }
                                cs_red_add(p.acc_c + (size_t)nd * p.cw + q, (double)val);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();  // s_dist is dead from here on: its storage may hold the dependency ring
        tc[4] = clock64();

        // ------------------------------------------------------------------ P5: dependencies, reverse settle order
        if (p.betweenness) {
            const double wt_d = (double)wt;
            const bool far_deps = s_maxgap + CS2_T >= WD;  // some successor may lie beyond the ring: write dependencies through
            for (int b0 = (int)((R - 1) / CS2_T * CS2_T); b0 >= 0; b0 -= (int)CS2_T) {
                const uint32_t r = (uint32_t)b0 + tid;
                const bool valid = r < R;
                __syncthreads();  // every reader of the previous chunk is done with the entries recycled below
#ifdef CS2_DEBUG_CLOCKS
                long long tq[6];
                tq[0] = clock64();
#endif
                double sigma_w = 1.0;
                float cost_w = 0.f;
                const long long SENT = -1ll;  // "not yet written" pattern of a ring word (a NaN no dependency can equal)
                if (valid) {
                    sigma_w = __ldcg(&g_sigma[r]);
                    cost_w = __fmul_rn(__uint_as_float(__ldcg(&g_agg[r])), p.speed);
                    volatile double* er0 = s_dep + (size_t)(r % WD) * ES;
                    er0[0] = sigma_w;
                    for (int i = 1; i <= D2; ++i) er0[i] = __longlong_as_double(SENT);
                }
                __syncthreads();
#ifdef CS2_DEBUG_CLOCKS
                tq[1] = clock64();
#endif
                uint32_t w = 0;
                // successors by in-list position (static indexing keeps them in registers); CS2_EMPTY = none
                uint32_t sr[CS2_MAX_DEG];
                double fr[CS2_MAX_DEG];
#pragma unroll
                for (int j = 0; j < CS2_MAX_DEG; ++j) sr[j] = CS2_EMPTY;
                double accf[DT], accbf[DT];  // contributions of successors beyond the ring (read from global once)
#pragma unroll
                for (int i = 0; i < DT; ++i) accf[i] = accbf[i] = 0.0;
                if (valid) {
                    const uint32_t slot = s_perm[r];
                    w = (s_keys[slot >> pb] << pb) | (slot & pm);
                    const uint4 nd = __ldg(&p.g.node2[w]);
                    const uint32_t eb = nd.x, deg = nd.z & 0xffu;
#pragma unroll
                    for (int j = 0; j < CS2_MAX_DEG; ++j) {
                        if ((uint32_t)j < deg) {
                            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&p.g.in2[eb + j]));
                            const uint32_t x = raw.x;
                            const uint32_t xpage = x == w ? CS2_EMPTY : cs2_map_find(s_keys, TB, x >> pb);
                            if (xpage != CS2_EMPTY) {
                                const uint32_t xslot = (xpage << pb) | (x & pm);
                                const uint32_t xr = s_rank[xslot];
                                // s_dist may be recycled: "reached" is decided by the permutation (rank valid and maps back)
                                if (xr < R && xr > r && s_perm[xr] == xslot && ((s_pmask[xr] >> (raw.w & 0xffu)) & 1u)) {
                                    const bool in_ring = xr < (uint32_t)b0 + WD;
                                    const double sx = in_ring ? s_dep[(size_t)(xr % WD) * ES] : __ldcg(&g_sigma[xr]);
                                    const double f = (sx == sigma_w) ? 1.0 : sigma_w / sx;  // centrality.rs:861-866
                                    if (in_ring) {
                                        sr[j] = xr;
                                        fr[j] = f;
                                    } else {
                                        const double* gx = g_dep + (size_t)xr * D2;
#pragma unroll
                                        for (int i = 0; i < DT; ++i) {
                                            if (i < D) {
                                                accf[i] += f * __ldcg(&gx[i]);
                                                accbf[i] += f * __ldcg(&gx[D + i]);
                                            }
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
#ifdef CS2_DEBUG_CLOCKS
                tq[2] = clock64();
#endif
                bool pending = valid;
                double cr[2 * DT];  // positive credits of this thread's node, slot 2 * i (plain) / 2 * i + 1 (beta-weighted)
#pragma unroll
                for (int q = 0; q < 2 * DT; ++q) cr[q] = 0.0;
                // seeds do not depend on the successors either (centrality.rs:1802-1806, f64 exp)
                const bool is_src = (w == src);
                double seed[DT], seedb[DT];
#pragma unroll
                for (int i = 0; i < DT; ++i) seed[i] = seedb[i] = 0.0;
                if (valid && !is_src) {
                    const double pc = __ldg(&p.eligible[w]) ? 0.5 : 1.0;
#pragma unroll
                    for (int i = 0; i < DT; ++i) {
                        if (i < D && cost_w <= p.dist_f[i]) {
                            seed[i] = pc;
                            seedb[i] = pc * exp(-p.beta_d[i] * (double)cost_w);
                        }
                    }
                }
                // Wait for the successors of this chunk.  A ring word is either the sentinel or final (8-byte stores are
                // single transactions), so readiness is read off the values themselves: no flags, no fences.
#ifdef CS2_DEBUG_CLOCKS
                tq[3] = clock64();
#endif
                // probe words (ring word 1) of the successors that sit in this chunk: the only ones that can be pending
                uint32_t pa0 = CS2_EMPTY, pa1 = CS2_EMPTY;
                bool pmany = false;
#pragma unroll
                for (int j = 0; j < CS2_MAX_DEG; ++j) {
                    if (sr[j] != CS2_EMPTY && sr[j] < (uint32_t)b0 + CS2_T) {
                        const uint32_t ix = (sr[j] % WD) * ES + 1u;
                        if (pa0 == CS2_EMPTY) pa0 = ix;
                        else if (pa1 == CS2_EMPTY) pa1 = ix;
                        else pmany = true;
                    }
                }
                for (;;) {
                    if (pending) {
                        bool ok = (pa0 == CS2_EMPTY || __double_as_longlong(s_dep[pa0]) != SENT) &&
                                  (pa1 == CS2_EMPTY || __double_as_longlong(s_dep[pa1]) != SENT);
                        if (ok && pmany) {
#pragma unroll
                            for (int j = 0; j < CS2_MAX_DEG; ++j) {
                                if (sr[j] != CS2_EMPTY && sr[j] < (uint32_t)b0 + CS2_T)
                                    ok = ok && (__double_as_longlong(s_dep[(size_t)(sr[j] % WD) * ES + 1]) != SENT);
                            }
                        }
                        if (ok) {
                            double acc[DT], accb[DT];
#pragma unroll
                            for (int i = 0; i < DT; ++i) {
                                acc[i] = accf[i];
                                accb[i] = accbf[i];
                            }
#pragma unroll
                            for (int j = 0; j < CS2_MAX_DEG; ++j) {
                                if (sr[j] != CS2_EMPTY) {
                                    const volatile double* e = s_dep + (size_t)(sr[j] % WD) * ES;
#pragma unroll
                                    for (int i = 0; i < DT; ++i) {
                                        if (i < D) {
                                            const double a0 = e[1 + i], a1 = e[1 + D + i];
                                            ok = ok && (__double_as_longlong(a0) != SENT) && (__double_as_longlong(a1) != SENT);
                                            acc[i] += fr[j] * a0;
                                            accb[i] += fr[j] * a1;
                                        }
                                    }
                                }
                            }
                            if (ok) {
                                volatile double* er = s_dep + (size_t)(r % WD) * ES;
#pragma unroll
                                for (int i = DT - 1; i >= 0; --i) {
                                    if (i < D) {
                                        er[1 + D + i] = seedb[i] + accb[i];
                                        er[1 + i] = seed[i] + acc[i];  // word 1 is written last: the cheap readiness probe
                                    }
                                }
                                pending = false;
                                // off the critical path: write-through for far readers and the credits
                                double* gr = g_dep + (size_t)r * D2;
#pragma unroll
                                for (int i = 0; i < DT; ++i) {
                                    if (i < D) {
                                        const double dpn = seed[i] + acc[i], dpb = seedb[i] + accb[i];
                                        if (far_deps) {
                                            __stcg(&gr[i], dpn);
                                            __stcg(&gr[D + i], dpb);
                                        }
                                        if (!is_src) {
                                            const double credit = dpn - seed[i], creditb = dpb - seedb[i];
                                            if (credit > 0.0 || creditb > 0.0) {
                                                ++n_ci;
                                                if (credit > 0.0) cr[2 * i] = credit * wt_d;
                                                if (creditb > 0.0) cr[2 * i + 1] = creditb * wt_d;
                                            }
                                        }
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (!__any_sync(CS_FULL, pending)) break;
                    __nanosleep(CS2_POLL_NS);  // waiting warps leave the issue slots to the warps that make progress
                }
#ifdef CS2_DEBUG_CLOCKS
                tq[4] = clock64();
#endif
                // packed credit scatter: 32/LPB nodes per warp instruction, LPB consecutive doubles each
                {
                    constexpr int NQB = 2 * DT;
                    constexpr int LPB = NQB <= 2 ? 2 : NQB <= 4 ? 4 : NQB <= 8 ? 8 : NQB <= 16 ? 16 : 32;
                    constexpr int GB = 32 / LPB;
                    const int q = (int)(lane & (LPB - 1));
                    const uint32_t wb0 = (uint32_t)b0 + wid * 32u;
                    const uint32_t cnt = wb0 < R ? min(32u, R - wb0) : 0u;
                    for (uint32_t g0 = 0; g0 < cnt; g0 += GB) {
                        const int sl = (int)(g0 + lane / LPB);
                        const uint32_t nd = __shfl_sync(CS_FULL, w, sl);
                        double val = 0.0;
#pragma unroll
                        for (int qq = 0; qq < NQB; ++qq) {
                            const double t = __shfl_sync(CS_FULL, cr[qq], sl);
                            if (q == qq) val = t;
                        }
                        if (val > 0.0) cs_red_add(p.acc_b + (size_t)nd * p.bw + q, val);
                    }
                }
#ifdef CS2_DEBUG_CLOCKS
                tq[5] = clock64();
                if (tid == 0)
                    for (int k = 0; k < 5; ++k) atomicAdd(&p.counters[CS_C_REACH0 + 10 + k], (unsigned long long)(tq[k + 1] - tq[k]));
#endif
            }
        }
        tc[5] = clock64();

        // ------------------------------------------------------------------ per-source counters
        edge_iters = cs_warp_sum(edge_iters);
        relax = cs_warp_sum(relax);
        n_ri = cs_warp_sum(n_ri);
        n_ci = cs_warp_sum(n_ci);
        if (lane == 0) {
            atomicAdd(&p.counters[CS_C_EDGE_ITERS], edge_iters);
            atomicAdd(&p.counters[CS_C_RELAX], relax);
            if (n_ri) atomicAdd(&p.counters[CS_C_SUM_RI], n_ri);
            if (n_ci) atomicAdd(&p.counters[CS_C_SUM_CI], n_ci);
        }
        if (tid == 0) {
            atomicAdd(&p.counters[CS_C_SOURCES], 1ull);
            atomicAdd(&p.counters[CS_C_SETTLED], (unsigned long long)R);
            atomicAdd(&p.counters[CS_C_PROGRESS], 1ull);
#pragma unroll
            for (int k = 0; k < 5; ++k) atomicAdd(&p.counters[CS_C_PHASE0 + k], (unsigned long long)(tc[k + 1] - tc[k]));
        }
    }
}

// Epilogue for the renumbered accumulators: new-id rows -> [7][D][node_bound] in original index order.
__global__ void cs_k_epilogue_shortest2(const double* __restrict__ acc_c, const double* __restrict__ acc_b, double* out,
                                        const uint32_t* __restrict__ orig_of_new, uint32_t n, int D, int cw, int bw,
                                        int closeness, int betweenness, int add) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t node = orig_of_new[v];
    if (closeness) {
        const double* row = acc_c + (size_t)v * cw;
        for (int i = 0; i < D; ++i)
            for (int m = 0; m < 5; ++m) {
                double* o = out + ((size_t)(m * D + i)) * n + node;
                const double val = row[5 * i + m];
                *o = add ? *o + val : val;
            }
    } else if (!add) {
        for (int q = 0; q < 5 * D; ++q) out[(size_t)q * n + node] = 0.0;
    }
    if (betweenness) {
        const double* row = acc_b + (size_t)v * bw;
        for (int i = 0; i < D; ++i)
            for (int b = 0; b < 2; ++b) {
                double* o = out + ((size_t)((5 + b) * D + i)) * n + node;
                const double val = row[2 * i + b];
                *o = add ? *o + val : val;
            }
    } else if (!add) {
        for (int q = 5 * D; q < 7 * D; ++q) out[(size_t)q * n + node] = 0.0;
    }
}
