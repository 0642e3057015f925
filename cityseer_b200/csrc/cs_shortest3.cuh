// centrality_shortest, chain-contracted kernel: one warp per source, search / order / predecessors / dependencies over
// JUNCTIONS only, chain interiors produced by walking contiguous per-chain seconds arrays (see cs_api_v3.inl for the
// graph layout).  Same arithmetic as cs_shortest.cuh, reference /root/reference/rust/src/centrality.rs:
//   * distances: the reference adds f32 seconds node by node (:1405); a walk over a chain performs the same additions
//     in the same order, so every interior distance is bit-identical, and min(walk from A, walk from B) is the fixed
//     point of the search because f32 `+` is monotone;
//   * predecessors: an interior has one candidate predecessor (the neighbour nearer its own end) except the single
//     last-settled node of a chain, where the two waves meet: there - and at junctions - the reference's sequential
//     epsilon rule (:1413-1437) or tolerance rule (:1457-1482) is evaluated on the actual candidates in settle order;
//   * sigma is constant along a one-predecessor run, and sigma_pred / sigma_succ == 1 exactly there (:861-866), so the
//     dependency of a run is the running f64 sum of its seeds (:823-873) - evaluated in the reference's order.
// Terminology: a LINK is one end of a chain at a junction; link j of junction v walks outward over nodes m_1 .. m_k to
// the far junction F.  sv[t] are v's outward steps (sv[t] leads to m_{t+1}), sF[t] those of F (sF[t] leads to m_{k-t}).
#pragma once
#include "cs_shortest.cuh"

#define CS3_KMAX 12       // interiors per chain (longer runs are cut at upload)
#define CS3_MAX_LINKS 8   // links per junction
// CTA shape.  The kernel's code (about 125 KB of SASS) is several times the 32 KB instruction cache next to the SM, so
// every warp of an SM should be in the same phase: up to four thresholds ONE 16-warp CTA per SM moves through the phases
// in lock step (measured on the 1M-node graph: 2 CTAs x 8 warps 1.22 M sources/s, 1 x 12: 1.25 M, 1 x 14: 1.30 M,
// 1 x 16: 1.36 M, 1 x 18 at 96 registers: 1.29 M); beyond four thresholds the per-warp shared memory only leaves room
// for 8-warp CTAs.
#ifndef CS3_WARPS_WIDE
#define CS3_WARPS_WIDE 16
#endif
#ifndef CS3_WARPS
#define CS3_WARPS 8
#endif
#ifndef CS3_MIN_BLOCKS
#define CS3_MIN_BLOCKS 2
#endif
#ifndef CS3_WORKERS_PER_SM
#define CS3_WORKERS_PER_SM CS3_WARPS_WIDE  // resident warps per SM of either shape
#endif
template <int DT>
__host__ __device__ constexpr uint32_t cs3_warps() { return DT <= 4 ? CS3_WARPS_WIDE : CS3_WARPS; }
template <int DT>
__host__ __device__ constexpr uint32_t cs3_min_blocks() { return DT <= 4 ? CS3_WORKERS_PER_SM / CS3_WARPS_WIDE : CS3_MIN_BLOCKS; }
#ifndef CS3_PHASE_SYNC
#define CS3_PHASE_SYNC 1
#endif
// Shared memory per warp is kept small on purpose: what the CTA does not take stays L1 cache (16 warps x 12.25 KB left
// 32 KB of L1; measured on the bench, staging 128 / 112 / 96 / 80 / 64 nodes: 1.35 / 1.42 / 1.41 / 1.36 / 1.29 M sources/s,
// 112 and 96 share the footprint of the chain-walk scratch).  The shared-memory carve-out comes in steps (.., 164, 196,
// 228 KB of the SM's 256 KB): 9.5 KB per warp puts the 16-warp CTA under the 164 KB step.
#ifndef CS3_NB3
#define CS3_NB3 112u  // staged nodes per sub-iteration of the dependency pass, up to three thresholds
#endif
#ifndef CS3_NBINS
#define CS3_NBINS 512u  // counting-sort bins of the junction order (the node-level kernels use CS_NBINS = 1024)
#endif
#define CS3_LIST 384      // staged (node, cost) entries per warp for the closeness scatter: 32 lanes x CS3_KMAX interiors

struct CsV3Graph {
    uint32_t J, I, n;
    const uint2* jinfo;           // [J] {first link, links | in-degree << 8}
    const uint4* links;           // {far junction, soff, ibase, k | dir << 4 | pos_at_far << 5 | canonical count << 9}
    const float* csec;            // per chain: fwd[0..k] then bwdr[0..k] (seconds)
    const uint4* ctab;            // per chain with interiors: {soff, ibase, k, A} {B, posA, posB, -}
    const uint32_t* int_chain;    // [I] chain of every interior
    const uint32_t* orig_of_new;  // [n]
    const uint32_t* new_of_orig;  // [n]
    const float* weight;          // [n] by new id
};

struct CsShortest3Params {
    CsV3Graph g;
    int D, closeness, betweenness, phase2;
    int beta_chain;  // every beta is exactly twice the next one (distances that double): one exp, then squarings
    float dist_f[CS_MAX_THRESHOLDS];
    float beta_f[CS_MAX_THRESHOLDS];
    double beta_d[CS_MAX_THRESHOLDS];
    float max_seconds, speed, tol;
    const uint32_t* sources;  // ORIGINAL indices
    const float* src_wt;
    unsigned long long n_sources;
    const uint8_t* eligible;  // by new id
    // betweenness_od_shortest (centrality.rs:2419-2540): the dependency seeds are the OD weights of sources[k]'s
    // destinations, od_dst / od_w [od_off[k], od_off[k + 1]) with od_dst in NEW ids, ascending per origin; NULL otherwise
    const unsigned long long* od_off;
    const uint32_t* od_dst;
    const float* od_w;
    double* acc_c;            // [5 * D][n] metric-major by new id: row 5 * i + m
    double* acc_b;            // [2 * D][n] metric-major by new id: row 2 * i (plain) / 2 * i + 1 (beta-weighted)
    unsigned long long* counters;
    int* error;
    uint8_t* arena;
    CsArenaLayout lay;        // ds sized J + 1 (slot J = a source that is an interior); linfo = per-rank link bytes
    float delta, bin_scale;
};

// Per-chain seconds block (16-byte aligned): fwd[0..k], padding to a multiple of four floats, then bwdr[0..k] - both
// direction arrays start on a 16-byte boundary, so a block is staged with 16-byte asynchronous copies.
__host__ __device__ __forceinline__ uint32_t cs3_pb(uint32_t k) { return (k + 4u) & ~3u; }

// Seed of a reached node (centrality.rs:1802-1806: half a pair when both ends are sources, a whole one otherwise); in an
// OD call the weight of the (origin, node) trip, zero when there is none (:2498-2507) - a binary search of the origin's
// ascending destination slice [lo, hi).
template <bool OD>
__device__ __forceinline__ double cs3_pc(const CsShortest3Params& p, unsigned long long lo, const unsigned long long hi,
                                         const uint32_t nid) {
    if (!OD) return __ldg(&p.eligible[nid]) ? 0.5 : 1.0;
    unsigned long long top = hi;
    while (lo < top) {
        const unsigned long long mid = (lo + top) >> 1;
        if (__ldg(&p.od_dst[mid]) < nid) lo = mid + 1;
        else top = mid;
    }
    return lo < hi && __ldg(&p.od_dst[lo]) == nid ? (double)__ldg(&p.od_w[lo]) : 0.0;
}

struct CsView {
    uint32_t far, sv, sF, k, id1, paf, cnt;  // sv / sF: offsets of the outward steps inside the chain's block
    uint32_t blk, nv;                        // the block: float offset into csec, number of floats
    int step;
};

struct CsSrc3 {
    uint32_t id, slot;  // new id of the source; its junction slot (J when the source is an interior)
    uint32_t interior, soff, ibase, k, p, A, B, posA, posB;
};

__device__ __forceinline__ CsView cs3_view(const CsV3Graph& g, const CsSrc3& S, uint32_t v, uint32_t off, uint32_t j) {
    CsView V;
    if (v == g.J) {
        // the source sits inside a chain at position p (1-based): two links, toward A (0) and toward B (1)
        V.cnt = 1;
        V.blk = S.soff;
        V.nv = cs3_pb(S.k) + S.k + 1;
        if (j == 0) {
            V.sv = cs3_pb(S.k) + (S.k - S.p + 1);
            V.sF = 0;
            V.k = S.p - 1;
            V.id1 = g.J + S.ibase + S.p - 2;
            V.step = -1;
            V.far = S.A;
            V.paf = S.posA;
        } else {
            V.sv = S.p;
            V.sF = cs3_pb(S.k);
            V.k = S.k - S.p;
            V.id1 = g.J + S.ibase + S.p;
            V.step = 1;
            V.far = S.B;
            V.paf = S.posB;
        }
        return V;
    }
    const uint4 L = __ldg(&g.links[off + j]);
    const uint32_t k = L.w & 15u, dir = (L.w >> 4) & 1u;
    V.far = L.x;
    V.k = k;
    V.paf = (L.w >> 5) & 15u;
    V.cnt = (L.w >> 9) & 3u;
    V.blk = L.y;
    V.nv = cs3_pb(k) + k + 1;
    if (dir == 0) {
        V.sv = 0;
        V.sF = cs3_pb(k);
        V.id1 = g.J + L.z;
        V.step = 1;
    } else {
        V.sv = cs3_pb(k);
        V.sF = 0;
        V.id1 = g.J + L.z + k - 1;
        V.step = -1;
    }
    if (S.interior && k > 0 && L.y == S.soff) {
        // the source's own chain, seen from one of its ends: the far end is the source
        V.far = g.J;
        if (dir == 0) {
            V.k = S.p - 1;
            V.sF = cs3_pb(k) + (k - S.p + 1);
            V.paf = 0;
        } else {
            V.k = k - S.p;
            V.sF = S.p;
            V.paf = 1;
        }
    }
    return V;
}

// The seconds of a chain (both directions, <= cs3_pb(CS3_KMAX) + CS3_KMAX + 1 = 29 floats) are staged with independent
// asynchronous 16-byte copies (LDGSTS.128, no registers) into the lane's column of 16-byte cells; the sequential f32 walks
// then run at shared-memory latency instead of one dependent L2 round trip per piece.  Float i of the block sits in cell
// i / 4 of the lane: cells[(i / 4) * 32 + lane], component i % 4.  (4-byte copies per float, round 1: -2.6 % on the bench.)
__device__ __forceinline__ void cs3_load_block(const CsV3Graph& g, const CsView& V, float* cb, uint32_t first, uint32_t count) {
    const uint32_t c0 = first >> 2, c1 = (first + count + 3u) >> 2;
    const float* src = g.csec + V.blk;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(cb);
    for (uint32_t c = c0; c < c1; ++c)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + c * 512u), "l"(src + 4u * c) : "memory");
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// counting-sort bin of a distance (quadratic in the seconds, so bins fill evenly on a 2-D network); bin_scale is set for
// CS3_NBINS bins by the host
__device__ __forceinline__ uint32_t cs3_bin(uint32_t abits, float bin_scale) {
    const float a = __uint_as_float(abits);
    return min((uint32_t)(CS3_NBINS - 1), (uint32_t)(__fmul_rn(__fmul_rn(a, a), bin_scale)));
}

// settle-order tie key of a node (new id): the source first, then ascending original index
__device__ __forceinline__ uint32_t cs3_key(const CsV3Graph& g, const CsSrc3& S, uint32_t id) {
    return id == S.id ? 0u : __ldg(&g.orig_of_new[id]) + 1u;
}
// does node a (distance bits da) settle before node b?
__device__ __forceinline__ bool cs3_before(const CsV3Graph& g, const CsSrc3& S, uint32_t da, uint32_t ida, uint32_t db, uint32_t idb) {
    if (da != db) return da < db;
    return cs3_key(g, S, ida) < cs3_key(g, S, idb);
}

// Two candidates reach a node: its own-side predecessor (c_own, never larger than c_oth) and the neighbour across
// the meeting point (c_oth).  Is the latter kept as a predecessor?  Sequential rule of centrality.rs:1413-1437 in
// arrival order, or the tolerance rule of :1457-1482 against the final distance.
__device__ __forceinline__ bool cs3_other_kept(float c_own, float c_oth, bool own_first, bool phase2, float one_plus_tol,
                                               float max_seconds) {
    if (phase2) return c_oth <= __fmul_rn(c_own, one_plus_tol);
    if (c_oth > max_seconds) return false;  // never relaxed (:1407)
    const float one_minus = 1.0f - CS_TIE_EPS, one_plus = 1.0f + CS_TIE_EPS;
    if (own_first) return c_oth <= __fmul_rn(c_own, one_plus);
    // the other candidate arrived first and is replaced only by a clearly smaller one
    if (c_own < c_oth) return !(c_own < __fmul_rn(c_oth, one_minus));
    return true;
}

// L2 prefetch of per-rank state a later chunk will read.  The per-warp arenas of the resident warps (about 150 KB per
// source, 2 368 warps) exceed the L2, so what the predecessor pass wrote has mostly gone to DRAM by the time the dependency
// pass walks back over it: each lane touches one rank PF ranks ahead, no registers, nothing waits on it.
#ifndef CS3_PREFETCH
#define CS3_PREFETCH 96
#endif
__device__ __forceinline__ void cs3_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// one copy of the f64 exponential (the beta-weighted seeds, centrality.rs:1805) instead of one per call site and
// threshold: the dependency phase has to stay inside the instruction cache
__device__ __noinline__ double cs3_exp(double x) { return exp(x); }

template <int DT>
__device__ __forceinline__ int cs3_first_threshold(const CsShortest3Params& p, float cost) {
    int ti = DT;
#pragma unroll
    for (int i = DT - 1; i >= 0; --i)
        if ((DT <= 4 || i < p.D) && cost <= p.dist_f[i]) ti = i;
    return ti;
}

// Closeness scatter of `total` (node, cost) pairs staged in shared memory, one lane per target (centrality.rs:1755-1777,
// f32 terms).  The accumulators are metric-major ([5 * i + m][n] by new id): the interiors of a chain have consecutive
// ids, so one warp-wide red.f64 covers runs of consecutive doubles.
template <int DT>
__device__ __forceinline__ void cs3_emit_closeness(const CsShortest3Params& p, const uint32_t* l_id, const float* l_cost,
                                                   uint32_t total, float wt, float cycles_wt, const float* rankf,
                                                   unsigned long long& n_ri) {
    const uint32_t lane = cs_lane();
    const size_t n = p.g.n;
    for (uint32_t b0 = 0; b0 < total; b0 += 32) {
        const uint32_t e = b0 + lane;
        if (e < total) {
            const uint32_t node = l_id[e];
            const float cost = l_cost[e];
            double* base = p.acc_c + node;
            const double far_t = (double)__fmul_rn(cost, wt);
            const double harm_t = (double)__fmul_rn(__fdiv_rn(1.0f, cost), wt);
            const double wt_t = (double)wt;
#pragma unroll
            for (int i = 0; i < DT; ++i) {
                if ((DT <= 4 || i < p.D) && cost <= p.dist_f[i]) {
                    ++n_ri;
                    double* q = base + (size_t)(5 * i) * n;
                    cs_red_add(q, wt_t);
                    cs_red_add(q + n, far_t);
                    cs_red_add(q + 2 * n, (double)__fmul_rn(rankf[i], cycles_wt));
                    cs_red_add(q + 3 * n, harm_t);
                    cs_red_add(q + 4 * n, (double)__fmul_rn(expf(__fmul_rn(-p.beta_f[i], cost)), wt));
                }
            }
        }
    }
}

// OD: the origin-destination variant (seeds looked up in the origin's trip list) is a separate instantiation, so that the
// plain kernel - at its register limit - does not carry the lookup state.
template <int DT, bool OD = false>
__global__ void __launch_bounds__(cs3_warps<DT>() * 32, cs3_min_blocks<DT>()) cs_k_shortest3(const CsShortest3Params p) {
    constexpr uint32_t WARPS = cs3_warps<DT>();
    // per-warp shared memory: region A (4 KB): P2 bins | P3 candidates | P4 staged (node, cost) list | P5 node ids / costs;
    // region B (6 KB): P3 walk values | P5 per-node seeds -> credits; region C: P3 / P5 link list, link bytes, P5 outflow
    constexpr uint32_t NB = DT <= 3 ? CS3_NB3 : DT == 4 ? 96u : 24u;  // staged nodes per P5 sub-iteration
    constexpr uint32_t WALK_BYTES = 8 * 32 * 16;  // the chain-block scratch: 8 cells of 16 bytes per lane
    constexpr uint32_t BYTES_A = CS3_NBINS * 4, BYTES_B = 2 * DT * NB * 8 > WALK_BYTES ? 2 * DT * NB * 8 : WALK_BYTES;
    constexpr uint32_t BYTES_C = 2 * DT * 32 * 8 + 256 * 2 + 256;
    constexpr uint32_t BYTES_W = BYTES_A + BYTES_B + BYTES_C;
    static_assert(BYTES_B >= WALK_BYTES && ((CS3_KMAX + 4) & ~3) + CS3_KMAX + 1 <= 32, "the chain block must fit region B");
    static_assert(3 * NB * 4 <= BYTES_A && NB >= CS3_KMAX, "P5 node staging must fit region A");
    static_assert(CS3_LIST * 4 <= BYTES_A && CS3_LIST * 4 <= BYTES_B && CS3_LIST >= 32 * CS3_KMAX && CS3_NBINS % 32 == 0,
                  "P4 list: ids in region A, costs in region B");
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ uint32_t s_hist_all[WARPS][2][CS_MAX_THRESHOLDS + 1];
    __shared__ float s_rank_all[WARPS][CS_MAX_THRESHOLDS];

    const uint32_t lane = cs_lane();
    const uint32_t ltmask = cs_lanemask_lt();
    const uint32_t wic = cs_warp_in_cta();
    const uint32_t worker = blockIdx.x * WARPS + wic;
    uint8_t* s_warp = s_dyn + (size_t)wic * BYTES_W;
    uint32_t* bins = reinterpret_cast<uint32_t*>(s_warp);
    uint32_t* l_id = bins;
    float* l_cost = reinterpret_cast<float*>(s_warp + BYTES_A);  // P4 only: region B is idle there (no chain walks)
    uint8_t* s_llist2 = reinterpret_cast<uint8_t*>(bins);  // P3: the chunk's links owned by their junction
    uint32_t* s_ids = bins;                        // P5 staged nodes
    float* s_cst = reinterpret_cast<float*>(bins + NB);
    float* s_pcs = reinterpret_cast<float*>(bins + 2 * NB);
    float* cblk = reinterpret_cast<float*>(s_warp + BYTES_A) + lane * 4;  // the lane's column of 16-byte cells (8 rows)
    uint16_t* ttab = reinterpret_cast<uint16_t*>(bins);  // P1: (owner lane | link << 8) of the batch's relaxations
    static_assert(32 * CS3_MAX_LINKS * 2 <= BYTES_A, "P1 task table must fit region A");
    double* s_crd = reinterpret_cast<double*>(s_warp + BYTES_A);
    double* s_acc = reinterpret_cast<double*>(s_warp + BYTES_A + BYTES_B);
    uint16_t* s_llist = reinterpret_cast<uint16_t*>(s_warp + BYTES_A + BYTES_B + 2 * DT * 32 * 8);
    uint8_t* s_info = s_warp + BYTES_A + BYTES_B + 2 * DT * 32 * 8 + 512;
    uint32_t* histN = s_hist_all[wic][0];
    uint32_t* histE = s_hist_all[wic][1];
    float* rankf = s_rank_all[wic];
#define CS3_CB(i) cblk[((i) >> 2) * 128u + ((i) & 3u)]
    const CsWarpArena A = cs_arena(p.arena, p.lay, worker);
    uint8_t* linfo = A.bdone;          // [rcap][8] link bytes: T | tie2 << 4 | yhas << 5
    uint32_t* minsucc = A.node_list;   // [rcap] after P2: smallest rank that has this junction as a predecessor
    // [rcap][8] per link: {candidate seconds bits, neighbour's seconds bits (inf = no candidate), neighbour id | pos << 28,
    // rank of the far junction}, written by the link's OWNER end (the earlier-settled one) for both ends
    uint4* cand = reinterpret_cast<uint4*>(p.arena + (size_t)worker * p.lay.stride + p.lay.frank);
    uint32_t* needm = reinterpret_cast<uint32_t*>(p.arena + (size_t)worker * p.lay.stride + p.lay.needm);  // [rcap]
    // [rcap] {first link, links | in-degree << 8} of the junction at each settle rank: the later phases read it next to
    // s_node / s_agg (one coalesced round trip) instead of chasing jinfo[s_node[r]]
    uint2* jrank = reinterpret_cast<uint2*>(p.arena + (size_t)worker * p.lay.stride + p.lay.jrank);
    const CsV3Graph& g = p.g;
    const uint32_t J = g.J;
    const int D = DT <= 4 ? DT : p.D;  // instantiations up to 4 thresholds are exact: the threshold loops unroll without tests
    const int D2 = 2 * D;
    const float one_minus = 1.0f - CS_TIE_EPS, one_plus = 1.0f + CS_TIE_EPS;
    const float one_plus_tol = 1.0f + p.tol;
    const uint32_t INF = CS_INF_BITS;

    // The warps of a CTA take WARPS consecutive sources and move through the phases together (one barrier per
    // phase): the kernel is far larger than the 32 KB instruction cache level next to the SM, and sixteen warps in
    // sixteen different phases starve on instruction fetch (profiles/r01l: no_instruction was the top stall).
    __shared__ unsigned long long s_base;
    __shared__ int s_err;
    for (;;) {
#if CS3_PHASE_SYNC
        __syncthreads();
        if (threadIdx.x == 0) {
            s_base = atomicAdd(&p.counters[CS_C_NEXT], (unsigned long long)WARPS);
            s_err = *reinterpret_cast<volatile int*>(p.error);
        }
        __syncthreads();
        if (s_base >= p.n_sources || s_err != 0) break;
        const unsigned long long si = cs_uni(s_base) + wic;
        bool run = si < p.n_sources;
#define CS3_BAR() __syncthreads()
#else
        unsigned long long si = 0;
        if (lane == 0) si = atomicAdd(&p.counters[CS_C_NEXT], 1ull);
        si = __shfl_sync(CS_FULL, si, 0);
        if (si >= p.n_sources) break;
        if (*reinterpret_cast<volatile int*>(p.error) != 0) break;
        bool run = true;
#define CS3_BAR()
#endif
        const float wt = cs_uni(run ? __ldg(&p.src_wt[si]) : 0.f);
        CsSrc3 S;
        S.id = cs_uni(run ? __ldg(&g.new_of_orig[__ldg(&p.sources[si])]) : 0u);  // per-source values: warp-uniform
        S.interior = S.id >= J ? 1u : 0u;
        S.slot = S.interior ? J : S.id;
        S.soff = S.ibase = S.k = S.p = S.A = S.B = S.posA = S.posB = 0;
        if (S.interior) {
            const uint32_t c = __ldg(&g.int_chain[S.id - J]);
            const uint4 c0 = __ldg(&g.ctab[2 * c]), c1 = __ldg(&g.ctab[2 * c + 1]);
            S.soff = cs_uni(c0.x);
            S.ibase = cs_uni(c0.y);
            S.k = cs_uni(c0.z);
            S.A = cs_uni(c0.w);
            S.B = cs_uni(c1.x);
            S.posA = cs_uni(c1.y);
            S.posB = cs_uni(c1.z);
            S.p = S.id - J - S.ibase + 1;
        }
        long long tc[7];
        tc[0] = clock64();

        // ------------------------------------------------------------------ P1: junction search (label-correcting)
        unsigned long long relax = 0, edge_iters = 0, n_interior = 0;
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
                            // the staleness test and the junction record are independent loads: one round trip, not two
                            const uint32_t cur = cs_ld(&A.ds[v].x);
                            uint2 ji = make_uint2(0u, 2u);
                            if (v != J) ji = __ldg(&g.jinfo[v]);
                            valid = cur == abits;
                            off = ji.x;
                            deg = valid ? (ji.y & 0xffu) : 0u;
                        }
                        // One lane per LINK: the frontier holds few junctions at a time (about nine per batch on the
                        // 1M-node street graph), so the (junction, link) pairs of the batch are spread over the lanes and
                        // relaxed together - link record, chain seconds and atomicMin are three dependent L2 round
                        // trips per batch instead of three per link of the widest junction.
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
                            float cand = 0.f;
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
                                    cand = a;
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
                        __syncwarp();  // the next batch rewrites the task table
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
                float mn = __uint_as_float(INF);
                for (uint32_t i = lane; i < nf; i += 32) {
                    const uint2 it = cs_ld(&far[i]);
                    if (cs_ld(&A.ds[it.x & CS_NODE_MASK].x) == it.y) mn = fminf(mn, __uint_as_float(it.y));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(CS_FULL, mn, o));
                if (!(mn < __uint_as_float(INF))) break;
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
        tc[1] = clock64();
        if (fail) {
            if (lane == 0) atomicCAS(p.error, 0, fail);
#if CS3_PHASE_SYNC
            run = false;
#else
            break;
#endif
        }
        if (!run) R = 0;  // an idle warp still meets the barriers; every loop below is empty for it
        R = cs_uni(R);
        CS3_BAR();

        // ------------------------------------------------------------------ P2: exact settle order of the junctions
        if (lane <= CS_MAX_THRESHOLDS) {
            histN[lane] = 0;
            histE[lane] = 0;
        }
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
                cs_st(&A.sigma[rank], 0.0);
                cs_st(&minsucc[rank], CS_NOSLOT);  // (node_list is dead from here on)
                cs_st(&needm[rank], 0u);
                const uint2 ji = node == J ? make_uint2(0u, 2u | (2u << 8)) : __ldg(&g.jinfo[node]);
                cs_st(&jrank[rank], ji);
                edge_iters += ji.y >> 8;
            }
            __syncwarp();
        }
        tc[2] = clock64();
        CS3_BAR();

        // ------------------------------------------------------------------ P3: predecessors, sigma, chain ownership
        // P3a, one lane per LINK of the chunk's junctions: walk the chain from both ends, decide which interiors belong
        // to this end, evaluate the meeting pair, and leave the link's candidate predecessor of the junction in shared
        // memory.  P3b, one lane per junction: the reference's sequential rule over its candidates, then sigma.
        // circuit-rank counts per threshold bin, 16 bits per bin in registers (up to three thresholds), flushed to the
        // shared-memory histograms before any lane can overflow
        unsigned long long cntN = 0, cntNE = 0, cntE = 0;
        auto flush_counts = [&]() {
            if constexpr (DT <= 3) {
#pragma unroll
                for (int b = 0; b <= DT; ++b) {
                    uint32_t cn = (uint32_t)(cntN >> (16 * b)) & 0xffffu, cne = (uint32_t)(cntNE >> (16 * b)) & 0xffffu;
                    uint32_t ce = (uint32_t)(cntE >> (16 * b)) & 0xffffu;
                    cn += cne;
                    ce += cne;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        cn += __shfl_xor_sync(CS_FULL, cn, o);
                        ce += __shfl_xor_sync(CS_FULL, ce, o);
                    }
                    if (lane == 0) {
                        histN[b] += cn;
                        histE[b] += ce;
                    }
                }
                cntN = cntNE = cntE = 0;
                __syncwarp();
            }
        };
        for (uint32_t b0 = 0; b0 < R; b0 += 32) {
            if ((b0 & (128u * 32u - 1u)) == 128u * 32u - 32u) flush_counts();
            const uint32_t r = b0 + lane;
            const bool valid = r < R;
            uint32_t v = 0, off = 0, deg = 0, vid = 0, avb = 0;
            if (valid) {
                v = cs_ld(&A.s_node[r]);
                vid = v == J ? S.id : v;
                avb = __float_as_uint(cs_ld(&A.s_agg[r]));
                if (p.closeness) {
                    const int th = cs3_first_threshold<DT>(p, __fmul_rn(__uint_as_float(avb), p.speed));
                    if constexpr (DT <= 3) cntN += 1ull << (16 * th);
                    else atomicAdd(&histN[th], 1u);
                }
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
            // every chain is evaluated once, by its OWNER end: the earlier-settled junction (or the only reached one);
            // first pass: find the owned links of the chunk
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
                if (!act) continue;
                const uint32_t lr = b0 + jl;
                const float av = __uint_as_float(lavb);
                const CsView V = cs3_view(g, S, lv, loff, j);
                const uint2 dF = cs_ld(&A.ds[V.far]);  // issued before the staging wait: the two round trips overlap
                cs3_load_block(g, V, cblk, 0, V.nv);
                const uint32_t k = V.k;
                const uint32_t fid = V.far == J ? S.id : V.far;
                // Both waves at once: my front starts at v, theirs at F; the smaller front advances (the settle order of
                // the chain's nodes), until the fronts meet or neither may advance (cutoff, centrality.rs:1407).  The owner
                // settles first, so it wins exact ties and counts the shared piece.
                const bool f_reached = dF.x != INF;
                float a = av, a_prev = av;                    // distance of m_T (v when T == 0) and of the node before it
                float b = __uint_as_float(dF.x), b_prev = b;  // distance of m_{k+1-jn} (F when jn == 0) and the one before
                float a_next = __fadd_rn(a, CS3_CB(V.sv));    // my candidate for m_{T+1} (for F itself once T == k)
                float b_next = f_reached ? __fadd_rn(b, CS3_CB(V.sF)) : __uint_as_float(INF);  // theirs for m_{k-jn}
                uint32_t T = 0, jn = 0;
                while (T + jn < k) {
                    const bool a_ok = !(a_next > p.max_seconds);
                    const bool b_ok = f_reached && !(b_next > p.max_seconds);
                    if (!a_ok && !b_ok) break;
                    bool take_a = a_ok;
                    if (a_ok && b_ok)  // the last unsettled node is claimed from both sides; otherwise two different nodes
                        take_a = a_next <= b_next;
                    float settled;
                    if (take_a) {
                        if (a_next == a) atomicCAS(p.error, 0, CS_ERR_ZERO_TIE);
                        a_prev = a;
                        a = a_next;
                        ++T;
                        settled = a;
                        a_next = __fadd_rn(a, CS3_CB(V.sv + T));
                    } else {
                        if (b_next == b) atomicCAS(p.error, 0, CS_ERR_ZERO_TIE);
                        b_prev = b;
                        b = b_next;
                        ++jn;
                        settled = b;
                        b_next = __fadd_rn(b, CS3_CB(V.sF + jn));
                    }
                    if (p.closeness) {
                        // a node and the piece towards its own end, whose larger cost is the node's
                        const int th = cs3_first_threshold<DT>(p, __fmul_rn(settled, p.speed));
                        if constexpr (DT <= 3) {
                            cntNE += 1ull << (16 * th);
                        } else {
                            atomicAdd(&histN[th], 1u);
                            atomicAdd(&histE[th], 1u);
                        }
                    }
                }
                n_interior += T + jn;
                uint32_t flags = 0;
                // meeting pair X = m_T (v when T == 0), Y = m_{T+1} (F when T == k): Y is reached iff the fronts met
                if (f_reached && T + jn == k) {
                    const uint32_t yd = __float_as_uint(b);
                    if (p.closeness && V.cnt) {
                        const float ec = __fmul_rn(fmaxf(a, b), p.speed);
                        const int th = cs3_first_threshold<DT>(p, ec);
                        if constexpr (DT <= 3) cntE += (unsigned long long)V.cnt << (16 * th);
                        else atomicAdd(&histE[th], V.cnt);
                    }
                    const uint32_t xid = T == 0 ? lvid : V.id1 + V.step * (int)(T - 1);
                    const uint32_t yid = T == k ? fid : V.id1 + V.step * (int)T;
                    const bool x_later = cs3_before(g, S, yd, yid, __float_as_uint(a), xid);
                    if (x_later && T >= 1) {
                        // X has its own-side predecessor m_{T-1} and, perhaps, Y (candidate: the next step of their wave)
                        const uint32_t pid = T == 1 ? lvid : V.id1 + V.step * (int)(T - 2);
                        const bool own_first = cs3_before(g, S, __float_as_uint(a_prev), pid, yd, yid);
                        if (cs3_other_kept(a, b_next, own_first, p.phase2 != 0, one_plus_tol, p.max_seconds)) flags |= 0x10u;
                    } else if (!x_later && T < k) {
                        // Y (an interior on F's side) has its own predecessor m_{T+2} / F and, perhaps, X
                        const uint32_t qid = T + 1 == k ? fid : V.id1 + V.step * (int)(T + 1);
                        const bool own_first = cs3_before(g, S, __float_as_uint(b_prev), qid, __float_as_uint(a), xid);
                        if (cs3_other_kept(b, a_next, own_first, p.phase2 != 0, one_plus_tol, p.max_seconds)) flags |= 0x20u;
                    }
                }
                // this end: the owner settles before F and before every node reached from F, so the link offers it no
                // candidate predecessor
                cs_st(&cand[(size_t)lr * 8 + j], make_uint4(0u, INF, 0u, dF.y));
                cs_st(&linfo[(size_t)lr * 8 + j], (uint8_t)(T | flags));
                if (f_reached) {
                    // the far end: it owns jn interiors, the roles of the two meeting flags swap, and the neighbour on its side
                    // of the link (m_k, or this junction) is its candidate predecessor when my wave took the whole chain
                    uint32_t c_ud = INF;
                    float c_c = 0.f;
                    const uint32_t nid = k == 0 ? lvid : V.id1 + V.step * (int)(k - 1);
                    if (T == k && dF.y != 0) {
                        const uint32_t ud = __float_as_uint(a);
                        const bool before = k == 0 ? true : cs3_before(g, S, ud, nid, dF.x, fid);
                        if (before && (p.phase2 || !(a_next > p.max_seconds))) {
                            c_ud = ud;
                            c_c = a_next;
                        }
                    }
                    const uint32_t fflags = ((flags & 0x10u) ? 0x20u : 0u) | ((flags & 0x20u) ? 0x10u : 0u);
                    cs_st(&cand[(size_t)dF.y * 8 + V.paf], make_uint4(__float_as_uint(c_c), c_ud, nid | (j << 28), lr));
                    cs_st(&linfo[(size_t)dF.y * 8 + V.paf], (uint8_t)(jn | fflags));
                }
            }
            __syncwarp();
            // P3b: the junction's candidates in settle order of the neighbours, then the sequential rule
            float cc[CS3_MAX_LINKS];
            uint32_t cu[CS3_MAX_LINKS], cd[CS3_MAX_LINKS], crk[CS3_MAX_LINKS], cj[CS3_MAX_LINKS];
            int ncand = 0;
            uint32_t pmask_c = 0;
            if (valid) {
                const float av = __uint_as_float(avb);
                for (uint32_t j0 = 0; j0 < deg; j0 += 4) {
                    // the records of four links are fetched together (independent addresses, one round trip per group)
                    uint4 grp[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        grp[t] = j0 + t < deg ? cs_ld(&cand[(size_t)r * 8 + j0 + t]) : make_uint4(0u, INF, 0u, 0u);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint4 cd4 = grp[t];
                        const uint32_t j = j0 + t;
                        const uint32_t ud = cd4.y;
                        if (ud == INF) continue;
                        const uint32_t uid = cd4.z & 0x0fffffffu, paf = cd4.z >> 28;
                        int q = ncand++;
                        while (q > 0) {
                            const bool gt = cd[q - 1] != ud ? cd[q - 1] > ud
                                            : cu[q - 1] != uid ? cs3_key(g, S, cu[q - 1]) > cs3_key(g, S, uid)
                                                               : (cj[q - 1] >> 8) > paf;
                            if (!gt) break;
                            cc[q] = cc[q - 1];
                            cu[q] = cu[q - 1];
                            cd[q] = cd[q - 1];
                            crk[q] = crk[q - 1];
                            cj[q] = cj[q - 1];
                            --q;
                        }
                        cc[q] = __uint_as_float(cd4.x);
                        cu[q] = uid;
                        cd[q] = ud;
                        crk[q] = cd4.w;
                        cj[q] = j | (paf << 8);
                    }
                }
                if (ncand == 1 && !p.phase2) {
                    pmask_c = 1u;
                } else if (!p.phase2) {
                    float old = __uint_as_float(INF);
                    for (int q = 0; q < ncand; ++q) {
                        const float c = cc[q];
                        if (c < old) {
                            if (c < __fmul_rn(old, one_minus)) pmask_c = 0;
                            pmask_c |= 1u << q;
                            old = c;
                        } else if (c <= __fmul_rn(old, one_plus)) {
                            bool dup = false;
                            for (uint32_t mm = pmask_c; mm; mm &= mm - 1) dup |= cu[__ffs(mm) - 1] == cu[q];
                            if (!dup) pmask_c |= 1u << q;
                        }
                    }
                } else {
                    const float lim = __fmul_rn(av, one_plus_tol);
                    for (int q = 0; q < ncand; ++q) {
                        if (cc[q] <= lim) {
                            bool dup = false;
                            for (uint32_t mm = pmask_c; mm; mm &= mm - 1) dup |= cu[__ffs(mm) - 1] == cu[q];
                            if (!dup) pmask_c |= 1u << q;
                        }
                    }
                }
                for (uint32_t mm = pmask_c; mm; mm &= mm - 1) {
                    const int q = __ffs(mm) - 1;
                    atomicMin(&minsucc[crk[q]], r);  // P5 forms chunks whose junctions do not depend on each other
                    atomicOr(&needm[crk[q]], 1u << (cj[q] >> 8));  // ... and that link of F carries this junction's dependency
                }
                if (r == 0) cs_st(&A.sigma[r], 1.0);
            }
            bool pending = valid && r != 0;
            for (;;) {
                if (pending) {
                    double sg_sum = 0.0;
                    bool ok = true;
                    for (uint32_t mm = pmask_c; mm; mm &= mm - 1) {
                        const double sg = cs_ld(&A.sigma[crk[__ffs(mm) - 1]]);
                        if (sg == 0.0) {
                            ok = false;
                            break;
                        }
                        sg_sum += sg;
                    }
                    if (ok) {
                        if (sg_sum == 0.0) {
                            atomicCAS(p.error, 0, CS_ERR_ZERO_TIE);  // reached without an earlier-settled predecessor
                            sg_sum = 1.0;
                        }
                        cs_st(&A.sigma[r], sg_sum);
                        pending = false;
                    }
                }
                __syncwarp();
                if (!__any_sync(CS_FULL, pending)) break;
            }
        }
        flush_counts();
        __syncwarp();
        tc[3] = clock64();
        CS3_BAR();

        // ------------------------------------------------------------------ P4: closeness scatter to targets
        unsigned long long n_ri = 0, n_ci = 0;
        uint32_t n_chunks = 0, n_batches = 0;
        float cycles_wt = 0.f;
        if (p.closeness) {
            if (lane < (uint32_t)D) {
                long long ncount = 0, ecount = 0;
                for (int t = 0; t <= (int)lane; ++t) {
                    ncount += histN[t];
                    ecount += histE[t];
                }
                rankf[lane] = ncount == 0 ? 0.0f : (float)max(ecount - ncount + 1ll, 0ll);
                atomicAdd(&p.counters[CS_C_REACH0 + lane], (unsigned long long)(ncount > 0 ? ncount - 1 : 0));
            }
            __syncwarp();
            cycles_wt = __fdiv_rn(wt, __ldg(&g.weight[S.id]));  // centrality.rs:1730
            for (uint32_t b0 = 0; b0 < R; b0 += 32) {
                const uint32_t r = b0 + lane;
#if CS3_PREFETCH
                if (r + CS3_PREFETCH < R) {
                    cs3_prefetch(&A.s_node[r + CS3_PREFETCH]);
                    cs3_prefetch(&A.s_agg[r + CS3_PREFETCH]);
                    cs3_prefetch(&jrank[r + CS3_PREFETCH]);
                    cs3_prefetch(linfo + (size_t)(r + CS3_PREFETCH) * 8);
                }
#endif
                uint32_t v = 0, off = 0, deg = 0;
                float av = 0.f;
                unsigned long long info8 = 0ull;
                if (r < R) {
                    v = cs_ld(&A.s_node[r]);
                    av = cs_ld(&A.s_agg[r]);
                    info8 = cs_ld(reinterpret_cast<const unsigned long long*>(linfo + (size_t)r * 8));
                    const uint2 ji = cs_ld(&jrank[r]);
                    off = ji.x;
                    deg = ji.y & 0xffu;
                }
                // the junctions of this chunk (the source is not a target)
                __syncwarp();
                l_id[lane] = v;
                l_cost[lane] = (r < R && r != 0) ? __fmul_rn(av, p.speed) : __uint_as_float(INF);
                __syncwarp();
                cs3_emit_closeness<DT>(p, l_id, l_cost, min(32u, R - b0), wt, cycles_wt, rankf, n_ri);
                // their own-side interiors, link by link (when betweenness runs too, P5 stages the same nodes and
                // scatters their closeness terms there)
                const uint32_t maxdeg = p.betweenness ? 0u : __reduce_max_sync(CS_FULL, deg);
                for (uint32_t j = 0; j < maxdeg; ++j) {
                    const uint32_t T = j < deg ? (uint32_t)(info8 >> (8 * j)) & 15u : 0u;
                    uint32_t inc = T;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(CS_FULL, inc, o);
                        if ((int)lane >= o) inc += t;
                    }
                    const uint32_t total = __shfl_sync(CS_FULL, inc, 31);
                    if (total == 0) continue;
                    __syncwarp();
                    if (T) {
                        const CsView V = cs3_view(g, S, v, off, j);
                        const uint32_t pos = inc - T;
                        float a = av;
                        const float* sec = g.csec + V.blk + V.sv;
                        for (uint32_t t0 = 0; t0 < T; t0 += 4) {  // the loads of four pieces are in flight together
                            float sx[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) sx[u] = t0 + u < T ? __ldg(sec + t0 + u) : 0.f;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (t0 + u < T) {
                                    a = __fadd_rn(a, sx[u]);
                                    l_id[pos + t0 + u] = V.id1 + V.step * (int)(t0 + u);
                                    l_cost[pos + t0 + u] = __fmul_rn(a, p.speed);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    cs3_emit_closeness<DT>(p, l_id, l_cost, total, wt, cycles_wt, rankf, n_ri);
                }
            }
        }
        __syncwarp();
        tc[4] = clock64();
        CS3_BAR();

        // ------------------------------------------------------------------ P5: dependencies, reverse settle order
        // Chunks are formed so that no junction of a chunk depends on another one of the same chunk (minsucc): nothing
        // waits.  Per chunk: one lane per LINK stages its own-side interiors; the seeds (f64 exp) are formed with one lane
        // per node; the lane of the link runs the sequential f64 recurrence of its chain (centrality.rs:823-873) over the
        // staged seeds and leaves the credits in shared memory; they are scattered with one lane per node (consecutive
        // ids, metric-major rows); the link's outflow is added to its junction in link order.
        if (p.betweenness) {
            const double wt_d = (double)wt;
            const unsigned long long od_lo = OD && run ? __ldg(&p.od_off[si]) : 0ull;
            const unsigned long long od_hi = OD && run ? __ldg(&p.od_off[si + 1]) : 0ull;
            int hi = (int)R - 1;
            while (hi >= 0) {
                ++n_chunks;
#if CS3_PREFETCH
                {
                    const int pr = hi - (int)CS3_PREFETCH - (int)lane;
                    if (pr >= 0) {
                        cs3_prefetch(&minsucc[pr]);
                        cs3_prefetch(&A.s_node[pr]);
                        cs3_prefetch(&A.s_agg[pr]);
                        cs3_prefetch(&A.sigma[pr]);
                        cs3_prefetch(&needm[pr]);
                        cs3_prefetch(&jrank[pr]);
                        cs3_prefetch(linfo + (size_t)pr * 8);
                        cs3_prefetch(&cand[(size_t)pr * 8]);
                    }
                }
#endif
                const int rr = hi - (int)lane;
                // the chunk boundary (minsucc) and the per-rank state of the 32 candidate junctions are loaded together;
                // lanes beyond the boundary discard what they fetched
                uint32_t ms = 0u, w = 0, off = 0, deg = 0, nm = 0;
                float aw = 0.f;
                double sigma_w = 1.0;
                unsigned long long info8 = 0ull;
                if (rr >= 0) {
                    ms = cs_ld(&minsucc[rr]);
                    w = cs_ld(&A.s_node[rr]);
                    aw = cs_ld(&A.s_agg[rr]);
                    sigma_w = cs_ld(&A.sigma[rr]);
                    nm = cs_ld(&needm[rr]);
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
                for (int i = 0; i < 2 * DT; ++i) s_acc[i * 32 + lane] = 0.0;
                __syncwarp();
                for (uint32_t base = 0; base < totalL; base += 32) {
                    ++n_batches;
                    const uint32_t e = base + lane;
                    const bool act = e < totalL;
                    const uint32_t code = act ? s_llist[e] : 0u;
                    const uint32_t jl = code & 31u, j = code >> 5;
                    const uint32_t lw = __shfl_sync(CS_FULL, w, jl), loff = __shfl_sync(CS_FULL, off, jl);
                    const float law = __shfl_sync(CS_FULL, aw, jl);
                    const double lsig = __shfl_sync(CS_FULL, sigma_w, jl);
                    const uint32_t lnm = __shfl_sync(CS_FULL, nm, jl);
                    const uint32_t lr = (uint32_t)(hi - (int)jl);
                    uint32_t T = 0;
                    bool tie2 = false, work = false;
                    double dl[DT], dlb[DT];  // dependency flowing toward the junction along this link
#pragma unroll
                    for (int i = 0; i < DT; ++i) dl[i] = dlb[i] = 0.0;
                    CsView V;
                    V.sv = V.id1 = 0;
                    V.step = 1;
                    double fT = 1.0;
                    if (act) {
                        const uint32_t ib = s_info[jl * 8 + j];
                        T = ib & 15u;
                        tie2 = (ib & 0x10u) != 0;
                        const bool yhas = (ib & 0x20u) != 0;
                        // F continues the path through this link iff F chose it as a predecessor (P3b left the bit)
                        const bool needF = (lnm >> j) & 1u;
                        work = T > 0 || needF || yhas;
                        if (work) V = cs3_view(g, S, lw, loff, j);
                        const uint32_t k = V.k;
                        const uint32_t rankF = (needF || yhas || tie2) ? cs_ld(&cand[(size_t)lr * 8 + j].w) : 0u;
                        double sigma_F = 0.0;
                        if (needF || yhas || tie2) sigma_F = cs_ld(&A.sigma[rankF]);
                        if (needF) {
                            // the whole chain is on this side and F continues the path (centrality.rs:861-866)
                            const double f = (sigma_F == lsig) ? 1.0 : lsig / sigma_F;
                            const double* dx = A.dep + (size_t)rankF * D2;
#pragma unroll
                            for (int i = 0; i < DT; ++i) {
                                if (i < D) {
                                    dl[i] = f * cs_ld(&dx[i]);
                                    dlb[i] = f * cs_ld(&dx[D + i]);
                                }
                            }
                        } else if (yhas) {
                            // Y = m_{T+1} was reached from F but keeps m_T (or this junction) as a second predecessor:
                            // it is the last-settled node of the chain, so its dependency is its seed
                            float b = cs_ld(&A.s_agg[rankF]);
                            for (uint32_t t = 0; t < k - T; ++t) b = __fadd_rn(b, __ldg(&g.csec[V.blk + V.sF + t]));
                            const float cost_y = __fmul_rn(b, p.speed);
                            const uint32_t yid = V.id1 + V.step * (int)T;
                            const double pc = cs3_pc<OD>(p, od_lo, od_hi, yid);
                            const double f = lsig / (sigma_F + lsig);
#pragma unroll
                            for (int i = 0; i < DT; ++i) {
                                if (i < D && cost_y <= p.dist_f[i]) {
                                    dl[i] = f * pc;
                                    dlb[i] = f * (pc * cs3_exp(-p.beta_d[i] * (double)cost_y));
                                }
                            }
                        }
                        // m_T with two predecessors carries sigma_w + sigma_F paths
                        if (tie2) fT = lsig / (lsig + sigma_F);
                    }
                    // sub-iterations: as many links as fit the node staging area
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
                            for (uint32_t t0 = 0; t0 < T; t0 += 4) {  // the loads of four pieces are in flight together
                                float sx[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) sx[u] = t0 + u < T ? __ldg(sec + t0 + u) : 0.f;
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    if (t0 + u < T) {
                                        a = __fadd_rn(a, sx[u]);
                                        s_ids[offs + t0 + u] = V.id1 + V.step * (int)(t0 + u);
                                        s_cst[offs + t0 + u] = __fmul_rn(a, p.speed);
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        // seeds, one lane per node (centrality.rs:1802-1806).  (Two nodes per lane and iteration, to give the
                        // scheduler independent instruction streams, cost 9 % on the bench: register pressure.)
                        auto seed_node = [&](const uint32_t en) {
                            const float cost = s_cst[en];
                            const uint32_t nid = s_ids[en];
                            const double pc = cs3_pc<OD>(p, od_lo, od_hi, nid);
                            s_pcs[en] = (float)pc;  // 0.5 / 1 or an f32 trip weight: exact
                            if (p.closeness) {
                                // the closeness terms of the same node (P4 leaves the interiors to this pass when both
                                // metric families run): centrality.rs:1755-1777, f32 terms
                                double* base = p.acc_c + nid;
                                const double far_t = (double)__fmul_rn(cost, wt);
                                const double harm_t = (double)__fmul_rn(__fdiv_rn(1.0f, cost), wt);
#pragma unroll
                                for (int i = 0; i < DT; ++i) {
                                    if (i < D && cost <= p.dist_f[i]) {
                                        ++n_ri;
                                        double* q = base + (size_t)(5 * i) * g.n;
                                        cs_red_add(q, (double)wt);
                                        cs_red_add(q + g.n, far_t);
                                        cs_red_add(q + 2 * (size_t)g.n, (double)__fmul_rn(rankf[i], cycles_wt));
                                        cs_red_add(q + 3 * (size_t)g.n, harm_t);
                                        cs_red_add(q + 4 * (size_t)g.n, (double)__fmul_rn(expf(__fmul_rn(-p.beta_f[i], cost)), wt));
                                    }
                                }
                            }
                            if (p.beta_chain) {
                                // beta_i = 2 * beta_{i+1} exactly (e.g. 500 / 1000 / 2000 m): exp(-beta_i c) is the square
                                // of exp(-beta_{i+1} c) to within an ulp or two
                                double ex = exp(-p.beta_d[D - 1] * (double)cost);
#pragma unroll
                                for (int i = DT - 1; i >= 0; --i) {
                                    if (i < D) {
                                        s_crd[(2 * i + 1) * NB + en] = cost <= p.dist_f[i] ? pc * ex : 0.0;
                                        ex *= ex;
                                    }
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < DT; ++i)
                                    if (i < D) s_crd[(2 * i + 1) * NB + en] = cost <= p.dist_f[i] ? pc * cs3_exp(-p.beta_d[i] * (double)cost) : 0.0;
                            }
                        };
                        for (uint32_t e0 = 0; e0 < total; e0 += 32) {
                            const uint32_t en = e0 + lane;
                            if (en < total) seed_node(en);
                        }
                        __syncwarp();
                        if (go) {
                            for (uint32_t t = T; t >= 1; --t) {
                                const uint32_t en = offs + t - 1;
                                const float cost = s_cst[en];
                                const double pc = (double)s_pcs[en];
                                const double f = t == T ? fT : 1.0;
#pragma unroll
                                for (int i = 0; i < DT; ++i) {
                                    if (i < D) {
                                        const double seed = cost <= p.dist_f[i] ? pc : 0.0;
                                        const double seedb = s_crd[(2 * i + 1) * NB + en];
                                        const double dpn = seed + dl[i], dpb = seedb + dlb[i];
                                        s_crd[(2 * i) * NB + en] = dpn - seed;
                                        s_crd[(2 * i + 1) * NB + en] = dpb - seedb;
                                        dl[i] = f * dpn;
                                        dlb[i] = f * dpb;
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        // credits, one lane per node
                        for (uint32_t e0 = 0; e0 < total; e0 += 32) {
                            const uint32_t en = e0 + lane;
                            if (en < total) {
                                double* col = p.acc_b + s_ids[en];
#pragma unroll
                                for (int i = 0; i < DT; ++i) {
                                    if (i < D) {
                                        const double credit = s_crd[(2 * i) * NB + en], creditb = s_crd[(2 * i + 1) * NB + en];
                                        if (credit > 0.0 || creditb > 0.0) {
                                            ++n_ci;
                                            if (credit > 0.0) cs_red_add(col + (size_t)(2 * i) * g.n, credit * wt_d);
                                            if (creditb > 0.0) cs_red_add(col + (size_t)(2 * i + 1) * g.n, creditb * wt_d);
                                        }
                                    }
                                }
                            }
                        }
                        // outflow into the junction (a handful of non-negative f64 terms per junction: the order of
                        // the shared-memory atomics does not matter beyond the last bit)
                        {
                            // the links of a junction sit in consecutive lanes (at most CS3_MAX_LINKS = 8): a segmented
                            // shuffle reduction and one plain add by the run's first lane replace per-link f64 atomics
                            // on shared memory (compare-and-swap loops: 6.6 % of the samples in profiles/r01z)
                            const uint32_t key = act ? jl : 64u + lane;
                            const uint32_t kprev = __shfl_up_sync(CS_FULL, key, 1);
                            const bool head = act && (lane == 0 || kprev != key);
                            double ov[2 * DT];
#pragma unroll
                            for (int i = 0; i < DT; ++i) {
                                ov[i] = go ? dl[i] : 0.0;
                                ov[DT + i] = go ? dlb[i] : 0.0;
                            }
#pragma unroll
                            for (int o = 1; o < (int)CS3_MAX_LINKS; o <<= 1) {
                                const bool same = __shfl_down_sync(CS_FULL, key, o) == key && lane + o < 32u;
#pragma unroll
                                for (int q = 0; q < 2 * DT; ++q) {
                                    const double t = __shfl_down_sync(CS_FULL, ov[q], o);
                                    if (same) ov[q] += t;
                                }
                            }
                            if (head) {
#pragma unroll
                                for (int i = 0; i < DT; ++i) {
                                    if (i < D) {
                                        s_acc[i * 32 + jl] += ov[i];
                                        s_acc[(DT + i) * 32 + jl] += ov[DT + i];
                                    }
                                }
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
                    const float cost_w = __fmul_rn(aw, p.speed);
                    const double pc = is_src ? 0.0 : cs3_pc<OD>(p, od_lo, od_hi, wid);
                    double* dr = A.dep + (size_t)r * D2;
                    double* col = p.acc_b + wid;
#pragma unroll
                    for (int i = 0; i < DT; ++i) {
                        if (i < D) {
                            double seed = 0.0, seedb = 0.0;
                            if (!is_src && cost_w <= p.dist_f[i]) {
                                seed = pc;
                                seedb = pc * cs3_exp(-p.beta_d[i] * (double)cost_w);
                            }
                            const double dpn = seed + s_acc[i * 32 + lane], dpb = seedb + s_acc[(DT + i) * 32 + lane];
                            cs_st(&dr[i], dpn);
                            cs_st(&dr[D + i], dpb);
                            if (!is_src) {
                                const double credit = dpn - seed, creditb = dpb - seedb;
                                if (credit > 0.0 || creditb > 0.0) {
                                    ++n_ci;
                                    if (credit > 0.0) cs_red_add(col + (size_t)(2 * i) * g.n, credit * wt_d);
                                    if (creditb > 0.0) cs_red_add(col + (size_t)(2 * i + 1) * g.n, creditb * wt_d);
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                hi -= (int)cnt;
            }
        }
        tc[5] = clock64();
        CS3_BAR();

        // ------------------------------------------------------------------ P6: reset the dense map
        cs_p6_reset(A, R);
        tc[6] = clock64();

        edge_iters = cs_warp_sum(edge_iters);
        relax = cs_warp_sum(relax);
        n_ri = cs_warp_sum(n_ri);
        n_ci = cs_warp_sum(n_ci);
        n_interior = cs_warp_sum(n_interior);
        if (lane == 0 && run) {
            atomicAdd(&p.counters[CS_C_SOURCES], 1ull);
            atomicAdd(&p.counters[CS_C_SETTLED], (unsigned long long)R + n_interior);
            atomicAdd(&p.counters[CS_C_EDGE_ITERS], edge_iters + 2ull * n_interior);
            atomicAdd(&p.counters[CS_C_RELAX], relax);
            if (n_ri) atomicAdd(&p.counters[CS_C_SUM_RI], n_ri);
            if (n_ci) atomicAdd(&p.counters[CS_C_SUM_CI], n_ci);
            atomicAdd(&p.counters[CS_C_PROGRESS], 1ull);
#pragma unroll
            for (int k = 0; k < 6; ++k) atomicAdd(&p.counters[CS_C_PHASE0 + k], (unsigned long long)(tc[k + 1] - tc[k]));
            atomicAdd(&p.counters[CS_C_PHASE0 + 6], (unsigned long long)n_chunks);   // dependency chunks
            atomicAdd(&p.counters[CS_C_PHASE0 + 7], (unsigned long long)n_batches);  // 32-link batches
        }
    }
#undef CS3_CB
#undef CS3_BAR
}

// Epilogue for the metric-major accumulators of the chain-contracted kernel: new-id columns -> [7][D][node_bound] in
// original index order.
__global__ void cs_k_epilogue_shortest3(const double* __restrict__ acc_c, const double* __restrict__ acc_b, double* out,
                                        const uint32_t* __restrict__ orig_of_new, uint32_t n, int D, int closeness,
                                        int betweenness, int add) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t node = orig_of_new[v];
    for (int i = 0; i < D; ++i) {
        for (int m = 0; m < 5; ++m) {
            double* o = out + ((size_t)(m * D + i)) * n + node;
            const double val = closeness ? acc_c[(size_t)(5 * i + m) * n + v] : 0.0;
            if (closeness || !add) *o = add ? *o + val : val;
        }
        for (int b = 0; b < 2; ++b) {
            double* o = out + ((size_t)((5 + b) * D + i)) * n + node;
            const double val = betweenness ? acc_b[(size_t)(2 * i + b) * n + v] : 0.0;
            if (betweenness || !add) *o = add ? *o + val : val;
        }
    }
}

template <int DT>
static constexpr uint32_t cs3_smem_bytes() {
    constexpr uint32_t NB = DT <= 3 ? CS3_NB3 : DT == 4 ? 96u : 24u;
    constexpr uint32_t WALK_BYTES = 8 * 32 * 16;
    constexpr uint32_t B = 2 * DT * NB * 8 > WALK_BYTES ? 2 * DT * NB * 8 : WALK_BYTES;
    return cs3_warps<DT>() * (CS3_NBINS * 4 + B + 2 * DT * 32 * 8 + 256 * 2 + 256);
}
