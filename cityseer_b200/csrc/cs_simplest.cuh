// centrality_simplest (angular) on the GPU (reference: /root/reference/rust/src/centrality.rs:533-566, :577-775,
// :793-821, :1986-2126).  One warp per source.
//
// The angular search is not a plain shortest-path problem: the priority is the angular cost, the cutoff is the travel
// time carried along whichever route currently wins under an order-dependent update rule (:658-703), so the settle
// ORDER is part of the result.  The search therefore replays the reference exactly: lane 0 owns a binary heap with the
// Rust std::collections::BinaryHeap sift rules (ties pop in the same order as upstream); for each popped state the 32
// lanes relax that state's outgoing edges in parallel (one edge per lane, degree <= 32) and lane 0 then pushes the
// improved targets in adjacency order.  Predecessor lists and sigma are built during the search like the reference
// does.  Closeness and the Brandes accumulation afterwards are data-parallel over the settled states.
#pragma once
#include "cs_common.cuh"
#include "cs_heap.cuh"

#define CS_ANG_MAXPRED 8
#define CS_VISITED 0x80000000u
#define CS_SLOT_MASK 0x7fffffffu

struct CsAngLayout {
    size_t ds, dn, st_state, st_secs, st_cost, st_sigma, st_np, st_preds, order, pos, delta, pending, heap, stride;
    uint32_t rcap, hcap;
};

struct CsSimplestParams {
    uint32_t n;
    const uint32_t* out_off;
    const CsEdge* ang_rec;  // {nbr | exit slot << 30 | entry slot << 31, seconds (no impedance), angle_sum, 1e-6 * length}
    int D, closeness, betweenness, phase2;
    float sec_f[CS_MAX_THRESHOLDS];
    float max_seconds, tol, unit, offset;
    const uint32_t* sources;
    const float* src_wt;
    unsigned long long n_sources;
    const uint8_t* eligible;
    double* out;  // [4][D][n]: density, farness, harmonic, betweenness
    unsigned long long* counters;
    int* error;
    uint8_t* arena;
    CsAngLayout lay;
};

enum { CS_ERR_PRED_OVERFLOW = 3 };

__global__ void cs_k_init_ang(uint8_t* arena, size_t stride, size_t ds_off, size_t n_states, size_t dn_off, size_t n_nodes) {
    uint8_t* base = arena + (size_t)blockIdx.y * stride;
    uint2* ds = reinterpret_cast<uint2*>(base + ds_off);
    uint2* dn = reinterpret_cast<uint2*>(base + dn_off);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (size_t k = i; k < n_states; k += step) ds[k] = make_uint2(CS_INF_BITS, CS_NOSLOT);
    for (size_t k = i; k < n_nodes; k += step) dn[k] = make_uint2(CS_INF_BITS, CS_INF_BITS);
}

// Resident CTAs per SM.  With the heap in shared memory the kernel is bound by the latency of its per-warp arena traffic
// (the arenas of all resident warps exceed the L2: ~3.9 MB of DRAM traffic per source, profiles/r02*), and more warps
// only enlarge that working set.  Measured on cfg #3 (bench, heap entries in shared memory / CTAs per SM): 512/3 319 k
// sources/s, 512/4 357 k, 512/5 329 k, 512/6 318 k, 256/5 336 k, 256/6 358 k, 256/8 358 k.
#ifndef CS_ANG_MIN_BLOCKS
#define CS_ANG_MIN_BLOCKS 4
#endif
#ifndef CS_ANG_HEAP_SMEM
#define CS_ANG_HEAP_SMEM 512  // heap entries per warp kept in shared memory (4 KB; 4 CTAs x 8 warps = 128 KB per SM)
#endif
template <int DT>
__global__ void __launch_bounds__(CS_WARPS_PER_CTA * 32, CS_ANG_MIN_BLOCKS) cs_k_simplest(const CsSimplestParams p) {
    const uint32_t lane = cs_lane();
    const uint32_t wic = threadIdx.x >> 5;  // (the broadcast form, cs_warp_in_cta, costs this kernel 1.6 %)
    const uint32_t worker = blockIdx.x * CS_WARPS_PER_CTA + wic;
    const uint32_t ltmask = cs_lanemask_lt();
    uint8_t* base = p.arena + (size_t)worker * p.lay.stride;
    uint2* ds = reinterpret_cast<uint2*>(base + p.lay.ds);  // per state {route cost bits, slot | VISITED}
    uint2* dn = reinterpret_cast<uint2*>(base + p.lay.dn);  // per node  {best route cost bits, best seconds bits}
    uint32_t* st_state = reinterpret_cast<uint32_t*>(base + p.lay.st_state);
    float* st_secs = reinterpret_cast<float*>(base + p.lay.st_secs);
    float* st_cost = reinterpret_cast<float*>(base + p.lay.st_cost);
    double* st_sigma = reinterpret_cast<double*>(base + p.lay.st_sigma);
    uint32_t* st_np = reinterpret_cast<uint32_t*>(base + p.lay.st_np);
    uint32_t* st_preds = reinterpret_cast<uint32_t*>(base + p.lay.st_preds);
    uint32_t* order = reinterpret_cast<uint32_t*>(base + p.lay.order);  // settle position -> slot
    uint32_t* posof = reinterpret_cast<uint32_t*>(base + p.lay.pos);    // slot -> settle position
    double* delta = reinterpret_cast<double*>(base + p.lay.delta);       // [slot][D]
    uint32_t* pending = reinterpret_cast<uint32_t*>(base + p.lay.pending);
    __shared__ uint2 s_heap[CS_WARPS_PER_CTA][CS_ANG_HEAP_SMEM];
    CsHeap heap;
    cs_heap_init(heap, s_heap[wic], CS_ANG_HEAP_SMEM, reinterpret_cast<uint2*>(base + p.lay.heap));
    const uint32_t rcap = p.lay.rcap, hcap = p.lay.hcap;
    const int D = p.D;
    const size_t n = p.n;
    const float one_minus = 1.0f - CS_TIE_EPS, one_plus = 1.0f + CS_TIE_EPS, one_plus_tol = 1.0f + p.tol;
    const float f_inf = __uint_as_float(CS_INF_BITS);

    for (;;) {
        unsigned long long si = 0;
        if (lane == 0) si = atomicAdd(&p.counters[CS_C_NEXT], 1ull);
        si = __shfl_sync(CS_FULL, si, 0);
        if (si >= p.n_sources) break;
        if (*reinterpret_cast<volatile int*>(p.error) != 0) break;
        const uint32_t src = __ldg(&p.sources[si]);
        const float wt = __ldg(&p.src_wt[si]);

        // ------------------------------------------------------------------ search in exact settle order (:598-705)
        uint32_t nslots = 2, nvisited = 0;
        int fail = 0;
        unsigned long long edge_iters = 0;
        cs_heap_clear(heap);
        if (lane == 0) {
            cs_st(&dn[src], make_uint2(0u, 0u));
            for (uint32_t slot = 0; slot < 2; ++slot) {
                cs_st(&ds[src * 2 + slot], make_uint2(0u, slot));
                cs_st(&st_state[slot], src * 2 + slot);
                cs_st(&st_secs[slot], 0.0f);
                cs_st(&st_cost[slot], 0.0f);
                cs_st(&st_sigma[slot], 1.0);
                cs_st(&st_np[slot], 0u);
                cs_heap_push(heap, src * 2 + slot, 0u);
            }
        }
        __syncwarp();
        for (;;) {
            uint32_t state = CS_NOSLOT, sslot = 0;
            if (lane == 0) {
                while (heap.len > 0) {  // lazy deletion: skip states that were settled through an earlier entry
                    const uint2 it = cs_heap_pop(heap);
                    const uint32_t y = cs_ld(&ds[it.x].y);
                    if (y & CS_VISITED) continue;
                    state = it.x;
                    sslot = y;
                    cs_st(&ds[state].y, sslot | CS_VISITED);
                    cs_st(&order[nvisited], sslot);
                    cs_st(&posof[sslot], nvisited);
                    break;
                }
            }
            state = __shfl_sync(CS_FULL, state, 0);
            if (state == CS_NOSLOT) break;
            sslot = __shfl_sync(CS_FULL, sslot, 0);
            nvisited++;
            const uint32_t cur = state >> 1, entry = state & 1u;
            const float Rs = cs_ld(&st_cost[sslot]);
            const float Ts = cs_ld(&st_secs[sslot]);
            const double sig_s = cs_ld(&st_sigma[sslot]);
            const uint32_t eb = __ldg(&p.out_off[cur]);
            const uint32_t deg = __ldg(&p.out_off[cur + 1]) - eb;
            edge_iters += deg;
            bool act = false;
            uint32_t ns = 0, nx = 0;
            float csec = 0.f, cr = 0.f;
            if (lane < deg) {
                const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&p.ang_rec[eb + lane]));
                const uint32_t cslot = (raw.x >> 30) & 1u, nslot = raw.x >> 31;
                nx = raw.x & 0x3fffffffu;
                ns = nx * 2 + nslot;
                csec = __fadd_rn(Ts, __uint_as_float(raw.y));
                cr = __fadd_rn(__fadd_rn(Rs, __uint_as_float(raw.z)), __uint_as_float(raw.w));
                act = (cslot == 1u - entry) && !(csec > p.max_seconds);
            }
            uint2 dsn = make_uint2(CS_INF_BITS, CS_NOSLOT);
            if (act) {
                dsn = cs_ld(&ds[ns]);
                act = !(dsn.y != CS_NOSLOT && (dsn.y & CS_VISITED));
            }
            // lanes that touch the same target node (parallel dual edges) must apply in adjacency order: serialise the pop
            const uint32_t actmask = __ballot_sync(CS_FULL, act);
            bool conflict = false;
            if (act) conflict = __match_any_sync(actmask, nx) != (1u << lane);
            const bool serial = __any_sync(CS_FULL, conflict);
            const uint32_t turns = serial ? deg : 1u;
            uint32_t push_state = CS_NOSLOT, push_bits = 0;
            for (uint32_t turn = 0; turn < turns; ++turn) {
                const bool mine = act && (!serial || lane == turn);
                if (serial && mine) dsn = cs_ld(&ds[ns]);
                const bool is_new = mine && dsn.y == CS_NOSLOT;
                const uint32_t newmask = __ballot_sync(CS_FULL, is_new);
                uint32_t slot = dsn.y & CS_SLOT_MASK;
                if (is_new) slot = nslots + __popc(newmask & ltmask);
                nslots += __popc(newmask);
                if (mine && slot < rcap) {
                    const float cur_cost = __uint_as_float(dsn.x);
                    if (is_new) {
                        cs_st(&st_state[slot], ns);
                        cs_st(&st_np[slot], 0u);
                        cs_st(&st_sigma[slot], 0.0);
                        cs_st(&st_secs[slot], f_inf);
                    }
                    if (cr < cur_cost) {  // improved (:673-686)
                        uint32_t np = is_new ? 0u : cs_ld(&st_np[slot]);
                        double sg = sig_s;
                        if (cr < __fmul_rn(cur_cost, one_minus)) np = 0;
                        else sg += cs_ld(&st_sigma[slot]);
                        if (np >= CS_ANG_MAXPRED) {
                            fail = CS_ERR_PRED_OVERFLOW;
                        } else {
                            cs_st(&st_preds[(size_t)slot * CS_ANG_MAXPRED + np], sslot);
                            cs_st(&st_np[slot], np + 1);
                        }
                        cs_st(&st_sigma[slot], sg);
                        cs_st(&st_cost[slot], cr);
                        cs_st(&st_secs[slot], csec);
                        cs_st(&ds[ns], make_uint2(__float_as_uint(cr), slot));
                        push_state = ns;
                        push_bits = __float_as_uint(cr);
                    } else if (cr <= __fmul_rn(cur_cost, one_plus)) {  // tied (:687-693)
                        const uint32_t np = cs_ld(&st_np[slot]);
                        bool dup = false;
                        for (uint32_t k = 0; k < np; ++k) dup |= cs_ld(&st_preds[(size_t)slot * CS_ANG_MAXPRED + k]) == sslot;
                        if (!dup) {
                            if (csec < cs_ld(&st_secs[slot])) cs_st(&st_secs[slot], csec);
                            if (np >= CS_ANG_MAXPRED) {
                                fail = CS_ERR_PRED_OVERFLOW;
                            } else {
                                cs_st(&st_preds[(size_t)slot * CS_ANG_MAXPRED + np], sslot);
                                cs_st(&st_np[slot], np + 1);
                            }
                            cs_st(&st_sigma[slot], cs_ld(&st_sigma[slot]) + sig_s);
                        }
                    }
                    // node-level bests, updated after every candidate that passed the cutoff / visited tests (:695-703)
                    const uint2 bn = cs_ld(&dn[nx]);
                    const float best = __uint_as_float(bn.x);
                    if (cr < __fmul_rn(best, one_minus)) {
                        cs_st(&dn[nx], make_uint2(__float_as_uint(cr), __float_as_uint(csec)));
                    } else if (cr <= __fmul_rn(best, one_plus)) {
                        cs_st(&dn[nx].y, __float_as_uint(fminf(__uint_as_float(bn.y), csec)));
                    }
                }
                if (serial) __syncwarp();
            }
            __syncwarp();
            // lane 0 pushes the improved targets in adjacency (= lane) order
            const uint32_t pm = __ballot_sync(CS_FULL, push_state != CS_NOSLOT);
            for (uint32_t mm = pm; mm; mm &= mm - 1) {
                const int l = __ffs(mm) - 1;
                const uint32_t s2 = __shfl_sync(CS_FULL, push_state, l);
                const uint32_t b2 = __shfl_sync(CS_FULL, push_bits, l);
                if (lane == 0) {
                    if (heap.len < hcap) cs_heap_push(heap, s2, b2);
                    else fail = CS_ERR_QUEUE_OVERFLOW;
                }
            }
            fail = __reduce_max_sync(CS_FULL, (unsigned)fail);
            if (nslots > rcap) fail = CS_ERR_REACH_OVERFLOW;
            if (fail) break;
        }
        if (fail) {
            if (lane == 0) atomicCAS(p.error, 0, fail);
            break;
        }
        __syncwarp();

        // ------------------------------------------------------------------ phase 2: tolerance predecessors (:710-764)
        if (p.phase2) {
            for (uint32_t k = lane; k < nslots; k += 32) {
                cs_st(&st_np[k], 0u);
                cs_st(&st_sigma[k], k < 2 ? 1.0 : 0.0);
            }
            __syncwarp();
            for (uint32_t pos = 0; pos < nvisited; ++pos) {
                const uint32_t us = cs_ld(&order[pos]);
                const uint32_t ustate = cs_ld(&st_state[us]);
                const uint32_t un = ustate >> 1, uentry = ustate & 1u;
                const float Ru = cs_ld(&st_cost[us]);
                const double sig_u = cs_ld(&st_sigma[us]);
                const uint32_t eb = __ldg(&p.out_off[un]);
                const uint32_t deg = __ldg(&p.out_off[un + 1]) - eb;
                bool act = false;
                uint32_t vslot = 0;
                float cr = 0.f;
                if (lane < deg) {
                    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&p.ang_rec[eb + lane]));
                    const uint32_t cslot = (raw.x >> 30) & 1u, nslot = raw.x >> 31;
                    if (cslot == 1u - uentry) {
                        const uint32_t y = cs_ld(&ds[(raw.x & 0x3fffffffu) * 2 + nslot].y);
                        if (y != CS_NOSLOT) {  // never-reached targets stay inert
                            vslot = y & CS_SLOT_MASK;
                            cr = __fadd_rn(__fadd_rn(Ru, __uint_as_float(raw.z)), __uint_as_float(raw.w));
                            act = cs_ld(&posof[vslot]) > pos && cr <= __fmul_rn(cs_ld(&st_cost[vslot]), one_plus_tol);
                        }
                    }
                }
                // same target from several lanes = duplicate predecessor: only the first (adjacency order) counts
                const uint32_t am = __ballot_sync(CS_FULL, act);
                if (act) act = (__ffs(__match_any_sync(am, vslot)) - 1) == (int)lane;
                if (act) {
                    const uint32_t np = cs_ld(&st_np[vslot]);
                    bool dup = false;
                    for (uint32_t k = 0; k < np; ++k) dup |= cs_ld(&st_preds[(size_t)vslot * CS_ANG_MAXPRED + k]) == us;
                    if (!dup) {
                        if (np >= CS_ANG_MAXPRED) {
                            fail = CS_ERR_PRED_OVERFLOW;
                        } else {
                            cs_st(&st_preds[(size_t)vslot * CS_ANG_MAXPRED + np], us);
                            cs_st(&st_np[vslot], np + 1);
                            cs_st(&st_sigma[vslot], cs_ld(&st_sigma[vslot]) + sig_u);
                        }
                    }
                }
                __syncwarp();
            }
            fail = __reduce_max_sync(CS_FULL, (unsigned)fail);
            if (fail) {
                if (lane == 0) atomicCAS(p.error, 0, fail);
                break;
            }
        }

        // ------------------------------------------------------------------ closeness, seconds thresholds (:1995-2034)
        unsigned long long n_ri = 0, n_ci = 0;
        if (p.closeness) {
            unsigned long long reach[DT];
#pragma unroll
            for (int i = 0; i < DT; ++i) reach[i] = 0;
            const double wt_d = (double)wt;
            const size_t ms = (size_t)D * n;
            for (uint32_t k = lane; k < nslots; k += 32) {
                const uint32_t state = cs_ld(&st_state[k]);
                const uint32_t node = state >> 1;
                if (node == src) continue;
                const uint32_t sib = cs_ld(&ds[state ^ 1u].y);
                if (sib != CS_NOSLOT && (sib & CS_SLOT_MASK) < k) continue;  // one visit per node
                const uint2 bn = cs_ld(&dn[node]);
                if (bn.x == CS_INF_BITS || bn.y == CS_INF_BITS) continue;
                const float simpl = __uint_as_float(bn.x), bsec = __uint_as_float(bn.y);
                const float ratio = __fdiv_rn(simpl, p.unit);
                const float far_t = __fmul_rn(__fadd_rn(p.offset, ratio), wt);
                const float harm_t = __fmul_rn(__fdiv_rn(1.0f, __fadd_rn(1.0f, ratio)), wt);
#pragma unroll
                for (int i = 0; i < DT; ++i) {
                    if (i < D && bsec <= p.sec_f[i]) {
                        double* o = p.out + (size_t)i * n + node;
                        cs_red_add(o, wt_d);
                        cs_red_add(o + ms, (double)far_t);
                        cs_red_add(o + 2 * ms, (double)harm_t);
                        ++n_ri;
                        ++reach[i];
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < DT; ++i) {
                if (i < D) {
                    const unsigned long long t = cs_warp_sum(reach[i]);
                    if (lane == 0 && t) atomicAdd(&p.counters[CS_C_REACH0 + i], t);
                }
            }
        }

        // ------------------------------------------------------------------ betweenness (:2037-2096, :793-873)
        if (p.betweenness) {
            for (uint32_t k = lane; k < nslots; k += 32) {
                cs_st(&pending[k], 0u);
                for (int i = 0; i < D; ++i) cs_st(&delta[(size_t)k * D + i], 0.0);
            }
            __syncwarp();
            // a successor feeds a predecessor only when it sorts strictly before it in the reference's stable
            // cost-descending order, i.e. when its cost is strictly larger (equal-cost pairs are dropped, :777-791)
            for (uint32_t k = lane; k < nslots; k += 32) {
                const float cw = cs_ld(&st_cost[k]);
                const uint32_t np = cs_ld(&st_np[k]);
                for (uint32_t j = 0; j < np; ++j) {
                    const uint32_t pr = cs_ld(&st_preds[(size_t)k * CS_ANG_MAXPRED + j]);
                    if (cw > cs_ld(&st_cost[pr])) atomicAdd(&pending[pr], 1u);
                }
            }
            __syncwarp();
            const double wt_d = (double)wt;
            for (int b0 = (int)((nvisited - 1) & ~31u); b0 >= 0; b0 -= 32) {
                const uint32_t pos = (uint32_t)b0 + lane;
                const bool valid = pos < nvisited;
                uint32_t w = 0, node = 0, np = 0;
                float cw = 0.f, Tw = 0.f;
                double sigma_w = 1.0;
                double seed[DT];
#pragma unroll
                for (int i = 0; i < DT; ++i) seed[i] = 0.0;
                if (valid) {
                    w = cs_ld(&order[pos]);
                    const uint32_t state = cs_ld(&st_state[w]);
                    node = state >> 1;
                    cw = cs_ld(&st_cost[w]);
                    Tw = cs_ld(&st_secs[w]);
                    sigma_w = cs_ld(&st_sigma[w]);
                    np = cs_ld(&st_np[w]);
                    if (node != src) {
                        // best_angular_target_states (:793-821): split the pair weight over the qualifying states by sigma
                        const uint2 bn = cs_ld(&dn[node]);
                        const float bcost = __uint_as_float(bn.x), bsec = __uint_as_float(bn.y);
                        const float lim = __fmul_rn(bcost, one_plus_tol);
                        const uint32_t sib = cs_ld(&ds[state ^ 1u].y);
                        float cs2 = f_inf, Ts2 = f_inf;
                        double sg2 = 0.0;
                        if (sib != CS_NOSLOT) {
                            const uint32_t s2 = sib & CS_SLOT_MASK;
                            cs2 = cs_ld(&st_cost[s2]);
                            Ts2 = cs_ld(&st_secs[s2]);
                            sg2 = cs_ld(&st_sigma[s2]);
                        }
                        const double pc = __ldg(&p.eligible[node]) ? 0.5 : 1.0;
                        if (bn.x != CS_INF_BITS && bn.y != CS_INF_BITS) {
#pragma unroll
                            for (int i = 0; i < DT; ++i) {
                                if (i < D && !(bsec > p.sec_f[i])) {
                                    const float thr = p.sec_f[i];
                                    const bool q1 = sigma_w != 0.0 && !(Tw > thr) && cw <= lim;
                                    const bool q2 = sg2 != 0.0 && !(Ts2 > thr) && cs2 <= lim;
                                    if (q1) {
                                        // sum in slot order (slot 0 first), as the reference iterates the two states
                                        const double tot = q2 ? ((state & 1u) ? sg2 + sigma_w : sigma_w + sg2) : sigma_w;
                                        if (tot != 0.0) seed[i] = pc * (sigma_w / tot);
                                    }
                                }
                            }
                        }
                    }
                }
                bool waiting = valid;
                for (;;) {
                    if (waiting && cs_ld(&pending[w]) == 0u) {
#pragma unroll
                        for (int i = 0; i < DT; ++i) {
                            if (i < D && !(Tw > p.sec_f[i])) {  // include_state: agg_seconds <= threshold (:2087)
                                const double dep = seed[i] + cs_ld(&delta[(size_t)w * D + i]);
                                if (dep != 0.0) {
                                    for (uint32_t j = 0; j < np; ++j) {
                                        const uint32_t pr = cs_ld(&st_preds[(size_t)w * CS_ANG_MAXPRED + j]);
                                        if (cw > cs_ld(&st_cost[pr]))
                                            cs_red_add(&delta[(size_t)pr * D + i], (cs_ld(&st_sigma[pr]) / sigma_w) * dep);
                                    }
                                    if (node != src) {
                                        const double credit = dep - seed[i];
                                        if (credit > 0.0) {
                                            ++n_ci;
                                            cs_red_add(p.out + ((size_t)(3 * D + i)) * n + node, credit * wt_d);
                                        }
                                    }
                                }
                            }
                        }
                        __threadfence_block();
                        for (uint32_t j = 0; j < np; ++j) {
                            const uint32_t pr = cs_ld(&st_preds[(size_t)w * CS_ANG_MAXPRED + j]);
                            if (cw > cs_ld(&st_cost[pr])) atomicSub(&pending[pr], 1u);
                        }
                        waiting = false;
                    }
                    __syncwarp();
                    if (!__any_sync(CS_FULL, waiting)) break;
                }
            }
        }

        // ------------------------------------------------------------------ reset the dense maps
        for (uint32_t k = lane; k < nslots; k += 32) {
            const uint32_t state = cs_ld(&st_state[k]);
            cs_st(&ds[state], make_uint2(CS_INF_BITS, CS_NOSLOT));
            cs_st(&dn[state >> 1], make_uint2(CS_INF_BITS, CS_INF_BITS));
        }
        __syncwarp();

        n_ri = cs_warp_sum(n_ri);
        n_ci = cs_warp_sum(n_ci);
        if (lane == 0) {
            atomicAdd(&p.counters[CS_C_SOURCES], 1ull);
            atomicAdd(&p.counters[CS_C_SETTLED], (unsigned long long)nvisited);
            atomicAdd(&p.counters[CS_C_EDGE_ITERS], edge_iters);
            if (n_ri) atomicAdd(&p.counters[CS_C_SUM_RI], n_ri);
            if (n_ci) atomicAdd(&p.counters[CS_C_SUM_CI], n_ci);
            atomicAdd(&p.counters[CS_C_PROGRESS], 1ull);
        }
    }
}
