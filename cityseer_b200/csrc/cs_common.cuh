// Shared device-side definitions for the centrality kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cityseer_b200.h"

#define CS_FULL 0xffffffffu
#define CS_INF_BITS 0x7f800000u
#define CS_NOSLOT 0xffffffffu
#ifndef CS_NBINS
#define CS_NBINS 1024          // counting-sort bins per source (shared memory, per warp)
#endif
#define CS_WARPS_PER_CTA 8
#define CS_TIE_EPS 1e-4f       // centrality.rs:28

// 16-byte directed-edge record, one 128-bit load per relaxation.
//   in-CSR  (edges nb->cur stored at cur): nbr = nb, sec = travel seconds nb->cur, aux = length of the twin cur->nb,
//            meta[7:0] = position of this edge in nb's out-list, meta[8] = twin exists, meta[9] = self-loop
//   out-CSR (edges v->u stored at v):      nbr = u,  sec = travel seconds v->u,   aux = edge length,
//            meta[7:0] = position of this edge in u's in-list,  meta[8] = canonical representative of its
//            (min,max,edge_idx) group (circuit rank, centrality.rs:500-507), meta[9] = self-loop
struct __align__(16) CsEdge {
    uint32_t nbr;
    float sec;
    float aux;
    uint32_t meta;
};

struct CsGraphDev {
    uint32_t n;             // node_bound
    const uint32_t* in_off; // [n+1]
    const CsEdge* in_rec;
    const uint32_t* out_off;
    const CsEdge* out_rec;
    const float* in_imp;    // [E] twin impedance per in-record (segment closeness)
    const float* weight;    // [n]
    const uint8_t* live;    // [n]
};

// Device counters (one block of 8 u64 + reach totals) accumulated with one atomic per source per field.
enum { CS_C_SOURCES = 0, CS_C_SETTLED, CS_C_EDGE_ITERS, CS_C_SUM_RI, CS_C_SUM_CI, CS_C_RELAX, CS_C_PROGRESS, CS_C_NEXT, CS_C_FALLBACK, CS_C_PHASE0, CS_C_REACH0 = CS_C_PHASE0 + 8 };
#define CS_NCOUNTERS (CS_C_REACH0 + CS_MAX_THRESHOLDS)

enum { CS_ERR_NONE = 0, CS_ERR_REACH_OVERFLOW = 1, CS_ERR_QUEUE_OVERFLOW = 2, CS_ERR_PRED_OVERFLOW_ = 3, CS_ERR_ZERO_TIE = 4 };

__device__ __forceinline__ uint32_t cs_lane() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t cs_lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// L2-only loads/stores for per-warp scratch that is rewritten during the kernel (L1 is not coherent with atomics).
template <class T>
__device__ __forceinline__ T cs_ld(const T* p) { return __ldcg(p); }
template <class T>
__device__ __forceinline__ void cs_st(T* p, T v) { __stcg(p, v); }

__device__ __forceinline__ void cs_red_add(double* p, double v) {
#ifdef CS_EXPERIMENT_NO_RED  // measurement only: how much of the time the f64 scatter costs
    if (v == -1.2345) *p = v;
#else
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
#endif
}
__device__ __forceinline__ unsigned long long cs_warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CS_FULL, v, o);
    return v;
}
__device__ __forceinline__ double cs_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CS_FULL, v, o);
    return v;
}
