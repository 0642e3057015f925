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

// every kernel here is launched with a one-dimensional block of whole warps: %laneid == threadIdx.x & 31.  The special
// register read is one instruction where the mask needs two, and the compiler re-reads it at most uses rather than
// keeping a register (8 % of the arena kernel's instructions were this line, profiles/r02z_ncu_shortest_arena.txt)
__device__ __forceinline__ uint32_t cs_lane() {
    uint32_t l;
    asm("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
// index of the warp inside its CTA, broadcast from lane 0 so that the compiler knows it is the same in every lane: what is
// derived from it (the warp's shared-memory window, its arena base) can then live in uniform registers instead of
// being recomputed from the thread index at every use
__device__ __forceinline__ uint32_t cs_warp_in_cta() { return __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0); }
// a value every lane holds alike (loaded from one address, read from a shared variable ...), restated as a broadcast from
// lane 0: the compiler then knows it is warp-uniform and may keep it in a uniform register
template <class T>
__device__ __forceinline__ T cs_uni(T v) { return __shfl_sync(0xffffffffu, v, 0); }
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
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// f32 exp as the reference's platform libm computes it.  Rust's f32::exp calls expf; on Linux that is glibc's table-driven
// algorithm (sysdeps/ieee754/flt-32/e_expf.c, from ARM optimized-routines): exp(x) = 2^(k/32) * 2^(r/32) with a 32-entry
// table and a cubic in double precision, rounded once to f32.  Restated here because segment_centrality forms DIFFERENCES
// of two f32 exponentials (centrality.rs:2281-2300, :2380-2391): a last-ulp disagreement in either one is amplified by the
// cancellation far beyond 1e-5.  The double result carries ~2^-34 relative error before the final rounding, so fused vs
// unfused multiply-adds on the host cannot change the f32 value (checked against libm: 0 mismatches in 4e5 samples).
__device__ const unsigned long long cs_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull,
    0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull,
    0x3feedea64c123422ull, 0x3feece086061892dull, 0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull,
    0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, 0x3feee89f995ad3adull,
    0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
    0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
__device__ __forceinline__ float cs_expf_libm(float x) {
    if (!(x >= -104.0f)) return x != x ? x : 0.0f;  // underflow (and NaN)
    if (x > 88.72284f) return __uint_as_float(CS_INF_BITS);
    const double z = 0x1.71547652b82fep+0 * 32.0 * (double)x;
    // round to nearest integer with the 1.5 * 2^52 shift (glibc's non-intrinsic path): the integer sits in the low
    // mantissa bits of kd + shift
    const double shift = 0x1.8p52;
    double kd = __dadd_rn(z, shift);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dadd_rn(kd, -shift);
    const double r = __dadd_rn(z, -kd);
    const unsigned long long t = __ldg(&cs_exp2f_tab[ki & 31ull]) + (ki << 47);
    const double s = __longlong_as_double((long long)t);
    const double c0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0, c1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
    const double c2 = 0x1.62e42ff0c52d6p-1 / 32.0;
    // fused or unfused does not matter for the f32 result (the double carries ~2^-34 relative error before rounding)
    const double zz = fma(c0, r, c1);
    const double r2 = r * r;
    double y = fma(c2, r, 1.0);
    y = fma(zz, r2, y);
    y = y * s;
    return (float)y;
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
