// Microbenchmark: throughput of red.global.add.f64 under the access patterns the closeness scatter can choose from.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mb/red_bench scripts/mb/red_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_add(double* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
// pattern 0: every lane its own random 128-B line anywhere in the buffer (node-major metric arrays, v1 layout)
// pattern 1: lanes 0-15 / 16-31 cover 16 consecutive doubles of two random nodes (node-interleaved layout)
// pattern 2: like 0 but random inside a 2 MB window per CTA (spatially coherent sources)
// pattern 3: like 1 but inside a 2 MB window per CTA
// pattern 4: all 32 lanes cover 32 consecutive doubles of one random node (window)
// pattern 5: like 3, only 5 of each 16 lanes active (one threshold admitted)
__global__ void k(double* buf, size_t n_lines, int pattern, int iters) {
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t win_lines = (2u << 20) / 128;
    const size_t win_base = (size_t)(mix(blockIdx.x * 7919u + 13u) % (uint32_t)(n_lines - win_lines));
    uint32_t s = warp * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
        s = mix(s + it);
        size_t line; uint32_t off; bool act = true;
        if (pattern == 0) { line = mix(s ^ (lane * 0x9e3779b9u)) % n_lines; off = lane & 15; }
        else if (pattern == 1) { line = mix(s ^ ((lane >> 4) * 0x9e3779b9u)) % n_lines; off = lane & 15; }
        else if (pattern == 2) { line = win_base + mix(s ^ (lane * 0x9e3779b9u)) % win_lines; off = lane & 15; }
        else if (pattern == 3) { line = win_base + mix(s ^ ((lane >> 4) * 0x9e3779b9u)) % win_lines; off = lane & 15; }
        else if (pattern == 4) { line = win_base + (mix(s) % (win_lines - 1)); off = lane; }
        else { line = win_base + mix(s ^ ((lane >> 4) * 0x9e3779b9u)) % win_lines; off = lane & 15; act = (lane & 15) < 5; }
        if (act) red_add(buf + line * 16 + off, 1.0);
    }
}
int main() {
    const size_t bytes = 192u << 20;
    double* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    const size_t n_lines = bytes / 128 - 2;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int threads : {256, 1024}) for (int pattern = 0; pattern < 6; ++pattern) {
        const int grid = 148 * (2048 / threads), iters = 2000;
        k<<<grid, threads>>>(buf, n_lines, pattern, 100);
        cudaEventRecord(a);
        k<<<grid, threads>>>(buf, n_lines, pattern, iters);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double lanes = (double)grid * threads * iters * (pattern == 5 ? 10.0 / 32.0 : 1.0);
        printf("threads %4d pattern %d: %.3f ms  %.1f G lane-reds/s  (%.2f cyc/lane/SM @1.9GHz)\n", threads, pattern, ms,
               lanes / ms / 1e6, 148.0 * 1.9e9 / (lanes / (ms * 1e-3)));
    }
    printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
