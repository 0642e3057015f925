// Host side of the shared-memory centrality_shortest kernel (cs_shortest2.cuh): Hilbert renumbering and the renumbered
// CSR built at upload, device-side staging of the source plan in new-id order, the capacity probe that sizes the
// shared-memory layout, the launch, and the fallback pass through the global-arena kernel.  Included by cs_api.cu.

// ------------------------------------------------------------------------------------------------ Hilbert order
static inline uint64_t hilbert_d(uint32_t x, uint32_t y, int bits) {
    uint64_t d = 0;
    for (uint32_t s = 1u << (bits - 1); s > 0; s >>= 1) {
        const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += (uint64_t)s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) {
                x = s - 1 - x;
                y = s - 1 - y;
            }
            std::swap(x, y);
        }
    }
    return d;
}

// new id of every original index: existing nodes along a Hilbert curve over (x, y) when coordinates are given (identity
// otherwise), absent StableGraph slots last
static void node_order(uint32_t n, const uint8_t* node_exists, const double* xs, const double* ys,
                       std::vector<uint32_t>& orig_of_new) {
    orig_of_new.resize(n);
    std::iota(orig_of_new.begin(), orig_of_new.end(), 0u);
    bool coords = xs != nullptr && ys != nullptr;
    double x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY;
    if (coords) {
        for (uint32_t i = 0; i < n; ++i) {
            if (!node_exists[i]) continue;
            if (!std::isfinite(xs[i]) || !std::isfinite(ys[i])) {
                coords = false;
                break;
            }
            x0 = std::min(x0, xs[i]);
            x1 = std::max(x1, xs[i]);
            y0 = std::min(y0, ys[i]);
            y1 = std::max(y1, ys[i]);
        }
    }
    std::vector<uint64_t> key(n, ~0ull);
    if (coords && x1 >= x0) {
        const double span = std::max(std::max(x1 - x0, y1 - y0), 1e-9);
        const double sc = 65535.0 / span;
        for (uint32_t i = 0; i < n; ++i)
            if (node_exists[i])
                key[i] = hilbert_d((uint32_t)((xs[i] - x0) * sc), (uint32_t)((ys[i] - y0) * sc), 16);
    } else {
        for (uint32_t i = 0; i < n; ++i)
            if (node_exists[i]) key[i] = i;
    }
    std::stable_sort(orig_of_new.begin(), orig_of_new.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
}

// Renumbered CSR for cs_k_shortest2 from the original-order CSR (same adjacency order inside every list).
static int build_v2_graph(cs_graph* g, uint32_t n, const uint8_t* node_exists, const double* xs, const double* ys,
                          const std::vector<uint32_t>& in_off, const std::vector<uint32_t>& out_off,
                          const std::vector<CsEdge>& in_rec, const std::vector<CsEdge>& out_rec,
                          const std::vector<float>& in_num, const std::vector<float>& out_num,
                          const std::vector<float>& weight, const std::vector<uint8_t>& live, uint32_t max_deg) {
    g->v2_ok = false;
    if (max_deg > CS2_MAX_DEG) return 0;
    std::vector<uint32_t> orig_of_new, new_of_orig(n);
    node_order(n, node_exists, xs, ys, orig_of_new);
    for (uint32_t v = 0; v < n; ++v) new_of_orig[orig_of_new[v]] = v;
    const size_t E = in_rec.size();
    std::vector<uint32_t> in2_off(n + 1, 0), out2_off(n + 1, 0);
    for (uint32_t v = 0; v < n; ++v) {
        const uint32_t o = orig_of_new[v];
        in2_off[v + 1] = in2_off[v] + (in_off[o + 1] - in_off[o]);
        out2_off[v + 1] = out2_off[v] + (out_off[o + 1] - out_off[o]);
    }
    std::vector<uint4> node2(n);
    std::vector<CsEdge> in2(E), out2(E);
    std::vector<float> in2_num(E), out2_num(E);
    for (uint32_t v = 0; v < n; ++v) {
        const uint32_t o = orig_of_new[v];
        const uint32_t ideg = in_off[o + 1] - in_off[o], odeg = out_off[o + 1] - out_off[o];
        uint32_t wbits;
        std::memcpy(&wbits, &weight[o], 4);
        node2[v] = make_uint4(in2_off[v], out2_off[v], ideg | (odeg << 8) | ((uint32_t)live[o] << 16), wbits);
        for (uint32_t j = 0; j < ideg; ++j) {
            const CsEdge& r = in_rec[in_off[o] + j];
            const uint32_t nb = new_of_orig[r.nbr];
            CsEdge& w = in2[in2_off[v] + j];
            w.nbr = nb;
            w.sec = 0.f;
            const uint32_t nb_eb = in2_off[nb];
            std::memcpy(&w.aux, &nb_eb, 4);
            const uint32_t nb_deg = in_off[r.nbr + 1] - in_off[r.nbr];
            w.meta = (r.meta & 0x3ffu) | (((r.meta >> 16) & 0xfu) << 16) | (nb_deg << 24);
            in2_num[in2_off[v] + j] = in_num[in_off[o] + j];
        }
        for (uint32_t j = 0; j < odeg; ++j) {
            const CsEdge& r = out_rec[out_off[o] + j];
            CsEdge& w = out2[out2_off[v] + j];
            w.nbr = new_of_orig[r.nbr];
            w.sec = 0.f;
            w.aux = 0.f;
            w.meta = r.meta & 0x3ffu;
            out2_num[out2_off[v] + j] = out_num[out_off[o] + j];
        }
    }
    int rc = 0;
    rc |= upload(&g->d_orig_of_new, orig_of_new);
    rc |= upload(&g->d_new_of_orig, new_of_orig);
    rc |= upload(&g->d_node2, node2);
    rc |= upload(&g->d_in2, in2);
    rc |= upload(&g->d_out2, out2);
    rc |= upload(&g->d_in2_num, in2_num);
    rc |= upload(&g->d_out2_num, out2_num);
    if (rc) return 1;
    CS_CUDA(cudaMalloc(&g->d_eligible2, n));
    CS_CUDA(cudaMalloc(&g->d_probe, 2 * sizeof(uint32_t)));
    g->v2_ok = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ staging (device)
__global__ void cs_k_map_sources(const uint32_t* sources, const uint32_t* new_of_orig, uint32_t* keys, uint32_t* vals,
                                 uint64_t m) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        keys[i] = new_of_orig[sources[i]];
        vals[i] = (uint32_t)i;
    }
}
__global__ void cs_k_gather_f32(const float* src, const uint32_t* idx, float* dst, uint64_t m) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) dst[i] = src[idx[i]];
}
__global__ void cs_k_permute_u8(const uint8_t* src, const uint32_t* orig_of_new, uint8_t* dst, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[orig_of_new[i]];
}
__global__ void cs_k_stride_sample(const uint32_t* src, const float* wt, uint64_t m, uint32_t count, uint32_t* dst,
                                   float* dwt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        const uint64_t k = (uint64_t)i * m / count;
        dst[i] = src[k];
        dwt[i] = wt[k];
    }
}
// fallback positions (into the staged new-id plan) -> original indices and weights for the global-arena kernel
__global__ void cs_k_gather_fallback(const uint32_t* pos, uint64_t m, const uint32_t* sources2, const float* wt2,
                                     const uint32_t* orig_of_new, uint32_t* dst, float* dwt) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        dst[i] = orig_of_new[sources2[pos[i]]];
        dwt[i] = wt2[pos[i]];
    }
}

// fallback positions -> new ids and weights (second pass of the shared-memory kernel at its largest layout)
__global__ void cs_k_gather_fallback_new(const uint32_t* pos, uint64_t m, const uint32_t* sources2, const float* wt2,
                                         uint32_t* dst, float* dwt) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        dst[i] = sources2[pos[i]];
        dwt[i] = wt2[pos[i]];
    }
}

// new-id source plan, ascending, from the staged original-index plan (d_sources / d_src_wt / d_eligible)
static int stage_sources_v2(cs_graph* g, uint64_t m, uint32_t* launches) {
    if (g->sources2_valid && g->n_sources2 == m) return 0;
    if (m > g->sources2_cap) {
        for (void* p : {(void*)g->d_sources2, (void*)g->d_src_wt2, (void*)g->d_sort_keys, (void*)g->d_sort_vals,
                        (void*)g->d_sort_vals2, (void*)g->d_fallback, (void*)g->d_fb_sources, (void*)g->d_fb_wt,
                        (void*)g->d_fb2_sources, (void*)g->d_fb2_wt})
            if (p) cudaFree(p);
        const size_t b = std::max<uint64_t>(m, 1) * 4;
        CS_CUDA(cudaMalloc(&g->d_sources2, b));
        CS_CUDA(cudaMalloc(&g->d_src_wt2, b));
        CS_CUDA(cudaMalloc(&g->d_sort_keys, b));
        CS_CUDA(cudaMalloc(&g->d_sort_vals, b));
        CS_CUDA(cudaMalloc(&g->d_sort_vals2, b));
        CS_CUDA(cudaMalloc(&g->d_fallback, b));
        CS_CUDA(cudaMalloc(&g->d_fb_sources, b));
        CS_CUDA(cudaMalloc(&g->d_fb_wt, b));
        CS_CUDA(cudaMalloc(&g->d_fb2_sources, b));
        CS_CUDA(cudaMalloc(&g->d_fb2_wt, b));
        g->sources2_cap = m;
    }
    if (m) {
        const int blocks = (int)((m + 255) / 256);
        cs_k_map_sources<<<blocks, 256, 0, g->stream>>>(g->d_sources, g->d_new_of_orig, g->d_sort_keys, g->d_sort_vals, m);
        int bits = 1;
        while (bits < 32 && (1ull << bits) < g->n) ++bits;
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, g->d_sort_keys, g->d_sources2, g->d_sort_vals, g->d_sort_vals2, (int)m,
                                        0, bits, g->stream);
        if (need > g->cub_tmp_bytes) {
            if (g->d_cub_tmp) cudaFree(g->d_cub_tmp);
            g->d_cub_tmp = nullptr;
            CS_CUDA(cudaMalloc(&g->d_cub_tmp, need));
            g->cub_tmp_bytes = need;
        }
        CS_CUDA(cub::DeviceRadixSort::SortPairs(g->d_cub_tmp, need, g->d_sort_keys, g->d_sources2, g->d_sort_vals,
                                                g->d_sort_vals2, (int)m, 0, bits, g->stream));
        cs_k_gather_f32<<<blocks, 256, 0, g->stream>>>(g->d_src_wt, g->d_sort_vals2, g->d_src_wt2, m);
        *launches += 4;
    }
    cs_k_permute_u8<<<(g->n + 255) / 256, 256, 0, g->stream>>>(g->d_eligible, g->d_orig_of_new, g->d_eligible2, g->n);
    *launches += 1;
    CS_CUDA(cudaGetLastError());
    g->sources2_valid = true;
    g->n_sources2 = m;
    return 0;
}

// ------------------------------------------------------------------------------------------------ layout / launch
static bool v2_layout(CsV2Smem& M, uint32_t rcap, uint32_t pages, uint32_t pb, int D, int T) {
    std::memset(&M, 0, sizeof(M));
    rcap = (uint32_t)align_up(std::max<uint32_t>(rcap, 64), 64);
    uint32_t TB = (uint32_t)std::ceil((double)std::max<uint32_t>(pages, 8) / 0.85);
    const uint32_t gran = std::max<uint32_t>(1, 32u >> pb);  // S must be a multiple of 32 slots
    TB = (uint32_t)align_up(TB, std::max<uint32_t>(gran, 2));
    M.TB = TB;
    M.pb = pb;
    M.S = TB << pb;
    if (M.S > 65536u || rcap > 65536u) return false;
    M.max_pages = (uint32_t)(TB * 0.92);
    M.rcap = rcap;
    M.NB = (uint32_t)std::max(512, T);
    M.WS = 4 * T;  // sigma ring: predecessors up to 2T ranks behind the chunk are read from shared memory
    M.WD = 2 * T;  // dependency ring: successors up to T ranks beyond the chunk
    const uint32_t ES = 2 * D + 1;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 16);
        return (uint32_t)o;
    };
    M.off_dist = take((size_t)M.S * 4);
    M.off_keys = take((size_t)TB * 4);
    M.off_defer = take((size_t)M.S / 8);
    M.off_rank = take((size_t)M.S * 2);
    M.off_perm = take((size_t)rcap * 2);
    // union region: P1 two queues of 3 words per item | P2 bins + bin-ordered slots | P3-P5 sigma ring + predecessor masks
    const size_t u_p2 = align_up((size_t)(M.NB + 1) * 4, 16) + (size_t)rcap * 2;
    const size_t u_p3 = (size_t)M.WS * 8 + rcap;
    size_t u = std::max(u_p2, u_p3);
    M.QC = (uint32_t)std::min<size_t>(1024, std::max<size_t>(u / 24, 64));
    u = std::max(u, (size_t)M.QC * 24);
    const size_t dep = (size_t)M.WD * (ES * 8) + 16;
    M.off_u = take(u);
    M.off_pmask = M.off_u + M.WS * 8;
    if (dep <= (size_t)M.S * 4) {
        M.off_dep = M.off_dist;  // distances are dead once the dependencies are accumulated
    } else {
        M.off_dep = take(dep);
    }
    M.total = (uint32_t)off;
    return true;
}

template <int DT, int T>
static cudaError_t v2_launch_t(const CsShortest2Params& p, uint32_t grid, cudaStream_t st, int* occ) {
    cudaError_t e = cudaFuncSetAttribute(cs_k_shortest2<DT, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.sm.total);
    if (e != cudaSuccess) return e;
    if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, cs_k_shortest2<DT, T>, T, p.sm.total);
    cs_k_shortest2<DT, T><<<grid, T, p.sm.total, st>>>(p);
    return cudaGetLastError();
}
template <int T>
static cudaError_t v2_launch_d(const CsShortest2Params& p, uint32_t grid, cudaStream_t st, int* occ) {
    switch (cs_shortest_dt(p.D)) {
        case 1: return v2_launch_t<1, T>(p, grid, st, occ);
        case 2: return v2_launch_t<2, T>(p, grid, st, occ);
        case 3: return v2_launch_t<3, T>(p, grid, st, occ);
        case 4: return v2_launch_t<4, T>(p, grid, st, occ);
        default: return v2_launch_t<8, T>(p, grid, st, occ);
    }
}
static cudaError_t v2_launch(const CsShortest2Params& p, int T, uint32_t grid, cudaStream_t st, int* occ) {
    if (T == 128) return v2_launch_d<128>(p, grid, st, occ);
    if (T == 256) return v2_launch_d<256>(p, grid, st, occ);
    return v2_launch_d<512>(p, grid, st, occ);
}

static const size_t CS2_SMEM_MAX = 227 * 1024 - 2048;  // opt-in limit per CTA minus static shared memory

static int v2_scratch(cs_graph* g, const CsV2Smem& M, int D, uint32_t grid, size_t* stride) {
    const size_t a = align_up((size_t)M.rcap * 4, 256);
    const size_t b = align_up((size_t)M.rcap, 32) * 8;
    const size_t c = (size_t)M.rcap * 2 * D * 8;
    *stride = align_up(a + b + c, 256);
    const size_t need = *stride * grid;
    if (need > g->scratch2_bytes) {
        if (g->d_scratch2) cudaFree(g->d_scratch2);
        g->d_scratch2 = nullptr;
        g->scratch2_bytes = 0;
        CS_CUDA(cudaMalloc(&g->d_scratch2, need));
        g->scratch2_bytes = need;
    }
    return 0;
}

// Decide whether the shared-memory kernel serves this call and with which capacities: probe a strided sample of the
// staged sources at the largest layout (one CTA per SM), then size pages / reached-node capacity with a little headroom
// so that two or more CTAs share an SM when the reach allows.  Sources that overflow the primary layout are re-run at
// the largest layout, and only what overflows that goes to the global-arena kernel.  Cached per (max_seconds, speed, D).
static int v2_plan(cs_graph* g, CsShortest2Params base, uint64_t m, uint32_t* launches) {
    auto& P = g->plan;
    if (P.valid && P.max_seconds == base.max_seconds && P.speed == base.speed && P.D == base.D && P.pb == g->opt_pb &&
        P.n_sources == m)
        return 0;
    P.valid = true;
    P.use = false;
    P.max_seconds = base.max_seconds;
    P.speed = base.speed;
    P.D = base.D;
    P.pb = g->opt_pb;
    P.n_sources = m;
    if (!g->v2_ok || base.D > 8 || m == 0) return 0;
    CsV2Smem M;
    // largest layout that fits one CTA per SM
    uint32_t rcap = 16384, pages = (uint32_t)(24576u >> g->opt_pb);
    while (!(v2_layout(M, rcap, pages, g->opt_pb, base.D, 256) && M.total <= CS2_SMEM_MAX)) {
        rcap -= 512;
        pages = pages * 31 / 32;
        if (rcap < 1024) return 0;
    }
    P.sm_big = M;
    const uint32_t count = (uint32_t)std::min<uint64_t>(m, 444);
    uint32_t* d_s = g->d_fb_sources;  // scratch for the sample
    float* d_w = g->d_fb_wt;
    cs_k_stride_sample<<<(count + 255) / 256, 256, 0, g->stream>>>(g->d_sources2, g->d_src_wt2, m, count, d_s, d_w);
    CS_CUDA(cudaMemsetAsync(g->d_probe, 0, 2 * sizeof(uint32_t), g->stream));
    CS_CUDA(cudaMemsetAsync(g->d_counters, 0, CS_NCOUNTERS * sizeof(unsigned long long), g->stream));
    CsShortest2Params p = base;
    p.sm = M;
    p.probe = 1;
    p.sources = d_s;
    p.src_wt = d_w;
    p.n_sources = count;
    p.probe_max = g->d_probe;
    p.bin_scale = (float)M.NB / ((base.max_seconds + 1.0f) * (base.max_seconds + 1.0f));
    const uint32_t grid = std::min<uint32_t>(count, (uint32_t)g->sm_count);
    size_t stride = 0;
    if (v2_scratch(g, M, base.D, grid, &stride)) return 1;
    p.scratch = g->d_scratch2;
    p.scratch_stride = stride;
    CS_CUDA(v2_launch(p, 256, grid, g->stream, nullptr));
    *launches += 2;
    uint32_t h[2] = {0, 0};
    CS_CUDA(cudaMemcpyAsync(h, g->d_probe, sizeof(h), cudaMemcpyDeviceToHost, g->stream));
    CS_CUDA(cudaStreamSynchronize(g->stream));
    CS_CUDA(cudaMemsetAsync(g->d_counters, 0, CS_NCOUNTERS * sizeof(unsigned long long), g->stream));
    P.probe_R = h[0];
    P.probe_pages = h[1];
    P.sm = M;
    P.T = 512;
    if (h[0] != 0xffffffffu && h[1] != 0xffffffffu) {
        // (a sample that overflows even the largest layout keeps that layout; its sources fall back individually)
        // Preference: several CTAs per SM (small reach: 128 threads; else 256 threads, shaving the capacity headroom
        // if that is what it takes to fit two CTAs), otherwise one 512-thread CTA per SM.
        const float heads[3] = {g->opt_headroom, 1.0f + (g->opt_headroom - 1.0f) * 0.6f, 1.0f + (g->opt_headroom - 1.0f) * 0.3f};
        bool done = false;
        for (int hi = 0; hi < 3 && !done; ++hi) {
            const uint32_t want_r = std::min<uint32_t>(M.rcap, (uint32_t)(h[0] * heads[hi]) + 32);
            const uint32_t want_p = (uint32_t)(h[1] * heads[hi]) + 4;
            const int T = g->opt_threads ? g->opt_threads : (h[0] <= 1536 ? 128 : 256);
            CsV2Smem M2;
            if (!(v2_layout(M2, want_r, want_p, g->opt_pb, base.D, T) && M2.total <= M.total)) continue;
            CsShortest2Params pt = base;
            pt.sm = M2;
            int occ2 = 0;
            CS_CUDA(v2_launch(pt, T, 0, g->stream, &occ2));
            if (occ2 >= 2 || g->opt_threads) {
                P.sm = M2;
                P.T = T;
                done = true;
            }
        }
        if (!done) {
            const uint32_t want_r = std::min<uint32_t>(M.rcap, (uint32_t)(h[0] * g->opt_headroom) + 32);
            const uint32_t want_p = (uint32_t)(h[1] * g->opt_headroom) + 4;
            CsV2Smem M2;
            if (v2_layout(M2, want_r, want_p, g->opt_pb, base.D, 512) && M2.total <= CS2_SMEM_MAX) P.sm = M2;
            else v2_layout(P.sm, M.rcap, M.TB * 85 / 100, g->opt_pb, base.D, 512);
        }
    } else {
        v2_layout(P.sm, M.rcap, M.TB * 85 / 100, g->opt_pb, base.D, 512);
    }
    if (g->opt_reach_limit) {
        CsV2Smem M3;
        if (v2_layout(M3, std::min(P.sm.rcap, g->opt_reach_limit), P.sm.TB, g->opt_pb, base.D, P.T) && M3.total <= CS2_SMEM_MAX) P.sm = M3;
    }
    if (g->opt_reach_limit2) {
        CsV2Smem M4;
        if (v2_layout(M4, std::min(P.sm_big.rcap, g->opt_reach_limit2), P.sm_big.TB, g->opt_pb, base.D, 256) && M4.total <= CS2_SMEM_MAX)
            P.sm_big = M4;
    }
    p.sm = P.sm;
    int occ = 0;
    CS_CUDA(v2_launch(p, P.T, 0, g->stream, &occ));
    if (occ < 1) return 0;
    P.ctas_per_sm = occ;
    P.use = true;
    return 0;
}
