// Host side of the chain-contracted segment kernel (cs_segment3.cuh).  Included by cs_api.cu.

template <int DT>
static cudaError_t seg3_launch_t(const CsSegment3Params& t, uint32_t workers, cudaStream_t st) {
    constexpr uint32_t smem = cs3s_smem_bytes<DT>();
    cudaError_t e = cudaFuncSetAttribute(cs_k_segment3<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    constexpr uint32_t W = cs3s_warps<DT>();
    static_assert(smem <= 227 * 1024, "the CTA must fit the SM's shared memory");
    const uint32_t grid = (uint32_t)std::min<uint64_t>(workers / W, (t.n_sources + W - 1) / W);
    if (grid == 0) return cudaSuccess;
    cs_k_segment3<DT><<<grid, W * 32, smem, st>>>(t);
    return cudaGetLastError();
}
static cudaError_t seg3_launch(const CsSegment3Params& t, uint32_t workers, cudaStream_t st) {
    switch (cs_shortest_dt(t.D)) {
        case 1: return seg3_launch_t<1>(t, workers, st);
        case 2: return seg3_launch_t<2>(t, workers, st);
        case 3: return seg3_launch_t<3>(t, workers, st);
        case 4: return seg3_launch_t<4>(t, workers, st);
        case 8: return seg3_launch_t<8>(t, workers, st);
        default: return seg3_launch_t<CS_MAX_THRESHOLDS>(t, workers, st);
    }
}

// A node-level arena of a few warps per SM for the heap-order replay of the sources the chain kernel set aside
// (exactly tied tree parents): kept beside the chain kernel's arena so that neither is re-allocated per call.
#define CS_REPLAY_SPC 4u
static int ensure_replay_arena(cs_graph* g, int D) {
    if (g->d_arena2 && g->arena2_D >= D) return 0;
    if (g->d_arena2) {
        CS_CUDA(cudaFree(g->d_arena2));
        g->d_arena2 = nullptr;
    }
    const size_t nstates = g->n;
    uint32_t rcap = g->cfg_rcap ? g->cfg_rcap : CS_DEFAULT_RCAP;
    rcap = (uint32_t)std::min<size_t>(std::max<size_t>(rcap * 8, 1u << 16), nstates);  // node-level reach, not junctions
    rcap = std::max(rcap, 32u);
    const uint32_t qcap = rcap * 2 + 64;
    CsArenaLayout L{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    L.ds = take(nstates * sizeof(uint2));
    L.node_list = take((size_t)rcap * 4);
    L.qa = take((size_t)qcap * 8);
    L.qb = take((size_t)qcap * 8);
    L.far = take((size_t)qcap * 8);
    L.s_node = take((size_t)rcap * 4);
    L.s_agg = take((size_t)rcap * 4);
    L.predmask = take((size_t)rcap * 8);
    L.sigma = take((size_t)rcap * 8);
    L.dep = take((size_t)rcap * 2 * D * 8);
    L.bdone = take((size_t)rcap * 8);
    L.erank = take((size_t)rcap * 16);
    L.stride = align_up(off, 4096);
    L.rcap = rcap;
    L.qcap = qcap;
    uint32_t workers = (uint32_t)g->sm_count * CS_REPLAY_SPC;
    size_t free_b = 0, total_b = 0;
    CS_CUDA(cudaMemGetInfo(&free_b, &total_b));
    while (workers > CS_REPLAY_SPC && (size_t)workers * L.stride > (size_t)((double)free_b * 0.5)) workers -= CS_REPLAY_SPC;
    if ((size_t)workers * L.stride > (size_t)((double)free_b * 0.5))
        return cs_fail("not enough device memory for the replay arena (%zu bytes per worker)", L.stride);
    CS_CUDA(cudaMalloc(&g->d_arena2, (size_t)workers * L.stride));
    dim3 grid((unsigned)std::min<size_t>((nstates + 255) / 256, 64), workers);
    cs_k_init_ds<<<grid, 256, 0, g->stream>>>(g->d_arena2, L.stride, L.ds, nstates);
    CS_CUDA(cudaGetLastError());
    CS_CUDA(cudaStreamSynchronize(g->stream));
    g->lay2 = L;
    g->workers2 = workers;
    g->arena2_D = D;
    return 0;
}

// segment_centrality on the chain-contracted copy of the graph; sources with exactly tied tree parents are replayed by the
// node-level kernel in heap order.  Returns -1 when the call must be served by the node-level kernel altogether.
static int run_segment3(cs_graph* g, int D, const uint32_t* distances, const float* betas, const uint32_t* seconds,
                        float speed_m_s, int compute_closeness, int compute_betweenness, uint64_t n_sources,
                        const uint32_t* sources, double* out, int out_on_device, int accumulate, cs_stats* stats) {
    CS_CUDA(cudaSetDevice(g->device));
    g->last_kernel = 3;
    if (ensure_arena(g, 3, D)) return 1;
    uint32_t launches = 0;
    CS_CUDA(cudaEventRecord(g->ev[0], g->stream));
    if (stage_sources(g, n_sources, sources, nullptr, nullptr)) return 1;
    if (prep_seconds(g, speed_m_s, false, &launches)) return 1;  // the replay reads the node-level records
    if (g->cached_speed3 != speed_m_s && g->v3_ncsec) {
        cs_k_prep_csec<<<(int)((g->v3_ncsec + 255) / 256), 256, 0, g->stream>>>(g->d3_csec, g->d3_cnum, g->v3_ncsec, speed_m_s);
        launches += 1;
        g->cached_speed3 = speed_m_s;
    }
    CS_CUDA(cudaMemsetAsync(g->d_counters, 0, CS_NCOUNTERS * sizeof(unsigned long long), g->stream));
    CS_CUDA(cudaMemsetAsync(g->d_error, 0, sizeof(int), g->stream));
    const size_t elems = (size_t)4 * D * g->n;
    double* d_out = nullptr;
    if (acquire_out(g, out, out_on_device, accumulate, elems, &d_out)) return 1;
    const size_t acc_elems = (size_t)g->n * D;
    if (g->acc_cap < acc_elems) {
        if (g->d_acc) cudaFree(g->d_acc);
        g->d_acc = nullptr;
        g->acc_cap = 0;
        CS_CUDA(cudaMalloc(&g->d_acc, acc_elems * sizeof(double)));
        g->acc_cap = acc_elems;
    }
    if (compute_betweenness) CS_CUDA(cudaMemsetAsync(g->d_acc, 0, acc_elems * sizeof(double), g->stream));
    if (g->redo_cap < n_sources) {
        if (g->d_redo) cudaFree(g->d_redo);
        g->d_redo = nullptr;
        g->redo_cap = 0;
        CS_CUDA(cudaMalloc(&g->d_redo, std::max<uint64_t>(n_sources, 1) * 4));
        g->redo_cap = n_sources;
    }
    uint32_t max_sec = 0;
    CsSegment3Params t{};
    t.g.J = g->v3_J;
    t.g.I = g->v3_I;
    t.g.n = g->n;
    t.g.jinfo = g->d3_jinfo;
    t.g.links = g->d3_links;
    t.g.csec = g->d3_csec;
    t.g.ctab = g->d3_ctab;
    t.g.int_chain = g->d3_int_chain;
    t.g.orig_of_new = g->d3_orig_of_new;
    t.g.new_of_orig = g->d3_new_of_orig;
    t.g.weight = g->d3_weight;
    t.clen = g->d3_clen;
    t.cimp = g->d3_cimp;
    t.D = D;
    t.closeness = compute_closeness;
    t.betweenness = compute_betweenness;
    for (int i = 0; i < D; ++i) {
        t.dist_f[i] = (float)distances[i];
        t.beta_f[i] = betas[i];
        max_sec = std::max(max_sec, seconds[i]);
    }
    t.max_seconds = (float)max_sec;
    t.speed = speed_m_s;
    t.sources = g->d_sources;
    t.n_sources = n_sources;
    t.out = d_out;
    t.acc_b = g->d_acc;
    t.counters = g->d_counters;
    t.error = g->d_error;
    t.arena = g->d_arena;
    t.lay = g->lay;
    t.delta = default_delta(g, speed_m_s);
    t.bin_scale = (float)CS3_NBINS / (((float)max_sec + 1.0f) * ((float)max_sec + 1.0f));
    t.redo_list = g->d_redo;
    CS_CUDA(cudaEventRecord(g->ev[1], g->stream));
    CS_CUDA(seg3_launch(t, g->workers, g->stream));
    launches += 1;
    if (compute_betweenness) {
        cs_k_epilogue_segment3<<<(g->n + 255) / 256, 256, 0, g->stream>>>(g->d_acc, d_out, g->d3_orig_of_new, g->n, D);
        launches += 1;
        CS_CUDA(cudaGetLastError());
    }
    unsigned long long n_redo = 0;
    int herr = 0;
    CS_CUDA(cudaMemcpyAsync(&n_redo, g->d_counters + CS_C_FALLBACK, sizeof(n_redo), cudaMemcpyDeviceToHost, g->stream));
    CS_CUDA(cudaMemcpyAsync(&herr, g->d_error, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    CS_CUDA(cudaStreamSynchronize(g->stream));
    if (n_redo && !herr) {
        // exactly tied tree parents (regular grids, equal pieces): the reference decides them by BinaryHeap pop order
        // (centrality.rs:1589); the node-level kernel replays its heap for these sources, adding into the same result
        if (ensure_replay_arena(g, D)) return 1;
        if (segment_smem_optin(g)) return 1;
        CS_CUDA(cudaMemsetAsync(g->d_counters + CS_C_NEXT, 0, sizeof(unsigned long long), g->stream));
        CsSegmentParams p{};
        p.g = graph_dev(g);
        p.D = D;
        p.closeness = compute_closeness;
        p.betweenness = compute_betweenness;
        for (int i = 0; i < D; ++i) {
            p.dist_f[i] = t.dist_f[i];
            p.beta_f[i] = t.beta_f[i];
        }
        p.max_seconds = t.max_seconds;
        p.speed = speed_m_s;
        p.sources = g->d_redo;
        p.n_sources = n_redo;
        p.out = d_out;
        p.counters = g->d_counters;
        p.error = g->d_error;
        p.arena = g->d_arena2;
        p.lay = g->lay2;
        p.delta = default_delta(g, speed_m_s);
        p.bin_scale = (float)CS_NBINS / (((float)max_sec + 1.0f) * ((float)max_sec + 1.0f));
        p.replay = 1;
        p.redo_list = nullptr;
        p.compact_arena = 1;
        const uint32_t ctas = g->workers2 / CS_REPLAY_SPC;
        p.src_per_cta = (uint32_t)std::min<uint64_t>(CS_REPLAY_SPC, (n_redo + ctas - 1) / ctas);
        const uint32_t grid = (uint32_t)std::min<uint64_t>(ctas, (n_redo + p.src_per_cta - 1) / p.src_per_cta);
        const int threads = CS_SEG_WARPS * 32;
        const size_t sm = CS_SEG_SMEM_BYTES;
        if (D == 1) cs_k_segment<1><<<grid, threads, sm, g->stream>>>(p);
        else if (D == 2) cs_k_segment<2><<<grid, threads, sm, g->stream>>>(p);
        else if (D == 3) cs_k_segment<3><<<grid, threads, sm, g->stream>>>(p);
        else if (D == 4) cs_k_segment<4><<<grid, threads, sm, g->stream>>>(p);
        else if (D <= 8) cs_k_segment<8><<<grid, threads, sm, g->stream>>>(p);
        else cs_k_segment<CS_MAX_THRESHOLDS><<<grid, threads, sm, g->stream>>>(p);
        launches += 1;
        CS_CUDA(cudaGetLastError());
    }
    CS_CUDA(cudaEventRecord(g->ev[2], g->stream));
    return finish_call(g, out, out_on_device, elems, d_out, stats, launches);
}
