// the segment kernel's per-warp bins exceed the 48 KB static limit: opt every instantiation in once per graph (= per device)
static int segment_smem_optin(cs_graph* g) {
    if (g->seg_optin) return 0;
    const int b = (int)CS_SEG_SMEM_BYTES;
    CS_CUDA(cudaFuncSetAttribute(cs_k_segment<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CS_CUDA(cudaFuncSetAttribute(cs_k_segment<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CS_CUDA(cudaFuncSetAttribute(cs_k_segment<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CS_CUDA(cudaFuncSetAttribute(cs_k_segment<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CS_CUDA(cudaFuncSetAttribute(cs_k_segment<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CS_CUDA(cudaFuncSetAttribute(cs_k_segment<CS_MAX_THRESHOLDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    g->seg_optin = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ single-source tree
// dijkstra_tree_shortest (centrality.rs:1141-1200, :1499-1508): the capped search of the segment kernel (distances by
// the label-correcting search, exact settle order, single predecessor = the first strict improvement in pop order) run
// for one source with the accumulation phases off, its per-node state downloaded.
extern "C" int cs_dijkstra_tree_shortest(cs_graph* g, uint32_t src, uint32_t max_seconds, float speed_m_s,
                                         uint32_t* n_visited, uint32_t* visited_order, int64_t* pred, float* agg_seconds) {
    if (!g) return cs_fail("null graph");
    if (!n_visited || !visited_order || !pred || !agg_seconds) return cs_fail("null output");
    if (src >= g->n) return cs_fail("src_idx %u out of range for network with node_bound %u", src, g->n);
    if (!(speed_m_s > 0.f) || !std::isfinite(speed_m_s)) return cs_fail("speed_m_s must be finite and positive, got %f", speed_m_s);
    CS_CUDA(cudaSetDevice(g->device));
    g->last_kernel = 0;
    if (ensure_arena(g, 1, 1)) return 1;
    uint32_t launches = 0;
    if (stage_sources(g, 1, &src, nullptr, nullptr)) return 1;
    if (prep_seconds(g, speed_m_s, false, &launches)) return 1;
    CS_CUDA(cudaMemsetAsync(g->d_counters, 0, CS_NCOUNTERS * sizeof(unsigned long long), g->stream));
    CS_CUDA(cudaMemsetAsync(g->d_error, 0, sizeof(int), g->stream));
    const size_t n = g->n;
    uint32_t *d_order = nullptr, *d_pred = nullptr, *d_count = nullptr;
    float* d_agg = nullptr;
    CS_CUDA(cudaMalloc(&d_order, n * 4));
    CS_CUDA(cudaMalloc(&d_pred, n * 4));
    CS_CUDA(cudaMalloc(&d_agg, n * 4));
    CS_CUDA(cudaMalloc(&d_count, 4));
    CS_CUDA(cudaMemsetAsync(d_pred, 0xff, n * 4, g->stream));
    CS_CUDA(cudaMemsetAsync(d_count, 0, 4, g->stream));
    {
        std::vector<float> inf(n, INFINITY);
        CS_CUDA(cudaMemcpyAsync(d_agg, inf.data(), n * 4, cudaMemcpyHostToDevice, g->stream));
        CS_CUDA(cudaStreamSynchronize(g->stream));
    }
    CsSegmentParams p{};
    p.g = graph_dev(g);
    p.D = 1;
    p.closeness = 0;
    p.betweenness = 0;
    p.dist_f[0] = 0.f;
    p.beta_f[0] = 0.f;
    p.max_seconds = (float)max_seconds;
    p.speed = speed_m_s;
    p.sources = g->d_sources;
    p.n_sources = 1;
    p.out = nullptr;
    p.counters = g->d_counters;
    p.error = g->d_error;
    p.arena = g->d_arena;
    p.lay = g->lay;
    p.delta = default_delta(g, speed_m_s);
    p.bin_scale = (float)CS_NBINS / (((float)max_seconds + 1.0f) * ((float)max_seconds + 1.0f));
    p.dump_order = d_order;
    p.dump_pred = d_pred;
    p.dump_agg = d_agg;
    p.dump_count = d_count;
    p.replay = 1;  // one source: always settle equal keys in the reference's heap order
    p.src_per_cta = 1;
    if (segment_smem_optin(g)) return 1;
    cs_k_segment<1><<<1, CS_SEG_WARPS * 32, CS_SEG_SMEM_BYTES, g->stream>>>(p);
    int rc = 0;
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(g->stream) != cudaSuccess) rc = cs_fail("CUDA error in the tree search");
    int herr = 0;
    if (!rc) {
        cudaMemcpy(&herr, g->d_error, sizeof(int), cudaMemcpyDeviceToHost);
        if (herr) {
            g->arena_kind = -1;  // the failing warp left its dense-map entries behind
            rc = cs_fail("search arena overflow: the source reached more than %u nodes; raise reach_capacity via cs_graph_configure", g->lay.rcap);
        }
    }
    if (!rc) {
        std::vector<uint32_t> hp(n);
        cudaMemcpy(n_visited, d_count, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(visited_order, d_order, (size_t)*n_visited * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hp.data(), d_pred, n * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(agg_seconds, d_agg, n * 4, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < n; ++i) pred[i] = hp[i] == 0xffffffffu ? -1 : (int64_t)hp[i];
    }
    cudaFree(d_order);
    cudaFree(d_pred);
    cudaFree(d_agg);
    cudaFree(d_count);
    return rc;
}

// Batched dijkstra_tree_shortest (the searches that feed the reference's data.rs aggregations run
// one source at a time, centrality.rs:1141-1200): many sources per launch, one warp each, every search replayed in the
// reference's heap order so that visit order and predecessors are exact under ties.  Outputs are [n_sources][capacity].
extern "C" int cs_dijkstra_trees_shortest(cs_graph* g, uint64_t n_sources, const uint32_t* sources, uint32_t max_seconds,
                                          float speed_m_s, uint32_t capacity, uint32_t* counts, uint32_t* visited_order,
                                          int64_t* pred, float* agg_seconds) {
    if (!g) return cs_fail("null graph");
    if (!sources || !counts || !visited_order || !pred || !agg_seconds) return cs_fail("null argument");
    if (capacity == 0) return cs_fail("capacity must be positive");
    if (!(speed_m_s > 0.f) || !std::isfinite(speed_m_s)) return cs_fail("speed_m_s must be finite and positive, got %f", speed_m_s);
    for (uint64_t i = 0; i < n_sources; ++i)
        if (sources[i] >= g->n) return cs_fail("src_idx %u out of range for network with node_bound %u", sources[i], g->n);
    if (n_sources == 0) return 0;
    CS_CUDA(cudaSetDevice(g->device));
    for (;;) {
        g->last_kernel = 0;
        if (ensure_arena(g, 1, 1)) return 1;
        uint32_t launches = 0;
        if (stage_sources(g, n_sources, sources, nullptr, nullptr)) return 1;
        if (prep_seconds(g, speed_m_s, false, &launches)) return 1;
        CS_CUDA(cudaMemsetAsync(g->d_counters, 0, CS_NCOUNTERS * sizeof(unsigned long long), g->stream));
        CS_CUDA(cudaMemsetAsync(g->d_error, 0, sizeof(int), g->stream));
        const size_t total = (size_t)n_sources * capacity;
        uint32_t *d_order = nullptr, *d_pred = nullptr, *d_count = nullptr;
        float* d_agg = nullptr;
        CS_CUDA(cudaMalloc(&d_order, total * 4));
        CS_CUDA(cudaMalloc(&d_pred, total * 4));
        CS_CUDA(cudaMalloc(&d_agg, total * 4));
        CS_CUDA(cudaMalloc(&d_count, n_sources * 4));
        CS_CUDA(cudaMemsetAsync(d_count, 0, n_sources * 4, g->stream));
        CsSegmentParams p{};
        p.g = graph_dev(g);
        p.D = 1;
        p.max_seconds = (float)max_seconds;
        p.speed = speed_m_s;
        p.sources = g->d_sources;
        p.n_sources = n_sources;
        p.counters = g->d_counters;
        p.error = g->d_error;
        p.arena = g->d_arena;
        p.lay = g->lay;
        p.delta = default_delta(g, speed_m_s);
        p.bin_scale = (float)CS_NBINS / (((float)max_seconds + 1.0f) * ((float)max_seconds + 1.0f));
        p.replay = 1;
        p.batch_cap = capacity;
        p.b_order = d_order;
        p.b_pred = d_pred;
        p.b_agg = d_agg;
        p.b_count = d_count;
        int rc = 0;
        if (segment_smem_optin(g)) rc = 1;
        if (!rc) {
            const uint32_t ctas = g->workers / CS_SEG_WARPS;
            p.src_per_cta = (uint32_t)std::min<uint64_t>(CS_SEG_WARPS, (n_sources + ctas - 1) / ctas);
            const uint32_t grid = (uint32_t)std::min<uint64_t>(ctas, (n_sources + p.src_per_cta - 1) / p.src_per_cta);
            cs_k_segment<1><<<grid, CS_SEG_WARPS * 32, CS_SEG_SMEM_BYTES, g->stream>>>(p);
            if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(g->stream) != cudaSuccess)
                rc = cs_fail("CUDA error in the batched tree search");
        }
        int herr = 0;
        if (!rc) {
            cudaMemcpy(&herr, g->d_error, sizeof(int), cudaMemcpyDeviceToHost);
            g->last_herr = herr;
            if (herr) {
                g->arena_kind = -1;
                rc = cs_fail("search arena overflow: a source reached more than %u nodes; raise reach_capacity via cs_graph_configure", g->lay.rcap);
            }
        }
        if (!rc) {
            cudaMemcpy(counts, d_count, n_sources * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(visited_order, d_order, total * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(agg_seconds, d_agg, total * 4, cudaMemcpyDeviceToHost);
            std::vector<uint32_t> hp(total);
            cudaMemcpy(hp.data(), d_pred, total * 4, cudaMemcpyDeviceToHost);
            for (uint64_t s2 = 0; s2 < n_sources; ++s2) {
                const uint32_t c = std::min(counts[s2], capacity);
                for (uint32_t k = 0; k < c; ++k) {
                    const size_t at = (size_t)s2 * capacity + k;
                    pred[at] = hp[at] == 0xffffffffu ? -1 : (int64_t)hp[at];
                }
            }
        }
        cudaFree(d_order);
        cudaFree(d_pred);
        cudaFree(d_agg);
        cudaFree(d_count);
        if (!rc) {
            for (uint64_t s2 = 0; s2 < n_sources; ++s2)
                if (counts[s2] > capacity)
                    return cs_fail("source %u settled %u nodes, more than the output capacity %u", sources[s2], counts[s2], capacity);
            return 0;
        }
        if (!grow_after_overflow(g, g->n, g->lay.rcap)) return rc;
    }
}

// ------------------------------------------------------------------------------------------------ segment
static int run_segment3(cs_graph* g, int D, const uint32_t* distances, const float* betas, const uint32_t* seconds,
                        float speed_m_s, int compute_closeness, int compute_betweenness, uint64_t n_sources,
                        const uint32_t* sources, double* out, int out_on_device, int accumulate, cs_stats* stats);

static int run_segment(cs_graph* g, int D, const uint32_t* distances, const float* betas, const uint32_t* seconds,
                       float speed_m_s, int compute_closeness, int compute_betweenness, uint64_t n_sources,
                       const uint32_t* sources, double* out, int out_on_device, int accumulate, cs_stats* stats) {
    if (!g) return cs_fail("null graph");
    if (!out) return cs_fail("null output");
    if (check_thresholds(D, seconds)) return 1;
    if (!compute_closeness && !compute_betweenness)
        return cs_fail("Either or both closeness and betweenness flags is required, but both parameters are False.");
    if (!(speed_m_s > 0.f) || !std::isfinite(speed_m_s)) return cs_fail("speed_m_s must be finite and positive, got %f", speed_m_s);
    if (g->twin_missing)
        return cs_fail("Edge not found: segment_centrality needs the reverse twin of every directed edge (graph.rs:1291)");
    CS_CUDA(cudaSetDevice(g->device));
    g->last_kernel = 0;
    if (ensure_arena(g, 1, D)) return 1;
    uint32_t launches = 0;
    CS_CUDA(cudaEventRecord(g->ev[0], g->stream));
    if (stage_sources(g, n_sources, sources, nullptr, nullptr)) return 1;
    if (prep_seconds(g, speed_m_s, false, &launches)) return 1;
    CS_CUDA(cudaMemsetAsync(g->d_counters, 0, CS_NCOUNTERS * sizeof(unsigned long long), g->stream));
    CS_CUDA(cudaMemsetAsync(g->d_error, 0, sizeof(int), g->stream));
    const size_t elems = (size_t)4 * D * g->n;
    double* d_out = nullptr;
    if (acquire_out(g, out, out_on_device, accumulate, elems, &d_out)) return 1;
    CsSegmentParams p{};
    p.g = graph_dev(g);
    p.D = D;
    p.closeness = compute_closeness;
    p.betweenness = compute_betweenness;
    uint32_t max_sec = 0;
    for (int i = 0; i < D; ++i) {
        p.dist_f[i] = (float)distances[i];
        p.beta_f[i] = betas[i];
        max_sec = std::max(max_sec, seconds[i]);
    }
    p.max_seconds = (float)max_sec;
    p.speed = speed_m_s;
    p.sources = g->d_sources;
    p.n_sources = n_sources;
    p.out = d_out;
    p.counters = g->d_counters;
    p.error = g->d_error;
    p.arena = g->d_arena;
    p.lay = g->lay;
    p.delta = default_delta(g, speed_m_s);
    p.bin_scale = (float)CS_NBINS / (((float)max_sec + 1.0f) * ((float)max_sec + 1.0f));
    if (segment_smem_optin(g)) return 1;
    if (g->redo_cap < n_sources) {
        if (g->d_redo) cudaFree(g->d_redo);
        g->d_redo = nullptr;
        g->redo_cap = 0;
        CS_CUDA(cudaMalloc(&g->d_redo, std::max<uint64_t>(n_sources, 1) * 4));
        g->redo_cap = n_sources;
    }
    p.redo_list = g->d_redo;
    auto launch = [&](uint64_t m) -> int {
        const uint32_t ctas = g->workers / CS_SEG_WARPS;
        if (ctas == 0 || m == 0) return 0;
        // few sources (the replay list): spread them over all CTAs, one warp each; else 32 consecutive sources per CTA
        p.src_per_cta = (uint32_t)std::min<uint64_t>(CS_SEG_WARPS, (m + ctas - 1) / ctas);
        const uint32_t grid = (uint32_t)std::min<uint64_t>(ctas, (m + p.src_per_cta - 1) / p.src_per_cta);
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
        return 0;
    };
    CS_CUDA(cudaEventRecord(g->ev[1], g->stream));
    p.replay = 0;
    if (launch(n_sources)) return 1;
    // sources where tree parents with bit-equal seconds compete (regular grids, equal pieces) were set aside: serve
    // them with the heap-order replay (centrality.rs:1589 resolves those ties by BinaryHeap pop order)
    unsigned long long n_redo = 0;
    CS_CUDA(cudaMemcpyAsync(&n_redo, g->d_counters + CS_C_FALLBACK, sizeof(n_redo), cudaMemcpyDeviceToHost, g->stream));
    CS_CUDA(cudaStreamSynchronize(g->stream));
    if (n_redo) {
        int herr = 0;
        CS_CUDA(cudaMemcpy(&herr, g->d_error, sizeof(int), cudaMemcpyDeviceToHost));
        if (!herr) {
            CS_CUDA(cudaMemsetAsync(g->d_counters + CS_C_NEXT, 0, sizeof(unsigned long long), g->stream));
            p.replay = 1;
            p.sources = g->d_redo;
            p.n_sources = n_redo;
            if (launch(n_redo)) return 1;
        }
    }
    CS_CUDA(cudaEventRecord(g->ev[2], g->stream));
    return finish_call(g, out, out_on_device, elems, d_out, stats, launches);
}

extern "C" int cs_segment_centrality(cs_graph* g, int D, const uint32_t* distances, const float* betas, const uint32_t* seconds,
                                     float speed_m_s, int compute_closeness, int compute_betweenness, uint64_t n_sources,
                                     const uint32_t* sources, double* out, int out_on_device, int accumulate, cs_stats* stats) {
    for (;;) {
        // the chain-contracted kernel serves graphs that are mostly chain interiors (decomposed / OSM-like street networks)
        const bool chain = g && out && g->v3_ok && !g->v3_loops && !g->twin_missing && n_sources > 0 && D >= 1 &&
                           D <= CS_MAX_THRESHOLDS && seconds && (compute_closeness || compute_betweenness) && speed_m_s > 0.f &&
                           std::isfinite(speed_m_s) && (g->opt_kernel == 3 || (g->opt_kernel == 0 && g->v3_I >= g->v3_J));
        const int rc = chain ? run_segment3(g, D, distances, betas, seconds, speed_m_s, compute_closeness,
                                            compute_betweenness, n_sources, sources, out, out_on_device, accumulate, stats)
                             : run_segment(g, D, distances, betas, seconds, speed_m_s, compute_closeness,
                                           compute_betweenness, n_sources, sources, out, out_on_device, accumulate, stats);
        // the segment kernel adds straight into `out`: a failed attempt can only be repeated when it started from zero
        if (!rc || accumulate || !g) return rc;
        if (!grow_after_overflow(g, chain ? (size_t)g->v3_J + 1 : (size_t)g->n, g->lay.rcap)) return rc;
    }
}

// ------------------------------------------------------------------------------------------------ simplest
static int run_simplest(cs_graph* g, int D, const uint32_t* distances, const uint32_t* seconds, float speed_m_s,
                        float tolerance, float angular_scaling_unit, float farness_scaling_offset, int compute_closeness,
                        int compute_betweenness, uint64_t n_sources, const uint32_t* sources, const float* source_wt,
                        const uint8_t* eligible, double* out, int out_on_device, int accumulate, cs_stats* stats) {
    (void)distances;
    if (!g) return cs_fail("null graph");
    if (!out) return cs_fail("null output");
    if (!g->is_dual)
        return cs_fail("centrality_simplest requires a dual graph for angular analysis. Convert the graph with "
                       "cityseer.tools.graphs.nx_to_dual(...) before ingesting it into NetworkStructure.");
    if (g->dual_status == 1) return cs_fail("dual edge is missing shared_primal_node_key metadata");
    if (g->dual_status == 2) return cs_fail("dual node references more than two primal endpoints");
    if (check_thresholds(D, seconds)) return 1;
    if (!compute_closeness && !compute_betweenness)
        return cs_fail("Either or both closeness and betweenness flags is required, but both parameters are False.");
    if (!(speed_m_s > 0.f) || !std::isfinite(speed_m_s)) return cs_fail("speed_m_s must be finite and positive, got %f", speed_m_s);
    if (!(tolerance >= CS_TIE_EPS)) return cs_fail("Tolerance must be >= TIE_EPSILON to avoid float-comparison bugs");
    CS_CUDA(cudaSetDevice(g->device));
    g->last_kernel = 0;
    if (ensure_arena_angular(g, D)) return 1;
    uint32_t launches = 0;
    CS_CUDA(cudaEventRecord(g->ev[0], g->stream));
    if (stage_sources(g, n_sources, sources, source_wt, eligible)) return 1;
    if (prep_seconds(g, speed_m_s, true, &launches)) return 1;
    CS_CUDA(cudaMemsetAsync(g->d_counters, 0, CS_NCOUNTERS * sizeof(unsigned long long), g->stream));
    CS_CUDA(cudaMemsetAsync(g->d_error, 0, sizeof(int), g->stream));
    const size_t elems = (size_t)4 * D * g->n;
    double* d_out = nullptr;
    if (acquire_out(g, out, out_on_device, accumulate, elems, &d_out)) return 1;
    CsSimplestParams p{};
    p.n = g->n;
    p.out_off = g->d_out_off;
    p.ang_rec = g->d_ang_rec;
    p.D = D;
    p.closeness = compute_closeness;
    p.betweenness = compute_betweenness;
    p.phase2 = tolerance > CS_TIE_EPS ? 1 : 0;
    uint32_t max_sec = 0;
    for (int i = 0; i < D; ++i) {
        p.sec_f[i] = (float)seconds[i];
        max_sec = std::max(max_sec, seconds[i]);
    }
    p.max_seconds = (float)max_sec;
    p.tol = tolerance;
    p.unit = angular_scaling_unit;
    p.offset = farness_scaling_offset;
    p.sources = g->d_sources;
    p.src_wt = g->d_src_wt;
    p.n_sources = n_sources;
    p.eligible = g->d_eligible;
    p.out = d_out;
    p.counters = g->d_counters;
    p.error = g->d_error;
    p.arena = g->d_arena;
    p.lay = g->ang_lay;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(g->workers / CS_WARPS_PER_CTA,
                                                       (n_sources + CS_WARPS_PER_CTA - 1) / CS_WARPS_PER_CTA);
    CS_CUDA(cudaEventRecord(g->ev[1], g->stream));
    if (grid > 0) {
        const int threads = CS_WARPS_PER_CTA * 32;
        if (D == 1) cs_k_simplest<1><<<grid, threads, 0, g->stream>>>(p);
        else if (D == 2) cs_k_simplest<2><<<grid, threads, 0, g->stream>>>(p);
        else if (D == 3) cs_k_simplest<3><<<grid, threads, 0, g->stream>>>(p);
        else if (D == 4) cs_k_simplest<4><<<grid, threads, 0, g->stream>>>(p);
        else if (D <= 8) cs_k_simplest<8><<<grid, threads, 0, g->stream>>>(p);
        else cs_k_simplest<CS_MAX_THRESHOLDS><<<grid, threads, 0, g->stream>>>(p);
        launches += 1;
        CS_CUDA(cudaGetLastError());
    }
    CS_CUDA(cudaEventRecord(g->ev[2], g->stream));
    return finish_call(g, out, out_on_device, elems, d_out, stats, launches);
}

extern "C" int cs_centrality_simplest(cs_graph* g, int D, const uint32_t* distances, const uint32_t* seconds, float speed_m_s,
                                      float tolerance, float angular_scaling_unit, float farness_scaling_offset,
                                      int compute_closeness, int compute_betweenness, uint64_t n_sources,
                                      const uint32_t* sources, const float* source_wt, const uint8_t* eligible, double* out,
                                      int out_on_device, int accumulate, cs_stats* stats) {
    for (;;) {
        const int rc = run_simplest(g, D, distances, seconds, speed_m_s, tolerance, angular_scaling_unit,
                                    farness_scaling_offset, compute_closeness, compute_betweenness, n_sources, sources,
                                    source_wt, eligible, out, out_on_device, accumulate, stats);
        if (!rc || accumulate || !g || !grow_after_overflow(g, (size_t)g->n * 2, g->ang_lay.rcap)) return rc;
    }
}
