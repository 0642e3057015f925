// ------------------------------------------------------------------------------------------------ tree dumps
// dijkstra_tree_segment / dijkstra_tree_simplest: one source, the reference's sequential relaxation rules replayed by
// cs_k_tree_segment / cs_k_tree_angular (cs_tree.cuh); per-node state downloaded.
namespace {
struct TreeBuffers {
    float *agg = nullptr, *simpl = nullptr, *st_metric = nullptr, *st_simpl = nullptr, *st_agg = nullptr;
    uint32_t *pred = nullptr, *origin = nullptr, *last = nullptr, *order = nullptr, *eorder = nullptr, *counts = nullptr;
    uint8_t *flags = nullptr, *st_flags = nullptr, *reached = nullptr;
    uint2* heap = nullptr;
    ~TreeBuffers() {
        for (void* p : {(void*)agg, (void*)simpl, (void*)st_metric, (void*)st_simpl, (void*)st_agg, (void*)pred, (void*)origin,
                        (void*)last, (void*)order, (void*)eorder, (void*)counts, (void*)flags, (void*)st_flags,
                        (void*)reached, (void*)heap})
            if (p) cudaFree(p);
    }
};
template <class T>
int tree_alloc_fill(cs_graph* g, T** p, size_t count, uint32_t fill_word, bool bytes) {
    const size_t sz = std::max<size_t>(count, 1) * sizeof(T);
    CS_CUDA(cudaMalloc(p, sz));
    if (bytes) {
        CS_CUDA(cudaMemsetAsync(*p, (int)(fill_word & 0xff), sz, g->stream));
    } else {
        cs_k_fill_u32<<<(unsigned)((count + 255) / 256 + 1), 256, 0, g->stream>>>(reinterpret_cast<uint32_t*>(*p), count, fill_word);
        CS_CUDA(cudaGetLastError());
    }
    return 0;
}
int tree_finish(cs_graph* g, const char* what) {
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(g->stream) != cudaSuccess)
        return cs_fail("CUDA error in %s", what);
    int herr = 0;
    CS_CUDA(cudaMemcpy(&herr, g->d_error, sizeof(int), cudaMemcpyDeviceToHost));
    if (herr) return cs_fail("%s: heap capacity exceeded", what);
    return 0;
}
}  // namespace

extern "C" int cs_dijkstra_tree_segment(cs_graph* g, uint32_t src, uint32_t max_seconds, float speed_m_s,
                                        uint32_t* n_visited, uint32_t* visited_nodes, uint64_t* n_visited_edges,
                                        uint32_t* visited_edges, int64_t* pred, float* agg_seconds, int64_t* origin_seg,
                                        int64_t* last_seg, uint8_t* flags) {
    if (!g) return cs_fail("null graph");
    if (!n_visited || !visited_nodes || !n_visited_edges || !visited_edges || !pred || !agg_seconds || !origin_seg ||
        !last_seg || !flags)
        return cs_fail("null output");
    if (src >= g->n) return cs_fail("src_idx %u out of range for network with node_bound %u", src, g->n);
    if (!(speed_m_s > 0.f) || !std::isfinite(speed_m_s)) return cs_fail("speed_m_s must be finite and positive, got %f", speed_m_s);
    CS_CUDA(cudaSetDevice(g->device));
    uint32_t launches = 0;
    if (prep_seconds(g, speed_m_s, false, &launches)) return 1;
    CS_CUDA(cudaMemsetAsync(g->d_error, 0, sizeof(int), g->stream));
    const size_t n = g->n, E = g->E;
    TreeBuffers b;
    if (tree_alloc_fill(g, &b.agg, n, CS_INF_BITS, false) || tree_alloc_fill(g, &b.pred, n, CS_NOSLOT, false) ||
        tree_alloc_fill(g, &b.origin, n, CS_NOSLOT, false) || tree_alloc_fill(g, &b.last, n, CS_NOSLOT, false) ||
        tree_alloc_fill(g, &b.flags, n, 0, true) || tree_alloc_fill(g, &b.order, n, 0, true) ||
        tree_alloc_fill(g, &b.eorder, E, 0, true) || tree_alloc_fill(g, &b.counts, 2, 0, true))
        return 1;
    CS_CUDA(cudaMalloc(&b.heap, (E + 2) * sizeof(uint2)));
    CsTreeParams p{};
    p.n = g->n;
    p.off = g->d_in_off;
    p.rec = g->d_in_rec;
    p.src = src;
    p.max_seconds = (float)max_seconds;
    p.agg = b.agg;
    p.pred = b.pred;
    p.origin = b.origin;
    p.last = b.last;
    p.flags = b.flags;
    p.order = b.order;
    p.eorder = b.eorder;
    p.counts = b.counts;
    p.heap = b.heap;
    p.heap_cap = (uint32_t)std::min<size_t>(E + 2, 0xffffffffu);
    p.error = g->d_error;
    cs_k_tree_segment<<<1, 32, 0, g->stream>>>(p);
    if (tree_finish(g, "dijkstra_tree_segment")) return 1;
    uint32_t counts[2] = {0, 0};
    CS_CUDA(cudaMemcpy(counts, b.counts, 8, cudaMemcpyDeviceToHost));
    *n_visited = counts[0];
    *n_visited_edges = counts[1];
    std::vector<uint32_t> hp(n), ho(n), hl(n), he(counts[1]);
    CS_CUDA(cudaMemcpy(visited_nodes, b.order, (size_t)counts[0] * 4, cudaMemcpyDeviceToHost));
    if (counts[1]) CS_CUDA(cudaMemcpy(he.data(), b.eorder, (size_t)counts[1] * 4, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(hp.data(), b.pred, n * 4, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(ho.data(), b.origin, n * 4, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(hl.data(), b.last, n * 4, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(agg_seconds, b.agg, n * 4, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(flags, b.flags, n, cudaMemcpyDeviceToHost));
    // in-CSR positions -> the container's edge ids (petgraph EdgeIndex)
    for (uint32_t i = 0; i < counts[1]; ++i) visited_edges[i] = g->in_edge_at[he[i]];
    for (size_t i = 0; i < n; ++i) {
        pred[i] = hp[i] == CS_NOSLOT ? -1 : (int64_t)hp[i];
        origin_seg[i] = ho[i] == CS_NOSLOT ? -1 : (int64_t)g->in_edge_at[ho[i]];
        last_seg[i] = hl[i] == CS_NOSLOT ? -1 : (int64_t)g->in_edge_at[hl[i]];
    }
    return 0;
}

extern "C" int cs_dijkstra_tree_simplest(cs_graph* g, uint32_t src, uint32_t max_seconds, float speed_m_s,
                                         uint32_t* n_visited, uint32_t* visited_nodes, int64_t* pred, float* simpl_dist,
                                         float* agg_seconds, uint8_t* flags) {
    if (!g) return cs_fail("null graph");
    if (!n_visited || !visited_nodes || !pred || !simpl_dist || !agg_seconds || !flags) return cs_fail("null output");
    if (!g->is_dual)
        return cs_fail("dijkstra_tree_simplest requires a dual graph for angular analysis. Convert the graph with "
                       "cityseer.tools.graphs.nx_to_dual(...) before ingesting it into NetworkStructure.");
    if (g->dual_status == 1) return cs_fail("dual edge is missing shared_primal_node_key metadata");
    if (g->dual_status == 2) return cs_fail("dual node references more than two primal endpoints");
    if (src >= g->n) return cs_fail("src_idx %u out of range for network with node_bound %u", src, g->n);
    if (!(speed_m_s > 0.f) || !std::isfinite(speed_m_s)) return cs_fail("speed_m_s must be finite and positive, got %f", speed_m_s);
    CS_CUDA(cudaSetDevice(g->device));
    uint32_t launches = 0;
    if (prep_seconds(g, speed_m_s, true, &launches)) return 1;
    CS_CUDA(cudaMemsetAsync(g->d_error, 0, sizeof(int), g->stream));
    const size_t n = g->n, E = g->E;
    TreeBuffers b;
    if (tree_alloc_fill(g, &b.agg, n, CS_INF_BITS, false) || tree_alloc_fill(g, &b.simpl, n, CS_INF_BITS, false) ||
        tree_alloc_fill(g, &b.pred, n, CS_NOSLOT, false) || tree_alloc_fill(g, &b.flags, n, 0, true) ||
        tree_alloc_fill(g, &b.reached, n, 0, true) || tree_alloc_fill(g, &b.order, n, 0, true) ||
        tree_alloc_fill(g, &b.st_metric, 2 * n, CS_INF_BITS, false) || tree_alloc_fill(g, &b.st_simpl, 2 * n, CS_INF_BITS, false) ||
        tree_alloc_fill(g, &b.st_agg, 2 * n, CS_INF_BITS, false) || tree_alloc_fill(g, &b.st_flags, 2 * n, 0, true) ||
        tree_alloc_fill(g, &b.counts, 2, 0, true))
        return 1;
    CS_CUDA(cudaMalloc(&b.heap, (E + 2) * sizeof(uint2)));
    CsTreeParams p{};
    p.n = g->n;
    p.off = g->d_out_off;
    p.rec = g->d_ang_rec;
    p.src = src;
    p.max_seconds = (float)max_seconds;
    p.agg = b.agg;
    p.simpl = b.simpl;
    p.pred = b.pred;
    p.flags = b.flags;
    p.st_metric = b.st_metric;
    p.st_simpl = b.st_simpl;
    p.st_agg = b.st_agg;
    p.st_flags = b.st_flags;
    p.reached = b.reached;
    p.order = b.order;
    p.counts = b.counts;
    p.heap = b.heap;
    p.heap_cap = (uint32_t)std::min<size_t>(E + 2, 0xffffffffu);
    p.error = g->d_error;
    cs_k_tree_angular<<<1, 32, 0, g->stream>>>(p);
    if (tree_finish(g, "dijkstra_tree_simplest")) return 1;
    uint32_t counts[2] = {0, 0};
    CS_CUDA(cudaMemcpy(counts, b.counts, 8, cudaMemcpyDeviceToHost));
    *n_visited = counts[0];
    std::vector<uint32_t> hp(n);
    CS_CUDA(cudaMemcpy(visited_nodes, b.order, (size_t)counts[0] * 4, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(hp.data(), b.pred, n * 4, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(simpl_dist, b.simpl, n * 4, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(agg_seconds, b.agg, n * 4, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(flags, b.flags, n, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) pred[i] = hp[i] == CS_NOSLOT ? -1 : (int64_t)hp[i];
    return 0;
}
