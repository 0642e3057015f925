extern "C" int cs_centrality_simplest(cs_graph* g, int D, const uint32_t* distances, const uint32_t* seconds, float speed_m_s,
                                      float tolerance, float angular_scaling_unit, float farness_scaling_offset,
                                      int compute_closeness, int compute_betweenness, uint64_t n_sources,
                                      const uint32_t* sources, const float* source_wt, const uint8_t* eligible, double* out,
                                      int out_on_device, int accumulate, cs_stats* stats) {
    (void)g; (void)D; (void)distances; (void)seconds; (void)speed_m_s; (void)tolerance; (void)angular_scaling_unit;
    (void)farness_scaling_offset; (void)compute_closeness; (void)compute_betweenness; (void)n_sources; (void)sources;
    (void)source_wt; (void)eligible; (void)out; (void)out_on_device; (void)accumulate; (void)stats;
    return cs_fail("cs_centrality_simplest: kernel not built yet");
}
extern "C" int cs_segment_centrality(cs_graph* g, int D, const uint32_t* distances, const float* betas, const uint32_t* seconds,
                                     float speed_m_s, int compute_closeness, int compute_betweenness, uint64_t n_sources,
                                     const uint32_t* sources, double* out, int out_on_device, int accumulate, cs_stats* stats) {
    (void)g; (void)D; (void)distances; (void)betas; (void)seconds; (void)speed_m_s; (void)compute_closeness;
    (void)compute_betweenness; (void)n_sources; (void)sources; (void)out; (void)out_on_device; (void)accumulate; (void)stats;
    return cs_fail("cs_segment_centrality: kernel not built yet");
}
