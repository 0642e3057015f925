/* cityseer_b200 — C ABI of the B200-native localized-centrality hot path.
 *
 * This is the drop-in seam beneath cityseer's `NetworkStructure` operator API.  The reference crosses exactly one
 * boundary on this path — Python -> PyO3 method on `rustalgos.graph.NetworkStructure`
 * (/root/reference/rust/src/lib.rs:37-77, stubs in pysrc/cityseer/rustalgos/graph.pyi:85-674) — and every entry
 * point below replaces one of those methods (file:line cited per function).  Plain pointers and sizes only; no
 * Python / torch types.  All functions return 0 on success, non-zero on error; `cs_last_error()` returns a
 * thread-local message.  There is no CPU fallback: every compute call requires a CUDA device (sm_100a build).
 *
 * Graph arrays are indexed by the reference's petgraph `StableGraph` indices (node index / edge index, gaps allowed):
 * that is what `NetworkStructure.node_indices()` / `edge_references()` expose (graph.rs:613, :987).
 */
#ifndef CITYSEER_B200_H
#define CITYSEER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cs_graph cs_graph;

#define CS_MAX_THRESHOLDS 16 /* D, number of distance thresholds per call */
#define CS_MAX_DEGREE 32     /* max in/out degree per node (predecessor sets are 32-bit adjacency masks) */

/* Counters and timings of the last call (all totals over the sources this call processed). */
typedef struct cs_stats {
    uint64_t sources;         /* sources searched */
    uint64_t settled;         /* nodes (states, for simplest) settled: sum over sources of R */
    uint64_t edge_iters;      /* directed edges iterated at settled nodes (the reference's edges_directed loop trips) */
    uint64_t sum_ri;          /* closeness: (source, target, threshold) triples accumulated */
    uint64_t sum_ci;          /* betweenness: positive credits accumulated */
    uint64_t relaxations;     /* device: successful distance decreases (work-efficiency of the label-correcting search) */
    uint64_t reach_totals[CS_MAX_THRESHOLDS]; /* per-threshold reachable-target totals (centrality.rs:1743-1753) */
    float kernel_ms;          /* CUDA-event time of the search/accumulate kernel on the launching stream */
    float total_ms;           /* CUDA-event time incl. uploads of per-call arrays, zeroing and result download */
    uint32_t gpu_launches;    /* kernels launched by this call */
    uint32_t workers;         /* resident workers (warps or CTAs, one source each at a time) */
    uint64_t phase_cycles[8]; /* SM clock cycles summed over workers per kernel phase (search, order, predecessors,
                                 closeness, dependencies, reset); chain-contracted kernel: [6] dependency chunks,
                                 [7] 32-link batches of the dependency pass */
    uint64_t fallback_sources; /* segment_centrality: sources whose tree has parents with bit-equal seconds competing for a
                                 node; they are served by the heap-order replay launch (centrality.rs:1589 resolves such
                                 ties by BinaryHeap pop order) */
    uint32_t smem_bytes;      /* reserved (0) */
    uint32_t ctas_per_sm;     /* reserved (0) */
    uint32_t reach_capacity;  /* nodes (junctions, for the chain-contracted kernel) a source may reach: arena capacity */
    uint32_t slot_capacity;   /* reserved (0) */
    uint32_t kernel_used;     /* centrality_shortest: 1 global-arena kernel, 3 chain-contracted kernel */
    uint32_t reserved;
} cs_stats;

const char* cs_last_error(void);
int cs_device_count(void);

/* Page-locked host buffers for results (the reference hands back numpy arrays it owns, common.rs:40-53; here the
 * [M][D][node_bound] result is downloaded straight into page-locked memory at full PCIe rate).  Freed buffers return
 * to a small pool and are reused by later calls of the same size.  `cs_host_alloc` returns NULL on failure. */
void* cs_host_alloc(uint64_t bytes);
void cs_host_free(void* ptr);

/* Replaces NetworkStructure construction + add_street_node / add_street_edge ingest (graph.rs:427-452, :728-889) and
 * validate() (:1035-1058): takes the container's payload fields as flat arrays, builds both CSR orientations with
 * 16-byte edge records in petgraph adjacency order (newest edge first), uploads once to `device`.
 *   node arrays [node_bound]: exists, live, weight, xs / ys (coordinates; may both be NULL), z (NaN = no elevation).
 *                             The coordinates only order the device copy of the graph along a Hilbert curve so that
 *                             the nodes one source reaches are contiguous in memory; results do not depend on them.
 *   edge arrays [edge_bound]: exists, src, dst, edge_idx (payload key), length, angle_sum, imp_factor,
 *                             seconds (NaN for street edges; finite >= 0 for transport edges, graph.rs:946-985: the
 *                             value edge_travel_seconds returns at any speed, centrality.rs:988-990, length NaN),
 *                             shared_key (dual: id of shared_primal_node_key, else -1),
 *                             stamp (insertion sequence; larger = newer)
 */
cs_graph* cs_graph_create(uint32_t node_bound, const uint8_t* node_exists, const uint8_t* live, const float* weight,
                          const double* xs, const double* ys, const double* z, uint64_t edge_bound, const uint8_t* edge_exists, const uint32_t* src,
                          const uint32_t* dst, const uint32_t* edge_idx, const float* length, const float* angle_sum,
                          const float* imp_factor, const float* seconds, const int32_t* shared_key,
                          const uint64_t* stamp, int is_dual, int device);
void cs_graph_destroy(cs_graph* g);

/* Tunables (0 = keep default): arena capacity in reached nodes per source, near/far bucket width in seconds,
 * resident warps.  */
int cs_graph_configure(cs_graph* g, uint32_t reach_capacity, float delta_seconds, uint32_t workers);

/* Named tunables of the search kernels: "kernel" (0 = choose per call: the chain-contracted kernel when the graph
 * qualifies, else the global-arena kernel; 1 = global-arena kernel only; 3 = require the chain-contracted kernel),
 * "delta_factor" (near/far bucket width in mean edge traversal times, default 12). */
int cs_graph_set_option(cs_graph* g, const char* name, double value);

/* Run subsequent calls on a caller-owned CUDA stream (e.g. torch's current stream) so that a collective enqueued by the
 * caller orders after the kernels; pass NULL to return to the library's own non-blocking stream. */
int cs_graph_set_stream(cs_graph* g, void* cuda_stream);

/* Stage a source plan in device memory ahead of a compute call; a following compute call that passes sources == NULL
 * (and the same n_sources) runs on the resident plan with no host-to-device copy (SourceSamplingPlan,
 * centrality.rs:447-456 — sources, per-source weight, source_eligible). */
int cs_stage_sources(cs_graph* g, uint64_t n_sources, const uint32_t* sources, const float* source_wt,
                     const uint8_t* eligible);

/* Replaces NetworkStructure.centrality_shortest (centrality.rs:1624-1874) after threshold pairing and source planning
 * (which stay on the host: common.rs:239-270, centrality.rs:1032-1139).
 *   distances/betas/seconds [D]   paired thresholds
 *   tolerance                     fraction, >= 1e-4 (validate_tolerance, centrality.rs:34-48)
 *   sources/source_wt [n_sources] eligible sources to search and their weight (node weight / sampling p)
 *   eligible [node_bound]         source_eligible mask (pair weight 0.5 vs 1.0, centrality.rs:1802-1806)
 *   out                           double [7][D][node_bound]: density, farness, cycles, harmonic, beta,
 *                                 betweenness, betweenness_beta.  Host pointer, or device pointer if out_on_device.
 *   accumulate                    0: out is overwritten; 1: added into (device pointer only; multi-call accumulation)
 */
int cs_centrality_shortest(cs_graph* g, int D, const uint32_t* distances, const float* betas, const uint32_t* seconds,
                           float speed_m_s, float tolerance, int compute_closeness, int compute_betweenness,
                           uint64_t n_sources, const uint32_t* sources, const float* source_wt, const uint8_t* eligible,
                           double* out, int out_on_device, int accumulate, cs_stats* stats);

/* Replaces NetworkStructure.centrality_simplest (centrality.rs:1880-2132). out: double [4][D][node_bound]:
 * density, farness, harmonic, betweenness. */
int cs_centrality_simplest(cs_graph* g, int D, const uint32_t* distances, const uint32_t* seconds, float speed_m_s,
                           float tolerance, float angular_scaling_unit, float farness_scaling_offset,
                           int compute_closeness, int compute_betweenness, uint64_t n_sources, const uint32_t* sources,
                           const float* source_wt, const uint8_t* eligible, double* out, int out_on_device,
                           int accumulate, cs_stats* stats);

/* Replaces NetworkStructure.segment_centrality (centrality.rs:2134-2407). out: double [4][D][node_bound]:
 * segment density, harmonic, beta, betweenness. */
int cs_segment_centrality(cs_graph* g, int D, const uint32_t* distances, const float* betas, const uint32_t* seconds,
                          float speed_m_s, int compute_closeness, int compute_betweenness, uint64_t n_sources,
                          const uint32_t* sources, double* out, int out_on_device, int accumulate, cs_stats* stats);

/* Replaces NetworkStructure.betweenness_od_shortest (centrality.rs:2419-2540): the dependency pass of centrality_shortest
 * seeded only at the OD destinations of every origin (weight w, beta seed w * exp(-beta * cost)); credits are not scaled
 * by a source weight.  `sources` are the live origins with outbound trips; the destinations / weights of sources[k] are
 * od_dst / od_w [od_off[k], od_off[k + 1]) (destinations unique per origin, as the reference's HashMap makes them).
 * `out` uses the [7][D][node_bound] layout of cs_centrality_shortest; only rows 5 (betweenness) and 6 (betweenness_beta)
 * are populated.  Served by the same kernel choice as cs_centrality_shortest (chain-contracted on decomposed graphs). */
int cs_betweenness_od_shortest(cs_graph* g, int D, const uint32_t* distances, const float* betas, const uint32_t* seconds,
                               float speed_m_s, float tolerance, uint64_t n_sources, const uint32_t* sources,
                               const uint64_t* od_off, const uint32_t* od_dst, const float* od_w, double* out,
                               int out_on_device, cs_stats* stats);

/* Replaces NetworkStructure.dijkstra_tree_shortest (centrality.rs:1141-1200, :1499-1508): one capped search from
 * `src`; `visited_order[0 .. *n_visited)` are the settled nodes in pop order, `pred[i]` the predecessor of node i
 * (-1 = none) and `agg_seconds[i]` its travel time (inf = not reached); pred / agg_seconds are sized node_bound,
 * visited_order at least node_bound. */
int cs_dijkstra_tree_shortest(cs_graph* g, uint32_t src, uint32_t max_seconds, float speed_m_s, uint32_t* n_visited,
                              uint32_t* visited_order, int64_t* pred, float* agg_seconds);

/* Batched form of cs_dijkstra_tree_shortest: the searches that feed the reference's data.rs aggregations
 * (data.rs:520-602 call dijkstra_tree_shortest once per data point, centrality.rs:1141-1200) for many sources in ONE
 * launch, one warp per source, every search replayed in the reference's heap order (visit order and predecessors exact
 * under ties).  Outputs are [n_sources][capacity]: for source slot s and k < counts[s], visited_order[s][k] is the k-th
 * settled node, pred[s][k] its predecessor (-1 = none, the source) and agg_seconds[s][k] its travel time.  Fails when a
 * source settles more than `capacity` nodes. */
int cs_dijkstra_trees_shortest(cs_graph* g, uint64_t n_sources, const uint32_t* sources, uint32_t max_seconds,
                               float speed_m_s, uint32_t capacity, uint32_t* counts, uint32_t* visited_order,
                               int64_t* pred, float* agg_seconds);

/* Replaces NetworkStructure.dijkstra_tree_segment (centrality.rs:1523-1611): the single-predecessor tree of one capped
 * search over incoming edges with the visited-edge list.  `visited_nodes[0 .. *n_visited)` in pop order;
 * `visited_edges[0 .. *n_visited_edges)` are container edge ids (petgraph EdgeIndex) in the order the reference pushes
 * them (the caller fills EdgeVisit.start_nd_idx = the edge's target, end_nd_idx = its source, edge_idx = its payload
 * key); per node [node_bound]: pred (-1 none), agg_seconds (inf = not reached), origin_seg / last_seg (edge ids,
 * -1 none), flags (bit 0 visited, bit 1 discovered).  visited_nodes is sized node_bound, visited_edges edge_bound. */
int cs_dijkstra_tree_segment(cs_graph* g, uint32_t src, uint32_t max_seconds, float speed_m_s, uint32_t* n_visited,
                             uint32_t* visited_nodes, uint64_t* n_visited_edges, uint32_t* visited_edges, int64_t* pred,
                             float* agg_seconds, int64_t* origin_seg, int64_t* last_seg, uint8_t* flags);

/* Replaces NetworkStructure.dijkstra_tree_simplest (centrality.rs:1510-1521, dijkstra_tree_angular :1202-1332): the
 * doubled-state angular search collapsed to its node-level tree.  `visited_nodes` lists nodes in first-reached order
 * (the source first); per node [node_bound]: pred (-1 none), simpl_dist (summed angle, inf = not reached),
 * agg_seconds, flags (bit 0 visited, bit 1 discovered).  Fails on a primal graph like the reference. */
int cs_dijkstra_tree_simplest(cs_graph* g, uint32_t src, uint32_t max_seconds, float speed_m_s, uint32_t* n_visited,
                              uint32_t* visited_nodes, int64_t* pred, float* simpl_dist, float* agg_seconds,
                              uint8_t* flags);

/* Replaces NetworkStructure.progress() (graph.rs:413): sources finished by the call in flight on this graph
 * (readable from another host thread while a compute call blocks). */
uint64_t cs_progress(cs_graph* g);

/* Per-source search dump for distance-level parity tests (no reference counterpart; mirrors what
 * dijkstra_brandes_shortest leaves in BrandesTraversal, centrality.rs:418-424): arrays sized node_bound. */
int cs_shortest_search(cs_graph* g, uint32_t src, uint32_t max_seconds, float speed_m_s, float tolerance,
                       float* agg_seconds, double* sigma, uint32_t* pred_count);

#ifdef __cplusplus
}
#endif
#endif /* CITYSEER_B200_H */
