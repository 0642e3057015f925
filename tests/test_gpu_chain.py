"""GPU parity of the chain-contracted centrality_shortest kernel (cs_shortest3.cuh) on the graph shapes that stress its
chain logic: sources inside chains, chains that loop back to their junction, interior-only rings, runs longer than the
per-chain cap, parallel chains that tie, waves meeting inside a chain under the epsilon and the tolerance rule.
Every case is checked against the CPU oracle on identical inputs: counts bit-exact, float metrics within 1e-5."""
import numpy as np
import pytest

import helpers as H
from cityseer_b200 import synth
from cityseer_b200.rustalgos.centrality import validate_tolerance
from cityseer_b200.tools import io

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(autouse=True)
def chain_kernel():
    from cityseer_b200 import _native

    _native.DEFAULT_OPTIONS["kernel"] = 3.0  # required: no silent fallback to another kernel
    yield
    _native.DEFAULT_OPTIONS.pop("kernel", None)


def check(oracle_mod, ns, distances, **kw):
    d, b, s = H.pair(distances=distances)
    res = ns.centrality_shortest(distances=distances, pbar_disabled=True, **kw)
    assert res.stats["kernel_used"] == 3
    og = oracle_mod.OracleGraph(ns.frozen())
    ref, cnt = og.centrality_shortest(d, b, s, H.SPEED, tol=validate_tolerance(kw.get("tolerance")), n_threads=8)
    got = res._out
    assert np.array_equal(got[0], ref[0]), "node_density not bit-exact"
    assert np.array_equal(got[2], ref[2]), "node_cycles not bit-exact"
    for m, name in enumerate(("density", "farness", "cycles", "harmonic", "beta", "betweenness", "betweenness_beta")):
        np.testing.assert_allclose(got[m], ref[m], rtol=RTOL, atol=1e-7, err_msg=name)
    for key in ("settled", "edge_iters", "sum_ri", "sum_ci"):
        assert res.stats[key] == cnt[key], key
    return res


def polyline(points, step):
    """Nodes every ~`step` metres along a polyline (slightly uneven, so that no two pieces are equal)."""
    xy = [points[0]]
    rng = np.random.default_rng(len(points) * 7919 + int(step))
    for a, b in zip(points[:-1], points[1:]):
        a, b = np.asarray(a, float), np.asarray(b, float)
        n = max(1, int(np.ceil(np.linalg.norm(b - a) / step)))
        ts = np.sort(np.concatenate([[1.0], (np.arange(1, n) + rng.uniform(-0.2, 0.2, n - 1)) / n])) if n > 1 else [1.0]
        xy += [tuple(a + (b - a) * t) for t in ts]
    return xy


def build(paths, step=20.0):
    """paths: list of point lists; shared end points (rounded coordinates) become shared nodes."""
    coords, edges, key_of = {}, [], {}

    def node(pt):
        k = (round(pt[0], 6), round(pt[1], 6))
        if k not in key_of:
            key_of[k] = f"n{len(key_of)}"
            coords[key_of[k]] = (float(pt[0]), float(pt[1]))
        return key_of[k]

    for pts in paths:
        xy = polyline(pts, step)
        for a, b in zip(xy[:-1], xy[1:]):
            edges.append((node(a), node(b)))
    g = H.graph_from_coords(coords, edges)
    return io.network_structure_from_nx(g)[2]


def test_long_path_is_cut_into_chains(oracle_mod):
    ns = build([[(0, 0), (1500, 0)]])  # 75 pieces between two dead ends: runs are cut every 12 interiors
    check(oracle_mod, ns, [200, 600, 1200])


def test_interior_only_ring(oracle_mod):
    c = [(300 * np.cos(t), 300 * np.sin(t)) for t in np.linspace(0, 2 * np.pi, 13)[:-1]]
    ns = build([c + [c[0]]])  # no junction at all: every node has two neighbours
    check(oracle_mod, ns, [300, 700, 1500])


def test_loop_chain_on_a_stem(oracle_mod):
    # a lollipop: the ring is one chain from the junction back to itself
    ring = [(400 + 150 * np.cos(t), 150 * np.sin(t)) for t in np.linspace(np.pi, 3 * np.pi, 10)]
    ns = build([[(0, 0), (250, 0)], ring])
    check(oracle_mod, ns, [150, 400, 900])
    check(oracle_mod, ns, [400, 900], tolerance=1.0)


def test_parallel_chains_meet_inside(oracle_mod):
    # three routes of nearly equal length between two junctions: the waves meet inside the chains
    a, b = (0.0, 0.0), (600.0, 0.0)
    ns = build([[a, (300, 40), b], [a, (300, -40.3), b], [a, (300, 120), b], [(-200, 0), a], [b, (800, 0)]])
    check(oracle_mod, ns, [300, 600, 1200])
    check(oracle_mod, ns, [300, 600, 1200], tolerance=0.5)
    check(oracle_mod, ns, [600, 1200], tolerance=2.0)


def test_symmetric_routes_tie_within_epsilon(oracle_mod):
    # mirror-image routes: the two candidate distances at the meeting nodes agree to ~1e-6 relative (epsilon ties)
    a, b = (0.0, 0.0), (500.0, 0.0)
    up = [a, (100, 80), (400, 80), b]
    dn = [a, (100, -80), (400, -80), b]
    ns = build([up, dn, [(-150, 0), a], [b, (650, 0)]], step=25.0)
    check(oracle_mod, ns, [250, 500, 1000])
    check(oracle_mod, ns, [500, 1000], tolerance=1.0)


@pytest.mark.parametrize("tolerance", [None, 0.3, 2.0])
def test_decomposed_grid_with_tolerance(oracle_mod, tolerance):
    ns, _ = synth.config("cfg4", 0.05)
    kw = {} if tolerance is None else {"tolerance": tolerance}
    check(oracle_mod, ns, [400, 800, 1600], **kw)


def test_decomposed_grid_single_threshold_and_many(oracle_mod):
    ns, _ = synth.config("cfg4", 0.04)
    check(oracle_mod, ns, [700])
    check(oracle_mod, ns, [200, 400, 600, 800, 1000])  # D = 5 runs the 8-threshold instantiation


def test_source_subset_inside_chains(oracle_mod):
    ns, _ = synth.config("cfg4", 0.05)
    f = ns.frozen()
    deg = np.bincount(f.src[f.edge_exists.astype(bool)], minlength=f.node_bound)
    interior = np.flatnonzero(deg == 2)[::7][:200]
    d, b, s = H.pair(distances=[500, 1000])
    res = ns.centrality_shortest(distances=[500, 1000], source_indices=interior.tolist(), sample_probability=1.0, pbar_disabled=True)
    assert res.stats["kernel_used"] == 3
    og = oracle_mod.OracleGraph(f)
    elig = np.zeros(f.node_bound, np.uint8)
    elig[interior] = 1  # an explicit source list is the eligible set (centrality.rs:1032-1139)
    ref, cnt = og.centrality_shortest(d, b, s, H.SPEED, sources=interior.astype(np.uint32),
                                      wt=np.ones(len(interior), np.float32), eligible=elig, n_threads=8)  # fmt: skip
    assert np.array_equal(res._out[0], ref[0]) and np.array_equal(res._out[2], ref[2])
    np.testing.assert_allclose(res._out, ref, rtol=RTOL, atol=1e-7)
    assert res.stats["settled"] == cnt["settled"] and res.stats["sum_ci"] == cnt["sum_ci"]


def test_one_way_piece_uses_the_arena_kernel(oracle_mod):
    # an edge without a twin disqualifies the graph for the chain kernel: "auto" serves it with the arena kernel,
    # requiring the chain kernel fails loudly
    from cityseer_b200 import _native
    from cityseer_b200.rustalgos.graph import NetworkStructure

    ns = NetworkStructure()
    for i in range(6):
        ns.add_street_node(f"k{i}", 100.0 * i, 0.0, True, 1.0)
    for i in range(5):
        ns.add_street_edge(i, i + 1, 0, f"k{i}", f"k{i + 1}", f"LINESTRING ({100.0 * i} 0, {100.0 * (i + 1)} 0)")
        if i != 2:
            ns.add_street_edge(i + 1, i, 0, f"k{i + 1}", f"k{i}", f"LINESTRING ({100.0 * (i + 1)} 0, {100.0 * i} 0)")
    with pytest.raises(ValueError, match="chain-contracted kernel cannot serve"):
        ns.centrality_shortest(distances=[300], pbar_disabled=True)
    _native.DEFAULT_OPTIONS["kernel"] = 0.0
    ns._invalidate()
    res = ns.centrality_shortest(distances=[300], pbar_disabled=True)
    assert res.stats["kernel_used"] == 1
    d, b, s = H.pair(distances=[300])
    ref, _ = oracle_mod.OracleGraph(ns.frozen()).centrality_shortest(d, b, s, H.SPEED, n_threads=2)
    np.testing.assert_allclose(res._out, ref, rtol=RTOL, atol=1e-7)


def test_full_size_cfg4_properties_and_kernel_agreement():
    """BASELINE config #4 at full size (1 027 753 nodes): the oracle would need hours, so the chain kernel is checked
    through size-independent properties and against the independently written arena kernel on the same sources."""
    from cityseer_b200 import _native

    ns, _ = synth.config("cfg4")
    f = ns.frozen()
    rng = np.random.default_rng(5)
    picks = rng.choice(f.node_indices, 12288, replace=False)
    a_src, b_src = np.sort(picks[:6144]), np.sort(picks[6144:])
    kw = dict(distances=[500, 1000, 2000], sample_probability=1.0, pbar_disabled=True)
    ra = ns.centrality_shortest(source_indices=a_src.tolist(), **kw)
    rb = ns.centrality_shortest(source_indices=b_src.tolist(), **kw)
    rab = ns.centrality_shortest(source_indices=np.sort(picks).tolist(), compute_betweenness=False, **kw)
    assert ra.stats["kernel_used"] == 3 and rab.stats["kernel_used"] == 3
    # linearity over disjoint source sets (closeness does not depend on which nodes are sources)
    assert np.array_equal(ra._out[0] + rb._out[0], rab._out[0])
    assert np.array_equal(ra._out[2] + rb._out[2], rab._out[2])
    np.testing.assert_allclose(ra._out[1] + rb._out[1], rab._out[1], rtol=1e-9)
    np.testing.assert_allclose(ra._out[3:5] + rb._out[3:5], rab._out[3:5], rtol=1e-9)
    # checksum: the density column sums are the per-threshold reachable-target totals counted on the device
    assert [int(x) for x in ra._out[0].sum(axis=1)] == ra.reachability_totals
    # thresholds nest: anything within 500 m is within 1000 m is within 2000 m
    assert np.all(ra._out[0][0] <= ra._out[0][1]) and np.all(ra._out[0][1] <= ra._out[0][2])
    # the arena kernel, written independently, on the same sources: counts bit-exact, floats to summation order
    _native.DEFAULT_OPTIONS["kernel"] = 1.0
    ns2, _ = synth.config("cfg4")
    r1 = ns2.centrality_shortest(source_indices=a_src.tolist(), **kw)
    assert r1.stats["kernel_used"] == 1
    assert np.array_equal(r1._out[0], ra._out[0]) and np.array_equal(r1._out[2], ra._out[2])
    np.testing.assert_allclose(r1._out, ra._out, rtol=1e-9, atol=1e-9)
    for key in ("settled", "edge_iters", "sum_ri", "sum_ci"):
        assert r1.stats[key] == ra.stats[key], key


def test_small_worker_counts_round_up_to_whole_ctas():
    # cs_graph_configure(workers=...) below one CTA of the widest kernel (16 / 32 warps) must still compute everything:
    # the resident-warp count is rounded up to whole CTAs per kernel, never down to an empty grid
    ns, _ = synth.config("cfg4", 0.04)
    full = ns.centrality_shortest(distances=[400, 800], pbar_disabled=True)
    seg = ns.segment_centrality(distances=[400, 800], pbar_disabled=True)
    assert full.stats["kernel_used"] == 3
    ns2, _ = synth.config("cfg4", 0.04)
    ns2.device_graph().configure(0, 0.0, 8)
    small = ns2.centrality_shortest(distances=[400, 800], pbar_disabled=True)
    seg2 = ns2.segment_centrality(distances=[400, 800], pbar_disabled=True)
    assert small.stats["kernel_used"] == 3 and small.stats["sources"] == full.stats["sources"]
    assert small.stats["workers"] >= 16
    assert np.array_equal(small._out[0], full._out[0]) and np.array_equal(small._out[2], full._out[2])
    np.testing.assert_allclose(small._out, full._out, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(seg2._out, seg._out, rtol=1e-12, atol=1e-9)
    ns2.device_graph().set_option("kernel", 1)
    arena = ns2.centrality_shortest(distances=[400, 800], pbar_disabled=True)
    assert arena.stats["kernel_used"] == 1
    np.testing.assert_allclose(arena._out, full._out, rtol=RTOL, atol=1e-7)


def test_full_size_cfg4_source_sample_vs_oracle(oracle_mod):
    """BASELINE config #4 at full size against the CPU oracle on a sample of sources the oracle finishes in seconds
    (its per-source cost is Theta(N), like the reference's): counts bit-exact, floats to rtol 1e-5, device counters
    equal - for centrality_shortest (chain kernel) and for segment_centrality (configs[3]: 400/800/1600 m)."""
    ns, _ = synth.config("cfg4")
    f = ns.frozen()
    og = oracle_mod.OracleGraph(f)
    rng = np.random.default_rng(11)
    src = np.sort(rng.choice(f.node_indices, 160, replace=False)).astype(np.uint32)
    # centrality_shortest, 500/1000/2000 m
    dist = [500, 1000, 2000]
    d, b, s = H.pair(distances=dist)
    res = ns.centrality_shortest(distances=dist, source_indices=src.tolist(), sample_probability=1.0, pbar_disabled=True)
    assert res.stats["kernel_used"] == 3
    elig = np.zeros(f.node_bound, np.uint8)
    elig[src] = 1
    ref, cnt = og.centrality_shortest(d, b, s, H.SPEED, sources=src, wt=np.ones(len(src), np.float32), eligible=elig,
                                      n_threads=8)  # fmt: skip
    assert np.array_equal(res._out[0], ref[0]) and np.array_equal(res._out[2], ref[2])
    np.testing.assert_allclose(res._out, ref, rtol=RTOL, atol=1e-7)
    for key in ("settled", "edge_iters", "sum_ri", "sum_ci"):
        assert res.stats[key] == cnt[key], key
    # the arena kernel against the same oracle result (not only against the chain kernel)
    dev = ns.device_graph()
    dev.set_option("kernel", 1)
    res1 = ns.centrality_shortest(distances=dist, source_indices=src.tolist(), sample_probability=1.0, pbar_disabled=True)
    dev.set_option("kernel", 3)
    assert res1.stats["kernel_used"] == 1
    assert np.array_equal(res1._out[0], ref[0]) and np.array_equal(res1._out[2], ref[2])
    np.testing.assert_allclose(res1._out, ref, rtol=RTOL, atol=1e-7)
    for key in ("settled", "edge_iters", "sum_ri", "sum_ci"):
        assert res1.stats[key] == cnt[key], key
    # segment_centrality, 400/800/1600 m, same sources through the C ABI's source list
    dist = [400, 800, 1600]
    d, b, s = H.pair(distances=dist)
    got, st = ns.device_graph().segment_centrality(d, b, s, float(np.float32(H.SPEED)), True, True, src, None, len(src))
    ref, cnt = og.segment_centrality(d, b, s, H.SPEED, sources=src, n_threads=8)
    # The exponential terms are differences of two f32 exponentials (centrality.rs:2281-2300, :2380-2391); the device
    # forms each exponential with the platform libm's own algorithm (cs_expf_libm), so the cancellation amplifies nothing:
    # every element holds rtol 1e-5 (atol 1e-6 only absorbs the order of the f64 sums on values of order 1e3).
    for m, name in enumerate(("density", "harmonic", "beta", "betweenness")):
        np.testing.assert_allclose(got[m], ref[m], rtol=RTOL, atol=1e-6, err_msg=name)
    assert st["settled"] == cnt["settled"] and st["edge_iters"] == cnt["edge_iters"]


# ------------------------------------------------------------------------------------------------ segment_centrality
# The chain-contracted segment kernel (cs_segment3.cuh) on the same shapes, against the CPU oracle.  Sources whose tree
# has exactly tied parents are set aside by the kernel and replayed in heap order by the node-level kernel; the counters
# of both launches add up to the oracle's.
def check_segment(oracle_mod, ns, distances, **kw):
    d, b, s = H.pair(distances=distances)
    res = ns.segment_centrality(distances=distances, pbar_disabled=True, **kw)
    assert res.stats["kernel_used"] == 3
    ref, cnt = oracle_mod.OracleGraph(ns.frozen()).segment_centrality(
        d, b, s, H.SPEED, closeness=kw.get("compute_closeness", True), betweenness=kw.get("compute_betweenness", True), n_threads=8)  # fmt: skip
    for m, name in enumerate(("density", "harmonic", "beta", "betweenness")):
        np.testing.assert_allclose(res._out[m], ref[m], rtol=RTOL, atol=1e-6, err_msg=name)
    assert res.stats["settled"] == cnt["settled"] and res.stats["edge_iters"] == cnt["edge_iters"]
    return res


def test_segment_long_path_and_ring(oracle_mod):
    check_segment(oracle_mod, build([[(0, 0), (1500, 0)]]), [200, 600, 1200])
    c = [(300 * np.cos(t), 300 * np.sin(t)) for t in np.linspace(0, 2 * np.pi, 13)[:-1]]
    check_segment(oracle_mod, build([c + [c[0]]]), [300, 700, 1500])


def test_segment_loop_chain_on_a_stem(oracle_mod):
    ring = [(400 + 150 * np.cos(t), 150 * np.sin(t)) for t in np.linspace(np.pi, 3 * np.pi, 10)]
    check_segment(oracle_mod, build([[(0, 0), (250, 0)], ring]), [150, 400, 900])


def test_segment_parallel_and_symmetric_chains(oracle_mod):
    a, b = (0.0, 0.0), (600.0, 0.0)
    ns = build([[a, (300, 40), b], [a, (300, -40.3), b], [a, (300, 120), b], [(-200, 0), a], [b, (800, 0)]])
    check_segment(oracle_mod, ns, [300, 600, 1200])
    a, b = (0.0, 0.0), (500.0, 0.0)
    ns = build([[a, (100, 80), (400, 80), b], [a, (100, -80), (400, -80), b], [(-150, 0), a], [b, (650, 0)]], step=25.0)
    check_segment(oracle_mod, ns, [250, 500, 1000])


@pytest.mark.parametrize("distances", [[700], [400, 800, 1600], [200, 400, 600, 800, 1000]])
def test_segment_decomposed_grid(oracle_mod, distances):
    ns, _ = synth.config("cfg4", 0.05)
    res = check_segment(oracle_mod, ns, distances)
    assert res.stats["fallback_sources"] < 0.05 * res.stats["sources"]  # jittered grid: exact ties are rare


@pytest.mark.parametrize("flags", [(True, False), (False, True)])
def test_segment_flag_combinations_and_slope(oracle_mod, flags):
    ns, _ = synth.config("cfg4", 0.05, hilly=True)
    check_segment(oracle_mod, ns, [300, 900], compute_closeness=flags[0], compute_betweenness=flags[1])


def test_segment_regular_grid_is_replayed_in_heap_order(oracle_mod):
    # equal pieces everywhere: nearly every source has exactly tied parents and goes through the heap-order replay
    gx, gy = np.meshgrid(np.arange(10), np.arange(10), indexing="xy")
    xy = np.stack([gx.ravel() * 100.0, gy.ravel() * 100.0], axis=1)
    idx = np.arange(100).reshape(10, 10)
    e = np.concatenate([np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], 1), np.stack([idx[:-1, :].ravel(), idx[1:, :].ravel()], 1)])
    xy2, e2 = synth.decompose(xy, e, 20.0)
    ns = synth.primal_network(xy2, e2)
    res = check_segment(oracle_mod, ns, [200, 400, 800])
    assert res.stats["fallback_sources"] > 0.5 * res.stats["sources"]


def test_segment_kernels_agree_on_the_sampled_full_size_graph(oracle_mod):
    # the node-level kernel on the same graph and sources as the chain kernel (both are compared with the oracle in
    # test_full_size_cfg4_source_sample_vs_oracle / tests/test_gpu_segment.py; this pins them against each other)
    ns, _ = synth.config("cfg4", 0.12)
    a = ns.segment_centrality(distances=[400, 800, 1600], pbar_disabled=True)
    assert a.stats["kernel_used"] == 3
    ns.device_graph().set_option("kernel", 1)
    b = ns.segment_centrality(distances=[400, 800, 1600], pbar_disabled=True)
    assert b.stats["kernel_used"] != 3
    np.testing.assert_allclose(a._out, b._out, rtol=1e-9, atol=1e-9)
    assert a.stats["settled"] == b.stats["settled"] and a.stats["sum_ci"] == b.stats["sum_ci"]
