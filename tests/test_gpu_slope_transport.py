"""GPU parity where the two directions of an edge differ, and where an edge carries its own travel time.

* Tobler slope penalty (centrality.rs:969-1007): node ``z`` on both ends makes uphill and downhill seconds differ, which
  feeds the in/out CSR orientations, the chain kernel's two per-direction seconds arrays and its two-front chain merge,
  the segment tree and the angular numerators.  Upstream test: tests/rustalgos/test_centrality.py:392-448.
* transport edges (graph.rs:946-985, centrality.rs:988-990): explicit ``seconds`` returned at any speed.
* a reach overflow must not poison the next call on the same graph (the failing warp leaves dense-map entries behind).
Everything is compared with the CPU oracle on identical inputs: counts bit-exact, floats within 1e-5."""
import numpy as np
import pytest

import helpers as H
from cityseer_b200 import synth
from cityseer_b200.rustalgos.graph import NetworkStructure
from cityseer_b200.tools import io

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def set_kernel(k):
    from cityseer_b200 import _native

    if k:
        _native.DEFAULT_OPTIONS["kernel"] = float(k)
    else:
        _native.DEFAULT_OPTIONS.pop("kernel", None)


@pytest.fixture(autouse=True)
def restore_kernel():
    yield
    set_kernel(0)


def check_shortest(oracle_mod, ns, distances, kernel, **kw):
    set_kernel(kernel)
    d, b, s = H.pair(distances=distances)
    res = ns.centrality_shortest(distances=distances, pbar_disabled=True, **kw)
    assert res.stats["kernel_used"] == kernel
    ref, cnt = oracle_mod.OracleGraph(ns.frozen()).centrality_shortest(d, b, s, H.SPEED, n_threads=8)
    got = res._out
    assert np.array_equal(got[0], ref[0]), "node_density not bit-exact"
    assert np.array_equal(got[2], ref[2]), "node_cycles not bit-exact"
    for m, name in enumerate(("density", "farness", "cycles", "harmonic", "beta", "betweenness", "betweenness_beta")):
        np.testing.assert_allclose(got[m], ref[m], rtol=RTOL, atol=1e-7, err_msg=name)
    for key in ("settled", "edge_iters", "sum_ri", "sum_ci"):
        assert res.stats[key] == cnt[key], key
    return res


def test_slope_directionality_reference_case():
    # tests/rustalgos/test_centrality.py:392-448: 5 m rise over 100 m; downhill < flat 75.0 s < uphill; a missing z on
    # either end means no penalty
    def two_nodes(z0, z1):
        g = H.graph_from_coords({"a": (0.0, 0.0), "b": (100.0, 0.0)}, [("a", "b")], z={"a": z0, "b": z1})
        return io.network_structure_from_nx(g)[2]

    ns = two_nodes(0.0, 5.0)
    flat = 100.0 / H.SPEED
    _o, tree_a = ns.dijkstra_tree_shortest(0, 1000, H.SPEED)  # from a: the search walks the incoming edge b -> a (downhill)
    _o, tree_b = ns.dijkstra_tree_shortest(1, 1000, H.SPEED)  # from b: the incoming edge a -> b (uphill)
    assert tree_a[1].agg_seconds < flat < tree_b[0].agg_seconds
    for z0, z1 in ((None, 5.0), (0.0, None), (None, None)):
        ns = two_nodes(z0, z1)
        _o, t = ns.dijkstra_tree_shortest(0, 1000, H.SPEED)
        assert t[1].agg_seconds == pytest.approx(flat, rel=1e-6)


@pytest.mark.parametrize("kernel", [3, 1], ids=["chain-kernel", "arena-kernel"])
def test_hilly_decomposed_grid_shortest(oracle_mod, kernel):
    # cfg #4 shape (89 % chain interiors) with elevation on every node: both seconds arrays of every chain differ
    ns, _ = synth.config("cfg4", scale=0.06, hilly=True)
    f = ns.frozen()
    fwd = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(f.src, f.dst))}
    asym = sum(1 for (a, b), i in fwd.items() if f.z[a] != f.z[b] and (b, a) in fwd)
    assert asym > 0.9 * len(fwd)
    check_shortest(oracle_mod, ns, [300, 600, 1200], kernel)


@pytest.mark.parametrize("kernel", [3, 1], ids=["chain-kernel", "arena-kernel"])
def test_hilly_sources_inside_chains_with_tolerance(oracle_mod, kernel):
    ns, _ = synth.config("cfg4", scale=0.05, hilly=True)
    set_kernel(kernel)
    d, b, s = H.pair(distances=[400, 800])
    rng = np.random.default_rng(11)
    src = np.sort(rng.choice(ns.node_bound(), 200, replace=False))
    res = ns.centrality_shortest(distances=[400, 800], tolerance=1.0, source_indices=src.tolist(), sample_probability=1.0,
                                 pbar_disabled=True)  # fmt: skip
    assert res.stats["kernel_used"] == kernel
    elig = np.zeros(ns.node_bound(), np.uint8)
    elig[src] = 1
    ref, _ = oracle_mod.OracleGraph(ns.frozen()).centrality_shortest(
        d, b, s, H.SPEED, tol=0.01, sources=src.astype(np.uint32), wt=np.ones(len(src), np.float32), eligible=elig, n_threads=8)  # fmt: skip
    got = res._out
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[2], ref[2])
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-7)


def test_hilly_segment_centrality(oracle_mod):
    ns, _ = synth.config("cfg4", scale=0.05, hilly=True)
    d, b, s = H.pair(distances=[200, 400, 800])
    res = ns.segment_centrality(distances=[200, 400, 800], pbar_disabled=True)
    ref, cnt = oracle_mod.OracleGraph(ns.frozen()).segment_centrality(d, b, s, H.SPEED, n_threads=8)
    np.testing.assert_allclose(res._out, ref, rtol=RTOL, atol=1e-6)
    assert res.stats["settled"] == cnt["settled"]


def test_hilly_dual_simplest(oracle_mod):
    ns, _ = synth.config("cfg3", scale=0.08, hilly=True)
    d, _b, s = H.pair(distances=[500, 1000])
    res = ns.centrality_simplest(distances=[500, 1000], angular_scaling_unit=90, farness_scaling_offset=1, pbar_disabled=True)
    ref, cnt = oracle_mod.OracleGraph(ns.frozen()).centrality_simplest(d, s, H.SPEED, unit=90.0, offset=1.0, n_threads=8)
    got = res._out
    assert np.array_equal(got[0], ref[0]), "density (seconds thresholds under slope) not bit-exact"
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-7)
    assert res.stats["settled"] == cnt["settled"] and res.stats["edge_iters"] == cnt["edge_iters"]


def transport_network(extra=()):
    """mock_graph plus two 'stops' joined by fast transport edges (explicit seconds, both directions)."""
    _g, _n, _e, base = H.primal_ns()
    f = base.frozen()
    ns = NetworkStructure()
    for i in range(f.node_bound):
        ns.add_street_node(i, float(f.xs[i]), float(f.ys[i]), True, 1.0)
    for e in range(f.edge_bound):
        a, b = int(f.src[e]), int(f.dst[e])
        ns.add_street_edge(a, b, int(f.edge_idx[e]), a, b,
                           f"LINESTRING({f.xs[a]} {f.ys[a]}, {f.xs[b]} {f.ys[b]})")  # fmt: skip
    far_a, far_b = 0, int(np.argmax(np.hypot(f.xs - f.xs[0], f.ys - f.ys[0])))
    for a, b, sec in ((far_a, far_b, 40.0), (far_b, far_a, 55.0), (3, 30, 0.5), (30, 3, 12.5)) + tuple(extra):
        ns.add_transport_edge(a, b, 7, a, b, sec)
    return ns


@pytest.mark.parametrize("kernel", [3, 1], ids=["chain-kernel", "arena-kernel"])
def test_transport_edges_shortest(oracle_mod, kernel):
    ns = transport_network()
    f = ns.frozen()
    assert np.isfinite(f.seconds).sum() == 4 and np.isnan(f.length[np.isfinite(f.seconds)]).all()
    check_shortest(oracle_mod, ns, [400, 800, 1600], kernel)
    # explicit seconds do not scale with the walking speed
    d, b, s = H.pair(distances=[800], speed=2.0)
    res = ns.centrality_shortest(distances=[800], speed_m_s=2.0, pbar_disabled=True)
    ref, _ = oracle_mod.OracleGraph(f).centrality_shortest(d, b, s, 2.0, n_threads=8)
    assert np.array_equal(res._out[0], ref[0])
    np.testing.assert_allclose(res._out, ref, rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("kernel", [0, 1], ids=["auto", "arena-kernel"])
def test_zero_second_edge_fails_loudly(kernel):
    # a zero-second edge makes its two ends tie on seconds; the reference orders them by heap pop order.  The device
    # refuses the call instead of guessing (and must not hang: the nodes behind the tie wait for its sigma)
    ns = transport_network(extra=((5, 40, 0.0),))
    set_kernel(kernel)
    with pytest.raises(ValueError, match="zero-second edge"):
        ns.centrality_shortest(distances=[400, 800], pbar_disabled=True)
    ns2 = transport_network()  # the same graph without it computes (fresh arena after the failure)
    ns2.centrality_shortest(distances=[400, 800], pbar_disabled=True)


def test_transport_edge_validation():
    ns = NetworkStructure()
    ns.add_street_node("a", 0.0, 0.0, True, 1.0)
    ns.add_street_node("b", 10.0, 0.0, True, 1.0)
    for bad in (-1.0, float("nan"), float("inf")):
        with pytest.raises(ValueError, match="Invalid seconds value"):
            ns.add_transport_edge(0, 1, 0, "a", "b", bad)
    ns.add_transport_edge(0, 1, 0, "a", "b", 3.0)
    assert ns.edge_count == 1


@pytest.mark.parametrize("kernel", [3, 1], ids=["chain-kernel", "arena-kernel"])
def test_overflow_then_retry_returns_clean_results(oracle_mod, kernel):
    """a reach overflow used to leave stale finite distances in the failing warp's dense map, and the
    next call on the same graph silently lost nodes."""
    ns, _ = synth.config("cfg4", scale=0.06)
    set_kernel(kernel)
    dev = ns.device_graph()
    dev.configure(reach_capacity=64)
    with pytest.raises(ValueError, match="overflow"):
        ns.centrality_shortest(distances=[2000], pbar_disabled=True)
    dev.configure(reach_capacity=4096)
    check_shortest(oracle_mod, ns, [150, 300], kernel)
    # and without reconfiguring: the overflowing call must not leave the arena dirty for a call that fits
    dev.configure(reach_capacity=256)
    with pytest.raises(ValueError, match="overflow"):
        ns.centrality_shortest(distances=[2000], pbar_disabled=True)
    check_shortest(oracle_mod, ns, [60], kernel)


def test_overflow_then_retry_segment(oracle_mod):
    ns, _ = synth.config("cfg4", scale=0.06)
    dev = ns.device_graph()
    dev.configure(reach_capacity=256)
    with pytest.raises(ValueError, match="overflow"):
        ns.segment_centrality(distances=[2000], pbar_disabled=True)
    d, b, s = H.pair(distances=[60])
    res = ns.segment_centrality(distances=[60], pbar_disabled=True)
    ref, _ = oracle_mod.OracleGraph(ns.frozen()).segment_centrality(d, b, s, H.SPEED, n_threads=8)
    np.testing.assert_allclose(res._out, ref, rtol=RTOL, atol=1e-6)


def test_reach_capacity_grows_automatically(oracle_mod):
    """The arena starts at 16 384 reached nodes per source; a call that needs more is repeated with a larger arena
    (the reference has no such limit: its 20 km Greater-London runs reach 69 k nodes)."""
    ns, _ = synth.config("cfg4", scale=0.3)
    set_kernel(1)
    f = ns.frozen()
    centre = np.array([f.xs.mean(), f.ys.mean()])
    src = np.sort(np.argsort(np.hypot(f.xs - centre[0], f.ys - centre[1]))[:6]).astype(np.uint32)
    d, b, s = H.pair(distances=[3500])
    res = ns.centrality_shortest(distances=[3500], source_indices=src.tolist(), sample_probability=1.0, pbar_disabled=True)
    assert res.stats["kernel_used"] == 1
    assert res.stats["settled"] / len(src) > 16384 and res.stats["reach_capacity"] > 16384
    elig = np.zeros(f.node_bound, np.uint8)
    elig[src] = 1
    ref, cnt = oracle_mod.OracleGraph(f).centrality_shortest(d, b, s, H.SPEED, sources=src, wt=np.ones(len(src), np.float32),
                                                             eligible=elig, n_threads=6)  # fmt: skip
    assert np.array_equal(res._out[0], ref[0]) and np.array_equal(res._out[2], ref[2])
    np.testing.assert_allclose(res._out, ref, rtol=RTOL, atol=1e-7)
    assert res.stats["settled"] == cnt["settled"]
