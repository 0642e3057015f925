"""Host-side NetworkStructure container: petgraph StableGraph index semantics, WKT measurement, bulk ingest, sampling plan
(reference: rust/src/graph.rs, tests/test_graph_mutation.py, tests/tools/test_io.py:244-290)."""
import math

import numpy as np
import pytest

import helpers as H
from cityseer_b200 import config, sampling, synth
from cityseer_b200.rustalgos import graph as G
from cityseer_b200.tools import graphs, io, mock


def test_ingest_conventions():
    g, nodes, edges, ns = H.primal_ns()
    assert ns.street_node_count() == g.number_of_nodes() == 57
    assert ns.edge_count == 2 * g.number_of_edges() == 158
    f = ns.frozen()
    assert np.all(f.imp == 1.0) and np.all(np.isnan(f.seconds)) and np.all(f.angle_sum == 0)
    assert list(nodes.index) == ns.node_keys_py()
    s, e, k = ns.edge_references()[0]
    assert ns.get_edge_length(s, e, k) == pytest.approx(math.hypot(90, 80), rel=1e-6)


def test_dual_construction():
    # tests/tools/test_graphs.py:638-682
    _g, nodes, _e, ns = H.diamond_ns(dual=True)
    assert ns.node_count() == 5 and ns.edge_count == 16 and ns.is_dual
    f = ns.frozen()
    assert np.allclose(f.length, 100.0, atol=1e-3)
    assert np.allclose(np.sort(np.unique(np.round(f.angle_sum))), [60.0, 120.0])
    assert list(nodes.index) == ["0_1_k0", "0_2_k0", "1_2_k0", "1_3_k0", "2_3_k0"]
    _g, _n, _e, nsd = H.dual_ns()
    assert nsd.node_count() == 79 and nsd.edge_count == 2 * 155


def test_wkt_metrics():
    coords = G.parse_linestring_wkt("LINESTRING (0 0, 100 0, 100 100)")
    length, angle, in_b, out_b = G.linestring_metrics(coords)
    assert length == 200.0 and angle == 90.0 and in_b == 0.0 and out_b == 90.0
    assert G.parse_linestring_wkt("LINESTRING Z (0 0 5, 3 4 6)") == [(0.0, 0.0), (3.0, 4.0)]
    coords = G.parse_linestring_wkt("LINESTRING(0 0, 1e2 0,2.5E2 -0.0)")
    assert G.linestring_metrics(coords)[:2] == (250.0, 0.0)
    with pytest.raises(ValueError):
        G.parse_linestring_wkt("POINT (0 0)")
    ns = G.NetworkStructure()
    a = ns.add_street_node("a", 0, 0, True, 1)
    b = ns.add_street_node("b", 10, 0, True, 1)
    with pytest.raises(ValueError, match="Failed to parse WKT"):
        ns.add_street_edge(a, b, 0, "a", "b", "LINESTRING (0 0, x y)")
    with pytest.raises(ValueError, match="at least 2 coordinates"):
        ns.add_street_edge(a, b, 0, "a", "b", "LINESTRING (0 0)")
    with pytest.raises(ValueError, match="Invalid impedance factor"):
        ns.add_street_edge(a, b, 0, "a", "b", "LINESTRING (0 0, 10 0)", imp_factor=0.0)
    with pytest.raises(ValueError, match="weight must be finite and non-negative"):
        ns.add_street_node("c", 0, 0, True, -1.0)


def test_mutation_semantics():
    # tests/test_graph_mutation.py
    _g, _n, _e, ns = H.primal_ns()
    idx = ns.street_node_indices()[0]
    ns.remove_street_node(idx)
    with pytest.raises(ValueError, match="does not exist"):
        ns.set_node_live(idx, True)
    for fn in (ns.get_node_payload_py, ns.get_node_weight, ns.is_node_live):
        with pytest.raises(ValueError, match="node_idx .* does not exist"):
            fn(idx)
    valid, removed = 1, idx
    with pytest.raises(ValueError, match="end_nd_idx .* does not exist"):
        ns.add_street_edge(valid, removed, 999999, "1", "0", "LINESTRING (0 0, 1 1)")
    s, e, k = ns.edge_references()[0]
    ns.remove_street_edge(s, e, k)
    for fn in (ns.get_edge_payload_py, ns.get_edge_length, ns.get_edge_impedance):
        with pytest.raises(ValueError, match="Edge not found"):
            fn(s, e, k)
    ns.remove_street_node(s)
    with pytest.raises(ValueError, match="start_nd_idx .* does not exist"):
        ns.remove_street_edge(s, e, k)


def test_stable_indices_and_free_lists():
    _g, _n, _e, ns = H.primal_ns()
    assert ns.node_bound() == 57
    ns.remove_street_node(56)  # last node: bound shrinks to the highest live index + 1
    assert ns.node_bound() == 56 and ns.node_count() == 56
    ns.remove_street_node(10)
    assert ns.node_bound() == 56 and ns.node_count() == 55 and 10 not in ns.node_indices()
    freed_edges = ns.edge_bound() - 0
    new = ns.add_street_node("new", 1.0, 2.0, True, 1.0)
    assert new == 10  # petgraph reuses the most recently vacated slot
    assert ns.node_count() == 56
    e = ns.add_street_edge(new, 0, 0, "new", "0", "LINESTRING (1 2, 700700 5719700)")
    assert e < freed_edges  # vacated edge slots are reused too
    f = ns.frozen()
    assert f.node_exists.sum() == 56 and f.node_bound == 56
    assert f.stamp[e] == f.stamp.max()  # newest edge iterates first


def test_bulk_ingest_equals_per_call_ingest():
    xy, e = synth.lattice(9, 9, seed=3)
    bulk = synth.primal_network(xy, e)
    import networkx as nx

    g = nx.MultiGraph()
    for i, (x, y) in enumerate(xy):
        g.add_node(str(i), x=float(x), y=float(y))
    for a, b in e:
        g.add_edge(str(int(a)), str(int(b)))
    g = graphs.nx_simple_geoms(g)
    _n, _e, ns = io.network_structure_from_nx(g)
    fa, fb = bulk.frozen(), ns.frozen()
    assert fa.node_bound == fb.node_bound and fa.edge_bound == fb.edge_bound
    ka = sorted(zip(fa.src.tolist(), fa.dst.tolist(), fa.length.tolist()))
    kb = sorted(zip(fb.src.tolist(), fb.dst.tolist(), fb.length.tolist()))
    assert ka == kb
    # (adjacency ORDER may differ: networkx re-orders neighbours when graphs are copied; both are valid insertion orders
    # and the oracle / device always see the same arrays)


def test_dual_bulk_matches_nx_to_dual():
    import networkx as nx

    xy, e = synth.lattice(6, 6, seed=5)
    bulk = synth.dual_network(xy, e).frozen()
    g = nx.MultiGraph()
    for i, (x, y) in enumerate(xy):
        g.add_node(str(i), x=float(x), y=float(y))
    for a, b in e:
        g.add_edge(str(int(a)), str(int(b)))
    _n, _e, ns = io.network_structure_from_nx(graphs.nx_to_dual(graphs.nx_simple_geoms(g)))
    ref = ns.frozen()
    assert bulk.node_bound == ref.node_bound and bulk.edge_bound == ref.edge_bound
    assert np.allclose(np.sort(bulk.length), np.sort(ref.length), rtol=1e-6)
    assert np.allclose(np.sort(bulk.angle_sum), np.sort(ref.angle_sum), atol=1e-3)


def test_sampling_plan():
    _g, _n, _e, ns = H.primal_ns()
    ns.set_node_live(5, False)
    src, wt, elig, n_prog, tracked, scale = ns._prepare_sources(None, None, None, None)
    assert len(src) == 56 and 5 not in src.tolist() and elig[5] == 0 and n_prog == 57 and not tracked and scale == 1.0
    src, wt, elig, n_prog, tracked, scale = ns._prepare_sources(None, None, None, [1, 2, 3])
    assert src.tolist() == [1, 2, 3] and elig.sum() == 3 and tracked and scale == pytest.approx(56 / 3)
    src, wt, *_ = ns._prepare_sources(0.5, None, 7, [1, 2, 3])
    assert np.allclose(wt, 2.0)
    a = ns._prepare_sources(0.3, None, 11, None)[0]
    b = ns._prepare_sources(0.3, None, 11, None)[0]
    assert a.tolist() == b.tolist() and 0 < len(a) < 56
    with pytest.raises(ValueError, match=r"sample_probability must be in \(0.0, 1.0\]"):
        ns._prepare_sources(0.0, None, None, None)
    with pytest.raises(ValueError, match="mutually exclusive"):
        ns._prepare_sources(0.5, np.ones(57, np.float32), None, [1])
    with pytest.raises(ValueError, match="must match node_count"):
        ns._prepare_sources(0.5, np.ones(3, np.float32), None, None)
    with pytest.raises(ValueError, match="out of range"):
        ns._prepare_sources(0.5, np.full(57, 1.5, np.float32), None, None)
    with pytest.raises(ValueError, match="does not exist"):
        ns._prepare_sources(None, None, None, [99])


def test_config_and_schedule():
    assert config.prep_gdf_key("density", 400) == "cc_density_400"
    assert config.prep_gdf_key("harmonic", 800, angular=True) == "cc_harmonic_800_ang"
    assert config.prep_gdf_key("seg_beta", 1600, weighted=True) == "cc_seg_beta_1600_wt"
    assert sampling.compute_distance_p(100) == 1.0
    p5, p20 = sampling.compute_distance_p(5000), sampling.compute_distance_p(20000)
    assert 0 < p20 < p5 < 1.0
    r = math.pi * 5000**2 / 175.0**2
    assert p5 == pytest.approx(math.log(2 * r / 0.1) / (2 * 0.06**2) / r)
    assert sampling.compute_hoeffding_p(float("nan")) == 1.0


def test_wrap_progress_polls_and_reraises():
    class Fake:
        def __init__(self):
            self.n = 0

        def progress(self):
            self.n += 1
            return min(self.n, 3)

    assert config.wrap_progress(3, Fake(), lambda: 42) == 42

    def boom():
        raise ValueError("dual graph")

    with pytest.raises(ValueError, match="dual graph"):
        config.wrap_progress(3, Fake(), boom)


def test_od_matrix_semantics():
    # centrality.rs:54-91: parallel arrays -> {origin: {destination: weight}}, a repeated pair keeps the last weight
    from cityseer_b200.rustalgos.centrality import OdMatrix

    od = OdMatrix([0, 0, 1, 0], [1, 2, 2, 1], [1.0, 2.0, 3.0, 5.5])
    assert od.len() == 3 and od.n_origins() == 2
    assert od.map[0][1] == 5.5 and od.map[1][2] == 3.0
    with pytest.raises(ValueError, match="must have equal length"):
        OdMatrix([0, 1], [1], [1.0, 2.0])
    with pytest.raises(OverflowError):
        OdMatrix([-1], [1], [1.0])
    assert OdMatrix([], [], []).len() == 0


def test_frozen_columns_equal_the_payloads_after_mutations():
    """frozen() copies the columnar mirror kept at mutation time: after adds, removals and slot reuse every column equals
    what a walk over the payload objects gives (the reference's container is walked like that, graph.rs:384-391)."""
    import math

    _g, _n, _e, ns = H.primal_ns()
    ns.remove_street_node(10)
    s, e, k = ns.edge_references()[3]
    ns.remove_street_edge(s, e, k)
    ns.set_node_live(7, False)
    new = ns.add_street_node("new", 123.0, 456.0, True, 2.5, z=7.0)
    ns.add_street_edge(new, 0, 4, "new", "0", "LINESTRING (123 456, 200 456, 200 500)", imp_factor=1.5)
    ns.add_transport_edge(0, new, 9, "0", "new", 33.0)
    f = ns.frozen()
    assert f.node_bound == ns.node_bound() and f.edge_bound == ns.edge_bound()
    for i, p in enumerate(ns._nodes[: f.node_bound]):
        assert f.node_exists[i] == (p is not None)
        if p is None:
            assert f.live[i] == 0 and f.weight[i] == 0 and math.isnan(f.z[i])
            continue
        assert (f.xs[i], f.ys[i]) == (p.x, p.y) and f.live[i] == p.live and f.weight[i] == np.float32(p.weight)
        assert (math.isnan(f.z[i]) and p.z is None) or f.z[i] == p.z
    for i, q in enumerate(ns._edges[: f.edge_bound]):
        assert f.edge_exists[i] == (q is not None)
        if q is None:
            assert math.isnan(f.seconds[i]) and f.imp[i] == 1 and f.shared_key[i] == -1
            continue
        assert (f.src[i], f.dst[i], f.edge_idx[i], f.stamp[i]) == (q._src, q._dst, q.edge_idx, q._stamp)
        assert f.imp[i] == np.float32(q.imp_factor)
        for col, val in ((f.length, q.length), (f.angle_sum, q.angle_sum), (f.seconds, q.seconds)):
            assert (math.isnan(col[i]) and math.isnan(val)) or col[i] == np.float32(val)
    assert f.node_indices.tolist() == ns.node_indices()
    # the transport edge: explicit seconds, NaN length (graph.rs:946-985)
    t = int(np.flatnonzero(np.isfinite(f.seconds))[0])
    assert f.seconds[t] == 33.0 and math.isnan(f.length[t]) and (f.src[t], f.dst[t]) == (0, new)
    with pytest.raises(ValueError, match="Invalid seconds value"):
        ns.add_transport_edge(0, new, 9, "0", "new", -1.0)
    with pytest.raises(NotImplementedError):
        ns.add_transport_node("stop", 0.0, 0.0)


def test_sharded_source_plan_blocks_cover_the_plan():
    """_prepare_sources(shard=(rank, ws)) returns that rank's contiguous block of the same plan; `eligible` always
    describes the whole source set (it decides the pair counts, centrality.rs:1802-1806)."""
    _g, _n, _e, ns = H.primal_ns()
    ns.set_node_live(5, False)
    for kw in ({"source_indices": None, "sample_probability": None}, {"source_indices": np.arange(3, 40, 2), "sample_probability": 0.5}):
        full = ns._prepare_sources(kw["sample_probability"], None, 3, kw["source_indices"])
        parts = [ns._prepare_sources(kw["sample_probability"], None, 3, kw["source_indices"], shard=(r, 3)) for r in range(3)]
        assert np.concatenate([p[0] for p in parts]).tolist() == full[0].tolist()
        assert np.allclose(np.concatenate([p[1] for p in parts]), full[1])
        for p in parts:
            assert np.array_equal(p[2], full[2]) and p[3] == full[3] and p[5] == full[5]


def test_od_lists_and_their_shards():
    """_prepare_od: live origins with trips in node order, CSR offsets; the shards of a multi-rank call are contiguous,
    disjoint, complete, and balanced by trip count."""
    from cityseer_b200.rustalgos.centrality import OdMatrix

    _g, _n, _e, ns = H.primal_ns()
    idx = ns.node_indices()
    rng = np.random.default_rng(2)
    o = rng.choice(idx, 400).tolist() + [idx[3]] * 150  # one origin with many trips
    t = rng.choice(idx, len(o)).tolist()
    w = rng.uniform(0.1, 2.0, len(o)).tolist()
    od = OdMatrix(o, t, w)
    src, off, dst, wt = ns._prepare_od(od)
    assert src.tolist() == sorted(od.map) and off[0] == 0 and off[-1] == len(dst) == od.len() == len(wt)
    for k, origin in enumerate(src.tolist()):
        a, b = int(off[k]), int(off[k + 1])
        assert dict(zip(dst[a:b].tolist(), wt[a:b].tolist())) == {d: np.float32(x) for d, x in od.map[origin].items()}
    for world in (1, 2, 3, 8):
        parts = [ns._prepare_od(od, shard=(r, world)) for r in range(world)]
        assert np.concatenate([p[0] for p in parts]).tolist() == src.tolist()
        assert np.concatenate([p[2] for p in parts]).tolist() == dst.tolist()
        assert all(p[1][0] == 0 and p[1][-1] == len(p[2]) and len(p[1]) == len(p[0]) + 1 for p in parts)
        if world == 2:
            assert abs(len(parts[0][2]) - len(parts[1][2])) <= max(len(x) for x in od.map.values())
    with pytest.raises(ValueError, match="out of range"):
        ns._prepare_od(OdMatrix([idx[0]], [10**6], [1.0]))
    # a dead origin is skipped, like the reference's is_node_live check (centrality.rs:2473)
    ns.set_node_live(idx[3], False)
    assert idx[3] not in ns._prepare_od(od)[0].tolist()


def test_od_matrix_columns_equal_the_insert_loop():
    """The columnar OdMatrix (sorted distinct pairs, last weight wins) against the plain HashMap-insert loop it restates
    (centrality.rs:66-80), for list, numpy and float-valued index inputs."""
    from cityseer_b200.rustalgos.centrality import OdMatrix

    rng = np.random.default_rng(8)
    o = rng.integers(0, 40, 3000)
    d = rng.integers(0, 60, 3000)  # many repeated pairs
    w = rng.uniform(0, 9, 3000)
    ref: dict[int, dict[int, float]] = {}
    for a, b, c in zip(o.tolist(), d.tolist(), w.tolist()):
        ref.setdefault(a, {})[b] = float(np.float32(c))
    for od in (OdMatrix(o.tolist(), d.tolist(), w.tolist()), OdMatrix(o, d, w.astype(np.float32)),
               OdMatrix(o.astype(np.uint32), d.astype(np.int16), w), OdMatrix(o.astype(float).tolist(), d.tolist(), w.tolist())):  # fmt: skip
        assert od.map == ref
        assert od.len() == sum(len(x) for x in ref.values()) and od.n_origins() == len(ref)
        assert np.all(np.diff(od._o) >= 0) and od._w.dtype == np.float32
    with pytest.raises(OverflowError):
        OdMatrix([3, 4], [1, -2], [1.0, 1.0])
    with pytest.raises((TypeError, ValueError)):
        OdMatrix(["a"], [1], [1.0])
    empty = OdMatrix([], [], [])
    assert empty.len() == 0 and empty.n_origins() == 0 and empty.map == {}
