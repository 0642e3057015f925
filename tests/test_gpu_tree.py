"""GPU parity of NetworkStructure.dijkstra_tree_shortest (centrality.rs:1141-1200) against the CPU oracle: settle order,
predecessor, seconds and distance of every node, bit for bit."""
import numpy as np
import pytest

import helpers as H
from cityseer_b200 import synth

pytestmark = pytest.mark.gpu


def compare(oracle_mod, ns, src, max_seconds):
    visited, tree = ns.dijkstra_tree_shortest(src, max_seconds, H.SPEED)
    ov, ot = oracle_mod.OracleGraph(ns.frozen()).dijkstra_tree_shortest(src, max_seconds, H.SPEED)
    # the single-source search is replayed in the reference's heap order (cs_seg_replay): the pop order is exact, nodes
    # with bit-equal seconds included
    assert visited == [int(x) for x in ov]
    assert len(tree) == len(ot)
    for i, (a, b) in enumerate(zip(tree, ot)):
        assert a.visited == b.visited and a.discovered == b.discovered, i
        assert a.pred == b.pred, i
        assert np.float32(a.agg_seconds) == np.float32(b.agg_seconds) or (np.isinf(a.agg_seconds) and np.isinf(b.agg_seconds)), i
        assert np.float32(a.short_dist) == np.float32(b.short_dist) or (np.isinf(a.short_dist) and np.isinf(b.short_dist)), i


def test_tree_mock_graph(oracle_mod):
    _g, _n, _e, ns = H.primal_ns()
    for src in (0, 7, 23, 49, 56):
        compare(oracle_mod, ns, src, 600)
    compare(oracle_mod, ns, 10, 5000)  # the whole component


def test_tree_decomposed_grid(oracle_mod):
    ns, _ = synth.config("cfg4", 0.05)
    f = ns.frozen()
    for src in f.node_indices[:: max(1, len(f.node_indices) // 6)][:6].tolist():
        compare(oracle_mod, ns, int(src), 900)


def test_tree_errors():
    _g, _n, _e, ns = H.primal_ns()
    with pytest.raises(ValueError, match="out of range"):
        ns.dijkstra_tree_shortest(9999, 600, H.SPEED)
    with pytest.raises(ValueError, match="finite and positive"):
        ns.dijkstra_tree_shortest(0, 600, 0.0)
    with pytest.raises(ValueError, match="out of range"):
        ns.dijkstra_tree_segment(9999, 600, H.SPEED)
    with pytest.raises(ValueError, match="dual graph"):
        ns.dijkstra_tree_simplest(0, 600, H.SPEED)


def _feq(a, b):
    return np.float32(a) == np.float32(b) or (np.isinf(a) and np.isinf(b))


def compare_segment(oracle_mod, ns, src, max_seconds):
    """dijkstra_tree_segment (centrality.rs:1523-1611): the device replays the reference's heap, so node order, edge
    order, origin / last segments and the edge map are compared exactly."""
    vn, ve, tree, emap = ns.dijkstra_tree_segment(src, max_seconds, H.SPEED)
    ovn, ove, otree, oemap = oracle_mod.OracleGraph(ns.frozen()).dijkstra_tree_segment(src, max_seconds, H.SPEED)
    assert vn == ovn
    assert ve == ove
    assert len(tree) == len(otree) and len(emap) == len(oemap)
    for i, (a, b) in enumerate(zip(tree, otree)):
        assert (a.visited, a.discovered, a.pred, a.origin_seg, a.last_seg) == (
            b.visited, b.discovered, b.pred, b.origin_seg, b.last_seg), i  # fmt: skip
        assert _feq(a.agg_seconds, b.agg_seconds) and _feq(a.short_dist, b.short_dist), i
        assert np.isinf(a.simpl_dist)
    for i, (a, b) in enumerate(zip(emap, oemap)):
        assert (a.visited, a.start_nd_idx, a.end_nd_idx, a.edge_idx) == b, i
    return len(vn), len(ve)


def compare_simplest(oracle_mod, ns, src, max_seconds):
    """dijkstra_tree_simplest (centrality.rs:1202-1332), exact replay."""
    vn, tree = ns.dijkstra_tree_simplest(src, max_seconds, H.SPEED)
    ovn, otree = oracle_mod.OracleGraph(ns.frozen()).dijkstra_tree_simplest(src, max_seconds, H.SPEED)
    assert vn == ovn
    assert len(tree) == len(otree)
    for i, (a, b) in enumerate(zip(tree, otree)):
        assert (a.visited, a.discovered, a.pred) == (b.visited, b.discovered, b.pred), i
        assert _feq(a.agg_seconds, b.agg_seconds) and _feq(a.simpl_dist, b.simpl_dist), i
        assert a.origin_seg is None and a.last_seg is None and np.isinf(a.short_dist)
    return len(vn)


def test_tree_segment_mock_graph(oracle_mod):
    _g, _n, _e, ns = H.primal_ns()
    for src in (0, 7, 23, 49, 56):
        compare_segment(oracle_mod, ns, src, 600)
    nn, ne = compare_segment(oracle_mod, ns, 10, 5000)  # the whole component
    assert nn == 50 and ne > 50
    assert compare_segment(oracle_mod, ns, 10, 0)[0] == 1  # nothing within reach: the source alone, its edges visited


def test_tree_segment_decomposed_grid(oracle_mod):
    ns, _ = synth.config("cfg4", 0.05)
    f = ns.frozen()
    for src in f.node_indices[:: max(1, len(f.node_indices) // 5)][:5].tolist():
        nn, _ne = compare_segment(oracle_mod, ns, int(src), 900)
        assert nn > 100


def test_tree_segment_after_edits(oracle_mod):
    # removed nodes / edges leave gaps in the edge ids the dump reports (StableGraph semantics, graph.rs:1017-1033)
    _g, _n, _e, ns = H.primal_ns()
    ns.remove_street_node(12)
    for src in (0, 10, 30):
        compare_segment(oracle_mod, ns, src, 1200)


def test_tree_simplest_mock_dual(oracle_mod):
    _g, _n, _e, ns = H.dual_ns()
    f = ns.frozen()
    for src in f.node_indices[::9].tolist():
        compare_simplest(oracle_mod, ns, int(src), 600)
    assert compare_simplest(oracle_mod, ns, int(f.node_indices[3]), 5000) > 40


def test_tree_simplest_dual_grid(oracle_mod):
    ns, _ = synth.config("cfg3", 0.05)
    f = ns.frozen()
    for src in f.node_indices[:: max(1, len(f.node_indices) // 5)][:5].tolist():
        assert compare_simplest(oracle_mod, ns, int(src), 900) > 50


def test_tree_simplest_diamond(oracle_mod):
    _g, _n, _e, ns = H.diamond_ns(dual=True)
    for src in ns.node_indices():
        compare_simplest(oracle_mod, ns, int(src), 1000)


def test_batched_tree_shortest_matches_the_oracle(oracle_mod):
    """cs_dijkstra_trees_shortest: many sources per launch, every search replayed in heap order - visit order,
    predecessors and seconds of each source equal the oracle's single-source tree, also on a tied (regular) lattice."""
    import numpy as np

    from cityseer_b200 import synth

    gx, gy = np.meshgrid(np.arange(12), np.arange(12), indexing="xy")
    xy = np.stack([gx.ravel() * 100.0, gy.ravel() * 100.0], axis=1)
    idx = np.arange(144).reshape(12, 12)
    e = np.concatenate([np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], 1), np.stack([idx[:-1, :].ravel(), idx[1:, :].ravel()], 1)])
    lattice = synth.primal_network(xy, e)
    jittered, _ = synth.config("cfg2", scale=0.1)
    for ns, max_seconds in ((lattice, 500), (jittered, 900)):
        og = oracle_mod.OracleGraph(ns.frozen())
        src = np.arange(0, ns.node_bound(), max(1, ns.node_bound() // 60))
        counts, order, pred, agg = ns.dijkstra_trees_shortest(src, max_seconds, H.SPEED)
        for i, s in enumerate(src.tolist()):
            o_ref, t_ref = og.dijkstra_tree_shortest(int(s), max_seconds, H.SPEED)
            c = int(counts[i])
            assert order[i, :c].tolist() == o_ref
            assert [None if p < 0 else int(p) for p in pred[i, :c]] == [t_ref[n].pred for n in o_ref]
            assert np.array_equal(agg[i, :c], np.array([t_ref[n].agg_seconds for n in o_ref], np.float32))
    with pytest.raises(ValueError, match="output capacity"):
        lattice.dijkstra_trees_shortest([0, 70], 2000, H.SPEED, capacity=8)
