"""GPU parity of NetworkStructure.dijkstra_tree_shortest (centrality.rs:1141-1200) against the CPU oracle: settle order,
predecessor, seconds and distance of every node, bit for bit (order of bit-equal keys aside)."""
import numpy as np
import pytest

import helpers as H
from cityseer_b200 import synth

pytestmark = pytest.mark.gpu


def compare(oracle_mod, ns, src, max_seconds):
    visited, tree = ns.dijkstra_tree_shortest(src, max_seconds, H.SPEED)
    ov, ot = oracle_mod.OracleGraph(ns.frozen()).dijkstra_tree_shortest(src, max_seconds, H.SPEED)
    # nodes with bit-equal seconds pop in heap-internal order upstream and in (seconds, index) order here:
    # the order must agree wherever the keys are distinct
    ov = [int(x) for x in ov]
    assert sorted(visited) == sorted(ov)
    assert [ot[v].agg_seconds for v in visited] == [ot[v].agg_seconds for v in ov]
    keys = np.array([ot[v].agg_seconds for v in ov], np.float32)
    distinct = np.ones(len(ov), bool)
    distinct[1:] &= keys[1:] != keys[:-1]
    distinct[:-1] &= keys[:-1] != keys[1:]
    assert [v for v, d in zip(visited, distinct) if d] == [v for v, d in zip(ov, distinct) if d]
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
    with pytest.raises(NotImplementedError):
        ns.dijkstra_tree_segment(0, 600, H.SPEED)
