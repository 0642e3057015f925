"""The reference's own known-answer tests for the hot path, replayed through the CUDA library (no oracle involved):
dual-graph routes (tests/rustalgos/test_centrality.py:171-234), tree paths against NetworkX (:139-163) and the
node-order invariance cases (:710-846)."""
import networkx as nx
import numpy as np
import pytest

import helpers as H
from cityseer_b200.tools import graphs, io

pytestmark = pytest.mark.gpu
ATOL = 0.01  # config.ATOL upstream


def _path(tm, keys, target, src):
    p, cur = [], target
    while True:
        p.append(cur)
        if cur == src:
            break
        cur = tm[cur].pred
    return [keys[i] for i in reversed(p)]


def test_dual_routes_on_gpu():
    # tests/rustalgos/test_centrality.py:171-234
    _g, nodes, _e, ns = H.dual_ns()
    keys = list(nodes.index)
    max_s = int(5000 / H.SPEED)
    src, tgt = keys.index("11_6_k0"), keys.index("39_40_k0")
    _v, tm = ns.dijkstra_tree_simplest(src, max_s, H.SPEED)
    assert _path(tm, keys, tgt, src) == ["11_6_k0", "11_14_k0", "10_14_k0", "10_43_k0", "43_44_k0", "40_44_k0", "39_40_k0"]
    _v, tm = ns.dijkstra_tree_shortest(src, max_s, H.SPEED)
    assert _path(tm, keys, tgt, src) == ["11_6_k0", "6_7_k0", "3_7_k0", "3_4_k0", "1_4_k0", "0_1_k0", "0_31_k0", "31_32_k0",
                                         "32_34_k0", "34_37_k0", "37_39_k0", "39_40_k0"]  # fmt: skip
    src, tgt = keys.index("10_43_k0"), keys.index("10_5_k0")
    _v, tm = ns.dijkstra_tree_simplest(src, max_s, H.SPEED)
    assert _path(tm, keys, tgt, src) == ["10_43_k0", "10_5_k0"]  # no side-stepping of the sharp turn


def test_tree_paths_vs_networkx_on_gpu():
    # tests/rustalgos/test_centrality.py:129-163 — the shortest and the segment tree give NetworkX's distances and
    # predecessor paths of equal length
    g, _nodes, _e, ns = H.primal_ns()
    gl = H.nx_length_graph(g)
    for max_dist in (0, 500, 2000, 5000):
        max_s = int(max_dist / H.SPEED)
        for src in range(0, 57, 4):
            nx_d = nx.single_source_dijkstra_path_length(gl, str(src), weight="length", cutoff=max_dist)
            _v, tm_a = ns.dijkstra_tree_shortest(src, max_s, H.SPEED)
            _v, _ve, tm_b, _em = ns.dijkstra_tree_segment(src, max_s, H.SPEED)
            for tm in (tm_a, tm_b):
                for k, v in nx_d.items():
                    if v > max_dist - 0.5:  # seconds cutoff is slightly tighter than the metre cutoff
                        continue
                    assert abs(tm[int(k)].short_dist - v) <= ATOL, (src, k)
                    # walking the predecessors back to the source adds up to the same length
                    cur, hops = int(k), 0
                    while cur != src:
                        cur = tm[cur].pred
                        hops += 1
                        assert cur is not None and hops <= 57
    with pytest.raises(ValueError, match="dual graph"):
        ns.dijkstra_tree_simplest(0, int(5000 / H.SPEED), H.SPEED)


def _label_graph(ordering, coords, edges, live=None):
    label_to_idx = {label: str(i) for i, label in enumerate(ordering)}
    g = nx.MultiGraph()
    g.graph["crs"] = 32630
    for label in ordering:
        x, y = coords[label]
        attrs = {"x": x, "y": y}
        if live is not None:
            attrs["live"] = live[label]
        g.add_node(label_to_idx[label], **attrs)
    for a, b in edges:
        g.add_edge(label_to_idx[a], label_to_idx[b])
    return graphs.nx_simple_geoms(g), label_to_idx


def _by_primal_edge(g_dual, nodes_df, values, idx_to_label):
    out = {}
    for pos, key in enumerate(nodes_df.index):
        d = g_dual.nodes[key]
        edge = tuple(sorted((idx_to_label[d["primal_edge_node_a"]], idx_to_label[d["primal_edge_node_b"]])))
        out[edge] = values[pos]
    return out


def test_simplest_betweenness_invariant_to_node_order():
    # tests/rustalgos/test_centrality.py:710-770: T-junction with an angled branch, three label orders
    coords = {"A": (500000.0, 0.0), "B": (500000.0, 100.0), "C": (500100.0, 100.0), "D": (500050.0, 170.0)}
    edges = [("A", "B"), ("B", "C"), ("B", "D")]
    results = []
    for ordering in (["A", "B", "C", "D"], ["D", "C", "B", "A"], ["C", "A", "D", "B"]):
        g, l2i = _label_graph(ordering, coords, edges)
        i2l = {v: k for k, v in l2i.items()}
        gd = graphs.nx_to_dual(g)
        nodes_df, _e, net = io.network_structure_from_nx(gd)
        res = net.centrality_simplest(compute_closeness=False, compute_betweenness=True, distances=[500])
        results.append(_by_primal_edge(gd, nodes_df, res.node_betweenness[500], i2l))
    for edge in (("A", "B"), ("B", "C"), ("B", "D")):
        vals = [r[edge] for r in results]
        assert all(abs(v - vals[0]) < ATOL for v in vals), (edge, vals)


def test_betweenness_mixed_live_non_live_invariant_to_node_order():
    # tests/rustalgos/test_centrality.py:773-846: corridor A-B-C-D, D non-live, interleaved by index
    coords = {"A": (500000.0, 0.0), "B": (500000.0, 100.0), "C": (500000.0, 200.0), "D": (500000.0, 300.0)}
    edges = [("A", "B"), ("B", "C"), ("C", "D")]
    live = {"A": True, "B": True, "C": True, "D": False}
    shortest, simplest = [], []
    for ordering in (["A", "B", "C", "D"], ["D", "A", "B", "C"], ["B", "D", "A", "C"]):
        g, l2i = _label_graph(ordering, coords, edges, live)
        i2l = {v: k for k, v in l2i.items()}
        _n, _e, net = io.network_structure_from_nx(g)
        gd = graphs.nx_to_dual(g)
        nodes_dual, _e2, net_dual = io.network_structure_from_nx(gd)
        rs = net.centrality_shortest(compute_closeness=False, compute_betweenness=True, distances=[1000])
        ra = net_dual.centrality_simplest(compute_closeness=False, compute_betweenness=True, distances=[1000])
        shortest.append({label: rs.node_betweenness[1000][int(l2i[label])] for label in ordering})
        simplest.append(_by_primal_edge(gd, nodes_dual, ra.node_betweenness[1000], i2l))
    assert any(v > 0 for v in shortest[0].values())
    for label in "ABCD":
        vals = [r[label] for r in shortest]
        assert all(abs(v - vals[0]) < ATOL for v in vals), (label, vals)
    for edge in (("A", "B"), ("B", "C"), ("C", "D")):
        vals = [r[edge] for r in simplest]
        assert all(abs(v - vals[0]) < ATOL for v in vals), (edge, vals)


def test_cfg1_fixture_on_gpu():
    # BASELINE.json configs[0]: mock_graph at 400/800/1600 m against the committed vectors (tests/golden/make_golden.py)
    import os

    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg1_mock_graph.npz"))
    dist = fx["distances"].tolist()
    _g, _n, _e, ns = H.primal_ns()
    r = ns.centrality_shortest(distances=dist, speed_m_s=H.SPEED, pbar_disabled=True)
    got = [r.node_density, r.node_farness, r.node_cycles, r.node_harmonic, r.node_beta, r.node_betweenness,
           r.node_betweenness_beta]  # fmt: skip
    for m, metric in enumerate(got):
        for i, d in enumerate(dist):
            if m in (0, 2):  # density and cycles are counts: bit-exact
                assert np.array_equal(metric[d], fx["shortest"][m][i]), (m, d)
            else:
                np.testing.assert_allclose(metric[d], fx["shortest"][m][i], rtol=1e-5, atol=1e-12, err_msg=f"{m} {d}")
    assert r.stats["settled"] == int(fx["settled"]) and r.stats["edge_iters"] == int(fx["edge_iters"])
    s = ns.segment_centrality(distances=dist, speed_m_s=H.SPEED, pbar_disabled=True)
    for m, metric in enumerate([s.segment_density, s.segment_harmonic, s.segment_beta, s.segment_betweenness]):
        for i, d in enumerate(dist):
            np.testing.assert_allclose(metric[d], fx["segment"][m][i], rtol=1e-5, atol=1e-9, err_msg=f"seg {m} {d}")
    _gd, _nd, _ed, nsd = H.dual_ns()
    a = nsd.centrality_simplest(distances=dist, speed_m_s=H.SPEED, angular_scaling_unit=90.0, farness_scaling_offset=1.0,
                                pbar_disabled=True)  # fmt: skip
    for m, metric in enumerate([a.node_density, a.node_farness, a.node_harmonic, a.node_betweenness]):
        for i, d in enumerate(dist):
            np.testing.assert_allclose(metric[d], fx["simplest"][m][i], rtol=1e-5, atol=1e-9, err_msg=f"ang {m} {d}")
