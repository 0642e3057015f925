"""Pins the CPU oracle (oracle/oracle.cpp) against the reference's own known-answer tests.

The Rust reference cannot be compiled here, so these vectors — hand constants and NetworkX cross-checks copied from the
reference's test-suite, cited per test — are what anchors the oracle; the GPU parity tests then compare against it."""
import networkx as nx
import numpy as np
import pytest

import helpers as H
from cityseer_b200 import rustalgos

ATOL, RTOL = 0.01, 1e-4  # config.ATOL / config.RTOL of the reference (config.py:63-64)


def _short(oracle_mod, ns, distances=None, betas=None, **kw):
    d, b, s = H.pair(distances=distances, betas=betas)
    og = oracle_mod.OracleGraph(ns.frozen())
    out, cnt = og.centrality_shortest(d, b, s, H.SPEED, **kw)
    return d, H.compact(out, ns.frozen()), cnt


def test_diamond_shortest_constants(oracle_mod):
    # tests/rustalgos/test_centrality.py:462-512
    _g, _n, _e, ns = H.diamond_ns()
    d, out, _ = _short(oracle_mod, ns, distances=[50, 150, 250], betweenness=False)
    dens, far, cyc, harm, beta = out[0], out[1], out[2], out[3], out[4]
    assert np.allclose(dens[0], [0, 0, 0, 0], atol=ATOL, rtol=RTOL)
    assert np.allclose(dens[1], [2, 3, 3, 2], atol=ATOL, rtol=RTOL)
    assert np.allclose(dens[2], [3, 3, 3, 3], atol=ATOL, rtol=RTOL)
    assert np.allclose(far[1], [200, 300, 300, 200], atol=ATOL, rtol=RTOL)
    assert np.allclose(far[2], [400, 300, 300, 400], atol=ATOL, rtol=RTOL)
    assert np.allclose(cyc[0], [0, 0, 0, 0], atol=ATOL, rtol=RTOL)
    assert np.allclose(cyc[1], [4, 4, 4, 4], atol=ATOL, rtol=RTOL)
    assert np.allclose(cyc[2], [6, 6, 6, 6], atol=ATOL, rtol=RTOL)
    assert np.allclose(harm[1], [0.02, 0.03, 0.03, 0.02], atol=ATOL, rtol=RTOL)
    assert np.allclose(harm[2], [0.025, 0.03, 0.03, 0.025], atol=ATOL, rtol=RTOL)
    assert np.allclose(beta[1], [0.1389669, 0.20845035, 0.20845035, 0.1389669], atol=ATOL, rtol=RTOL)
    assert np.allclose(beta[2], [0.44455525, 0.6056895, 0.6056895, 0.44455522], atol=ATOL, rtol=RTOL)


def test_diamond_simplest_constants(oracle_mod):
    # tests/rustalgos/test_centrality.py:520-535 (Rust defaults: unit 180, offset 1)
    _g, nodes, _e, ns = H.diamond_ns(dual=True)
    assert list(nodes.index) == ["0_1_k0", "0_2_k0", "1_2_k0", "1_3_k0", "2_3_k0"]
    d, b, s = H.pair(distances=[50, 150, 250])
    og = oracle_mod.OracleGraph(ns.frozen())
    out, _ = og.centrality_simplest(d, s, H.SPEED, betweenness=False)
    harm = H.compact(out, ns.frozen())[2]
    assert np.allclose(harm[0], [0, 0, 0, 0, 0], atol=ATOL, rtol=RTOL)
    assert np.allclose(harm[1], [1.95, 1.95, 2.4, 1.95, 1.95], atol=ATOL, rtol=RTOL)
    assert np.allclose(harm[2], [2.45, 2.45, 2.4, 2.45, 2.45], atol=ATOL, rtol=RTOL)


def test_diamond_segment_constants(oracle_mod):
    # tests/rustalgos/test_centrality.py:537-598
    _g, _n, _e, ns = H.diamond_ns()
    d, b, s = H.pair(distances=[50, 150, 250])
    og = oracle_mod.OracleGraph(ns.frozen())
    out, _ = og.segment_centrality(d, b, s, H.SPEED)
    dens, harm, beta, betw = H.compact(out, ns.frozen())
    assert np.allclose(dens[0], [100, 150, 150, 100], atol=ATOL, rtol=RTOL)
    assert np.allclose(dens[1], [400, 500, 500, 400], atol=ATOL, rtol=RTOL)
    assert np.allclose(dens[2], [500, 500, 500, 500], atol=ATOL, rtol=RTOL)
    assert np.allclose(harm[0], [7.824046, 11.736069, 11.736069, 7.824046], atol=ATOL, rtol=RTOL)
    assert np.allclose(harm[1], [10.832201, 15.437371, 15.437371, 10.832201], atol=ATOL, rtol=RTOL)
    assert np.allclose(harm[2], [11.407564, 15.437371, 15.437371, 11.407565], atol=ATOL, rtol=RTOL)
    assert np.allclose(beta[0], [24.54211, 36.813164, 36.813164, 24.54211], atol=ATOL, rtol=RTOL)
    assert np.allclose(beta[1], [77.45388, 112.34476, 112.34476, 77.45388], atol=ATOL, rtol=RTOL)
    assert np.allclose(beta[2], [133.80203, 177.439, 177.439, 133.80203], atol=ATOL, rtol=RTOL)
    # the credit lands at node 2 only with newest-first adjacency + the Rust heap's tie order
    assert np.allclose(betw[0], [0, 0, 24.542109, 0], atol=ATOL, rtol=RTOL)
    assert np.allclose(betw[1], [0, 0, 69.78874, 0], atol=ATOL, rtol=RTOL)
    assert np.allclose(betw[2], [0, 0, 99.76293, 0], atol=ATOL, rtol=RTOL)


def test_mock_closeness_vs_networkx(oracle_mod):
    # tests/rustalgos/test_centrality.py:247-341
    g, _n, _e, ns = H.primal_ns()
    betas = [0.02, 0.01, 0.005, 0.0008]
    d, out, _ = _short(oracle_mod, ns, betas=betas, betweenness=False)
    assert d == [200, 400, 800, 5000]
    assert set(np.unique(out[0][3]).tolist()) <= {49.0, 3.0, 1.0, 0.0}
    gl = H.nx_length_graph(g)
    nx_harm = nx.harmonic_centrality(gl, distance="length")
    for i in range(57):
        assert abs(nx_harm[str(i)] - out[3][3][i]) < ATOL
    # full restatement of the five metrics from per-source NetworkX distances (cycles are target-aggregated, :329)
    n = 57
    dens = np.zeros((4, n)); far = np.zeros((4, n)); cyc = np.zeros((4, n)); harm = np.zeros((4, n)); grav = np.zeros((4, n))
    for src in range(n):
        dists = nx.single_source_dijkstra_path_length(gl, str(src), weight="length")
        for di, cutoff in enumerate(d):
            inside = {k for k, v in dists.items() if v <= cutoff}
            eids = set()
            for u, v, k in gl.edges(keys=True):
                if u != v and u in inside and v in inside:
                    eids.add(tuple(sorted((u, v))) + (k,))
            score = max(0, len(eids) - len(inside) + 1) if inside else 0
            for to, dist in dists.items():
                ti = int(to)
                if ti == src or dist > cutoff:
                    continue
                dens[di][src] += 1
                far[di][src] += dist
                harm[di][src] += 1 / dist
                grav[di][src] += np.exp(-betas[di] * dist)
                cyc[di][ti] += score
    for di in range(4):
        assert np.allclose(out[0][di], dens[di], atol=ATOL, rtol=RTOL)
        assert np.allclose(out[1][di], far[di], atol=ATOL, rtol=1e-3)
        assert np.allclose(out[2][di], cyc[di], atol=ATOL, rtol=RTOL)
        assert np.allclose(out[3][di], harm[di], atol=ATOL, rtol=RTOL)
        assert np.allclose(out[4][di], grav[di], atol=ATOL, rtol=RTOL)


def test_mock_betweenness_vs_networkx(oracle_mod):
    # tests/rustalgos/test_centrality.py:652-674
    g, _n, _e, ns = H.primal_ns()
    d, out, _ = _short(oracle_mod, ns, distances=[5000], closeness=False)
    nx_b = nx.betweenness_centrality(H.nx_length_graph(g), normalized=False, weight="length")
    for i in range(57):
        assert abs(nx_b[str(i)] - out[5][0][i]) < ATOL


def test_node_weights_scale_linearly(oracle_mod):
    # tests/rustalgos/test_centrality.py:343-389
    from cityseer_b200.tools import graphs, io, mock

    g = graphs.nx_simple_geoms(mock.mock_graph())
    _n, _e, ns = io.network_structure_from_nx(g)
    _d, base, _ = _short(oracle_mod, ns, distances=[400, 800], betweenness=False)
    for wt in (0.5, 2):
        gw = g.copy()
        for nd in gw.nodes():
            gw.nodes[nd]["weight"] = wt
        _n2, _e2, nsw = io.network_structure_from_nx(gw)
        _d, outw, _ = _short(oracle_mod, nsw, distances=[400, 800], betweenness=False)
        for m in (0, 1, 3, 4):
            assert np.allclose(outw[m], base[m] * wt, rtol=1e-5, atol=1e-6)
        assert np.allclose(outw[2], base[2])  # cycles unchanged


def test_slope_directionality(oracle_mod):
    # tests/rustalgos/test_centrality.py:392-448 — 5 m rise over 100 m: downhill < flat (75 s) < uphill
    g = H.graph_from_coords({"0": (0.0, 0.0), "1": (100.0, 0.0)}, [("0", "1")], z={"0": 0.0, "1": 5.0})
    from cityseer_b200.tools import io

    _n, _e, ns = io.network_structure_from_nx(g)
    og = oracle_mod.OracleGraph(ns.frozen())
    # search from node 1 walks the incoming edge 0->1 (uphill); from node 0 the edge 1->0 (downhill)
    agg_up, _, _ = og.shortest_distances(1, 1000, H.SPEED)
    agg_dn, _, _ = og.shortest_distances(0, 1000, H.SPEED)
    flat = 100 / H.SPEED
    assert agg_dn[1] < flat < agg_up[0]
    g2 = H.graph_from_coords({"0": (0.0, 0.0), "1": (100.0, 0.0)}, [("0", "1")], z={"0": 0.0})
    _n, _e, ns2 = io.network_structure_from_nx(g2)
    agg_flat, _, _ = oracle_mod.OracleGraph(ns2.frozen()).shortest_distances(1, 1000, H.SPEED)
    assert abs(agg_flat[0] - flat) < 1e-3


def test_dual_routes(oracle_mod):
    # tests/rustalgos/test_centrality.py:171-234
    _g, nodes, _e, ns = H.dual_ns()
    keys = list(nodes.index)
    og = oracle_mod.OracleGraph(ns.frozen())
    max_s = int(5000 / H.SPEED)

    def path(tm, target, src):
        p, cur = [], target
        while True:
            p.append(cur)
            if cur == src:
                break
            cur = tm[cur].pred
        return [keys[i] for i in reversed(p)]

    src, tgt = keys.index("11_6_k0"), keys.index("39_40_k0")
    _v, tm = og.dijkstra_tree_simplest(src, max_s, H.SPEED)
    assert path(tm, tgt, src) == ["11_6_k0", "11_14_k0", "10_14_k0", "10_43_k0", "43_44_k0", "40_44_k0", "39_40_k0"]
    _v, tm = og.dijkstra_tree_shortest(src, max_s, H.SPEED)
    assert path(tm, tgt, src) == ["11_6_k0", "6_7_k0", "3_7_k0", "3_4_k0", "1_4_k0", "0_1_k0", "0_31_k0", "31_32_k0",
                                  "32_34_k0", "34_37_k0", "37_39_k0", "39_40_k0"]  # fmt: skip
    src, tgt = keys.index("10_43_k0"), keys.index("10_5_k0")
    _v, tm = og.dijkstra_tree_simplest(src, max_s, H.SPEED)
    assert path(tm, tgt, src) == ["10_43_k0", "10_5_k0"]


def test_mock_tree_shortest_vs_networkx(oracle_mod):
    # tests/rustalgos/test_centrality.py:129-163
    g, _n, _e, ns = H.primal_ns()
    gl = H.nx_length_graph(g)
    og = oracle_mod.OracleGraph(ns.frozen())
    for max_dist in (0, 500, 2000, 5000):
        for src in range(57):
            nx_d = nx.single_source_dijkstra_path_length(gl, str(src), weight="length", cutoff=max_dist)
            _v, tm = og.dijkstra_tree_shortest(src, int(max_dist / H.SPEED), H.SPEED)
            for k, v in nx_d.items():
                if v > max_dist - 0.5:  # seconds cutoff is slightly tighter than the metre cutoff
                    continue
                assert abs(tm[int(k)].short_dist - v) <= ATOL


def test_plateau_ratio(oracle_mod):
    # tests/rustalgos/test_centrality.py:872-890
    from cityseer_b200.tools import graphs, io

    gd = graphs.nx_to_dual(H.plateau_graph())
    nodes, _e, ns = io.network_structure_from_nx(gd)
    d, b, s = H.pair(distances=[1000])
    og = oracle_mod.OracleGraph(ns.frozen())
    out, _ = og.centrality_simplest(d, s, H.SPEED, closeness=False)
    betw = dict(zip(nodes.index, H.compact(out, ns.frozen())[3][0]))
    assert betw["C_D_k0"] > 0
    assert abs(betw["B_C_k0"] / betw["C_D_k0"] - 1.8) < 1e-6


def test_tolerance_drift(oracle_mod):
    # tests/rustalgos/test_centrality.py:893-930
    from cityseer_b200.tools import io
    from cityseer_b200.rustalgos.centrality import validate_tolerance

    g = H.tolerance_drift_graph()
    nodes, _e, ns = io.network_structure_from_nx(g)
    idx = {k: i for i, k in enumerate(nodes.index)}
    f = ns.frozen()
    d, b, s = H.pair(distances=[20])
    og = oracle_mod.OracleGraph(f)
    res = {}
    for tol_pct in (0.0, 10.0):
        sources = np.array([idx["S"]], np.uint32)
        wt = np.ones(1, np.float32)
        elig = np.zeros(f.node_bound, np.uint8)
        elig[idx["S"]] = 1
        out, _ = og.centrality_shortest(d, b, s, H.SPEED, tol=validate_tolerance(tol_pct), closeness=False,
                                        sources=sources, wt=wt, eligible=elig)  # fmt: skip
        res[tol_pct] = out[5][0]
    assert res[0.0][idx["A"]] == 0 and res[0.0][idx["B"]] == 0 and res[0.0][idx["C"]] > 0
    assert res[10.0][idx["A"]] == 0 and res[10.0][idx["B"]] > 0 and res[10.0][idx["C"]] > 0


def test_threshold_pairing_tables():
    # tests/rustalgos/test_common.py:32-185
    assert rustalgos.distances_from_betas([0.04, 0.0025]) == [100, 1600]
    assert np.allclose(rustalgos.betas_from_distances([173], min_threshold_wt=0.001), [0.0399292], atol=1e-5)
    dist = [400, 600, 800, 1600, 2000, 10000, 20000]
    betas = [0.01, 0.00667, 0.005, 0.0025, 0.002, 0.0004, 0.0002]
    secs = [300, 450, 600, 1200, 1500, 7500, 15000]
    d, b, s = rustalgos.pair_distances_betas_time(H.SPEED, distances=dist)
    assert d == dist and s == secs and np.allclose(b, betas, rtol=1e-6)
    d, b, s = rustalgos.pair_distances_betas_time(H.SPEED, betas=betas)
    assert s == secs and np.allclose(d, dist, rtol=2e-3)
    d, b, s = rustalgos.pair_distances_betas_time(H.SPEED, minutes=[x / 60 for x in secs])
    assert s == secs and d == dist
    for bad in ({"distances": [400], "betas": [0.01]}, {}):
        with pytest.raises(ValueError):
            rustalgos.pair_distances_betas_time(H.SPEED, **bad)
    with pytest.raises(ValueError):
        rustalgos.betas_from_distances([400, 400])
    with pytest.raises(ValueError):
        rustalgos.distances_from_betas([0.01, 0.02])
    with pytest.raises(TypeError):
        rustalgos.betas_from_distances("boo")
    # SURVEY.md A.1 f32 table
    table = {50: (0.08, 38), 150: (0.02667, 113), 250: (0.016, 188), 500: (0.008, 375), 1000: (0.004, 750), 5000: (0.0008, 3750)}
    for dd, (bb, ss) in table.items():
        d, b, s = rustalgos.pair_distances_betas_time(H.SPEED, distances=[dd])
        assert s == [ss] and abs(b[0] - bb) < 1e-7


def test_committed_fixture_matches_the_oracle(oracle_mod):
    # tests/golden/cfg1_mock_graph.npz (made by tests/golden/make_golden.py): BASELINE.json configs[0]
    import os

    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg1_mock_graph.npz"))
    d, b, s = H.pair(distances=[400, 800, 1600])
    assert fx["distances"].tolist() == list(d) and fx["seconds"].tolist() == list(s)
    assert np.array_equal(fx["betas"], np.array(b, np.float32))
    _g, _n, _e, ns = H.primal_ns()
    f = ns.frozen()
    og = oracle_mod.OracleGraph(f)
    out, cnt = og.centrality_shortest(d, b, s, H.SPEED)
    assert np.array_equal(H.compact(out, f), fx["shortest"])  # same code, same machine arithmetic: bit for bit
    assert cnt["settled"] == int(fx["settled"]) and cnt["edge_iters"] == int(fx["edge_iters"])
    seg, _ = og.segment_centrality(d, b, s, H.SPEED)
    np.testing.assert_allclose(H.compact(seg, f), fx["segment"], rtol=1e-12, atol=0)
    _gd, _nd, _ed, nsd = H.dual_ns()
    fd = nsd.frozen()
    simp, _ = oracle_mod.OracleGraph(fd).centrality_simplest(d, s, H.SPEED, unit=90.0, offset=1.0)
    np.testing.assert_allclose(H.compact(simp, fd), fx["simplest"], rtol=1e-12, atol=0)


def test_optimised_cpu_variant_equals_the_faithful_port(oracle_mod):
    """The sparse-reset "optimised CPU" baseline performs the same arithmetic in the
    same order as the faithful restatement: single-threaded results and counters are identical bit for bit, with and
    without the tolerance pass, on mock_graph (exact-run plan) and on a cfg #4 lattice with a non-live border."""
    from cityseer_b200 import synth

    d, b, s = H.pair(distances=[400, 800, 1600])
    _g, _n, _e, ns = H.primal_ns()
    og = oracle_mod.OracleGraph(ns.frozen())
    for tol in (1e-4, 0.02):
        ref, c_ref = og.centrality_shortest(d, b, s, H.SPEED, tol=tol)
        opt, c_opt = og.centrality_shortest(d, b, s, H.SPEED, tol=tol, optimised=True)
        assert np.array_equal(ref, opt, equal_nan=True) and c_ref == c_opt
    ns4, _ = synth.config("cfg4", scale=0.08)
    f = ns4.frozen()
    og4 = oracle_mod.OracleGraph(f)
    d, b, s = H.pair(distances=[500, 1000, 2000])
    rng = np.random.default_rng(3)
    src = np.sort(rng.choice(f.node_bound, 64, replace=False)).astype(np.uint32)
    elig = (rng.random(f.node_bound) < 0.8).astype(np.uint8)
    wt = rng.uniform(0.5, 2.0, len(src)).astype(np.float32)
    for tol in (1e-4, 0.01):
        ref, c_ref = og4.centrality_shortest(d, b, s, H.SPEED, tol=tol, sources=src, wt=wt, eligible=elig)
        opt, c_opt = og4.centrality_shortest(d, b, s, H.SPEED, tol=tol, sources=src, wt=wt, eligible=elig, optimised=True)
        assert np.array_equal(ref, opt) and c_ref == c_opt
    # threads only change the order of the f64 atomic adds
    par, _ = og4.centrality_shortest(d, b, s, H.SPEED, sources=src, wt=wt, eligible=elig, optimised=True, n_threads=4)
    exact, _ = og4.centrality_shortest(d, b, s, H.SPEED, sources=src, wt=wt, eligible=elig)
    np.testing.assert_allclose(par, exact, rtol=1e-12, atol=0)
