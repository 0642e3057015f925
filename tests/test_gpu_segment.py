"""GPU parity: segment_centrality through the C ABI vs the CPU oracle (f32 integrals, f64 sums: rtol 1e-5)."""
import numpy as np
import pytest

import helpers as H
from cityseer_b200 import synth
from cityseer_b200.tools import graphs, io, mock

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def run_both(oracle_mod, ns, distances, **kw):
    d, b, s = H.pair(distances=distances)
    res = ns.segment_centrality(distances=distances, pbar_disabled=True, **kw)
    og = oracle_mod.OracleGraph(ns.frozen())
    ref, cnt = og.segment_centrality(d, b, s, H.SPEED, closeness=kw.get("compute_closeness", True),
                                     betweenness=kw.get("compute_betweenness", True), n_threads=8)  # fmt: skip
    return res, ref, cnt


def check(got, ref):
    for m, name in enumerate(("density", "harmonic", "beta", "betweenness")):
        np.testing.assert_allclose(got[m], ref[m], rtol=RTOL, atol=1e-6, err_msg=name)


def test_diamond_constants_on_gpu():
    # tests/rustalgos/test_centrality.py:537-598, every row.  The betweenness vector [0, 0, x, 0] (credit at node 2, not
    # node 1) is decided by an exact 200 m tie between the two routes 0-1-3 and 0-2-3 that the reference resolves by the
    # pop order of its BinaryHeap: the device sets such sources aside and replays the heap (cs_seg_replay).
    _g, _n, _e, ns = H.diamond_ns()
    r = ns.segment_centrality(distances=[50, 150, 250], pbar_disabled=True)
    A = 0.01
    assert np.allclose(r.segment_density[50], [100, 150, 150, 100], atol=A)
    assert np.allclose(r.segment_density[150], [400, 500, 500, 400], atol=A)
    assert np.allclose(r.segment_density[250], [500, 500, 500, 500], atol=A)
    assert np.allclose(r.segment_harmonic[50], [7.824046, 11.736069, 11.736069, 7.824046], atol=A)
    assert np.allclose(r.segment_harmonic[150], [10.832201, 15.437371, 15.437371, 10.832201], atol=A)
    assert np.allclose(r.segment_harmonic[250], [11.407564, 15.437371, 15.437371, 11.407565], atol=A)
    assert np.allclose(r.segment_beta[50], [24.54211, 36.813164, 36.813164, 24.54211], atol=A)
    assert np.allclose(r.segment_beta[150], [77.45388, 112.34476, 112.34476, 77.45388], atol=A)
    assert np.allclose(r.segment_beta[250], [133.80203, 177.439, 177.439, 133.80203], atol=A)
    assert np.allclose(r.segment_betweenness[50], [0, 0, 24.542109, 0], atol=A)
    assert np.allclose(r.segment_betweenness[150], [0, 0, 69.78874, 0], atol=A)
    assert np.allclose(r.segment_betweenness[250], [0, 0, 99.76293, 0], atol=A)
    assert r.stats["fallback_sources"] > 0  # the tied sources went through the heap-order replay


def regular_lattice(side=14, spacing=100.0, pieces=1):
    """Unjittered lattice: every street has the same length, so equal-seconds ties between competing tree parents are
    everywhere (the case the reference decides by heap order).  ``pieces`` > 1 decomposes every street into equal parts."""
    gx, gy = np.meshgrid(np.arange(side), np.arange(side), indexing="xy")
    xy = np.stack([gx.ravel() * spacing, gy.ravel() * spacing], axis=1).astype(np.float64)
    idx = np.arange(side * side).reshape(side, side)
    e = np.concatenate([np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], 1), np.stack([idx[:-1, :].ravel(), idx[1:, :].ravel()], 1)])
    if pieces > 1:
        xy, e = synth.decompose(xy, e, spacing / pieces)
    return synth.primal_network(xy, e)


@pytest.mark.parametrize("pieces", [1, 4])
def test_regular_lattice_heap_order_ties(oracle_mod, pieces):
    ns = regular_lattice(pieces=pieces)
    res, ref, cnt = run_both(oracle_mod, ns, [200, 400, 800])
    assert res.stats["fallback_sources"] > 0.5 * res.stats["sources"]  # nearly every source has tied parents
    check(res._out, ref)
    assert res.stats["settled"] == cnt["settled"] and res.stats["edge_iters"] == cnt["edge_iters"]


def test_regular_lattice_tree_matches_oracle(oracle_mod):
    # dijkstra_tree_shortest / dijkstra_tree_segment on the tied lattice: pop order and predecessors as the reference's heap
    ns = regular_lattice(side=9)
    og = oracle_mod.OracleGraph(ns.frozen())
    for src in (0, 40, 80):
        order, tm = ns.dijkstra_tree_shortest(src, 600, H.SPEED)
        o_ref, t_ref = og.dijkstra_tree_shortest(src, 600, H.SPEED)
        assert order == o_ref
        assert [t.pred for t in tm] == [t.pred for t in t_ref]


def test_mock_graph(oracle_mod):
    _g, _n, _e, ns = H.primal_ns()
    res, ref, cnt = run_both(oracle_mod, ns, [200, 400, 800, 5000])
    check(res._out, ref)
    assert res.stats["settled"] == cnt["settled"] and res.stats["edge_iters"] == cnt["edge_iters"]


@pytest.mark.parametrize("flags", [(True, False), (False, True)])
def test_flag_combinations(oracle_mod, flags):
    _g, _n, _e, ns = H.primal_ns()
    res, ref, _ = run_both(oracle_mod, ns, [400, 1600], compute_closeness=flags[0], compute_betweenness=flags[1])
    check(res._out, ref)


def test_decomposed_cfg4_small(oracle_mod):
    ns, _ = synth.config("cfg4", 0.06)
    res, ref, _ = run_both(oracle_mod, ns, [400, 800, 1600])
    check(res._out, ref)


def test_perturbed_grid_with_impedance(oracle_mod):
    xy, e = synth.lattice(28, 28, seed=11)
    n = len(xy)
    src, dst = synth._directed_in_ingest_order(n, e)
    dd = xy[dst] - xy[src]
    rng = np.random.default_rng(2)
    # symmetric impedance per undirected edge keeps twins consistent but != 1 exercises the *_imp terms
    key = np.minimum(src, dst).astype(np.int64) * n + np.maximum(src, dst)
    imp = (0.8 + (key % 7) * 0.1).astype(np.float32)
    live = np.ones(n, np.uint8)
    live[rng.choice(n, 40, replace=False)] = 0
    from cityseer_b200.rustalgos.graph import NetworkStructure

    ns = NetworkStructure.from_arrays(live=live, weight=np.ones(n, np.float32), src=src, dst=dst,
                                      edge_idx=np.zeros(len(src), np.uint32),
                                      length=np.hypot(dd[:, 0], dd[:, 1]).astype(np.float32), imp_factor=imp)  # fmt: skip
    res, ref, _ = run_both(oracle_mod, ns, [300, 900])
    check(res._out, ref)
    assert np.all(res._out[0][:, live == 0] == 0)  # non-live nodes are never sources


def test_decomposition_invariance():
    # tests/rustalgos/test_centrality.py:622-649: closeness sums over the original nodes survive 20 m decomposition
    g = graphs.nx_simple_geoms(mock.mock_graph())
    _n, _e, ns = io.network_structure_from_nx(g)
    _n2, _e2, nsd = io.network_structure_from_nx(graphs.nx_decompose(g, 20))
    a = ns.segment_centrality(distances=[200, 400, 800, 5000], pbar_disabled=True)
    b = nsd.segment_centrality(distances=[200, 400, 800, 5000], pbar_disabled=True)
    for name in ("segment_density", "segment_beta", "segment_harmonic"):
        assert np.isclose(getattr(a, name)[400].sum(), getattr(b, name)[400][:57].sum(), rtol=1e-4)


def test_one_way_edge_is_rejected():
    _g, _n, _e, ns = H.primal_ns()
    s, e, k = ns.edge_references()[0]
    ns.remove_street_edge(s, e, k)
    with pytest.raises(ValueError, match="Edge not found"):
        ns.segment_centrality(distances=[400], pbar_disabled=True)
