"""GPU parity of betweenness_od_shortest (centrality.rs:2419-2540) against the CPU oracle, plus the reference's own OD
tests (tests/test_sampling.py:665-735) restated through the GPU path."""
import numpy as np
import pandas as pd
import pytest

import helpers as H
from cityseer_b200 import rustalgos, synth
from cityseer_b200.metrics import networks

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def od_arrays(ns, od):
    f = ns.frozen()
    live = f.live.astype(bool) & f.node_exists.astype(bool)
    sources, off, dst, w = [], [0], [], []
    for src in f.node_indices.tolist():
        dests = od.map.get(src)
        if not live[src] or not dests:
            continue
        for d_, w_ in dests.items():
            dst.append(d_)
            w.append(w_)
        sources.append(src)
        off.append(len(dst))
    return sources, off, dst, w


def run_both(oracle_mod, ns, od, distances, **kw):
    d, b, s = H.pair(distances=distances)
    res = ns.betweenness_od_shortest(od_matrix=od, distances=distances, pbar_disabled=True, **kw)
    tol = rustalgos.centrality.validate_tolerance(kw.get("tolerance"))
    sources, off, dst, w = od_arrays(ns, od)
    ref = oracle_mod.OracleGraph(ns.frozen()).betweenness_od(d, b, s, H.SPEED, sources, off, dst, w, tol=tol, n_threads=4)
    return res, ref


def test_od_matrix_construction():
    od = rustalgos.centrality.OdMatrix([0, 0, 1], [1, 2, 2], [1.0, 2.0, 3.0])
    assert od.len() == 3 and od.n_origins() == 2
    with pytest.raises(ValueError, match="must have equal length"):
        rustalgos.centrality.OdMatrix([0], [1, 2], [1.0])


def test_od_betweenness_mock_graph(oracle_mod):
    _g, _n, _e, ns = H.primal_ns()
    idx = ns.street_node_indices()
    rng = np.random.default_rng(3)
    o = rng.choice(idx, 300)
    t = rng.choice(idx, 300)
    w = rng.uniform(0.1, 5.0, 300).astype(np.float32)
    od = rustalgos.centrality.OdMatrix(o.tolist(), t.tolist(), w.tolist())
    res, ref = run_both(oracle_mod, ns, od, [400, 800, 1600])
    np.testing.assert_allclose(res._out[5], ref[0], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(res._out[6], ref[1], rtol=RTOL, atol=1e-9)
    assert np.all(res._out[:5] == 0) and res._out[5].max() > 0
    res2, ref2 = run_both(oracle_mod, ns, od, [800], tolerance=2.0)
    np.testing.assert_allclose(res2._out[5], ref2[0], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(res2._out[6], ref2[1], rtol=RTOL, atol=1e-9)


def test_od_betweenness_decomposed_grid(oracle_mod):
    ns, _ = synth.config("cfg4", 0.05)
    f = ns.frozen()
    rng = np.random.default_rng(11)
    o = rng.choice(f.node_indices, 400)
    t = rng.choice(f.node_indices, 400)
    od = rustalgos.centrality.OdMatrix(o.tolist(), t.tolist(), rng.uniform(0.5, 3.0, 400).tolist())
    res, ref = run_both(oracle_mod, ns, od, [500, 1000])
    np.testing.assert_allclose(res._out[5], ref[0], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(res._out[6], ref[1], rtol=RTOL, atol=1e-9)


def _dense_od(ns, seed, n_origins, n_dests):
    """Many destinations per origin (the chain kernel looks reached nodes up by binary search), the origin itself among
    them, origins both at junctions and inside chains."""
    f = ns.frozen()
    rng = np.random.default_rng(seed)
    o, t, w = [], [], []
    for src in rng.choice(f.node_indices, min(n_origins, len(f.node_indices)), replace=False).tolist():
        dests = rng.choice(f.node_indices, min(n_dests, len(f.node_indices)), replace=False).tolist()
        for dst in [src] + [x for x in dests if x != src]:
            o.append(src)
            t.append(dst)
            w.append(float(rng.uniform(0.25, 4.0)))
    return rustalgos.centrality.OdMatrix(o, t, w)


def _od_on_kernel(oracle_mod, ns, od, distances, kernel, **kw):
    ns.device_graph().set_option("kernel", float(kernel))  # required: no silent fallback to another kernel
    try:
        res, ref = run_both(oracle_mod, ns, od, distances, **kw)
    finally:
        ns.device_graph().set_option("kernel", 0.0)
    assert res.stats["kernel_used"] == kernel
    np.testing.assert_allclose(res._out[5], ref[0], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(res._out[6], ref[1], rtol=RTOL, atol=1e-9)
    assert np.all(res._out[:5] == 0) and res._out[5].max() > 0
    return res


@pytest.mark.parametrize("tolerance", [None, 1.0])
def test_od_on_the_chain_kernel_decomposed_grid(oracle_mod, tolerance):
    ns, _ = synth.config("cfg4", 0.05)
    od = _dense_od(ns, 5, 120, 700)
    kw = {} if tolerance is None else {"tolerance": tolerance}
    chain = _od_on_kernel(oracle_mod, ns, od, [400, 800, 1600], 3, **kw)
    arena = _od_on_kernel(oracle_mod, ns, od, [400, 800, 1600], 1, **kw)
    np.testing.assert_allclose(chain._out[5:], arena._out[5:], rtol=1e-9, atol=1e-12)
    assert chain.stats["sources"] == arena.stats["sources"] == od.n_origins()
    # the automatic choice on a decomposed graph is the chain kernel, for OD calls too
    auto = ns.betweenness_od_shortest(od_matrix=od, distances=[800], pbar_disabled=True)
    assert auto.stats["kernel_used"] == 3


def test_od_on_the_chain_kernel_chain_shapes(oracle_mod):
    """Loop chains, interior-only rings, runs cut at the chain cap, waves meeting inside a chain: every node a
    destination of every origin."""
    import test_gpu_chain as C

    a, b = (0.0, 0.0), (600.0, 0.0)
    ring = [(400 + 150 * np.cos(t), 150 * np.sin(t)) for t in np.linspace(np.pi, 3 * np.pi, 10)]
    circle = [(300 * np.cos(t), 300 * np.sin(t)) for t in np.linspace(0, 2 * np.pi, 13)[:-1]]
    shapes = [
        [[(0, 0), (1500, 0)]],
        [circle + [circle[0]]],
        [[(0, 0), (250, 0)], ring],
        [[a, (300, 40), b], [a, (300, -40.3), b], [a, (300, 120), b], [(-200, 0), a], [b, (800, 0)]],
    ]
    for k, paths in enumerate(shapes):
        ns = C.build(paths)
        od = _dense_od(ns, 40 + k, 25, 10_000)
        _od_on_kernel(oracle_mod, ns, od, [300, 700, 1500], 3)
        _od_on_kernel(oracle_mod, ns, od, [700, 1500], 3, tolerance=1.0)


def test_od_on_the_chain_kernel_five_thresholds_and_one(oracle_mod):
    ns, _ = synth.config("cfg4", 0.04)
    od = _dense_od(ns, 9, 60, 300)
    _od_on_kernel(oracle_mod, ns, od, [200, 400, 600, 800, 1000], 3)  # the 8-threshold instantiation
    _od_on_kernel(oracle_mod, ns, od, [900], 3)


def test_reference_od_tests_through_the_gpu():
    # tests/test_sampling.py:681-735
    _g, nodes, _e, ns = H.primal_ns()
    idx = ns.street_node_indices()
    o, t, w = [], [], []
    for i in range(5):
        for j in range(5):
            if i != j:
                o.append(idx[i])
                t.append(idx[j])
                w.append(1.0)
    result = ns.betweenness_od_shortest(od_matrix=rustalgos.centrality.OdMatrix(o, t, w), distances=[500], pbar_disabled=True)
    betw = np.array(result.node_betweenness[500])
    assert len(betw) == len(idx) and np.all(betw >= 0) and np.any(betw > 0)
    zero = rustalgos.centrality.OdMatrix([idx[0], idx[1]], [idx[1], idx[2]], [0.0, 0.0])
    r0 = ns.betweenness_od_shortest(od_matrix=zero, distances=[500], pbar_disabled=True)
    assert np.allclose(np.array(r0.node_betweenness[500]), 0.0)
    with pytest.raises(TypeError):
        ns.betweenness_od_shortest(od_matrix={"a": 1}, distances=[500])
    # wrapper columns (networks.py:443-461)
    df = networks.betweenness_od(ns, pd.DataFrame(index=nodes.index), rustalgos.centrality.OdMatrix(o, t, w), distances=[500])
    assert "cc_betweenness_500" in df.columns and "cc_betweenness_beta_500" in df.columns
    assert np.allclose(df["cc_betweenness_500"].to_numpy(), betw)
