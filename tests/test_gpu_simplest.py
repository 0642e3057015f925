"""GPU parity: centrality_simplest (angular, dual graph) through the C ABI vs the CPU oracle.

The device search replays the reference's settle order exactly (Rust BinaryHeap tie order included), so density is
bit-exact and the float metrics agree to f64 summation order (rtol 1e-5 stated, typically ~1e-15)."""
import numpy as np
import pytest

import helpers as H
from cityseer_b200 import synth
from cityseer_b200.rustalgos.centrality import validate_tolerance
from cityseer_b200.tools import graphs, io

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def run_both(oracle_mod, ns, distances, **kw):
    d, b, s = H.pair(distances=distances)
    res = ns.centrality_simplest(distances=distances, pbar_disabled=True, **kw)
    og = oracle_mod.OracleGraph(ns.frozen())
    ref, cnt = og.centrality_simplest(
        d, s, H.SPEED, tol=validate_tolerance(kw.get("tolerance")), unit=kw.get("angular_scaling_unit", 180.0),
        offset=kw.get("farness_scaling_offset", 1.0), closeness=kw.get("compute_closeness", True),
        betweenness=kw.get("compute_betweenness", True), n_threads=8)  # fmt: skip
    return res, ref, cnt


def check(got, ref):
    assert np.array_equal(got[0], ref[0]), "node_density not bit-exact"
    for m, name in enumerate(("density", "farness", "harmonic", "betweenness")):
        np.testing.assert_allclose(got[m], ref[m], rtol=RTOL, atol=1e-9, err_msg=name)


def test_diamond_dual_constants_on_gpu():
    # tests/rustalgos/test_centrality.py:520-535
    _g, nodes, _e, ns = H.diamond_ns(dual=True)
    r = ns.centrality_simplest(distances=[50, 150, 250], compute_betweenness=False, pbar_disabled=True)
    assert np.allclose(r.node_harmonic[50], [0, 0, 0, 0, 0])
    assert np.allclose(r.node_harmonic[150], [1.95, 1.95, 2.4, 1.95, 1.95], atol=0.01)
    assert np.allclose(r.node_harmonic[250], [2.45, 2.45, 2.4, 2.45, 2.45], atol=0.01)


def test_requires_dual_graph():
    _g, _n, _e, ns = H.primal_ns()
    with pytest.raises(ValueError, match="dual graph"):
        ns.centrality_simplest(distances=[500], pbar_disabled=True)


def test_mock_dual(oracle_mod):
    _g, _n, _e, ns = H.dual_ns()
    res, ref, cnt = run_both(oracle_mod, ns, [400, 800, 1600, 5000])
    check(res._out, ref)
    assert res.stats["settled"] == cnt["settled"] and res.stats["edge_iters"] == cnt["edge_iters"]
    assert res.stats["sum_ri"] == cnt["sum_ri"] and res.stats["sum_ci"] == cnt["sum_ci"]


def test_wrapper_defaults_unit_90(oracle_mod):
    _g, _n, _e, ns = H.dual_ns()
    res, ref, _ = run_both(oracle_mod, ns, [1000, 2000], angular_scaling_unit=90.0, farness_scaling_offset=1.0)
    check(res._out, ref)


def test_plateau_ratio_on_gpu():
    # tests/rustalgos/test_centrality.py:872-890 — zero-angle plateau: equal-cost predecessor pairs must be dropped
    gd = graphs.nx_to_dual(H.plateau_graph())
    nodes, _e, ns = io.network_structure_from_nx(gd)
    r = ns.centrality_simplest(distances=[1000], compute_closeness=False, pbar_disabled=True)
    betw = dict(zip(nodes.index, r.node_betweenness[1000]))
    assert betw["C_D_k0"] > 0
    assert abs(betw["B_C_k0"] / betw["C_D_k0"] - 1.8) < 1e-6


def test_cfg3_small(oracle_mod):
    ns, _ = synth.config("cfg3", 0.12)  # same generator as config #3 (28 x 28 lattice -> ~1.3k dual nodes)
    res, ref, cnt = run_both(oracle_mod, ns, [1000, 2000], angular_scaling_unit=90.0)
    check(res._out, ref)
    assert res.stats["settled"] == cnt["settled"]


def test_tolerance_phase2(oracle_mod):
    ns, _ = synth.config("cfg3", 0.1)
    res, ref, _ = run_both(oracle_mod, ns, [1500], tolerance=2.0)
    check(res._out, ref)


def test_regular_grid_ties(oracle_mod):
    # an unjittered lattice is full of exact angular ties (0 / 90 degree turns): tie order must match the reference heap
    xy, e = synth.lattice(12, 12, jitter=0.0, drop=0.0, seed=1)
    ns = synth.dual_network(xy, e)
    res, ref, _ = run_both(oracle_mod, ns, [400, 800])
    check(res._out, ref)


def test_source_subset_and_nonlive(oracle_mod):
    _g, _n, _e, ns = H.dual_ns()
    ns.set_node_live(3, False)
    ns.set_node_live(40, False)
    res, ref, _ = run_both(oracle_mod, ns, [800, 2000])
    check(res._out, ref)


def test_full_size_cfg3_source_sample_vs_oracle(oracle_mod):
    """BASELINE config #3 at full size (99 868 dual nodes), 1000/2000 m, wrapper defaults (unit 90, offset 1): a
    600-source sample against the oracle; density bit-exact, floats to rtol 1e-5, settled-state counts equal."""
    ns, _ = synth.config("cfg3")
    f = ns.frozen()
    rng = np.random.default_rng(17)
    src = np.sort(rng.choice(f.node_indices, 600, replace=False)).astype(np.uint32)
    dist = [1000, 2000]
    d, _b, s = H.pair(distances=dist)
    res = ns.centrality_simplest(distances=dist, source_indices=src.tolist(), sample_probability=1.0,
                                 angular_scaling_unit=90.0, farness_scaling_offset=1.0, pbar_disabled=True)  # fmt: skip
    elig = np.zeros(f.node_bound, np.uint8)
    elig[src] = 1
    ref, cnt = oracle_mod.OracleGraph(f).centrality_simplest(
        d, s, H.SPEED, unit=90.0, offset=1.0, sources=src, wt=np.ones(len(src), np.float32), eligible=elig, n_threads=8)  # fmt: skip
    check(res._out, ref)
    assert res.stats["settled"] == cnt["settled"]
    assert res.stats["settled"] > 600 * 1000
