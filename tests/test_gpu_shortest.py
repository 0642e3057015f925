"""GPU parity: centrality_shortest through the C ABI vs the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): node_density / node_cycles / reachable sets bit-exact; float metrics within 1e-5
relative.  f64 accumulation order differs between the two (atomics), hence the tiny tolerance on sums."""
import numpy as np
import pytest

import helpers as H
from cityseer_b200 import rustalgos, synth
from cityseer_b200.rustalgos.centrality import validate_tolerance
from cityseer_b200.tools import graphs, io, mock

pytestmark = pytest.mark.gpu
RTOL = 1e-5  # stated tolerance for floating-point metrics


@pytest.fixture(autouse=True, params=[3, 1], ids=["chain-kernel", "arena-kernel"])
def kernel_choice(request):
    """Every test runs against both search kernels: 3 = chain-contracted kernel (required, no silent fallback),
    1 = global-arena kernel."""
    from cityseer_b200 import _native

    _native.DEFAULT_OPTIONS["kernel"] = float(request.param)
    yield request.param
    _native.DEFAULT_OPTIONS.pop("kernel", None)


def check(got, ref, names=("density", "farness", "cycles", "harmonic", "beta", "betweenness", "betweenness_beta")):
    assert got.shape == ref.shape
    assert np.array_equal(got[0], ref[0]), "node_density not bit-exact"
    assert np.array_equal(got[2], ref[2]), "node_cycles not bit-exact"
    for m, name in enumerate(names):
        np.testing.assert_allclose(got[m], ref[m], rtol=RTOL, atol=1e-7, err_msg=name)


def run_both(oracle_mod, ns, distances, **kw):
    d, b, s = H.pair(distances=distances)
    res = ns.centrality_shortest(distances=distances, pbar_disabled=True, **kw)
    tol = validate_tolerance(kw.get("tolerance"))
    og = oracle_mod.OracleGraph(ns.frozen())
    ref, cnt = og.centrality_shortest(d, b, s, H.SPEED, tol=tol, closeness=kw.get("compute_closeness", True),
                                      betweenness=kw.get("compute_betweenness", True), n_threads=8)  # fmt: skip
    return res, ref, cnt


def test_mock_graph_cfg1(oracle_mod):
    _g, _n, _e, ns = H.primal_ns()
    res, ref, cnt = run_both(oracle_mod, ns, [400, 800, 1600])
    check(res._out, ref)
    assert res.stats["settled"] == cnt["settled"] and res.stats["edge_iters"] == cnt["edge_iters"]
    assert res.stats["sum_ri"] == cnt["sum_ri"] and res.stats["sum_ci"] == cnt["sum_ci"]


def test_diamond_constants_on_gpu():
    # the reference's hand constants (tests/rustalgos/test_centrality.py:462-512) straight through the GPU path
    _g, _n, _e, ns = H.diamond_ns()
    r = ns.centrality_shortest(distances=[50, 150, 250], compute_betweenness=False, pbar_disabled=True)
    assert np.allclose(r.node_density[150], [2, 3, 3, 2]) and np.allclose(r.node_density[250], [3, 3, 3, 3])
    assert np.allclose(r.node_farness[250], [400, 300, 300, 400], rtol=1e-4)
    assert np.allclose(r.node_cycles[150], [4, 4, 4, 4]) and np.allclose(r.node_cycles[250], [6, 6, 6, 6])
    assert np.allclose(r.node_harmonic[250], [0.025, 0.03, 0.03, 0.025], rtol=1e-4)
    assert np.allclose(r.node_beta[250], [0.44455525, 0.6056895, 0.6056895, 0.44455522], atol=0.01)


@pytest.mark.parametrize("flags", [(True, False), (False, True), (True, True)])
def test_flag_combinations(oracle_mod, flags):
    _g, _n, _e, ns = H.primal_ns()
    res, ref, _ = run_both(oracle_mod, ns, [200, 400, 800, 5000], compute_closeness=flags[0], compute_betweenness=flags[1])
    check(res._out, ref)


def test_search_state_exact(oracle_mod):
    # distances (f32 seconds) and sigma per node must be bit-identical to the reference search
    ns, _ = synth.config("cfg2", 0.15)
    f = ns.frozen()
    og = oracle_mod.OracleGraph(f)
    dev = ns.device_graph()
    rng = np.random.default_rng(5)
    for src in rng.choice(f.node_bound, 12, replace=False).tolist():
        agg, sig, _np = dev.shortest_search(src, 1500, H.SPEED)
        ragg, _rc, rsig = og.shortest_distances(src, 1500, H.SPEED)
        assert np.array_equal(agg, ragg)
        assert np.array_equal(sig, rsig)


def test_perturbed_grid_cfg2_small(oracle_mod):
    ns, _ = synth.config("cfg2", 0.2)  # 63 x 63 lattice, same generator as config #2
    res, ref, cnt = run_both(oracle_mod, ns, [500, 1000, 2000])
    check(res._out, ref)
    assert res.stats["settled"] == cnt["settled"]
    assert res.stats["sum_ri"] == cnt["sum_ri"]


def test_decomposed_cfg4_small(oracle_mod):
    ns, _ = synth.config("cfg4", 0.06)  # ~20 x 20 lattice cut into 20 m segments
    res, ref, cnt = run_both(oracle_mod, ns, [500, 1000, 2000])
    check(res._out, ref)
    assert res.stats["edge_iters"] == cnt["edge_iters"]


def test_tolerance_phase2(oracle_mod):
    ns, _ = synth.config("cfg2", 0.12)
    res, ref, _ = run_both(oracle_mod, ns, [1000, 2000], tolerance=1.5)
    check(res._out, ref)
    base = ns.centrality_shortest(distances=[1000, 2000], pbar_disabled=True)
    assert not np.allclose(base._out[5], res._out[5])  # tolerance spreads betweenness


def test_tolerance_drift_graph(oracle_mod):
    # tests/rustalgos/test_centrality.py:893-930 through the GPU path
    nodes, _e, ns = io.network_structure_from_nx(H.tolerance_drift_graph())
    idx = {k: i for i, k in enumerate(nodes.index)}
    r0 = ns.centrality_shortest(distances=[20], compute_closeness=False, source_indices=[idx["S"]], tolerance=0.0, pbar_disabled=True)
    r10 = ns.centrality_shortest(distances=[20], compute_closeness=False, source_indices=[idx["S"]], tolerance=10.0, pbar_disabled=True)
    b0, b10 = r0.node_betweenness[20], r10.node_betweenness[20]
    assert b0[idx["A"]] == 0 and b0[idx["B"]] == 0 and b0[idx["C"]] > 0
    assert b10[idx["A"]] == 0 and b10[idx["B"]] > 0 and b10[idx["C"]] > 0


def test_weights_and_nonlive_border(oracle_mod):
    xy, e = synth.lattice(30, 30, seed=7)
    n = len(xy)
    rng = np.random.default_rng(3)
    live = np.ones(n, np.uint8)
    border = (xy[:, 0] < synth.X0 + 500) | (xy[:, 1] < synth.Y0 + 500)
    live[border] = 0
    src, dst = synth._directed_in_ingest_order(n, e)
    dd = xy[dst] - xy[src]
    from cityseer_b200.rustalgos.graph import NetworkStructure

    ns = NetworkStructure.from_arrays(
        live=live, weight=rng.uniform(0.5, 2.0, n).astype(np.float32), src=src, dst=dst,
        edge_idx=np.zeros(len(src), np.uint32), length=np.hypot(dd[:, 0], dd[:, 1]).astype(np.float32),
        imp_factor=rng.uniform(0.8, 1.3, len(src)).astype(np.float32),
    )  # fmt: skip
    d, b, s = H.pair(distances=[400, 1200])
    res = ns.centrality_shortest(distances=[400, 1200], pbar_disabled=True)
    og = oracle_mod.OracleGraph(ns.frozen())
    ref, _ = og.centrality_shortest(d, b, s, H.SPEED, n_threads=8)
    got = res._out
    # weighted sums are no longer integers: density / cycles compare at float tolerance here
    for m in range(7):
        np.testing.assert_allclose(got[m], ref[m], rtol=RTOL, atol=1e-7)
    assert np.all(got[:, :, :][5][:, live == 0] >= 0)


def test_source_indices_and_scaling(oracle_mod):
    _g, _n, _e, ns = H.primal_ns()
    subset = [0, 5, 17, 30, 44]
    res = ns.centrality_shortest(distances=[800], source_indices=subset, pbar_disabled=True)
    f = ns.frozen()
    d, b, s = H.pair(distances=[800])
    elig = np.zeros(f.node_bound, np.uint8)
    elig[subset] = 1
    ref, _ = oracle_mod.OracleGraph(f).centrality_shortest(
        d, b, s, H.SPEED, sources=np.array(subset, np.uint32), wt=np.ones(len(subset), np.float32), eligible=elig)
    ref[5:7] *= 57 / len(subset)  # n_live / n_sources post-scale (centrality.rs:1852-1868)
    check(res._out, ref)
    assert res.sampled_source_count == len(subset)
    assert res.reachability_totals == [int(ref[0][0].sum())]
    with pytest.raises(ValueError, match="does not exist"):
        ns.centrality_shortest(distances=[800], source_indices=[9999], pbar_disabled=True)


def test_removed_node_gaps(oracle_mod):
    # StableGraph gaps: node_bound > node_count; results are compacted over node_indices (common.rs:40-53)
    g = graphs.nx_simple_geoms(mock.mock_graph())
    _n, _e, ns = io.network_structure_from_nx(g)
    ns.remove_street_node(7)
    ns.remove_street_node(20)
    assert ns.node_bound() == 57 and ns.node_count() == 55
    res, ref, _ = run_both(oracle_mod, ns, [400, 1600])
    check(res._out, ref)
    assert len(res.node_density[400]) == 55
    with pytest.raises(ValueError, match="does not exist"):
        ns.centrality_shortest(distances=[500], source_indices=[7], sample_probability=1.0, pbar_disabled=True)


def test_sampling_semantics():
    # tests/test_sampling.py:45-107 — same seed reproducible; p = 1.0 equals the exact run
    _g, _n, _e, ns = H.primal_ns()
    full = ns.centrality_shortest(distances=[800], pbar_disabled=True)
    p1 = ns.centrality_shortest(distances=[800], sample_probability=1.0, random_seed=1, pbar_disabled=True)
    np.testing.assert_allclose(p1._out, full._out, rtol=1e-12)
    a = ns.centrality_shortest(distances=[800], sample_probability=0.5, random_seed=42, pbar_disabled=True)
    b = ns.centrality_shortest(distances=[800], sample_probability=0.5, random_seed=42, pbar_disabled=True)
    np.testing.assert_allclose(a._out, b._out, rtol=1e-12)
    assert 0 < a.sampled_source_count < 57


def test_progress_and_errors():
    _g, _n, _e, ns = H.primal_ns()
    ns.centrality_shortest(distances=[400])
    assert ns.progress() == 57
    with pytest.raises(ValueError, match="both parameters are False"):
        ns.centrality_shortest(distances=[400], compute_closeness=False, compute_betweenness=False)
    with pytest.raises(ValueError):
        ns.centrality_shortest(distances=[400], tolerance=-1)
    with pytest.raises(ValueError, match="exactly one"):
        ns.centrality_shortest(distances=[400], betas=[0.01])


def test_linearity_property_full_size_graph():
    # size-independent property at a large size: running two disjoint halves of the sources and adding equals one run
    ns, _ = synth.config("cfg2", 0.5)
    f = ns.frozen()
    all_src = f.node_indices
    a = ns.centrality_shortest(distances=[500, 1000], source_indices=all_src[::2].tolist(), sample_probability=1.0, pbar_disabled=True)
    b = ns.centrality_shortest(distances=[500, 1000], source_indices=all_src[1::2].tolist(), sample_probability=1.0, pbar_disabled=True)
    full = ns.centrality_shortest(distances=[500, 1000], compute_betweenness=False, pbar_disabled=True)
    # closeness does not depend on source_eligible, so halves add up exactly for the integer metrics
    assert np.array_equal(a._out[0] + b._out[0], full._out[0])
    assert np.array_equal(a._out[2] + b._out[2], full._out[2])
    np.testing.assert_allclose(a._out[1] + b._out[1], full._out[1], rtol=1e-9)


def test_full_size_cfg5_source_sample_vs_oracle(oracle_mod):
    """BASELINE config #5 (4 000 000 nodes, 14.4 M directed edges) at 5 km: a sample of sources the oracle finishes in
    seconds; counts bit-exact, floats to rtol 1e-5, device counters equal - on both kernels (the graph has no chains: "auto"
    picks the global-arena kernel; the chain kernel then runs with every node a junction)."""
    ns, _ = synth.config("cfg5")
    f = ns.frozen()
    rng = np.random.default_rng(13)
    src = np.sort(rng.choice(f.node_indices, 48, replace=False)).astype(np.uint32)
    dist = [5000]
    d, b, s = H.pair(distances=dist)
    res = ns.centrality_shortest(distances=dist, source_indices=src.tolist(), sample_probability=1.0, pbar_disabled=True)
    elig = np.zeros(f.node_bound, np.uint8)
    elig[src] = 1
    og = oracle_mod.OracleGraph(f)
    ref, cnt = og.centrality_shortest(d, b, s, H.SPEED, sources=src, wt=np.ones(len(src), np.float32), eligible=elig,
                                      n_threads=8)  # fmt: skip
    assert np.array_equal(res._out[0], ref[0]) and np.array_equal(res._out[2], ref[2])
    np.testing.assert_allclose(res._out, ref, rtol=1e-5, atol=1e-7)
    for key in ("settled", "edge_iters", "sum_ri", "sum_ci"):
        assert res.stats[key] == cnt[key], key
    assert res.stats["settled"] > 48 * 4000  # about 5 400 nodes within 5 km of a source


def test_full_size_cfg2_source_sample_vs_oracle(oracle_mod):
    """BASELINE config #2 at full size (99 856 nodes), 500/1000/2000 m: a 1500-source sample against the oracle."""
    ns, _ = synth.config("cfg2")
    f = ns.frozen()
    rng = np.random.default_rng(19)
    src = np.sort(rng.choice(f.node_indices, 1500, replace=False)).astype(np.uint32)
    dist = [500, 1000, 2000]
    d, b, s = H.pair(distances=dist)
    res = ns.centrality_shortest(distances=dist, source_indices=src.tolist(), sample_probability=1.0, pbar_disabled=True)
    elig = np.zeros(f.node_bound, np.uint8)
    elig[src] = 1
    ref, cnt = oracle_mod.OracleGraph(f).centrality_shortest(d, b, s, H.SPEED, sources=src, wt=np.ones(len(src), np.float32),
                                                             eligible=elig, n_threads=8)  # fmt: skip
    assert np.array_equal(res._out[0], ref[0]) and np.array_equal(res._out[2], ref[2])
    np.testing.assert_allclose(res._out, ref, rtol=RTOL, atol=1e-7)
    for key in ("settled", "edge_iters", "sum_ri", "sum_ci"):
        assert res.stats[key] == cnt[key], key
