"""Source sampling through the GPU path: the behaviours the reference pins in tests/test_sampling.py —
seed reproducibility, p = 1 equals the exact run, inverse-probability weighting is unbiased on average, sampling
weights are validated and scale the inclusion probability, zero-weight nodes are never sources, the betweenness and
angular paths sample the same way, tolerance 0 means the default epsilon.  The random stream itself is numpy's PCG64,
not rand::StdRng (unpinned upstream), so only stream-independent properties are asserted."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
D = 500


@pytest.fixture(scope="module")
def primal():
    _g, nodes, _e, ns = H.primal_ns()
    return ns, nodes


@pytest.fixture(scope="module")
def dual():
    _g, nodes, _e, ns = H.dual_ns()
    return ns, nodes


def density(ns, **kw):
    r = ns.centrality_shortest(compute_closeness=True, compute_betweenness=False, distances=[D], pbar_disabled=True, **kw)
    return np.array(r.node_density[D])


def test_same_seed_same_result_other_seed_other_result(primal):
    ns, _ = primal
    a = density(ns, sample_probability=0.3, random_seed=42)
    assert np.array_equal(a, density(ns, sample_probability=0.3, random_seed=42))
    assert not np.allclose(a, density(ns, sample_probability=0.3, random_seed=43))


def test_probability_one_is_the_exact_run(primal):
    ns, _ = primal
    assert np.array_equal(density(ns), density(ns, sample_probability=1.0, random_seed=42))
    full = ns.centrality_shortest(compute_closeness=False, distances=[D], pbar_disabled=True)
    samp = ns.centrality_shortest(compute_closeness=False, distances=[D], sample_probability=1.0, random_seed=7, pbar_disabled=True)
    np.testing.assert_allclose(full.node_betweenness[D], samp.node_betweenness[D], rtol=1e-12)
    np.testing.assert_allclose(full.node_betweenness_beta[D], samp.node_betweenness_beta[D], rtol=1e-12)
    assert samp.sampled_source_count == ns.street_node_count()


def test_inverse_probability_weighting_is_unbiased_on_average(primal):
    ns, _ = primal
    full = density(ns)
    mask = full > 0
    avg = np.mean([density(ns, sample_probability=0.5, random_seed=s) for s in range(24)], axis=0)
    assert np.mean(np.abs(avg[mask] - full[mask]) / full[mask]) < 0.15
    for prob in (0.3, 0.7):
        means = []
        for s in range(24):
            d_ = density(ns, sample_probability=prob, random_seed=s)
            m = (d_ > 0) & mask
            means.append(d_[m].mean())
        assert abs(np.mean(means) - full[mask].mean()) / full[mask].mean() < 0.25


def test_sampling_weights_are_validated(primal):
    ns, nodes = primal
    n = len(nodes)
    for bad, msg in (([1.5] + [1.0] * (n - 1), "out of range"), ([-0.1] + [1.0] * (n - 1), "out of range"),
                     ([1.0] * (n - 1), "must match node_count")):  # fmt: skip
        with pytest.raises(ValueError, match=msg):
            density(ns, sample_probability=0.5, sampling_weights=bad)
    with pytest.raises(ValueError, match="mutually exclusive"):
        density(ns, source_indices=[0, 1], sampling_weights=[1.0] * n)
    with pytest.raises(ValueError, match=r"\(0.0, 1.0\]"):
        density(ns, sample_probability=0.0)


def test_zero_weight_nodes_are_never_sources(primal):
    ns, nodes = primal
    n = len(nodes)
    w = [1.0 if i < n // 2 else 0.0 for i in range(n)]
    r = ns.centrality_shortest(compute_betweenness=False, distances=[D], sample_probability=1.0, sampling_weights=w,
                               pbar_disabled=True)  # fmt: skip
    assert r.sampled_source_count == n // 2
    # targets still aggregate from the remaining sources, and equal the explicit-source run over the first half
    ref = ns.centrality_shortest(compute_betweenness=False, distances=[D], source_indices=list(range(n // 2)),
                                 sample_probability=1.0, pbar_disabled=True)  # fmt: skip
    assert np.array_equal(np.array(r.node_density[D]), np.array(ref.node_density[D]))
    assert np.count_nonzero(np.array(r.node_density[D])) > 0


def test_uniform_weight_scales_the_inclusion_probability(primal):
    ns, nodes = primal
    n = len(nodes)
    # weight 0.5 at p = 1 samples like p = 0.5 without weights: same draws, same inclusion test, same IPW
    for seed in range(3):
        a = density(ns, sample_probability=1.0, sampling_weights=[0.5] * n, random_seed=seed)
        b = density(ns, sample_probability=0.5, random_seed=seed)
        assert np.array_equal(a, b)


def test_betweenness_sampling_reproducible_and_converging(primal):
    ns, _ = primal
    kw = dict(compute_closeness=False, compute_betweenness=True, distances=[D], pbar_disabled=True)
    a = ns.centrality_shortest(sample_probability=0.3, random_seed=42, **kw)
    b = ns.centrality_shortest(sample_probability=0.3, random_seed=42, **kw)
    np.testing.assert_allclose(a.node_betweenness[D], b.node_betweenness[D], rtol=1e-12)
    full = np.array(ns.centrality_shortest(**kw).node_betweenness[D])
    avg = np.mean([np.array(ns.centrality_shortest(sample_probability=0.5, random_seed=s, **kw).node_betweenness[D])
                   for s in range(24)], axis=0)  # fmt: skip
    top = full >= np.percentile(full, 60)
    assert np.corrcoef(avg[top], full[top])[0, 1] > 0.9


def test_simplest_sampling_reproducible_and_converging(dual):
    ns, _ = dual
    kw = dict(compute_closeness=True, compute_betweenness=False, distances=[D], pbar_disabled=True)
    a = ns.centrality_simplest(sample_probability=0.4, random_seed=11, **kw)
    b = ns.centrality_simplest(sample_probability=0.4, random_seed=11, **kw)
    assert np.array_equal(np.array(a.node_density[D]), np.array(b.node_density[D]))
    full = np.array(ns.centrality_simplest(**kw).node_density[D])
    avg = np.mean([np.array(ns.centrality_simplest(sample_probability=0.5, random_seed=s, **kw).node_density[D])
                   for s in range(24)], axis=0)  # fmt: skip
    mask = full > 0
    assert np.mean(np.abs(avg[mask] - full[mask]) / full[mask]) < 0.2


def test_tolerance_zero_is_the_default_and_tolerance_spreads_betweenness(primal, dual):
    ns, _ = primal
    kw = dict(compute_closeness=False, distances=[D], pbar_disabled=True)
    base = np.array(ns.centrality_shortest(**kw).node_betweenness[D])
    np.testing.assert_allclose(np.array(ns.centrality_shortest(tolerance=0.0, **kw).node_betweenness[D]), base, rtol=1e-12)
    assert not np.allclose(np.array(ns.centrality_shortest(tolerance=10.0, **kw).node_betweenness[D]), base)
    nd, _ = dual
    base_a = np.array(nd.centrality_simplest(**kw).node_betweenness[D])
    np.testing.assert_allclose(np.array(nd.centrality_simplest(tolerance=0.0, **kw).node_betweenness[D]), base_a, rtol=1e-12)
    assert not np.allclose(np.array(nd.centrality_simplest(tolerance=10.0, **kw).node_betweenness[D]), base_a)
