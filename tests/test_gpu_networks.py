"""Wrapper parity (reference: tests/metrics/test_networks.py:11-195): the ``metrics.networks`` functions write exactly
what the direct NetworkStructure calls return, under the reference's column names."""
import numpy as np
import pytest

import helpers as H
from cityseer_b200 import config
from cityseer_b200.metrics import networks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("flags", [(True, True), (True, False), (False, True)])
def test_node_centrality_shortest_wrapper(flags):
    _g, nodes, _e, ns = H.primal_ns()
    distances = [400, 800]
    out = networks.node_centrality_shortest(ns, nodes.copy(), distances=distances, compute_closeness=flags[0], compute_betweenness=flags[1])
    direct = ns.centrality_shortest(distances=distances, compute_closeness=flags[0], compute_betweenness=flags[1], pbar_disabled=True)
    for d in distances:
        if flags[0]:
            for key, attr in [("beta", "node_beta"), ("cycles", "node_cycles"), ("density", "node_density"),
                              ("farness", "node_farness"), ("harmonic", "node_harmonic")]:  # fmt: skip
                assert np.allclose(out[config.prep_gdf_key(key, d)], getattr(direct, attr)[d], rtol=1e-9, equal_nan=True)
            hill = direct.node_density[d] ** 2 / direct.node_farness[d]
            assert np.allclose(out[config.prep_gdf_key("hillier", d)], hill, rtol=1e-9, equal_nan=True)
        else:
            assert config.prep_gdf_key("density", d) not in out.columns
        if flags[1]:
            assert np.allclose(out[f"cc_betweenness_{d}"], direct.node_betweenness[d], rtol=1e-9)
            assert np.allclose(out[f"cc_betweenness_beta_{d}"], direct.node_betweenness_beta[d], rtol=1e-9)


def test_betas_and_minutes_inputs_key_by_distance():
    _g, nodes, _e, ns = H.primal_ns()
    out = networks.node_centrality_shortest(ns, nodes.copy(), betas=[0.01, 0.005], compute_betweenness=False)
    assert "cc_density_400" in out.columns and "cc_density_800" in out.columns
    out = networks.node_centrality_shortest(ns, nodes.copy(), minutes=[5.0], compute_betweenness=False)
    assert "cc_density_400" in out.columns


def test_simplest_and_segment_wrappers():
    _g, nodes, _e, ns = H.dual_ns()
    out = networks.node_centrality_simplest(ns, nodes.copy(), distances=[800])
    direct = ns.centrality_simplest(distances=[800], angular_scaling_unit=90, farness_scaling_offset=1, pbar_disabled=True)
    assert np.allclose(out["cc_harmonic_800_ang"], direct.node_harmonic[800], rtol=1e-9)
    assert np.allclose(out["cc_betweenness_800_ang"], direct.node_betweenness[800], rtol=1e-9)
    assert {"cc_density_800_ang", "cc_farness_800_ang", "cc_hillier_800_ang"} <= set(out.columns)
    _g, nodes_p, _e, ns_p = H.primal_ns()
    with pytest.raises(ValueError, match="dual graph"):
        networks.node_centrality_simplest(ns_p, nodes_p.copy(), distances=[800])
    seg = networks.segment_centrality(ns_p, nodes_p.copy(), distances=[400])
    direct = ns_p.segment_centrality(distances=[400], pbar_disabled=True)
    for key, attr in [("seg_density", "segment_density"), ("seg_harmonic", "segment_harmonic"),
                      ("seg_beta", "segment_beta"), ("seg_betweenness", "segment_betweenness")]:  # fmt: skip
        assert np.allclose(seg[config.prep_gdf_key(key, 400)], getattr(direct, attr)[400], rtol=1e-9)


def test_sampled_mode_is_seed_deterministic():
    # tests/metrics/test_networks.py:198-244
    _g, nodes, _e, ns = H.primal_ns()
    a = networks.node_centrality_shortest(ns, nodes.copy(), distances=[2000], sample=True, random_seed=3, epsilon=0.2)
    b = networks.node_centrality_shortest(ns, nodes.copy(), distances=[2000], sample=True, random_seed=3, epsilon=0.2)
    assert np.allclose(a["cc_density_2000"], b["cc_density_2000"])
