"""Shared graph builders for the test-suite (fixtures of /root/reference/tests/conftest.py restated without shapely)."""
from __future__ import annotations

import networkx as nx
import numpy as np

from cityseer_b200 import rustalgos
from cityseer_b200.tools import graphs, io, mock

SPEED = 1.33333


def primal_ns():
    g = graphs.nx_simple_geoms(mock.mock_graph())
    return (g,) + io.network_structure_from_nx(g)


def dual_ns():
    g = graphs.nx_to_dual(graphs.nx_simple_geoms(mock.mock_graph()))
    return (g,) + io.network_structure_from_nx(g)


def diamond_ns(dual=False):
    g = graphs.nx_simple_geoms(mock.diamond_graph())
    if dual:
        g = graphs.nx_to_dual(g)
    return (g,) + io.network_structure_from_nx(g)


def graph_from_coords(coords: dict, edges, z: dict | None = None):
    g = nx.MultiGraph()
    g.graph["crs"] = 32630
    for k, (x, y) in coords.items():
        attrs = {"x": x, "y": y}
        if z and k in z and z[k] is not None:
            attrs["z"] = z[k]
        g.add_node(k, **attrs)
    for a, b in edges:
        g.add_edge(a, b)
    return graphs.nx_simple_geoms(g)


def plateau_graph():
    # /root/reference/tests/rustalgos/test_centrality.py:70-97
    coords = {"A": (0.0, 0.0), "B": (100.0, 0.0), "C": (200.0, 0.0), "D": (300.0, 0.0), "E": (400.0, 0.0),
              "BU": (100.0, 100.0), "BD": (100.0, -100.0), "CU": (200.0, 100.0)}  # fmt: skip
    edges = [("A", "B"), ("B", "C"), ("C", "D"), ("D", "E"), ("B", "BU"), ("B", "BD"), ("C", "CU")]
    return graph_from_coords(coords, edges)


def tolerance_drift_graph():
    # /root/reference/tests/rustalgos/test_centrality.py:100-126
    coords = {"S": (0.0, 0.0), "T": (8.0, 0.0), "A": (4.0, 3.0), "B": (4.0, 2.831960451701259),
              "C": (4.0, 2.0615528128088303)}  # fmt: skip
    edges = [("S", "A"), ("A", "T"), ("S", "B"), ("B", "T"), ("S", "C"), ("C", "T")]
    return graph_from_coords(coords, edges)


def pair(distances=None, betas=None, minutes=None, speed=SPEED):
    return rustalgos.pair_distances_betas_time(speed, distances, betas, minutes)


def compact(out, f):
    """[M][D][node_bound] -> [M][D][node_count] over node_indices."""
    return out[:, :, f.node_indices]


def nx_length_graph(g):
    h = nx.MultiGraph()
    for n, d in g.nodes(data=True):
        h.add_node(n, **d)
    for s, e, k, d in g.edges(keys=True, data=True):
        h.add_edge(s, e, key=k, length=float(np.float32(graphs.coords_length(d["geom"]))))
    return h
