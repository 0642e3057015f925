"""NetworkX → NetworkStructure ingest with the call sequence of ``io.network_structure_from_nx``
(/root/reference/pysrc/cityseer/tools/io.py:1018-1210): nodes in graph order, then for each start node, each neighbour,
each parallel key one ``add_street_edge`` with the start-aligned geometry — so every undirected edge is stored as two
directed edges and adjacency order matches the reference.  Returns plain pandas frames (geopandas is not available)."""
from __future__ import annotations

import pandas as pd

from ..rustalgos.graph import NetworkStructure
from .graphs import align_coords, coords_wkt


def network_structure_from_nx(g):
    ns = NetworkStructure()
    ns.set_is_dual(bool(g.graph.get("is_dual", False)))
    node_rows = {}
    for key, data in g.nodes(data=True):
        if not isinstance(key, str):
            raise TypeError(f"Node key must be of type string but encountered {type(key)}")
        x, y = float(data["x"]), float(data["y"])
        z = data.get("z")
        live = bool(data["live"]) if "live" in data else True
        weight = data.get("weight", 1)
        idx = ns.add_street_node(key, x, y, live, weight, z=z)
        node_rows[key] = (idx, x, y, z, live, weight)
    edge_rows = {}
    for s in g.nodes():
        s_idx, sx, sy = node_rows[s][0], node_rows[s][1], node_rows[s][2]
        for e in g.neighbors(s):
            e_idx = node_rows[e][0]
            for k, data in g[s][e].items():
                geom = data.get("geom")
                if geom is None:
                    raise ValueError(f"Edge has no geometry: start_node={s}, end_node={e}, edge_idx={k}")
                coords = align_coords(geom, (sx, sy))
                ns_e = ns.add_street_edge(
                    s_idx, e_idx, int(k), s, e, coords_wkt(coords), float(data.get("imp_factor", 1.0)),
                    shared_primal_node_key=data.get("primal_node_id"),
                )  # fmt: skip
                edge_rows[f"{s}-{e}-{k}"] = (ns_e, s_idx, e_idx, k, s, e, data.get("imp_factor", 1.0))
    nodes_df = pd.DataFrame.from_dict(node_rows, orient="index", columns=["ns_node_idx", "x", "y", "z", "live", "weight"])
    edges_df = pd.DataFrame.from_dict(
        edge_rows, orient="index",
        columns=["ns_edge_idx", "start_ns_node_idx", "end_ns_node_idx", "edge_idx", "nx_start_node_key", "nx_end_node_key", "imp_factor"],
    )  # fmt: skip
    ns.validate()
    ns.build_edge_rtree()
    return nodes_df, edges_df, ns
