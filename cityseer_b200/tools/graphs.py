"""Shapely-free restatements of the graph-preparation steps that define the kernels' inputs:
``nx_simple_geoms`` (graphs.py:58-107), ``nx_to_dual`` (graphs.py:1965-2147) and ``nx_decompose`` (graphs.py:1829-1962)
of /root/reference/pysrc/cityseer/tools/graphs.py.  Edge geometry is a list of (x, y) tuples under ``geom``."""
from __future__ import annotations

import math

import networkx as nx
import numpy as np


def coords_length(coords) -> float:
    return sum(math.hypot(coords[i + 1][0] - coords[i][0], coords[i + 1][1] - coords[i][1]) for i in range(len(coords) - 1))


def coords_wkt(coords) -> str:
    return "LINESTRING (" + ", ".join(f"{float(x)!r} {float(y)!r}" for x, y in coords) + ")"


def align_coords(coords, xy, tolerance: float = 0.5):
    """util.align_linestring_coords (util.py:264-312): orient ``coords`` to start at ``xy``."""
    coords = [tuple(c[:2]) for c in coords]
    a = math.hypot(coords[0][0] - xy[0], coords[0][1] - xy[1])
    b = math.hypot(coords[-1][0] - xy[0], coords[-1][1] - xy[1])
    if a > b:
        coords = coords[::-1]
    tol = math.hypot(coords[0][0] - xy[0], coords[0][1] - xy[1])
    if tol > tolerance:
        raise ValueError(f"Closest side of edge geom {coords} is {tol} from node {xy}, exceeding tolerance of {tolerance}.")
    return coords


def substring(coords, start_frac: float, end_frac: float):
    """Sub-line between two normalised positions (shapely.ops.substring, normalized=True)."""
    total = coords_length(coords)
    s, e = start_frac * total, end_frac * total
    out = []
    acc = 0.0
    for i in range(len(coords) - 1):
        (x0, y0), (x1, y1) = coords[i], coords[i + 1]
        seg = math.hypot(x1 - x0, y1 - y0)
        nxt = acc + seg
        if seg > 0:
            if acc <= s <= nxt and not out:
                t = (s - acc) / seg
                out.append((x0 + (x1 - x0) * t, y0 + (y1 - y0) * t))
            if out and s < nxt < e:
                out.append((x1, y1))
            if out and acc <= e <= nxt:
                t = (e - acc) / seg
                p = (x0 + (x1 - x0) * t, y0 + (y1 - y0) * t)
                if p != out[-1]:
                    out.append(p)
                break
        acc = nxt
    return out


def nx_simple_geoms(g: nx.MultiGraph) -> nx.MultiGraph:
    g = g.copy()
    remove = []
    for s, e, k in g.edges(keys=True):
        a = (float(g.nodes[s]["x"]), float(g.nodes[s]["y"]))
        b = (float(g.nodes[e]["x"]), float(g.nodes[e]["y"]))
        if a == b:
            remove.append((s, e, k))
        else:
            g[s][e][k]["geom"] = [a, b]
    for s, e, k in remove:
        g.remove_edge(s, e, key=k)
    return g


def _dual_key(a, b, k) -> str:
    s = sorted([str(a), str(b)])
    return f"{s[0]}_{s[1]}_k{k}"


def nx_to_dual(g: nx.MultiGraph) -> nx.MultiGraph:
    """Primal → dual: one dual node per primal edge (at its midpoint), one dual edge per pair of primal edges sharing a
    node, geometry = the two welded half-geoms, ``primal_node_id`` = the shared primal node (graphs.py:2077-2147)."""
    d = nx.MultiGraph()
    d.graph["crs"] = g.graph.get("crs")
    d.graph["is_dual"] = True

    def half_geoms(a, b, k):
        axy = (float(g.nodes[a]["x"]), float(g.nodes[a]["y"]))
        bxy = (float(g.nodes[b]["x"]), float(g.nodes[b]["y"]))
        coords = align_coords(g[a][b][k]["geom"], axy)
        ah = substring(coords, 0.0, 0.5)
        bh = substring(coords, 0.5, 1.0)
        ah[0] = axy
        mid = ah[-1]
        bh[0] = mid
        bh[-1] = bxy
        return ah, bh

    for s, e, k, data in g.edges(keys=True, data=True):
        coords = data["geom"]
        mid = substring(coords, 0.0, 0.5)[-1]
        key = _dual_key(s, e, k)
        d.add_node(key, x=mid[0], y=mid[1], primal_edge_node_a=s, primal_edge_node_b=e, primal_edge_idx=k)
        if "live" in g.nodes[s] and "live" in g.nodes[e]:
            d.nodes[key]["live"] = bool(g.nodes[s]["live"] or g.nodes[e]["live"])
    for s, e, k in g.edges(keys=True):
        hub = _dual_key(s, e, k)
        s_half, e_half = half_geoms(s, e, k)
        for n_side, m_side, half in ((s, e, s_half), (e, s, e_half)):
            for nb in nx.neighbors(g, n_side):
                if nb == m_side:
                    continue
                for k2 in g[n_side][nb]:
                    spoke = _dual_key(n_side, nb, k2)
                    if d.has_edge(hub, spoke):
                        continue
                    spoke_half, _ = half_geoms(n_side, nb, k2)
                    # weld: hub midpoint -> shared node -> spoke midpoint
                    hub_part = half if half[-1] == spoke_half[0] else half[::-1]
                    merged = list(hub_part) + list(spoke_half[1:])
                    d.add_edge(hub, spoke, primal_node_id=n_side, geom=merged)
    return d


def nx_decompose(g: nx.MultiGraph, decompose_max: float) -> nx.MultiGraph:
    """Split every edge into ``ceil(length / decompose_max)`` equal pieces (graphs.py:1911-1960); new nodes are keyed
    ``{start}_{n}_{end}``; the original nodes keep their keys and come first."""
    out = nx.MultiGraph()
    out.graph.update(g.graph)
    for nd, data in g.nodes(data=True):
        out.add_node(nd, **data)
    for s, e, k, data in g.edges(keys=True, data=True):
        sxy = (float(g.nodes[s]["x"]), float(g.nodes[s]["y"]))
        coords = align_coords(data["geom"], sxy)
        total = coords_length(coords)
        n = int(np.ceil(total / decompose_max))
        prev = s
        for i in range(n):
            sub = substring(coords, i / n, (i + 1) / n)
            if i == n - 1:
                nxt = e
            else:
                nxt = f"{s}_{i}_{e}_k{k}"
                out.add_node(nxt, x=sub[-1][0], y=sub[-1][1])
                for attr in ("live", "weight"):
                    if attr in g.nodes[s] and attr in g.nodes[e]:
                        out.nodes[nxt][attr] = g.nodes[s][attr] if attr == "weight" else (g.nodes[s][attr] or g.nodes[e][attr])
            out.add_edge(prev, nxt, geom=sub)
            prev = nxt
    return out
