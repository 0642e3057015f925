"""Synthetic street-network generators for the BASELINE.json configs.

Pure numpy (no shapely): a jittered lattice with 10 % of the edges dropped, optionally decomposed into ≈20 m segments
(config #4) or converted to its dual (config #3).  Every graph goes through ``NetworkStructure.from_arrays`` with a
directed-edge insertion order of the same shape ``io.network_structure_from_nx`` produces (for each start node, each of
its neighbours; both directions of every undirected edge), and lengths (f64 hypot → f32) and angle sums follow the
reference's ingest rules (graph.rs:774, :857).
"""
from __future__ import annotations

import numpy as np

from .rustalgos.graph import NetworkStructure

X0, Y0 = 500000.0, 5700000.0  # UTM-like offsets so the f64 → f32 length conversion is exercised


def lattice(nx_: int, ny_: int, spacing: float = 100.0, jitter: float = 25.0, drop: float = 0.10, seed: int = 42):
    """Jittered lattice: returns (xy float64 [n,2], undirected edges int64 [m,2]).  Every edge is ≥ 30 m."""
    rng = np.random.default_rng(seed)
    gx, gy = np.meshgrid(np.arange(nx_), np.arange(ny_), indexing="xy")
    xy = np.stack([gx.ravel() * spacing, gy.ravel() * spacing], axis=1).astype(np.float64)
    xy += rng.uniform(-jitter, jitter, size=xy.shape)
    xy[:, 0] += X0
    xy[:, 1] += Y0
    idx = np.arange(nx_ * ny_).reshape(ny_, nx_)
    horiz = np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], axis=1)
    vert = np.stack([idx[:-1, :].ravel(), idx[1:, :].ravel()], axis=1)
    edges = np.concatenate([horiz, vert], axis=0)
    keep = np.random.default_rng(seed + 1).random(len(edges)) >= drop
    edges = edges[keep]
    ln = np.hypot(*(xy[edges[:, 0]] - xy[edges[:, 1]]).T)
    edges = edges[ln >= 30.0]
    return xy, edges.astype(np.int64)


def decompose(xy: np.ndarray, edges: np.ndarray, max_len: float = 20.0):
    """Cut every edge into ceil(len / max_len) equal straight pieces (graphs.py:1911-1912).  Original nodes keep their
    indices; chain nodes are appended edge by edge."""
    a, b = xy[edges[:, 0]], xy[edges[:, 1]]
    ln = np.hypot(*(a - b).T)
    pieces = np.ceil(ln / max_len).astype(np.int64)
    n_new = int((pieces - 1).sum())
    n0 = len(xy)
    new_xy = np.empty((n_new, 2), np.float64)
    first_new = n0 + np.concatenate([[0], np.cumsum(pieces - 1)[:-1]])
    # per-piece endpoints
    total = int(pieces.sum())
    eid = np.repeat(np.arange(len(edges)), pieces)
    k = np.arange(total) - np.repeat(np.cumsum(pieces) - pieces, pieces)
    p = pieces[eid]
    start = np.where(k == 0, edges[eid, 0], first_new[eid] + k - 1)
    end = np.where(k == p - 1, edges[eid, 1], first_new[eid] + k)
    inner = k < p - 1
    t = ((k + 1) / p)[inner, None]
    new_xy[(first_new[eid] + k - n0)[inner]] = a[eid[inner]] + (b[eid[inner]] - a[eid[inner]]) * t
    return np.concatenate([xy, new_xy], axis=0), np.stack([start, end], axis=1)


def _directed_in_ingest_order(n: int, edges: np.ndarray):
    """Directed edge list grouped by start node 0..n-1 (the loop shape of io.network_structure_from_nx), neighbours in
    edge-list order as seen from that node."""
    m = len(edges)
    src = np.concatenate([edges[:, 0], edges[:, 1]])
    dst = np.concatenate([edges[:, 1], edges[:, 0]])
    seq = np.concatenate([np.arange(m), np.arange(m)])
    order = np.lexsort((seq, src))
    return src[order], dst[order]


def terrain(xy: np.ndarray, relief: float = 30.0) -> np.ndarray:
    """Smooth synthetic elevation (metres) over the lattice: gradients of a few percent, so the Tobler slope penalty
    (centrality.rs:969-984) makes the two directions of every edge differ."""
    x, y = xy[:, 0] - X0, xy[:, 1] - Y0
    return relief * np.sin(x / 700.0) * np.cos(y / 500.0) + 0.004 * x - 0.002 * y


def primal_network(xy: np.ndarray, edges: np.ndarray, live: np.ndarray | None = None,
                   z: np.ndarray | None = None) -> NetworkStructure:  # fmt: skip
    n = len(xy)
    src, dst = _directed_in_ingest_order(n, edges)
    d = xy[dst] - xy[src]
    length = np.hypot(d[:, 0], d[:, 1]).astype(np.float32)
    return NetworkStructure.from_arrays(
        live=np.ones(n, np.uint8) if live is None else live,
        weight=np.ones(n, np.float32),
        src=src,
        dst=dst,
        edge_idx=np.zeros(len(src), np.uint32),
        length=length,
        z=z,
        x=xy[:, 0],
        y=xy[:, 1],
    )


def _turn_angle(a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """|turn| in degrees at b for a→b→c (graph.rs:326-346), f64."""
    a1 = np.degrees(np.arctan2(a[:, 1] - b[:, 1], a[:, 0] - b[:, 0]))
    a2 = np.degrees(np.arctan2(b[:, 1] - c[:, 1], b[:, 0] - c[:, 0]))
    return np.abs(np.mod(a2 - a1 + 180.0, 360.0) - 180.0)


def dual_network(xy: np.ndarray, edges: np.ndarray, z: np.ndarray | None = None) -> NetworkStructure:
    """Dual of a straight-edged primal graph (graphs.py:2077-2147): dual node per primal edge at its midpoint, dual edge
    per pair of primal edges sharing a node with geometry [mid_a, shared, mid_b]."""
    m = len(edges)
    mid = (xy[edges[:, 0]] + xy[edges[:, 1]]) / 2.0
    # incident primal edges per primal node
    ends = np.concatenate([edges[:, 0], edges[:, 1]])
    eids = np.concatenate([np.arange(m), np.arange(m)])
    order = np.lexsort((eids, ends))
    ends, eids = ends[order], eids[order]
    starts = np.searchsorted(ends, np.arange(len(xy)))
    stops = np.searchsorted(ends, np.arange(len(xy)), side="right")
    pa, pb, shared = [], [], []
    for node in np.nonzero(stops - starts >= 2)[0]:
        inc = eids[starts[node] : stops[node]]
        for i in range(len(inc)):
            for j in range(i + 1, len(inc)):
                pa.append(inc[i])
                pb.append(inc[j])
                shared.append(node)
    pa, pb, shared = np.asarray(pa, np.int64), np.asarray(pb, np.int64), np.asarray(shared, np.int64)
    dedges = np.stack([pa, pb], axis=1)
    # directed in ingest order, carrying the shared node alongside
    k = len(dedges)
    src = np.concatenate([pa, pb])
    dst = np.concatenate([pb, pa])
    sh = np.concatenate([shared, shared])
    seq = np.concatenate([np.arange(k), np.arange(k)])
    order = np.lexsort((seq, src))
    src, dst, sh = src[order], dst[order], sh[order]
    p0, p1, p2 = mid[src], xy[sh], mid[dst]
    length = (np.hypot(*(p1 - p0).T) + np.hypot(*(p2 - p1).T)).astype(np.float32)
    angle = _turn_angle(p0, p1, p2).astype(np.float32)
    return NetworkStructure.from_arrays(
        live=np.ones(m, np.uint8),
        weight=np.ones(m, np.float32),
        src=src,
        dst=dst,
        edge_idx=np.zeros(len(src), np.uint32),
        length=length,
        angle_sum=angle,
        shared_key=sh.astype(np.int32),
        is_dual=True,
        z=None if z is None else (z[edges[:, 0]] + z[edges[:, 1]]) / 2.0,
        x=mid[:, 0],
        y=mid[:, 1],
    )


def config(name: str, scale: float = 1.0, hilly: bool = False):
    """Named BASELINE.json workloads → (NetworkStructure, description dict).  ``scale`` < 1 shrinks the lattice side;
    ``hilly`` gives every node an elevation (slope-penalised, direction-dependent edge seconds)."""
    if name == "cfg2":  # 100k-node perturbed grid, primal
        side = max(8, int(316 * scale))
        xy, e = lattice(side, side, seed=42)
        return primal_network(xy, e, z=terrain(xy) if hilly else None), {"workload": "cfg2-100k-primal-grid", "lattice": side}
    if name == "cfg3":  # 100k-node dual
        side = max(8, int(236 * scale))
        xy, e = lattice(side, side, seed=42)
        return dual_network(xy, e, z=terrain(xy) if hilly else None), {"workload": "cfg3-100k-dual", "lattice": side}
    if name == "cfg4":  # 1M-node decomposed (20 m segments)
        side = max(8, int(333 * scale))
        xy, e = lattice(side, side, seed=42)
        xy2, e2 = decompose(xy, e, 20.0)
        return primal_network(xy2, e2, z=terrain(xy2) if hilly else None), {"workload": "cfg4-1M-decomposed-20m", "lattice": side}
    if name == "cfg5":  # 4M-node metro
        side = max(8, int(2000 * scale))
        xy, e = lattice(side, side, seed=42)
        return primal_network(xy, e), {"workload": "cfg5-4M-metro", "lattice": side}
    raise ValueError(f"unknown config {name}")
