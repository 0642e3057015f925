"""Host-side ``NetworkStructure`` container — drop-in for ``cityseer.rustalgos.graph.NetworkStructure`` on the
centrality hot path (stub: /root/reference/pysrc/cityseer/rustalgos/graph.pyi:85-674; Rust: rust/src/graph.rs).

The container keeps the node / edge payload fields the kernels read (graph.rs:25-38, :89-115) in plain Python lists
with petgraph ``StableGraph`` index semantics (stable indices, LIFO free-lists, newest-first adjacency), freezes them
into flat arrays on the first compute call, and hands those arrays to the CUDA library through the C ABI
(``include/cityseer_b200.h``).  Compute methods never run on the CPU: without the CUDA library / a GPU they raise.
"""
from __future__ import annotations

import math
import re
import threading
from array import array
from typing import Any

import numpy as np

from .. import _native
from . import centrality as _centrality

_NUM_RE = re.compile(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?|[-+]?(?:inf|nan)", re.IGNORECASE)


class NodeVisit:
    """graph.rs:252-284"""

    __slots__ = ("visited", "discovered", "pred", "short_dist", "simpl_dist", "origin_seg", "last_seg", "agg_seconds")

    def __init__(self):
        self.visited = False
        self.discovered = False
        self.pred = None
        self.short_dist = math.inf
        self.simpl_dist = math.inf
        self.origin_seg = None
        self.last_seg = None
        self.agg_seconds = math.inf


class EdgeVisit:
    """graph.rs:287-315"""

    __slots__ = ("visited", "start_nd_idx", "end_nd_idx", "edge_idx")

    def __init__(self):
        self.visited = False
        self.start_nd_idx = None
        self.end_nd_idx = None
        self.edge_idx = None


class NodePayload:
    """graph.rs:25-38 (street nodes only; transport nodes are out of scope)."""

    __slots__ = ("node_key", "x", "y", "z", "live", "weight", "is_transport")

    def __init__(self, node_key, x, y, z, live, weight):
        self.node_key = node_key
        self.x = x
        self.y = y
        self.z = z
        self.live = live
        self.weight = weight
        self.is_transport = False

    @property
    def coord(self):
        return (self.x, self.y)

    @property
    def coord_z(self):
        return (self.x, self.y, self.z)

    def validate(self):
        if not math.isfinite(self.weight) or self.weight < 0.0:
            raise ValueError(
                "Invalid street node payload: weight must be finite and non-negative (>= 0.0). "
                f"Found {self.weight}. Node key: {self.node_key!r}"
            )


class EdgePayload:
    """graph.rs:89-115 (street edges, and transport edges defined by their travel time)."""

    __slots__ = (
        "start_nd_key_py", "end_nd_key_py", "shared_primal_node_key", "edge_idx", "length", "angle_sum", "imp_factor",
        "in_bearing", "out_bearing", "seconds", "geom_wkt", "is_transport", "_src", "_dst", "_stamp",
    )  # fmt: skip

    def validate(self):
        if not math.isfinite(self.imp_factor) or self.imp_factor <= 0.0:
            raise ValueError(
                f"Invalid edge payload : imp_factor must be finite and positive (> 0.0). Found {self.imp_factor}. "
                f"Start key: {self.start_nd_key_py!r}, End key: {self.end_nd_key_py!r}"
            )
        if getattr(self, "is_transport", False) and not math.isnan(self.seconds):
            # graph.rs:152-160
            if not math.isfinite(self.seconds) or self.seconds < 0.0:
                raise ValueError(
                    "Invalid transport edge payload : seconds must be finite and non-negative. "
                    f"Start key: {self.start_nd_key_py!r}, End key: {self.end_nd_key_py!r}"
                )
            return
        if not math.isfinite(self.length):
            raise ValueError(f"Invalid street edge payload : length must be finite. Found {self.length}.")
        if not math.isfinite(self.angle_sum):
            raise ValueError(f"Invalid street edge payload : angle_sum must be finite. Found {self.angle_sum}.")


def parse_linestring_wkt(geom_wkt: str) -> list[tuple[float, float]]:
    """Minimal WKT ``LINESTRING`` reader (2D or Z; Z is dropped, as geo's ``LineString<f64>`` is 2D)."""
    s = geom_wkt.strip()
    head = s[:12].upper()
    if not head.startswith("LINESTRING"):
        raise ValueError(f"expected LINESTRING, found: {s[:40]}")
    lp = s.find("(")
    rp = s.rfind(")")
    if lp < 0 or rp < lp:
        if "EMPTY" in s.upper():
            return []
        raise ValueError("malformed WKT")
    has_z = " Z" in s[:lp].upper() or "Z(" in s[: lp + 1].upper()
    coords = []
    for part in s[lp + 1 : rp].split(","):
        nums = _NUM_RE.findall(part)
        if len(nums) < 2:
            raise ValueError(f"malformed coordinate '{part.strip()}'")
        if not has_z and len(nums) > 3:
            raise ValueError(f"malformed coordinate '{part.strip()}'")
        coords.append((float(nums[0]), float(nums[1])))
    return coords


def _bearing(ax, ay, bx, by) -> float:
    # graph.rs:326-332
    if ax == bx and ay == by:
        return 0.0
    return math.degrees(math.atan2(by - ay, bx - ax))


def _coords_angle(a, b, c) -> float:
    # graph.rs:336-346
    if a == b or b == c:
        return 0.0
    a1 = _bearing(b[0], b[1], a[0], a[1])
    a2 = _bearing(c[0], c[1], b[0], b[1])
    diff = a2 - a1
    # rem_euclid(360)
    norm = math.fmod(diff + 180.0, 360.0)
    if norm < 0.0:
        norm += 360.0
    return abs(norm - 180.0)


def linestring_metrics(coords: list[tuple[float, float]]) -> tuple[float, float, float, float]:
    """(length f32, angle_sum f32, in_bearing f32, out_bearing f32) per graph.rs:774-857."""
    length = 0.0
    for i in range(len(coords) - 1):
        length += math.hypot(coords[i + 1][0] - coords[i][0], coords[i + 1][1] - coords[i][1])
    angle_sum = 0.0
    for i in range(1, len(coords) - 1):
        angle_sum += _coords_angle(coords[i - 1], coords[i], coords[i + 1])
    in_b = math.nan
    first = coords[0]
    for i in range(1, len(coords)):
        if coords[i] != first:
            in_b = _bearing(first[0], first[1], coords[i][0], coords[i][1])
            break
    if math.isnan(in_b):
        in_b = 0.0
    out_b = math.nan
    last = coords[-1]
    for i in range(len(coords) - 2, -1, -1):
        if coords[i] != last:
            out_b = _bearing(coords[i][0], coords[i][1], last[0], last[1])
            break
    if math.isnan(out_b):
        out_b = 0.0
    f32 = np.float32
    return float(f32(length)), float(f32(angle_sum)), float(f32(in_b)), float(f32(out_b))


class FrozenGraph:
    """Flat arrays handed to the C ABI (and, in tests, to the oracle). Indexed by petgraph node / edge index."""

    __slots__ = (
        "node_bound", "node_exists", "live", "weight", "z", "edge_bound", "edge_exists", "src", "dst", "edge_idx",
        "length", "angle_sum", "imp", "seconds", "shared_key", "stamp", "is_dual", "node_indices", "key_names",
        "xs", "ys", "_plan",
    )  # fmt: skip


class NetworkStructure:
    """Drop-in for ``cityseer.rustalgos.graph.NetworkStructure`` (centrality surface)."""

    def __init__(self):
        self._nodes: list[NodePayload | None] = []
        self._free_nodes: list[int] = []
        self._edges: list[EdgePayload | None] = []
        self._free_edges: list[int] = []
        self._out: list[list[int]] = []  # per node: edge ids, oldest first (iterate reversed = petgraph order)
        self._in: list[list[int]] = []
        self._stamp = 0
        # columnar mirror of the payload fields the kernels read, maintained at mutation time: frozen() is a handful of
        # list -> ndarray conversions instead of a Python loop over every node and edge
        self._ncol = {k: array(t) for k, t in (("x", "d"), ("y", "d"), ("z", "d"), ("live", "B"), ("weight", "f"), ("exists", "B"))}
        self._ecol = {k: array(t) for k, t in (("src", "I"), ("dst", "I"), ("edge_idx", "I"), ("length", "f"), ("angle_sum", "f"),
                                               ("imp", "f"), ("seconds", "f"), ("stamp", "Q"), ("key", "i"), ("exists", "B"))}  # fmt: skip
        self._key_ids: dict[str, int] = {}
        self._is_dual = False
        self._progress = _native.ProgressCounter()
        self._frozen: FrozenGraph | None = None
        self._bulk: FrozenGraph | None = None  # set by from_arrays(); immutable bulk-ingested graph
        self._bulk_keys = None
        self._keys_cache = None
        self._device = None  # _native.DeviceGraph
        self._lock = threading.Lock()

    # ------------------------------------------------------------------ basic state
    def progress_init(self) -> None:
        self._progress.set(0)

    def progress(self) -> int:
        return self._progress.get()

    @property
    def is_dual(self) -> bool:
        return self._is_dual

    def set_is_dual(self, is_dual: bool) -> None:
        self._is_dual = bool(is_dual)
        self._invalidate()

    def _invalidate(self) -> None:
        if self._bulk is not None:
            raise ValueError("NetworkStructure built with from_arrays() is immutable.")
        self._frozen = None
        self._keys_cache = None
        if self._device is not None:
            self._device.close()
            self._device = None

    # ------------------------------------------------------------------ nodes
    def add_street_node(self, node_key, x: float, y: float, live: bool, weight: float, z: float | None = None) -> int:
        """graph.rs:427-452"""
        payload = NodePayload(node_key, float(x), float(y), None if z is None else float(z), bool(live),
                              float(np.float32(weight)))  # fmt: skip
        payload.validate()
        self._invalidate()
        return self._add_node_internal(payload)

    def _add_node_internal(self, payload: NodePayload) -> int:
        c = self._ncol
        row = (payload.x, payload.y, math.nan if payload.z is None else payload.z, 1 if payload.live else 0,
               payload.weight, 1)  # fmt: skip
        if self._free_nodes:
            idx = self._free_nodes.pop()
            self._nodes[idx] = payload
            for k, v in zip(("x", "y", "z", "live", "weight", "exists"), row):
                c[k][idx] = v
        else:
            idx = len(self._nodes)
            self._nodes.append(payload)
            self._out.append([])
            self._in.append([])
            for k, v in zip(("x", "y", "z", "live", "weight", "exists"), row):
                c[k].append(v)
        return idx

    def add_transport_node(self, *args, **kwargs):
        """graph.rs:455-555 links the new node to nearby street nodes through the edge R-tree and the data-assignment
        search (graph.rs:1064-1232), which is outside this build's scope (land-use assignment).  Add the
        stop with ``add_street_node(..., live=False, weight=0)`` and connect it with ``add_transport_edge`` instead."""
        raise NotImplementedError(
            "add_transport_node needs the edge R-tree / data-assignment search (out of scope); add the stop as a "
            "non-live street node and connect it with add_transport_edge"
        )

    def add_transport_edge(self, start_nd_idx: int, end_nd_idx: int, edge_idx: int, start_nd_key_py, end_nd_key_py,
                           seconds: float, imp_factor: float | None = None) -> int:  # fmt: skip
        """graph.rs:946-985 — an abstract edge defined by its travel time: ``seconds`` is what
        ``edge_travel_seconds`` returns for it at any speed (centrality.rs:988-990); length / angle are NaN."""
        sec = float(np.float32(seconds))
        if not math.isfinite(sec) or sec < 0.0:
            raise ValueError(
                f"Invalid seconds value ({seconds}) for transport edge (idx {edge_idx}) between nodes {start_nd_idx} "
                f"and {end_nd_idx}."
            )
        p = EdgePayload()
        p.start_nd_key_py = start_nd_key_py
        p.end_nd_key_py = end_nd_key_py
        p.shared_primal_node_key = None
        p.edge_idx = int(edge_idx)
        p.length = math.nan
        p.angle_sum = math.nan
        p.imp_factor = 1.0 if imp_factor is None else float(np.float32(imp_factor))
        p.in_bearing = math.nan
        p.out_bearing = math.nan
        p.seconds = sec
        p.geom_wkt = None
        p.is_transport = True
        return self._add_edge_internal(start_nd_idx, end_nd_idx, p)

    def _require_node(self, node_idx: int, param: str) -> NodePayload:
        # graph.rs:1248-1258
        if self._bulk is not None:
            raise ValueError("per-node payload access is not available on a from_arrays() graph")
        if node_idx < 0 or node_idx >= len(self._nodes) or self._nodes[node_idx] is None:
            raise ValueError(f"{param} {node_idx} does not exist in the graph.")
        return self._nodes[node_idx]  # type: ignore[return-value]

    def get_node_payload_py(self, node_idx: int) -> NodePayload:
        return self._require_node(node_idx, "node_idx")

    def get_node_weight(self, node_idx: int) -> float:
        if self._bulk is not None:
            self._check_bulk_node(node_idx)
            return float(self._bulk.weight[node_idx])
        return self._require_node(node_idx, "node_idx").weight

    def is_node_live(self, node_idx: int) -> bool:
        if self._bulk is not None:
            self._check_bulk_node(node_idx)
            return bool(self._bulk.live[node_idx])
        return self._require_node(node_idx, "node_idx").live

    def _check_bulk_node(self, node_idx: int) -> None:
        if node_idx < 0 or node_idx >= self._bulk.node_bound:
            raise ValueError(f"node_idx {node_idx} does not exist in the graph.")

    def set_node_live(self, node_idx: int, live: bool) -> None:
        if self._bulk is not None:
            raise ValueError("NetworkStructure built with from_arrays() is immutable.")
        if node_idx < 0 or node_idx >= len(self._nodes) or self._nodes[node_idx] is None:
            raise ValueError(f"Node index {node_idx} does not exist in the graph.")
        self._nodes[node_idx].live = bool(live)  # type: ignore[union-attr]
        self._ncol["live"][node_idx] = 1 if live else 0
        self._invalidate()

    def remove_street_node(self, node_idx: int) -> None:
        """graph.rs:893-913; petgraph StableGraph::remove_node (out-edges then in-edges, newest first)."""
        if node_idx < 0 or node_idx >= len(self._nodes) or self._nodes[node_idx] is None:
            raise ValueError(f"Node index {node_idx} does not exist in the graph.")
        self._invalidate()
        for lst in (self._out, self._in):
            while lst[node_idx]:
                self._remove_edge_id(lst[node_idx][-1])
        self._nodes[node_idx] = None
        self._ncol["exists"][node_idx] = 0
        self._ncol["live"][node_idx] = 0
        self._free_nodes.append(node_idx)

    def node_count(self) -> int:
        if self._bulk is not None:
            return int(self._bulk.node_exists.sum())
        return len(self._nodes) - len(self._free_nodes)

    def node_bound(self) -> int:
        if self._bulk is not None:
            return self._bulk.node_bound
        n = len(self._nodes)
        while n > 0 and self._nodes[n - 1] is None:
            n -= 1
        return n

    def edge_bound(self) -> int:
        if self._bulk is not None:
            return self._bulk.edge_bound
        n = len(self._edges)
        while n > 0 and self._edges[n - 1] is None:
            n -= 1
        return n

    def street_node_count(self) -> int:
        return self.node_count()

    def node_indices(self) -> list[int]:
        if self._bulk is not None:
            return self._bulk.node_indices.tolist()
        return [i for i, p in enumerate(self._nodes) if p is not None]

    def street_node_indices(self) -> list[int]:
        return self.node_indices()

    def node_keys_py(self) -> list[Any]:
        return list(self._node_keys_shared())

    def _node_keys_shared(self) -> list[Any]:
        """The key list, built once per graph state and shared (read-only) by the result objects: rebuilding a
        million-element Python list per call costs more than the GPU spends on small runs."""
        if self._keys_cache is None:
            if self._bulk is not None:
                self._keys_cache = self._bulk.node_indices.tolist() if self._bulk_keys is None else list(self._bulk_keys)
            else:
                self._keys_cache = [p.node_key for p in self._nodes if p is not None]
        return self._keys_cache

    @property
    def node_xs(self) -> list[float]:
        return [p.x for p in self._nodes if p is not None]

    @property
    def node_ys(self) -> list[float]:
        return [p.y for p in self._nodes if p is not None]

    @property
    def node_xys(self) -> list[tuple[float, float]]:
        return [(p.x, p.y) for p in self._nodes if p is not None]

    @property
    def node_zs(self) -> list[float | None]:
        return [p.z for p in self._nodes if p is not None]

    @property
    def node_lives(self) -> list[bool]:
        if self._bulk is not None:
            return self._bulk.live[self._bulk.node_indices].astype(bool).tolist()
        return [p.live for p in self._nodes if p is not None]

    @property
    def street_node_lives(self) -> list[bool]:
        return self.node_lives

    # ------------------------------------------------------------------ edges
    @property
    def edge_count(self) -> int:
        if self._bulk is not None:
            return int(self._bulk.edge_exists.sum())
        return len(self._edges) - len(self._free_edges)

    def add_street_edge(
        self,
        start_nd_idx: int,
        end_nd_idx: int,
        edge_idx: int,
        start_nd_key_py,
        end_nd_key_py,
        geom_wkt: str,
        imp_factor: float | None = None,
        shared_primal_node_key: str | None = None,
    ) -> int:
        """graph.rs:728-889 — WKT → length / bearings / angle_sum (f64 → f32); ``seconds`` = NaN."""
        wkt_preview = geom_wkt if len(geom_wkt) <= 200 else f"{geom_wkt[:200]}... (truncated, {len(geom_wkt)} chars total)"
        try:
            coords = parse_linestring_wkt(geom_wkt)
        except ValueError as e:
            raise ValueError(
                f"Failed to parse WKT for street edge (idx {edge_idx}) between nodes {start_nd_idx} and {end_nd_idx}.\n"
                f"Parse error: {e}\nWKT: {wkt_preview}"
            ) from None
        if len(coords) < 2:
            raise ValueError(
                f"Street edge geometry (idx {edge_idx}) between nodes {start_nd_idx} and {end_nd_idx} must have at "
                f"least 2 coordinates. Found {len(coords)}.\nWKT: {wkt_preview}"
            )
        length, angle_sum, in_b, out_b = linestring_metrics(coords)
        imp = 1.0 if imp_factor is None else float(np.float32(imp_factor))
        if not math.isfinite(imp) or imp <= 0.0:
            raise ValueError(
                f"Invalid impedance factor ({imp}) for edge (idx {edge_idx}) between nodes {start_nd_idx} and "
                f"{end_nd_idx}.\nImpedance must be finite and positive (> 0.0).\n"
                f"Edge length: {length:.4f}m, num_coords: {len(coords)}"
            )
        p = EdgePayload()
        p.start_nd_key_py = start_nd_key_py
        p.end_nd_key_py = end_nd_key_py
        p.shared_primal_node_key = None if shared_primal_node_key is None else str(shared_primal_node_key)
        p.edge_idx = int(edge_idx)
        p.length = length
        p.angle_sum = angle_sum
        p.imp_factor = imp
        p.in_bearing = in_b
        p.out_bearing = out_b
        p.seconds = math.nan
        p.geom_wkt = geom_wkt
        p.is_transport = False
        return self._add_edge_internal(start_nd_idx, end_nd_idx, p)

    def _add_edge_internal(self, start_nd_idx: int, end_nd_idx: int, p: EdgePayload) -> int:
        self._require_node(start_nd_idx, "start_nd_idx")
        self._require_node(end_nd_idx, "end_nd_idx")
        p.validate()
        self._invalidate()
        p._src = start_nd_idx
        p._dst = end_nd_idx
        self._stamp += 1
        p._stamp = self._stamp
        key = -1
        if p.shared_primal_node_key is not None:
            key = self._key_ids.setdefault(p.shared_primal_node_key, len(self._key_ids))
        row = (start_nd_idx, end_nd_idx, p.edge_idx, p.length, p.angle_sum, p.imp_factor, p.seconds, p._stamp, key, 1)
        names = ("src", "dst", "edge_idx", "length", "angle_sum", "imp", "seconds", "stamp", "key", "exists")
        c = self._ecol
        if self._free_edges:
            eid = self._free_edges.pop()
            self._edges[eid] = p
            for k, v in zip(names, row):
                c[k][eid] = v
        else:
            eid = len(self._edges)
            self._edges.append(p)
            for k, v in zip(names, row):
                c[k].append(v)
        self._out[start_nd_idx].append(eid)
        self._in[end_nd_idx].append(eid)
        return eid

    def _remove_edge_id(self, eid: int) -> None:
        p = self._edges[eid]
        self._out[p._src].remove(eid)
        self._in[p._dst].remove(eid)
        self._edges[eid] = None
        self._ecol["exists"][eid] = 0
        self._free_edges.append(eid)

    def _find_edge(self, start_nd_idx: int, end_nd_idx: int, edge_idx: int) -> int | None:
        # petgraph edges_connecting: out-list order, newest first
        for eid in reversed(self._out[start_nd_idx]):
            p = self._edges[eid]
            if p._dst == end_nd_idx and p.edge_idx == edge_idx:
                return eid
        return None

    def remove_street_edge(self, start_nd_idx: int, end_nd_idx: int, edge_idx: int) -> None:
        """graph.rs:917-943"""
        self._require_node(start_nd_idx, "start_nd_idx")
        self._require_node(end_nd_idx, "end_nd_idx")
        eid = self._find_edge(start_nd_idx, end_nd_idx, edge_idx)
        if eid is None:
            raise ValueError(
                f"No edge with edge_idx {edge_idx} found from node {start_nd_idx} to node {end_nd_idx}."
            )
        self._invalidate()
        self._remove_edge_id(eid)

    def edge_references(self) -> list[tuple[int, int, int]]:
        if self._bulk is not None:
            b = self._bulk
            return list(zip(b.src.tolist(), b.dst.tolist(), b.edge_idx.tolist()))
        return [(p._src, p._dst, p.edge_idx) for p in self._edges if p is not None]

    def _edge_payload_checked(self, s: int, e: int, k: int) -> EdgePayload:
        self._require_node(s, "start_nd_idx")
        self._require_node(e, "end_nd_idx")
        eid = self._find_edge(s, e, k)
        if eid is None:
            raise ValueError("Edge not found")
        return self._edges[eid]  # type: ignore[return-value]

    def get_edge_payload_py(self, start_nd_idx: int, end_nd_idx: int, edge_idx: int) -> EdgePayload:
        return self._edge_payload_checked(start_nd_idx, end_nd_idx, edge_idx)

    def get_edge_length(self, start_nd_idx: int, end_nd_idx: int, edge_idx: int) -> float:
        return self._edge_payload_checked(start_nd_idx, end_nd_idx, edge_idx).length

    def get_edge_impedance(self, start_nd_idx: int, end_nd_idx: int, edge_idx: int) -> float:
        return self._edge_payload_checked(start_nd_idx, end_nd_idx, edge_idx).imp_factor

    def validate(self) -> None:
        """graph.rs:1035-1058"""
        if self.node_count() == 0:
            raise ValueError("NetworkStructure contains no nodes.")
        if self._bulk is not None:
            return
        for p in self._nodes:
            if p is not None:
                p.validate()
        for e in self._edges:
            if e is not None:
                e.validate()

    def build_edge_rtree(self) -> None:
        """The edge R-tree serves land-use assignment only (graph.rs:1064-1232) — not on this path; no-op."""

    # ------------------------------------------------------------------ bulk ingest
    @classmethod
    def from_arrays(
        cls,
        *,
        live,
        weight,
        src,
        dst,
        edge_idx,
        length,
        angle_sum=None,
        imp_factor=None,
        z=None,
        shared_key=None,
        is_dual: bool = False,
        node_keys=None,
        x=None,
        y=None,
    ) -> "NetworkStructure":
        """Array ingest: the same container state as ``add_street_node`` × N then ``add_street_edge`` × E called in
        array order (directed edges; pass both directions), with ``length`` / ``angle_sum`` already measured.
        ``shared_key`` is an int32 id per edge (dual graphs; equal ids = equal ``shared_primal_node_key``)."""
        self = cls()
        f = FrozenGraph()
        n = len(live)
        e = len(src)
        f.node_bound = n
        f.node_exists = np.ones(n, dtype=np.uint8)
        f.live = np.ascontiguousarray(live, dtype=np.uint8)
        f.weight = np.ascontiguousarray(weight, dtype=np.float32)
        f.z = np.full(n, np.nan, dtype=np.float64) if z is None else np.ascontiguousarray(z, dtype=np.float64)
        f.xs = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
        f.ys = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        if (f.xs is None) != (f.ys is None) or (f.xs is not None and (len(f.xs) != n or len(f.ys) != n)):
            raise ValueError("x and y must both be given with one value per node")
        f.edge_bound = e
        f.edge_exists = np.ones(e, dtype=np.uint8)
        f.src = np.ascontiguousarray(src, dtype=np.uint32)
        f.dst = np.ascontiguousarray(dst, dtype=np.uint32)
        f.edge_idx = np.ascontiguousarray(edge_idx, dtype=np.uint32)
        f.length = np.ascontiguousarray(length, dtype=np.float32)
        f.angle_sum = np.zeros(e, np.float32) if angle_sum is None else np.ascontiguousarray(angle_sum, dtype=np.float32)
        f.imp = np.ones(e, np.float32) if imp_factor is None else np.ascontiguousarray(imp_factor, dtype=np.float32)
        f.seconds = np.full(e, np.nan, dtype=np.float32)
        f.shared_key = np.full(e, -1, np.int32) if shared_key is None else np.ascontiguousarray(shared_key, dtype=np.int32)
        f.stamp = np.arange(1, e + 1, dtype=np.uint64)
        f.is_dual = bool(is_dual)
        f.node_indices = np.arange(n, dtype=np.int64)
        f.key_names = None
        if e and (int(f.src.max()) >= n or int(f.dst.max()) >= n):
            raise ValueError("edge endpoint index out of range")
        if not np.all(np.isfinite(f.weight)) or np.any(f.weight < 0):
            raise ValueError("Invalid street node payload: weight must be finite and non-negative (>= 0.0).")
        if not np.all(np.isfinite(f.imp)) or np.any(f.imp <= 0):
            raise ValueError("Invalid edge payload : imp_factor must be finite and positive (> 0.0).")
        if not np.all(np.isfinite(f.length)) or not np.all(np.isfinite(f.angle_sum)):
            raise ValueError("Invalid street edge payload : length and angle_sum must be finite.")
        self._is_dual = bool(is_dual)
        self._bulk = f
        self._frozen = f
        self._bulk_keys = None if node_keys is None else list(node_keys)
        return self

    # ------------------------------------------------------------------ freeze → flat arrays
    def frozen(self) -> FrozenGraph:
        """Flat arrays in petgraph index order; cached until the next mutation."""
        if self._frozen is not None:
            return self._frozen
        f = FrozenGraph()
        nb = self.node_bound()
        eb = self.edge_bound()
        def col(c, name, m, dtype):
            return np.frombuffer(c[name], dtype=dtype)[:m].copy() if m else np.zeros(0, dtype)

        nc, ec = self._ncol, self._ecol
        f.node_bound = nb
        f.node_exists = col(nc, "exists", nb, np.uint8)
        gone = f.node_exists == 0
        f.live = col(nc, "live", nb, np.uint8)
        f.weight = col(nc, "weight", nb, np.float32)
        f.z = col(nc, "z", nb, np.float64)
        f.xs = col(nc, "x", nb, np.float64)
        f.ys = col(nc, "y", nb, np.float64)
        if gone.any():  # removed nodes read as the defaults
            f.live[gone] = 0
            f.weight[gone] = 0
            f.z[gone] = np.nan
            f.xs[gone] = 0
            f.ys[gone] = 0
        f.edge_bound = eb
        f.edge_exists = col(ec, "exists", eb, np.uint8)
        egone = f.edge_exists == 0
        f.src = col(ec, "src", eb, np.uint32)
        f.dst = col(ec, "dst", eb, np.uint32)
        f.edge_idx = col(ec, "edge_idx", eb, np.uint32)
        f.length = col(ec, "length", eb, np.float32)
        f.angle_sum = col(ec, "angle_sum", eb, np.float32)
        f.imp = col(ec, "imp", eb, np.float32)
        f.seconds = col(ec, "seconds", eb, np.float32)
        f.shared_key = col(ec, "key", eb, np.int32)
        f.stamp = col(ec, "stamp", eb, np.uint64)
        if egone.any():
            for a, v in ((f.src, 0), (f.dst, 0), (f.edge_idx, 0), (f.length, 0), (f.angle_sum, 0), (f.imp, 1),
                         (f.seconds, np.nan), (f.shared_key, -1), (f.stamp, 0)):  # fmt: skip
                a[egone] = v
        key_ids = dict(self._key_ids)
        f.is_dual = self._is_dual
        f.node_indices = np.nonzero(f.node_exists)[0].astype(np.int64)
        f.key_names = key_ids
        self._frozen = f
        return f

    def device_graph(self):
        """Upload (once) and return the device-resident graph handle. Raises if the CUDA library / GPU is missing."""
        with self._lock:
            if self._device is None:
                self._device = _native.DeviceGraph(self.frozen())
            return self._device

    # ------------------------------------------------------------------ sampling plan (centrality.rs:1032-1139)
    def _expand_sampling_weights(self, weights) -> np.ndarray:
        w = np.asarray(weights, dtype=np.float32)
        nc, nb = self.node_count(), self.node_bound()
        if len(w) != nb and len(w) != nc:
            raise ValueError(f"sampling_weights length ({len(w)}) must match node_count ({nc}) or node_bound ({nb})")
        bad = np.nonzero((w < 0.0) | (w > 1.0))[0]
        if len(bad):
            i = int(bad[0])
            raise ValueError(f"sampling_weights[{i}] = {w[i]} is out of range [0.0, 1.0]")
        if len(w) == nb:
            return w.copy()
        out = np.zeros(nb, np.float32)
        out[np.asarray(self.node_indices(), dtype=np.int64)] = w
        return out

    def _prepare_sources(self, sample_probability, sampling_weights, random_seed, source_indices, shard=None):
        """``shard = (rank, world_size)``: return only that rank's contiguous block of the sources (and their weights);
        ``eligible`` still describes the whole source set (it decides the pair counts 0.5 / 1.0, centrality.rs:1802-1806).

        Returns (sources u32[], wt f32[], eligible u8[node_bound], n_visited_for_progress, is_sampled, scale)
        — prepare_source_sampling / sample_source_weight of centrality.rs:1032-1139.  A handful of numpy passes: this runs
        inside every timed call (a million-source plan costs a few milliseconds)."""
        f = self.frozen()
        nb = f.node_bound
        sw = None if sampling_weights is None else self._expand_sampling_weights(sampling_weights)
        if sample_probability is not None:
            sample_probability = float(np.float32(sample_probability))
            if sample_probability <= 0.0 or sample_probability > 1.0:
                raise ValueError("sample_probability must be in (0.0, 1.0]")
        if source_indices is not None and sw is not None:
            raise ValueError("source_indices and sampling_weights are mutually exclusive")
        cache = getattr(f, "_plan", None)
        if cache is None:  # per frozen graph: live mask, live count, whether every slot holds a node
            exists = f.node_exists.astype(bool)
            live_mask = f.live.astype(bool) & exists
            cache = {"live_mask": live_mask, "n_live": int(live_mask.sum()), "all_exist": bool(exists.all()),
                     "live_u8": live_mask.astype(np.uint8), "node_indices_u32": f.node_indices.astype(np.uint32)}  # fmt: skip
            try:
                f._plan = cache
            except AttributeError:
                pass
        n_live = cache["n_live"]
        if source_indices is not None:
            src = source_indices if isinstance(source_indices, np.ndarray) else np.asarray(list(source_indices), dtype=np.int64)
            if src.dtype.kind not in "iu":
                src = src.astype(np.int64)
            if len(src):
                lo, hi = int(src.min()), int(src.max())
                if lo < 0 or hi >= nb:
                    bad = src[(src < 0) | (src >= nb)]
                    raise ValueError(f"node index {int(bad[0])} does not exist in the graph")
                if not cache["all_exist"]:
                    missing = f.node_exists[src] == 0
                    if missing.any():
                        raise ValueError(f"node index {int(src[np.nonzero(missing)[0][0]])} does not exist in the graph")
            sources = np.ascontiguousarray(src, dtype=np.uint32)
            eligible = np.zeros(nb, np.uint8)
            eligible[sources] = 1
            n_all = len(sources)
            if shard is not None:
                from ..parallel import shard_bounds

                lo_i, hi_i = shard_bounds(n_all, shard[0], shard[1])
                sources = sources[lo_i:hi_i]
            wt = f.weight[sources]
            if sample_probability is not None and sample_probability != 1.0:
                wt = (wt / np.float32(sample_probability)).astype(np.float32)
            scale = 1.0
            if sample_probability is None:
                scale = n_live / n_all if n_all else 1.0
            return np.ascontiguousarray(sources), np.ascontiguousarray(wt, dtype=np.float32), eligible, n_all, True, scale
        node_indices = f.node_indices
        live_mask = cache["live_mask"]
        eligible = cache["live_u8"].copy()
        sources_all = node_indices
        n_sources = len(sources_all)
        wt_all = f.weight[sources_all].astype(np.float32)
        keep = live_mask[sources_all]
        if sample_probability is not None:
            # Bernoulli draw per node_bound slot. The reference draws from rand::StdRng (ChaCha12), whose stream is not
            # pinned by any upstream test (only same-seed self-consistency is); we use numpy's PCG64 — SURVEY.md §8c.
            rng = np.random.default_rng(random_seed)
            randoms = rng.random(nb, dtype=np.float32)
            p = np.full(nb, np.float32(sample_probability), np.float32)
            if sw is not None:
                p = (p * sw).astype(np.float32)
            ps = p[sources_all]
            keep = keep & (ps > 0.0) & (randoms[sources_all] < ps)
            with np.errstate(divide="ignore", invalid="ignore"):
                wt_all = (wt_all / ps).astype(np.float32)
        sources = np.ascontiguousarray(sources_all[keep], dtype=np.uint32)
        wt = np.ascontiguousarray(wt_all[keep], dtype=np.float32)
        if shard is not None:
            from ..parallel import shard_bounds

            lo_i, hi_i = shard_bounds(len(sources), shard[0], shard[1])
            sources, wt = np.ascontiguousarray(sources[lo_i:hi_i]), np.ascontiguousarray(wt[lo_i:hi_i])
        return sources, wt, eligible, n_sources, sample_probability is not None, 1.0

    # ------------------------------------------------------------------ compute entry points (CUDA only)
    def centrality_shortest(
        self,
        distances=None,
        betas=None,
        minutes=None,
        compute_closeness=None,
        compute_betweenness=None,
        min_threshold_wt=None,
        speed_m_s=None,
        tolerance=None,
        sample_probability=None,
        sampling_weights=None,
        random_seed=None,
        source_indices=None,
        pbar_disabled=None,
    ) -> "_centrality.CentralityShortestResult":
        """centrality.rs:1624-1874 — one capped search per source on the GPU; closeness scattered to targets,
        betweenness by reverse dependency accumulation; results summed into ``[7][D][node_bound]`` f64."""
        from . import pair_distances_betas_time, WALKING_SPEED

        compute_closeness = True if compute_closeness is None else bool(compute_closeness)
        compute_betweenness = True if compute_betweenness is None else bool(compute_betweenness)
        if not compute_closeness and not compute_betweenness:
            raise ValueError(
                "Either or both closeness and betweenness flags is required, but both parameters are False."
            )
        speed = float(WALKING_SPEED if speed_m_s is None else np.float32(speed_m_s))
        tol = _centrality.validate_tolerance(tolerance)
        d, b, s = pair_distances_betas_time(speed, distances, betas, minutes, min_threshold_wt)
        sources, wt, eligible, n_prog, tracked, scale = self._prepare_sources(
            sample_probability, sampling_weights, random_seed, source_indices
        )
        self.progress_init()
        dev = self.device_graph()
        out, stats = dev.centrality_shortest(
            d, b, s, speed, tol, compute_closeness, compute_betweenness, sources, wt, eligible,
            None if pbar_disabled else self._progress, n_prog,
        )  # fmt: skip
        if compute_betweenness and scale != 1.0:
            out[5:7] *= scale
        f = self.frozen()
        res = _centrality.CentralityShortestResult(d, self._node_keys_shared(), f.node_indices, out, stats)
        if tracked:
            res.sampled_source_count = int(len(sources))
            res.reachability_totals = [int(x) for x in stats["reach_totals"]] if compute_closeness else [0] * len(d)
        return res

    def _prepare_od(self, od_matrix, shard=None):
        """Flat trip lists of an OD call: the live origins with outbound trips in node order (the reference walks
        ``node_indices`` and skips the rest, centrality.rs:2470-2481), per origin its ``(destination, weight)`` pairs with
        CSR offsets.  ``shard=(rank, world_size)``: only this rank's contiguous block of the origins, cut so that the
        blocks hold about the same number of trips."""
        f = self.frozen()
        live = f.live.astype(bool) & f.node_exists.astype(bool)
        o, d, w = od_matrix._o, od_matrix._d, od_matrix._w  # sorted by (origin, destination), pairs distinct
        ok = o < f.node_bound
        ok[ok] = live[o[ok]]
        o, d, w = o[ok], d[ok], w[ok]
        if len(d) and d.max() >= f.node_bound:  # checked on the whole list, so that every rank of a sharded call raises
            raise ValueError(f"OD destination {int(d[d >= f.node_bound][0])} is out of range for node_bound {f.node_bound}")
        origins, counts = np.unique(o, return_counts=True)
        ends = np.cumsum(counts)
        lo, hi = 0, len(origins)
        if shard is not None and len(origins):
            rank, world_size = shard
            cuts = np.searchsorted(ends, ends[-1] * np.arange(1, world_size) / world_size, side="left") + 1
            cuts = np.concatenate([[0], np.minimum(cuts, len(origins)), [len(origins)]])
            lo, hi = int(cuts[rank]), int(cuts[rank + 1])
        first = int(ends[lo - 1]) if lo > 0 else 0
        last = int(ends[hi - 1]) if hi > 0 else 0
        od_off = np.zeros(hi - lo + 1, np.uint64)
        od_off[1:] = ends[lo:hi] - first
        return (origins[lo:hi].astype(np.uint32), od_off, np.ascontiguousarray(d[first:last], dtype=np.uint32),
                np.ascontiguousarray(w[first:last], dtype=np.float32))  # fmt: skip

    def betweenness_od_shortest(
        self,
        od_matrix,
        distances=None,
        betas=None,
        minutes=None,
        min_threshold_wt=None,
        speed_m_s=None,
        tolerance=None,
        pbar_disabled=None,
    ) -> "_centrality.BetweennessShortestResult":
        """centrality.rs:2419-2540 — OD-weighted betweenness: one capped search per live origin with outbound trips, the
        dependency pass seeded at its destinations only (weight w, beta seed ``w * exp(-beta * cost)``)."""
        from . import pair_distances_betas_time, WALKING_SPEED

        if not isinstance(od_matrix, _centrality.OdMatrix):
            raise TypeError("argument 'od_matrix': expected OdMatrix")
        speed = float(WALKING_SPEED if speed_m_s is None else np.float32(speed_m_s))
        d, b, s = pair_distances_betas_time(speed, distances, betas, minutes, min_threshold_wt)
        tol = _centrality.validate_tolerance(tolerance)
        f = self.frozen()
        sources, od_off, od_dst, od_w = self._prepare_od(od_matrix)
        self.progress_init()
        dev = self.device_graph()
        out, stats = dev.betweenness_od_shortest(
            d, b, s, speed, tol, sources, od_off, od_dst, od_w, None if pbar_disabled else self._progress,
            len(f.node_indices),
        )  # fmt: skip
        return _centrality.BetweennessShortestResult(d, self._node_keys_shared(), f.node_indices, out, stats)

    def centrality_simplest(
        self,
        distances=None,
        betas=None,
        minutes=None,
        compute_closeness=None,
        compute_betweenness=None,
        min_threshold_wt=None,
        speed_m_s=None,
        tolerance=None,
        angular_scaling_unit=None,
        farness_scaling_offset=None,
        sample_probability=None,
        sampling_weights=None,
        random_seed=None,
        source_indices=None,
        pbar_disabled=None,
    ) -> "_centrality.CentralitySimplestResult":
        """centrality.rs:1880-2132 — angular (simplest-path) centrality on the dual graph."""
        from . import pair_distances_betas_time, WALKING_SPEED

        if not self._is_dual:
            raise ValueError(
                "centrality_simplest requires a dual graph for angular analysis. Convert the graph with "
                "cityseer.tools.graphs.nx_to_dual(...) before ingesting it into NetworkStructure."
            )
        compute_closeness = True if compute_closeness is None else bool(compute_closeness)
        compute_betweenness = True if compute_betweenness is None else bool(compute_betweenness)
        tol = _centrality.validate_tolerance(tolerance)
        if not compute_closeness and not compute_betweenness:
            raise ValueError(
                "Either or both closeness and betweenness flags is required, but both parameters are False."
            )
        speed = float(WALKING_SPEED if speed_m_s is None else np.float32(speed_m_s))
        unit = float(np.float32(180.0 if angular_scaling_unit is None else angular_scaling_unit))
        offset = float(np.float32(1.0 if farness_scaling_offset is None else farness_scaling_offset))
        d, _b, s = pair_distances_betas_time(speed, distances, betas, minutes, min_threshold_wt)
        sources, wt, eligible, n_prog, tracked, scale = self._prepare_sources(
            sample_probability, sampling_weights, random_seed, source_indices
        )
        self.progress_init()
        dev = self.device_graph()
        out, stats = dev.centrality_simplest(
            d, s, speed, tol, unit, offset, compute_closeness, compute_betweenness, sources, wt, eligible,
            None if pbar_disabled else self._progress, n_prog,
        )  # fmt: skip
        if compute_betweenness and scale != 1.0:
            out[3:4] *= scale
        f = self.frozen()
        res = _centrality.CentralitySimplestResult(d, self._node_keys_shared(), f.node_indices, out, stats)
        if tracked:
            res.sampled_source_count = int(len(sources))
            res.reachability_totals = [int(x) for x in stats["reach_totals"]] if compute_closeness else [0] * len(d)
        return res

    def segment_centrality(
        self,
        distances=None,
        betas=None,
        minutes=None,
        compute_closeness=None,
        compute_betweenness=None,
        min_threshold_wt=None,
        speed_m_s=None,
        pbar_disabled=None,
    ) -> "_centrality.CentralitySegmentResult":
        """centrality.rs:2134-2407 — continuous (segment) closeness at the source + tree betweenness."""
        from . import pair_distances_betas_time, WALKING_SPEED

        speed = float(WALKING_SPEED if speed_m_s is None else np.float32(speed_m_s))
        d, b, s = pair_distances_betas_time(speed, distances, betas, minutes, min_threshold_wt)
        compute_closeness = True if compute_closeness is None else bool(compute_closeness)
        compute_betweenness = True if compute_betweenness is None else bool(compute_betweenness)
        if not compute_closeness and not compute_betweenness:
            raise ValueError(
                "Either or both closeness and betweenness flags is required, but both parameters are False."
            )
        f = self.frozen()
        node_indices = f.node_indices
        live = f.live[node_indices].astype(bool)
        sources = np.ascontiguousarray(node_indices[live], dtype=np.uint32)
        self.progress_init()
        dev = self.device_graph()
        out, stats = dev.segment_centrality(
            d, b, s, speed, compute_closeness, compute_betweenness, sources,
            None if pbar_disabled else self._progress, len(node_indices),
        )  # fmt: skip
        return _centrality.CentralitySegmentResult(d, self._node_keys_shared(), f.node_indices, out, stats)

    # ------------------------------------------------------------------ single-source tree searches
    def _validate_dijkstra_inputs(self, src_idx: int, speed_m_s: float) -> None:
        # centrality.rs:1009-1030
        f = self.frozen()
        if src_idx < 0 or src_idx >= f.node_bound:
            raise ValueError(f"src_idx {src_idx} out of range for network with node_bound {f.node_bound}")
        if not f.node_exists[src_idx]:
            raise ValueError(f"src_idx {src_idx} does not exist in the graph")
        if not math.isfinite(speed_m_s) or speed_m_s <= 0.0:
            raise ValueError(f"speed_m_s must be finite and positive, got {speed_m_s}")

    def dijkstra_tree_shortest(self, src_idx: int, max_seconds: int, speed_m_s: float):
        """centrality.rs:1499-1508 — device search; returns ``(visited_nodes, tree_map)``."""
        self._validate_dijkstra_inputs(src_idx, float(speed_m_s))
        speed = np.float32(speed_m_s)
        order, pred, agg = self.device_graph().dijkstra_tree(0, src_idx, int(max_seconds), float(speed))
        tree_map = [NodeVisit() for _ in range(len(pred))]
        short = agg * speed  # f32 product, as total_seconds * speed_m_s (:1188)
        for i in order.tolist():
            t = tree_map[i]
            t.visited = True
            t.discovered = True
            t.pred = None if pred[i] < 0 else int(pred[i])
            t.agg_seconds = float(agg[i])
            t.short_dist = float(short[i])
        return order.tolist(), tree_map

    def dijkstra_trees_shortest(self, src_indices, max_seconds: int, speed_m_s: float, capacity: int | None = None):
        """Extension: ``dijkstra_tree_shortest`` for many sources in one device launch — the call
        shape of the reference's data-layer consumers (data.rs:520-602 run one tree search per data point).  Returns
        ``(counts, visited_nodes, preds, agg_seconds)``: for source ``i`` the first ``counts[i]`` entries of row ``i`` are
        the settled nodes in the reference's pop order, each node's tree predecessor (-1 for the source) and its travel
        seconds (``short_dist = agg_seconds * speed_m_s``).  ``capacity`` bounds the nodes one source may settle."""
        src = np.asarray(list(src_indices) if not isinstance(src_indices, np.ndarray) else src_indices, dtype=np.int64)
        for i in src.tolist():
            self._validate_dijkstra_inputs(int(i), float(speed_m_s))
        speed = np.float32(speed_m_s)
        dev = self.device_graph()
        cap = int(capacity) if capacity else min(self.node_bound(), 16384)
        src32 = src.astype(np.uint32)
        while True:
            # a launch is bounded to 2^26 output entries (0.8 GB on the device); longer source lists go in slices
            step = max(1, (1 << 26) // cap)
            try:
                parts = [dev.dijkstra_trees_shortest(src32[i : i + step], int(max_seconds), float(speed), cap)
                         for i in range(0, max(len(src32), 1), step)]  # fmt: skip
                break
            except ValueError as e:
                if capacity or "output capacity" not in str(e) or cap >= self.node_bound():
                    raise
                cap = min(self.node_bound(), cap * 4)
        counts = np.concatenate([p[0] for p in parts])
        width = int(counts.max()) if len(counts) else 0
        order, pred, agg = (np.concatenate([p[k][:, :width] for p in parts]) for k in (1, 2, 3))
        return counts, order, pred, agg

    def dijkstra_tree_simplest(self, src_idx: int, max_seconds: int, speed_m_s: float):
        """centrality.rs:1510-1521"""
        if not self._is_dual:
            raise ValueError(
                "dijkstra_tree_simplest requires a dual graph for angular analysis. Convert the graph with "
                "cityseer.tools.graphs.nx_to_dual(...) before ingesting it into NetworkStructure."
            )
        self._validate_dijkstra_inputs(src_idx, float(speed_m_s))
        order, pred, simpl, agg, flags = self.device_graph().dijkstra_tree(
            1, src_idx, int(max_seconds), float(np.float32(speed_m_s))
        )
        tree_map = [NodeVisit() for _ in range(len(pred))]
        for i in np.nonzero(flags)[0].tolist():
            t = tree_map[i]
            t.visited = bool(flags[i] & 1)
            t.discovered = bool(flags[i] & 2)
            t.pred = None if pred[i] < 0 else int(pred[i])
            t.simpl_dist = float(simpl[i])
            t.agg_seconds = float(agg[i])
        return order.tolist(), tree_map

    def dijkstra_tree_segment(self, src_idx: int, max_seconds: int, speed_m_s: float):
        """centrality.rs:1523-1611 — returns ``(visited_nodes, visited_edges, tree_map, edge_map)``."""
        self._validate_dijkstra_inputs(src_idx, float(speed_m_s))
        speed = np.float32(speed_m_s)
        order, eorder, pred, agg, origin, last, flags = self.device_graph().dijkstra_tree(
            2, src_idx, int(max_seconds), float(speed)
        )
        tree_map = [NodeVisit() for _ in range(len(pred))]
        short = agg * speed  # f32 product, as total_seconds * speed_m_s (:1593)
        for i in np.nonzero(flags)[0].tolist():
            t = tree_map[i]
            t.visited = bool(flags[i] & 1)
            t.discovered = bool(flags[i] & 2)
            t.pred = None if pred[i] < 0 else int(pred[i])
            t.agg_seconds = float(agg[i])
            t.short_dist = float(short[i])
            t.origin_seg = None if origin[i] < 0 else int(origin[i])
            t.last_seg = None if last[i] < 0 else int(last[i])
        f = self.frozen()
        edge_map = [EdgeVisit() for _ in range(f.edge_bound)]
        for eid in eorder.tolist():
            ev = edge_map[eid]
            ev.visited = True
            ev.start_nd_idx = int(f.dst[eid])  # the settled node the edge points into (:1561, :1570)
            ev.end_nd_idx = int(f.src[eid])
            ev.edge_idx = int(f.edge_idx[eid])
        return order.tolist(), eorder.tolist(), tree_map, edge_map
