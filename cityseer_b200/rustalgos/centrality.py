"""Result objects — drop-in for ``cityseer.rustalgos.centrality`` (stub: /root/reference/pysrc/cityseer/rustalgos/
centrality.pyi:10-118; Rust: rust/src/centrality.rs:152-355, rust/src/common.rs:18-54).

Each metric getter returns ``dict[distance:int -> np.ndarray float64[node_count]]`` compacted over ``node_indices``
(StableGraph gaps removed), a fresh copy per access, like ``MetricResult::load_compact``.
"""
from __future__ import annotations

import logging

import numpy as np

logger = logging.getLogger(__name__)

TIE_EPSILON = float(np.float32(1e-4))  # centrality.rs:28
TOLERANCE_WARN_PCT = 2.0  # centrality.rs:30


def validate_tolerance(tolerance) -> float:
    """centrality.rs:34-48 — user percent → fraction, clamped to at least TIE_EPSILON (all in f32)."""
    pct = np.float32(0.0 if tolerance is None else tolerance)
    if pct < 0.0 or pct > 100.0:
        raise ValueError(f"Tolerance must be between 0 and 100 (percent), got {float(pct)}")
    if pct > TOLERANCE_WARN_PCT:
        logger.warning(
            f"Tolerance {float(pct):.1f}% is high — values above {TOLERANCE_WARN_PCT}% increasingly "
            "diffuse route concentration, especially at larger distance thresholds."
        )
    return float(max(pct / np.float32(100.0), np.float32(1e-4)))


class _ResultBase:
    _metrics: tuple[str, ...] = ()

    def __init__(self, distances, node_keys_py, node_indices, out, stats):
        self.distances = [int(d) for d in distances]
        self._node_keys = node_keys_py  # shared with the graph, never mutated
        self._node_indices = np.asarray(node_indices, dtype=np.int64)
        self._out = out  # float64 [M][D][node_bound]
        self.stats = stats  # device counters / timings (extension; not part of the reference surface)
        self.reachability_totals: list[int] = []
        self.sampled_source_count: int = 0

    @property
    def node_keys_py(self) -> list:
        return list(self._node_keys)

    @property
    def node_indices(self) -> list[int]:
        return self._node_indices.tolist()

    def _metric(self, m: int) -> dict[int, np.ndarray]:
        return {d: np.ascontiguousarray(self._out[m, i, self._node_indices]) for i, d in enumerate(self.distances)}


def _getter(idx: int):
    return property(lambda self: self._metric(idx))


class CentralityShortestResult(_ResultBase):
    """centrality.rs:152-233"""

    node_density = _getter(0)
    node_farness = _getter(1)
    node_cycles = _getter(2)
    node_harmonic = _getter(3)
    node_beta = _getter(4)
    node_betweenness = _getter(5)
    node_betweenness_beta = _getter(6)


class CentralitySimplestResult(_ResultBase):
    """centrality.rs:236-297"""

    node_density = _getter(0)
    node_farness = _getter(1)
    node_harmonic = _getter(2)
    node_betweenness = _getter(3)


class CentralitySegmentResult(_ResultBase):
    """centrality.rs:300-355"""

    segment_density = _getter(0)
    segment_harmonic = _getter(1)
    segment_beta = _getter(2)
    segment_betweenness = _getter(3)


class BetweennessShortestResult(_ResultBase):
    """centrality.rs:93-150 — result of ``betweenness_od_shortest``."""

    node_betweenness = _getter(5)
    node_betweenness_beta = _getter(6)


class OdMatrix:
    """Sparse origin-destination trip weights (centrality.rs:54-91): ``{origin: {destination: weight}}`` built from
    parallel arrays; a repeated (origin, destination) pair keeps the last weight, like the reference's HashMap insert.

    Held as three columns sorted by (origin, destination), one row per distinct pair - what the device call consumes
    (``NetworkStructure._prepare_od``); the nested-dict view ``map`` is built on first use."""

    def __init__(self, origins, destinations, weights):
        origins, destinations, weights = list(origins), list(destinations), list(weights)
        if len(origins) != len(destinations) or len(origins) != len(weights):
            raise ValueError(
                f"origins ({len(origins)}), destinations ({len(destinations)}), and weights ({len(weights)}) "
                "must have equal length"
            )
        o, d = np.asarray(origins), np.asarray(destinations)
        if not (len(origins) and o.dtype.kind in "iu" and d.dtype.kind in "iu"):
            # anything but plain integer columns: element by element, with Python's own conversions and errors
            o = np.asarray([int(x) for x in origins], dtype=object)
            d = np.asarray([int(x) for x in destinations], dtype=object)
        if len(origins) and (min(o.min(), d.min()) < 0):
            raise OverflowError("can't convert negative int to unsigned")
        o, d = o.astype(np.int64), d.astype(np.int64)
        w = np.asarray(weights, dtype=np.float32)
        order = np.lexsort((d, o))  # stable: among equal pairs the input order survives, the last one is kept
        o, d, w = o[order], d[order], w[order]
        keep = np.ones(len(o), bool)
        keep[:-1] = (o[1:] != o[:-1]) | (d[1:] != d[:-1])
        self._o, self._d, self._w = o[keep], d[keep], w[keep]
        self._map: dict[int, dict[int, float]] | None = None

    @property
    def map(self) -> dict[int, dict[int, float]]:
        if self._map is None:
            m: dict[int, dict[int, float]] = {}
            for o, d, w in zip(self._o.tolist(), self._d.tolist(), self._w.tolist()):
                m.setdefault(o, {})[d] = w
            self._map = m
        return self._map

    def len(self) -> int:
        """Number of non-zero OD pairs."""
        return int(len(self._o))

    def n_origins(self) -> int:
        """Number of unique origin nodes."""
        return int(len(np.unique(self._o)))
