"""ctypes wrapper of the CPU oracle (``oracle/oracle.cpp``) — TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs.  Takes the flat arrays
of ``NetworkStructure.frozen()`` so that the oracle and the CUDA library see byte-identical inputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_lib = None

COUNTER_NAMES = ("sources", "settled", "edge_iters", "sum_ri", "sum_ci", "key_ties", "multi_pred")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_graph_create.restype = C.c_void_p
        L.orc_graph_create.argtypes = [C.c_uint32, _u8p, _u8p, _f32p, _f64p, C.c_uint64, _u8p, _u32p, _u32p, _u32p, _f32p,
                                       _f32p, _f32p, _f32p, _i32p, _u64p, C.c_int]  # fmt: skip
        L.orc_graph_destroy.argtypes = [C.c_void_p]
        L.orc_centrality_shortest.argtypes = [C.c_void_p, C.c_int, _u32p, _f32p, _u32p, C.c_float, C.c_float, C.c_int,
                                              C.c_int, C.c_uint64, _u32p, _f32p, _u8p, _f64p, _u64p, _u64p, C.c_int]  # fmt: skip
        L.orc_centrality_shortest_opt.argtypes = L.orc_centrality_shortest.argtypes
        L.orc_centrality_simplest.argtypes = [C.c_void_p, C.c_int, _u32p, _u32p, C.c_float, C.c_float, C.c_float,
                                              C.c_float, C.c_int, C.c_int, C.c_uint64, _u32p, _f32p, _u8p, _f64p, _u64p,
                                              _u64p, C.c_int]  # fmt: skip
        L.orc_segment_centrality.argtypes = [C.c_void_p, C.c_int, _u32p, _f32p, _u32p, C.c_float, C.c_int, C.c_int,
                                             C.c_uint64, _u32p, _f64p, _u64p, C.c_int]  # fmt: skip
        tree_args = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, _u32p, _u64p, _i64p, _f32p, _f32p, _f32p, _i64p, _i64p, _u8p]
        L.orc_dijkstra_tree_shortest.argtypes = tree_args
        L.orc_dijkstra_tree_simplest.argtypes = tree_args
        L.orc_dijkstra_tree_segment.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, _u32p, _u64p, _u32p, _u64p,
                                                _i64p, _f32p, _f32p, _f32p, _i64p, _i64p, _u8p, _i64p, _i64p, _i64p, _u8p]  # fmt: skip
        L.orc_shortest_distances.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, _f32p, _f32p, _f64p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


class TreeNode:
    __slots__ = ("visited", "discovered", "pred", "short_dist", "simpl_dist", "origin_seg", "last_seg", "agg_seconds")


class OracleGraph:
    """CPU oracle over a ``FrozenGraph`` (see cityseer_b200.rustalgos.graph.NetworkStructure.frozen)."""

    def __init__(self, frozen):
        L = lib()
        f = frozen
        self.f = f
        self.nb = int(f.node_bound)
        self.eb = int(f.edge_bound)
        self._h = L.orc_graph_create(
            f.node_bound, _p(f.node_exists, _u8p), _p(f.live, _u8p), _p(f.weight, _f32p), _p(f.z, _f64p), f.edge_bound,
            _p(f.edge_exists, _u8p), _p(f.src, _u32p), _p(f.dst, _u32p), _p(f.edge_idx, _u32p), _p(f.length, _f32p),
            _p(f.angle_sum, _f32p), _p(f.imp, _f32p), _p(f.seconds, _f32p), _p(f.shared_key, _i32p), _p(f.stamp, _u64p),
            1 if f.is_dual else 0,
        )  # fmt: skip

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_graph_destroy(self._h)
            self._h = None

    def default_plan(self):
        """(sources, wt, eligible) for an exact run over all live nodes (centrality.rs:1066-1082)."""
        f = self.f
        live = f.live.astype(bool) & f.node_exists.astype(bool)
        sources = np.ascontiguousarray(np.nonzero(live)[0], dtype=np.uint32)
        wt = np.ascontiguousarray(f.weight[sources], dtype=np.float32)
        eligible = live.astype(np.uint8)
        return sources, wt, eligible

    def centrality_shortest(self, d, b, s, speed, tol=1e-4, closeness=True, betweenness=True, sources=None, wt=None,
                            eligible=None, n_threads=1, optimised=False):  # fmt: skip
        """``optimised=True`` runs the sparse-reset variant (same arithmetic, no Theta(N) work per source)."""
        D = len(d)
        if sources is None:
            sources, wt, eligible = self.default_plan()
        out = np.zeros((7, D, self.nb), np.float64)
        cnt = np.zeros(7, np.uint64)
        reach = np.zeros(D, np.uint64)
        da, ba, sa = (np.ascontiguousarray(d, np.uint32), np.ascontiguousarray(b, np.float32), np.ascontiguousarray(s, np.uint32))
        sources = np.ascontiguousarray(sources, np.uint32)
        wt = np.ascontiguousarray(wt, np.float32)
        eligible = np.ascontiguousarray(eligible, np.uint8)
        fn = lib().orc_centrality_shortest_opt if optimised else lib().orc_centrality_shortest
        rc = fn(self._h, D, _p(da, _u32p), _p(ba, _f32p), _p(sa, _u32p), speed, tol,
                                           int(closeness), int(betweenness), len(sources), _p(sources, _u32p),
                                           _p(wt, _f32p), _p(eligible, _u8p), _p(out, _f64p), _p(cnt, _u64p),
                                           _p(reach, _u64p), n_threads)  # fmt: skip
        assert rc == 0
        return out, dict(zip(COUNTER_NAMES, cnt.tolist()), reach_totals=reach.tolist())

    def betweenness_od(self, d, b, s, speed, sources, od_off, od_dst, od_w, tol=1e-4, n_threads=1):
        """centrality.rs:2419-2540 — returns float64 [2][D][node_bound] (betweenness, betweenness_beta)."""
        D = len(d)
        out = np.zeros((2, D, self.nb), np.float64)
        da, ba, sa = (np.ascontiguousarray(d, np.uint32), np.ascontiguousarray(b, np.float32), np.ascontiguousarray(s, np.uint32))
        sources = np.ascontiguousarray(sources, np.uint32)
        od_off = np.ascontiguousarray(od_off, np.uint64)
        od_dst = np.ascontiguousarray(od_dst, np.uint32)
        od_w = np.ascontiguousarray(od_w, np.float32)
        L = lib()
        L.orc_betweenness_od.restype = C.c_int
        rc = L.orc_betweenness_od(self._h, C.c_int(D), _p(da, _u32p), _p(ba, _f32p), _p(sa, _u32p), C.c_float(speed),
                                  C.c_float(tol), C.c_uint64(len(sources)), _p(sources, _u32p), _p(od_off, _u64p),
                                  _p(od_dst, _u32p), _p(od_w, _f32p), _p(out, _f64p), C.c_int(n_threads))  # fmt: skip
        assert rc == 0
        return out

    def centrality_simplest(self, d, s, speed, tol=1e-4, unit=180.0, offset=1.0, closeness=True, betweenness=True,
                            sources=None, wt=None, eligible=None, n_threads=1):  # fmt: skip
        D = len(d)
        if sources is None:
            sources, wt, eligible = self.default_plan()
        out = np.zeros((4, D, self.nb), np.float64)
        cnt = np.zeros(7, np.uint64)
        reach = np.zeros(D, np.uint64)
        da, sa = np.ascontiguousarray(d, np.uint32), np.ascontiguousarray(s, np.uint32)
        sources = np.ascontiguousarray(sources, np.uint32)
        wt = np.ascontiguousarray(wt, np.float32)
        eligible = np.ascontiguousarray(eligible, np.uint8)
        rc = lib().orc_centrality_simplest(self._h, D, _p(da, _u32p), _p(sa, _u32p), speed, tol, unit, offset,
                                           int(closeness), int(betweenness), len(sources), _p(sources, _u32p),
                                           _p(wt, _f32p), _p(eligible, _u8p), _p(out, _f64p), _p(cnt, _u64p),
                                           _p(reach, _u64p), n_threads)  # fmt: skip
        if rc == 1:
            raise ValueError("dual edge is missing shared_primal_node_key metadata")
        if rc == 2:
            raise ValueError("dual node references more than two primal endpoints")
        return out, dict(zip(COUNTER_NAMES, cnt.tolist()), reach_totals=reach.tolist())

    def segment_centrality(self, d, b, s, speed, closeness=True, betweenness=True, sources=None, n_threads=1):
        D = len(d)
        if sources is None:
            sources, _, _ = self.default_plan()
        out = np.zeros((4, D, self.nb), np.float64)
        cnt = np.zeros(7, np.uint64)
        da, ba, sa = (np.ascontiguousarray(d, np.uint32), np.ascontiguousarray(b, np.float32), np.ascontiguousarray(s, np.uint32))
        sources = np.ascontiguousarray(sources, np.uint32)
        rc = lib().orc_segment_centrality(self._h, D, _p(da, _u32p), _p(ba, _f32p), _p(sa, _u32p), speed, int(closeness),
                                          int(betweenness), len(sources), _p(sources, _u32p), _p(out, _f64p),
                                          _p(cnt, _u64p), n_threads)  # fmt: skip
        if rc:
            raise RuntimeError("Edge not found (reverse twin missing)")
        return out, dict(zip(COUNTER_NAMES, cnt.tolist()))

    def _tree(self, fn, src, max_seconds, speed):
        nb = self.nb
        visited = np.zeros(max(nb, 1), np.uint32)
        nv = C.c_uint64(0)
        pred = np.zeros(nb, np.int64)
        sd = np.zeros(nb, np.float32)
        sm = np.zeros(nb, np.float32)
        ag = np.zeros(nb, np.float32)
        os_ = np.zeros(nb, np.int64)
        ls = np.zeros(nb, np.int64)
        fl = np.zeros(nb, np.uint8)
        rc = fn(self._h, src, max_seconds, speed, _p(visited, _u32p), C.byref(nv), _p(pred, _i64p), _p(sd, _f32p),
                _p(sm, _f32p), _p(ag, _f32p), _p(os_, _i64p), _p(ls, _i64p), _p(fl, _u8p))  # fmt: skip
        if rc:
            raise ValueError("invalid dual metadata")
        return visited[: nv.value].tolist(), self._tree_map(pred, sd, sm, ag, os_, ls, fl)

    @staticmethod
    def _tree_map(pred, sd, sm, ag, os_, ls, fl):
        tm = []
        for i in range(len(pred)):
            t = TreeNode()
            t.visited = bool(fl[i] & 1)
            t.discovered = bool(fl[i] & 2)
            t.pred = None if pred[i] < 0 else int(pred[i])
            t.short_dist = float(sd[i])
            t.simpl_dist = float(sm[i])
            t.agg_seconds = float(ag[i])
            t.origin_seg = None if os_[i] < 0 else int(os_[i])
            t.last_seg = None if ls[i] < 0 else int(ls[i])
            tm.append(t)
        return tm

    def dijkstra_tree_shortest(self, src, max_seconds, speed):
        return self._tree(lib().orc_dijkstra_tree_shortest, src, max_seconds, speed)

    def dijkstra_tree_simplest(self, src, max_seconds, speed):
        return self._tree(lib().orc_dijkstra_tree_simplest, src, max_seconds, speed)

    def dijkstra_tree_segment(self, src, max_seconds, speed):
        """(visited_nodes, visited_edges, tree_map, edge_map) of centrality.rs:1523-1611; edge_map entries are
        (visited, start_nd_idx, end_nd_idx, edge_idx) tuples with None for unset fields."""
        nb, eb = self.nb, self.eb
        visited = np.zeros(max(nb, 1), np.uint32)
        ve = np.zeros(max(eb, 1), np.uint32)
        nv, ne = C.c_uint64(0), C.c_uint64(0)
        pred, os_, ls = (np.zeros(nb, np.int64) for _ in range(3))
        sd, sm, ag = (np.zeros(nb, np.float32) for _ in range(3))
        fl = np.zeros(nb, np.uint8)
        es, ee, ei = (np.zeros(max(eb, 1), np.int64) for _ in range(3))
        ev = np.zeros(max(eb, 1), np.uint8)
        lib().orc_dijkstra_tree_segment(self._h, src, max_seconds, speed, _p(visited, _u32p), C.byref(nv), _p(ve, _u32p),
                                        C.byref(ne), _p(pred, _i64p), _p(sd, _f32p), _p(sm, _f32p), _p(ag, _f32p),
                                        _p(os_, _i64p), _p(ls, _i64p), _p(fl, _u8p), _p(es, _i64p), _p(ee, _i64p),
                                        _p(ei, _i64p), _p(ev, _u8p))  # fmt: skip
        opt = lambda x: None if x < 0 else int(x)  # noqa: E731
        edge_map = [(bool(ev[i]), opt(es[i]), opt(ee[i]), opt(ei[i])) for i in range(eb)]
        return (visited[: nv.value].tolist(), ve[: ne.value].tolist(), self._tree_map(pred, sd, sm, ag, os_, ls, fl),
                edge_map)  # fmt: skip

    def shortest_distances(self, src, max_seconds, speed):
        nb = self.nb
        agg = np.zeros(nb, np.float32)
        rc_ = np.zeros(nb, np.float32)
        sig = np.zeros(nb, np.float64)
        lib().orc_shortest_distances(self._h, src, max_seconds, speed, _p(agg, _f32p), _p(rc_, _f32p), _p(sig, _f64p))
        return agg, rc_, sig
