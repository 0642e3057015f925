"""ctypes binding of ``libcityseer_b200.so`` (C ABI in ``include/cityseer_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is deliberately no fallback:
a missing library or a missing GPU raises ``RuntimeError`` from every compute entry point.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CITYSEER_B200_LIB") or os.path.join(_HERE, "libcityseer_b200.so")  # env: A/B builds of the same ABI
MAX_THRESHOLDS = 16

_lib = None
_lib_lock = threading.Lock()
# options applied to every new DeviceGraph (cs_graph_set_option); CITYSEER_B200_KERNEL=1|2 pins the search kernel
DEFAULT_OPTIONS: dict = {}
if os.environ.get("CITYSEER_B200_KERNEL"):
    DEFAULT_OPTIONS["kernel"] = float(os.environ["CITYSEER_B200_KERNEL"])


class CsStats(C.Structure):
    _fields_ = [
        ("sources", C.c_uint64),
        ("settled", C.c_uint64),
        ("edge_iters", C.c_uint64),
        ("sum_ri", C.c_uint64),
        ("sum_ci", C.c_uint64),
        ("relaxations", C.c_uint64),
        ("reach_totals", C.c_uint64 * MAX_THRESHOLDS),
        ("kernel_ms", C.c_float),
        ("total_ms", C.c_float),
        ("gpu_launches", C.c_uint32),
        ("workers", C.c_uint32),
        ("phase_cycles", C.c_uint64 * 8),
        ("fallback_sources", C.c_uint64),
        ("smem_bytes", C.c_uint32),
        ("ctas_per_sm", C.c_uint32),
        ("reach_capacity", C.c_uint32),
        ("slot_capacity", C.c_uint32),
        ("kernel_used", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)

# every symbol include/cityseer_b200.h declares, with its signature (tests check the library exports all of them)
SIGNATURES = {
    "cs_last_error": (C.c_char_p, []),
    "cs_device_count": (C.c_int, []),
    "cs_host_alloc": (C.c_void_p, [C.c_uint64]),
    "cs_host_free": (None, [C.c_void_p]),
    "cs_graph_create": (
        C.c_void_p,
        [C.c_uint32, _u8p, _u8p, _f32p, _f64p, _f64p, _f64p, C.c_uint64, _u8p, _u32p, _u32p, _u32p, _f32p, _f32p, _f32p, _f32p, _i32p,
         _u64p, C.c_int, C.c_int],
    ),  # fmt: skip
    "cs_graph_destroy": (None, [C.c_void_p]),
    "cs_graph_configure": (C.c_int, [C.c_void_p, C.c_uint32, C.c_float, C.c_uint32]),
    "cs_graph_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "cs_graph_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cs_stage_sources": (C.c_int, [C.c_void_p, C.c_uint64, _u32p, _f32p, _u8p]),
    "cs_centrality_shortest": (
        C.c_int,
        [C.c_void_p, C.c_int, _u32p, _f32p, _u32p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint64, _u32p, _f32p,
         _u8p, C.c_void_p, C.c_int, C.c_int, C.POINTER(CsStats)],
    ),  # fmt: skip
    "cs_centrality_simplest": (
        C.c_int,
        [C.c_void_p, C.c_int, _u32p, _u32p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint64,
         _u32p, _f32p, _u8p, C.c_void_p, C.c_int, C.c_int, C.POINTER(CsStats)],
    ),  # fmt: skip
    "cs_segment_centrality": (
        C.c_int,
        [C.c_void_p, C.c_int, _u32p, _f32p, _u32p, C.c_float, C.c_int, C.c_int, C.c_uint64, _u32p, C.c_void_p, C.c_int,
         C.c_int, C.POINTER(CsStats)],
    ),  # fmt: skip
    "cs_betweenness_od_shortest": (
        C.c_int,
        [C.c_void_p, C.c_int, _u32p, _f32p, _u32p, C.c_float, C.c_float, C.c_uint64, _u32p, _u64p, _u32p, _f32p,
         C.c_void_p, C.c_int, C.POINTER(CsStats)],
    ),  # fmt: skip
    "cs_dijkstra_tree_shortest": (
        C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, _u32p, _u32p, C.POINTER(C.c_int64), _f32p],
    ),
    "cs_dijkstra_trees_shortest": (
        C.c_int,
        [C.c_void_p, C.c_uint64, _u32p, C.c_uint32, C.c_float, C.c_uint32, _u32p, _u32p, C.POINTER(C.c_int64), _f32p],
    ),
    "cs_dijkstra_tree_segment": (
        C.c_int,
        [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, _u32p, _u32p, _u64p, _u32p, C.POINTER(C.c_int64), _f32p,
         C.POINTER(C.c_int64), C.POINTER(C.c_int64), _u8p],
    ),  # fmt: skip
    "cs_dijkstra_tree_simplest": (
        C.c_int,
        [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, _u32p, _u32p, C.POINTER(C.c_int64), _f32p, _f32p, _u8p],
    ),  # fmt: skip
    "cs_progress": (C.c_uint64, [C.c_void_p]),
    "cs_shortest_search": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, _f32p, _f64p, _u32p]),
}


def load_library():
    """Load the CUDA library; raises RuntimeError when it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "cityseer_b200 has no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def pinned_empty(lib, shape, dtype=np.float64) -> np.ndarray:
    """A numpy array over a page-locked buffer from the library's pool (cs_host_alloc); the buffer returns to the pool
    when the array (and every view of it) is garbage-collected.  Device-to-host copies into it run at full PCIe rate."""
    import weakref

    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = lib.cs_host_alloc(max(nbytes, 8))
    if not ptr:
        raise MemoryError(_err(lib))
    buf = (C.c_uint8 * max(nbytes, 8)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lib.cs_host_free, C.c_void_p(ptr))
    return arr


def _err(lib) -> str:
    msg = lib.cs_last_error()
    return msg.decode("utf-8", "replace") if msg else "unknown error"


def _ptr(arr: np.ndarray, typ):
    return arr.ctypes.data_as(typ)


class ProgressCounter:
    """Host-visible progress value; while a device call is in flight it reads the device counter."""

    def __init__(self):
        self._base = 0
        self._reader = None

    def set(self, v: int) -> None:
        self._base = int(v)

    def attach(self, reader) -> None:
        self._reader = reader

    def detach(self, final: int) -> None:
        self._reader = None
        self._base = int(final)

    def get(self) -> int:
        r = self._reader
        if r is None:
            return self._base
        try:
            return self._base + int(r())
        except Exception:  # noqa: BLE001
            return self._base


def current_device() -> int:
    """Device for this process: LOCAL_RANK under torchrun, else CITYSEER_B200_DEVICE, else 0."""
    for key in ("CITYSEER_B200_DEVICE", "LOCAL_RANK"):
        if key in os.environ:
            return int(os.environ[key])
    return 0


class DeviceGraph:
    """Device-resident frozen graph (CSR, both orientations) + the per-warp search arena."""

    def __init__(self, frozen, device: int | None = None):
        lib = load_library()
        if lib.cs_device_count() <= 0:
            raise RuntimeError("no CUDA device available: cityseer_b200 has no CPU fallback")
        self._lib = lib
        self._frozen = frozen
        self.device = current_device() if device is None else int(device)
        f = frozen
        self.node_bound = int(f.node_bound)
        self.edge_bound = int(f.edge_bound)
        self._h = lib.cs_graph_create(
            f.node_bound, _ptr(f.node_exists, _u8p), _ptr(f.live, _u8p), _ptr(f.weight, _f32p),
            None if getattr(f, "xs", None) is None else _ptr(f.xs, _f64p),
            None if getattr(f, "ys", None) is None else _ptr(f.ys, _f64p), _ptr(f.z, _f64p),
            f.edge_bound, _ptr(f.edge_exists, _u8p), _ptr(f.src, _u32p), _ptr(f.dst, _u32p), _ptr(f.edge_idx, _u32p),
            _ptr(f.length, _f32p), _ptr(f.angle_sum, _f32p), _ptr(f.imp, _f32p), _ptr(f.seconds, _f32p),
            _ptr(f.shared_key, _i32p), _ptr(f.stamp, _u64p), 1 if f.is_dual else 0, self.device,
        )  # fmt: skip
        if not self._h:
            raise ValueError(_err(lib))
        self._call_lock = threading.Lock()
        for k, v in DEFAULT_OPTIONS.items():
            self.set_option(k, v)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.cs_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def configure(self, reach_capacity: int = 0, delta_seconds: float = 0.0, workers: int = 0) -> None:
        if self._lib.cs_graph_configure(self._h, int(reach_capacity), float(delta_seconds), int(workers)):
            raise ValueError(_err(self._lib))

    def set_option(self, name: str, value: float) -> None:
        """Named tunables (``kernel``, ``delta_factor``); see include/cityseer_b200.h."""
        if self._lib.cs_graph_set_option(self._h, name.encode(), float(value)):
            raise ValueError(_err(self._lib))

    def set_stream(self, cuda_stream: int | None) -> None:
        """Run on a caller-owned stream (``torch.cuda.current_stream().cuda_stream``); ``None`` = library stream."""
        self._lib.cs_graph_set_stream(self._h, C.c_void_p(int(cuda_stream)) if cuda_stream else None)

    def stage_sources(self, sources: np.ndarray, wt: np.ndarray, eligible: np.ndarray | None) -> int:
        """Upload a source plan; pass ``resident=True`` and the returned count to the next compute call."""
        sources = np.ascontiguousarray(sources, np.uint32)
        wt = np.ascontiguousarray(wt, np.float32)
        ep = None if eligible is None else _ptr(np.ascontiguousarray(eligible, np.uint8), _u8p)
        if self._lib.cs_stage_sources(self._h, len(sources), _ptr(sources, _u32p), _ptr(wt, _f32p), ep):
            raise ValueError(_err(self._lib))
        return len(sources)

    def progress(self) -> int:
        return int(self._lib.cs_progress(self._h))

    @staticmethod
    def _stats(st: CsStats, D: int) -> dict:
        return {
            "sources": int(st.sources),
            "settled": int(st.settled),
            "edge_iters": int(st.edge_iters),
            "sum_ri": int(st.sum_ri),
            "sum_ci": int(st.sum_ci),
            "relaxations": int(st.relaxations),
            "reach_totals": [int(st.reach_totals[i]) for i in range(D)],
            "kernel_ms": float(st.kernel_ms),
            "total_ms": float(st.total_ms),
            "gpu_launches": int(st.gpu_launches),
            "workers": int(st.workers),
            "phase_cycles": [int(st.phase_cycles[i]) for i in range(8)],
            "fallback_sources": int(st.fallback_sources),
            "smem_bytes": int(st.smem_bytes),
            "ctas_per_sm": int(st.ctas_per_sm),
            "reach_capacity": int(st.reach_capacity),
            "slot_capacity": int(st.slot_capacity),
            "kernel_used": int(st.kernel_used),
        }

    def _thresholds(self, d, b, s):
        D = len(d)
        if D < 1 or D > MAX_THRESHOLDS:
            raise ValueError(f"number of thresholds must be in [1, {MAX_THRESHOLDS}], got {D}")
        da = np.ascontiguousarray(d, dtype=np.uint32)
        ba = np.ascontiguousarray(b if b is not None else np.zeros(D), dtype=np.float32)
        sa = np.ascontiguousarray(s, dtype=np.uint32)
        return D, da, ba, sa

    def _run(self, fn, progress, n_prog, n_run):
        """Call ``fn`` with the progress counter wired to the device for the duration of the call."""
        with self._call_lock:
            skipped = int(n_prog) - int(n_run)
            if progress is not None:
                progress.set(skipped)
                progress.attach(self.progress)
            try:
                rc = fn()
            finally:
                if progress is not None:
                    progress.detach(int(n_prog))
            if rc:
                raise ValueError(_err(self._lib))

    def centrality_shortest(self, d, b, s, speed, tol, closeness, betweenness, sources, wt, eligible, progress, n_prog,
                            out_device_ptr: int | None = None, accumulate: bool = False, resident: bool = False):  # fmt: skip
        D, da, ba, sa = self._thresholds(d, b, s)
        st = CsStats()
        if out_device_ptr is None:
            out = pinned_empty(self._lib, (7, D, self.node_bound))
            optr, on_dev = out.ctypes.data_as(C.c_void_p), 0
        else:
            out, optr, on_dev = None, C.c_void_p(int(out_device_ptr)), 1
        n_src = len(sources) if not resident else int(sources)
        sp = None if resident else _ptr(sources, _u32p)
        wp = None if resident else _ptr(wt, _f32p)
        ep = None if (resident or eligible is None) else _ptr(eligible, _u8p)

        def call():
            return self._lib.cs_centrality_shortest(
                self._h, D, _ptr(da, _u32p), _ptr(ba, _f32p), _ptr(sa, _u32p), speed, tol, int(closeness),
                int(betweenness), n_src, sp, wp, ep, optr, on_dev, int(accumulate), C.byref(st),
            )  # fmt: skip

        self._run(call, progress, n_prog, n_src)
        return out, self._stats(st, D)

    def betweenness_od_shortest(self, d, b, s, speed, tol, sources, od_off, od_dst, od_w, progress, n_prog,
                                out_device_ptr: int | None = None):  # fmt: skip
        """Returns float64 [7][D][node_bound] (rows 5, 6 populated) and the device counters; with ``out_device_ptr`` the
        result stays in that device buffer (the sharded call merges it there) and ``None`` is returned for the array."""
        D, da, ba, sa = self._thresholds(d, b, s)
        st = CsStats()
        if out_device_ptr is None:
            out = pinned_empty(self._lib, (7, D, self.node_bound))
            optr, on_dev = out.ctypes.data_as(C.c_void_p), 0
        else:
            out, optr, on_dev = None, C.c_void_p(int(out_device_ptr)), 1
        sources = np.ascontiguousarray(sources, np.uint32)
        od_off = np.ascontiguousarray(od_off, np.uint64)
        od_dst = np.ascontiguousarray(od_dst, np.uint32)
        od_w = np.ascontiguousarray(od_w, np.float32)

        def call():
            return self._lib.cs_betweenness_od_shortest(
                self._h, D, _ptr(da, _u32p), _ptr(ba, _f32p), _ptr(sa, _u32p), speed, tol, len(sources),
                _ptr(sources, _u32p), _ptr(od_off, _u64p), _ptr(od_dst, _u32p), _ptr(od_w, _f32p), optr, on_dev,
                C.byref(st),
            )  # fmt: skip

        self._run(call, progress, n_prog, len(sources))
        return out, self._stats(st, D)

    def centrality_simplest(self, d, s, speed, tol, unit, offset, closeness, betweenness, sources, wt, eligible, progress,
                            n_prog, out_device_ptr: int | None = None, accumulate: bool = False):  # fmt: skip
        D, da, _ba, sa = self._thresholds(d, None, s)
        st = CsStats()
        if out_device_ptr is None:
            out = pinned_empty(self._lib, (4, D, self.node_bound))
            optr, on_dev = out.ctypes.data_as(C.c_void_p), 0
        else:
            out, optr, on_dev = None, C.c_void_p(int(out_device_ptr)), 1

        def call():
            return self._lib.cs_centrality_simplest(
                self._h, D, _ptr(da, _u32p), _ptr(sa, _u32p), speed, tol, unit, offset, int(closeness), int(betweenness),
                len(sources), _ptr(sources, _u32p), _ptr(wt, _f32p), _ptr(eligible, _u8p), optr, on_dev,
                int(accumulate), C.byref(st),
            )  # fmt: skip

        self._run(call, progress, n_prog, len(sources))
        return out, self._stats(st, D)

    def segment_centrality(self, d, b, s, speed, closeness, betweenness, sources, progress, n_prog,
                           out_device_ptr: int | None = None, accumulate: bool = False):  # fmt: skip
        D, da, ba, sa = self._thresholds(d, b, s)
        st = CsStats()
        if out_device_ptr is None:
            out = pinned_empty(self._lib, (4, D, self.node_bound))
            optr, on_dev = out.ctypes.data_as(C.c_void_p), 0
        else:
            out, optr, on_dev = None, C.c_void_p(int(out_device_ptr)), 1

        def call():
            return self._lib.cs_segment_centrality(
                self._h, D, _ptr(da, _u32p), _ptr(ba, _f32p), _ptr(sa, _u32p), speed, int(closeness), int(betweenness),
                len(sources), _ptr(sources, _u32p), optr, on_dev, int(accumulate), C.byref(st),
            )  # fmt: skip

        self._run(call, progress, n_prog, len(sources))
        return out, self._stats(st, D)

    def shortest_search(self, src: int, max_seconds: int, speed: float, tol: float = 1e-4):
        """Per-source dump (agg_seconds f32, sigma f64, pred_count u32), each sized node_bound."""
        n = self.node_bound
        agg = np.empty(n, np.float32)
        sig = np.empty(n, np.float64)
        npred = np.empty(n, np.uint32)
        with self._call_lock:
            rc = self._lib.cs_shortest_search(self._h, int(src), int(max_seconds), float(speed), float(tol),
                                              _ptr(agg, _f32p), _ptr(sig, _f64p), _ptr(npred, _u32p))  # fmt: skip
            if rc:
                raise ValueError(_err(self._lib))
        return agg, sig, npred

    def dijkstra_trees_shortest(self, sources: np.ndarray, max_seconds: int, speed: float, capacity: int):
        """Batched dijkstra_tree_shortest: (counts [n], visited_order [n, cap], pred [n, cap] (-1 none), seconds [n, cap])."""
        sources = np.ascontiguousarray(sources, np.uint32)
        n = len(sources)
        counts = np.zeros(max(n, 1), np.uint32)
        order = np.zeros((max(n, 1), capacity), np.uint32)
        pred = np.full((max(n, 1), capacity), -1, np.int64)
        agg = np.full((max(n, 1), capacity), np.inf, np.float32)
        with self._call_lock:
            rc = self._lib.cs_dijkstra_trees_shortest(self._h, n, _ptr(sources, _u32p), int(max_seconds), float(speed),
                                                      int(capacity), _ptr(counts, _u32p), _ptr(order, _u32p),
                                                      pred.ctypes.data_as(C.POINTER(C.c_int64)), _ptr(agg, _f32p))  # fmt: skip
            if rc:
                raise ValueError(_err(self._lib))
        return counts[:n], order[:n], pred[:n], agg[:n]

    def dijkstra_tree(self, kind: int, src_idx: int, max_seconds: int, speed: float):
        """Single-source tree dumps.  kind 0 = dijkstra_tree_shortest: (visited order, pred per node (-1 none), seconds
        per node); kind 1 = dijkstra_tree_simplest: (visited nodes, pred, simpl_dist, seconds, flags); kind 2 =
        dijkstra_tree_segment: (visited nodes, visited edge ids, pred, seconds, origin_seg, last_seg, flags)."""
        n = self.node_bound
        i64p = C.POINTER(C.c_int64)
        nv = C.c_uint32(0)
        order = np.zeros(max(n, 1), np.uint32)
        pred = np.zeros(max(n, 1), np.int64)
        agg = np.zeros(max(n, 1), np.float32)
        with self._call_lock:
            if kind == 0:
                rc = self._lib.cs_dijkstra_tree_shortest(self._h, int(src_idx), int(max_seconds), float(speed), C.byref(nv),
                                                         _ptr(order, _u32p), pred.ctypes.data_as(i64p), _ptr(agg, _f32p))  # fmt: skip
                if rc:
                    raise ValueError(_err(self._lib))
                return order[: nv.value], pred[:n], agg[:n]
            flags = np.zeros(max(n, 1), np.uint8)
            if kind == 1:
                simpl = np.zeros(max(n, 1), np.float32)
                rc = self._lib.cs_dijkstra_tree_simplest(self._h, int(src_idx), int(max_seconds), float(speed), C.byref(nv),
                                                         _ptr(order, _u32p), pred.ctypes.data_as(i64p), _ptr(simpl, _f32p),
                                                         _ptr(agg, _f32p), _ptr(flags, _u8p))  # fmt: skip
                if rc:
                    raise ValueError(_err(self._lib))
                return order[: nv.value], pred[:n], simpl[:n], agg[:n], flags[:n]
            if kind == 2:
                ne = C.c_uint64(0)
                eorder = np.zeros(max(self.edge_bound, 1), np.uint32)
                origin = np.zeros(max(n, 1), np.int64)
                last = np.zeros(max(n, 1), np.int64)
                rc = self._lib.cs_dijkstra_tree_segment(self._h, int(src_idx), int(max_seconds), float(speed), C.byref(nv),
                                                        _ptr(order, _u32p), C.byref(ne), _ptr(eorder, _u32p),
                                                        pred.ctypes.data_as(i64p), _ptr(agg, _f32p), origin.ctypes.data_as(i64p),
                                                        last.ctypes.data_as(i64p), _ptr(flags, _u8p))  # fmt: skip
                if rc:
                    raise ValueError(_err(self._lib))
                return order[: nv.value], eorder[: ne.value], pred[:n], agg[:n], origin[:n], last[:n], flags[:n]
        raise ValueError(f"unknown tree kind {kind}")
