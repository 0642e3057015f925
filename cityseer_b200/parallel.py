"""Multi-GPU execution of the centrality path: one process per GPU, sources sharded, outputs summed.

The reference has a single data-parallel axis — sources (rayon ``par_iter`` over the source list,
/root/reference/rust/src/centrality.rs:1703, :1967, :2190) — with one shared additive ``[M][D][node_bound]`` result.
Here each rank owns a contiguous block of the source list and a private device-resident partial result.  The partials
are merged with ONE ``reduce_scatter`` (sum, f64) over NCCL / NVLink: rank r ends up with the r-th slice of the summed
result, downloads only that slice — into a host buffer that all ranks of the node share (POSIX shared memory,
page-locked by every rank) — and after a barrier every rank reads the complete result from that buffer.  Against an
all-reduce followed by a full download per rank this moves 1/N of the bytes over each GPU's PCIe link and half the
bytes over NVLink.  All arrays need the sum because the reference scatters closeness to the *target*
(centrality.rs:1754-1776); integer-valued metrics stay bit-exact under any summation order (< 2^53).

There is no compute step that consumes the merged data on the device, hence no fused compute+collective kernel.
"""
from __future__ import annotations

import atexit
import os
from collections.abc import Callable

import numpy as np


def world() -> tuple[int, int]:
    """(rank, world_size) from the torchrun environment; (0, 1) when not distributed."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_bounds(n_items: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of ``n_items`` owned by ``rank``; block sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_sources(sources: np.ndarray, wt: np.ndarray, rank: int, world_size: int):
    lo, hi = shard_bounds(len(sources), rank, world_size)
    return np.ascontiguousarray(sources[lo:hi]), np.ascontiguousarray(wt[lo:hi])


def _dist_state(group=None) -> tuple[int, int]:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def sharded_sum(compute_shard: Callable[[np.ndarray, np.ndarray], "object"], sources: np.ndarray, wt: np.ndarray,
                group=None):  # fmt: skip
    """Run ``compute_shard(sources_block, wt_block)`` on this rank's block and all-reduce (sum) the returned tensor.

    ``compute_shard`` returns a torch tensor (CUDA for the product path, CPU under gloo in tests) holding this rank's
    partial ``[M][D][node_bound]`` result; the reduced tensor is returned on every rank.  (The ``*_sharded`` entry
    points below use the cheaper reduce-scatter merge; this helper is the plain form.)"""
    import torch.distributed as dist

    rank, ws = _dist_state(group)
    s, w = shard_sources(sources, wt, rank, ws)
    part = compute_shard(s, w)
    if ws > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    return part


# ---------------------------------------------------------------------------------------------- shared host result
class _SharedHost:
    """A host buffer shared by the ranks of one node: a file in ``/dev/shm`` mapped by every rank (and page-locked in
    every rank that has a GPU).  ``array`` is None when the node has no room for it."""

    def __init__(self, nbytes: int, group, pin: bool):
        import mmap
        import uuid

        import torch.distributed as dist

        rank, ws = _dist_state(group)
        self.nbytes = int(nbytes)
        self._pinned_ptr = None
        self._map = None
        self.array = None
        self.path = None
        self.owner = False
        size = max(self.nbytes, mmap.PAGESIZE)
        name = [None]
        if rank == 0 and not os.environ.get("CITYSEER_B200_NO_SHM"):
            try:  # tmpfs is sparse: check the room first, a full /dev/shm would only show as SIGBUS on first touch
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize > 2 * size + (64 << 20):
                    path = f"/dev/shm/cityseer_b200_{os.getpid()}_{uuid.uuid4().hex[:12]}"
                    fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                    try:
                        os.ftruncate(fd, size)
                        self._map = mmap.mmap(fd, size)
                    finally:
                        os.close(fd)
                    name[0] = path
                    self.owner = True
            except OSError:
                self._map = None
                name[0] = None
        if ws > 1:
            dist.broadcast_object_list(name, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if name[0] is None:
            return  # no room for a shared segment: merge_to_host falls back to all-reduce + one full download per rank
        self.path = name[0]
        if rank != 0:
            fd = os.open(self.path, os.O_RDWR)
            try:
                self._map = mmap.mmap(fd, size)
            finally:
                os.close(fd)
        self.array = np.frombuffer(self._map, dtype=np.uint8, count=self.nbytes)
        if pin:
            import torch

            ptr = self.array.ctypes.data
            err = torch.cuda.cudart().cudaHostRegister(ptr, size, 0)
            if int(err) == 0:
                self._pinned_ptr = ptr
        if ws > 1:
            dist.barrier(group=group)  # every rank has mapped the file
        if self.owner:
            os.unlink(self.path)  # the name goes away now; the pages live until the last mapping is dropped

    def close(self):
        try:
            if self._pinned_ptr is not None:
                import torch

                torch.cuda.cudart().cudaHostUnregister(self._pinned_ptr)
                self._pinned_ptr = None
        except Exception:  # noqa: BLE001
            pass
        # result arrays handed out earlier may still view the mapping: dropping our references is enough, the mapping is
        # released with the last of them (or at process exit)
        self.array = None
        self._map = None


_shared_cache: dict[tuple, _SharedHost] = {}
_partial_cache: dict[tuple, object] = {}


@atexit.register
def _release_shared():
    for b in _shared_cache.values():
        b.close()
    _shared_cache.clear()
    _partial_cache.clear()
    _toggle.clear()


_toggle: dict[tuple, int] = {}


def _shared_host(nbytes: int, group, pin: bool) -> _SharedHost:
    """Two buffers per size, used alternately: a returned result stays intact through the next merge of the same size
    (a rank that runs ahead writes into the other buffer)."""
    base = (int(nbytes), id(group), bool(pin))
    _toggle[base] = 1 - _toggle.get(base, 1)
    key = base + (_toggle[base],)
    if key not in _shared_cache:
        _shared_cache[key] = _SharedHost(nbytes, group, pin)
    return _shared_cache[key]


def merge_to_host(part, group=None) -> np.ndarray:
    """Sum the ranks' partial results and return the full result as a host array visible to every rank.

    ``part``: this rank's flat-able torch tensor (same shape on every rank).  NCCL: reduce-scatter, each rank downloads
    its slice into the node-shared page-locked buffer; gloo (CPU tests): all-reduce, each rank stores its slice.  The
    returned array is a view of a shared buffer that is reused by the second-next merge of the same size."""
    import torch
    import torch.distributed as dist

    rank, ws = _dist_state(group)
    shape = tuple(part.shape)
    total = int(part.numel())
    if ws == 1:
        if part.is_cuda:
            from . import _native

            host = _native.pinned_empty(_native.load_library(), shape)  # pooled page-locked buffer
        else:
            host = np.empty(shape, np.float64)
        torch.from_numpy(host).copy_(part)
        return host
    chunk = (total + ws - 1) // ws
    flat = part.reshape(-1)
    on_gpu = part.is_cuda
    buf = _shared_host(total * 8, group, pin=on_gpu)
    if buf.array is None:
        # no shared segment (tiny /dev/shm): every rank takes the whole sum - all-reduce, one full download per rank
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if on_gpu:
            from . import _native

            host = _native.pinned_empty(_native.load_library(), shape)
        else:
            host = np.empty(shape, np.float64)
        torch.from_numpy(host).copy_(part)
        return host
    if on_gpu and dist.get_backend(group) == "nccl":
        if chunk * ws != total:
            padded = torch.zeros(chunk * ws, dtype=part.dtype, device=part.device)
            padded[:total] = flat
            flat = padded
        mine = torch.empty(chunk, dtype=part.dtype, device=part.device)
        dist.reduce_scatter_tensor(mine, flat, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        mine = flat[rank * chunk : min(total, (rank + 1) * chunk)]
    host = buf.array.view(np.float64)
    lo, hi = rank * chunk, min(total, (rank + 1) * chunk)
    if hi > lo:
        torch.from_numpy(host[lo:hi]).copy_(mine[: hi - lo])
    if on_gpu:
        torch.cuda.current_stream(part.device).synchronize()
    dist.barrier(group=group)  # every slice is in the shared buffer
    return host.reshape(shape)


def _partial_buffer(shape, device):
    """This rank's device-resident partial result, reused across calls (the library zeroes it at the start of a call)."""
    import torch

    key = (tuple(shape), str(device))
    t = _partial_cache.get(key)
    if t is None:
        t = torch.empty(shape, dtype=torch.float64, device=device)
        _partial_cache[key] = t
    return t


def _device_of(dev):
    import torch

    return torch.device("cuda", dev.device)


def _run_on_current_stream(dev, fn):
    import torch

    dev.set_stream(torch.cuda.current_stream(_device_of(dev)).cuda_stream)
    try:
        return fn()
    finally:
        dev.set_stream(None)


# ---------------------------------------------------------------------------------------------- public entry points
def centrality_shortest_sharded(ns, distances=None, betas=None, minutes=None, compute_closeness=True,
                                compute_betweenness=True, min_threshold_wt=None, speed_m_s=None, tolerance=None,
                                source_indices=None, sample_probability=None, group=None):  # fmt: skip
    """``NetworkStructure.centrality_shortest`` over all ranks of the process group (centrality.rs:1624-1874).

    Every rank holds the same graph (replicated upload) and searches its block of the sources on its own GPU; the result
    object is identical on every rank (its arrays view the node-shared host buffer, valid until the next sharded call)."""
    from .rustalgos import WALKING_SPEED, pair_distances_betas_time
    from .rustalgos import centrality as _c

    speed = float(WALKING_SPEED if speed_m_s is None else np.float32(speed_m_s))
    tol = _c.validate_tolerance(tolerance)
    d, b, s = pair_distances_betas_time(speed, distances, betas, minutes, min_threshold_wt)
    rank, ws = _dist_state(group)
    src_block, wt_block, eligible, n_all, tracked, scale = ns._prepare_sources(sample_probability, None, None, source_indices,
                                                                               shard=(rank, ws))  # fmt: skip
    dev = ns.device_graph()
    part = _partial_buffer((7, len(d), dev.node_bound), _device_of(dev))
    _o, st = _run_on_current_stream(dev, lambda: dev.centrality_shortest(
        d, b, s, speed, tol, compute_closeness, compute_betweenness, src_block, wt_block, eligible, None, len(src_block),
        out_device_ptr=part.data_ptr()))  # fmt: skip
    host = merge_to_host(part, group)
    if compute_betweenness and scale != 1.0:
        host = host.copy()
        host[5:7] *= scale
    res = _c.CentralityShortestResult(d, ns._node_keys_shared(), ns.frozen().node_indices, host, st)
    if tracked:
        res.sampled_source_count = int(n_all) if source_indices is not None else 0
    return res


def centrality_simplest_sharded(ns, distances=None, betas=None, minutes=None, compute_closeness=True,
                                compute_betweenness=True, min_threshold_wt=None, speed_m_s=None, tolerance=None,
                                angular_scaling_unit=None, farness_scaling_offset=None, source_indices=None,
                                sample_probability=None, group=None):  # fmt: skip
    """``NetworkStructure.centrality_simplest`` (dual graph, centrality.rs:1880-2132) over all ranks."""
    from .rustalgos import WALKING_SPEED, pair_distances_betas_time
    from .rustalgos import centrality as _c

    if not ns.is_dual:
        raise ValueError("centrality_simplest requires a dual graph for angular analysis.")
    speed = float(WALKING_SPEED if speed_m_s is None else np.float32(speed_m_s))
    tol = _c.validate_tolerance(tolerance)
    unit = float(np.float32(180.0 if angular_scaling_unit is None else angular_scaling_unit))
    offset = float(np.float32(1.0 if farness_scaling_offset is None else farness_scaling_offset))
    d, _b, s = pair_distances_betas_time(speed, distances, betas, minutes, min_threshold_wt)
    rank, ws = _dist_state(group)
    src_block, wt_block, eligible, n_all, tracked, scale = ns._prepare_sources(sample_probability, None, None, source_indices,
                                                                               shard=(rank, ws))  # fmt: skip
    dev = ns.device_graph()
    part = _partial_buffer((4, len(d), dev.node_bound), _device_of(dev))
    _o, st = _run_on_current_stream(dev, lambda: dev.centrality_simplest(
        d, s, speed, tol, unit, offset, compute_closeness, compute_betweenness, src_block, wt_block, eligible, None,
        len(src_block), out_device_ptr=part.data_ptr()))  # fmt: skip
    host = merge_to_host(part, group)
    if compute_betweenness and scale != 1.0:
        host = host.copy()
        host[3:4] *= scale
    res = _c.CentralitySimplestResult(d, ns._node_keys_shared(), ns.frozen().node_indices, host, st)
    if tracked:
        res.sampled_source_count = int(n_all) if source_indices is not None else 0
    return res


def segment_centrality_sharded(ns, distances=None, betas=None, minutes=None, compute_closeness=True,
                               compute_betweenness=True, min_threshold_wt=None, speed_m_s=None, group=None):  # fmt: skip
    """``NetworkStructure.segment_centrality`` (centrality.rs:2134-2407) over all ranks: the live nodes are the sources
    (:2194); closeness lands at the source, betweenness at the tree ancestors, both merged by the same sum."""
    from .rustalgos import WALKING_SPEED, pair_distances_betas_time
    from .rustalgos import centrality as _c

    speed = float(WALKING_SPEED if speed_m_s is None else np.float32(speed_m_s))
    d, b, s = pair_distances_betas_time(speed, distances, betas, minutes, min_threshold_wt)
    f = ns.frozen()
    live = f.live[f.node_indices].astype(bool)
    sources = np.ascontiguousarray(f.node_indices[live], dtype=np.uint32)
    rank, ws = _dist_state(group)
    lo, hi = shard_bounds(len(sources), rank, ws)
    dev = ns.device_graph()
    part = _partial_buffer((4, len(d), dev.node_bound), _device_of(dev))
    _o, st = _run_on_current_stream(dev, lambda: dev.segment_centrality(
        d, b, s, speed, compute_closeness, compute_betweenness, np.ascontiguousarray(sources[lo:hi]), None, hi - lo,
        out_device_ptr=part.data_ptr()))  # fmt: skip
    host = merge_to_host(part, group)
    return _c.CentralitySegmentResult(d, ns._node_keys_shared(), f.node_indices, host, st)


def betweenness_od_shortest_sharded(ns, od_matrix, distances=None, betas=None, minutes=None, min_threshold_wt=None,
                                    speed_m_s=None, tolerance=None, group=None):  # fmt: skip
    """``NetworkStructure.betweenness_od_shortest`` (centrality.rs:2419-2540) over all ranks: the origins with outbound
    trips shard in contiguous blocks of about equal trip counts, every rank seeds its own origins' destinations, and
    the betweenness rows are merged by the same sum as the other calls."""
    from .rustalgos import WALKING_SPEED, pair_distances_betas_time
    from .rustalgos import centrality as _c

    if not isinstance(od_matrix, _c.OdMatrix):
        raise TypeError("argument 'od_matrix': expected OdMatrix")
    speed = float(WALKING_SPEED if speed_m_s is None else np.float32(speed_m_s))
    tol = _c.validate_tolerance(tolerance)
    d, b, s = pair_distances_betas_time(speed, distances, betas, minutes, min_threshold_wt)
    rank, ws = _dist_state(group)
    sources, od_off, od_dst, od_w = ns._prepare_od(od_matrix, shard=(rank, ws))
    dev = ns.device_graph()
    part = _partial_buffer((7, len(d), dev.node_bound), _device_of(dev))
    _o, st = _run_on_current_stream(dev, lambda: dev.betweenness_od_shortest(
        d, b, s, speed, tol, sources, od_off, od_dst, od_w, None, len(sources), out_device_ptr=part.data_ptr()))  # fmt: skip
    host = merge_to_host(part, group)
    return _c.BetweennessShortestResult(d, ns._node_keys_shared(), ns.frozen().node_indices, host, st)
