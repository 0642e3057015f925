"""Multi-GPU execution of the centrality path: one process per GPU, sources sharded, outputs summed.

The reference has a single data-parallel axis — sources (rayon ``par_iter`` over ``sampling_plan.sources``,
/root/reference/rust/src/centrality.rs:1703) — with one shared additive ``[M][D][node_bound]`` result.  Here each rank
owns a contiguous block of the source list and a private device-resident result; one all-reduce (sum, f64) over
NCCL / NVLink merges them.  All seven arrays need the sum because the reference scatters closeness to the *target*
(centrality.rs:1754-1776); integer-valued metrics stay bit-exact under any summation order (< 2^53).
"""
from __future__ import annotations

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


def sharded_sum(compute_shard: Callable[[np.ndarray, np.ndarray], "object"], sources: np.ndarray, wt: np.ndarray,
                group=None):  # fmt: skip
    """Run ``compute_shard(sources_block, wt_block)`` on this rank's block and all-reduce (sum) the returned tensor.

    ``compute_shard`` returns a torch tensor (CUDA for the product path, CPU under gloo in tests) holding this rank's
    partial ``[M][D][node_bound]`` result; the reduced tensor is returned on every rank."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        rank, ws = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, ws = 0, 1
    s, w = shard_sources(sources, wt, rank, ws)
    part = compute_shard(s, w)
    if ws > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    return part


def centrality_shortest_sharded(ns, distances=None, betas=None, minutes=None, compute_closeness=True,
                                compute_betweenness=True, speed_m_s=None, tolerance=None, group=None):  # fmt: skip
    """``NetworkStructure.centrality_shortest`` over all ranks of the default process group (exact mode).

    Every rank holds the same graph (replicated upload), searches its block of the live sources on its own GPU with a
    device-resident f64 result, then one NCCL all-reduce sums the blocks.  Returns a ``CentralityShortestResult`` whose
    arrays are identical on every rank."""
    import torch

    from .rustalgos import WALKING_SPEED, pair_distances_betas_time
    from .rustalgos import centrality as _c

    speed = float(WALKING_SPEED if speed_m_s is None else np.float32(speed_m_s))
    tol = _c.validate_tolerance(tolerance)
    d, b, s = pair_distances_betas_time(speed, distances, betas, minutes)
    sources, wt, eligible, _n_prog, _tracked, _scale = ns._prepare_sources(None, None, None, None)
    dev = ns.device_graph()
    device = torch.device("cuda", dev.device)
    stats_box = {}

    def compute(src_block, wt_block):
        out = torch.zeros((7, len(d), dev.node_bound), dtype=torch.float64, device=device)
        dev.set_stream(torch.cuda.current_stream(device).cuda_stream)
        try:
            _o, st = dev.centrality_shortest(d, b, s, speed, tol, compute_closeness, compute_betweenness, src_block,
                                             wt_block, eligible, None, len(src_block), out_device_ptr=out.data_ptr())  # fmt: skip
        finally:
            dev.set_stream(None)
        stats_box.update(st)
        return out

    total = sharded_sum(compute, sources, wt, group)
    # download into a pooled page-locked buffer (full PCIe rate; every rank has its own link)
    from . import _native

    host = _native.pinned_empty(_native.load_library(), tuple(total.shape))
    torch.from_numpy(host).copy_(total)
    return _c.CentralityShortestResult(d, ns._node_keys_shared(), ns.frozen().node_indices, host, stats_box)
