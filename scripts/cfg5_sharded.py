"""torchrun --nproc-per-node N scripts/cfg5_sharded.py [--km 5 10] — BASELINE.json configs[4]: the 4M-node metro graph,
``centrality_shortest`` at 5 km and 10 km sharded over N GPUs (graph replicated, sources in contiguous blocks, one
reduce-scatter of the f64 result, slices assembled in the node-shared page-locked buffer).

Prints one JSON line per distance: whole-graph wall time through ``parallel.centrality_shortest_sharded`` (host source
plan in, host result out), sources/s, GTEPS, and two size-independent checks: the density column sums equal the
per-threshold reachable-target totals counted on the devices, and a block of sources run on rank 0 alone equals the same
block through the sharded call."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cityseer_b200 import parallel, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--km", type=float, nargs="+", default=[5.0, 10.0])
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--sources", type=int, default=0, help="0 = every node (exact run)")
a = ap.parse_args()

rank, local, ws = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
os.environ["CITYSEER_B200_DEVICE"] = str(local)
if ws > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t0 = time.time()
ns, info = synth.config("cfg5", a.scale)
N = ns.node_bound()
dev = ns.device_graph()
t_build = time.time() - t0
for km in a.km:
    d = [int(km * 1000)]
    kw = {}
    n_src = N
    if a.sources:
        rng = np.random.default_rng(7)
        kw = dict(source_indices=np.sort(rng.choice(N, a.sources, replace=False)), sample_probability=1.0)
        n_src = a.sources
    for _ in range(2):  # warm-up: arena, page-locked result buffers (the merge alternates between two)
        parallel.centrality_shortest_sharded(ns, distances=d, source_indices=np.arange(0, N, max(1, N // 4096)), sample_probability=1.0)
    torch.cuda.synchronize()
    if ws > 1:
        dist.barrier()
    t = time.perf_counter()
    res = parallel.centrality_shortest_sharded(ns, distances=d, **kw)
    torch.cuda.synchronize()
    wall = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
    st = res.stats
    tot = torch.tensor([st["sources"], st["edge_iters"], st["settled"], st["reach_totals"][0]], dtype=torch.float64, device="cuda")
    kms = torch.tensor([st["kernel_ms"]], dtype=torch.float64, device="cuda")
    if ws > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    sources, edges, settled, reach = tot.tolist()
    dens_sum = float(res._out[0].sum())
    # a block of sources alone on this rank's GPU == the same block through the sharded call
    blk = np.arange(N // 2, N // 2 + 4096 * ws, dtype=np.int64)
    sh = parallel.centrality_shortest_sharded(ns, distances=d, source_indices=blk, sample_probability=1.0)._out.copy()
    ok_block = None
    if rank == 0:
        solo = ns.centrality_shortest(distances=d, source_indices=blk, sample_probability=1.0, pbar_disabled=True)._out
        ok_block = bool(np.array_equal(solo[0], sh[0]) and np.array_equal(solo[2], sh[2]) and np.allclose(solo, sh, rtol=1e-12, atol=1e-12))
        print(json.dumps({
            "workload": info["workload"], "nodes": int(ns.node_count()), "directed_edges": int(ns.edge_count), "n_gpus": ws,
            "distance_m": d[0], "sources": int(sources), "wall_s": wall.item(), "sources_per_s_e2e": sources / wall.item(),
            "kernel_ms_max": kms.item(), "sources_per_s_kernel": sources / (kms.item() / 1e3),
            "gteps_kernel": edges / (kms.item() / 1e3) / 1e9, "reach_per_source": settled / max(1.0, sources),
            "kernel_used": st["kernel_used"], "workers_per_gpu": st["workers"], "reach_capacity": st["reach_capacity"],
            "density_sum_equals_reach_totals": bool(dens_sum == reach), "sharded_block_equals_single_gpu": ok_block,
            "graph_build_upload_s": t_build}), flush=True)
    if ws > 1:
        dist.barrier()
if ws > 1:
    dist.destroy_process_group()
