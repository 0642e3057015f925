"""torchrun --nproc-per-node N scripts/check_sharded.py — the N-rank sharded runs (cityseer_b200.parallel: sources
sharded, reduce-scatter, slices assembled in the node-shared host buffer) equal the single-GPU runs of the same calls for
all three functions and the OD call: counts bit-exact, float metrics to f64 summation order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cityseer_b200 import parallel, synth  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ["CITYSEER_B200_DEVICE"] = str(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ws = dist.get_world_size()

ns, _ = synth.config("cfg4", 0.08)
for rep in range(2):  # twice: the second call reuses the cached partial / alternates the shared host buffer
    res = parallel.centrality_shortest_sharded(ns, distances=[500, 1000, 2000])
    full = ns.centrality_shortest(distances=[500, 1000, 2000], pbar_disabled=True)
    assert np.array_equal(res._out[0], full._out[0]) and np.array_equal(res._out[2], full._out[2])
    np.testing.assert_allclose(res._out, full._out, rtol=1e-12, atol=1e-12)
seg = parallel.segment_centrality_sharded(ns, distances=[400, 800, 1600])
seg_full = ns.segment_centrality(distances=[400, 800, 1600], pbar_disabled=True)
np.testing.assert_allclose(seg._out, seg_full._out, rtol=1e-12, atol=1e-9)
# an explicit source list with sampling weights (the sampled / IPW path) shards the same way
src = np.arange(0, ns.node_bound(), 3)
sub = parallel.centrality_shortest_sharded(ns, distances=[800], source_indices=src, sample_probability=0.5)
sub_full = ns.centrality_shortest(distances=[800], source_indices=src, sample_probability=0.5, pbar_disabled=True)
np.testing.assert_allclose(sub._out, sub_full._out, rtol=1e-12, atol=1e-12)
# OD betweenness: the origins with trips shard, destinations stay with their origin
from cityseer_b200.rustalgos.centrality import OdMatrix  # noqa: E402

rng = np.random.default_rng(21)
idx = np.asarray(ns.node_indices())
o = np.repeat(rng.choice(idx, 300, replace=False), 40)
od = OdMatrix(o.tolist(), rng.choice(idx, len(o)).tolist(), rng.uniform(0.5, 3.0, len(o)).tolist())
odr = parallel.betweenness_od_shortest_sharded(ns, od, distances=[500, 1000, 2000])
od_full = ns.betweenness_od_shortest(od_matrix=od, distances=[500, 1000, 2000], pbar_disabled=True)
assert od_full._out[5].max() > 0
np.testing.assert_allclose(odr._out, od_full._out, rtol=1e-12, atol=1e-12)

nd, _ = synth.config("cfg3", 0.1)
ang = parallel.centrality_simplest_sharded(nd, distances=[1000, 2000], angular_scaling_unit=90, farness_scaling_offset=1)
ang_full = nd.centrality_simplest(distances=[1000, 2000], angular_scaling_unit=90, farness_scaling_offset=1, pbar_disabled=True)
assert np.array_equal(ang._out[0], ang_full._out[0])
np.testing.assert_allclose(ang._out, ang_full._out, rtol=1e-12, atol=1e-12)
dist.barrier()
if rank == 0:
    print(f"sharded == single over {ws} ranks: shortest (kernel {res.stats['kernel_used']}), segment, OD, simplest: ok", flush=True)
dist.destroy_process_group()
