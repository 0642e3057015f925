"""torchrun --nproc-per-node N scripts/check_sharded.py — the N-rank sharded run (cityseer_b200.parallel) equals the
single-GPU run of the same call: counts bit-exact, float metrics to f64 summation order."""
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
ns, _ = synth.config("cfg4", 0.08)
res = parallel.centrality_shortest_sharded(ns, distances=[500, 1000, 2000])
full = ns.centrality_shortest(distances=[500, 1000, 2000], pbar_disabled=True)
assert np.array_equal(res._out[0], full._out[0]) and np.array_equal(res._out[2], full._out[2])
np.testing.assert_allclose(res._out, full._out, rtol=1e-12, atol=1e-12)
dist.barrier()
if rank == 0:
    print(f"sharded == single over {dist.get_world_size()} ranks: ok; kernel {res.stats['kernel_used']}", flush=True)
dist.destroy_process_group()
