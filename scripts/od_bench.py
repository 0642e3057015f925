"""betweenness_od_shortest on the 1M-node decomposed workload (cfg #4): the chain-contracted kernel against the global-arena
kernel on the same origin-destination lists, device time only.  usage: od_bench.py [origins] [destinations per origin] [lattice side]"""
import json
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cityseer_b200 import rustalgos, synth

n_orig = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n_dest = int(sys.argv[2]) if len(sys.argv) > 2 else 64
side = int(sys.argv[3]) if len(sys.argv) > 3 else 333
xy, e = synth.lattice(side, side, seed=42)
xy, e = synth.decompose(xy, e, 20.0)
ns = synth.primal_network(xy, e)
info = {"workload": "cfg4-1M-decomposed-20m" if side == 333 else f"cfg4-decomposed-lattice-{side}"}
f = ns.frozen()
rng = np.random.default_rng(5)
sources = np.sort(rng.choice(f.node_indices, min(n_orig, len(f.node_indices)), replace=False)).astype(np.uint32)
n_orig = len(sources)
# destinations: the node nearest to a random point within 1.5 km of the origin (most lie within the 2 km threshold)
r = 1500.0 * np.sqrt(rng.uniform(0, 1, (n_orig, n_dest)))
a = rng.uniform(0, 2 * np.pi, (n_orig, n_dest))
pts = xy[sources][:, None, :] + np.stack([r * np.cos(a), r * np.sin(a)], axis=2)
dst = np.sort(cKDTree(xy).query(pts.reshape(-1, 2), workers=-1)[1].reshape(n_orig, n_dest), axis=1)
keep = np.ones(dst.shape, bool)
keep[:, 1:] = dst[:, 1:] != dst[:, :-1]  # unique per origin, like the reference's map
od_off = np.concatenate([[0], np.cumsum(keep.sum(1))]).astype(np.uint64)
od_dst = dst[keep].astype(np.uint32)
od_w = rng.uniform(0.5, 2.0, len(od_dst)).astype(np.float32)
d, b, s = rustalgos.pair_distances_betas_time(1.33333, distances=[500, 1000, 2000])
dev = ns.device_graph()
res = {}
for kernel in (3, 1):
    dev.set_option("kernel", float(kernel))
    best = None
    for _ in range(3):
        out, st = dev.betweenness_od_shortest(d, b, s, 1.33333, rustalgos.centrality.validate_tolerance(None), sources, od_off,
                                              od_dst, od_w, None, 0)  # fmt: skip
        assert st["kernel_used"] == kernel
        best = st["kernel_ms"] if best is None else min(best, st["kernel_ms"])
    res[kernel] = (best, out[5:].copy())
np.testing.assert_allclose(res[3][1], res[1][1], rtol=1e-9, atol=1e-12)
print(json.dumps({
    "workload": info["workload"], "origins": n_orig, "pairs": int(len(od_dst)), "distances_m": [500, 1000, 2000],
    "chain_kernel_ms": res[3][0], "arena_kernel_ms": res[1][0],
    "chain_origins_per_s": n_orig / res[3][0] * 1e3, "arena_origins_per_s": n_orig / res[1][0] * 1e3,
    "betweenness_sum": float(res[3][1][0].sum()),
}))  # fmt: skip
