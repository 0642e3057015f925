"""Quick device probe: time centrality_shortest on a named synthetic config (optionally a source subset)."""
import argparse
import json
import time

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cityseer_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="cfg2")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--nsrc", type=int, default=0)
ap.add_argument("--delta", type=float, default=0.0)
ap.add_argument("--workers", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--closeness", type=int, default=1)
ap.add_argument("--betweenness", type=int, default=1)
ap.add_argument("--distances", default="500,1000,2000")
ap.add_argument("--fn", default="shortest", choices=["shortest", "segment", "simplest"])
ap.add_argument("--opt", action="append", default=[], help="name=value device option (cs_graph_set_option)")
a = ap.parse_args()
t = time.time()
ns, info = synth.config(a.cfg, a.scale)
print("graph", info, ns.node_count(), ns.edge_count, f"{time.time() - t:.2f}s", flush=True)
t = time.time()
dev = ns.device_graph()
print(f"upload {time.time() - t:.2f}s", flush=True)
if a.delta or a.workers:
    dev.configure(0, a.delta, a.workers)
for o in a.opt:
    k, v = o.split("=")
    dev.set_option(k, float(v))
kw = {}
if a.nsrc:
    rng = np.random.default_rng(7)
    kw = dict(source_indices=np.sort(rng.choice(ns.node_bound(), a.nsrc, replace=False)).tolist(), sample_probability=1.0)
dist = [int(x) for x in a.distances.split(",")]
for rep in range(a.reps):
    t = time.time()
    if a.fn == "shortest":
        r = ns.centrality_shortest(distances=dist, compute_closeness=bool(a.closeness), compute_betweenness=bool(a.betweenness),
                                   pbar_disabled=True, **kw)
    elif a.fn == "simplest":
        r = ns.centrality_simplest(distances=dist, compute_closeness=bool(a.closeness), compute_betweenness=bool(a.betweenness),
                                   angular_scaling_unit=90, pbar_disabled=True, **kw)
    else:
        r = ns.segment_centrality(distances=dist, compute_closeness=bool(a.closeness), compute_betweenness=bool(a.betweenness),
                                  pbar_disabled=True)
    wall = time.time() - t
    s = r.stats
    print(json.dumps({"rep": rep, "wall_s": round(wall, 3), "kernel_ms": s["kernel_ms"], "total_ms": s["total_ms"],
                      "sources": s["sources"], "src_per_s_kernel": s["sources"] / (s["kernel_ms"] / 1e3),
                      "gteps": s["edge_iters"] / (s["kernel_ms"] / 1e3) / 1e9, "R": s["settled"] / max(1, s["sources"]),
                      "relax_per_settled": s["relaxations"] / max(1, s["settled"]), "workers": s["workers"],
                      "fallback": s["fallback_sources"],
                      "phase_pct": [round(100.0 * c / max(1, sum(s["phase_cycles"][:6])), 1) for c in s["phase_cycles"][:6]],
                      "cycles_per_source": sum(s["phase_cycles"][:5]) / max(1, s["sources"]),
                      "phase_kcyc": [round(c / max(1, s["sources"]) / 1e3, 1) for c in s["phase_cycles"][:6]],
                      "p1_iters": s["phase_cycles"][6] / max(1, s["sources"]), "p1_splits": s["phase_cycles"][7] / max(1, s["sources"]),
                      "kernel_used": s["kernel_used"], "smem": s["smem_bytes"], "ctas_per_sm": s["ctas_per_sm"], "rcap": s["reach_capacity"],
                      "slots": s["slot_capacity"]}), flush=True)
