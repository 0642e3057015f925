#!/usr/bin/env python
"""Benchmark of the hot path.

Default (the headline line, the configuration BASELINE.json's metric is quoted on): ``centrality_shortest`` (closeness +
betweenness, 500/1000/2000 m) on the synthetic 1M-node decomposed street graph of BASELINE.json config #4.
``--function segment`` times ``segment_centrality`` (400/800/1600 m) on the same graph (configs[3]); ``--function
simplest`` times ``centrality_simplest`` (1000/2000 m) on the 100k-node dual graph (configs[2]).

A step = one pass of the hot path over one batch of ``--batch`` sources per rank (default 131072; a different block of
the node range each step).  ``value`` counts sources of all ranks / device time, inputs resident in HBM; ``e2e`` is the
same through the public API with host buffers — ``NetworkStructure.centrality_*`` at N=1, the public
``cityseer_b200.parallel.*_sharded`` call at N>1 (H2D of the source plan, reduce-scatter, D2H of the f64 result inside
the timed region).  ``--impl reference`` times the CPU restatement of the reference's algorithm (oracle/, the Rust crate
cannot be built here) on all host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPEED = 1.33333
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent

SPECS = {
    "shortest": {"cfg": "cfg4", "distances": [500, 1000, 2000], "M": 7, "api": "centrality_shortest",
                 "metric": "node_centrality_shortest sources/sec"},
    "segment": {"cfg": "cfg4", "distances": [400, 800, 1600], "M": 4, "api": "segment_centrality",
                "metric": "segment_centrality sources/sec"},
    "simplest": {"cfg": "cfg3", "distances": [1000, 2000], "M": 4, "api": "centrality_simplest",
                 "metric": "node_centrality_simplest sources/sec", "unit": 90.0, "offset": 1.0},
}  # fmt: skip


def alg_bytes(fn: str, st: dict, D: int) -> float:
    """Algorithmic bytes of one launch: CSR row pair + one 16-byte edge record per iterated edge over the
    settled nodes, one distance write + read-back per reached node, f64 read-modify-write (16 B) per accumulated value."""
    R, E, ri, ci, n = st["settled"], st["edge_iters"], st["sum_ri"], st["sum_ci"], st["sources"]
    base = (8.0 * R + 16.0 * E) + 8.0 * R
    if fn == "shortest":
        return base + 16.0 * (5.0 * ri + 2.0 * ci)
    if fn == "simplest":
        return base + 16.0 * (3.0 * ri + 1.0 * ci)
    return base + 16.0 * ci + 8.0 * 3.0 * D * n  # segment: closeness is one source-owned 8 B write per metric/threshold


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBs"):
                if k in j:
                    return float(j[k]), "measured"
        except Exception:  # noqa: BLE001
            pass
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)  # fmt: skip
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}  # fmt: skip


def build_graph(fn: str):
    from cityseer_b200 import synth

    return synth.config(SPECS[fn]["cfg"])


def cpu_sample(ns, fn: str, n_sample: int, n_threads: int, seed: int = 7, optimised: bool = False):
    """Time the CPU restatement of the reference algorithm on a bounded random sample of sources."""
    from cityseer_b200 import rustalgos
    from oracle import oracle

    oracle.build()
    spec = SPECS[fn]
    f = ns.frozen()
    og = oracle.OracleGraph(f)
    d, b, s = rustalgos.pair_distances_betas_time(SPEED, distances=spec["distances"])
    rng = np.random.default_rng(seed)
    src = np.sort(rng.choice(f.node_bound, min(n_sample, f.node_bound), replace=False)).astype(np.uint32)
    elig = np.ones(f.node_bound, np.uint8)
    wt = np.ones(len(src), np.float32)
    t = time.perf_counter()
    if fn == "shortest":
        _out, cnt = og.centrality_shortest(d, b, s, SPEED, sources=src, wt=wt, eligible=elig, n_threads=n_threads,
                                           optimised=optimised)  # fmt: skip
    elif fn == "segment":
        _out, cnt = og.segment_centrality(d, b, s, SPEED, sources=src, n_threads=n_threads)
    else:
        _out, cnt = og.centrality_simplest(d, s, SPEED, unit=spec["unit"], offset=spec["offset"], sources=src, wt=wt,
                                           eligible=elig, n_threads=n_threads)  # fmt: skip
    dt = time.perf_counter() - t
    return len(src) / dt, cnt["edge_iters"] / dt / 1e9, dt


def host_threads() -> int:
    n = os.cpu_count() or 1
    return n - 1 if n > 2 else n  # the reference's rayon rule (rust/src/lib.rs:26-32)


def cpu_sample_size(fn: str, threads: int, secs: float) -> int:
    """Sources for roughly ``secs`` seconds of CPU work: the faithful port pays Theta(N) (shortest, simplest) or
    Theta(N + E) (segment) allocations per source, roughly 15 / 12 / 130 sources per second and thread on the three
    workloads (1M-node graph for shortest and segment, 100k-node dual for simplest; measured on the GPU box's host)."""
    per_thread = {"shortest": 15.0, "segment": 12.0, "simplest": 120.0}[fn]
    return max(threads, int(per_thread * threads * secs))


def config_block(fn, info, ns, batch, ws):
    spec = SPECS[fn]
    c = {**info, "function": spec["api"], "nodes": int(ns.node_count()), "directed_edges": int(ns.edge_count),
         "distances_m": spec["distances"], "closeness": True, "betweenness": True, "sources_per_step_per_gpu": int(batch),
         "parallelism": f"sources x{ws}",
         "l2": "per-step working set (per-warp search arenas + f64 accumulators) exceeds the 126 MB L2; no flush"}  # fmt: skip
    if fn == "simplest":
        c["angular_scaling_unit"], c["farness_scaling_offset"] = spec["unit"], spec["offset"]
        c["l2"] = "per-step working set (per-warp search arenas of the resident warps) exceeds the 126 MB L2; no flush"
    return c


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fn = args.function
    spec = SPECS[fn]
    ns, info = build_graph(fn)
    threads = host_threads()
    n_sample = args.cpu_sample or cpu_sample_size(fn, threads, 4.0)  # ~4 s of CPU work per step
    for _ in range(args.warmup):
        cpu_sample(ns, fn, max(threads, n_sample // 8), threads, seed=1)
    t0 = time.perf_counter()
    teps = []
    for k in range(args.steps):
        _r, g, _dt = cpu_sample(ns, fn, n_sample, threads, seed=100 + k)
        teps.append(g)
    total = time.perf_counter() - t0
    value = args.steps * n_sample / total
    sample = f"{n_sample} random sources per step of the {ns.node_count()}-node graph, {threads} threads"
    cfg = config_block(fn, info, ns, n_sample, 1)
    cfg["sources_per_step"] = cfg.pop("sources_per_step_per_gpu")
    line = {
        "impl": "reference", "metric": spec["metric"], "value": value, "unit": "sources/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 paths / f64 accumulators", "data": "synthetic",
        "config": cfg, "gteps": float(np.mean(teps)),
        "cpu_baseline": {"value": value, "unit": "sources/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sources/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from cityseer_b200 import parallel, rustalgos

    fn = args.function
    spec = SPECS[fn]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if ws > 1:
        dist.init_process_group("nccl", device_id=device)
    os.environ["CITYSEER_B200_DEVICE"] = str(local_rank)

    ns, info = build_graph(fn)
    f = ns.frozen()
    N = f.node_bound
    dev = ns.device_graph()
    DIST = spec["distances"]
    d, b, s = rustalgos.pair_distances_betas_time(SPEED, distances=DIST)
    D, M = len(d), spec["M"]
    batch = min(args.batch, N)
    nblocks = max(1, N // batch)
    eligible = np.ones(N, np.uint8)
    tol = 1e-4

    def block(step: int, r: int = rank) -> np.ndarray:
        k = (step * ws + r) % nblocks
        return np.arange(k * batch, k * batch + batch, dtype=np.uint32)

    out = torch.zeros((M, D, N), dtype=torch.float64, device=device)
    stream = torch.cuda.current_stream(device)
    dev.set_stream(stream.cuda_stream)
    agg = {"settled": 0, "edge_iters": 0, "sum_ri": 0, "sum_ci": 0, "kernel_ms": 0.0, "launches": 0, "sources": 0,
           "kernel_used": 0, "fallback_sources": 0}  # fmt: skip
    step_events = []
    ones = np.ones(batch, np.float32)

    def step(k: int, record: bool):
        src = block(k)
        n_res = 0
        if fn == "shortest":
            n_res = dev.stage_sources(src, ones, eligible)  # untimed: the source plan is resident in HBM
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        if fn == "shortest":
            _o, st = dev.centrality_shortest(d, b, s, SPEED, tol, True, True, n_res, None, None, None, n_res,
                                             out_device_ptr=out.data_ptr(), resident=True)  # fmt: skip
        elif fn == "segment":
            _o, st = dev.segment_centrality(d, b, s, SPEED, True, True, src, None, len(src), out_device_ptr=out.data_ptr())
        else:
            _o, st = dev.centrality_simplest(d, s, SPEED, tol, spec["unit"], spec["offset"], True, True, src, ones, eligible,
                                             None, len(src), out_device_ptr=out.data_ptr())  # fmt: skip
        if ws > 1:
            dist.all_reduce(out)
        eb.record(stream)
        if record:
            step_events.append((ea, eb))
            for key in ("settled", "edge_iters", "sum_ri", "sum_ci", "sources", "fallback_sources"):
                agg[key] += st[key]
            agg["kernel_ms"] += st["kernel_ms"]
            agg["launches"] += st["gpu_launches"]
            agg["kernel_used"] = st["kernel_used"]
        return st

    for k in range(args.warmup):
        step(k, False)
    torch.cuda.synchronize()
    if ws > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    for k in range(args.steps):
        step(args.warmup + k, True)
    torch.cuda.synchronize()
    if ws > 1:
        dist.barrier()
    # device time of the K steps: CUDA events on the launching stream around each step (kernel + all-reduce when N > 1);
    # the plan upload between steps is outside the events (inputs resident when the timed region starts)
    dev_ms = agg["kernel_ms"]
    span_ms = sum(a.elapsed_time(e) for a, e in step_events)
    t = torch.tensor([dev_ms, span_ms], dtype=torch.float64, device=device)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, span_ms_max = t.tolist()
    clocks = sampler.stop() if rank == 0 else None
    tot = torch.tensor([agg["sources"], agg["edge_iters"]], dtype=torch.float64, device=device)
    if ws > 1:
        dist.all_reduce(tot)
    total_sources, total_edges = tot.tolist()

    # ---- end-to-end through the public API: host source plan in, host result out, every step
    dev.set_stream(None)
    e2e_steps = max(1, min(args.steps, 5))
    kw = dict(distances=DIST)
    if fn == "simplest":
        kw.update(angular_scaling_unit=spec["unit"], farness_scaling_offset=spec["offset"])
    whole_graph = fn == "segment"  # the reference's segment_centrality takes no source subset: every live node, sharded
    per_step_sources = int(f.live.sum()) if whole_graph else batch * ws
    h2d = per_step_sources * (4 if whole_graph else 8) + (0 if whole_graph else N * ws)
    d2h = M * D * N * 8

    def e2e_step(k: int):
        if whole_graph:
            if ws == 1:
                ns.segment_centrality(pbar_disabled=True, **kw)
            else:
                parallel.segment_centrality_sharded(ns, **kw)
            return
        if ws == 1:
            getattr(ns, spec["api"])(source_indices=block(k), sample_probability=1.0, pbar_disabled=True, **kw)
        else:
            src_all = np.concatenate([block(k, r) for r in range(ws)])  # rank r's shard of this list is its own block
            getattr(parallel, spec["api"] + "_sharded")(ns, source_indices=src_all, sample_probability=1.0, **kw)

    e2e_step(args.warmup + args.steps)  # untimed warm-up of this path (page-locked / shared result buffers, key list)
    if ws > 1:
        e2e_step(args.warmup + args.steps)  # ... both alternating shared host buffers
    torch.cuda.synchronize()
    if ws > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(args.warmup + args.steps + 1 + k)
    torch.cuda.synchronize()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if ws > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = e2e_steps * per_step_sources / e2e_t.item()

    if rank == 0:
        peak, peak_kind = hbm_peak()
        ab = alg_bytes(fn, agg, D)
        achieved = ab / (agg["kernel_ms"] / 1e3) / 1e9 if agg["kernel_ms"] > 0 else 0.0
        kernel = {"shortest": {1: "cs_k_shortest", 3: "cs_k_shortest3"}.get(agg["kernel_used"], "?"),
                  "segment": {3: "cs_k_segment3"}.get(agg["kernel_used"], "cs_k_segment"), "simplest": "cs_k_simplest"}[fn]  # fmt: skip
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = (tj.get("kernels", {}).get(kernel) or {}).get("dram_bytes_per_launch")
            except Exception:  # noqa: BLE001
                traffic = None
        cpu = None
        if ws == 1 and not args.no_cpu:
            threads = host_threads()
            n_sample = args.cpu_sample or cpu_sample_size(fn, threads, 10.0)  # bounded sample: ~10 s of CPU work
            r, g, dt = cpu_sample(ns, fn, n_sample, threads)
            cpu = {"value": r, "unit": "sources/s", "cores": threads, "kind": "port", "gteps": g,
                   "sample": f"{n_sample} random sources (seed 7) of the same graph and thresholds, {dt:.1f} s"}  # fmt: skip
            if fn == "shortest":
                # the same arithmetic without the reference's Theta(N) per-source allocations (sparse reset): the GPU is
                # not only compared with the slow formulation
                n_opt = args.cpu_sample or max(threads, int(120.0 * threads * 10.0))
                ro, go, dto = cpu_sample(ns, fn, n_opt, threads, optimised=True)
                cpu["optimised"] = {"value": ro, "unit": "sources/s", "cores": threads, "kind": "port, sparse reset",
                                    "gteps": go, "sample": f"{n_opt} random sources (seed 7), {dto:.1f} s"}  # fmt: skip
        # value: whole-job throughput with inputs resident in HBM, on the device clock (span of the K steps, which
        # includes the all-reduce when N > 1), max over ranks
        value = total_sources / (span_ms_max / 1e3)
        line = {
            "metric": spec["metric"], "value": value, "unit": "sources/s", "n_gpus": ws, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": span_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 paths / f64 accumulators", "data": "synthetic",
            "config": config_block(fn, info, ns, batch, ws),
            "gteps": total_edges / (span_ms_max / 1e3) / 1e9,
            "kernel_ms_per_step": dev_ms_max / args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_kind, "kernel": kernel,
                         "algorithmic_bytes_per_source": ab / max(1, agg["sources"])},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "sources/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": ("NetworkStructure." if ws == 1 else "parallel.") + spec["api"] + ("" if ws == 1 else "_sharded"),
                    "sources_per_step": per_step_sources},
            "gpu_launches": agg["launches"],
            "clocks": clocks,
        }  # fmt: skip
        if agg["fallback_sources"]:
            line["config"]["heap_order_replays"] = agg["fallback_sources"]
        print(json.dumps(line), flush=True)
    if ws > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--function", default="shortest", choices=sorted(SPECS))
    ap.add_argument("--batch", type=int, default=131072)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import __graft_entry__

    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        __graft_entry__.build()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
