#!/usr/bin/env python
"""Benchmark of the hot path: ``centrality_shortest`` (closeness + betweenness, 500/1000/2000 m) on the synthetic
1M-node decomposed street graph of BASELINE.json config #4 — the configuration the metric
("node_centrality_shortest sources/sec & GTEPS, 1M-node graph, d<=2km, 1-8 GPU") is quoted on.

A step = one pass of the hot path over one batch of ``--batch`` sources per rank (default 131072; a different block of
the node range each step).  ``value`` counts sources of all ranks / device time, inputs resident in HBM; ``e2e`` is the
same through the public ``NetworkStructure.centrality_shortest`` call with host buffers (H2D of the source plan, D2H of
the [7][D][N] f64 result inside the timed region).  ``--impl reference`` times the CPU restatement of the reference's
algorithm (oracle/, the Rust crate cannot be built here) on all host cores over a bounded sample of the same workload.
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

DISTANCES = [500, 1000, 2000]
SPEED = 1.33333
METRIC = "node_centrality_shortest sources/sec"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def alg_bytes(st: dict) -> float:
    """Algorithmic bytes of a centrality_shortest launch: CSR rows + 16-byte edge records over the
    settled nodes, one distance write + read-back per reached node, f64 read-modify-write per accumulated metric."""
    R, E, ri, ci = st["settled"], st["edge_iters"], st["sum_ri"], st["sum_ci"]
    return (8.0 * R + 16.0 * E) + 8.0 * R + 16.0 * (5.0 * ri + 2.0 * ci)


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


def build_graph():
    from cityseer_b200 import synth

    ns, info = synth.config("cfg4")
    return ns, info


def cpu_sample(ns, n_sample: int, n_threads: int, seed: int = 7):
    """Time the CPU restatement of the reference algorithm on a bounded random sample of sources."""
    from cityseer_b200 import rustalgos
    from oracle import oracle

    oracle.build()
    f = ns.frozen()
    og = oracle.OracleGraph(f)
    d, b, s = rustalgos.pair_distances_betas_time(SPEED, distances=DISTANCES)
    rng = np.random.default_rng(seed)
    src = np.sort(rng.choice(f.node_bound, n_sample, replace=False)).astype(np.uint32)
    elig = np.ones(f.node_bound, np.uint8)
    t = time.perf_counter()
    _out, cnt = og.centrality_shortest(d, b, s, SPEED, sources=src, wt=np.ones(len(src), np.float32), eligible=elig,
                                       n_threads=n_threads)  # fmt: skip
    dt = time.perf_counter() - t
    return len(src) / dt, cnt["edge_iters"] / dt / 1e9, dt


def host_threads() -> int:
    n = os.cpu_count() or 1
    return n - 1 if n > 2 else n  # the reference's rayon rule (rust/src/lib.rs:26-32)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ns, info = build_graph()
    threads = host_threads()
    n_sample = args.cpu_sample or max(64, 64 * threads)  # ~4 s of CPU work per step on the GPU box's host
    for _ in range(args.warmup):
        cpu_sample(ns, max(threads, n_sample // 8), threads, seed=1)
    t0 = time.perf_counter()
    rates, teps = [], []
    for k in range(args.steps):
        r, g, _dt = cpu_sample(ns, n_sample, threads, seed=100 + k)
        rates.append(r)
        teps.append(g)
    total = time.perf_counter() - t0
    value = args.steps * n_sample / total
    sample = f"{n_sample} random sources per step of the {ns.node_count()}-node graph, {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "sources/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 paths / f64 accumulators", "data": "synthetic",
        "config": {**info, "function": "centrality_shortest", "distances_m": DISTANCES, "closeness": True,
                   "betweenness": True, "sources_per_step": n_sample},
        "gteps": float(np.mean(teps)),
        "cpu_baseline": {"value": value, "unit": "sources/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sources/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from cityseer_b200 import _native, rustalgos

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

    ns, info = build_graph()
    f = ns.frozen()
    N = f.node_bound
    dev = ns.device_graph()
    d, b, s = rustalgos.pair_distances_betas_time(SPEED, distances=DISTANCES)
    D = len(d)
    batch = min(args.batch, N)
    nblocks = max(1, N // batch)
    eligible = np.ones(N, np.uint8)
    tol = 1e-4

    def block(step: int) -> np.ndarray:
        k = (step * ws + rank) % nblocks
        return np.arange(k * batch, k * batch + batch, dtype=np.uint32)

    out = torch.zeros((7, D, N), dtype=torch.float64, device=device)
    stream = torch.cuda.current_stream(device)
    dev.set_stream(stream.cuda_stream)
    agg = {"settled": 0, "edge_iters": 0, "sum_ri": 0, "sum_ci": 0, "kernel_ms": 0.0, "launches": 0, "sources": 0,
           "kernel_used": 0}  # fmt: skip

    step_events = []

    def step(k: int, record: bool):
        src = block(k)
        n_res = dev.stage_sources(src, np.ones(len(src), np.float32), eligible)  # untimed: plan resident in HBM
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        _o, st = dev.centrality_shortest(d, b, s, SPEED, tol, True, True, n_res, None, None, None, n_res,
                                         out_device_ptr=out.data_ptr(), resident=True)  # fmt: skip
        if ws > 1:
            dist.all_reduce(out)
        eb.record(stream)
        if record:
            step_events.append((ea, eb))
            for key in ("settled", "edge_iters", "sum_ri", "sum_ci", "sources"):
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

    # ---- end-to-end through the public API: host plan in, host result out, every step
    dev.set_stream(None)
    e2e_steps = max(1, min(args.steps, 5))
    h2d = batch * 8 + N
    d2h = 7 * D * N * 8

    def e2e_step(k: int):
        src = block(k)
        if ws == 1:
            ns.centrality_shortest(distances=DISTANCES, source_indices=src, sample_probability=1.0, pbar_disabled=True)
        else:
            # the path of cityseer_b200.parallel.centrality_shortest_sharded with this rank's block of sources: device-resident
            # partial result, one all-reduce, download into a pooled page-locked buffer
            part = torch.zeros((7, D, N), dtype=torch.float64, device=device)
            dev.centrality_shortest(d, b, s, SPEED, tol, True, True, src, np.ones(len(src), np.float32), eligible, None,
                                    len(src), out_device_ptr=part.data_ptr())  # fmt: skip
            dist.all_reduce(part)
            host = _native.pinned_empty(_native.load_library(), (7, D, N))
            torch.from_numpy(host).copy_(part)

    e2e_step(args.warmup + args.steps)  # untimed warm-up of this path (page-locked result buffer, key list)
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
    e2e_value = e2e_steps * batch * ws / e2e_t.item()

    if rank == 0:
        peak, peak_kind = hbm_peak()
        achieved = alg_bytes(agg) / (agg["kernel_ms"] / 1e3) / 1e9 if agg["kernel_ms"] > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:  # noqa: BLE001
                traffic = None
        cpu = None
        if ws == 1 and not args.no_cpu:
            threads = host_threads()
            n_sample = args.cpu_sample or max(256, 160 * threads)  # bounded sample: ~10 s of CPU work
            r, g, dt = cpu_sample(ns, n_sample, threads)
            cpu = {"value": r, "unit": "sources/s", "cores": threads, "kind": "port", "gteps": g,
                   "sample": f"{n_sample} random sources (seed 7) of the same graph and thresholds, {dt:.1f} s"}  # fmt: skip
        # value: whole-job throughput with inputs resident in HBM, on the device clock (span of the K steps, which
        # includes the all-reduce when N > 1), max over ranks
        value = total_sources / (span_ms_max / 1e3)
        line = {
            "metric": METRIC, "value": value, "unit": "sources/s", "n_gpus": ws, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": span_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 paths / f64 accumulators", "data": "synthetic",
            "config": {**info, "function": "centrality_shortest", "nodes": int(ns.node_count()),
                       "directed_edges": int(ns.edge_count), "distances_m": DISTANCES, "closeness": True,
                       "betweenness": True, "sources_per_step_per_gpu": batch, "parallelism": f"sources x{ws}",
                       "l2": "per-step working set (per-warp search arenas + 172 MB of f64 accumulators) exceeds the 126 MB L2; no flush"},
            "gteps": total_edges / (span_ms_max / 1e3) / 1e9,
            "kernel_ms_per_step": dev_ms_max / args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_kind,
                         "kernel": {1: "cs_k_shortest", 2: "cs_k_shortest2", 3: "cs_k_shortest3"}.get(agg["kernel_used"], "?"),
                         "algorithmic_bytes_per_source": alg_bytes(agg) / max(1, agg["sources"])},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "sources/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": agg["launches"],
            "clocks": clocks,
        }  # fmt: skip
        print(json.dumps(line), flush=True)
    if ws > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
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
