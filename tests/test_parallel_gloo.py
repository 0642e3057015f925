"""Multi-rank HOST logic on CPU (no device path here: the GPU side of the sharded calls is covered by the `-m gpu` test
that runs scripts/check_sharded.py under torchrun).  Two gloo ranks shard the sources, each computes its block with the
CPU oracle standing in for the device call, and the partial results are merged (a) by the plain all-reduce helper and
(b) by ``merge_to_host`` — slices assembled in the node-shared host buffer — both equal to the single-rank result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from cityseer_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _od_matrix(ns):
    from cityseer_b200.rustalgos.centrality import OdMatrix

    idx = ns.node_indices()
    rng = np.random.default_rng(17)
    o = rng.choice(idx, 250)
    return OdMatrix(o.tolist(), rng.choice(idx, 250).tolist(), rng.uniform(0.2, 3.0, 250).tolist())


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle

    _g, _n, _e, ns = H.primal_ns()
    f = ns.frozen()
    og = oracle.OracleGraph(f)
    d, b, s = H.pair(distances=[400, 1600])
    sources, wt, eligible, *_ = ns._prepare_sources(None, None, None, None)

    def compute(src_block, wt_block):
        out, _ = og.centrality_shortest(d, b, s, H.SPEED, sources=src_block, wt=wt_block, eligible=eligible)
        return torch.from_numpy(out)

    total = parallel.sharded_sum(compute, sources, wt)
    lo, hi = parallel.shard_bounds(len(sources), rank, world)
    # the product merge: every rank's slice of the sum lands in one host buffer shared by the ranks of the node
    merged = []
    for _rep in range(3):  # alternating buffers: three merges in a row stay consistent
        part = compute(*parallel.shard_sources(sources, wt, rank, world))
        merged.append(parallel.merge_to_host(part).copy())
    assert all(np.array_equal(m, merged[0]) for m in merged)
    assert np.array_equal(merged[0], total.numpy())
    # OD betweenness: each rank's block of the origins (cut by trip count), merged the same way
    od = _od_matrix(ns)
    src_b, off_b, dst_b, w_b = ns._prepare_od(od, shard=(rank, world))
    od_part = np.zeros((7, len(d), f.node_bound))
    od_part[5], od_part[6] = og.betweenness_od(d, b, s, H.SPEED, src_b.tolist(), off_b.tolist(), dst_b.tolist(), w_b.tolist())
    od_merged = parallel.merge_to_host(torch.from_numpy(od_part)).copy()
    q.put((rank, total.numpy(), (lo, hi), od_merged, len(src_b)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("no_shm", [False, True], ids=["shared-buffer", "no-shm-fallback"])
def test_two_rank_sharded_sum_matches_single_rank(oracle_mod, no_shm, monkeypatch):
    # no_shm: the merge without a node-shared segment (a tiny /dev/shm): all-reduce + one full copy per rank
    if no_shm:
        monkeypatch.setenv("CITYSEER_B200_NO_SHM", "1")
    else:
        monkeypatch.delenv("CITYSEER_B200_NO_SHM", raising=False)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _g, _n, _e, ns = H.primal_ns()
    d, b, s = H.pair(distances=[400, 1600])
    ref, _ = oracle_mod.OracleGraph(ns.frozen()).centrality_shortest(d, b, s, H.SPEED)
    bounds = sorted(x[2] for x in got)
    assert bounds[0][0] == 0 and bounds[0][1] == bounds[1][0] and bounds[1][1] == 57
    od = _od_matrix(ns)
    src, off, dst, w = ns._prepare_od(od)
    od_ref = oracle_mod.OracleGraph(ns.frozen()).betweenness_od(d, b, s, H.SPEED, src.tolist(), off.tolist(), dst.tolist(), w.tolist())
    assert sum(x[4] for x in got) == len(src) and all(x[4] > 0 for x in got)
    for _rank, total, _b, od_merged, _n in got:
        assert np.array_equal(total[0], ref[0]) and np.array_equal(total[2], ref[2])
        np.testing.assert_allclose(total, ref, rtol=1e-12, atol=1e-12)
        assert od_merged[5].max() > 0 and np.all(od_merged[:5] == 0)
        np.testing.assert_allclose(od_merged[5], od_ref[0], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(od_merged[6], od_ref[1], rtol=1e-12, atol=1e-12)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1000003):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
