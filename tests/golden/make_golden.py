"""Regenerates the committed fixtures of this directory.

The reference (a Rust crate) cannot be built or imported in this environment (no cargo / rustc), so the
vectors come from the CPU oracle ``oracle/oracle.cpp`` - the restatement that ``tests/test_oracle_golden.py`` pins on
the reference's own known-answer tests (diamond constants, NetworkX closeness / betweenness, dual routes, plateau ratio,
tolerance drift).  The fixtures freeze its output on BASELINE.json's configs[0] (``mock_graph``, distances
400/800/1600) for all three centrality functions, so that neither the oracle nor the CUDA path can drift unnoticed.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import helpers as H  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    oracle.build()
    _g, _n, _e, ns = H.primal_ns()
    f = ns.frozen()
    og = oracle.OracleGraph(f)
    d, b, s = H.pair(distances=[400, 800, 1600])
    shortest, cnt = og.centrality_shortest(d, b, s, H.SPEED)
    segment, _ = og.segment_centrality(d, b, s, H.SPEED)
    _gd, _nd, _ed, nsd = H.dual_ns()
    fd = nsd.frozen()
    simplest, _ = oracle.OracleGraph(fd).centrality_simplest(d, s, H.SPEED, unit=90.0, offset=1.0)
    np.savez_compressed(
        os.path.join(HERE, "cfg1_mock_graph.npz"),
        distances=np.array(d, np.uint32), betas=np.array(b, np.float32), seconds=np.array(s, np.uint32),
        speed=np.float32(H.SPEED),
        shortest=H.compact(shortest, f), segment=H.compact(segment, f), simplest=H.compact(simplest, fd),
        settled=np.uint64(cnt["settled"]), edge_iters=np.uint64(cnt["edge_iters"]),
    )  # fmt: skip
    print("wrote cfg1_mock_graph.npz", shortest.shape, segment.shape, simplest.shape, cnt)


if __name__ == "__main__":
    main()
