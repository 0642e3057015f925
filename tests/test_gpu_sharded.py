"""The multi-GPU entry points (cityseer_b200.parallel.*_sharded) on real devices: scripts/check_sharded.py under torchrun
on two GPUs compares every sharded call with the single-GPU call.  Skipped on a one-GPU box (the host-side sharding and
merge logic is covered on CPU by tests/test_parallel_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_equals_single_gpu_under_torchrun():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "scripts", "check_sharded.py")]  # fmt: skip
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "sharded == single over 2 ranks" in out.stdout


def test_single_rank_sharded_calls_equal_the_plain_calls():
    # world size 1 (no process group): the sharded entry points degenerate to the plain calls
    import numpy as np

    from cityseer_b200 import parallel, synth

    ns, _ = synth.config("cfg4", 0.05)
    a = parallel.centrality_shortest_sharded(ns, distances=[400, 800])
    b = ns.centrality_shortest(distances=[400, 800], pbar_disabled=True)
    assert np.array_equal(a._out[0], b._out[0])
    np.testing.assert_allclose(a._out, b._out, rtol=1e-12, atol=1e-12)
    s1 = parallel.segment_centrality_sharded(ns, distances=[400])
    s2 = ns.segment_centrality(distances=[400], pbar_disabled=True)
    np.testing.assert_allclose(s1._out, s2._out, rtol=1e-12, atol=1e-9)
