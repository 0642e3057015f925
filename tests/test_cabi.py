"""The C-ABI library loads on a CPU-only box and exports every symbol ``include/cityseer_b200.h`` declares; compute entry
points fail loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers as H
from cityseer_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__

    __graft_entry__.build()
    return _native.load_library()


def header_functions():
    text = open(os.path.join(ROOT, "include", "cityseer_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cs_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = header_functions()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_native.SIGNATURES) == names


def test_stats_struct_layout_matches_header():
    # cs_stats: 6 x u64 + 16 x u64 + 2 x f32 + 2 x u32 + 8 x u64 (phase cycles) + u64 (fallback sources) + 6 x u32 (layout, kernel)
    assert ctypes.sizeof(_native.CsStats) == 6 * 8 + 16 * 8 + 2 * 4 + 2 * 4 + 8 * 8 + 8 + 6 * 4
    text = open(os.path.join(ROOT, "include", "cityseer_b200.h")).read()
    assert "#define CS_MAX_THRESHOLDS 16" in text and _native.MAX_THRESHOLDS == 16


def test_no_cpu_fallback(lib):
    if lib.cs_device_count() > 0:
        pytest.skip("a GPU is present")
    _g, _n, _e, ns = H.primal_ns()
    for call in (
        lambda: ns.centrality_shortest(distances=[400]),
        lambda: ns.segment_centrality(distances=[400]),
    ):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    f = ns.frozen()
    h = lib.cs_graph_create(
        f.node_bound, f.node_exists.ctypes.data_as(_native._u8p), f.live.ctypes.data_as(_native._u8p),
        f.weight.ctypes.data_as(_native._f32p), f.xs.ctypes.data_as(_native._f64p), f.ys.ctypes.data_as(_native._f64p),
        f.z.ctypes.data_as(_native._f64p), f.edge_bound,
        f.edge_exists.ctypes.data_as(_native._u8p), f.src.ctypes.data_as(_native._u32p), f.dst.ctypes.data_as(_native._u32p),
        f.edge_idx.ctypes.data_as(_native._u32p), f.length.ctypes.data_as(_native._f32p),
        f.angle_sum.ctypes.data_as(_native._f32p), f.imp.ctypes.data_as(_native._f32p),
        f.seconds.ctypes.data_as(_native._f32p), f.shared_key.ctypes.data_as(_native._i32p),
        f.stamp.ctypes.data_as(_native._u64p), 0, 0)  # fmt: skip
    assert not h
    assert b"no CUDA device" in lib.cs_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cityseer_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.lower() or fn == "graph.py", f"{fn} mentions the oracle"
    # graph.py mentions it only in a docstring; make sure there is no import
    g = open(os.path.join(pkg, "rustalgos", "graph.py")).read()
    assert "import oracle" not in g and "from oracle" not in g
