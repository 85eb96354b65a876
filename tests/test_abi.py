"""CPU tier: the C-ABI library loads and exports every symbol include/pano_b200.h declares;
without a GPU it refuses to create a context (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, load_pkg


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pano_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pano_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    pkg = load_pkg()
    if not os.path.exists(pkg.LIB_PATH):
        pkg.build()
    lib = C.CDLL(pkg.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    assert sorted(pkg.EXPORTED_SYMBOLS) == syms


def test_struct_layouts_match_header():
    pkg = load_pkg()
    assert C.sizeof(pkg.HarrisCornerOptions) == 32
    assert C.sizeof(pkg.RansacOptions) == 16
    assert C.sizeof(pkg.CanvasInfo) == 16 + 72
    assert C.sizeof(pkg.PairResult) == 24 + 72 + 88 + 20 + 4  # padded to 8
    assert pkg.MATCH_DTYPE.itemsize == 12


def test_defaults_are_the_reference_defaults():
    """ref: src/serial/main.cpp:28-40, :428-435"""
    pkg = load_pkg()
    lib = pkg.load_library()
    h, r = pkg.HarrisCornerOptions(0, 0, 0, 0, 0), pkg.RansacOptions(0, 0, 0)
    lib.pano_default_harris_opts(C.byref(h)); lib.pano_default_ransac_opts(C.byref(r))
    assert (h.k_, h.nmsThresh_, h.nmsNeighborhood_, h.patchSize_, h.maxSSDThresh_) == (0.04, 1e6, 3, 5, 1e8)
    assert (r.numIterations_, r.numSamples_, r.distanceThreshold_) == (1000, 4, 3.0)


def test_no_gpu_means_no_context():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pkg = load_pkg()
    with pytest.raises(pkg.PanoError) as e:
        pkg.Engine()
    assert e.value.status == pkg.PANO_ERR_NO_DEVICE


def test_canvas_geometry_is_host_side_and_matches_oracle(oracle):
    import numpy as np
    pkg = load_pkg()
    lib = pkg.load_library()
    H = np.array([[1.0, 0.0, 2244.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    info = pkg.CanvasInfo()
    st = lib.pano_canvas_geometry(4156, 3117, 4156, 3117, H.ctypes.data_as(C.c_void_p), C.byref(info))
    assert st == 0 and (info.canvas_w, info.canvas_h, info.left_x, info.left_y) == (6400, 3117, 0, 0)
    ok, g, TH = oracle.canvas_geometry(4156, 3117, 4156, 3117, H)
    assert ok and g == (6400, 3117, 0, 0)


def test_replay_work_estimate_is_host_side_and_consistent_with_the_plan():
    """pano_replay_work_estimate (measurement aid for bench.py's replay floor): planning only, callable without a GPU;
    the numbers follow the plan's definition (ref: the shuffle being replayed is src/serial/main.cpp:264-275)"""
    pkg = load_pkg()
    lib = pkg.load_library()
    e = pkg.Engine.__new__(pkg.Engine)
    e.lib = lib
    for m, target in ((10833, 0.0), (10833, 4000.0), (37, 0.0), (4, 0.0), (70000, 16000.0)):
        w = e.replayWork(m, 1000, target)
        assert w["chunks"] == -(-1000 // w["chunk_iterations"]) and w["chunk_iterations"] >= 1
        assert w["diagonals_per_chunk"] >= w["candidates_per_chunk"] >= w["chunk_iterations"]
        assert w["diagonals_per_chunk"] % 32 == 0
        assert w["cells"] == w["chunks"] * w["diagonals_per_chunk"] * (-(-w["steps"] // 32) * 32)
        assert w["candidates_per_chunk"] <= (target or 50000.0) + 1 or w["chunk_iterations"] == 8
    # libstdc++ draws two swap positions per engine output while the range product fits 32 bits
    assert e.replayWork(10833, 1000)["steps"] == 10833 // 2 and e.replayWork(70000, 1000)["steps"] == 69999
    # smaller chunks: more sequential phases, less speculative work
    a, b = e.replayWork(10833, 1000, 0.0), e.replayWork(10833, 1000, 4000.0)
    assert b["chunks"] > a["chunks"] and b["cells"] < a["cells"]
    for bad in ((3, 1000, 0.0), (100, 0, 0.0), (100, 1000, 10.0), (100, 1000, 1e6)):
        with pytest.raises(pkg.PanoError) as ex:
            e.replayWork(*bad)
        assert ex.value.status == pkg.PANO_ERR_INVALID
