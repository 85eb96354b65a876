"""CPU tier: the match stage's device code - csrc/match_kernels.cuh (in-border flags, warp-per-keypoint descriptor
gather with its shuffle-reduced norms, SIMT SSD matcher, match emission with and without the threshold filter, the
incremental fold's keypoint carry-over) plus the flag compaction - compiled UNCHANGED by g++ on the CPU emulation of the
CUDA execution model (tests/hostsim/cuda_emu.hpp) against the oracle, record for record.  The
descriptor gather and the emission are the kernels the tensor-core matcher runs between, too; the tensor-core GEMM
itself has no CPU tier (tests/test_gpu_parity.py compares it with this SIMT kernel on a B200)."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_synth

MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("distance", "<f4")])


def p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def match(lib, kq, kt, imq, imt, patch=5, max_ssd=1e8, offset=0, splits=0, norms=False):
    kq = np.ascontiguousarray(kq, np.int32).reshape(-1, 2)
    kt = np.ascontiguousarray(kt, np.int32).reshape(-1, 2)
    imq, imt = np.ascontiguousarray(imq), np.ascontiguousarray(imt)
    out = np.zeros(max(len(kq), 1), MATCH_DTYPE)
    nrm = np.zeros(max(len(kq), 1), np.uint32)
    n = lib.memu_match(p(kq, C.c_int32), len(kq), p(kt, C.c_int32), len(kt), p(imq, C.c_uint8), imq.shape[1], imq.shape[0],
                       C.c_size_t(imq.strides[0]), p(imt, C.c_uint8), imt.shape[1], imt.shape[0], C.c_size_t(imt.strides[0]),
                       patch, C.c_double(max_ssd), offset, splits, out.ctypes.data_as(C.c_void_p), len(out),
                       p(nrm, C.c_uint32) if norms else None)
    assert n >= 0, (n, lib.memu_last_error())
    return (out[:n], nrm) if norms else out[:n]


@pytest.fixture(scope="module")
def scene(oracle):
    left, right, _ = load_synth().make_pair(480, 300, seed=5)
    kl, kr = oracle.detect(left), oracle.detect(right)
    kl = np.concatenate([kl, [[0, 0], [1, 299], [479, 5]], kl[:7]]).astype(np.int32)   # border points, duplicates (ties)
    kr = np.concatenate([[[2, 1]], kr, kr[3:5]]).astype(np.int32)
    return left, right, kl, kr


@pytest.mark.parametrize("splits", [0, 1, 3, 9])
def test_emulated_matcher_equals_oracle(match_emu, oracle, scene, splits):
    left, right, kl, kr = scene
    m = match(match_emu, kr, kl, right, left, splits=splits)
    mo = oracle.match(kr, kl, right, left)
    assert len(mo) > 100 and m.tobytes() == np.ascontiguousarray(mo).tobytes()


def test_emulated_matcher_threshold_offset_and_patch_sizes(match_emu, oracle, scene):
    left, right, kl, kr = scene
    for patch, max_ssd, offset in ((5, 3000.0, 0), (5, 1e8, 17), (3, 1e8, 0), (3, 500.0, 5), (1, 1e8, 0), (5, 0.0, 0)):
        m = match(match_emu, kr, kl, right, left, patch=patch, max_ssd=max_ssd, offset=offset)
        mo = oracle.match(kr, kl, right, left, patch=patch, max_ssd=max_ssd, offset=offset)
        assert m.tobytes() == np.ascontiguousarray(mo).tobytes(), (patch, max_ssd, offset)
    assert len(match(match_emu, kr, kl, right, left, max_ssd=3000.0)) < len(match(match_emu, kr, kl, right, left))


def test_emulated_matcher_edge_cases(match_emu, oracle, scene):
    left, right, kl, kr = scene
    empty = np.zeros((0, 2), np.int32)
    for kq, kt in ((empty, kl), (kr, empty), (kr[:1], kl), (kr, kl[-10:]), (kr[:129], kl[:65]), (kr[:33], kl[:1])):
        m = match(match_emu, kq, kt, right, left)
        mo = oracle.match(kq, kt, right, left)
        assert m.tobytes() == np.ascontiguousarray(mo).tobytes()
    flat = np.full((40, 50, 3), 9, np.uint8)       # every SSD is 0: the first train keypoint wins every query
    k = np.array([[x, y] for y in range(4, 36, 5) for x in range(4, 46, 6)], np.int32)
    m = match(match_emu, k, k, flat, flat)
    assert len(m) == len(k) and (m["trainIdx"] == 0).all() and (m["distance"] == 0).all()


def test_emulated_descriptor_norms(match_emu, scene):
    """gather_desc_kernel's warp-shuffle reduction: squared norm of every in-border query patch"""
    left, right, kl, kr = scene
    m, nrm = match(match_emu, kr, kl, right, left, norms=True)
    inb = ~((kr[:, 0] < 2) | (kr[:, 1] < 2) | (kr[:, 0] + 2 >= 480) | (kr[:, 1] + 2 >= 300))
    want = [int((right[y - 2:y + 3, x - 2:x + 3].astype(np.int64) ** 2).sum()) for x, y in kr[inb]]
    assert list(nrm[:len(want)]) == want


def test_emulated_keypoint_carry_over_of_the_incremental_fold(match_emu, oracle):
    rng = np.random.default_rng(3)
    old = rng.integers(0, 300, (50, 2)).astype(np.int32)
    old[5] = (-1, -1)                                   # a point that left the canvas earlier stays out
    new = rng.integers(0, 200, (70, 2)).astype(np.int32)
    TH = np.array([[1.01, 0.02, 150.5], [-0.015, 0.99, 12.25], [1e-5, -2e-5, 1.0]])
    out = np.zeros((120, 2), np.int32)
    n = match_emu.memu_update_keypoints(p(old, C.c_int32), 50, 7, 3, p(new, C.c_int32), 70, p(TH, C.c_double), 400, 260,
                                        p(out, C.c_int32))
    assert n == 120
    want_old = np.where(old[:, :1] < 0, -1, old + [7, 3])
    assert np.array_equal(out[:50], want_old)
    pts = oracle.perspective_transform(new.astype(np.float32), TH)      # cv::perspectiveTransform arithmetic
    r = np.rint(pts).astype(np.int32)
    inside = (r[:, 0] >= 0) & (r[:, 1] >= 0) & (r[:, 0] < 400) & (r[:, 1] < 260)
    assert np.array_equal(out[50:], np.where(inside[:, None], r, -1))
