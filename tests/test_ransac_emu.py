"""CPU tier: the RANSAC stage's device code - csrc/ransac_kernels.cuh, compiled UNCHANGED by g++ on the CPU emulation of
the CUDA execution model (tests/hostsim/cuda_emu.hpp) - against the oracle, i.e. against real libstdc++ std::shuffle /
std::mt19937 and the cv2-pinned findHomography, bit for bit.  Covered: the mt19937 stream kernel,
the shuffle replay in its chunked form (rejection cells, candidate walks, chain, segment replay, sample combination; one
and many chunks, a deliberately missed speculation window and its wider re-run) and in its resident one-CTA form, the
point builder with its index check, the warp-per-hypothesis DLT, scoring, selection and the inlier mask.  The flow in
tests/hostsim/ransac_emu.cpp mirrors ransac.cu's host code.  The same kernels run on a B200 in tests/test_gpu_parity.py."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_synth

MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("distance", "<f4")])
CHUNKED, RESIDENT = 0, 1


def p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def ransac(lib, kp1, kp2, m, iters=100, thr=3.0, seed=12345, mode=CHUNKED, target=50000.0, z=4.2, scale=1):
    kp1 = np.ascontiguousarray(kp1, np.int32).reshape(-1, 2)
    kp2 = np.ascontiguousarray(kp2, np.int32).reshape(-1, 2)
    m = np.ascontiguousarray(m, MATCH_DTYPE)
    H = np.zeros((3, 3))
    best, best_iter, errw, chunks = C.c_int(0), C.c_int(-1), C.c_int(0), C.c_int(0)
    samples = np.full((iters, 4), -1, np.int32)
    counts = np.full(iters, -2, np.int32)
    mask = np.zeros(max(len(m), 1), np.uint8)
    st = lib.remu_ransac(p(kp1, C.c_int32), len(kp1), p(kp2, C.c_int32), len(kp2), m.ctypes.data_as(C.c_void_p), len(m), iters,
                         C.c_double(thr), C.c_uint32(seed), mode, C.c_double(target), C.c_double(z), scale, p(H, C.c_double),
                         C.byref(best), C.byref(best_iter), p(samples, C.c_int32), p(counts, C.c_int32), p(mask, C.c_uint8),
                         C.byref(errw), C.byref(chunks))
    assert st != -100, lib.remu_last_error()
    return dict(status=st, H=H, best_count=best.value, best_iter=best_iter.value, samples=samples, counts=counts,
                mask=mask[:len(m)].astype(bool), errw=errw.value, chunks=chunks.value)


def same(r, o, iters):
    assert r["status"] == (0 if o["ok"] else 5)                       # PANO_OK / PANO_ERR_NO_HOMOGRAPHY
    assert np.array_equal(r["samples"], o["samples"][:iters])         # the shuffle replay: every iteration's four draws
    assert np.array_equal(r["counts"], o["counts"][:iters])           # DLT + scoring: every hypothesis' inlier count
    assert (r["best_count"], r["best_iter"]) == (o["best_count"], o["best_iter"])
    if o["ok"]:
        assert np.array_equal(bits(r["H"]), bits(o["H"]))
        assert np.array_equal(r["mask"], o["inlier_mask"])


@pytest.fixture(scope="module")
def scene(oracle):
    left, right, _ = load_synth().make_pair(480, 300, seed=5)
    kl, kr = oracle.detect(left), oracle.detect(right)
    m = oracle.match(kr, kl, right, left)
    assert len(m) > 150
    return kr, kl, m


def test_emulated_mt19937_stream(ransac_emu, oracle):
    for seed, n in ((12345, 3000), (1, 624), (267, 1500)):
        out = np.zeros(n, np.uint32)
        ransac_emu.remu_mt19937(C.c_uint32(seed), n, p(out, C.c_uint32))
        assert np.array_equal(out, oracle.mt19937(seed, n))


@pytest.mark.parametrize("mode,target", [(CHUNKED, 50000.0), (CHUNKED, 700.0), (RESIDENT, 50000.0)])
def test_emulated_ransac_equals_oracle(ransac_emu, oracle, scene, mode, target):
    kr, kl, m = scene
    iters = 120
    for seed in (12345, 7):
        o = oracle.ransac(kr, kl, m, iters=iters, seed=seed)
        r = ransac(ransac_emu, kr, kl, m, iters=iters, seed=seed, mode=mode, target=target)
        same(r, o, iters)
    if mode == CHUNKED and target < 1000:
        assert r["chunks"] >= 2                                        # the small chunk target really made several chunks


@pytest.mark.parametrize("count", [4, 5, 8, 33, 64, 97])
def test_emulated_ransac_small_match_counts(ransac_emu, oracle, scene, count):
    kr, kl, m = scene
    for mode in (CHUNKED, RESIDENT):
        o = oracle.ransac(kr, kl, m[:count], iters=50, seed=3)
        r = ransac(ransac_emu, kr, kl, m[:count], iters=50, seed=3, mode=mode)
        same(r, o, 50)


def synthetic_matches(n, seed):
    rng = np.random.default_rng(seed)
    kp1 = rng.integers(0, 4000, (n, 2)).astype(np.int32)
    kp2 = (kp1 + rng.integers(-2, 3, (n, 2))).astype(np.int32)
    m = np.zeros(n, MATCH_DTYPE)
    m["queryIdx"] = np.arange(n)
    m["trainIdx"] = np.arange(n)
    return kp1, kp2, m


def test_emulated_missed_window_is_reported_and_the_wider_rerun_is_exact(ransac_emu, oracle):
    """ransac_retry's contract: a speculation window that misses the true offset is never guessed around - the status
    says so (bit 0) and the re-run with wider windows reproduces the reference.  20 000 matches: ~300 rejected draws
    per shuffle, so that a window of +-0.15 sigma around the expected offset does miss."""
    kp1, kp2, m = synthetic_matches(20000, 2)
    iters = 12
    o = oracle.ransac(kp1, kp2, m, iters=iters, seed=12345)
    r = ransac(ransac_emu, kp1, kp2, m, iters=iters, seed=12345, z=0.15)
    assert r["status"] < 0 and (-r["status"]) & 1                      # window miss flagged
    scale = 2
    while True:
        r = ransac(ransac_emu, kp1, kp2, m, iters=iters, seed=12345, z=0.15, scale=scale)
        if r["status"] >= 0 or scale >= 256:
            break
        scale *= 2
    same(r, o, iters)
    # and with the product's window (4.2 sigma) the first attempt is exact, chunked and resident
    same(ransac(ransac_emu, kp1, kp2, m, iters=iters, seed=12345), o, iters)
    same(ransac(ransac_emu, kp1, kp2, m, iters=iters, seed=12345, mode=RESIDENT), o, iters)


def test_emulated_threshold_and_bad_indices(ransac_emu, oracle, scene):
    kr, kl, m = scene
    for thr in (0.5, 1.0, 10.0):
        o = oracle.ransac(kr, kl, m, iters=40, thr=thr, seed=9)
        same(ransac(ransac_emu, kr, kl, m, iters=40, thr=thr, seed=9), o, 40)
    bad = m.copy()
    bad["trainIdx"][7] = len(kl) + 5                                   # build_points_kernel flags it in the error word
    r = ransac(ransac_emu, kr, kl, bad, iters=10)
    assert r["errw"] & 4 and r["status"] == 2


def test_emulated_ransac_large_shuffle_uses_single_draws(ransac_emu, oracle):
    """libstdc++'s std::shuffle draws two positions from one engine output only while n * n fits the engine's range
    (n <= 65535); above that every swap costs one output.  Few iterations, synthetic points."""
    kp1, kp2, m = synthetic_matches(65600, 1)
    o = oracle.ransac(kp1, kp2, m, iters=3, seed=5)
    r = ransac(ransac_emu, kp1, kp2, m, iters=3, seed=5)
    same(r, o, 3)
