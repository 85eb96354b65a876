"""CPU tier: the kernels' shared order-exact arithmetic (csrc/pano_core.cuh) and the shuffle
replay's windowed-speculation scheme (csrc/replay_plan.hpp), compiled for the host and run as
plain loops (tests/hostsim), against the oracle / real libstdc++."""
import ctypes as C

import numpy as np
import pytest


def p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def test_find_homography4_host_compile_matches_oracle(hostsim, oracle, pins):
    for src, dst, H, ok in zip(pins["fh_src"], pins["fh_dst"], pins["fh_H"], pins["fh_ok"]):
        out = np.zeros((3, 3))
        r = hostsim.hs_find_homography4(p(np.ascontiguousarray(src), C.c_float), p(np.ascontiguousarray(dst), C.c_float),
                                        p(out, C.c_double))
        assert bool(r) == bool(ok)
        if ok:
            assert np.array_equal(out.view(np.uint64), H.view(np.uint64))


def test_warp_model_matches_cv2_fixtures(hostsim, pins):
    src = np.ascontiguousarray(pins["warp_src"])
    for i in range(4):
        ref = pins["warp_out%d" % i]
        M = np.ascontiguousarray(pins["warp_M%d" % i])
        out = np.zeros_like(ref)
        hostsim.hs_warp_perspective(p(src, C.c_uint8), src.shape[1], src.shape[0], C.c_size_t(src.strides[0]),
                                    p(M, C.c_double), p(out, C.c_uint8), ref.shape[1], ref.shape[0],
                                    C.c_size_t(out.strides[0]))
        assert np.array_equal(out, ref)


def test_canvas_geometry_matches_oracle(hostsim, oracle):
    rng = np.random.default_rng(3)
    for t in range(200):
        H = np.array([[1 + rng.normal() * 0.01, rng.normal() * 0.01, rng.uniform(-500, 2500)],
                      [rng.normal() * 0.01, 1 + rng.normal() * 0.01, rng.uniform(-200, 200)],
                      [rng.normal() * 1e-6, rng.normal() * 1e-6, 1.0]])
        geom = np.zeros(6, np.int32); TH = np.zeros((3, 3)); Mi = np.zeros((3, 3))
        hostsim.hs_canvas_geometry(1000, 700, 900, 650, p(H, C.c_double), p(geom, C.c_int32), p(TH, C.c_double),
                                   p(Mi, C.c_double))
        ok, g, TH2 = oracle.canvas_geometry(1000, 700, 900, 650, H)
        assert bool(geom[5]) == ok
        assert tuple(int(v) for v in geom[:4]) == g
        assert np.array_equal(TH.view(np.uint64), TH2.view(np.uint64))


@pytest.mark.parametrize("n,iters", [(4, 64), (5, 64), (6, 64), (7, 64), (8, 64), (9, 33), (10, 100), (33, 100),
                                     (256, 300), (257, 300), (1000, 1000), (2500, 1000), (4097, 600),
                                     (65535, 6), (65536, 6), (70000, 12)])
def test_replay_matches_std_shuffle(hostsim, oracle, n, iters):
    """first 4 elements of every iteration's shuffled copy == real std::shuffle with one
    continuing mt19937 (both regimes: paired draws for n <= 65535, single draws above)."""
    ref = oracle.shuffle_iota(12345, n, reps=iters)
    samples = np.zeros((iters, 4), np.int32)
    end = C.c_uint64(0)
    st = hostsim.hs_replay(C.c_uint32(12345), C.c_uint32(n), iters, 1, p(samples, C.c_int32), C.byref(end), None)
    assert st == 0
    assert np.array_equal(samples, ref)


def test_replay_other_seeds(hostsim, oracle):
    for seed in (0, 1, 267, 0xFFFFFFFF):
        ref = oracle.shuffle_iota(seed, 3001, reps=200)
        samples = np.zeros((200, 4), np.int32)
        st = hostsim.hs_replay(C.c_uint32(seed), C.c_uint32(3001), 200, 1, p(samples, C.c_int32), None, None)
        assert st == 0 and np.array_equal(samples, ref)


def test_replay_total_draws_match_oracle_ransac(hostsim, oracle, small_pair):
    left, right, _ = small_pair
    kl, kr = oracle.detect(left), oracle.detect(right)
    m = oracle.match(kr, kl, right, left)
    r = oracle.ransac(kr, kl, m, seed=12345)
    samples = np.zeros((1000, 4), np.int32)
    end = C.c_uint64(0)
    st = hostsim.hs_replay(C.c_uint32(12345), C.c_uint32(len(m)), 1000, 1, p(samples, C.c_int32), C.byref(end), None)
    assert st == 0 and end.value == r["draws"]
    assert np.array_equal(samples, r["samples"])


@pytest.mark.parametrize("n,iters", [(4, 64), (5, 64), (6, 64), (7, 64), (8, 64), (9, 33), (10, 100), (33, 100),
                                     (256, 300), (257, 300), (1000, 1000), (2500, 1000), (4097, 600),
                                     (10774, 300), (10837, 200), (12288, 60)])
def test_resident_replay_matches_std_shuffle(hostsim, oracle, n, iters):
    """replay_resident_kernel's formulation (banded rejection cells from the exact start of every
    iteration, segment walks, chain, max-element tracking) == real std::shuffle."""
    ref = oracle.shuffle_iota(12345, n, reps=iters)
    samples = np.zeros((iters, 4), np.int32)
    end = C.c_uint64(0)
    end2 = C.c_uint64(0)
    stats = np.zeros(8)
    st = hostsim.hs_replay_resident(C.c_uint32(12345), C.c_uint32(n), iters, 1, C.c_double(0.0), p(samples, C.c_int32),
                                    C.byref(end), p(stats, C.c_double))
    assert st == 0, (st, stats)
    assert np.array_equal(samples, ref)
    # same total number of engine outputs as the chunked formulation
    s2 = np.zeros((iters, 4), np.int32)
    assert hostsim.hs_replay(C.c_uint32(12345), C.c_uint32(n), iters, 1, p(s2, C.c_int32), C.byref(end2), None) == 0
    assert end.value == end2.value and np.array_equal(s2, samples)


def test_resident_replay_band_miss_is_detected(hostsim, oracle):
    """a band that is too narrow must be reported (status 1), never silently produce samples;
    doubling the margins (window_scale) recovers the exact result."""
    n, iters = 9000, 400
    samples = np.zeros((iters, 4), np.int32)
    st = hostsim.hs_replay_resident(C.c_uint32(7), C.c_uint32(n), iters, 1, C.c_double(0.6), p(samples, C.c_int32), None, None)
    assert st == 1
    for scale in (2, 4, 8, 16):
        st = hostsim.hs_replay_resident(C.c_uint32(7), C.c_uint32(n), iters, scale, C.c_double(0.6), p(samples, C.c_int32), None, None)
        if st == 0:
            break
    assert st == 0
    assert np.array_equal(samples, oracle.shuffle_iota(7, n, reps=iters))


def test_resident_plan_limits(hostsim):
    """the single-draw regime (n > 65535), shuffles longer than the CTA's 16 x 12 blocks of 32 steps and bands wider
    than a byte of diagonals stay on the chunked path"""
    samples = np.zeros((4, 4), np.int32)
    for n in (70000, 40000, 20001, 12290):
        assert hostsim.hs_replay_resident(C.c_uint32(1), C.c_uint32(n), 4, 1, C.c_double(0.0), p(samples, C.c_int32), None, None) == 32


def test_inlier_limit_is_exact_sqrt_threshold(hostsim):
    """inlier_d2_limit(thr): d2 < limit  <=>  sqrt(d2) < thr for the doubles around the limit (score kernel's
    square-root-free predicate)"""
    import struct
    hostsim.hs_inlier_d2_limit.restype = C.c_double
    for thr in (3.0, 1.0, 0.5, 2.9999999, 1e-3, 7.25, 1e6):
        lim = hostsim.hs_inlier_d2_limit(C.c_double(thr))
        bits_ = struct.unpack("<Q", struct.pack("<d", lim))[0]
        for k in range(-3, 4):
            x = struct.unpack("<d", struct.pack("<Q", bits_ + k))[0]
            assert (np.sqrt(np.float64(x)) < thr) == (x < lim), (thr, k)
    assert hostsim.hs_inlier_d2_limit(C.c_double(0.0)) == 0.0 and hostsim.hs_inlier_d2_limit(C.c_double(-1.0)) == 0.0
    assert np.isnan(hostsim.hs_inlier_d2_limit(C.c_double(float("nan"))))


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("seed_mode", [0, 1, 2, 3])
def test_warp_fast_coordinates_never_differ_from_exact(hostsim, seed_mode, variant):
    """warp.cu's fast coordinate path (checked Newton reciprocal + magic rounding, pano_core.cuh warp_coord_fast) either
    reproduces OpenCV's exact  rint((X0 + M0 x1) * (32 / W))  or asks for the exact path - for a good seed, a poor
    seed and a useless seed (which must send every pixel to the exact path), over whole canvases"""
    rng = np.random.default_rng(5 + seed_mode)
    out = np.zeros(3, np.uint64)
    total_need = total = 0
    mats = [np.array([[1.0046794925793106, -0.009686860045644952, 1923.8011703686257],
                      [0.006395609059110648, 0.999516018010321, 0.0],
                      [8.627383134815669e-07, -1.129114095386824e-06, 1.0]])]
    for _ in range(6):
        mats.append(np.array([[1 + rng.normal() * 0.05, rng.normal() * 0.05, rng.uniform(0, 2500)],
                              [rng.normal() * 0.05, 1 + rng.normal() * 0.05, rng.uniform(0, 300)],
                              [rng.normal() * 3e-6, rng.normal() * 3e-6, 1.0]]))
    for H in mats:
        Minv = np.ascontiguousarray(np.linalg.inv(H))
        hostsim.hs_warp_fast_check(p(Minv, C.c_double), 5763, 2182, 7, seed_mode, p(out, C.c_uint64), variant)
        assert out[0] == 0, (seed_mode, H, out)
        total_need += int(out[1]); total += int(out[2])
    if seed_mode == 2:
        assert total_need == total            # a useless seed must never be trusted
    else:
        assert total_need < total * 1e-3      # the exact path stays rare (about 6e-5 of the pixels)
