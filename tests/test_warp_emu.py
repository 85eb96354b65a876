"""CPU tier: the warp stage's device code - csrc/warp_kernels.cuh, compiled UNCHANGED by g++ on the CPU emulation of
the CUDA execution model (tests/hostsim/cuda_emu.hpp) - against the oracle and against real cv2 output, byte for byte
(SURVEY 8 rows a11-a13).  Covered: the production quad kernel (four canvas pixels per thread, word stores, checked
Newton reciprocal, DP4A bilinear), the round-1 fast kernel, the general kernel (any matrix, unaligned sources), the
footprint box / fast-path admission on the host, the accumulate mode of chain bands and the row packer.  The same
kernels run on a B200 in tests/test_gpu_parity.py; this tier needs no GPU."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_synth

QUAD, FAST1, GENERAL = 0, 1, 2


def p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def overlay(lib, left, right, H, kernel, pitch_align=256):
    left, right = np.ascontiguousarray(left), np.ascontiguousarray(right)
    H = np.ascontiguousarray(H, np.float64)
    cap = 64 << 20
    canvas = np.zeros(cap, np.uint8)
    geom = (C.c_int * 5)()
    st = lib.wemu_overlay(p(left, C.c_uint8), left.shape[1], left.shape[0], C.c_size_t(left.strides[0]),
                          p(right, C.c_uint8), right.shape[1], right.shape[0], C.c_size_t(right.strides[0]),
                          p(H, C.c_double), kernel, pitch_align, p(canvas, C.c_uint8), C.c_size_t(cap), geom)
    assert st >= 0, (st, lib.wemu_last_error())
    if st == 0:
        return None, False
    cw, ch = geom[0], geom[1]
    return canvas[:cw * ch * 3].reshape(ch, cw, 3), bool(geom[4])


@pytest.fixture(scope="module")
def pair(oracle):
    left, right, Ht = load_synth().make_pair(320, 200, seed=9)
    H = oracle.pair_homography(left, right, seed=12345)
    assert H is not None
    return left, right, H


@pytest.mark.parametrize("kernel", [QUAD, FAST1, GENERAL])
def test_emulated_pair_canvas_equals_oracle(warp_emu, oracle, pair, kernel):
    left, right, H = pair
    canvas, fast = overlay(warp_emu, left, right, H, kernel)
    assert fast == (kernel != GENERAL)                       # the pair's matrix is admitted to the fast path
    assert np.array_equal(canvas, oracle.compose(left, right, H))


def test_emulated_unaligned_source_takes_the_general_kernel(warp_emu, oracle, pair):
    left, right, H = pair
    l2, r2 = np.ascontiguousarray(left[:, :317]), np.ascontiguousarray(right[:, :315])     # 3 * w not a multiple of 4
    canvas, fast = overlay(warp_emu, l2, r2, H, QUAD, pitch_align=1)
    assert not fast and np.array_equal(canvas, oracle.compose(l2, r2, H))
    canvas, fast = overlay(warp_emu, l2, r2, H, QUAD)                                      # pitched like the engine's uploads
    assert fast and np.array_equal(canvas, oracle.compose(l2, r2, H))


def random_h(rng, w, h):
    a, s = rng.uniform(-0.25, 0.25), rng.uniform(0.75, 1.3)
    return np.array([[s * np.cos(a), -s * np.sin(a), rng.uniform(-0.5, 0.9) * w],
                     [s * np.sin(a), s * np.cos(a), rng.uniform(-0.4, 0.4) * h],
                     [rng.uniform(-4e-4, 4e-4), rng.uniform(-4e-4, 4e-4), 1.0]])


@pytest.mark.parametrize("kernel", [QUAD, FAST1, GENERAL])
def test_emulated_random_homographies_equal_oracle(warp_emu, oracle, kernel):
    rng = np.random.default_rng(11 + kernel)
    n_fast = 0
    for case in range(12):
        wl, hl, wr, hr = (int(v) for v in rng.integers(20, 150, 4))
        left = rng.integers(0, 256, (hl, wl, 3), dtype=np.uint8)
        right = rng.integers(0, 256, (hr, wr, 3), dtype=np.uint8)
        if case % 3 == 0:
            right[rng.integers(0, hr):, :] = 0               # black regions: the overlay keeps the left image there
        H = random_h(rng, wl, hl)
        want = oracle.compose(left, right, H)
        canvas, fast = overlay(warp_emu, left, right, H, kernel)
        if want is None:
            assert canvas is None
            continue
        n_fast += fast
        assert canvas.shape == want.shape and np.array_equal(canvas, want), case
    assert kernel == GENERAL or n_fast >= 3


@pytest.mark.parametrize("kernel", [QUAD, FAST1, GENERAL])
def test_emulated_warp_perspective_equals_cv2_fixtures(warp_emu, pins, kernel):
    """the committed cv2.warpPerspective outputs (tests/golden/opencv_pins.npz)"""
    src = np.ascontiguousarray(pins["warp_src"])
    for i in range(4):
        ref = pins["warp_out%d" % i]
        M = np.ascontiguousarray(pins["warp_M%d" % i])
        out = np.zeros_like(ref)
        st = warp_emu.wemu_warp_perspective(p(src, C.c_uint8), src.shape[1], src.shape[0], C.c_size_t(src.strides[0]),
                                            p(M, C.c_double), kernel, p(out, C.c_uint8), ref.shape[1], ref.shape[0])
        assert st == 1 and np.array_equal(out, ref)


def test_emulated_warp_perspective_equals_live_cv2(warp_emu):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for case in range(8):
        w, h, dw, dh = (int(v) for v in rng.integers(16, 180, 4))
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        M = random_h(rng, w, h)
        want = cv2.warpPerspective(src, M, (dw, dh))
        out = np.zeros_like(want)
        for kernel in (QUAD, GENERAL):
            st = warp_emu.wemu_warp_perspective(p(src, C.c_uint8), w, h, C.c_size_t(src.strides[0]), p(M, C.c_double), kernel,
                                                p(out, C.c_uint8), dw, dh)
            assert st == 1 and np.array_equal(out, want), (case, kernel)


@pytest.mark.parametrize("kernel", [QUAD, GENERAL])
def test_emulated_band_accumulation_tiles_the_canvas(warp_emu, oracle, pair, kernel):
    """chain mode: rendering a canvas band by band gives the rows of the whole canvas, and non-black warped
    pixels overwrite what the band already holds"""
    left, right, H = pair
    ok, (cw, ch, x0, y0), T = oracle.canvas_geometry(320, 200, 320, 200, H)
    assert ok
    M0 = np.array([[1, 0, x0], [0, 1, y0], [0, 0, 1.0]])
    M1 = np.ascontiguousarray(oracle.mul33(T, H))
    whole = np.zeros((ch, cw, 3), np.uint8)
    for im, M in ((left, M0), (right, M1)):
        w = oracle.warp_perspective(im, M, (cw, ch))
        nz = w.any(axis=2)
        whole[nz] = w[nz]
    bands = [(0, 37), (37, ch - 37 - 50), (ch - 50, 50)]
    got = np.zeros_like(whole)
    for yb, bh in bands:
        band = np.zeros((bh, cw, 3), np.uint8)
        for im, M in ((left, M0), (right, M1)):
            im = np.ascontiguousarray(im)
            Mc = np.ascontiguousarray(M, np.float64)
            st = warp_emu.wemu_accumulate(p(im, C.c_uint8), 320, 200, C.c_size_t(im.strides[0]), p(Mc, C.c_double), kernel,
                                          p(band, C.c_uint8), cw, ch, yb, bh)
            assert st == 1
        got[yb:yb + bh] = band
    assert np.array_equal(got, whole)
