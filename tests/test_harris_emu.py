"""CPU tier: the detector's device code - csrc/harris_kernels.cuh, compiled UNCHANGED by g++ on the CPU emulation of
the CUDA execution model (tests/hostsim/cuda_emu.hpp: threads as fibers, real barriers, ballots and shuffles) - against
the oracle, bit for bit.  Covered: the production fused response + NMS kernel in both
instantiations (plain-load tile, and the TMA tile load with the TMA unit modelled as a zero-filled box copy), the
two-kernel path used for other NMS neighbourhoods, the single-block scan, the ordered scatter, the stand-alone response
plane, the generic FP64 correlation and the flag compaction the matcher uses.  The same kernels run on a B200 in
tests/test_gpu_parity.py; this tier needs no GPU."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_synth

FUSED, FUSED_TMA, TWO_KERNEL = 0, 1, 2


def p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def detect(lib, img, k=0.04, thresh=1e6, nbhd=3, path=FUSED, order=0):
    h, w = img.shape[:2]
    xy = np.zeros((w * h // 2 + 16, 2), np.int32)
    n = lib.hemu_detect(p(img, C.c_uint8), w, h, C.c_size_t(img.strides[0]), C.c_double(k), C.c_double(thresh), nbhd, path, order,
                        p(xy, C.c_int32), len(xy))
    assert n >= 0, (n, lib.hemu_last_error())
    return xy[:n]


@pytest.fixture(scope="module")
def images():
    rng = np.random.default_rng(7)
    left, right, _ = load_synth().make_pair(320, 200, seed=9)
    noise = rng.integers(0, 256, (70, 97, 3), dtype=np.uint8)           # ragged: not a multiple of any tile
    tiny = rng.integers(0, 256, (9, 11, 3), dtype=np.uint8)
    tall = rng.integers(0, 256, (131, 37, 3), dtype=np.uint8)           # two tile rows, one and a bit tile columns
    padded = np.zeros((64, 100 + 7, 3), np.uint8)                       # a view with a pitch that is not 3 * w
    padded[:] = rng.integers(0, 256, padded.shape)
    view = padded[:, :100]
    flat = np.full((50, 60, 3), 128, np.uint8)
    return {"left": left, "right": right, "noise": noise, "tiny": tiny, "tall": tall, "view": view, "flat": flat}


def test_taps_are_the_reference_gaussian(harris_emu, oracle):
    t = np.zeros(25)
    harris_emu.hemu_taps(p(t, C.c_double))
    assert np.array_equal(t.view(np.uint64), np.ascontiguousarray(oracle.gaussian_kernel(5, 1.0)).reshape(-1).view(np.uint64))


@pytest.mark.parametrize("path", [FUSED, FUSED_TMA, TWO_KERNEL])
@pytest.mark.parametrize("name", ["left", "noise", "tiny", "tall", "view", "flat"])
def test_emulated_detector_equals_oracle(harris_emu, oracle, images, name, path):
    img = images[name]
    for thresh, order in ((1e6, 0), (1e4, 2)):
        k = detect(harris_emu, img, thresh=thresh, path=path, order=order)
        ko = oracle.detect(np.ascontiguousarray(img), thresh=thresh)
        assert k.shape == ko.shape and np.array_equal(k, ko)      # same keypoints in the reference's row-major order
    if name in ("left", "noise"):
        assert len(ko) > 20
    if name == "flat":
        assert len(ko) == 0


def test_emulated_detector_other_k_and_right_image(harris_emu, oracle, images):
    for path in (FUSED, FUSED_TMA):
        k = detect(harris_emu, images["right"], k=0.06, thresh=5e5, path=path, order=1)
        assert np.array_equal(k, oracle.detect(images["right"], k=0.06, thresh=5e5))


@pytest.mark.parametrize("nbhd", [3, 5, 7])
def test_emulated_two_kernel_path_other_neighbourhoods(harris_emu, oracle, images, nbhd):
    k = detect(harris_emu, images["left"], nbhd=nbhd, path=TWO_KERNEL, order=1)
    ko = oracle.detect(images["left"], nbhd=nbhd)
    assert len(ko) > 5 and np.array_equal(k, ko)


def test_emulated_response_plane_bit_exact(harris_emu, oracle, images):
    for name in ("noise", "view", "tiny"):
        img = images[name]
        h, w = img.shape[:2]
        r = np.zeros((h, w))
        assert harris_emu.hemu_response(p(img, C.c_uint8), w, h, C.c_size_t(img.strides[0]), C.c_double(0.04), p(r, C.c_double)) == 0
        ro = oracle.harris_response(np.ascontiguousarray(img))
        assert np.array_equal(r.view(np.uint64), np.ascontiguousarray(ro).view(np.uint64))


@pytest.mark.parametrize("ksize", [3, 5, 7])
def test_emulated_convolution_equals_oracle(harris_emu, oracle, ksize):
    rng = np.random.default_rng(ksize)
    plane = rng.normal(size=(41, 53)) * 1000
    kern = rng.normal(size=(ksize, ksize))
    out = np.zeros_like(plane)
    assert harris_emu.hemu_convolve(p(plane, C.c_double), 53, 41, p(kern, C.c_double), ksize, p(out, C.c_double)) == 0
    assert np.array_equal(out.view(np.uint64), np.ascontiguousarray(oracle.convolve(plane, kern)).view(np.uint64))


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 5000])
def test_emulated_flag_compaction_is_stable(harris_emu, n):
    rng = np.random.default_rng(n)
    flags = (rng.random(max(n, 1)) < 0.3).astype(np.uint8)
    out = np.full(max(n, 1), -1, np.int32)
    m = harris_emu.hemu_compact(p(flags, C.c_uint8), n, p(out, C.c_int32))
    want = np.flatnonzero(flags[:n])
    assert m == len(want) and np.array_equal(out[:m], want)
