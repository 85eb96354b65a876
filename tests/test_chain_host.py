"""CPU tier of gpu_stitching's multi-GPU chain mode (PANO_MODE=chain: one host process, several devices; SURVEY 8e2 /
8e3).  host/chain_multi_gpu.hpp - the very code gpu_stitching.cpp instantiates with CUDA memory - is instantiated with
host memory and linked against a CPU stand-in of the C ABI (tests/hostsim/abi_standin.cpp, on the oracle): pair
sharding over the worker threads, composition of the homographies, canvas geometry and band tiling must give the
oracle's chain panorama for every device count; a pair that fails ends the chain.  The same executable on a real B200
is checked in tests/test_zz2_chain_cli_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT, load_synth


@pytest.fixture(scope="module")
def chain_host():
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libchain_host.so")
    oracle_so = os.path.join(ROOT, "oracle", "libpano_oracle.so")
    if not os.path.exists(oracle_so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libpano_oracle.so"], stdout=subprocess.DEVNULL)
    srcs = [os.path.join(d, "chain_host.cpp"), os.path.join(d, "abi_standin.cpp"),
            os.path.join(ROOT, PKG, "host", "chain_multi_gpu.hpp"), os.path.join(ROOT, "include", "pano_b200.h"), oracle_so]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
                               "-o", so, srcs[0], srcs[1], "-L" + os.path.join(ROOT, "oracle"), "-l:libpano_oracle.so",
                               "-Wl,-rpath,$ORIGIN/../../oracle", "-lpthread"])
    return C.CDLL(so)


def run_chain(lib, images, n_dev, seed=12345):
    n = len(images)
    imgs = [np.ascontiguousarray(im) for im in images]
    ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
    ws = (C.c_int * n)(*[im.shape[1] for im in imgs])
    hs = (C.c_int * n)(*[im.shape[0] for im in imgs])
    cap = 64 << 20
    canvas = np.zeros(cap, np.uint8)
    geom = (C.c_int * 3)()
    st_pairs = (C.c_int * (n - 1))()
    dev_pairs = (C.c_int * (n - 1))()
    st = lib.hs_chain_multi(ptrs, ws, hs, n, n_dev, C.c_uint32(seed), canvas.ctypes.data_as(C.c_void_p), C.c_size_t(cap), geom,
                            st_pairs, dev_pairs)
    w, h, used = geom[0], geom[1], geom[2]
    return st, (canvas[:w * h * 3].reshape(h, w, 3) if st == 0 else None), used, list(st_pairs), list(dev_pairs)


@pytest.fixture(scope="module")
def strip():
    return load_synth().make_strip(n=4, w=320, h=200, seed=21)


@pytest.mark.parametrize("n_dev", [1, 2, 3, 5])
def test_chain_multi_device_equals_oracle_chain(chain_host, oracle, strip, n_dev):
    pano, pair_H = oracle.stitch_chain(strip, seed=12345)
    assert all(H is not None for H in pair_H)
    st, canvas, used, st_pairs, dev_pairs = run_chain(chain_host, strip, n_dev)
    assert st == 0 and used == len(strip) and st_pairs == [0] * (len(strip) - 1)
    assert dev_pairs == [i % n_dev for i in range(len(strip) - 1)]        # pair i on device i mod D
    assert canvas.shape == pano.shape and np.array_equal(canvas, pano)    # bands tile the oracle's canvas exactly


def test_chain_two_images_is_the_references_pair(chain_host, oracle, strip):
    """for two images chain mode is stitchTwoImages (ref: src/serial/main.cpp:311-391)"""
    o = oracle.stitch_pair(strip[0], strip[1], seed=7)
    st, canvas, used, _, _ = run_chain(chain_host, strip[:2], 2, seed=7)
    assert o["status"] == 1 and st == 0 and used == 2 and np.array_equal(canvas, o["canvas"])


def test_chain_broken_pair_ends_the_chain(chain_host, oracle, strip):
    flat = np.full_like(strip[2], 90)                      # no corners: pair (1, 2) has no matches
    images = [strip[0], strip[1], flat, strip[3]]
    st, canvas, used, st_pairs, _ = run_chain(chain_host, images, 3)
    pano, pair_H = oracle.stitch_chain(images, seed=12345)
    assert st == 0 and used == 2 and st_pairs[0] == 0 and st_pairs[1] != 0
    assert pair_H[1] is None and np.array_equal(canvas, pano)


def test_band_rows_partition():
    """every canvas row belongs to exactly one worker, also when the rows do not divide evenly"""
    dist = __import__("importlib").import_module(PKG + ".dist")
    for ch in (1, 7, 200, 1501):
        for world in (1, 2, 3, 8):
            rows = [dist.band_rows(ch, r, world) for r in range(world)]
            assert rows[0][0] == 0 and sum(b for _, b in rows) == ch
            assert all(rows[r][0] + rows[r][1] == rows[r + 1][0] for r in range(world - 1))
