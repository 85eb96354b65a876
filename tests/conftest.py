import ctypes
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "ucb-cs267-parallel-panoramic-image-stitching_b200"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_pkg():
    return importlib.import_module(PKG)


def load_synth():
    return importlib.import_module(PKG + ".synth")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def hostsim():
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libhostsim.so")
    src = os.path.join(d, "hostsim.cpp")
    hdrs = [os.path.join(ROOT, PKG, "csrc", f) for f in ("pano_core.cuh", "replay_plan.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src])
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def knn_emu():
    """the engine's kNN kernels (csrc/knn_kernels.cuh, unchanged) compiled by g++ on the CPU emulation of the CUDA
    execution model (tests/hostsim/cuda_emu.hpp)"""
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libknn_emu.so")
    srcs = [os.path.join(d, "knn_emu.cpp"), os.path.join(d, "cuda_emu.hpp")]
    srcs += [os.path.join(ROOT, PKG, "csrc", f) for f in ("knn_kernels.cuh", "knn_core.cuh", "pano_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-o", so, srcs[0]])
    lib = ctypes.CDLL(so)
    lib.emu_last_error.restype = ctypes.c_char_p
    return lib


@pytest.fixture(scope="session")
def harris_emu():
    """the detector's device code (csrc/harris_kernels.cuh, incl. the production fused kernel) compiled by g++ on the CPU
    emulation of the CUDA execution model (tests/hostsim/cuda_emu.hpp)"""
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libharris_emu.so")
    srcs = [os.path.join(d, "harris_emu.cpp"), os.path.join(d, "cuda_emu.hpp")]
    srcs += [os.path.join(ROOT, PKG, "csrc", f) for f in ("harris_kernels.cuh", "pano_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-o", so, srcs[0]])
    lib = ctypes.CDLL(so)
    lib.hemu_last_error.restype = ctypes.c_char_p
    return lib


@pytest.fixture(scope="session")
def warp_emu():
    """the warp stage's device code (csrc/warp_kernels.cuh, incl. the production quad kernel) compiled by g++ on the CPU
    emulation of the CUDA execution model (tests/hostsim/cuda_emu.hpp)"""
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libwarp_emu.so")
    srcs = [os.path.join(d, "warp_emu.cpp"), os.path.join(d, "cuda_emu.hpp")]
    srcs += [os.path.join(ROOT, PKG, "csrc", f) for f in ("warp_kernels.cuh", "pano_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-Wno-unused-function", "-o", so, srcs[0]])
    lib = ctypes.CDLL(so)
    lib.wemu_last_error.restype = ctypes.c_char_p
    return lib


@pytest.fixture(scope="session")
def match_emu():
    """the match stage's device code (csrc/match_kernels.cuh + the flag compaction) compiled by g++ on the CPU emulation
    of the CUDA execution model (tests/hostsim/cuda_emu.hpp)"""
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libmatch_emu.so")
    srcs = [os.path.join(d, "match_emu.cpp"), os.path.join(d, "cuda_emu.hpp")]
    srcs += [os.path.join(ROOT, PKG, "csrc", f) for f in ("match_kernels.cuh", "harris_kernels.cuh", "pano_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-Wno-unused-function", "-o", so, srcs[0]])
    lib = ctypes.CDLL(so)
    lib.memu_last_error.restype = ctypes.c_char_p
    return lib


@pytest.fixture(scope="session")
def ransac_emu():
    """the RANSAC stage's device code (csrc/ransac_kernels.cuh: shuffle replay, DLT, scoring ...) compiled by g++ on the
    CPU emulation of the CUDA execution model (tests/hostsim/cuda_emu.hpp)"""
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libransac_emu.so")
    srcs = [os.path.join(d, "ransac_emu.cpp"), os.path.join(d, "cuda_emu.hpp")]
    srcs += [os.path.join(ROOT, PKG, "csrc", f) for f in ("ransac_kernels.cuh", "replay_plan.hpp", "pano_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-Wno-unused-function", "-o", so, srcs[0]])
    lib = ctypes.CDLL(so)
    lib.remu_last_error.restype = ctypes.c_char_p
    return lib


@pytest.fixture(scope="session")
def match_tc_emu():
    """the tensor-core matcher's kernel body (csrc/match_tc_kernels.cuh) compiled by g++ on the CPU emulation of the CUDA
    execution model plus a host model of mbarriers / TMA / tcgen05 / tensor memory (tests/hostsim/tcgen05_emu.hpp)"""
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libmatch_tc_emu.so")
    srcs = [os.path.join(d, "match_tc_emu.cpp"), os.path.join(d, "cuda_emu.hpp"), os.path.join(d, "tcgen05_emu.hpp")]
    srcs += [os.path.join(ROOT, PKG, "csrc", f) for f in ("match_tc_kernels.cuh", "knn_core.cuh", "pano_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-Wno-unused-function", "-o", so, srcs[0]])
    lib = ctypes.CDLL(so)
    lib.tcemu_last_error.restype = ctypes.c_char_p
    return lib


@pytest.fixture(scope="session")
def pins():
    return np.load(os.path.join(GOLDEN, "opencv_pins.npz"))


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    pkg = load_pkg()
    e = pkg.Engine(device=0, seed=12345)   # raises if the .so is missing: no fallback
    yield e
    e.close()


@pytest.fixture(scope="session")
def small_pair():
    return load_synth().make_pair(960, 540, seed=267)


@pytest.fixture(scope="session")
def mid_pair():
    return load_synth().make_pair(1920, 1080, seed=31)
