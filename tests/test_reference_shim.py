"""Drop-in check of SURVEY 8 b2 without a GPU: the reference's own GPU executable - src/gpu/main.cpp compiled
UNMODIFIED from /root/reference - with its four .cu stage files replaced by the binding a maintainer would add
(examples/reference_shim/pano_b200_shim.cpp, the code INTEGRATION.md shows).  Here the C ABI underneath the shim is a
test-only stand-in on the CPU oracle (tests/hostsim/abi_standin.cpp), so what is tested is the shim's marshalling of the
reference's C++ types and the reference main's flow on top of it: the panorama must be the one the reference's serial
code produces for the same seed.  The same executable linked against the real libpano_b200.so is
oracle/_ref/gpu_stitching_refmain (GPU tier: tests/test_zz1_reference_gpu_main.py)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_synth

REF = "/root/reference"


@pytest.fixture(scope="module")
def standin_exe(tmp_path_factory):
    if not os.path.exists(os.path.join(REF, "src", "gpu", "main.cpp")):
        pytest.skip("/root/reference not present (the GPU box): nothing to compile")
    oracle_so = os.path.join(ROOT, "oracle", "libpano_oracle.so")
    if not os.path.exists(oracle_so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libpano_oracle.so"], stdout=subprocess.DEVNULL)
    out = str(tmp_path_factory.mktemp("shim") / "gpu_stitching_on_standin")
    o = os.path.join(ROOT, "oracle")
    subprocess.check_call(
        ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-w", "-I" + os.path.join(o, "cvshim"), "-I" + REF + "/src",
         "-I" + REF + "/src/reader", "-I" + REF + "/src/gpu", "-I" + os.path.join(ROOT, "include"), "-o", out,
         REF + "/src/gpu/main.cpp", REF + "/src/reader/reader.cpp",
         os.path.join(ROOT, "examples", "reference_shim", "pano_b200_shim.cpp"),
         os.path.join(ROOT, "tests", "hostsim", "abi_standin.cpp"), os.path.join(o, "cvshim", "cvshim.cpp"),
         oracle_so, "-Wl,-rpath," + o])
    return out


@pytest.mark.parametrize("seed", [12345, 7])
def test_reference_gpu_main_on_the_shim_equals_reference_serial(standin_exe, tmp_path, seed):
    cv2 = pytest.importorskip("cv2")
    from oracle import ref as refmod
    if not refmod.available():
        pytest.skip("oracle/_ref not built")
    left, right, _ = load_synth().make_pair(480, 300, seed=5)
    a, b, out = str(tmp_path / "a.ppm"), str(tmp_path / "b.ppm"), str(tmp_path / "pano.ppm")
    assert cv2.imwrite(a, left) and cv2.imwrite(b, right)
    r = subprocess.run([standin_exe, a, b, "--out", out], capture_output=True, text=True, env=dict(os.environ, PANO_SEED=str(seed)))
    assert r.returncode == 0, r.stderr
    # the reference GPU main's own stage lines (ref: src/gpu/main.cpp:333,351)
    assert "Harris Corner Matching (GPU):" in r.stdout and "RANSAC Homography Estimation (GPU):" in r.stdout
    assert "falling back" not in r.stderr
    ref = refmod.Reference().stitch_pair(left, right, seed=seed)
    assert ref["status"] == 1
    assert np.array_equal(cv2.imread(out), ref["canvas"])


def test_shim_builds_against_the_engine_and_fails_loudly_without_a_gpu(tmp_path):
    """oracle/_ref/gpu_stitching_refmain = the same sources linked against the real libpano_b200.so: without an sm_100
    GPU it must stop with the engine's error, not produce a panorama by some other route"""
    import torch
    exe = os.path.join(ROOT, "oracle", "_ref", "gpu_stitching_refmain")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/gpu_stitching_refmain not built (make -C oracle, needs /root/reference and the engine)")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the GPU tier")
    cv2 = pytest.importorskip("cv2")
    left, right, _ = load_synth().make_pair(320, 200, seed=5)
    a, b, out = str(tmp_path / "a.ppm"), str(tmp_path / "b.ppm"), str(tmp_path / "pano.ppm")
    assert cv2.imwrite(a, left) and cv2.imwrite(b, right)
    r = subprocess.run([exe, a, b, "--out", out], capture_output=True, text=True)
    assert r.returncode != 0 and "pano_create failed" in r.stderr and not os.path.exists(out)
