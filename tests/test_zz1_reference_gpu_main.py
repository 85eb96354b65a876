"""GPU tier, drop-in check of SURVEY 8 b2: oracle/_ref/gpu_stitching_refmain is the reference's own GPU executable -
src/gpu/main.cpp compiled UNMODIFIED - whose four .cu stage files are replaced by the binding a maintainer would add
(examples/reference_shim/pano_b200_shim.cpp -> the C ABI of libpano_b200.so).  The detector, matcher and RANSAC it
calls are this engine's kernels; canvas geometry, warp and overlay are the reference's own code.  Its panorama must be
the one the reference's serial executable produces for the same RANSAC seed.  (The same sources on a CPU stand-in of the
ABI are checked in tests/test_reference_shim.py; this file sorts last so that it runs after the parity suite.)"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_synth

pytestmark = pytest.mark.gpu

EXE = os.path.join(ROOT, "oracle", "_ref", "gpu_stitching_refmain")


@pytest.mark.parametrize("w,h,scene,seed", [(640, 360, 11, 12345), (800, 450, 3, 7)])
def test_reference_gpu_main_on_the_engine_equals_reference_serial(tmp_path, w, h, scene, seed):
    cv2 = pytest.importorskip("cv2")
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import ref as refmod
    if not os.path.exists(EXE) or not refmod.available():
        pytest.skip("oracle/_ref (built where /root/reference is present) did not travel to this box")
    left, right, _ = load_synth().make_pair(w, h, seed=scene)
    a, b, out = str(tmp_path / "a.ppm"), str(tmp_path / "b.ppm"), str(tmp_path / "pano.ppm")
    assert cv2.imwrite(a, left) and cv2.imwrite(b, right)
    r = subprocess.run([EXE, a, b, "--out", out], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, PANO_SEED=str(seed)))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Harris Corner Matching (GPU):" in r.stdout and "RANSAC Homography Estimation (GPU):" in r.stdout
    assert "falling back" not in r.stderr          # the reference's CPU-RANSAC fallback was not taken
    ref = refmod.Reference().stitch_pair(left, right, seed=seed)
    assert ref["status"] == 1
    assert np.array_equal(cv2.imread(out), ref["canvas"])
