"""GPU tier: gpu_stitching's opt-in chain mode (PANO_MODE=chain) - one host process driving every
visible GPU (PANO_GPUS of them): adjacent pairs sharded over the devices, each device rendering its band of canvas rows
(host/chain_multi_gpu.hpp).  The panorama file must be the oracle's chain panorama, whatever the device count, and the
default (no PANO_MODE) must stay the reference's fold.  The control flow is covered on the CPU tier by
tests/test_chain_host.py; this file sorts last and was written after the round's GPU budget was spent (first run on a
B200 is the driver's)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_synth
from test_cli import exe

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def strip_files(tmp_path_factory):
    cv2 = pytest.importorskip("cv2")
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    d = tmp_path_factory.mktemp("chaincli")
    views = load_synth().make_strip(n=4, w=640, h=400, seed=21)
    paths = []
    for i, v in enumerate(views):
        p = str(d / ("v%d.ppm" % i))
        assert cv2.imwrite(p, v)
        paths.append(p)
    return d, views, paths


def test_gpu_stitching_chain_mode_equals_oracle_chain(strip_files, oracle):
    import cv2
    import torch
    d, views, paths = strip_files
    e = exe("gpu")
    pano, pair_H = oracle.stitch_chain(views, seed=12345)
    assert all(H is not None for H in pair_H)
    counts = sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()})
    for gpus in counts:
        out = str(d / ("chain_%d.png" % gpus))
        p = subprocess.run([e] + paths + ["--out", out], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, PANO_MODE="chain", PANO_GPUS=str(gpus)))
        assert p.returncode == 0, p.stderr[-2000:]
        assert "Chain mode: 4 images, adjacent pairs sharded over %d GPU(s)" % min(gpus, 3) in p.stdout
        for needle in ("Harris Corner Detection (GPU): ", "RANSAC Homography Estimation (GPU): ", "Image Stitching: ",
                       "Total Stitching Process: ", "Stitched result saved to " + out, "Total Execution Time: "):
            assert needle in p.stdout, needle
        assert np.array_equal(cv2.imread(out), pano)


def test_gpu_stitching_default_is_still_the_references_fold(strip_files, oracle):
    import cv2
    d, views, paths = strip_files
    out = str(d / "fold.png")
    p = subprocess.run([exe("gpu")] + paths[:3] + ["--out", out], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "Chain mode" not in p.stdout
    fold = oracle.stitch_fold(views[:3], seed=12345)
    canvas = fold[0] if isinstance(fold, tuple) else fold
    assert np.array_equal(cv2.imread(out), np.asarray(canvas))
