"""Process boundary: the executables found by pano.sh keep the reference's command
line, output lines and exit codes.  CPU tier: serial_stitching / openmp_stitching against the
oracle; GPU tier: gpu_stitching produces the same panorama file as the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_synth

PANO = os.path.join(ROOT, "pano.sh")


def exe(impl):
    p = os.path.join(ROOT, "build", "src", impl, impl + "_stitching")
    if not os.path.exists(p):
        args = [PANO, "build"] + (["--no-gpu"] if impl != "gpu" else [])
        subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return p


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    cv2 = pytest.importorskip("cv2")
    d = tmp_path_factory.mktemp("cli")
    left, right, _ = load_synth().make_pair(800, 450, seed=3)
    paths = {}
    for name, img in (("a_left", left), ("b_right", right)):
        for ext in ("png", "ppm", "bmp"):
            p = str(d / ("%s.%s" % (name, ext)))
            assert cv2.imwrite(p, img)
            paths[(name, ext)] = p
    return d, left, right, paths


def run(args, **kw):
    return subprocess.run(args, capture_output=True, text=True, **kw)


@pytest.mark.parametrize("impl,suffix", [("serial", ""), ("openmp", " (OpenMP)")])
def test_cpu_executables_match_oracle_and_print_reference_lines(files, oracle, impl, suffix):
    import cv2
    d, left, right, paths = files
    exe(impl)
    out = str(d / ("out_%s.png" % impl))
    p = run([PANO, "run", impl, paths[("a_left", "png")], paths[("b_right", "ppm")], "--out", out])
    assert p.returncode == 0 and "Stitching completed successfully!" in p.stdout
    for needle in ("Stitching image 2 of 2...", "Harris Corner Detection%s: " % suffix, "Harris Corner Matching%s: " % suffix,
                   "RANSAC Homography Estimation%s: " % suffix, "Image Stitching%s: " % suffix,
                   "Total Stitching Process%s: " % suffix, "Stitched result saved to " + out,
                   "Total Execution Time%s: " % suffix):
        assert needle in p.stdout, needle
    o = oracle.stitch_pair(left, right, seed=12345)
    assert np.array_equal(cv2.imread(out), o["canvas"])


def test_reader_contract(files):
    d, _, _, paths = files
    e = exe("serial")
    p = run([e])                                            # no arguments: usage, exit(-1)
    assert p.returncode == 255 and "Usage:" in p.stderr
    p = run([e, "--dir"])
    assert p.returncode == 255 and "--dir requires" in p.stderr
    p = run([e, paths[("a_left", "bmp")]])                 # one image only
    assert p.returncode == 255 and "At least two images are required" in p.stderr
    p = run([e, paths[("a_left", "bmp")], "/nonexistent.png", paths[("b_right", "bmp")], "--out", str(d / "o.bmp")])
    assert p.returncode == 0 and "Warning: Unable to open image file: /nonexistent.png" in p.stderr
    p = run([e, "--dir", "/definitely/not/a/dir"])
    assert p.returncode == 255 and "is not a valid directory" in p.stderr


def test_dir_mode_default_output(files):
    d, left, right, paths = files
    sub = d / "only_two"
    sub.mkdir(exist_ok=True)
    import shutil
    shutil.copy(paths[("a_left", "png")], sub / "a.png")
    shutil.copy(paths[("b_right", "png")], sub / "b.png")
    p = run([exe("serial"), "--dir", str(sub)], cwd=str(sub), env=dict(os.environ, PANO_SORT_DIR="1"))
    # the default output name is result.jpg (ref: src/reader/reader.cpp:16); without OpenCV the JPEG
    # encoder is nvJPEG, which needs a GPU: on a CPU-only box the run stitches, then reports that
    assert "Total Stitching Process: " in p.stdout
    assert (p.returncode == 0 and os.path.exists(sub / "result.jpg")) or "Failed to write result.jpg" in p.stderr


def test_jpeg_inputs_are_decoded_like_cv_imread(files, oracle, tmp_path):
    """the reference reads its inputs with cv::imread (ref: src/reader/reader.cpp:61,72); without OpenCV C++ the
    executables decode JPEG through the Python cv2 wheel (same decoder), so the pipeline sees the same pixels"""
    import cv2
    d, left, right, paths = files
    jl, jr = str(tmp_path / "l.jpg"), str(tmp_path / "r.jpg")
    cv2.imwrite(jl, left); cv2.imwrite(jr, right)
    out = str(tmp_path / "o.png")
    p = run([exe("serial"), jl, jr, "--out", out], env=dict(os.environ, PANO_DUMP_DECODED=str(tmp_path)))
    assert p.returncode == 0, p.stderr[-500:]
    for i, f in enumerate((jl, jr)):
        assert np.array_equal(cv2.imread(str(tmp_path / ("decoded_%d.ppm" % i))), cv2.imread(f))
    o = oracle.stitch_pair(cv2.imread(jl), cv2.imread(jr), seed=12345)
    assert o["status"] == 1 and np.array_equal(cv2.imread(out), o["canvas"])


@pytest.mark.gpu
def test_nvjpeg_decode_difference_is_reported(files, tmp_path):
    """PANO_JPEG=nvjpeg decodes on the GPU instead; nvJPEG is not bit-identical to libjpeg-turbo: measure it on the
    reference's sample photographs when they are staged, else on a synthetic JPEG"""
    import cv2
    d, left, right, paths = files
    mount = [os.path.join(ROOT, "baseline", "_ref", "images", "mountain", "mountain%d.jpg" % i) for i in (1, 2)]
    if all(os.path.exists(m) for m in mount):
        jl, jr = mount
    else:
        jl, jr = str(tmp_path / "l.jpg"), str(tmp_path / "r.jpg")
        cv2.imwrite(jl, left); cv2.imwrite(jr, right)
    e = exe("gpu")
    stats = {}
    for mode in ("cv2", "nvjpeg"):
        sub = tmp_path / mode
        sub.mkdir()
        p = run([e, jl, jr, "--out", str(sub / "o.png")], env=dict(os.environ, PANO_DUMP_DECODED=str(sub), PANO_JPEG=mode))
        assert p.returncode == 0, p.stderr[-500:]
        stats[mode] = [cv2.imread(str(sub / ("decoded_%d.ppm" % i))) for i in range(2)]
    for i, f in enumerate((jl, jr)):
        assert np.array_equal(stats["cv2"][i], cv2.imread(f))          # default path = the reference's decoder
        diff = np.abs(stats["nvjpeg"][i].astype(np.int16) - stats["cv2"][i].astype(np.int16))
        print("nvJPEG vs cv2.imread on %s: max %d LSB, mean %.4f LSB, %.2f %% of bytes differ"
              % (os.path.basename(f), diff.max(), diff.mean(), 100.0 * (diff > 0).mean()))
        assert diff.max() <= 40 and diff.mean() < 2.0      # a decoder difference (measured: max 18, mean 0.52 LSB), not a different image


@pytest.mark.gpu
def test_gpu_stitching_executable(files, oracle):
    import cv2
    d, left, right, paths = files
    exe("gpu")
    out = str(d / "out_gpu.png")
    p = run([PANO, "run", "gpu", paths[("a_left", "png")], paths[("b_right", "png")], "--out", out])
    assert p.returncode == 0, p.stderr
    for needle in ("Stitching image 2 of 2...", "Harris Corner Detection (GPU): ", "Harris Corner Matching (GPU): ",
                   "RANSAC Homography Estimation (GPU): ", "Image Stitching: ", "Total Stitching Process: ",
                   "Stitched result saved to " + out, "Total Execution Time: "):
        assert needle in p.stdout, needle
    o = oracle.stitch_pair(left, right, seed=12345)
    assert np.array_equal(cv2.imread(out), o["canvas"])
    # JPEG in / out through nvJPEG
    jl, jr = str(d / "l.jpg"), str(d / "r.jpg")
    cv2.imwrite(jl, left); cv2.imwrite(jr, right)
    outj = str(d / "out_gpu.jpg")
    p = run([PANO, "run", "gpu", jl, jr, "--out", outj])
    assert p.returncode == 0, p.stderr
    pj = cv2.imread(outj)
    assert pj is not None and abs(pj.shape[1] - o["canvas"].shape[1]) < 40


def test_benchmark_harness_datasets_mode(files, tmp_path):
    """tools/benchmark_scaling.py: per-dataset runs through pano.sh, parsed from the stage lines
    the reference's benchmark scripts grep, written as CSV"""
    import csv
    import shutil
    import sys
    d, _, _, paths = files
    root = tmp_path / "sets"
    (root / "pairA").mkdir(parents=True)
    shutil.copy(paths[("a_left", "png")], root / "pairA" / "a.png")
    shutil.copy(paths[("b_right", "png")], root / "pairA" / "b.png")
    exe("serial")
    out_csv = str(tmp_path / "datasets.csv")
    p = run([sys.executable, os.path.join(ROOT, "tools", "benchmark_scaling.py"), "datasets", "--root", str(root),
             "--impls", "serial", "--out-dir", str(tmp_path), "--csv", out_csv])
    assert p.returncode == 0, p.stderr[-800:]
    rows = list(csv.DictReader(open(out_csv)))
    assert len(rows) == 1 and rows[0]["dataset"] == "pairA" and rows[0]["impl"] == "serial" and rows[0]["ok"] == "1"
    assert float(rows[0]["image_stitching_ms"]) > 0 and float(rows[0]["total_execution_ms"]) > 0
