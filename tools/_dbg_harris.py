import importlib, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("ucb-cs267-parallel-panoramic-image-stitching_b200")
from oracle.oracle import Oracle
rng = np.random.default_rng(1)
img = rng.integers(0, 256, (1300, 5000, 3), dtype=np.uint8)
img[50:120, 60:200] = (250, 20, 30)
eng = pkg.Engine(0, 12345)
k = eng.gpuHarrisCornerDetectorDetect(img)
print("gpu", len(k), "cpu", len(Oracle().detect(img)), np.array_equal(k, Oracle().detect(img)))
