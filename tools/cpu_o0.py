"""What the reference's own build flags cost: its CMake sets no build type, i.e. -O0 (ref: CMakeLists.txt:1-22).
Times the reference's serial pipeline (oracle/_ref: src/serial/main.cpp unmodified + cvshim) compiled with -O2 and
with -O0 on one 1920x1080 synthetic pair (a bounded sample: the 4K pair takes a minute at -O2) and prints one JSON line."""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "ucb-cs267-parallel-panoramic-image-stitching_b200"


def main():
    from oracle import ref as refmod
    synth = importlib.import_module(PKG + ".synth")
    left, right, _ = synth.make_pair(1920, 1080, seed=31)
    out = {"tool": "cpu_o0", "pair": "synthetic 1920x1080, seed 31", "input_MP": 2 * 1920 * 1080 / 1e6}
    for variant, name in (("", "O2"), ("O0", "O0")):
        if not refmod.available(variant):
            continue
        R = refmod.Reference(variant)
        t0 = time.perf_counter()
        s = R.stitch_pair(left, right, seed=12345)
        dt = time.perf_counter() - t0
        out[name] = {"seconds": dt, "MP_per_s": out["input_MP"] / dt, "status": s["status"], "stage_ms": s["times_ms"]}
    if "O2" in out and "O0" in out:
        out["O0_over_O2"] = out["O0"]["seconds"] / out["O2"]["seconds"]
    for l in open("/proc/cpuinfo"):
        if l.startswith("model name"):
            out["cpu_model"] = l.split(":", 1)[1].strip()
            break
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
