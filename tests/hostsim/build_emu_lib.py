"""TEST INFRASTRUCTURE ONLY: builds tests/hostsim/libpano_b200_emu.so = the engine's own sources (csrc/*.cu, *.cuh, *.hpp,
include/pano_b200.h), UNCHANGED except for one mechanical rewrite - every `kernel<<<grid, block, smem, stream>>>(args);`
becomes `PANO_CUDA(emu_launch(grid, block, smem, [&] { kernel(args); }));` -, compiled by g++ against the fake CUDA runtime
in tests/hostsim/fake_cuda on the CPU emulation of the CUDA execution model (cuda_emu.hpp, tcgen05_emu.hpp).  The result
exports the real C ABI, so the package's Python binding and the GPU-tier test code can drive the WHOLE engine - host
orchestration, every kernel - without a GPU (tests/test_engine_emu.py).  It is never shipped and never loaded by the package.

    python tests/hostsim/build_emu_lib.py        # rebuilds when a source is newer than the library
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = "ucb-cs267-parallel-panoramic-image-stitching_b200"
CSRC = os.path.join(ROOT, PKG, "csrc")
OUT = os.path.join(HERE, "libpano_b200_emu.so")
EXE = os.path.join(HERE, "gpu_stitching_emu")
BUILD = os.path.join(HERE, "_emu_build")
# PANO_EMU_CXXFLAGS="-fsanitize=address -fno-omit-frame-pointer -g": the same build under a sanitizer (run the tests with
# LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0)
EXTRA = os.environ.get("PANO_EMU_CXXFLAGS", "").split()
UNITS = ["pano_api", "harris", "match", "match_tc", "knn", "ransac", "warp"]


def rewrite_launches(text):
    """kernel<<<cfg>>>(args);  ->  PANO_CUDA(emu_launch(cfg..., [&] { kernel(args); }));"""
    out, i = [], 0
    while True:
        j = text.find("<<<", i)
        if j < 0:
            out.append(text[i:])
            break
        # kernel expression: identifier (with ::) and an optional template argument list, scanning backwards
        k = j
        if text[k - 1] == ">":                       # template arguments
            depth = 0
            while True:
                k -= 1
                depth += text[k] == ">"
                depth -= text[k] == "<"
                if depth == 0:
                    break
        while k > 0 and (text[k - 1].isalnum() or text[k - 1] in "_:"):
            k -= 1
        kernel = text[k:j]
        if not kernel:                               # a "<<<" in a comment, not a launch
            out.append(text[i:j + 3])
            i = j + 3
            continue
        e = text.index(">>>", j)
        cfg = [c.strip() for c in split_top(text[j + 3:e])]
        while len(cfg) < 4:
            cfg.append("0")
        a0 = text.index("(", e)
        depth, a1 = 0, a0
        while True:
            depth += text[a1] == "("
            depth -= text[a1] == ")"
            if depth == 0:
                break
            a1 += 1
        args = text[a0 + 1:a1]
        semi = text.index(";", a1)
        out.append(text[i:k])
        out.append("PANO_CUDA(emu_launch(%s, %s, %s, [&] { %s(%s); }))" % (cfg[0], cfg[1], cfg[2], kernel, args))
        i = semi
    return "".join(out)


def split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "(<[":
            depth += 1
        if ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def sources():
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".hpp"))]
    files.append(os.path.join(ROOT, "include", "pano_b200.h"))
    files += [os.path.join(HERE, f) for f in ("cuda_emu.hpp", "tcgen05_emu.hpp", "build_emu_lib.py")]
    files += [os.path.join(HERE, "fake_cuda", f) for f in ("cuda_runtime.h", "cuda.h")]
    host = os.path.join(ROOT, PKG, "host")
    files += [os.path.join(host, f) for f in sorted(os.listdir(host))]
    return files


def build(force=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(f) <= os.path.getmtime(OUT) for f in sources()):
        return OUT
    shutil.rmtree(BUILD, ignore_errors=True)
    csrc_out = os.path.join(BUILD, PKG, "csrc")
    os.makedirs(csrc_out)
    os.makedirs(os.path.join(BUILD, "include"))
    shutil.copy(os.path.join(ROOT, "include", "pano_b200.h"), os.path.join(BUILD, "include"))
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".hpp")):
            text = rewrite_launches(open(os.path.join(CSRC, f)).read())
            name = f[:-3] + ".cpp" if f.endswith(".cu") else f
            open(os.path.join(csrc_out, name), "w").write(text)
    objs = []
    procs = []
    for u in UNITS:
        obj = os.path.join(BUILD, u + ".o")
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-unused-function"] + EXTRA + [
               "-DPANO_CUDA_EMU_LIB=1", "-I" + os.path.join(HERE, "fake_cuda"), "-c", os.path.join(csrc_out, u + ".cpp"), "-o", obj]
        procs.append((u, subprocess.Popen(cmd, stderr=subprocess.PIPE, text=True)))
        objs.append(obj)
    for u, pr in procs:
        err = pr.communicate()[1]
        if pr.returncode != 0:
            sys.stderr.write(err[-6000:])
            raise RuntimeError("emulated build of %s failed" % u)
    subprocess.check_call(["g++", "-shared", "-o", OUT] + EXTRA + objs + ["-lpthread"])
    # the gpu_stitching executable itself (host/gpu_stitching.cpp, reader, image codecs without nvJPEG) on the emulated
    # library: the command-line paths (the reference's fold, PANO_MODE=chain over PANO_EMU_DEVICES "devices")
    host = os.path.join(ROOT, PKG, "host")
    zlib = ["-DPANO_WITH_ZLIB", "-lz"] if os.path.exists("/usr/include/zlib.h") else []
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-unused-function"] + EXTRA +
                          ["-DPANO_HAVE_CUDA", "-include", "cuda_runtime.h", "-I" + os.path.join(HERE, "fake_cuda"), "-I" + host, "-o", EXE,
                           os.path.join(host, "gpu_stitching.cpp"), os.path.join(host, "reader.cpp"), os.path.join(host, "image_io.cpp"),
                           OUT, "-Wl,-rpath,$ORIGIN", "-lpthread"] + zlib)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
