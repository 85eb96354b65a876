#!/usr/bin/env bash
# The CPU emulation tier under sanitizers (the host-side counterpart of compute-sanitizer memcheck): builds the kernel
# drivers of tests/hostsim and the whole-engine emulation library with AddressSanitizer, then with UBSan
# (-fno-sanitize-recover), and runs the emulation test files on them.  Usage: tools/emu_sanitize.sh [address|undefined]
# Every global / shared / "device" buffer access of every kernel and of the host orchestration is checked for
# out-of-bounds and (UBSan) misaligned vector accesses, shifts and signed overflow.  Build artefacts are removed
# afterwards so that the normal fixtures rebuild the plain libraries.
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
TESTS="tests/test_knn.py tests/test_harris_emu.py tests/test_warp_emu.py tests/test_match_emu.py tests/test_match_tc_emu.py tests/test_ransac_emu.py tests/test_engine_emu.py"
for SAN in ${1:-address undefined}; do
  FLAGS="-fsanitize=$SAN -fno-omit-frame-pointer -g"
  [ "$SAN" = undefined ] && FLAGS="$FLAGS -fno-sanitize-recover=undefined"
  rm -f tests/hostsim/lib*_emu.so tests/hostsim/gpu_stitching_emu
  for f in knn_emu harris_emu warp_emu match_emu ransac_emu match_tc_emu; do
    g++ -O1 -std=c++17 -fPIC -shared -ffp-contract=off $FLAGS -Wno-unknown-pragmas -Wno-unused-function \
        -o tests/hostsim/lib$f.so tests/hostsim/$f.cpp
  done
  PANO_EMU_CXXFLAGS="$FLAGS" python tests/hostsim/build_emu_lib.py --force > /dev/null
  PRE=""
  [ "$SAN" = address ] && PRE=$(gcc -print-file-name=libasan.so)
  echo "== $SAN"
  # (the two allocation-failure tests make the engine throw C++ exceptions internally; a preloaded libasan inside the
  # Python process cannot intercept __cxa_throw of a library loaded later, so they run in the plain and UBSan builds only)
  SKIP=""
  [ "$SAN" = address ] && SKIP="not failed_allocation"
  LD_PRELOAD="$PRE" ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:verify_asan_link_order=0 \
    UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 python -m pytest $TESTS -x -q -k "$SKIP" 2>&1 | tail -3
done
rm -f tests/hostsim/lib*_emu.so tests/hostsim/gpu_stitching_emu
# Scheduling-order independence: the same kernels with the runnable threads of a block taking their turns in descending
# and in pseudo-random order (PANO_EMU_THREAD_ORDER).  A result that depends on the order is a race (missing barrier,
# warp-lockstep assumption).  The tensor-core matcher is left out: under adversarial scheduling a consumer warp can be
# held back until its operand stage's barrier has advanced two phases - reported as a deadlock by the
# emulation, bounded by the spin limit and the error word on the device.
for ORDER in 1 2; do
  echo "== thread order $ORDER"
  PANO_EMU_THREAD_ORDER=$ORDER python -m pytest tests/test_knn.py tests/test_harris_emu.py tests/test_warp_emu.py tests/test_match_emu.py \
    tests/test_ransac_emu.py -q 2>&1 | tail -2
done
