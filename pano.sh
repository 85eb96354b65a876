#!/bin/bash
# pano.sh — build / run / perf / eval driver with the reference's interface (ref: pano.sh):
#   ./pano.sh build [--no-gpu] [--build-dir=<path>]
#   ./pano.sh run  <serial|openmp|gpu> [--build-dir=<path>] [--dir <d>] [--out <f>] [img ...]
#   ./pano.sh perf <serial|openmp|gpu> ...      (perf record -g + report, if perf is installed)
#   ./pano.sh eval <generated> <reference>
# gpu = the B200 engine (sm_100a); serial / openmp = CPU baselines.  The reference's fourth
# implementation, `opencv` (cv::Stitcher), is a different algorithm and is not provided.
SCRIPT_DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")" &>/dev/null && pwd)"
BUILD_DIR="${SCRIPT_DIR}/build"

exe_name() {
  case "$1" in
    serial) echo serial_stitching ;;
    openmp) echo openmp_stitching ;;
    gpu) echo gpu_stitching ;;
    opencv) echo "The 'opencv' implementation (cv::Stitcher) is out of scope of this engine" >&2; return 1 ;;
    *) echo "Unknown implementation: $1 (supported: serial, openmp, gpu)" >&2; return 1 ;;
  esac
}

find_exe() {  # impl exe -> path; same three locations the reference searches
  for p in "${BUILD_DIR}/$2" "${BUILD_DIR}/src/$1/$2" "./$2"; do
    [ -f "$p" ] && { echo "$p"; return 0; }
  done
  echo "Executable not found: $2 (looked in ${BUILD_DIR}/, ${BUILD_DIR}/src/$1/, ./). Try: $0 build" >&2
  return 1
}

usage() {
  sed -n '2,9p' "${BASH_SOURCE[0]}" | sed 's/^# \{0,1\}//'
  echo "Options for run/perf: --build-dir=<path>  --dir <directory>  --out <file> (default result.jpg)"
  exit 1
}

prepare_run() {  # sets IMPL, EXEC; leaves program arguments in ARGS
  [ $# -ge 1 ] || { echo "Error: Missing implementation"; usage; }
  IMPL=$1; shift
  while [[ $1 =~ ^--build-dir= ]]; do BUILD_DIR="${1#*=}"; shift; done
  if [ $# -lt 1 ]; then echo "Error: No image files specified and --dir option not used"; usage; fi
  local name; name=$(exe_name "$IMPL") || exit 1
  EXEC=$(find_exe "$IMPL" "$name") || exit 1
  ARGS=("$@")
}

[ $# -ge 1 ] || usage
COMMAND=$1; shift
case $COMMAND in
  build)
    NO_GPU=false
    while [[ $1 =~ ^-- ]]; do
      case $1 in
        --no-gpu) NO_GPU=true; shift ;;
        --build-dir=*) BUILD_DIR="${1#*=}"; shift ;;
        *) echo "Unknown option for build command: $1"; usage ;;
      esac
    done
    echo "=== Building project in $BUILD_DIR ==="
    mkdir -p "$BUILD_DIR" && cd "$BUILD_DIR" || { echo "Failed to enter build directory"; exit 1; }
    CMAKE_ARGS="-DBUILD_GPU=ON"; $NO_GPU && { CMAKE_ARGS="-DBUILD_GPU=OFF"; echo "Building without GPU support"; }
    # use the system compilers unless told otherwise (an inherited CXX may lack OpenMP support)
    export CXX="${PANO_CXX:-$(command -v g++)}" CC="${PANO_CC:-$(command -v gcc)}"
    cmake $CMAKE_ARGS "$SCRIPT_DIR" || { echo "CMake failed"; exit 1; }
    make -j"$(nproc)" || { echo "Make failed"; exit 1; }
    echo "=== Build completed successfully ==="
    ;;
  run)
    prepare_run "$@"
    echo "Running $IMPL implementation using $EXEC..."
    "$EXEC" "${ARGS[@]}"; rc=$?
    if [ $rc -eq 0 ]; then echo "Stitching completed successfully!"; else echo "Stitching failed with error code $rc"; fi
    ;;
  perf)
    prepare_run "$@"
    command -v perf >/dev/null || { echo "perf is not installed on this machine"; exit 1; }
    echo "Running performance profiling on $IMPL implementation using $EXEC..."
    perf record -g "$EXEC" "${ARGS[@]}"
    perf report --stdio > "${IMPL}_perf_report.txt" && echo "Performance report saved to ${IMPL}_perf_report.txt"
    ;;
  eval)
    [ $# -ge 2 ] || { echo "Usage: $0 eval <generated_panorama> <reference_panorama>"; exit 1; }
    [ -f "$1" ] || { echo "Error: Generated panorama file not found: $1"; exit 1; }
    [ -f "$2" ] || { echo "Error: Reference panorama file not found: $2"; exit 1; }
    echo "Evaluating panorama quality..."
    python3 "${SCRIPT_DIR}/tools/evaluate_panorama.py" "$1" "$2" && echo "Evaluation completed successfully!"
    ;;
  help) usage ;;
  *) echo "Unknown command: $COMMAND"; usage ;;
esac
