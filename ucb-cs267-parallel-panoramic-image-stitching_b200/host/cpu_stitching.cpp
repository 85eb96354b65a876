// cpu_stitching.cpp — the `serial_stitching` and `openmp_stitching` executables: the reference's
// CPU pipelines kept as REPORTED BASELINES next to gpu_stitching (same command line and output
// lines; ref: src/serial/main.cpp:417-452, src/openmp/main.cpp:563-606).  They are built from the
// CPU oracle (oracle/pano_oracle.cpp, an OpenCV-free restatement of the serial path; compiled
// with -fopenmp and -DPANO_CPU_OPENMP for the OpenMP flavour) and are not part of the engine.
// PANO_SEED seeds RANSAC (default 12345; the reference seeds from std::random_device).
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <vector>

#include "reader.hpp"

extern "C" int orc_stitch_pair(const uint8_t* left, int wl, int hl, size_t sl, const uint8_t* right, int wr, int hr,
                               size_t sr, uint32_t seed, uint8_t* canvas, size_t cap, int* geom, double* H,
                               int* stats, double* times_ms);
extern "C" int orc_num_threads();

#ifdef PANO_CPU_OPENMP
#define SUFFIX " (OpenMP)"
#else
#define SUFFIX ""
#endif

int main(int argc, char** argv) {
  auto t_all = std::chrono::high_resolution_clock::now();
  auto ms_since = [](std::chrono::high_resolution_clock::time_point t) {
    return std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t).count();
  };
  ImageReaderResult rr = readImagesFromArgs(argc, argv);
  if (rr.images.size() < 2) {
    std::cerr << "At least two images are required for stitching!" << std::endl;
    return -1;
  }
#ifdef PANO_CPU_OPENMP
  std::cout << "Using OpenMP with " << orc_num_threads() << " threads" << std::endl;
#endif
  const char* sd = std::getenv("PANO_SEED");
  uint32_t seed = sd ? (uint32_t)std::strtoul(sd, nullptr, 10) : 12345u;
  auto t_fold = std::chrono::high_resolution_clock::now();
  pano_io::Image pano = rr.images[0];
  std::cout << std::fixed << std::setprecision(3);
  for (size_t i = 1; i < rr.images.size(); i++) {
    std::cout << "Stitching image " << i + 1 << " of " << rr.images.size() << "..." << std::endl;
    const pano_io::Image& R = rr.images[i];
    size_t cap = 6 * ((size_t)pano.w * pano.h + (size_t)R.w * R.h) * 3;
    std::vector<uint8_t> canvas(cap);
    int geom[4] = {0, 0, 0, 0}, stats[4];
    double H[9], times[4];
    auto t_pair = std::chrono::high_resolution_clock::now();
    int st = orc_stitch_pair(pano.bgr.data(), pano.w, pano.h, pano.stride(), R.bgr.data(), R.w, R.h, R.stride(), seed,
                             canvas.data(), cap, geom, H, stats, times);
    std::cout << "Harris Corner Detection" SUFFIX ": " << times[0] << " ms" << std::endl;
    std::cout << "Harris Corner Matching" SUFFIX ": " << times[1] << " ms" << std::endl;
    if (st == 0) std::cerr << "Not enough matched corners for stitching!" << std::endl;
    else {
      std::cout << "RANSAC Homography Estimation" SUFFIX ": " << times[2] << " ms" << std::endl;
      if (st == -2) std::cerr << "RANSAC failed to estimate a homography matrix!" << std::endl;
    }
    if (st != 1) {
      std::cerr << "Failed to stitch image " << i << "!" << std::endl;
      continue;
    }
    std::cout << "Image Stitching" SUFFIX ": " << ms_since(t_pair) << " ms" << std::endl;
    pano.w = geom[0]; pano.h = geom[1];
    canvas.resize((size_t)pano.w * pano.h * 3);
    pano.bgr.swap(canvas);
  }
  std::cout << "Total Stitching Process" SUFFIX ": " << ms_since(t_fold) << " ms" << std::endl;
  if (pano.empty()) {
    std::cerr << "Panoramic stitching failed!" << std::endl;
    return -1;
  }
  if (!pano_io::write_image(rr.outputFile, pano.bgr.data(), pano.w, pano.h, pano.stride())) {
    std::cerr << "Failed to write " << rr.outputFile << std::endl;
    return -1;
  }
  std::cout << "Stitched result saved to " << rr.outputFile << std::endl;
  std::cout << "\nTotal Execution Time" SUFFIX ": " << ms_since(t_all) << " ms" << std::endl;
  return 0;
}
