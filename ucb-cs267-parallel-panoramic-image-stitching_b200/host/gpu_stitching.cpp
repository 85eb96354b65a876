// gpu_stitching.cpp — the `gpu_stitching` executable: same command line, same output lines and
// exit codes as the reference's (ref: src/gpu/main.cpp:452-487, src/serial/main.cpp:395-452), with
// every stage running in the B200 engine through the C ABI (include/pano_b200.h).  The panorama
// stays on the device between fold steps.  No CPU fallback: RANSAC failure is reported, not
// retried on the host (the reference's gpu main falls back to its CPU RANSAC, :355-367).
//
// Environment: PANO_DEVICE (GPU ordinal, default 0), PANO_SEED (RANSAC seed, default 12345; the
// reference seeds from std::random_device).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>

#include <cuda_runtime.h>

#include "../../include/pano_b200.h"
#include "reader.hpp"

namespace {
class Timer {
 public:
  Timer() : start_(std::chrono::high_resolution_clock::now()) {}
  double elapsed() const {
    return std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - start_).count();
  }
 private:
  std::chrono::time_point<std::chrono::high_resolution_clock> start_;
};

void line(const char* what, double ms) {
  std::cout << what << std::fixed << std::setprecision(3) << ms << " ms" << std::endl;
}
}  // namespace

int main(int argc, char** argv) {
  Timer totalTimer;
  ImageReaderResult rr = readImagesFromArgs(argc, argv);
  if (rr.images.size() < 2) {
    std::cerr << "At least two images are required for stitching!" << std::endl;
    return -1;
  }
  pano_harris_opts harrisOpts;
  pano_default_harris_opts(&harrisOpts);
  harrisOpts.nms_thresh = 1e6;
  harrisOpts.max_ssd_thresh = 1e8;
  pano_ransac_opts ransacOpts;
  pano_default_ransac_opts(&ransacOpts);

  const char* dv = std::getenv("PANO_DEVICE");
  const char* sd = std::getenv("PANO_SEED");
  int device = dv ? std::atoi(dv) : 0;
  uint32_t seed = sd ? (uint32_t)std::strtoul(sd, nullptr, 10) : 12345u;
  pano_ctx* ctx = nullptr;
  int st = pano_create(device, seed, &ctx);
  if (st != PANO_OK) {
    std::cerr << "gpu_stitching: cannot create the B200 engine (status " << st
              << "): an sm_100 GPU is required, there is no CPU path" << std::endl;
    return -1;
  }

  // stitchAllImages: left fold, the panorama stays in device memory
  Timer foldTimer;
  const size_t n = rr.images.size();
  std::vector<uint8_t*> dev(n, nullptr);
  for (size_t i = 0; i < n; i++) {
    const pano_io::Image& im = rr.images[i];
    if (cudaMalloc((void**)&dev[i], im.bgr.size()) != cudaSuccess ||
        cudaMemcpy(dev[i], im.bgr.data(), im.bgr.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      std::cerr << "gpu_stitching: device allocation/copy failed" << std::endl;
      return -1;
    }
  }
  const uint8_t* pano = dev[0];
  int pw = rr.images[0].w, ph = rr.images[0].h;
  size_t pstride = rr.images[0].stride();
  for (size_t i = 1; i < n; i++) {
    std::cout << "Stitching image " << i + 1 << " of " << n << "..." << std::endl;
    pano_pair_result r;
    st = pano_stitch_pair(ctx, pano, pw, ph, pstride, dev[i], rr.images[i].w, rr.images[i].h, rr.images[i].stride(),
                          PANO_MEM_DEVICE, &harrisOpts, &ransacOpts, &r);
    line("Harris Corner Detection (GPU): ", r.ms_detect);
    line("Harris Corner Matching (GPU): ", r.ms_match);
    if (st == PANO_ERR_NO_MATCHES) {
      std::cerr << "Not enough matched corners for stitching!" << std::endl;
    } else {
      line("RANSAC Homography Estimation (GPU): ", r.ms_ransac);
      if (st == PANO_ERR_TOO_FEW_MATCHES || st == PANO_ERR_NO_HOMOGRAPHY)
        std::cerr << "RANSAC failed to estimate a homography matrix!" << std::endl;
      else if (st == PANO_ERR_ROI)
        std::cerr << "Left image does not fit the canvas (the reference would throw here)!" << std::endl;
      else if (st != PANO_OK)
        std::cerr << "Engine error " << st << ": " << pano_last_error(ctx) << std::endl;
    }
    if (st != PANO_OK) {
      std::cerr << "Failed to stitch image " << i << "!" << std::endl;
      if (st == PANO_ERR_CUDA) return -1;
      continue;  // keep the previous panorama (ref: src/serial/main.cpp:404-407)
    }
    line("Image Stitching: ", r.ms_total);
    pano_canvas_device(ctx, &pano, &pstride, &pw, &ph);
  }
  line("Total Stitching Process: ", foldTimer.elapsed());

  if (!pano || pw <= 0 || ph <= 0) {
    std::cerr << "Panoramic stitching failed!" << std::endl;
    return -1;
  }
  if (!pano_io::write_image_device(rr.outputFile, pano, pw, ph, pstride)) {
    std::cerr << "Failed to write " << rr.outputFile << std::endl;
    return -1;
  }
  std::cout << "Stitched result saved to " << rr.outputFile << std::endl;
  std::cout << "\nTotal Execution Time: " << std::fixed << std::setprecision(3) << totalTimer.elapsed() << " ms" << std::endl;
  for (auto p : dev) cudaFree(p);
  pano_destroy(ctx);
  return 0;
}
