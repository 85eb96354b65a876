// gpu_stitching.cpp — the `gpu_stitching` executable: same command line, same output lines and
// exit codes as the reference's (ref: src/gpu/main.cpp:452-487, src/serial/main.cpp:395-452), with
// every stage running in the B200 engine through the C ABI (include/pano_b200.h).  The panorama
// stays on the device between fold steps.  No CPU fallback: RANSAC failure is reported, not
// retried on the host (the reference's gpu main falls back to its CPU RANSAC, :355-367).
//
// Environment: PANO_DEVICE (GPU ordinal, default 0), PANO_SEED (RANSAC seed, default 12345; the
// reference seeds from std::random_device).
// Opt-in, behaviour changing: PANO_MODE=chain stitches in chain mode (adjacent pairs estimated independently, SURVEY
// 8e2) over PANO_GPUS devices of this box (default: all visible) inside this one process - pairs sharded over the
// devices, canvas bands rendered per device (host/chain_multi_gpu.hpp).  The default is the reference's fold.
// PANO_MATCH=knn[:ratio] makes the fold use the 2-NN / Lowe-ratio matcher (pano_set_match_mode).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/pano_b200.h"
#include "chain_multi_gpu.hpp"
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

struct CudaMem {   // device memory for host/chain_multi_gpu.hpp
  static void set_device(int d) { cudaSetDevice(d); }
  static void* alloc(size_t n) { void* p = nullptr; return cudaMalloc(&p, n) == cudaSuccess ? p : nullptr; }
  static void free(void* p) { if (p) cudaFree(p); }
  static bool zero(void* p, size_t n) { return cudaMemset(p, 0, n) == cudaSuccess; }
  static void sync() { cudaDeviceSynchronize(); }
  static bool h2d_2d(void* d, size_t dp, const void* s, size_t sp, size_t row_bytes, int rows) {
    return cudaMemcpy2D(d, dp, s, sp, row_bytes, (size_t)rows, cudaMemcpyHostToDevice) == cudaSuccess;
  }
  static bool d2h_2d(void* d, size_t dp, const void* s, size_t sp, size_t row_bytes, int rows) {
    return cudaMemcpy2D(d, dp, s, sp, row_bytes, (size_t)rows, cudaMemcpyDeviceToHost) == cudaSuccess;
  }
};

// PANO_MODE=chain: one process, PANO_GPUS devices
int run_chain(const ImageReaderResult& rr, uint32_t seed, const pano_harris_opts& harrisOpts, const pano_ransac_opts& ransacOpts,
              const Timer& totalTimer) {
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) {
    std::cerr << "gpu_stitching: no CUDA device: an sm_100 GPU is required, there is no CPU path" << std::endl;
    return -1;
  }
  const char* ng = std::getenv("PANO_GPUS");
  int D = ng ? std::atoi(ng) : visible;
  if (D < 1) D = 1;
  if (D > visible) D = visible;
  const int n = (int)rr.images.size();
  if (D > n - 1) D = n - 1;          // at most one device per adjacent pair
  std::vector<int> devices;
  for (int d = 0; d < D; d++) devices.push_back(d);
  std::vector<pano_host::ImageView> views;
  for (const pano_io::Image& im : rr.images) views.push_back({im.bgr.data(), im.w, im.h, im.stride()});
  std::cout << "Chain mode: " << n << " images, adjacent pairs sharded over " << D << " GPU(s)" << std::endl;
  Timer foldTimer;
  pano_host::ChainOutput out;
  const int st = pano_host::stitch_chain_multi_gpu<CudaMem>(views, devices, seed, harrisOpts, ransacOpts, &out);
  for (size_t i = 0; i < out.pairs.size(); i++) {
    const pano_pair_result& r = out.pairs[i];
    std::cout << "Stitching image " << i + 2 << " of " << n << "... (GPU " << out.pair_device[i] << ")" << std::endl;
    line("Harris Corner Detection (GPU): ", r.ms_detect);
    line("Harris Corner Matching (GPU): ", r.ms_match);
    if (r.status == PANO_ERR_NO_MATCHES) {
      std::cerr << "Not enough matched corners for stitching!" << std::endl;
    } else {
      line("RANSAC Homography Estimation (GPU): ", r.ms_ransac);
      if (r.status == PANO_ERR_TOO_FEW_MATCHES || r.status == PANO_ERR_NO_HOMOGRAPHY)
        std::cerr << "RANSAC failed to estimate a homography matrix!" << std::endl;
    }
    if (r.status != PANO_OK) std::cerr << "Failed to stitch image " << i + 1 << "!" << std::endl;
  }
  line("Image Stitching: ", foldTimer.elapsed());
  line("Total Stitching Process: ", foldTimer.elapsed());
  if (st != PANO_OK || out.w <= 0 || out.h <= 0) {
    std::cerr << "Panoramic stitching failed! (status " << st << ") " << out.error << std::endl;
    return -1;
  }
  if (out.n_used < n)
    std::cerr << "Chain broken after image " << out.n_used << ": the panorama covers images 1.." << out.n_used << std::endl;
  if (!pano_io::write_image(rr.outputFile, out.canvas.data(), out.w, out.h, (size_t)out.w * 3)) {
    std::cerr << "Failed to write " << rr.outputFile << std::endl;
    return -1;
  }
  std::cout << "Stitched result saved to " << rr.outputFile << std::endl;
  std::cout << "\nTotal Execution Time: " << std::fixed << std::setprecision(3) << totalTimer.elapsed() << " ms" << std::endl;
  return 0;
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
  if (const char* mode = std::getenv("PANO_MODE"))
    if (std::string(mode) == "chain") return run_chain(rr, seed, harrisOpts, ransacOpts, totalTimer);
  // (PANO_MATCH below applies to the fold; chain mode keeps the reference's matcher)
  pano_ctx* ctx = nullptr;
  int st = pano_create(device, seed, &ctx);
  if (st != PANO_OK) {
    std::cerr << "gpu_stitching: cannot create the B200 engine (status " << st
              << "): an sm_100 GPU is required, there is no CPU path" << std::endl;
    return -1;
  }

  // opt-in, behaviour changing: PANO_MATCH=knn[:ratio] - 2 nearest neighbours + Lowe's ratio test (default ratio 0.75)
  // instead of the reference's nearest-patch matcher (pano_set_match_mode)
  if (const char* mm = std::getenv("PANO_MATCH")) {
    const std::string v(mm);
    if (v.rfind("knn", 0) == 0) {
      const double ratio = v.size() > 4 && v[3] == ':' ? std::atof(v.c_str() + 4) : 0.75;
      if (pano_set_match_mode(ctx, 1, ratio, PANO_KNN_PATCH_SSD) != PANO_OK) {
        std::cerr << "gpu_stitching: bad PANO_MATCH value '" << v << "': " << pano_last_error(ctx) << std::endl;
        return -1;
      }
      std::cout << "Matcher: 2-NN + ratio test " << ratio << " (opt-in; not the reference's matcher)" << std::endl;
    }
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
