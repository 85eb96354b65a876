// pano_b200_shim.cpp — what a maintainer of the reference would add to src/gpu/ to run its own, unmodified
// src/gpu/main.cpp on libpano_b200.so.  It defines the four stage entry points that
// main.cpp links (ref: src/gpu/harris_detector.cuh:5-9, harris_matcher.cuh:5-9, ransac.cuh:8-36, convolution.cuh:5)
// with their exact C++ signatures and forwards to the C ABI; it replaces
// src/gpu/{convolution,harris_detector,harris_matcher,ransac}.cu in src/gpu/CMakeLists.txt
// (target_link_libraries(gpu_stitching PRIVATE reader ${OpenCV_LIBS} pano_b200)).  OpenCV types stay on this side.
//
// oracle/Makefile builds it together with the reference's main.cpp + reader.cpp (compiled where they lie) into
// oracle/_ref/gpu_stitching_refmain, the reference's GPU executable running on this engine.
#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <vector>

#include <opencv2/core.hpp>

#include "pano_b200.h"
#include "convolution.cuh"
#include "harris_detector.cuh"
#include "harris_matcher.cuh"
#include "ransac.cuh"

namespace {
pano_ctx* ctx() {   // one context per process: the reference is single-threaded (ref: src/gpu/main.cpp:440-489)
  static pano_ctx* c = [] {
    pano_ctx* p = nullptr;
    const char* sd = std::getenv("PANO_SEED");   // the reference seeds RANSAC from std::random_device
    const int st = pano_create(0, sd ? (uint32_t)std::strtoul(sd, nullptr, 10) : 12345u, &p);
    if (st != PANO_OK) {   // no CPU fallback
      std::cerr << "pano_b200: pano_create failed with status " << st << " (needs an sm_100 GPU)" << std::endl;
      std::abort();
    }
    return p;
  }();
  return c;
}
std::vector<int32_t> xy_of(const std::vector<cv::KeyPoint>& k) {
  std::vector<int32_t> v(2 * std::max<size_t>(k.size(), 1));
  for (size_t i = 0; i < k.size(); i++) {
    v[2 * i] = (int32_t)k[i].pt.x;
    v[2 * i + 1] = (int32_t)k[i].pt.y;
  }
  return v;
}
void report(const char* what, int st) {
  if (st != PANO_OK) std::cerr << what << ": pano status " << st << " " << pano_last_error(ctx()) << std::endl;
}
}  // namespace

std::vector<cv::KeyPoint> gpuHarrisCornerDetectorDetect(const cv::Mat& image, double k, double nmsThresh,
                                                        int nmsNeighborhood) {
  pano_harris_opts o;
  pano_default_harris_opts(&o);
  o.k = k;
  o.nms_thresh = nmsThresh;
  o.nms_neighborhood = nmsNeighborhood;
  // one call with a generous buffer; a second one only if the image has more corners than that
  static std::vector<int32_t> xy(2 * (size_t)(1 << 18));
  int n = 0;
  int st = pano_detect(ctx(), image.data, image.cols, image.rows, (size_t)image.step, PANO_MEM_HOST, &o, xy.data(),
                       (int)(xy.size() / 2), &n);
  if (st == PANO_ERR_CAPACITY) {
    xy.resize(2 * (size_t)n);
    st = pano_detect(ctx(), image.data, image.cols, image.rows, (size_t)image.step, PANO_MEM_HOST, &o, xy.data(), n, &n);
  }
  report("gpuHarrisCornerDetectorDetect", st);
  std::vector<cv::KeyPoint> out;
  if (st != PANO_OK) return out;
  out.reserve(n);
  for (int i = 0; i < n; i++) out.emplace_back((float)xy[2 * i], (float)xy[2 * i + 1], 1.f);   // as ref: src/serial/main.cpp:175
  return out;
}

std::vector<cv::DMatch> gpuHarrisMatchKeyPoints(const std::vector<cv::KeyPoint>& kq, const std::vector<cv::KeyPoint>& kt,
                                                const cv::Mat& imgQ, const cv::Mat& imgT, int patchSize, double maxSSD,
                                                int offset) {
  pano_harris_opts o;
  pano_default_harris_opts(&o);
  o.patch_size = patchSize;
  o.max_ssd_thresh = maxSSD;
  const std::vector<int32_t> q = xy_of(kq), t = xy_of(kt);
  std::vector<pano_dmatch> m(std::max<size_t>(kq.size(), 1));
  int n = 0;
  const int st = pano_match(ctx(), q.data(), (int)kq.size(), t.data(), (int)kt.size(), imgQ.data, imgQ.cols, imgQ.rows,
                            (size_t)imgQ.step, imgT.data, imgT.cols, imgT.rows, (size_t)imgT.step, PANO_MEM_HOST, &o, offset,
                            m.data(), (int)m.size(), &n);
  report("gpuHarrisMatchKeyPoints", st);
  std::vector<cv::DMatch> out;
  if (st != PANO_OK) return out;
  out.reserve(n);
  for (int i = 0; i < n; i++) out.emplace_back(m[i].query_idx, m[i].train_idx, m[i].distance);
  return out;
}

GpuRansacHomographyCalculator::GpuRansacHomographyCalculator(const Options& o) : options_(o) {}

cv::Mat GpuRansacHomographyCalculator::computeHomography(const std::vector<cv::KeyPoint>& k1,
                                                         const std::vector<cv::KeyPoint>& k2,
                                                         const std::vector<cv::DMatch>& matches) {
  pano_ransac_opts o;
  pano_default_ransac_opts(&o);
  o.num_iterations = options_.numIterations_;
  o.num_samples = options_.numSamples_;
  o.distance_threshold = options_.distanceThreshold_;
  const std::vector<int32_t> a = xy_of(k1), b = xy_of(k2);
  std::vector<pano_dmatch> m(std::max<size_t>(matches.size(), 1));
  for (size_t i = 0; i < matches.size(); i++) m[i] = {matches[i].queryIdx, matches[i].trainIdx, matches[i].distance};
  double H[9];
  int best = 0, it = -1;
  const int st = pano_ransac(ctx(), a.data(), (int)k1.size(), b.data(), (int)k2.size(), m.data(), (int)matches.size(),
                             PANO_MEM_HOST, &o, H, &best, &it, nullptr, nullptr, nullptr);
  if (st != PANO_OK) return cv::Mat();   // empty Mat = the reference's failure value (ref: src/gpu/ransac.cuh:29-35)
  cv::Mat out(3, 3, CV_64F);
  for (int i = 0; i < 9; i++) out.at<double>(i / 3, i % 3) = H[i];
  return out;
}

void convolveCUDA(const cv::Mat& in, cv::Mat& out, const std::vector<std::vector<double>>& k) {
  std::vector<double> flat;
  for (const auto& r : k) flat.insert(flat.end(), r.begin(), r.end());
  out.create(in.rows, in.cols, CV_64FC1);
  report("convolveCUDA", pano_convolve_f64(ctx(), in.ptr<double>(), in.cols, in.rows, flat.data(), (int)k.size(),
                                           PANO_MEM_HOST, out.ptr<double>()));
}
