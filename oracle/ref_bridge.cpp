// ref_bridge.cpp — builds the UNMODIFIED reference into oracle/_ref/.  TEST INFRASTRUCTURE ONLY.
//
// This translation unit #includes the reference's own source file from where it lies
// (/root/reference/src/serial/main.cpp, or src/openmp/main.cpp with -DREF_OPENMP; found through
// -I$(REF)/src — nothing is copied into this repository) and exports its functions through a
// small C ABI for ctypes (oracle/ref.py).  OpenCV is replaced by oracle/cvshim (cv2-pinned
// arithmetic; see cvshim.hpp).  Exactly two things are adjusted from outside, by macros that are
// active only while the reference file is being read:
//   * `main`               -> `ref_cli_main` (so the file can live in a shared library; the
//                             executable targets of oracle/Makefile compile it untouched);
//   * `std::random_device` -> a device that returns the seed set with ref_set_seed(), which is
//                             the one deliberate deviation of the whole project: the reference
//                             seeds std::mt19937 from std::random_device
//                             (ref: src/serial/main.cpp:264-265, src/openmp/main.cpp:385-386).
// Everything else — convolution loops, Harris response, NMS scan, matcher, the RANSAC loop with
// the real libstdc++ std::shuffle / std::sample, canvas geometry, ROI copy, overlay,
// stitchAllImages fold, the Timer lines on stdout — is the reference's code, executed.
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <iomanip>
#include <iostream>
#include <limits>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <opencv2/core.hpp>

namespace pano_ref {
unsigned g_seed = 12345u;
struct seeded_device {  // stands in for std::random_device: operator() returns the pinned seed
  typedef unsigned result_type;
  unsigned operator()() { return g_seed; }
};
}  // namespace pano_ref
namespace std { typedef ::pano_ref::seeded_device pano_ref_seeded_device; }

#define random_device pano_ref_seeded_device
#define main ref_cli_main
#ifdef REF_OPENMP
#include "openmp/main.cpp"   // the reference's file, unmodified (-I/root/reference/src)
#else
#include "serial/main.cpp"   // the reference's file, unmodified (-I/root/reference/src)
#endif
#undef main
#undef random_device

namespace {

cv::Mat wrap_bgr(const uint8_t* p, int w, int h, size_t stride) {
  // The reference indexes images through Mat::at (serial) or assumes continuity (openmp :249), so
  // hand it a continuous copy, exactly what cv::imread would have produced.
  cv::Mat m(h, w, CV_8UC3);
  for (int y = 0; y < h; y++) std::memcpy(m.ptr<uint8_t>(y), p + (size_t)y * stride, (size_t)w * 3);
  return m;
}

std::vector<cv::KeyPoint> wrap_kp(const int32_t* xy, int n) {
  std::vector<cv::KeyPoint> v;
  v.reserve(n);
  for (int i = 0; i < n; i++) v.push_back(cv::KeyPoint((float)xy[2 * i], (float)xy[2 * i + 1], 1.f));
  return v;
}

// The reference prints its stage timings on std::cout (ref: src/serial/main.cpp:183,242,302,389,412);
// the bridge captures them so that callers get the reference's own Timer readings.
std::string g_log;
std::mutex g_mu;
struct Capture {
  std::ostringstream os;
  std::streambuf* old;
  Capture() : old(std::cout.rdbuf(os.rdbuf())) {}
  ~Capture() { std::cout.rdbuf(old); g_log += os.str(); }
};

HarrisCornerOptions harris_opts(double k, double thresh, int nbhd, int patch, double maxssd) {
  HarrisCornerOptions o;
  o.k_ = k; o.nmsThresh_ = thresh; o.nmsNeighborhood_ = nbhd; o.patchSize_ = patch; o.maxSSDThresh_ = maxssd;
  return o;
}

int export_mat(const cv::Mat& m, uint8_t* out, size_t cap, int* wh) {
  if (m.empty()) return 0;
  wh[0] = m.cols; wh[1] = m.rows;
  size_t rb = (size_t)m.cols * 3;
  if (rb * m.rows > cap) return -1;
  for (int y = 0; y < m.rows; y++) std::memcpy(out + (size_t)y * rb, m.ptr<uint8_t>(y), rb);
  return 1;
}

}  // namespace

extern "C" {

struct ref_dmatch { int32_t queryIdx, trainIdx; float distance; };

void ref_set_seed(unsigned seed) { pano_ref::g_seed = seed; }

int ref_is_openmp() {
#ifdef REF_OPENMP
  return 1;
#else
  return 0;
#endif
}

int ref_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// copies and clears the captured stdout of the reference functions; returns the length
int ref_take_log(char* buf, int cap) {
  std::lock_guard<std::mutex> lk(g_mu);
  int n = (int)std::min<size_t>(g_log.size(), cap > 0 ? (size_t)cap - 1 : 0);
  if (buf && cap > 0) { std::memcpy(buf, g_log.data(), n); buf[n] = 0; }
  g_log.clear();
  return n;
}

// ref: getGaussianKernel (src/serial/main.cpp:73-91)
void ref_gaussian_kernel(int ksize, double sigma, double* out) {
  auto k = getGaussianKernel(ksize, sigma);
  for (int i = 0; i < ksize; i++)
    for (int j = 0; j < ksize; j++) out[i * ksize + j] = k[i][j];
}

// ref: convolveSequential (src/serial/main.cpp:96-116) / convolveParallel (src/openmp/main.cpp:105-126)
void ref_convolve(const double* in, int w, int h, const double* kern, int ksize, double* out) {
  cv::Mat m(h, w, CV_64FC1, (void*)in);
  std::vector<std::vector<double>> k(ksize, std::vector<double>(ksize));
  for (int i = 0; i < ksize; i++)
    for (int j = 0; j < ksize; j++) k[i][j] = kern[i * ksize + j];
#ifdef REF_OPENMP
  cv::Mat r = convolveParallel(m, k);
#else
  cv::Mat r = convolveSequential(m, k);
#endif
  for (int y = 0; y < h; y++) std::memcpy(out + (size_t)y * w, r.ptr<double>(y), sizeof(double) * w);
}

// ref: seqHarrisCornerDetectorDetect (src/serial/main.cpp:119-185); returns the keypoint count
int ref_detect(const uint8_t* bgr, int w, int h, size_t stride, double k, double thresh, int nbhd,
               int32_t* xy, int cap) {
  std::lock_guard<std::mutex> lk(g_mu);
  Capture c;
  cv::Mat img = wrap_bgr(bgr, w, h, stride);
#ifdef REF_OPENMP
  auto kp = ompHarrisCornerDetectorDetect(img, harris_opts(k, thresh, nbhd, 5, 1e8));
#else
  auto kp = seqHarrisCornerDetectorDetect(img, harris_opts(k, thresh, nbhd, 5, 1e8));
#endif
  int n = (int)kp.size();
  for (int i = 0; i < n && i < cap; i++) { xy[2 * i] = (int32_t)kp[i].pt.x; xy[2 * i + 1] = (int32_t)kp[i].pt.y; }
  return n;
}

// ref: seqHarrisMatchKeyPoints (src/serial/main.cpp:188-244); returns the match count
int ref_match(const int32_t* kq, int nq, const int32_t* kt, int nt, const uint8_t* imq, int wq, int hq,
              size_t sq, const uint8_t* imt, int wt, int ht, size_t st, int patch, double maxSSD,
              int offset, ref_dmatch* out, int cap) {
  std::lock_guard<std::mutex> lk(g_mu);
  Capture c;
  cv::Mat a = wrap_bgr(imq, wq, hq, sq), b = wrap_bgr(imt, wt, ht, st);
  auto vq = wrap_kp(kq, nq), vt = wrap_kp(kt, nt);
#ifdef REF_OPENMP
  auto m = ompHarrisMatchKeyPoints(vq, vt, a, b, harris_opts(0.04, 1e6, 3, patch, maxSSD), offset);
#else
  auto m = seqHarrisMatchKeyPoints(vq, vt, a, b, harris_opts(0.04, 1e6, 3, patch, maxSSD), offset);
#endif
  int n = (int)m.size();
  for (int i = 0; i < n && i < cap; i++) out[i] = ref_dmatch{m[i].queryIdx, m[i].trainIdx, m[i].distance};
  return n;
}

// ref: SeqRansacHomographyCalculator::computeHomography (src/serial/main.cpp:247-307).
// Returns 1 and H (row-major 3x3) or 0 for the reference's empty Mat.
int ref_ransac(const int32_t* kp1, int n1, const int32_t* kp2, int n2, const ref_dmatch* matches, int m,
               int iters, int nsamples, double thr, unsigned seed, double* H) {
  std::lock_guard<std::mutex> lk(g_mu);
  Capture c;
  pano_ref::g_seed = seed;
  auto v1 = wrap_kp(kp1, n1), v2 = wrap_kp(kp2, n2);
  std::vector<cv::DMatch> mv;
  mv.reserve(m);
  for (int i = 0; i < m; i++) mv.push_back(cv::DMatch(matches[i].queryIdx, matches[i].trainIdx, matches[i].distance));
  RansacOptions o;
  o.numIterations_ = iters; o.numSamples_ = nsamples; o.distanceThreshold_ = thr;
#ifdef REF_OPENMP
  OmpRansacHomographyCalculator r(o);
#else
  SeqRansacHomographyCalculator r(o);
#endif
  cv::Mat Hm = r.computeHomography(v1, v2, mv);
  if (Hm.empty()) return 0;
  for (int i = 0; i < 9; i++) H[i] = Hm.at<double>(i / 3, i % 3);
  return 1;
}

// ref: stitchTwoImages (src/serial/main.cpp:311-391).  1 = canvas written (wh = its size),
// 0 = the reference returned an empty Mat (no matches / RANSAC failed), -1 = cap too small,
// -3 = the reference threw (cv::Exception from the left-image ROI, ref :376).
int ref_stitch_pair(const uint8_t* left, int wl, int hl, size_t sl, const uint8_t* right, int wr, int hr,
                    size_t sr, unsigned seed, uint8_t* canvas, size_t cap, int* wh) {
  std::lock_guard<std::mutex> lk(g_mu);
  Capture c;
  pano_ref::g_seed = seed;
  cv::Mat L = wrap_bgr(left, wl, hl, sl), R = wrap_bgr(right, wr, hr, sr);
  try {
    cv::Mat out = stitchTwoImages(L, R, HarrisCornerOptions(), RansacOptions());
    return export_mat(out, canvas, cap, wh);
  } catch (const cv::Exception& e) {
    std::cerr << e.what() << std::endl;
    return -3;
  }
}

// ref: stitchAllImages (src/serial/main.cpp:395-414): n tightly packed BGR8 images.
int ref_stitch_all(const uint8_t* const* imgs, const int* ws, const int* hs, int n, unsigned seed,
                   uint8_t* canvas, size_t cap, int* wh) {
  std::lock_guard<std::mutex> lk(g_mu);
  Capture c;
  pano_ref::g_seed = seed;
  std::vector<cv::Mat> v;
  for (int i = 0; i < n; i++) v.push_back(wrap_bgr(imgs[i], ws[i], hs[i], (size_t)ws[i] * 3));
  try {
    cv::Mat out = stitchAllImages(v, HarrisCornerOptions(), RansacOptions());
    return export_mat(out, canvas, cap, wh);
  } catch (const cv::Exception& e) {
    std::cerr << e.what() << std::endl;
    return -3;
  }
}

}  // extern "C"
