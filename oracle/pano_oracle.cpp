// pano_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A plain C++17 restatement, without OpenCV, of the serial stitching path of the
// reference (`src/serial/main.cpp`, cited per function below as "ref:").  It exists only
// so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs can check and time the CUDA engine against the reference's semantics.  Nothing in
// the product path (the package's csrc/, the C ABI, the CLI) may link or call this file.
//
// One deliberate deviation from the reference: `std::mt19937 rng(rd())`
// (ref: src/serial/main.cpp:264-265) becomes `std::mt19937 rng(seed)` with the seed passed
// in, so runs are reproducible.  The shuffle itself is the real libstdc++ std::shuffle.
//
// The reference cannot be compiled here (every target needs OpenCV C++, which is absent),
// so the OpenCV routines on the path are restated from their published algorithms
// (OpenCV 4.x: cvtColor BGR2GRAY, findHomography 4-point path = normalised DLT + Jacobi
// eigen, gemm small-matrix path, Mat /= scalar, cv::norm(Point2f), perspectiveTransform,
// invert 3x3, warpPerspective INTER_LINEAR/BORDER_CONSTANT).  Parity pin: each restated
// routine is checked bit-for-bit against the real routines of Python cv2 4.13.0 by
// oracle/gen_golden.py and tests/test_oracle_vs_golden.py (fixtures in tests/golden/).
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off; no -march so no FMA is emitted,
// matching the reference's x86-64 baseline build).

#include <algorithm>
#include <cfloat>
#include <climits>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "cv_pinned.hpp"

namespace {

struct Match {  // the three cv::DMatch fields the path uses (ref: src/serial/main.cpp:237)
  int32_t queryIdx;
  int32_t trainIdx;
  float distance;
};

using namespace cvpin;  // the cv2-pinned OpenCV arithmetic (oracle/cv_pinned.hpp)


// ref: src/serial/main.cpp:73-91  getGaussianKernel(5, 1.0)
void gaussian_kernel(int ksize, double sigma, std::vector<double>& k) {
  k.assign((size_t)ksize * ksize, 0.0);
  double sum = 0.0;
  int half = ksize / 2;
  for (int i = 0; i < ksize; ++i) {
    int x = i - half;
    for (int j = 0; j < ksize; ++j) {
      int y = j - half;
      k[(size_t)i * ksize + j] = exp(-(x * x + y * y) / (2 * sigma * sigma));
      sum += k[(size_t)i * ksize + j];
    }
  }
  for (auto& e : k) e /= sum;
}

// ref: src/serial/main.cpp:96-116  convolveSequential — correlation, zero border of k px,
// single accumulator, i (rows) outer, j (cols) inner, separate rounded multiply and add.
void convolve(const double* in, int w, int h, const double* kern, int ksize, double* out) {
  int k = ksize / 2;
  std::fill(out, out + (size_t)w * h, 0.0);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int y = k; y < h - k; y++) {
    for (int x = k; x < w - k; x++) {
      double sum = 0.0;
      for (int i = -k; i <= k; i++)
        for (int j = -k; j <= k; j++)
          sum += in[(size_t)(y + i) * w + (x + j)] * kern[(k + i) * ksize + (k + j)];
      out[(size_t)y * w + x] = sum;
    }
  }
}

// ref: src/serial/main.cpp:119-155 — gray, Sobel, products, Gaussian, response.
void harris_response(const uint8_t* bgr, int w, int h, size_t stride, double kparam,
                     std::vector<double>& resp) {
  size_t n = (size_t)w * h;
  std::vector<double> gray(n), gx(n), gy(n), xx(n), yy(n), xy(n), t(n);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) gray[(size_t)y * w + x] = gray_of(bgr + y * stride + 3 * x);
  const double sobx[9] = {-1, 0, 1, -2, 0, 2, -1, 0, 1};
  const double soby[9] = {-1, -2, -1, 0, 0, 0, 1, 2, 1};
  std::vector<double> g5;
  gaussian_kernel(5, 1.0, g5);
  convolve(gray.data(), w, h, sobx, 3, gx.data());
  convolve(gray.data(), w, h, soby, 3, gy.data());
  for (size_t i = 0; i < n; i++) {
    xx[i] = gx[i] * gx[i];
    yy[i] = gy[i] * gy[i];
    xy[i] = gx[i] * gy[i];
  }
  convolve(xx.data(), w, h, g5.data(), 5, t.data()); xx.swap(t);
  convolve(yy.data(), w, h, g5.data(), 5, t.data()); yy.swap(t);
  convolve(xy.data(), w, h, g5.data(), 5, t.data()); xy.swap(t);
  resp.assign(n, 0.0);
  for (size_t i = 0; i < n; i++) {
    double det = xx[i] * yy[i] - xy[i] * xy[i];
    double trace = xx[i] + yy[i];
    resp[i] = det - kparam * trace * trace;
  }
}

// ref: src/serial/main.cpp:157-180 — threshold + strict NMS, row-major order.
void nms(const double* resp, int w, int h, double thresh, int nbhd, std::vector<int32_t>& xy) {
  int halfLen = nbhd / 2;
  for (int y = halfLen; y < h - halfLen; y++) {
    for (int x = halfLen; x < w - halfLen; x++) {
      double r = resp[(size_t)y * w + x];
      if (r <= thresh) continue;
      double max_resp = std::numeric_limits<double>::lowest();
      bool skip = false;
      for (int i = -halfLen; i <= halfLen && !skip; i++) {
        for (int j = -halfLen; j <= halfLen; j++) {
          if (i == 0 && j == 0) continue;
          max_resp = std::max(max_resp, resp[(size_t)(y + i) * w + (x + j)]);
          if (max_resp > r) { skip = true; break; }
        }
      }
      if (skip) continue;
      if (r > max_resp) { xy.push_back(x); xy.push_back(y); }
    }
  }
}

// ref: src/serial/main.cpp:188-244 — brute-force SSD over patch x patch x 3 u8, first strict
// minimum, in-border test on both sides, threshold on the best SSD.
void match_keypoints(const int32_t* kq, int nq, const int32_t* kt, int nt,
                     const uint8_t* imq, int wq, int hq, size_t sq,
                     const uint8_t* imt, int wt, int ht, size_t st,
                     int patch, double maxSSD, int offset, std::vector<Match>& out) {
  int border = patch / 2;
  // pre-gather in-border train patches (pure re-ordering of loads; same arithmetic)
  std::vector<int> tj;
  std::vector<uint8_t> tp;
  int plen = patch * patch * 3;
  for (int j = 0; j < nt; j++) {
    int x = kt[2 * j], y = kt[2 * j + 1];
    if (x < border || y < border || x + border >= wt || y + border >= ht) continue;
    tj.push_back(j);
    size_t o = tp.size();
    tp.resize(o + plen);
    uint8_t* d = &tp[o];
    for (int dy = -border; dy <= border; dy++)
      for (int dx = -border; dx <= border; dx++)
        for (int c = 0; c < 3; c++) *d++ = imt[(size_t)(y + dy) * st + 3 * (x + dx) + c];
  }
  std::vector<Match> res((size_t)nq);
  std::vector<uint8_t> has((size_t)nq, 0);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64)
#endif
  for (int i = 0; i < nq; i++) {
    int x = kq[2 * i], y = kq[2 * i + 1];
    if (x < border || y < border || x + border >= wq || y + border >= hq) continue;
    std::vector<uint8_t> qp((size_t)plen);
    {
      uint8_t* d = qp.data();
      for (int dy = -border; dy <= border; dy++)
        for (int dx = -border; dx <= border; dx++)
          for (int c = 0; c < 3; c++) *d++ = imq[(size_t)(y + dy) * sq + 3 * (x + dx) + c];
    }
    int best = -1;
    uint64_t bestSSD = std::numeric_limits<uint64_t>::max();
    for (size_t jj = 0; jj < tj.size(); jj++) {
      const uint8_t* t = &tp[jj * plen];
      uint32_t ssd = 0;
      for (int e = 0; e < plen; e++) {
        int d = (int)qp[e] - (int)t[e];
        ssd += (uint32_t)(d * d);
      }
      if ((uint64_t)ssd < bestSSD) { bestSSD = ssd; best = tj[jj]; }
    }
    if ((double)bestSSD < maxSSD) {
      res[i] = Match{i + offset, best, (float)bestSSD};
      has[i] = 1;
    }
  }
  for (int i = 0; i < nq; i++)
    if (has[i]) out.push_back(res[i]);
}


// ---- NOT the reference: checker of the engine's opt-in pano_match_knn (north star item (c)) ----------------------
// The published algorithm it restates is OpenCV's BFMatcher::knnMatch(k = 2) followed by Lowe's ratio test
// (Lowe 2004; the OpenCV feature-matching tutorial), on the candidates and with the tie rule of the reference's
// matcher above: in-border keypoints only, neighbours ordered by (distance, position in the train list).
// descriptor 0: the reference's patch, distance = SSD (exact), test  (double)ssd1 < (ratio * ratio) * (double)ssd2,
//              i.e. d1 < ratio * d2 on the L2 distances.
// descriptor 1: 256-bit intensity-comparison descriptor of the 5 x 5 gray patch (gray = cvtColor's formula): bit k =
//              gray[a] < gray[b] for pair number (37 k mod 300) of the lexicographic list of the 300 position pairs
//              a < b (positions row-major); distance = Hamming; test  (double)h1 < ratio * (double)h2.
// Pinned against real cv2.BFMatcher (NORM_L2SQR / NORM_HAMMING) in tests/test_knn.py.
struct KnnMatch { int32_t queryIdx, trainIdx; float distance, second; };

void knn_binary_descriptor(const uint8_t* im, size_t stride, int x, int y, uint32_t* bits) {
  static std::vector<std::pair<int, int>> pairs;
  if (pairs.empty())
    for (int a = 0; a < 25; a++)
      for (int b = a + 1; b < 25; b++) pairs.push_back({a, b});
  int g[25];
  for (int dy = -2; dy <= 2; dy++)
    for (int dx = -2; dx <= 2; dx++) {
      const uint8_t* p = im + (size_t)(y + dy) * stride + 3 * (size_t)(x + dx);
      g[(dy + 2) * 5 + dx + 2] = (p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15;
    }
  for (int w = 0; w < 8; w++) bits[w] = 0;
  for (int k = 0; k < 256; k++) {
    const std::pair<int, int>& pr = pairs[(size_t)((37 * k) % 300)];
    if (g[pr.first] < g[pr.second]) bits[k >> 5] |= 1u << (k & 31);
  }
}

void match_knn(const int32_t* kq, int nq, const int32_t* kt, int nt, const uint8_t* imq, int wq, int hq, size_t sq,
               const uint8_t* imt, int wt, int ht, size_t st, int patch, int descriptor, double ratio,
               std::vector<KnnMatch>& out) {
  const int border = patch / 2, plen = patch * patch * 3;
  const int dlen = descriptor == 1 ? 32 : plen;   // bytes per descriptor
  auto describe = [&](const uint8_t* im, size_t stride, int x, int y, uint8_t* d) {
    if (descriptor == 1) {
      uint32_t bits[8];
      knn_binary_descriptor(im, stride, x, y, bits);
      memcpy(d, bits, 32);
    } else {
      for (int dy = -border; dy <= border; dy++)
        for (int dx = -border; dx <= border; dx++)
          for (int c = 0; c < 3; c++) *d++ = im[(size_t)(y + dy) * stride + 3 * (x + dx) + c];
    }
  };
  auto distance = [&](const uint8_t* a, const uint8_t* b) -> uint32_t {
    uint32_t s = 0;
    if (descriptor == 1) {
      for (int e = 0; e < 32; e++) s += (uint32_t)__builtin_popcount((unsigned)(a[e] ^ b[e]));
    } else {
      for (int e = 0; e < plen; e++) { const int d = (int)a[e] - (int)b[e]; s += (uint32_t)(d * d); }
    }
    return s;
  };
  std::vector<int> tj;
  std::vector<uint8_t> tp;
  for (int j = 0; j < nt; j++) {
    const int x = kt[2 * j], y = kt[2 * j + 1];
    if (x < border || y < border || x + border >= wt || y + border >= ht) continue;
    tj.push_back(j);
    tp.resize(tp.size() + dlen);
    describe(imt, st, x, y, &tp[tp.size() - dlen]);
  }
  const double factor = descriptor == 1 ? ratio : ratio * ratio;
  std::vector<KnnMatch> res((size_t)nq);
  std::vector<uint8_t> has((size_t)nq, 0);
  if (descriptor == 1) { uint32_t warm[8]; knn_binary_descriptor(imt, st, 2, 2, warm); }   // (builds the pair table before threads start)
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64)
#endif
  for (int i = 0; i < nq; i++) {
    const int x = kq[2 * i], y = kq[2 * i + 1];
    if (x < border || y < border || x + border >= wq || y + border >= hq) continue;
    std::vector<uint8_t> qp((size_t)dlen);
    describe(imq, sq, x, y, qp.data());
    int j1 = -1, j2 = -1;
    uint32_t d1 = 0, d2 = 0;
    for (size_t jj = 0; jj < tj.size(); jj++) {   // strict '<' twice: the earlier train keypoint keeps its rank on ties
      const uint32_t d = distance(qp.data(), &tp[jj * dlen]);
      if (j1 < 0 || d < d1) { j2 = j1; d2 = d1; j1 = (int)jj; d1 = d; }
      else if (j2 < 0 || d < d2) { j2 = (int)jj; d2 = d; }
    }
    if (j2 < 0) continue;   // fewer than two candidates: no runner-up, no match
    if ((double)d1 < factor * (double)d2) { res[(size_t)i] = KnnMatch{i, tj[(size_t)j1], (float)d1, (float)d2}; has[(size_t)i] = 1; }
  }
  for (int i = 0; i < nq; i++)
    if (has[(size_t)i]) out.push_back(res[(size_t)i]);
}

// Inlier predicate of ref: src/serial/main.cpp:285-293.
//   pt2Transformed = H * (x, y, 1)      -> gemm small path, (h0*x + h1*y) + h2*1
//   pt2Transformed /= w                  -> Mat::convertTo(-1, 1./w): multiply by reciprocal
//   Point2f est(X, Y)                    -> double -> float casts
//   cv::norm(est - pt2) < thr            -> float subtraction, sqrt((double)dx*dx + (double)dy*dy)
// div_mode 0 = reciprocal multiply (OpenCV's Mat /= double), 1 = true division (diagnostic).
inline bool is_inlier(const double* H, float x, float y, float qx, float qy, double thr,
                      int div_mode) {
  double X = (H[0] * x + H[1] * y) + H[2] * 1.0;
  double Y = (H[3] * x + H[4] * y) + H[5] * 1.0;
  double Wd = (H[6] * x + H[7] * y) + H[8] * 1.0;
  float ex, ey;
  if (div_mode == 0) {
    double s = 1. / Wd;
    ex = (float)(X * s);
    ey = (float)(Y * s);
  } else {
    ex = (float)(X / Wd);
    ey = (float)(Y / Wd);
  }
  float dx = ex - qx, dy = ey - qy;
  return std::sqrt((double)dx * dx + (double)dy * dy) < thr;
}

// ref: src/serial/main.cpp:247-307 (SeqRansacHomographyCalculator::computeHomography),
// seeded.  kp1 = query-side keypoints (right image), kp2 = train-side (left image).
// Optional outputs: samples[iters*4] (match indices drawn, in draw order), counts[iters]
// (inlier count, -1 where findHomography returned empty), inlier_mask[m] for the best H,
// draws = number of 32-bit engine outputs consumed.
int ransac(const int32_t* kp1, const int32_t* kp2, const Match* matches, int m, int iters,
           int nsamples, double thr, uint32_t seed, int div_mode, double* Hbest, int* best_count,
           int32_t* samples, int32_t* counts, uint8_t* inlier_mask, uint64_t* draws,
           int* best_iter) {
  struct CountingRng {  // forwards to mt19937, counts outputs (observer only)
    typedef std::mt19937::result_type result_type;
    std::mt19937 e;
    uint64_t n = 0;
    explicit CountingRng(uint32_t s) : e(s) {}
    static constexpr result_type min() { return std::mt19937::min(); }
    static constexpr result_type max() { return std::mt19937::max(); }
    result_type operator()() { ++n; return e(); }
  };
  CountingRng rng(seed);
  int bestInlierCount = 0;
  bool have = false;
  if (best_iter) *best_iter = -1;
  std::vector<std::pair<Match, int32_t>> local((size_t)m);
  for (int iter = 0; iter < iters; ++iter) {
    if (m < nsamples) break;
    for (int i = 0; i < m; i++) local[i] = {matches[i], i};
    std::shuffle(local.begin(), local.end(), rng);
    float src[8], dst[8];
    for (int j = 0; j < nsamples && j < 4; j++) {
      const Match& mm = local[j].first;
      src[2 * j] = (float)kp1[2 * mm.queryIdx];
      src[2 * j + 1] = (float)kp1[2 * mm.queryIdx + 1];
      dst[2 * j] = (float)kp2[2 * mm.trainIdx];
      dst[2 * j + 1] = (float)kp2[2 * mm.trainIdx + 1];
      if (samples) samples[iter * 4 + j] = local[j].second;
    }
    double H[9];
    if (!find_homography4(src, dst, 4, H)) {
      if (counts) counts[iter] = -1;
      continue;
    }
    int inlierCount = 0;
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : inlierCount) schedule(static)
#endif
    for (int i = 0; i < m; i++) {
      const Match& mm = matches[i];
      if (is_inlier(H, (float)kp1[2 * mm.queryIdx], (float)kp1[2 * mm.queryIdx + 1],
                    (float)kp2[2 * mm.trainIdx], (float)kp2[2 * mm.trainIdx + 1], thr, div_mode))
        inlierCount++;
    }
    if (counts) counts[iter] = inlierCount;
    if (inlierCount > bestInlierCount) {
      bestInlierCount = inlierCount;
      memcpy(Hbest, H, sizeof H);
      have = true;
      if (best_iter) *best_iter = iter;
    }
  }
  if (draws) *draws = rng.n;
  if (best_count) *best_count = bestInlierCount;
  if (have && inlier_mask) {
    for (int i = 0; i < m; i++) {
      const Match& mm = matches[i];
      inlier_mask[i] = is_inlier(Hbest, (float)kp1[2 * mm.queryIdx], (float)kp1[2 * mm.queryIdx + 1],
                                 (float)kp2[2 * mm.trainIdx], (float)kp2[2 * mm.trainIdx + 1], thr,
                                 div_mode);
    }
  }
  return have ? 1 : 0;
}


struct Canvas {
  int cw, ch, offx, offy;  // canvas size; left image ROI origin = (int)(-minX), (int)(-minY)
  double TH[9];            // translation * H
};

// ref: src/serial/main.cpp:335-369.  Returns 0 if the left ROI would not fit the canvas
// (the reference would throw from cv::Mat::operator()(Rect)).
int canvas_geometry(int wl, int hl, int wr, int hr, const double* H, Canvas* c) {
  float rc[8] = {0.f, 0.f, (float)wr, 0.f, (float)wr, (float)hr, 0.f, (float)hr};
  float wc[8];
  perspective_transform(rc, 4, H, wc);
  float lc[8] = {0.f, 0.f, (float)wl, 0.f, (float)wl, (float)hl, 0.f, (float)hl};
  float minX = 0, minY = 0, maxX = (float)wl, maxY = (float)hl;
  for (int i = 0; i < 4; i++) {
    minX = std::min(minX, wc[2 * i]); minY = std::min(minY, wc[2 * i + 1]);
    maxX = std::max(maxX, wc[2 * i]); maxY = std::max(maxY, wc[2 * i + 1]);
  }
  for (int i = 0; i < 4; i++) {
    minX = std::min(minX, lc[2 * i]); minY = std::min(minY, lc[2 * i + 1]);
    maxX = std::max(maxX, lc[2 * i]); maxY = std::max(maxY, lc[2 * i + 1]);
  }
  double T[9] = {1, 0, (double)(-minX), 0, 1, (double)(-minY), 0, 0, 1};
  mul33(T, H, c->TH);
  c->cw = (int)std::ceil(maxX - minX);
  c->ch = (int)std::ceil(maxY - minY);
  c->offx = (int)(-minX);
  c->offy = (int)(-minY);
  if (c->cw <= 0 || c->ch <= 0) return 0;
  if (c->offx < 0 || c->offy < 0 || c->offx + wl > c->cw || c->offy + hl > c->ch) return 0;
  return 1;
}


// ref: src/serial/main.cpp:371-386 — warp, left copy, "non-black overwrites" overlay.
void compose(const uint8_t* left, int wl, int hl, size_t sl, const uint8_t* right, int wr, int hr,
             size_t sr, const Canvas& c, uint8_t* canvas /* cw*ch*3 tightly packed */) {
  size_t cs = (size_t)c.cw * 3;
  std::vector<uint8_t> warped((size_t)c.ch * cs);
  warp_perspective(right, wr, hr, sr, c.TH, warped.data(), c.cw, c.ch, cs);
  memset(canvas, 0, (size_t)c.ch * cs);
  for (int y = 0; y < hl; y++) memcpy(canvas + (size_t)(y + c.offy) * cs + 3 * (size_t)c.offx, left + y * sl, (size_t)wl * 3);
  for (int y = 0; y < c.ch; y++)
    for (int x = 0; x < c.cw; x++) {
      const uint8_t* p = &warped[(size_t)y * cs + 3 * (size_t)x];
      if (p[0] | p[1] | p[2]) memcpy(canvas + (size_t)y * cs + 3 * (size_t)x, p, 3);
    }
}

}  // namespace

// =========================================================================================
// C ABI for ctypes (tests / bench only)
// =========================================================================================
extern "C" {

struct orc_dmatch { int32_t queryIdx, trainIdx; float distance; };

void orc_gray(const uint8_t* bgr, int w, int h, size_t stride, uint8_t* gray) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) gray[(size_t)y * w + x] = gray_of(bgr + y * stride + 3 * x);
}

void orc_gaussian_kernel(int ksize, double sigma, double* out) {
  std::vector<double> k;
  gaussian_kernel(ksize, sigma, k);
  memcpy(out, k.data(), k.size() * sizeof(double));
}

void orc_convolve(const double* in, int w, int h, const double* kern, int ksize, double* out) {
  convolve(in, w, h, kern, ksize, out);
}

void orc_harris_response(const uint8_t* bgr, int w, int h, size_t stride, double k, double* resp) {
  std::vector<double> r;
  harris_response(bgr, w, h, stride, k, r);
  memcpy(resp, r.data(), r.size() * sizeof(double));
}

// returns the number of keypoints; writes at most cap of them as (x, y) pairs
int orc_detect(const uint8_t* bgr, int w, int h, size_t stride, double k, double thresh, int nbhd,
               int32_t* xy, int cap) {
  std::vector<double> r;
  harris_response(bgr, w, h, stride, k, r);
  std::vector<int32_t> v;
  nms(r.data(), w, h, thresh, nbhd, v);
  int n = (int)(v.size() / 2);
  if (xy) memcpy(xy, v.data(), sizeof(int32_t) * 2 * (size_t)std::min(n, cap));
  return n;
}

int orc_match(const int32_t* kq, int nq, const int32_t* kt, int nt, const uint8_t* imq, int wq,
              int hq, size_t sq, const uint8_t* imt, int wt, int ht, size_t st, int patch,
              double maxSSD, int offset, orc_dmatch* out, int cap) {
  std::vector<Match> v;
  match_keypoints(kq, nq, kt, nt, imq, wq, hq, sq, imt, wt, ht, st, patch, maxSSD, offset, v);
  int n = (int)v.size();
  if (out) memcpy(out, v.data(), sizeof(Match) * (size_t)std::min(n, cap));
  return n;
}

// checker of pano_match_knn (not a reference function; see match_knn above).  second[] = runner-up distances.
int orc_match_knn(const int32_t* kq, int nq, const int32_t* kt, int nt, const uint8_t* imq, int wq, int hq, size_t sq,
                  const uint8_t* imt, int wt, int ht, size_t st, int patch, int descriptor, double ratio,
                  orc_dmatch* out, float* second, int cap) {
  std::vector<KnnMatch> v;
  match_knn(kq, nq, kt, nt, imq, wq, hq, sq, imt, wt, ht, st, patch, descriptor, ratio, v);
  const int n = (int)v.size();
  for (int i = 0; i < std::min(n, cap); i++) {
    if (out) out[i] = orc_dmatch{v[(size_t)i].queryIdx, v[(size_t)i].trainIdx, v[(size_t)i].distance};
    if (second) second[i] = v[(size_t)i].second;
  }
  return n;
}

// the 256-bit descriptor of one keypoint (8 words), for the descriptor-level tests
void orc_knn_binary_descriptor(const uint8_t* im, size_t stride, int x, int y, uint32_t* bits) {
  knn_binary_descriptor(im, stride, x, y, bits);
}

// cv::eigen(A symmetric n x n) -> W (descending), V (rows).  A is not modified.
void orc_eigen_sym(const double* A, int n, double* W, double* V) {
  std::vector<double> a(A, A + (size_t)n * n);
  jacobi(a.data(), n, W, V, n, n);
}

int orc_find_homography4(const float* src, const float* dst, double* H) {
  return find_homography4(src, dst, 4, H);
}

int orc_ransac(const int32_t* kp1, const int32_t* kp2, const orc_dmatch* matches, int m, int iters,
               int nsamples, double thr, uint32_t seed, int div_mode, double* H, int* best_count,
               int32_t* samples, int32_t* counts, uint8_t* inlier_mask, uint64_t* draws,
               int* best_iter) {
  return ransac(kp1, kp2, (const Match*)matches, m, iters, nsamples, thr, seed, div_mode, H,
                best_count, samples, counts, inlier_mask, draws, best_iter);
}

// libstdc++ known-answer helpers: first n outputs of mt19937(seed); shuffle(iota(n)) with a
// continuing engine (skip = outputs discarded first).
void orc_mt19937(uint32_t seed, int n, uint32_t* out) {
  std::mt19937 e(seed);
  for (int i = 0; i < n; i++) out[i] = e();
}
void orc_shuffle_iota(uint32_t seed, uint64_t skip, int n, int reps, int32_t* out_first4) {
  std::mt19937 e(seed);
  e.discard(skip);
  std::vector<int32_t> v((size_t)n);
  for (int r = 0; r < reps; r++) {
    for (int i = 0; i < n; i++) v[i] = i;
    std::shuffle(v.begin(), v.end(), e);
    for (int j = 0; j < 4; j++) out_first4[r * 4 + j] = j < n ? v[j] : -1;
  }
}

void orc_perspective_transform(const float* pts, int n, const double* H, float* out) {
  perspective_transform(pts, n, H, out);
}

int orc_invert33(const double* s, double* d) { return invert33(s, d); }
void orc_mul33(const double* a, const double* b, double* d) { mul33(a, b, d); }

// out: cw, ch, offx, offy; TH[9]
int orc_canvas_geometry(int wl, int hl, int wr, int hr, const double* H, int* geom, double* TH) {
  Canvas c;
  int ok = canvas_geometry(wl, hl, wr, hr, H, &c);
  geom[0] = c.cw; geom[1] = c.ch; geom[2] = c.offx; geom[3] = c.offy;
  memcpy(TH, c.TH, sizeof c.TH);
  return ok;
}

void orc_warp_perspective(const uint8_t* src, int sw, int sh, size_t sstride, const double* M,
                          uint8_t* dst, int dw, int dh, size_t dstride) {
  warp_perspective(src, sw, sh, sstride, M, dst, dw, dh, dstride);
}

int orc_compose(const uint8_t* left, int wl, int hl, size_t sl, const uint8_t* right, int wr, int hr,
                size_t sr, const double* H, uint8_t* canvas, size_t cap, int* geom) {
  Canvas c;
  if (!canvas_geometry(wl, hl, wr, hr, H, &c)) return 0;
  geom[0] = c.cw; geom[1] = c.ch; geom[2] = c.offx; geom[3] = c.offy;
  if ((size_t)c.cw * c.ch * 3 > cap) return -1;
  compose(left, wl, hl, sl, right, wr, hr, sr, c, canvas);
  return 1;
}

// ref: src/serial/main.cpp:311-391 stitchTwoImages.  status: 1 ok, 0 no matches,
// -2 RANSAC failed, -3 ROI does not fit, -1 canvas buffer too small.
// times_ms[4] = detect(both), match, ransac, warp+overlay (wall clock).
int orc_stitch_pair(const uint8_t* left, int wl, int hl, size_t sl, const uint8_t* right, int wr,
                    int hr, size_t sr, uint32_t seed, uint8_t* canvas, size_t cap, int* geom,
                    double* H, int* stats /* kl, kr, m, best_count */, double* times_ms) {
  auto now = [] { return std::chrono::high_resolution_clock::now(); };
  auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  auto t0 = now();
  std::vector<double> r;
  std::vector<int32_t> kl, kr;
  harris_response(left, wl, hl, sl, 0.04, r);
  nms(r.data(), wl, hl, 1e6, 3, kl);
  harris_response(right, wr, hr, sr, 0.04, r);
  nms(r.data(), wr, hr, 1e6, 3, kr);
  auto t1 = now();
  std::vector<Match> mv;
  match_keypoints(kr.data(), (int)kr.size() / 2, kl.data(), (int)kl.size() / 2, right, wr, hr, sr,
                  left, wl, hl, sl, 5, 1e8, 0, mv);
  auto t2 = now();
  if (stats) { stats[0] = (int)kl.size() / 2; stats[1] = (int)kr.size() / 2; stats[2] = (int)mv.size(); stats[3] = 0; }
  if (times_ms) { times_ms[0] = ms(t0, t1); times_ms[1] = ms(t1, t2); times_ms[2] = times_ms[3] = 0; }
  if (mv.empty()) return 0;
  int best = 0;
  int ok = ransac(kr.data(), kl.data(), mv.data(), (int)mv.size(), 1000, 4, 3.0, seed, 0, H, &best,
                  nullptr, nullptr, nullptr, nullptr, nullptr);
  auto t3 = now();
  if (stats) stats[3] = best;
  if (times_ms) times_ms[2] = ms(t2, t3);
  if (!ok) return -2;
  Canvas c;
  if (!canvas_geometry(wl, hl, wr, hr, H, &c)) return -3;
  geom[0] = c.cw; geom[1] = c.ch; geom[2] = c.offx; geom[3] = c.offy;
  if ((size_t)c.cw * c.ch * 3 > cap) return -1;
  compose(left, wl, hl, sl, right, wr, hr, sr, c, canvas);
  if (times_ms) times_ms[3] = ms(t3, now());
  return 1;
}

int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
