// cvshim.cpp — bodies of the minimal OpenCV stand-in (see cvshim.hpp).  TEST INFRASTRUCTURE ONLY.
// Every arithmetic routine forwards to oracle/cv_pinned.hpp, whose functions are pinned
// bit-for-bit against Python cv2 4.13.0 (tests/test_oracle_golden.py).
#include "cvshim.hpp"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "../cv_pinned.hpp"

namespace cv {

void error(const char* expr, const char* func, const char* file, int line) {
  std::ostringstream os;
  os << "cvshim: (-215:Assertion failed) " << expr << " in function '" << func << "' (" << file << ":" << line << ")";
  throw Exception(os.str());
}

void Mat::create(int rows_, int cols_, int type_) {
  CV_Assert(rows_ >= 0 && cols_ >= 0);
  flags = type_;
  rows = rows_;
  cols = cols_;
  step = (size_t)cols_ * elemSize();
  size_t bytes = step * (size_t)rows_;
  if (bytes == 0) { owner_.reset(); data = nullptr; return; }
  owner_ = std::shared_ptr<uchar>((uchar*)std::malloc(bytes), std::free);
  CV_Assert(owner_ != nullptr);
  data = owner_.get();
}

Mat Mat::clone() const {
  Mat m;
  if (empty()) return m;
  m.create(rows, cols, type());
  size_t rb = (size_t)cols * elemSize();
  for (int y = 0; y < rows; y++) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, rb);
  return m;
}

void Mat::copyTo(Mat& dst) const {
  if (empty()) { dst.release(); return; }
  if (dst.data == nullptr || dst.rows != rows || dst.cols != cols || dst.type() != type())
    dst.create(rows, cols, type());
  size_t rb = (size_t)cols * elemSize();
  for (int y = 0; y < rows; y++) std::memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, rb);
}

Mat& Mat::setTo(const Scalar& s) {
  int cn = channels();
  for (int y = 0; y < rows; y++) {
    if (depth() == CV_8U) {
      uchar* p = ptr<uchar>(y);
      for (int x = 0; x < cols; x++)
        for (int c = 0; c < cn; c++) p[(size_t)x * cn + c] = (uchar)std::min(255.0, std::max(0.0, std::nearbyint(s[c])));
    } else if (depth() == CV_64F) {
      double* p = ptr<double>(y);
      for (int x = 0; x < cols; x++)
        for (int c = 0; c < cn; c++) p[(size_t)x * cn + c] = s[c];
    } else if (depth() == CV_32F) {
      float* p = ptr<float>(y);
      for (int x = 0; x < cols; x++)
        for (int c = 0; c < cn; c++) p[(size_t)x * cn + c] = (float)s[c];
    } else {
      CV_Assert(!"setTo: unsupported depth");
    }
  }
  return *this;
}

// Mat::convertTo: dst = saturate_cast<rtype>(src * alpha + beta).  Paths the reference takes:
// 8U -> 64F with alpha 1, beta 0 (an exact integer -> double conversion, ref: :129) and
// 64F -> 64F scaling (Mat /= s, ref: :289).
void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const {
  int ddepth = rtype < 0 ? depth() : (rtype & 7);
  int cn = channels();
  Mat out(rows, cols, CV_MAKETYPE(ddepth, cn));
  bool noscale = alpha == 1 && beta == 0;
  for (int y = 0; y < rows; y++) {
    size_t n = (size_t)cols * cn;
    if (depth() == CV_8U && ddepth == CV_64F) {
      const uchar* s = ptr<uchar>(y);
      double* d = out.ptr<double>(y);
      for (size_t i = 0; i < n; i++) d[i] = noscale ? (double)s[i] : s[i] * alpha + beta;
    } else if (depth() == CV_64F && ddepth == CV_64F) {
      const double* s = ptr<double>(y);
      double* d = out.ptr<double>(y);
      for (size_t i = 0; i < n; i++) d[i] = noscale ? s[i] : s[i] * alpha + beta;
    } else if (depth() == CV_8U && ddepth == CV_8U && noscale) {
      std::memcpy(out.ptr<uchar>(y), ptr<uchar>(y), n);
    } else {
      CV_Assert(!"convertTo: conversion not on the stitching path");
    }
  }
  dst = out;
}

Mat Mat::mul(const Mat& m, double scale) const {
  CV_Assert(type() == CV_64FC1 && m.type() == CV_64FC1 && rows == m.rows && cols == m.cols && scale == 1);
  Mat out(rows, cols, CV_64FC1);
  for (int y = 0; y < rows; y++) {
    const double* a = ptr<double>(y);
    const double* b = m.ptr<double>(y);
    double* d = out.ptr<double>(y);
    for (int x = 0; x < cols; x++) d[x] = a[x] * b[x];
  }
  return out;
}

Mat Mat::operator()(const Rect& roi) const {
  // core/src/matrix.cpp Mat::Mat(const Mat& m, const Rect& roi)
  CV_Assert(0 <= roi.x && 0 <= roi.width && roi.x + roi.width <= cols && 0 <= roi.y && 0 <= roi.height &&
            roi.y + roi.height <= rows);
  Mat r = *this;  // shares the pixels
  r.data = data + (size_t)roi.y * step + (size_t)roi.x * elemSize();
  r.rows = roi.height;
  r.cols = roi.width;
  return r;
}

// cv::gemm, CV_64F, alpha 1, no C: the small-matrix path taken for an inner length of 2..4
// (core/src/matmul.dispatch.cpp) sums the products left to right with separately rounded
// multiplies and adds.  Other sizes are not on the stitching path.
Mat operator*(const Mat& a, const Mat& b) {
  CV_Assert(a.type() == CV_64FC1 && b.type() == CV_64FC1 && a.cols == b.rows);
  int len = a.cols;
  CV_Assert(2 <= len && len <= 4 && (len == b.cols || len == a.rows));
  Mat d(a.rows, b.cols, CV_64FC1);
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < b.cols; j++) {
      double t = a.at<double>(i, 0) * b.at<double>(0, j);
      for (int k = 1; k < len; k++) t = t + a.at<double>(i, k) * b.at<double>(k, j);
      d.at<double>(i, j) = t;
    }
  return d;
}

// core/mat.inl.hpp: Mat& operator /= (Mat& a, double b) { a.convertTo(a, -1, 1./b); return a; }
Mat& operator/=(Mat& a, double s) {
  a.convertTo(a, -1, 1. / s);
  return a;
}

void cvtColor(const Mat& src, Mat& dst, int code) {
  CV_Assert(code == COLOR_BGR2GRAY && src.type() == CV_8UC3);
  Mat out(src.rows, src.cols, CV_8UC1);
  for (int y = 0; y < src.rows; y++) {
    const uchar* s = src.ptr<uchar>(y);
    uchar* d = out.ptr<uchar>(y);
    for (int x = 0; x < src.cols; x++) d[x] = cvpin::gray_of(s + 3 * (size_t)x);
  }
  dst = out;
}

// cv::findHomography(src, dst) with method 0: for exactly 4 points OpenCV runs the minimal
// solver once and does no refinement (calib3d/src/fundam.cpp: `method == 0 || npoints == 4`,
// LM only if npoints > 4).  Other point counts are not on the stitching path (numSamples_ = 4).
Mat findHomography(const std::vector<Point2f>& srcPoints, const std::vector<Point2f>& dstPoints, int method, double) {
  CV_Assert(method == 0 && srcPoints.size() == dstPoints.size() && srcPoints.size() == 4);
  float s[8], d[8];
  for (int i = 0; i < 4; i++) {
    s[2 * i] = srcPoints[i].x; s[2 * i + 1] = srcPoints[i].y;
    d[2 * i] = dstPoints[i].x; d[2 * i + 1] = dstPoints[i].y;
  }
  double H[9];
  if (!cvpin::find_homography4(s, d, 4, H)) return Mat();
  Mat out(3, 3, CV_64FC1);
  std::memcpy(out.data, H, sizeof H);
  return out;
}

void perspectiveTransform(const std::vector<Point2f>& src, std::vector<Point2f>& dst, const Mat& m) {
  CV_Assert(m.type() == CV_64FC1 && m.rows == 3 && m.cols == 3 && m.isContinuous());
  std::vector<Point2f> out(src.size());
  static_assert(sizeof(Point2f) == 2 * sizeof(float), "Point2f layout");
  cvpin::perspective_transform((const float*)src.data(), (int)src.size(), m.ptr<double>(0), (float*)out.data());
  dst.swap(out);
}

void warpPerspective(const Mat& src, Mat& dst, const Mat& M, Size dsize, int flags, int borderMode, const Scalar&) {
  CV_Assert(src.type() == CV_8UC3 && M.type() == CV_64FC1 && M.rows == 3 && M.cols == 3 && M.isContinuous());
  CV_Assert(flags == INTER_LINEAR && borderMode == BORDER_CONSTANT && dsize.width > 0 && dsize.height > 0);
  Mat out(dsize, src.type());
  cvpin::warp_perspective(src.data, src.cols, src.rows, src.step, M.ptr<double>(0), out.data, out.cols, out.rows, out.step);
  dst = out;
}

// ---- image files: binary PPM/PGM and uncompressed 24-bit BMP (no codec libraries in this image) ----
static bool ends_with(const std::string& s, const char* suf) {
  size_t n = std::strlen(suf);
  if (s.size() < n) return false;
  for (size_t i = 0; i < n; i++)
    if (std::tolower((unsigned char)s[s.size() - n + i]) != suf[i]) return false;
  return true;
}

static int pnm_int(std::istream& f) {
  int c;
  for (;;) {
    c = f.peek();
    if (c == '#') { std::string l; std::getline(f, l); }
    else if (std::isspace(c)) f.get();
    else break;
  }
  int v = -1;
  f >> v;
  return v;
}

Mat imread(const std::string& filename, int) {
  std::ifstream f(filename, std::ios::binary);
  if (!f) return Mat();
  char m0 = 0, m1 = 0;
  f.get(m0); f.get(m1);
  if (m0 == 'P' && (m1 == '6' || m1 == '5')) {
    int w = pnm_int(f), h = pnm_int(f), mx = pnm_int(f);
    if (w <= 0 || h <= 0 || mx != 255) return Mat();
    f.get();  // the single whitespace after maxval
    int cn = m1 == '6' ? 3 : 1;
    std::vector<uchar> row((size_t)w * cn);
    Mat img(h, w, CV_8UC3);
    for (int y = 0; y < h; y++) {
      f.read((char*)row.data(), row.size());
      if (!f) return Mat();
      uchar* d = img.ptr<uchar>(y);
      for (int x = 0; x < w; x++) {
        if (cn == 3) { d[3 * x] = row[3 * x + 2]; d[3 * x + 1] = row[3 * x + 1]; d[3 * x + 2] = row[3 * x]; }  // RGB -> BGR
        else d[3 * x] = d[3 * x + 1] = d[3 * x + 2] = row[x];
      }
    }
    return img;
  }
  if (m0 == 'B' && m1 == 'M') {
    uchar hd[52];
    f.read((char*)hd, sizeof hd);
    if (!f) return Mat();
    auto u32 = [&](int o) { return (uint32_t)hd[o] | (uint32_t)hd[o + 1] << 8 | (uint32_t)hd[o + 2] << 16 | (uint32_t)hd[o + 3] << 24; };
    uint32_t off = u32(8);
    int w = (int)u32(16), h = (int)u32(20);
    int bpp = hd[26] | hd[27] << 8;
    if (bpp != 24 || u32(28) != 0 || w <= 0 || h == 0) return Mat();
    bool flip = h > 0;
    if (h < 0) h = -h;
    size_t rb = ((size_t)w * 3 + 3) & ~(size_t)3;
    std::vector<uchar> row(rb);
    Mat img(h, w, CV_8UC3);
    f.seekg(off);
    for (int y = 0; y < h; y++) {
      f.read((char*)row.data(), rb);
      if (!f) return Mat();
      std::memcpy(img.ptr<uchar>(flip ? h - 1 - y : y), row.data(), (size_t)w * 3);
    }
    return img;
  }
  return Mat();
}

bool imwrite(const std::string& filename, const Mat& img) {
  if (img.empty() || img.type() != CV_8UC3) return false;
  if (ends_with(filename, ".bmp")) {
    size_t rb = ((size_t)img.cols * 3 + 3) & ~(size_t)3;
    uint32_t size = 54 + (uint32_t)(rb * img.rows);
    uchar hd[54] = {'B', 'M'};
    auto put = [&](int o, uint32_t v) { hd[o] = v & 255; hd[o + 1] = (v >> 8) & 255; hd[o + 2] = (v >> 16) & 255; hd[o + 3] = v >> 24; };
    put(2, size); put(10, 54); put(14, 40); put(18, (uint32_t)img.cols); put(22, (uint32_t)img.rows);
    hd[26] = 1; hd[28] = 24; put(34, (uint32_t)(rb * img.rows));
    std::ofstream f(filename, std::ios::binary);
    if (!f) return false;
    f.write((char*)hd, 54);
    std::vector<uchar> row(rb, 0);
    for (int y = img.rows - 1; y >= 0; y--) {
      std::memcpy(row.data(), img.ptr<uchar>(y), (size_t)img.cols * 3);
      f.write((char*)row.data(), rb);
    }
    return (bool)f;
  }
  if (!ends_with(filename, ".ppm") && !ends_with(filename, ".pnm"))
    std::fprintf(stderr, "cvshim: no codec for '%s' here; writing binary PPM data under that name\n", filename.c_str());
  std::ofstream f(filename, std::ios::binary);
  if (!f) return false;
  f << "P6\n" << img.cols << " " << img.rows << "\n255\n";
  std::vector<uchar> row((size_t)img.cols * 3);
  for (int y = 0; y < img.rows; y++) {
    const uchar* s = img.ptr<uchar>(y);
    for (int x = 0; x < img.cols; x++) { row[3 * x] = s[3 * x + 2]; row[3 * x + 1] = s[3 * x + 1]; row[3 * x + 2] = s[3 * x]; }
    f.write((char*)row.data(), row.size());
  }
  return (bool)f;
}

}  // namespace cv
