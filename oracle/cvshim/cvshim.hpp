// cvshim.hpp — a minimal stand-in for the OpenCV C++ API, TEST INFRASTRUCTURE ONLY.
//
// OpenCV C++ (headers, libraries) is not installed in this image and cannot be fetched, so the
// reference (`/root/reference/src/{serial,openmp}/main.cpp`, `src/reader/reader.cpp`) cannot be
// compiled against the real thing.  This header provides exactly the slice of `cv::` those three
// files use, so that they compile UNMODIFIED, from where they lie, into `oracle/_ref/`
// (recipe: oracle/Makefile, target `_ref`).  What that buys: the reference's own control flow —
// its convolution loops, NMS scan, matcher loops, RANSAC loop with the real libstdc++
// std::shuffle, canvas geometry, ROI copy and overlay loop — is EXECUTED, not restated.
//
// The OpenCV routines the reference calls are implemented in cvshim.cpp on top of
// oracle/cv_pinned.hpp, i.e. by the restatements that are pinned bit-for-bit against Python
// cv2 4.13.0 (tests/test_oracle_golden.py):
//   cvtColor(BGR2GRAY)          ref call: src/serial/main.cpp:125
//   findHomography (4 points)   ref call: src/serial/main.cpp:279
//   Mat * Mat (gemm)            ref call: src/serial/main.cpp:288, :371 (translation * H)
//   Mat /= double               ref call: src/serial/main.cpp:289   (OpenCV: convertTo(-1, 1./s))
//   cv::norm(Point2f)           ref call: src/serial/main.cpp:292   (sqrt((double)x*x + (double)y*y))
//   perspectiveTransform        ref call: src/serial/main.cpp:342
//   warpPerspective             ref call: src/serial/main.cpp:372
//   Mat(Rect) / copyTo          ref call: src/serial/main.cpp:376-377
//   imread / imwrite            ref call: src/reader/reader.cpp:61,72, src/serial/main.cpp:445
//                               (binary PPM / PGM and uncompressed 24-bit BMP only: no codecs here)
// Types follow OpenCV's public layout where the reference touches it (Mat::rows/cols/data/step,
// KeyPoint::pt, DMatch::queryIdx/trainIdx/distance, Point_/Size_/Rect_/Vec/Scalar_).
#pragma once
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;  // core/hal/interface.h defines it at global scope

namespace cv {

using ::uchar;

// ---- type codes (core/hal/interface.h) ---------------------------------------------------
enum { CV_8U_ = 0, CV_32F_ = 5, CV_64F_ = 6 };
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

class Exception : public std::runtime_error {
 public:
  explicit Exception(const std::string& m) : std::runtime_error(m) {}
};
[[noreturn]] void error(const char* expr, const char* func, const char* file, int line);
#define CV_Assert(expr) \
  do { if (!(expr)) ::cv::error(#expr, __func__, __FILE__, __LINE__); } while (0)

// ---- small value types (core/types.hpp, core/matx.hpp) -----------------------------------
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> explicit Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
};
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) {
  return Point_<T>((T)(a.x - b.x), (T)(a.y - b.y));  // saturate_cast<float> of a float difference
}
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) {
  return Point_<T>((T)(a.x + b.x), (T)(a.y + b.y));
}
template <typename T> inline bool operator==(const Point_<T>& a, const Point_<T>& b) {
  return a.x == b.x && a.y == b.y;
}
// core/types.hpp: template<typename _Tp> double norm(const Point_<_Tp>& pt)
template <typename T> inline double norm(const Point_<T>& pt) {
  return std::sqrt((double)pt.x * pt.x + (double)pt.y * pt.y);
}
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
  bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size_& o) const { return !(*this == o); }
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
};
typedef Rect_<int> Rect;

template <typename T, int N> struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
  Vec(T a, T b, T c) { static_assert(N == 3, "3-element constructor"); val[0] = a; val[1] = b; val[2] = c; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
  bool operator==(const Vec& o) const { for (int i = 0; i < N; i++) if (val[i] != o.val[i]) return false; return true; }
  bool operator!=(const Vec& o) const { return !(*this == o); }
};
typedef Vec<uchar, 3> Vec3b;

template <typename T> struct Scalar_ {
  T val[4];
  Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
  Scalar_(T v0, T v1 = 0, T v2 = 0, T v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
  static Scalar_ all(T v) { return Scalar_(v, v, v, v); }
  T operator[](int i) const { return val[i]; }
};
typedef Scalar_<double> Scalar;

struct KeyPoint {  // features2d / core/types.hpp
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

struct DMatch {  // core/types.hpp
  int queryIdx, trainIdx, imgIdx;
  float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38f) {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};

// ---- Mat ---------------------------------------------------------------------------------
template <typename T> class MatCommaInitializer_;

class Mat {
 public:
  int flags = 0;  // the type code only
  int rows = 0, cols = 0;
  uchar* data = nullptr;
  size_t step = 0;  // bytes per row

  Mat() {}
  Mat(int rows_, int cols_, int type_) { create(rows_, cols_, type_); }
  Mat(Size sz, int type_) { create(sz.height, sz.width, type_); }
  Mat(int rows_, int cols_, int type_, const Scalar& s) { create(rows_, cols_, type_); setTo(s); }
  Mat(Size sz, int type_, const Scalar& s) { create(sz.height, sz.width, type_); setTo(s); }
  // user-allocated data (not owned, never freed) — used by the bridge to wrap caller buffers
  Mat(int rows_, int cols_, int type_, void* data_, size_t step_ = 0)
      : flags(type_), rows(rows_), cols(cols_), data((uchar*)data_) {
    step = step_ ? step_ : (size_t)cols_ * elemSize();
  }
  template <typename T> Mat(const MatCommaInitializer_<T>& ci);

  void create(int rows_, int cols_, int type_);
  void release() { *this = Mat(); }
  int type() const { return flags; }
  int depth() const { return flags & 7; }
  int channels() const { return (flags >> CV_CN_SHIFT) + 1; }
  size_t elemSize1() const { int d = depth(); return d == CV_8U ? 1 : d == CV_32F ? 4 : 8; }
  size_t elemSize() const { return elemSize1() * (size_t)channels(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  Size size() const { return Size(cols, rows); }
  size_t total() const { return (size_t)rows * cols; }
  bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }

  template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
  template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }

  Mat clone() const;
  void copyTo(Mat& dst) const;  // writes in place when dst already has the size and type (ROI copy)
  void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;
  Mat mul(const Mat& m, double scale = 1) const;
  Mat operator()(const Rect& roi) const;  // shares the pixels; throws cv::Exception if out of range
  Mat& setTo(const Scalar& s);

 private:
  std::shared_ptr<uchar> owner_;  // atomic refcount: the OpenMP reference copies Mats across threads
};

template <typename T> class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int rows_, int cols_);
  T& operator()(int y, int x) { return this->template at<T>(y, x); }
  const T& operator()(int y, int x) const { return this->template at<T>(y, x); }
};
template <> inline Mat_<double>::Mat_(int rows_, int cols_) : Mat(rows_, cols_, CV_64FC1) {}
template <> inline Mat_<float>::Mat_(int rows_, int cols_) : Mat(rows_, cols_, CV_32FC1) {}
template <> inline Mat_<uchar>::Mat_(int rows_, int cols_) : Mat(rows_, cols_, CV_8UC1) {}

// (Mat_<T>(r, c) << a, b, c ...)  — core/mat.hpp MatCommaInitializer_
template <typename T> class MatCommaInitializer_ {
 public:
  explicit MatCommaInitializer_(const Mat_<T>& m) : m_(m), i_(0) {}
  template <typename U> MatCommaInitializer_& operator,(U v) {
    CV_Assert(i_ < m_.total());
    m_.template at<T>((int)(i_ / m_.cols), (int)(i_ % m_.cols)) = (T)v;
    ++i_;
    return *this;
  }
  const Mat_<T>& mat() const { CV_Assert(i_ == m_.total()); return m_; }
  operator Mat_<T>() const { return mat(); }

 private:
  Mat_<T> m_;
  size_t i_;
};
template <typename T, typename U> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v) {
  MatCommaInitializer_<T> ci(m);
  return (ci, v);
}
template <typename T> inline Mat::Mat(const MatCommaInitializer_<T>& ci) { *this = (const Mat&)ci.mat(); }

Mat operator*(const Mat& a, const Mat& b);  // gemm (CV_64F), OpenCV's small-matrix order
Mat& operator/=(Mat& a, double s);          // OpenCV: a.convertTo(a, -1, 1./s)

// ---- imgproc / calib3d / core routines on the path ------------------------------------------
enum { COLOR_BGR2GRAY = 6 };
enum { INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0 };
enum { IMREAD_COLOR = 1 };

void cvtColor(const Mat& src, Mat& dst, int code);
Mat findHomography(const std::vector<Point2f>& srcPoints, const std::vector<Point2f>& dstPoints,
                   int method = 0, double ransacReprojThreshold = 3);
void perspectiveTransform(const std::vector<Point2f>& src, std::vector<Point2f>& dst, const Mat& m);
void warpPerspective(const Mat& src, Mat& dst, const Mat& M, Size dsize, int flags = INTER_LINEAR,
                     int borderMode = BORDER_CONSTANT, const Scalar& borderValue = Scalar());
Mat imread(const std::string& filename, int flags = IMREAD_COLOR);
bool imwrite(const std::string& filename, const Mat& img);

}  // namespace cv
