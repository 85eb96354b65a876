// cv_pinned.hpp — the OpenCV / third-party arithmetic on the stitching path, restated
// once and PINNED bit-for-bit against Python cv2 4.13.0 (oracle/gen_golden.py ->
// tests/golden/opencv_pins.npz, tests/test_oracle_golden.py).  TEST INFRASTRUCTURE ONLY.
//
// Two users, both under oracle/:
//   * pano_oracle.cpp  — the OpenCV-free restatement of the reference's serial path;
//   * cvshim/          — the minimal `opencv2/` stand-in that lets the UNMODIFIED reference
//                        sources (/root/reference/src/{serial,openmp}/main.cpp) compile into
//                        oracle/_ref/ (OpenCV C++ is not installed in this image).
// Nothing in the product path (the package's csrc/, the C ABI, the CLI) includes this file.
//
// Routines (OpenCV 4.x sources they restate):
//   gray_of               cvtColor(BGR2GRAY), 8-bit: imgproc/src/color_rgb.simd.hpp (15-bit fixed point)
//   cv_hypot, jacobi      cv::eigen without Eigen/LAPACK: core/src/lapack.cpp JacobiImpl_
//   mul33, gemm_small     cv::gemm small-matrix path: core/src/matmul.simd.hpp
//   find_homography4      cv::findHomography, 4 points: calib3d/src/fundam.cpp runKernel
//   perspective_transform cv::perspectiveTransform (32f points, 64f matrix): core/src/matmul.simd.hpp
//   invert33              cv::invert 3x3 CV_64F: core/src/lapack.cpp
//   warp_perspective      cv::warpPerspective INTER_LINEAR / BORDER_CONSTANT(0), 8UC3:
//                         imgproc/src/imgwarp.cpp (WarpPerspectiveInvoker + remapBilinear)
#pragma once
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace cvpin {

// ---------------------------------------------------------------------------------------
// cv::cvtColor(BGR2GRAY) for 8-bit input: 15-bit fixed point, coefficients B 3735, G 19235,
// R 9798 (sum 32768), rounding constant 1<<14.   ref: src/serial/main.cpp:123-129
// ---------------------------------------------------------------------------------------
inline uint8_t gray_of(const uint8_t* p) {
  return (uint8_t)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15);
}

// ---------------------------------------------------------------------------------------
// cv::eigen for a symmetric n x n double matrix when OpenCV is built without Eigen/LAPACK
// eigen solvers: cyclic-by-pivot Jacobi (OpenCV modules/core/src/lapack.cpp JacobiImpl_).
// Eigenvalues sorted descending, eigenvectors as rows of V.
// ---------------------------------------------------------------------------------------
// OpenCV's own hypot (core/src/lapack.cpp), used by its Jacobi instead of libm's.
inline double cv_hypot(double a, double b) {
  a = std::abs(a);
  b = std::abs(b);
  if (a > b) {
    b /= a;
    return a * std::sqrt(1 + b * b);
  }
  if (b > 0) {
    a /= b;
    return b * std::sqrt(1 + a * a);
  }
  return 0;
}

inline void jacobi(double* A, int astep, double* W, double* V, int vstep, int n) {
  const double eps = std::numeric_limits<double>::epsilon();
  int i, j, k, m;
  for (i = 0; i < n; i++) {
    for (j = 0; j < n; j++) V[i * vstep + j] = 0;
    V[i * vstep + i] = 1;
  }
  int iters, maxIters = n * n * 30;
  std::vector<int> indRv(n), indCv(n);
  int* indR = indRv.data();
  int* indC = indCv.data();
  double mv = 0;
  for (k = 0; k < n; k++) {
    W[k] = A[(astep + 1) * k];
    if (k < n - 1) {
      for (m = k + 1, mv = std::abs(A[astep * k + m]), i = k + 2; i < n; i++) {
        double val = std::abs(A[astep * k + i]);
        if (mv < val) mv = val, m = i;
      }
      indR[k] = m;
    }
    if (k > 0) {
      for (m = 0, mv = std::abs(A[k]), i = 1; i < k; i++) {
        double val = std::abs(A[astep * i + k]);
        if (mv < val) mv = val, m = i;
      }
      indC[k] = m;
    }
  }
  if (n > 1)
    for (iters = 0; iters < maxIters; iters++) {
      for (k = 0, mv = std::abs(A[indR[0]]), i = 1; i < n - 1; i++) {
        double val = std::abs(A[astep * i + indR[i]]);
        if (mv < val) mv = val, k = i;
      }
      int l = indR[k];
      for (i = 1; i < n; i++) {
        double val = std::abs(A[astep * indC[i] + i]);
        if (mv < val) mv = val, k = indC[i], l = i;
      }
      double p = A[astep * k + l];
      if (std::abs(p) <= eps) break;
      double y = (W[l] - W[k]) * 0.5;
      double t = std::abs(y) + cv_hypot(p, y);
      double s = cv_hypot(p, t);
      double c = t / s;
      s = p / s;
      t = (p / t) * p;
      if (y < 0) s = -s, t = -t;
      A[astep * k + l] = 0;
      W[k] -= t;
      W[l] += t;
      double a0, b0;
#define ROT(v0, v1) a0 = v0, b0 = v1, v0 = a0 * c - b0 * s, v1 = a0 * s + b0 * c
      for (i = 0; i < k; i++) ROT(A[astep * i + k], A[astep * i + l]);
      for (i = k + 1; i < l; i++) ROT(A[astep * k + i], A[astep * i + l]);
      for (i = l + 1; i < n; i++) ROT(A[astep * k + i], A[astep * l + i]);
      for (i = 0; i < n; i++) ROT(V[vstep * k + i], V[vstep * l + i]);
#undef ROT
      for (j = 0; j < 2; j++) {
        int idx = j == 0 ? k : l;
        if (idx < n - 1) {
          for (m = idx + 1, mv = std::abs(A[astep * idx + m]), i = idx + 2; i < n; i++) {
            double val = std::abs(A[astep * idx + i]);
            if (mv < val) mv = val, m = i;
          }
          indR[idx] = m;
        }
        if (idx > 0) {
          for (m = 0, mv = std::abs(A[idx]), i = 1; i < idx; i++) {
            double val = std::abs(A[astep * i + idx]);
            if (mv < val) mv = val, m = i;
          }
          indC[idx] = m;
        }
      }
    }
  for (k = 0; k < n - 1; k++) {
    m = k;
    for (i = k + 1; i < n; i++)
      if (W[m] < W[i]) m = i;
    if (k != m) {
      std::swap(W[m], W[k]);
      for (i = 0; i < n; i++) std::swap(V[vstep * m + i], V[vstep * k + i]);
    }
  }
}

// OpenCV gemm small-matrix path (3x3 * 3x3, alpha 1, no C): left-to-right, no FMA.
inline void mul33(const double* a, const double* b, double* d) {
  double r[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      r[i * 3 + j] = (a[i * 3 + 0] * b[0 + j] + a[i * 3 + 1] * b[3 + j]) + a[i * 3 + 2] * b[6 + j];
  memcpy(d, r, sizeof r);
}

// cv::findHomography(src, dst) with method 0 and exactly 4 points == one call of
// HomographyEstimatorCallback::runKernel (OpenCV calib3d/src/fundam.cpp): normalised DLT,
// 9x9 LtL, smallest eigenvector via cv::eigen, de-normalise, scale by 1/h22.
// Returns 0 (empty Mat) iff a mean-abs-deviation sum is < DBL_EPSILON.
// ref call site: src/serial/main.cpp:279.  M = src (points1), m = dst (points2).
inline int find_homography4(const float* M, const float* m, int count, double* Hout) {
  double LtL[9][9], W[9], V[9][9];
  double cMx = 0, cMy = 0, cmx = 0, cmy = 0, sMx = 0, sMy = 0, smx = 0, smy = 0;
  for (int i = 0; i < count; i++) {
    cmx += m[2 * i]; cmy += m[2 * i + 1];
    cMx += M[2 * i]; cMy += M[2 * i + 1];
  }
  cmx /= count; cmy /= count; cMx /= count; cMy /= count;
  for (int i = 0; i < count; i++) {
    smx += fabs(m[2 * i] - cmx);
    smy += fabs(m[2 * i + 1] - cmy);
    sMx += fabs(M[2 * i] - cMx);
    sMy += fabs(M[2 * i + 1] - cMy);
  }
  if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON ||
      fabs(sMy) < DBL_EPSILON)
    return 0;
  smx = count / smx; smy = count / smy;
  sMx = count / sMx; sMy = count / sMy;
  double invHnorm[9] = {1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1};
  double Hnorm2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
  memset(LtL, 0, sizeof LtL);
  for (int i = 0; i < count; i++) {
    double x = (m[2 * i] - cmx) * smx, y = (m[2 * i + 1] - cmy) * smy;
    double X = (M[2 * i] - cMx) * sMx, Y = (M[2 * i + 1] - cMy) * sMy;
    double Lx[] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
    double Ly[] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
    for (int j = 0; j < 9; j++)
      for (int k = j; k < 9; k++) LtL[j][k] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
  }
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < i; j++) LtL[i][j] = LtL[j][i];  // completeSymm (upper -> lower)
  jacobi(&LtL[0][0], 9, W, &V[0][0], 9, 9);
  double Htemp[9], H0[9];
  mul33(invHnorm, V[8], Htemp);
  mul33(Htemp, Hnorm2, H0);
  double sc = 1. / H0[8];
  for (int i = 0; i < 9; i++) Hout[i] = H0[i] * sc;
  return 1;
}

// cv::perspectiveTransform for Point2f input and a 3x3 double matrix
// (OpenCV core/src/matmul.simd.hpp perspectiveTransform_32f).  ref: src/serial/main.cpp:342
inline void perspective_transform(const float* src, int n, const double* m, float* dst) {
  const double eps = FLT_EPSILON;
  for (int i = 0; i < n; i++) {
    float x = src[2 * i], y = src[2 * i + 1];
    double w = x * m[6] + y * m[7] + m[8];
    if (fabs(w) > eps) {
      w = 1. / w;
      dst[2 * i] = (float)((x * m[0] + y * m[1] + m[2]) * w);
      dst[2 * i + 1] = (float)((x * m[3] + y * m[4] + m[5]) * w);
    } else
      dst[2 * i] = dst[2 * i + 1] = 0.f;
  }
}

// cv::invert for 3x3 CV_64F (DECOMP_LU fast path, OpenCV core/src/lapack.cpp).
inline int invert33(const double* s, double* d) {
  double det = s[0] * (s[4] * s[8] - s[5] * s[7]) - s[1] * (s[3] * s[8] - s[5] * s[6]) +
               s[2] * (s[3] * s[7] - s[4] * s[6]);
  if (det == 0.) return 0;
  det = 1. / det;
  double t[9];
  t[0] = (s[4] * s[8] - s[5] * s[7]) * det;
  t[1] = (s[2] * s[7] - s[1] * s[8]) * det;
  t[2] = (s[1] * s[5] - s[2] * s[4]) * det;
  t[3] = (s[5] * s[6] - s[3] * s[8]) * det;
  t[4] = (s[0] * s[8] - s[2] * s[6]) * det;
  t[5] = (s[2] * s[3] - s[0] * s[5]) * det;
  t[6] = (s[3] * s[7] - s[4] * s[6]) * det;
  t[7] = (s[1] * s[6] - s[0] * s[7]) * det;
  t[8] = (s[0] * s[4] - s[1] * s[3]) * det;
  memcpy(d, t, sizeof t);
  return 1;
}

inline int cv_round(double v) { return (int)lrint(v); }  // SSE2 cvtsd2si, ties-to-even
inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

// cv::warpPerspective(src, dst, M, dsize) with INTER_LINEAR, BORDER_CONSTANT(0), 8UC3
// (OpenCV imgproc/src/imgwarp.cpp WarpPerspectiveInvoker + remapBilinear fixed point):
// M is inverted; the destination is processed in blocks (bw x bh); per block row the
// coordinate numerators are X0 = M0*x + M1*(y+y1) + M2 at the block origin x, then
// (X0 + M0*x1) * (32 / (W0 + M6*x1)) is rounded to an integer in 1/32-px units; bilinear
// weights are the exact products (32-fx)(32-fy)*32 ... (15-bit), result (sum + 2^14) >> 15;
// taps outside the source read 0.      ref call site: src/serial/main.cpp:371-372
inline void warp_perspective(const uint8_t* src, int sw, int sh, size_t sstride, const double* M0,
                      uint8_t* dst, int dw, int dh, size_t dstride) {
  double M[9];
  if (!invert33(M0, M)) { memset(M, 0, sizeof M); }
  const int BLOCK_SZ = 32;
  int bh0 = std::min(BLOCK_SZ / 2, dh);
  int bw0 = std::min(BLOCK_SZ * BLOCK_SZ / bh0, dw);
  bh0 = std::min(BLOCK_SZ * BLOCK_SZ / bw0, dh);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int y = 0; y < dh; y += bh0) {
    for (int x = 0; x < dw; x += bw0) {
      int bw = std::min(bw0, dw - x);
      int bh = std::min(bh0, dh - y);
      for (int y1 = 0; y1 < bh; y1++) {
        double X0 = M[0] * x + M[1] * (y + y1) + M[2];
        double Y0 = M[3] * x + M[4] * (y + y1) + M[5];
        double W0 = M[6] * x + M[7] * (y + y1) + M[8];
        uint8_t* drow = dst + (size_t)(y + y1) * dstride + 3 * (size_t)x;
        for (int x1 = 0; x1 < bw; x1++) {
          double W = W0 + M[6] * x1;
          W = W ? 32. / W : 0;
          double fX = std::max((double)INT_MIN, std::min((double)INT_MAX, (X0 + M[0] * x1) * W));
          double fY = std::max((double)INT_MIN, std::min((double)INT_MAX, (Y0 + M[3] * x1) * W));
          int X = cv_round(fX), Y = cv_round(fY);
          int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
          int fx = X & 31, fy = Y & 31;
          int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
          int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
          for (int c = 0; c < 3; c++) {
            auto tap = [&](int yy, int xx) -> int {
              return (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[(size_t)yy * sstride + 3 * xx + c] : 0;
            };
            int v = tap(sy, sx) * w00 + tap(sy, sx + 1) * w01 + tap(sy + 1, sx) * w10 +
                    tap(sy + 1, sx + 1) * w11;
            drow[3 * x1 + c] = (uint8_t)((v + (1 << 14)) >> 15);
          }
        }
      }
    }
  }
}

}  // namespace cvpin
