// pano_core.cuh — order-exact arithmetic shared by the CUDA kernels.
//
// Every function here reproduces, operation by operation, a piece of the reference's serial
// path (ref: src/serial/main.cpp) or of the OpenCV routine it calls, so that device results
// are bit-identical to the CPU reference.  All FP64 steps go through PANO_DMUL/PANO_DADD/...
// which map to the round-to-nearest intrinsics on the device (never contracted into FMA).
// The same header compiles for the host (tests/hostsim) where the macros are plain
// operators and the translation unit is built with -ffp-contract=off.
#pragma once
#include <cstdint>
#include <cstring>
#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define PANO_HD __host__ __device__ __forceinline__
#else
#define PANO_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define PANO_DMUL(a, b) __dmul_rn((a), (b))
#define PANO_DADD(a, b) __dadd_rn((a), (b))
#define PANO_DSUB(a, b) __dsub_rn((a), (b))
#define PANO_DDIV(a, b) __ddiv_rn((a), (b))
#define PANO_DSQRT(a) __dsqrt_rn((a))
#define PANO_FSUB(a, b) __fsub_rn((a), (b))
#define PANO_D2F(a) __double2float_rn((a))
#define PANO_RINT_I(a) __double2int_rn((a))
#define PANO_DFMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define PANO_DMUL(a, b) ((a) * (b))
#define PANO_DADD(a, b) ((a) + (b))
#define PANO_DSUB(a, b) ((a) - (b))
#define PANO_DDIV(a, b) ((a) / (b))
#define PANO_DSQRT(a) (std::sqrt((a)))
#define PANO_FSUB(a, b) ((a) - (b))
#define PANO_D2F(a) ((float)(a))
#define PANO_RINT_I(a) ((int)lrint((a)))
#define PANO_DFMA(a, b, c) (std::fma((a), (b), (c)))
#endif

namespace pano {

// ----------------------------------------------------------------------------------------
// cvtColor(BGR2GRAY), 8-bit: 15-bit fixed point (OpenCV color_yuv / RGB2Gray<uchar>).
// ref: src/serial/main.cpp:125
// ----------------------------------------------------------------------------------------
PANO_HD int gray_u8(int b, int g, int r) { return (b * 3735 + g * 19235 + r * 9798 + 16384) >> 15; }

// Harris response from the three Gaussian-smoothed products.  ref: src/serial/main.cpp:151-153
PANO_HD double harris_resp(double xx, double yy, double xy, double k) {
  double det = PANO_DSUB(PANO_DMUL(xx, yy), PANO_DMUL(xy, xy));
  double trace = PANO_DADD(xx, yy);
  return PANO_DSUB(det, PANO_DMUL(PANO_DMUL(k, trace), trace));
}

// ----------------------------------------------------------------------------------------
// OpenCV's Jacobi eigen-solver for a symmetric 9x9 (core/src/lapack.cpp JacobiImpl_<double>)
// with OpenCV's own hypot.  A is destroyed; W gets eigenvalues (descending), V rows are the
// eigenvectors.  Used by find_homography4 below.
// ----------------------------------------------------------------------------------------
PANO_HD double cv_hypot(double a, double b) {
  a = fabs(a);
  b = fabs(b);
  if (a > b) {
    b = PANO_DDIV(b, a);
    return PANO_DMUL(a, PANO_DSQRT(PANO_DADD(1.0, PANO_DMUL(b, b))));
  }
  if (b > 0) {
    a = PANO_DDIV(a, b);
    return PANO_DMUL(b, PANO_DSQRT(PANO_DADD(1.0, PANO_DMUL(a, a))));
  }
  return 0;
}

#define PANO_ROT(v0, v1)                                         \
  {                                                              \
    double a0_ = (v0), b0_ = (v1);                               \
    (v0) = PANO_DSUB(PANO_DMUL(a0_, c), PANO_DMUL(b0_, s));      \
    (v1) = PANO_DADD(PANO_DMUL(a0_, s), PANO_DMUL(b0_, c));      \
  }

PANO_HD void jacobi9(double* A, double* W, double* V) {
  const int n = 9;
  const double eps = DBL_EPSILON;
  int i, j, k, m;
  for (i = 0; i < n; i++) {
    for (j = 0; j < n; j++) V[i * n + j] = 0;
    V[i * n + i] = 1;
  }
  int iters, maxIters = n * n * 30;
  int indR[9], indC[9];
  double mv = 0;
  for (k = 0; k < n; k++) {
    W[k] = A[(n + 1) * k];
    if (k < n - 1) {
      for (m = k + 1, mv = fabs(A[n * k + m]), i = k + 2; i < n; i++) {
        double val = fabs(A[n * k + i]);
        if (mv < val) mv = val, m = i;
      }
      indR[k] = m;
    }
    if (k > 0) {
      for (m = 0, mv = fabs(A[k]), i = 1; i < k; i++) {
        double val = fabs(A[n * i + k]);
        if (mv < val) mv = val, m = i;
      }
      indC[k] = m;
    }
  }
  for (iters = 0; iters < maxIters; iters++) {
    for (k = 0, mv = fabs(A[indR[0]]), i = 1; i < n - 1; i++) {
      double val = fabs(A[n * i + indR[i]]);
      if (mv < val) mv = val, k = i;
    }
    int l = indR[k];
    for (i = 1; i < n; i++) {
      double val = fabs(A[n * indC[i] + i]);
      if (mv < val) mv = val, k = indC[i], l = i;
    }
    double p = A[n * k + l];
    if (fabs(p) <= eps) break;
    double y = PANO_DMUL(PANO_DSUB(W[l], W[k]), 0.5);
    double t = PANO_DADD(fabs(y), cv_hypot(p, y));
    double s = cv_hypot(p, t);
    double c = PANO_DDIV(t, s);
    s = PANO_DDIV(p, s);
    t = PANO_DMUL(PANO_DDIV(p, t), p);
    if (y < 0) s = -s, t = -t;
    A[n * k + l] = 0;
    W[k] = PANO_DSUB(W[k], t);
    W[l] = PANO_DADD(W[l], t);
    for (i = 0; i < k; i++) PANO_ROT(A[n * i + k], A[n * i + l]);
    for (i = k + 1; i < l; i++) PANO_ROT(A[n * k + i], A[n * i + l]);
    for (i = l + 1; i < n; i++) PANO_ROT(A[n * k + i], A[n * l + i]);
    for (i = 0; i < n; i++) PANO_ROT(V[n * k + i], V[n * l + i]);
    for (j = 0; j < 2; j++) {
      int idx = j == 0 ? k : l;
      if (idx < n - 1) {
        for (m = idx + 1, mv = fabs(A[n * idx + m]), i = idx + 2; i < n; i++) {
          double val = fabs(A[n * idx + i]);
          if (mv < val) mv = val, m = i;
        }
        indR[idx] = m;
      }
      if (idx > 0) {
        for (m = 0, mv = fabs(A[idx]), i = 1; i < idx; i++) {
          double val = fabs(A[n * i + idx]);
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
      double tw = W[m]; W[m] = W[k]; W[k] = tw;
      for (i = 0; i < n; i++) {
        double tv = V[n * m + i]; V[n * m + i] = V[n * k + i]; V[n * k + i] = tv;
      }
    }
  }
}

// OpenCV gemm small-matrix path for 3x3 * 3x3 (no FMA, left to right).
PANO_HD void mul33(const double* a, const double* b, double* d) {
  double r[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      r[i * 3 + j] = PANO_DADD(PANO_DADD(PANO_DMUL(a[i * 3], b[j]), PANO_DMUL(a[i * 3 + 1], b[3 + j])),
                               PANO_DMUL(a[i * 3 + 2], b[6 + j]));
  for (int i = 0; i < 9; i++) d[i] = r[i];
}

// cv::findHomography(src, dst) for exactly 4 points (method 0): one runKernel call of
// HomographyEstimatorCallback (OpenCV calib3d/src/fundam.cpp).  M = src, m = dst, (x, y)
// interleaved floats.  Returns 0 for the "empty Mat" case.  ref: src/serial/main.cpp:279
PANO_HD int find_homography4(const float* M, const float* m, double* Hout, double* LtL /*81*/,
                             double* V /*81*/) {
  const int count = 4;
  double W[9];
  double cMx = 0, cMy = 0, cmx = 0, cmy = 0, sMx = 0, sMy = 0, smx = 0, smy = 0;
  for (int i = 0; i < count; i++) {
    cmx = PANO_DADD(cmx, (double)m[2 * i]);
    cmy = PANO_DADD(cmy, (double)m[2 * i + 1]);
    cMx = PANO_DADD(cMx, (double)M[2 * i]);
    cMy = PANO_DADD(cMy, (double)M[2 * i + 1]);
  }
  cmx = PANO_DDIV(cmx, (double)count); cmy = PANO_DDIV(cmy, (double)count);
  cMx = PANO_DDIV(cMx, (double)count); cMy = PANO_DDIV(cMy, (double)count);
  for (int i = 0; i < count; i++) {
    smx = PANO_DADD(smx, fabs(PANO_DSUB((double)m[2 * i], cmx)));
    smy = PANO_DADD(smy, fabs(PANO_DSUB((double)m[2 * i + 1], cmy)));
    sMx = PANO_DADD(sMx, fabs(PANO_DSUB((double)M[2 * i], cMx)));
    sMy = PANO_DADD(sMy, fabs(PANO_DSUB((double)M[2 * i + 1], cMy)));
  }
  if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON ||
      fabs(sMy) < DBL_EPSILON)
    return 0;
  smx = PANO_DDIV((double)count, smx); smy = PANO_DDIV((double)count, smy);
  sMx = PANO_DDIV((double)count, sMx); sMy = PANO_DDIV((double)count, sMy);
  double invHnorm[9] = {PANO_DDIV(1., smx), 0, cmx, 0, PANO_DDIV(1., smy), cmy, 0, 0, 1};
  double Hnorm2[9] = {sMx, 0, PANO_DMUL(-cMx, sMx), 0, sMy, PANO_DMUL(-cMy, sMy), 0, 0, 1};
  for (int i = 0; i < 81; i++) LtL[i] = 0;
  for (int i = 0; i < count; i++) {
    double x = PANO_DMUL(PANO_DSUB((double)m[2 * i], cmx), smx);
    double y = PANO_DMUL(PANO_DSUB((double)m[2 * i + 1], cmy), smy);
    double X = PANO_DMUL(PANO_DSUB((double)M[2 * i], cMx), sMx);
    double Y = PANO_DMUL(PANO_DSUB((double)M[2 * i + 1], cMy), sMy);
    double Lx[9] = {X, Y, 1, 0, 0, 0, PANO_DMUL(-x, X), PANO_DMUL(-x, Y), -x};
    double Ly[9] = {0, 0, 0, X, Y, 1, PANO_DMUL(-y, X), PANO_DMUL(-y, Y), -y};
    for (int j = 0; j < 9; j++)
      for (int k = j; k < 9; k++)
        LtL[j * 9 + k] = PANO_DADD(LtL[j * 9 + k], PANO_DADD(PANO_DMUL(Lx[j], Lx[k]), PANO_DMUL(Ly[j], Ly[k])));
  }
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < i; j++) LtL[i * 9 + j] = LtL[j * 9 + i];
  jacobi9(LtL, W, V);
  double Htemp[9], H0[9];
  mul33(invHnorm, V + 72, Htemp);
  mul33(Htemp, Hnorm2, H0);
  double sc = PANO_DDIV(1., H0[8]);
  for (int i = 0; i < 9; i++) Hout[i] = PANO_DMUL(H0[i], sc);
  return 1;
}

// Inlier predicate.  ref: src/serial/main.cpp:285-293 (gemm small path, Mat /= w as a
// multiply by 1./w, double->float casts, float subtraction, cv::norm(Point2f) in double).
PANO_HD bool is_inlier(const double* H, float x, float y, float qx, float qy, double thr) {
  double xd = (double)x, yd = (double)y;
  double X = PANO_DADD(PANO_DADD(PANO_DMUL(H[0], xd), PANO_DMUL(H[1], yd)), H[2]);
  double Y = PANO_DADD(PANO_DADD(PANO_DMUL(H[3], xd), PANO_DMUL(H[4], yd)), H[5]);
  double Wd = PANO_DADD(PANO_DADD(PANO_DMUL(H[6], xd), PANO_DMUL(H[7], yd)), H[8]);
  double s = PANO_DDIV(1., Wd);
  float ex = PANO_D2F(PANO_DMUL(X, s));
  float ey = PANO_D2F(PANO_DMUL(Y, s));
  float dx = PANO_FSUB(ex, qx), dy = PANO_FSUB(ey, qy);
  double d2 = PANO_DADD(PANO_DMUL((double)dx, (double)dx), PANO_DMUL((double)dy, (double)dy));
  return PANO_DSQRT(d2) < thr;
}

// The same predicate without the square root: sqrt is monotone and correctly rounded, so  sqrt(d2) < thr  holds
// exactly for  d2 < lim  where lim is the smallest double whose rounded square root reaches thr (inlier_d2_limit,
// computed once on the host by bisection over the doubles).  NaN / infinite d2 fail both forms.
PANO_HD bool is_inlier_lim(const double* H, float x, float y, float qx, float qy, double lim) {
  double xd = (double)x, yd = (double)y;
  double X = PANO_DADD(PANO_DADD(PANO_DMUL(H[0], xd), PANO_DMUL(H[1], yd)), H[2]);
  double Y = PANO_DADD(PANO_DADD(PANO_DMUL(H[3], xd), PANO_DMUL(H[4], yd)), H[5]);
  double Wd = PANO_DADD(PANO_DADD(PANO_DMUL(H[6], xd), PANO_DMUL(H[7], yd)), H[8]);
  double s = PANO_DDIV(1., Wd);
  float ex = PANO_D2F(PANO_DMUL(X, s));
  float ey = PANO_D2F(PANO_DMUL(Y, s));
  float dx = PANO_FSUB(ex, qx), dy = PANO_FSUB(ey, qy);
  double d2 = PANO_DADD(PANO_DMUL((double)dx, (double)dx), PANO_DMUL((double)dy, (double)dy));
  return d2 < lim;
}
// (host only) smallest non-negative double x with sqrt(x) >= thr (so sqrt(d2) < thr  <=>  d2 < x); 0 if thr <= 0, NaN if thr is NaN
inline double inlier_d2_limit(double thr) {
  if (thr != thr) return thr;
  if (!(thr > 0.0)) return 0.0;
  if (std::isinf(thr)) return thr;               // every finite d2 qualifies, infinity does not
  uint64_t lo = 0, hi = 0x7ff0000000000000ull;   // bit patterns of non-negative doubles are ordered like the values
  while (lo < hi) {                              // invariant: sqrt(hi) >= thr
    const uint64_t mid = lo + (hi - lo) / 2;
    double x;
    memcpy(&x, &mid, sizeof x);
    if (std::sqrt(x) >= thr) hi = mid; else lo = mid + 1;
  }
  double x;
  memcpy(&x, &lo, sizeof x);
  return x;
}

// cv::perspectiveTransform, Point2f with a 3x3 double matrix.  ref: src/serial/main.cpp:342
PANO_HD void persp_point(const double* m, float x, float y, float* ox, float* oy) {
  double xd = (double)x, yd = (double)y;
  double w = PANO_DADD(PANO_DADD(PANO_DMUL(xd, m[6]), PANO_DMUL(yd, m[7])), m[8]);
  if (fabs(w) > (double)FLT_EPSILON) {
    w = PANO_DDIV(1., w);
    *ox = PANO_D2F(PANO_DMUL(PANO_DADD(PANO_DADD(PANO_DMUL(xd, m[0]), PANO_DMUL(yd, m[1])), m[2]), w));
    *oy = PANO_D2F(PANO_DMUL(PANO_DADD(PANO_DADD(PANO_DMUL(xd, m[3]), PANO_DMUL(yd, m[4])), m[5]), w));
  } else {
    *ox = 0.f;
    *oy = 0.f;
  }
}

struct CanvasGeom {
  int cw, ch, offx, offy;  // canvas size and the left image's origin inside it
  double TH[9];            // translation * H
  double Minv[9];          // inverse of TH (what warpPerspective iterates with)
  int bw0;                 // warpPerspective's block width for this canvas (coordinate rounding)
  int ok;
};

// cv::invert 3x3 (closed form).  Returns 0 when det == 0 (OpenCV then leaves zeros).
PANO_HD int invert33(const double* s, double* d) {
  double c0 = PANO_DSUB(PANO_DMUL(s[4], s[8]), PANO_DMUL(s[5], s[7]));
  double c1 = PANO_DSUB(PANO_DMUL(s[3], s[8]), PANO_DMUL(s[5], s[6]));
  double c2 = PANO_DSUB(PANO_DMUL(s[3], s[7]), PANO_DMUL(s[4], s[6]));
  double det = PANO_DADD(PANO_DSUB(PANO_DMUL(s[0], c0), PANO_DMUL(s[1], c1)), PANO_DMUL(s[2], c2));
  if (det == 0.) {
    for (int i = 0; i < 9; i++) d[i] = 0;
    return 0;
  }
  det = PANO_DDIV(1., det);
  double t[9];
  t[0] = PANO_DMUL(c0, det);
  t[1] = PANO_DMUL(PANO_DSUB(PANO_DMUL(s[2], s[7]), PANO_DMUL(s[1], s[8])), det);
  t[2] = PANO_DMUL(PANO_DSUB(PANO_DMUL(s[1], s[5]), PANO_DMUL(s[2], s[4])), det);
  t[3] = PANO_DMUL(PANO_DSUB(PANO_DMUL(s[5], s[6]), PANO_DMUL(s[3], s[8])), det);
  t[4] = PANO_DMUL(PANO_DSUB(PANO_DMUL(s[0], s[8]), PANO_DMUL(s[2], s[6])), det);
  t[5] = PANO_DMUL(PANO_DSUB(PANO_DMUL(s[2], s[3]), PANO_DMUL(s[0], s[5])), det);
  t[6] = PANO_DMUL(c2, det);
  t[7] = PANO_DMUL(PANO_DSUB(PANO_DMUL(s[1], s[6]), PANO_DMUL(s[0], s[7])), det);
  t[8] = PANO_DMUL(PANO_DSUB(PANO_DMUL(s[0], s[4]), PANO_DMUL(s[1], s[3])), det);
  for (int i = 0; i < 9; i++) d[i] = t[i];
  return 1;
}

// Canvas geometry of stitchTwoImages.  ref: src/serial/main.cpp:335-369 and :376 (ROI).
PANO_HD void canvas_geometry(int wl, int hl, int wr, int hr, const double* H, CanvasGeom* g) {
  float cx[4] = {0.f, (float)wr, (float)wr, 0.f};
  float cy[4] = {0.f, 0.f, (float)hr, (float)hr};
  float minX = 0, minY = 0, maxX = (float)wl, maxY = (float)hl;
  for (int i = 0; i < 4; i++) {
    float px, py;
    persp_point(H, cx[i], cy[i], &px, &py);
    minX = fminf(minX, px); minY = fminf(minY, py);
    maxX = fmaxf(maxX, px); maxY = fmaxf(maxY, py);
  }
  // (the left corners (0,0),(wl,0),(wl,hl),(0,hl) are already covered by the initial values)
  double T[9] = {1, 0, (double)(-minX), 0, 1, (double)(-minY), 0, 0, 1};
  mul33(T, H, g->TH);
  float fw = PANO_FSUB(maxX, minX), fh = PANO_FSUB(maxY, minY);
  g->cw = (int)ceilf(fw);
  g->ch = (int)ceilf(fh);
  g->offx = (int)(-minX);
  g->offy = (int)(-minY);
  g->ok = 1;
  if (!(fw == fw) || !(fh == fh) || g->cw <= 0 || g->ch <= 0) g->ok = 0;
  else if (g->offx < 0 || g->offy < 0 || g->offx + wl > g->cw || g->offy + hl > g->ch) g->ok = 0;
  invert33(g->TH, g->Minv);
  // block shape of OpenCV's WarpPerspectiveInvoker (BLOCK_SZ = 32)
  int bh0 = g->ch < 16 ? g->ch : 16;
  if (bh0 < 1) bh0 = 1;
  int bw0 = 1024 / bh0;
  if (bw0 > g->cw) bw0 = g->cw;
  if (bw0 < 1) bw0 = 1;
  g->bw0 = bw0;
}

// Destination pixel -> source coordinate of cv::warpPerspective (INTER_LINEAR), in 1/32-px fixed
// point, exactly as OpenCV's WarpPerspectiveInvoker evaluates it: the three numerators are
// computed once per row at the block origin xb (= x - x % bw0) and advanced by x1 = x - xb.
struct WarpRow {
  double X0, Y0, W0;
};
PANO_HD WarpRow warp_row_origin(const double* M, int xb, int y) {
  const double xbd = (double)xb, yd = (double)y;
  WarpRow r;
  r.X0 = PANO_DADD(PANO_DADD(PANO_DMUL(M[0], xbd), PANO_DMUL(M[1], yd)), M[2]);
  r.Y0 = PANO_DADD(PANO_DADD(PANO_DMUL(M[3], xbd), PANO_DMUL(M[4], yd)), M[5]);
  r.W0 = PANO_DADD(PANO_DADD(PANO_DMUL(M[6], xbd), PANO_DMUL(M[7], yd)), M[8]);
  return r;
}
PANO_HD void warp_coord_from(const double* M, const WarpRow& r, int x1, int* Xo, int* Yo) {
  const double x1d = (double)x1;
  double W = PANO_DADD(r.W0, PANO_DMUL(M[6], x1d));
  W = (W != 0.) ? PANO_DDIV(32., W) : 0.;
  double fX = PANO_DMUL(PANO_DADD(r.X0, PANO_DMUL(M[0], x1d)), W);
  double fY = PANO_DMUL(PANO_DADD(r.Y0, PANO_DMUL(M[3], x1d)), W);
  fX = fmax(-2147483648.0, fmin(2147483647.0, fX));
  fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
  *Xo = PANO_RINT_I(fX);
  *Yo = PANO_RINT_I(fY);
}
PANO_HD void warp_coord(const double* M, int x, int y, int bw0, int* Xo, int* Yo) {
  const int xb = (x / bw0) * bw0;
  const WarpRow r = warp_row_origin(M, xb, y);
  warp_coord_from(M, r, x - xb, Xo, Yo);
}

// ----------------------------------------------------------------------------------------
// Fast evaluation of warp_coord_from for the warp kernel's fast path (warp.cu).  X0s, Y0s are the row-origin
// numerators scaled by 32 (a power of two commutes with every rounding), M0s = 32 M[0], M3s = 32 M[3], M6 = M[6].
//   OpenCV:  W = 32 / (W0 + M6 x1);  X = rint((X0 + M0 x1) * W)
// Here 1 / Wd comes from two Newton steps on a seed r0 (the device's MUFU.RCP64H; ANY seed is safe: the second
// step's residual e1 bounds the relative error of r2 by e1^2 + 2^-52), the product is rounded at 2^-20 with a magic
// constant (mantissa of v + 1.5 * 2^32 = 2^51 + round(v * 2^20)), and the result is accepted only if its fraction
// is at least 16 * 2^-20 away from one half and |e1| < 2^-22 (then |approx - exact| < 2^-22 * 2^-43.9 * ... is far
// below the distance to the rounding boundary for |v| < 2^22, which the caller guarantees).  Otherwise *exact_needed
// is set and the caller evaluates the exact expression (warp_coord).  Returns X, Y in 1/32 px.
// ----------------------------------------------------------------------------------------
PANO_HD void pano_d2words(double v, uint32_t* lo, uint32_t* hi) {
#if defined(__CUDA_ARCH__)
  *lo = (uint32_t)__double2loint(v);
  *hi = (uint32_t)__double2hiint(v);
#else
  uint64_t b;
  memcpy(&b, &v, sizeof b);
  *lo = (uint32_t)b;
  *hi = (uint32_t)(b >> 32);
#endif
}
PANO_HD void warp_coord_fast(double X0s, double Y0s, double W0, double M0s, double M3s, double M6, double x1d, double r0,
                             int* Xo, int* Yo, bool* exact_needed) {
  // (fused multiply-adds throughout: this is the APPROXIMATE evaluation, one ulp here or there is far inside the
  // tolerance that the acceptance test below enforces; the exact path keeps OpenCV's separately rounded operations)
  const double Wd = PANO_DFMA(M6, x1d, W0);
  const double Xn = PANO_DFMA(M0s, x1d, X0s);
  const double Yn = PANO_DFMA(M3s, x1d, Y0s);
  const double e0 = PANO_DFMA(-Wd, r0, 1.0);
  const double r1 = PANO_DFMA(r0, e0, r0);
  const double e1 = PANO_DFMA(-Wd, r1, 1.0);
  const double r2 = PANO_DFMA(r1, e1, r1);
  const double mx = PANO_DFMA(Xn, r2, 6442450944.0);
  const double my = PANO_DFMA(Yn, r2, 6442450944.0);
  uint32_t xl, xh, yl, yh;
  pano_d2words(mx, &xl, &xh);
  pano_d2words(my, &yl, &yh);
  // fraction within 16 * 2^-20 of one half?  ((frac + 2^19 + 16) mod 2^20 <= 32)
  const uint32_t tx = (xl + 0x80010u) & 0xFFFFFu, ty = (yl + 0x80010u) & 0xFFFFFu;
  *exact_needed = (tx < ty ? tx : ty) <= 32u || !(fabs(e1) < 2.384185791015625e-07);   // |e1| < 2^-22
  const uint32_t xl2 = xl + 0x80000u, yl2 = yl + 0x80000u;
  const uint32_t xh2 = xh + (xl2 < xl ? 1u : 0u), yh2 = yh + (yl2 < yl ? 1u : 0u);
  *Xo = (int)(((xl2 >> 20) | (xh2 << 12)) ^ 0x80000000u);
  *Yo = (int)(((yl2 >> 20) | (yh2 << 12)) ^ 0x80000000u);
}

// Second formulation of the same check, all in the FP64 pipe (the quad warp kernel: integer issue slots are what
// bounds it).  mx = fma(Xn, r2, 1.5 * 2^52) holds rint(Xn * r2) in its low mantissa word (round to nearest even of the
// exact product, |v| < 2^31); rnd = mx - 1.5 * 2^52 is that integer as a double (exact), d = fma(Xn, r2, -rnd) the
// signed distance of the approximate value from it (exact up to one rounding of a number below 1).  The result is
// accepted only if |d| <= 0.5 - 2^-16, i.e. the approximate value is at least 2^-16 away from the rounding boundary:
// the same margin as above, against an approximation error below 2^22 * (2^-44 + 3 * 2^-53) < 2^-21.
PANO_HD void warp_coord_fast2(double X0s, double Y0s, double W0, double M0s, double M3s, double M6, double x1d, double r0,
                              int* Xo, int* Yo, bool* exact_needed) {
  const double Wd = PANO_DFMA(M6, x1d, W0);
  const double Xn = PANO_DFMA(M0s, x1d, X0s);
  const double Yn = PANO_DFMA(M3s, x1d, Y0s);
  const double e0 = PANO_DFMA(-Wd, r0, 1.0);
  const double r1 = PANO_DFMA(r0, e0, r0);
  const double e1 = PANO_DFMA(-Wd, r1, 1.0);
  const double r2 = PANO_DFMA(r1, e1, r1);
  const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52
  const double mx = PANO_DFMA(Xn, r2, MAGIC), my = PANO_DFMA(Yn, r2, MAGIC);
  const double dx = PANO_DFMA(Xn, r2, -PANO_DSUB(mx, MAGIC)), dy = PANO_DFMA(Yn, r2, -PANO_DSUB(my, MAGIC));
  const double lim = 0.4999847412109375;     // 0.5 - 2^-16
  // (written so that a NaN anywhere - e.g. a zero denominator - fails the test and takes the exact path)
  *exact_needed = !(fabs(dx) <= lim && fabs(dy) <= lim && fabs(e1) < 2.384185791015625e-07);
  uint32_t lo, hi;
  pano_d2words(mx, &lo, &hi);
  *Xo = (int)lo;
  pano_d2words(my, &lo, &hi);
  *Yo = (int)lo;
}

PANO_HD int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

// One destination pixel of the fixed-point bilinear remap (OpenCV remapBilinear, 8UC3,
// BORDER_CONSTANT 0) for the 1/32-px source coordinate (X, Y): returns b | g<<8 | r<<16.
PANO_HD uint32_t warp_pixel(const uint8_t* src, size_t sstride, int ws, int hs,
                                               int X, int Y) {
  const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
  const int fx = X & 31, fy = Y & 31;
  // fully outside -> 0 (also the common case on the left part of the canvas)
  if (sx >= ws || sx + 1 < 0 || sy >= hs || sy + 1 < 0) return 0u;
  const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
  const int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  const bool x0 = sx >= 0, x1 = sx + 1 < ws, y0 = sy >= 0, y1 = sy + 1 < hs;
  const uint8_t* r0 = src + (size_t)(y0 ? sy : 0) * sstride;
  const uint8_t* r1 = src + (size_t)(y1 ? sy + 1 : 0) * sstride;
  const int cx0 = 3 * (x0 ? sx : 0), cx1 = 3 * (x1 ? sx + 1 : 0);
  const bool v00 = x0 && y0, v01 = x1 && y0, v10 = x0 && y1, v11 = x1 && y1;
  uint32_t out = 0;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    int p00 = v00 ? r0[cx0 + c] : 0, p01 = v01 ? r0[cx1 + c] : 0;
    int p10 = v10 ? r1[cx0 + c] : 0, p11 = v11 ? r1[cx1 + c] : 0;
    int v = p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11;
    out |= (uint32_t)((v + (1 << 14)) >> 15) << (8 * c);
  }
  return out;
}


// ----------------------------------------------------------------------------------------
// libstdc++ 13 std::shuffle + uniform_int_distribution (Lemire) on a 32-bit engine,
// restated per step.  ref: src/serial/main.cpp:270-271; /usr/include/c++/13/bits/stl_algo.h
// (shuffle, __gen_two_uniform_ints) and bits/uniform_int_dist.h (_S_nd).
//
// "Pair" regime (n <= 65535... precisely 0xFFFFFFFF / n >= n): after an optional single
// draw for even n, elements idx, idx+1 are swapped with positions drawn from ONE Lemire draw
// over range b0*b1, b0 = idx+1, b1 = idx+2:  pos_first = r / b1, pos_second = r % b1 where
// r = hi32(x * b0*b1).  With u = x*b0 (64-bit) and v = lo32(u)*b1 (64-bit) this is
//   pos_first = hi32(u), pos_second = hi32(v), and Lemire's low word is lo32(v).
// A draw is rejected iff lo32(v) < (2^32 mod (b0*b1)).
// "Single" regime (larger n): element idx swaps with hi32(x * (idx+1)), rejected iff
// lo32(x * (idx+1)) < (2^32 mod (idx+1)).
// ----------------------------------------------------------------------------------------
PANO_HD bool shuffle_uses_pairs(uint32_t n) { return n > 0 && (0xFFFFFFFFu / n) >= n; }

// number of Lemire draws (steps) one shuffle of n elements makes, excluding rejections
PANO_HD uint32_t shuffle_steps(uint32_t n) {
  if (n < 2) return 0;
  if (shuffle_uses_pairs(n)) return (n % 2 == 0) ? 1 + (n - 2) / 2 : (n - 1) / 2;
  return n - 1;
}

// 2^32 mod r for 1 <= r < 2^32
PANO_HD uint32_t lemire_threshold(uint32_t r) { return (uint32_t)(0u - r) % r; }

// ----------------------------------------------------------------------------------------
// Walks of one shuffle through the engine's output stream X (offsets relative to X).
// rt[k] = (range r_k, Lemire threshold T_k) of step k: a draw x is rejected iff
// lo32(x * r_k) < T_k (for paired steps r_k = b0*b1 < 2^32, and lo32(lo32(x*b0)*b1) is the same
// low word).  Step 0 of an even-sized paired shuffle is the d{0,1} draw: (2, 0), never rejects.
//
// Both walks process 8 steps at a time under the assumption that none of them rejects (and,
// for the tracking walk, that none touches positions 0..3); the 8 loads and tests are
// independent, which is what hides the memory latency.  On a flagged step they fall back to
// the exact one-step path from the first flagged step on.
// ----------------------------------------------------------------------------------------
struct alignas(8) RT {
  uint32_t r, T;
};

PANO_HD int first_bit8(uint32_t m) {
  int i = 0;
  while (!((m >> i) & 1u)) i++;
  return i;
}

// Steps per pass-2 segment.  Pass 1 records each candidate's offset at every segment
// boundary so that pass 2 can walk all segments of all iterations in parallel.
#define PANO_SEG_STEPS 256u

PANO_HD uint32_t ctz32(uint32_t w) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)(__ffs((int)w) - 1);
#else
  return (uint32_t)__builtin_ctz(w);
#endif
}

// Pass 1, grid formulation.  For one iteration, cell (d, k) says whether the draw at stream
// position pos0 + d + k is rejected when used for step k; d ("diagonal") = candidate start index
// + rejections so far.  The cells are evaluated independently (replay_cells_kernel) and packed
// 32 steps per word: word(kb, d) = bits[kb * D + d].  A walk then only scans words: it moves
// along its diagonal until the next set bit (a rejection at step k), hops to diagonal d + 1 and
// re-examines the same step there.  Returns the end offset pos0 + d_final + steps (relative to
// the chunk base) or 0xffffffff if the walk leaves the D diagonals that were evaluated.
// Segment-boundary offsets are recorded as in walk_offsets.
PANO_HD uint32_t walk_bits(const uint32_t* bits, uint32_t D, uint32_t nkb, uint32_t d0, uint32_t steps,
                           uint32_t pos0, uint32_t* seg_off, size_t seg_stride) {
  // Every round fetches the next AHEAD words of the current diagonal AND of the one above it (independent loads,
  // adjacent addresses): the first rejection of a round continues on the second diagonal without another L2 round
  // trip; only the second rejection (or the end of the batch) starts a new round.  A hop discards what is left of
  // a batch, so a deeper look-ahead only adds L2 traffic.
  constexpr uint32_t AHEAD = 8;
  constexpr uint32_t SEGW = PANO_SEG_STEPS / 32u;      // words per pass-2 segment
  uint32_t d = d0, kb = 0, mask = ~0u;
  while (kb < nkb) {
    uint32_t w[AHEAD], v[AHEAD];
    const uint32_t* p = bits + (size_t)kb * D + d;
    const uint32_t left = nkb - kb;
    const bool up = d + 1u < D;                        // the diagonal above exists in the evaluated grid
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t i = 0; i < AHEAD; i++) {
      w[i] = i < left ? p[(size_t)i * D] : 0u;
      v[i] = (i < left && up) ? p[(size_t)i * D + 1u] : 0u;
    }
    w[0] &= mask;
    mask = ~0u;
    const uint32_t lim = left < AHEAD ? left : AHEAD;
    uint32_t adv = lim, hit = 0;                       // words without a rejection, first word with one
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = (int)AHEAD - 1; i >= 0; i--)
      if (w[i]) { adv = (uint32_t)i; hit = w[i]; }
    // segment boundaries (multiples of SEGW words) in (kb, kb + adv]: at most AHEAD / SEGW of them
    if (seg_off) {
      for (uint32_t nb = (kb / SEGW + 1u) * SEGW; nb <= kb + adv && nb < nkb; nb += SEGW)
        seg_off[(size_t)(nb / SEGW - 1u) * seg_stride] = pos0 + d + nb * 32u;
    }
    kb += adv;
    if (!hit) continue;
    d++;
    if (d >= D) return 0xffffffffu;
    // same step again on the diagonal above, out of the words already fetched: indices adv .. lim - 1
    const uint32_t m2 = ~0u << ctz32(hit);
    uint32_t adv2 = lim, hit2 = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = (int)AHEAD - 1; i >= 0; i--) {
      const uint32_t x = (uint32_t)i == adv ? (v[i] & m2) : v[i];
      if ((uint32_t)i >= adv && x) { adv2 = (uint32_t)i; hit2 = x; }
    }
    if (seg_off) {
      for (uint32_t nb = (kb / SEGW + 1u) * SEGW; nb <= kb + (adv2 - adv) && nb < nkb; nb += SEGW)
        seg_off[(size_t)(nb / SEGW - 1u) * seg_stride] = pos0 + d + nb * 32u;
    }
    kb += adv2 - adv;
    if (!hit2) continue;                               // the rest of the batch is clean on this diagonal
    d++;
    if (d >= D) return 0xffffffffu;
    mask = ~0u << ctz32(hit2);
  }
  return pos0 + d + steps;
}

// pass 1: end offset (and, if seg_off != nullptr, the offset at the start of segments 1, 2, ...
// written to seg_off[(s - 1) * seg_stride])
PANO_HD uint32_t walk_offsets(const uint32_t* X, uint32_t o, uint32_t steps, const RT* rt, uint32_t* seg_off,
                              size_t seg_stride) {
  const uint32_t* px = X + o;  // advancing pointers: loads use immediate offsets
  uint32_t k0 = 0;
  while (k0 < steps) {
    const uint32_t k1 = (steps - k0 > PANO_SEG_STEPS) ? k0 + PANO_SEG_STEPS : steps;
    const RT* pr = rt + k0;
    const RT* const pr_end = rt + k1;
    const RT* const pr_end8 = (k1 - k0 >= 8) ? pr_end - 7 : pr;  // pr < pr_end8: 8 steps remain
    while (pr < pr_end8) {
      uint32_t rej = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int i = 0; i < 8; i++) {
        const RT q = pr[i];
        rej |= (px[i] * q.r < q.T) ? (1u << i) : 0u;
      }
      if (rej == 0) { px += 8; pr += 8; continue; }
      const int i0 = first_bit8(rej);
      px += i0; pr += i0;
      const RT q = *pr;
      uint32_t x0 = *px++;
      while (x0 * q.r < q.T) x0 = *px++;
      pr++;
    }
    for (; pr < pr_end; pr++) {
      const RT q = *pr;
      uint32_t x0 = *px++;
      while (x0 * q.r < q.T) x0 = *px++;
    }
    k0 = k1;
    if (seg_off && k0 < steps) seg_off[(size_t)(k0 / PANO_SEG_STEPS - 1) * seg_stride] = (uint32_t)(px - X);
  }
  return (uint32_t)(px - X);
}

PANO_HD void track_swap(int (&a)[4], uint32_t idx, uint32_t p) {
  if (p < 4u) {
    if (idx < 4u) {
      int t = a[idx];
      a[idx] = a[p];
      a[p] = t;
    } else {
      a[p] = (int)idx;
    }
  }
}

// exact single step of the tracking walk (consumes draws until one is accepted)
template <bool PAIRS>
PANO_HD void track_step(const uint32_t* X, uint32_t& o, uint32_t k, uint32_t n, const RT* rt, int (&a)[4]) {
  const uint32_t r = rt[k].r, T = rt[k].T;
  uint32_t x = X[o++];
  while (x * r < T) x = X[o++];
  if (PAIRS) {
    const uint32_t odd = n & 1u;
    if (!odd && k == 0) {  // d{0,1}: element 1 swaps with position x >> 31
      track_swap(a, 1u, x >> 31);
      return;
    }
    const uint32_t idx = 2u * k + odd, b0 = idx + 1u, b1 = idx + 2u;
    const unsigned long long u = (unsigned long long)x * b0;
    const unsigned long long v = (unsigned long long)(uint32_t)u * b1;
    track_swap(a, idx, (uint32_t)(u >> 32));
    track_swap(a, idx + 1u, (uint32_t)(v >> 32));
  } else {
    const uint32_t idx = k + 1u;
    track_swap(a, idx, (uint32_t)(((unsigned long long)x * r) >> 32));
  }
}

// pass 2, one segment [k0, k1) of one iteration starting at offset o: returns the end offset.
// w[p] receives the element that the segment leaves in position p (p = 0..3), or -1 if the
// segment never writes p.  Segment 0 (k0 == 0) also carries the initial identity and the swaps
// among the first four elements, so its w[] is always complete.  Later segments only ever
// assign "position p <- element idx" (idx >= 4), so segments compose by last-writer-wins.
template <bool PAIRS>
PANO_HD uint32_t walk_track_segment(const uint32_t* X, uint32_t o, uint32_t n, uint32_t k0, uint32_t k1,
                                    const RT* rt, int (&w)[4]) {
  int a[4];
  uint32_t k = k0;
  const uint32_t odd = n & 1u;
  if (k0 == 0) {
    a[0] = 0; a[1] = 1; a[2] = 2; a[3] = 3;
    // exact steps until every swap partner index is >= 4 (idx < 4 swaps exchange tracked slots)
    while (k < k1 && (PAIRS ? 2u * k + odd : k + 1u) < 4u) { track_step<PAIRS>(X, o, k, n, rt, a); k++; }
  } else {
    a[0] = a[1] = a[2] = a[3] = -1;
  }
  while (k + 8 <= k1) {
    const uint32_t* px = X + o;
    const RT* pr = rt + k;
    uint32_t flag = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 8; i++) {
      const uint32_t x = px[i];
      const RT q = pr[i];
      bool f = x * q.r < q.T;
      if (PAIRS) {
        const uint32_t idx = 2u * (k + i) + odd, b0 = idx + 1u, b1 = idx + 2u;
        const unsigned long long u = (unsigned long long)x * b0;
        const unsigned long long v = (unsigned long long)(uint32_t)u * b1;
        const uint32_t p1 = (uint32_t)(u >> 32), p2 = (uint32_t)(v >> 32);
        f = f || ((p1 < p2 ? p1 : p2) < 4u);
      } else {
        f = f || ((uint32_t)(((unsigned long long)x * q.r) >> 32) < 4u);
      }
      flag |= f ? (1u << i) : 0u;
    }
    if (flag == 0) { o += 8; k += 8; continue; }
    int i0 = first_bit8(flag);
    o += i0; k += i0;
    track_step<PAIRS>(X, o, k, n, rt, a);
    k++;
  }
  for (; k < k1; k++) track_step<PAIRS>(X, o, k, n, rt, a);
  w[0] = a[0]; w[1] = a[1]; w[2] = a[2]; w[3] = a[3];
  return o;
}

}  // namespace pano
