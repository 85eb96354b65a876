// warp.cu — K8 inverse-homography bilinear warp + left copy + "non-black overwrites" overlay,
// one pass over the output canvas.
//
// Semantics: ref src/serial/main.cpp:371-386:
//   warpedRight = cv::warpPerspective(right, T*H, canvasSize)   (INTER_LINEAR, BORDER_CONSTANT 0)
//   canvas(Rect(-minX, -minY, wl, hl)) = left
//   canvas(y, x) = warpedRight(y, x) wherever warpedRight(y, x) != (0, 0, 0)
// The warp reproduces OpenCV's fixed-point path bit for bit (see pano_core.cuh warp_coord):
// 1/32-px coordinates computed per 64-px block origin, 15-bit bilinear weights
// (32-fx)(32-fy)*32 ..., result (sum + 2^14) >> 15, taps outside the source are 0.
// Roofline: HBM bound — 3 B/px of each source read once, 3 B/px of canvas written once.
#include "common.cuh"

namespace pano {

namespace {

struct WarpParams {
  double M[9];  // inverse of T*H
  int bw0;      // OpenCV block width used for coordinate evaluation
  int cw, ch;
  int offx, offy, wl, hl;  // left ROI
  int ws, hs;              // source (right) size
  int bx0, by0, bx1, by1;  // canvas pixels outside this box provably map outside the source
  size_t src_bytes;        // bytes of the source buffer that may be read with word loads
  int y0;                  // first canvas row of the band being rendered (0 for a whole canvas)
};

// 6 consecutive bytes starting at p (two adjacent BGR pixels) from aligned 32-bit loads
__device__ __forceinline__ void load6(const uint8_t* __restrict__ p, uint32_t& lo, uint32_t& hi) {
  const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
  const uint32_t* q = reinterpret_cast<const uint32_t*>(p - a);
  const uint32_t w0 = q[0], w1 = q[1];
  const uint32_t w2 = a >= 3u ? q[2] : 0u;
  lo = __funnelshift_r(w0, w1, 8u * a);   // bytes p[0..3]
  hi = __funnelshift_r(w1, w2, 8u * a);   // bytes p[4..7]
}

// fixed-point bilinear tap combination for one pixel; fast path when all four taps are inside
__device__ __forceinline__ uint32_t warp_pixel_fast(const uint8_t* __restrict__ src, size_t sstride, int ws, int hs,
                                                    size_t src_bytes, int X, int Y) {
  const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
  if (sx >= ws || sx + 1 < 0 || sy >= hs || sy + 1 < 0) return 0u;
  const size_t off = (size_t)sy * sstride + 3 * (size_t)sx;
  if (sx >= 0 && sx + 1 < ws && sy >= 0 && sy + 1 < hs && off + sstride + 12 <= src_bytes) {
    const int fx = X & 31, fy = Y & 31;
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
    const int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    uint32_t a_lo, a_hi, b_lo, b_hi;
    load6(src + off, a_lo, a_hi);
    load6(src + off + sstride, b_lo, b_hi);
    uint32_t out = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int p00 = (a_lo >> (8 * c)) & 255;
      const int p01 = c == 0 ? (a_lo >> 24) : ((a_hi >> (8 * (c - 1))) & 255);
      const int p10 = (b_lo >> (8 * c)) & 255;
      const int p11 = c == 0 ? (b_lo >> 24) : ((b_hi >> (8 * (c - 1))) & 255);
      const int v = p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11;
      out |= (uint32_t)((v + (1 << 14)) >> 15) << (8 * c);
    }
    return out;
  }
  return warp_pixel(src, sstride, ws, hs, X, Y);  // border taps: general path
}

// Each thread produces 4 horizontally adjacent canvas pixels (12 bytes = three 32-bit stores;
// the canvas pitch is a multiple of 4).  MODE 0: plain warpPerspective; 1: + left copy (pair);
// 2: accumulate — non-black warped pixels overwrite what the canvas band already holds.
template <int MODE>
__global__ void __launch_bounds__(256)
warp_overlay_kernel(const uint8_t* __restrict__ left, size_t lstride, const uint8_t* __restrict__ right,
                    size_t rstride, WarpParams P, uint8_t* __restrict__ canvas, size_t cstride) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int yb = blockIdx.y * blockDim.y + threadIdx.y;   // row inside the band
  const int y = yb + P.y0;                                // row of the whole canvas
  if (x0 >= P.cw || yb >= P.ch) return;
  uint32_t px[4] = {0u, 0u, 0u, 0u};
  if (!(x0 + 3 < P.bx0 || x0 > P.bx1 || y < P.by0 || y > P.by1)) {
    // the four pixels share OpenCV's row-origin numerators when they lie in one bw0-block
    const int xb = (x0 / P.bw0) * P.bw0;
    const bool same = (x0 + 3) / P.bw0 == x0 / P.bw0;
    const WarpRow r = warp_row_origin(P.M, xb, y);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int x = x0 + i;
      if (x < P.cw) {
        int X, Y;
        if (same) warp_coord_from(P.M, r, x - xb, &X, &Y);
        else warp_coord(P.M, x, y, P.bw0, &X, &Y);
        px[i] = warp_pixel_fast(right, rstride, P.ws, P.hs, P.src_bytes, X, Y);
      }
    }
  }
  if (MODE == 1) {
    const int ly = y - P.offy;
    if (ly >= 0 && ly < P.hl) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int lx = x0 + i - P.offx;
        if (px[i] == 0u && lx >= 0 && lx < P.wl) {
          const uint8_t* p = left + (size_t)ly * lstride + 3 * (size_t)lx;
          px[i] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
        }
      }
    }
  }
  uint8_t* row = canvas + (size_t)yb * cstride + 3 * (size_t)x0;
  if (MODE == 2) {
    if ((px[0] | px[1] | px[2] | px[3]) == 0u) return;
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (px[i] != 0u && x0 + i < P.cw) {
        row[3 * i] = (uint8_t)px[i];
        row[3 * i + 1] = (uint8_t)(px[i] >> 8);
        row[3 * i + 2] = (uint8_t)(px[i] >> 16);
      }
    return;
  }
  if (x0 + 3 < P.cw && (reinterpret_cast<uintptr_t>(row) & 3) == 0) {
    uint32_t* o = reinterpret_cast<uint32_t*>(row);  // 3*x0 is a multiple of 12
    o[0] = px[0] | (px[1] << 24);
    o[1] = (px[1] >> 8) | (px[2] << 16);
    o[2] = (px[2] >> 16) | (px[3] << 8);
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (x0 + i < P.cw) {
        row[3 * i] = (uint8_t)px[i];
        row[3 * i + 1] = (uint8_t)(px[i] >> 8);
        row[3 * i + 2] = (uint8_t)(px[i] >> 16);
      }
  }
}

// Canvas box outside which no pixel can receive a source tap: forward image (by `fwd`, the matrix
// whose inverse is iterated) of the source rectangle grown by 2 px, grown again by 2 px.  Only
// valid when the inverse map's denominator is positive on the whole canvas (then it is a proper
// projective bijection there and the image of the rectangle is the convex quad of its corners);
// otherwise the box is the whole canvas and every pixel is evaluated.
void footprint_box(const double* fwd, const double* Minv, int ws, int hs, int cw, int ch, WarpParams& P) {
  P.bx0 = 0; P.by0 = 0; P.bx1 = cw - 1; P.by1 = ch - 1;
  const double cx[4] = {0.0, (double)cw, (double)cw, 0.0}, cy[4] = {0.0, 0.0, (double)ch, (double)ch};
  for (int i = 0; i < 4; i++) {
    double wd = Minv[6] * cx[i] + Minv[7] * cy[i] + Minv[8];
    if (!(wd > 1e-12)) return;
  }
  const double sx[4] = {-2.0, ws + 1.0, ws + 1.0, -2.0}, sy[4] = {-2.0, -2.0, hs + 1.0, hs + 1.0};
  double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
  for (int i = 0; i < 4; i++) {
    double wd = fwd[6] * sx[i] + fwd[7] * sy[i] + fwd[8];
    if (!(wd > 1e-12)) return;
    double X = (fwd[0] * sx[i] + fwd[1] * sy[i] + fwd[2]) / wd, Y = (fwd[3] * sx[i] + fwd[4] * sy[i] + fwd[5]) / wd;
    x0 = fmin(x0, X); x1 = fmax(x1, X); y0 = fmin(y0, Y); y1 = fmax(y1, Y);
  }
  if (!(x0 == x0) || !(x1 == x1) || !(y0 == y0) || !(y1 == y1)) return;
  auto clampi = [](double v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : (int)v); };
  P.bx0 = clampi(floor(x0) - 2, 0, cw - 1);
  P.by0 = clampi(floor(y0) - 2, 0, ch - 1);
  P.bx1 = clampi(ceil(x1) + 2, 0, cw - 1);
  P.by1 = clampi(ceil(y1) + 2, 0, ch - 1);
}

}  // namespace

void warp_overlay_device(cudaStream_t st, const DevImage& left, const DevImage& right, const CanvasGeom& g,
                         uint8_t* canvas, size_t canvas_stride) {
  WarpParams P;
  memcpy(P.M, g.Minv, sizeof P.M);
  P.bw0 = g.bw0;
  P.cw = g.cw; P.ch = g.ch;
  P.offx = g.offx; P.offy = g.offy; P.wl = left.w; P.hl = left.h;
  P.ws = right.w; P.hs = right.h;
  P.y0 = 0;
  P.src_bytes = (size_t)(right.h - 1) * right.stride + (size_t)right.w * 3;
  footprint_box(g.TH, g.Minv, right.w, right.h, g.cw, g.ch, P);
  dim3 block(32, 8), grid(((g.cw + 3) / 4 + 31) / 32, (g.ch + 7) / 8);
  warp_overlay_kernel<1><<<grid, block, 0, st>>>(left.p, left.stride, right.p, right.stride, P, canvas,
                                                   canvas_stride);
  PANO_LAUNCH_CHECK();
}

void warp_only_device(cudaStream_t st, const DevImage& src, const double* Minv, int bw0, uint8_t* dst, int dw,
                      int dh, size_t dstride) {
  WarpParams P;
  memcpy(P.M, Minv, sizeof P.M);
  P.bw0 = bw0;
  P.cw = dw; P.ch = dh;
  P.offx = P.offy = 0; P.wl = P.hl = 0;
  P.ws = src.w; P.hs = src.h;
  P.y0 = 0;
  P.src_bytes = (size_t)(src.h - 1) * src.stride + (size_t)src.w * 3;
  {
    double fwd[9];
    if (invert33(Minv, fwd)) footprint_box(fwd, Minv, src.w, src.h, dw, dh, P);
    else { P.bx0 = 0; P.by0 = 0; P.bx1 = dw - 1; P.by1 = dh - 1; }
  }
  dim3 block(32, 8), grid(((dw + 3) / 4 + 31) / 32, (dh + 7) / 8);
  warp_overlay_kernel<0><<<grid, block, 0, st>>>(nullptr, 0, src.p, src.stride, P, dst, dstride);
  PANO_LAUNCH_CHECK();
}

void warp_accumulate_device(cudaStream_t st, const DevImage& src, const double* M, uint8_t* band, int canvas_w,
                            int canvas_h, int y0, int band_h, size_t band_stride) {
  WarpParams P;
  double Minv[9];
  invert33(M, Minv);
  memcpy(P.M, Minv, sizeof P.M);
  int bh0 = canvas_h < 16 ? canvas_h : 16;          // OpenCV's block shape depends on the WHOLE canvas
  int bw0 = 1024 / (bh0 < 1 ? 1 : bh0);
  if (bw0 > canvas_w) bw0 = canvas_w;
  P.bw0 = bw0 < 1 ? 1 : bw0;
  P.cw = canvas_w; P.ch = band_h;
  P.offx = P.offy = 0; P.wl = P.hl = 0;
  P.ws = src.w; P.hs = src.h;
  P.y0 = y0;
  P.src_bytes = (size_t)(src.h - 1) * src.stride + (size_t)src.w * 3;
  footprint_box(M, Minv, src.w, src.h, canvas_w, canvas_h, P);   // box in whole-canvas coordinates
  dim3 block(32, 8), grid(((canvas_w + 3) / 4 + 31) / 32, (band_h + 7) / 8);
  warp_overlay_kernel<2><<<grid, block, 0, st>>>(nullptr, 0, src.p, src.stride, P, band, band_stride);
  PANO_LAUNCH_CHECK();
}

}  // namespace pano
