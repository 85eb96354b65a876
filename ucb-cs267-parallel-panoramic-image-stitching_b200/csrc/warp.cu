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
};

// Each thread produces 4 horizontally adjacent canvas pixels (12 bytes = three 32-bit stores;
// the canvas pitch is a multiple of 4).  OVERLAY = false: plain warpPerspective.
template <bool OVERLAY>
__global__ void __launch_bounds__(256)
warp_overlay_kernel(const uint8_t* __restrict__ left, size_t lstride, const uint8_t* __restrict__ right,
                    size_t rstride, WarpParams P, uint8_t* __restrict__ canvas, size_t cstride) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x0 >= P.cw || y >= P.ch) return;
  uint32_t px[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int x = x0 + i;
    uint32_t v = 0;
    if (x < P.cw) {
      int X, Y;
      warp_coord(P.M, x, y, P.bw0, &X, &Y);
      v = warp_pixel(right, rstride, P.ws, P.hs, X, Y);
      if (OVERLAY && v == 0u) {
        const int lx = x - P.offx, ly = y - P.offy;
        if (lx >= 0 && lx < P.wl && ly >= 0 && ly < P.hl) {
          const uint8_t* p = left + (size_t)ly * lstride + 3 * (size_t)lx;
          v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
        }
      }
    }
    px[i] = v;
  }
  uint8_t* row = canvas + (size_t)y * cstride + 3 * (size_t)x0;
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

}  // namespace

void warp_overlay_device(cudaStream_t st, const DevImage& left, const DevImage& right, const CanvasGeom& g,
                         uint8_t* canvas, size_t canvas_stride) {
  WarpParams P;
  memcpy(P.M, g.Minv, sizeof P.M);
  P.bw0 = g.bw0;
  P.cw = g.cw; P.ch = g.ch;
  P.offx = g.offx; P.offy = g.offy; P.wl = left.w; P.hl = left.h;
  P.ws = right.w; P.hs = right.h;
  dim3 block(32, 8), grid(((g.cw + 3) / 4 + 31) / 32, (g.ch + 7) / 8);
  warp_overlay_kernel<true><<<grid, block, 0, st>>>(left.p, left.stride, right.p, right.stride, P, canvas,
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
  dim3 block(32, 8), grid(((dw + 3) / 4 + 31) / 32, (dh + 7) / 8);
  warp_overlay_kernel<false><<<grid, block, 0, st>>>(nullptr, 0, src.p, src.stride, P, dst, dstride);
  PANO_LAUNCH_CHECK();
}

}  // namespace pano
