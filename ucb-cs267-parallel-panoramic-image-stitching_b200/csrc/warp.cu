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

#include "warp_kernels.cuh"

}  // namespace

void pack_rows_device(cudaStream_t st, const uint8_t* src, size_t pitch, size_t row_bytes, int rows, uint8_t* dst) {
  const unsigned long long total = (unsigned long long)row_bytes * (unsigned long long)rows;
  if (total == 0) return;
  const unsigned long long words = (total + 3) / 4;
  pack_rows_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(src, pitch, (uint32_t)row_bytes, total, dst);
  PANO_LAUNCH_CHECK();
}

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
  const bool box = footprint_box(g.TH, g.Minv, right.w, right.h, g.cw, g.ch, P);
  ProfScope ps(PROF_WARP, st);
  if (fast_path_ok(P, right.p, right.stride, box) && warp_fast_enabled()) {
    launch_fast<1>(st, left.p, left.stride, right.p, right.stride, P, canvas, canvas_stride);
  } else {
    dim3 block(32, 8), grid(((g.cw + 3) / 4 + 31) / 32, (g.ch + 7) / 8);
    warp_overlay_kernel<1><<<grid, block, 0, st>>>(left.p, left.stride, right.p, right.stride, P, canvas,
                                                     canvas_stride);
  }
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
  bool box = false;
  {
    double fwd[9];
    if (invert33(Minv, fwd)) box = footprint_box(fwd, Minv, src.w, src.h, dw, dh, P);
    else { P.bx0 = 0; P.by0 = 0; P.bx1 = dw - 1; P.by1 = dh - 1; }
  }
  if (fast_path_ok(P, src.p, src.stride, box) && warp_fast_enabled()) {
    launch_fast<0>(st, nullptr, 0, src.p, src.stride, P, dst, dstride);
  } else {
    dim3 block(32, 8), grid(((dw + 3) / 4 + 31) / 32, (dh + 7) / 8);
    warp_overlay_kernel<0><<<grid, block, 0, st>>>(nullptr, 0, src.p, src.stride, P, dst, dstride);
  }
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
  const bool box = footprint_box(M, Minv, src.w, src.h, canvas_w, canvas_h, P);   // box in whole-canvas coordinates
  if (fast_path_ok(P, src.p, src.stride, box) && warp_fast_enabled()) {
    launch_fast<2>(st, nullptr, 0, src.p, src.stride, P, band, band_stride);
  } else {
    dim3 block(32, 8), grid(((canvas_w + 3) / 4 + 31) / 32, (band_h + 7) / 8);
    warp_overlay_kernel<2><<<grid, block, 0, st>>>(nullptr, 0, src.p, src.stride, P, band, band_stride);
  }
  PANO_LAUNCH_CHECK();
}

}  // namespace pano
