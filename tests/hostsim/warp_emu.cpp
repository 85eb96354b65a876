// warp_emu.cpp — runs the device code of the warp stage (csrc/warp_kernels.cuh: the production quad kernel, the
// round-1 fast kernel, the general kernel, the row packer, plus the host-side footprint box / fast-path admission they
// depend on) on the CPU emulation of the CUDA execution model (cuda_emu.hpp), for the no-GPU test tier.
//
// TEST INFRASTRUCTURE ONLY.  The launch arithmetic mirrors warp.cu's launchers (warp_overlay_device, warp_only_device,
// warp_accumulate_device, launch_fast, pack_rows_device); the kernels are the product's source compiled unchanged.
#include "cuda_emu.hpp"

#include <cmath>
#include <cstdlib>

#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/pano_core.cuh"

namespace pano {
namespace {
#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/warp_kernels.cuh"
}  // namespace
}  // namespace pano

using namespace pano;

namespace {
struct DevImg {   // a "device" copy: 256-byte aligned base, 256-byte pitch (what the engine's to_device produces)
  uint8_t* p = nullptr;
  int w = 0, h = 0;
  size_t stride = 0;
  DevImg(const uint8_t* src, int w_, int h_, size_t sstride, size_t pitch_align = 256) : w(w_), h(h_) {
    stride = ((size_t)w * 3 + pitch_align - 1) / pitch_align * pitch_align;
    p = static_cast<uint8_t*>(aligned_alloc(256, (stride * h + 255) / 256 * 256 + 256));
    memset(p, 0xEE, stride * h);   // pitch padding holds garbage, as after a cudaMalloc
    for (int y = 0; y < h; y++) memcpy(p + (size_t)y * stride, src + (size_t)y * sstride, (size_t)w * 3);
  }
  ~DevImg() { free(p); }
  DevImg(const DevImg&) = delete;
};
const char* g_error = nullptr;
void run(dim3 grid, dim3 block, const std::function<void()>& body) {
  const char* e = emu::launch(grid, block, body, emu::SHUFFLED);
  if (e) g_error = e;
}

// launch_fast of warp.cu (kernel: 0 = quad, 1 = round-1 fast kernel)
template <int MODE>
void emu_launch_fast(int kernel, const uint8_t* left, size_t lstride, const uint8_t* right, size_t rstride, const WarpParams& P,
                     uint8_t* canvas, size_t cstride) {
  FastParams F;
  for (int i = 0; i < 3; i++) { F.Mx[i] = 32.0 * P.M[i]; F.My[i] = 32.0 * P.M[3 + i]; F.Mw[i] = P.M[6 + i]; }
  F.W = P;
  F.copy_words = ((reinterpret_cast<uintptr_t>(canvas) & 3u) == 0 && (cstride & 3u) == 0) ? 1 : 0;
  const bool left_words = MODE != 1 || ((reinterpret_cast<uintptr_t>(left) & 3u) == 0 && (lstride & 3u) == 0);
  if (kernel == 0 && left_words) {
    run(dim3((P.cw + 127) / 128, (P.ch + 7) / 8), dim3(256),
        [&] { warp_quad_kernel<MODE>(left, lstride, right, rstride, F, canvas, cstride); });
    return;
  }
  run(dim3((P.cw + 255) / 256, (P.ch + 7) / 8), dim3(256),
      [&] { warp_fast_kernel<MODE, 5>(left, lstride, right, rstride, F, canvas, cstride); });
}
}  // namespace

extern "C" {

const char* wemu_last_error() { return g_error ? g_error : ""; }

// stitchTwoImages' canvas for a homography (warp_overlay_device).  kernel: 0 quad (production), 1 round-1 fast kernel,
// 2 general kernel.  geom_out: cw, ch, offx, offy, used_fast.  Returns 1, 0 if the geometry is not ok, -1 on an
// emulation error, -2 if the canvas does not fit.
int wemu_overlay(const uint8_t* left, int wl, int hl, size_t sl, const uint8_t* right, int wr, int hr, size_t sr, const double* H,
                 int kernel, int src_pitch_align, uint8_t* canvas_out, size_t cap, int* geom_out) {
  g_error = nullptr;
  CanvasGeom g;
  canvas_geometry(wl, hl, wr, hr, H, &g);
  geom_out[0] = g.cw; geom_out[1] = g.ch; geom_out[2] = g.offx; geom_out[3] = g.offy; geom_out[4] = 0;
  if (!g.ok) return 0;
  if ((size_t)g.cw * 3 * g.ch > cap) return -2;
  DevImg L(left, wl, hl, sl, (size_t)src_pitch_align), R(right, wr, hr, sr, (size_t)src_pitch_align);
  const size_t cstride = ((size_t)g.cw * 3 + 255) / 256 * 256;
  uint8_t* canvas = static_cast<uint8_t*>(aligned_alloc(256, cstride * g.ch + 256));
  memset(canvas, 0xAB, cstride * g.ch);
  WarpParams P;
  memcpy(P.M, g.Minv, sizeof P.M);
  P.bw0 = g.bw0;
  P.cw = g.cw; P.ch = g.ch;
  P.offx = g.offx; P.offy = g.offy; P.wl = wl; P.hl = hl;
  P.ws = wr; P.hs = hr;
  P.y0 = 0;
  P.src_bytes = (size_t)(hr - 1) * R.stride + (size_t)wr * 3;
  const bool box = footprint_box(g.TH, g.Minv, wr, hr, g.cw, g.ch, P);
  if (kernel != 2 && fast_path_ok(P, R.p, R.stride, box)) {
    geom_out[4] = 1;
    emu_launch_fast<1>(kernel, L.p, L.stride, R.p, R.stride, P, canvas, cstride);
  } else {
    run(dim3(((g.cw + 3) / 4 + 31) / 32, (g.ch + 7) / 8), dim3(32, 8),
        [&] { warp_overlay_kernel<1>(L.p, L.stride, R.p, R.stride, P, canvas, cstride); });
  }
  // pack_rows_device: pitched canvas -> tightly packed rows
  const unsigned long long total = (unsigned long long)g.cw * 3 * g.ch, words = (total + 3) / 4;
  uint8_t* tight = static_cast<uint8_t*>(aligned_alloc(256, (total + 255) / 256 * 256 + 256));
  run(dim3((unsigned)((words + 255) / 256)), dim3(256),
      [&] { pack_rows_kernel(canvas, cstride, (uint32_t)((size_t)g.cw * 3), total, tight); });
  memcpy(canvas_out, tight, total);
  free(tight);
  free(canvas);
  return g_error ? -1 : 1;
}

// cv::warpPerspective(src, M, (dw, dh)) through warp_only_device's flow
int wemu_warp_perspective(const uint8_t* src, int w, int h, size_t stride, const double* M, int kernel, uint8_t* dst, int dw, int dh) {
  g_error = nullptr;
  DevImg S(src, w, h, stride);
  double Minv[9];
  invert33(M, Minv);
  const int bh0 = dh < 16 ? dh : 16;
  int bw0 = 1024 / (bh0 < 1 ? 1 : bh0);
  if (bw0 > dw) bw0 = dw;
  const size_t dstride = ((size_t)dw * 3 + 255) / 256 * 256;
  uint8_t* d = static_cast<uint8_t*>(aligned_alloc(256, dstride * dh + 256));
  memset(d, 0xAB, dstride * dh);
  WarpParams P;
  memcpy(P.M, Minv, sizeof P.M);
  P.bw0 = bw0 < 1 ? 1 : bw0;
  P.cw = dw; P.ch = dh;
  P.offx = P.offy = 0; P.wl = P.hl = 0;
  P.ws = w; P.hs = h;
  P.y0 = 0;
  P.src_bytes = (size_t)(h - 1) * S.stride + (size_t)w * 3;
  bool box = footprint_box(M, Minv, w, h, dw, dh, P);
  if (kernel != 2 && fast_path_ok(P, S.p, S.stride, box))
    emu_launch_fast<0>(kernel, nullptr, 0, S.p, S.stride, P, d, dstride);
  else
    run(dim3(((dw + 3) / 4 + 31) / 32, (dh + 7) / 8), dim3(32, 8), [&] { warp_overlay_kernel<0>(nullptr, 0, S.p, S.stride, P, d, dstride); });
  for (int y = 0; y < dh; y++) memcpy(dst + (size_t)y * dw * 3, d + (size_t)y * dstride, (size_t)dw * 3);
  free(d);
  return g_error ? -1 : 1;
}

// pano_warp_accumulate on a band of canvas rows (warp_accumulate_device): band = tightly packed rows [y0, y0 + bh)
int wemu_accumulate(const uint8_t* src, int w, int h, size_t stride, const double* M, int kernel, uint8_t* band, int cw, int ch,
                    int y0, int bh) {
  g_error = nullptr;
  DevImg S(src, w, h, stride);
  const size_t bstride = ((size_t)cw * 3 + 255) / 256 * 256;
  uint8_t* b = static_cast<uint8_t*>(aligned_alloc(256, bstride * bh + 256));
  for (int y = 0; y < bh; y++) memcpy(b + (size_t)y * bstride, band + (size_t)y * cw * 3, (size_t)cw * 3);
  WarpParams P;
  double Minv[9];
  invert33(M, Minv);
  memcpy(P.M, Minv, sizeof P.M);
  const int bh0 = ch < 16 ? ch : 16;
  int bw0 = 1024 / (bh0 < 1 ? 1 : bh0);
  if (bw0 > cw) bw0 = cw;
  P.bw0 = bw0 < 1 ? 1 : bw0;
  P.cw = cw; P.ch = bh;
  P.offx = P.offy = 0; P.wl = P.hl = 0;
  P.ws = w; P.hs = h;
  P.y0 = y0;
  P.src_bytes = (size_t)(h - 1) * S.stride + (size_t)w * 3;
  const bool box = footprint_box(M, Minv, w, h, cw, ch, P);
  if (kernel != 2 && fast_path_ok(P, S.p, S.stride, box))
    emu_launch_fast<2>(kernel, nullptr, 0, S.p, S.stride, P, b, bstride);
  else
    run(dim3(((cw + 3) / 4 + 31) / 32, (bh + 7) / 8), dim3(32, 8), [&] { warp_overlay_kernel<2>(nullptr, 0, S.p, S.stride, P, b, bstride); });
  for (int y = 0; y < bh; y++) memcpy(band + (size_t)y * cw * 3, b + (size_t)y * bstride, (size_t)cw * 3);
  free(b);
  return g_error ? -1 : 1;
}

}  // extern "C"
