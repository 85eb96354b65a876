// chain_multi_gpu.hpp — chain-mode panorama over several GPUs inside ONE host process (the
// C++ counterpart of dist.py's stitch_chain_distributed, which does the same with one process per GPU).
//
// The reference folds its images sequentially and re-detects on the growing panorama (ref: src/serial/main.cpp:
// 395-414), which cannot be sharded.  Chain mode estimates H(i <- i+1) of every adjacent pair independently - pair i
// on device i mod D - composes them into the frame of image 0, computes the canvas of all images exactly as the
// reference does for two (ref: :335-369 via pano_chain_geometry) and lets every device render its own band of
// canvas rows: image 0 at its integer offset, every other image warped by T * H(0 <- i) with the reference's
// "non-black pixels overwrite" rule (ref: :380-386).  The per-pair homographies (the only data the devices exchange)
// travel through host memory - one process, so no collective is needed; the bands go device -> host canvas directly.
// A pair that fails (the reference's empty-Mat cases) ends the chain: the panorama covers images 0 .. k.
//
// One worker thread per device and phase; a device's context is only ever used by one thread at a time
// (pano_b200.h: a context is not thread-safe).  `Mem` supplies device memory (CudaMem in gpu_stitching.cpp; the CPU test tier
// instantiates the same code with host memory on a CPU stand-in of the C ABI, tests/hostsim/chain_host.cpp).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pano_b200.h"

namespace pano_host {

struct ImageView {   // 8-bit BGR, interleaved, host memory
  const uint8_t* p;
  int w, h;
  size_t stride;
};

struct ChainOutput {
  std::vector<uint8_t> canvas;   // canvas_h rows of 3 * canvas_w bytes
  int w = 0, h = 0;
  int n_used = 0;                            // images 0 .. n_used - 1 are in the panorama
  std::vector<pano_pair_result> pairs;       // one record per adjacent pair
  std::vector<int> pair_device;              // device ordinal that estimated it
  std::string error;
};

// rows of the canvas rendered by worker `d` of `D`: contiguous bands, remainder to the first workers
inline void band_rows(int canvas_h, int d, int D, int* y0, int* bh) {
  const int base = canvas_h / D, rem = canvas_h % D;
  *y0 = d * base + (d < rem ? d : rem);
  *bh = base + (d < rem ? 1 : 0);
}

template <class Mem>
int stitch_chain_multi_gpu(const std::vector<ImageView>& images, const std::vector<int>& devices, uint32_t seed,
                           const pano_harris_opts& hopts, const pano_ransac_opts& ropts, ChainOutput* out) {
  const int n = (int)images.size(), D = (int)devices.size();
  if (n < 2 || D < 1 || !out) return PANO_ERR_INVALID;
  const int n_pairs = n - 1;
  out->pairs.assign((size_t)n_pairs, pano_pair_result());
  out->pair_device.assign((size_t)n_pairs, -1);
  for (auto& r : out->pairs) r.status = PANO_ERR_INVALID;

  struct Worker {
    pano_ctx* ctx = nullptr;
    std::vector<uint8_t*> dimg;   // every input image on this device, rows pitched to 256 bytes (TMA / word loads)
    int status = PANO_OK;
    std::string err;
  };
  std::vector<Worker> W((size_t)D);
  auto pitch_of = [&](int i) { return ((size_t)images[(size_t)i].w * 3 + 255) / 256 * 256; };

  // ---- phase 1: contexts, inputs, the adjacent pairs of each device -------------------------------------------
  auto phase1 = [&](int d) {
    Worker& k = W[(size_t)d];
    Mem::set_device(devices[(size_t)d]);
    k.status = pano_create(devices[(size_t)d], seed, &k.ctx);
    if (k.status != PANO_OK) { k.err = "pano_create failed"; return; }
    k.dimg.assign((size_t)n, nullptr);
    for (int i = 0; i < n; i++) {
      const ImageView& im = images[(size_t)i];
      k.dimg[(size_t)i] = static_cast<uint8_t*>(Mem::alloc(pitch_of(i) * (size_t)im.h));
      if (!k.dimg[(size_t)i] || !Mem::h2d_2d(k.dimg[(size_t)i], pitch_of(i), im.p, im.stride, (size_t)im.w * 3, im.h)) {
        k.status = PANO_ERR_CUDA;
        k.err = "device allocation / upload failed";
        return;
      }
    }
    Mem::sync();   // the uploads have landed before the engine's own (non-blocking) stream reads them
    for (int i = d; i < n_pairs; i += D) {
      const ImageView& L = images[(size_t)i];
      const ImageView& R = images[(size_t)i + 1];
      pano_pair_result r;
      memset(&r, 0, sizeof r);
      const int st = pano_pair_homography(k.ctx, k.dimg[(size_t)i], L.w, L.h, pitch_of(i), k.dimg[(size_t)i + 1], R.w, R.h,
                                          pitch_of(i + 1), PANO_MEM_DEVICE, &hopts, &ropts, &r);
      r.status = st;
      out->pairs[(size_t)i] = r;
      out->pair_device[(size_t)i] = devices[(size_t)d];
      if (st == PANO_ERR_CUDA || st == PANO_ERR_INVALID) {   // an engine failure, not one of the reference's empty results
        k.status = st;
        k.err = pano_last_error(k.ctx);
        return;
      }
    }
  };
  auto run_all = [&](auto fn) {
    std::vector<std::thread> th;
    for (int d = 1; d < D; d++) th.emplace_back(fn, d);
    fn(0);
    for (auto& t : th) t.join();
  };
  auto cleanup = [&] {
    run_all([&](int d) {
      Worker& k = W[(size_t)d];
      Mem::set_device(devices[(size_t)d]);
      for (uint8_t* p : k.dimg) Mem::free(p);
      k.dimg.clear();
      if (k.ctx) pano_destroy(k.ctx);
      k.ctx = nullptr;
    });
  };
  run_all(phase1);
  for (const Worker& k : W)
    if (k.status != PANO_OK) {
      out->error = k.err;
      const int st = k.status;
      cleanup();
      return st;
    }

  // ---- compose H(0 <- i) over the connected prefix, canvas of all its images --------------------------------------
  std::vector<double> Hs(9, 0.0);
  Hs[0] = Hs[4] = Hs[8] = 1.0;
  int used = 1;
  for (int i = 0; i < n_pairs && out->pairs[(size_t)i].status == PANO_OK; i++, used++) {
    double next[9];
    pano_mul33(&Hs[9 * (size_t)i], out->pairs[(size_t)i].H, next);
    Hs.insert(Hs.end(), next, next + 9);
  }
  out->n_used = used;
  std::vector<int> ws((size_t)used), hs((size_t)used);
  for (int i = 0; i < used; i++) { ws[(size_t)i] = images[(size_t)i].w; hs[(size_t)i] = images[(size_t)i].h; }
  pano_canvas_info geom;
  const int gst = pano_chain_geometry(used, ws.data(), hs.data(), Hs.data(), &geom);
  if (gst != PANO_OK) {
    out->error = "chain canvas geometry failed";
    cleanup();
    return gst;
  }
  const int cw = geom.canvas_w, ch = geom.canvas_h;
  out->w = cw;
  out->h = ch;
  out->canvas.assign((size_t)cw * 3 * ch, 0);

  // ---- phase 2: every device renders its band of canvas rows straight into the host canvas --------------------
  auto phase2 = [&](int d) {
    Worker& k = W[(size_t)d];
    int y0, bh;
    band_rows(ch, d, D, &y0, &bh);
    if (bh <= 0) return;
    Mem::set_device(devices[(size_t)d]);
    const size_t pitch = ((size_t)cw * 3 + 255) / 256 * 256;
    uint8_t* band = static_cast<uint8_t*>(Mem::alloc(pitch * (size_t)bh));
    if (!band || !Mem::zero(band, pitch * (size_t)bh)) { k.status = PANO_ERR_CUDA; k.err = "band allocation failed"; Mem::free(band); return; }
    Mem::sync();
    for (int i = 0; i < used && k.status == PANO_OK; i++) {
      double M[9] = {1, 0, (double)geom.left_x, 0, 1, (double)geom.left_y, 0, 0, 1};
      if (i > 0) pano_mul33(geom.TH, &Hs[9 * (size_t)i], M);
      const ImageView& im = images[(size_t)i];
      k.status = pano_warp_accumulate(k.ctx, k.dimg[(size_t)i], im.w, im.h, pitch_of(i), PANO_MEM_DEVICE, M, band, cw, ch, y0, bh,
                                      pitch);
      if (k.status != PANO_OK) k.err = pano_last_error(k.ctx);
    }
    if (k.status == PANO_OK &&
        !Mem::d2h_2d(out->canvas.data() + (size_t)y0 * cw * 3, (size_t)cw * 3, band, pitch, (size_t)cw * 3, bh)) {
      k.status = PANO_ERR_CUDA;
      k.err = "band download failed";
    }
    Mem::free(band);
  };
  run_all(phase2);
  int st = PANO_OK;
  for (const Worker& k : W)
    if (k.status != PANO_OK) { st = k.status; out->error = k.err; }
  cleanup();
  return st;
}

}  // namespace pano_host
