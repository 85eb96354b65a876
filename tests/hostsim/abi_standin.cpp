// abi_standin.cpp — TEST INFRASTRUCTURE ONLY: the handful of libpano_b200 entry points that
// examples/reference_shim/pano_b200_shim.cpp calls, implemented on the CPU oracle (liboracle's orc_* functions).
// tests/test_reference_shim.py links the reference's unmodified src/gpu/main.cpp + that shim against this stand-in to
// check, without a GPU, that the shim marshals the reference's C++ types correctly and that the reference's GPU
// executable then produces what its serial one produces.  It is never built into or loaded by the product.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pano_b200.h"

extern "C" {
struct orc_dmatch { int32_t queryIdx, trainIdx; float distance; };
void orc_convolve(const double* in, int w, int h, const double* kern, int ksize, double* out);
int orc_detect(const uint8_t* bgr, int w, int h, size_t stride, double k, double thresh, int nbhd, int32_t* xy, int cap);
int orc_match(const int32_t* kq, int nq, const int32_t* kt, int nt, const uint8_t* imq, int wq, int hq, size_t sq,
              const uint8_t* imt, int wt, int ht, size_t st, int patch, double maxSSD, int offset, orc_dmatch* out, int cap);
int orc_ransac(const int32_t* kp1, const int32_t* kp2, const orc_dmatch* matches, int m, int iters, int nsamples,
               double thr, uint32_t seed, int div_mode, double* H, int* best_count, int32_t* samples, int32_t* counts,
               uint8_t* inlier_mask, uint64_t* draws, int* best_iter);
int orc_match_knn(const int32_t* kq, int nq, const int32_t* kt, int nt, const uint8_t* imq, int wq, int hq, size_t sq,
                  const uint8_t* imt, int wt, int ht, size_t st, int patch, int descriptor, double ratio, orc_dmatch* out,
                  float* second, int cap);
void orc_mul33(const double* a, const double* b, double* d);
void orc_perspective_transform(const float* pts, int n, const double* H, float* out);
void orc_warp_perspective(const uint8_t* src, int sw, int sh, size_t sstride, const double* M, uint8_t* dst, int dw, int dh,
                          size_t dstride);
}

struct pano_ctx { uint32_t seed; std::string err; int pending = 0, pending_status = 0; };

extern "C" {
void pano_default_harris_opts(pano_harris_opts* o) { *o = {0.04, 1e6, 3, 5, 1e8}; }
void pano_default_ransac_opts(pano_ransac_opts* o) { *o = {1000, 4, 3.0}; }
int pano_create(int, uint32_t seed, pano_ctx** out) { *out = new pano_ctx{seed, "", 0, 0}; return PANO_OK; }
void pano_destroy(pano_ctx* c) { delete c; }
const char* pano_last_error(const pano_ctx* c) { return c ? c->err.c_str() : ""; }

int pano_detect(pano_ctx*, const uint8_t* bgr, int w, int h, size_t stride, int, const pano_harris_opts* o, int32_t* xy,
                int cap, int* count) {
  *count = orc_detect(bgr, w, h, stride, o->k, o->nms_thresh, o->nms_neighborhood, xy, cap);
  return (xy && *count > cap) ? PANO_ERR_CAPACITY : PANO_OK;
}

int pano_match(pano_ctx*, const int32_t* kq, int nq, const int32_t* kt, int nt, const uint8_t* imq, int wq, int hq,
               size_t sq, const uint8_t* imt, int wt, int ht, size_t st, int, const pano_harris_opts* o, int offset,
               pano_dmatch* out, int cap, int* count) {
  static_assert(sizeof(pano_dmatch) == sizeof(orc_dmatch), "layout");
  *count = orc_match(kq, nq, kt, nt, imq, wq, hq, sq, imt, wt, ht, st, o->patch_size, o->max_ssd_thresh, offset,
                     (orc_dmatch*)out, cap);
  return *count > cap ? PANO_ERR_CAPACITY : PANO_OK;
}

int pano_ransac(pano_ctx* c, const int32_t* kp1, int, const int32_t* kp2, int, const pano_dmatch* m, int n, int,
                const pano_ransac_opts* o, double H[9], int* best, int* best_it, int32_t*, int32_t*, uint8_t*) {
  if (n < o->num_samples) return PANO_ERR_TOO_FEW_MATCHES;
  const int ok = orc_ransac(kp1, kp2, (const orc_dmatch*)m, n, o->num_iterations, o->num_samples, o->distance_threshold,
                            c->seed, 0, H, best, nullptr, nullptr, nullptr, nullptr, best_it);
  return ok == 1 ? PANO_OK : PANO_ERR_NO_HOMOGRAPHY;
}

int pano_convolve_f64(pano_ctx*, const double* in, int w, int h, const double* k, int ksize, int, double* out) {
  orc_convolve(in, w, h, k, ksize, out);
  return PANO_OK;
}
// ---- chain mode (host/chain_multi_gpu.hpp on the CPU tier: tests/hostsim/chain_host.cpp) --------------------------
// "device" memory is host memory here; the device ordinal is ignored.
int pano_pair_homography(pano_ctx* c, const uint8_t* left, int wl, int hl, size_t sl, const uint8_t* right, int wr, int hr,
                         size_t sr, int, const pano_harris_opts* ho, const pano_ransac_opts* ro, pano_pair_result* res) {
  std::vector<int32_t> kl(2 * (size_t)wl * hl / 4 + 16), kr(2 * (size_t)wr * hr / 4 + 16);
  const int nl = orc_detect(left, wl, hl, sl, ho->k, ho->nms_thresh, ho->nms_neighborhood, kl.data(), (int)kl.size() / 2);
  const int nr = orc_detect(right, wr, hr, sr, ho->k, ho->nms_thresh, ho->nms_neighborhood, kr.data(), (int)kr.size() / 2);
  std::vector<orc_dmatch> m((size_t)std::max(nr, 1));
  const int nm = orc_match(kr.data(), nr, kl.data(), nl, right, wr, hr, sr, left, wl, hl, sl, ho->patch_size,
                           ho->max_ssd_thresh, 0, m.data(), (int)m.size());
  memset(res, 0, sizeof *res);
  res->n_kp_left = nl; res->n_kp_right = nr; res->n_matches = nm; res->best_iteration = -1;
  if (nm == 0) return res->status = PANO_ERR_NO_MATCHES;
  if (nm < ro->num_samples) return res->status = PANO_ERR_TOO_FEW_MATCHES;
  const int ok = orc_ransac(kr.data(), kl.data(), m.data(), nm, ro->num_iterations, ro->num_samples, ro->distance_threshold,
                            c->seed, 0, res->H, &res->best_inliers, nullptr, nullptr, nullptr, nullptr, &res->best_iteration);
  return res->status = (ok == 1 ? PANO_OK : PANO_ERR_NO_HOMOGRAPHY);
}

void pano_mul33(const double A[9], const double B[9], double out[9]) { orc_mul33(A, B, out); }

// ref: src/serial/main.cpp:335-369 with every image i >= 1 in the role of "right" (float min / max, ceil in float)
int pano_chain_geometry(int n, const int* ws, const int* hs, const double* Hs, pano_canvas_info* out) {
  float minX = 0, minY = 0, maxX = (float)ws[0], maxY = (float)hs[0];
  for (int i = 1; i < n; i++) {
    const float pts[8] = {0.f, 0.f, (float)ws[i], 0.f, (float)ws[i], (float)hs[i], 0.f, (float)hs[i]};
    float q[8];
    orc_perspective_transform(pts, 4, Hs + 9 * (size_t)i, q);
    for (int k = 0; k < 4; k++) {
      minX = std::min(minX, q[2 * k]); maxX = std::max(maxX, q[2 * k]);
      minY = std::min(minY, q[2 * k + 1]); maxY = std::max(maxY, q[2 * k + 1]);
    }
  }
  const double T[9] = {1, 0, (double)(-minX), 0, 1, (double)(-minY), 0, 0, 1};
  memcpy(out->TH, T, sizeof T);
  out->canvas_w = (int)std::ceil(maxX - minX);
  out->canvas_h = (int)std::ceil(maxY - minY);
  out->left_x = (int)(-minX);
  out->left_y = (int)(-minY);
  return out->canvas_w > 0 && out->canvas_h > 0 ? PANO_OK : PANO_ERR_ROI;
}

// cv::warpPerspective of the whole canvas, then the reference's overlay rule (ref: :380-386) on rows [y0, y0 + band_h)
int pano_warp_accumulate(pano_ctx*, const uint8_t* src, int w, int h, size_t stride, int, const double M[9], uint8_t* band,
                         int canvas_w, int canvas_h, int y0, int band_h, size_t band_stride) {
  std::vector<uint8_t> full((size_t)canvas_w * 3 * canvas_h);
  orc_warp_perspective(src, w, h, stride, M, full.data(), canvas_w, canvas_h, (size_t)canvas_w * 3);
  for (int y = 0; y < band_h; y++)
    for (int x = 0; x < canvas_w; x++) {
      const uint8_t* p = &full[((size_t)(y0 + y) * canvas_w + x) * 3];
      if (p[0] | p[1] | p[2]) memcpy(band + (size_t)y * band_stride + 3 * (size_t)x, p, 3);
    }
  return PANO_OK;
}
// ---- opt-in calls and the asynchronous forms (the Python binding on the CPU tier: tests/test_python_binding.py) -------
int pano_set_matcher(pano_ctx*, int) { return PANO_OK; }
void pano_default_knn_opts(pano_knn_opts* o) { *o = {5, PANO_KNN_PATCH_SSD, 0.75}; }

int pano_match_knn(pano_ctx*, const int32_t* kq, int nq, const int32_t* kt, int nt, const uint8_t* imq, int wq, int hq, size_t sq,
                   const uint8_t* imt, int wt, int ht, size_t st, int, const pano_knn_opts* o, pano_dmatch* out, float* second,
                   int cap, int* count) {
  if (!(o->ratio > 0.0) || o->ratio > 1.0) return PANO_ERR_INVALID;
  if (o->descriptor != PANO_KNN_PATCH_SSD && o->descriptor != PANO_KNN_BINARY) return PANO_ERR_UNSUPPORTED;
  if ((o->patch_size != 1 && o->patch_size != 3 && o->patch_size != 5) || (o->descriptor == PANO_KNN_BINARY && o->patch_size != 5))
    return PANO_ERR_UNSUPPORTED;
  *count = orc_match_knn(kq, nq, kt, nt, imq, wq, hq, sq, imt, wt, ht, st, o->patch_size, o->descriptor, o->ratio,
                         (orc_dmatch*)out, second, cap);
  return *count > cap ? PANO_ERR_CAPACITY : PANO_OK;
}

// the worker of the real library is replaced by an immediate call; completion is reported by query / wait as there
static int finish_async(pano_ctx* c, int status) {
  if (c->pending) return PANO_ERR_BUSY;
  c->pending = 1;
  c->pending_status = status;
  return PANO_OK;
}
int pano_detect_async(pano_ctx* c, const uint8_t* bgr, int w, int h, size_t stride, int mem, const pano_harris_opts* o, int32_t* xy,
                      int cap, int* count, void*) {
  if (c->pending) return PANO_ERR_BUSY;
  return finish_async(c, pano_detect(c, bgr, w, h, stride, mem, o, xy, cap, count));
}
int pano_match_async(pano_ctx* c, const int32_t* kq, int nq, const int32_t* kt, int nt, const uint8_t* imq, int wq, int hq,
                     size_t sq, const uint8_t* imt, int wt, int ht, size_t st, int mem, const pano_harris_opts* o, int offset,
                     pano_dmatch* out, int cap, int* count, void*) {
  if (c->pending) return PANO_ERR_BUSY;
  return finish_async(c, pano_match(c, kq, nq, kt, nt, imq, wq, hq, sq, imt, wt, ht, st, mem, o, offset, out, cap, count));
}
int pano_ransac_async(pano_ctx* c, const int32_t* kp1, int n1, const int32_t* kp2, int n2, const pano_dmatch* m, int n, int mem,
                      const pano_ransac_opts* o, double H[9], int* best, int* best_it, int32_t* s, int32_t* cn, uint8_t* mask, void*) {
  if (c->pending) return PANO_ERR_BUSY;
  return finish_async(c, pano_ransac(c, kp1, n1, kp2, n2, m, n, mem, o, H, best, best_it, s, cn, mask));
}
int pano_pair_query(pano_ctx* c) {
  if (!c->pending) return PANO_ERR_INVALID;
  c->pending = 0;
  return c->pending_status;
}
int pano_pair_wait(pano_ctx* c) { return pano_pair_query(c); }
}
