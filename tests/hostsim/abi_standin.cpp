// abi_standin.cpp — TEST INFRASTRUCTURE ONLY: the handful of libpano_b200 entry points that
// examples/reference_shim/pano_b200_shim.cpp calls, implemented on the CPU oracle (liboracle's orc_* functions).
// tests/test_reference_shim.py links the reference's unmodified src/gpu/main.cpp + that shim against this stand-in to
// check, without a GPU, that the shim marshals the reference's C++ types correctly and that the reference's GPU
// executable then produces what its serial one produces.  It is never built into or loaded by the product.
#include <algorithm>
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
}

struct pano_ctx { uint32_t seed; std::string err; };

extern "C" {
void pano_default_harris_opts(pano_harris_opts* o) { *o = {0.04, 1e6, 3, 5, 1e8}; }
void pano_default_ransac_opts(pano_ransac_opts* o) { *o = {1000, 4, 3.0}; }
int pano_create(int, uint32_t seed, pano_ctx** out) { *out = new pano_ctx{seed, ""}; return PANO_OK; }
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
}
