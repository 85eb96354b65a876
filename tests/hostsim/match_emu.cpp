// match_emu.cpp — runs the device code of the match stage (csrc/match_kernels.cuh: in-border flags, descriptor gather,
// SIMT SSD matcher, match emission, keypoint carry-over of the incremental fold) together with the flag compaction of
// harris_kernels.cuh on the CPU emulation of the CUDA execution model (cuda_emu.hpp), for the no-GPU test tier.
//
// TEST INFRASTRUCTURE ONLY.  The flow mirrors match.cu's host code (build_descriptors_device, match_simt_device,
// emit_matches_device); the kernels are the product's source compiled unchanged by g++.
#include "cuda_emu.hpp"

#include <cmath>
#include <memory>

#include "../../include/pano_b200.h"
#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/pano_core.cuh"

namespace pano {
constexpr int PANO_DESC_STRIDE = 128;     // as in common.cuh
constexpr int PANO_ERRW_NO_BEST = 2;
namespace {
#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/harris_kernels.cuh"
#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/match_kernels.cuh"
}  // namespace
}  // namespace pano

using namespace pano;

namespace {
template <typename T>
struct Aligned {
  T* p = nullptr;
  explicit Aligned(size_t n, int fill = 0) {
    const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) / 256 * 256;
    p = static_cast<T*>(aligned_alloc(256, bytes));
    memset(p, fill, bytes);
  }
  ~Aligned() { free(p); }
  Aligned(const Aligned&) = delete;
};
const char* g_error = nullptr;
void run(dim3 grid, dim3 block, const std::function<void()>& body, int order = emu::SHUFFLED) {
  const char* e = emu::launch(grid, block, body, order);
  if (e) g_error = e;
}
int compact(const uint8_t* flags, int n, int32_t* out_idx) {   // compact_flagged of harris.cu
  int nb = (n + 255) / 256;
  if (nb < 1) nb = 1;
  Aligned<uint32_t> bc((size_t)nb), bo((size_t)nb), cnt(1);
  run(dim3(nb), dim3(256), [&] { flag_count_kernel(flags, n, bc.p); });
  run(dim3(1), dim3(1024), [&] { scan_kernel(bc.p, bo.p, nb, cnt.p); }, 0);
  run(dim3(nb), dim3(256), [&] { flag_scatter_kernel(flags, n, bo.p, out_idx); });
  return (int)cnt.p[0];
}
struct Side {
  std::unique_ptr<Aligned<int32_t>> orig;
  std::unique_ptr<Aligned<uint8_t>> desc;
  std::unique_ptr<Aligned<uint32_t>> norm;
  int count = 0;
};
// build_descriptors_device
void build(const uint8_t* img, int w, int h, size_t stride, const int32_t* xy_host, int n, int patch, Side& s) {
  s.count = 0;
  if (n <= 0) return;
  Aligned<int32_t> xy((size_t)2 * n);
  memcpy(xy.p, xy_host, sizeof(int32_t) * 2 * (size_t)n);
  Aligned<uint8_t> flags((size_t)n);
  s.orig.reset(new Aligned<int32_t>((size_t)n));
  run(dim3((n + 255) / 256), dim3(256), [&] { border_flags_kernel(xy.p, n, w, h, patch / 2, flags.p); });
  const int n_in = compact(flags.p, n, s.orig->p);
  s.count = n_in;
  if (n_in == 0) return;
  const size_t rows = ((size_t)n_in + 255) / 256 * 256;
  s.desc.reset(new Aligned<uint8_t>(rows * PANO_DESC_STRIDE));
  s.norm.reset(new Aligned<uint32_t>(rows));
  const int wpb = 8;
  run(dim3((n_in + wpb - 1) / wpb), dim3(wpb * 32),
      [&] { gather_desc_kernel(img, w, h, stride, xy.p, s.orig->p, n_in, patch, s.desc->p, s.norm->p); });
}
}  // namespace

extern "C" {

const char* memu_last_error() { return g_error ? g_error : ""; }

// pano_match with the SIMT matcher.  Returns the number of matches, -1 on an emulation error, -2 if the device error
// word was raised.  norms_out (optional, n_query entries): squared norm of every in-border query descriptor.
int memu_match(const int32_t* kq, int nq, const int32_t* kt, int nt, const uint8_t* imq, int wq, int hq, size_t sq,
               const uint8_t* imt, int wt, int ht, size_t st, int patch, double max_ssd, int offset, int force_splits,
               pano_dmatch* out, int cap, uint32_t* norms_out) {
  g_error = nullptr;
  Side Q, T;
  build(imq, wq, hq, sq, kq, nq, patch, Q);
  build(imt, wt, ht, st, kt, nt, patch, T);
  if (Q.count == 0 || T.count == 0) return 0;
  if (norms_out) memcpy(norms_out, Q.norm->p, sizeof(uint32_t) * (size_t)Q.count);
  Aligned<unsigned long long> best((size_t)Q.count, 0xff);
  {   // match_simt_device
    const int gx = (Q.count + MQ - 1) / MQ;
    int splits = force_splits > 0 ? force_splits : (148 * 4 + gx - 1) / gx;
    const int max_splits = (T.count + MT_TILE - 1) / MT_TILE;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int per = ((T.count + splits - 1) / splits + MT_TILE - 1) / MT_TILE * MT_TILE;
    splits = (T.count + per - 1) / per;
    run(dim3(gx, splits), dim3(MQ), [&] { match_simt_kernel(Q.desc->p, Q.count, T.desc->p, T.count, per, best.p); });
  }
  // emit_matches_device
  Aligned<int> errw(1);
  Aligned<pano_dmatch> rec((size_t)Q.count), outd((size_t)Q.count);
  const double ssd_bound = (double)patch * patch * 3 * 255.0 * 255.0;
  int m = Q.count;
  if (max_ssd > ssd_bound) {
    run(dim3((Q.count + 255) / 256), dim3(256),
        [&] { emit_matches_kernel(best.p, Q.count, Q.orig->p, T.orig->p, max_ssd, offset, outd.p, nullptr, errw.p); });
  } else {
    Aligned<uint8_t> flags((size_t)Q.count);
    Aligned<int32_t> idx((size_t)Q.count);
    run(dim3((Q.count + 255) / 256), dim3(256),
        [&] { emit_matches_kernel(best.p, Q.count, Q.orig->p, T.orig->p, max_ssd, offset, rec.p, flags.p, errw.p); });
    m = compact(flags.p, Q.count, idx.p);
    if (m > 0) run(dim3((m + 255) / 256), dim3(256), [&] { gather_matches_kernel(rec.p, idx.p, m, outd.p); });
  }
  if (g_error) return -1;
  if (errw.p[0]) return -2;
  memcpy(out, outd.p, sizeof(pano_dmatch) * (size_t)std::min(m, cap));
  return m;
}

// update_pano_keypoints_kernel (incremental fold): out has n_old + n_new (x, y) pairs
int memu_update_keypoints(const int32_t* old_xy, int n_old, int offx, int offy, const int32_t* new_xy, int n_new, const double* TH,
                          int cw, int ch, int32_t* out) {
  g_error = nullptr;
  const int n = n_old + n_new;
  if (n <= 0) return 0;
  Mat33 m;
  memcpy(m.m, TH, sizeof m.m);
  run(dim3((n + 255) / 256), dim3(256), [&] { update_pano_keypoints_kernel(old_xy, n_old, offx, offy, new_xy, n_new, m, cw, ch, out); });
  return g_error ? -1 : n;
}

}  // extern "C"
