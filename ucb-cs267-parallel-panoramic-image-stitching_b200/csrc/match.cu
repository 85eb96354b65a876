// match.cu — K3 descriptor gather (warp per keypoint) and the SIMT SSD matcher, plus the
// conversion of per-query (ssd, j) minima into the reference's match list.
//
// Semantics: ref src/serial/main.cpp:188-244 (seqHarrisMatchKeyPoints):
//  - a keypoint takes part only if its patch lies inside its own image
//    (x >= b, y >= b, x + b < cols, y + b < rows with b = patch/2), on both sides;
//  - SSD over patch*patch*3 uint8, exact integer;
//  - the best train keypoint is the FIRST strict minimum in train order;
//  - a match is emitted, in ascending query order, iff best SSD < maxSSDThresh.
// The tensor-core matcher (match_tc.cu) produces the same packed minima; this SIMT kernel is
// the cross-check kernel for it and the matcher for debugging (pano_set_matcher(ctx, 1)).
#include "common.cuh"

namespace pano {

namespace {

#include "match_kernels.cuh"

}  // namespace

void update_pano_keypoints_device(cudaStream_t st, const int32_t* old_xy, int n_old, int offx, int offy, const int32_t* new_xy,
                                  int n_new, const double* TH, int cw, int ch, int32_t* out) {
  const int n = n_old + n_new;
  if (n <= 0) return;
  Mat33 m;
  memcpy(m.m, TH, sizeof m.m);
  update_pano_keypoints_kernel<<<(n + 255) / 256, 256, 0, st>>>(old_xy, n_old, offx, offy, new_xy, n_new, m, cw, ch, out);
  PANO_LAUNCH_CHECK();
}

int build_descriptors_device(cudaStream_t st, const DevImage& img, const int32_t* xy, int n, int patch,
                             MatchScratch& s, DevDescriptors& d, PinnedBuf& pin) {
  d.count = 0;
  if (n <= 0) return 0;
  s.flags.reserve((size_t)n);
  s.cnt.reserve(sizeof(uint32_t));
  d.orig.reserve(sizeof(int32_t) * (size_t)n);
  pin.reserve(64);
  ProfScope ps(PROF_DESC, st);
  border_flags_kernel<<<(n + 255) / 256, 256, 0, st>>>(xy, n, img.w, img.h, patch / 2, s.flags.as<uint8_t>());
  PANO_LAUNCH_CHECK();
  compact_flagged(st, s.flags.as<uint8_t>(), n, d.orig.as<int32_t>(), s.cnt.as<uint32_t>(), s.tmp);
  PANO_CUDA(cudaMemcpyAsync(pin.p, s.cnt.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  PANO_CUDA(stream_wait(st));
  int n_in = (int)*pin.as<uint32_t>();
  d.count = n_in;
  if (n_in == 0) return 0;
  // rows padded to a multiple of 256 so tensor-core tiles never read past the allocation
  size_t rows = ((size_t)n_in + 255) / 256 * 256;
  d.desc.reserve(rows * PANO_DESC_STRIDE);
  d.norm.reserve(rows * sizeof(uint32_t));
  PANO_CUDA(cudaMemsetAsync(d.desc.p, 0, rows * PANO_DESC_STRIDE, st));
  PANO_CUDA(cudaMemsetAsync(d.norm.p, 0, rows * sizeof(uint32_t), st));
  int wpb = 8;
  gather_desc_kernel<<<(n_in + wpb - 1) / wpb, wpb * 32, 0, st>>>(img.p, img.w, img.h, img.stride, xy,
                                                                 d.orig.as<int32_t>(), n_in, patch,
                                                                 d.desc.as<uint8_t>(), d.norm.as<uint32_t>());
  PANO_LAUNCH_CHECK();
  return n_in;
}

void match_simt_device(cudaStream_t st, const DevDescriptors& q, const DevDescriptors& t,
                       unsigned long long* best) {
  PANO_CUDA(cudaMemsetAsync(best, 0xff, sizeof(unsigned long long) * (size_t)q.count, st));
  if (q.count == 0 || t.count == 0) return;
  int gx = (q.count + MQ - 1) / MQ;
  // split the train set so the grid covers the 148 SMs a few times over
  int splits = (148 * 4 + gx - 1) / gx;
  int max_splits = (t.count + MT_TILE - 1) / MT_TILE;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int per = ((t.count + splits - 1) / splits + MT_TILE - 1) / MT_TILE * MT_TILE;
  splits = (t.count + per - 1) / per;
  match_simt_kernel<<<dim3(gx, splits), MQ, 0, st>>>(q.desc.as<uint8_t>(), q.count, t.desc.as<uint8_t>(),
                                                     t.count, per, best);
  PANO_LAUNCH_CHECK();
}

int emit_matches_device(cudaStream_t st, const DevDescriptors& q, const DevDescriptors& t,
                        const unsigned long long* best, double max_ssd, int offset, int patch,
                        MatchScratch& s, pano_dmatch* out_dev, PinnedBuf& pin, int* errw) {
  const int nq = q.count;
  if (nq == 0 || t.count == 0) return 0;
  // every SSD is <= patch^2*3*255^2; above that bound the threshold can never reject
  const double ssd_bound = (double)patch * patch * 3 * 255.0 * 255.0;
  const bool need_filter = !(max_ssd > ssd_bound);
  if (!need_filter) {
    emit_matches_kernel<<<(nq + 255) / 256, 256, 0, st>>>(best, nq, q.orig.as<int32_t>(), t.orig.as<int32_t>(),
                                                         max_ssd, offset, out_dev, nullptr, errw);
    PANO_LAUNCH_CHECK();
    return nq;
  }
  s.mflags.reserve((size_t)nq);
  s.midx.reserve(sizeof(int32_t) * (size_t)nq);
  s.mtmp.reserve(sizeof(pano_dmatch) * (size_t)nq);
  s.cnt.reserve(sizeof(uint32_t));
  pin.reserve(64);
  emit_matches_kernel<<<(nq + 255) / 256, 256, 0, st>>>(best, nq, q.orig.as<int32_t>(), t.orig.as<int32_t>(),
                                                       max_ssd, offset, s.mtmp.as<pano_dmatch>(),
                                                       s.mflags.as<uint8_t>(), errw);
  PANO_LAUNCH_CHECK();
  compact_flagged(st, s.mflags.as<uint8_t>(), nq, s.midx.as<int32_t>(), s.cnt.as<uint32_t>(), s.tmp);
  PANO_CUDA(cudaMemcpyAsync(pin.p, s.cnt.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  PANO_CUDA(stream_wait(st));
  int m = (int)*pin.as<uint32_t>();
  if (m > 0) {
    gather_matches_kernel<<<(m + 255) / 256, 256, 0, st>>>(s.mtmp.as<pano_dmatch>(), s.midx.as<int32_t>(), m, out_dev);
    PANO_LAUNCH_CHECK();
  }
  return m;
}

}  // namespace pano
