// knn.cu — host side of the opt-in 2-NN / Lowe-ratio matcher (pano_match_knn): launches the kernels of
// knn_kernels.cuh (and match_tc.cu's top-2 variant of the tensor-core matcher) and turns their per-query
// (nearest, runner-up) keys into the match list.  Semantics: knn_core.cuh.  Not on the reference's path
// (ref: src/serial/main.cpp:188-244 is pano_match); nothing in pano_stitch_* calls into this file.
#include "common.cuh"
#include "knn_kernels.cuh"

namespace pano {

static_assert(KNN_DESC_STRIDE == PANO_DESC_STRIDE, "patch descriptor row size");
static_assert(KNN_ERRW_NO_BEST == PANO_ERRW_NO_BEST, "error word bit");

namespace {

const KnnBinPairs& bin_pairs() {
  static const KnnBinPairs p = [] {
    KnnBinPairs t;
    for (int k = 0; k < 32 * KNN_BIN_WORDS; k++) {
      int a, b;
      knn_bin_bit_positions(k, &a, &b);
      t.a[k] = (uint8_t)a;
      t.b[k] = (uint8_t)b;
    }
    return t;
  }();
  return p;
}

void knn_ssd_simt_device(cudaStream_t st, const DevDescriptors& q, const DevDescriptors& t, unsigned long long* best1,
                         unsigned long long* best2) {
  PANO_CUDA(cudaMemsetAsync(best1, 0xff, sizeof(unsigned long long) * (size_t)q.count, st));
  PANO_CUDA(cudaMemsetAsync(best2, 0xff, sizeof(unsigned long long) * (size_t)q.count, st));
  if (q.count == 0 || t.count == 0) return;
  const int gx = (q.count + KQ - 1) / KQ;
  // split the train set so that the grid covers the 148 SMs a few times over
  int splits = (148 * 4 + gx - 1) / gx;
  const int max_splits = (t.count + KT_TILE - 1) / KT_TILE;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int per = ((t.count + splits - 1) / splits + KT_TILE - 1) / KT_TILE * KT_TILE;
  splits = (t.count + per - 1) / per;
  knn_ssd_simt_kernel<<<dim3(gx, splits), KQ, 0, st>>>(q.desc.as<uint8_t>(), q.count, t.desc.as<uint8_t>(), t.count, per,
                                                       best1, best2);
  PANO_LAUNCH_CHECK();
}

void knn_binary_device(cudaStream_t st, const DevImage& iq, const DevImage& it, const int32_t* kq, const int32_t* kt,
                       const DevDescriptors& q, const DevDescriptors& t, KnnScratch& ks, unsigned long long* best1,
                       unsigned long long* best2) {
  PANO_CUDA(cudaMemsetAsync(best1, 0xff, sizeof(unsigned long long) * (size_t)q.count, st));
  PANO_CUDA(cudaMemsetAsync(best2, 0xff, sizeof(unsigned long long) * (size_t)q.count, st));
  if (q.count == 0 || t.count == 0) return;
  ks.qbits.reserve(sizeof(uint32_t) * KNN_BIN_WORDS * (size_t)q.count);
  ks.tbits.reserve(sizeof(uint32_t) * KNN_BIN_WORDS * (size_t)t.count);
  const int wpb = 8;   // warps per block
  knn_bin_desc_kernel<<<(q.count + wpb - 1) / wpb, wpb * 32, 0, st>>>(iq.p, iq.stride, kq, q.orig.as<int32_t>(), q.count,
                                                                     bin_pairs(), ks.qbits.as<uint32_t>());
  PANO_LAUNCH_CHECK();
  knn_bin_desc_kernel<<<(t.count + wpb - 1) / wpb, wpb * 32, 0, st>>>(it.p, it.stride, kt, t.orig.as<int32_t>(), t.count,
                                                                     bin_pairs(), ks.tbits.as<uint32_t>());
  PANO_LAUNCH_CHECK();
  knn_hamming_kernel<<<(q.count + wpb - 1) / wpb, wpb * 32, 0, st>>>(ks.qbits.as<uint32_t>(), q.count,
                                                                    ks.tbits.as<uint32_t>(), t.count, best1, best2);
  PANO_LAUNCH_CHECK();
}

}  // namespace

// Leaves the matches (ascending query order) in ks.out / ks.out2 on the device and returns their number.
// q / t: the in-border keypoints of both sides with their patch descriptors (build_descriptors_device).
int match_knn_device(cudaStream_t st, const DevImage& iq, const DevImage& it, const int32_t* kq, const int32_t* kt,
                     const DevDescriptors& q, const DevDescriptors& t, const pano_knn_opts& o, bool use_tc,
                     MatchScratch& ms, DevBuf& best1, KnnScratch& ks, PinnedBuf& pin, int* errw) {
  const int nq = q.count;
  if (nq == 0 || t.count == 0) return 0;
  best1.reserve(sizeof(unsigned long long) * (size_t)nq);
  ks.best2.reserve(sizeof(unsigned long long) * (size_t)nq);
  unsigned long long* b1 = best1.as<unsigned long long>();
  unsigned long long* b2 = ks.best2.as<unsigned long long>();
  double factor;
  if (o.descriptor == PANO_KNN_BINARY) {
    knn_binary_device(st, iq, it, kq, kt, q, t, ks, b1, b2);
    factor = o.ratio;
  } else {
    if (use_tc) match_tc_device(st, q, t, b1, ms.tc_err, errw, b2);
    else knn_ssd_simt_device(st, q, t, b1, b2);
    factor = o.ratio * o.ratio;   // test on L2 distances, evaluated on the squared ones (knn_core.cuh)
  }
  ks.rec.reserve(sizeof(pano_dmatch) * (size_t)nq);
  ks.second.reserve(sizeof(float) * (size_t)nq);
  ks.out.reserve(sizeof(pano_dmatch) * (size_t)nq);
  ks.out2.reserve(sizeof(float) * (size_t)nq);
  ks.flags.reserve((size_t)nq);
  ks.idx.reserve(sizeof(int32_t) * (size_t)nq);
  ks.cnt.reserve(sizeof(uint32_t));
  pin.reserve(64);
  knn_emit_kernel<<<(nq + 255) / 256, 256, 0, st>>>(b1, b2, nq, q.orig.as<int32_t>(), t.orig.as<int32_t>(), factor,
                                                   ks.rec.as<pano_dmatch>(), ks.second.as<float>(),
                                                   ks.flags.as<uint8_t>(), errw);
  PANO_LAUNCH_CHECK();
  compact_flagged(st, ks.flags.as<uint8_t>(), nq, ks.idx.as<int32_t>(), ks.cnt.as<uint32_t>(), ks.tmp);
  PANO_CUDA(cudaMemcpyAsync(pin.p, ks.cnt.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  PANO_CUDA(stream_wait(st));
  const int m = (int)*pin.as<uint32_t>();
  if (m > 0) {
    knn_gather_kernel<<<(m + 255) / 256, 256, 0, st>>>(ks.rec.as<pano_dmatch>(), ks.second.as<float>(), ks.idx.as<int32_t>(),
                                                      m, ks.out.as<pano_dmatch>(), ks.out2.as<float>());
    PANO_LAUNCH_CHECK();
  }
  return m;
}

}  // namespace pano
