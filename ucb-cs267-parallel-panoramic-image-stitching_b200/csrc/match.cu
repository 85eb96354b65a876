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

__global__ void border_flags_kernel(const int32_t* __restrict__ xy, int n, int w, int h, int b,
                                    uint8_t* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int x = xy[2 * i], y = xy[2 * i + 1];
  flags[i] = !(x < b || y < b || x + b >= w || y + b >= h);
}

// one warp per in-border keypoint: lanes 0..p*p-1 each fetch one BGR pixel of the patch
__global__ void gather_desc_kernel(const uint8_t* __restrict__ img, int w, int h, size_t stride,
                                   const int32_t* __restrict__ xy, const int32_t* __restrict__ idx, int n_in,
                                   int patch, uint8_t* __restrict__ desc, uint32_t* __restrict__ norm) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= n_in) return;
  const int i = idx[k];
  const int x = xy[2 * i], y = xy[2 * i + 1];
  const int b = patch / 2, pp = patch * patch;
  uint32_t s = 0;
  if (lane < pp) {
    int dy = lane / patch - b, dx = lane % patch - b;
    const uint8_t* p = img + (size_t)(y + dy) * stride + 3 * (size_t)(x + dx);
    uint8_t c0 = p[0], c1 = p[1], c2 = p[2];
    uint8_t* d = desc + (size_t)k * PANO_DESC_STRIDE + 3 * lane;
    d[0] = c0; d[1] = c1; d[2] = c2;
    s = (uint32_t)c0 * c0 + (uint32_t)c1 * c1 + (uint32_t)c2 * c2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) norm[k] = s;
}

constexpr int MQ = 128;      // queries per block (one per thread)
constexpr int MT_TILE = 64;  // train descriptors staged per smem tile
constexpr int DW = 20;       // 32-bit words of a descriptor that can be non-zero (80 B >= 75)

__global__ void __launch_bounds__(MQ)
match_simt_kernel(const uint8_t* __restrict__ qd, int nq, const uint8_t* __restrict__ td, int nt,
                  int t_per_split, unsigned long long* __restrict__ best) {
  __shared__ uint4 stile[MT_TILE][DW / 4];
  const int qi = blockIdx.x * MQ + threadIdx.x;
  uint32_t q[DW];
  if (qi < nq) {
    const uint4* src = reinterpret_cast<const uint4*>(qd + (size_t)qi * PANO_DESC_STRIDE);
#pragma unroll
    for (int k = 0; k < DW / 4; k++) {
      uint4 v = src[k];
      q[4 * k] = v.x; q[4 * k + 1] = v.y; q[4 * k + 2] = v.z; q[4 * k + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < DW; k++) q[k] = 0;
  }
  const int t0 = blockIdx.y * t_per_split;
  const int t1 = min(nt, t0 + t_per_split);
  uint32_t bs = 0xffffffffu, bj = 0xffffffffu;
  for (int tb = t0; tb < t1; tb += MT_TILE) {
    __syncthreads();
    for (int e = threadIdx.x; e < MT_TILE * (DW / 4); e += MQ) {
      int r = e / (DW / 4), c = e % (DW / 4);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (tb + r < t1) v = reinterpret_cast<const uint4*>(td + (size_t)(tb + r) * PANO_DESC_STRIDE)[c];
      stile[r][c] = v;
    }
    __syncthreads();
    const int lim = min(MT_TILE, t1 - tb);
    for (int r = 0; r < lim; r++) {
      uint32_t ssd = 0;
#pragma unroll
      for (int k = 0; k < DW / 4; k++) {
        uint4 v = stile[r][k];
        uint32_t d;
        d = __vabsdiffu4(q[4 * k], v.x);     ssd = __dp4a(d, d, ssd);
        d = __vabsdiffu4(q[4 * k + 1], v.y); ssd = __dp4a(d, d, ssd);
        d = __vabsdiffu4(q[4 * k + 2], v.z); ssd = __dp4a(d, d, ssd);
        d = __vabsdiffu4(q[4 * k + 3], v.w); ssd = __dp4a(d, d, ssd);
      }
      if (ssd < bs) { bs = ssd; bj = (uint32_t)(tb + r); }  // strict <: first minimum wins
    }
  }
  if (qi < nq && bj != 0xffffffffu) {
    unsigned long long key = ((unsigned long long)bs << 32) | bj;
    atomicMin(&best[qi], key);
  }
}

__global__ void emit_matches_kernel(const unsigned long long* __restrict__ best, int nq,
                                    const int32_t* __restrict__ qorig, const int32_t* __restrict__ torig,
                                    double max_ssd, int offset, pano_dmatch* __restrict__ out,
                                    uint8_t* __restrict__ flags, int* __restrict__ errw) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  unsigned long long key = best[i];
  uint32_t ssd = (uint32_t)(key >> 32), j = (uint32_t)key;
  // every query row has a minimum when train descriptors exist; a row without one means the matcher did not
  // finish (aborted CTA): flag it, the host fails the call when it reads the error word
  if (key == ~0ull) atomicOr(errw, PANO_ERRW_NO_BEST);
  bool ok = key != ~0ull && (double)ssd < max_ssd;
  pano_dmatch m;
  m.query_idx = qorig[i] + offset;
  m.train_idx = ok ? torig[j] : -1;
  m.distance = (float)ssd;
  out[i] = m;
  if (flags) flags[i] = ok;
}

__global__ void gather_matches_kernel(const pano_dmatch* __restrict__ in, const int32_t* __restrict__ idx, int n,
                                      pano_dmatch* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}

// Incremental fold (opt-in): the panorama's keypoint list after a step = the old list shifted by the left
// image's offset in the new canvas, followed by the new image's keypoints mapped through T*H
// (cv::perspectiveTransform arithmetic, rounded to the nearest pixel, ties to even); points that leave the canvas
// become (-1, -1), which the matcher's in-border test skips.
struct Mat33 { double m[9]; };
__global__ void update_pano_keypoints_kernel(const int32_t* __restrict__ old_xy, int n_old, int offx, int offy,
                                             const int32_t* __restrict__ new_xy, int n_new, Mat33 TH, int cw, int ch,
                                             int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_old) {
    const int x = old_xy[2 * i], y = old_xy[2 * i + 1];
    out[2 * i] = x < 0 ? -1 : x + offx;
    out[2 * i + 1] = x < 0 ? -1 : y + offy;
  } else if (i < n_old + n_new) {
    const int j = i - n_old;
    float px, py;
    persp_point(TH.m, (float)new_xy[2 * j], (float)new_xy[2 * j + 1], &px, &py);
    const int x = __float2int_rn(px), y = __float2int_rn(py);
    const bool in = x >= 0 && y >= 0 && x < cw && y < ch;
    out[2 * i] = in ? x : -1;
    out[2 * i + 1] = in ? y : -1;
  }
}

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
