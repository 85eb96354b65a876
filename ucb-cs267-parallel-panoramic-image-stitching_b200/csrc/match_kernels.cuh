// match_kernels.cuh — the device code of match.cu: in-border flags, the warp-per-keypoint descriptor gather (K3), the
// SIMT SSD matcher (cross-check of the tensor-core matcher), the conversion of per-query minima into the reference's
// match list, and the keypoint carry-over of the incremental fold.
//
// Included by match.cu INSIDE `namespace pano { namespace {` (no includes or namespaces of its own), and by the CPU
// emulation tier (tests/hostsim/match_emu.cpp on tests/hostsim/cuda_emu.hpp), which compiles the same source with g++
// and runs it thread by thread against the oracle.  Needs PANO_DESC_STRIDE, PANO_ERRW_NO_BEST, pano_dmatch and
// pano_core.cuh's persp_point in scope.
// Semantics: see the header of match.cu (ref src/serial/main.cpp:188-244).

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

