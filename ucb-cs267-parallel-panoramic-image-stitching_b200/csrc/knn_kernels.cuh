// knn_kernels.cuh — the SIMT kernels of the opt-in 2-NN / Lowe-ratio matcher (pano_match_knn; semantics in
// knn_core.cuh).  Kept in a header of their own so that the very same source is compiled twice: by nvcc into
// libpano_b200.so (knn.cu) and by g++ on top of a CPU emulation of the CUDA execution model
// (tests/hostsim/cuda_emu.hpp: fibers for threads, real barriers and warp shuffles) for the no-GPU test tier.
// The kernels therefore use only: thread/block indices, static __shared__ arrays, __syncthreads, warp shuffles,
// __popc, __vabsdiffu4, __dp4a, 64-bit atomicMin and 32-bit atomicOr.
#pragma once
#include "../../include/pano_b200.h"
#include "knn_core.cuh"

namespace pano {

constexpr int KNN_DESC_STRIDE = 128;   // bytes per patch-descriptor row (= PANO_DESC_STRIDE, checked in knn.cu)
constexpr int KNN_ERRW_NO_BEST = 2;    // (= PANO_ERRW_NO_BEST)
constexpr int KQ = 128;                // queries per block of the SSD kernel (one per thread)
constexpr int KT_TILE = 64;            // train descriptors staged per shared-memory tile
constexpr int KDW = 20;                // 32-bit words of a patch descriptor that can be non-zero (80 B >= 75)

// ---- patch descriptors: exact SSD, one query per thread, train tiles staged in shared memory -------------------
// grid (ceil(nq / KQ), splits): block (x, y) scans train rows [y * t_per_split, (y + 1) * t_per_split)
__global__ void __launch_bounds__(KQ)
knn_ssd_simt_kernel(const uint8_t* __restrict__ qd, int nq, const uint8_t* __restrict__ td, int nt, int t_per_split,
                    unsigned long long* __restrict__ best1, unsigned long long* __restrict__ best2) {
  __shared__ uint4 stile[KT_TILE][KDW / 4];
  const int qi = blockIdx.x * KQ + threadIdx.x;
  uint32_t q[KDW];
  if (qi < nq) {
    const uint4* src = reinterpret_cast<const uint4*>(qd + (size_t)qi * KNN_DESC_STRIDE);
#pragma unroll
    for (int k = 0; k < KDW / 4; k++) {
      const uint4 v = src[k];
      q[4 * k] = v.x; q[4 * k + 1] = v.y; q[4 * k + 2] = v.z; q[4 * k + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < KDW; k++) q[k] = 0;
  }
  const int t0 = blockIdx.y * t_per_split;
  const int t1 = min(nt, t0 + t_per_split);
  Top2 top = top2_empty();
  for (int tb = t0; tb < t1; tb += KT_TILE) {
    __syncthreads();
    for (int e = threadIdx.x; e < KT_TILE * (KDW / 4); e += KQ) {
      const int r = e / (KDW / 4), c = e % (KDW / 4);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (tb + r < t1) v = reinterpret_cast<const uint4*>(td + (size_t)(tb + r) * KNN_DESC_STRIDE)[c];
      stile[r][c] = v;
    }
    __syncthreads();
    const int lim = min(KT_TILE, t1 - tb);
    for (int r = 0; r < lim; r++) {
      uint32_t ssd = 0;
#pragma unroll
      for (int k = 0; k < KDW / 4; k++) {
        const uint4 v = stile[r][k];
        uint32_t d;
        d = __vabsdiffu4(q[4 * k], v.x);     ssd = __dp4a(d, d, ssd);
        d = __vabsdiffu4(q[4 * k + 1], v.y); ssd = __dp4a(d, d, ssd);
        d = __vabsdiffu4(q[4 * k + 2], v.z); ssd = __dp4a(d, d, ssd);
        d = __vabsdiffu4(q[4 * k + 3], v.w); ssd = __dp4a(d, d, ssd);
      }
      top2_insert(top, knn_key(ssd, (uint32_t)(tb + r)));
    }
  }
  if (qi < nq) knn_publish(best1, best2, qi, top);
}

// ---- binary descriptors ------------------------------------------------------------------------------------------
struct KnnBinPairs {   // positions compared by bit k (knn_bin_bit_positions), filled on the host
  uint8_t a[32 * KNN_BIN_WORDS], b[32 * KNN_BIN_WORDS];
};

// one warp per in-border keypoint: lanes 0..24 fetch one pixel of the 5 x 5 patch and convert it to gray; lane L
// then evaluates bits 8L .. 8L+7 on gray values fetched from the other lanes by shuffle; four lanes make a word
__global__ void knn_bin_desc_kernel(const uint8_t* __restrict__ img, size_t stride, const int32_t* __restrict__ xy,
                                    const int32_t* __restrict__ idx, int n_in, KnnBinPairs pairs,
                                    uint32_t* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= n_in) return;   // (the whole warp leaves together)
  const int i = idx[k];
  const int x = xy[2 * i], y = xy[2 * i + 1];
  int g = 0;
  if (lane < 25) {
    const int dy = lane / 5 - 2, dx = lane % 5 - 2;
    const uint8_t* p = img + (size_t)(y + dy) * stride + 3 * (size_t)(x + dx);
    g = gray_u8(p[0], p[1], p[2]);
  }
  uint32_t byte = 0;
#pragma unroll
  for (int i8 = 0; i8 < 8; i8++) {
    const int bit = 8 * lane + i8;
    const int ga = __shfl_sync(0xffffffffu, g, pairs.a[bit]);
    const int gb = __shfl_sync(0xffffffffu, g, pairs.b[bit]);
    byte |= (uint32_t)(ga < gb) << i8;
  }
  uint32_t v = byte << (8 * (lane & 3));
  v |= __shfl_xor_sync(0xffffffffu, v, 1);
  v |= __shfl_xor_sync(0xffffffffu, v, 2);
  if ((lane & 3) == 0) bits[(size_t)k * KNN_BIN_WORDS + (lane >> 2)] = v;
}

// one warp per query: the lanes stride over the train descriptors (XOR + __popc), keep their own two smallest keys
// and merge them with a butterfly of warp shuffles (the lane groups merged at each step are disjoint: no key twice)
__global__ void knn_hamming_kernel(const uint32_t* __restrict__ qb, int nq, const uint32_t* __restrict__ tb, int nt,
                                   unsigned long long* __restrict__ best1, unsigned long long* __restrict__ best2) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  const uint4* qs = reinterpret_cast<const uint4*>(qb + (size_t)q * KNN_BIN_WORDS);
  const uint4 q0 = qs[0], q1 = qs[1];
  Top2 top = top2_empty();
  for (int j = lane; j < nt; j += 32) {
    const uint4* ts = reinterpret_cast<const uint4*>(tb + (size_t)j * KNN_BIN_WORDS);
    const uint4 a = ts[0], b = ts[1];
    const uint32_t d = __popc(q0.x ^ a.x) + __popc(q0.y ^ a.y) + __popc(q0.z ^ a.z) + __popc(q0.w ^ a.w) +
                       __popc(q1.x ^ b.x) + __popc(q1.y ^ b.y) + __popc(q1.z ^ b.z) + __popc(q1.w ^ b.w);
    top2_insert(top, knn_key(d, (uint32_t)j));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Top2 other;
    other.k1 = __shfl_xor_sync(0xffffffffu, top.k1, o);
    other.k2 = __shfl_xor_sync(0xffffffffu, top.k2, o);
    top2_merge(top, other);
  }
  if (lane == 0) { best1[q] = top.k1; best2[q] = top.k2; }
}

// ---- (nearest, runner-up) keys -> match records + Lowe's test -----------------------------------------------------
// factor: ratio^2 for SSD keys, ratio for Hamming keys (knn_core.cuh).  Row i always gets a record; flags[i] says
// whether it passed (the host compacts the passing rows in ascending query order).
__global__ void knn_emit_kernel(const unsigned long long* __restrict__ best1, const unsigned long long* __restrict__ best2,
                                int nq, const int32_t* __restrict__ qorig, const int32_t* __restrict__ torig, double factor,
                                pano_dmatch* __restrict__ out, float* __restrict__ second, uint8_t* __restrict__ flags,
                                int* __restrict__ errw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const unsigned long long k1 = best1[i], k2 = best2[i];
  // every query row has a nearest neighbour when train descriptors exist; a row without one means the matcher did
  // not finish: flag it, the host fails the call when it reads the error word
  if (k1 == KNN_NONE) atomicOr(errw, KNN_ERRW_NO_BEST);
  const uint32_t d1 = (uint32_t)(k1 >> 32), d2 = (uint32_t)(k2 >> 32);
  const bool ok = k1 != KNN_NONE && k2 != KNN_NONE && lowe_accept(d1, d2, factor);
  pano_dmatch m;
  m.query_idx = qorig[i];
  m.train_idx = ok ? torig[(uint32_t)k1] : -1;
  m.distance = (float)d1;
  out[i] = m;
  second[i] = k2 != KNN_NONE ? (float)d2 : 0.f;
  flags[i] = ok;
}

__global__ void knn_gather_kernel(const pano_dmatch* __restrict__ in, const float* __restrict__ second_in,
                                  const int32_t* __restrict__ idx, int n, pano_dmatch* __restrict__ out,
                                  float* __restrict__ second_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    out[i] = in[idx[i]];
    second_out[i] = second_in[idx[i]];
  }
}

}  // namespace pano
