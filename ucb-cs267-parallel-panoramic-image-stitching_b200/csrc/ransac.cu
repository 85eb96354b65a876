// ransac.cu — K5 sample replay, K6 batched 4-point DLT, K7 inlier scoring + ordered argmax.
//
// Semantics: ref src/serial/main.cpp:247-307 (SeqRansacHomographyCalculator), seeded.
//
// K5 — replaying `std::shuffle(all M matches)` x num_iterations with ONE continuing
// std::mt19937 (ref :264-271).  Only the first four elements of each shuffled copy are used
// (:274-277), and the copy starts from the identity each iteration (:270), so an iteration's
// sample depends only on where in the engine's output stream it starts.  Lemire's rejection
// sampling makes every step consume a data-dependent number of outputs, so the start offset
// of iteration t depends on every earlier rejection.  The replay is made parallel by
// windowed speculation:
//   * the mt19937(seed) output stream is generated once per context and kept in HBM;
//   * iterations are processed in chunks of G; inside a chunk, iteration g is walked from
//     EVERY start offset in a window around its expected offset (the window grows like
//     sqrt(g) * sigma, sigma = std-dev of the rejections of one shuffle, from the exact
//     per-step rejection probabilities), one thread per (iteration, candidate offset);
//   * a single block then chains the per-candidate end offsets from the chunk's exact start
//     and picks each iteration's true candidate (exact: a miss is detected, never guessed,
//     and the caller re-runs with wider windows).
// The walk itself tracks only what lands in positions 0..3 (see pano_core.cuh).
//
// K6 — one thread per hypothesis runs OpenCV's 4-point findHomography path (normalised DLT,
// Jacobi eigen of the 9x9 LtL) in FP64 in OpenCV's exact operation order.
// K7 — one block per hypothesis counts inliers with the reference's mixed f64/f32 predicate;
// the first iteration with the strictly largest count wins (:295-298).
#include "common.cuh"
#include "replay_plan.hpp"

#include <algorithm>
#include <cmath>

namespace pano {

namespace {

// ---------------------------------------------------------------------------------------
// mt19937 output stream (std::mt19937: w=32 n=624 m=397 r=31 a=0x9908b0df u=11 s=7
// b=0x9d2c5680 t=15 c=0xefc60000 l=18, init multiplier 1812433253)
// ---------------------------------------------------------------------------------------
constexpr int MT_N = 624, MT_M = 397;

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

// One block.  state[624] persists in global memory between calls so the stream can be
// extended.  Generates `gens` blocks of 624 outputs into out[].
__global__ void __launch_bounds__(256) mt_generate_kernel(uint32_t* __restrict__ state, int init, uint32_t seed,
                                                         uint32_t* __restrict__ out, int gens) {
  __shared__ uint32_t mt[MT_N];
  const int tid = threadIdx.x;
  if (init) {
    if (tid == 0) {
      uint32_t x = seed;
      mt[0] = x;
      for (int i = 1; i < MT_N; i++) {
        x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i;
        mt[i] = x;
      }
    }
  } else {
    for (int i = tid; i < MT_N; i += blockDim.x) mt[i] = state[i];
  }
  __syncthreads();
  for (int g = 0; g < gens; g++) {
    // new[i] = twist(old[i], old[i+1], z) with z = old[i+397] for i < 227 and new[i-227]
    // after that.  Thread tid owns i = tid, 227+tid, 454+tid, so new[i-227] is its own
    // previous result; every old value is read before any write.
    uint32_t a1 = 0, b1 = 0, c1 = 0, a2 = 0, b2 = 0, a3 = 0, b3 = 0;
    if (tid < 227) {
      a1 = mt[tid]; b1 = mt[tid + 1]; c1 = mt[tid + MT_M];
      a2 = mt[227 + tid]; b2 = mt[228 + tid];
    }
    if (tid < 170) {
      a3 = mt[454 + tid];
      b3 = tid < 169 ? mt[455 + tid] : 0u;
    }
    __syncthreads();
    uint32_t n2 = 0;
    if (tid < 227) {
      uint32_t n1 = mt_twist(a1, b1, c1);
      n2 = mt_twist(a2, b2, n1);
      mt[tid] = n1;
      mt[227 + tid] = n2;
    }
    if (tid < 169) mt[454 + tid] = mt_twist(a3, b3, n2);
    __syncthreads();
    if (tid == 169) mt[623] = mt_twist(a3, mt[0], n2);  // old[623], new[0], new[396]
    __syncthreads();
    uint32_t* o = out + (size_t)g * MT_N;
    for (int i = tid; i < MT_N; i += blockDim.x) o[i] = mt_temper(mt[i]);
  }
  __syncthreads();
  for (int i = tid; i < MT_N; i += blockDim.x) state[i] = mt[i];
}

// ---------------------------------------------------------------------------------------
// K5 walk: one full shuffle of n elements starting at stream offset o.  Returns the end
// offset and the elements that end up in positions 0..3.
// ---------------------------------------------------------------------------------------
template <bool PAIRS>
__global__ void __launch_bounds__(64)
replay_walk_kernel(const uint32_t* __restrict__ X, uint32_t n, uint32_t steps, const uint32_t* __restrict__ thr,
                   const WinEntry* __restrict__ win, const unsigned long long* __restrict__ base_ptr,
                   uint32_t* __restrict__ cand_end, int4* __restrict__ cand_samp, unsigned long long stream_len,
                   int* __restrict__ status) {
  const int g = blockIdx.y;
  const WinEntry we = win[g];
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= we.width) return;
  const unsigned long long base = *base_ptr;
  const unsigned long long start = base + (unsigned long long)g * steps + we.lo + j;
  // a walk never consumes more than `steps` + (drift within one shuffle); the host sized the
  // stream with a large margin — refuse rather than read past the end
  if (start + 2ull * steps + 64ull >= stream_len) {
    atomicOr(status, 2);
    cand_end[we.first + j] = 0xffffffffu;
    return;
  }
  int a[4];
  // offsets relative to the chunk base fit 32 bits by construction (chunk span << 2^32)
  const uint32_t rel0 = (uint32_t)(start - base);
  uint32_t end = walk_shuffle<PAIRS>(X + base, rel0, n, steps, thr, a);
  cand_end[we.first + j] = end;  // relative to chunk base
  cand_samp[we.first + j] = make_int4(a[0], a[1], a[2], a[3]);
}

// chain the chunk: from the exact base offset, follow each iteration's true candidate
__global__ void __launch_bounds__(1024)
replay_chain_kernel(const WinEntry* __restrict__ win, int G, uint32_t steps, const uint32_t* __restrict__ cand_end,
                    const int4* __restrict__ cand_samp, int n_cand, unsigned long long* __restrict__ base_ptr,
                    int4* __restrict__ samples /* for this chunk */, int* __restrict__ status) {
  extern __shared__ uint32_t s_end[];
  const bool in_smem = n_cand * sizeof(uint32_t) <= 200 * 1024;
  if (in_smem) {
    for (int i = threadIdx.x; i < n_cand; i += blockDim.x) s_end[i] = cand_end[i];
  }
  __shared__ int s_pick[1024];
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t rel = 0;  // offset relative to chunk base
    bool bad = (*status != 0);
    for (int g = 0; g < G; g++) {
      int pick = -1;
      if (!bad) {
        const WinEntry we = win[g];
        long long j = (long long)rel - ((long long)g * steps + we.lo);
        if (j < 0 || j >= (long long)we.width) {
          bad = true;
          atomicOr(status, 1);
        } else {
          pick = (int)(we.first + (uint32_t)j);
          uint32_t e = in_smem ? s_end[pick] : cand_end[pick];
          if (e == 0xffffffffu) { bad = true; pick = -1; } else rel = e;
        }
      }
      if (g < 1024) s_pick[g] = pick; else if (pick >= 0) samples[g] = cand_samp[pick];
    }
    if (!bad) *base_ptr += rel;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G && g < 1024; g += blockDim.x) {
    int pick = s_pick[g];
    samples[g] = pick >= 0 ? cand_samp[pick] : make_int4(-1, -1, -1, -1);
  }
}

// ---------------------------------------------------------------------------------------
// K6 / K7
// ---------------------------------------------------------------------------------------
__global__ void build_points_kernel(const int32_t* __restrict__ kp1, const int32_t* __restrict__ kp2,
                                    const pano_dmatch* __restrict__ m, int n, float4* __restrict__ pts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pano_dmatch mm = m[i];
  pts[i] = make_float4((float)kp1[2 * mm.query_idx], (float)kp1[2 * mm.query_idx + 1],
                       (float)kp2[2 * mm.train_idx], (float)kp2[2 * mm.train_idx + 1]);
}

__global__ void __launch_bounds__(32)
dlt_kernel(const float4* __restrict__ pts, const int4* __restrict__ samples, int iters, double* __restrict__ Hs,
           int* __restrict__ valid) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= iters) return;
  int4 s = samples[t];
  int idx[4] = {s.x, s.y, s.z, s.w};
  if (s.x < 0 || s.y < 0 || s.z < 0 || s.w < 0) {  // replay did not resolve this iteration
    valid[t] = 0;
    for (int i = 0; i < 9; i++) Hs[(size_t)t * 9 + i] = 0.0;
    return;
  }
  float src[8], dst[8];
  for (int j = 0; j < 4; j++) {
    float4 p = pts[idx[j]];
    src[2 * j] = p.x; src[2 * j + 1] = p.y;
    dst[2 * j] = p.z; dst[2 * j + 1] = p.w;
  }
  double LtL[81], V[81], H[9];
  int ok = find_homography4(src, dst, H, LtL, V);
  valid[t] = ok;
  for (int i = 0; i < 9; i++) Hs[(size_t)t * 9 + i] = ok ? H[i] : 0.0;
}

__global__ void __launch_bounds__(256)
score_kernel(const float4* __restrict__ pts, int m, const double* __restrict__ Hs, const int* __restrict__ valid,
             double thr, int* __restrict__ counts) {
  const int t = blockIdx.x;
  if (!valid[t]) {
    if (threadIdx.x == 0) counts[t] = -1;
    return;
  }
  __shared__ double sH[9];
  __shared__ int wc[8];
  if (threadIdx.x < 9) sH[threadIdx.x] = Hs[(size_t)t * 9 + threadIdx.x];
  __syncthreads();
  double H[9];
#pragma unroll
  for (int i = 0; i < 9; i++) H[i] = sH[i];
  int c = 0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    float4 p = pts[i];
    c += is_inlier(H, p.x, p.y, p.z, p.w, thr) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wc[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < 8; i++) s += wc[i];
    counts[t] = s;
  }
}

struct SelectOut {
  double H[9];
  int best_count, best_iter, status, pad;
};

// first iteration with the strictly largest positive count (ref :295-298, bestInlierCount = 0)
__global__ void __launch_bounds__(1024)
select_kernel(const int* __restrict__ counts, int iters, const double* __restrict__ Hs, SelectOut* __restrict__ out) {
  __shared__ unsigned long long wbest[32];
  unsigned long long best = 0;  // key = count << 32 | (0xffffffff - iter): max picks lowest iter
  for (int t = threadIdx.x; t < iters; t += blockDim.x) {
    int c = counts[t];
    if (c > 0) {
      unsigned long long key = ((unsigned long long)(uint32_t)c << 32) | (0xffffffffu - (uint32_t)t);
      best = key > best ? key : best;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long b2 = __shfl_xor_sync(0xffffffffu, best, o);
    best = b2 > best ? b2 : best;
  }
  if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) best = wbest[i] > best ? wbest[i] : best;
    if (best == 0) {
      out->best_count = 0;
      out->best_iter = -1;
      out->status = PANO_ERR_NO_HOMOGRAPHY;
      for (int i = 0; i < 9; i++) out->H[i] = 0;
    } else {
      int it = (int)(0xffffffffu - (uint32_t)best);
      out->best_count = (int)(best >> 32);
      out->best_iter = it;
      out->status = PANO_OK;
      for (int i = 0; i < 9; i++) out->H[i] = Hs[(size_t)it * 9 + i];
    }
  }
}

__global__ void inlier_mask_kernel(const float4* __restrict__ pts, int m, const SelectOut* __restrict__ sel,
                                   double thr, uint8_t* __restrict__ mask) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  if (sel->status != PANO_OK) { mask[i] = 0; return; }
  double H[9];
#pragma unroll
  for (int k = 0; k < 9; k++) H[k] = sel->H[k];
  float4 p = pts[i];
  mask[i] = is_inlier(H, p.x, p.y, p.z, p.w, thr) ? 1 : 0;
}

}  // namespace

void mt_ensure(cudaStream_t st, MtStream& mt, uint32_t seed, uint64_t need, uint64_t guard) {
  const bool fresh = !mt.valid || mt.seed != seed;
  if (!fresh && mt.len >= need && mt.guard >= guard) return;
  if (!fresh && mt.len >= need) need = mt.len;  // only the guard has to grow
  uint64_t gens_total = (need + MT_N - 1) / MT_N + 1;
  uint64_t new_len = gens_total * MT_N;
  if (fresh) {
    mt.state.reserve(sizeof(uint32_t) * MT_N);
    mt.x.reserve(sizeof(uint32_t) * (new_len + guard));
    mt.len = 0;
    mt.seed = seed;
    mt.valid = true;
  } else if (sizeof(uint32_t) * (new_len + guard) > mt.x.cap) {
    // grow geometrically, keeping what was generated
    DevBuf bigger;
    bigger.reserve(sizeof(uint32_t) * (std::max<uint64_t>(new_len, 2 * mt.len) + guard));
    PANO_CUDA(cudaMemcpyAsync(bigger.p, mt.x.p, sizeof(uint32_t) * mt.len, cudaMemcpyDeviceToDevice, st));
    PANO_CUDA(cudaStreamSynchronize(st));
    mt.x.release();
    mt.x = bigger;
  }
  uint64_t gens = (new_len - mt.len) / MT_N;
  uint64_t done = 0;
  while (done < gens) {
    int g = (int)std::min<uint64_t>(gens - done, 1u << 20);
    mt_generate_kernel<<<1, 256, 0, st>>>(mt.state.as<uint32_t>(), (fresh && done == 0 && mt.len == 0) ? 1 : 0, seed,
                                          mt.x.as<uint32_t>() + mt.len, g);
    PANO_LAUNCH_CHECK();
    mt.len += (uint64_t)g * MT_N;
    done += g;
  }
  // guard region: 0xFFFFFFFF is never rejected by Lemire's test (its low word is 2^32 - r,
  // which is >= 2^32 mod r), so a speculative walk that runs off the generated stream
  // terminates after at most `steps` more reads instead of spinning on stale memory
  PANO_CUDA(cudaMemsetAsync(mt.x.as<uint32_t>() + mt.len, 0xff, sizeof(uint32_t) * guard, st));
  mt.guard = guard;
}

RansacResult ransac_device(cudaStream_t st, const int32_t* kp1_dev, const int32_t* kp2_dev,
                           const pano_dmatch* matches_dev, int m, const pano_ransac_opts& o, uint32_t seed,
                           MtStream& mt, RansacScratch& s, PinnedBuf& pin, int32_t* samples_out_host,
                           int32_t* counts_out_host, uint8_t* mask_out_host, int window_scale) {
  RansacResult res;
  memset(&res, 0, sizeof res);
  res.best_iter = -1;
  const int iters = o.num_iterations;
  if (m < o.num_samples || iters <= 0) {  // ref :268-269: the loop breaks at once -> empty H
    res.status = PANO_ERR_TOO_FEW_MATCHES;
    return res;
  }
  const uint32_t n = (uint32_t)m;
  const bool pairs = shuffle_uses_pairs(n);
  const uint32_t steps = shuffle_steps(n);

  // ---- per-step Lemire thresholds, rejection statistics, chunk/window plan (host, O(steps)
  //      integer work: launch-parameter planning, like the Gaussian taps) ------------------
  ReplayPlan plan = plan_replay(n, iters, window_scale);
  const std::vector<uint32_t>& thr = plan.thr;
  const std::vector<WinEntry>& win = plan.win;
  const int G = plan.G;
  const uint32_t n_cand = plan.n_cand, max_w = plan.max_w;
  const int n_chunks = (iters + G - 1) / G;

  // ---- stream: everything the walks can touch, with margin ---------------------------------
  uint64_t need = plan.stream_need;
  mt_ensure(st, mt, seed, need, (uint64_t)steps + 4096);

  // ---- buffers ---------------------------------------------------------------------------
  s.thr.reserve(sizeof(uint32_t) * thr.size());
  s.plan.reserve(sizeof(WinEntry) * win.size());
  s.cand_off.reserve(sizeof(uint32_t) * (size_t)n_cand);
  s.cand_samp.reserve(sizeof(int4) * (size_t)n_cand);
  s.base.reserve(sizeof(unsigned long long) + 2 * sizeof(int));
  s.samples.reserve(sizeof(int4) * (size_t)(n_chunks * G));
  s.pts.reserve(sizeof(float4) * (size_t)m);
  s.Hs.reserve(sizeof(double) * 9 * (size_t)iters);
  s.valid.reserve(sizeof(int) * (size_t)iters);
  s.counts.reserve(sizeof(int) * (size_t)iters);
  s.result.reserve(sizeof(SelectOut));
  s.mask.reserve((size_t)m);
  pin.reserve(sizeof(SelectOut) + 64);

  PANO_CUDA(cudaMemcpyAsync(s.thr.p, thr.data(), sizeof(uint32_t) * thr.size(), cudaMemcpyHostToDevice, st));
  PANO_CUDA(cudaMemcpyAsync(s.plan.p, win.data(), sizeof(WinEntry) * win.size(), cudaMemcpyHostToDevice, st));
  PANO_CUDA(cudaMemsetAsync(s.base.p, 0, sizeof(unsigned long long) + 2 * sizeof(int), st));
  unsigned long long* base_ptr = s.base.as<unsigned long long>();
  int* status_ptr = reinterpret_cast<int*>(base_ptr + 1);

  build_points_kernel<<<(m + 255) / 256, 256, 0, st>>>(kp1_dev, kp2_dev, matches_dev, m, s.pts.as<float4>());
  PANO_LAUNCH_CHECK();

  size_t chain_smem = (size_t)n_cand * sizeof(uint32_t) <= 200 * 1024 ? (size_t)n_cand * sizeof(uint32_t) : 0;
  PANO_CUDA(cudaFuncSetAttribute(replay_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int c = 0; c < n_chunks; c++) {
    int Gc = std::min(G, iters - c * G);
    dim3 grid((max_w + 63) / 64, Gc);
    if (pairs)
      replay_walk_kernel<true><<<grid, 64, 0, st>>>(mt.x.as<uint32_t>(), n, steps, s.thr.as<uint32_t>(),
                                                   s.plan.as<WinEntry>(), base_ptr, s.cand_off.as<uint32_t>(),
                                                   s.cand_samp.as<int4>(), mt.len, status_ptr);
    else
      replay_walk_kernel<false><<<grid, 64, 0, st>>>(mt.x.as<uint32_t>(), n, steps, s.thr.as<uint32_t>(),
                                                    s.plan.as<WinEntry>(), base_ptr, s.cand_off.as<uint32_t>(),
                                                    s.cand_samp.as<int4>(), mt.len, status_ptr);
    PANO_LAUNCH_CHECK();
    replay_chain_kernel<<<1, 1024, chain_smem, st>>>(s.plan.as<WinEntry>(), Gc, steps, s.cand_off.as<uint32_t>(),
                                                    s.cand_samp.as<int4>(), (int)n_cand, base_ptr,
                                                    s.samples.as<int4>() + (size_t)c * G, status_ptr);
    PANO_LAUNCH_CHECK();
  }

  dlt_kernel<<<(iters + 31) / 32, 32, 0, st>>>(s.pts.as<float4>(), s.samples.as<int4>(), iters, s.Hs.as<double>(),
                                               s.valid.as<int>());
  PANO_LAUNCH_CHECK();
  score_kernel<<<iters, 256, 0, st>>>(s.pts.as<float4>(), m, s.Hs.as<double>(), s.valid.as<int>(),
                                      o.distance_threshold, s.counts.as<int>());
  PANO_LAUNCH_CHECK();
  select_kernel<<<1, 1024, 0, st>>>(s.counts.as<int>(), iters, s.Hs.as<double>(), s.result.as<SelectOut>());
  PANO_LAUNCH_CHECK();
  if (mask_out_host) {
    inlier_mask_kernel<<<(m + 255) / 256, 256, 0, st>>>(s.pts.as<float4>(), m, s.result.as<SelectOut>(),
                                                       o.distance_threshold, s.mask.as<uint8_t>());
    PANO_LAUNCH_CHECK();
  }
  char* pp = pin.as<char>();
  PANO_CUDA(cudaMemcpyAsync(pp, s.result.p, sizeof(SelectOut), cudaMemcpyDeviceToHost, st));
  PANO_CUDA(cudaMemcpyAsync(pp + sizeof(SelectOut), status_ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
  PANO_CUDA(cudaStreamSynchronize(st));
  SelectOut so;
  memcpy(&so, pp, sizeof so);
  int replay_status;
  memcpy(&replay_status, pp + sizeof(SelectOut), sizeof(int));
  if (replay_status != 0) {
    // a speculation window was missed (bit 0) or the stream margin was too small (bit 1):
    // nothing was guessed; tell the caller to re-run with wider windows
    res.status = -replay_status;
    return res;
  }
  res.status = so.status;
  memcpy(res.H, so.H, sizeof so.H);
  res.best_count = so.best_count;
  res.best_iter = so.best_iter;
  if (samples_out_host)
    PANO_CUDA(cudaMemcpyAsync(samples_out_host, s.samples.p, sizeof(int4) * (size_t)iters, cudaMemcpyDeviceToHost, st));
  if (counts_out_host)
    PANO_CUDA(cudaMemcpyAsync(counts_out_host, s.counts.p, sizeof(int) * (size_t)iters, cudaMemcpyDeviceToHost, st));
  if (mask_out_host)
    PANO_CUDA(cudaMemcpyAsync(mask_out_host, s.mask.p, (size_t)m, cudaMemcpyDeviceToHost, st));
  if (samples_out_host || counts_out_host || mask_out_host) PANO_CUDA(cudaStreamSynchronize(st));
  return res;
}

}  // namespace pano
