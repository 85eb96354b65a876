// replay_plan.hpp — host-side planning of the windowed speculative shuffle replay (K5).
// Pure C++ (no CUDA) so the CPU emulation in tests/hostsim can share it with ransac.cu.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "pano_core.cuh"

namespace pano {

// per-iteration speculation window inside a chunk
struct WinEntry {
  uint32_t lo;     // window lower bound, relative to (chunk base offset + g * steps)
  uint32_t width;  // number of candidate start offsets
  uint32_t first;  // index of candidate 0 in the chunk's candidate arrays
  uint32_t dfirst; // index of diagonal 0 in the chunk's diagonal space (multiple of 32)
};

struct ReplayPlan {
  std::vector<RT> rt;          // (range, Lemire threshold 2^32 mod range) per step of one shuffle
  std::vector<WinEntry> win;   // G entries
  int G = 0;                   // iterations per chunk
  uint32_t n_cand = 0, max_w = 0;
  double mu = 0, sigma = 0;    // mean / std-dev of the rejections of one shuffle
  uint64_t stream_need = 0;    // engine outputs the replay may touch
  uint32_t dextra = 0;         // diagonals beyond a window: room for one walk's own rejections
  uint32_t n_diag = 0;         // diagonals per chunk (each window padded to a multiple of 32)
  std::vector<int> diag_block_iter;  // iteration of every 32-diagonal block
};

// n = number of shuffled elements (matches), iters = RANSAC iterations,
// window_scale = 1, 2, 4, ... (doubled by the caller after a detected window miss).
inline ReplayPlan plan_replay(uint32_t n, int iters, int window_scale, double target_cand = 50000.0,
                              double z_sigma = 4.2) {
  ReplayPlan P;
  const bool pairs = shuffle_uses_pairs(n);
  const uint32_t steps = shuffle_steps(n);
  P.rt.assign((size_t)(steps ? steps : 1) + 8, RT{2u, 0u});  // +8: vector loads past the end stay in bounds
  double mu = 0, var = 0;
  const uint32_t odd = n & 1u;
  for (uint32_t k = 0; k < steps; k++) {
    uint32_t r;
    if (pairs) {
      if (!odd && k == 0) { P.rt[k] = RT{2u, 0u}; continue; }  // d{0,1}: range 2 never rejects
      uint32_t idx = 2u * k + odd;
      r = (idx + 1u) * (idx + 2u);
    } else {
      r = k + 2u;
    }
    uint32_t T = lemire_threshold(r);
    P.rt[k] = RT{r, T};
    double p = (double)T / 4294967296.0;
    mu += p / (1.0 - p);                   // extra draws of a step are geometric
    var += p / ((1.0 - p) * (1.0 - p));
  }
  P.mu = mu;
  P.sigma = std::sqrt(var);
  // windows of +-(4.2 sigma sqrt(g) + 2) * scale around g * mu.  A true start outside its window
  // is detected and the replay re-run with doubled windows, never guessed, so the width only trades
  // speculative work against the (rare: ~2e-5 per iteration) cost of a re-run.  The chunk length is
  // chosen so that a chunk has about `target` candidate walks.
  const double zs = z_sigma * window_scale, pad = 2.0 * window_scale;
  int G = 8;
  for (int cand = 8; cand <= 1024; cand *= 2) {
    double tot = 0;
    for (int g = 0; g < cand; g++) tot += 2.0 * (zs * P.sigma * std::sqrt((double)g) + pad) + 1.0;
    if (tot <= target_cand) G = cand; else break;  // (<= 51k: the chain kernel stages end offsets in smem)
  }
  if (G > iters) G = iters;
  if (G < 1) G = 1;
  P.G = G;
  P.win.resize((size_t)G);
  double max_hi = 0;
  for (int g = 0; g < G; g++) {
    double c = g * mu, hw = (g == 0) ? 0.0 : zs * P.sigma * std::sqrt((double)g) + pad;
    double lo = std::floor(c - hw), hi = std::ceil(c + hw);
    if (lo < 0) lo = 0;
    P.win[g].lo = (uint32_t)lo;
    P.win[g].width = (uint32_t)(hi - lo) + 1u;
    P.win[g].first = P.n_cand;
    P.win[g].dfirst = 0;
    P.n_cand += P.win[g].width;
    P.max_w = std::max(P.max_w, P.win[g].width);
    max_hi = std::max(max_hi, hi);
  }
  // a walk moves up one diagonal per rejection: mu + 6 sigma (+ margin) extra diagonals; a walk
  // that would leave them is flagged and the caller re-plans wider, like a window miss
  P.dextra = (uint32_t)std::ceil(mu + (6.0 * window_scale) * P.sigma + 4.0 * window_scale);
  for (int g = 0; g < G; g++) {
    P.win[g].dfirst = P.n_diag;
    uint32_t D = (P.win[g].width + P.dextra + 31u) / 32u * 32u;
    for (uint32_t b = 0; b < D / 32u; b++) P.diag_block_iter.push_back(g);
    P.n_diag += D;
  }
  P.stream_need = (uint64_t)((double)iters * ((double)steps + mu) +
                             12.0 * window_scale * P.sigma * std::sqrt((double)iters) + max_hi + 4.0 * steps + 4096.0);
  return P;
}

}  // namespace pano
