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


// ---------------------------------------------------------------------------------------
// Resident replay (replay_resident_kernel): one CTA replays the iterations strictly in order,
// so every iteration starts from its EXACT stream offset and the only speculation left is on
// the number of rejections so far inside the iteration ("diagonal" d = rejections before the
// current step).  Per 32-step block b only the diagonals [dlo, dlo + w) around the expected
// count are evaluated (mean -+ z sigma of the rejections before / through the block, from the
// exact per-step probabilities); a path that leaves its band is DETECTED (status bit 0) and the
// caller re-plans with doubled margins, exactly like a window miss of the chunked replay.
// ---------------------------------------------------------------------------------------
struct alignas(8) ResBlock {
  uint16_t dlo;   // first diagonal evaluated for this block
  uint16_t w;     // number of diagonals
  uint16_t woff;  // index of (diagonal dlo) in the iteration's word array: word(b, d) = bits[woff + d - dlo]
  uint8_t seg;    // walk segment of this block
  uint8_t boff;   // index of the block inside its segment
};

struct ResidentPlan {
  bool ok = false;                // fits the kernel's shared-memory budget
  uint32_t steps = 0, nkb = 0;    // Lemire steps per shuffle, 32-step blocks
  uint32_t nwords = 0;            // bitmap words per iteration
  uint32_t dmax = 0;              // diagonals are < dmax (max rejections per iteration the band admits + 1)
  uint32_t segb = 0, nseg = 0;    // blocks per walk segment, segments
  uint32_t n_entries = 0;         // sum over segments of the entry-band widths
  uint32_t xcap = 0;              // stream words per shared-memory window
  std::vector<ResBlock> blk;      // nkb entries
  std::vector<uint32_t> seg_eoff; // nseg + 1: first entry index of each segment
  size_t smem_bytes = 0;
};

constexpr uint32_t RES_MAX_DIAG = 250;        // diagonals are stored in bytes (255 = miss)
constexpr uint32_t RES_WARPS = 16;            // warps of the resident CTA (512 threads: half an SM's registers, so it
                                              // can be placed next to other kernels' CTAs instead of waiting for an idle SM)
constexpr uint32_t RES_MAXB = 12;             // 32-step blocks per warp of the kernel: steps <= 32 * RES_WARPS * RES_MAXB
constexpr size_t RES_SMEM_BUDGET = 200 * 1024;


constexpr uint32_t RES_MISS = 255u;

// Walk of blocks [b0, b1) over the evaluated rejection bits, entered on diagonal d: along a diagonal until
// the next set bit (a rejection at that step), then the same step again on the next diagonal.  dout[i] =
// diagonal on entering block b0 + i.  Returns the exit diagonal, or RES_MISS if the path leaves the band.
// The words of RES_LOOK consecutive blocks on the current diagonal are fetched together (most blocks hold no
// rejection on a given diagonal), so the dependent-load chain is per batch, not per block.
constexpr int RES_LOOK = 4;
PANO_HD uint32_t res_walk_segment(const uint32_t* bits, const ResBlock* blk, uint32_t b0, uint32_t b1, uint32_t d,
                                  uint8_t* dout) {
  uint32_t b = b0, mask = ~0u;
  bool fresh = true;   // block b has not been entered yet (its dout is still to be written)
  while (b < b1) {
    uint32_t wv[RES_LOOK];
    bool inb[RES_LOOK];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < RES_LOOK; i++) {
      const uint32_t bb = b + (uint32_t)i < b1 ? b + (uint32_t)i : b1 - 1u;
      const ResBlock B = blk[bb];
      const uint32_t r = d - (uint32_t)B.dlo;
      inb[i] = r < (uint32_t)B.w;
      wv[i] = inb[i] ? bits[B.woff + r] : 0u;
    }
    wv[0] &= mask;
    int stop = RES_LOOK;   // first block of the batch that ends the straight run along diagonal d
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = RES_LOOK - 1; i >= 0; i--)
      if (b + (uint32_t)i < b1 && (!inb[i] || wv[i] != 0u)) stop = i;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < RES_LOOK; i++)
      if (i <= stop && b + (uint32_t)i < b1 && (i > 0 || fresh)) dout[b + (uint32_t)i - b0] = (uint8_t)d;
    if (stop == RES_LOOK) {
      b += RES_LOOK;
      fresh = true;
      mask = ~0u;
      continue;
    }
    uint32_t hit = 0u;
    bool in_band = true;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < RES_LOOK; i++)
      if (i == stop) { hit = wv[i]; in_band = inb[i]; }
    if (!in_band) return RES_MISS;
    b += (uint32_t)stop;
    fresh = false;
    d++;
    {
      const ResBlock B = blk[b];
      if (d - (uint32_t)B.dlo >= (uint32_t)B.w) return RES_MISS;
    }
    mask = ~0u << ctz32(hit);
  }
  return d;
}

// Rejections at steps <= (block start + kbit) for a path that entered the block on diagonal d (a path the
// segment walk has already validated): the accepted draw of that step is word (start + step + result).
PANO_HD uint32_t res_diag_at(const uint32_t* bits, const ResBlock B, uint32_t d, uint32_t kbit) {
  const uint32_t upto = kbit == 31u ? ~0u : ((2u << kbit) - 1u);
  uint32_t mask = upto;
  for (;;) {
    const uint32_t w = bits[B.woff + d - B.dlo] & mask;
    if (!w) break;
    d++;
    mask = (~0u << ctz32(w)) & upto;
  }
  return d;
}

inline size_t resident_smem_bytes(const ResidentPlan& R) {
  size_t b = 0;
  b += sizeof(uint32_t) * 2 * (size_t)R.xcap;                 // stream windows (double buffered)
  b += sizeof(uint32_t) * (size_t)R.nwords;                   // rejection bitmap
  b += (sizeof(ResBlock) + 16) * (size_t)R.nkb;               // band table + phase-1 view of it
  b += (size_t)R.n_entries * (R.segb + 1);                    // per entry: diagonal at each block + exit
  b += sizeof(uint32_t) * 4 * (size_t)(R.nseg + 1);           // per segment: entry base, band, chosen entry
  return b + 256;
}

inline ResidentPlan plan_resident(const ReplayPlan& P, uint32_t n, int window_scale, double z_sigma = 4.5) {
  ResidentPlan R;
  const uint32_t steps = shuffle_steps(n);
  R.steps = steps;
  if (!shuffle_uses_pairs(n) || steps == 0) return R;   // the single-draw regime (n > 65535) stays on the chunked path
  const uint32_t nkb = (steps + 31u) / 32u;
  R.nkb = nkb;
  // cumulative mean / variance of the rejections before step k
  std::vector<double> cm((size_t)steps + 1, 0.0), cv((size_t)steps + 1, 0.0);
  for (uint32_t k = 0; k < steps; k++) {
    const double p = (double)P.rt[k].T / 4294967296.0;
    cm[k + 1] = cm[k] + p / (1.0 - p);
    cv[k + 1] = cv[k] + p / ((1.0 - p) * (1.0 - p));
  }
  const double z = z_sigma * window_scale, pad = 1.0 * window_scale;
  R.blk.resize(nkb);
  uint32_t prev_lo = 0, prev_hi = 0, woff = 0;
  for (uint32_t b = 0; b < nkb; b++) {
    const uint32_t k0 = b * 32u, k1 = std::min(steps, k0 + 32u);
    double lo = std::floor(cm[k0] - z * std::sqrt(cv[k0]) - pad);
    double hi = std::ceil(cm[k1] + z * std::sqrt(cv[k1]) + pad);
    if (lo < 0) lo = 0;
    uint32_t ulo = std::max(prev_lo, (uint32_t)lo), uhi = std::max(prev_hi, (uint32_t)hi);
    if (b == 0) ulo = 0;
    uhi = ulo + ((uhi - ulo + 1u + 3u) & ~3u) - 1u;   // width a multiple of 4: phase 1 evaluates 4 diagonals per trip
    if (uhi >= RES_MAX_DIAG) return R;
    R.blk[b].dlo = (uint16_t)ulo;
    R.blk[b].w = (uint16_t)(uhi - ulo + 1u);
    if (woff > 0xffffu) return R;
    R.blk[b].woff = (uint16_t)woff;
    woff += R.blk[b].w;
    prev_lo = ulo; prev_hi = uhi;
  }
  R.nwords = woff;
  R.dmax = prev_hi + 1u;
  // walk segments: about 16 per iteration (the chain over segments is sequential, the walks of a
  // segment's entries are parallel)
  R.segb = std::max(1u, (nkb + 15u) / 16u);
  R.nseg = (nkb + R.segb - 1u) / R.segb;
  for (uint32_t b = 0; b < nkb; b++) { R.blk[b].seg = (uint8_t)(b / R.segb); R.blk[b].boff = (uint8_t)(b % R.segb); }
  R.seg_eoff.assign((size_t)R.nseg + 1, 0u);
  for (uint32_t sgm = 0; sgm < R.nseg; sgm++) R.seg_eoff[sgm + 1] = R.seg_eoff[sgm] + R.blk[sgm * R.segb].w;
  R.n_entries = R.seg_eoff[R.nseg];
  // window of iteration t+1 is prefetched before iteration t's rejections are known:
  // [s + steps, s + steps + dmax) are its possible starts, each needs steps + dmax + 32 words
  R.xcap = (steps + 2u * R.dmax + 64u + 3u) / 4u * 4u + 8u;
  R.smem_bytes = resident_smem_bytes(R);
  R.ok = R.smem_bytes <= RES_SMEM_BUDGET && R.n_entries <= 32u * RES_WARPS && nkb <= RES_WARPS * RES_MAXB && R.segb <= 255u;
  return R;
}

}  // namespace pano
