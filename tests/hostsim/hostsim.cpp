// hostsim.cpp — CPU emulation of the engine's kernels' LOGIC, for the no-GPU test tier.
//
// TEST INFRASTRUCTURE ONLY.  It compiles the very headers the CUDA kernels use
// (csrc/pano_core.cuh, csrc/replay_plan.hpp) with g++ and re-runs the kernels' control flow
// (grid of candidate walks + chain for the shuffle replay, per-hypothesis DLT, per-pixel warp)
// in plain loops, so that the order-exact arithmetic and the speculation scheme can be checked
// against the oracle without a GPU.  It is not linked into, or reachable from, the product.
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/pano_core.cuh"
#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/replay_plan.hpp"

using namespace pano;

extern "C" {

int hs_find_homography4(const float* src, const float* dst, double* H) {
  double LtL[81], V[81];
  return find_homography4(src, dst, H, LtL, V);
}

int hs_is_inlier(const double* H, float x, float y, float qx, float qy, double thr) {
  return is_inlier(H, x, y, qx, qy, thr) ? 1 : 0;
}

// out: cw, ch, offx, offy, bw0, ok ; TH[9]; Minv[9]
void hs_canvas_geometry(int wl, int hl, int wr, int hr, const double* H, int* geom, double* TH, double* Minv) {
  CanvasGeom g;
  canvas_geometry(wl, hl, wr, hr, H, &g);
  geom[0] = g.cw; geom[1] = g.ch; geom[2] = g.offx; geom[3] = g.offy; geom[4] = g.bw0; geom[5] = g.ok;
  memcpy(TH, g.TH, sizeof g.TH);
  memcpy(Minv, g.Minv, sizeof g.Minv);
}

// emulates warp_overlay_kernel<false>: cv::warpPerspective(src, M, (dw, dh))
void hs_warp_perspective(const uint8_t* src, int w, int h, size_t stride, const double* M, uint8_t* dst, int dw,
                         int dh, size_t dstride) {
  double Minv[9];
  invert33(M, Minv);
  int bh0 = dh < 16 ? dh : 16;
  int bw0 = 1024 / bh0;
  if (bw0 > dw) bw0 = dw;
  for (int y = 0; y < dh; y++)
    for (int x = 0; x < dw; x++) {
      int X, Y;
      warp_coord(Minv, x, y, bw0, &X, &Y);
      uint32_t v = warp_pixel(src, stride, w, h, X, Y);
      uint8_t* o = dst + (size_t)y * dstride + 3 * (size_t)x;
      o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)(v >> 16);
    }
}

// emulates replay_offsets_kernel + replay_chain_kernel chunk by chunk, then replay_samples_kernel.
// returns 0 ok, bit 0 window miss, bit 1 stream too short, bit 2 pass-1/pass-2 disagreement.
// stats: G, n_cand, max_w, chunks, mu, sigma
int hs_replay(uint32_t seed, uint32_t n, int iters, int window_scale, int32_t* samples, uint64_t* end_offset,
              double* stats) {
  ReplayPlan P = plan_replay(n, iters, window_scale);
  const uint32_t steps = shuffle_steps(n);
  const bool pairs = shuffle_uses_pairs(n);
  uint64_t guard = steps + 4096;
  std::vector<uint32_t> X(P.stream_need + guard, 0xffffffffu);
  {
    std::mt19937 e(seed);
    for (uint64_t i = 0; i < P.stream_need; i++) X[i] = e();
  }
  if (stats) { stats[0] = P.G; stats[1] = P.n_cand; stats[2] = P.max_w; stats[3] = (iters + P.G - 1) / P.G; stats[4] = P.mu; stats[5] = P.sigma; }
  const int nseg = (int)((steps + PANO_SEG_STEPS - 1) / PANO_SEG_STEPS);
  std::vector<uint32_t> cand_end(P.n_cand);
  std::vector<uint32_t> seg_off((size_t)P.n_cand * (size_t)std::max(nseg - 1, 1));
  std::vector<uint64_t> seg_tab((size_t)iters * nseg);
  uint64_t base = 0;
  for (int c0 = 0; c0 < iters; c0 += P.G) {
    int Gc = std::min(P.G, iters - c0);
    // replay_cells_kernel: every (diagonal, step) rejection test of the chunk, packed 32 steps per word
    const uint32_t nkb = (steps + 31u) / 32u;
    std::vector<uint32_t> bits((size_t)P.n_diag * nkb, 0u);
    for (int g = 0; g < Gc; g++) {
      const WinEntry& we = P.win[g];
      const uint32_t D = (we.width + P.dextra + 31u) / 32u * 32u;
      uint32_t* bg = bits.data() + (size_t)we.dfirst * nkb;
      const uint64_t pos0 = base + (uint64_t)g * steps + we.lo;
      for (uint32_t d = 0; d < D; d++)
        for (uint32_t k = 0; k < steps; k++) {
          uint64_t pos = pos0 + d + k;
          if (pos >= X.size()) continue;
          if (X[pos] * P.rt[k].r < P.rt[k].T) bg[(size_t)(k / 32u) * D + d] |= 1u << (k & 31u);
        }
    }
    for (int g = 0; g < Gc; g++) {          // replay_walk_bits_kernel
      const WinEntry& we = P.win[g];
      const uint32_t D = (we.width + P.dextra + 31u) / 32u * 32u;
      for (uint32_t j = 0; j < we.width; j++) {
        uint64_t start = base + (uint64_t)g * steps + we.lo + j;
        if (start + 2ull * steps + 64ull >= P.stream_need) return 2;
        uint32_t end = walk_bits(bits.data() + (size_t)we.dfirst * nkb, D, nkb, j, steps, (uint32_t)g * steps + we.lo,
                                 seg_off.data() + we.first + j, (size_t)P.n_cand);
        if (end == 0xffffffffu) return 8;
        // cross-check against the direct walk (the formulation the grid replaces)
        if (end != walk_offsets(X.data() + base, (uint32_t)(start - base), steps, P.rt.data(), nullptr, 0)) return 16;
        cand_end[we.first + j] = end;
      }
    }
    uint32_t rel = 0;                        // replay_chain_kernel
    for (int g = 0; g < Gc; g++) {
      const WinEntry& we = P.win[g];
      long long j = (long long)rel - ((long long)g * steps + we.lo);
      if (j < 0 || j >= (long long)we.width) return 1;
      uint32_t pick = we.first + (uint32_t)j;
      seg_tab[((size_t)c0 + g) * nseg] = base + rel;
      for (int sg = 1; sg < nseg; sg++) seg_tab[((size_t)c0 + g) * nseg + sg] = base + seg_off[(size_t)(sg - 1) * P.n_cand + pick];
      rel = cand_end[pick];
    }
    base += rel;
  }
  std::vector<int> seg_w((size_t)iters * nseg * 4);
  for (int i = 0; i < iters * nseg; i++) {   // replay_segments_kernel
    int sg = i % nseg;
    uint32_t k0 = (uint32_t)sg * PANO_SEG_STEPS, k1 = std::min(steps, k0 + PANO_SEG_STEPS);
    int w[4];
    uint32_t end = pairs ? walk_track_segment<true>(X.data() + seg_tab[i], 0u, n, k0, k1, P.rt.data(), w)
                         : walk_track_segment<false>(X.data() + seg_tab[i], 0u, n, k0, k1, P.rt.data(), w);
    memcpy(&seg_w[(size_t)i * 4], w, sizeof w);
    uint64_t next = i + 1 < iters * nseg ? seg_tab[i + 1] : base;
    if (seg_tab[i] + end != next) return 4;
  }
  for (int t = 0; t < iters; t++) {          // combine_samples_kernel
    int a[4] = {-1, -1, -1, -1};
    for (int sg = nseg - 1; sg >= 0; sg--)
      for (int p = 0; p < 4; p++)
        if (a[p] < 0) a[p] = seg_w[((size_t)t * nseg + sg) * 4 + p];
    memcpy(&samples[(size_t)t * 4], a, sizeof a);
  }
  if (end_offset) *end_offset = base;
  return 0;
}

}  // extern "C"
