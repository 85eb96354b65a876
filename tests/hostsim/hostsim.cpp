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

double hs_inlier_d2_limit(double thr) { return inlier_d2_limit(thr); }

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

// emulates replay_resident_kernel phase by phase (cells -> segment walks -> chain -> tracking).
// returns 0 ok, 1 band miss, 2 stream too short, 32 plan does not fit.  stats: nwords, dmax, nseg, n_entries, smem
int hs_replay_resident(uint32_t seed, uint32_t n, int iters, int window_scale, double z_sigma, int32_t* samples,
                       uint64_t* end_offset, double* stats) {
  ReplayPlan P = plan_replay(n, iters, window_scale);
  ResidentPlan R = plan_resident(P, n, window_scale, z_sigma > 0 ? z_sigma : 4.5);
  if (stats) { stats[0] = R.nwords; stats[1] = R.dmax; stats[2] = R.nseg; stats[3] = R.n_entries; stats[4] = (double)R.smem_bytes; }
  if (!R.ok) return 32;
  const uint32_t steps = R.steps, odd = n & 1u;
  const uint64_t x_limit = P.stream_need + steps + 4096;
  std::vector<uint32_t> X(x_limit, 0xffffffffu);
  {
    std::mt19937 e(seed);
    for (uint64_t i = 0; i < P.stream_need; i++) X[i] = e();
  }
  std::vector<uint32_t> bits(R.nwords);
  std::vector<uint8_t> dtab((size_t)R.n_entries * (R.segb + 1));
  std::vector<uint32_t> seg_entry(R.nseg);
  uint64_t s = 0;
  for (int t = 0; t < iters; t++) {
    if (s + R.xcap + steps > x_limit) return 2;
    const uint32_t* Xs = X.data() + s;   // (the kernel's shared-memory window, rel = 0)
    // phase 1: cells
    for (uint32_t b = 0; b < R.nkb; b++) {
      const ResBlock B = R.blk[b];
      for (uint32_t j = 0; j < B.w; j++) {
        uint32_t word = 0;
        for (uint32_t lane = 0; lane < 32; lane++) {
          const uint32_t k = b * 32u + lane;
          const RT q = k < steps ? P.rt[k] : RT{2u, 0u};
          if (Xs[k + B.dlo + j] * q.r < q.T) word |= 1u << lane;
        }
        bits[B.woff + j] = word;
      }
    }
    // phase 2: one walk per (segment, entry diagonal)
    for (uint32_t sgm = 0; sgm < R.nseg; sgm++) {
      const uint32_t b0 = sgm * R.segb, b1 = std::min(R.nkb, b0 + R.segb);
      for (uint32_t e = 0; e < R.blk[b0].w; e++) {
        uint8_t* out = dtab.data() + (size_t)(R.seg_eoff[sgm] + e) * (R.segb + 1);
        out[R.segb] = (uint8_t)res_walk_segment(bits.data(), R.blk.data(), b0, b1, R.blk[b0].dlo + e, out);
      }
    }
    // chain
    uint32_t d = 0;
    for (uint32_t sgm = 0; sgm < R.nseg; sgm++) {
      const ResBlock B0 = R.blk[sgm * R.segb];
      const uint32_t e = d - B0.dlo;
      if (e >= B0.w) return 1;
      seg_entry[sgm] = R.seg_eoff[sgm] + e;
      d = dtab[(size_t)seg_entry[sgm] * (R.segb + 1) + R.segb];
      if (d == RES_MISS) return 1;
    }
    // tracking: exact first steps, then "largest element written to position p wins"
    int a[4] = {0, 1, 2, 3};
    {
      uint32_t o = 0, k = 0;
      while (k < steps && 2u * k + odd < 4u) { track_step<true>(Xs, o, k, n, P.rt.data(), a); k++; }
    }
    int slot[4] = {-1, -1, -1, -1};
    for (uint32_t k = 0; k < steps; k++) {
      const uint32_t idx = 2u * k + odd;
      if (idx < 4u) continue;
      const uint32_t b = k >> 5, sgm = b / R.segb;
      const uint32_t d0 = dtab[(size_t)seg_entry[sgm] * (R.segb + 1) + (b - sgm * R.segb)];
      const uint32_t dk = res_diag_at(bits.data(), R.blk[b], d0, k & 31u);
      const uint32_t x = Xs[k + dk];
      const unsigned long long u = (unsigned long long)x * (idx + 1u);
      const unsigned long long v = (unsigned long long)(uint32_t)u * (idx + 2u);
      const uint32_t p1 = (uint32_t)(u >> 32), p2 = (uint32_t)(v >> 32);
      if (p1 < 4u) slot[p1] = std::max(slot[p1], (int)idx);
      if (p2 < 4u) slot[p2] = std::max(slot[p2], (int)idx + 1);
    }
    for (int p = 0; p < 4; p++) samples[(size_t)t * 4 + p] = slot[p] >= 0 ? slot[p] : a[p];
    s += steps + d;
  }
  if (end_offset) *end_offset = s;
  return 0;
}

// warp_coord_fast (the warp kernel's fast coordinate path) against the exact warp_coord over a whole canvas.
// seed_mode: 0 = reciprocal truncated to 20 mantissa bits (what MUFU.RCP64H delivers), 1 = 2^-12 relative error,
// 2 = a useless seed (half the reciprocal), 3 = the correctly rounded reciprocal.
// out[0] = pixels whose accepted fast result differs from the exact one (must be 0), out[1] = pixels sent to the
// exact path, out[2] = pixels checked.
void hs_warp_fast_check(const double* Minv, int cw, int ch, int step, int seed_mode, uint64_t* out, int variant) {
  out[0] = out[1] = out[2] = 0;
  double Mx[3] = {32.0 * Minv[0], 32.0 * Minv[1], 32.0 * Minv[2]};
  double My[3] = {32.0 * Minv[3], 32.0 * Minv[4], 32.0 * Minv[5]};
  for (int y = 0; y < ch; y += step)
    for (int xb = 0; xb < cw; xb += 64) {
      const double xbd = (double)xb, yd = (double)y;
      const double X0 = (Mx[0] * xbd + Mx[1] * yd) + Mx[2];
      const double Y0 = (My[0] * xbd + My[1] * yd) + My[2];
      const double W0 = (Minv[6] * xbd + Minv[7] * yd) + Minv[8];
      for (int x1 = 0; x1 < 64 && xb + x1 < cw; x1++) {
        const double Wd = W0 + Minv[6] * (double)x1;
        double r0 = 1.0 / Wd;
        if (seed_mode == 0) { uint64_t b; memcpy(&b, &r0, 8); b &= ~0xffffffffull; memcpy(&r0, &b, 8); }
        else if (seed_mode == 1) r0 *= 1.000244140625;
        else if (seed_mode == 2) r0 *= 0.5;
        int Xf, Yf, Xe, Ye;
        bool need;
        if (variant == 2) warp_coord_fast2(X0, Y0, W0, Mx[0], My[0], Minv[6], (double)x1, r0, &Xf, &Yf, &need);
        else warp_coord_fast(X0, Y0, W0, Mx[0], My[0], Minv[6], (double)x1, r0, &Xf, &Yf, &need);
        warp_coord(Minv, xb + x1, y, 64, &Xe, &Ye);
        out[2]++;
        if (need) out[1]++;
        else if (Xf != Xe || Yf != Ye) out[0]++;
      }
    }
}

}  // extern "C"
