// warp_kernels.cuh — the device code of warp.cu (general warp + overlay kernel, the round-1 fast kernel, the
// production quad kernel, the row packer) together with the host-side decisions that select and parameterise them
// (footprint box, fast-path admission).
//
// Included by warp.cu INSIDE `namespace pano { namespace {` (no includes or namespaces of its own), and by the CPU
// emulation tier (tests/hostsim/warp_emu.cpp on tests/hostsim/cuda_emu.hpp), which compiles the same source with g++ and
// runs every kernel thread by thread against the oracle / cv2.  Under PANO_CUDA_EMU two things differ: the reciprocal
// seed (`rcp.approx` PTX on the device, 1 / w on the host - the acceptance test of the Newton step makes the result
// independent of the seed) and launch_fast (the <<< >>> launches), which the emulation mirrors.
// Semantics and roofline: see the header of warp.cu.

struct WarpParams {
  double M[9];  // inverse of T*H
  int bw0;      // OpenCV block width used for coordinate evaluation
  int cw, ch;
  int offx, offy, wl, hl;  // left ROI
  int ws, hs;              // source (right) size
  int bx0, by0, bx1, by1;  // canvas pixels outside this box provably map outside the source
  size_t src_bytes;        // bytes of the source buffer that may be read with word loads
  int y0;                  // first canvas row of the band being rendered (0 for a whole canvas)
};

// 6 consecutive bytes starting at p (two adjacent BGR pixels) from aligned 32-bit loads
__device__ __forceinline__ void load6(const uint8_t* __restrict__ p, uint32_t& lo, uint32_t& hi) {
  const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
  const uint32_t* q = reinterpret_cast<const uint32_t*>(p - a);
  const uint32_t w0 = q[0], w1 = q[1];
  const uint32_t w2 = a >= 3u ? q[2] : 0u;
  lo = __funnelshift_r(w0, w1, 8u * a);   // bytes p[0..3]
  hi = __funnelshift_r(w1, w2, 8u * a);   // bytes p[4..7]
}

// fixed-point bilinear tap combination for one pixel; fast path when all four taps are inside
__device__ __forceinline__ uint32_t warp_pixel_fast(const uint8_t* __restrict__ src, size_t sstride, int ws, int hs,
                                                    size_t src_bytes, int X, int Y) {
  const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
  if (sx >= ws || sx + 1 < 0 || sy >= hs || sy + 1 < 0) return 0u;
  const size_t off = (size_t)sy * sstride + 3 * (size_t)sx;
  if (sx >= 0 && sx + 1 < ws && sy >= 0 && sy + 1 < hs && off + sstride + 12 <= src_bytes) {
    const int fx = X & 31, fy = Y & 31;
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
    const int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    uint32_t a_lo, a_hi, b_lo, b_hi;
    load6(src + off, a_lo, a_hi);
    load6(src + off + sstride, b_lo, b_hi);
    uint32_t out = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int p00 = (a_lo >> (8 * c)) & 255;
      const int p01 = c == 0 ? (a_lo >> 24) : ((a_hi >> (8 * (c - 1))) & 255);
      const int p10 = (b_lo >> (8 * c)) & 255;
      const int p11 = c == 0 ? (b_lo >> 24) : ((b_hi >> (8 * (c - 1))) & 255);
      const int v = p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11;
      out |= (uint32_t)((v + (1 << 14)) >> 15) << (8 * c);
    }
    return out;
  }
  return warp_pixel(src, sstride, ws, hs, X, Y);  // border taps: general path
}

// Each thread produces 4 horizontally adjacent canvas pixels (12 bytes = three 32-bit stores;
// the canvas pitch is a multiple of 4).  MODE 0: plain warpPerspective; 1: + left copy (pair);
// 2: accumulate — non-black warped pixels overwrite what the canvas band already holds.
template <int MODE>
__global__ void __launch_bounds__(256)
warp_overlay_kernel(const uint8_t* __restrict__ left, size_t lstride, const uint8_t* __restrict__ right,
                    size_t rstride, WarpParams P, uint8_t* __restrict__ canvas, size_t cstride) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int yb = blockIdx.y * blockDim.y + threadIdx.y;   // row inside the band
  const int y = yb + P.y0;                                // row of the whole canvas
  if (x0 >= P.cw || yb >= P.ch) return;
  uint32_t px[4] = {0u, 0u, 0u, 0u};
  if (!(x0 + 3 < P.bx0 || x0 > P.bx1 || y < P.by0 || y > P.by1)) {
    // the four pixels share OpenCV's row-origin numerators when they lie in one bw0-block
    const int xb = (x0 / P.bw0) * P.bw0;
    const bool same = (x0 + 3) / P.bw0 == x0 / P.bw0;
    const WarpRow r = warp_row_origin(P.M, xb, y);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int x = x0 + i;
      if (x < P.cw) {
        int X, Y;
        if (same) warp_coord_from(P.M, r, x - xb, &X, &Y);
        else warp_coord(P.M, x, y, P.bw0, &X, &Y);
        px[i] = warp_pixel_fast(right, rstride, P.ws, P.hs, P.src_bytes, X, Y);
      }
    }
  }
  if (MODE == 1) {
    const int ly = y - P.offy;
    if (ly >= 0 && ly < P.hl) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int lx = x0 + i - P.offx;
        if (px[i] == 0u && lx >= 0 && lx < P.wl) {
          const uint8_t* p = left + (size_t)ly * lstride + 3 * (size_t)lx;
          px[i] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
        }
      }
    }
  }
  uint8_t* row = canvas + (size_t)yb * cstride + 3 * (size_t)x0;
  if (MODE == 2) {
    if ((px[0] | px[1] | px[2] | px[3]) == 0u) return;
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (px[i] != 0u && x0 + i < P.cw) {
        row[3 * i] = (uint8_t)px[i];
        row[3 * i + 1] = (uint8_t)(px[i] >> 8);
        row[3 * i + 2] = (uint8_t)(px[i] >> 16);
      }
    return;
  }
  if (x0 + 3 < P.cw && (reinterpret_cast<uintptr_t>(row) & 3) == 0) {
    uint32_t* o = reinterpret_cast<uint32_t*>(row);  // 3*x0 is a multiple of 12
    o[0] = px[0] | (px[1] << 24);
    o[1] = (px[1] >> 8) | (px[2] << 16);
    o[2] = (px[2] >> 16) | (px[3] << 8);
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (x0 + i < P.cw) {
        row[3 * i] = (uint8_t)px[i];
        row[3 * i + 1] = (uint8_t)(px[i] >> 8);
        row[3 * i + 2] = (uint8_t)(px[i] >> 16);
      }
  }
}

// ---------------------------------------------------------------------------------------
// Fast path (the common case: a proper projective map with a positive denominator on the whole
// canvas, 64-px coordinate blocks).  Same results, about a third of the instructions:
//   * lane = pixel (3-byte stride), so a warp's tap loads touch one or two cache lines; a warp
//     renders 256 px of one row, a block 8 adjacent rows (L1 reuse of the source rows);
//   * OpenCV's  X = rint((X0 + M0 x1) * (32 / W))  is evaluated with a checked Newton reciprocal
//     (MUFU.RCP64H seed, two steps, the second step's residual bounds the error) and a magic-number
//     rounding at 2^-20; whenever the approximation could round differently from the exact
//     expression (fraction within 2^-16 of .5, residual too large) the pixel takes the exact path
//     (IEEE division, pano_core.cuh warp_coord), so the fast path never changes a result;
//   * the factor 32 is folded into the numerator rows (a power of two commutes with rounding);
//   * separable fixed-point bilinear: h = (32-fx) p0 + fx p1 per row with DP4A on the raw BGRBGR
//     bytes, then ((32-fy) h0 + fy h1 + 512) >> 10, algebraically identical to OpenCV's
//     (sum w_i p_i + 2^14) >> 15 with w = 32 (32-fx)(32-fy) ... (no intermediate rounding);
//   * groups of 32 px outside the source footprint are straight copies of the left image.
// ---------------------------------------------------------------------------------------
struct FastParams {
  double Mx[3], My[3], Mw[3];  // 32 * M[0..2], 32 * M[3..5], M[6..8]  (M = inverse of T*H)
  WarpParams W;                // everything the exact path needs
  int copy_words;              // canvas base and pitch are 4-byte aligned: left-only groups are copied as words
};

__device__ __forceinline__ double rcp_seed(double w) {
#ifdef PANO_CUDA_EMU
  return 1.0 / w;   // (CPU emulation tier: any seed is admissible - the Newton step's acceptance test guards exactness)
#else
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w));
  return r;
#endif
}

// fixed-point bilinear, all four taps inside the source (sx in [0, ws-2], sy in [0, hs-3])
__device__ __forceinline__ uint32_t bilinear_dp4a(const uint8_t* __restrict__ src, uint32_t sstride, int sx, int sy,
                                                  int fx, int fy) {
  const uint32_t off = (uint32_t)sy * sstride + 3u * (uint32_t)sx;    // sources are < 4 GB
  const uint32_t a = off & 3u, sh = 8u * a;                           // base and pitch are multiples of 4
  const uint32_t* q0 = reinterpret_cast<const uint32_t*>(src + (off - a));
  const uint32_t* q1 = reinterpret_cast<const uint32_t*>(src + (off - a + sstride));
  const uint32_t u0 = q0[0], u1 = q0[1], u2 = q0[2];
  const uint32_t v0 = q1[0], v1 = q1[1], v2 = q1[2];
  const uint32_t alo = __funnelshift_r(u0, u1, sh), ahi = __funnelshift_r(u1, u2, sh);   // B0 G0 R0 B1 | G1 R1 . .
  const uint32_t blo = __funnelshift_r(v0, v1, sh), bhi = __funnelshift_r(v1, v2, sh);
  const uint32_t wx = (uint32_t)(32 - fx) | ((uint32_t)fx << 24);                        // bytes 0 and 3
  const int gy = 32 - fy;
  uint32_t out = 0;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const uint32_t ta = c == 0 ? alo : __funnelshift_r(alo, ahi, 8 * c);   // p0[c] . . p1[c]
    const uint32_t tb = c == 0 ? blo : __funnelshift_r(blo, bhi, 8 * c);
    const int h0 = (int)__dp4a(ta, wx, 0u), h1 = (int)__dp4a(tb, wx, 0u);
    const int sv = gy * h0 + fy * h1 + 512;
    out |= (uint32_t)(sv >> 10) << (8 * c);
  }
  return out;
}

template <int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB)
warp_fast_kernel(const uint8_t* __restrict__ left, size_t lstride, const uint8_t* __restrict__ right, size_t rstride,
                 const FastParams F, uint8_t* __restrict__ canvas, size_t cstride) {
  const WarpParams& P = F.W;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int yb = blockIdx.y * 8 + wrp;            // row inside the band
  if (yb >= P.ch) return;
  const int y = yb + P.y0;
  const int xc = blockIdx.x * 256;
  uint8_t* crow = canvas + (size_t)yb * cstride;
  const double yd = (double)y;
  const double cX = __dmul_rn(F.Mx[1], yd), cY = __dmul_rn(F.My[1], yd), cW = __dmul_rn(F.Mw[1], yd);
  const bool row_in = y >= P.by0 && y <= P.by1;
  const int ly = y - P.offy;
  const bool lrow = MODE == 1 && ly >= 0 && ly < P.hl;
  const uint8_t* lrowp = left + (size_t)(lrow ? ly : 0) * lstride;
  const uint32_t rs = (uint32_t)rstride;
#pragma unroll 1
  for (int blk = 0; blk < 4; blk++) {
    const int xb = xc + blk * 64;
    if (xb >= P.cw) break;
    const double xbd = (double)xb;
    // row-origin numerators of this 64-px block (ref arithmetic of warp_row_origin, X and Y scaled by 32)
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(F.Mx[0], xbd), cX), F.Mx[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(F.My[0], xbd), cY), F.My[2]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(F.Mw[0], xbd), cW), F.Mw[2]);
#pragma unroll
    for (int g = 0; g < 2; g++) {
      const int x0 = xb + 32 * g, x = x0 + lane;
      if (x0 >= P.cw) break;
      uint32_t px = 0u;
      const bool grp_in = row_in && !(x0 + 31 < P.bx0 || x0 > P.bx1);   // warp-uniform
      if (MODE == 1 && F.copy_words && !grp_in && lrow && x0 >= P.offx && x0 + 32 <= P.offx + P.wl && x0 + 32 <= P.cw &&
          ly + 1 < P.hl) {
        // 32 px of the left image, nothing of the right one: 24 aligned words (the canvas row and x0 * 3 are
        // 4-byte aligned; the source words are realigned with a funnel shift; the row below keeps the last
        // word's over-read inside the image)
        if (lane < 24) {
          const uint8_t* sp = lrowp + 3u * (uint32_t)(x0 - P.offx) + 4u * (uint32_t)lane;
          const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(sp) & 3u);
          const uint32_t* q = reinterpret_cast<const uint32_t*>(sp - a);
          const uint32_t w0 = q[0], w1 = q[1];
          reinterpret_cast<uint32_t*>(crow + 3u * (uint32_t)x0)[lane] = __funnelshift_r(w0, w1, 8u * a);
        }
        continue;
      }
      if (grp_in) {
        // OpenCV's X = rint((X0 + M0 x1) * (32 / W)) with a checked Newton reciprocal and magic rounding
        // (pano_core.cuh warp_coord_fast; the host tier tests it against the exact expression)
        const double Wseed = __dadd_rn(W0, __dmul_rn(F.Mw[0], (double)(32 * g + lane)));
        int X, Y;
        bool exact_needed;
        warp_coord_fast(X0, Y0, W0, F.Mx[0], F.My[0], F.Mw[0], (double)(32 * g + lane), rcp_seed(Wseed), &X, &Y, &exact_needed);
        if (x < P.cw) {
          if (exact_needed) warp_coord(P.M, x, y, P.bw0, &X, &Y);
          const int sx = X >> 5, sy = Y >> 5;
          if ((unsigned)sx < (unsigned)(P.ws - 1) && (unsigned)sy < (unsigned)(P.hs - 2))
            px = bilinear_dp4a(right, rs, sx, sy, X & 31, Y & 31);
          else
            px = warp_pixel(right, rstride, P.ws, P.hs, X, Y);   // border ring / outside: general path
        }
      }
      if (x < P.cw) {
        if (MODE == 1 && px == 0u && lrow) {
          const int lx = x - P.offx;
          if (lx >= 0 && lx < P.wl) {
            const uint8_t* lp = lrowp + 3u * (uint32_t)lx;
            px = (uint32_t)lp[0] | ((uint32_t)lp[1] << 8) | ((uint32_t)lp[2] << 16);
          }
        }
        if (MODE != 2 || px != 0u) {
          uint8_t* o = crow + 3u * (uint32_t)x;
          o[0] = (uint8_t)px;
          o[1] = (uint8_t)(px >> 8);
          o[2] = (uint8_t)(px >> 16);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Quad kernel (round 2): the fast path with FOUR adjacent canvas pixels per thread.
//   * a thread owns 12 contiguous canvas bytes = three aligned 32-bit words: no byte stores, a warp writes 384
//     contiguous bytes of one canvas row; the left image is read as words too (funnel-shifted to the canvas phase);
//   * the four pixels are independent chains (coordinates, 6 tap words each, DP4A bilinear): 24 tap loads in flight
//     per thread hide the L1/L2 latency that stalled the one-pixel-per-lane kernel (ncu r01: long-scoreboard bound);
//   * OpenCV's 64-px coordinate block, the row-origin numerators and every bounds decision are computed once per
//     thread instead of once per pixel; a 128-px warp span is two coordinate blocks (lanes 0-15 / 16-31);
//   * a block is 8 warps = 8 adjacent canvas rows of the same 128 columns, so the source rows are reused from L1.
// Arithmetic per pixel is unchanged (warp_coord_fast with its exact fallback, bilinear_dp4a, warp_pixel on the
// border ring), so results are identical to the other two kernels; tests/test_gpu_parity.py compares all three.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t px_of(const uint8_t* __restrict__ p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
}

template <int MODE>
__global__ void __launch_bounds__(256)
warp_quad_kernel(const uint8_t* __restrict__ left, size_t lstride, const uint8_t* __restrict__ right, size_t rstride,
                 const FastParams F, uint8_t* __restrict__ canvas, size_t cstride) {
  const WarpParams& P = F.W;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int yb = blockIdx.y * 8 + wrp;            // row inside the band
  const int x0 = blockIdx.x * 128 + lane * 4;     // first of this thread's four pixels
  if (yb >= P.ch || x0 >= P.cw) return;
  const int y = yb + P.y0;
  uint8_t* crow = canvas + (size_t)yb * cstride;
  uint32_t px[4] = {0u, 0u, 0u, 0u};
  const bool full = x0 + 3 < P.cw;                // all four pixels exist
  if (y >= P.by0 && y <= P.by1 && !(x0 + 3 < P.bx0 || x0 > P.bx1)) {
    const int xb = x0 & ~63;
    const double xbd = (double)xb, yd = (double)y;
    // row-origin numerators of the 64-px block (ref arithmetic of warp_row_origin, X and Y scaled by 32)
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(F.Mx[0], xbd), __dmul_rn(F.Mx[1], yd)), F.Mx[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(F.My[0], xbd), __dmul_rn(F.My[1], yd)), F.My[2]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(F.Mw[0], xbd), __dmul_rn(F.Mw[1], yd)), F.Mw[2]);
    const double x1d0 = (double)(x0 - xb);
    const uint32_t rs = (uint32_t)rstride;
    int X[4], Y[4];
    bool need[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const double x1d = x1d0 + (double)i;        // small integers: exact
      warp_coord_fast2(X0, Y0, W0, F.Mx[0], F.My[0], F.Mw[0], x1d, rcp_seed(__fma_rn(F.Mw[0], x1d, W0)), &X[i], &Y[i],
                       &need[i]);
    }
    bool fast4 = full & !(need[0] | need[1] | need[2] | need[3]);
#pragma unroll
    for (int i = 0; i < 4; i++)
      fast4 = fast4 & ((unsigned)(X[i] >> 5) < (unsigned)(P.ws - 1)) & ((unsigned)(Y[i] >> 5) < (unsigned)(P.hs - 2));
    if (fast4) {
      // the common case, branch-free: all 24 tap words are requested before the first one is used
      uint32_t u[4][6];
      uint32_t sh[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint32_t off = (uint32_t)(Y[i] >> 5) * rs + 3u * (uint32_t)(X[i] >> 5);    // sources are < 4 GB
        const uint32_t a = off & 3u;                                                   // base and pitch are multiples of 4
        sh[i] = 8u * a;
        const uint32_t* q0 = reinterpret_cast<const uint32_t*>(right + (off - a));
        const uint32_t* q1 = reinterpret_cast<const uint32_t*>(right + (off - a + rs));
        u[i][0] = q0[0]; u[i][1] = q0[1]; u[i][2] = q0[2];
        u[i][3] = q1[0]; u[i][4] = q1[1]; u[i][5] = q1[2];
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int fx = X[i] & 31, fy = Y[i] & 31;
        const uint32_t alo = __funnelshift_r(u[i][0], u[i][1], sh[i]), ahi = __funnelshift_r(u[i][1], u[i][2], sh[i]);
        const uint32_t blo = __funnelshift_r(u[i][3], u[i][4], sh[i]), bhi = __funnelshift_r(u[i][4], u[i][5], sh[i]);
        const uint32_t wx = (uint32_t)(32 - fx) | ((uint32_t)fx << 24);                 // bytes 0 and 3
        const int gy = 32 - fy;
        uint32_t out = 0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const uint32_t ta = c == 0 ? alo : __funnelshift_r(alo, ahi, 8 * c);          // p0[c] . . p1[c]
          const uint32_t tb = c == 0 ? blo : __funnelshift_r(blo, bhi, 8 * c);
          const int h0 = (int)__dp4a(ta, wx, 0u), h1 = (int)__dp4a(tb, wx, 0u);
          out |= (uint32_t)((gy * h0 + fy * h1 + 512) >> 10) << (8 * c);
        }
        px[i] = out;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (x0 + i < P.cw) {
          int Xi = X[i], Yi = Y[i];
          if (need[i]) warp_coord(P.M, x0 + i, y, P.bw0, &Xi, &Yi);
          const int sx = Xi >> 5, sy = Yi >> 5;
          uint32_t v;
          if ((unsigned)sx < (unsigned)(P.ws - 1) && (unsigned)sy < (unsigned)(P.hs - 2))
            v = bilinear_dp4a(right, rs, sx, sy, Xi & 31, Yi & 31);
          else
            v = warp_pixel(right, rstride, P.ws, P.hs, Xi, Yi);   // border ring / outside: general path
          px[i] = v;
        }
      }
    }
  }
  if (MODE == 1) {
    const int ly = y - P.offy, lx0 = x0 - P.offx;
    if (ly >= 0 && ly < P.hl && lx0 + 3 >= 0 && lx0 < P.wl && ((px[0] == 0u) | (px[1] == 0u) | (px[2] == 0u) | (px[3] == 0u))) {
      const uint8_t* lrowp = left + (size_t)ly * lstride;
      if (lx0 >= 0 && lx0 + 4 <= P.wl && ly + 1 < P.hl) {
        // 12 bytes of the left row as words (base and pitch of the engine's images are 4-byte aligned; the row below
        // keeps the last word's over-read inside the image)
        const uint8_t* sp = lrowp + 3u * (uint32_t)lx0;
        const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(sp) & 3u), sh = 8u * a;
        const uint32_t* q = reinterpret_cast<const uint32_t*>(sp - a);
        const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
        const uint32_t l0 = __funnelshift_r(w0, w1, sh), l1 = __funnelshift_r(w1, w2, sh), l2 = __funnelshift_r(w2, w3, sh);
        if (px[0] == 0u) px[0] = l0 & 0xffffffu;
        if (px[1] == 0u) px[1] = __funnelshift_r(l0, l1, 24) & 0xffffffu;
        if (px[2] == 0u) px[2] = __funnelshift_r(l1, l2, 16) & 0xffffffu;
        if (px[3] == 0u) px[3] = l2 >> 8;
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int lx = lx0 + i;
          if (px[i] == 0u && lx >= 0 && lx < P.wl) px[i] = px_of(lrowp + 3u * (uint32_t)lx);
        }
      }
    }
  }
  uint8_t* o = crow + 3u * (uint32_t)x0;
  if (MODE == 2) {
    // accumulate: only non-black warped pixels overwrite what the band already holds
    if ((px[0] | px[1] | px[2] | px[3]) == 0u) return;
    if (!(full && F.copy_words && px[0] != 0u && px[1] != 0u && px[2] != 0u && px[3] != 0u)) {
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (px[i] != 0u && x0 + i < P.cw) {
          o[3 * i] = (uint8_t)px[i];
          o[3 * i + 1] = (uint8_t)(px[i] >> 8);
          o[3 * i + 2] = (uint8_t)(px[i] >> 16);
        }
      return;
    }
  }
  if (full && F.copy_words) {
    uint32_t* ow = reinterpret_cast<uint32_t*>(o);   // 3 * x0 = 12 * (x0 / 4): word aligned
    ow[0] = px[0] | (px[1] << 24);
    ow[1] = (px[1] >> 8) | (px[2] << 16);
    ow[2] = (px[2] >> 16) | (px[3] << 8);
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (x0 + i < P.cw) {
        o[3 * i] = (uint8_t)px[i];
        o[3 * i + 1] = (uint8_t)(px[i] >> 8);
        o[3 * i + 2] = (uint8_t)(px[i] >> 16);
      }
  }
}

// PANO_WARP_KERNEL: 0 = quad kernel (default), 1 = the round-1 one-pixel-per-lane fast kernel (kept for A/B runs
// and as a second implementation the tests compare against)
int warp_kernel_choice() {
  const char* e = getenv("PANO_WARP_KERNEL");
  return e ? atoi(e) : 0;
}

// PANO_WARP_FAST=0 forces the general kernel (tests compare the two)
bool warp_fast_enabled() {
  const char* e = getenv("PANO_WARP_FAST");
  return !(e && atoi(e) == 0);
}

// Can the fast kernel be used?  Needs the footprint box (positive denominators), OpenCV's 64-px coordinate
// blocks, |coordinates| * 32 < 2^22 inside the canvas (magic rounding range) and a 4-byte aligned source.
bool fast_path_ok(const WarpParams& P, const uint8_t* src, size_t sstride, bool box_valid) {
  if (!box_valid || P.bw0 != 64 || P.ws < 4 || P.hs < 4) return false;
  if ((reinterpret_cast<uintptr_t>(src) & 3u) != 0 || (sstride & 3u) != 0) return false;
  const double cx[4] = {0.0, (double)P.cw, (double)P.cw, 0.0};
  const double cy[4] = {(double)P.y0, (double)P.y0, (double)(P.y0 + P.ch), (double)(P.y0 + P.ch)};
  double wmin = 1e300, nmax = 0;
  for (int i = 0; i < 4; i++) {
    const double w = P.M[6] * cx[i] + P.M[7] * cy[i] + P.M[8];
    wmin = fmin(wmin, w);
    nmax = fmax(nmax, fabs(P.M[0] * cx[i] + P.M[1] * cy[i] + P.M[2]));
    nmax = fmax(nmax, fabs(P.M[3] * cx[i] + P.M[4] * cy[i] + P.M[5]));
  }
  if (!(wmin > 1e-9) || !(nmax == nmax)) return false;
  return 32.0 * nmax / wmin < 4000000.0;   // < 2^22 with margin
}

#if !defined(PANO_CUDA_EMU) || defined(PANO_CUDA_EMU_LIB)   // (kernel launches: device build, and the whole-library emulation build)
template <int MODE>
void launch_fast(cudaStream_t st, const uint8_t* left, size_t lstride, const uint8_t* right, size_t rstride,
                 const WarpParams& P, uint8_t* canvas, size_t cstride) {
  FastParams F;
  for (int i = 0; i < 3; i++) { F.Mx[i] = 32.0 * P.M[i]; F.My[i] = 32.0 * P.M[3 + i]; F.Mw[i] = P.M[6 + i]; }
  F.W = P;
  F.copy_words = ((reinterpret_cast<uintptr_t>(canvas) & 3u) == 0 && (cstride & 3u) == 0) ? 1 : 0;
  const bool left_words = MODE != 1 || ((reinterpret_cast<uintptr_t>(left) & 3u) == 0 && (lstride & 3u) == 0);
  if (warp_kernel_choice() == 0 && left_words) {
    dim3 qgrid((P.cw + 127) / 128, (P.ch + 7) / 8);
    warp_quad_kernel<MODE><<<qgrid, dim3(256), 0, st>>>(left, lstride, right, rstride, F, canvas, cstride);
    return;
  }
  dim3 block(256), grid((P.cw + 255) / 256, (P.ch + 7) / 8);
  static const int minb = [] { const char* e = getenv("PANO_WARP_MINB"); return e ? atoi(e) : 5; }();
  if (minb >= 6)
    warp_fast_kernel<MODE, 6><<<grid, block, 0, st>>>(left, lstride, right, rstride, F, canvas, cstride);
  else if (minb == 5)
    warp_fast_kernel<MODE, 5><<<grid, block, 0, st>>>(left, lstride, right, rstride, F, canvas, cstride);
  else
    warp_fast_kernel<MODE, 4><<<grid, block, 0, st>>>(left, lstride, right, rstride, F, canvas, cstride);
}
#endif

// Canvas box outside which no pixel can receive a source tap: forward image (by `fwd`, the matrix
// whose inverse is iterated) of the source rectangle grown by 2 px, grown again by 2 px.  Only
// valid when the inverse map's denominator is positive on the whole canvas (then it is a proper
// projective bijection there and the image of the rectangle is the convex quad of its corners);
// otherwise the box is the whole canvas and every pixel is evaluated.
bool footprint_box(const double* fwd, const double* Minv, int ws, int hs, int cw, int ch, WarpParams& P) {
  P.bx0 = 0; P.by0 = 0; P.bx1 = cw - 1; P.by1 = ch - 1;
  const double cx[4] = {0.0, (double)cw, (double)cw, 0.0}, cy[4] = {0.0, 0.0, (double)ch, (double)ch};
  for (int i = 0; i < 4; i++) {
    double wd = Minv[6] * cx[i] + Minv[7] * cy[i] + Minv[8];
    if (!(wd > 1e-12)) return false;
  }
  const double sx[4] = {-2.0, ws + 1.0, ws + 1.0, -2.0}, sy[4] = {-2.0, -2.0, hs + 1.0, hs + 1.0};
  double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
  for (int i = 0; i < 4; i++) {
    double wd = fwd[6] * sx[i] + fwd[7] * sy[i] + fwd[8];
    if (!(wd > 1e-12)) return false;
    double X = (fwd[0] * sx[i] + fwd[1] * sy[i] + fwd[2]) / wd, Y = (fwd[3] * sx[i] + fwd[4] * sy[i] + fwd[5]) / wd;
    x0 = fmin(x0, X); x1 = fmax(x1, X); y0 = fmin(y0, Y); y1 = fmax(y1, Y);
  }
  if (!(x0 == x0) || !(x1 == x1) || !(y0 == y0) || !(y1 == y1)) return false;
  auto clampi = [](double v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : (int)v); };
  P.bx0 = clampi(floor(x0) - 2, 0, cw - 1);
  P.by0 = clampi(floor(y0) - 2, 0, ch - 1);
  P.bx1 = clampi(ceil(x1) + 2, 0, cw - 1);
  P.by1 = clampi(ceil(y1) + 2, 0, ch - 1);
  return true;
}

// pitched rows -> tightly packed rows (device to device), one 32-bit word of the destination per thread.  A canvas is
// rendered with a 256-byte aligned pitch (word stores, aligned rows) but handed to the host tightly packed
// (3 * canvas_w bytes per row, usually not a multiple of 4): a pitched 2-D device-to-host copy of ~2200 rows of
// 17 289 bytes runs far below PCIe rate, a flat copy of the packed buffer does not.
__global__ void __launch_bounds__(256)
pack_rows_kernel(const uint8_t* __restrict__ src, size_t pitch, uint32_t row_bytes, unsigned long long total_bytes,
                 uint8_t* __restrict__ dst) {
  const unsigned long long i = 4ull * ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x);   // first byte of this word
  if (i >= total_bytes) return;
  const uint32_t row = (uint32_t)(i / row_bytes), col = (uint32_t)(i - (unsigned long long)row * row_bytes);
  if (col + 4u <= row_bytes && i + 4ull <= total_bytes) {
    const uint8_t* sp = src + (size_t)row * pitch + col;      // pitch and base are multiples of 4: alignment = col & 3
    const uint32_t a = col & 3u;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(sp - a);
    const uint32_t w0 = q[0], w1 = a ? q[1] : 0u;
    *reinterpret_cast<uint32_t*>(dst + i) = __funnelshift_r(w0, w1, 8u * a);
  } else {
    for (uint32_t b = 0; b < 4u && i + b < total_bytes; b++) {  // the word straddles two rows (once per row) or the end
      const unsigned long long j = i + b;
      const uint32_t r = (uint32_t)(j / row_bytes), c = (uint32_t)(j - (unsigned long long)r * row_bytes);
      dst[j] = src[(size_t)r * pitch + c];
    }
  }
}

