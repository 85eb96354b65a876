// harris_kernels.cuh — the device code of harris.cu (K1 response, K1+K2 fused detector, K2 NMS / scan / scatter,
// the generic FP64 correlation and the flag compaction) and the host-computed Gaussian taps.
//
// Included by harris.cu INSIDE `namespace pano { namespace {` (no includes or namespaces of its own), and by the CPU
// emulation tier (tests/hostsim/harris_emu.cpp on tests/hostsim/cuda_emu.hpp) which compiles the same source with g++
// and runs it thread by thread against the oracle.  Two places differ under PANO_CUDA_EMU: the dynamic shared memory
// declaration and the TMA tile load of the fused kernel (inline PTX on the device, a model of the TMA unit's zero-filled
// box copy on the host); everything else - tile geometry, gray, Sobel, the order-exact FP64 Gaussian, response,
// threshold, strict NMS, ballots, mask / count updates, scan, ordered scatter - is shared.
// Semantics and roofline: see the header of harris.cu.

constexpr int TX = 32;        // tile width  (one warp spans a tile row)
constexpr int TY = 32;        // tile height
constexpr int RPT = 4;        // output rows per thread (vertical strip)
constexpr int BY = TY / RPT;  // 8 thread rows -> 256 threads
constexpr int PW = TX + 4, PH = TY + 4;  // product planes incl. Gaussian halo
constexpr int GW = TX + 6, GH = TY + 6;  // gray incl. Sobel halo

struct GaussTaps {
  double g[25];
};

__global__ void __launch_bounds__(TX* BY)
harris_response_kernel(const uint8_t* __restrict__ img, int w, int h, size_t stride, double kparam,
                       GaussTaps taps, double* __restrict__ resp, double thresh, uint32_t* __restrict__ cand,
                       int cand_stride) {
  __shared__ uint8_t sgray[GH][GW + 2];
  __shared__ double sxx[PH][PW];
  __shared__ double syy[PH][PW];
  __shared__ double sxy[PH][PW];

  const int tid = threadIdx.y * TX + threadIdx.x;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;

  // phase 1: gray tile with a 3-px halo (values outside the image are never used)
  for (int i = tid; i < GH * GW; i += TX * BY) {
    int gy = i / GW, gx = i - gy * GW;
    int X = x0 - 3 + gx, Y = y0 - 3 + gy;
    int v = 0;
    if (X >= 0 && X < w && Y >= 0 && Y < h) {
      const uint8_t* p = img + (size_t)Y * stride + 3 * (size_t)X;
      v = gray_u8(p[0], p[1], p[2]);
    }
    sgray[gy][gx] = (uint8_t)v;
  }
  __syncthreads();

  // phase 2: Sobel (integer, exact) and the three products on the tile + 2-px halo
  for (int i = tid; i < PH * PW; i += TX * BY) {
    int py = i / PW, px = i - py * PW;
    int X = x0 - 2 + px, Y = y0 - 2 + py;
    int gx = 0, gy = 0;
    if (X >= 1 && X <= w - 2 && Y >= 1 && Y <= h - 2) {
      int r = py + 1, c = px + 1;
      int a00 = sgray[r - 1][c - 1], a01 = sgray[r - 1][c], a02 = sgray[r - 1][c + 1];
      int a10 = sgray[r][c - 1], a12 = sgray[r][c + 1];
      int a20 = sgray[r + 1][c - 1], a21 = sgray[r + 1][c], a22 = sgray[r + 1][c + 1];
      gx = (a02 - a00) + 2 * (a12 - a10) + (a22 - a20);
      gy = (a20 - a00) + 2 * (a21 - a01) + (a22 - a02);
    }
    sxx[py][px] = (double)(gx * gx);
    syy[py][px] = (double)(gy * gy);
    sxy[py][px] = (double)(gx * gy);
  }
  __syncthreads();

  // phase 3: 5x5 Gaussian of the three planes for a vertical strip of RPT pixels.  Input
  // rows are visited in increasing order so every output accumulates its 25 terms in the
  // reference's (row, column) order.
  const int tx = threadIdx.x, ty = threadIdx.y;
  double axx[RPT], ayy[RPT], axy[RPT];
#pragma unroll
  for (int o = 0; o < RPT; o++) axx[o] = ayy[o] = axy[o] = 0.0;
#pragma unroll
  for (int r = 0; r < RPT + 4; r++) {
    double vxx[5], vyy[5], vxy[5];
#pragma unroll
    for (int j = 0; j < 5; j++) {
      vxx[j] = sxx[ty * RPT + r][tx + j];
      vyy[j] = syy[ty * RPT + r][tx + j];
      vxy[j] = sxy[ty * RPT + r][tx + j];
    }
#pragma unroll
    for (int o = 0; o < RPT; o++) {
      int i = r - o;  // kernel row for output o
      if (i >= 0 && i < 5) {
#pragma unroll
        for (int j = 0; j < 5; j++) {
          double g = taps.g[i * 5 + j];
          axx[o] = __dadd_rn(axx[o], __dmul_rn(vxx[j], g));
          ayy[o] = __dadd_rn(ayy[o], __dmul_rn(vyy[j], g));
          axy[o] = __dadd_rn(axy[o], __dmul_rn(vxy[j], g));
        }
      }
    }
  }
  const int X = x0 + tx;
#pragma unroll
  for (int o = 0; o < RPT; o++) {
    int Y = y0 + ty * RPT + o;
    double r = 0.0;
    if (X < w && Y < h) {
      if (X >= 2 && X <= w - 3 && Y >= 2 && Y <= h - 3) r = harris_resp(axx[o], ayy[o], axy[o], kparam);
      resp[(size_t)Y * w + X] = r;
    }
    // one word per 32-px row segment: which pixels exceed the NMS threshold at all (the NMS kernel then only
    // loads the response around those; a warp is one row of the tile, the tile is 32 px wide)
    if (cand != nullptr) {
      const unsigned bits = __ballot_sync(0xffffffffu, X < w && Y < h && r > thresh);
      if (tx == 0 && Y < h) cand[(size_t)Y * cand_stride + blockIdx.x] = bits;
    }
  }
}

// ---------------------------------------------------------------------------------------
// K1+K2 fused (round 2): response AND strict 3x3 NMS in one kernel; the FP64 response plane (66 MB per 4K image)
// is never written.  A CTA computes the response of a 32 x 64 tile in shared memory and decides the inner
// 30 x 62 pixels (tiles overlap by one response pixel on every side: +10 % FP64 work instead of a 66 MB write,
// a second kernel and its re-read; under 24 concurrent lanes that plane did not fit the L2 and the detect stage
// went from 0.34 to 2.7 ms per pair).  Data flow per CTA:
//   TMA      one 2-D tile load (cp.async.bulk.tensor.2d) of 144 bytes x 70 rows of the interleaved BGR image
//            (38 pixels = 32 + the 3-px halo on each side, starting at the 16-byte boundary below the first byte;
//            out-of-image bytes are zero-filled by the TMA unit);
//            images whose base / pitch are not 16-byte aligned take a plain-load fallback of the same tile
//   phase 1  gray (15-bit fixed point) of 38 x 70 pixels
//   phase 2  Sobel (integer) and the three products on 36 x 68
//   phase 3  5x5 Gaussian in the reference's summation order, 8-row vertical strip per thread (each loaded
//            product row feeds up to five outputs), response -> shared memory
//   phase 4  threshold + strict NMS on the inner pixels, one ballot word per tile row, OR-ed into the 1-bit/px
//            mask (keypoints are sparse: almost every word is zero and skipped) + per-row counts
// Tried and dropped (round 2): a persistent warp-specialised form (4 producer warps running TMA + gray + Sobel of tile
// k + 1 while 8 consumer warps run the FP64 phase of tile k, one CTA per SM): 146.7 us per 4K image against 137.3 us
// for this kernel - with one CTA per SM only two FP64 warps per scheduler remain, and the stencil's dependent
// DADD chains need the four that two co-resident CTAs of this kernel provide (ncu: 26 % "wait" stalls).
// Semantics: ref src/serial/main.cpp:119-180, bit-identical keypoints (the response values are the same doubles).
// ---------------------------------------------------------------------------------------
constexpr int FX = 32, FY = 64;           // response tile
constexpr int FRPT = 8;                   // output rows per thread
constexpr int FBY = FY / FRPT;            // 8 thread rows -> 256 threads
constexpr int FPW = FX + 4, FPH = FY + 4; // product planes incl. Gaussian halo
constexpr int FGW = FX + 6, FGH = FY + 6; // gray incl. Sobel halo
constexpr int FRAW = 144;                 // bytes per raw tile row: 3 * FGW = 114 plus up to 15 bytes of alignment slack (the
                                          // innermost TMA coordinate of a byte tensor must be a multiple of 16: measured, an
                                          // unaligned one raises 'illegal instruction'; tools/tma_probe.cu), multiple of 16

struct FusedSmem {
  double xx[FPH][FPW];
  double yy[FPH][FPW];
  double xy[FPH][FPW];
  double resp[FY][FX + 1];
  __align__(128) uint8_t raw[FGH][FRAW];
  uint8_t gray[FGH][FGW + 2];
  unsigned long long bar;
};

template <bool USE_TMA>
__global__ void __launch_bounds__(FX* FBY, 2)
harris_fused_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ img, int w, int h,
                    size_t stride, double kparam, GaussTaps taps, double thresh, uint32_t* __restrict__ mask,
                    int mask_stride, uint32_t* __restrict__ rowcnt) {
#ifdef PANO_CUDA_EMU
  uint8_t* fused_smem_raw = emu::dyn_smem();   // (CPU emulation tier: the launch's dynamic shared memory)
#else
  extern __shared__ __align__(128) uint8_t fused_smem_raw[];
#endif
  FusedSmem& S = *reinterpret_cast<FusedSmem*>(fused_smem_raw);
  const int tid = threadIdx.y * FX + threadIdx.x;
  // response tile origin: the inner (decided) pixels are x0 + 1 .. x0 + FX - 2, y0 + 1 .. y0 + FY - 2
  const int x0 = blockIdx.x * (FX - 2) - 1, y0 = blockIdx.y * (FY - 2) - 1;
  const int braw = 3 * (x0 - 3);       // first image byte of the tile row (may be negative)
  const int araw = braw & ~15;         // 16-byte boundary at or below it (floor, also for negative values)
  const int boff = braw - araw;        // 0 .. 15: where pixel 0 of the tile sits inside a raw row

  // ---- raw BGR tile: pixels x0 - 3 .. x0 - 3 + 37 inside 144 bytes from the aligned start, rows y0 - 3 .. y0 - 3 + 69
  if (USE_TMA) {
#ifdef PANO_CUDA_EMU
    // CPU emulation tier: the TMA unit's 2-D tile load as a model (box FRAW bytes x FGH rows from byte `araw`, row
    // y0 - 3; bytes outside the tensor are zero-filled), issued by one thread like the real one
    if (tid == 0) emu::tma_tile_load_2d(&S.raw[0][0], tmap, araw, y0 - 3, FRAW, FGH);
    __syncthreads();
#else
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&S.bar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(FRAW * FGH)) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
          ::"r"((uint32_t)__cvta_generic_to_shared(&S.raw[0][0])), "l"(reinterpret_cast<uint64_t>(&tmap)),
            "r"(araw), "r"(y0 - 3), "r"(bar) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(bar) : "memory");
    }
#endif
  } else {
    for (int i = tid; i < FGH * (FRAW / 4); i += FX * FBY) {
      const int ry = i / (FRAW / 4), rw = i - ry * (FRAW / 4);
      const int Y = y0 - 3 + ry;
      uint32_t v = 0;
      if (Y >= 0 && Y < h) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const int B = araw + 4 * rw + b;
          if (B >= 0 && B < 3 * w) v |= (uint32_t)img[(size_t)Y * stride + B] << (8 * b);
        }
      }
      reinterpret_cast<uint32_t*>(&S.raw[ry][0])[rw] = v;
    }
    __syncthreads();
  }

  // ---- phase 1: gray (values outside the image are zero bytes -> gray 0, never used by a decided pixel)
  for (int i = tid; i < FGH * FGW; i += FX * FBY) {
    const int gy = i / FGW, gx = i - gy * FGW;
    const uint8_t* p = &S.raw[gy][boff + 3 * gx];
    S.gray[gy][gx] = (uint8_t)gray_u8(p[0], p[1], p[2]);
  }
  __syncthreads();

  // ---- phase 2: Sobel (integer, exact) and the three products on the tile + 2-px halo
  for (int i = tid; i < FPH * FPW; i += FX * FBY) {
    const int py = i / FPW, px = i - py * FPW;
    const int X = x0 - 2 + px, Y = y0 - 2 + py;
    int gx = 0, gy = 0;
    if (X >= 1 && X <= w - 2 && Y >= 1 && Y <= h - 2) {
      const int r = py + 1, c = px + 1;
      const int a00 = S.gray[r - 1][c - 1], a01 = S.gray[r - 1][c], a02 = S.gray[r - 1][c + 1];
      const int a10 = S.gray[r][c - 1], a12 = S.gray[r][c + 1];
      const int a20 = S.gray[r + 1][c - 1], a21 = S.gray[r + 1][c], a22 = S.gray[r + 1][c + 1];
      gx = (a02 - a00) + 2 * (a12 - a10) + (a22 - a20);
      gy = (a20 - a00) + 2 * (a21 - a01) + (a22 - a02);
    }
    S.xx[py][px] = (double)(gx * gx);
    S.yy[py][px] = (double)(gy * gy);
    S.xy[py][px] = (double)(gx * gy);
  }
  __syncthreads();

  // ---- phase 3: 5x5 Gaussian of the three planes for a vertical strip of FRPT pixels (input rows in increasing
  // order, so every output accumulates its 25 terms in the reference's (row, column) order), response
  const int tx = threadIdx.x, ty = threadIdx.y;
  {
    double axx[FRPT], ayy[FRPT], axy[FRPT];
#pragma unroll
    for (int o = 0; o < FRPT; o++) axx[o] = ayy[o] = axy[o] = 0.0;
#pragma unroll
    for (int r = 0; r < FRPT + 4; r++) {
      double vxx[5], vyy[5], vxy[5];
#pragma unroll
      for (int j = 0; j < 5; j++) {
        vxx[j] = S.xx[ty * FRPT + r][tx + j];
        vyy[j] = S.yy[ty * FRPT + r][tx + j];
        vxy[j] = S.xy[ty * FRPT + r][tx + j];
      }
#pragma unroll
      for (int o = 0; o < FRPT; o++) {
        const int i = r - o;  // kernel row for output o
        if (i >= 0 && i < 5) {
#pragma unroll
          for (int j = 0; j < 5; j++) {
            const double g = taps.g[i * 5 + j];
            axx[o] = __dadd_rn(axx[o], __dmul_rn(vxx[j], g));
            ayy[o] = __dadd_rn(ayy[o], __dmul_rn(vyy[j], g));
            axy[o] = __dadd_rn(axy[o], __dmul_rn(vxy[j], g));
          }
        }
      }
    }
    const int X = x0 + tx;
#pragma unroll
    for (int o = 0; o < FRPT; o++) {
      const int Y = y0 + ty * FRPT + o;
      double r = 0.0;
      if (X >= 2 && X <= w - 3 && Y >= 2 && Y <= h - 3) r = harris_resp(axx[o], ayy[o], axy[o], kparam);
      S.resp[ty * FRPT + o][tx] = r;
    }
  }
  __syncthreads();

  // ---- phase 4: threshold + strict 3x3 NMS of the inner pixels (ref :157-180: y, x in [1, size - 2])
  {
    const int X = x0 + tx;
    // word / shift of this tile row's 32 ballot bits inside the mask row (x0 may be -1)
    const int wb = (x0 >= 0) ? (x0 >> 5) : -1;
    const int sh = x0 - 32 * wb;
#pragma unroll
    for (int o = 0; o < FRPT; o++) {
      const int ry = ty * FRPT + o, Y = y0 + ry;
      bool keep = false;
      if (tx >= 1 && tx <= FX - 2 && ry >= 1 && ry <= FY - 2 && X >= 1 && X <= w - 2 && Y >= 1 && Y <= h - 2) {
        const double r = S.resp[ry][tx];
        if (r > thresh) {
          keep = r > S.resp[ry - 1][tx - 1] && r > S.resp[ry - 1][tx] && r > S.resp[ry - 1][tx + 1] &&
                 r > S.resp[ry][tx - 1] && r > S.resp[ry][tx + 1] &&
                 r > S.resp[ry + 1][tx - 1] && r > S.resp[ry + 1][tx] && r > S.resp[ry + 1][tx + 1];
        }
      }
      const unsigned bits = __ballot_sync(0xffffffffu, keep);
      if (bits != 0u && tx == 0) {
        const unsigned long long v = (unsigned long long)bits << sh;
        const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
        uint32_t* mrow = mask + (size_t)Y * mask_stride;
        if (lo != 0u && wb >= 0) atomicOr(&mrow[wb], lo);
        if (hi != 0u) atomicOr(&mrow[wb + 1], hi);
        atomicAdd(&rowcnt[Y], (uint32_t)__popc(bits));
      }
    }
  }
}

// K2a: threshold + strict NMS over a (2*half+1)^2 neighbourhood -> bit mask + per-row counts.
// ref: src/serial/main.cpp:157-180 (keep iff resp > thresh and resp > every neighbour).
template <int HALF_T>   // HALF_T > 0: neighbourhood known at compile time (3x3: the reference's setting); 0: runtime `half`
__global__ void nms_mask_kernel(const double* __restrict__ resp, int w, int h, double thresh, int half,
                                uint32_t* __restrict__ mask, int mask_stride,
                                uint32_t* __restrict__ rowcnt) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (y >= h) return;
  if (HALF_T > 0) half = HALF_T;
  bool keep = false;
  // the response kernel left "response > threshold" bits in the mask word this warp is about to overwrite
  const uint32_t cword = mask[(size_t)y * mask_stride + blockIdx.x];
  if (cword == 0u) return;   // (the word already holds the result: no keypoint in these 32 px)
  if (((cword >> threadIdx.x) & 1u) && x >= half && x < w - half && y >= half && y < h - half) {
    const double* c = resp + (size_t)y * w + x;
    const double r = *c;
    if (r > thresh) {
      if (HALF_T > 0) {
        // all neighbours requested at once (independent loads), then one AND: "keep iff r > every neighbour" does
        // not depend on the order the reference visits them in
        double nb[(2 * HALF_T + 1) * (2 * HALF_T + 1)];
#pragma unroll
        for (int i = -HALF_T; i <= HALF_T; i++)
#pragma unroll
          for (int j = -HALF_T; j <= HALF_T; j++) nb[(i + HALF_T) * (2 * HALF_T + 1) + (j + HALF_T)] = c[(ptrdiff_t)i * w + j];
        keep = true;
#pragma unroll
        for (int q = 0; q < (2 * HALF_T + 1) * (2 * HALF_T + 1); q++)
          if (q != (2 * HALF_T + 1) * HALF_T + HALF_T) keep = keep && (r > nb[q]);
      } else {
        keep = true;
        for (int i = -half; i <= half && keep; i++)
          for (int j = -half; j <= half; j++) {
            if (i == 0 && j == 0) continue;
            if (!(r > c[(ptrdiff_t)i * w + j])) { keep = false; break; }
          }
      }
    }
  }
  unsigned b = __ballot_sync(0xffffffffu, keep);
  if (threadIdx.x == 0) {
    mask[(size_t)y * mask_stride + blockIdx.x] = b;
    if (b) atomicAdd(&rowcnt[y], __popc(b));
  }
}

// K2b: ordered scatter, one warp per image row.
__global__ void scatter_keypoints_kernel(const uint32_t* __restrict__ mask, int mask_stride, int h,
                                         const uint32_t* __restrict__ rowoff, int32_t* __restrict__ xy) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= h) return;
  uint32_t base = rowoff[row];
  for (int w0 = 0; w0 < mask_stride; w0 += 32) {
    uint32_t word = (w0 + lane < mask_stride) ? mask[(size_t)row * mask_stride + w0 + lane] : 0u;
    uint32_t c = __popc(word), incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    uint32_t pos = base + incl - c;
    while (word) {
      int b = __ffs(word) - 1;
      word &= word - 1;
      xy[2 * (size_t)pos] = (w0 + lane) * 32 + b;
      xy[2 * (size_t)pos + 1] = row;
      pos++;
    }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// single-block exclusive scan with a running carry over chunks of blockDim.x
__global__ void scan_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n,
                            uint32_t* __restrict__ total) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t woff[32];
  __shared__ uint32_t chunk_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint32_t carry = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    int i = base + threadIdx.x;
    uint32_t v = i < n ? in[i] : 0u, incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      uint32_t s = lane < nw ? wsum[lane] : 0u, si = s;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, si, d);
        if (lane >= d) si += t;
      }
      woff[lane] = si - s;
      if (lane == 31) chunk_total = si;
    }
    __syncthreads();
    if (i < n) out[i] = carry + woff[wid] + incl - v;
    carry += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

// generic FP64 correlation with a zero border (ref: convolveSequential / convolveCUDA)
__global__ void convolve_f64_kernel(const double* __restrict__ in, int w, int h,
                                    const double* __restrict__ kern, int ksize, double* __restrict__ out) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  int k = ksize / 2;
  double sum = 0.0;
  if (x >= k && x < w - k && y >= k && y < h - k) {
    for (int i = -k; i <= k; i++)
      for (int j = -k; j <= k; j++)
        sum = __dadd_rn(sum, __dmul_rn(in[(size_t)(y + i) * w + (x + j)], kern[(k + i) * ksize + (k + j)]));
  }
  out[(size_t)y * w + x] = sum;
}

// block counts of flagged items (256 per block)
__global__ void flag_count_kernel(const uint8_t* __restrict__ flags, int n, uint32_t* __restrict__ bc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int f = (i < n && flags[i]) ? 1 : 0;
  int c = __syncthreads_count(f);
  if (threadIdx.x == 0) bc[blockIdx.x] = c;
}

__global__ void flag_scatter_kernel(const uint8_t* __restrict__ flags, int n, const uint32_t* __restrict__ boff,
                                    int32_t* __restrict__ out) {
  __shared__ uint32_t wcnt[8];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  bool f = (i < n && flags[i]);
  unsigned b = __ballot_sync(0xffffffffu, f);
  if (lane == 0) wcnt[wid] = __popc(b);
  __syncthreads();
  uint32_t off = boff[blockIdx.x];
  for (int k = 0; k < wid; k++) off += wcnt[k];
  if (f) out[off + __popc(b & ((1u << lane) - 1))] = i;
}

// getGaussianKernel(5, 1.0) exactly as the reference computes it on the host
// (ref: src/serial/main.cpp:73-91; libm exp, row-major running sum, divide by the sum).
GaussTaps make_taps() {
  GaussTaps t;
  const int ks = 5, half = 2;
  const double sigma = 1.0;
  double sum = 0.0;
  for (int i = 0; i < ks; ++i) {
    int x = i - half;
    for (int j = 0; j < ks; ++j) {
      int y = j - half;
      t.g[i * ks + j] = exp(-(x * x + y * y) / (2 * sigma * sigma));
      sum += t.g[i * ks + j];
    }
  }
  for (int i = 0; i < 25; i++) t.g[i] /= sum;
  return t;
}

