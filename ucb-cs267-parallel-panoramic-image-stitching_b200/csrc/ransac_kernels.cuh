// ransac_kernels.cuh — the device code of ransac.cu: the mt19937 stream generator, the shuffle replay (rejection
// cells, candidate walks, chain, segment replay, sample combination; the resident one-CTA variant), the point builder,
// the warp-per-hypothesis DLT (OpenCV's 4-point findHomography), scoring, selection and the inlier mask.
//
// Included by ransac.cu INSIDE `namespace pano { namespace {` (no includes or namespaces of its own), and by the CPU
// emulation tier (tests/hostsim/ransac_emu.cpp on tests/hostsim/cuda_emu.hpp), which compiles the same source with g++
// and runs it thread by thread against the oracle's real std::shuffle / cv2-pinned findHomography.  Under PANO_CUDA_EMU
// the few inline-PTX spots are spelled out in C++ (the add-with-carry of cell_step, cp.async as an immediate copy with
// no-op commit / wait groups) and the dynamic shared memory comes from the emulation; pdl_wait() is a no-op there.
// Needs replay_plan.hpp, pano_core.cuh, pano_dmatch and PANO_ERRW_BAD_INDEX in scope.
// Semantics: see the header of ransac.cu (ref src/serial/main.cpp:247-307).


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
// Pass 1 for one chunk of G iterations: thread (g, j) walks iteration g from candidate start
// offset j of its window and records where the walk ends.  The last block to finish then
// chains the chunk from its exact base offset: iteration g's true candidate is the one that
// starts where iteration g-1's true candidate ended.  Exact: a true start outside its window
// is detected (status bit 0), never guessed.
struct ReplayCtl {
  unsigned long long base;  // exact stream offset of the chunk's first iteration
  int status;               // bit 0: window miss, bit 1: stream too short
  unsigned int done;        // blocks finished (last-block-done hand-over)
};

constexpr int RW_THREADS = 128;
constexpr size_t CHAIN_SMEM_MAX = 200 * 1024;  // + 20 KB static: under the 227 KB per-block limit

// Pass 1a: rejection cells of a whole chunk.  One warp = one (iteration g, 32-step block kb) task; it keeps the
// block's 32 (range, threshold) pairs in registers and sweeps the iteration's diagonals in tiles of 128.
// Lane l owns the four diagonals 4l .. 4l+3 of a tile: the cell (diagonal d, step i) reads stream word
// pos + d + i, so ONE aligned 16-byte shared-memory load (words 4l + 4q .. 4l + 4q + 3) feeds 16 cells - four
// steps of each of the lane's four diagonals - and the kernel is bound by its three ALU instructions per cell,
// not by shared-memory loads (the earlier lane = diagonal mapping spent one LDS per cell and was LSU bound).
// The test bit is shifted into the diagonal's word with an add-with-carry pair: ~lo = x*(-r) - 1 (one IMAD), and
// T + ~lo carries out of 32 bits exactly when lo32(x*r) < T, so "add.cc; addc w, w, w" is w = 2w + rejected.
// Steps are visited from 31 down to 0 so that bit k is step k; a lane's four words are one 16-byte store.
__device__ __forceinline__ void cell_step(uint32_t& w, uint32_t x, uint32_t neg_r, uint32_t T) {
  const uint32_t nlo = x * neg_r + 0xffffffffu;   // ~(x * r)
#ifdef PANO_CUDA_EMU
  w = 2u * w + (uint32_t)(((unsigned long long)T + nlo) >> 32);   // (CPU emulation tier: the carry, spelled out)
#else
  asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, %0;\n\t}" : "+r"(w) : "r"(T), "r"(nlo));
#endif
}

constexpr int CELL_WIN = 192;   // stream words staged per 128-diagonal tile (128 + 32 + 3, rounded up to 6 x 32)

__global__ void __launch_bounds__(256)
replay_cells_kernel(const uint32_t* __restrict__ X, uint32_t steps, const RT* __restrict__ rt,
                    const WinEntry* __restrict__ win, int G, uint32_t nkb, uint32_t dextra, const ReplayCtl* ctl,
                    uint32_t* __restrict__ bits, unsigned long long x_limit) {
  pdl_wait();
  __shared__ __align__(16) uint32_t s_x[8][CELL_WIN];
  __shared__ RT s_rt[8][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + wid;
  if (wg >= (long long)G * nkb) return;
  const int g = (int)(wg / nkb);
  const uint32_t kb = (uint32_t)(wg - (long long)g * nkb);
  const WinEntry we = win[g];
  const uint32_t D = (we.width + dextra + 31u) / 32u * 32u;
  {
    const uint32_t k = kb * 32u + lane;   // steps past the end carry (0, 0): lo < 0 never holds
    RT q;
    q.r = 0; q.T = 0;
    if (k < steps) q = rt[k];
    s_rt[wid][lane] = q;
  }
  __syncwarp();
  uint32_t rr[32], tt[32];
#pragma unroll
  for (int i = 0; i < 32; i++) { rr[i] = 0u - s_rt[wid][i].r; tt[i] = s_rt[wid][i].T; }
  const unsigned long long pos0 = ctl->base + (unsigned long long)g * steps + we.lo + (unsigned long long)kb * 32u;
  uint32_t* out = bits + (size_t)we.dfirst * nkb + (size_t)kb * D;   // 16-byte aligned: dfirst, D multiples of 32
  for (uint32_t d0 = 0; d0 < D; d0 += 128u) {
    const unsigned long long pos = pos0 + d0;   // cell (d0 + d, step i) reads word pos + d + i
    const bool in_range = pos + CELL_WIN < x_limit;
    __syncwarp();
#pragma unroll
    for (int u = 0; u < CELL_WIN / 32; u++) s_x[wid][32 * u + lane] = in_range ? X[pos + 32 * u + lane] : 0xffffffffu;
    __syncwarp();
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    const uint4* xq = reinterpret_cast<const uint4*>(&s_x[wid][4 * lane]);
#pragma unroll
    for (int q = 8; q >= 0; q--) {
      const uint4 v = xq[q];                    // words 4 lane + 4 q + (0..3)
      const uint32_t xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 3; j >= 0; j--) {            // diagonal 4 lane + r uses word j at step i = 4 q + j - r
        if (4 * q + j - 0 >= 0 && 4 * q + j - 0 < 32) cell_step(w0, xv[j], rr[(4 * q + j - 0) & 31], tt[(4 * q + j - 0) & 31]);
        if (4 * q + j - 1 >= 0 && 4 * q + j - 1 < 32) cell_step(w1, xv[j], rr[(4 * q + j - 1) & 31], tt[(4 * q + j - 1) & 31]);
        if (4 * q + j - 2 >= 0 && 4 * q + j - 2 < 32) cell_step(w2, xv[j], rr[(4 * q + j - 2) & 31], tt[(4 * q + j - 2) & 31]);
        if (4 * q + j - 3 >= 0 && 4 * q + j - 3 < 32) cell_step(w3, xv[j], rr[(4 * q + j - 3) & 31], tt[(4 * q + j - 3) & 31]);
      }
    }
    const uint32_t d = d0 + 4u * lane;
    if (d < D) *reinterpret_cast<uint4*>(out + d) = make_uint4(w0, w1, w2, w3);   // D is a multiple of 32: all or none
  }
}

// Pass 1b: thread (g, j) scans the cell words of iteration g from candidate start j.
__global__ void __launch_bounds__(RW_THREADS)
replay_walk_bits_kernel(uint32_t steps, const WinEntry* __restrict__ win, uint32_t nkb, uint32_t dextra,
                        ReplayCtl* ctl, const uint32_t* __restrict__ bits, uint32_t* __restrict__ cand_end,
                        uint32_t* __restrict__ seg_off, int n_cand, unsigned long long stream_len) {
  pdl_wait();
  const int g = blockIdx.y;
  const WinEntry we = win[g];
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= we.width) return;
  const uint32_t D = (we.width + dextra + 31u) / 32u * 32u;
  const unsigned long long start = ctl->base + (unsigned long long)g * steps + we.lo + j;
  uint32_t end = 0xffffffffu;
  if (start + 2ull * steps + 64ull < stream_len) {
    end = walk_bits(bits + (size_t)we.dfirst * nkb, D, nkb, j, steps, (uint32_t)g * steps + we.lo,
                    seg_off + we.first + j, (size_t)n_cand);
    if (end == 0xffffffffu) atomicOr(&ctl->status, 8);  // left the evaluated diagonals: re-plan wider
  } else {
    atomicOr(&ctl->status, 2);
  }
  cand_end[we.first + j] = end;
}

// One block: stage the chunk's candidate end offsets and windows in shared memory, chain
// sequentially (G dependent shared-memory lookups), then copy the true candidates' segment
// offsets into the per-iteration table used by pass 2.
// seg_tab[(t * nseg + s)] = absolute stream offset at which segment s of iteration t starts.
__global__ void __launch_bounds__(1024)
replay_chain_kernel(const WinEntry* __restrict__ win, int G, uint32_t steps, const uint32_t* __restrict__ cand_end,
                    const uint32_t* __restrict__ seg_off, int n_cand, int nseg, ReplayCtl* ctl,
                    unsigned long long* __restrict__ seg_tab /* this chunk's slice */) {
  pdl_wait();
#ifdef PANO_CUDA_EMU
  uint32_t* s_end = reinterpret_cast<uint32_t*>(emu::dyn_smem());   // (CPU emulation tier: the launch's dynamic shared memory)
#else
  extern __shared__ __align__(16) uint32_t s_end[];
#endif
  __shared__ WinEntry s_win[1024];
  __shared__ int s_pick[1024];
  __shared__ unsigned long long s_base;
  const bool in_smem = (size_t)n_cand * sizeof(uint32_t) <= CHAIN_SMEM_MAX;
  if (in_smem) {
    // 16-byte loads, 4 in flight per thread (cand_end is a cudaMalloc'ed array: 16-byte aligned)
    const int n4 = n_cand >> 2;
    const uint4* src4 = reinterpret_cast<const uint4*>(cand_end);
    uint4* dst4 = reinterpret_cast<uint4*>(s_end);
    for (int i = threadIdx.x; i < n4; i += 4 * blockDim.x) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = i + u * (int)blockDim.x;
        v[u] = e < n4 ? src4[e] : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = i + u * (int)blockDim.x;
        if (e < n4) dst4[e] = v[u];
      }
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n_cand; i += blockDim.x) s_end[i] = cand_end[i];
  }
  for (int i = threadIdx.x; i < G; i += blockDim.x) s_win[i] = win[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long base = ctl->base;
    s_base = base;
    uint32_t rel = 0;
    bool bad = ctl->status != 0;
    for (int g = 0; g < G; g++) {
      s_pick[g] = -1;
      if (bad) continue;
      const WinEntry we = s_win[g];
      long long j = (long long)rel - ((long long)g * steps + we.lo);
      if (j < 0 || j >= (long long)we.width) {
        bad = true;
        atomicOr(&ctl->status, 1);
        continue;
      }
      const uint32_t pick = we.first + (uint32_t)j;
      const uint32_t e = in_smem ? s_end[pick] : cand_end[pick];
      if (e == 0xffffffffu) { bad = true; continue; }
      s_pick[g] = (int)pick;
      seg_tab[(size_t)g * nseg] = base + rel;
      rel = e;
    }
    if (!bad) ctl->base = base + rel;
  }
  __syncthreads();
  const unsigned long long base = s_base;
  for (int i = threadIdx.x; i < G * nseg; i += blockDim.x) {
    const int g = i / nseg, sgm = i - g * nseg;
    const int pick = s_pick[g];
    if (pick < 0) seg_tab[(size_t)g * nseg + sgm] = ~0ull;
    else if (sgm > 0) seg_tab[(size_t)g * nseg + sgm] = base + seg_off[(size_t)(sgm - 1) * n_cand + pick];
  }
}

// Pass 2: every segment of every iteration has a known start offset; walk them all in
// parallel, each recording what it leaves in positions 0..3, and check that each segment ends
// where the next one starts.
template <bool PAIRS>
__global__ void __launch_bounds__(64)
replay_segments_kernel(const uint32_t* __restrict__ X, uint32_t n, uint32_t steps, const RT* __restrict__ rt,
                       const unsigned long long* __restrict__ seg_tab, int iters, int nseg, ReplayCtl* ctl,
                       int4* __restrict__ seg_w) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= iters * nseg) return;
  const int t = i / nseg, sgm = i - t * nseg;
  const unsigned long long start = seg_tab[i];
  if (start == ~0ull) {
    seg_w[i] = make_int4(-2, -2, -2, -2);
    return;
  }
  const uint32_t k0 = (uint32_t)sgm * PANO_SEG_STEPS;
  const uint32_t k1 = min(steps, k0 + PANO_SEG_STEPS);
  int w[4];
  const uint32_t end = walk_track_segment<PAIRS>(X + start, 0u, n, k0, k1, rt, w);
  seg_w[i] = make_int4(w[0], w[1], w[2], w[3]);
  const unsigned long long next = (i + 1 < iters * nseg) ? seg_tab[i + 1] : ctl->base;
  if (next != ~0ull && start + end != next) atomicOr(&ctl->status, 4);  // passes disagree: never expected
}

// last writer wins across an iteration's segments
__global__ void combine_samples_kernel(const int4* __restrict__ seg_w, int iters, int nseg, int4* __restrict__ samples) {
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= iters) return;
  int a[4] = {-1, -1, -1, -1};
  for (int sgm = nseg - 1; sgm >= 0; sgm--) {
    const int4 w = seg_w[(size_t)t * nseg + sgm];
    if (w.x == -2) { a[0] = a[1] = a[2] = a[3] = -1; break; }
    if (a[0] < 0) a[0] = w.x;
    if (a[1] < 0) a[1] = w.y;
    if (a[2] < 0) a[2] = w.z;
    if (a[3] < 0) a[3] = w.w;
  }
  samples[t] = make_int4(a[0], a[1], a[2], a[3]);
}

// ---------------------------------------------------------------------------------------
// K5, resident formulation: ONE CTA replays all iterations strictly in order.  Every iteration
// starts from its exact stream offset, so the only unknown is the number of rejections so far
// inside the iteration; the cells of a narrow band of diagonals around its expectation are
// evaluated (about 20x fewer than the chunked replay's windows over start offsets), at the price
// of ~num_iterations sequential phases inside the CTA.  It occupies one SM, which is what makes
// it the throughput-mode replay: independent pairs on other lanes fill the rest of the GPU.
//   phase 1  lane = step, loop over the band's diagonals: ballot packs 32 steps of one diagonal
//            into a word (same bit layout as the chunked cell grid)
//   phase 2  one thread per (walk segment, entry diagonal): exit diagonal + diagonal at every block
//   phase 3  thread 0 chains the segments from diagonal 0 (exact; leaving the band is detected)
//   phase 4  thread per step: accepted draw -> swap targets; position p ends up holding the LARGEST
//            element index written to it (writes happen in increasing element order), the swaps
//            among the first four elements are replayed exactly by one thread
// The stream window of the next iteration is prefetched with cp.async while the current one runs.
// ---------------------------------------------------------------------------------------
constexpr int RES_THREADS = 32 * (int)RES_WARPS;

struct ResParams {
  const uint32_t* X;
  unsigned long long x_limit;   // readable stream words (generated + guard)
  const RT* rt;
  const ResBlock* blk;
  const uint32_t* seg_eoff;
  uint32_t n, steps, nkb, nwords, dmax, segb, nseg, n_entries, xcap;
  int iters;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
#ifdef PANO_CUDA_EMU
  memcpy(smem_dst, gsrc, 16);   // (CPU emulation tier: the copy completes at once; commit / wait groups are no-ops)
#else
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc) : "memory");
#endif
}

__global__ void __launch_bounds__(RES_THREADS, 2)
replay_resident_kernel(ResParams P, ReplayCtl* ctl, int4* __restrict__ samples) {
#ifdef PANO_CUDA_EMU
  uint8_t* res_smem = emu::dyn_smem();
#else
  extern __shared__ __align__(16) uint8_t res_smem[];
#endif
  uint32_t* Xs0 = reinterpret_cast<uint32_t*>(res_smem);
  uint32_t* Xs1 = Xs0 + P.xcap;
  uint32_t* bits = Xs1 + P.xcap;
  ResBlock* sblk = reinterpret_cast<ResBlock*>(bits + P.nwords + (P.nwords & 1u));
  uint4* s_p1 = reinterpret_cast<uint4*>(sblk + P.nkb + (P.nkb & 1u));   // phase-1 view: (x offset, word offset, w / 4, -)
  int2* s_segc = reinterpret_cast<int2*>(s_p1 + P.nkb);          // per segment: (entry base - dlo, dlo | w << 16)
  uint32_t* s_entry = reinterpret_cast<uint32_t*>(s_segc + P.nseg + 1);
  uint8_t* dtab = reinterpret_cast<uint8_t*>(s_entry + P.nseg + 1);
  __shared__ int s_slot[4], s_init[4];
  __shared__ RT s_rt4[4];
  __shared__ uint32_t s_rej;
  __shared__ int s_flag;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t steps = P.steps, odd = P.n & 1u, stride = P.segb + 1u;
  for (uint32_t i = tid; i < P.nkb; i += RES_THREADS) {
    const ResBlock B = P.blk[i];
    sblk[i] = B;
    s_p1[i] = make_uint4(i * 32u + B.dlo, B.woff, B.w >> 2, 0u);
  }
  for (uint32_t i = tid; i < P.nseg; i += RES_THREADS) {
    const ResBlock B0 = P.blk[i * P.segb];
    s_segc[i] = make_int2((int)P.seg_eoff[i] - (int)B0.dlo, (int)((uint32_t)B0.dlo | ((uint32_t)B0.w << 16)));
  }
  if (tid < 4) { s_slot[tid] = -1; s_rt4[tid] = P.rt[tid]; }   // (rt is padded by 8 entries)
  if (tid == 0) s_flag = 0;
  // this thread's (segment, entry diagonal) of phase 2: the plan is the same for every iteration
  int my_seg = -1;
  uint32_t my_d = 0;
  if ((uint32_t)tid < P.n_entries) {
    uint32_t sgm = 0;
    while (P.seg_eoff[sgm + 1] <= (uint32_t)tid) sgm++;
    my_seg = (int)sgm;
    my_d = P.blk[sgm * P.segb].dlo + ((uint32_t)tid - P.seg_eoff[sgm]);
  }
  // this lane's steps of phase 1 (blocks warp, warp + 32, ...): range / threshold stay in registers
  RT q[RES_MAXB];
#pragma unroll
  for (uint32_t i = 0; i < RES_MAXB; i++) {
    const uint32_t k = (warp + RES_WARPS * i) * 32u + lane;
    q[i].r = 2u; q[i].T = 0u;                // steps past the end never reject
    if (k < steps) q[i] = P.rt[k];
  }

  unsigned long long s = ctl->base;          // exact stream offset of the current iteration
  unsigned long long xb = s & ~3ull;         // stream offset of word 0 of the current window
  uint32_t* Xc = Xs0;
  uint32_t* Xn = Xs1;
  if (xb + P.xcap > P.x_limit) {
    if (tid == 0) atomicOr(&ctl->status, 2);
    return;
  }
  for (uint32_t c = tid; c < P.xcap / 4u; c += RES_THREADS) cp_async16(Xc + 4u * c, P.X + xb + 4u * c);
#ifndef PANO_CUDA_EMU
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
  __syncthreads();

  for (int t = 0; t < P.iters; t++) {
    const uint32_t rel = (uint32_t)(s - xb);
    // ---- prefetch: the next iteration starts in [s + steps, s + steps + dmax)
    const unsigned long long nb = (s + steps) & ~3ull;
    if (t + 1 < P.iters) {
      if (nb + P.xcap > P.x_limit) {
        if (tid == 0) { atomicOr(&ctl->status, 2); s_flag = 1; }
      } else {
        for (uint32_t c = tid; c < P.xcap / 4u; c += RES_THREADS) cp_async16(Xn + 4u * c, P.X + nb + 4u * c);
      }
    }
#ifndef PANO_CUDA_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif

    // ---- phase 1: rejection cells of the band (lane = step, one ballot word per diagonal, 4 per trip)
    {
      const uint32_t* xl = Xc + rel + lane;
      const bool lead = lane == 0;
#pragma unroll
      for (uint32_t i = 0; i < RES_MAXB; i++) {
        const uint32_t b = warp + RES_WARPS * i;
        if (b < P.nkb) {
          const uint4 B = s_p1[b];
          const uint32_t* px = xl + B.x;
          uint32_t* out = bits + B.y;
          const uint32_t r = q[i].r, T = q[i].T;
          for (uint32_t j4 = B.z; j4 != 0; j4--, px += 4, out += 4) {
            const uint32_t x0 = px[0], x1 = px[1], x2 = px[2], x3 = px[3];
            const uint32_t w0 = __ballot_sync(0xffffffffu, x0 * r < T);
            const uint32_t w1 = __ballot_sync(0xffffffffu, x1 * r < T);
            const uint32_t w2 = __ballot_sync(0xffffffffu, x2 * r < T);
            const uint32_t w3 = __ballot_sync(0xffffffffu, x3 * r < T);
            if (lead) { out[0] = w0; out[1] = w1; out[2] = w2; out[3] = w3; }
          }
        }
      }
    }
    __syncthreads();

    // ---- phase 2: segment walks for every entry diagonal
    if (my_seg >= 0) {
      const uint32_t b0 = (uint32_t)my_seg * P.segb, b1 = min(P.nkb, b0 + P.segb);
      uint8_t* out = dtab + (size_t)tid * stride;
      out[P.segb] = (uint8_t)res_walk_segment(bits, sblk, b0, b1, my_d, out);
    }
    __syncthreads();

    // ---- phase 3: chain from diagonal 0 (thread 0); exact swaps among the first four elements (thread 32)
    if (tid == 0) {
      uint32_t d = 0;
      bool bad = false;
      for (uint32_t sgm = 0; sgm < P.nseg; sgm++) {
        const int2 c = s_segc[sgm];
        const uint32_t dlo = (uint32_t)c.y & 0xffffu, w = (uint32_t)c.y >> 16;
        if (d - dlo >= w) { bad = true; break; }
        const uint32_t ent = (uint32_t)(c.x + (int)d);
        s_entry[sgm] = ent;
        d = dtab[(size_t)ent * stride + P.segb];
        if (d == RES_MISS) { bad = true; break; }
      }
      s_rej = d;
      if (bad) { atomicOr(&ctl->status, 1); s_flag = 1; }
    } else if (tid == 32) {
      int a0 = 0, a1 = 1, a2 = 2, a3 = 3;
      uint32_t o = rel;
      for (uint32_t k = 0; k < steps && 2u * k + odd < 4u; k++) {
        int a[4] = {a0, a1, a2, a3};
        track_step<true>(Xc, o, k, P.n, s_rt4, a);
        a0 = a[0]; a1 = a[1]; a2 = a[2]; a3 = a[3];
      }
      s_init[0] = a0; s_init[1] = a1; s_init[2] = a2; s_init[3] = a3;
    }
    __syncthreads();
    if (s_flag) break;

    // ---- phase 4: swap targets of every step's accepted draw (a warp works on one block at a time)
    {
      const uint32_t* xl = Xc + rel + lane;
      const uint32_t upto = lane == 31 ? ~0u : ((2u << lane) - 1u);
#pragma unroll 2
      for (uint32_t b = warp; b < P.nkb; b += RES_THREADS / 32) {
        const ResBlock B = sblk[b];
        uint32_t d = dtab[s_entry[B.seg] * stride + B.boff];      // diagonal on entering the block (warp-uniform)
        const uint32_t* wp = bits + B.woff - B.dlo;
        uint32_t m = wp[d] & upto;
        while (m) {                                               // rejections at steps <= mine: rare
          d++;
          m = wp[d] & upto & (~0u << (__ffs((int)m) - 1));
        }
        const uint32_t k = b * 32u + lane, idx = 2u * k + odd;
        const uint32_t x = xl[b * 32u + d];
        const uint32_t p1 = __umulhi(x, idx + 1u), p2 = __umulhi(x * (idx + 1u), idx + 2u);
        if (min(p1, p2) < 4u && k < steps && idx >= 4u) {
          if (p1 < 4u) atomicMax(&s_slot[p1], (int)idx);
          if (p2 < 4u) atomicMax(&s_slot[p2], (int)idx + 1);
        }
      }
    }
#ifndef PANO_CUDA_EMU
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
    __syncthreads();
    if (tid == 0) {
      int4 r;
      r.x = s_slot[0] >= 0 ? s_slot[0] : s_init[0];
      r.y = s_slot[1] >= 0 ? s_slot[1] : s_init[1];
      r.z = s_slot[2] >= 0 ? s_slot[2] : s_init[2];
      r.w = s_slot[3] >= 0 ? s_slot[3] : s_init[3];
      samples[t] = r;
      s_slot[0] = s_slot[1] = s_slot[2] = s_slot[3] = -1;
    }
    s += (unsigned long long)steps + s_rej;
    xb = nb;
    uint32_t* tmp = Xc; Xc = Xn; Xn = tmp;
  }
#ifndef PANO_CUDA_EMU
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
  if (tid == 0 && !s_flag) ctl->base = s;
}

// ---------------------------------------------------------------------------------------
// K6 / K7
// ---------------------------------------------------------------------------------------
__global__ void build_points_kernel(const int32_t* __restrict__ kp1, int n1, const int32_t* __restrict__ kp2, int n2,
                                    const pano_dmatch* __restrict__ m, int n, float4* __restrict__ pts,
                                    int* __restrict__ errw) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pano_dmatch mm = m[i];
  // n1 / n2 > 0: the caller's keypoint counts; an index outside them is reported, never dereferenced
  if (mm.query_idx < 0 || mm.train_idx < 0 || (n1 > 0 && mm.query_idx >= n1) || (n2 > 0 && mm.train_idx >= n2)) {
    atomicOr(errw, PANO_ERRW_BAD_INDEX);
    pts[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  pts[i] = make_float4((float)kp1[2 * mm.query_idx], (float)kp1[2 * mm.query_idx + 1],
                       (float)kp2[2 * mm.train_idx], (float)kp2[2 * mm.train_idx + 1]);
}

// K6: one warp per hypothesis.  Same arithmetic, in the same order, as pano_core.cuh
// find_homography4 / jacobi9 (= OpenCV's runKernel + JacobiImpl_), with the independent
// element updates of each Jacobi rotation spread over lanes 0..8:
//   S   full symmetric copy of OpenCV's upper-triangular working matrix A (S[a][b] == A[min][max])
//   rotation (k,l): for every i not in {k,l} OpenCV rotates the pair (A[.][k-side], A[.][l-side]),
//       which in S is always (S[i][k], S[i][l]) -> lane i; eigenvector columns -> lane i.
//   pivot: OpenCV's scan over |A[i][indR[i]]| (i = 0..7) then |A[indC[i]][i]| (i = 1..8) with a
//       strict '<' keeps the FIRST maximum -> warp arg-max over 16 candidates, ties to the lowest
//       scan position.  indR / indC are maintained exactly as OpenCV does (only rows k and l are
//       refreshed after a rotation, so stale entries behave identically).
constexpr int DLT_WARPS = 4;

struct DltSmem {
  double S[81];
  double V[81];
  double W[9];
  int indR[9], indC[9];
};

__device__ __forceinline__ void warp_argmax_first(double& v, int& pos, int& k, int& l) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    double v2 = __shfl_xor_sync(0xffffffffu, v, o);
    int p2 = __shfl_xor_sync(0xffffffffu, pos, o);
    int k2 = __shfl_xor_sync(0xffffffffu, k, o);
    int l2 = __shfl_xor_sync(0xffffffffu, l, o);
    if (v2 > v || (v2 == v && p2 < pos)) { v = v2; pos = p2; k = k2; l = l2; }
  }
}

__global__ void __launch_bounds__(DLT_WARPS * 32)
dlt_kernel(const float4* __restrict__ pts, const int4* __restrict__ samples, int iters, double* __restrict__ Hs,
           int* __restrict__ valid) {
  pdl_wait();
  __shared__ DltSmem sm_all[DLT_WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int t = blockIdx.x * DLT_WARPS + wid;
  if (t >= iters) return;
  DltSmem& sm = sm_all[wid];
  const int4 smp = samples[t];
  if (smp.x < 0 || smp.y < 0 || smp.z < 0 || smp.w < 0) {  // replay did not resolve this iteration
    if (lane == 0) valid[t] = 0;
    if (lane < 9) Hs[(size_t)t * 9 + lane] = 0.0;
    return;
  }
  const int idx[4] = {smp.x, smp.y, smp.z, smp.w};
  float M[8], m[8];  // M = src (query side), m = dst
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float4 p = pts[idx[j]];
    M[2 * j] = p.x; M[2 * j + 1] = p.y;
    m[2 * j] = p.z; m[2 * j + 1] = p.w;
  }
  // ---- normalisation (every lane computes the same scalars) -----------------------------
  const int count = 4;
  double cMx = 0, cMy = 0, cmx = 0, cmy = 0, sMx = 0, sMy = 0, smx = 0, smy = 0;
#pragma unroll
  for (int i = 0; i < count; i++) {
    cmx = __dadd_rn(cmx, (double)m[2 * i]);
    cmy = __dadd_rn(cmy, (double)m[2 * i + 1]);
    cMx = __dadd_rn(cMx, (double)M[2 * i]);
    cMy = __dadd_rn(cMy, (double)M[2 * i + 1]);
  }
  cmx = __ddiv_rn(cmx, 4.0); cmy = __ddiv_rn(cmy, 4.0);
  cMx = __ddiv_rn(cMx, 4.0); cMy = __ddiv_rn(cMy, 4.0);
#pragma unroll
  for (int i = 0; i < count; i++) {
    smx = __dadd_rn(smx, fabs(__dsub_rn((double)m[2 * i], cmx)));
    smy = __dadd_rn(smy, fabs(__dsub_rn((double)m[2 * i + 1], cmy)));
    sMx = __dadd_rn(sMx, fabs(__dsub_rn((double)M[2 * i], cMx)));
    sMy = __dadd_rn(sMy, fabs(__dsub_rn((double)M[2 * i + 1], cMy)));
  }
  if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON) {
    if (lane == 0) valid[t] = 0;
    if (lane < 9) Hs[(size_t)t * 9 + lane] = 0.0;
    return;
  }
  smx = __ddiv_rn(4.0, smx); smy = __ddiv_rn(4.0, smy);
  sMx = __ddiv_rn(4.0, sMx); sMy = __ddiv_rn(4.0, sMy);
  // ---- LtL: lane j accumulates row j (k >= j), points in order --------------------------
  if (lane < 9) {
    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; k++) acc[k] = 0.0;
#pragma unroll
    for (int i = 0; i < count; i++) {
      double x = __dmul_rn(__dsub_rn((double)m[2 * i], cmx), smx);
      double y = __dmul_rn(__dsub_rn((double)m[2 * i + 1], cmy), smy);
      double X = __dmul_rn(__dsub_rn((double)M[2 * i], cMx), sMx);
      double Y = __dmul_rn(__dsub_rn((double)M[2 * i + 1], cMy), sMy);
      double Lx[9] = {X, Y, 1, 0, 0, 0, __dmul_rn(-x, X), __dmul_rn(-x, Y), -x};
      double Ly[9] = {0, 0, 0, X, Y, 1, __dmul_rn(-y, X), __dmul_rn(-y, Y), -y};
      double lxj = 0, lyj = 0;
#pragma unroll
      for (int k = 0; k < 9; k++)
        if (k == lane) { lxj = Lx[k]; lyj = Ly[k]; }
#pragma unroll
      for (int k = 0; k < 9; k++)
        if (k >= lane) acc[k] = __dadd_rn(acc[k], __dadd_rn(__dmul_rn(lxj, Lx[k]), __dmul_rn(lyj, Ly[k])));
    }
#pragma unroll
    for (int k = 0; k < 9; k++)
      if (k >= lane) { sm.S[lane * 9 + k] = acc[k]; sm.S[k * 9 + lane] = acc[k]; }
#pragma unroll
    for (int k = 0; k < 9; k++) sm.V[lane * 9 + k] = (k == lane) ? 1.0 : 0.0;
  }
  __syncwarp();
  // ---- Jacobi init -----------------------------------------------------------------------
  if (lane < 9) {
    const int k = lane;
    sm.W[k] = sm.S[10 * k];
    if (k < 8) {
      int mm = k + 1;
      double mv = fabs(sm.S[9 * k + mm]);
      for (int i = k + 2; i < 9; i++) {
        double val = fabs(sm.S[9 * k + i]);
        if (mv < val) mv = val, mm = i;
      }
      sm.indR[k] = mm;
    }
    if (k > 0) {
      int mm = 0;
      double mv = fabs(sm.S[k]);
      for (int i = 1; i < k; i++) {
        double val = fabs(sm.S[9 * i + k]);
        if (mv < val) mv = val, mm = i;
      }
      sm.indC[k] = mm;
    }
  }
  __syncwarp();
  const int maxIters = 9 * 9 * 30;
  for (int iter = 0; iter < maxIters; iter++) {
    // pivot: candidates 0..7 from indR, 8..15 from indC (scan order), first maximum wins
    double v = -1.0;
    int pos = 64, k = 0, l = 0;
    if (lane < 8) {
      k = lane; l = sm.indR[lane];
      v = fabs(sm.S[9 * k + l]); pos = lane;
    } else if (lane < 16) {
      const int i = lane - 7;
      k = sm.indC[i]; l = i;
      v = fabs(sm.S[9 * k + l]); pos = lane;
    }
    warp_argmax_first(v, pos, k, l);
    k = __shfl_sync(0xffffffffu, k, 0);
    l = __shfl_sync(0xffffffffu, l, 0);
    const double p = sm.S[9 * k + l];
    if (fabs(p) <= DBL_EPSILON) break;
    const double Wk = sm.W[k], Wl = sm.W[l];
    double y = __dmul_rn(__dsub_rn(Wl, Wk), 0.5);
    double tt = __dadd_rn(fabs(y), cv_hypot(p, y));
    double s = cv_hypot(p, tt);
    double c = __ddiv_rn(tt, s);
    s = __ddiv_rn(p, s);
    tt = __dmul_rn(__ddiv_rn(p, tt), p);
    if (y < 0) s = -s, tt = -tt;
    __syncwarp();
    if (lane == 0) {
      sm.S[9 * k + l] = 0; sm.S[9 * l + k] = 0;
      sm.W[k] = __dsub_rn(Wk, tt);
      sm.W[l] = __dadd_rn(Wl, tt);
    }
    if (lane < 9) {
      const int i = lane;
      if (i != k && i != l) {
        double a0 = sm.S[9 * i + k], b0 = sm.S[9 * i + l];
        double n0 = __dsub_rn(__dmul_rn(a0, c), __dmul_rn(b0, s));
        double n1 = __dadd_rn(__dmul_rn(a0, s), __dmul_rn(b0, c));
        sm.S[9 * i + k] = n0; sm.S[9 * k + i] = n0;
        sm.S[9 * i + l] = n1; sm.S[9 * l + i] = n1;
      }
      double a0 = sm.V[9 * k + i], b0 = sm.V[9 * l + i];
      sm.V[9 * k + i] = __dsub_rn(__dmul_rn(a0, c), __dmul_rn(b0, s));
      sm.V[9 * l + i] = __dadd_rn(__dmul_rn(a0, s), __dmul_rn(b0, c));
    }
    __syncwarp();
    // refresh indR / indC of rows k and l only (as OpenCV does): four independent "first maximum" scans of at most
    // 8 elements, one per group of 8 lanes (group 0: indR[k], 1: indR[l], 2: indC[k], 3: indC[l]), reduced with
    // three xor-shuffles inside the group; ties keep the lowest index = OpenCV's strict '<' scan
    {
      const int grp = lane >> 3, j = lane & 7;
      const int idx = (grp & 1) ? l : k;
      const bool rowscan = grp < 2;
      // row scan: elements idx+1 .. 8 of row idx; column scan: elements 0 .. idx-1 of column idx (S is symmetric)
      const int e = rowscan ? idx + 1 + j : j;
      const bool valid = rowscan ? (idx < 8 && e < 9) : (idx > 0 && e < idx);
      double v = valid ? fabs(sm.S[9 * idx + (valid ? e : 0)]) : -1.0;
      int pos = j;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
        const int p2 = __shfl_xor_sync(0xffffffffu, pos, o);
        if (v2 > v || (v2 == v && p2 < pos)) { v = v2; pos = p2; }
      }
      if (j == 0) {
        if (rowscan) { if (idx < 8) sm.indR[idx] = idx + 1 + pos; }
        else         { if (idx > 0) sm.indC[idx] = pos; }
      }
    }
    __syncwarp();
  }
  __syncwarp();
  // ---- eigenvalue selection sort (descending, first maximum) -> row of the smallest ------
  if (lane == 0) {
    double Wv[9];
    int perm[9];
    for (int i = 0; i < 9; i++) { Wv[i] = sm.W[i]; perm[i] = i; }
    for (int k = 0; k < 8; k++) {
      int mm = k;
      for (int i = k + 1; i < 9; i++)
        if (Wv[mm] < Wv[i]) mm = i;
      if (k != mm) {
        double tw = Wv[mm]; Wv[mm] = Wv[k]; Wv[k] = tw;
        int tp = perm[mm]; perm[mm] = perm[k]; perm[k] = tp;
      }
    }
    const double* h0 = &sm.V[9 * perm[8]];
    double invHnorm[9] = {__ddiv_rn(1., smx), 0, cmx, 0, __ddiv_rn(1., smy), cmy, 0, 0, 1};
    double Hnorm2[9] = {sMx, 0, __dmul_rn(-cMx, sMx), 0, sMy, __dmul_rn(-cMy, sMy), 0, 0, 1};
    double Htemp[9], H0[9];
    mul33(invHnorm, h0, Htemp);
    mul33(Htemp, Hnorm2, H0);
    double sc = __ddiv_rn(1., H0[8]);
    for (int i = 0; i < 9; i++) Hs[(size_t)t * 9 + i] = __dmul_rn(H0[i], sc);
    valid[t] = 1;
  }
}

__global__ void __launch_bounds__(256)
score_kernel(const float4* __restrict__ pts, int m, const double* __restrict__ Hs, const int* __restrict__ valid,
             double thr, int* __restrict__ counts) {
  pdl_wait();
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
    c += is_inlier_lim(H, p.x, p.y, p.z, p.w, thr) ? 1 : 0;   // thr = inlier_d2_limit(distance threshold)
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

