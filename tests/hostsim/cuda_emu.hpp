// cuda_emu.hpp — a small CPU emulation of the CUDA execution model, for the no-GPU test tier.
//
// TEST INFRASTRUCTURE ONLY (never linked into the product).  It lets g++ compile a header of real __global__
// kernels (csrc/knn_kernels.cuh) unchanged and run them block by block on the host:
//   * every CUDA thread of a block is a fiber (ucontext) with its own stack; threadIdx / blockIdx / blockDim /
//     gridDim are globals that the scheduler sets whenever it resumes a fiber;
//   * __syncthreads() and the warp collectives (__shfl_sync, __shfl_xor_sync, __shfl_down_sync, __ballot_sync,
//     __syncwarp) are REAL rendezvous points: a fiber yields until every live participant has arrived, values are
//     exchanged through per-warp slots.  A barrier that can never complete (divergent participants) is reported as
//     an error instead of hanging;
//   * __shared__ becomes `static` (blocks run one after another); a thread that returns early counts as "exited"
//     for later barriers, as on the hardware;
//   * atomics are plain read-modify-write (one fiber runs at a time); blocks of a grid run in the order given by
//     emu::Launch::block_order (forward, reverse or shuffled) so that order-dependent results show up.
// What it cannot show: data races between threads of a warp that the hardware runs in lockstep, memory-model
// effects, launch-configuration limits (registers, shared memory) - ptxas -v and the GPU tier cover those.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define PANO_CUDA_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__

struct uint4 {
  uint32_t x, y, z, w;
} __attribute__((aligned(16)));
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct int4 {
  int x, y, z, w;
} __attribute__((aligned(16)));

static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
struct int2 {
  int x, y;
} __attribute__((aligned(8)));
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
struct float4 {
  float x, y, z, w;
} __attribute__((aligned(16)));
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct emu_idx {
  unsigned x, y, z;
};
inline emu_idx threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

using std::max;
using std::min;

namespace emu {

// Context switch between fibers.  On x86-64 a dozen instructions (callee-saved registers + stack pointer; the
// kernels neither change the FP control state nor throw), because a block of 1024 threads meeting at a few thousand
// barriers switches millions of times and ucontext's swapcontext costs a system call each; ucontext elsewhere.
#if defined(__x86_64__)
#define PANO_EMU_FAST_SWITCH 1
extern "C" void pano_emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.p2align 4
.weak pano_emu_switch
.type pano_emu_switch,@function
pano_emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size pano_emu_switch,.-pano_emu_switch
)");
#endif

struct Fiber {
#ifdef PANO_EMU_FAST_SWITCH
  void* sp = nullptr;
#else
  ucontext_t ctx;
#endif
  char* stack = nullptr;       // (malloc'ed, not zero-filled: 1024 threads x 256 KB would cost more than the kernel)
  size_t stack_size = 0;
  bool done = false;
  unsigned tid = 0;
};

struct Warp {
  unsigned arrived = 0;   // lanes waiting at a warp rendezvous
  unsigned released = 0;  // lanes whose rendezvous has completed and that have not resumed yet
  unsigned alive = 0;
  uint64_t slot[32];
};

struct State {
#ifdef PANO_EMU_FAST_SWITCH
  void* sched_sp = nullptr;
#else
  ucontext_t sched;
#endif
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  Fiber* cur = nullptr;
  unsigned block_arrived = 0, block_gen = 0, live = 0;
  std::function<void()> body;
  const char* error = nullptr;
  uint64_t progress = 0;
  std::vector<uint8_t> dyn;      // the launch's dynamic shared memory (one block at a time)
  int count_acc = 0;             // __syncthreads_count accumulator
};
inline State* S = nullptr;

// the dynamic shared memory of the running block (128-byte aligned like the kernels ask for)
inline uint8_t* dyn_smem() {
  uintptr_t a = reinterpret_cast<uintptr_t>(S->dyn.data());
  return reinterpret_cast<uint8_t*>((a + 1023) & ~uintptr_t(1023));
}

#ifdef PANO_EMU_FAST_SWITCH
inline void to_scheduler(Fiber* f) { pano_emu_switch(&f->sp, S->sched_sp); }
inline void to_fiber(Fiber* f) { pano_emu_switch(&S->sched_sp, f->sp); }
#else
inline void to_scheduler(Fiber* f) { swapcontext(&f->ctx, &S->sched); }
inline void to_fiber(Fiber* f) { swapcontext(&S->sched, &f->ctx); }
#endif
inline void yield() { to_scheduler(S->cur); }

inline void trampoline() {
  S->body();
  Fiber* f = S->cur;
  f->done = true;
  S->live--;
  S->warps[f->tid >> 5].alive &= ~(1u << (f->tid & 31));
  S->progress++;
  to_scheduler(f);   // never resumed
}

inline void prepare_fiber(Fiber& f) {
#ifdef PANO_EMU_FAST_SWITCH
  // initial frame: six callee-saved registers, the entry point as return address, a null return address above it
  // (the entry never returns); the stack pointer is 8 modulo 16 when the entry starts, as after a call
  uintptr_t top = (reinterpret_cast<uintptr_t>(f.stack) + f.stack_size) & ~uintptr_t(15);
  void** sp = reinterpret_cast<void**>(top);
  *--sp = nullptr;
  *--sp = reinterpret_cast<void*>(&trampoline);
  for (int i = 0; i < 6; i++) *--sp = nullptr;
  f.sp = sp;
#else
  getcontext(&f.ctx);
  f.ctx.uc_stack.ss_sp = f.stack;
  f.ctx.uc_stack.ss_size = f.stack_size;
  f.ctx.uc_link = &S->sched;
  makecontext(&f.ctx, (void (*)())trampoline, 0);
#endif
}

inline void sync_block() {
  State& s = *S;
  const unsigned gen = s.block_gen;
  s.block_arrived++;
  s.progress++;
  // (exited threads no longer count: the barrier completes when every live thread has arrived)
  while (s.block_gen == gen) {
    if (s.block_arrived >= s.live) { s.block_arrived = 0; s.block_gen++; s.progress++; break; }
    yield();
  }
}

// rendezvous of the live lanes of `mask` (several disjoint groups of one warp may meet independently: a lane is
// released by the completion of ITS group only)
inline void sync_warp(unsigned mask) {
  State& s = *S;
  Warp& w = s.warps[s.cur->tid >> 5];
  const unsigned lane = s.cur->tid & 31, bit = 1u << lane;
  w.arrived |= bit;
  s.progress++;
  for (;;) {
    if (w.released & bit) { w.released &= ~bit; break; }
    const unsigned group = mask & w.alive;
    if ((w.arrived & group) == group) {       // the last one in: release the others, go on
      w.arrived &= ~group;
      w.released |= group & ~bit;
      s.progress++;
      break;
    }
    yield();
  }
}

template <typename T>
inline T exchange(unsigned mask, T v, unsigned src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  State& s = *S;
  Warp& w = s.warps[s.cur->tid >> 5];
  const unsigned lane = s.cur->tid & 31;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  w.slot[lane] = raw;
  sync_warp(mask);
  // a source lane outside the mask / already exited returns the caller's own value (undefined on hardware)
  uint64_t got = ((mask >> (src_lane & 31)) & 1u) && ((w.alive >> (src_lane & 31)) & 1u) ? w.slot[src_lane & 31] : raw;
  sync_warp(mask);   // nobody overwrites a slot before everyone has read
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}

// bar.sync id, count: a barrier among `count` threads of the block, identified by `id`
struct NamedBarrier { unsigned arrived = 0, gen = 0; };
inline NamedBarrier named[16];
inline void named_barrier(int id, unsigned count) {
  NamedBarrier& b = named[id & 15];
  const unsigned gen = b.gen;
  b.arrived++;
  S->progress++;
  while (b.gen == gen) {
    if (b.arrived >= count) { b.arrived = 0; b.gen++; S->progress++; break; }
    yield();
  }
}

enum Order { FORWARD = 0, REVERSE = 1, SHUFFLED = 2 };

// runs `body` (a call of a __global__ function) for every thread of every block of the grid
inline const char* launch(dim3 grid, dim3 block, const std::function<void()>& body, int order = FORWARD,
                          size_t dyn_smem_bytes = 0, size_t stack_bytes = 256 * 1024) {
  const unsigned nthreads = block.x * block.y * block.z;
  State st;
  S = &st;
  st.body = body;
  st.dyn.assign(dyn_smem_bytes + 1024, 0xCD);   // (not zero: shared memory is uninitialised on the device)
  gridDim = grid;
  blockDim = block;
  std::vector<unsigned> blocks(grid.x * grid.y * grid.z);
  for (unsigned i = 0; i < blocks.size(); i++) blocks[i] = i;
  if (order == REVERSE) std::reverse(blocks.begin(), blocks.end());
  if (order == SHUFFLED) {
    uint32_t x = 2463534242u;
    for (size_t i = blocks.size(); i > 1; i--) {
      x ^= x << 13; x ^= x >> 17; x ^= x << 5;
      std::swap(blocks[i - 1], blocks[x % i]);
    }
  }
  st.fibers.resize(nthreads);
  char* stacks = static_cast<char*>(malloc((size_t)nthreads * stack_bytes));
  for (unsigned t = 0; t < nthreads; t++) { st.fibers[t].stack = stacks + (size_t)t * stack_bytes; st.fibers[t].stack_size = stack_bytes; }
  for (unsigned b : blocks) {
    const emu_idx bi = {b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y)};
    st.warps.assign((nthreads + 31) / 32, Warp());
    st.block_arrived = 0;
    st.live = nthreads;
    for (auto& nb : named) nb = NamedBarrier();
    for (unsigned t = 0; t < nthreads; t++) {
      Fiber& f = st.fibers[t];
      f.done = false;
      f.tid = t;
      st.warps[t >> 5].alive |= 1u << (t & 31);
      prepare_fiber(f);
    }
    // PANO_EMU_THREAD_ORDER: the order in which the runnable threads of a block get their turn between two rendezvous
    // points - 0 ascending (default), 1 descending, 2 a new pseudo-random permutation every round.  A kernel whose result
    // depends on it has a race (a missing __syncthreads / __syncwarp, an assumption of warp lockstep): the emulation
    // tests are run under all three (tools/emu_sanitize.sh).
    static const int thread_order = [] { const char* e = getenv("PANO_EMU_THREAD_ORDER"); return e ? atoi(e) : 0; }();
    std::vector<unsigned> turn(nthreads);
    for (unsigned t = 0; t < nthreads; t++) turn[t] = thread_order == 1 ? nthreads - 1 - t : t;
    uint32_t rs = 88172645u + b;
    while (st.live > 0) {
      const uint64_t before = st.progress;
      if (thread_order == 2)
        for (unsigned i = nthreads; i > 1; i--) {
          rs ^= rs << 13; rs ^= rs >> 17; rs ^= rs << 5;
          std::swap(turn[i - 1], turn[rs % i]);
        }
      for (unsigned ti = 0; ti < nthreads; ti++) {
        const unsigned t = turn[ti];
        Fiber& f = st.fibers[t];
        if (f.done) continue;
        st.cur = &f;
        threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
        blockIdx = bi;
        to_fiber(&f);
      }
      if (st.progress == before && st.live > 0) {   // a full round in which nobody moved: a barrier cannot complete
        S = nullptr;
        free(stacks);
        return "deadlock: a barrier / warp collective is waiting for threads that never arrive";
      }
    }
  }
  S = nullptr;
  free(stacks);
  return nullptr;
}

}  // namespace emu

// ---- the CUDA device API surface the kernels use -----------------------------------------------------------------
inline void __syncthreads() { emu::sync_block(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::sync_warp(mask); }
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src_lane) { return emu::exchange(mask, v, (unsigned)src_lane); }
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
  return emu::exchange(mask, v, (emu::S->cur->tid & 31) ^ (unsigned)lane_mask);
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
  const unsigned lane = emu::S->cur->tid & 31;
  return emu::exchange(mask, v, lane + delta < 32 ? lane + delta : lane);
}
// every lane of the mask contributes its predicate; one exchange round (the slots hold the predicates)
inline unsigned __ballot_sync(unsigned mask, int pred) {
  emu::State& s = *emu::S;
  emu::Warp& w = s.warps[s.cur->tid >> 5];
  const unsigned lane = s.cur->tid & 31;
  w.slot[lane] = pred ? 1u : 0u;
  emu::sync_warp(mask);
  unsigned r = 0;
  for (unsigned l = 0; l < 32; l++)
    if (((mask >> l) & 1u) && ((w.alive >> l) & 1u) && w.slot[l]) r |= 1u << l;
  emu::sync_warp(mask);
  return r;
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
  const unsigned lane = emu::S->cur->tid & 31;
  return emu::exchange(mask, v, lane >= delta ? lane - delta : lane);
}
// number of threads of the block whose predicate is non-zero (a barrier like __syncthreads)
inline int __syncthreads_count(int pred) {
  emu::State& s = *emu::S;
  if (pred) s.count_acc++;
  emu::sync_block();                       // every live thread has added its predicate
  const int r = s.count_acc;
  emu::sync_block();                       // every thread has read the sum
  s.count_acc = 0;                         // (idempotent; nobody adds again before the third barrier)
  emu::sync_block();
  return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
// separately rounded FP64 operations (this translation unit is built with -ffp-contract=off: no FMA is formed)
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline unsigned __vabsdiffu4(unsigned a, unsigned b) {
  unsigned r = 0;
  for (int i = 0; i < 4; i++) {
    const int x = (a >> (8 * i)) & 255, y = (b >> (8 * i)) & 255;
    r |= (unsigned)(x > y ? x - y : y - x) << (8 * i);
  }
  return r;
}
inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
  for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 255) * ((b >> (8 * i)) & 255);
  return c;
}
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
  const unsigned long long old = *p;
  if (v < old) *p = v;
  return old;
}
inline int atomicOr(int* p, int v) {
  const int old = *p;
  *p = old | v;
  return old;
}
inline uint32_t atomicOr(uint32_t* p, uint32_t v) {
  const uint32_t old = *p;
  *p = old | v;
  return old;
}
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) {
  const uint32_t old = *p;
  *p = old + v;
  return old;
}

// ---- a model of the TMA unit's 2-D byte-tensor tile load (cp.async.bulk.tensor.2d, no swizzle) --------------------
// CUtensorMap here is the host-side description the real encoder would be given: the tile is box_bytes x box_rows
// from (byte coordinate c0, row coordinate r0); elements outside [0, row_bytes) x [0, rows) are zero-filled, as the
// hardware does for out-of-bounds coordinates (negative ones included).
struct CUtensorMap {
  const uint8_t* base;
  uint64_t row_bytes, rows, pitch;
  uint32_t box_rows = 0;   // (swizzled tile maps of the matcher model, tcgen05_emu.hpp)
};
namespace emu {
inline void tma_tile_load_2d(uint8_t* dst, const CUtensorMap& m, int c0, int r0, int box_bytes, int box_rows) {
  for (int r = 0; r < box_rows; r++)
    for (int c = 0; c < box_bytes; c++) {
      const long long R = (long long)r0 + r, Cc = (long long)c0 + c;
      uint8_t v = 0;
      if (R >= 0 && R < (long long)m.rows && Cc >= 0 && Cc < (long long)m.row_bytes) v = m.base[(size_t)R * m.pitch + (size_t)Cc];
      dst[(size_t)r * box_bytes + c] = v;
    }
}
}  // namespace emu

// ---- further device intrinsics (warp kernels) -----------------------------------------------------------------------
#include <cmath>
// low 32 bits of (hi:lo) >> (shift mod 32)
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
  const unsigned s = shift & 31u;
  return s ? (lo >> s) | (hi << (32u - s)) : lo;
}
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline int __float2int_rn(float v) { return (int)lrintf(v); }   // round to nearest even (default rounding mode)
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline int atomicMax(int* p, int v) {
  const int old = *p;
  if (v > old) *p = v;
  return old;
}
