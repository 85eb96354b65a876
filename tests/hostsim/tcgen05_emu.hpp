// tcgen05_emu.hpp — a host MODEL of the Blackwell machinery the tensor-core matcher drives through inline PTX, for the
// CPU emulation tier (include after cuda_emu.hpp, before csrc/match_tc_kernels.cuh; TEST INFRASTRUCTURE ONLY).
//
// It provides, under the names and signatures of the kernel's PTX wrappers:
//   mbarriers      init / arrive / arrive.expect_tx / complete_tx / test_wait.parity with phases: a phase completes when
//                  the pending arrival count AND the outstanding transaction bytes reach zero; test_wait(parity P) is
//                  true once the phase of parity P has completed (so a fresh barrier passes a wait on parity 1);
//   TMA            cp.async.bulk.tensor.2d of a (128 bytes x box_rows) box with SWIZZLE_128B - rows outside the tensor
//                  are zero-filled, 16-byte chunk c of row r lands at chunk (c xor (r mod 8)) of its 128-byte line, 8-row
//                  groups 1024 bytes apart - and the 1-D bulk copy; both complete their bytes on the mbarrier at once;
//   tcgen05.mma    kind::i8, M = 128, N = 128, K = 32 per instruction, u8 x u8 -> s32, operands read from shared memory
//                  through K-major SWIZZLE_128B descriptors (start address >> 4 in the low 14 bits; the K step is the
//                  descriptor's byte offset inside the 128-byte line), accumulator in tensor memory at (lane = row,
//                  column = base column + n); executed synchronously when issued;
//   tcgen05.commit an arrival on the mbarrier (the MMAs above have already completed);
//   tcgen05.ld     32x32b.x32: the calling thread's lane of its warp's 32-lane quarter, 32 consecutive columns;
//   tensor memory  128 lanes x 512 columns of 32 bits, one CTA at a time; alloc returns address 0.
// Waits are polling loops that yield to the other fibers; every state change counts as progress for the emulation's
// deadlock detection, so a pipeline that cannot advance is reported instead of hanging.
// The model is self-consistent (what the TMA model writes is what the MMA model reads).  Whether the hardware agrees
// with it is established on the GPU, where the default instantiation of the same kernel body is validated against the
// SIMT matcher and the oracle.
#pragma once
#include <cstdint>
#include <cstring>

namespace pano {

struct MBarModel {   // the 8 bytes of an mbarrier object in shared memory
  int32_t tx;        // outstanding transaction bytes
  uint16_t pending;  // arrivals still expected in this phase
  uint8_t expected;  // arrival count per phase
  uint8_t phase;     // parity of the phase in progress
};
static_assert(sizeof(MBarModel) == 8, "mbarrier object");

inline uint32_t tmem_model[128][512];

inline uint32_t smem_u32(const void* p) {
  return (uint32_t)(reinterpret_cast<const uint8_t*>(p) - emu::dyn_smem());
}
inline void mbar_check(MBarModel* b) {
  if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = b->expected; }
  emu::S->progress++;
}
inline void mbar_init(unsigned long long* bar, uint32_t count) {
  MBarModel* b = reinterpret_cast<MBarModel*>(bar);
  b->tx = 0; b->pending = (uint16_t)count; b->expected = (uint8_t)count; b->phase = 0;
}
inline void mbar_arrive(unsigned long long* bar) {
  MBarModel* b = reinterpret_cast<MBarModel*>(bar);
  b->pending--;
  mbar_check(b);
}
// Fault injection (tests only): PANO_EMU_TC_DROP_LOAD=N makes the TMA model drop the N-th tile load of the next matcher
// kernel - no data, no complete_tx, so that stage's barrier never completes, like a lost transaction would on the
// device.  From then on the waits are bounded the way the device's are (a wait that keeps failing raises the abort flag
// and gives up), so that the host's recovery path - error word, failed call, tensor-core matcher disabled, SIMT matcher
// taking over - can be exercised.
inline long emu_tc_drop_countdown = -1;
inline bool emu_tc_fault = false;
inline bool mbar_wait(unsigned long long* bar, uint32_t parity, volatile int* abort_flag) {
  const MBarModel* b = reinterpret_cast<const MBarModel*>(bar);
  unsigned spins = 0;
  while (b->phase == parity) {          // the phase of parity `parity` has not completed yet
    if (*abort_flag) return false;
    if (emu_tc_fault) {                 // (bounded like the device's SPIN_LIMIT; polling counts as activity here)
      emu::S->progress++;
      if (++spins > 4096) { *abort_flag = 1; return false; }
    }
    emu::yield();
  }
  emu::S->progress++;
  return true;
}
inline void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  MBarModel* b = reinterpret_cast<MBarModel*>(bar);
  b->tx += (int32_t)bytes;
  b->pending--;
  mbar_check(b);
}
inline void mbar_complete_tx(unsigned long long* bar, uint32_t bytes) {
  MBarModel* b = reinterpret_cast<MBarModel*>(bar);
  b->tx -= (int32_t)bytes;
  mbar_check(b);
}
// byte k of row `row` of a K-major SWIZZLE_128B tile whose descriptor / TMA destination starts at byte offset `start`
inline uint32_t swizzle128_offset(uint32_t start, int row, int k) {
  const uint32_t tile = start & ~1023u;
  const int kk = (int)(start & 127u) + k;
  return tile + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((((kk >> 4) ^ (row & 7)) & 7) << 4) +
         (uint32_t)(kk & 15);
}
inline void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int row0, unsigned long long* bar) {
  if (emu_tc_drop_countdown >= 0 && emu_tc_drop_countdown-- == 0) { emu_tc_fault = true; return; }   // (fault injection)
  uint8_t* smem = emu::dyn_smem();
  const uint32_t start = smem_u32(smem_dst);
  for (int r = 0; r < (int)map->box_rows; r++)
    for (int k = 0; k < 128; k++) {
      const long long R = (long long)row0 + r, Cc = (long long)c0 + k;
      uint8_t v = 0;
      if (R >= 0 && R < (long long)map->rows && Cc >= 0 && Cc < (long long)map->row_bytes) v = map->base[(size_t)R * map->pitch + (size_t)Cc];
      smem[swizzle128_offset(start, r, k)] = v;
    }
  mbar_complete_tx(bar, 128u * map->box_rows);
}
inline void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
  memcpy(smem_dst, gsrc, bytes);
  mbar_complete_tx(bar, bytes);
}
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void fence_async_smem() {}
inline void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  const uint8_t* smem = emu::dyn_smem();
  const uint32_t a0 = (uint32_t)(adesc & 0x3FFFu) << 4, b0 = (uint32_t)(bdesc & 0x3FFFu) << 4;
  const uint32_t col0 = d_tmem & 0xFFFFu, lane0 = d_tmem >> 16;
  uint8_t A[128][32], B[128][32];
  for (int r = 0; r < 128; r++)
    for (int k = 0; k < 32; k++) {
      A[r][k] = smem[swizzle128_offset(a0, r, k)];
      B[r][k] = smem[swizzle128_offset(b0, r, k)];
    }
  for (int m = 0; m < 128; m++)
    for (int n = 0; n < 128; n++) {
      uint32_t acc = accumulate ? tmem_model[lane0 + m][col0 + n] : 0u;
      for (int k = 0; k < 32; k++) acc += (uint32_t)A[m][k] * B[n][k];
      tmem_model[lane0 + m][col0 + n] = acc;
    }
  emu::S->progress++;
}
inline void mma_commit(unsigned long long* bar) { mbar_arrive(bar); }
inline void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  const uint32_t lane = (taddr >> 16) + (emu::S->cur->tid & 31u), col = taddr & 0xFFFFu;
  for (int i = 0; i < 32; i++) r[i] = tmem_model[lane][col + i];
}
inline void tmem_ld_wait() {}
inline void tmem_alloc_emu(uint32_t* slot, uint32_t) {
  *slot = 0;
  if (emu::S->cur->tid % 32 == 0 && blockIdx.x == 0) {      // once per launch: arm the fault injection, if asked for
    const char* e = getenv("PANO_EMU_TC_DROP_LOAD");
    emu_tc_drop_countdown = e ? atol(e) : -1;
    emu_tc_fault = false;
  }
}

}  // namespace pano
