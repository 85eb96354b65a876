// match_tc_kernels.cuh — the device code of match_tc.cu: the tensor-core matcher (tcgen05 / TMEM / TMA pipeline with
// the fused arg-min epilogue, and its top-2 variant for pano_match_knn) and the column-constant kernel.
//
// Included by match_tc.cu INSIDE `namespace pano { namespace {` (no includes or namespaces of its own), and by the CPU
// emulation tier (tests/hostsim/match_tc_emu.cpp), which compiles the same kernel body with g++ on
// tests/hostsim/cuda_emu.hpp plus tests/hostsim/tcgen05_emu.hpp - a host MODEL of the pieces of Blackwell the kernel
// drives through inline PTX: mbarriers (phases, expect-tx / complete-tx), the TMA 2-D tile load with 128-byte swizzle,
// the 1-D bulk copy, tcgen05.mma kind::i8 on K-major swizzled shared-memory descriptors into tensor memory,
// tcgen05.commit, tcgen05.ld 32x32b.x32, named barriers.  Under PANO_CUDA_EMU the PTX wrappers below are replaced by
// that model (same names and signatures); everything else - tile / super-tile scheduling, pipeline stages and phases,
// who waits for and who releases what, descriptors, the epilogue's key arithmetic and reductions, the cross-CTA merge -
// is the product's code.  The model is consistent with itself (the TMA model writes the swizzle the MMA model reads);
// that the real hardware behaves like it is what the GPU tier establishes.
// Design and roofline: see the header of match_tc.cu.

constexpr int TM = 128;            // query rows per tile (UMMA M)
constexpr int TN = 128;            // train rows per tile (UMMA N)
constexpr int QT = 4;              // query tiles per super-tile = TMEM accumulators (4 x 128 columns = 512)
constexpr int KB = PANO_DESC_STRIDE;  // 128 bytes of K per row = one 128B swizzle atom
constexpr int NSTAGE = 4;
constexpr int NCV = NSTAGE + 2;      // column-constant slots (see the header comment)
constexpr int K_STEPS = 3;         // 3 x 32 = 96 >= 75 descriptor bytes; bytes 96..127 of a row are zero padding
constexpr int A_BYTES = TM * KB;   // 16 KB
constexpr int B_BYTES = TN * KB;   // 16 KB
constexpr int EPI_WARPS = 4 * QT;  // group g = warps 4g .. 4g + 3 drains accumulator g
constexpr int TC_THREADS = 32 * (EPI_WARPS + 2);
constexpr uint32_t SPIN_LIMIT = 1u << 26;

// instruction descriptor, kind::i8: D = S32 (2 << 4), A = B = UINT8 (0), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24   (cute::UMMA::InstrDescriptor layout)
constexpr uint32_t IDESC = (2u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

struct Smem {
  // tiles first: 1024-byte aligned (SWIZZLE_128B requirement)
  uint8_t A[2][QT][A_BYTES];
  uint8_t B[NSTAGE][B_BYTES];
  int cvec[NCV][TN];               // per unit in flight: |t_j|^2 * 256 + (j mod 128), INT_MAX for padding
  unsigned long long b_full[NSTAGE], b_empty[NSTAGE];
  unsigned long long a_full[2], a_empty[2];
  unsigned long long t_full[QT];   // accumulator g complete (tcgen05.commit of its group's issuing lane)
  uint32_t tmem_base;
  int abort_flag;
};

#ifndef PANO_CUDA_EMU   // the PTX below has a host model of the same names in tests/hostsim/tcgen05_emu.hpp
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded polling wait (test_wait never blocks, so the spin bound is a real time bound);
// returns false if it gave up (or another role already aborted)
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, uint32_t parity, volatile int* abort_flag) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < SPIN_LIMIT; spin++) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return true;
    if ((spin & 1023u) == 1023u && *abort_flag) return false;
  }
  *abort_flag = 1;
  return false;
}

__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 2-D TMA tile load: box (128 bytes of K) x (rows) starting at row `row0`, completes on `bar`
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int row0,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(row0), "r"(smem_u32(bar))
      : "memory");
}

// 1-D bulk copy global -> shared (16-byte aligned, size a multiple of 16), completes on `bar`
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 [0,14), LBO >> 4 [16,30) (unused for swizzled K-major: 1), SBO >> 4 [32,46)
// = 1024 bytes between 8-row groups, version 1 at [46,48), layout type 2 (SWIZZLE_128B) at [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

#ifndef PANO_CUDA_EMU
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
#endif

// 32 columns of one accumulator row: key = cvec[j] - 512 * acc[j], four independent min chains
__device__ __forceinline__ void chunk_min(const int4* __restrict__ cv, const uint32_t (&acc)[32], int& km0, int& km1,
                                          int& km2, int& km3) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int4 c4 = cv[i];
    km0 = min(km0, (int)((uint32_t)c4.x - 512u * acc[4 * i]));
    km1 = min(km1, (int)((uint32_t)c4.y - 512u * acc[4 * i + 1]));
    km2 = min(km2, (int)((uint32_t)c4.z - 512u * acc[4 * i + 2]));
    km3 = min(km3, (int)((uint32_t)c4.w - 512u * acc[4 * i + 3]));
  }
}

// the same with the runner-up of each chain (top-2 epilogue of pano_match_knn): 1 IMAD + 3 VIMNMX per element
__device__ __forceinline__ void chunk_min2(const int4* __restrict__ cv, const uint32_t (&acc)[32], int (&km)[4], int (&ks)[4]) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int4 c4 = cv[i];
    top2_insert_i32(km[0], ks[0], (int)((uint32_t)c4.x - 512u * acc[4 * i]));
    top2_insert_i32(km[1], ks[1], (int)((uint32_t)c4.y - 512u * acc[4 * i + 1]));
    top2_insert_i32(km[2], ks[2], (int)((uint32_t)c4.z - 512u * acc[4 * i + 2]));
    top2_insert_i32(km[3], ks[3], (int)((uint32_t)c4.w - 512u * acc[4 * i + 3]));
  }
}

// the CTA's next run of units inside one super-row: units [unit, unit + n) of the flattened grid
struct Seg { int sr, t0, t1; };
__device__ __forceinline__ Seg next_seg(int unit, int hi, int n_ttiles) {
  Seg g;
  g.sr = unit / n_ttiles;
  g.t0 = unit - g.sr * n_ttiles;
  g.t1 = min(n_ttiles, g.t0 + (hi - unit));
  return g;
}

// TOP2 = false: the reference's matcher (nearest neighbour only; `best2` unused).  TOP2 = true: pano_match_knn's
// variant - the epilogue also keeps the runner-up of every query row and publishes both (knn_publish).  Pipeline,
// barriers and tensor-memory traffic are the same; only the per-element reduction and the final atomics differ.
template <bool TOP2>
__device__ __forceinline__ void match_tc_body(const CUtensorMap& tmap_q, const CUtensorMap& tmap_t,
                                              const uint32_t* __restrict__ qn, int nq, const int* __restrict__ tkey, int nt,
                                              int n_qtiles, int n_ttiles, int n_units, unsigned long long* __restrict__ best,
                                              unsigned long long* __restrict__ best2, int* __restrict__ err) {
#ifdef PANO_CUDA_EMU
  uint8_t* smem_raw = emu::dyn_smem();   // (CPU emulation tier: the launch's dynamic shared memory)
#else
  extern __shared__ __align__(1024) uint8_t smem_raw[];
#endif
  // keep the pointer in the shared address space (LDS/STS, not generic loads)
  Smem& S = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  volatile int* abort_flag = &S.abort_flag;

  if (threadIdx.x == 0) {
    // every group issues the MMAs of its own accumulator, so a B stage / an A super-row is released by QT arrivals
    for (int i = 0; i < NSTAGE; i++) { mbar_init(&S.b_full[i], 1); mbar_init(&S.b_empty[i], QT); }
    for (int i = 0; i < 2; i++) { mbar_init(&S.a_full[i], 1); mbar_init(&S.a_empty[i], QT); }
    for (int i = 0; i < QT; i++) mbar_init(&S.t_full[i], 1);
    S.abort_flag = 0;
#ifndef PANO_CUDA_EMU
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
  }
  if (warp == EPI_WARPS + 1) {  // TMEM: all 512 columns (four 128 x 128 s32 accumulators)
#ifdef PANO_CUDA_EMU
    tmem_alloc_emu(&S.tmem_base, 512u);
#else
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
#endif
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_base;

  // this CTA's run of the flattened (super-row, train tile) grid
  const int lo = (int)((long long)n_units * blockIdx.x / gridDim.x);
  const int hi = (int)((long long)n_units * (blockIdx.x + 1) / gridDim.x);

  if (warp < EPI_WARPS) {
    // ===== group g = warp >> 2 owns accumulator g (query tile 4 sr + g): its first lane issues the MMAs, all four
    // warps drain them.  (Round 2, first version: one MMA-issuing thread for all four accumulators; ncu showed the
    // epilogue warps polling its hand-over barriers for most of the kernel - 4200 clk per unit against 768 clk of
    // MMA - so the issue moved into the groups: no cross-role round trip is left on an accumulator's cycle.)
    const int g = warp >> 2;
    const bool leader = (threadIdx.x & 127) == 0;
    uint32_t unit_ctr = 0;   // units of this CTA so far (B stage / column-constant slot)
    uint32_t use_ctr = 0;    // units in which accumulator g was used (t_full's phase)
    uint32_t seg_ctr = 0;
    bool ok = true;
    for (int unit = lo; unit < hi && ok; seg_ctr++) {
      const Seg sg = next_seg(unit, hi, n_ttiles);
      const int n_seg = sg.t1 - sg.t0;
      unit += n_seg;
      const uint32_t ab = seg_ctr & 1u, aph = (seg_ctr >> 1) & 1u;
      const int qt = sg.sr * QT + g;
      if (qt >= n_qtiles) {
        // this super-row has fewer than QT query tiles: nothing to compute, but the producer counts QT releases
        // (paced by the "full" barriers: one release per phase, never two of this group inside one phase)
        if (leader) {
          ok = mbar_wait(&S.a_full[ab], aph, abort_flag);
          for (int i = 0; i < n_seg && ok; i++) {
            const uint32_t uc = unit_ctr + (uint32_t)i;
            ok = mbar_wait(&S.b_full[uc % NSTAGE], (uc / NSTAGE) & 1u, abort_flag);
            if (ok) mbar_arrive(&S.b_empty[uc % NSTAGE]);
          }
          if (ok) mbar_arrive(&S.a_empty[ab]);
        }
        unit_ctr += (uint32_t)n_seg;
        continue;
      }
      ok = mbar_wait(&S.a_full[ab], aph, abort_flag);
      if (!ok) break;
      const uint64_t adesc = make_desc(smem_u32(S.A[ab][g]));
      const int qrow = qt * TM + ((int)threadIdx.x & 127);
      const int myqn = qrow < nq ? (int)qn[qrow] : 0;
      int best_ssd = 0x7fffffff, best_j = -1;
      Top2 top = top2_empty();   // (TOP2 only)
      for (int tt = sg.t0; tt < sg.t1 && ok; tt++, unit_ctr++, use_ctr++) {
        const uint32_t s = unit_ctr % NSTAGE, sph = (unit_ctr / NSTAGE) & 1u;
        const uint32_t cs = unit_ctr % NCV;
        // every thread acquires the stage (descriptor tile + column constants written by the async proxy)
        ok = mbar_wait(&S.b_full[s], sph, abort_flag);
        if (!ok) break;
#ifdef PANO_TC_BARRIER_AFTER_STAGE_WAIT
        // (variant, off by default - DESIGN section 9: the group meets AFTER every thread has observed this unit's stage
        // instead of right after draining the previous one.  Same number of barriers; the leader's MMAs - whose commit
        // lets the producer refill the stage - then cannot be issued before all 128 threads have passed their wait, so no
        // thread can find the stage barrier two phases on.  The drain of the previous unit still precedes the barrier.)
#ifdef PANO_CUDA_EMU
        emu::named_barrier(1 + g, 128);
#else
        asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(128) : "memory");
#endif
#endif
        if (leader) {
          // accumulator g is free: the whole group passed the named barrier below after draining the previous unit
          tc_fence_after();
          const uint64_t bdesc = make_desc(smem_u32(S.B[s]));
          const uint32_t d_tmem = tmem_base + (uint32_t)(g * TN);
#pragma unroll
          for (int ks = 0; ks < K_STEPS; ks++)  // 32 bytes of K per MMA: descriptor start advances 32 B
            mma_i8(d_tmem, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), ks > 0 ? 1u : 0u);
          mma_commit(&S.t_full[g]);    // accumulator ready
          mma_commit(&S.b_empty[s]);   // this group's read of the stage is done once these MMAs have completed
          if (tt + 1 == sg.t1) mma_commit(&S.a_empty[ab]);
        }
        ok = mbar_wait(&S.t_full[g], use_ctr & 1u, abort_flag);
        if (!ok) break;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * TN);
        const int4* cv = reinterpret_cast<const int4*>(&S.cvec[cs][0]);
        int km0 = 0x7fffffff, km1 = 0x7fffffff, km2 = 0x7fffffff, km3 = 0x7fffffff;
        int km[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff}, ks[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
        // two register buffers: the tcgen05.ld of chunk c + 1 is in flight while chunk c is reduced
        uint32_t ra[32], rb[32];
        tmem_ld32(taddr, ra);
        tmem_ld_wait();
#pragma unroll
        for (int cb = 0; cb < TN / 32; cb += 2) {
          tmem_ld32(taddr + (cb + 1) * 32, rb);
          if constexpr (TOP2) chunk_min2(cv + cb * 8, ra, km, ks);
          else chunk_min(cv + cb * 8, ra, km0, km1, km2, km3);
          tmem_ld_wait();
          if (cb + 2 < TN / 32) tmem_ld32(taddr + (cb + 2) * 32, ra);
          if constexpr (TOP2) chunk_min2(cv + (cb + 1) * 8, rb, km, ks);
          else chunk_min(cv + (cb + 1) * 8, rb, km0, km1, km2, km3);
          tmem_ld_wait();
        }
        tc_fence_before();
        // the group's 128 threads have read the accumulator: its leader may overwrite it (named barrier 1 + g)
#ifndef PANO_TC_BARRIER_AFTER_STAGE_WAIT
#ifdef PANO_CUDA_EMU
        emu::named_barrier(1 + g, 128);
#else
        asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(128) : "memory");
#endif
#endif
        if constexpr (TOP2) {
          // the tile's two smallest keys out of the four chains' (valid keys of a tile are distinct: the column sits
          // in their low bits; 0x7fffffff = padding column / nothing), then into the row's running pair in
          // (SSD, train index) order - tiles come in train order, so earlier columns win ties as in the reference
          knn_fold_tile(top, km, ks, myqn, tt * TN);
        } else {
          const int kmin = min(min(km0, km1), min(km2, km3));
          const int v = kmin >> 8, jl = kmin & 255;
          const int ssd = v + myqn;
          if (kmin != 0x7fffffff && ssd < best_ssd) { best_ssd = ssd; best_j = tt * TN + jl; }
        }
      }
      if constexpr (TOP2) {
        if (ok && qrow < nq) knn_publish(best, best2, qrow, top);
      } else {
        if (ok && qrow < nq && best_j >= 0 && best_j < nt)
          atomicMin(&best[qrow], ((unsigned long long)(uint32_t)best_ssd << 32) | (uint32_t)best_j);
      }
    }
  } else if (warp == EPI_WARPS) {
    // ================= producer: TMA tile loads + bulk copies of the column constants =================
    uint32_t unit_ctr = 0, seg_ctr = 0;
    bool ok = true;
    for (int unit = lo; unit < hi && ok; seg_ctr++) {
      const Seg sg = next_seg(unit, hi, n_ttiles);
      unit += sg.t1 - sg.t0;
      const uint32_t ab = seg_ctr & 1u, aph = (seg_ctr >> 1) & 1u;
      ok = mbar_wait(&S.a_empty[ab], aph ^ 1u, abort_flag);
      if (!ok) break;
      const int nvalid = min(QT, n_qtiles - sg.sr * QT);
      if (lane == 0) {
        mbar_arrive_expect_tx(&S.a_full[ab], (uint32_t)nvalid * A_BYTES);
        for (int q = 0; q < nvalid; q++) tma_load_2d(S.A[ab][q], &tmap_q, 0, (sg.sr * QT + q) * TM, &S.a_full[ab]);
      }
      for (int tt = sg.t0; tt < sg.t1; tt++, unit_ctr++) {
        const uint32_t s = unit_ctr % NSTAGE, sph = (unit_ctr / NSTAGE) & 1u;
        const uint32_t cs = unit_ctr % NCV;
        ok = mbar_wait(&S.b_empty[s], sph ^ 1u, abort_flag);
        if (!ok) break;
        if (lane == 0) {   // descriptor tile (TMA) + the tile's 128 column constants (bulk copy), one barrier
          mbar_arrive_expect_tx(&S.b_full[s], B_BYTES + TN * (uint32_t)sizeof(int));
          tma_load_2d(S.B[s], &tmap_t, 0, tt * TN, &S.b_full[s]);
          bulk_load_1d(S.cvec[cs], tkey + (size_t)tt * TN, TN * (uint32_t)sizeof(int), &S.b_full[s]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && S.abort_flag) atomicOr(err, PANO_ERRW_TC_ABORT);
  if (warp == EPI_WARPS + 1) {
#ifndef PANO_CUDA_EMU
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
#endif
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
match_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_t,
                const uint32_t* __restrict__ qn, int nq, const int* __restrict__ tkey, int nt, int n_qtiles, int n_ttiles,
                int n_units, unsigned long long* __restrict__ best, int* __restrict__ err) {
  match_tc_body<false>(tmap_q, tmap_t, qn, nq, tkey, nt, n_qtiles, n_ttiles, n_units, best, nullptr, err);
}

// pano_match_knn: nearest neighbour and runner-up of every query row (keys (ssd << 32 | j) in best / best2)
__global__ void __launch_bounds__(TC_THREADS, 1)
match_tc_top2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_t,
                     const uint32_t* __restrict__ qn, int nq, const int* __restrict__ tkey, int nt, int n_qtiles,
                     int n_ttiles, int n_units, unsigned long long* __restrict__ best,
                     unsigned long long* __restrict__ best2, int* __restrict__ err) {
  match_tc_body<true>(tmap_q, tmap_t, qn, nq, tkey, nt, n_qtiles, n_ttiles, n_units, best, best2, err);
}

// column constants of the train side: |t_j|^2 * 256 + (j mod 256); INT_MAX for the padding columns
__global__ void tkey_kernel(const uint32_t* __restrict__ tn, int nt, int n_pad, int* __restrict__ tkey) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_pad) tkey[j] = j < nt ? (int)(tn[j] * 256u + (uint32_t)(j & (TN - 1))) : 0x7fffffff;
}

