// ransac_emu.cpp — runs the device code of the RANSAC stage (csrc/ransac_kernels.cuh: mt19937 stream, the shuffle
// replay in its chunked and resident forms, point builder, warp-per-hypothesis DLT, scoring, selection, inlier mask)
// on the CPU emulation of the CUDA execution model (cuda_emu.hpp), for the no-GPU test tier.
//
// TEST INFRASTRUCTURE ONLY.  emu_ransac mirrors ransac.cu's host flow (mt_ensure, ransac_device: plan from
// replay_plan.hpp, buffers, the per-chunk kernel sequence, solve) with the same launch arithmetic; the kernels are the
// product's source compiled unchanged by g++ (-ffp-contract=off).
#include "cuda_emu.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <ctime>
#include <memory>

#include "../../include/pano_b200.h"
#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/pano_core.cuh"
#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/replay_plan.hpp"

namespace pano {
constexpr int PANO_ERRW_BAD_INDEX = 4;    // as in common.cuh
inline void pdl_wait() {}                 // programmatic dependent launch: kernels run one after another here
inline void pdl_trigger() {}
namespace {
#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/ransac_kernels.cuh"
}  // namespace
}  // namespace pano

using namespace pano;

namespace {
template <typename T>
struct Aligned {
  T* p = nullptr;
  size_t n = 0;
  explicit Aligned(size_t n_, int fill = 0) : n(n_) {
    const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) / 256 * 256 + 256;
    p = static_cast<T*>(aligned_alloc(256, bytes));
    memset(p, fill, bytes);
  }
  ~Aligned() { free(p); }
  Aligned(const Aligned&) = delete;
};
const char* g_error = nullptr;
void run(dim3 grid, dim3 block, const std::function<void()>& body, size_t dyn = 0, int order = emu::SHUFFLED) {
  static const bool timing = getenv("REMU_TIMING") != nullptr;
  const clock_t t0 = clock();
  const char* e = emu::launch(grid, block, body, order, dyn);
  if (timing)
    fprintf(stderr, "[remu] grid %u x %u, block %u: %.2f s\n", grid.x, grid.y, block.x * block.y, (double)(clock() - t0) / CLOCKS_PER_SEC);
  if (e) g_error = e;
}
}  // namespace

extern "C" {

const char* remu_last_error() { return g_error ? g_error : ""; }

// std::mt19937(seed) outputs through mt_generate_kernel (mt_ensure's flow: whole generations of 624 words)
void remu_mt19937(uint32_t seed, int n, uint32_t* out) {
  g_error = nullptr;
  const int gens = (n + MT_N - 1) / MT_N + 1;
  Aligned<uint32_t> state(MT_N), x((size_t)gens * MT_N);
  // two launches, to exercise the "continue from the saved state" path as well
  const int g1 = gens / 2;
  if (g1 > 0) run(dim3(1), dim3(256), [&] { mt_generate_kernel(state.p, 1, seed, x.p, g1); });
  run(dim3(1), dim3(256), [&] { mt_generate_kernel(state.p, g1 > 0 ? 0 : 1, seed, x.p + (size_t)g1 * MT_N, gens - g1); });
  memcpy(out, x.p, sizeof(uint32_t) * (size_t)n);
}

// pano_ransac on the emulation.  replay_mode: 0 chunked, 1 resident.  target_cand: candidate walks per chunk (small
// values = many chunks); z_sigma: window half-width (small values force window misses).  window_scale as in
// ransac_retry.  Returns the status of ransac_device (PANO_OK, PANO_ERR_*, or -replay_status for a missed window);
// -100 on an emulation error.  samples_out[iters * 4], counts_out[iters], mask_out[m] may be null.
int remu_ransac(const int32_t* kp1, int n1, const int32_t* kp2, int n2, const pano_dmatch* matches, int m, int iters, double thr,
                uint32_t seed, int replay_mode, double target_cand, double z_sigma, int window_scale, double* H_out,
                int* best_count, int* best_iter, int32_t* samples_out, int32_t* counts_out, uint8_t* mask_out, int* errw_out,
                int* n_chunks_out) {
  g_error = nullptr;
  if (m < 4 || iters <= 0) return PANO_ERR_TOO_FEW_MATCHES;
  const uint32_t n = (uint32_t)m;
  const bool pairs = shuffle_uses_pairs(n);
  const uint32_t steps = shuffle_steps(n);
  const ReplayPlan plan = plan_replay(n, iters, window_scale, target_cand, z_sigma);
  const ResidentPlan rplan = plan_resident(plan, n, window_scale, z_sigma + 0.3);
  const bool resident = replay_mode == 1 && rplan.ok;
  const std::vector<WinEntry>& win = plan.win;
  const int G = plan.G;
  const uint32_t n_cand = plan.n_cand, max_w = plan.max_w;
  const int n_chunks = (iters + G - 1) / G;
  if (n_chunks_out) *n_chunks_out = resident ? 0 : n_chunks;

  // ---- mt_ensure ----------------------------------------------------------------------------------------------
  const uint64_t need = plan.stream_need, guard = (uint64_t)steps + 4096;
  const uint64_t gens_total = (need + MT_N - 1) / MT_N + 1, mt_len = gens_total * MT_N;
  Aligned<uint32_t> state(MT_N), X((size_t)(mt_len + guard));
  run(dim3(1), dim3(256), [&] { mt_generate_kernel(state.p, 1, seed, X.p, (int)gens_total); });
  memset(X.p + mt_len, 0xff, sizeof(uint32_t) * guard);

  // ---- buffers (ransac_device) ----------------------------------------------------------------------------------
  const int nseg = (int)((steps + PANO_SEG_STEPS - 1) / PANO_SEG_STEPS);
  Aligned<RT> rt_dev(plan.rt.size());
  memcpy(rt_dev.p, plan.rt.data(), sizeof(RT) * plan.rt.size());
  Aligned<uint8_t> planbuf(std::max(sizeof(WinEntry) * win.size(),
                                    sizeof(ResBlock) * (size_t)rplan.nkb + sizeof(uint32_t) * ((size_t)rplan.nseg + 2)));
  Aligned<uint32_t> cand_off((size_t)n_cand * (size_t)std::max(nseg, 1));
  Aligned<unsigned long long> seg_tab_buf((size_t)n_chunks * G * nseg + 1);
  Aligned<ReplayCtl> ctlbuf(1);
  const uint32_t nkb = (steps + 31u) / 32u;
  const int n_dblocks = (int)plan.diag_block_iter.size();
  Aligned<uint8_t> bitsbuf(sizeof(uint32_t) * (size_t)plan.n_diag * nkb + sizeof(int) * (size_t)n_dblocks + 256);
  Aligned<int4> samples(std::max((size_t)n_chunks * G * (nseg + 1), (size_t)iters));
  Aligned<float4> pts((size_t)m);
  Aligned<double> Hs((size_t)9 * iters);
  Aligned<int> valid((size_t)iters), counts((size_t)iters), errw(1);
  Aligned<SelectOut> result(1);
  Aligned<uint8_t> mask((size_t)m);
  Aligned<int32_t> k1((size_t)2 * std::max(n1, 1)), k2((size_t)2 * std::max(n2, 1));
  Aligned<pano_dmatch> md((size_t)m);
  memcpy(k1.p, kp1, sizeof(int32_t) * 2 * (size_t)n1);
  memcpy(k2.p, kp2, sizeof(int32_t) * 2 * (size_t)n2);
  memcpy(md.p, matches, sizeof(pano_dmatch) * (size_t)m);
  if (!resident) memcpy(planbuf.p, win.data(), sizeof(WinEntry) * win.size());
  uint32_t* bits = reinterpret_cast<uint32_t*>(bitsbuf.p);
  int* dbi = reinterpret_cast<int*>(bits + (size_t)plan.n_diag * nkb);
  if (!resident) memcpy(dbi, plan.diag_block_iter.data(), sizeof(int) * (size_t)n_dblocks);
  ReplayCtl* ctl = ctlbuf.p;
  uint32_t* cand_end = cand_off.p;
  uint32_t* seg_off = cand_end + n_cand;
  unsigned long long* seg_tab = seg_tab_buf.p;
  int4* samples_dev = samples.p;
  int4* seg_w = samples_dev + (size_t)n_chunks * G;

  run(dim3((m + 255) / 256), dim3(256), [&] { build_points_kernel(k1.p, n1, k2.p, n2, md.p, m, pts.p, errw.p); });

  if (resident) {
    ResBlock* blk_dev = reinterpret_cast<ResBlock*>(planbuf.p);
    uint32_t* eoff_dev = reinterpret_cast<uint32_t*>(blk_dev + rplan.nkb);
    memcpy(blk_dev, rplan.blk.data(), sizeof(ResBlock) * rplan.nkb);
    memcpy(eoff_dev, rplan.seg_eoff.data(), sizeof(uint32_t) * (rplan.nseg + 1));
    ResParams rp;
    rp.X = X.p;
    rp.x_limit = mt_len + guard;
    rp.rt = rt_dev.p;
    rp.blk = blk_dev;
    rp.seg_eoff = eoff_dev;
    rp.n = n; rp.steps = steps; rp.nkb = rplan.nkb; rp.nwords = rplan.nwords; rp.dmax = rplan.dmax;
    rp.segb = rplan.segb; rp.nseg = rplan.nseg; rp.n_entries = rplan.n_entries; rp.xcap = rplan.xcap;
    rp.iters = iters;
    run(dim3(1), dim3(RES_THREADS), [&] { replay_resident_kernel(rp, ctl, samples_dev); }, rplan.smem_bytes);
  } else {
    const size_t chain_smem = (size_t)n_cand * sizeof(uint32_t) <= CHAIN_SMEM_MAX ? (size_t)n_cand * sizeof(uint32_t) : 0;
    for (int c = 0; c < n_chunks; c++) {
      const int Gc = std::min(G, iters - c * G);
      const long long warps = (long long)Gc * nkb;
      run(dim3((unsigned)((warps + 7) / 8)), dim3(256), [&] {
        replay_cells_kernel(X.p, steps, rt_dev.p, reinterpret_cast<WinEntry*>(planbuf.p), Gc, nkb, plan.dextra, ctl, bits,
                            mt_len + guard - 64);
      });
      run(dim3((max_w + RW_THREADS - 1) / RW_THREADS, Gc), dim3(RW_THREADS), [&] {
        replay_walk_bits_kernel(steps, reinterpret_cast<WinEntry*>(planbuf.p), nkb, plan.dextra, ctl, bits, cand_end, seg_off,
                                (int)n_cand, mt_len);
      });
      run(dim3(1), dim3(1024), [&] {
        replay_chain_kernel(reinterpret_cast<WinEntry*>(planbuf.p), Gc, steps, cand_end, seg_off, (int)n_cand, nseg, ctl,
                            seg_tab + (size_t)c * G * nseg);
      }, chain_smem, emu::FORWARD);
    }
    const int nthr = iters * nseg;
    if (pairs)
      run(dim3((nthr + 63) / 64), dim3(64), [&] { replay_segments_kernel<true>(X.p, n, steps, rt_dev.p, seg_tab, iters, nseg, ctl, seg_w); });
    else
      run(dim3((nthr + 63) / 64), dim3(64), [&] { replay_segments_kernel<false>(X.p, n, steps, rt_dev.p, seg_tab, iters, nseg, ctl, seg_w); });
    run(dim3((iters + 127) / 128), dim3(128), [&] { combine_samples_kernel(seg_w, iters, nseg, samples_dev); });
  }

  run(dim3((iters + DLT_WARPS - 1) / DLT_WARPS), dim3(DLT_WARPS * 32), [&] { dlt_kernel(pts.p, samples.p, iters, Hs.p, valid.p); });
  run(dim3(iters), dim3(256), [&] { score_kernel(pts.p, m, Hs.p, valid.p, inlier_d2_limit(thr), counts.p); });
  run(dim3(1), dim3(1024), [&] { select_kernel(counts.p, iters, Hs.p, result.p); }, 0, emu::FORWARD);
  if (mask_out) run(dim3((m + 255) / 256), dim3(256), [&] { inlier_mask_kernel(pts.p, m, result.p, thr, mask.p); });
  if (g_error) return -100;
  if (errw_out) *errw_out = errw.p[0];
  if (errw.p[0] != 0) return PANO_ERR_CUDA;
  if (ctl->status != 0) return -ctl->status;
  const SelectOut& so = *result.p;
  memcpy(H_out, so.H, sizeof so.H);
  *best_count = so.best_count;
  *best_iter = so.best_iter;
  if (samples_out) memcpy(samples_out, samples.p, sizeof(int4) * (size_t)iters);
  if (counts_out) memcpy(counts_out, counts.p, sizeof(int) * (size_t)iters);
  if (mask_out) memcpy(mask_out, mask.p, (size_t)m);
  return so.status;
}

}  // extern "C"
