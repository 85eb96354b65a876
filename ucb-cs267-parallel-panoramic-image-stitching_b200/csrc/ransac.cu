// ransac.cu — K5 sample replay, K6 batched 4-point DLT, K7 inlier scoring + ordered argmax.
//
// Semantics: ref src/serial/main.cpp:247-307 (SeqRansacHomographyCalculator), seeded.
//
// K5 — replaying `std::shuffle(all M matches)` x num_iterations with ONE continuing
// std::mt19937 (ref :264-271).  Only the first four elements of each shuffled copy are used
// (:274-277), and the copy starts from the identity each iteration (:270), so an iteration's
// sample depends only on where in the engine's output stream it starts.  Lemire's rejection
// sampling makes every step consume a data-dependent number of outputs, so the start offset
// of iteration t depends on every earlier rejection.  The replay is made parallel by
// windowed speculation:
//   * the mt19937(seed) output stream is generated once per context and kept in HBM;
//   * iterations are processed in chunks of G; inside a chunk, iteration g is walked from
//     EVERY start offset in a window around its expected offset (the window grows like
//     sqrt(g) * sigma, sigma = std-dev of the rejections of one shuffle, from the exact
//     per-step rejection probabilities), one thread per (iteration, candidate offset);
//   * a single block then chains the per-candidate end offsets from the chunk's exact start
//     and picks each iteration's true candidate (exact: a miss is detected, never guessed,
//     and the caller re-runs with wider windows).
// The walk itself tracks only what lands in positions 0..3 (see pano_core.cuh).
//
// K6 — one thread per hypothesis runs OpenCV's 4-point findHomography path (normalised DLT,
// Jacobi eigen of the 9x9 LtL) in FP64 in OpenCV's exact operation order.
// K7 — one block per hypothesis counts inliers with the reference's mixed f64/f32 predicate;
// the first iteration with the strictly largest count wins (:295-298).
#include "common.cuh"
#include "replay_plan.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace pano {

namespace {

#include "ransac_kernels.cuh"

}  // namespace

void mt_ensure(cudaStream_t st, MtStream& mt, uint32_t seed, uint64_t need, uint64_t guard) {
  const bool fresh = !mt.valid || mt.seed != seed;
  if (!fresh && mt.len >= need && mt.guard >= guard) return;
  if (!fresh && mt.len >= need) need = mt.len;  // only the guard has to grow
  uint64_t gens_total = (need + MT_N - 1) / MT_N + 1;
  uint64_t new_len = gens_total * MT_N;
  if (fresh) {
    mt.state.reserve(sizeof(uint32_t) * MT_N);
    mt.x.reserve(sizeof(uint32_t) * (new_len + guard));
    mt.len = 0;
    mt.seed = seed;
    mt.valid = true;
  } else if (sizeof(uint32_t) * (new_len + guard) > mt.x.cap) {
    // grow geometrically, keeping what was generated
    DevBuf bigger;
    bigger.reserve(sizeof(uint32_t) * (std::max<uint64_t>(new_len, 2 * mt.len) + guard));
    PANO_CUDA(cudaMemcpyAsync(bigger.p, mt.x.p, sizeof(uint32_t) * mt.len, cudaMemcpyDeviceToDevice, st));
    PANO_CUDA(stream_wait(st));
    mt.x.release();
    mt.x = bigger;
  }
  uint64_t gens = (new_len - mt.len) / MT_N;
  uint64_t done = 0;
  while (done < gens) {
    int g = (int)std::min<uint64_t>(gens - done, 1u << 20);
    mt_generate_kernel<<<1, 256, 0, st>>>(mt.state.as<uint32_t>(), (fresh && done == 0 && mt.len == 0) ? 1 : 0, seed,
                                          mt.x.as<uint32_t>() + mt.len, g);
    PANO_LAUNCH_CHECK();
    mt.len += (uint64_t)g * MT_N;
    done += g;
  }
  // guard region: 0xFFFFFFFF is never rejected by Lemire's test (its low word is 2^32 - r,
  // which is >= 2^32 mod r), so a speculative walk that runs off the generated stream
  // terminates after at most `steps` more reads instead of spinning on stale memory
  PANO_CUDA(cudaMemsetAsync(mt.x.as<uint32_t>() + mt.len, 0xff, sizeof(uint32_t) * guard, st));
  mt.guard = guard;
}

RansacResult ransac_device(cudaStream_t st, const int32_t* kp1_dev, const int32_t* kp2_dev,
                           const pano_dmatch* matches_dev, int m, const pano_ransac_opts& o, uint32_t seed,
                           MtStream& mt, RansacScratch& s, PinnedBuf& pin, int32_t* samples_out_host,
                           int32_t* counts_out_host, uint8_t* mask_out_host, int window_scale, double replay_target,
                           int replay_mode, int phase) {
  // phase 0: everything; 1: only the shuffle replay (it depends on nothing but the match COUNT m, so a caller can
  // enqueue it on another stream while the matches are still being computed); 2: only the solve (points, DLT,
  // scoring, selection) on top of a replay enqueued by phase 1 with the same m / options / scale
  const bool do_replay = phase != 2, do_solve = phase != 1;
  RansacResult res;
  memset(&res, 0, sizeof res);
  res.best_iter = -1;
  const int iters = o.num_iterations;
  if (m < o.num_samples || iters <= 0) {  // ref :268-269: the loop breaks at once -> empty H
    res.status = PANO_ERR_TOO_FEW_MATCHES;
    return res;
  }
  const uint32_t n = (uint32_t)m;
  const bool pairs = shuffle_uses_pairs(n);
  const uint32_t steps = shuffle_steps(n);

  // ---- per-step Lemire thresholds, rejection statistics, chunk/window plan (host, O(steps)
  //      integer work: launch-parameter planning, like the Gaussian taps) ------------------
  // candidate walks per chunk: larger chunks = fewer sequential phases (better single-pair
  // latency), smaller chunks = less speculative work (better throughput when pairs overlap)
  static const double env_target = [] {
    const char* e = getenv("PANO_REPLAY_TARGET");
    double v = e ? atof(e) : 0.0;
    return (v >= 64.0 && v <= 51000.0) ? v : 0.0;
  }();
  const double target_cand = env_target > 0 ? env_target : (replay_target > 0 ? replay_target : 50000.0);
  static const double z_sigma = [] {   // window half-width in sigmas (test hook: small values force re-runs)
    const char* e = getenv("PANO_REPLAY_Z");
    double v = e ? atof(e) : 0.0;
    return (v > 0.0 && v < 20.0) ? v : 4.2;
  }();
  // plans depend only on (M, iterations, window scale, chunk target): keep the most recent ones
  struct PlanKey { uint32_t n; int iters, scale; double target; };
  struct Plans { ReplayPlan chunked; ResidentPlan resident; };
  typedef std::vector<std::pair<PlanKey, Plans>> PlanCache;   // lives in the context: lanes keep theirs across batch calls
  if (!s.plan_cache) s.plan_cache = std::make_shared<PlanCache>();
  PlanCache& plan_cache = *static_cast<PlanCache*>(s.plan_cache.get());
  const Plans* plan_ptr = nullptr;
  for (auto& e : plan_cache)
    if (e.first.n == n && e.first.iters == iters && e.first.scale == window_scale && e.first.target == target_cand)  // (z_sigma is process-constant)
      plan_ptr = &e.second;
  if (!plan_ptr) {
    if (plan_cache.size() >= 8) plan_cache.erase(plan_cache.begin());
    Plans np;
    np.chunked = plan_replay(n, iters, window_scale, target_cand, z_sigma);
    np.resident = plan_resident(np.chunked, n, window_scale, z_sigma + 0.3);
    plan_cache.emplace_back(PlanKey{n, iters, window_scale, target_cand}, std::move(np));
    plan_ptr = &plan_cache.back().second;
  }
  const ReplayPlan& plan = plan_ptr->chunked;
  const ResidentPlan& rplan = plan_ptr->resident;
  static const int env_mode = [] {   // PANO_REPLAY_MODE: 0 chunked, 1 resident (overrides the context's choice)
    const char* e = getenv("PANO_REPLAY_MODE");
    return e ? atoi(e) : -1;
  }();
  const bool resident = ((env_mode >= 0 ? env_mode : replay_mode) == 1) && rplan.ok;
  const std::vector<WinEntry>& win = plan.win;
  const int G = plan.G;
  const uint32_t n_cand = plan.n_cand, max_w = plan.max_w;
  const int n_chunks = (iters + G - 1) / G;

  // ---- stream: everything the walks can touch, with margin ---------------------------------
  uint64_t need = plan.stream_need;
  MtStream* mtp = &mt;
  if (s.shared_mt && s.shared_mt->valid && s.shared_mt->seed == seed && s.shared_mt->len >= need &&
      s.shared_mt->guard >= (uint64_t)steps + 4096)
    mtp = const_cast<MtStream*>(s.shared_mt);   // (never modified through this pointer)
  else if (do_replay)
    mt_ensure(st, mt, seed, need, (uint64_t)steps + 4096);

  // ---- buffers ---------------------------------------------------------------------------
  const int nseg = (int)((steps + PANO_SEG_STEPS - 1) / PANO_SEG_STEPS);
  s.thr.reserve(sizeof(RT) * plan.rt.size());
  s.plan.reserve(std::max(sizeof(WinEntry) * win.size(),
                          sizeof(ResBlock) * (size_t)rplan.nkb + sizeof(uint32_t) * ((size_t)rplan.nseg + 2)));
  s.cand_off.reserve(sizeof(uint32_t) * (size_t)n_cand * (size_t)std::max(nseg, 1));  // end + (nseg-1) boundaries
  s.cand_samp.reserve(sizeof(unsigned long long) * ((size_t)n_chunks * G * nseg + 1));  // segment start table
  s.base.reserve(sizeof(ReplayCtl));
  const uint32_t nkb = (steps + 31u) / 32u;
  const int n_dblocks = (int)plan.diag_block_iter.size();
  s.pts_bits.reserve(sizeof(uint32_t) * (size_t)plan.n_diag * nkb + sizeof(int) * (size_t)n_dblocks + 256);
  s.samples.reserve(sizeof(int4) * std::max((size_t)n_chunks * G * (nseg + 1), (size_t)iters));
  s.pts.reserve(sizeof(float4) * (size_t)m);
  s.Hs.reserve(sizeof(double) * 9 * (size_t)iters);
  s.valid.reserve(sizeof(int) * (size_t)iters);
  s.counts.reserve(sizeof(int) * (size_t)iters);
  s.result.reserve(sizeof(SelectOut));
  s.mask.reserve((size_t)m);
  pin.reserve(sizeof(SelectOut) + 64);

  if (do_replay) {
    PANO_CUDA(cudaMemcpyAsync(s.thr.p, plan.rt.data(), sizeof(RT) * plan.rt.size(), cudaMemcpyHostToDevice, st));
    if (!resident)
      PANO_CUDA(cudaMemcpyAsync(s.plan.p, win.data(), sizeof(WinEntry) * win.size(), cudaMemcpyHostToDevice, st));
    PANO_CUDA(cudaMemsetAsync(s.base.p, 0, sizeof(ReplayCtl), st));
  }
  uint32_t* bits = s.pts_bits.as<uint32_t>();
  int* dbi = reinterpret_cast<int*>(bits + (size_t)plan.n_diag * nkb);
  if (!resident && do_replay)
    PANO_CUDA(cudaMemcpyAsync(dbi, plan.diag_block_iter.data(), sizeof(int) * (size_t)n_dblocks, cudaMemcpyHostToDevice, st));
  ReplayCtl* ctl = s.base.as<ReplayCtl>();
  int* status_ptr = &ctl->status;
  uint32_t* cand_end = s.cand_off.as<uint32_t>();
  uint32_t* seg_off = cand_end + n_cand;
  unsigned long long* seg_tab = s.cand_samp.as<unsigned long long>();
  int4* samples_dev = s.samples.as<int4>();
  int4* seg_w = samples_dev + (size_t)n_chunks * G;

  if (do_solve) {
    build_points_kernel<<<(m + 255) / 256, 256, 0, st>>>(kp1_dev, s.n1, kp2_dev, s.n2, matches_dev, m, s.pts.as<float4>(),
                                                         s.errw);
    PANO_LAUNCH_CHECK();
  }

  ProfScope* prep = do_replay ? new ProfScope(PROF_REPLAY, st) : nullptr;
  if (!do_replay) {
    // (samples were produced by an earlier phase-1 call)
  } else if (resident) {
    // one CTA, iterations in order from exact offsets (throughput mode: leaves the other SMs to other pairs)
    ResBlock* blk_dev = s.plan.as<ResBlock>();
    uint32_t* eoff_dev = reinterpret_cast<uint32_t*>(blk_dev + rplan.nkb);
    PANO_CUDA(cudaMemcpyAsync(blk_dev, rplan.blk.data(), sizeof(ResBlock) * rplan.nkb, cudaMemcpyHostToDevice, st));
    PANO_CUDA(cudaMemcpyAsync(eoff_dev, rplan.seg_eoff.data(), sizeof(uint32_t) * (rplan.nseg + 1), cudaMemcpyHostToDevice, st));
    ResParams rp;
    rp.X = mtp->x.as<uint32_t>();
    rp.x_limit = mtp->len + mtp->guard;
    rp.rt = s.thr.as<RT>();
    rp.blk = blk_dev;
    rp.seg_eoff = eoff_dev;
    rp.n = n; rp.steps = steps; rp.nkb = rplan.nkb; rp.nwords = rplan.nwords; rp.dmax = rplan.dmax;
    rp.segb = rplan.segb; rp.nseg = rplan.nseg; rp.n_entries = rplan.n_entries; rp.xcap = rplan.xcap;
    rp.iters = iters;
    PANO_CUDA(cudaFuncSetAttribute(replay_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)RES_SMEM_BUDGET));
    replay_resident_kernel<<<1, RES_THREADS, rplan.smem_bytes, st>>>(rp, ctl, samples_dev);
    PANO_LAUNCH_CHECK();
  } else {
    size_t chain_smem = (size_t)n_cand * sizeof(uint32_t) <= CHAIN_SMEM_MAX ? (size_t)n_cand * sizeof(uint32_t) : 0;
    PANO_CUDA(cudaFuncSetAttribute(replay_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHAIN_SMEM_MAX));
    for (int c = 0; c < n_chunks; c++) {
      int Gc = std::min(G, iters - c * G);
      dim3 grid((max_w + RW_THREADS - 1) / RW_THREADS, Gc);
      {
        long long warps = (long long)Gc * nkb;
        launch_pdl(replay_cells_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, st, mtp->x.as<uint32_t>(), steps,
                   s.thr.as<RT>(), s.plan.as<WinEntry>(), Gc, nkb, plan.dextra, ctl, bits, mtp->len + mtp->guard - 64);
        PANO_LAUNCH_CHECK();
      }
      launch_pdl(replay_walk_bits_kernel, grid, dim3(RW_THREADS), 0, st, steps, s.plan.as<WinEntry>(), nkb, plan.dextra,
                 ctl, bits, cand_end, seg_off, (int)n_cand, mtp->len);
      PANO_LAUNCH_CHECK();
      launch_pdl(replay_chain_kernel, dim3(1), dim3(1024), chain_smem, st, s.plan.as<WinEntry>(), Gc, steps, cand_end,
                 seg_off, (int)n_cand, nseg, ctl, seg_tab + (size_t)c * G * nseg);
      PANO_LAUNCH_CHECK();
    }
    {
      int nthr = iters * nseg;
      if (pairs)
        launch_pdl(replay_segments_kernel<true>, dim3((nthr + 63) / 64), dim3(64), 0, st, mtp->x.as<uint32_t>(), n, steps,
                   s.thr.as<RT>(), seg_tab, iters, nseg, ctl, seg_w);
      else
        launch_pdl(replay_segments_kernel<false>, dim3((nthr + 63) / 64), dim3(64), 0, st, mtp->x.as<uint32_t>(), n, steps,
                   s.thr.as<RT>(), seg_tab, iters, nseg, ctl, seg_w);
      PANO_LAUNCH_CHECK();
      launch_pdl(combine_samples_kernel, dim3((iters + 127) / 128), dim3(128), 0, st, seg_w, iters, nseg, samples_dev);
      PANO_LAUNCH_CHECK();
    }
  }

  delete prep;
  if (!do_solve) {
    res.status = PANO_OK;
    return res;
  }
  {
    ProfScope ps(PROF_DLT, st);
  launch_pdl(dlt_kernel, dim3((iters + DLT_WARPS - 1) / DLT_WARPS), dim3(DLT_WARPS * 32), 0, st, s.pts.as<float4>(),
             s.samples.as<int4>(), iters, s.Hs.as<double>(), s.valid.as<int>());
  }
  PANO_LAUNCH_CHECK();
  {
    ProfScope ps(PROF_SCORE, st);
  launch_pdl(score_kernel, dim3(iters), dim3(256), 0, st, s.pts.as<float4>(), m, s.Hs.as<double>(), s.valid.as<int>(),
             inlier_d2_limit(o.distance_threshold), s.counts.as<int>());
  }
  PANO_LAUNCH_CHECK();
  select_kernel<<<1, 1024, 0, st>>>(s.counts.as<int>(), iters, s.Hs.as<double>(), s.result.as<SelectOut>());
  PANO_LAUNCH_CHECK();
  if (mask_out_host) {
    inlier_mask_kernel<<<(m + 255) / 256, 256, 0, st>>>(s.pts.as<float4>(), m, s.result.as<SelectOut>(),
                                                       o.distance_threshold, s.mask.as<uint8_t>());
    PANO_LAUNCH_CHECK();
  }
  char* pp = pin.as<char>();
  PANO_CUDA(cudaMemcpyAsync(pp, s.result.p, sizeof(SelectOut), cudaMemcpyDeviceToHost, st));
  PANO_CUDA(cudaMemcpyAsync(pp + sizeof(SelectOut), status_ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
  PANO_CUDA(cudaMemcpyAsync(pp + sizeof(SelectOut) + sizeof(int), s.errw, sizeof(int), cudaMemcpyDeviceToHost, st));
  PANO_CUDA(stream_wait(st));
  SelectOut so;
  memcpy(&so, pp, sizeof so);
  int replay_status;
  memcpy(&replay_status, pp + sizeof(SelectOut), sizeof(int));
  memcpy(&res.errw, pp + sizeof(SelectOut) + sizeof(int), sizeof(int));
  if (res.errw != 0) {   // matcher abort / missing minimum / bad match index: the caller fails the call
    res.status = PANO_ERR_CUDA;
    return res;
  }
  if (replay_status != 0) {
    // a speculation window was missed (bit 0) or the stream margin was too small (bit 1):
    // nothing was guessed; tell the caller to re-run with wider windows
    res.status = -replay_status;
    return res;
  }
  res.status = so.status;
  memcpy(res.H, so.H, sizeof so.H);
  res.best_count = so.best_count;
  res.best_iter = so.best_iter;
  if (samples_out_host)
    PANO_CUDA(cudaMemcpyAsync(samples_out_host, s.samples.p, sizeof(int4) * (size_t)iters, cudaMemcpyDeviceToHost, st));
  if (counts_out_host)
    PANO_CUDA(cudaMemcpyAsync(counts_out_host, s.counts.p, sizeof(int) * (size_t)iters, cudaMemcpyDeviceToHost, st));
  if (mask_out_host)
    PANO_CUDA(cudaMemcpyAsync(mask_out_host, s.mask.p, (size_t)m, cudaMemcpyDeviceToHost, st));
  if (samples_out_host || counts_out_host || mask_out_host) PANO_CUDA(stream_wait(st));
  return res;
}

}  // namespace pano
