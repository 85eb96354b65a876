// pano_api.cu — the C ABI (include/pano_b200.h): context, stage entry points, and the fused
// device-resident pair / fold pipelines.  Host-side orchestration mirrors
// ref src/serial/main.cpp:311-414 (stitchTwoImages / stitchAllImages) and the stage
// interfaces of ref src/gpu/*.cuh.  No CPU fallback anywhere: without an sm_100 device
// pano_create fails and nothing else can be called.
#include "common.cuh"
#include "replay_plan.hpp"
#include <array>

#include <algorithm>
#include <exception>
#include <new>
#include <future>
#include <mutex>
#include <thread>

namespace pano {
thread_local int t_yield_wait = 0;
thread_local Prof* t_prof = nullptr;
// (The library does not touch the process environment.  Batch lanes are independent streams; with the default of 8
// hardware work queues they alias, so callers that batch should export CUDA_DEVICE_MAX_CONNECTIONS=32 before the
// CUDA context exists - bench.py and the executables do; see INTEGRATION.md.)
std::atomic<uint64_t> g_kernel_launches{0};
}

using namespace pano;

struct pano_ctx {
  int device = 0;
  uint32_t seed = 0;
  int matcher = 0;  // 0 tensor-core, 1 SIMT
  int match_mode = 0;        // fused calls: 0 the reference's matcher, 1 pano_match_knn (opt-in; pano_set_match_mode)
  pano_knn_opts knn_mode = {5, PANO_KNN_PATCH_SSD, 0.75};
  double replay_target = 0;  // candidate walks per replay chunk (0 = default)
  bool overlap_replay = true;  // stitch: start the shuffle replay as soon as the match count is known (own stream)
  int replay_mode = 0;       // 0: chunked speculative replay (lowest latency), 1: resident one-CTA replay (least work)
  cudaStream_t st = nullptr;
  bool owns_stream = true;
  std::string err;
  PinnedBuf pin;
  Prof prof;     // per-kernel event timing (pano_set_profile)
  DevBuf errw;   // device error word (PANO_ERRW_* bits), read back with every result the host waits for
  DevBuf up[2];  // staging for host images
  // batch lanes with host buffers: the next pair's images are uploaded, and the previous canvas downloaded, on
  // their own streams while this pair is being stitched
  DevBuf upq[2][2];
  cudaStream_t st_up = nullptr, st_down = nullptr;
  cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_down[2] = {nullptr, nullptr};
  DevBuf kpup[2], mup;
  HarrisScratch hs;
  DevKeypoints kpL, kpR;
  MatchScratch ms;
  KnnScratch ks;     // pano_match_knn only
  DevDescriptors dQ, dT;
  DevBuf best, matches;
  RansacScratch rs;
  MtStream mt;
  DevBuf tmp[3];
  DevBuf canvas[2];
  DevBuf tight;          // the current canvas tightly packed (3 * w bytes per row) for flat device-to-host copies
  bool pack_tight = false;   // stage B also produces `tight` (batch slots with host canvases)
  int cur = 0;
  int cw = 0, ch = 0;
  size_t cstride = 0;
  bool has_canvas = false;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  struct Job {             // the pair between its stage A and its stage B (see pair_stage_a / pair_stage_b)
    DevImage L, R;
    int m = 0;
    bool prelaunched = false;
    int n_left = 0;                 // left-side keypoint list the matches refer to
    const int32_t* left_xy = nullptr;
  } job;
  std::vector<pano_ctx*> slots;  // a batch lane's pipeline slots (child contexts): pairs in flight between A and B
  int fold_mode = 0;             // 0: the reference's fold (re-detects on the growing panorama); 1: incremental (opt-in)
  DevKeypoints kpP[2];           // incremental fold: the panorama's carried keypoint list (double buffered)
  const DevKeypoints* left_kp = nullptr;   // when set, stage A uses this list for the left image instead of detecting
  std::future<int> async_job;    // pano_stitch_pair_async: the pair running on the context's worker
  cudaEvent_t ev_async = nullptr;
  void* async_stream = nullptr;
  std::vector<pano_ctx*> lanes;  // child contexts (own stream + scratch) used by pano_stitch_batch
};

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int fail(pano_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}

#define API_TRY(c)                                  \
  if (!(c)) return PANO_ERR_INVALID;                \
  t_prof = &(c)->prof;                              \
  try {                                             \
    PANO_CUDA(cudaSetDevice((c)->device));

#define API_CATCH(c)                                                                        \
  }                                                                                         \
  catch (const CudaError& e) {                                                              \
    char buf[512];                                                                          \
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e.e,                  \
             cudaGetErrorString(e.e), e.file, e.line, e.what);                              \
    (c)->err = buf;                                                                         \
    cudaGetLastError();                                                                     \
    return PANO_ERR_CUDA;                                                                   \
  }                                                                                         \
  catch (const std::exception& e) {                                                         \
    (c)->err = e.what();                                                                    \
    return PANO_ERR_CUDA;                                                                   \
  }

// Returns a device view of an image; host images are copied into staging slot `slot`.
DevImage to_device(pano_ctx* c, const uint8_t* p, int w, int h, size_t stride, int mem, int slot) {
  DevImage d;
  d.w = w;
  d.h = h;
  if (mem == PANO_MEM_DEVICE) {
    d.p = p;
    d.stride = stride;
    return d;
  }
  size_t pitch = align_up((size_t)w * 3, 256);
  c->up[slot].reserve(pitch * h);
  PANO_CUDA(cudaMemcpy2DAsync(c->up[slot].p, pitch, p, stride, (size_t)w * 3, h, cudaMemcpyHostToDevice, c->st));
  d.p = c->up[slot].as<uint8_t>();
  d.stride = pitch;
  return d;
}

// 2-D image copy; one flat cudaMemcpyAsync when source and destination rows are contiguous and equally pitched
// (3840 x 3 = 11520 bytes is already a multiple of the engine's 256-byte pitch), which spares the driver the
// per-row descriptor work of cudaMemcpy2DAsync on the end-to-end path
void copy_image_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t row_bytes, int rows,
                      cudaMemcpyKind kind, cudaStream_t st) {
  // diagnosis only (PANO_DEBUG_SKIP_COPY=h2d|d2h): which direction of the end-to-end path costs what
  static const int skip = [] { const char* e = getenv("PANO_DEBUG_SKIP_COPY"); return !e ? 0 : (e[0] == 'h' ? 1 : (e[0] == 'd' ? 2 : 0)); }();
  if ((skip == 1 && kind == cudaMemcpyHostToDevice) || (skip == 2 && kind == cudaMemcpyDeviceToHost)) return;
  if (dpitch == spitch && rows > 0 && (spitch == row_bytes || rows == 1)) {
    PANO_CUDA(cudaMemcpyAsync(dst, src, spitch * (size_t)(rows - 1) + row_bytes, kind, st));
  } else {
    PANO_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, rows, kind, st));
  }
}

const void* upload(pano_ctx* c, DevBuf& buf, const void* p, size_t bytes, int mem) {
  if (mem == PANO_MEM_DEVICE || bytes == 0) return p;
  buf.reserve(bytes);
  PANO_CUDA(cudaMemcpyAsync(buf.p, p, bytes, cudaMemcpyHostToDevice, c->st));
  return buf.p;
}

bool valid_image(const uint8_t* p, int w, int h, size_t stride) {
  return p && w > 0 && h > 0 && stride >= (size_t)w * 3;
}

int check_harris(const pano_harris_opts& o) {
  if (o.nms_neighborhood < 1 || o.nms_neighborhood % 2 == 0 || o.nms_neighborhood > 15) return PANO_ERR_UNSUPPORTED;
  if (o.patch_size != 1 && o.patch_size != 3 && o.patch_size != 5) return PANO_ERR_UNSUPPORTED;
  return PANO_OK;
}

// fused calls with the opt-in 2-NN matcher: the binary descriptor is defined on 5 x 5 patches only
int check_match_mode(const pano_ctx* c, const pano_harris_opts& o) {
  if (c->match_mode == 1 && c->knn_mode.descriptor == PANO_KNN_BINARY && o.patch_size != 5) return PANO_ERR_UNSUPPORTED;
  return PANO_OK;
}

void run_matcher(pano_ctx* c, const DevDescriptors& q, const DevDescriptors& t) {
  c->best.reserve(sizeof(unsigned long long) * (size_t)std::max(q.count, 1));
  if (c->matcher == 0 && match_tc_available())
    match_tc_device(c->st, q, t, c->best.as<unsigned long long>(), c->ms.tc_err, c->errw.as<int>());
  else
    match_simt_device(c->st, q, t, c->best.as<unsigned long long>());
}

// match stage on device-resident inputs; leaves matches in c->matches; returns count.
// query_ready: c->dQ already holds the query descriptors (nqi of them).
int match_on_device(pano_ctx* c, const int32_t* kq, int nq, const int32_t* kt, int nt, const DevImage& iq,
                    const DevImage& it, const pano_harris_opts& o, int offset, bool query_ready = false, int nqi_ready = 0) {
  int nqi = query_ready ? nqi_ready : build_descriptors_device(c->st, iq, kq, nq, o.patch_size, c->ms, c->dQ, c->pin);
  int nti = build_descriptors_device(c->st, it, kt, nt, o.patch_size, c->ms, c->dT, c->pin);
  if (nqi == 0 || nti == 0) return 0;
  run_matcher(c, c->dQ, c->dT);
  c->matches.reserve(sizeof(pano_dmatch) * (size_t)nqi);
  return emit_matches_device(c->st, c->dQ, c->dT, c->best.as<unsigned long long>(), o.max_ssd_thresh, offset,
                             o.patch_size, c->ms, c->matches.as<pano_dmatch>(), c->pin, c->errw.as<int>());
}

// Turns a non-zero device error word into a failed call: message, word cleared, and a matcher whose pipeline
// aborted is not used again by this process (the SIMT matcher takes over on the next call).
int fail_errw(pano_ctx* c, int errw) {
  char buf[256];
  snprintf(buf, sizeof buf, "device error word 0x%x:%s%s%s", errw,
           (errw & PANO_ERRW_TC_ABORT) ? " tensor-core matcher pipeline timed out;" : "",
           (errw & PANO_ERRW_NO_BEST) ? " a query row has no nearest neighbour (matcher did not finish);" : "",
           (errw & PANO_ERRW_BAD_INDEX) ? " a match index is outside the keypoint lists;" : "");
  c->err = buf;
  if (errw & (PANO_ERRW_TC_ABORT | PANO_ERRW_NO_BEST)) match_tc_disable();
  cudaMemsetAsync(c->errw.p, 0, sizeof(int), c->st);
  return (errw & ~PANO_ERRW_BAD_INDEX) ? PANO_ERR_CUDA : PANO_ERR_INVALID;
}

// waits for the context's stream and returns the device error word as it was at that point
int wait_errw(pano_ctx* c) {
  int* slot = reinterpret_cast<int*>(c->pin.as<char>() + 1024);
  PANO_CUDA(cudaMemcpyAsync(slot, c->errw.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  PANO_CUDA(stream_wait(c->st));
  return *slot;
}

RansacResult ransac_retry(pano_ctx* c, const int32_t* kp1, const int32_t* kp2, const pano_dmatch* m, int n,
                          const pano_ransac_opts& o, int32_t* samples, int32_t* counts, uint8_t* mask,
                          bool replay_prelaunched = false) {
  RansacResult r;
  int scale = 1;
  if (replay_prelaunched) {
    // the replay for exactly n matches is running (or done) on the side stream: join it and solve on top of it
    PANO_CUDA(cudaEventRecord(c->rs.ev_join, c->rs.side));
    PANO_CUDA(cudaStreamWaitEvent(c->st, c->rs.ev_join, 0));
    r = ransac_device(c->st, kp1, kp2, m, n, o, c->seed, c->mt, c->rs, c->pin, samples, counts, mask, 1,
                      c->replay_target, c->replay_mode, /*phase=*/2);
    if (r.status >= 0 || r.errw) return r;
    scale = 2;   // a speculation window was missed: re-run the whole thing wider, in order
  }
  for (;;) {
    r = ransac_device(c->st, kp1, kp2, m, n, o, c->seed, c->mt, c->rs, c->pin, samples, counts, mask, scale,
                      c->replay_target, c->replay_mode);
    if (r.status >= 0 || r.errw || scale >= 16) break;
    scale *= 2;  // a speculation window was missed: re-run wider (exactness is never traded)
  }
  if (r.status < 0 && !r.errw) {
    c->err = "shuffle replay could not be resolved";
    r.status = PANO_ERR_CUDA;
  }
  return r;
}

// Enqueues the shuffle replay for n matches on the context's side stream (it needs the COUNT only).
void prelaunch_replay(pano_ctx* c, int n, const pano_ransac_opts& o) {
  if (!c->rs.side) {
    PANO_CUDA(cudaStreamCreateWithFlags(&c->rs.side, cudaStreamNonBlocking));
    PANO_CUDA(cudaEventCreateWithFlags(&c->rs.ev_fork, cudaEventDisableTiming));
    PANO_CUDA(cudaEventCreateWithFlags(&c->rs.ev_join, cudaEventDisableTiming));
  }
  PANO_CUDA(cudaEventRecord(c->rs.ev_fork, c->st));          // earlier users of the replay scratch have finished
  PANO_CUDA(cudaStreamWaitEvent(c->rs.side, c->rs.ev_fork, 0));
  ransac_device(c->rs.side, nullptr, nullptr, nullptr, n, o, c->seed, c->mt, c->rs, c->pin, nullptr, nullptr, nullptr, 1,
                c->replay_target, c->replay_mode, /*phase=*/1);
}

void fill_canvas_info(const CanvasGeom& g, pano_canvas_info* info) {
  info->canvas_w = g.cw;
  info->canvas_h = g.ch;
  info->left_x = g.offx;
  info->left_y = g.offy;
  memcpy(info->TH, g.TH, sizeof g.TH);
}

// ref: src/serial/main.cpp:311-391, in two halves so that a batch lane can keep several pairs in flight:
//   stage A  detection of both images, descriptors, matching - and, as soon as the right image's in-border keypoints
//            are counted, the shuffle replay on the context's side stream (it depends on nothing but that count);
//   stage B  joins the replay, solves (DLT, scoring, selection), reads H back, canvas geometry, warp + overlay.
// A single pair runs A then B back to back (the replay then overlaps the left image's detection and the matching);
// pano_stitch_batch runs B of an earlier pair after A of a later one, so a slow but cheap replay (the resident,
// one-CTA formulation) is never waited for.  left/right are device views that must stay valid until B returns.
// On success the new canvas is in c->canvas[c->cur].
int pair_stage_a(pano_ctx* c, const DevImage& L, const DevImage& R, const pano_harris_opts& ho,
                 const pano_ransac_opts& ro, pano_pair_result* res) {
  memset(res, 0, sizeof *res);
  res->best_iteration = -1;
  cudaStream_t st = c->st;
  c->job.L = L;
  c->job.R = R;
  c->job.m = 0;
  c->job.prelaunched = false;
  PANO_CUDA(cudaEventRecord(c->ev[0], st));
  // 1. corner detection (ref :316-317) and 2. matching: right = query, left = train (ref :320).  The right image
  // goes first: once its in-border keypoints are counted the number of matches M is known (every query keypoint
  // gets its nearest neighbour; the max-SSD filter is vacuous at the reference's 1e8).  If the filter does remove
  // matches the pre-launched replay is discarded and RANSAC runs in order.
  res->n_kp_right = harris_detect_device(st, R, ho, c->hs, c->kpR, c->pin);
  bool prelaunched = false;
  int nqi = 0;
  // (with per-kernel profiling on, the replay is not overlapped: every kernel is then timed alone on the GPU; with the
  // opt-in 2-NN matcher the match count is only known after the ratio test)
  const bool try_overlap = c->overlap_replay && !c->prof.on && ro.num_samples == 4 && c->match_mode == 0;
  if (try_overlap) {
    nqi = build_descriptors_device(st, R, c->kpR.xy.as<int32_t>(), c->kpR.count, ho.patch_size, c->ms, c->dQ, c->pin);
    if (nqi >= ro.num_samples && ro.num_iterations > 0) {
      prelaunch_replay(c, nqi, ro);
      prelaunched = true;
    }
  }
  const DevKeypoints* kl = c->left_kp;     // incremental fold: the carried list, no detection on the panorama
  if (!kl) {
    res->n_kp_left = harris_detect_device(st, L, ho, c->hs, c->kpL, c->pin);
    kl = &c->kpL;
  } else {
    res->n_kp_left = kl->count;
  }
  c->job.n_left = kl->count;
  c->job.left_xy = kl->xy.as<int32_t>();
  PANO_CUDA(cudaEventRecord(c->ev[1], st));
  int m;
  if (c->match_mode == 0) {
    m = match_on_device(c, c->kpR.xy.as<int32_t>(), c->kpR.count, kl->xy.as<int32_t>(), kl->count, R, L, ho, 0, try_overlap,
                        nqi);
  } else {
    // opt-in: 2 nearest neighbours + Lowe's ratio test (knn.cu); the ratio-tested matches become RANSAC's input
    pano_knn_opts ko = c->knn_mode;
    ko.patch_size = ho.patch_size;
    const int nq = build_descriptors_device(st, R, c->kpR.xy.as<int32_t>(), c->kpR.count, ko.patch_size, c->ms, c->dQ, c->pin);
    const int nt = build_descriptors_device(st, L, kl->xy.as<int32_t>(), kl->count, ko.patch_size, c->ms, c->dT, c->pin);
    m = 0;
    if (nq > 0 && nt > 0)
      m = match_knn_device(st, R, L, c->kpR.xy.as<int32_t>(), kl->xy.as<int32_t>(), c->dQ, c->dT, ko,
                           c->matcher == 0 && match_tc_available(), c->ms, c->best, c->ks, c->pin, c->errw.as<int>());
    c->matches.reserve(sizeof(pano_dmatch) * (size_t)std::max(m, 1));
    if (m > 0)
      PANO_CUDA(cudaMemcpyAsync(c->matches.p, c->ks.out.p, sizeof(pano_dmatch) * (size_t)m, cudaMemcpyDeviceToDevice, st));
  }
  res->n_matches = m;
  PANO_CUDA(cudaEventRecord(c->ev[2], st));
  if (prelaunched && m != nqi) {   // not the count the replay was started for: let it drain, then ignore it
    PANO_CUDA(cudaEventRecord(c->rs.ev_join, c->rs.side));
    PANO_CUDA(cudaStreamWaitEvent(st, c->rs.ev_join, 0));
    prelaunched = false;
  }
  c->job.m = m;
  c->job.prelaunched = prelaunched;
  return PANO_OK;
}

int pair_stage_b(pano_ctx* c, const pano_harris_opts& ho, const pano_ransac_opts& ro, pano_pair_result* res,
                 bool homography_only) {
  (void)ho;
  cudaStream_t st = c->st;
  const DevImage& L = c->job.L;
  const DevImage& R = c->job.R;
  const int m = c->job.m;
  PANO_CUDA(cudaEventRecord(c->ev[5], st));
  auto finish = [&](int status) {
    int* ew_slot = reinterpret_cast<int*>(c->pin.as<char>() + 1024);
    PANO_CUDA(cudaMemcpyAsync(ew_slot, c->errw.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    PANO_CUDA(cudaEventRecord(c->ev[4], st));
    PANO_CUDA(cudaEventSynchronize(c->ev[4]));
    if (*ew_slot) status = fail_errw(c, *ew_slot);
    float ta = 0, tb = 0;
    cudaEventElapsedTime(&res->ms_detect, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&res->ms_match, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&ta, c->ev[0], c->ev[2]);
    cudaEventElapsedTime(&tb, c->ev[5], c->ev[4]);
    res->ms_total = ta + tb;     // (a batch lane runs other pairs' stages between A and B: not counted here)
    res->status = status;
    return status;
  };
  if (m == 0) return finish(PANO_ERR_NO_MATCHES);
  // 3. RANSAC (ref :327-332)
  c->rs.n1 = c->kpR.count;
  c->rs.n2 = c->job.n_left;
  RansacResult rr = ransac_retry(c, c->kpR.xy.as<int32_t>(), c->job.left_xy, c->matches.as<pano_dmatch>(),
                                 m, ro, nullptr, nullptr, nullptr, c->job.prelaunched);
  PANO_CUDA(cudaEventRecord(c->ev[3], st));
  if (rr.errw) rr.status = fail_errw(c, rr.errw);
  if (rr.status != PANO_OK) {
    int s = finish(rr.status);
    cudaEventElapsedTime(&res->ms_ransac, c->ev[5], c->ev[3]);
    return s;
  }
  memcpy(res->H, rr.H, sizeof rr.H);
  res->best_inliers = rr.best_count;
  res->best_iteration = rr.best_iter;
  // 4. canvas geometry (ref :335-369), warp + overlay (ref :371-386)
  CanvasGeom g;
  canvas_geometry(L.w, L.h, R.w, R.h, rr.H, &g);
  fill_canvas_info(g, &res->canvas);
  if (homography_only) {  // chain mode: the canvas is composed later from all homographies
    int s = finish(PANO_OK);
    cudaEventElapsedTime(&res->ms_ransac, c->ev[5], c->ev[3]);
    return s;
  }
  if (!g.ok) {
    int s = finish(PANO_ERR_ROI);
    cudaEventElapsedTime(&res->ms_ransac, c->ev[5], c->ev[3]);
    return s;
  }
  int nxt = 1 - c->cur;
  size_t pitch = align_up((size_t)g.cw * 3, 256);
  c->canvas[nxt].reserve(pitch * (size_t)g.ch);
  warp_overlay_device(st, L, R, g, c->canvas[nxt].as<uint8_t>(), pitch);
  if (c->pack_tight && pitch != (size_t)g.cw * 3) {
    c->tight.reserve((size_t)g.cw * 3 * (size_t)g.ch + 16);
    pack_rows_device(st, c->canvas[nxt].as<uint8_t>(), pitch, (size_t)g.cw * 3, g.ch, c->tight.as<uint8_t>());
  }
  if (int s = finish(PANO_OK)) return s;
  cudaEventElapsedTime(&res->ms_ransac, c->ev[5], c->ev[3]);
  cudaEventElapsedTime(&res->ms_warp, c->ev[3], c->ev[4]);
  c->cur = nxt;
  c->cw = g.cw;
  c->ch = g.ch;
  c->cstride = pitch;
  c->has_canvas = true;
  return PANO_OK;
}

int stitch_pair_device(pano_ctx* c, const DevImage& L, const DevImage& R, const pano_harris_opts& ho,
                       const pano_ransac_opts& ro, pano_pair_result* res, bool homography_only = false) {
  int s = pair_stage_a(c, L, R, ho, ro, res);
  if (s != PANO_OK) return s;
  return pair_stage_b(c, ho, ro, res, homography_only);
}

}  // namespace

// worker launch shared by the asynchronous forms of the stage / fold / batch calls (below)
namespace {
template <class F>
int start_async(pano_ctx* c, void* stream, F fn) {
  API_TRY(c)
  if (c->async_job.valid()) return PANO_ERR_BUSY;
  if (!c->ev_async) PANO_CUDA(cudaEventCreateWithFlags(&c->ev_async, cudaEventDisableTiming));
  c->async_stream = stream;
  if (stream) {
    PANO_CUDA(cudaEventRecord(c->ev_async, (cudaStream_t)stream));
    PANO_CUDA(cudaStreamWaitEvent(c->st, c->ev_async, 0));
  }
  c->async_job = std::async(std::launch::async, fn);
  return PANO_OK;
  API_CATCH(c)
}
}  // namespace

extern "C" {

void pano_default_harris_opts(pano_harris_opts* o) {
  o->k = 0.04;
  o->nms_thresh = 1e6;
  o->nms_neighborhood = 3;
  o->patch_size = 5;
  o->max_ssd_thresh = 1e8;
}

void pano_default_knn_opts(pano_knn_opts* o) {
  o->patch_size = 5;
  o->descriptor = PANO_KNN_PATCH_SSD;
  o->ratio = 0.75;
}

void pano_default_ransac_opts(pano_ransac_opts* o) {
  o->num_iterations = 1000;
  o->num_samples = 4;
  o->distance_threshold = 3.0;
}

const char* pano_version(void) { return "pano_b200 0.1 (sm_100a)"; }

int pano_create(int device, uint32_t seed, pano_ctx** out) {
  if (!out) return PANO_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    cudaGetLastError();
    return PANO_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
    cudaGetLastError();
    return PANO_ERR_NO_DEVICE;  // kernels are built for sm_100a only
  }
  pano_ctx* c = new (std::nothrow) pano_ctx();
  if (!c) return PANO_ERR_INVALID;
  c->device = device;
  c->seed = seed;
  try {
    PANO_CUDA(cudaSetDevice(device));
    PANO_CUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    for (auto& e : c->ev) PANO_CUDA(cudaEventCreate(&e));
    c->pin.reserve(4096);
    c->errw.reserve(256);
    PANO_CUDA(cudaMemsetAsync(c->errw.p, 0, 256, c->st));
    c->rs.errw = c->errw.as<int>();
  } catch (const CudaError&) {
    cudaGetLastError();
    delete c;
    return PANO_ERR_CUDA;
  }
  *out = c;
  return PANO_OK;
}

void pano_destroy(pano_ctx* c) {
  if (!c) return;
  for (pano_ctx* l : c->lanes) pano_destroy(l);
  c->lanes.clear();
  for (pano_ctx* l : c->slots) pano_destroy(l);
  c->slots.clear();
  if (c->async_job.valid()) c->async_job.wait();
  cudaSetDevice(c->device);
  if (c->ev_async) cudaEventDestroy(c->ev_async);
  for (auto& r : c->prof.pool) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  if (c->st) cudaStreamSynchronize(c->st);
  if (c->rs.side) { cudaStreamSynchronize(c->rs.side); cudaStreamDestroy(c->rs.side); }
  if (c->rs.ev_fork) cudaEventDestroy(c->rs.ev_fork);
  if (c->rs.ev_join) cudaEventDestroy(c->rs.ev_join);
  if (c->st_up) { cudaStreamSynchronize(c->st_up); cudaStreamDestroy(c->st_up); }
  if (c->st_down) { cudaStreamSynchronize(c->st_down); cudaStreamDestroy(c->st_down); }
  for (int q = 0; q < 2; q++) {
    if (c->ev_up[q]) cudaEventDestroy(c->ev_up[q]);
    if (c->ev_down[q]) cudaEventDestroy(c->ev_down[q]);
    c->upq[q][0].release();
    c->upq[q][1].release();
  }
  DevBuf* bufs[] = {&c->errw, &c->up[0], &c->up[1], &c->kpup[0], &c->kpup[1], &c->mup, &c->hs.resp, &c->hs.mask, &c->hs.rowcnt,
                    &c->hs.rowoff, &c->hs.total, &c->kpL.xy, &c->kpR.xy, &c->kpP[0].xy, &c->kpP[1].xy, &c->ms.flags, &c->ms.tmp, &c->ms.best,
                    &c->ms.cnt, &c->ms.mflags, &c->ms.midx, &c->ms.mtmp, &c->ms.tc_err, &c->dQ.desc, &c->dQ.norm, &c->dQ.orig,
                    &c->dT.desc, &c->dT.norm, &c->dT.orig, &c->best, &c->matches, &c->rs.pts, &c->rs.thr,
                    &c->rs.cand_off, &c->rs.cand_samp, &c->rs.base, &c->rs.samples, &c->rs.Hs, &c->rs.valid,
                    &c->rs.counts, &c->rs.result, &c->rs.mask, &c->rs.plan, &c->rs.pts_bits, &c->mt.x, &c->mt.state, &c->tmp[0],
                    &c->tmp[1], &c->tmp[2], &c->canvas[0], &c->canvas[1], &c->tight, &c->ks.best2, &c->ks.qbits, &c->ks.tbits,
                    &c->ks.rec, &c->ks.second, &c->ks.out, &c->ks.out2, &c->ks.flags, &c->ks.idx, &c->ks.cnt, &c->ks.tmp};
  for (DevBuf* b : bufs) b->release();
  c->pin.release();
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  if (c->st && c->owns_stream) cudaStreamDestroy(c->st);
  delete c;
}

int pano_set_seed(pano_ctx* c, uint32_t seed) {
  if (!c) return PANO_ERR_INVALID;
  c->seed = seed;
  return PANO_OK;
}

int pano_set_matcher(pano_ctx* c, int which) {
  if (!c || (which != 0 && which != 1)) return PANO_ERR_INVALID;
  c->matcher = which;
  return PANO_OK;
}

int pano_set_match_mode(pano_ctx* c, int mode, double ratio, int descriptor) {
  if (!c || (mode != 0 && mode != 1)) return PANO_ERR_INVALID;
  if (mode == 1) {
    if (!(ratio > 0.0) || ratio > 1.0) return fail(c, PANO_ERR_INVALID, "pano_set_match_mode: ratio outside (0, 1]");
    if (descriptor != PANO_KNN_PATCH_SSD && descriptor != PANO_KNN_BINARY)
      return fail(c, PANO_ERR_UNSUPPORTED, "pano_set_match_mode: unknown descriptor");
    c->knn_mode.ratio = ratio;
    c->knn_mode.descriptor = descriptor;
  }
  c->match_mode = mode;
  return PANO_OK;
}

int pano_set_fold_mode(pano_ctx* c, int mode) {
  if (!c || (mode != 0 && mode != 1)) return PANO_ERR_INVALID;
  c->fold_mode = mode;
  return PANO_OK;
}

int pano_set_replay_mode(pano_ctx* c, int mode) {
  if (!c || (mode != 0 && mode != 1)) return PANO_ERR_INVALID;
  c->replay_mode = mode;
  return PANO_OK;
}

const char* pano_last_error(const pano_ctx* c) { return c ? c->err.c_str() : "null context"; }

uint64_t pano_kernel_launches(const pano_ctx*) { return g_kernel_launches.load(); }

int pano_detect(pano_ctx* c, const uint8_t* bgr, int w, int h, size_t stride, int mem, const pano_harris_opts* opts,
                int32_t* xy_out, int cap, int* count) {
  API_TRY(c)
  if (!valid_image(bgr, w, h, stride) || !opts || !count) return fail(c, PANO_ERR_INVALID, "pano_detect: bad argument");
  if (int e = check_harris(*opts)) return fail(c, e, "pano_detect: unsupported option");
  DevImage img = to_device(c, bgr, w, h, stride, mem, 0);
  int n = harris_detect_device(c->st, img, *opts, c->hs, c->kpL, c->pin);
  *count = n;
  int ncopy = std::min(n, cap);
  if (xy_out && ncopy > 0) {
    PANO_CUDA(cudaMemcpyAsync(xy_out, c->kpL.xy.p, sizeof(int32_t) * 2 * (size_t)ncopy,
                              mem == PANO_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->st));
  }
  PANO_CUDA(stream_wait(c->st));
  return (xy_out && n > cap) ? PANO_ERR_CAPACITY : PANO_OK;
  API_CATCH(c)
}

int pano_harris_response(pano_ctx* c, const uint8_t* bgr, int w, int h, size_t stride, int mem, double k,
                         double* resp_out) {
  API_TRY(c)
  if (!valid_image(bgr, w, h, stride) || !resp_out) return fail(c, PANO_ERR_INVALID, "pano_harris_response: bad argument");
  DevImage img = to_device(c, bgr, w, h, stride, mem, 0);
  size_t bytes = sizeof(double) * (size_t)w * h;
  double* dst = resp_out;
  if (mem == PANO_MEM_HOST) {
    c->hs.resp.reserve(bytes);
    dst = c->hs.resp.as<double>();
  }
  harris_response_device(c->st, img, k, dst);
  if (mem == PANO_MEM_HOST) PANO_CUDA(cudaMemcpyAsync(resp_out, dst, bytes, cudaMemcpyDeviceToHost, c->st));
  PANO_CUDA(stream_wait(c->st));
  return PANO_OK;
  API_CATCH(c)
}

int pano_convolve_f64(pano_ctx* c, const double* in, int w, int h, const double* kernel, int ksize, int mem,
                      double* out) {
  API_TRY(c)
  if (!in || !out || !kernel || w <= 0 || h <= 0 || ksize < 1 || ksize % 2 == 0)
    return fail(c, PANO_ERR_INVALID, "pano_convolve_f64: bad argument");
  size_t bytes = sizeof(double) * (size_t)w * h;
  const double* din = (const double*)upload(c, c->tmp[0], in, bytes, mem);
  const double* dk = (const double*)upload(c, c->tmp[1], kernel, sizeof(double) * ksize * ksize, mem);
  double* dout = out;
  if (mem == PANO_MEM_HOST) {
    c->tmp[2].reserve(bytes);
    dout = c->tmp[2].as<double>();
  }
  convolve_f64_device(c->st, din, w, h, dk, ksize, dout);
  if (mem == PANO_MEM_HOST) PANO_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, c->st));
  PANO_CUDA(stream_wait(c->st));
  return PANO_OK;
  API_CATCH(c)
}

int pano_match(pano_ctx* c, const int32_t* kp_query, int n_query, const int32_t* kp_train, int n_train,
               const uint8_t* img_query, int wq, int hq, size_t stride_q, const uint8_t* img_train, int wt, int ht,
               size_t stride_t, int mem, const pano_harris_opts* opts, int offset, pano_dmatch* out, int cap,
               int* count) {
  API_TRY(c)
  if (!valid_image(img_query, wq, hq, stride_q) || !valid_image(img_train, wt, ht, stride_t) || !opts || !count ||
      n_query < 0 || n_train < 0 || (n_query > 0 && !kp_query) || (n_train > 0 && !kp_train))
    return fail(c, PANO_ERR_INVALID, "pano_match: bad argument");
  if (int e = check_harris(*opts)) return fail(c, e, "pano_match: unsupported option");
  DevImage iq = to_device(c, img_query, wq, hq, stride_q, mem, 0);
  DevImage it = to_device(c, img_train, wt, ht, stride_t, mem, 1);
  const int32_t* kq = (const int32_t*)upload(c, c->kpup[0], kp_query, sizeof(int32_t) * 2 * (size_t)n_query, mem);
  const int32_t* kt = (const int32_t*)upload(c, c->kpup[1], kp_train, sizeof(int32_t) * 2 * (size_t)n_train, mem);
  int m = match_on_device(c, kq, n_query, kt, n_train, iq, it, *opts, offset);
  *count = m;
  int ncopy = std::min(m, cap);
  if (out && ncopy > 0)
    PANO_CUDA(cudaMemcpyAsync(out, c->matches.p, sizeof(pano_dmatch) * (size_t)ncopy,
                              mem == PANO_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->st));
  if (int ew = wait_errw(c)) return fail_errw(c, ew);
  return (out && m > cap) ? PANO_ERR_CAPACITY : PANO_OK;
  API_CATCH(c)
}

// north star item (c), opt-in: 2 nearest neighbours + Lowe's ratio test (knn.cu).  Not the reference's matcher.
int pano_match_knn(pano_ctx* c, const int32_t* kp_query, int n_query, const int32_t* kp_train, int n_train,
                   const uint8_t* img_query, int wq, int hq, size_t stride_q, const uint8_t* img_train, int wt, int ht,
                   size_t stride_t, int mem, const pano_knn_opts* opts, pano_dmatch* out, float* second_out, int cap,
                   int* count) {
  API_TRY(c)
  if (!valid_image(img_query, wq, hq, stride_q) || !valid_image(img_train, wt, ht, stride_t) || !opts || !count ||
      n_query < 0 || n_train < 0 || (n_query > 0 && !kp_query) || (n_train > 0 && !kp_train) || cap < 0)
    return fail(c, PANO_ERR_INVALID, "pano_match_knn: bad argument");
  if (!(opts->ratio > 0.0) || opts->ratio > 1.0) return fail(c, PANO_ERR_INVALID, "pano_match_knn: ratio outside (0, 1]");
  if (opts->descriptor != PANO_KNN_PATCH_SSD && opts->descriptor != PANO_KNN_BINARY)
    return fail(c, PANO_ERR_UNSUPPORTED, "pano_match_knn: unknown descriptor");
  if ((opts->patch_size != 1 && opts->patch_size != 3 && opts->patch_size != 5) ||
      (opts->descriptor == PANO_KNN_BINARY && opts->patch_size != 5))
    return fail(c, PANO_ERR_UNSUPPORTED, "pano_match_knn: unsupported patch size");
  *count = 0;
  DevImage iq = to_device(c, img_query, wq, hq, stride_q, mem, 0);
  DevImage it = to_device(c, img_train, wt, ht, stride_t, mem, 1);
  const int32_t* kq = (const int32_t*)upload(c, c->kpup[0], kp_query, sizeof(int32_t) * 2 * (size_t)n_query, mem);
  const int32_t* kt = (const int32_t*)upload(c, c->kpup[1], kp_train, sizeof(int32_t) * 2 * (size_t)n_train, mem);
  // the candidates of both sides: in-border keypoints and their patch descriptors, exactly as in pano_match
  const int nqi = build_descriptors_device(c->st, iq, kq, n_query, opts->patch_size, c->ms, c->dQ, c->pin);
  const int nti = build_descriptors_device(c->st, it, kt, n_train, opts->patch_size, c->ms, c->dT, c->pin);
  int m = 0;
  if (nqi > 0 && nti > 0)
    m = match_knn_device(c->st, iq, it, kq, kt, c->dQ, c->dT, *opts, c->matcher == 0 && match_tc_available(), c->ms, c->best,
                         c->ks, c->pin, c->errw.as<int>());
  *count = m;
  const int ncopy = std::min(m, cap);
  const cudaMemcpyKind kind = mem == PANO_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (out && ncopy > 0) PANO_CUDA(cudaMemcpyAsync(out, c->ks.out.p, sizeof(pano_dmatch) * (size_t)ncopy, kind, c->st));
  if (second_out && ncopy > 0) PANO_CUDA(cudaMemcpyAsync(second_out, c->ks.out2.p, sizeof(float) * (size_t)ncopy, kind, c->st));
  if (int ew = wait_errw(c)) return fail_errw(c, ew);
  return (out && m > cap) ? PANO_ERR_CAPACITY : PANO_OK;
  API_CATCH(c)
}

int pano_ransac(pano_ctx* c, const int32_t* kp1, int n1, const int32_t* kp2, int n2, const pano_dmatch* matches,
                int n_matches, int mem, const pano_ransac_opts* opts, double H_out[9], int* best_inliers,
                int* best_iteration, int32_t* samples_out, int32_t* counts_out, uint8_t* inlier_mask_out) {
  API_TRY(c)
  if (!opts || !H_out || n1 < 0 || n2 < 0 || n_matches < 0 || (n_matches > 0 && (!kp1 || !kp2 || !matches)))
    return fail(c, PANO_ERR_INVALID, "pano_ransac: bad argument");
  if (opts->num_samples != 4) return fail(c, PANO_ERR_UNSUPPORTED, "pano_ransac: only num_samples == 4 is supported");
  if (best_inliers) *best_inliers = 0;
  if (best_iteration) *best_iteration = -1;
  if (n_matches < opts->num_samples || opts->num_iterations <= 0) return PANO_ERR_TOO_FEW_MATCHES;
  const int32_t* d1 = (const int32_t*)upload(c, c->kpup[0], kp1, sizeof(int32_t) * 2 * (size_t)n1, mem);
  const int32_t* d2 = (const int32_t*)upload(c, c->kpup[1], kp2, sizeof(int32_t) * 2 * (size_t)n2, mem);
  const pano_dmatch* dm = (const pano_dmatch*)upload(c, c->mup, matches, sizeof(pano_dmatch) * (size_t)n_matches, mem);
  c->rs.n1 = n1;   // the indices of the caller's matches are checked against the caller's counts on the device
  c->rs.n2 = n2;
  RansacResult r = ransac_retry(c, d1, d2, dm, n_matches, *opts, samples_out, counts_out, inlier_mask_out);
  if (r.errw) return fail_errw(c, r.errw);
  if (best_inliers) *best_inliers = r.best_count;
  if (best_iteration) *best_iteration = r.best_iter;
  if (r.status == PANO_OK) memcpy(H_out, r.H, sizeof r.H);
  return r.status;
  API_CATCH(c)
}

int pano_canvas_geometry(int wl, int hl, int wr, int hr, const double H[9], pano_canvas_info* out) {
  if (!H || !out || wl <= 0 || hl <= 0 || wr <= 0 || hr <= 0) return PANO_ERR_INVALID;
  CanvasGeom g;
  canvas_geometry(wl, hl, wr, hr, H, &g);
  fill_canvas_info(g, out);
  return g.ok ? PANO_OK : PANO_ERR_ROI;
}

int pano_warp_overlay(pano_ctx* c, const uint8_t* left, int wl, int hl, size_t stride_l, const uint8_t* right, int wr,
                      int hr, size_t stride_r, int mem, const double H[9], uint8_t* canvas_out, size_t canvas_stride,
                      size_t canvas_cap_bytes, pano_canvas_info* info) {
  API_TRY(c)
  if (!valid_image(left, wl, hl, stride_l) || !valid_image(right, wr, hr, stride_r) || !H || !canvas_out)
    return fail(c, PANO_ERR_INVALID, "pano_warp_overlay: bad argument");
  CanvasGeom g;
  canvas_geometry(wl, hl, wr, hr, H, &g);
  if (info) fill_canvas_info(g, info);
  if (!g.ok) return PANO_ERR_ROI;
  if (canvas_stride < (size_t)g.cw * 3 || canvas_cap_bytes < canvas_stride * (size_t)(g.ch - 1) + (size_t)g.cw * 3)
    return PANO_ERR_CAPACITY;
  DevImage L = to_device(c, left, wl, hl, stride_l, mem, 0);
  DevImage R = to_device(c, right, wr, hr, stride_r, mem, 1);
  if (mem == PANO_MEM_DEVICE) {
    warp_overlay_device(c->st, L, R, g, canvas_out, canvas_stride);
  } else {
    size_t pitch = align_up((size_t)g.cw * 3, 256);
    c->tmp[0].reserve(pitch * (size_t)g.ch);
    warp_overlay_device(c->st, L, R, g, c->tmp[0].as<uint8_t>(), pitch);
    PANO_CUDA(cudaMemcpy2DAsync(canvas_out, canvas_stride, c->tmp[0].p, pitch, (size_t)g.cw * 3, g.ch,
                                cudaMemcpyDeviceToHost, c->st));
  }
  PANO_CUDA(stream_wait(c->st));
  return PANO_OK;
  API_CATCH(c)
}

int pano_warp_perspective(pano_ctx* c, const uint8_t* src, int w, int h, size_t stride, int mem, const double M[9],
                          uint8_t* dst, int dw, int dh, size_t dstride) {
  API_TRY(c)
  if (!valid_image(src, w, h, stride) || !M || !dst || dw <= 0 || dh <= 0 || dstride < (size_t)dw * 3)
    return fail(c, PANO_ERR_INVALID, "pano_warp_perspective: bad argument");
  double Minv[9];
  invert33(M, Minv);
  int bh0 = std::min(16, dh);
  int bw0 = std::min(1024 / bh0, dw);
  DevImage S = to_device(c, src, w, h, stride, mem, 0);
  if (mem == PANO_MEM_DEVICE) {
    warp_only_device(c->st, S, Minv, bw0, dst, dw, dh, dstride);
  } else {
    size_t pitch = align_up((size_t)dw * 3, 256);
    c->tmp[0].reserve(pitch * (size_t)dh);
    warp_only_device(c->st, S, Minv, bw0, c->tmp[0].as<uint8_t>(), dw, dh, pitch);
    PANO_CUDA(cudaMemcpy2DAsync(dst, dstride, c->tmp[0].p, pitch, (size_t)dw * 3, dh, cudaMemcpyDeviceToHost, c->st));
  }
  PANO_CUDA(stream_wait(c->st));
  return PANO_OK;
  API_CATCH(c)
}

int pano_stitch_pair(pano_ctx* c, const uint8_t* left, int wl, int hl, size_t stride_l, const uint8_t* right, int wr,
                     int hr, size_t stride_r, int mem, const pano_harris_opts* hopts, const pano_ransac_opts* ropts,
                     pano_pair_result* res) {
  API_TRY(c)
  if (!valid_image(left, wl, hl, stride_l) || !valid_image(right, wr, hr, stride_r) || !hopts || !ropts || !res)
    return fail(c, PANO_ERR_INVALID, "pano_stitch_pair: bad argument");
  if (int e = check_harris(*hopts)) return fail(c, e, "pano_stitch_pair: unsupported option");
  if (int e = check_match_mode(c, *hopts)) return fail(c, e, "pano_stitch_pair: the binary descriptor needs patch size 5");
  if (ropts->num_samples != 4) return fail(c, PANO_ERR_UNSUPPORTED, "only num_samples == 4 is supported");
  DevImage L = to_device(c, left, wl, hl, stride_l, mem, 0);
  DevImage R = to_device(c, right, wr, hr, stride_r, mem, 1);
  return stitch_pair_device(c, L, R, *hopts, *ropts, res);
  API_CATCH(c)
}

int pano_stitch_pair_async(pano_ctx* c, const uint8_t* left, int wl, int hl, size_t stride_l, const uint8_t* right,
                           int wr, int hr, size_t stride_r, int mem, const pano_harris_opts* hopts,
                           const pano_ransac_opts* ropts, void* stream, pano_pair_result* res) {
  API_TRY(c)
  if (c->async_job.valid()) return PANO_ERR_BUSY;
  if (!valid_image(left, wl, hl, stride_l) || !valid_image(right, wr, hr, stride_r) || !hopts || !ropts || !res)
    return fail(c, PANO_ERR_INVALID, "pano_stitch_pair_async: bad argument");
  if (int e = check_harris(*hopts)) return fail(c, e, "pano_stitch_pair_async: unsupported option");
  if (int e = check_match_mode(c, *hopts)) return fail(c, e, "pano_stitch_pair_async: the binary descriptor needs patch size 5");
  if (ropts->num_samples != 4) return fail(c, PANO_ERR_UNSUPPORTED, "only num_samples == 4 is supported");
  if (!c->ev_async) PANO_CUDA(cudaEventCreateWithFlags(&c->ev_async, cudaEventDisableTiming));
  c->async_stream = stream;
  if (stream) {   // the pair starts after what the caller has enqueued on its stream
    PANO_CUDA(cudaEventRecord(c->ev_async, (cudaStream_t)stream));
    PANO_CUDA(cudaStreamWaitEvent(c->st, c->ev_async, 0));
  }
  const pano_harris_opts ho = *hopts;
  const pano_ransac_opts ro = *ropts;
  c->async_job = std::async(std::launch::async, [=]() -> int {
    return pano_stitch_pair(c, left, wl, hl, stride_l, right, wr, hr, stride_r, mem, &ho, &ro, res);
  });
  return PANO_OK;
  API_CATCH(c)
}

static int async_finish(pano_ctx* c) {
  const int s = c->async_job.get();
  if (c->async_stream) {   // later work on the caller's stream is ordered after the pair (its kernels have completed)
    cudaSetDevice(c->device);
    cudaEventRecord(c->ev_async, c->st);
    cudaStreamWaitEvent((cudaStream_t)c->async_stream, c->ev_async, 0);
  }
  return s;
}

int pano_pair_query(pano_ctx* c) {
  if (!c || !c->async_job.valid()) return PANO_ERR_INVALID;
  if (c->async_job.wait_for(std::chrono::seconds(0)) != std::future_status::ready) return PANO_ERR_BUSY;
  return async_finish(c);
}

int pano_pair_wait(pano_ctx* c) {
  if (!c || !c->async_job.valid()) return PANO_ERR_INVALID;
  return async_finish(c);
}

// ---- asynchronous forms of the stage / fold / batch calls ------------------------------------------
// Same mechanism and contract as pano_stitch_pair_async: the blocking call runs on a worker owned by the context
// (its stage synchronisations happen there), its device work is ordered after what the caller has enqueued on
// `stream`, completion and status come from pano_pair_query / pano_pair_wait.  Argument errors of the underlying call
// are therefore reported at completion.

int pano_detect_async(pano_ctx* c, const uint8_t* bgr, int w, int h, size_t stride, int mem, const pano_harris_opts* opts,
                      int32_t* xy_out, int cap, int* count, void* stream) {
  if (!c || !opts) return PANO_ERR_INVALID;
  const pano_harris_opts o = *opts;
  return start_async(c, stream, [=]() -> int { return pano_detect(c, bgr, w, h, stride, mem, &o, xy_out, cap, count); });
}

int pano_match_async(pano_ctx* c, const int32_t* kp_query, int n_query, const int32_t* kp_train, int n_train,
                     const uint8_t* img_query, int wq, int hq, size_t stride_q, const uint8_t* img_train, int wt, int ht,
                     size_t stride_t, int mem, const pano_harris_opts* opts, int offset, pano_dmatch* out, int cap,
                     int* count, void* stream) {
  if (!c || !opts) return PANO_ERR_INVALID;
  const pano_harris_opts o = *opts;
  return start_async(c, stream, [=]() -> int {
    return pano_match(c, kp_query, n_query, kp_train, n_train, img_query, wq, hq, stride_q, img_train, wt, ht, stride_t, mem,
                      &o, offset, out, cap, count);
  });
}

int pano_ransac_async(pano_ctx* c, const int32_t* kp1, int n1, const int32_t* kp2, int n2, const pano_dmatch* matches,
                      int n_matches, int mem, const pano_ransac_opts* opts, double H_out[9], int* best_inliers,
                      int* best_iteration, int32_t* samples_out, int32_t* counts_out, uint8_t* inlier_mask_out,
                      void* stream) {
  if (!c || !opts) return PANO_ERR_INVALID;
  const pano_ransac_opts o = *opts;
  return start_async(c, stream, [=]() -> int {
    return pano_ransac(c, kp1, n1, kp2, n2, matches, n_matches, mem, &o, H_out, best_inliers, best_iteration, samples_out,
                       counts_out, inlier_mask_out);
  });
}

int pano_warp_overlay_async(pano_ctx* c, const uint8_t* left, int wl, int hl, size_t stride_l, const uint8_t* right, int wr,
                            int hr, size_t stride_r, int mem, const double H[9], uint8_t* canvas_out, size_t canvas_stride,
                            size_t canvas_cap_bytes, pano_canvas_info* info, void* stream) {
  if (!c || !H) return PANO_ERR_INVALID;
  std::array<double, 9> h9;
  memcpy(h9.data(), H, sizeof(double) * 9);
  return start_async(c, stream, [=]() -> int {
    return pano_warp_overlay(c, left, wl, hl, stride_l, right, wr, hr, stride_r, mem, h9.data(), canvas_out, canvas_stride,
                             canvas_cap_bytes, info);
  });
}

int pano_stitch_fold_async(pano_ctx* c, const uint8_t* const* images, const int* ws, const int* hs, const size_t* strides,
                           int n, int mem, const pano_harris_opts* hopts, const pano_ransac_opts* ropts,
                           pano_pair_result* results, void* stream) {
  if (!c || !hopts || !ropts) return PANO_ERR_INVALID;
  const pano_harris_opts ho = *hopts;
  const pano_ransac_opts ro = *ropts;
  return start_async(c, stream, [=]() -> int { return pano_stitch_fold(c, images, ws, hs, strides, n, mem, &ho, &ro, results); });
}

int pano_stitch_batch_async(pano_ctx* c, int n, const uint8_t* const* lefts, const uint8_t* const* rights, int wl, int hl,
                            size_t stride_l, int wr, int hr, size_t stride_r, int mem, const pano_harris_opts* hopts,
                            const pano_ransac_opts* ropts, pano_pair_result* results, uint8_t* const* canvases_out,
                            size_t canvas_cap_bytes, float* ms_batch, void* stream) {
  if (!c || !hopts || !ropts) return PANO_ERR_INVALID;
  const pano_harris_opts ho = *hopts;
  const pano_ransac_opts ro = *ropts;
  return start_async(c, stream, [=]() -> int {
    return pano_stitch_batch(c, n, lefts, rights, wl, hl, stride_l, wr, hr, stride_r, mem, &ho, &ro, results, canvases_out,
                             canvas_cap_bytes, ms_batch);
  });
}

int pano_canvas_device(pano_ctx* c, const uint8_t** ptr, size_t* stride, int* w, int* h) {
  if (!c || !c->has_canvas) return PANO_ERR_INVALID;
  if (ptr) *ptr = c->canvas[c->cur].as<uint8_t>();
  if (stride) *stride = c->cstride;
  if (w) *w = c->cw;
  if (h) *h = c->ch;
  return PANO_OK;
}

int pano_get_canvas(pano_ctx* c, uint8_t* out, size_t out_stride, size_t cap_bytes, int mem, int* w, int* h) {
  API_TRY(c)
  if (!c->has_canvas) return fail(c, PANO_ERR_INVALID, "pano_get_canvas: no canvas");
  if (w) *w = c->cw;
  if (h) *h = c->ch;
  if (!out) return PANO_OK;
  if (out_stride < (size_t)c->cw * 3 || cap_bytes < out_stride * (size_t)(c->ch - 1) + (size_t)c->cw * 3)
    return PANO_ERR_CAPACITY;
  const size_t row = (size_t)c->cw * 3;
  if (mem == PANO_MEM_HOST && out_stride == row && c->cstride != row) {   // pack on the device, then one flat copy
    c->tight.reserve(row * (size_t)c->ch + 16);
    pack_rows_device(c->st, c->canvas[c->cur].as<uint8_t>(), c->cstride, row, c->ch, c->tight.as<uint8_t>());
    PANO_CUDA(cudaMemcpyAsync(out, c->tight.p, row * (size_t)c->ch, cudaMemcpyDeviceToHost, c->st));
  } else {
    PANO_CUDA(cudaMemcpy2DAsync(out, out_stride, c->canvas[c->cur].p, c->cstride, row, c->ch,
                                mem == PANO_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->st));
  }
  PANO_CUDA(stream_wait(c->st));
  return PANO_OK;
  API_CATCH(c)
}

int pano_stitch_fold(pano_ctx* c, const uint8_t* const* images, const int* ws, const int* hs, const size_t* strides,
                     int n, int mem, const pano_harris_opts* hopts, const pano_ransac_opts* ropts,
                     pano_pair_result* results) {
  API_TRY(c)
  if (!images || !ws || !hs || !strides || n < 1 || !hopts || !ropts)
    return fail(c, PANO_ERR_INVALID, "pano_stitch_fold: bad argument");
  for (int i = 0; i < n; i++)
    if (!valid_image(images[i], ws[i], hs[i], strides[i])) return fail(c, PANO_ERR_INVALID, "pano_stitch_fold: bad image");
  if (int e = check_harris(*hopts)) return fail(c, e, "pano_stitch_fold: unsupported option");
  if (int e = check_match_mode(c, *hopts)) return fail(c, e, "pano_stitch_fold: the binary descriptor needs patch size 5");
  if (ropts->num_samples != 4) return fail(c, PANO_ERR_UNSUPPORTED, "only num_samples == 4 is supported");
  // panorama = images[0] (ref :400): place it in the current canvas slot
  {
    size_t pitch = align_up((size_t)ws[0] * 3, 256);
    c->canvas[c->cur].reserve(pitch * (size_t)hs[0]);
    PANO_CUDA(cudaMemcpy2DAsync(c->canvas[c->cur].p, pitch, images[0], strides[0], (size_t)ws[0] * 3, hs[0],
                                mem == PANO_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->st));
    c->cw = ws[0];
    c->ch = hs[0];
    c->cstride = pitch;
    c->has_canvas = true;
  }
  int pcur = 0;   // incremental mode: which kpP buffer holds the panorama's keypoints
  if (c->fold_mode == 1 && n > 1) {
    DevImage P0;
    P0.p = c->canvas[c->cur].as<uint8_t>(); P0.w = c->cw; P0.h = c->ch; P0.stride = c->cstride;
    harris_detect_device(c->st, P0, *hopts, c->hs, c->kpP[0], c->pin);
  }
  for (int i = 1; i < n; i++) {
    DevImage L;
    L.p = c->canvas[c->cur].as<uint8_t>();
    L.w = c->cw;
    L.h = c->ch;
    L.stride = c->cstride;
    DevImage R = to_device(c, images[i], ws[i], hs[i], strides[i], mem, 1);
    pano_pair_result r;
    c->left_kp = c->fold_mode == 1 ? &c->kpP[pcur] : nullptr;
    int s = PANO_ERR_CUDA;
    try {
      s = stitch_pair_device(c, L, R, *hopts, *ropts, &r);
    } catch (...) {
      c->left_kp = nullptr;
      throw;
    }
    c->left_kp = nullptr;
    if (results) results[i - 1] = r;
    if (s == PANO_ERR_CUDA) return s;
    // any other failure: the reference logs and keeps the previous panorama (ref :404-407)
    if (s == PANO_OK && c->fold_mode == 1) {
      // carry the keypoints instead of re-detecting on the grown panorama: old ones shifted, new ones through T*H
      const DevKeypoints& oldk = c->kpP[pcur];
      DevKeypoints& newk = c->kpP[1 - pcur];
      const int total = oldk.count + c->kpR.count;
      newk.xy.reserve(sizeof(int32_t) * 2 * (size_t)std::max(total, 1));
      update_pano_keypoints_device(c->st, oldk.xy.as<int32_t>(), oldk.count, r.canvas.left_x, r.canvas.left_y,
                                   c->kpR.xy.as<int32_t>(), c->kpR.count, r.canvas.TH, r.canvas.canvas_w, r.canvas.canvas_h,
                                   newk.xy.as<int32_t>());
      newk.count = total;
      pcur = 1 - pcur;
    }
  }
  PANO_CUDA(stream_wait(c->st));
  return PANO_OK;
  API_CATCH(c)
}

int pano_set_profile(pano_ctx* c, int on) {
  if (!c) return PANO_ERR_INVALID;
  c->prof.on = on != 0;
  c->prof.used = 0;
  for (int i = 0; i < PROF_N; i++) { c->prof.ms[i] = 0; c->prof.n[i] = 0; }
  return PANO_OK;
}

int pano_get_profile(pano_ctx* c, double* ms_out, int* count_out, int cap) {
  if (!c || !ms_out || !count_out) return PANO_ERR_INVALID;
  cudaSetDevice(c->device);
  if (c->st) cudaStreamSynchronize(c->st);
  if (c->rs.side) cudaStreamSynchronize(c->rs.side);
  prof_collect(c->prof);
  for (int i = 0; i < cap && i < PROF_N; i++) { ms_out[i] = c->prof.ms[i]; count_out[i] = c->prof.n[i]; }
  return PROF_N;
}

void* pano_stream(pano_ctx* c) { return c ? (void*)c->st : nullptr; }

int pano_set_stream(pano_ctx* c, void* stream) {
  if (!c) return PANO_ERR_INVALID;
  cudaSetDevice(c->device);
  if (c->st) cudaStreamSynchronize(c->st);
  if (c->owns_stream && c->st) cudaStreamDestroy(c->st);
  if (stream) {
    c->st = (cudaStream_t)stream;
    c->owns_stream = false;
  } else {
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) return PANO_ERR_CUDA;
    c->owns_stream = true;
  }
  return PANO_OK;
}

int pano_pair_homography(pano_ctx* c, const uint8_t* left, int wl, int hl, size_t stride_l, const uint8_t* right,
                         int wr, int hr, size_t stride_r, int mem, const pano_harris_opts* hopts,
                         const pano_ransac_opts* ropts, pano_pair_result* res) {
  API_TRY(c)
  if (!valid_image(left, wl, hl, stride_l) || !valid_image(right, wr, hr, stride_r) || !hopts || !ropts || !res)
    return fail(c, PANO_ERR_INVALID, "pano_pair_homography: bad argument");
  if (int e = check_harris(*hopts)) return fail(c, e, "pano_pair_homography: unsupported option");
  if (int e = check_match_mode(c, *hopts)) return fail(c, e, "pano_pair_homography: the binary descriptor needs patch size 5");
  if (ropts->num_samples != 4) return fail(c, PANO_ERR_UNSUPPORTED, "only num_samples == 4 is supported");
  DevImage L = to_device(c, left, wl, hl, stride_l, mem, 0);
  DevImage R = to_device(c, right, wr, hr, stride_r, mem, 1);
  return stitch_pair_device(c, L, R, *hopts, *ropts, res, true);
  API_CATCH(c)
}

void pano_mul33(const double A[9], const double B[9], double out[9]) { mul33(A, B, out); }

// measurement aid: the plan ransac.cu makes for this match count (default window scale and width), as numbers
int pano_replay_work_estimate(int n_matches, int iterations, double target_candidates, pano_replay_work* out) {
  if (n_matches < 4 || iterations < 1 || !out || target_candidates < 0 || (target_candidates > 0 && target_candidates < 64) ||
      target_candidates > 51000.0)
    return PANO_ERR_INVALID;
  try {
    const ReplayPlan P = plan_replay((uint32_t)n_matches, iterations, 1, target_candidates > 0 ? target_candidates : 50000.0);
    const uint32_t steps = shuffle_steps((uint32_t)n_matches);
    const uint32_t nkb = (steps + 31u) / 32u;
    memset(out, 0, sizeof *out);
    out->chunk_iterations = P.G;
    out->chunks = (iterations + P.G - 1) / P.G;
    out->steps = steps;
    out->candidates_per_chunk = P.n_cand;
    out->diagonals_per_chunk = P.n_diag;
    out->rejections_mean = P.mu;
    out->rejections_sigma = P.sigma;
    out->cells = (double)out->chunks * (double)P.n_diag * (double)nkb * 32.0;
    return PANO_OK;
  } catch (const std::exception&) {
    return PANO_ERR_CUDA;
  }
}

int pano_chain_geometry(int n, const int* ws, const int* hs, const double* Hs, pano_canvas_info* out) {
  if (n < 1 || !ws || !hs || !Hs || !out) return PANO_ERR_INVALID;
  // ref: src/serial/main.cpp:351-369 with every image i >= 1 in the role of "right"
  float minX = 0, minY = 0, maxX = (float)ws[0], maxY = (float)hs[0];
  for (int i = 1; i < n; i++) {
    const float cx[4] = {0.f, (float)ws[i], (float)ws[i], 0.f};
    const float cy[4] = {0.f, 0.f, (float)hs[i], (float)hs[i]};
    for (int k = 0; k < 4; k++) {
      float px, py;
      persp_point(Hs + 9 * (size_t)i, cx[k], cy[k], &px, &py);
      minX = fminf(minX, px); minY = fminf(minY, py);
      maxX = fmaxf(maxX, px); maxY = fmaxf(maxY, py);
    }
  }
  const double T[9] = {1, 0, (double)(-minX), 0, 1, (double)(-minY), 0, 0, 1};
  memcpy(out->TH, T, sizeof T);
  const float fw = maxX - minX, fh = maxY - minY;
  out->canvas_w = (int)ceilf(fw);
  out->canvas_h = (int)ceilf(fh);
  out->left_x = (int)(-minX);
  out->left_y = (int)(-minY);
  if (!(fw == fw) || !(fh == fh) || out->canvas_w <= 0 || out->canvas_h <= 0) return PANO_ERR_ROI;
  if (out->left_x < 0 || out->left_y < 0 || out->left_x + ws[0] > out->canvas_w || out->left_y + hs[0] > out->canvas_h)
    return PANO_ERR_ROI;
  return PANO_OK;
}

int pano_warp_accumulate(pano_ctx* c, const uint8_t* src, int w, int h, size_t stride, int mem, const double M[9],
                         uint8_t* band, int canvas_w, int canvas_h, int y0, int band_h, size_t band_stride) {
  API_TRY(c)
  if (!valid_image(src, w, h, stride) || !M || !band || canvas_w <= 0 || canvas_h <= 0 || y0 < 0 || band_h <= 0 ||
      y0 + band_h > canvas_h || band_stride < (size_t)canvas_w * 3)
    return fail(c, PANO_ERR_INVALID, "pano_warp_accumulate: bad argument");
  DevImage S = to_device(c, src, w, h, stride, mem, 0);
  if (mem == PANO_MEM_DEVICE) {
    warp_accumulate_device(c->st, S, M, band, canvas_w, canvas_h, y0, band_h, band_stride);
  } else {
    size_t pitch = align_up((size_t)canvas_w * 3, 256);
    c->tmp[0].reserve(pitch * (size_t)band_h);
    PANO_CUDA(cudaMemcpy2DAsync(c->tmp[0].p, pitch, band, band_stride, (size_t)canvas_w * 3, band_h,
                                cudaMemcpyHostToDevice, c->st));
    warp_accumulate_device(c->st, S, M, c->tmp[0].as<uint8_t>(), canvas_w, canvas_h, y0, band_h, pitch);
    PANO_CUDA(cudaMemcpy2DAsync(band, band_stride, c->tmp[0].p, pitch, (size_t)canvas_w * 3, band_h,
                                cudaMemcpyDeviceToHost, c->st));
  }
  PANO_CUDA(stream_wait(c->st));
  return PANO_OK;
  API_CATCH(c)
}

int pano_stitch_batch(pano_ctx* c, int n, const uint8_t* const* lefts, const uint8_t* const* rights, int wl, int hl,
                      size_t stride_l, int wr, int hr, size_t stride_r, int mem, const pano_harris_opts* hopts,
                      const pano_ransac_opts* ropts, pano_pair_result* results, uint8_t* const* canvases_out,
                      size_t canvas_cap_bytes, float* ms_batch) {
  API_TRY(c)
  if (n < 0 || !lefts || !rights || !hopts || !ropts || !results)
    return fail(c, PANO_ERR_INVALID, "pano_stitch_batch: bad argument");
  for (int i = 0; i < n; i++)
    if (!valid_image(lefts[i], wl, hl, stride_l) || !valid_image(rights[i], wr, hr, stride_r))
      return fail(c, PANO_ERR_INVALID, "pano_stitch_batch: bad image");
  if (int e = check_harris(*hopts)) return fail(c, e, "pano_stitch_batch: unsupported option");
  if (int e = check_match_mode(c, *hopts)) return fail(c, e, "pano_stitch_batch: the binary descriptor needs patch size 5");
  if (ropts->num_samples != 4) return fail(c, PANO_ERR_UNSUPPORTED, "only num_samples == 4 is supported");
  // Pairs are independent: run them on several lanes (child contexts, each with its own stream,
  // scratch and host thread) so that one pair's host synchronisations, copies and low-occupancy
  // kernels overlap with another pair's work.  PANO_BATCH_LANES (default 10, 1 = sequential); lanes beyond the
  // host cores poll-and-sleep instead of spinning inside the driver (t_yield_wait, set per lane thread).
  int n_lanes = 10;
  if (const char* e = getenv("PANO_BATCH_LANES")) n_lanes = atoi(e);
  if (n_lanes < 1) n_lanes = 1;
  if (n_lanes > 32) n_lanes = 32;
  if (n_lanes > n) n_lanes = n > 0 ? n : 1;
  // lanes beyond the host cores (or PANO_YIELD_WAIT=1): waits poll and sleep instead of spinning in the driver
  bool yield_wait = false;
  {
    const unsigned hc = std::thread::hardware_concurrency();
    const char* e = getenv("PANO_YIELD_WAIT");
    // (one process per GPU: the other ranks of this node have as many lane threads on the same cores)
    const char* lw = getenv("LOCAL_WORLD_SIZE");
    const int ranks = lw && atoi(lw) > 0 ? atoi(lw) : 1;
    yield_wait = e ? atoi(e) != 0 : (hc > 0 && n_lanes * ranks + 1 > (int)hc);
  }
  // one mt19937 output stream for all slots: long enough for shuffles of up to 16 384 matches (longer ones make the
  // slot generate its own)
  if (ropts->num_iterations > 0) {
    const uint64_t steps_cap = 8192;
    mt_ensure(c->st, c->mt, c->seed, (uint64_t)ropts->num_iterations * (steps_cap + steps_cap / 32 + 256) + 65536, steps_cap + 4096);
  }
  cudaEvent_t e0, e1;
  PANO_CUDA(cudaEventCreate(&e0));
  PANO_CUDA(cudaEventCreate(&e1));
  PANO_CUDA(stream_wait(c->st));
  PANO_CUDA(cudaEventRecord(e0, c->st));
  std::vector<int> lane_rc((size_t)n_lanes, PANO_OK);
  std::vector<std::string> lane_err((size_t)n_lanes);   // merged into c->err after the join (no shared writes)
  // A lane is a software pipeline over N_SLOTS child contexts: stage A (detect, match, replay launched on the
  // slot's side stream) of pair k, then stage B (solve, warp, download) of pair k - DEPTH.  With the resident
  // replay (one CTA, ~20x less work than the chunked one but milliseconds long) nobody waits for a replay: it
  // finishes while the lane's next DEPTH pairs go through stage A.  The upload of pair k + 1 is enqueued before
  // stage A of pair k, so N_SLOTS = DEPTH + 2 slots are in use.
  static const int DEPTH = [] { const char* e = getenv("PANO_BATCH_DEPTH"); int v = e ? atoi(e) : 1; return v < 0 ? 0 : (v > 6 ? 6 : v); }();
  const int N_SLOTS = DEPTH + 2;
  const bool host_io = mem != PANO_MEM_DEVICE;
  // the slots of all lanes: one child context (stream + scratch) each; no other streams than these and the two
  // copy streams exist, so up to 30 slots run without hardware work-queue aliasing (CUDA_DEVICE_MAX_CONNECTIONS = 32)
  while ((int)c->slots.size() < n_lanes * N_SLOTS) {
    pano_ctx* sl = nullptr;
    int s = pano_create(c->device, c->seed, &sl);
    if (s != PANO_OK) return fail(c, s, "pano_stitch_batch: cannot create a slot context");
    c->slots.push_back(sl);
  }
  std::mutex down_order;
  if (host_io && !c->st_up) {
    // ONE upload and ONE download stream per context: the copy engines serialise transfers anyway, and every extra
    // stream costs one of the (at most 32) hardware work queues that the slots' compute streams need
    PANO_CUDA(cudaStreamCreateWithFlags(&c->st_up, cudaStreamNonBlocking));
    PANO_CUDA(cudaStreamCreateWithFlags(&c->st_down, cudaStreamNonBlocking));
  }
  auto work = [&](int li) {
    t_yield_wait = yield_wait ? 1 : 0;
    try {
      PANO_CUDA(cudaSetDevice(c->device));
      pano_ctx* const* lane_slots = &c->slots[(size_t)li * N_SLOTS];   // (created below, before the lane threads start)
      for (int q = 0; q < N_SLOTS; q++) {
        pano_ctx* sl = lane_slots[q];
        sl->seed = c->seed;
        sl->matcher = c->matcher;
        sl->match_mode = c->match_mode;
        sl->knn_mode = c->knn_mode;
        sl->rs.shared_mt = &c->mt;
        sl->pack_tight = host_io && canvases_out != nullptr;
        // replay: chunked with small chunks (least speculative work) unless PANO_BATCH_REPLAY=1 asks for the resident
        // one-CTA replay (measured slower in throughput mode: profiles/r02_batch_experiments.md)
        static const int batch_replay = [] { const char* e = getenv("PANO_BATCH_REPLAY"); return e ? atoi(e) : 0; }();
        sl->replay_mode = (DEPTH > 0 && n_lanes > 1) ? batch_replay : c->replay_mode;
        // chunk size of the chunked replay: small chunks = least speculative GPU work (resident inputs: 21.9 k MP/s at
        // 4000 candidates per chunk against 19.6 k at 16000); with host buffers the step is PCIe bound and fewer, larger
        // chunks (a third of the launches) leave the copy engines better fed (13.4 k against 12.9 k MP/s end to end)
        sl->replay_target = n_lanes > 1 ? (host_io ? 16000.0 : 4000.0) : 0.0;
        // the replay is pre-launched on the slot's side stream as soon as the match count is known (measured: 21.5 k
        // against 18.3 k MP/s resident without it, in spite of the second stream per slot)
        sl->overlap_replay = DEPTH > 0;
        if (host_io && !sl->ev_up[0]) {
          PANO_CUDA(cudaEventCreateWithFlags(&sl->ev_up[0], cudaEventDisableTiming));
          PANO_CUDA(cudaEventCreateWithFlags(&sl->ev_down[0], cudaEventDisableTiming));
        }
      }
      const int n_mine = li < n ? (n - li + n_lanes - 1) / n_lanes : 0;      // pairs li, li + n_lanes, ...
      std::vector<DevImage> Ls((size_t)N_SLOTS), Rs((size_t)N_SLOTS);
      std::vector<char> down_pending((size_t)N_SLOTS, 0);
      // both images of pair i -> the slot's upload buffers, on the slot's upload stream
      auto enqueue_upload = [&](int k) {
        const int i = li + k * n_lanes, q = k % N_SLOTS;
        pano_ctx* sl = lane_slots[q];
        const size_t pl = align_up((size_t)wl * 3, 256), pr = align_up((size_t)wr * 3, 256);
        sl->upq[0][0].reserve(pl * hl);
        sl->upq[0][1].reserve(pr * hr);
        {
          // a pair needs both images: keep its two uploads adjacent in the copy engine's queue, otherwise the
          // lanes' copies interleave (all lefts, then all rights) and no lane can start until most have landed
          static std::mutex upload_order;
          std::lock_guard<std::mutex> lk(upload_order);
          copy_image_async(sl->upq[0][0].p, pl, lefts[i], stride_l, (size_t)wl * 3, hl, cudaMemcpyHostToDevice, c->st_up);
          copy_image_async(sl->upq[0][1].p, pr, rights[i], stride_r, (size_t)wr * 3, hr, cudaMemcpyHostToDevice, c->st_up);
          PANO_CUDA(cudaEventRecord(sl->ev_up[0], c->st_up));
        }
        Ls[q].p = sl->upq[0][0].as<uint8_t>(); Ls[q].w = wl; Ls[q].h = hl; Ls[q].stride = pl;
        Rs[q].p = sl->upq[0][1].as<uint8_t>(); Rs[q].w = wr; Rs[q].h = hr; Rs[q].stride = pr;
      };
      auto stage_b = [&](int k) -> bool {
        const int i = li + k * n_lanes, q = k % N_SLOTS;
        pano_ctx* sl = lane_slots[q];
        // this pair overwrites the canvas buffer that the download of the slot's previous pair read
        if (host_io && down_pending[q]) PANO_CUDA(cudaStreamWaitEvent(sl->st, sl->ev_down[0], 0));
        int s = pair_stage_b(sl, *hopts, *ropts, &results[i], false);
        if (s == PANO_ERR_CUDA) { lane_rc[li] = s; lane_err[li] = sl->err; return false; }
        if (s == PANO_OK && canvases_out && canvases_out[i]) {
          const size_t row = (size_t)sl->cw * 3;
          if (row * (size_t)sl->ch > canvas_cap_bytes) {
            results[i].status = PANO_ERR_CAPACITY;
          } else if (host_io) {
            // (the pair's kernels have completed: no dependency to express)
            // one download stream for the whole context: copies have no dependencies (the pair's kernels have
            // completed), the D2H engine takes them in submission order
            std::lock_guard<std::mutex> lk(down_order);
            if (sl->cstride != row)   // packed on the device by stage B: one flat copy
              copy_image_async(canvases_out[i], row, sl->tight.p, row, row, sl->ch, cudaMemcpyDeviceToHost, c->st_down);
            else
              copy_image_async(canvases_out[i], row, sl->canvas[sl->cur].p, sl->cstride, row, sl->ch, cudaMemcpyDeviceToHost,
                               c->st_down);
            PANO_CUDA(cudaEventRecord(sl->ev_down[0], c->st_down));
            down_pending[q] = 1;
          } else {
            copy_image_async(canvases_out[i], row, sl->canvas[sl->cur].p, sl->cstride, row, sl->ch, cudaMemcpyDeviceToDevice,
                             sl->st);
          }
        }
        return true;
      };
      if (host_io && n_mine > 0) enqueue_upload(0);
      bool ok = true;
      for (int k = 0; k < n_mine && ok; k++) {
        const int i = li + k * n_lanes, q = k % N_SLOTS;
        pano_ctx* sl = lane_slots[q];
        if (host_io) {
          if (k + 1 < n_mine) enqueue_upload(k + 1);     // slot (k + 1) % N_SLOTS: its stage B (pair k + 1 - N_SLOTS) is done
          PANO_CUDA(cudaStreamWaitEvent(sl->st, sl->ev_up[0], 0));
        } else {
          Ls[q] = to_device(sl, lefts[i], wl, hl, stride_l, mem, 0);
          Rs[q] = to_device(sl, rights[i], wr, hr, stride_r, mem, 1);
        }
        int s = pair_stage_a(sl, Ls[q], Rs[q], *hopts, *ropts, &results[i]);
        if (s == PANO_ERR_CUDA) { lane_rc[li] = s; lane_err[li] = sl->err; ok = false; break; }
        if (k >= DEPTH) ok = stage_b(k - DEPTH);
      }
      for (int k = std::max(0, n_mine - DEPTH); k < n_mine && ok; k++) ok = stage_b(k);
      for (int q = 0; q < N_SLOTS; q++) {
        pano_ctx* sl = lane_slots[q];
        if (sl->rs.side) PANO_CUDA(stream_wait(sl->rs.side));
        PANO_CUDA(stream_wait(sl->st));
        if (host_io && down_pending[q]) PANO_CUDA(cudaEventSynchronize(sl->ev_down[0]));
      }
    } catch (const CudaError& e) {
      char buf[512];
      snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e.e, cudaGetErrorString(e.e), e.file, e.line,
               e.what);
      lane_err[li] = buf;
      cudaGetLastError();
      lane_rc[li] = PANO_ERR_CUDA;
    } catch (const std::exception& e) {   // (an exception leaving a std::thread would terminate the host process)
      lane_err[li] = std::string("lane failed: ") + e.what();
      lane_rc[li] = PANO_ERR_CUDA;
    }
    t_yield_wait = 0;
  };
  if (n_lanes == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int li = 0; li < n_lanes; li++) th.emplace_back(work, li);
    for (auto& t : th) t.join();
  }
  int rc = PANO_OK;
  for (int li = 0; li < n_lanes; li++)
    if (lane_rc[li] != PANO_OK) { rc = lane_rc[li]; c->err = lane_err[li]; }
  PANO_CUDA(cudaEventRecord(e1, c->st));
  PANO_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_batch) *ms_batch = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
  API_CATCH(c)
}

}  // extern "C"
