// common.cuh — internal declarations shared by the engine's translation units.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include <chrono>
#include <thread>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/pano_b200.h"
#include "pano_core.cuh"

namespace pano {

struct CudaError {
  cudaError_t e;
  const char* what;
  const char* file;
  int line;
};

#define PANO_CUDA(expr)                                                  \
  do {                                                                   \
    cudaError_t e__ = (expr);                                            \
    if (e__ != cudaSuccess) throw ::pano::CudaError{e__, #expr, __FILE__, __LINE__}; \
  } while (0)

// Host-side wait for a stream.  With more batch lanes than host cores (g_yield_wait > 0) the lanes must not
// spin inside the driver, and must not poll it either (every cudaStreamQuery takes the driver lock that the other
// lanes' kernel launches need): the thread blocks on a blocking-sync event until the GPU interrupts.
// The flag is per host thread (a batch call sets it in its own lane threads), so concurrent batch calls on other
// contexts / GPUs do not flip each other's wait mode.
extern thread_local int t_yield_wait;
inline cudaError_t stream_wait(cudaStream_t st) {
  if (t_yield_wait <= 0) return cudaStreamSynchronize(st);
  struct Ev {
    cudaEvent_t e = nullptr;
    int dev = -1;
    ~Ev() { if (e) cudaEventDestroy(e); }
  };
  thread_local Ev ev;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!ev.e || ev.dev != dev) {
    if (ev.e) cudaEventDestroy(ev.e);
    ev.e = nullptr;
    cudaError_t r = cudaEventCreateWithFlags(&ev.e, cudaEventBlockingSync | cudaEventDisableTiming);
    if (r != cudaSuccess) return r;
    ev.dev = dev;
  }
  cudaError_t r = cudaEventRecord(ev.e, st);
  if (r != cudaSuccess) return r;
  return cudaEventSynchronize(ev.e);
}

// Programmatic dependent launch: the next kernel of a stream is set up while the previous one drains; every kernel
// launched this way starts with pdl_wait() (= all earlier grids complete and visible) before it touches memory.
#ifdef PANO_CUDA_EMU   // (CPU emulation tier, tests/hostsim: kernels run one after another)
inline void pdl_wait() {}
inline void pdl_trigger() {}
#else
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("PANO_PDL"); return !(e && atoi(e) == 0); }();
  return on;
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  PANO_CUDA(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
}

// ---- optional per-kernel timing (pano_set_profile): CUDA events recorded on the launching stream right around
// selected launches, read after the call's final synchronisation.  bench.py's roofline numbers come from here, i.e.
// they are measured live in the run that prints them, not copied from a profiler report.
enum ProfId { PROF_HARRIS = 0, PROF_NMS_COMPACT, PROF_DESC, PROF_MATCH_TC, PROF_EMIT, PROF_REPLAY, PROF_DLT, PROF_SCORE,
              PROF_WARP, PROF_N };
struct Prof {
  bool on = false;
  struct Rec { int id; cudaEvent_t a, b; };
  std::vector<Rec> pool;
  size_t used = 0;
  double ms[PROF_N] = {0};
  int n[PROF_N] = {0};
};
extern thread_local Prof* t_prof;
struct ProfScope {
  Prof::Rec* r = nullptr;
  cudaStream_t st;
  ProfScope(int id, cudaStream_t s) : st(s) {
    Prof* p = t_prof;
    if (!p || !p->on) return;
    if (p->used == p->pool.size()) {
      Prof::Rec nr;
      nr.id = id;
      if (cudaEventCreate(&nr.a) != cudaSuccess || cudaEventCreate(&nr.b) != cudaSuccess) return;
      p->pool.push_back(nr);
    }
    r = &p->pool[p->used++];
    r->id = id;
    cudaEventRecord(r->a, st);
  }
  ~ProfScope() { if (r) cudaEventRecord(r->b, st); }
};
// (call with every stream of the context idle)
inline void prof_collect(Prof& p) {
  for (size_t i = 0; i < p.used; i++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.pool[i].a, p.pool[i].b) == cudaSuccess) { p.ms[p.pool[i].id] += ms; p.n[p.pool[i].id]++; }
  }
  p.used = 0;
}

// every engine kernel launch goes through this: counts it and surfaces launch errors
extern std::atomic<uint64_t> g_kernel_launches;
#define PANO_LAUNCH_CHECK()            \
  do {                                 \
    ++::pano::g_kernel_launches;       \
    PANO_CUDA(cudaGetLastError());     \
  } while (0)

// grow-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) PANO_CUDA(cudaFree(p));
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    PANO_CUDA(cudaMalloc(&p, want));
    cap = want;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
  void* p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) PANO_CUDA(cudaFreeHost(p));
    p = nullptr;
    cap = 0;
    PANO_CUDA(cudaMallocHost(&p, bytes + bytes / 4 + 256));
    cap = bytes + bytes / 4 + 256;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// A BGR8 image resident on the device.
struct DevImage {
  const uint8_t* p = nullptr;
  int w = 0, h = 0;
  size_t stride = 0;
};

// Keypoints of one image, resident on the device, in the reference's row-major order.
struct DevKeypoints {
  DevBuf xy;       // int32 (x, y) pairs
  int count = 0;
};

// Descriptors of the in-border keypoints of one image (patch bytes, zero padded rows).
struct DevDescriptors {
  DevBuf desc;     // [count_pad][PANO_DESC_STRIDE] u8
  DevBuf norm;     // u32 sum of squares per row
  DevBuf orig;     // int32 index into the full keypoint list
  int count = 0;   // in-border keypoints
};

constexpr int PANO_DESC_STRIDE = 128;  // bytes per descriptor row (75 used for a 5x5x3 patch)

// ---- kernels' host launchers (one per .cu) ---------------------------------------------
struct HarrisScratch {
  DevBuf resp, mask, rowcnt, rowoff, total;
};
// Runs gray->Sobel->Gaussian->response (+optional copy-out), NMS mask, ordered compaction.
// Leaves the keypoints in kp (device) and returns their number.
int harris_detect_device(cudaStream_t st, const DevImage& img, const pano_harris_opts& o,
                         HarrisScratch& s, DevKeypoints& kp, PinnedBuf& pin);
// cand_dev (optional): one word per 32-px row segment, bit = response > thresh (consumed by the NMS kernel)
void harris_response_device(cudaStream_t st, const DevImage& img, double k, double* resp_dev, double thresh = 0.0,
                            uint32_t* cand_dev = nullptr, int cand_stride = 0);
void convolve_f64_device(cudaStream_t st, const double* in, int w, int h, const double* kern_dev,
                         int ksize, double* out);

// exclusive scan of n uint32 (single block); writes total to *total_dev if non-null
void exclusive_scan_u32(cudaStream_t st, const uint32_t* in, uint32_t* out, int n, uint32_t* total_dev);

// stable compaction of indices i in [0,n) with flags[i] != 0; out_idx gets the indices,
// *count_dev the count.  tmp must hold 2*ceil(n/256)+2 uint32.
void compact_flagged(cudaStream_t st, const uint8_t* flags, int n, int32_t* out_idx,
                     uint32_t* count_dev, DevBuf& tmp);

struct MatchScratch {
  DevBuf flags, tmp, best, cnt, mflags, midx, mtmp, tc_err;
};
// Device-side error word of a context (bits OR-ed in by kernels, read back with every result the host waits for):
constexpr int PANO_ERRW_TC_ABORT = 1;    // tensor-core matcher: a pipeline wait gave up (CTA aborted)
constexpr int PANO_ERRW_NO_BEST = 2;     // a query row has no minimum although train descriptors exist
constexpr int PANO_ERRW_BAD_INDEX = 4;   // a match refers to a keypoint outside [0, n1) x [0, n2)
int build_descriptors_device(cudaStream_t st, const DevImage& img, const int32_t* xy, int n, int patch,
                             MatchScratch& s, DevDescriptors& d, PinnedBuf& pin);
// best[i] = (ssd << 32 | j) over all train descriptors, lowest j on ties
void match_simt_device(cudaStream_t st, const DevDescriptors& q, const DevDescriptors& t,
                       unsigned long long* best);
// best2 != nullptr: the top-2 variant (pano_match_knn) - nearest neighbour in best, runner-up in best2
void match_tc_device(cudaStream_t st, const DevDescriptors& q, const DevDescriptors& t,
                     unsigned long long* best, DevBuf& keybuf, int* errw, unsigned long long* best2 = nullptr);
bool match_tc_available();
void match_tc_disable();
// 2-D byte tensor map (CUtensorMap, 128 bytes) over a pitched image for TMA tile loads; false if not describable
bool make_tmap_bytes_2d(void* map_out, const void* base, size_t row_bytes, size_t rows, size_t pitch, uint32_t box_bytes,
                        uint32_t box_rows);
// turns best[] into pano_match records (ascending query order), applying maxSSD; returns count
int emit_matches_device(cudaStream_t st, const DevDescriptors& q, const DevDescriptors& t,
                        const unsigned long long* best, double max_ssd, int offset, int patch,
                        MatchScratch& s, pano_dmatch* out_dev, PinnedBuf& pin, int* errw);

// opt-in 2-NN / Lowe-ratio matcher (knn.cu; semantics in knn_core.cuh)
struct KnnScratch {
  DevBuf best2, qbits, tbits, rec, second, out, out2, flags, idx, cnt, tmp;
};
// matches (ascending query order) are left in ks.out (records) / ks.out2 (runner-up distances); returns the count
int match_knn_device(cudaStream_t st, const DevImage& iq, const DevImage& it, const int32_t* kq, const int32_t* kt,
                     const DevDescriptors& q, const DevDescriptors& t, const pano_knn_opts& o, bool use_tc,
                     MatchScratch& ms, DevBuf& best1, KnnScratch& ks, PinnedBuf& pin, int* errw);

struct MtStream;
void update_pano_keypoints_device(cudaStream_t st, const int32_t* old_xy, int n_old, int offx, int offy, const int32_t* new_xy,
                                  int n_new, const double* TH, int cw, int ch, int32_t* out);

struct RansacScratch {
  const MtStream* shared_mt = nullptr;       // read-only mt19937 stream generated by the parent context (batch slots share
                                             // it instead of generating 96 private copies); used when long enough
  DevBuf pts, thr, cand_off, cand_samp, base, samples, Hs, valid, counts, result, mask, plan, pts_bits;
  std::shared_ptr<void> plan_cache;          // host-side replay plans of this context (ransac.cu), keyed by (M, iterations, ...)
  int* errw = nullptr;                       // the context's device error word (see PANO_ERRW_*)
  int n1 = 0, n2 = 0;                        // keypoint counts the match indices are checked against (0 = unknown)
  cudaStream_t side = nullptr;               // a pair's replay runs here while its matches are still being computed
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};
struct RansacResult {
  int errw;    // the context's device error word at the time of the read-back (0 = clean)
  int status;  // PANO_OK / PANO_ERR_*
  double H[9];
  int best_count, best_iter;
};
struct MtStream {
  DevBuf x;           // uint32 outputs of std::mt19937(seed)
  uint64_t len = 0;   // outputs generated so far
  uint64_t guard = 0; // never-rejecting guard words after the stream
  uint32_t seed = 0;
  bool valid = false;
  DevBuf state;       // 624-word engine state to continue from
};
void mt_ensure(cudaStream_t st, MtStream& mt, uint32_t seed, uint64_t need, uint64_t guard);
RansacResult ransac_device(cudaStream_t st, const int32_t* kp1_dev, const int32_t* kp2_dev,
                           const pano_dmatch* matches_dev, int m, const pano_ransac_opts& o,
                           uint32_t seed, MtStream& mt, RansacScratch& s, PinnedBuf& pin,
                           int32_t* samples_out_host, int32_t* counts_out_host,
                           uint8_t* mask_out_host, int window_scale, double replay_target,
                           int replay_mode, int phase = 0);

void warp_overlay_device(cudaStream_t st, const DevImage& left, const DevImage& right,
                         const CanvasGeom& g, uint8_t* canvas, size_t canvas_stride);
void warp_accumulate_device(cudaStream_t st, const DevImage& src, const double* M, uint8_t* band, int canvas_w,
                            int canvas_h, int y0, int band_h, size_t band_stride);
// pitched rows -> tightly packed rows on the device (src base and pitch multiples of 4, dst 4-byte aligned)
void pack_rows_device(cudaStream_t st, const uint8_t* src, size_t pitch, size_t row_bytes, int rows, uint8_t* dst);
void warp_only_device(cudaStream_t st, const DevImage& src, const double* Minv, int bw0, uint8_t* dst,
                      int dw, int dh, size_t dstride);

}  // namespace pano
