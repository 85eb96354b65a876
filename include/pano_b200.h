/* pano_b200.h — C ABI of the B200-native stitching engine (libpano_b200.so).
 *
 * Drop-in boundary for the stitching hot path of
 * Albus-Tan/UCB-CS267-Parallel-Panoramic-Image-Stitching.  Each entry point replaces one
 * of the stage functions the reference's gpu_stitching executable links against; the
 * reference interface it replaces is cited as "ref:" (paths relative to the reference
 * repository).  Plain pointers and sizes only — no OpenCV, torch or CUDA types.
 *
 * Semantics are those of the reference's SERIAL path (src/serial/main.cpp), bit-exact for
 * keypoints, matches, RANSAC samples / inlier counts / inlier sets, with one deliberate
 * deviation: the RANSAC engine std::mt19937 is seeded from the context's seed instead of
 * std::random_device (ref: src/serial/main.cpp:264-265), so results are reproducible.
 *
 * Conventions
 *  - Images are 8-bit BGR, interleaved, row stride in bytes (what cv::imread returns).
 *  - `mem` says where caller buffers live: PANO_MEM_HOST (pageable or pinned host memory;
 *    the call copies in/out) or PANO_MEM_DEVICE (device pointers on the context's device;
 *    no host copies of bulk data).
 *  - Keypoints are int32 (x, y) pairs in the reference's row-major detection order
 *    (cv::KeyPoint::pt of the reference is integral; size/angle are unused on the path).
 *  - Matches mirror the three cv::DMatch fields the path uses.
 *  - Every function returns a pano_status; none throws, none falls back to the CPU.
 *    The reference's "empty result" cases map to distinct codes.
 *  - A context is bound to one device and one stream and is not thread-safe (the reference
 *    is single-threaded); use one context per device per thread.
 */
#ifndef PANO_B200_H
#define PANO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pano_ctx pano_ctx;

typedef enum {
  PANO_OK = 0,
  PANO_ERR_INVALID = 1,          /* bad argument */
  PANO_ERR_CUDA = 2,             /* CUDA runtime error; see pano_last_error */
  PANO_ERR_NO_MATCHES = 3,       /* ref: "Not enough matched corners" (src/serial/main.cpp:321-324) */
  PANO_ERR_TOO_FEW_MATCHES = 4,  /* matches < numSamples: RANSAC loop breaks at once (:268-269) */
  PANO_ERR_NO_HOMOGRAPHY = 5,    /* no iteration produced an inlier: empty Mat (:329-332) */
  PANO_ERR_ROI = 6,              /* left ROI outside canvas: the reference throws at :376 */
  PANO_ERR_CAPACITY = 7,         /* caller buffer too small; required size is reported */
  PANO_ERR_UNSUPPORTED = 8,      /* option outside what the engine implements */
  PANO_ERR_NO_DEVICE = 9,        /* no CUDA device / not an sm_100 device */
  PANO_ERR_BUSY = 10             /* an asynchronous call on this context has not completed yet */
} pano_status;

enum { PANO_MEM_HOST = 0, PANO_MEM_DEVICE = 1 };

/* ref: src/serial/main.cpp:28-34 HarrisCornerOptions (same defaults). */
typedef struct {
  double k;                 /* 0.04 */
  double nms_thresh;        /* 1e6  */
  int nms_neighborhood;     /* 3 (odd) */
  int patch_size;           /* 5 (odd; 1, 3 or 5 supported) */
  double max_ssd_thresh;    /* 1e8  */
} pano_harris_opts;

/* ref: src/serial/main.cpp:36-40 RansacOptions / src/gpu/ransac.cuh:10-14 Options. */
typedef struct {
  int num_iterations;       /* 1000 */
  int num_samples;          /* 4 (only 4 is supported: cv::findHomography's 4-point path) */
  double distance_threshold;/* 3.0 */
} pano_ransac_opts;

/* ref: cv::DMatch(queryIdx, trainIdx, distance) as built at src/serial/main.cpp:237. */
typedef struct {
  int32_t query_idx;
  int32_t train_idx;
  float distance;
} pano_dmatch;

typedef struct {
  int canvas_w, canvas_h;   /* ref: canvasSize, src/serial/main.cpp:369 */
  int left_x, left_y;       /* ref: cv::Rect(-minX, -minY, ...), :376 */
  double TH[9];             /* translation * H, :366-372 */
} pano_canvas_info;

typedef struct {
  int status;               /* pano_status of this pair */
  int n_kp_left, n_kp_right;
  int n_matches;
  int best_inliers;
  int best_iteration;
  double H[9];              /* right -> left homography (row-major), ref: :328 */
  pano_canvas_info canvas;
  float ms_detect, ms_match, ms_ransac, ms_warp, ms_total;  /* device time, CUDA events */
} pano_pair_result;

void pano_default_harris_opts(pano_harris_opts* o);
void pano_default_ransac_opts(pano_ransac_opts* o);

/* Context: device ordinal, RANSAC seed.  Fails (never falls back) without an sm_100 GPU. */
int pano_create(int device, uint32_t seed, pano_ctx** out);
void pano_destroy(pano_ctx* ctx);
int pano_set_seed(pano_ctx* ctx, uint32_t seed);
const char* pano_last_error(const pano_ctx* ctx);
const char* pano_version(void);
/* number of engine kernels launched by this context so far (bench.py's gpu_launches) */
uint64_t pano_kernel_launches(const pano_ctx* ctx);
/* 0 = tensor-core matcher (default), 1 = SIMT cross-check kernel (debug / tests) */
int pano_set_matcher(pano_ctx* ctx, int which);
/* shuffle replay of pano_ransac / pano_stitch_*: 0 = chunked speculative replay over the whole GPU (lowest
 * single-pair latency, default), 1 = resident replay (one CTA walks the iterations in order from exact
 * offsets: ~20x less work, used by pano_stitch_batch's lanes).  Both are bit-exact; results are identical. */
int pano_set_replay_mode(pano_ctx* ctx, int mode);

/* ---- stage entry points -------------------------------------------------------------- */

/* ref: gpuHarrisCornerDetectorDetect(image, k, nmsThresh, nmsNeighborhood)
 *      (src/gpu/harris_detector.cuh:5-9) with the semantics of seqHarrisCornerDetectorDetect
 *      (src/serial/main.cpp:119-185).  Writes min(*count, cap) keypoints; *count is the
 *      full number (PANO_ERR_CAPACITY if cap was too small). */
int pano_detect(pano_ctx* ctx, const uint8_t* bgr, int w, int h, size_t stride, int mem,
                const pano_harris_opts* opts, int32_t* xy_out, int cap, int* count);

/* Harris response plane (FP64, w*h, row-major), for inspection and parity tests.
 * ref: src/serial/main.cpp:123-155. */
int pano_harris_response(pano_ctx* ctx, const uint8_t* bgr, int w, int h, size_t stride, int mem,
                         double k, double* resp_out);

/* ref: convolveCUDA(input64f, output64f, kernel) (src/gpu/convolution.cuh:5) with the
 *      semantics of convolveSequential (src/serial/main.cpp:96-116). */
int pano_convolve_f64(pano_ctx* ctx, const double* in, int w, int h, const double* kernel,
                      int ksize, int mem, double* out);

/* ref: gpuHarrisMatchKeyPoints(kpQuery, kpTrain, imgQuery, imgTrain, patchSize, maxSSD,
 *      offset) (src/gpu/harris_matcher.cuh:5-9) with the semantics of
 *      seqHarrisMatchKeyPoints (src/serial/main.cpp:188-244). */
int pano_match(pano_ctx* ctx, const int32_t* kp_query, int n_query, const int32_t* kp_train,
               int n_train, const uint8_t* img_query, int wq, int hq, size_t stride_q,
               const uint8_t* img_train, int wt, int ht, size_t stride_t, int mem,
               const pano_harris_opts* opts, int offset, pano_dmatch* out, int cap, int* count);

/* ---- 2-nearest-neighbour matching with Lowe's ratio test (north star item (c)); opt-in -----------------
 * NOT the reference's matcher and never used by pano_stitch_*: the reference keeps the single nearest patch
 * under a fixed SSD threshold (pano_match above).  Same candidates (in-border keypoints of both sides), same
 * tie rule extended to the runner-up: neighbours are ordered by (distance, position in the train list).
 *  PANO_KNN_PATCH_SSD  the path's raw BGR patch; distance = exact SSD (what pano_match reports); tensor-core
 *                      distance GEMM with a fused top-2 epilogue (tcgen05), or the SIMT kernel under
 *                      pano_set_matcher(ctx, 1).  Test on L2 distances: ssd1 < ratio^2 * ssd2.
 *  PANO_KNN_BINARY     256-bit intensity-comparison descriptor of the 5 x 5 gray patch (bit k compares the two
 *                      positions of pair number 37 k mod 300 of the 300 position pairs); distance = Hamming
 *                      (XOR + popcount, warp-shuffle reduction).  Test: h1 < ratio * h2.
 * A query whose train side offers fewer than two candidates yields no match. */
enum { PANO_KNN_PATCH_SSD = 0, PANO_KNN_BINARY = 1 };
typedef struct {
  int patch_size;           /* 5 (patch descriptor: 1, 3 or 5; binary descriptor: 5 only) */
  int descriptor;           /* PANO_KNN_PATCH_SSD */
  double ratio;             /* 0.75 (0 < ratio <= 1) */
} pano_knn_opts;
void pano_default_knn_opts(pano_knn_opts* o);
/* Writes min(*count, cap) matches in ascending query order: (query_idx, train_idx of the nearest neighbour,
 * its distance); second_out (may be NULL) receives the runner-up's distance of each written match. */
int pano_match_knn(pano_ctx* ctx, const int32_t* kp_query, int n_query, const int32_t* kp_train,
                   int n_train, const uint8_t* img_query, int wq, int hq, size_t stride_q,
                   const uint8_t* img_train, int wt, int ht, size_t stride_t, int mem,
                   const pano_knn_opts* opts, pano_dmatch* out, float* second_out, int cap, int* count);

/* Matcher of the fused calls (pano_stitch_pair / _fold / _batch, pano_pair_homography and their asynchronous forms).
 * mode 0 (default): the reference's matcher (pano_match) - results are the reference's, bit for bit.
 * mode 1 (opt-in, behaviour changing): pano_match_knn with `ratio` and `descriptor`, patch size taken from the call's
 *         pano_harris_opts - RANSAC then draws from the ratio-tested matches only.  Results differ from the reference's
 *         by design (on its own oilseed sample the reference's outcome depends on the RANSAC seed because 77 % of its
 *         matches are outliers; the ratio test leaves 9 %).  The shuffle replay cannot be started before the match count
 *         is known in this mode, so a pair takes slightly longer. */
int pano_set_match_mode(pano_ctx* ctx, int mode, double ratio, int descriptor);

/* ref: GpuRansacHomographyCalculator::computeHomography(kp1, kp2, matches)
 *      (src/gpu/ransac.cuh:8-36) with the semantics of
 *      SeqRansacHomographyCalculator::computeHomography (src/serial/main.cpp:247-307).
 *      Optional outputs (may be NULL): samples_out[num_iterations*4] match indices drawn per
 *      iteration, counts_out[num_iterations] inlier counts (-1 where findHomography was
 *      empty), inlier_mask_out[n_matches] for the returned H.  Host pointers always. */
int pano_ransac(pano_ctx* ctx, const int32_t* kp1, int n1, const int32_t* kp2, int n2,
                const pano_dmatch* matches, int n_matches, int mem, const pano_ransac_opts* opts,
                double H_out[9], int* best_inliers, int* best_iteration, int32_t* samples_out,
                int32_t* counts_out, uint8_t* inlier_mask_out);

/* Canvas geometry for a homography.  ref: src/serial/main.cpp:335-369, :376. */
int pano_canvas_geometry(int wl, int hl, int wr, int hr, const double H[9], pano_canvas_info* out);

/* ref: cv::warpPerspective + left copy + overlay, src/serial/main.cpp:371-386.
 *      canvas_out: canvas_h rows of canvas_stride bytes (>= 3*canvas_w). */
int pano_warp_overlay(pano_ctx* ctx, const uint8_t* left, int wl, int hl, size_t stride_l,
                      const uint8_t* right, int wr, int hr, size_t stride_r, int mem,
                      const double H[9], uint8_t* canvas_out, size_t canvas_stride,
                      size_t canvas_cap_bytes, pano_canvas_info* info);

/* cv::warpPerspective alone (INTER_LINEAR, BORDER_CONSTANT 0), for parity tests. */
int pano_warp_perspective(pano_ctx* ctx, const uint8_t* src, int w, int h, size_t stride, int mem,
                          const double M[9], uint8_t* dst, int dw, int dh, size_t dstride);

/* ---- fused pipeline ------------------------------------------------------------------ */

/* ref: stitchTwoImages(left, right, harrisOpts, ransacOpts) (src/serial/main.cpp:311-391;
 *      src/gpu/main.cpp:322-426).  Everything stays on the device between stages.  The
 *      canvas is kept in the context; fetch it with pano_get_canvas or feed it to the next
 *      fold step.  res->status carries the reference's failure cases. */
int pano_stitch_pair(pano_ctx* ctx, const uint8_t* left, int wl, int hl, size_t stride_l,
                     const uint8_t* right, int wr, int hr, size_t stride_r, int mem,
                     const pano_harris_opts* hopts, const pano_ransac_opts* ropts,
                     pano_pair_result* res);

/* Asynchronous, stream-ordered variant of pano_stitch_pair (the reference's stage calls block on
 * the default stream, ref: src/gpu/convolution.cu:35-53).  Returns at once; the pair runs on a worker owned by the
 * context.  `stream` (a cudaStream_t on the context's device, may be NULL): the pair's device work starts after
 * everything already enqueued on `stream`, and work enqueued on `stream` after pano_pair_wait / a successful
 * pano_pair_query sees the canvas (pano_canvas_device).  Inputs and `res` must stay valid until completion; the
 * context accepts no other call meanwhile (PANO_ERR_BUSY), except pano_pair_query / pano_pair_wait.
 * pano_pair_query: PANO_ERR_BUSY while running, else the pair's status (as pano_stitch_pair would have returned).
 * pano_pair_wait: blocks the calling host thread until completion and returns that status. */
int pano_stitch_pair_async(pano_ctx* ctx, const uint8_t* left, int wl, int hl, size_t stride_l,
                           const uint8_t* right, int wr, int hr, size_t stride_r, int mem,
                           const pano_harris_opts* hopts, const pano_ransac_opts* ropts,
                           void* stream, pano_pair_result* res);
int pano_pair_query(pano_ctx* ctx);
int pano_pair_wait(pano_ctx* ctx);

/* Asynchronous forms of the stage, fold and batch calls ("each with an explicit cudaStream_t-carrying
 * async variant").  Same contract as pano_stitch_pair_async: the call returns at once; the work runs on the context's
 * worker, ordered after everything already enqueued on `stream` (may be NULL); every pointer argument (inputs,
 * outputs, *count, results) must stay valid until completion; one pending call per context (PANO_ERR_BUSY
 * otherwise); pano_pair_query / pano_pair_wait report completion and return the status the blocking call would have
 * returned (argument errors included). */
int pano_detect_async(pano_ctx* ctx, const uint8_t* bgr, int w, int h, size_t stride, int mem,
                      const pano_harris_opts* opts, int32_t* xy_out, int cap, int* count, void* stream);
int pano_match_async(pano_ctx* ctx, const int32_t* kp_query, int n_query, const int32_t* kp_train,
                     int n_train, const uint8_t* img_query, int wq, int hq, size_t stride_q,
                     const uint8_t* img_train, int wt, int ht, size_t stride_t, int mem,
                     const pano_harris_opts* opts, int offset, pano_dmatch* out, int cap, int* count,
                     void* stream);
int pano_ransac_async(pano_ctx* ctx, const int32_t* kp1, int n1, const int32_t* kp2, int n2,
                      const pano_dmatch* matches, int n_matches, int mem, const pano_ransac_opts* opts,
                      double H_out[9], int* best_inliers, int* best_iteration, int32_t* samples_out,
                      int32_t* counts_out, uint8_t* inlier_mask_out, void* stream);
int pano_warp_overlay_async(pano_ctx* ctx, const uint8_t* left, int wl, int hl, size_t stride_l,
                            const uint8_t* right, int wr, int hr, size_t stride_r, int mem,
                            const double H[9], uint8_t* canvas_out, size_t canvas_stride,
                            size_t canvas_cap_bytes, pano_canvas_info* info, void* stream);
int pano_stitch_fold_async(pano_ctx* ctx, const uint8_t* const* images, const int* ws, const int* hs,
                           const size_t* strides, int n, int mem, const pano_harris_opts* hopts,
                           const pano_ransac_opts* ropts, pano_pair_result* results, void* stream);
int pano_stitch_batch_async(pano_ctx* ctx, int n, const uint8_t* const* lefts, const uint8_t* const* rights,
                            int wl, int hl, size_t stride_l, int wr, int hr, size_t stride_r, int mem,
                            const pano_harris_opts* hopts, const pano_ransac_opts* ropts,
                            pano_pair_result* results, uint8_t* const* canvases_out,
                            size_t canvas_cap_bytes, float* ms_batch, void* stream);

/* Copies the context's current canvas (result of the last successful pair / fold step). */
int pano_get_canvas(pano_ctx* ctx, uint8_t* out, size_t out_stride, size_t cap_bytes, int mem,
                    int* w, int* h);
/* Device pointer/stride of the context's current canvas (valid until the next stitch call). */
int pano_canvas_device(pano_ctx* ctx, const uint8_t** ptr, size_t* stride, int* w, int* h);

/* ref: stitchAllImages(images, ...) (src/serial/main.cpp:395-414): left fold over n images;
 *      a failed step keeps the previous panorama.  results[n-1] receives one record per step
 *      (may be NULL).  Returns PANO_OK if a panorama exists at the end. */
int pano_stitch_fold(pano_ctx* ctx, const uint8_t* const* images, const int* ws, const int* hs,
                     const size_t* strides, int n, int mem, const pano_harris_opts* hopts,
                     const pano_ransac_opts* ropts, pano_pair_result* results);

/* Fold restructuring (opt-in, behaviour changing).  mode 0 (default) = the reference's fold: every
 * step re-detects keypoints on the whole grown panorama (ref: src/serial/main.cpp:400-409).  mode 1 = incremental:
 * only the new image is detected; the panorama's keypoints are carried - the previous list shifted by the left
 * image's offset in the new canvas, followed by the new image's keypoints mapped through T*H
 * (cv::perspectiveTransform arithmetic, rounded to the nearest pixel) - and matched with patches taken from the
 * current panorama.  Results differ from the reference's by design; a failed step keeps panorama and list. */
int pano_set_fold_mode(pano_ctx* ctx, int mode);

/* ---- measurement aid (host only, no GPU work) ----
 * The work the chunked shuffle replay plans for one RANSAC run over n_matches matches: pass 1
 * evaluates `cells` (candidate diagonal, shuffle step) tests of 3 integer-ALU instructions each -
 * the operation count bench.py turns into the replay's floor (cells / measured cell rate).
 * target_candidates: candidate walks per chunk (0 = the single-pair default of 50 000;
 * pano_stitch_batch uses 4000 for resident inputs and 16 000 for host buffers).
 * ref: what is being replayed is the per-iteration std::shuffle of src/serial/main.cpp:264-275. */
typedef struct pano_replay_work {
  int32_t chunk_iterations;      /* G: iterations replayed per chunk */
  int32_t chunks;                /* ceil(iterations / G): sequential phases */
  uint32_t steps;                /* engine draws of one rejection-free shuffle */
  uint32_t candidates_per_chunk; /* speculative start offsets (walks) per chunk */
  uint32_t diagonals_per_chunk;  /* diagonals pass 1 evaluates per chunk */
  uint32_t reserved;
  double rejections_mean;        /* mu: expected rejected draws per shuffle */
  double rejections_sigma;
  double cells;                  /* chunks x diagonals_per_chunk x steps rounded up to 32 */
} pano_replay_work;
int pano_replay_work_estimate(int n_matches, int iterations, double target_candidates, pano_replay_work* out);

/* ---- chain mode (multi-image panoramas whose adjacent pairs are independent work items) ----
 * The reference folds images sequentially and re-detects on the growing panorama, which cannot
 * be sharded.  Chain mode instead estimates H(i <- i+1) for every adjacent pair
 * independently (shardable across GPUs), composes them into the frame of image 0 and renders the
 * canvas, optionally band by band.  For two images it reproduces pano_stitch_pair. */

/* Steps 1-3 of stitchTwoImages only (detect both, match, RANSAC): the homography right -> left.
 * ref: src/serial/main.cpp:316-332.  No canvas is produced. */
int pano_pair_homography(pano_ctx* ctx, const uint8_t* left, int wl, int hl, size_t stride_l,
                         const uint8_t* right, int wr, int hr, size_t stride_r, int mem,
                         const pano_harris_opts* hopts, const pano_ransac_opts* ropts,
                         pano_pair_result* res);

/* out = A * B with OpenCV's small-matrix gemm operation order (what `translation * H` does at
 * ref: src/serial/main.cpp:372); used to compose chain homographies reproducibly. */
void pano_mul33(const double A[9], const double B[9], double out[9]);

/* Canvas of n images given H[i] (image i -> image 0 frame, H[0] = identity): bounds over image
 * 0 and all transformed corners exactly as ref :335-369 does for two images.  out->TH receives
 * the translation T (to be multiplied with each H[i] by pano_mul33). */
int pano_chain_geometry(int n, const int* ws, const int* hs, const double* Hs, pano_canvas_info* out);

/* Warps src by M (cv::warpPerspective, INTER_LINEAR, BORDER_CONSTANT 0) into rows
 * [y0, y0 + band_h) of a canvas of width canvas_w and overlays it with the reference's rule
 * (non-black warped pixels overwrite, ref :380-386) on what `band` already holds. */
int pano_warp_accumulate(pano_ctx* ctx, const uint8_t* src, int w, int h, size_t stride, int mem,
                         const double M[9], uint8_t* band, int canvas_w, int canvas_h, int y0,
                         int band_h, size_t band_stride);

/* Throughput mode: n independent pairs (BASELINE config "batch of 4K pairs"), each exactly
 * pano_stitch_pair.  lefts/rights/canvases_out are arrays of n pointers in `mem`; all pairs
 * share the given image geometry.  canvases_out may be NULL (canvases are then produced and
 * discarded on the device); otherwise canvas i is written tightly packed (stride 3*canvas_w)
 * if it fits canvas_cap_bytes, else results[i].status = PANO_ERR_CAPACITY.
 * *ms_batch = device time of the whole batch (CUDA events on the context's stream). */
int pano_stitch_batch(pano_ctx* ctx, int n, const uint8_t* const* lefts, const uint8_t* const* rights,
                      int wl, int hl, size_t stride_l, int wr, int hr, size_t stride_r, int mem,
                      const pano_harris_opts* hopts, const pano_ransac_opts* ropts,
                      pano_pair_result* results, uint8_t* const* canvases_out, size_t canvas_cap_bytes,
                      float* ms_batch);

/* Per-kernel device timing (measurement aid for bench.py's roofline numbers): when on, CUDA events are recorded
 * right around the engine's main kernel launches.  pano_get_profile waits for the context's streams and returns
 * the accumulated milliseconds and launch counts per kernel class since pano_set_profile(ctx, 1), in this order:
 * 0 harris (fused response + NMS), 1 scan/scatter, 2 descriptor gather, 3 tensor-core matcher, 4 emit,
 * 5 shuffle replay (all its kernels), 6 DLT, 7 scoring, 8 warp + overlay.  Returns the number of classes. */
int pano_set_profile(pano_ctx* ctx, int on);
int pano_get_profile(pano_ctx* ctx, double* ms_out, int* count_out, int cap);

/* The context's CUDA stream (a cudaStream_t), so callers can bracket calls with their own
 * events or order their own work against the engine's. */
void* pano_stream(pano_ctx* ctx);
/* Makes the context enqueue all its work on the caller's stream (a cudaStream_t on the context's
 * device; NULL = back to a private stream).  Calls still return after their results are valid. */
int pano_set_stream(pano_ctx* ctx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PANO_B200_H */
