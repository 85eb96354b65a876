// harris.cu — K1 Harris response (gray -> Sobel -> products -> 5x5 Gaussian -> response),
// K2 strict NMS into a bit mask, ordered (row-major) compaction into a keypoint list.
//
// Semantics: ref src/serial/main.cpp:119-185 (seqHarrisCornerDetectorDetect) and :96-116
// (convolveSequential).  The response is bit-identical to the reference's FP64 pipeline:
//  - gray is OpenCV's 15-bit fixed point BGR2GRAY;
//  - Sobel sums and the three products are small integers, exact in FP64 in any order;
//  - the 5x5 Gaussian is accumulated in the reference's order (rows outer, columns inner,
//    one accumulator from 0.0) with separately rounded multiply and add (__dmul_rn /
//    __dadd_rn, never FMA);
//  - borders: Sobel output is 0 on the outer ring, Gaussian output 0 on the outer 2 px, and
//    the Gaussian reads the zero Sobel ring.
// Roofline: FP64-pipe bound (157 non-fusable FP64 ops per pixel against 3 B/px of input).
#include "common.cuh"

#include <cuda.h>  // CUtensorMap (type only)

namespace pano {

namespace {

#include "harris_kernels.cuh"

}  // namespace

void exclusive_scan_u32(cudaStream_t st, const uint32_t* in, uint32_t* out, int n, uint32_t* total_dev) {
  scan_kernel<<<1, 1024, 0, st>>>(in, out, n, total_dev);
  PANO_LAUNCH_CHECK();
}

void compact_flagged(cudaStream_t st, const uint8_t* flags, int n, int32_t* out_idx, uint32_t* count_dev,
                     DevBuf& tmp) {
  int nb = (n + 255) / 256;
  if (nb < 1) nb = 1;
  tmp.reserve(sizeof(uint32_t) * (2 * (size_t)nb + 2));
  uint32_t* bc = tmp.as<uint32_t>();
  uint32_t* bo = bc + nb;
  flag_count_kernel<<<nb, 256, 0, st>>>(flags, n, bc);
  PANO_LAUNCH_CHECK();
  exclusive_scan_u32(st, bc, bo, nb, count_dev);
  flag_scatter_kernel<<<nb, 256, 0, st>>>(flags, n, bo, out_idx);
  PANO_LAUNCH_CHECK();
}

void harris_response_device(cudaStream_t st, const DevImage& img, double k, double* resp_dev, double thresh,
                            uint32_t* cand_dev, int cand_stride) {
  static const GaussTaps taps = make_taps();
  dim3 grid((img.w + TX - 1) / TX, (img.h + TY - 1) / TY), block(TX, BY);
  harris_response_kernel<<<grid, block, 0, st>>>(img.p, img.w, img.h, img.stride, k, taps, resp_dev, thresh, cand_dev,
                                                 cand_stride);
  PANO_LAUNCH_CHECK();
}

void convolve_f64_device(cudaStream_t st, const double* in, int w, int h, const double* kern_dev, int ksize,
                         double* out) {
  dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
  convolve_f64_kernel<<<grid, block, 0, st>>>(in, w, h, kern_dev, ksize, out);
  PANO_LAUNCH_CHECK();
}

int harris_detect_device(cudaStream_t st, const DevImage& img, const pano_harris_opts& o, HarrisScratch& s,
                         DevKeypoints& kp, PinnedBuf& pin) {
  const int w = img.w, h = img.h;
  const int mask_stride = (w + 31) / 32;
  s.mask.reserve(sizeof(uint32_t) * (size_t)mask_stride * h);
  s.rowcnt.reserve(sizeof(uint32_t) * (size_t)h);
  s.rowoff.reserve(sizeof(uint32_t) * (size_t)h);
  s.total.reserve(sizeof(uint32_t));
  pin.reserve(64);

  static const bool fused_on = [] { const char* e = getenv("PANO_HARRIS_FUSED"); return !(e && atoi(e) == 0); }();
  PANO_CUDA(cudaMemsetAsync(s.rowcnt.p, 0, sizeof(uint32_t) * (size_t)h, st));
  if (o.nms_neighborhood == 3 && fused_on) {
    // fused response + NMS: no response plane (the reference's setting; other neighbourhoods take the two-kernel path)
    static const GaussTaps taps = make_taps();
    PANO_CUDA(cudaMemsetAsync(s.mask.p, 0, sizeof(uint32_t) * (size_t)mask_stride * h, st));
    dim3 grid((w + (FX - 2) - 1) / (FX - 2), (h + (FY - 2) - 1) / (FY - 2)), block(FX, FBY);
    const size_t smem = sizeof(FusedSmem);
    static const bool attr = [&] {
      PANO_CUDA(cudaFuncSetAttribute(harris_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PANO_CUDA(cudaFuncSetAttribute(harris_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      return true;
    }();
    (void)attr;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    ProfScope ps(PROF_HARRIS, st);
    static const bool tma_on = [] { const char* e = getenv("PANO_HARRIS_TMA"); return !(e && atoi(e) == 0); }();
    if (tma_on && make_tmap_bytes_2d(&tmap, img.p, (size_t)w * 3, (size_t)h, img.stride, FRAW, FGH))
      harris_fused_kernel<true><<<grid, block, smem, st>>>(tmap, img.p, w, h, img.stride, o.k, taps, o.nms_thresh,
                                                            s.mask.as<uint32_t>(), mask_stride, s.rowcnt.as<uint32_t>());
    else
      harris_fused_kernel<false><<<grid, block, smem, st>>>(tmap, img.p, w, h, img.stride, o.k, taps, o.nms_thresh,
                                                             s.mask.as<uint32_t>(), mask_stride, s.rowcnt.as<uint32_t>());
    PANO_LAUNCH_CHECK();
  } else {
    s.resp.reserve(sizeof(double) * (size_t)w * h);
    harris_response_device(st, img, o.k, s.resp.as<double>(), o.nms_thresh, s.mask.as<uint32_t>(), mask_stride);
    dim3 block(32, 8), grid(mask_stride, (h + 7) / 8);
    if (o.nms_neighborhood == 3)
      nms_mask_kernel<1><<<grid, block, 0, st>>>(s.resp.as<double>(), w, h, o.nms_thresh, 1, s.mask.as<uint32_t>(),
                                                 mask_stride, s.rowcnt.as<uint32_t>());
    else
      nms_mask_kernel<0><<<grid, block, 0, st>>>(s.resp.as<double>(), w, h, o.nms_thresh, o.nms_neighborhood / 2,
                                                 s.mask.as<uint32_t>(), mask_stride, s.rowcnt.as<uint32_t>());
    PANO_LAUNCH_CHECK();
  }
  ProfScope* pc = new ProfScope(PROF_NMS_COMPACT, st);
  exclusive_scan_u32(st, s.rowcnt.as<uint32_t>(), s.rowoff.as<uint32_t>(), h, s.total.as<uint32_t>());
  delete pc;
  PANO_CUDA(cudaMemcpyAsync(pin.p, s.total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  PANO_CUDA(stream_wait(st));
  int n = (int)*pin.as<uint32_t>();
  kp.count = n;
  kp.xy.reserve(sizeof(int32_t) * 2 * (size_t)(n > 0 ? n : 1));
  if (n > 0) {
    int wpb = 8;
    scatter_keypoints_kernel<<<(h + wpb - 1) / wpb, wpb * 32, 0, st>>>(s.mask.as<uint32_t>(), mask_stride, h,
                                                                     s.rowoff.as<uint32_t>(), kp.xy.as<int32_t>());
    PANO_LAUNCH_CHECK();
  }
  return n;
}

}  // namespace pano
