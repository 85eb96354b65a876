// harris_emu.cpp — runs the device code of the detector (csrc/harris_kernels.cuh: the production fused
// response + NMS kernel in both of its instantiations, the two-kernel path for other NMS neighbourhoods, scan,
// ordered scatter, the generic FP64 correlation and the flag compaction) on the CPU emulation of the CUDA execution
// model (cuda_emu.hpp), for the no-GPU test tier.
//
// TEST INFRASTRUCTURE ONLY.  The launch arithmetic below mirrors harris.cu's host launchers; the kernels are the
// product's source, compiled unchanged by g++ (-ffp-contract=off: __dmul_rn / __dadd_rn stay separately rounded).
#include "cuda_emu.hpp"

#include <cmath>
#include <memory>

#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/pano_core.cuh"

namespace pano {
namespace {
#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/harris_kernels.cuh"
}  // namespace
}  // namespace pano

using namespace pano;

namespace {
template <typename T>
struct Aligned {   // cudaMalloc-like alignment (256 bytes)
  T* p = nullptr;
  explicit Aligned(size_t n, int fill = 0) {
    const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) / 256 * 256;
    p = static_cast<T*>(aligned_alloc(256, bytes));
    memset(p, fill, bytes);
  }
  ~Aligned() { free(p); }
  Aligned(const Aligned&) = delete;
};
const char* g_error = nullptr;
void run(dim3 grid, dim3 block, const std::function<void()>& body, int order, size_t dyn = 0) {
  const char* e = emu::launch(grid, block, body, order, dyn);
  if (e) g_error = e;
}
}  // namespace

extern "C" {

const char* hemu_last_error() { return g_error ? g_error : ""; }

void hemu_taps(double* out) {
  const GaussTaps t = make_taps();
  memcpy(out, t.g, sizeof t.g);
}

// harris_detect_device on the emulation.  path: 0 = fused kernel with the plain-load tile (harris_fused_kernel<false>),
// 1 = fused kernel with the TMA tile load (harris_fused_kernel<true>, the TMA unit modelled by cuda_emu.hpp),
// 2 = two-kernel path (harris_response_kernel + nms_mask_kernel; any odd neighbourhood).
// Returns the number of keypoints (row-major order in xy_out), -1 on an emulation error.
int hemu_detect(const uint8_t* img, int w, int h, size_t stride, double k, double thresh, int nbhd, int path, int order,
                int32_t* xy_out, int cap) {
  g_error = nullptr;
  const int mask_stride = (w + 31) / 32;
  Aligned<uint32_t> mask((size_t)mask_stride * h), rowcnt((size_t)h), rowoff((size_t)h), total(1);
  const GaussTaps taps = make_taps();
  if (path < 2) {
    if (nbhd != 3) return -3;
    dim3 grid((w + (FX - 2) - 1) / (FX - 2), (h + (FY - 2) - 1) / (FY - 2)), block(FX, FBY);
    CUtensorMap tmap{img, (uint64_t)w * 3, (uint64_t)h, (uint64_t)stride};
    if (path == 1)
      run(grid, block, [&] { harris_fused_kernel<true>(tmap, img, w, h, stride, k, taps, thresh, mask.p, mask_stride, rowcnt.p); },
          order, sizeof(FusedSmem));
    else
      run(grid, block, [&] { harris_fused_kernel<false>(tmap, img, w, h, stride, k, taps, thresh, mask.p, mask_stride, rowcnt.p); },
          order, sizeof(FusedSmem));
  } else {
    Aligned<double> resp((size_t)w * h);
    run(dim3((w + TX - 1) / TX, (h + TY - 1) / TY), dim3(TX, BY),
        [&] { harris_response_kernel(img, w, h, stride, k, taps, resp.p, thresh, mask.p, mask_stride); }, order);
    dim3 block(32, 8), grid(mask_stride, (h + 7) / 8);
    if (nbhd == 3)
      run(grid, block, [&] { nms_mask_kernel<1>(resp.p, w, h, thresh, 1, mask.p, mask_stride, rowcnt.p); }, order);
    else
      run(grid, block, [&] { nms_mask_kernel<0>(resp.p, w, h, thresh, nbhd / 2, mask.p, mask_stride, rowcnt.p); }, order);
  }
  run(dim3(1), dim3(1024), [&] { scan_kernel(rowcnt.p, rowoff.p, h, total.p); }, 0);
  const int n = (int)total.p[0];
  if (g_error) return -1;
  if (n > 0) {
    Aligned<int32_t> xy((size_t)2 * n);
    const int wpb = 8;
    run(dim3((h + wpb - 1) / wpb), dim3(wpb * 32), [&] { scatter_keypoints_kernel(mask.p, mask_stride, h, rowoff.p, xy.p); }, order);
    if (g_error) return -1;
    memcpy(xy_out, xy.p, sizeof(int32_t) * 2 * (size_t)std::min(n, cap));
  }
  return n;
}

// the response plane of harris_response_kernel (the stage entry point pano_harris_response)
int hemu_response(const uint8_t* img, int w, int h, size_t stride, double k, double* resp_out) {
  g_error = nullptr;
  const GaussTaps taps = make_taps();
  Aligned<double> resp((size_t)w * h);
  run(dim3((w + TX - 1) / TX, (h + TY - 1) / TY), dim3(TX, BY),
      [&] { harris_response_kernel(img, w, h, stride, k, taps, resp.p, 0.0, nullptr, 0); }, 2);
  memcpy(resp_out, resp.p, sizeof(double) * (size_t)w * h);
  return g_error ? -1 : 0;
}

int hemu_convolve(const double* in, int w, int h, const double* kern, int ksize, double* out) {
  g_error = nullptr;
  run(dim3((w + 31) / 32, (h + 7) / 8), dim3(32, 8), [&] { convolve_f64_kernel(in, w, h, kern, ksize, out); }, 1);
  return g_error ? -1 : 0;
}

// compact_flagged: flag_count_kernel + scan_kernel + flag_scatter_kernel; returns the count
int hemu_compact(const uint8_t* flags, int n, int32_t* out_idx) {
  g_error = nullptr;
  int nb = (n + 255) / 256;
  if (nb < 1) nb = 1;
  Aligned<uint32_t> bc((size_t)nb), bo((size_t)nb), cnt(1);
  run(dim3(nb), dim3(256), [&] { flag_count_kernel(flags, n, bc.p); }, 2);
  run(dim3(1), dim3(1024), [&] { scan_kernel(bc.p, bo.p, nb, cnt.p); }, 0);
  run(dim3(nb), dim3(256), [&] { flag_scatter_kernel(flags, n, bo.p, out_idx); }, 1);
  return g_error ? -1 : (int)cnt.p[0];
}

}  // extern "C"
