// chain_host.cpp — TEST INFRASTRUCTURE ONLY: host/chain_multi_gpu.hpp (the multi-GPU chain mode of gpu_stitching,
// one process, several devices) instantiated with HOST memory and linked against the CPU stand-in of the C ABI
// (abi_standin.cpp, on the oracle), so that its control flow - pair sharding over the workers, composition of the
// homographies, canvas geometry, band tiling, a broken chain - is checked without a GPU (tests/test_chain_host.py).
#include <cstdlib>
#include <cstring>

#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/host/chain_multi_gpu.hpp"

namespace {
struct HostMem {
  static void set_device(int) {}
  static void* alloc(size_t n) { return std::malloc(n ? n : 1); }
  static void free(void* p) { std::free(p); }
  static bool zero(void* p, size_t n) { std::memset(p, 0, n); return true; }
  static void sync() {}
  static bool copy2d(void* d, size_t dp, const void* s, size_t sp, size_t row_bytes, int rows) {
    for (int y = 0; y < rows; y++) std::memcpy((char*)d + (size_t)y * dp, (const char*)s + (size_t)y * sp, row_bytes);
    return true;
  }
  static bool h2d_2d(void* d, size_t dp, const void* s, size_t sp, size_t rb, int rows) { return copy2d(d, dp, s, sp, rb, rows); }
  static bool d2h_2d(void* d, size_t dp, const void* s, size_t sp, size_t rb, int rows) { return copy2d(d, dp, s, sp, rb, rows); }
};
}  // namespace

extern "C" {
// images: n pointers to tightly packed BGR8 images.  geom_out: w, h, n_used; pair_status / pair_device: n - 1 entries.
// Returns the status of stitch_chain_multi_gpu; the canvas is copied if it fits cap bytes.
int hs_chain_multi(const uint8_t* const* images, const int* ws, const int* hs, int n, int n_dev, uint32_t seed,
                   uint8_t* canvas, size_t cap, int* geom_out, int* pair_status, int* pair_device) {
  std::vector<pano_host::ImageView> views;
  for (int i = 0; i < n; i++) views.push_back({images[i], ws[i], hs[i], (size_t)ws[i] * 3});
  std::vector<int> devices;
  for (int d = 0; d < n_dev; d++) devices.push_back(d);
  pano_harris_opts ho;
  pano_default_harris_opts(&ho);
  pano_ransac_opts ro;
  pano_default_ransac_opts(&ro);
  pano_host::ChainOutput out;
  const int st = pano_host::stitch_chain_multi_gpu<HostMem>(views, devices, seed, ho, ro, &out);
  geom_out[0] = out.w; geom_out[1] = out.h; geom_out[2] = out.n_used;
  for (size_t i = 0; i < out.pairs.size(); i++) { pair_status[i] = out.pairs[i].status; pair_device[i] = out.pair_device[i]; }
  if (st == PANO_OK && out.canvas.size() <= cap) std::memcpy(canvas, out.canvas.data(), out.canvas.size());
  return st;
}
}
