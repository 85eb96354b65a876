// match_tc_emu.cpp — runs the tensor-core matcher's kernel body (csrc/match_tc_kernels.cuh: both instantiations, the
// reference's arg-min matcher and the top-2 variant of pano_match_knn) on the CPU emulation of the CUDA execution model
// (cuda_emu.hpp) plus a host model of the Blackwell machinery it drives (tcgen05_emu.hpp).
//
// TEST INFRASTRUCTURE ONLY.  The launch arithmetic mirrors match_tc_device (tkey_kernel, one persistent CTA per "SM",
// equal contiguous runs of the flattened (super-row, train tile) grid, dynamic shared memory = sizeof(Smem) + 1024).
#include "cuda_emu.hpp"
#include "tcgen05_emu.hpp"

#include <memory>

#include "../../include/pano_b200.h"
#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/knn_core.cuh"

namespace pano {
constexpr int PANO_DESC_STRIDE = 128;   // as in common.cuh
constexpr int PANO_ERRW_TC_ABORT = 1;
namespace {
#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/match_tc_kernels.cuh"
static_assert(TM == 128 && TN == 128 && KB == 128, "tcgen05_emu.hpp models 128 x 128 x 32 MMAs on 128-byte rows");
}  // namespace
}  // namespace pano

using namespace pano;

namespace {
template <typename T>
struct Aligned {
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
}  // namespace

extern "C" {

const char* tcemu_last_error() { return g_error ? g_error : ""; }

// qd / td: nq / nt descriptor rows of 128 bytes (75 used).  ctas: number of persistent CTAs (the product uses the SM
// count; small values make a CTA's run cross super-row boundaries, large ones make several CTAs share a query row).
// top2 = 0: match_tc_kernel, best2 untouched; 1: match_tc_top2_kernel.  Returns 0, -1 on an emulation error (deadlock),
// -2 if the kernel raised its abort word.
int tcemu_match(const uint8_t* qd, int nq, const uint8_t* td, int nt, int ctas, int top2, int block_order,
                unsigned long long* best1, unsigned long long* best2) {
  g_error = nullptr;
  const size_t qrows = ((size_t)nq + 255) / 256 * 256, trows = ((size_t)nt + 255) / 256 * 256;
  Aligned<uint8_t> Q(qrows * 128), T(trows * 128);      // zero padded like build_descriptors_device leaves them
  memcpy(Q.p, qd, (size_t)nq * 128);
  memcpy(T.p, td, (size_t)nt * 128);
  Aligned<uint32_t> qn(qrows), tn(trows);
  for (int i = 0; i < nq; i++) for (int e = 0; e < 128; e++) qn.p[i] += (uint32_t)Q.p[(size_t)i * 128 + e] * Q.p[(size_t)i * 128 + e];
  for (int i = 0; i < nt; i++) for (int e = 0; e < 128; e++) tn.p[i] += (uint32_t)T.p[(size_t)i * 128 + e] * T.p[(size_t)i * 128 + e];
  Aligned<unsigned long long> b1((size_t)nq, 0xff), b2((size_t)nq, 0xff);
  Aligned<int> err(1);
  const int n_qtiles = (nq + TM - 1) / TM, n_ttiles = (nt + TN - 1) / TN;
  Aligned<int> tkey((size_t)n_ttiles * TN);
  const char* e = emu::launch(dim3((n_ttiles * TN + 255) / 256), dim3(256), [&] { tkey_kernel(tn.p, nt, n_ttiles * TN, tkey.p); });
  if (e) { g_error = e; return -1; }
  const int n_units = ((n_qtiles + QT - 1) / QT) * n_ttiles;
  const int grid = n_units < ctas ? n_units : ctas;
  CUtensorMap tmap_q{Q.p, 128, qrows, 128, (uint32_t)TM}, tmap_t{T.p, 128, trows, 128, (uint32_t)TN};
  const size_t smem = sizeof(Smem) + 1024;
  if (top2)
    e = emu::launch(dim3(grid), dim3(TC_THREADS), [&] {
      match_tc_top2_kernel(tmap_q, tmap_t, qn.p, nq, tkey.p, nt, n_qtiles, n_ttiles, n_units, b1.p, b2.p, err.p);
    }, block_order, smem);
  else
    e = emu::launch(dim3(grid), dim3(TC_THREADS), [&] {
      match_tc_kernel(tmap_q, tmap_t, qn.p, nq, tkey.p, nt, n_qtiles, n_ttiles, n_units, b1.p, err.p);
    }, block_order, smem);
  if (e) { g_error = e; return -1; }
  memcpy(best1, b1.p, sizeof(unsigned long long) * (size_t)nq);
  if (top2) memcpy(best2, b2.p, sizeof(unsigned long long) * (size_t)nq);
  return err.p[0] ? -2 : 0;
}

}  // extern "C"
