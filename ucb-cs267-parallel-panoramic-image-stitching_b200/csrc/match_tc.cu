// match_tc.cu — K4 tensor-core matcher: brute-force 1-NN over 5x5x3 u8 patches as an integer
// distance GEMM on the 5th-gen tensor cores (tcgen05.mma kind::i8, u8 x u8 -> s32 in TMEM)
// with a fused arg-min epilogue.  The Nq x Nt distance matrix is never written.
//
// Semantics: ref src/serial/main.cpp:188-244.  SSD(q, t) = |q|^2 + |t|^2 - 2 q.t, all exact
// integers (q.t <= 75 * 255^2 < 2^23).  For a query row the epilogue minimises the packed key
//      key(j) = (|t_j|^2 - 2 q.t_j) * 256 + (j mod 256)         (fits a signed 32-bit word)
// over the 256 columns of a tile with one IMAD + one signed MIN per element; lexicographic order
// on (partial SSD, column) gives "first strict minimum in train order" (:230-233) inside a tile,
// tiles are visited in train order with a strict '<', and CTAs that split the train range merge
// through atomicMin on (ssd << 32 | j).
//
// Work unit = a SUPER-TILE: QT = 4 query tiles (4 x 128 rows) against one train tile of 128 rows.  The four
// 128 x 128 s32 accumulators fill the CTA's 512 TMEM columns; one B tile (16 KB) brought in by TMA feeds all four,
// so the train descriptors stream from L2 once per 512 query rows.  (Round 1 / early round 2 streamed a 32 KB B tile
// per 128 x 256 tile: 122 MB of L2 reads per 11k x 11k match in 22 us = 5.5 TB/s, the L2 bandwidth - that, not
// tensor-memory read bandwidth, was the bound; tools/microbench.cu measures ~900 B/clk/SM for tcgen05.ld from
// 16 warps.)  The (super-row, train tile) grid is flattened super-row-major and every CTA takes one contiguous,
// equally long run of units (a run that crosses a super-row boundary reloads A and flushes its rows' minima), so all
// SMs finish together.
// Structure (one persistent CTA per SM, 576 threads):
//   warps 0-15 epilogue, four groups of four warps (a warp can only read its own 32-lane quarter of TMEM, so a
//              group of four covers the 128 rows): group g drains accumulator g (query tile g of the super-row):
//              tcgen05.ld 32x32b.x32 -> registers, key/min (1 IMAD + 1/2 VIMNMX3 per element), atomicMin.
//   warp 16    producer: one lane issues TMA loads (cp.async.bulk.tensor.2d, 128B swizzle) of the descriptor
//              tiles straight into the K-major layout the UMMA descriptors expect, and bulk copies
//              (cp.async.bulk) of the train tile's 128 precomputed column constants
//              The group's first lane also ISSUES its accumulator's tcgen05.mma (3 K-steps of 32 bytes) and commits:
//              no cross-role hand-over sits on an accumulator's cycle (issue -> commit -> drain -> named barrier).
//   warp 17    allocates / frees tensor memory
// Pipelines (mbarriers): B stages full/empty (4 deep, TMA complete_tx; released by QT commits, one per group), A
// super-rows full/empty (2 deep), one "full" barrier per accumulator; the four groups interleave on the tensor pipe.
// The column constants live in NCV = NSTAGE + 2 slots: the producer reaches unit m only after the MMAs of unit
// m - NSTAGE have completed, which needed every epilogue of unit m - NSTAGE - 1 to have released its accumulator.
// Every wait is bounded: a bring-up bug raises an error flag instead of hanging the device.
#include "common.cuh"
#include "knn_core.cuh"

#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint

namespace pano {

namespace {

#include "match_tc_kernels.cuh"

std::atomic<int> g_tc_state{0};  // 0 unknown, 1 usable, -1 disabled after a failure

}  // namespace

bool match_tc_available() { return g_tc_state >= 0; }
void match_tc_disable() { g_tc_state = -1; }

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tmap_encoder() {
  static EncodeTiledFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return (EncodeTiledFn)fn;
  }();
  return encode;
}

// descriptor matrix [rows][128 B] u8 -> 2-D tensor map, box = 128 bytes x box_rows, 128B swizzle
void make_tmap(CUtensorMap* map, const void* base, size_t rows, uint32_t box_rows) {
  EncodeTiledFn encode = tmap_encoder();
  if (!encode) throw CudaError{cudaErrorNotSupported, "cuTensorMapEncodeTiled unavailable", __FILE__, __LINE__};
  const cuuint64_t dims[2] = {(cuuint64_t)KB, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)KB};
  const cuuint32_t box[2] = {(cuuint32_t)KB, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw CudaError{cudaErrorInvalidValue, "cuTensorMapEncodeTiled failed", __FILE__, __LINE__};
}
}  // namespace

// plain (unswizzled) 2-D byte tensor map over a pitched image: [rows][row_bytes], box = box_bytes x box_rows.
// Returns false when the layout cannot be described (base or pitch not 16-byte aligned, encoder unavailable).
bool make_tmap_bytes_2d(void* map_out, const void* base, size_t row_bytes, size_t rows, size_t pitch, uint32_t box_bytes,
                        uint32_t box_rows) {
  EncodeTiledFn encode = tmap_encoder();
  if (!encode || (reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (pitch & 15u) != 0 || row_bytes == 0 || rows == 0)
    return false;
  const cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch};
  const cuuint32_t box[2] = {box_bytes, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(reinterpret_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base),
                      dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

void match_tc_device(cudaStream_t st, const DevDescriptors& q, const DevDescriptors& t, unsigned long long* best,
                     DevBuf& keybuf, int* errw, unsigned long long* best2) {
  PANO_CUDA(cudaMemsetAsync(best, 0xff, sizeof(unsigned long long) * (size_t)q.count, st));
  if (best2) PANO_CUDA(cudaMemsetAsync(best2, 0xff, sizeof(unsigned long long) * (size_t)q.count, st));
  if (q.count == 0 || t.count == 0) return;
  const int n_qtiles = (q.count + TM - 1) / TM;
  const int n_ttiles = (t.count + TN - 1) / TN;
  // keybuf: the train side's column constants
  keybuf.reserve(256 + sizeof(int) * (size_t)n_ttiles * TN);
  int* tkey = reinterpret_cast<int*>(keybuf.as<uint8_t>() + 256);
  tkey_kernel<<<(n_ttiles * TN + 255) / 256, 256, 0, st>>>(t.norm.as<uint32_t>(), t.count, n_ttiles * TN, tkey);
  PANO_LAUNCH_CHECK();
  static const int sms = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  // flattened (super-row of QT query tiles, train tile) grid split into equal contiguous runs, one per SM
  const int n_units = ((n_qtiles + QT - 1) / QT) * n_ttiles;
  const int grid = n_units < sms ? n_units : sms;
  const size_t smem = sizeof(Smem) + 1024;
  static const bool attr_set = [&] {
    PANO_CUDA(cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return true;
  }();
  (void)attr_set;
  // descriptor buffers are padded to a multiple of 256 rows (build_descriptors_device)
  CUtensorMap tmap_q, tmap_t;
  make_tmap(&tmap_q, q.desc.p, ((size_t)q.count + 255) / 256 * 256, TM);
  make_tmap(&tmap_t, t.desc.p, ((size_t)t.count + 255) / 256 * 256, TN);
  // A CTA whose pipeline wait gives up ORs PANO_ERRW_TC_ABORT into the context's error word; the host sees it
  // with the next result it waits for (every call, not only the first), fails the call and disables this path.
  if (best2) {
    static const bool attr2_set = [&] {   // (kept apart from the default kernel's set-up: the opt-in variant cannot fail it)
      PANO_CUDA(cudaFuncSetAttribute(match_tc_top2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      return true;
    }();
    (void)attr2_set;
    match_tc_top2_kernel<<<grid, TC_THREADS, smem, st>>>(tmap_q, tmap_t, q.norm.as<uint32_t>(), q.count, tkey, t.count,
                                                         n_qtiles, n_ttiles, n_units, best, best2, errw);
  } else {
    ProfScope ps(PROF_MATCH_TC, st);
    match_tc_kernel<<<grid, TC_THREADS, smem, st>>>(tmap_q, tmap_t, q.norm.as<uint32_t>(), q.count,
                                                    tkey, t.count, n_qtiles, n_ttiles, n_units, best, errw);
  }
  PANO_LAUNCH_CHECK();
  if (g_tc_state == 0) g_tc_state = 1;
}

}  // namespace pano
