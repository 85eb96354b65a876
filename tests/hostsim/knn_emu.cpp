// knn_emu.cpp — runs the engine's kNN kernels (csrc/knn_kernels.cuh, compiled UNCHANGED by g++) on the CPU
// emulation of the CUDA execution model in cuda_emu.hpp, for the no-GPU test tier.
//
// TEST INFRASTRUCTURE ONLY.  emu_match_knn mirrors the host flow of pano_match_knn / knn.cu (candidate selection,
// kernel launches with the same grid arithmetic, flag compaction, gather); the kernels themselves are the product's
// source.  emu_tc_top2 replays the arithmetic of the tensor-core matcher's top-2 epilogue (match_tc.cu: tile keys,
// four chains, knn_fold_tile, knn_publish) on integer dot products computed here, with the train range cut into
// runs the way CTAs share a query row.
#include "cuda_emu.hpp"

#include "../../ucb-cs267-parallel-panoramic-image-stitching_b200/csrc/knn_kernels.cuh"

#include <climits>
#include <memory>

using namespace pano;

namespace {

template <typename T>
struct Aligned {   // cudaMalloc-like alignment (256 bytes), zero filled
  T* p = nullptr;
  explicit Aligned(size_t n) {
    const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) / 256 * 256;
    p = static_cast<T*>(aligned_alloc(256, bytes));
    memset(p, 0, bytes);
  }
  ~Aligned() { free(p); }
  Aligned(const Aligned&) = delete;
};

struct Side {   // in-border keypoints of one image and their patch descriptors (what build_descriptors_device leaves)
  std::vector<int32_t> orig;
  std::unique_ptr<Aligned<uint8_t>> desc;
  int count = 0;
};

void build_side(const int32_t* xy, int n, const uint8_t* im, int w, int h, size_t stride, int patch, Side& s) {
  const int b = patch / 2;
  for (int i = 0; i < n; i++) {
    const int x = xy[2 * i], y = xy[2 * i + 1];
    if (!(x < b || y < b || x + b >= w || y + b >= h)) s.orig.push_back(i);
  }
  s.count = (int)s.orig.size();
  const size_t rows = ((size_t)s.count + 255) / 256 * 256;
  s.desc.reset(new Aligned<uint8_t>(rows * KNN_DESC_STRIDE));
  for (int k = 0; k < s.count; k++) {
    const int x = xy[2 * s.orig[k]], y = xy[2 * s.orig[k] + 1];
    uint8_t* d = s.desc->p + (size_t)k * KNN_DESC_STRIDE;
    for (int dy = -b; dy <= b; dy++)
      for (int dx = -b; dx <= b; dx++)
        for (int c = 0; c < 3; c++) *d++ = im[(size_t)(y + dy) * stride + 3 * (size_t)(x + dx) + c];
  }
}

const char* g_error = nullptr;
void run(dim3 grid, dim3 block, const std::function<void()>& body, int order) {
  const char* e = emu::launch(grid, block, body, order);
  if (e) g_error = e;
}

}  // namespace

// ---- self-tests of the emulation itself ---------------------------------------------------------------------------
namespace {
// block-wide sum: warp shuffle reduction, one partial per warp in shared memory, second reduction by warp 0
__global__ void selftest_reduce_kernel(const int* in, int n, int* out) {
  __shared__ int part[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int v = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v += in[i];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) part[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (int)(blockDim.x >> 5) ? part[lane] : 0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) out[blockIdx.x] = v;
  }
}
// half of the block never reaches the second barrier while the other half waits for thread 0, which waits for them
__global__ void selftest_deadlock_kernel(int* flag) {
  if (threadIdx.x & 1) {
    __syncthreads();
  } else {
    const int x = __shfl_sync(0xffffffffu, (int)threadIdx.x, 1);   // lanes 1, 3, ... never arrive
    flag[0] = x;
  }
}
// the two halves of a warp meet independently (different masks, different numbers of collectives), then together
__global__ void selftest_half_warps_kernel(int* out) {
  const int lane = threadIdx.x & 31;
  int v = lane;
  if (lane < 16) {
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0x0000ffffu, v, o);      // sum of lanes 0..15 = 120
  } else {
    v = __shfl_sync(0xffff0000u, v, 31);                                          // lane 31's value
  }
  const unsigned b = __ballot_sync(0xffffffffu, v == 120);
  out[threadIdx.x] = v + (int)(b & 0xffffu);
}
}  // namespace

extern "C" {

int emu_selftest_half_warps(int* out64) {
  g_error = nullptr;
  run(dim3(1), dim3(64), [&] { selftest_half_warps_kernel(out64); }, 0);
  return g_error ? 1 : 0;
}

int emu_selftest_reduce(const int* in, int n, int blocks, int threads) {
  g_error = nullptr;
  std::vector<int> out((size_t)blocks, 0);
  run(dim3(blocks), dim3(threads), [&] { selftest_reduce_kernel(in, n, out.data()); }, 2);
  if (g_error) return INT32_MIN;
  int s = 0;
  for (int v : out) s += v;
  return s;
}

int emu_selftest_deadlock() {
  g_error = nullptr;
  int flag = 0;
  run(dim3(1), dim3(64), [&] { selftest_deadlock_kernel(&flag); }, 0);
  return g_error ? 1 : 0;
}

const char* emu_last_error() { return g_error ? g_error : ""; }

// returns the number of matches, or -1 on an emulation error / -2 if the device error word was raised.
// force_splits > 0 overrides the number of train-range splits of the SSD kernel; block_order: emu::Order.
int emu_match_knn(const int32_t* kq, int nq, const int32_t* kt, int nt, const uint8_t* imq, int wq, int hq, size_t sq,
                  const uint8_t* imt, int wt, int ht, size_t st, int patch, int descriptor, double ratio,
                  int force_splits, int block_order, pano_dmatch* out, float* second, int cap) {
  g_error = nullptr;
  Side Q, T;
  build_side(kq, nq, imq, wq, hq, sq, patch, Q);
  build_side(kt, nt, imt, wt, ht, st, patch, T);
  if (Q.count == 0 || T.count == 0) return 0;
  Aligned<unsigned long long> b1((size_t)Q.count), b2((size_t)Q.count);
  memset(b1.p, 0xff, sizeof(unsigned long long) * (size_t)Q.count);
  memset(b2.p, 0xff, sizeof(unsigned long long) * (size_t)Q.count);
  Aligned<int32_t> qorig((size_t)Q.count), torig((size_t)T.count);
  memcpy(qorig.p, Q.orig.data(), sizeof(int32_t) * (size_t)Q.count);
  memcpy(torig.p, T.orig.data(), sizeof(int32_t) * (size_t)T.count);
  double factor;
  if (descriptor == PANO_KNN_BINARY) {
    KnnBinPairs pairs;
    for (int k = 0; k < 32 * KNN_BIN_WORDS; k++) {
      int a, b;
      knn_bin_bit_positions(k, &a, &b);
      pairs.a[k] = (uint8_t)a;
      pairs.b[k] = (uint8_t)b;
    }
    Aligned<uint32_t> qbits((size_t)Q.count * KNN_BIN_WORDS), tbits((size_t)T.count * KNN_BIN_WORDS);
    // device copies of the keypoint lists (aligned)
    Aligned<int32_t> kqd((size_t)2 * nq), ktd((size_t)2 * nt);
    memcpy(kqd.p, kq, sizeof(int32_t) * 2 * (size_t)nq);
    memcpy(ktd.p, kt, sizeof(int32_t) * 2 * (size_t)nt);
    const int wpb = 8;
    run(dim3((Q.count + wpb - 1) / wpb), dim3(wpb * 32),
        [&] { knn_bin_desc_kernel(imq, sq, kqd.p, qorig.p, Q.count, pairs, qbits.p); }, block_order);
    run(dim3((T.count + wpb - 1) / wpb), dim3(wpb * 32),
        [&] { knn_bin_desc_kernel(imt, st, ktd.p, torig.p, T.count, pairs, tbits.p); }, block_order);
    run(dim3((Q.count + wpb - 1) / wpb), dim3(wpb * 32),
        [&] { knn_hamming_kernel(qbits.p, Q.count, tbits.p, T.count, b1.p, b2.p); }, block_order);
    factor = ratio;
  } else {
    const int gx = (Q.count + KQ - 1) / KQ;
    int splits = force_splits > 0 ? force_splits : (148 * 4 + gx - 1) / gx;
    const int max_splits = (T.count + KT_TILE - 1) / KT_TILE;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int per = ((T.count + splits - 1) / splits + KT_TILE - 1) / KT_TILE * KT_TILE;
    splits = (T.count + per - 1) / per;
    run(dim3(gx, splits), dim3(KQ),
        [&] { knn_ssd_simt_kernel(Q.desc->p, Q.count, T.desc->p, T.count, per, b1.p, b2.p); }, block_order);
    factor = ratio * ratio;
  }
  Aligned<pano_dmatch> rec((size_t)Q.count), outd((size_t)Q.count);
  Aligned<float> sec((size_t)Q.count), sec2((size_t)Q.count);
  Aligned<uint8_t> flags((size_t)Q.count);
  Aligned<int> errw(1);
  run(dim3((Q.count + 255) / 256), dim3(256),
      [&] { knn_emit_kernel(b1.p, b2.p, Q.count, qorig.p, torig.p, factor, rec.p, sec.p, flags.p, errw.p); }, block_order);
  Aligned<int32_t> idx((size_t)Q.count);
  int m = 0;
  for (int i = 0; i < Q.count; i++)   // compact_flagged: stable compaction of the flagged rows
    if (flags.p[i]) idx.p[m++] = i;
  if (m > 0)
    run(dim3((m + 255) / 256), dim3(256), [&] { knn_gather_kernel(rec.p, sec.p, idx.p, m, outd.p, sec2.p); }, block_order);
  if (g_error) return -1;
  if (errw.p[0]) return -2;
  for (int i = 0; i < std::min(m, cap); i++) {
    out[i] = outd.p[i];
    second[i] = sec2.p[i];
  }
  return m;
}

// 256-bit descriptors of the in-border keypoints through knn_bin_desc_kernel; returns their number
int emu_binary_descriptors(const int32_t* xy, int n, const uint8_t* im, int w, int h, size_t stride, uint32_t* bits_out,
                           int32_t* orig_out) {
  g_error = nullptr;
  Side S;
  build_side(xy, n, im, w, h, stride, 5, S);
  if (S.count == 0) return 0;
  KnnBinPairs pairs;
  for (int k = 0; k < 32 * KNN_BIN_WORDS; k++) {
    int a, b;
    knn_bin_bit_positions(k, &a, &b);
    pairs.a[k] = (uint8_t)a;
    pairs.b[k] = (uint8_t)b;
  }
  Aligned<uint32_t> bits((size_t)S.count * KNN_BIN_WORDS);
  Aligned<int32_t> orig((size_t)S.count), xyd((size_t)2 * n);
  memcpy(orig.p, S.orig.data(), sizeof(int32_t) * (size_t)S.count);
  memcpy(xyd.p, xy, sizeof(int32_t) * 2 * (size_t)n);
  run(dim3((S.count + 7) / 8), dim3(256), [&] { knn_bin_desc_kernel(im, stride, xyd.p, orig.p, S.count, pairs, bits.p); }, 0);
  if (g_error) return -1;
  memcpy(bits_out, bits.p, sizeof(uint32_t) * KNN_BIN_WORDS * (size_t)S.count);
  memcpy(orig_out, orig.p, sizeof(int32_t) * (size_t)S.count);
  return S.count;
}

// The tensor-core matcher's top-2 epilogue on one set of descriptors (rows of 128 bytes, zero padded): for every
// query row, train tiles of 128 columns are reduced exactly as match_tc_body<true> does - tile key
// (|t|^2 - 2 q.t) * 256 + (j mod 128) in a signed 32-bit word, 0x7fffffff for padding columns, element i of a
// 32-column chunk feeding chain (i mod 4), knn_fold_tile per tile, knn_publish per run - with the row's train tiles
// cut into `runs` contiguous runs (the CTAs that share a query row), published in `order`.
void emu_tc_top2(const uint8_t* qd, int nq, const uint8_t* td, int nt, int runs, int order, unsigned long long* best1,
                 unsigned long long* best2) {
  const int TN = 128;
  const int n_tt = (nt + TN - 1) / TN;
  std::vector<int> tkey((size_t)n_tt * TN);
  for (int j = 0; j < n_tt * TN; j++) {
    uint32_t tn = 0;
    if (j < nt)
      for (int e = 0; e < KNN_DESC_STRIDE; e++) tn += (uint32_t)td[(size_t)j * KNN_DESC_STRIDE + e] * td[(size_t)j * KNN_DESC_STRIDE + e];
    tkey[(size_t)j] = j < nt ? (int)(tn * 256u + (uint32_t)(j & (TN - 1))) : 0x7fffffff;
  }
  for (int q = 0; q < nq; q++) { best1[q] = KNN_NONE; best2[q] = KNN_NONE; }
  if (runs < 1) runs = 1;
  if (runs > n_tt) runs = n_tt;
  std::vector<int> run_ids(runs);
  for (int r = 0; r < runs; r++) run_ids[r] = r;
  if (order == 1) std::reverse(run_ids.begin(), run_ids.end());
  if (order == 2) std::rotate(run_ids.begin(), run_ids.begin() + runs / 2, run_ids.end());
  for (int q = 0; q < nq; q++) {
    uint32_t qn = 0;
    for (int e = 0; e < KNN_DESC_STRIDE; e++) qn += (uint32_t)qd[(size_t)q * KNN_DESC_STRIDE + e] * qd[(size_t)q * KNN_DESC_STRIDE + e];
    for (int r : run_ids) {
      const int t0 = (int)((long long)n_tt * r / runs), t1 = (int)((long long)n_tt * (r + 1) / runs);
      Top2 top = top2_empty();
      for (int tt = t0; tt < t1; tt++) {
        int km[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff}, ks[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
        for (int c = 0; c < TN; c++) {
          const int j = tt * TN + c;
          uint32_t acc = 0;   // the s32 accumulator: q . t_j (zero rows beyond nt)
          if (j < nt)
            for (int e = 0; e < KNN_DESC_STRIDE; e++) acc += (uint32_t)qd[(size_t)q * KNN_DESC_STRIDE + e] * td[(size_t)j * KNN_DESC_STRIDE + e];
          const int key = (int)((uint32_t)tkey[(size_t)j] - 512u * acc);
          top2_insert_i32(km[c & 3], ks[c & 3], key);
        }
        knn_fold_tile(top, km, ks, (int)qn, tt * TN);
      }
      knn_publish(best1, best2, q, top);
    }
  }
}

}  // extern "C"
