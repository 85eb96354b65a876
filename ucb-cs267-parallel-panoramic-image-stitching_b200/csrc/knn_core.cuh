// knn_core.cuh — exact arithmetic of the opt-in 2-NN / Lowe-ratio matcher (pano_match_knn), shared by its CUDA
// kernels (knn_kernels.cuh, match_tc.cu) and by the host tier that runs those kernels on a CPU (tests/hostsim).
//
// NOT part of the reference's path: the reference's matcher is the 1-NN SSD with a threshold
// (ref: src/serial/main.cpp:188-244, pano_match).  This is the north star's item (c) - "brute-force kNN matching
// with Lowe's ratio test", a fused top-2 epilogue for the patch descriptors and an XOR / popcount path for binary
// descriptors - offered beside it, off by default and never used by pano_stitch_*.
//
// Definitions (restated by the checker, oracle/pano_oracle.cpp orc_match_knn):
//  * candidates: the in-border keypoints of both sides, exactly as in pano_match;
//  * a neighbour of query i is identified with the 64-bit key  (distance << 32) | j  (j = position of the train
//    keypoint among the in-border train keypoints).  Keys of one query are distinct, their order is "smaller
//    distance first, earlier train keypoint on ties" - the reference's "first strict minimum" rule extended to the
//    runner-up.  nearest = smallest key, runner-up = second smallest key;
//  * patch descriptor: distance = SSD over patch x patch x 3 bytes (exact integer, the reference's distance);
//    Lowe's test on L2 distances  sqrt(ssd1) < ratio * sqrt(ssd2)  is evaluated as
//    (double)ssd1 < (ratio * ratio) * (double)ssd2  (one rounded product; ssd < 2^23 is exact in a double);
//  * binary descriptor (5 x 5 patches only): 256 bits, bit k = gray[a_k] < gray[b_k] for the gray values of the
//    patch (the path's gray formula, gray_u8) and the position pair number (37 k mod 300) of the lexicographic list
//    of the 300 pairs a < b of the 25 patch positions (row-major); distance = Hamming distance (XOR + popcount);
//    Lowe's test  (double)h1 < ratio * (double)h2;
//  * a query with fewer than two candidates on the train side has no runner-up and yields no match.
#pragma once
#include <cstdint>

#include "pano_core.cuh"

namespace pano {

constexpr unsigned long long KNN_NONE = ~0ull;   // "no neighbour yet" (larger than every real key)
constexpr int KNN_BIN_WORDS = 8;                 // 256-bit binary descriptor
constexpr int KNN_BIN_PAIRS = 300;               // unordered pairs of the 25 positions of a 5 x 5 patch

// running two smallest keys of a set of distinct keys
struct Top2 {
  unsigned long long k1, k2;
};
PANO_HD Top2 top2_empty() { Top2 t; t.k1 = KNN_NONE; t.k2 = KNN_NONE; return t; }
PANO_HD void top2_insert(Top2& t, unsigned long long key) {
  const unsigned long long hi = key > t.k1 ? key : t.k1;   // the larger of (key, current nearest)
  t.k1 = key < t.k1 ? key : t.k1;
  t.k2 = hi < t.k2 ? hi : t.k2;
}
PANO_HD void top2_merge(Top2& t, const Top2& o) {
  top2_insert(t, o.k1);
  if (o.k2 != KNN_NONE) top2_insert(t, o.k2);
}
PANO_HD unsigned long long knn_key(uint32_t dist, uint32_t j) { return ((unsigned long long)dist << 32) | j; }

// the same for the signed 32-bit keys of one tensor-core tile (match_tc.cu: key = partial SSD * 256 + column;
// 0x7fffffff = padding column, may repeat)
PANO_HD void top2_insert_i32(int& a, int& b, int key) {
  const int hi = key > a ? key : a;
  a = key < a ? key : a;
  b = hi < b ? hi : b;
}

// One tile of the tensor-core matcher's top-2 epilogue: the four chains' (smallest, runner-up) tile keys -> the
// tile's two smallest -> the row's running pair.  A tile key is (|t_j|^2 - 2 q.t_j) * 256 + (j mod 128); adding the
// query's |q|^2 to its upper part gives the SSD, tile_base + low byte the train index.
PANO_HD void knn_fold_tile(Top2& top, const int (&km)[4], const int (&ks)[4], int qnorm, int tile_base) {
  int a = km[0], b = ks[0];
  for (int c = 1; c < 4; c++) { top2_insert_i32(a, b, km[c]); top2_insert_i32(a, b, ks[c]); }
  if (a != 0x7fffffff) top2_insert(top, knn_key((uint32_t)((a >> 8) + qnorm), (uint32_t)(tile_base + (a & 255))));
  if (b != 0x7fffffff) top2_insert(top, knn_key((uint32_t)((b >> 8) + qnorm), (uint32_t)(tile_base + (b & 255))));
}

// Lowe's ratio test.  ssd: r2 = ratio * ratio (computed once on the host in double); hamming: r = ratio.
PANO_HD bool lowe_accept(uint32_t d1, uint32_t d2, double factor) {
  return (double)d1 < PANO_DMUL(factor, (double)d2);
}

// pair number p (0 <= p < 300) of the lexicographic list of position pairs a < b, 0 <= a, b < 25
PANO_HD void knn_bin_pair(int p, int* a, int* b) {
  int aa = 0;
  while (p >= 24 - aa) { p -= 24 - aa; aa++; }
  *a = aa;
  *b = aa + 1 + p;
}
// positions compared by bit k of the binary descriptor
PANO_HD void knn_bin_bit_positions(int k, int* a, int* b) { knn_bin_pair((37 * k) % KNN_BIN_PAIRS, a, b); }

#if defined(__CUDACC__) || defined(PANO_CUDA_EMU)
// Publishes one thread's two smallest keys of query row q into the row's two global slots.  Every key other than
// the row's final nearest one reaches slot 2: a key that loses (or later loses) slot 1 is handed to slot 2 by the
// thread whose atomicMin received it as the old value, so slot 2 ends as the smallest of all the others = the
// runner-up, whatever the interleaving of the CTAs that share the row.
__device__ __forceinline__ void knn_publish(unsigned long long* best1, unsigned long long* best2, int q, const Top2& t) {
  if (t.k1 != KNN_NONE) {
    const unsigned long long old = atomicMin(&best1[q], t.k1);
    const unsigned long long loser = old > t.k1 ? old : t.k1;
    if (loser != KNN_NONE) atomicMin(&best2[q], loser);
  }
  if (t.k2 != KNN_NONE) atomicMin(&best2[q], t.k2);
}
#endif

}  // namespace pano
