// probe: which 2-D byte-tensor TMA tile loads are legal (coordinate alignment, box shape, negative / OOB coordinates)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int bytes, uint32_t* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ unsigned long long bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((uint32_t)bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(sm)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(c1), "r"(b) : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b) : "memory");
  }
  uint32_t s = 0;
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) s += sm[i] * (uint32_t)(i + 1);
  atomicAdd(out, s);
}
int main(int argc, char** argv) {
  int c0 = atoi(argv[1]), c1 = atoi(argv[2]), box0 = atoi(argv[3]), box1 = atoi(argv[4]);
  const int W = 1500, H = 300, pitch = 1536;
  std::vector<uint8_t> h((size_t)pitch * H);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 9));
  uint8_t* d; uint32_t* out;
  cudaMalloc(&d, h.size()); cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, str[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1}, es[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("c0=%d c1=%d box=%dx%d: encode failed %d\n", c0, c1, box0, box1, (int)r); return 0; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<<<1, 128, box0 * box1>>>(m, c0, c1, box0 * box1, out);
  cudaError_t e = cudaDeviceSynchronize();
  uint32_t got = 0, want = 0;
  if (e == cudaSuccess) cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost);
  for (int y = 0; y < box1; y++) for (int x = 0; x < box0; x++) {
    int X = c0 + x, Y = c1 + y;
    uint32_t v = (X >= 0 && X < W && Y >= 0 && Y < H) ? h[(size_t)Y * pitch + X] : 0;
    want += v * (uint32_t)(y * box0 + x + 1);
  }
  printf("c0=%d c1=%d box=%dx%d: %s %s\n", c0, c1, box0, box1, cudaGetErrorString(e), e == cudaSuccess ? (got == want ? "DATA OK" : "DATA MISMATCH") : "");
  return 0;
}
