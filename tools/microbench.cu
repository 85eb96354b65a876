// microbench.cu — measured pipe ceilings the rooflines in DESIGN.md / bench.py are quoted against
// ("no FP64 or int8 peak is recorded - measure both before quoting K1/K4 pipe fractions").
//
//   tmem   tcgen05.ld-only read bandwidth of tensor memory per SM: 4 / 8 / 16 warps, .x32 / .x64 / .x128,
//          1 .. 4 loads in flight before tcgen05.wait::ld.  The matcher's epilogue has to read a 128 x 256 s32
//          accumulator (128 KB) per 384-cycle MMA tile; this is the number that bounds it.
//   fp64   issue rate of dependent-free DADD / DMUL / DFMA streams (the Harris stencil is 157 non-fusable FP64
//          operations per pixel).
//   ialu   IMAD + add-with-carry rate (the replay cell kernel's 3 instructions per cell).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Run:   tools/microbench [json-out]        (prints one JSON object)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int W>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<64>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
        "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
        "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Each warp reads its own 32-lane quarter (warp % 4) of the CTA's 512 TMEM columns, `reps` times over, with
// `INFLIGHT` loads of width W issued between two waits.  The values are folded so nothing is dead code.
template <int W, int INFLIGHT>
__global__ void __launch_bounds__(512) tmem_ld_kernel(int reps, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t r[INFLIGHT][W];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < reps; it++) {
    // warps sharing a lane quarter start at different columns so they do not read the same words
    uint32_t col = ((uint32_t)(warp >> 2) * 128u + (uint32_t)it * (uint32_t)(W * INFLIGHT)) & 511u;
#pragma unroll
    for (int q = 0; q < INFLIGHT; q++) tmem_ld<W>(base + ((col + q * W) & 511u & ~(uint32_t)(W - 1)), r[q]);
    tmem_wait();
#pragma unroll
    for (int q = 0; q < INFLIGHT; q++)
#pragma unroll
      for (int i = 0; i < W; i += 8) acc ^= r[q][i];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u));
}

// tcgen05.mma issue-to-completion rate: one thread issues `n_mma` MMAs of shape M128 x N x K32B (kind::i8, u8 x u8 -> s32)
// or M128 x N x K16 (kind::f16, bf16 x bf16 -> f32) on zero-filled shared-memory operands (K-major, 128B swizzle, the
// matcher's descriptors), commits, and waits for the commit's mbarrier.  cycles / n_mma is the tensor pipe's time per MMA.
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
template <int KIND, int N>   // KIND 0 = i8, 1 = bf16
__global__ void __launch_bounds__(128) mma_rate_kernel(int n_mma, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
  __shared__ uint32_t s_tmem;
  __shared__ unsigned long long bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint64_t adesc = mk_desc(smem_u32(sm)), bdesc = mk_desc(smem_u32(sm + 128 * 128));
    // instruction descriptor: D format bits [4,6) (1 = f32, 2 = s32), A/B format bits [7,10) / [10,13) (i8: 0 = u8; f16 kind: 1 = bf16)
    const uint32_t idesc = KIND == 0 ? ((2u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24))
                                     : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i++) {
      const uint32_t d = s_tmem + (uint32_t)((i & 1) * 256);
      if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u));
}

template <int KIND, int N>
static void run_mma(int sms, std::string& out, long long* d_cycles) {
  const int n_mma = 2048;
  const size_t smem = (128 + 256) * 128 + 1024;
  CK(cudaFuncSetAttribute(mma_rate_kernel<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mma_rate_kernel<KIND, N><<<sms, 128, smem>>>(64, d_cycles);
  CK(cudaDeviceSynchronize());
  mma_rate_kernel<KIND, N><<<sms, 128, smem>>>(n_mma, d_cycles);
  CK(cudaDeviceSynchronize());
  std::vector<long long> cyc(sms);
  CK(cudaMemcpy(cyc.data(), d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  double mean = 0;
  for (long long c : cyc) mean += (double)c;
  mean /= sms;
  const double k_elems = KIND == 0 ? 32.0 : 16.0;
  char buf[256];
  snprintf(buf, sizeof buf, "%s{\"kind\": \"%s\", \"M\": 128, \"N\": %d, \"K\": %d, \"cycles_per_mma\": %.1f, \"mac_per_clk_per_sm\": %.0f}",
           out.empty() ? "" : ", ", KIND == 0 ? "i8" : "bf16", N, (int)k_elems, mean / n_mma, 128.0 * N * k_elems / (mean / n_mma));
  out += buf;
}

template <int OP>   // 0 DADD, 1 DMUL, 2 DFMA, 3 alternating DMUL + DADD (the stencil's mix)
__global__ void __launch_bounds__(256) fp64_kernel(int reps, double seed, double* sink) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = seed + i + threadIdx.x * 1e-9;
  const double m = 1.0000001, c = 1e-7;
  for (int it = 0; it < reps; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (OP == 0) a[i] = __dadd_rn(a[i], c);
      if (OP == 1) a[i] = __dmul_rn(a[i], m);
      if (OP == 2) a[i] = __fma_rn(a[i], m, c);
      if (OP == 3) a[i] = __dadd_rn(__dmul_rn(a[i], m), c);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i];
  if (s == 42.0) sink[0] = s;
}

__global__ void __launch_bounds__(256) ialu_kernel(int reps, uint32_t seed, uint32_t* sink) {
  uint32_t w[8], x = seed + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = i;
  const uint32_t nr = 0u - (seed | 1u), T = seed >> 3;
  for (int it = 0; it < reps; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {   // the replay cell: IMAD, add.cc, addc
      const uint32_t nlo = (x + i) * nr + 0xffffffffu;
      asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, %0;\n\t}" : "+r"(w[i]) : "r"(T), "r"(nlo));
    }
    x += 8;
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= w[i];
  if (s == 0x9e3779b9u) sink[0] = s;
}

static float time_ms(cudaEvent_t e0, cudaEvent_t e1) {
  float ms = 0;
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms;
}

template <int W, int INF>
static void run_tmem(int warps, int sms, std::string& out, long long* d_cycles, uint32_t* d_sink) {
  const int reps = 4096;
  tmem_ld_kernel<W, INF><<<sms, warps * 32>>>(64, d_cycles, d_sink);   // warm-up
  CK(cudaDeviceSynchronize());
  tmem_ld_kernel<W, INF><<<sms, warps * 32>>>(reps, d_cycles, d_sink);
  CK(cudaDeviceSynchronize());
  std::vector<long long> cyc(sms);
  CK(cudaMemcpy(cyc.data(), d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  double mean = 0;
  for (long long c : cyc) mean += (double)c;
  mean /= sms;
  const double bytes = (double)reps * warps * INF * W * 32 * 4;   // per CTA = per SM
  char buf[256];
  snprintf(buf, sizeof buf, "%s{\"warps\": %d, \"width\": %d, \"in_flight\": %d, \"bytes_per_clk_per_sm\": %.2f}",
           out.empty() ? "" : ", ", warps, W, INF, bytes / mean);
  out += buf;
}

int main(int argc, char** argv) {
  int dev = 0, sms = 0, khz = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  long long* d_cycles;
  uint32_t* d_sink;
  double* d_dsink;
  CK(cudaMalloc(&d_cycles, sizeof(long long) * sms));
  CK(cudaMalloc(&d_sink, 64));
  CK(cudaMalloc(&d_dsink, 64));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));

  std::string tm;
  for (int warps : {4, 8, 16}) {
    run_tmem<32, 1>(warps, sms, tm, d_cycles, d_sink);
    run_tmem<32, 2>(warps, sms, tm, d_cycles, d_sink);
    run_tmem<32, 4>(warps, sms, tm, d_cycles, d_sink);
    run_tmem<64, 1>(warps, sms, tm, d_cycles, d_sink);
    run_tmem<64, 2>(warps, sms, tm, d_cycles, d_sink);
  }

  std::string mm;
  run_mma<0, 256>(sms, mm, d_cycles);
  run_mma<0, 128>(sms, mm, d_cycles);
  run_mma<1, 256>(sms, mm, d_cycles);
  run_mma<1, 128>(sms, mm, d_cycles);

  // FP64 / integer issue rates: 8 resident blocks of 256 threads per SM, 8 independent chains per thread
  const int blocks = sms * 8, reps = 20000;
  double fp[4];
  const char* fpn[4] = {"dadd", "dmul", "dfma", "dmul_dadd_pair"};
  for (int op = 0; op < 4; op++) {
    for (int pass = 0; pass < 2; pass++) {
      CK(cudaEventRecord(e0));
      if (op == 0) fp64_kernel<0><<<blocks, 256>>>(reps, 1.0, d_dsink);
      if (op == 1) fp64_kernel<1><<<blocks, 256>>>(reps, 1.0, d_dsink);
      if (op == 2) fp64_kernel<2><<<blocks, 256>>>(reps, 1.0, d_dsink);
      if (op == 3) fp64_kernel<3><<<blocks, 256>>>(reps, 1.0, d_dsink);
      CK(cudaEventRecord(e1));
      const float ms = time_ms(e0, e1);
      const double instr = (double)blocks * 256 * reps * 8 * (op == 3 ? 2 : 1);
      fp[op] = instr / (ms * 1e-3);
    }
  }
  double ia = 0;
  for (int pass = 0; pass < 2; pass++) {
    CK(cudaEventRecord(e0));
    ialu_kernel<<<blocks, 256>>>(reps, 12345u, d_sink);
    CK(cudaEventRecord(e1));
    const float ms = time_ms(e0, e1);
    ia = (double)blocks * 256 * reps * 8 / (ms * 1e-3);   // cells per second (3 instructions each)
  }

  std::string js = "{\"device_sms\": " + std::to_string(sms) + ", \"clock_khz_attr\": " + std::to_string(khz);
  js += ", \"tmem_ld\": [" + tm + "]";
  js += ", \"tcgen05_mma\": [" + mm + "]";
  char buf[512];
  snprintf(buf, sizeof buf,
           ", \"fp64_instr_per_s\": {\"%s\": %.4e, \"%s\": %.4e, \"%s\": %.4e, \"%s\": %.4e}, \"replay_cells_per_s\": %.4e}",
           fpn[0], fp[0], fpn[1], fp[1], fpn[2], fp[2], fpn[3], fp[3], ia);
  js += buf;
  printf("%s\n", js.c_str());
  if (argc > 1) {
    FILE* f = fopen(argv[1], "w");
    if (f) { fprintf(f, "%s\n", js.c_str()); fclose(f); }
  }
  return 0;
}
