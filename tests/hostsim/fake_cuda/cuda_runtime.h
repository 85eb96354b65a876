// fake_cuda/cuda_runtime.h — TEST INFRASTRUCTURE ONLY: the slice of the CUDA runtime API that libpano_b200's host code
// uses, as a synchronous single-"device" CPU model on top of cuda_emu.hpp.  tests/hostsim/build_emu_lib.py compiles the
// engine's UNCHANGED sources (kernel launches rewritten mechanically into emu_launch calls) against this header into
// tests/hostsim/libpano_b200_emu.so, so that the whole engine - host orchestration included - runs in the no-GPU test
// tier.  Never shipped, never loaded by the package.
//   memory     cudaMalloc = aligned_alloc(256) (poisoned), device pointers are host pointers, every copy is a memcpy
//   streams    handles only: all work completes inside the call that enqueues it, so waits and queries succeed at once
//   events     a timestamp at record time
//   launches   run block by block, thread by thread, by cuda_emu.hpp (one launch at a time per process)
//   device     one sm_100 "device" with 148 SMs
#pragma once
#include "../cuda_emu.hpp"

#include <chrono>
#include <cstdio>
#include <mutex>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef struct emu_stream* cudaStream_t;
struct emu_event { std::chrono::steady_clock::time_point t; };
typedef emu_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventBlockingSync = 1, cudaEnableDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; struct { int programmaticStreamSerializationAllowed; } val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream; cudaLaunchAttribute* attrs; unsigned numAttrs; };

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
// PANO_EMU_DEVICES "devices" (default 1) share the host's memory: enough for the multi-device host logic
inline int emu_device_count() { static const int n = [] { const char* e = getenv("PANO_EMU_DEVICES"); int v = e ? atoi(e) : 1; return v < 1 ? 1 : (v > 16 ? 16 : v); }(); return n; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = emu_device_count(); return cudaSuccess; }
inline cudaError_t cudaSetDevice(int d) { return d >= 0 && d < emu_device_count() ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof *p);
  snprintf(p->name, sizeof p->name, "emulated sm_100 device (tests/hostsim)");
  p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
  return cudaSuccess;
}
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 148; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }

// Fault injection (tests only): pano_emu_fail_malloc(n) makes the n-th cudaMalloc / cudaMallocHost from now on fail
// (n = 0: the next one; negative: never), to exercise the engine's error paths under memory exhaustion.
inline long emu_fail_malloc_countdown = -1;
extern "C" __attribute__((weak)) void pano_emu_fail_malloc(long n) { emu_fail_malloc_countdown = n; }
inline bool emu_malloc_should_fail() { return emu_fail_malloc_countdown >= 0 && emu_fail_malloc_countdown-- == 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) {
  if (emu_malloc_should_fail()) { *p = nullptr; return cudaErrorMemoryAllocation; }
  const size_t bytes = (n + 255) / 256 * 256 + 256;
  *p = aligned_alloc(256, bytes);
  if (!*p) return cudaErrorMemoryAllocation;
  memset(*p, 0xA5, bytes);   // device memory is not zeroed
  return cudaSuccess;
}
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t n) {
  if (emu_malloc_should_fail()) { *p = nullptr; return cudaErrorMemoryAllocation; }
  *p = malloc(n ? n : 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
  for (size_t y = 0; y < h; y++) memmove((char*)d + y * dp, (const char*)s + y * sp, w);
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2D(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind k) {
  return cudaMemcpy2DAsync(d, dp, s, sp, w, h, k, nullptr);
}

inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(malloc(8)); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event(); (*e)->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- kernel launches ---------------------------------------------------------------------------------------------
// one launch at a time per process (the emulation's scheduler state is global): contexts driven from several host
// threads (async worker, batch lanes) take turns here
inline std::mutex& emu_launch_mutex() { static std::mutex m; return m; }
inline cudaError_t emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  if (grid.x == 0 || grid.y == 0 || grid.z == 0 || block.x * block.y * block.z == 0 || block.x * block.y * block.z > 1024)
    return cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lk(emu_launch_mutex());
  static const int order = [] { const char* e = getenv("PANO_EMU_BLOCK_ORDER"); return e ? atoi(e) : (int)emu::SHUFFLED; }();
  const char* err = emu::launch(grid, block, body, order, smem);
  if (err) { fprintf(stderr, "[cuda emulation] %s\n", err); return cudaErrorInvalidValue; }
  return cudaSuccess;
}
template <typename... KArgs, typename... Args>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kern)(KArgs...), Args... args) {
  return emu_launch(cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, [&] { kern(args...); });
}

// cudaGetDriverEntryPoint("cuTensorMapEncodeTiled"): the model's encoder (fake cuda.h)
cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long flags, cudaDriverEntryPointQueryResult* q);
