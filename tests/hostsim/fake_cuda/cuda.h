// fake_cuda/cuda.h — TEST INFRASTRUCTURE ONLY: the driver-API types the engine names (tensor map encoder) and, for
// the translation units that contain the tensor-core matcher, the host model of the machinery its PTX drives.
#pragma once
#include "cuda_runtime.h"
#include "../tcgen05_emu.hpp"

typedef int CUresult;
enum { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1 };
typedef unsigned long long cuuint64_t;
typedef unsigned int cuuint32_t;
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_UINT8 = 0 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_128B = 3 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };

// cuTensorMapEncodeTiled for 2-D byte tensors: the model keeps the description itself (cuda_emu.hpp CUtensorMap)
inline CUresult emu_cuTensorMapEncodeTiled(CUtensorMap* map, CUtensorMapDataType, cuuint32_t rank, void* base, const cuuint64_t* dims,
                                           const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  if (rank != 2 || (reinterpret_cast<uintptr_t>(base) & 15u) || (strides[0] & 15u)) return CUDA_ERROR_INVALID_VALUE;
  map->base = static_cast<const uint8_t*>(base);
  map->row_bytes = dims[0];
  map->rows = dims[1];
  map->pitch = strides[0];
  map->box_rows = box[1];
  return CUDA_SUCCESS;
}
inline cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q) {
  if (strcmp(name, "cuTensorMapEncodeTiled") == 0) {
    *fn = reinterpret_cast<void*>(&emu_cuTensorMapEncodeTiled);
    if (q) *q = cudaDriverEntryPointSuccess;
    return cudaSuccess;
  }
  *fn = nullptr;
  if (q) *q = cudaDriverEntryPointSymbolNotFound;
  return cudaErrorNotSupported;
}
