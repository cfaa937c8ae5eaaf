// Host-side TMA descriptor for the bf16 operand matrices of the full-sort kernels (shared by fullsort.cu / fullsort_bwd.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace bdlru {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda).
inline EncodeTiledFn tma_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !ptr) {
    cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// [rows, D] bf16 row-major -> boxes of box_rows rows x 64 channels, 128-byte swizzle, zero fill out of bounds.
inline int make_rows_map(CUtensorMap* m, const void* base, long rows, int D, int box_rows) {
  EncodeTiledFn fn = tma_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return BDLRU_ERR_CUDA;
  }
  // The driver entry point needs the primary context bound on THIS thread; PyTorch's autograd threads only bind it
  // lazily through runtime calls (CUDA_ERROR_INVALID_CONTEXT otherwise).  cudaFree(nullptr) binds it and is a no-op.
  static thread_local bool bound = false;
  if (!bound) {
    cudaFree(nullptr);
    bound = true;
  }
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%ld D=%d)", (int)r, rows, D);
    return BDLRU_ERR_CUDA;
  }
  return BDLRU_OK;
}

// [rows, 16] bf16 row-major (32-byte rows) -> boxes of box_rows rows, 32-byte swizzle, zero fill out of bounds.
inline int make_rows16_map(CUtensorMap* m, const void* base, long rows, int box_rows) {
  EncodeTiledFn fn = tma_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return BDLRU_ERR_CUDA;
  }
  cuuint64_t dims[2] = {16, (cuuint64_t)rows};
  cuuint64_t strides[1] = {32};
  cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (32-byte rows) failed with CUresult %d (rows=%ld)", (int)r, rows);
    return BDLRU_ERR_CUDA;
  }
  return BDLRU_OK;
}

// [rows, D] fp32 row-major -> boxes of box_rows rows x 32 columns (128 bytes), 128-byte swizzle: the target of TMA stores.
inline int make_rows_map_f32(CUtensorMap* m, const void* base, long rows, int D, int box_rows) {
  EncodeTiledFn fn = tma_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return BDLRU_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)D * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32) failed with CUresult %d (rows=%ld D=%d)", (int)r, rows, D);
    return BDLRU_ERR_CUDA;
  }
  return BDLRU_OK;
}

}  // namespace bdlru
