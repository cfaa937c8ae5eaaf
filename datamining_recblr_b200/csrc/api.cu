// Library-wide plumbing of the C ABI (include/bdlru.h): version, thread-local error text, launch counter.
#include <stdarg.h>

#include "common.cuh"

namespace bdlru {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();  // no device (build container): sizes are computed for a B200
    return 148;
  }
  cached = n;
  return n;
}

__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ part, int n_rows, int row_stride,
                                                     int n_cols, int mode, float* __restrict__ out0,
                                                     float* __restrict__ out1, int split, const float* __restrict__ aux) {
  constexpr int RL = 32;  // row lanes per CTA
  __shared__ float red[RL][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < n_cols) {
    float s0 = 0.f, s1 = 0.f;  // two independent chains: the loads of consecutive rows overlap
    int r = ty;
    for (; r + RL < n_rows; r += 2 * RL) {
      s0 += part[(size_t)r * row_stride + c];
      s1 += part[(size_t)(r + RL) * row_stride + c];
    }
    if (r < n_rows) s0 += part[(size_t)r * row_stride + c];
    s = s0 + s1;
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < n_cols) {
    float a = red[0][tx];
#pragma unroll
    for (int k = 1; k < RL; ++k) a += red[k][tx];
    if (mode == COLSUM_SPLIT) {
      if (c < split) out0[c] = a;
      else if (out1) out1[c - split] = a;
    } else if (mode == COLSUM_SIGMOID) {
      out0[c] = a * (1.0f / (1.0f + expf(-aux[c])));
    } else {
      const int ch = c / (split + 1), j = c % (split + 1);
      if (j < split) out0[ch * split + j] = a;
      else if (out1) out1[ch] = a;
    }
  }
}

int launch_colsum(const float* part, int n_rows, int row_stride, int n_cols, int mode, float* out0, float* out1,
                  int split, const float* aux, cudaStream_t st) {
  colsum_kernel<<<(n_cols + 31) / 32, 1024, 0, st>>>(part, n_rows, row_stride, n_cols, mode, out0, out1, split, aux);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

}  // namespace bdlru

extern "C" BDLRU_API int bdlru_version(void) { return 4; }
extern "C" BDLRU_API const char* bdlru_last_error(void) { return bdlru::g_err; }
extern "C" BDLRU_API uint64_t bdlru_launch_count(void) { return bdlru::g_launches.load(); }

#ifndef BDLRU_SOURCE_DIGEST
#define BDLRU_SOURCE_DIGEST "unknown"
#endif
#ifdef BDLRU_TUNING
#define BDLRU_TUNING_TAG " tuning"
#else
#define BDLRU_TUNING_TAG ""
#endif
extern "C" BDLRU_API const char* bdlru_build_info(void) { return "src=" BDLRU_SOURCE_DIGEST BDLRU_TUNING_TAG; }
