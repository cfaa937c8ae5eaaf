// Column sums of a row-major [n_rows, n_cols] matrix (fp32 or bf16) -> fp32 [n_cols]: the bias gradient of the
// reference's nn.Linear layers with bias (gates RecBLR.py:165, FFN w_1 / w_2 RecBLR.py:213-214).  ATen's generic
// reduce kernel needs ~60 us for a [102400, 256] bf16 matrix; this streams it once at HBM speed: threads cover a row
// with 16-byte vectors, row-lanes of a CTA stride over the rows with 4 loads in flight, one smem reduction per CTA,
// then the library's deterministic second pass.
#include "common.cuh"

namespace bdlru {

template <typename T, int VW>
__device__ __forceinline__ void ld_vec(const T* p, float (&v)[VW]);
template <>
__device__ __forceinline__ void ld_vec<float, 4>(const float* p, float (&v)[4]) {
  const float4 f = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}
template <>
__device__ __forceinline__ void ld_vec<__nv_bfloat16, 8>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

template <typename T, int VW>
__global__ void __launch_bounds__(256) colsum_rows_kernel(const T* __restrict__ x, long n_rows, int n_cols,
                                                          long row_stride, int tcn, float* __restrict__ part) {
  extern __shared__ float red[];  // [ny][n_cols]
  const int tx = threadIdx.x % tcn, ty = threadIdx.x / tcn, ny = blockDim.x / tcn;
  float acc[VW];
#pragma unroll
  for (int e = 0; e < VW; ++e) acc[e] = 0.f;
  const long step = (long)gridDim.x * ny;
  long r = (long)blockIdx.x * ny + ty;
  for (; r + 3 * step < n_rows; r += 4 * step) {
    float v[4][VW];
#pragma unroll
    for (int u = 0; u < 4; ++u) ld_vec<T, VW>(x + (r + u * step) * row_stride + tx * VW, v[u]);
#pragma unroll
    for (int e = 0; e < VW; ++e) acc[e] += (v[0][e] + v[1][e]) + (v[2][e] + v[3][e]);
  }
  for (; r < n_rows; r += step) {
    float v[VW];
    ld_vec<T, VW>(x + r * row_stride + tx * VW, v);
#pragma unroll
    for (int e = 0; e < VW; ++e) acc[e] += v[e];
  }
#pragma unroll
  for (int e = 0; e < VW; ++e) red[(size_t)ty * n_cols + tx * VW + e] = acc[e];
  __syncthreads();
  for (int c = threadIdx.x; c < n_cols; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < ny; ++k) s += red[(size_t)k * n_cols + c];
    part[(size_t)blockIdx.x * n_cols + c] = s;
  }
}

static int colsum_grid(long n_rows, int ny) {
  long g = (n_rows + ny - 1) / ny;
  const long cap = (long)sm_count() * 4;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace bdlru

using namespace bdlru;

extern "C" BDLRU_API size_t bdlru_colsum_workspace_bytes(int64_t n_rows, int n_cols) {
  (void)n_rows;
  return (size_t)sm_count() * 4 * (size_t)n_cols * sizeof(float);
}

extern "C" BDLRU_API int bdlru_colsum(const void* x, int64_t n_rows, int n_cols, int64_t row_stride, int dtype,
                                      float* out, void* workspace, size_t workspace_bytes, void* stream) {
  BDLRU_REQUIRE(x && out && workspace, "colsum: null pointer");
  BDLRU_REQUIRE(dtype == BDLRU_F32 || dtype == BDLRU_BF16, "colsum: bad dtype %d", dtype);
  const int vw = dtype == BDLRU_F32 ? 4 : 8;
  BDLRU_REQUIRE(n_rows >= 1 && n_cols >= vw && n_cols % vw == 0 && n_cols / vw <= 256,
                "colsum: n_cols=%d must be a multiple of %d and <= %d", n_cols, vw, 256 * vw);
  BDLRU_REQUIRE(row_stride >= n_cols && row_stride % vw == 0 && aligned(x, 16), "colsum: rows must be 16-byte aligned");
  const int tcn = n_cols / vw;
  int ny = 256 / tcn;
  const int threads = ny * tcn;
  const int grid = colsum_grid(n_rows, ny);
  BDLRU_REQUIRE(workspace_bytes >= (size_t)grid * n_cols * sizeof(float), "colsum: workspace too small");
  const size_t smem = (size_t)ny * n_cols * sizeof(float);
  BDLRU_REQUIRE(smem <= 48 * 1024, "colsum: n_cols=%d too wide", n_cols);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* part = reinterpret_cast<float*>(workspace);
  if (dtype == BDLRU_F32)
    colsum_rows_kernel<float, 4><<<grid, threads, smem, st>>>((const float*)x, n_rows, n_cols, row_stride, tcn, part);
  else
    colsum_rows_kernel<__nv_bfloat16, 8><<<grid, threads, smem, st>>>((const __nv_bfloat16*)x, n_rows, n_cols,
                                                                      row_stride, tcn, part);
  BDLRU_LAUNCHED();
  return launch_colsum(part, grid, n_cols, n_cols, COLSUM_SPLIT, out, nullptr, n_cols, nullptr, st);
}
