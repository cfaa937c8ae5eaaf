// Residual epilogue of RecurrentLayer.forward / FeedForward.forward (RecBLR.py:142, 221-225):
//     out = LayerNorm(dropout(x) + residual) * gamma + beta          eps = 1e-12, fp32 statistics
// as ONE kernel (the reference runs dropout, add and an ATen LayerNorm that spends ~150 us on a [102400, 64] tensor).
// One warp per row, lanes hold 4-channel vectors (128-bit I/O), mean/variance by warp shuffles; dropout uses the same
// counter-based hash stream as the front end (common.cuh keep_mask4), keyed by (seed [+ device counter], row, vector) and
// regenerated in the backward.  Backward: LayerNorm backward per row -> d(sum) written as dresidual and, masked, as dx;
// dgamma/dbeta accumulated in registers over a grid-stride loop, block-reduced, finished by a deterministic second pass.
#include "common.cuh"

namespace bdlru {

constexpr int kAV = 4;  // vectors per lane: D <= 512

__device__ __forceinline__ void keep_scale4(uint64_t seed, long row, int vec, float p, float (&m)[4]) {
  keep_mask4(seed ^ 0x2545F4914F6CDD1Dull, ((uint64_t)row << 8) | (uint32_t)vec, p, m);  // vec < 128
}

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LPR lanes cooperate on one row (32 / LPR rows per warp, so D = 64 keeps all lanes busy with two rows per warp);
// each lane holds VPL 4-channel vectors: D <= 4 * LPR * VPL.
template <int LPR>
__device__ __forceinline__ float gsum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, int LPR, int VPL>
__global__ void __launch_bounds__(256) add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         T* __restrict__ out, float* __restrict__ mean_out,
                                                         float* __restrict__ rstd_out, long n_rows, int D, float eps,
                                                         float p, uint64_t seed, const uint64_t* __restrict__ seed_dev) {
  if (seed_dev) seed += *seed_dev;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
  const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  const int nvec = D / 4;
  for (long n0 = gw * RPW; n0 < n_rows; n0 += nw * RPW) {
    const long n = n0 + sub;
    const bool live = n < n_rows;
    float v[VPL][4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vec = sl + LPR * k;
      v[k][0] = v[k][1] = v[k][2] = v[k][3] = 0.f;
      if (live && vec < nvec) {
        float a[4], r[4];
        IO<T>::load(x + n * D + vec * 4, a);
        IO<T>::load(res + n * D + vec * 4, r);
        if (p > 0.f) {
          float m[4];
          keep_scale4(seed, n, vec, p, m);
#pragma unroll
          for (int e = 0; e < 4; ++e) a[e] *= m[e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) v[k][e] = a[e] + r[e];
        s += (v[k][0] + v[k][1]) + (v[k][2] + v[k][3]);
      }
    }
    const float mean = gsum<LPR>(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      if (sl + LPR * k < nvec) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float d = v[k][e] - mean;
          q = fmaf(d, d, q);
        }
      }
    const float rstd = rsqrtf(gsum<LPR>(q) / (float)D + eps);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vec = sl + LPR * k;
      if (live && vec < nvec) {
        const float4 g4 = *reinterpret_cast<const float4*>(gamma + vec * 4);
        const float4 b4 = *reinterpret_cast<const float4*>(beta + vec * 4);
        float o[4];
        o[0] = fmaf((v[k][0] - mean) * rstd, g4.x, b4.x);
        o[1] = fmaf((v[k][1] - mean) * rstd, g4.y, b4.y);
        o[2] = fmaf((v[k][2] - mean) * rstd, g4.z, b4.z);
        o[3] = fmaf((v[k][3] - mean) * rstd, g4.w, b4.w);
        IO<T>::store(out + n * D + vec * 4, o);
      }
    }
    if (live && sl == 0) {
      mean_out[n] = mean;
      rstd_out[n] = rstd;
    }
  }
}

// Backward reads x, res (to rebuild the normalised row), dy; writes dres (= d sum) and dx (= dres * mask).
template <typename T, int LPR, int VPL>
__global__ void __launch_bounds__(256) add_ln_bwd_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                         const float* __restrict__ gamma, const T* __restrict__ dy,
                                                         const float* __restrict__ mean_in,
                                                         const float* __restrict__ rstd_in, T* __restrict__ dx,
                                                         T* __restrict__ dres, float* __restrict__ part, long n_rows,
                                                         int D, float p, uint64_t seed,
                                                         const uint64_t* __restrict__ seed_dev) {
  extern __shared__ float red[];  // [warps * RPW][2][D]
  if (seed_dev) seed += *seed_dev;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int sub = lane / LPR, sl = lane % LPR;
  const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  const int nvec = D / 4;
  float dg[VPL][4], db[VPL][4];
#pragma unroll
  for (int k = 0; k < VPL; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) dg[k][e] = db[k][e] = 0.f;
  for (long n0 = gw * RPW; n0 < n_rows; n0 += nw * RPW) {
    const long n = n0 + sub;
    const bool live = n < n_rows;
    const float mean = live ? mean_in[n] : 0.f, rstd = live ? rstd_in[n] : 0.f;
    float xh[VPL][4], dxh[VPL][4], msk[VPL][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vec = sl + LPR * k;
#pragma unroll
      for (int e = 0; e < 4; ++e) xh[k][e] = dxh[k][e] = 0.f, msk[k][e] = 1.f;
      if (live && vec < nvec) {
        float a[4], r[4], g[4];
        IO<T>::load(x + n * D + vec * 4, a);
        IO<T>::load(res + n * D + vec * 4, r);
        IO<T>::load(dy + n * D + vec * 4, g);
        if (p > 0.f) keep_scale4(seed, n, vec, p, msk[k]);
        const float4 g4 = *reinterpret_cast<const float4*>(gamma + vec * 4);
        const float gm[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          xh[k][e] = (fmaf(a[e], msk[k][e], r[e]) - mean) * rstd;
          dxh[k][e] = g[e] * gm[e];
          dg[k][e] = fmaf(g[e], xh[k][e], dg[k][e]);
          db[k][e] += g[e];
          s1 += dxh[k][e];
          s2 = fmaf(dxh[k][e], xh[k][e], s2);
        }
      }
    }
    s1 = gsum<LPR>(s1) / (float)D;
    s2 = gsum<LPR>(s2) / (float)D;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vec = sl + LPR * k;
      if (live && vec < nvec) {
        float ds[4], dxx[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ds[e] = rstd * (dxh[k][e] - s1 - xh[k][e] * s2);
          dxx[e] = ds[e] * msk[k][e];
        }
        IO<T>::store(dres + n * D + vec * 4, ds);
        if (dx != dres) IO<T>::store(dx + n * D + vec * 4, dxx);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vec = sl + LPR * k;
    if (vec < nvec) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        red[((size_t)(warp * RPW + sub) * 2 + 0) * D + vec * 4 + e] = dg[k][e];
        red[((size_t)(warp * RPW + sub) * 2 + 1) * D + vec * 4 + e] = db[k][e];
      }
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * D; idx += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nwarp * RPW; ++w) s += red[(size_t)w * 2 * D + idx];
    part[(size_t)blockIdx.x * 2 * D + idx] = s;
  }
}

// (LPR, VPL) for a row of D channels
#define ADD_LN_DISPATCH(D, CALL)                         \
  do {                                                   \
    const int nv_ = (D) / 4;                             \
    if (nv_ <= 8) { CALL(8, 1); }                        \
    else if (nv_ <= 16) { CALL(16, 1); }                 \
    else if (nv_ <= 32) { CALL(32, 1); }                 \
    else if (nv_ <= 64) { CALL(32, 2); }                 \
    else { CALL(32, 4); }                                \
  } while (0)

static int add_ln_rpw(int D) { const int nv = D / 4; return nv <= 8 ? 4 : (nv <= 16 ? 2 : 1); }

static int add_ln_grid(long n_rows, int D) {
  const int rpb = 8 * add_ln_rpw(D);
  long blocks = (n_rows + rpb - 1) / rpb;
  const long cap = (long)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

static int add_ln_check(long n_rows, int D, int dtype, float p) {
  BDLRU_REQUIRE(n_rows >= 1, "add_ln: n_rows=%ld", n_rows);
  BDLRU_REQUIRE(D >= 4 && D % 4 == 0 && D <= 128 * kAV, "add_ln: D=%d must be a multiple of 4 and <= %d", D, 128 * kAV);
  BDLRU_REQUIRE(dtype == BDLRU_F32 || dtype == BDLRU_BF16, "add_ln: bad dtype %d", dtype);
  BDLRU_REQUIRE(p >= 0.f && p < 1.f, "add_ln: dropout_p=%f not in [0, 1)", p);
  return BDLRU_OK;
}

}  // namespace bdlru

using namespace bdlru;

extern "C" BDLRU_API int bdlru_add_ln_fwd(const void* x, const void* residual, const float* gamma, const float* beta,
                                          void* out, float* mean, float* rstd, int64_t n_rows, int D, float eps,
                                          float dropout_p, uint64_t seed, const uint64_t* seed_device, int dtype,
                                          void* stream) {
  int rc = add_ln_check(n_rows, D, dtype, dropout_p);
  if (rc) return rc;
  BDLRU_REQUIRE(x && residual && gamma && beta && out && mean && rstd, "add_ln_fwd: null pointer");
  BDLRU_REQUIRE(aligned(x, 8) && aligned(residual, 8) && aligned(out, 8) && aligned(gamma, 16) && aligned(beta, 16),
                "add_ln_fwd: misaligned pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = add_ln_grid(n_rows, D);
  if (dtype == BDLRU_F32) {
    BDLRU_REQUIRE(aligned(x, 16) && aligned(residual, 16) && aligned(out, 16), "add_ln_fwd: misaligned fp32 pointer");
#define FWD_F32(LPR, VPL)                                                                                          \
  add_ln_fwd_kernel<float, LPR, VPL><<<grid, 256, 0, st>>>((const float*)x, (const float*)residual, gamma, beta,   \
                                                           (float*)out, mean, rstd, n_rows, D, eps, dropout_p, seed, \
                                                           seed_device)
    ADD_LN_DISPATCH(D, FWD_F32);
#undef FWD_F32
  } else {
#define FWD_BF16(LPR, VPL)                                                                                       \
  add_ln_fwd_kernel<__nv_bfloat16, LPR, VPL><<<grid, 256, 0, st>>>(                                                \
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)residual, gamma, beta, (__nv_bfloat16*)out, mean, rstd, n_rows, \
      D, eps, dropout_p, seed, seed_device)
    ADD_LN_DISPATCH(D, FWD_BF16);
#undef FWD_BF16
  }
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

extern "C" BDLRU_API size_t bdlru_add_ln_bwd_workspace_bytes(int64_t n_rows, int D) {
  (void)n_rows;
  return (size_t)sm_count() * 8 * 2 * (size_t)D * sizeof(float);
}

extern "C" BDLRU_API int bdlru_add_ln_bwd(const void* x, const void* residual, const float* gamma, const void* grad_out,
                                          const float* mean, const float* rstd, void* dx, void* dresidual, float* dgamma,
                                          float* dbeta, void* workspace, size_t workspace_bytes, int64_t n_rows, int D,
                                          float dropout_p, uint64_t seed, const uint64_t* seed_device, int dtype,
                                          void* stream) {
  int rc = add_ln_check(n_rows, D, dtype, dropout_p);
  if (rc) return rc;
  BDLRU_REQUIRE(x && residual && gamma && grad_out && mean && rstd && dx && dresidual && dgamma && dbeta,
                "add_ln_bwd: null pointer");
  const int grid = add_ln_grid(n_rows, D);
  const size_t need = (size_t)grid * 2 * D * sizeof(float);
  BDLRU_REQUIRE(workspace && workspace_bytes >= need, "add_ln_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* part = reinterpret_cast<float*>(workspace);
  const size_t smem = (size_t)8 * add_ln_rpw(D) * 2 * D * sizeof(float);
#define BWD_F32(LPR, VPL)                                                                                            \
  add_ln_bwd_kernel<float, LPR, VPL><<<grid, 256, smem, st>>>((const float*)x, (const float*)residual, gamma,        \
                                                              (const float*)grad_out, mean, rstd, (float*)dx,        \
                                                              (float*)dresidual, part, n_rows, D, dropout_p, seed,   \
                                                              seed_device)
#define BWD_BF16(LPR, VPL)                                                                                           \
  add_ln_bwd_kernel<__nv_bfloat16, LPR, VPL><<<grid, 256, smem, st>>>(                                               \
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)residual, gamma, (const __nv_bfloat16*)grad_out, mean, rstd,    \
      (__nv_bfloat16*)dx, (__nv_bfloat16*)dresidual, part, n_rows, D, dropout_p, seed, seed_device)
  if (dtype == BDLRU_F32) {
    ADD_LN_DISPATCH(D, BWD_F32);
  } else {
    ADD_LN_DISPATCH(D, BWD_BF16);
  }
#undef BWD_F32
#undef BWD_BF16
  BDLRU_LAUNCHED();
  return launch_colsum(part, grid, 2 * D, 2 * D, COLSUM_SPLIT, dgamma, dbeta, D, nullptr, st);
}
