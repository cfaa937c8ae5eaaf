// Residual epilogue of RecurrentLayer.forward / FeedForward.forward (RecBLR.py:142, 221-225):
//     out = LayerNorm(dropout(x) + residual) * gamma + beta          eps = 1e-12, fp32 statistics
// as ONE kernel (the reference runs dropout, add and an ATen LayerNorm that spends ~150 us on a [102400, 64] tensor).
// One warp per row, lanes hold 4-channel vectors (128-bit I/O), mean/variance by warp shuffles; dropout uses the same
// counter-based hash stream as the front end (common.cuh keep_mask4), keyed by (seed [+ device counter], row, vector) and
// regenerated in the backward.  Backward: LayerNorm backward per row -> d(sum) written as dresidual and, masked, as dx;
// dgamma/dbeta accumulated in registers over a grid-stride loop, block-reduced, finished by a deterministic second pass.
#include "common.cuh"

namespace bdlru {

#ifndef BDLRU_ADD_LN_FWD_MINB
#define BDLRU_ADD_LN_FWD_MINB 1
#endif
#ifndef BDLRU_ADD_LN_BWD_MINB
#define BDLRU_ADD_LN_BWD_MINB 1
#endif
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

// CPV channels per 16-byte vector (4 fp32 / 8 bf16; bf16 rows with D % 8 != 0 fall back to 4 = 8-byte accesses: at 8 bytes
// per lane the kernels were bound by load/store instruction issue at ~3 TB/s).  The dropout stream is keyed per 4
// channels in every case, so the mask does not depend on the vector width.
template <int CPV>
__device__ __forceinline__ void keep_scale(uint64_t seed, long row, int vec, float p, float (&m)[CPV]) {
#pragma unroll
  for (int h = 0; h < CPV / 4; ++h) {
    float t[4];
    keep_scale4(seed, row, vec * (CPV / 4) + h, p, t);
#pragma unroll
    for (int e = 0; e < 4; ++e) m[h * 4 + e] = t[e];
  }
}

template <int CPV>
__device__ __forceinline__ void load_f32(const float* __restrict__ p, float (&v)[CPV]) {
#pragma unroll
  for (int h = 0; h < CPV / 4; ++h) {
    const float4 f = *reinterpret_cast<const float4*>(p + h * 4);
    v[h * 4 + 0] = f.x, v[h * 4 + 1] = f.y, v[h * 4 + 2] = f.z, v[h * 4 + 3] = f.w;
  }
}

// R rows per warp iteration have their loads issued before any arithmetic.
template <typename T, int CPV, int LPR, int VPL, int R>
__global__ void __launch_bounds__(256, BDLRU_ADD_LN_FWD_MINB) add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         T* __restrict__ out, float* __restrict__ mean_out,
                                                         float* __restrict__ rstd_out, long n_rows, int D, float eps,
                                                         float p, uint64_t seed, const uint64_t* __restrict__ seed_dev) {
  if (seed_dev) seed += *seed_dev;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
  const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  const int nvec = D / CPV;
  for (long n0 = gw * (RPW * R); n0 < n_rows; n0 += nw * (RPW * R)) {
    float v[R][VPL][CPV], rr[R][VPL][CPV];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long n = n0 + j * RPW + sub;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vec = sl + LPR * k;
#pragma unroll
        for (int e = 0; e < CPV; ++e) v[j][k][e] = rr[j][k][e] = 0.f;
        if (n < n_rows && vec < nvec) {
          IOV<T, CPV>::load(x + n * D + vec * CPV, v[j][k]);
          IOV<T, CPV>::load(res + n * D + vec * CPV, rr[j][k]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long n = n0 + j * RPW + sub;
      const bool live = n < n_rows;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vec = sl + LPR * k;
        if (live && vec < nvec) {
          if (p > 0.f) {
            float m[CPV];
            keep_scale<CPV>(seed, n, vec, p, m);
#pragma unroll
            for (int e = 0; e < CPV; ++e) v[j][k][e] *= m[e];
          }
#pragma unroll
          for (int e = 0; e < CPV; ++e) {
            v[j][k][e] += rr[j][k][e];
            s += v[j][k][e];
          }
        }
      }
      const float mean = gsum<LPR>(s) / (float)D;
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k)
        if (sl + LPR * k < nvec) {
#pragma unroll
          for (int e = 0; e < CPV; ++e) {
            const float d = v[j][k][e] - mean;
            q = fmaf(d, d, q);
          }
        }
      const float rstd = rsqrtf(gsum<LPR>(q) / (float)D + eps);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vec = sl + LPR * k;
        if (live && vec < nvec) {
          float gm[CPV], bt[CPV], o[CPV];
          load_f32<CPV>(gamma + vec * CPV, gm);
          load_f32<CPV>(beta + vec * CPV, bt);
#pragma unroll
          for (int e = 0; e < CPV; ++e) o[e] = fmaf((v[j][k][e] - mean) * rstd, gm[e], bt[e]);
          IOV<T, CPV>::store(out + n * D + vec * CPV, o);
        }
      }
      if (live && sl == 0) {
        mean_out[n] = mean;
        rstd_out[n] = rstd;
      }
    }
  }
}

// Backward reads x, res (to rebuild the normalised row), dy; writes dres (= d sum) and dx (= dres * mask).
template <typename T, int CPV, int LPR, int VPL, int R>
__global__ void __launch_bounds__(256, BDLRU_ADD_LN_BWD_MINB) add_ln_bwd_kernel(const T* __restrict__ x, const T* __restrict__ res,
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
  const int nvec = D / CPV;
  float dg[VPL][CPV], db[VPL][CPV], gm[VPL][CPV];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vec = sl + LPR * k;
#pragma unroll
    for (int e = 0; e < CPV; ++e) dg[k][e] = db[k][e] = gm[k][e] = 0.f;
    if (vec < nvec) load_f32<CPV>(gamma + vec * CPV, gm[k]);
  }
  for (long n0 = gw * (RPW * R); n0 < n_rows; n0 += nw * (RPW * R)) {
    float a[R][VPL][CPV], r[R][VPL][CPV], g[R][VPL][CPV], mean[R], rstd[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long n = n0 + j * RPW + sub;
      const bool live = n < n_rows;
      mean[j] = live ? mean_in[n] : 0.f;
      rstd[j] = live ? rstd_in[n] : 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vec = sl + LPR * k;
#pragma unroll
        for (int e = 0; e < CPV; ++e) a[j][k][e] = r[j][k][e] = g[j][k][e] = 0.f;
        if (live && vec < nvec) {
          IOV<T, CPV>::load(x + n * D + vec * CPV, a[j][k]);
          IOV<T, CPV>::load(res + n * D + vec * CPV, r[j][k]);
          IOV<T, CPV>::load(dy + n * D + vec * CPV, g[j][k]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long n = n0 + j * RPW + sub;
      const bool live = n < n_rows;
      float msk[VPL][CPV];
      float s1 = 0.f, s2 = 0.f;
      // a <- normalised row, g <- dy * gamma
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vec = sl + LPR * k;
#pragma unroll
        for (int e = 0; e < CPV; ++e) msk[k][e] = 1.f;
        if (live && vec < nvec) {
          if (p > 0.f) keep_scale<CPV>(seed, n, vec, p, msk[k]);
#pragma unroll
          for (int e = 0; e < CPV; ++e) {
            const float xh = (fmaf(a[j][k][e], msk[k][e], r[j][k][e]) - mean[j]) * rstd[j];
            const float gy = g[j][k][e];
            dg[k][e] = fmaf(gy, xh, dg[k][e]);
            db[k][e] += gy;
            a[j][k][e] = xh;
            g[j][k][e] = gy * gm[k][e];
            s1 += g[j][k][e];
            s2 = fmaf(g[j][k][e], xh, s2);
          }
        }
      }
      s1 = gsum<LPR>(s1) / (float)D;
      s2 = gsum<LPR>(s2) / (float)D;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vec = sl + LPR * k;
        if (live && vec < nvec) {
          float ds[CPV], dxx[CPV];
#pragma unroll
          for (int e = 0; e < CPV; ++e) {
            ds[e] = rstd[j] * (g[j][k][e] - s1 - a[j][k][e] * s2);
            dxx[e] = ds[e] * msk[k][e];
          }
          IOV<T, CPV>::store(dres + n * D + vec * CPV, ds);
          if (dx != dres) IOV<T, CPV>::store(dx + n * D + vec * CPV, dxx);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vec = sl + LPR * k;
    if (vec < nvec) {
#pragma unroll
      for (int e = 0; e < CPV; ++e) {
        red[((size_t)(warp * RPW + sub) * 2 + 0) * D + vec * CPV + e] = dg[k][e];
        red[((size_t)(warp * RPW + sub) * 2 + 1) * D + vec * CPV + e] = db[k][e];
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

#ifndef BDLRU_ADD_LN_FWD_ROWS
#define BDLRU_ADD_LN_FWD_ROWS 1
#endif
#ifndef BDLRU_ADD_LN_BWD_ROWS
#define BDLRU_ADD_LN_BWD_ROWS 1
#endif
#ifndef BDLRU_ADD_LN_BPSM
#define BDLRU_ADD_LN_BPSM 0
#endif
constexpr int kAddLnBlocksPerSm = BDLRU_ADD_LN_BPSM;  // tuning: extra grid cap in blocks per SM (0 = occupancy only)
constexpr int kAddLnFwdRows = BDLRU_ADD_LN_FWD_ROWS;  // rows in flight per warp iteration (1 for the widest rows)
constexpr int kAddLnBwdRows = BDLRU_ADD_LN_BWD_ROWS;

// (LPR, VPL, R) for a row of D / CPV vectors
#define ADD_LN_DISPATCH(NV, ROWS, CALL)                          \
  do {                                                           \
    if ((NV) <= 8) { CALL(8, 1, ROWS); }                         \
    else if ((NV) <= 16) { CALL(16, 1, ROWS); }                  \
    else if ((NV) <= 32) { CALL(32, 1, ROWS); }                  \
    else if ((NV) <= 64) { CALL(32, 2, (ROWS > 2 ? 2 : ROWS)); } \
    else { CALL(32, 4, 1); }                                     \
  } while (0)

static int add_ln_cpv(int D, int dtype) { return (dtype == BDLRU_BF16 && D % 8 == 0) ? 8 : 4; }

static int add_ln_rpw(int nv) { return nv <= 8 ? 4 : (nv <= 16 ? 2 : 1); }

// Grid-stride kernels: the grid is capped at exactly the blocks that are resident at once (occupancy of that
// instantiation x SMs) — a cap above residency runs a ragged second wave (measured: 8 per SM with 3 resident cost 6 %).
template <typename K>
static int add_ln_grid(K kernel, size_t smem, long n_rows, int nv) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  if (kAddLnBlocksPerSm > 0 && per_sm > kAddLnBlocksPerSm) per_sm = kAddLnBlocksPerSm;
  const int rpb = 8 * add_ln_rpw(nv);
  long blocks = (n_rows + rpb - 1) / rpb;
  const long cap = (long)sm_count() * per_sm;
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
  const bool a16 = aligned(x, 16) && aligned(residual, 16) && aligned(out, 16);
  if (dtype == BDLRU_F32) {
    BDLRU_REQUIRE(a16, "add_ln_fwd: misaligned fp32 pointer");
    const int nv = D / 4;
#define FWD_F32(LPR, VPL, R)                                                                                     \
  {                                                                                                              \
    auto kern = add_ln_fwd_kernel<float, 4, LPR, VPL, R>;                                                        \
    kern<<<add_ln_grid(kern, 0, n_rows, nv), 256, 0, st>>>((const float*)x, (const float*)residual, gamma, beta, \
                                                           (float*)out, mean, rstd, n_rows, D, eps, dropout_p,   \
                                                           seed, seed_device);                                   \
  }
    ADD_LN_DISPATCH(nv, kAddLnFwdRows, FWD_F32);
#undef FWD_F32
  } else {
#define FWD_BF16(CPV, LPR, VPL, R)                                                                               \
  {                                                                                                              \
    auto kern = add_ln_fwd_kernel<__nv_bfloat16, CPV, LPR, VPL, R>;                                              \
    kern<<<add_ln_grid(kern, 0, n_rows, nv), 256, 0, st>>>(                                                      \
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)residual, gamma, beta, (__nv_bfloat16*)out, mean, rstd,   \
        n_rows, D, eps, dropout_p, seed, seed_device);                                                           \
  }
#define FWD_BF16_8(LPR, VPL, R) FWD_BF16(8, LPR, VPL, R)
#define FWD_BF16_4(LPR, VPL, R) FWD_BF16(4, LPR, VPL, R)
    const int cpv = a16 ? add_ln_cpv(D, dtype) : 4;
    const int nv = D / cpv;
    if (cpv == 8) {
      ADD_LN_DISPATCH(nv, kAddLnFwdRows, FWD_BF16_8);
    } else {
      ADD_LN_DISPATCH(nv, kAddLnFwdRows, FWD_BF16_4);
    }
#undef FWD_BF16_8
#undef FWD_BF16_4
#undef FWD_BF16
  }
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

extern "C" BDLRU_API size_t bdlru_add_ln_bwd_workspace_bytes(int64_t n_rows, int D) {
  (void)n_rows;
  return (size_t)sm_count() * 8 * 2 * (size_t)D * sizeof(float);  // <= 8 resident 256-thread blocks per SM
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
  const bool a16 = aligned(x, 16) && aligned(residual, 16) && aligned(grad_out, 16) && aligned(dx, 16) &&
                   aligned(dresidual, 16);
  BDLRU_REQUIRE(aligned(x, 8) && aligned(residual, 8) && aligned(grad_out, 8) && aligned(dx, 8) &&
                aligned(dresidual, 8) && aligned(gamma, 16) && (a16 || dtype == BDLRU_BF16),
                "add_ln_bwd: misaligned pointer");
  const int cpv = a16 ? add_ln_cpv(D, dtype) : 4;
  const int nv = D / cpv;
  const size_t need = bdlru_add_ln_bwd_workspace_bytes(n_rows, D);
  BDLRU_REQUIRE(workspace && workspace_bytes >= need, "add_ln_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* part = reinterpret_cast<float*>(workspace);
  const size_t smem = (size_t)8 * add_ln_rpw(nv) * 2 * D * sizeof(float);
  int grid = 1;
#define BWD_F32(LPR, VPL, R)                                                                                      \
  {                                                                                                               \
    auto kern = add_ln_bwd_kernel<float, 4, LPR, VPL, R>;                                                         \
    grid = add_ln_grid(kern, smem, n_rows, nv);                                                                   \
    kern<<<grid, 256, smem, st>>>((const float*)x, (const float*)residual, gamma, (const float*)grad_out, mean,   \
                                  rstd, (float*)dx, (float*)dresidual, part, n_rows, D, dropout_p, seed,          \
                                  seed_device);                                                                   \
  }
#define BWD_BF16(CPV, LPR, VPL, R)                                                                                \
  {                                                                                                               \
    auto kern = add_ln_bwd_kernel<__nv_bfloat16, CPV, LPR, VPL, R>;                                               \
    grid = add_ln_grid(kern, smem, n_rows, nv);                                                                   \
    kern<<<grid, 256, smem, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)residual, gamma,                 \
                                  (const __nv_bfloat16*)grad_out, mean, rstd, (__nv_bfloat16*)dx,                 \
                                  (__nv_bfloat16*)dresidual, part, n_rows, D, dropout_p, seed, seed_device);      \
  }
#define BWD_BF16_8(LPR, VPL, R) BWD_BF16(8, LPR, VPL, R)
#define BWD_BF16_4(LPR, VPL, R) BWD_BF16(4, LPR, VPL, R)
  if (dtype == BDLRU_F32) {
    ADD_LN_DISPATCH(nv, kAddLnBwdRows, BWD_F32);
  } else if (cpv == 8) {
    ADD_LN_DISPATCH(nv, kAddLnBwdRows, BWD_BF16_8);
  } else {
    ADD_LN_DISPATCH(nv, kAddLnBwdRows, BWD_BF16_4);
  }
#undef BWD_F32
#undef BWD_BF16_8
#undef BWD_BF16_4
#undef BWD_BF16
  BDLRU_LAUNCHED();
  return launch_colsum(part, grid, 2 * D, 2 * D, COLSUM_SPLIT, dgamma, dbeta, D, nullptr, st);
}
