// Front end of RecBLR.forward (RecBLR.py:76-78): item-embedding gather -> dropout -> LayerNorm(eps),
// fused so the [B, L, D] activation makes one HBM round trip (one 4*D-byte random row read + one write).
//
// One warp per token; lanes hold 4-channel vectors (128-bit row reads), mean/variance by warp shuffles,
// fp32 statistics.  Dropout uses a counter-based hash (common.cuh keep_mask4) keyed by (seed, token, channel vector) so the backward
// regenerates the mask instead of storing it.  Backward: LayerNorm backward per token, scatter-add of the
// row gradient into dtable with vector atomics (red.global.add.v4.f32), dgamma/dbeta accumulated in
// registers across a grid-stride loop and reduced deterministically in a second pass.
#include "common.cuh"

namespace bdlru {

constexpr int kMaxV = 4;  // 4-channel vectors per lane: D <= 32 * 4 * kMaxV = 512

// keep-mask * 1/(1-p) for the 4 channels of vector `vec` of token `tok`
__device__ __forceinline__ void dropout_scale4(uint64_t seed, long tok, int vec, float p, float (&m)[4]) {
  keep_mask4(seed, ((uint64_t)tok << 8) | (uint32_t)vec, p, m);  // vec < 128
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LPR lanes cooperate on one token (32 / LPR tokens per warp: D = 64 keeps every lane busy with two tokens per warp);
// each lane holds VPL 4-channel vectors: D <= 4 * LPR * VPL.
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// R tokens per warp iteration: their ids are loaded first, then their rows, then the arithmetic — the row read depends on
// the id read and lands anywhere in the table, so one token per warp left ~8 KB per SM in flight (0.26 of HBM at 10 M rows).
template <typename T, typename TO, int LPR, int VPL, int R>
__global__ void __launch_bounds__(256) embed_ln_fwd_kernel(const int64_t* __restrict__ ids, const T* __restrict__ table,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, TO* __restrict__ out,
                                                           float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                           long n_tokens, long n_items, int D, float eps, float p,
                                                           uint64_t seed, const uint64_t* __restrict__ seed_dev) {
  if (seed_dev) seed += *seed_dev;  // device-side step counter: lets a captured CUDA graph draw a new mask per replay
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
  const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  const int nvec = D / 4;
  for (long n0 = gw * (RPW * R); n0 < n_tokens; n0 += nw * (RPW * R)) {
    long idv[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long n = n0 + j * RPW + sub;
      long id = n < n_tokens ? ids[n] : 0;
      idv[j] = id < 0 ? 0 : (id >= n_items ? n_items - 1 : id);
    }
    float x[R][VPL][4];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long n = n0 + j * RPW + sub;
      const T* row = table + idv[j] * D;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vec = sl + LPR * v;
        x[j][v][0] = x[j][v][1] = x[j][v][2] = x[j][v][3] = 0.f;
        if (n < n_tokens && vec < nvec) IO<T>::load(row + vec * 4, x[j][v]);
      }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long n = n0 + j * RPW + sub;
      const bool live = n < n_tokens;
      float s = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vec = sl + LPR * v;
        if (live && vec < nvec) {
          if (p > 0.f) {
            float m[4];
            dropout_scale4(seed, n, vec, p, m);
#pragma unroll
            for (int e = 0; e < 4; ++e) x[j][v][e] *= m[e];
          }
          s += (x[j][v][0] + x[j][v][1]) + (x[j][v][2] + x[j][v][3]);
        }
      }
      const float mean = group_sum<LPR>(s) / (float)D;
      float q = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        if (sl + LPR * v < nvec) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float d = x[j][v][e] - mean;
            q = fmaf(d, d, q);
          }
        }
      }
      const float rstd = rsqrtf(group_sum<LPR>(q) / (float)D + eps);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vec = sl + LPR * v;
        if (live && vec < nvec) {
          const float4 g4 = *reinterpret_cast<const float4*>(gamma + vec * 4);
          const float4 b4 = *reinterpret_cast<const float4*>(beta + vec * 4);
          float o[4];
          o[0] = fmaf((x[j][v][0] - mean) * rstd, g4.x, b4.x);
          o[1] = fmaf((x[j][v][1] - mean) * rstd, g4.y, b4.y);
          o[2] = fmaf((x[j][v][2] - mean) * rstd, g4.z, b4.z);
          o[3] = fmaf((x[j][v][3] - mean) * rstd, g4.w, b4.w);
          IO<TO>::store(out + n * D + vec * 4, o);
        }
      }
      if (live && sl == 0) {
        mean_out[n] = mean;
        rstd_out[n] = rstd;
      }
    }
  }
}

template <typename T, typename TO, int LPR, int VPL, int R>
__global__ void __launch_bounds__(256) embed_ln_bwd_kernel(const int64_t* __restrict__ ids, const T* __restrict__ table,
                                                           const float* __restrict__ gamma, const TO* __restrict__ dy,
                                                           const float* __restrict__ mean_in,
                                                           const float* __restrict__ rstd_in, float* __restrict__ dtable,
                                                           float* __restrict__ part /* [grid][2][D] */, long n_tokens,
                                                           long n_items, int D, float p, uint64_t seed,
                                                           const uint64_t* __restrict__ seed_dev, long padding_idx,
                                                           TO* __restrict__ drows /* nullable: [n_tokens][D] */) {
  extern __shared__ float red[];  // [warps * RPW][2][D]
  if (seed_dev) seed += *seed_dev;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int sub = lane / LPR, sl = lane % LPR;
  const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  const int nvec = D / 4;
  float dg[VPL][4], db[VPL][4], gm[VPL][4];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int vec = sl + LPR * v;
    const float4 g4 = vec < nvec ? *reinterpret_cast<const float4*>(gamma + vec * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    gm[v][0] = g4.x, gm[v][1] = g4.y, gm[v][2] = g4.z, gm[v][3] = g4.w;
#pragma unroll
    for (int e = 0; e < 4; ++e) dg[v][e] = db[v][e] = 0.f;
  }

  for (long n0 = gw * (RPW * R); n0 < n_tokens; n0 += nw * (RPW * R)) {
    // ids, then the rows / dy / statistics of R tokens, then the arithmetic (the row read depends on the id read)
    long idv[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long n = n0 + j * RPW + sub;
      long id = n < n_tokens ? ids[n] : 0;
      idv[j] = id < 0 ? 0 : (id >= n_items ? n_items - 1 : id);
    }
    float x[R][VPL][4], g[R][VPL][4], mean[R], rstd[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long n = n0 + j * RPW + sub;
      const bool live = n < n_tokens;
      const T* row = table + idv[j] * D;
      mean[j] = live ? mean_in[n] : 0.f;
      rstd[j] = live ? rstd_in[n] : 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vec = sl + LPR * v;
#pragma unroll
        for (int e = 0; e < 4; ++e) x[j][v][e] = g[j][v][e] = 0.f;
        if (live && vec < nvec) {
          IO<T>::load(row + vec * 4, x[j][v]);
          IO<TO>::load(dy + n * D + vec * 4, g[j][v]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const long n = n0 + j * RPW + sub;
      const bool live = n < n_tokens;
      const long id = idv[j];
      float msk[VPL][4];
      float s1 = 0.f, s2 = 0.f;
      // x <- normalised row, g <- dy * gamma
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vec = sl + LPR * v;
#pragma unroll
        for (int e = 0; e < 4; ++e) msk[v][e] = 1.f;
        if (live && vec < nvec) {
          if (p > 0.f) dropout_scale4(seed, n, vec, p, msk[v]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float xh = (x[j][v][e] * msk[v][e] - mean[j]) * rstd[j];
            const float gy = g[j][v][e];
            dg[v][e] = fmaf(gy, xh, dg[v][e]);
            db[v][e] += gy;
            x[j][v][e] = xh;
            g[j][v][e] = gy * gm[v][e];
            s1 += g[j][v][e];
            s2 = fmaf(g[j][v][e], xh, s2);
          }
        }
      }
      s1 = group_sum<LPR>(s1) / (float)D;
      s2 = group_sum<LPR>(s2) / (float)D;
      if (live && (drows || id != padding_idx)) {
        const bool pad = id == padding_idx;   // row-gradient mode keeps the slot (zeros): the exchange is dense per token
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int vec = sl + LPR * v;
          if (vec < nvec) {
            float4 d;
            d.x = rstd[j] * (g[j][v][0] - s1 - x[j][v][0] * s2) * msk[v][0];
            d.y = rstd[j] * (g[j][v][1] - s1 - x[j][v][1] * s2) * msk[v][1];
            d.z = rstd[j] * (g[j][v][2] - s1 - x[j][v][2] * s2) * msk[v][2];
            d.w = rstd[j] * (g[j][v][3] - s1 - x[j][v][3] * s2) * msk[v][3];
            if (drows) {
              const float o[4] = {pad ? 0.f : d.x, pad ? 0.f : d.y, pad ? 0.f : d.z, pad ? 0.f : d.w};
              IO<TO>::store(drows + n * D + vec * 4, o);
            } else {
              atomicAdd(reinterpret_cast<float4*>(dtable + id * D + vec * 4), d);
            }
          }
        }
      }
    }
  }
  // block reduce dgamma / dbeta over warps and row groups -> one partial row per CTA
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int vec = sl + LPR * v;
    if (vec < nvec) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        red[((size_t)(warp * RPW + sub) * 2 + 0) * D + vec * 4 + e] = dg[v][e];
        red[((size_t)(warp * RPW + sub) * 2 + 1) * D + vec * 4 + e] = db[v][e];
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

// dst[(id - row_lo) * D + :] += rows[n * D + :] for every token n whose id lies in [row_lo, row_hi) and is not
// padding_idx: the owner-side half of the sharded embedding gradient (the rows arrive by all-gather from every rank;
// a rank adds only what it owns).  One lane per 4-channel vector, vector red.global.add.
template <typename TO>
__global__ void __launch_bounds__(256) scatter_rows_kernel(const int64_t* __restrict__ ids, const TO* __restrict__ rows,
                                                           long n_tokens, int D, long row_lo, long row_hi,
                                                           long padding_idx, float* __restrict__ dst) {
  const int nvec = D / 4;
  const long total = n_tokens * nvec;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / nvec;
    const int vec = (int)(i - n * nvec);
    const long id = ids[n];
    if (id < row_lo || id >= row_hi || id == padding_idx) continue;
    float v[4];
    IO<TO>::load(rows + n * D + vec * 4, v);
    atomicAdd(reinterpret_cast<float4*>(dst + (id - row_lo) * D + vec * 4), make_float4(v[0], v[1], v[2], v[3]));
  }
}

#define EMBED_DISPATCH(D, CALL)                          \
  do {                                                   \
    const int nv_ = (D) / 4;                             \
    if (nv_ <= 8) { CALL(8, 1); }                        \
    else if (nv_ <= 16) { CALL(16, 1); }                 \
    else if (nv_ <= 32) { CALL(32, 1); }                 \
    else if (nv_ <= 64) { CALL(32, 2); }                 \
    else { CALL(32, 4); }                                \
  } while (0)

#ifndef BDLRU_EMBED_FWD_ROWS
#define BDLRU_EMBED_FWD_ROWS 4
#endif
#ifndef BDLRU_EMBED_BWD_ROWS
#define BDLRU_EMBED_BWD_ROWS 1
#endif
constexpr int kEmbedBwdRows = BDLRU_EMBED_BWD_ROWS;   // measured: 2 tokens in flight made the backward slower (0.57 -> 0.68 ms)
constexpr int kEmbedFwdRows = BDLRU_EMBED_FWD_ROWS;   // tokens in flight per warp iteration (2 / 1 for the widest rows)
#define EMBED_DISPATCH_R(D, CALL)                                          \
  do {                                                                     \
    const int nv_ = (D) / 4;                                               \
    if (nv_ <= 8) { CALL(8, 1, kEmbedFwdRows); }                           \
    else if (nv_ <= 16) { CALL(16, 1, kEmbedFwdRows); }                    \
    else if (nv_ <= 32) { CALL(32, 1, kEmbedFwdRows); }                    \
    else if (nv_ <= 64) { CALL(32, 2, (kEmbedFwdRows > 2 ? 2 : kEmbedFwdRows)); } \
    else { CALL(32, 4, 1); }                                               \
  } while (0)

// grid-stride kernels: exactly the CTAs that are resident at once
template <typename K>
static int embed_grid_occ(K kernel, size_t smem, long n_tokens, int D) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int nv = D / 4;
  const int rpb = 8 * (nv <= 8 ? 4 : (nv <= 16 ? 2 : 1));
  long blocks = (n_tokens + rpb - 1) / rpb;
  const long cap = (long)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

static int embed_rpw(int D) { const int nv = D / 4; return nv <= 8 ? 4 : (nv <= 16 ? 2 : 1); }

static int embed_check(long n_tokens, long n_items, int D, int dtype, float p) {
  BDLRU_REQUIRE(n_tokens >= 1 && n_items >= 1, "embed_ln: bad sizes n_tokens=%ld n_items=%ld", n_tokens, n_items);
  BDLRU_REQUIRE(D >= 4 && D % 4 == 0 && D <= 128 * kMaxV, "embed_ln: D=%d must be a multiple of 4 and <= %d", D,
                128 * kMaxV);
  BDLRU_REQUIRE(dtype == BDLRU_F32 || dtype == BDLRU_BF16, "embed_ln: bad dtype %d", dtype);
  BDLRU_REQUIRE(p >= 0.f && p < 1.f, "embed_ln: dropout_p=%f not in [0, 1)", p);
  return BDLRU_OK;
}

}  // namespace bdlru

using namespace bdlru;

extern "C" BDLRU_API int bdlru_embed_ln_fwd(const int64_t* ids, const void* table, const float* gamma,
                                            const float* beta, void* out, float* mean, float* rstd, int64_t n_tokens,
                                            int64_t n_items, int D, float eps, float dropout_p, uint64_t seed,
                                            const uint64_t* seed_device, int dtype, int out_dtype, void* stream) {
  int rc = embed_check(n_tokens, n_items, D, dtype, dropout_p);
  if (rc) return rc;
  BDLRU_REQUIRE(out_dtype == dtype || (dtype == BDLRU_F32 && out_dtype == BDLRU_BF16),
                "embed_ln_fwd: out_dtype must equal the table dtype, or be bf16 for an fp32 table");
  BDLRU_REQUIRE(ids && table && gamma && beta && out && mean && rstd, "embed_ln_fwd: null pointer");
  BDLRU_REQUIRE(aligned(table, 16) && aligned(out, 16) && aligned(gamma, 16) && aligned(beta, 16),
                "embed_ln_fwd: pointers must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define EFWD(TT, TTO, LPR, VPL, R)                                                                             \
  {                                                                                                            \
    auto kern = embed_ln_fwd_kernel<TT, TTO, LPR, VPL, R>;                                                     \
    kern<<<embed_grid_occ(kern, 0, n_tokens, D), 256, 0, st>>>(ids, (const TT*)table, gamma, beta, (TTO*)out,  \
                                                               mean, rstd, n_tokens, n_items, D, eps,          \
                                                               dropout_p, seed, seed_device);                  \
  }
#define EFWD_FF(LPR, VPL, R) EFWD(float, float, LPR, VPL, R)
#define EFWD_BB(LPR, VPL, R) EFWD(__nv_bfloat16, __nv_bfloat16, LPR, VPL, R)
#define EFWD_FB(LPR, VPL, R) EFWD(float, __nv_bfloat16, LPR, VPL, R)
  if (dtype == BDLRU_F32 && out_dtype == BDLRU_F32) {
    EMBED_DISPATCH_R(D, EFWD_FF);
  } else if (dtype == BDLRU_BF16) {
    EMBED_DISPATCH_R(D, EFWD_BB);
  } else {
    EMBED_DISPATCH_R(D, EFWD_FB);
  }
#undef EFWD_FF
#undef EFWD_BB
#undef EFWD_FB
#undef EFWD
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

extern "C" BDLRU_API size_t bdlru_embed_ln_bwd_workspace_bytes(int64_t n_tokens, int D) {
  (void)n_tokens;
  return (size_t)sm_count() * 8 * 2 * (size_t)D * sizeof(float);
}

static int embed_bwd_impl(const int64_t* ids, const void* table, const float* gamma, const void* grad_out,
                          const float* mean, const float* rstd, float* dtable, void* drows, float* dgamma, float* dbeta,
                          void* workspace, size_t workspace_bytes, int64_t n_tokens, int64_t n_items, int D,
                          float dropout_p, uint64_t seed, const uint64_t* seed_device, int64_t padding_idx, int dtype,
                          int out_dtype, void* stream) {
  int rc = embed_check(n_tokens, n_items, D, dtype, dropout_p);
  if (rc) return rc;
  BDLRU_REQUIRE(out_dtype == dtype || (dtype == BDLRU_F32 && out_dtype == BDLRU_BF16),
                "embed_ln_bwd: out_dtype must equal the table dtype, or be bf16 for an fp32 table");
  BDLRU_REQUIRE(ids && table && gamma && grad_out && mean && rstd && (dtable || drows) && dgamma && dbeta,
                "embed_ln_bwd: null pointer");
  BDLRU_REQUIRE(aligned(table, 16) && aligned(grad_out, 16) && aligned(dtable, 16) && aligned(gamma, 16) &&
                    aligned(drows, 16),
                "embed_ln_bwd: pointers must be 16-byte aligned");
  const size_t need = bdlru_embed_ln_bwd_workspace_bytes(n_tokens, D);   // <= 8 resident 256-thread CTAs per SM
  BDLRU_REQUIRE(workspace && workspace_bytes >= need, "embed_ln_bwd: workspace too small (%zu < %zu)", workspace_bytes,
                need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* part = reinterpret_cast<float*>(workspace);
  const size_t smem = (size_t)8 * embed_rpw(D) * 2 * D * sizeof(float);
  int grid = 1;
#define EBWD(TT, TTO, LPR, VPL, R)                                                                                    \
  {                                                                                                                   \
    auto kern = embed_ln_bwd_kernel<TT, TTO, LPR, VPL, (R > kEmbedBwdRows ? kEmbedBwdRows : R)>;                                              \
    grid = embed_grid_occ(kern, smem, n_tokens, D);                                                                   \
    kern<<<grid, 256, smem, st>>>(ids, (const TT*)table, gamma, (const TTO*)grad_out, mean, rstd, dtable, part,       \
                                  n_tokens, n_items, D, dropout_p, seed, seed_device, padding_idx, (TTO*)drows);      \
  }
#define EBWD_FF(LPR, VPL, R) EBWD(float, float, LPR, VPL, R)
#define EBWD_BB(LPR, VPL, R) EBWD(__nv_bfloat16, __nv_bfloat16, LPR, VPL, R)
#define EBWD_FB(LPR, VPL, R) EBWD(float, __nv_bfloat16, LPR, VPL, R)
  if (dtype == BDLRU_F32 && out_dtype == BDLRU_F32) {
    EMBED_DISPATCH_R(D, EBWD_FF);
  } else if (dtype == BDLRU_BF16) {
    EMBED_DISPATCH_R(D, EBWD_BB);
  } else {
    EMBED_DISPATCH_R(D, EBWD_FB);
  }
#undef EBWD_FF
#undef EBWD_BB
#undef EBWD_FB
#undef EBWD
  BDLRU_LAUNCHED();
  return launch_colsum(part, grid, 2 * D, 2 * D, COLSUM_SPLIT, dgamma, dbeta, D, nullptr, st);
}

extern "C" BDLRU_API int bdlru_embed_ln_bwd(const int64_t* ids, const void* table, const float* gamma,
                                            const void* grad_out, const float* mean, const float* rstd, float* dtable,
                                            float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                                            int64_t n_tokens, int64_t n_items, int D, float dropout_p, uint64_t seed,
                                            const uint64_t* seed_device, int64_t padding_idx, int dtype, int out_dtype,
                                            void* stream) {
  BDLRU_REQUIRE(dtable, "embed_ln_bwd: null dtable");
  return embed_bwd_impl(ids, table, gamma, grad_out, mean, rstd, dtable, nullptr, dgamma, dbeta, workspace,
                        workspace_bytes, n_tokens, n_items, D, dropout_p, seed, seed_device, padding_idx, dtype, out_dtype,
                        stream);
}

extern "C" BDLRU_API int bdlru_embed_ln_bwd_rows(const int64_t* ids, const void* table, const float* gamma,
                                                 const void* grad_out, const float* mean, const float* rstd,
                                                 void* drows, float* dgamma, float* dbeta, void* workspace,
                                                 size_t workspace_bytes, int64_t n_tokens, int64_t n_items, int D,
                                                 float dropout_p, uint64_t seed, const uint64_t* seed_device,
                                                 int64_t padding_idx, int dtype, int out_dtype, void* stream) {
  BDLRU_REQUIRE(drows, "embed_ln_bwd_rows: null drows");
  return embed_bwd_impl(ids, table, gamma, grad_out, mean, rstd, nullptr, drows, dgamma, dbeta, workspace,
                        workspace_bytes, n_tokens, n_items, D, dropout_p, seed, seed_device, padding_idx, dtype, out_dtype,
                        stream);
}

extern "C" BDLRU_API int bdlru_scatter_add_rows(const int64_t* ids, const void* rows, int64_t n_tokens, int D,
                                                int rows_dtype, int64_t row_lo, int64_t row_hi, int64_t padding_idx,
                                                float* dst, void* stream) {
  BDLRU_REQUIRE(ids && rows && dst, "scatter_add_rows: null pointer");
  BDLRU_REQUIRE(n_tokens >= 1 && D >= 4 && D % 4 == 0, "scatter_add_rows: bad sizes n_tokens=%ld D=%d", (long)n_tokens, D);
  BDLRU_REQUIRE(rows_dtype == BDLRU_F32 || rows_dtype == BDLRU_BF16, "scatter_add_rows: bad dtype %d", rows_dtype);
  BDLRU_REQUIRE(row_lo >= 0 && row_hi >= row_lo, "scatter_add_rows: bad row range");
  BDLRU_REQUIRE(aligned(rows, 16) && aligned(dst, 16), "scatter_add_rows: pointers must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long total = n_tokens * (D / 4);
  long blocks = (total + 255) / 256;
  const long cap = (long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (rows_dtype == BDLRU_F32)
    scatter_rows_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(ids, (const float*)rows, n_tokens, D, row_lo, row_hi,
                                                                 padding_idx, dst);
  else
    scatter_rows_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(ids, (const __nv_bfloat16*)rows, n_tokens, D,
                                                                         row_lo, row_hi, padding_idx, dst);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}
