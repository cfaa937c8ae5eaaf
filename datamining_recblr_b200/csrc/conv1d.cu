// Causal depthwise conv1d (+bias, +SiLU) on channel-last [B, T, C] views, forward and backward.
// Replaces the causal_conv1d_fn call of RecBLR.py:188-193 (semantics pinned by the fallback, line 185).
//
// Threads form [ny, tcn]: tcn threads cover a channel tile with one 4-channel vector each (coalesced
// along C), each thread takes S consecutive time steps of one batch row per work item and keeps the
// W-1 halo in registers, so every input row is fetched from HBM once (halo re-reads hit L1/L2).
// The grid is (n_ctile, GY) with a grid-stride loop over (batch, time chunk) work items, which lets the
// backward accumulate dweight/dbias in registers across its whole loop and finish with one block
// reduction per CTA plus a tiny deterministic second pass (no atomics).
// HBM traffic: fwd reads x, writes y (2 units); bwd reads x, dy, writes dx (3 units; SURVEY §8d counts 6
// for fwd+bwd because the reference kernel also re-reads its saved pre-activation).
#include "common.cuh"

namespace bdlru {

struct CView {
  const unsigned char* p;
  long bs, rs;
};
template <typename T>
__device__ __forceinline__ const unsigned char* cat(const CView& v, long b, long t, int c) {
  return v.p + ((b * v.bs + t * v.rs + c) * (long)sizeof(T));
}

struct ConvParams {
  CView x, y, dy, dx;
  const float* w;     // [C, W]
  const float* bias;  // [C] or null
  float* part;        // bwd: [GY, C, W+1]
  int B, T, C, tcn, n_chunk;
  long n_items;  // B * n_chunk
};

#ifndef BDLRU_CONV_FWD_MINB
#define BDLRU_CONV_FWD_MINB 2
#endif
// V channels per thread: 4, or 8 for bf16 rows whose channel count, strides and base addresses allow 16-byte accesses
// (at 8 bytes per lane the kernel is bound by the number of load / store instructions, not by bytes).
template <typename T, int W, int S, int V, bool SILU>
__global__ void __launch_bounds__(256, BDLRU_CONV_FWD_MINB) conv_fwd_kernel(const ConvParams p) {
  const int tc = threadIdx.x, ty = threadIdx.y;
  const int c = (blockIdx.x * p.tcn + tc) * V;
  float wv[V][W], bv[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
#pragma unroll
    for (int j = 0; j < W; ++j) wv[e][j] = p.w[(c + e) * W + j];
    bv[e] = p.bias ? p.bias[c + e] : 0.f;
  }
  for (long item = (long)blockIdx.y * blockDim.y + ty; item < p.n_items; item += (long)gridDim.y * blockDim.y) {
    const long b = item / p.n_chunk;
    const int t0 = (int)(item % p.n_chunk) * S;
    float xs[S + W - 1][V];
#pragma unroll
    for (int s = 0; s < S + W - 1; ++s) {
      const int t = t0 + s - (W - 1);
      if (t >= 0 && t < p.T) {
        IOV<T, V>::load(cat<T>(p.x, b, t, c), xs[s]);
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) xs[s][e] = 0.f;
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (t0 + s < p.T) {
        float o[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
          float pre = bv[e];
#pragma unroll
          for (int j = 0; j < W; ++j) pre = fmaf(wv[e][j], xs[s + j][e], pre);
          o[e] = SILU ? silu_f(pre) : pre;
        }
        IOV<T, V>::store(const_cast<unsigned char*>(cat<T>(p.y, b, t0 + s, c)), o);
      }
    }
  }
}

// Backward.  A work item is (batch row, run of RUN consecutive time steps); the thread walks the run from its
// last chunk of S steps down to its first, carrying in registers the W-1 lowest x rows and the W-1 lowest
// dpre rows of the chunk above, so that each x / dy row is loaded once and each activation derivative is
// computed once (only the W-1 rows above a run are recomputed).
template <typename T, int W, int S, int RUN, bool SILU>
__global__ void __launch_bounds__(256, 2) conv_bwd_kernel(const ConvParams p) {
  extern __shared__ float red[];  // [ny][tcn][4*(W+1)]
  constexpr int H = W - 1;
  const int tc = threadIdx.x, ty = threadIdx.y;
  const int c = (blockIdx.x * p.tcn + tc) * 4;
  float wv[4][W], bv[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
#pragma unroll
    for (int j = 0; j < W; ++j) wv[e][j] = p.w[(c + e) * W + j];
    bv[e] = p.bias ? p.bias[c + e] : 0.f;
  }
  float dw[4][W], db[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    db[e] = 0.f;
#pragma unroll
    for (int j = 0; j < W; ++j) dw[e][j] = 0.f;
  }
  const long xrb = p.x.rs * (long)sizeof(T), dyrb = p.dy.rs * (long)sizeof(T), dxrb = p.dx.rs * (long)sizeof(T);

  auto ld_row = [&](const unsigned char* base, long rb, int t, float (&v)[4]) {
    if (t >= 0 && t < p.T) {
      IO<T>::load(base + t * rb, v);
    } else {
      v[0] = v[1] = v[2] = v[3] = 0.f;
    }
  };
  // dpre_t = dy_t * act'(pre_t) with pre_t = bias + sum_j w_j x_{t-H+j}; xw[j] = x_{t-H+j}
  auto dpre_of = [&](const float (&dyv)[4], const float (*xw)[4], float (&out)[4]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (SILU) {
        float pre = bv[e];
#pragma unroll
        for (int j = 0; j < W; ++j) pre = fmaf(wv[e][j], xw[j][e], pre);
        const float sg = sigmoid_f(pre);
        out[e] = dyv[e] * silu_grad_f(pre, sg);
      } else {
        out[e] = dyv[e];
      }
    }
  };

  for (long item = (long)blockIdx.y * blockDim.y + ty; item < p.n_items; item += (long)gridDim.y * blockDim.y) {
    const long b = item / p.n_chunk;
    const int t_lo = (int)(item % p.n_chunk) * RUN;
    const int t_hi = min(p.T, t_lo + RUN);
    const unsigned char* xb = cat<T>(p.x, b, 0, c);
    const unsigned char* dyb = cat<T>(p.dy, b, 0, c);
    unsigned char* dxb = const_cast<unsigned char*>(cat<T>(p.dx, b, 0, c));
    const int n_ch = (t_hi - t_lo + S - 1) / S;
    int t0 = t_lo + (n_ch - 1) * S;  // top chunk [t0, t0+S)
    // rows carried from "above": xs[S .. S+2H-1] = x_{t0+S-H .. t0+S+H-1}, dp[S .. S+H-1] = dpre_{t0+S .. t0+S+H-1}
    float xs[S + 2 * H][4], dp[S + H][4];
#pragma unroll
    for (int j = 0; j < 2 * H; ++j) ld_row(xb, xrb, t0 + S - H + j, xs[S + j]);
#pragma unroll
    for (int j = 0; j < H; ++j) {
      float dyv[4];
      ld_row(dyb, dyrb, t0 + S + j, dyv);  // zero beyond T => dpre = 0
      dpre_of(dyv, &xs[S + j], dp[S + j]);
    }
    for (int ch = 0; ch < n_ch; ++ch, t0 -= S) {
      // new rows of this chunk: xs[0 .. S-1] = x_{t0-H .. t0+S-H-1}, dy rows t0 .. t0+S-1
      float dyv[S][4];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        ld_row(xb, xrb, t0 - H + s, xs[s]);
        ld_row(dyb, dyrb, t0 + s, dyv[s]);
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        dpre_of(dyv[s], &xs[s], dp[s]);  // window of x_{t-H .. t} for t = t0+s is xs[s .. s+H]
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          db[e] += dp[s][e];
#pragma unroll
          for (int j = 0; j < W; ++j) dw[e][j] = fmaf(dp[s][e], xs[s + j][e], dw[e][j]);
        }
      }
      // dx_t = sum_j w_j * dpre_{t + H - j}
#pragma unroll
      for (int s = 0; s < S; ++s) {
        if (t0 + s < p.T) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < W; ++j) acc = fmaf(wv[e][j], dp[s + H - j][e], acc);
            o[e] = acc;
          }
          IO<T>::store(dxb + (long)(t0 + s) * dxrb, o);
        }
      }
      // hand the lowest rows down to the next (earlier) chunk
#pragma unroll
      for (int j = 0; j < 2 * H; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) xs[S + j][e] = xs[j][e];
#pragma unroll
      for (int j = 0; j < H; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) dp[S + j][e] = dp[j][e];
    }
  }
  // block reduction over ty, then one partial row per CTA
  constexpr int NV = 4 * (W + 1);
  float* mine = red + ((size_t)ty * p.tcn + tc) * NV;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
#pragma unroll
    for (int j = 0; j < W; ++j) mine[e * (W + 1) + j] = dw[e][j];
    mine[e * (W + 1) + W] = db[e];
  }
  __syncthreads();
  if (ty == 0) {
    for (int v = 0; v < NV; ++v) {
      float s = 0.f;
      for (int y = 0; y < (int)blockDim.y; ++y) s += red[((size_t)y * p.tcn + tc) * NV + v];
      // part[gy][c + e][j]  with v = e*(W+1) + j
      p.part[((size_t)blockIdx.y * p.C + c) * (W + 1) + v] = s;
    }
  }
}

constexpr int kConvS = 8;     // steps per thread, forward
constexpr int kConvSB = 4;    // steps per chunk, backward
constexpr int kConvRun = 32;  // time steps per backward work item (walked chunk by chunk with carried halos)

struct ConvTiling {
  int tcn, ny, n_ctile, n_chunk, gy;
};
static ConvTiling conv_tiling(int B, int T, int C, int ctas_per_sm, int S, int V = 4) {
  ConvTiling t;
  const int cvec = C / V;
  t.tcn = 1;
  while (t.tcn < 32 && cvec % (t.tcn * 2) == 0) t.tcn *= 2;
  t.n_ctile = cvec / t.tcn;
  t.ny = 256 / t.tcn;
  if (t.ny > 32) t.ny = 32;
  t.n_chunk = (T + S - 1) / S;
  const long items = (long)B * t.n_chunk;
  long gy = ((long)sm_count() * ctas_per_sm + t.n_ctile - 1) / t.n_ctile;
  const long need = (items + t.ny - 1) / t.ny;
  if (gy > need) gy = need;
  if (gy < 1) gy = 1;
  if (gy > 65535) gy = 65535;
  t.gy = (int)gy;
  return t;
}

static int conv_check(const char* name, const bdlru_view& v, int es) {
  BDLRU_REQUIRE(v.ptr, "%s: null pointer", name);
  BDLRU_REQUIRE(aligned(v.ptr, 4 * es), "%s: pointer not aligned to %d bytes", name, 4 * es);
  BDLRU_REQUIRE(v.bstride % 4 == 0 && v.rstride % 4 == 0, "%s: strides must be multiples of 4 elements", name);
  return BDLRU_OK;
}
static CView cmk(const bdlru_view& v) { return CView{reinterpret_cast<const unsigned char*>(v.ptr), v.bstride, v.rstride}; }

// Measured at 8192 x 200 x 256 bf16 (tools/conv_bench.py, ms): 8 channels x S steps per thread, grid = resident CTAs
//   S=1 0.91   S=2 0.50   S=3 0.79   S=4 0.72-0.76   S=8 0.74      (4 channels x 8 steps, the previous kernel: 0.71-0.73)
#ifndef BDLRU_CONV_FWD_S8
#define BDLRU_CONV_FWD_S8 2
#endif
#ifndef BDLRU_CONV_FWD_CTAS
#define BDLRU_CONV_FWD_CTAS 2
#endif
constexpr int kConvS8 = BDLRU_CONV_FWD_S8;   // steps per thread of the 8-channel forward

template <typename T, int W, int S, int V>
static int conv_fwd_launch_v(ConvParams& p, bool silu, cudaStream_t st) {
  ConvTiling t = conv_tiling(p.B, p.T, p.C, BDLRU_CONV_FWD_CTAS, S, V);
  p.tcn = t.tcn; p.n_chunk = t.n_chunk; p.n_items = (long)p.B * t.n_chunk;
  dim3 grid(t.n_ctile, t.gy), block(t.tcn, t.ny);
  if (silu)
    conv_fwd_kernel<T, W, S, V, true><<<grid, block, 0, st>>>(p);
  else
    conv_fwd_kernel<T, W, S, V, false><<<grid, block, 0, st>>>(p);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

static bool conv_wide_ok(const CView& v) {
  return aligned(v.p, 16) && v.bs % 8 == 0 && v.rs % 8 == 0;
}

template <typename T, int W>
static int conv_fwd_launch(ConvParams& p, bool silu, cudaStream_t st) {
  if constexpr (sizeof(T) == 2) {
    if (p.C % 8 == 0 && conv_wide_ok(p.x) && conv_wide_ok(p.y)) return conv_fwd_launch_v<T, W, kConvS8, 8>(p, silu, st);
  }
  return conv_fwd_launch_v<T, W, kConvS, 4>(p, silu, st);
}

template <typename T, int W>
static int conv_bwd_launch(ConvParams& p, bool silu, float* dweight, float* dbias, void* ws, size_t ws_bytes,
                           cudaStream_t st) {
  ConvTiling t = conv_tiling(p.B, p.T, p.C, 4, kConvRun);
  p.tcn = t.tcn; p.n_chunk = t.n_chunk; p.n_items = (long)p.B * t.n_chunk;
  const size_t need = (size_t)t.gy * p.C * (W + 1) * sizeof(float);
  BDLRU_REQUIRE(ws && ws_bytes >= need, "conv1d_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
  p.part = reinterpret_cast<float*>(ws);
  dim3 grid(t.n_ctile, t.gy), block(t.tcn, t.ny);
  const size_t smem = (size_t)t.ny * t.tcn * 4 * (W + 1) * sizeof(float);
  if (silu)
    conv_bwd_kernel<T, W, kConvSB, kConvRun, true><<<grid, block, smem, st>>>(p);
  else
    conv_bwd_kernel<T, W, kConvSB, kConvRun, false><<<grid, block, smem, st>>>(p);
  BDLRU_LAUNCHED();
  const int n = p.C * (W + 1);
  return launch_colsum(p.part, t.gy, n, n, COLSUM_CONV, dweight, dbias, W, nullptr, st);
}

}  // namespace bdlru

using namespace bdlru;

#define CONV_CHECK(name, v, es)          \
  do {                                   \
    int rc_ = conv_check(name, v, es);   \
    if (rc_) return rc_;                 \
  } while (0)

#define CONV_DISPATCH(FN, ...)                                                       \
  do {                                                                               \
    if (dtype == BDLRU_F32) {                                                        \
      switch (W) {                                                                   \
        case 1: return FN<float, 1>(__VA_ARGS__);                                    \
        case 2: return FN<float, 2>(__VA_ARGS__);                                    \
        case 3: return FN<float, 3>(__VA_ARGS__);                                    \
        default: return FN<float, 4>(__VA_ARGS__);                                   \
      }                                                                              \
    } else {                                                                         \
      switch (W) {                                                                   \
        case 1: return FN<__nv_bfloat16, 1>(__VA_ARGS__);                            \
        case 2: return FN<__nv_bfloat16, 2>(__VA_ARGS__);                            \
        case 3: return FN<__nv_bfloat16, 3>(__VA_ARGS__);                            \
        default: return FN<__nv_bfloat16, 4>(__VA_ARGS__);                           \
      }                                                                              \
    }                                                                                \
  } while (0)

static int conv_common(int B, int T, int C, int W, int dtype) {
  BDLRU_REQUIRE(B >= 1 && T >= 1 && C >= 4 && C % 4 == 0, "conv1d: bad shape B=%d T=%d C=%d (C %% 4 == 0)", B, T, C);
  BDLRU_REQUIRE(W >= 1 && W <= 4, "conv1d: kernel width %d not in [1, 4]", W);
  BDLRU_REQUIRE(dtype == BDLRU_F32 || dtype == BDLRU_BF16, "conv1d: bad dtype %d", dtype);
  return BDLRU_OK;
}

extern "C" BDLRU_API int bdlru_conv1d_fwd(bdlru_view x, const float* weight, const float* bias, bdlru_view y, int B, int T,
                                int C, int W, int silu, int dtype, void* stream) {
  int rc = conv_common(B, T, C, W, dtype);
  if (rc) return rc;
  const int es = dtype == BDLRU_F32 ? 4 : 2;
  CONV_CHECK("x", x, es); CONV_CHECK("y", y, es);
  BDLRU_REQUIRE(weight, "conv1d: weight is null");
  ConvParams p{};
  p.x = cmk(x); p.y = cmk(y); p.w = weight; p.bias = bias; p.B = B; p.T = T; p.C = C;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CONV_DISPATCH(conv_fwd_launch, p, silu != 0, st);
}

extern "C" BDLRU_API size_t bdlru_conv1d_bwd_workspace_bytes(int B, int T, int C, int W) {
  (void)B; (void)T;
  // gy <= SMs * 4 / n_ctile + 1 partial rows of C*(W+1) floats
  return ((size_t)sm_count() * 4 + 2) * (size_t)C * (W + 1) * sizeof(float);
}

extern "C" BDLRU_API int bdlru_conv1d_bwd(bdlru_view x, const float* weight, const float* bias, bdlru_view grad_y, bdlru_view dx,
                                float* dweight, float* dbias, void* workspace, size_t workspace_bytes, int B, int T,
                                int C, int W, int silu, int dtype, void* stream) {
  int rc = conv_common(B, T, C, W, dtype);
  if (rc) return rc;
  const int es = dtype == BDLRU_F32 ? 4 : 2;
  CONV_CHECK("x", x, es); CONV_CHECK("grad_y", grad_y, es); CONV_CHECK("dx", dx, es);
  BDLRU_REQUIRE(weight && dweight, "conv1d_bwd: weight/dweight is null");
  ConvParams p{};
  p.x = cmk(x); p.dy = cmk(grad_y); p.dx = cmk(dx); p.w = weight; p.bias = bias; p.B = B; p.T = T; p.C = C;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CONV_DISPATCH(conv_bwd_launch, p, silu != 0, dweight, dbias, workspace, workspace_bytes, st);
}
