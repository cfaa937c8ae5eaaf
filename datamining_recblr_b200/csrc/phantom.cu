// Closed form of the state the reference's LEFT zero-pad leaves in front of the first real step (RecBLR.py:177-199,
// SURVEY §3.4): on the P phantom steps the conv emits s = silu(conv bias), so with g = W_gates s + b_gates,
//   a = exp(-softplus(Lambda) * sigmoid(g_rec)),  b' = sqrt(1 - a^2 + 1e-8) * sigmoid(g_in) * s,
//   h0 = b' * (1 - a^P) / (1 - a)            (batch independent, [C])
// and its backward (dh0 -> d conv bias, d W_gates, d b_gates, d Lambda).  As torch ops this is ~15 tiny kernels forward
// and ~30 backward PER LAYER on 128-element vectors — ~120 launches per L = 50 training step; here it is one small kernel each way.
#include "common.cuh"

namespace bdlru {

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

// G = sum_{k<P} a^k and dG/da = sum_{k=1}^{P-1} k a^(k-1) for a = exp(-u), u > 0.  Explicit sums for short pads (exact
// to rounding), closed forms built from expm1 for long ones (1 - a and 1 - a^P without cancellation).
__device__ __forceinline__ void geo_sums(float a, float u, int P, float& G, float& dG) {
  if (P <= 256) {
    float g = 0.f, d = 0.f, ak = 1.f;
    for (int k = 0; k < P; ++k) {
      g += ak;
      if (k + 1 < P) d += (float)(k + 1) * ak;
      ak *= a;
    }
    G = g; dG = d;
  } else {
    const float om = -expm1f(-u);                 // 1 - a
    const float omP = -expm1f(-(float)P * u);     // 1 - a^P
    const float aPm1 = expf(-(float)(P - 1) * u); // a^(P-1)
    G = omP / om;
    dG = (omP - (float)P * aPm1 * om) / (om * om);
  }
}

constexpr int kPhCh = 4;  // channels per CTA in the forward (two gate rows each -> 8 warps)

// Forward: CTA b handles channels [4b, 4b+4); warp w computes the dot product of gate row (w odd: input, w even:
// recurrence) of channel 4b + w/2 with s = silu(conv bias) (coalesced row read, shuffle reduce).
__global__ void __launch_bounds__(256) phantom_fwd_kernel(const float* __restrict__ cb, const float* __restrict__ W,
                                                          const float* __restrict__ gb, const float* __restrict__ lam,
                                                          int C, int P, float* __restrict__ h0,
                                                          float* __restrict__ save /* [5][C]: s, g_rec, g_in, a, q */) {
  __shared__ float gsh[2 * kPhCh];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * kPhCh + (warp >> 1);
  if (c < C) {
    const int j = (warp & 1) ? C + c : c;
    float acc = 0.f;
    for (int k = lane; k < C; k += 32) {
      const float x = cb[k];
      acc = fmaf(W[(size_t)j * C + k], x * sigm(x), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) gsh[warp] = acc + gb[j];
  }
  __syncthreads();
  if (threadIdx.x < kPhCh) {
    const int cc = blockIdx.x * kPhCh + threadIdx.x;
    if (cc < C) {
      const float grec = gsh[2 * threadIdx.x], gin = gsh[2 * threadIdx.x + 1];
      const float x = cb[cc], sc = x * sigm(x);
      const float cs = softplus_acc(lam[cc]);
      const float sr = sigm(grec), si = sigm(gin);
      const float u = cs * sr, a = expf(-u);
      const float q = sqrtf(1.0f - a * a + 1e-8f);
      float G, dG;
      geo_sums(a, u, P, G, dG);
      h0[cc] = q * si * sc * G;
      save[0 * C + cc] = sc;
      save[1 * C + cc] = grec;
      save[2 * C + cc] = gin;
      save[3 * C + cc] = a;
      save[4 * C + cc] = q;
    }
  }
}

// Backward.  Every CTA first rebuilds the per-channel scalars (dg [2C], s [C], direct ds [C]) in shared memory, then
//   CTAs [0, nb_rows): 8 gate rows each -> dW[j][:] = dg[j] * s (coalesced), d b_gates[j] = dg[j]; CTA 0 also dLambda;
//   CTAs [nb_rows, ..): 32 channels each -> ds[k] = sum_j dg[j] W[j][k] (8 warps split the 2C rows, fixed-order smem
//                       reduction) -> d conv bias[k] = (ds[k] + direct) * silu'(conv bias[k]).
__global__ void __launch_bounds__(256) phantom_bwd_kernel(const float* __restrict__ cb, const float* __restrict__ W,
                                                          const float* __restrict__ lam, const float* __restrict__ save,
                                                          const float* __restrict__ dh0, int C, int P, int nb_rows,
                                                          float* __restrict__ dcb, float* __restrict__ dW,
                                                          float* __restrict__ dgb, float* __restrict__ dlam) {
  extern __shared__ float sm[];  // s[C], dg[2C], ds_direct[C], red[8][32]
  float* s = sm;
  float* dg = sm + C;
  float* dsd = sm + 3 * C;
  float* red = sm + 4 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float sc = save[c], grec = save[C + c], gin = save[2 * C + c], a = save[3 * C + c], q = save[4 * C + c];
    const float cs = softplus_acc(lam[c]);
    const float sr = sigm(grec), si = sigm(gin);
    float G, dG;
    geo_sums(a, cs * sr, P, G, dG);
    const float d = dh0[c];
    const float bp = q * si * sc;
    const float dbp = d * G;
    const float da = d * bp * dG + dbp * si * sc * (-a / q);   // through G and through q
    const float dsi = dbp * q * sc;
    const float dsr = da * (-cs * a);
    s[c] = sc;
    dsd[c] = dbp * q * si;                                     // direct dependence of b' on s_c
    dg[c] = dsr * sr * (1.0f - sr);
    dg[C + c] = dsi * si * (1.0f - si);
    if (blockIdx.x == 0) dlam[c] = da * (-sr * a) * sigm(lam[c]);  // d softplus(L)/dL = sigmoid(L)
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((int)blockIdx.x < nb_rows) {
    const int j = blockIdx.x * 8 + warp;
    if (j < 2 * C) {
      const float g = dg[j];
      for (int k = lane; k < C; k += 32) dW[(size_t)j * C + k] = g * s[k];
      if (lane == 0) dgb[j] = g;
    }
  } else {
    const int k = (blockIdx.x - nb_rows) * 32 + lane;
    float acc = 0.f;
    if (k < C) {
      int j = warp;
      for (; j + 24 < 2 * C; j += 32) {  // 4 independent loads in flight
        const float w0 = W[(size_t)j * C + k], w1 = W[(size_t)(j + 8) * C + k];
        const float w2 = W[(size_t)(j + 16) * C + k], w3 = W[(size_t)(j + 24) * C + k];
        acc = fmaf(dg[j], w0, acc);
        acc = fmaf(dg[j + 8], w1, acc);
        acc = fmaf(dg[j + 16], w2, acc);
        acc = fmaf(dg[j + 24], w3, acc);
      }
      for (; j < 2 * C; j += 8) acc = fmaf(dg[j], W[(size_t)j * C + k], acc);
    }
    red[warp * 32 + lane] = acc;
    __syncthreads();
    if (warp == 0 && k < C) {
      float t = dsd[k];
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w * 32 + lane];
      const float x = cb[k], sg = sigm(x);
      dcb[k] = t * sg * (1.0f + x * (1.0f - sg));
    }
  }
}

}  // namespace bdlru

using namespace bdlru;

extern "C" BDLRU_API int bdlru_phantom_h0_fwd(const float* conv_bias, const float* gates_w, const float* gates_b,
                                              const float* Lambda, int C, int pad_len, float* h0, float* saved,
                                              void* stream) {
  BDLRU_REQUIRE(conv_bias && gates_w && gates_b && Lambda && h0 && saved, "phantom_h0_fwd: null pointer");
  BDLRU_REQUIRE(C >= 1 && C <= 2048 && pad_len >= 1 && pad_len <= (1 << 20), "phantom_h0_fwd: bad C=%d pad_len=%d", C, pad_len);
  phantom_fwd_kernel<<<(C + kPhCh - 1) / kPhCh, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      conv_bias, gates_w, gates_b, Lambda, C, pad_len, h0, saved);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

extern "C" BDLRU_API int bdlru_phantom_h0_bwd(const float* conv_bias, const float* gates_w, const float* Lambda,
                                              const float* saved, const float* dh0, int C, int pad_len, float* dconv_bias,
                                              float* dgates_w, float* dgates_b, float* dLambda, void* stream) {
  BDLRU_REQUIRE(conv_bias && gates_w && Lambda && saved && dh0 && dconv_bias && dgates_w && dgates_b && dLambda,
                "phantom_h0_bwd: null pointer");
  BDLRU_REQUIRE(C >= 1 && C <= 2048 && pad_len >= 1, "phantom_h0_bwd: bad C=%d pad_len=%d", C, pad_len);
  const int nb_rows = (2 * C + 7) / 8, nb_cols = (C + 31) / 32;
  phantom_bwd_kernel<<<nb_rows + nb_cols, 256, (size_t)(4 * C + 256) * sizeof(float),
                       reinterpret_cast<cudaStream_t>(stream)>>>(conv_bias, gates_w, Lambda, saved, dh0, C, pad_len,
                                                                   nb_rows, dconv_bias, dgates_w, dgates_b, dLambda);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}
