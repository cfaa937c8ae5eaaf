// Closed form of the state the reference's LEFT zero-pad leaves in front of the first real step (RecBLR.py:177-199,
// SURVEY §3.4): on the P phantom steps the conv emits s = silu(conv bias), so with g = W_gates s + b_gates,
//   a = exp(-softplus(Lambda) * sigmoid(g_rec)),  b' = sqrt(1 - a^2 + 1e-8) * sigmoid(g_in) * s,
//   h0 = b' * (1 - a^P) / (1 - a)            (batch independent, [C])
// and its backward (dh0 -> d conv bias, d W_gates, d b_gates, d Lambda).  As torch ops this is ~15 tiny kernels forward
// and ~30 backward PER LAYER on 128-element vectors — ~10 % of the L = 50 training step; here it is one CTA each way.
#include "common.cuh"

namespace bdlru {

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

// one CTA, blockDim.x >= 2C is not required: threads stride over rows
__global__ void __launch_bounds__(256) phantom_fwd_kernel(const float* __restrict__ cb, const float* __restrict__ W,
                                                          const float* __restrict__ gb, const float* __restrict__ lam,
                                                          int C, int P, float* __restrict__ h0,
                                                          float* __restrict__ save /* [5][C]: s, g_rec, g_in, a, q */) {
  extern __shared__ float sm[];  // s[C], g[2C]
  float* s = sm;
  float* g = sm + C;
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    const float x = cb[k];
    s[k] = x * sigm(x);
  }
  __syncthreads();
  // g[j] = W[j,:] . s + gb[j]: one warp per row (coalesced row reads, shuffle reduce)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int j = warp; j < 2 * C; j += nwarp) {
    float acc = 0.f;
    for (int k = lane; k < C; k += 32) acc = fmaf(W[(size_t)j * C + k], s[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) g[j] = acc + gb[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float cs = softplus_acc(lam[c]);
    const float sr = sigm(g[c]), si = sigm(g[C + c]);
    const float a = expf(-cs * sr);
    const float q = sqrtf(1.0f - a * a + 1e-8f);
    float G = 0.f, ak = 1.f;  // geometric sum_{k<P} a^k, summed explicitly (P <= a few thousand, a < 1)
    for (int k = 0; k < P; ++k) { G += ak; ak *= a; }
    h0[c] = q * si * s[c] * G;
    save[0 * C + c] = s[c];
    save[1 * C + c] = g[c];
    save[2 * C + c] = g[C + c];
    save[3 * C + c] = a;
    save[4 * C + c] = q;
  }
}

__global__ void __launch_bounds__(256) phantom_bwd_kernel(const float* __restrict__ cb, const float* __restrict__ W,
                                                          const float* __restrict__ lam, const float* __restrict__ save,
                                                          const float* __restrict__ dh0, int C, int P,
                                                          float* __restrict__ dcb, float* __restrict__ dW,
                                                          float* __restrict__ dgb, float* __restrict__ dlam) {
  extern __shared__ float sm[];  // s[C], dg[2C], ds_direct[C]
  float* s = sm;
  float* dg = sm + C;
  float* dsd = sm + 3 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float sc = save[c], grec = save[C + c], gin = save[2 * C + c], a = save[3 * C + c], q = save[4 * C + c];
    const float cs = softplus_acc(lam[c]);
    const float sr = sigm(grec), si = sigm(gin);
    float G = 0.f, dG = 0.f, ak = 1.f;  // G = sum_{k<P} a^k,  dG/da = sum_{k=1}^{P-1} k a^(k-1)
    for (int k = 0; k < P; ++k) {
      G += ak;
      if (k + 1 < P) dG += (float)(k + 1) * ak;
      ak *= a;
    }
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
    dlam[c] = da * (-sr * a) * sigm(lam[c]);                   // d softplus(L)/dL = sigmoid(L)
  }
  __syncthreads();
  // d b_gates = dg;  d W[j][k] = dg[j] * s[k]
  for (int j = threadIdx.x; j < 2 * C; j += blockDim.x) dgb[j] = dg[j];
  for (int idx = threadIdx.x; idx < 2 * C * C; idx += blockDim.x) dW[idx] = dg[idx / C] * s[idx % C];
  // ds[k] = sum_j dg[j] W[j][k] + direct;  d conv bias = ds * silu'(cb)
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    float acc = dsd[k];
    for (int j = 0; j < 2 * C; ++j) acc = fmaf(dg[j], W[(size_t)j * C + k], acc);
    const float x = cb[k], sg = sigm(x);
    dcb[k] = acc * sg * (1.0f + x * (1.0f - sg));
  }
}

}  // namespace bdlru

using namespace bdlru;

extern "C" BDLRU_API int bdlru_phantom_h0_fwd(const float* conv_bias, const float* gates_w, const float* gates_b,
                                              const float* Lambda, int C, int pad_len, float* h0, float* saved,
                                              void* stream) {
  BDLRU_REQUIRE(conv_bias && gates_w && gates_b && Lambda && h0 && saved, "phantom_h0_fwd: null pointer");
  BDLRU_REQUIRE(C >= 1 && C <= 2048 && pad_len >= 1 && pad_len <= (1 << 20), "phantom_h0_fwd: bad C=%d pad_len=%d", C, pad_len);
  phantom_fwd_kernel<<<1, 256, (size_t)3 * C * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
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
  phantom_bwd_kernel<<<1, 256, (size_t)4 * C * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
      conv_bias, gates_w, Lambda, saved, dh0, C, pad_len, dconv_bias, dgates_w, dgates_b, dLambda);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}
