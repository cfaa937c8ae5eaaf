// FeedForward's activation (RecBLR.py:219-221): out = dropout(silu(x)) as one elementwise kernel, and its backward
// dx = dy * mask * silu'(x) (the mask is regenerated from the same Philox counter stream: seed [+ device step counter],
// vector index).  The reference runs silu, dropout (and in backward masked_scale, silu_backward) as four ATen kernels
// over the [B*L, 4*D] activation.
#include "common.cuh"

namespace bdlru {

__device__ __forceinline__ void philox_s(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
  uint32_t c2 = 0x1234567u, c3 = 0x9abcdefu;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// keep-mask * 1/(1-p) for 8 consecutive elements from ONE Philox call (16 random bits per element)
__device__ __forceinline__ void mask8(uint64_t seed, long vec8, float p, float inv, float (&m)[8]) {
  uint32_t r[4];
  philox_s((uint32_t)vec8, (uint32_t)((uint64_t)vec8 >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const uint32_t thr = (uint32_t)(p * 65536.0f);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    m[2 * e] = (r[e] & 0xffffu) >= thr ? inv : 0.f;
    m[2 * e + 1] = (r[e] >> 16) >= thr ? inv : 0.f;
  }
}

// one thread = 8 consecutive elements per iteration (16-byte bf16 / 2 x 16-byte fp32 accesses);  BWD: out = dy * mask * silu'(x)
template <typename T, bool BWD>
__global__ void __launch_bounds__(256) silu_dropout_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                           T* __restrict__ out, long n_vec8, float p, uint64_t seed,
                                                           const uint64_t* __restrict__ seed_dev) {
  if (seed_dev) seed += *seed_dev;
  const float inv = 1.0f / (1.0f - p);
  const long stride = (long)gridDim.x * blockDim.x;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec8; v += stride) {
    float a[8], g[8], o[8], m[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    if constexpr (sizeof(T) == 2) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + v * 8);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[2 * i] = __uint_as_float(w[i] << 16);
        a[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      }
      if (BWD) {
        const uint4 d = *reinterpret_cast<const uint4*>(dy + v * 8);
        const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          g[2 * i] = __uint_as_float(dw[i] << 16);
          g[2 * i + 1] = __uint_as_float(dw[i] & 0xffff0000u);
        }
      }
    } else {
      float lo[4], hi[4];
      IO<T>::load(x + v * 8, lo);
      IO<T>::load(x + v * 8 + 4, hi);
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = lo[i], a[4 + i] = hi[i];
      if (BWD) {
        IO<T>::load(dy + v * 8, lo);
        IO<T>::load(dy + v * 8 + 4, hi);
#pragma unroll
        for (int i = 0; i < 4; ++i) g[i] = lo[i], g[4 + i] = hi[i];
      }
    }
    if (p > 0.f) mask8(seed, v, p, inv, m);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float s = sigmoid_f(a[e]);
      o[e] = BWD ? g[e] * m[e] * silu_grad_f(a[e], s) : a[e] * s * m[e];
    }
    if constexpr (sizeof(T) == 2) {
      IOV<T, 8>::store(out + v * 8, o);   // one 16-byte store
    } else {
      const float o0[4] = {o[0], o[1], o[2], o[3]}, o1[4] = {o[4], o[5], o[6], o[7]};
      IO<T>::store(out + v * 8, o0);
      IO<T>::store(out + v * 8 + 4, o1);
    }
  }
}

static int sd_launch(const void* x, const void* dy, void* out, int64_t n, float p, uint64_t seed,
                     const uint64_t* seed_dev, int dtype, bool bwd, void* stream) {
  BDLRU_REQUIRE(x && out && (!bwd || dy), "silu_dropout: null pointer");
  BDLRU_REQUIRE(n >= 8 && n % 8 == 0, "silu_dropout: n=%ld must be a positive multiple of 8", (long)n);
  BDLRU_REQUIRE(dtype == BDLRU_F32 || dtype == BDLRU_BF16, "silu_dropout: bad dtype %d", dtype);
  BDLRU_REQUIRE(p >= 0.f && p < 1.f, "silu_dropout: dropout_p=%f not in [0, 1)", p);
  const size_t al = 16;
  BDLRU_REQUIRE(aligned(x, al) && aligned(out, al) && (!bwd || aligned(dy, al)), "silu_dropout: misaligned pointer");
  const long n_vec = n / 8;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // grid-stride loop capped at two full waves of the 8 resident 256-thread CTAs per SM (32 registers)
  auto launch = [&](auto kern, auto xp, auto dyp, auto op) {
    long blocks = (n_vec + 255) / 256;
    const long cap = (long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    kern<<<(int)blocks, 256, 0, st>>>(xp, dyp, op, n_vec, p, seed, seed_dev);
  };
  if (dtype == BDLRU_F32) {
    if (bwd) launch(silu_dropout_kernel<float, true>, (const float*)x, (const float*)dy, (float*)out);
    else launch(silu_dropout_kernel<float, false>, (const float*)x, (const float*)nullptr, (float*)out);
  } else {
    if (bwd) launch(silu_dropout_kernel<__nv_bfloat16, true>, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)out);
    else launch(silu_dropout_kernel<__nv_bfloat16, false>, (const __nv_bfloat16*)x, (const __nv_bfloat16*)nullptr, (__nv_bfloat16*)out);
  }
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

}  // namespace bdlru

using namespace bdlru;

extern "C" BDLRU_API int bdlru_silu_dropout_fwd(const void* x, void* out, int64_t n, float dropout_p, uint64_t seed,
                                                const uint64_t* seed_device, int dtype, void* stream) {
  return sd_launch(x, nullptr, out, n, dropout_p, seed, seed_device, dtype, false, stream);
}

extern "C" BDLRU_API int bdlru_silu_dropout_bwd(const void* x, const void* grad_out, void* dx, int64_t n, float dropout_p,
                                                uint64_t seed, const uint64_t* seed_device, int dtype, void* stream) {
  return sd_launch(x, grad_out, dx, n, dropout_p, seed, seed_device, dtype, true, stream);
}
