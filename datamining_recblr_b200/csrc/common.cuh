// Shared device/host helpers for libbdlru.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/bdlru.h"

#define BDLRU_API __attribute__((visibility("default")))

namespace bdlru {

// ----------------------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define BDLRU_REQUIRE(cond, ...)              \
  do {                                        \
    if (!(cond)) {                            \
      ::bdlru::set_error(__VA_ARGS__);        \
      return BDLRU_ERR_INVALID;               \
    }                                         \
  } while (0)

#define BDLRU_CUDA(call)                                                                   \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      ::bdlru::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                         __LINE__);                                                        \
      return BDLRU_ERR_CUDA;                                                               \
    }                                                                                      \
  } while (0)

// call after every <<<>>> launch
#define BDLRU_LAUNCHED()                                   \
  do {                                                     \
    ::bdlru::g_launches.fetch_add(1);                      \
    BDLRU_CUDA(cudaGetLastError());                        \
  } while (0)

int sm_count();  // SMs of the current device (cached)

// Tuning switches (BDLRU_FS_DEBUG, BDLRU_FS_NT, BDLRU_CE_SETS, BDLRU_GSCAN_NS) and the clock64 phase counters exist only
// in builds made with -DBDLRU_TUNING (tools/ce_variants.py); the shipped library ignores the environment, so a timed
// kernel can never be told to skip its work.
#ifdef BDLRU_TUNING
#include <stdlib.h>
inline int tuning_env(const char* name) {
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
}
#define FS_DBG(p) ((p).dbg)
#define FS_CLOCK() clock64()
#else
inline int tuning_env(const char*) { return 0; }
#define FS_DBG(p) 0
#define FS_CLOCK() 0ll
#endif

// Deterministic second pass of every "per-CTA partials" reduction in this library:
//     acc[c] = sum_{r < n_rows} part[r * row_stride + c],  c < n_cols
// 32 columns x 32 row-lanes per CTA (coalesced 128-byte reads, fixed summation order), then an epilogue:
//   COLSUM_SPLIT   out0[c] = acc[c] for c < split, out1[c - split] = acc[c] otherwise (out1 may be NULL)
//   COLSUM_SIGMOID out0[c] = acc[c] * sigmoid(aux[c])        (dLambda: d softplus(L)/dL)
//   COLSUM_CONV    c = ch * (W+1) + j: j < W -> out0[ch * W + j], j == W -> out1[ch]   (split = W)
enum { COLSUM_SPLIT = 0, COLSUM_SIGMOID = 1, COLSUM_CONV = 2 };
int launch_colsum(const float* part, int n_rows, int row_stride, int n_cols, int mode, float* out0, float* out1,
                  int split, const float* aux, cudaStream_t st);

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ----------------------------------------------------------------------------- 4-wide I/O
template <typename T>
struct IO;

template <>
struct IO<float> {
  static constexpr int BYTES = 16;
  __device__ __forceinline__ static void load(const void* p, float (&v)[4]) {
    float4 f = *reinterpret_cast<const float4*>(p);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
  __device__ __forceinline__ static void store(void* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <>
struct IO<__nv_bfloat16> {
  static constexpr int BYTES = 8;
  __device__ __forceinline__ static void load(const void* p, float (&v)[4]) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  }
  __device__ __forceinline__ static void store(void* p, const float (&v)[4]) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// V-wide I/O: V = 4 as above, or 8 bf16 channels in one 16-byte access (the bf16 kernels are issue-bound, not
// bandwidth-bound: half as many load/store/cp.async instructions and address computations per element)
template <typename T, int V>
struct IOV;
template <>
struct IOV<float, 4> : IO<float> {};
template <>
struct IOV<__nv_bfloat16, 4> : IO<__nv_bfloat16> {};
template <>
struct IOV<__nv_bfloat16, 2> {
  static constexpr int BYTES = 4;
  __device__ __forceinline__ static void load(const void* p, float (&v)[2]) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
    v[0] = __uint_as_float(u << 16); v[1] = __uint_as_float(u & 0xffff0000u);
  }
  __device__ __forceinline__ static void store(void* p, const float (&v)[2]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(&a);
  }
};
template <>
struct IOV<__nv_bfloat16, 8> {
  static constexpr int BYTES = 16;
  __device__ __forceinline__ static void load(const void* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
    v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
  }
  __device__ __forceinline__ static void store(void* p, const float (&v)[8]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    const __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a); u.y = *reinterpret_cast<const uint32_t*>(&b);
    u.z = *reinterpret_cast<const uint32_t*>(&c); u.w = *reinterpret_cast<const uint32_t*>(&d);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

// ----------------------------------------------------------------------------- cp.async (LDGSTS)
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  if constexpr (BYTES == 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(s), "l"(gmem_src), "n"(BYTES) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ----------------------------------------------------------------------------- dropout masks
// Counter-based keep-mask for 4 consecutive elements: 16 random bits per element from three rounds of the lowbias32
// integer hash keyed by (seed, 64-bit vector index).  ~18 integer instructions per 4 elements; a Philox4x32-10 call (~60)
// made the LayerNorm kernels instruction-bound at ~50 % of HBM bandwidth.  Stateless, so the backward regenerates it.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ void keep_mask4(uint64_t seed, uint64_t vec_index, float p, float (&m)[4]) {
  const uint32_t h = mix32(mix32((uint32_t)vec_index ^ (uint32_t)seed) ^ (uint32_t)(vec_index >> 32) ^ (uint32_t)(seed >> 32));
  const uint32_t r0 = h, r1 = mix32(h ^ 0x68bc21ebu);
  const uint32_t thr = (uint32_t)(p * 65536.0f);
  const float inv = 1.0f / (1.0f - p);
  m[0] = (r0 & 0xffffu) >= thr ? inv : 0.f;
  m[1] = (r0 >> 16) >= thr ? inv : 0.f;
  m[2] = (r1 & 0xffffu) >= thr ? inv : 0.f;
  m[3] = (r1 >> 16) >= thr ? inv : 0.f;
}

// ----------------------------------------------------------------------------- math
// Single-instruction SFU approximations (MUFU.EX2 / RCP / RSQ, flush-to-zero): ~1e-7 relative error, far
// inside the 1e-4 budget, and none of the fix-up sequences the IEEE-rounded intrinsics expand to.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA/ALU pipes instead of the SFU (the softmax epilogues of the fused CE kernels are MUFU.EX2-bound at 16
// results/clk/SM; routing a fraction of the elements through this ~9-instruction sequence balances the two pipes).
// Round-to-nearest split x = n + f, f in [-0.5, 0.5] (the magic-number add leaves n in the low mantissa bits), minimax
// polynomial for 2^f (relative error 7.5e-5 at degree 3, 2.7e-6 at degree 4), n added straight into the exponent field.
// x is clamped at -125 (result ~2e-38 instead of a denormal/0); x up to +126 is fine.
template <int DEG>
__device__ __forceinline__ float ex2_poly(float x) {
  static_assert(DEG == 3 || DEG == 4, "degree");
  x = fmaxf(x, -125.f);
  const float xf = x + 12582912.f;
  const float f = x - (xf - 12582912.f);
  float p;
  if constexpr (DEG == 3) {
    p = fmaf(0.05517132207751274f, f, 0.24261054396629333f);
    p = fmaf(p, f, 0.6932609677314758f);
    p = fmaf(p, f, 0.9999281167984009f);
  } else {
    p = fmaf(0.009570068679749966f, f, 0.055917806923389435f);
    p = fmaf(p, f, 0.240247443318367f);
    p = fmaf(p, f, 0.6931218504905701f);
    p = fmaf(p, f, 0.9999992847442627f);
  }
  return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));
}
// ex2 of element `i` of an unrolled 32-wide chunk: SFU, or the polynomial where bit i of MASK is set (compile time)
template <uint32_t MASK, int DEG>
__device__ __forceinline__ float ex2_mixed(float x, int i) {
  return ((MASK >> (i & 31)) & 1u) ? ex2_poly<DEG>(x) : ex2_ftz(x);
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_ftz(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float exp_f(float x) { return ex2_ftz(x * 1.4426950408889634f); }
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_ftz(1.0f + ex2_ftz(x * -1.4426950408889634f)); }
// One-MUFU sigmoid (tanh.approx, abs error ~5e-4): used by the bf16-I/O instantiations only, where the result is rounded
// to 8 bits anyway and the kernels are SFU-bound rather than HBM-bound (6-8 SFU ops per element with the exact form).
template <bool FAST>
__device__ __forceinline__ float sigmoid_t(float x) {
  if constexpr (FAST) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
  } else {
    return sigmoid_f(x);
  }
}
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }
// d/dx silu(x) given s = sigmoid(x)
__device__ __forceinline__ float silu_grad_f(float x, float s) { return s * (1.0f + x * (1.0f - s)); }
__device__ __forceinline__ float softplus_acc(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

// 1 - exp(-u) for u >= 0 without cancellation; e = exp(-u) already computed.  Below 1/16 the 4-term series
// (truncation u^4/120 < 1.3e-7 relative); above it 1 - e loses at most eps/0.06 ~ 1e-6 relative.
__device__ __forceinline__ float one_minus_exp_neg(float u, float e) {
  float p = fmaf(u, -1.0f / 24.0f, 1.0f / 6.0f);
  p = fmaf(u, p, -0.5f);
  p = fmaf(u, p, 1.0f);
  return u < 0.0625f ? u * p : 1.0f - e;
}

// The BD-LRU gates (RecBLR.py:197-198) for one element.  c = softplus(Lambda).
struct Gate {
  float sr, si, a, rq, q;  // sigmoid(r), sigmoid(i), alpha, 1/sqrt(1-a^2+1e-8), sqrt(1-a^2+1e-8)
};
template <bool FAST = false>
__device__ __forceinline__ float gate_alpha(float c, float r, float& sr) {
  sr = sigmoid_t<FAST>(r);
  return exp_f(-c * sr);
}
// 1 - a^2 + 1e-8 with a = exp(-c * sr).  Exact path: cancellation-free series for small exponents.  FAST (bf16 I/O, whose
// sigmoid is already a 5e-4 approximation and whose outputs are rounded to 2^-9): FMA + add; its absolute error ~6e-8 is
// below the bf16 rounding of the result for every a <= 1 - 3e-5.  The 1e-8 is added AFTERWARDS (1 + 1e-8 == 1 in fp32):
// a saturated gate (a == 1) must give sqrt(1e-8), not rsqrt(0) = inf.
template <bool FAST>
__device__ __forceinline__ float one_minus_a2(float c, float sr, float a) {
  if constexpr (FAST) return fmaf(-a, a, 1.0f) + 1e-8f;
  else return one_minus_exp_neg(2.0f * c * sr, a * a) + 1e-8f;
}
template <bool FAST = false>
__device__ __forceinline__ Gate gate_full(float c, float r, float i) {
  Gate g;
  g.a = gate_alpha<FAST>(c, r, g.sr);
  g.si = sigmoid_t<FAST>(i);
  float v = one_minus_a2<FAST>(c, g.sr, g.a);
  g.rq = rsqrt_ftz(v);
  g.q = v * g.rq;
  return g;
}

}  // namespace bdlru
