// S0: raw first-order scan on [rows = B*C, T] contiguous fp32 (T contiguous) — the exact operator
// boundary of the reference's parallel_scan(gates, tokens) (parallel_scan.py:85-95, 98-114).
//
// One warp walks one row in chunks of 32*VEC steps: each lane scans VEC consecutive steps in
// registers (128-bit loads when T % 4 == 0), the 32 chunk aggregates (A, H) are combined with a
// Kogge-Stone warp-shuffle scan, and the row carry moves from chunk to chunk in a register.  Any
// T >= 1 is accepted (no power-of-two padding).  The next chunk's loads are issued before the current
// chunk is scanned.  Backward = the same thing in reverse time with u_t = a_t * dh~_t carried.
// HBM traffic: fwd reads a, b, writes h (3 units); bwd reads a, h, g, writes da, db (5 units).
#include "common.cuh"

namespace bdlru {

template <int VEC>
__device__ __forceinline__ void ldv(const float* base, long t, long T, float fill, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    if (t < T) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(base + t));
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
      v[0] = v[1] = v[2] = v[3] = fill;
    }
  } else {
    v[0] = t < T ? __ldg(base + t) : fill;
  }
}
template <int VEC>
__device__ __forceinline__ void stv(float* base, long t, long T, const float (&v)[VEC]) {
  if (t >= T) return;
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(base + t) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    base[t] = v[0];
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) scan_bct_fwd_kernel(const float* __restrict__ gates,
                                                           const float* __restrict__ tokens,
                                                           float* __restrict__ states, long rows, long T) {
  const int lane = threadIdx.x & 31;
  const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  constexpr int CH = 32 * VEC;
  for (long row = gw; row < rows; row += nw) {
    const float* ar = gates + row * T;
    const float* br = tokens + row * T;
    float* hr = states + row * T;
    float carry = 0.f;
    float av[VEC], bv[VEC];
    ldv<VEC>(ar, (long)lane * VEC, T, 1.f, av);
    ldv<VEC>(br, (long)lane * VEC, T, 0.f, bv);
    for (long t0 = 0; t0 < T; t0 += CH) {
      const long t = t0 + (long)lane * VEC;
      float an[VEC] = {}, bn[VEC] = {};
      if (t0 + CH < T) {  // prefetch the next chunk (warp-uniform branch)
        ldv<VEC>(ar, t + CH, T, 1.f, an);
        ldv<VEC>(br, t + CH, T, 0.f, bn);
      }
      float A = 1.f, H = 0.f, hl[VEC], cl[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        H = fmaf(av[e], H, bv[e]);
        A *= av[e];
        hl[e] = H;
        cl[e] = A;
      }
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float Ap = __shfl_up_sync(0xffffffffu, A, d);
        const float Hp = __shfl_up_sync(0xffffffffu, H, d);
        if (lane >= d) {
          H = fmaf(A, Hp, H);
          A *= Ap;
        }
      }
      float Aex = __shfl_up_sync(0xffffffffu, A, 1);
      float Hex = __shfl_up_sync(0xffffffffu, H, 1);
      if (lane == 0) { Aex = 1.f; Hex = 0.f; }
      const float cin = fmaf(Aex, carry, Hex);
      float out[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) out[e] = fmaf(cl[e], cin, hl[e]);
      stv<VEC>(hr, t, T, out);
      carry = fmaf(__shfl_sync(0xffffffffu, A, 31), carry, __shfl_sync(0xffffffffu, H, 31));
#pragma unroll
      for (int e = 0; e < VEC; ++e) { av[e] = an[e]; bv[e] = bn[e]; }
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) scan_bct_bwd_kernel(const float* __restrict__ gates,
                                                           const float* __restrict__ states,
                                                           const float* __restrict__ grad,
                                                           float* __restrict__ d_gates,
                                                           float* __restrict__ d_tokens, long rows, long T) {
  const int lane = threadIdx.x & 31;
  const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  constexpr int CH = 32 * VEC;
  const long nchunk = (T + CH - 1) / CH;
  for (long row = gw; row < rows; row += nw) {
    const float* ar = gates + row * T;
    const float* hr = states + row * T;
    const float* gr = grad + row * T;
    float carry = 0.f;  // u entering the current chunk from later time steps
    float av[VEC], hv[VEC], gv[VEC], hb = 0.f;
    {
      const long t = (nchunk - 1) * CH + (long)lane * VEC;
      ldv<VEC>(ar, t, T, 1.f, av);
      ldv<VEC>(hr, t, T, 0.f, hv);
      ldv<VEC>(gr, t, T, 0.f, gv);
      if (lane == 0 && nchunk > 1) hb = __ldg(hr + (nchunk - 1) * CH - 1);
    }
    for (long ci = nchunk - 1; ci >= 0; --ci) {
      const long t0 = ci * CH, t = t0 + (long)lane * VEC;
      float an[VEC] = {}, hn[VEC] = {}, gn[VEC] = {}, hbn = 0.f;
      if (ci > 0) {
        ldv<VEC>(ar, t - CH, T, 1.f, an);
        ldv<VEC>(hr, t - CH, T, 0.f, hn);
        ldv<VEC>(gr, t - CH, T, 0.f, gn);
        if (lane == 0 && ci > 1) hbn = __ldg(hr + t0 - CH - 1);
      }
      // local reverse scan from u_in = 0
      float U = 0.f, A = 1.f, dl[VEC], pm[VEC];
#pragma unroll
      for (int e = VEC - 1; e >= 0; --e) {
        const float d = gv[e] + U;
        dl[e] = d;
        pm[e] = A;
        U = av[e] * d;
        A *= av[e];
      }
      // inclusive suffix scan of (A, U) over lanes
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float Aq = __shfl_down_sync(0xffffffffu, A, d);
        const float Uq = __shfl_down_sync(0xffffffffu, U, d);
        if (lane + d < 32) {
          U = fmaf(A, Uq, U);
          A *= Aq;
        }
      }
      float Aex = __shfl_down_sync(0xffffffffu, A, 1);
      float Uex = __shfl_down_sync(0xffffffffu, U, 1);
      if (lane == 31) { Aex = 1.f; Uex = 0.f; }
      const float uin = fmaf(Aex, carry, Uex);
      // h_{t-1} of the lane's first element comes from the previous lane (lane 0: element before the chunk)
      float hprev0 = __shfl_up_sync(0xffffffffu, hv[VEC - 1], 1);
      if (lane == 0) hprev0 = hb;
      float dtok[VEC], dgat[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        dtok[e] = fmaf(pm[e], uin, dl[e]);
        dgat[e] = (e == 0 ? hprev0 : hv[e > 0 ? e - 1 : 0]) * dtok[e];
      }
      stv<VEC>(d_tokens + row * T, t, T, dtok);
      stv<VEC>(d_gates + row * T, t, T, dgat);
      carry = fmaf(__shfl_sync(0xffffffffu, A, 0), carry, __shfl_sync(0xffffffffu, U, 0));
#pragma unroll
      for (int e = 0; e < VEC; ++e) { av[e] = an[e]; hv[e] = hn[e]; gv[e] = gn[e]; }
      hb = hbn;
    }
  }
}

static int scan_grid(long rows) {
  const long warps_per_block = 8;
  long blocks = (rows + warps_per_block - 1) / warps_per_block;
  const long cap = (long)sm_count() * 8;  // 8 CTAs x 8 warps = 64 warps / SM, grid-stride beyond that
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace bdlru

using namespace bdlru;

extern "C" BDLRU_API int bdlru_scan_fwd(const float* gates, const float* tokens, float* states, int64_t rows, int64_t T,
                              void* stream) {
  BDLRU_REQUIRE(gates && tokens && states, "scan_fwd: null pointer");
  BDLRU_REQUIRE(rows >= 1 && T >= 1, "scan_fwd: bad shape rows=%lld T=%lld", (long long)rows, (long long)T);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = (T % 4 == 0) && aligned(gates, 16) && aligned(tokens, 16) && aligned(states, 16);
  const int grid = scan_grid(rows);
  if (vec)
    scan_bct_fwd_kernel<4><<<grid, 256, 0, st>>>(gates, tokens, states, rows, T);
  else
    scan_bct_fwd_kernel<1><<<grid, 256, 0, st>>>(gates, tokens, states, rows, T);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

extern "C" BDLRU_API int bdlru_scan_bwd(const float* gates, const float* states, const float* grad_out, float* d_gates,
                              float* d_tokens, int64_t rows, int64_t T, void* stream) {
  BDLRU_REQUIRE(gates && states && grad_out && d_gates && d_tokens, "scan_bwd: null pointer");
  BDLRU_REQUIRE(rows >= 1 && T >= 1, "scan_bwd: bad shape rows=%lld T=%lld", (long long)rows, (long long)T);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = (T % 4 == 0) && aligned(gates, 16) && aligned(states, 16) && aligned(grad_out, 16) &&
                   aligned(d_gates, 16) && aligned(d_tokens, 16);
  const int grid = scan_grid(rows);
  if (vec)
    scan_bct_bwd_kernel<4><<<grid, 256, 0, st>>>(gates, states, grad_out, d_gates, d_tokens, rows, T);
  else
    scan_bct_bwd_kernel<1><<<grid, 256, 0, st>>>(gates, states, grad_out, d_gates, d_tokens, rows, T);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}
