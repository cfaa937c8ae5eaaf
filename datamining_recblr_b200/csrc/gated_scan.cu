// S1: fused BD-LRU gate math + chunked scan on channel-last [B, T, C] views (and the same tiling
// without gate math = channel-last raw scan).  Replaces RecBLR.py:197-200 and, through h0/dh0, the
// left-pad of RecBLR.py:177-179,203-204.
//
// Tiling.  A persistent CTA owns (batch b, channel tile) units.  Its threads form a [NS, tcn] grid:
// tcn threads cover the channel tile with one 4-channel vector each (coalesced 16 B / 8 B per thread
// along C), NS "time slices" each take S consecutive steps, so one iteration covers NS*S time steps.
// Inputs are staged through shared memory with cp.async (LDGSTS) into THREAD-PRIVATE slots — a thread
// only ever reads back what it copied itself — so the two-deep pipeline needs no block barrier for the
// data, only cp.async.wait_group.  Per iteration:
//   pass 1  each slice scans its S steps from a zero state, keeping the local states and the running
//           product of gates in registers (chunk aggregate (A, H));
//   combine aggregates go through shared memory, every thread walks the NS aggregates of its channel
//           vector to get its carry-in and the CTA state after the block (one __syncthreads);
//   pass 2  h_t = local_t + cumprod_t * carry_in, written straight to HBM.
// The backward mirrors this in reverse time with u_t = a_t * dh~_t as the carried quantity.
// HBM traffic is the algorithmic minimum: fwd reads x', r, i and writes h (4 units);
// bwd reads x', r, i, h, g and writes dx', dr, di (8 units).   [SURVEY §8d: 12*E*s bytes fwd+bwd]
#include <stdlib.h>

#include "common.cuh"

namespace bdlru {

struct View {
  const unsigned char* p;
  long bs, rs;  // element strides
};

struct GScanParams {
  View x, r, i, z;      // GATED: x', r, i (+z).  RAW: x = tokens b, r = gates a.
  View h, y;            // fwd outputs (h is an input in bwd); y iff HAS_Z
  View g;               // bwd: upstream grad (dL/dh, or dL/dy when HAS_Z)
  View dx, dr, di, dz;  // bwd outputs (RAW: dx = d_tokens, dr = d_gates)
  const float* lambda;
  const float* h0;
  long h0_bs;
  float* part_dc;   // [grid, V*tcn] partial sums of dL/dc (c = softplus(Lambda))
  float* part_dh0;  // [grid, V*tcn] partial sums of dh0 when h0 is broadcast
  float* dh0;       // direct output when h0 is per batch
  int B, T, C;
  int tcn, NS, n_ctile, n_iter, n_units;
  bool wide_ok;     // every view is 16-byte aligned with strides that are multiples of 8 elements (bf16: 8 channels/thread)
};

template <typename T>
__device__ __forceinline__ const unsigned char* at(const View& v, long b, long t, int c) {
  return v.p + ((b * v.bs + t * v.rs + c) * (long)sizeof(T));
}
template <typename T>
__device__ __forceinline__ long row_bytes(const View& v) { return v.rs * (long)sizeof(T); }

// NTC: threads per CTA when known at compile time (256: every smem offset folds into an immediate),
// 0: read blockDim.x (rare small-C shapes).
template <int NTC>
__device__ __forceinline__ int cta_threads() { return NTC ? NTC : (int)blockDim.x; }

// ------------------------------------------------------------------------------------------ forward
// V = channels per thread (one 16-byte access: 4 x fp32 or 8 x bf16; 4 x bf16 when C is not a multiple of 8).
// Per-thread vectors of V floats go through shared memory as V/4 float4.
template <int V>
__device__ __forceinline__ void st_vec(float* dst, const float (&v)[V]) {
#pragma unroll
  for (int k = 0; k < V / 4; ++k)
    reinterpret_cast<float4*>(dst)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}
template <int V>
__device__ __forceinline__ void ld_vec(const float* src, float (&v)[V]) {
#pragma unroll
  for (int k = 0; k < V / 4; ++k) {
    const float4 f = reinterpret_cast<const float4*>(src)[k];
    v[4 * k] = f.x; v[4 * k + 1] = f.y; v[4 * k + 2] = f.z; v[4 * k + 3] = f.w;
  }
}

template <typename T, int V, int S, bool GATED, bool HAS_Z, int NTC>
__global__ void __launch_bounds__(256) gscan_fwd_kernel(const GScanParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  using IOx = IOV<T, V>;
  constexpr int EB = IOx::BYTES;
  constexpr bool FAST = sizeof(T) == 2;
  constexpr int NARR = (GATED ? 3 : 2) + (HAS_Z ? 1 : 0);
  const int NT = cta_threads<NTC>(), tid = threadIdx.x;
  const int tc = tid % p.tcn, ts = tid / p.tcn;
  const int ROW = NT * EB;                      // bytes between consecutive staged vectors of a thread
  const int stage_bytes = NARR * S * ROW;
  unsigned char* mine = smem + tid * EB;        // thread-private column of the staging area
  float* agg = reinterpret_cast<float*>(smem + 2 * stage_bytes);  // [2 parity][2 (A,H)][NT][V]

  const int my_units = (p.n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int my_total = my_units * p.n_iter;
  const int Tb = p.NS * S;
  // channel offset is fixed per CTA (the host sizes the grid as a multiple of n_ctile); the producer
  // side keeps its own (unit, iteration) cursor so that no division happens per iteration.
  const int c = (((int)blockIdx.x % p.n_ctile) * p.tcn + tc) * V;
  const long xrb = row_bytes<T>(p.x), rrb = row_bytes<T>(p.r), irb = row_bytes<T>(p.i), zrb = row_bytes<T>(p.z);
  const long hrb = row_bytes<T>(p.h), yrb = row_bytes<T>(p.y);

  int pu = blockIdx.x, pit = 0;  // producer cursor
  auto issue = [&](int buf) {
    if (pu < p.n_units) {
      const long b = pu / p.n_ctile;
      const int t0 = pit * Tb + ts * S;
      unsigned char* dst = mine + buf * stage_bytes;
      const unsigned char* px = at<T>(p.x, b, t0, c);
      const unsigned char* pr = at<T>(p.r, b, t0, c);
      const unsigned char* pi = GATED ? at<T>(p.i, b, t0, c) : nullptr;
      const unsigned char* pz = HAS_Z ? at<T>(p.z, b, t0, c) : nullptr;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        if (t0 + s < p.T) {
          cp_async<EB>(dst + (0 * S + s) * ROW, px + s * xrb);
          cp_async<EB>(dst + (1 * S + s) * ROW, pr + s * rrb);
          if (GATED) cp_async<EB>(dst + (2 * S + s) * ROW, pi + s * irb);
          if (HAS_Z) cp_async<EB>(dst + ((NARR - 1) * S + s) * ROW, pz + s * zrb);
        }
      }
      if (++pit == p.n_iter) { pit = 0; pu += gridDim.x; }
    }
    cp_async_commit();
  };

  issue(0);
  issue(1);

  float state[V], csp[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    state[e] = 0.f;
    csp[e] = GATED ? softplus_acc(p.lambda[c + e]) : 0.f;
  }
  int cu = blockIdx.x, it = 0, buf = 0;  // consumer cursor
  long b = cu / p.n_ctile;
  for (int w = 0; w < my_total; ++w, buf ^= 1) {
    const int t0 = it * Tb + ts * S;
    const unsigned char* src = mine + buf * stage_bytes;
    if (it == 0) {
#pragma unroll
      for (int e = 0; e < V; ++e) state[e] = p.h0 ? p.h0[b * p.h0_bs + c + e] : 0.f;
    }
    cp_async_wait<1>();

    // pass 1: local scan from a zero state
    float hloc[S][V], cum[S][V];
    float hl[V], ca[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { hl[e] = 0.f; ca[e] = 1.f; }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (t0 + s < p.T) {
        float xv[V], rv[V], iv[V];
        IOx::load(src + (0 * S + s) * ROW, xv);
        IOx::load(src + (1 * S + s) * ROW, rv);
        if (GATED) IOx::load(src + (2 * S + s) * ROW, iv);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          float a, bb;
          if (GATED) {
            Gate gt = gate_full<FAST>(csp[e], rv[e], iv[e]);
            a = gt.a;
            bb = gt.q * gt.si * xv[e];
          } else {
            a = rv[e];
            bb = xv[e];
          }
          hl[e] = fmaf(a, hl[e], bb);
          ca[e] *= a;
        }
      }
#pragma unroll
      for (int e = 0; e < V; ++e) { hloc[s][e] = hl[e]; cum[s][e] = ca[e]; }
    }
    if (!HAS_Z) issue(buf);  // slots of this buffer are consumed: refill (thread-private)

    // combine chunk aggregates across the NS slices
    float* aggA = agg + (size_t)(buf * 2 + 0) * NT * V;
    float* aggH = agg + (size_t)(buf * 2 + 1) * NT * V;
    st_vec<V>(aggA + tid * V, ca);
    st_vec<V>(aggH + tid * V, hl);
    __syncthreads();
    float cin[V];
#pragma unroll
    for (int e = 0; e < V; ++e) cin[e] = 0.f;
    for (int j = 0; j < p.NS; ++j) {
      float A[V], H[V];
      ld_vec<V>(aggA + (j * p.tcn + tc) * V, A);
      ld_vec<V>(aggH + (j * p.tcn + tc) * V, H);
      if (j == ts) {
#pragma unroll
        for (int e = 0; e < V; ++e) cin[e] = state[e];
      }
#pragma unroll
      for (int e = 0; e < V; ++e) state[e] = fmaf(A[e], state[e], H[e]);
    }

    // pass 2: apply the carry-in and write
    unsigned char* ph = const_cast<unsigned char*>(at<T>(p.h, b, t0, c));
    unsigned char* py = HAS_Z ? const_cast<unsigned char*>(at<T>(p.y, b, t0, c)) : nullptr;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (t0 + s < p.T) {
        float hv[V];
#pragma unroll
        for (int e = 0; e < V; ++e) hv[e] = fmaf(cum[s][e], cin[e], hloc[s][e]);
        IOx::store(ph + s * hrb, hv);
        if (HAS_Z) {
          float zv[V], yv[V];
          IOx::load(src + ((NARR - 1) * S + s) * ROW, zv);
#pragma unroll
          for (int e = 0; e < V; ++e) yv[e] = zv[e] * sigmoid_t<FAST>(zv[e]) * hv[e];
          IOx::store(py + s * yrb, yv);
        }
      }
    }
    if (HAS_Z) issue(buf);
    if (++it == p.n_iter) { it = 0; cu += gridDim.x; b = cu / p.n_ctile; }
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------ backward
// Staged vectors per thread: GATED: x'[S], r[S], i[S], g[S], h[t0-1 ..][NH] (+ z[S] and one more h entry
// when HAS_Z);  RAW: a[S] (in p.r), g[S], h[NH].
template <typename T, int V, int S, bool GATED, bool HAS_Z, int NTC>
__global__ void __launch_bounds__(256) gscan_bwd_kernel(const GScanParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  using IOx = IOV<T, V>;
  constexpr int EB = IOx::BYTES;
  constexpr bool FAST = sizeof(T) == 2;
  constexpr int NH = S + (HAS_Z ? 1 : 0);                          // staged h entries: h[t0-1 .. t0+NH-2]
  constexpr int NVEC = (GATED ? 4 : 2) * S + NH + (HAS_Z ? S : 0);  // vectors staged per thread
  constexpr int OFF_X = 0, OFF_R = GATED ? S : 0, OFF_I = 2 * S, OFF_G = GATED ? 3 * S : S;
  constexpr int OFF_H = OFF_G + S, OFF_Z = OFF_H + NH;
  const int NT = cta_threads<NTC>(), tid = threadIdx.x;
  const int tc = tid % p.tcn, ts = tid / p.tcn;
  const int ROW = NT * EB;
  const int stage_bytes = NVEC * ROW;
  unsigned char* mine = smem + tid * EB;
  float* agg = reinterpret_cast<float*>(smem + 2 * stage_bytes);  // [2][2][NT][V]
  // GATED pass 1 leaves sigmoid(r) and alpha for pass 2 in thread-private fp32 scratch: for fp32 I/O it
  // overwrites the consumed r and g staging slots, for bf16 I/O (smaller slots) it has its own array.
  constexpr bool INPLACE = sizeof(T) == 4;
  float* scr = agg + (size_t)4 * NT * V + tid * V;  // !INPLACE: [2 (sr, a)][S][NT][V]

  const int my_units = (p.n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int my_total = my_units * p.n_iter;
  const int Tb = p.NS * S;
  const int c = (((int)blockIdx.x % p.n_ctile) * p.tcn + tc) * V;  // fixed per CTA
  const long xrb = row_bytes<T>(p.x), rrb = row_bytes<T>(p.r), irb = row_bytes<T>(p.i), zrb = row_bytes<T>(p.z);
  const long hrb = row_bytes<T>(p.h), grb = row_bytes<T>(p.g);
  const long dxrb = row_bytes<T>(p.dx), drrb = row_bytes<T>(p.dr), dirb = row_bytes<T>(p.di), dzrb = row_bytes<T>(p.dz);

  int pu = blockIdx.x, pit = p.n_iter - 1;  // producer cursor (time blocks run backwards)
  auto issue = [&](int buf) {
    if (pu < p.n_units) {
      const long b = pu / p.n_ctile;
      const int t0 = pit * Tb + ts * S;
      unsigned char* dst = mine + buf * stage_bytes;
      const unsigned char* px = GATED ? at<T>(p.x, b, t0, c) : nullptr;
      const unsigned char* pr = at<T>(p.r, b, t0, c);
      const unsigned char* pi = GATED ? at<T>(p.i, b, t0, c) : nullptr;
      const unsigned char* pg = at<T>(p.g, b, t0, c);
      const unsigned char* pz = HAS_Z ? at<T>(p.z, b, t0, c) : nullptr;
      const unsigned char* ph = at<T>(p.h, b, t0 - 1, c);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        if (t0 + s < p.T) {
          if (GATED) cp_async<EB>(dst + (OFF_X + s) * ROW, px + s * xrb);
          cp_async<EB>(dst + (OFF_R + s) * ROW, pr + s * rrb);
          if (GATED) cp_async<EB>(dst + (OFF_I + s) * ROW, pi + s * irb);
          cp_async<EB>(dst + (OFF_G + s) * ROW, pg + s * grb);
          if (HAS_Z) cp_async<EB>(dst + (OFF_Z + s) * ROW, pz + s * zrb);
        }
      }
#pragma unroll
      for (int s = 0; s < NH; ++s) {
        const int t = t0 + s - 1;
        if (t >= 0 && t < p.T) cp_async<EB>(dst + (OFF_H + s) * ROW, ph + s * hrb);
      }
      if (--pit < 0) { pit = p.n_iter - 1; pu += gridDim.x; }
    }
    cp_async_commit();
  };

  issue(0);
  issue(1);

  float ustate[V], csp[V], h0v[V], dc_acc[V], dh0_acc[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    ustate[e] = h0v[e] = dc_acc[e] = dh0_acc[e] = 0.f;
    csp[e] = GATED ? softplus_acc(p.lambda[c + e]) : 0.f;
  }

  int cu = blockIdx.x, it = p.n_iter - 1, buf = 0;  // consumer cursor
  long b = cu / p.n_ctile;
  for (int w = 0; w < my_total; ++w, buf ^= 1) {
    const int t0 = it * Tb + ts * S;
    unsigned char* src = mine + buf * stage_bytes;
    auto scr_sr = [&](int s) -> float* {
      return INPLACE ? reinterpret_cast<float*>(src + (OFF_R + s) * ROW) : scr + (size_t)(0 * S + s) * NT * V;
    };
    auto scr_a = [&](int s) -> float* {
      return INPLACE ? reinterpret_cast<float*>(src + (OFF_G + s) * ROW) : scr + (size_t)(1 * S + s) * NT * V;
    };
    if (it == p.n_iter - 1) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        h0v[e] = p.h0 ? p.h0[b * p.h0_bs + c + e] : 0.f;
        ustate[e] = 0.f;
      }
    }
    cp_async_wait<1>();

    // pass 1 (reverse time): local dh~ from u_in = 0; pm = product of the gates AFTER step s
    float dloc[S][V], pm[S][V];
    float ul[V], pc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { ul[e] = 0.f; pc[e] = 1.f; }
    unsigned char* pdz = HAS_Z ? const_cast<unsigned char*>(at<T>(p.dz, b, t0, c)) : nullptr;
#pragma unroll
    for (int s = S - 1; s >= 0; --s) {
      float av[V], gv[V];
#pragma unroll
      for (int e = 0; e < V; ++e) { av[e] = 1.f; gv[e] = 0.f; }
      if (t0 + s < p.T) {
        float rv[V];
        IOx::load(src + (OFF_R + s) * ROW, rv);
        IOx::load(src + (OFF_G + s) * ROW, gv);
        if (HAS_Z) {
          // upstream is dL/dy with y = silu(z) * h:  dz = dy * h_t * silu'(z),  g = dy * silu(z)
          float zv[V], hv[V], dzv[V];
          IOx::load(src + (OFF_Z + s) * ROW, zv);
          IOx::load(src + (OFF_H + s + 1) * ROW, hv);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const float sz = sigmoid_t<FAST>(zv[e]);
            dzv[e] = gv[e] * hv[e] * silu_grad_f(zv[e], sz);
            gv[e] *= zv[e] * sz;
          }
          IOx::store(pdz + s * dzrb, dzv);
        }
        if (GATED) {
          float srv[V];
#pragma unroll
          for (int e = 0; e < V; ++e) av[e] = gate_alpha<FAST>(csp[e], rv[e], srv[e]);
          st_vec<V>(scr_sr(s), srv);
          st_vec<V>(scr_a(s), av);
        } else {
#pragma unroll
          for (int e = 0; e < V; ++e) av[e] = rv[e];
        }
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float d = gv[e] + ul[e];
        dloc[s][e] = d;
        pm[s][e] = pc[e];
        ul[e] = av[e] * d;
        pc[e] *= av[e];
      }
    }

    // combine (carry flows from later slices to earlier ones)
    float* aggA = agg + (size_t)(buf * 2 + 0) * NT * V;
    float* aggU = agg + (size_t)(buf * 2 + 1) * NT * V;
    st_vec<V>(aggA + tid * V, pc);
    st_vec<V>(aggU + tid * V, ul);
    __syncthreads();
    float uin[V];
#pragma unroll
    for (int e = 0; e < V; ++e) uin[e] = 0.f;
    for (int j = p.NS - 1; j >= 0; --j) {
      float A[V], U[V];
      ld_vec<V>(aggA + (j * p.tcn + tc) * V, A);
      ld_vec<V>(aggU + (j * p.tcn + tc) * V, U);
      if (j == ts) {
#pragma unroll
        for (int e = 0; e < V; ++e) uin[e] = ustate[e];
      }
#pragma unroll
      for (int e = 0; e < V; ++e) ustate[e] = fmaf(A[e], ustate[e], U[e]);
    }
    if (it == 0 && ts == 0) {  // u flowing out of t = 0 is dL/dh0 for this (b, channel vector)
      if (p.dh0) {
#pragma unroll
        for (int e = 0; e < V; ++e) p.dh0[b * p.C + c + e] = ustate[e];
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) dh0_acc[e] += ustate[e];
      }
    }

    // pass 2: true dh~ and the chain rule through the gate math
    unsigned char* pdx = const_cast<unsigned char*>(at<T>(p.dx, b, t0, c));
    unsigned char* pdr = const_cast<unsigned char*>(at<T>(p.dr, b, t0, c));
    unsigned char* pdi = GATED ? const_cast<unsigned char*>(at<T>(p.di, b, t0, c)) : nullptr;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int t = t0 + s;
      if (t < p.T) {
        float d[V], hp[V];
#pragma unroll
        for (int e = 0; e < V; ++e) d[e] = fmaf(pm[s][e], uin[e], dloc[s][e]);
        if (t == 0) {
#pragma unroll
          for (int e = 0; e < V; ++e) hp[e] = h0v[e];
        } else {
          IOx::load(src + (OFF_H + s) * ROW, hp);
        }
        if (GATED) {
          float xv[V], iv[V], dxv[V], drv[V], div[V], srv[V], av[V];
          IOx::load(src + (OFF_X + s) * ROW, xv);
          IOx::load(src + (OFF_I + s) * ROW, iv);
          ld_vec<V>(scr_sr(s), srv);
          ld_vec<V>(scr_a(s), av);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const float si = sigmoid_t<FAST>(iv[e]);
            const float v = one_minus_a2<FAST>(csp[e], srv[e], av[e]);
            const float rq = rsqrt_ftz(v), q = v * rq;
            const float dbeta = d[e] * xv[e];
            dxv[e] = d[e] * q * si;
            div[e] = dbeta * q * si * (1.0f - si);
            const float da = fmaf(hp[e], d[e], -dbeta * si * av[e] * rq);
            const float daa = da * av[e];
            drv[e] = -csp[e] * daa * srv[e] * (1.0f - srv[e]);
            dc_acc[e] = fmaf(-daa, srv[e], dc_acc[e]);
          }
          IOx::store(pdx + s * dxrb, dxv);
          IOx::store(pdr + s * drrb, drv);
          IOx::store(pdi + s * dirb, div);
        } else {
          float dgv[V];
#pragma unroll
          for (int e = 0; e < V; ++e) dgv[e] = hp[e] * d[e];
          IOx::store(pdx + s * dxrb, d);     // d_tokens
          IOx::store(pdr + s * drrb, dgv);   // d_gates
        }
      }
    }
    issue(buf);
    if (--it < 0) { it = p.n_iter - 1; cu += gridDim.x; b = cu / p.n_ctile; }
  }
  cp_async_wait<0>();

  // per-CTA partial sums of dL/dc and (broadcast) dh0: reduce over the NS slices, one row per CTA.
  // The host sizes the grid as a multiple of n_ctile, so a CTA sees a single channel tile.
  __syncthreads();
  float* red = agg;  // reuse: [NT][V]
  if (GATED) {
    st_vec<V>(red + tid * V, dc_acc);
    __syncthreads();
    if (ts == 0) {
      float s4[V];
      ld_vec<V>(red + tc * V, s4);
      for (int j = 1; j < p.NS; ++j) {
        float v[V];
        ld_vec<V>(red + (j * p.tcn + tc) * V, v);
#pragma unroll
        for (int e = 0; e < V; ++e) s4[e] += v[e];
      }
      st_vec<V>(p.part_dc + ((size_t)blockIdx.x * p.tcn + tc) * V, s4);
    }
  }
  if (p.part_dh0 && ts == 0) st_vec<V>(p.part_dh0 + ((size_t)blockIdx.x * p.tcn + tc) * V, dh0_acc);
}

// ------------------------------------------------------------------------------------------ sequential variant
// Large B*C (the training shapes: 8 192 x 256, 2 048 x 128): there are enough (batch row, channel vector) pairs to fill the
// machine with ONE thread per pair walking the whole sequence — no time slices, no chunk aggregates, no second pass, no
// shared-memory staging, no block barrier.  ncu on the chunked bf16 kernels showed them issue-bound (62.6 thread
// instructions per element forward against a budget of 66 at full HBM rate, a quarter of the stall samples on the
// aggregate walk and its barrier); this form needs ~35.  Lanes of a warp hold consecutive channel vectors of one row, so
// every step is one coalesced row segment per array; U steps are loaded before they are consumed (U x 3..6 independent
// loads in flight per thread) and occupancy hides the rest.  GATED, bf16 only; everything else stays on the chunked kernels.
template <typename T, int V, int U, bool HAS_Z>
__global__ void __launch_bounds__(256) gscan_seq_fwd_kernel(const GScanParams p) {
  using IOx = IOV<T, V>;
  constexpr bool FAST = sizeof(T) == 2;
  const int lanes = p.C / V, rpb = (int)blockDim.x / lanes;
  const int li = (int)threadIdx.x % lanes, rw = (int)threadIdx.x / lanes;
  const int c = li * V;
  const long xrb = row_bytes<T>(p.x), rrb = row_bytes<T>(p.r), irb = row_bytes<T>(p.i), zrb = row_bytes<T>(p.z);
  const long hrb = row_bytes<T>(p.h), yrb = row_bytes<T>(p.y);
  float csp[V];
#pragma unroll
  for (int e = 0; e < V; ++e) csp[e] = softplus_acc(p.lambda[c + e]);
  for (long b = (long)blockIdx.x * rpb + rw; b < p.B; b += (long)gridDim.x * rpb) {
    float h[V];
#pragma unroll
    for (int e = 0; e < V; ++e) h[e] = p.h0 ? p.h0[b * p.h0_bs + c + e] : 0.f;
    const unsigned char* px = at<T>(p.x, b, 0, c);
    const unsigned char* pr = at<T>(p.r, b, 0, c);
    const unsigned char* pi = at<T>(p.i, b, 0, c);
    const unsigned char* pz = HAS_Z ? at<T>(p.z, b, 0, c) : nullptr;
    unsigned char* ph = const_cast<unsigned char*>(at<T>(p.h, b, 0, c));
    unsigned char* py = HAS_Z ? const_cast<unsigned char*>(at<T>(p.y, b, 0, c)) : nullptr;
    for (int t0 = 0; t0 < p.T; t0 += U) {
      float xv[U][V], rv[U][V], iv[U][V], zv[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (t0 + u < p.T) {
          IOx::load(px + (long)(t0 + u) * xrb, xv[u]);
          IOx::load(pr + (long)(t0 + u) * rrb, rv[u]);
          IOx::load(pi + (long)(t0 + u) * irb, iv[u]);
          if (HAS_Z) IOx::load(pz + (long)(t0 + u) * zrb, zv[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (t0 + u < p.T) {
          float hv[V], yv[V];
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const Gate g = gate_full<FAST>(csp[e], rv[u][e], iv[u][e]);
            h[e] = fmaf(g.a, h[e], g.q * g.si * xv[u][e]);
            hv[e] = h[e];
            if (HAS_Z) yv[e] = zv[u][e] * sigmoid_t<FAST>(zv[u][e]) * h[e];
          }
          IOx::store(ph + (long)(t0 + u) * hrb, hv);
          if (HAS_Z) IOx::store(py + (long)(t0 + u) * yrb, yv);
        }
      }
    }
  }
}

template <typename T, int V, int U, bool HAS_Z>
__global__ void __launch_bounds__(256) gscan_seq_bwd_kernel(const GScanParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  using IOx = IOV<T, V>;
  constexpr bool FAST = sizeof(T) == 2;
  const int lanes = p.C / V, rpb = (int)blockDim.x / lanes;
  const int li = (int)threadIdx.x % lanes, rw = (int)threadIdx.x / lanes;
  const int c = li * V;
  const long xrb = row_bytes<T>(p.x), rrb = row_bytes<T>(p.r), irb = row_bytes<T>(p.i), zrb = row_bytes<T>(p.z);
  const long hrb = row_bytes<T>(p.h), grb = row_bytes<T>(p.g);
  const long dxrb = row_bytes<T>(p.dx), drrb = row_bytes<T>(p.dr), dirb = row_bytes<T>(p.di), dzrb = row_bytes<T>(p.dz);
  float csp[V], dc_acc[V], dh0_acc[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    csp[e] = softplus_acc(p.lambda[c + e]);
    dc_acc[e] = dh0_acc[e] = 0.f;
  }
  for (long b = (long)blockIdx.x * rpb + rw; b < p.B; b += (long)gridDim.x * rpb) {
    float u_[V], h0v[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      u_[e] = 0.f;
      h0v[e] = p.h0 ? p.h0[b * p.h0_bs + c + e] : 0.f;
    }
    const unsigned char* px = at<T>(p.x, b, 0, c);
    const unsigned char* pr = at<T>(p.r, b, 0, c);
    const unsigned char* pi = at<T>(p.i, b, 0, c);
    const unsigned char* pg = at<T>(p.g, b, 0, c);
    const unsigned char* pz = HAS_Z ? at<T>(p.z, b, 0, c) : nullptr;
    const unsigned char* ph = at<T>(p.h, b, 0, c);
    unsigned char* pdx = const_cast<unsigned char*>(at<T>(p.dx, b, 0, c));
    unsigned char* pdr = const_cast<unsigned char*>(at<T>(p.dr, b, 0, c));
    unsigned char* pdi = const_cast<unsigned char*>(at<T>(p.di, b, 0, c));
    unsigned char* pdz = HAS_Z ? const_cast<unsigned char*>(at<T>(p.dz, b, 0, c)) : nullptr;
    // time chunks in reverse: chunk k covers steps [k*U, k*U + U)
    for (int t0 = ((p.T - 1) / U) * U; t0 >= 0; t0 -= U) {
      float xv[U][V], rv[U][V], iv[U][V], gv[U][V], zv[U][V], hv[U + 1][V];   // hv[j] = h_{t0 + j - 1}
#pragma unroll
      for (int j = 0; j <= U; ++j) {
        const int t = t0 + j - 1;
        if (t < 0) {
#pragma unroll
          for (int e = 0; e < V; ++e) hv[j][e] = h0v[e];
        } else if (t < p.T) {
          IOx::load(ph + (long)t * hrb, hv[j]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (t0 + u < p.T) {
          IOx::load(px + (long)(t0 + u) * xrb, xv[u]);
          IOx::load(pr + (long)(t0 + u) * rrb, rv[u]);
          IOx::load(pi + (long)(t0 + u) * irb, iv[u]);
          IOx::load(pg + (long)(t0 + u) * grb, gv[u]);
          if (HAS_Z) IOx::load(pz + (long)(t0 + u) * zrb, zv[u]);
        }
      }
#pragma unroll
      for (int u = U - 1; u >= 0; --u) {
        if (t0 + u < p.T) {
          float dxv[V], drv[V], div[V], dzv[V];
#pragma unroll
          for (int e = 0; e < V; ++e) {
            float g = gv[u][e];
            if (HAS_Z) {
              const float sz = sigmoid_t<FAST>(zv[u][e]);
              dzv[e] = g * hv[u + 1][e] * silu_grad_f(zv[u][e], sz);
              g *= zv[u][e] * sz;
            }
            float sr;
            const float a = gate_alpha<FAST>(csp[e], rv[u][e], sr);
            const float si = sigmoid_t<FAST>(iv[u][e]);
            const float v = one_minus_a2<FAST>(csp[e], sr, a);
            const float rq = rsqrt_ftz(v), q = v * rq;
            const float d = g + u_[e];
            const float dbeta = d * xv[u][e];
            dxv[e] = d * q * si;
            div[e] = dbeta * q * si * (1.0f - si);
            const float da = fmaf(hv[u][e], d, -dbeta * si * a * rq);
            const float daa = da * a;
            drv[e] = -csp[e] * daa * sr * (1.0f - sr);
            dc_acc[e] = fmaf(-daa, sr, dc_acc[e]);
            u_[e] = a * d;
          }
          IOx::store(pdx + (long)(t0 + u) * dxrb, dxv);
          IOx::store(pdr + (long)(t0 + u) * drrb, drv);
          IOx::store(pdi + (long)(t0 + u) * dirb, div);
          if (HAS_Z) IOx::store(pdz + (long)(t0 + u) * dzrb, dzv);
        }
      }
    }
    if (p.dh0) {
#pragma unroll
      for (int e = 0; e < V; ++e) p.dh0[b * p.C + c + e] = u_[e];
    } else {
#pragma unroll
      for (int e = 0; e < V; ++e) dh0_acc[e] += u_[e];
    }
  }
  // per-CTA partial rows [C] of dL/dc and (broadcast) dh0: reduce over the rpb batch rows of the block
  float* red = reinterpret_cast<float*>(smem);   // [2][blockDim.x][V]
#pragma unroll
  for (int e = 0; e < V; ++e) {
    red[(size_t)threadIdx.x * V + e] = dc_acc[e];
    red[(size_t)(blockDim.x + threadIdx.x) * V + e] = dh0_acc[e];
  }
  __syncthreads();
  if (rw == 0) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      float s0 = 0.f, s1 = 0.f;
      for (int j = 0; j < rpb; ++j) {
        s0 += red[(size_t)(j * lanes + li) * V + e];
        s1 += red[(size_t)(blockDim.x + j * lanes + li) * V + e];
      }
      p.part_dc[(size_t)blockIdx.x * p.C + c + e] = s0;
      if (p.part_dh0) p.part_dh0[(size_t)blockIdx.x * p.C + c + e] = s1;
    }
  }
}

// sequential variant: which vector width (0 = use the chunked kernels)
// Measured on B200 (tools/scan_bench.py, bf16): SLOWER than the chunked kernels — 2 048 x 200 x 128 fwd+bwd 0.68 vs 0.35 ms,
// 8 192 x 200 x 256 (z-gated) fwd 1.30 vs 1.10 ms, bwd 2.81 vs 2.24 ms: with 94-128 registers per thread only 16-20 warps
// per SM are resident and each walks its row with dependent strided loads; the chunked kernels' cp.async staging hides
// that latency better than occupancy does.  Kept as a build-time variant (-DBDLRU_GSCAN_SEQ), parity-tested, off.
static int seq_vector_width(const GScanParams& p) {
#ifndef BDLRU_GSCAN_SEQ
  return 0;
#endif
  const long min_warps = 4096;   // ~28 warps per SM
  for (int V : {4, 2}) {
    if (p.C % V) continue;
    const int lanes = p.C / V;
    if (lanes > 256 || 256 % lanes) continue;
    if ((long)p.B * lanes / 32 >= min_warps) return V;
  }
  return 0;
}

template <typename T, int V, bool HAS_Z>
static int launch_seq_fwd(GScanParams& p, cudaStream_t st) {
  const int lanes = p.C / V, rpb = 256 / lanes;
  long grid = ((long)p.B + rpb - 1) / rpb;
  const long cap = (long)sm_count() * 8;
  if (grid > cap) grid = cap;
  gscan_seq_fwd_kernel<T, V, 4, HAS_Z><<<(unsigned)grid, 256, 0, st>>>(p);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

template <typename T, int V, bool HAS_Z>
static int launch_seq_bwd(GScanParams& p, float* dLambda, float* dh0_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int lanes = p.C / V, rpb = 256 / lanes;
  long grid = ((long)p.B + rpb - 1) / rpb;
  const long cap = (long)sm_count() * 6;
  if (grid > cap) grid = cap;
  const size_t need = 2 * (size_t)grid * p.C * sizeof(float);
  BDLRU_REQUIRE(ws && ws_bytes >= need, "gated_scan_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
  p.part_dc = reinterpret_cast<float*>(ws);
  const bool bcast_h0 = dh0_out && p.h0_bs == 0;
  p.part_dh0 = bcast_h0 ? p.part_dc + (size_t)grid * p.C : nullptr;
  p.dh0 = (dh0_out && !bcast_h0) ? dh0_out : nullptr;
  const size_t smem = 2 * 256 * (size_t)V * sizeof(float);
  gscan_seq_bwd_kernel<T, V, 2, HAS_Z><<<(unsigned)grid, 256, smem, st>>>(p);
  BDLRU_LAUNCHED();
  int rc = launch_colsum(p.part_dc, (int)grid, p.C, p.C, COLSUM_SIGMOID, dLambda, nullptr, 0, p.lambda, st);
  if (rc) return rc;
  if (bcast_h0) {
    rc = launch_colsum(p.part_dh0, (int)grid, p.C, p.C, COLSUM_SPLIT, dh0_out, nullptr, p.C, nullptr, st);
    if (rc) return rc;
  }
  return BDLRU_OK;
}

// ------------------------------------------------------------------------------------------ host side
struct Tiling {
  int tcn, NS, NT, n_ctile, n_iter, n_units, grid;
};

template <int S, int V>
static Tiling make_tiling(int B, int T, int C, int esize) {
  Tiling t;
  const int cvec = C / V;
  t.tcn = 1;
  while (t.tcn < 32 && cvec % (t.tcn * 2) == 0) t.tcn *= 2;
  t.n_ctile = cvec / t.tcn;
  // time slices per CTA: enough to cover T when it is short, at most 256 threads / 32 slices
  int ns = 256 / t.tcn;
  if (ns > 32) ns = 32;  // tcn < 8 (C/4 not a multiple of 8): fewer threads rather than a long combine loop
  // bf16 I/O is issue-bound (half the bytes, the same instructions): smaller CTAs = a shorter aggregate walk and cheaper
  // barrier per element, more CTAs per SM.  Measured on B200 (tools/scan_bench.py, S1 fwd+bwd, fraction of HBM peak):
  // 2 048 x 200 x 128: NS 8 / 4 / 2 = 0.54 / 0.63 / 0.58;  256 x 4096 x 256: 0.53 / 0.58 / 0.50;  8 192 x 200 x 256 (z-gated)
  // fwd 0.69 / 0.76 / 0.82, bwd 0.57 / 0.60 / 0.63.  fp32 I/O (0.82-0.95) is bandwidth-bound and keeps 8.
  if (esize == 2 && t.tcn == 32) ns = ((long)B * t.n_ctile >= 8192) ? 2 : 4;
  static const int force_ns = tuning_env("BDLRU_GSCAN_NS");
  if (force_ns > 0 && force_ns < ns) ns = force_ns;
  t.NS = ns;
  (void)T;
  t.NT = t.tcn * t.NS;
  t.n_iter = (T + t.NS * S - 1) / (t.NS * S);
  t.n_units = B * t.n_ctile;
  return t;
}

static int pick_grid(const Tiling& t, int ctas_per_sm) {
  long g = (long)sm_count() * ctas_per_sm;
  if (g > t.n_units) g = t.n_units;
  g = (g / t.n_ctile) * t.n_ctile;  // multiple of n_ctile: a CTA always sees the same channel tile
  if (g < t.n_ctile) g = t.n_ctile;
  return (int)g;
}

static int check_view(const char* name, const bdlru_view& v, int esize, bool required) {
  if (!v.ptr) {
    BDLRU_REQUIRE(!required, "%s: null pointer", name);
    return BDLRU_OK;
  }
  BDLRU_REQUIRE(aligned(v.ptr, 4 * esize), "%s: pointer not aligned to %d bytes", name, 4 * esize);
  BDLRU_REQUIRE(v.bstride % 4 == 0 && v.rstride % 4 == 0, "%s: strides must be multiples of 4 elements", name);
  return BDLRU_OK;
}
static View mk(const bdlru_view& v) { return View{reinterpret_cast<const unsigned char*>(v.ptr), v.bstride, v.rstride}; }
static bool view_wide(const View& v) { return !v.p || (aligned(v.p, 16) && v.bs % 8 == 0 && v.rs % 8 == 0); }
static void set_wide(GScanParams& p) {
  p.wide_ok = view_wide(p.x) && view_wide(p.r) && view_wide(p.i) && view_wide(p.z) && view_wide(p.h) && view_wide(p.y) &&
              view_wide(p.g) && view_wide(p.dx) && view_wide(p.dr) && view_wide(p.di) && view_wide(p.dz);
}

// (V channels, S steps) per thread and iteration: fp32 4 x 4; bf16 8 x 2 when C is a multiple of 8 (same 16 elements
// per thread and iteration, same register footprint, half the memory instructions), else 4 x 4.
template <typename T, int V, int kS, bool GATED, bool HAS_Z>
static int launch_fwd_v(GScanParams& p, cudaStream_t st) {
  Tiling t = make_tiling<kS, V>(p.B, p.T, p.C, (int)sizeof(T));
  p.tcn = t.tcn; p.NS = t.NS; p.n_ctile = t.n_ctile; p.n_iter = t.n_iter; p.n_units = t.n_units;
  constexpr int NARR = (GATED ? 3 : 2) + (HAS_Z ? 1 : 0);
  const size_t smem = 2 * (size_t)NARR * kS * t.NT * IOV<T, V>::BYTES + 4 * (size_t)t.NT * V * sizeof(float);
  auto kern = t.NT == 256 ? gscan_fwd_kernel<T, V, kS, GATED, HAS_Z, 256> : gscan_fwd_kernel<T, V, kS, GATED, HAS_Z, 0>;
  BDLRU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  BDLRU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, t.NT, smem));
  if (occ < 1) occ = 1;
  const int grid = pick_grid(t, occ);
  kern<<<grid, t.NT, smem, st>>>(p);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

// Measured on B200 (tools/scan_bench.py, bf16): 8 channels x 2 steps per thread is SLOWER than 4 x 4 at C = 128
// (2 048 x 200 x 128 fwd+bwd 0.64 ms vs 0.38 ms: only 16 channel vectors, so 16 time slices per CTA and a 16-long
// aggregate walk per 16 elements; with 2 / 4 / 8 slices forced: 0.57 / 0.57 / 0.46 of HBM against 0.63).  With a full warp
// of channel vectors (C a multiple of 256) and the short slice counts of make_tiling it is FASTER: 8 192 x 200 x 256 z-gated
// fwd 0.94 -> 0.89 ms, bwd 2.06 -> 1.95 ms in tools/scan_bench.py (graph replay); inside the configs[4] training step the
// forward went 1.12 -> 1.04 ms but the backward 2.12 -> 2.20 ms.  So: wide exactly when C / 8 fills whole warps, forward only
// (bit 0 = forward, bit 1 = backward).
#ifndef BDLRU_GSCAN_WIDE
#define BDLRU_GSCAN_WIDE 1
#endif
static bool use_wide(const GScanParams& p, int dir) {
  // ... and there are enough (row, channel-tile) units for the 2-slice regime of make_tiling: at 256 x 4096 x 256 (256 wide
  // units on 148 SMs) the wide forward pulled the S1 fwd+bwd line from 0.58 to 0.49 of HBM
  return ((BDLRU_GSCAN_WIDE >> dir) & 1) && p.C % 256 == 0 && p.wide_ok && (long)p.B * (p.C / 256) >= 8192;
}

template <typename T, bool GATED, bool HAS_Z>
static int launch_fwd(GScanParams& p, cudaStream_t st) {
  if constexpr (sizeof(T) == 2 && GATED) {
    const int V = seq_vector_width(p);
    if (V == 4) return launch_seq_fwd<T, 4, HAS_Z>(p, st);
    if (V == 2) return launch_seq_fwd<T, 2, HAS_Z>(p, st);
  }
  if constexpr (sizeof(T) == 2) {
    if (use_wide(p, 0)) return launch_fwd_v<T, 8, 2, GATED, HAS_Z>(p, st);
  }
  return launch_fwd_v<T, 4, 4, GATED, HAS_Z>(p, st);
}

template <typename T, int V, int kS, bool GATED, bool HAS_Z>
static int launch_bwd_v(GScanParams& p, float* dLambda, float* dh0_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  Tiling t = make_tiling<kS, V>(p.B, p.T, p.C, (int)sizeof(T));
  p.tcn = t.tcn; p.NS = t.NS; p.n_ctile = t.n_ctile; p.n_iter = t.n_iter; p.n_units = t.n_units;
  constexpr int NH = kS + (HAS_Z ? 1 : 0);
  constexpr int NVEC = (GATED ? 4 : 2) * kS + NH + (HAS_Z ? kS : 0);
  const size_t smem = 2 * (size_t)NVEC * t.NT * IOV<T, V>::BYTES + 4 * (size_t)t.NT * V * sizeof(float) +
                      ((GATED && sizeof(T) != 4) ? 2 * (size_t)kS * t.NT * V * sizeof(float) : 0);
  auto kern = t.NT == 256 ? gscan_bwd_kernel<T, V, kS, GATED, HAS_Z, 256> : gscan_bwd_kernel<T, V, kS, GATED, HAS_Z, 0>;
  BDLRU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  BDLRU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, t.NT, smem));
  if (occ < 1) occ = 1;
  const int grid = pick_grid(t, occ);
  const int ctw = V * t.tcn;
  const size_t need = 2 * (size_t)grid * ctw * sizeof(float);
  BDLRU_REQUIRE(ws && ws_bytes >= need, "gated_scan_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
  p.part_dc = reinterpret_cast<float*>(ws);
  const bool bcast_h0 = dh0_out && p.h0_bs == 0;
  p.part_dh0 = bcast_h0 ? p.part_dc + (size_t)grid * ctw : nullptr;
  p.dh0 = (dh0_out && !bcast_h0) ? dh0_out : nullptr;
  kern<<<grid, t.NT, smem, st>>>(p);
  BDLRU_LAUNCHED();
  // partial row blk holds channel tile blk % n_ctile: viewed as [grid / n_ctile][C] it is a plain column sum
  if (GATED) {
    int rc = launch_colsum(p.part_dc, grid / t.n_ctile, p.C, p.C, COLSUM_SIGMOID, dLambda, nullptr, 0, p.lambda, st);
    if (rc) return rc;
  }
  if (bcast_h0) {
    int rc = launch_colsum(p.part_dh0, grid / t.n_ctile, p.C, p.C, COLSUM_SPLIT, dh0_out, nullptr, p.C, nullptr, st);
    if (rc) return rc;
  }
  return BDLRU_OK;
}

template <typename T, bool GATED, bool HAS_Z>
static int launch_bwd(GScanParams& p, float* dLambda, float* dh0_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if constexpr (sizeof(T) == 2 && GATED) {
    const int V = seq_vector_width(p);
    if (V == 4) return launch_seq_bwd<T, 4, HAS_Z>(p, dLambda, dh0_out, ws, ws_bytes, st);
    if (V == 2) return launch_seq_bwd<T, 2, HAS_Z>(p, dLambda, dh0_out, ws, ws_bytes, st);
  }
  if constexpr (sizeof(T) == 2) {
    if (use_wide(p, 1)) return launch_bwd_v<T, 8, 2, GATED, HAS_Z>(p, dLambda, dh0_out, ws, ws_bytes, st);
  }
  return launch_bwd_v<T, 4, 4, GATED, HAS_Z>(p, dLambda, dh0_out, ws, ws_bytes, st);
}

static int check_common(int B, int T, int C, int dtype) {
  BDLRU_REQUIRE(B >= 1 && T >= 1 && C >= 4, "bad shape B=%d T=%d C=%d", B, T, C);
  BDLRU_REQUIRE(C % 4 == 0, "C=%d must be a multiple of 4", C);
  BDLRU_REQUIRE(dtype == BDLRU_F32 || dtype == BDLRU_BF16, "bad dtype %d", dtype);
  return BDLRU_OK;
}

}  // namespace bdlru

using namespace bdlru;

#define CHECK_VIEW(name, v, es, req)                        \
  do {                                                      \
    int rc_ = check_view(name, v, es, req);                 \
    if (rc_) return rc_;                                    \
  } while (0)

extern "C" BDLRU_API size_t bdlru_gated_scan_bwd_workspace_bytes(int B, int T, int C) {
  (void)B; (void)T;
  // 2 partial arrays of [grid, V*tcn] floats; grid <= 32 CTAs/SM * SMs, V*tcn <= 256
  const size_t chunked = 2 * (size_t)32 * (size_t)sm_count() * 256 * sizeof(float) + 2 * (size_t)(C / 4 + 1) * 256 * sizeof(float);
  const size_t seq = 2 * (size_t)6 * (size_t)sm_count() * (size_t)C * sizeof(float);   // sequential variant: [grid, C] x 2
  return chunked > seq ? chunked : seq;
}

extern "C" BDLRU_API int bdlru_gated_scan_fwd(bdlru_view xp, bdlru_view r, bdlru_view i, const float* Lambda, const float* h0,
                                    int64_t h0_bstride, bdlru_view z, bdlru_view h, bdlru_view y, int B, int T,
                                    int C, int dtype, void* stream) {
  int rc = check_common(B, T, C, dtype);
  if (rc) return rc;
  const int es = dtype == BDLRU_F32 ? 4 : 2;
  CHECK_VIEW("xp", xp, es, true); CHECK_VIEW("r", r, es, true); CHECK_VIEW("i", i, es, true);
  CHECK_VIEW("h", h, es, true); CHECK_VIEW("z", z, es, false);
  if (z.ptr) CHECK_VIEW("y", y, es, true);
  BDLRU_REQUIRE(Lambda, "Lambda: null pointer");
  BDLRU_REQUIRE(h0_bstride == 0 || h0_bstride == C, "h0_bstride must be 0 or C");
  GScanParams p{};
  p.x = mk(xp); p.r = mk(r); p.i = mk(i); p.z = mk(z); p.h = mk(h); p.y = mk(y);
  p.lambda = Lambda; p.h0 = h0; p.h0_bs = h0_bstride; p.B = B; p.T = T; p.C = C;
  set_wide(p);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == BDLRU_F32)
    return z.ptr ? launch_fwd<float, true, true>(p, st) : launch_fwd<float, true, false>(p, st);
  return z.ptr ? launch_fwd<__nv_bfloat16, true, true>(p, st) : launch_fwd<__nv_bfloat16, true, false>(p, st);
}

extern "C" BDLRU_API int bdlru_gated_scan_bwd(bdlru_view xp, bdlru_view r, bdlru_view i, const float* Lambda, const float* h0,
                                    int64_t h0_bstride, bdlru_view z, bdlru_view h, bdlru_view grad, bdlru_view dxp,
                                    bdlru_view dr, bdlru_view di, bdlru_view dz, float* dLambda, float* dh0,
                                    void* workspace, size_t workspace_bytes, int B, int T, int C, int dtype,
                                    void* stream) {
  int rc = check_common(B, T, C, dtype);
  if (rc) return rc;
  const int es = dtype == BDLRU_F32 ? 4 : 2;
  CHECK_VIEW("xp", xp, es, true); CHECK_VIEW("r", r, es, true); CHECK_VIEW("i", i, es, true);
  CHECK_VIEW("h", h, es, true); CHECK_VIEW("grad", grad, es, true); CHECK_VIEW("dxp", dxp, es, true);
  CHECK_VIEW("dr", dr, es, true); CHECK_VIEW("di", di, es, true); CHECK_VIEW("z", z, es, false);
  if (z.ptr) CHECK_VIEW("dz", dz, es, true);
  BDLRU_REQUIRE(Lambda && dLambda, "Lambda/dLambda: null pointer");
  BDLRU_REQUIRE(h0_bstride == 0 || h0_bstride == C, "h0_bstride must be 0 or C");
  GScanParams p{};
  p.x = mk(xp); p.r = mk(r); p.i = mk(i); p.z = mk(z); p.h = mk(h); p.g = mk(grad);
  p.dx = mk(dxp); p.dr = mk(dr); p.di = mk(di); p.dz = mk(dz);
  p.lambda = Lambda; p.h0 = h0; p.h0_bs = h0_bstride; p.B = B; p.T = T; p.C = C;
  set_wide(p);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == BDLRU_F32)
    return z.ptr ? launch_bwd<float, true, true>(p, dLambda, dh0, workspace, workspace_bytes, st)
                 : launch_bwd<float, true, false>(p, dLambda, dh0, workspace, workspace_bytes, st);
  return z.ptr ? launch_bwd<__nv_bfloat16, true, true>(p, dLambda, dh0, workspace, workspace_bytes, st)
               : launch_bwd<__nv_bfloat16, true, false>(p, dLambda, dh0, workspace, workspace_bytes, st);
}

extern "C" BDLRU_API int bdlru_scan_cl_fwd(bdlru_view a, bdlru_view b, const float* h0, int64_t h0_bstride, bdlru_view h, int B,
                                 int T, int C, int dtype, void* stream) {
  int rc = check_common(B, T, C, dtype);
  if (rc) return rc;
  const int es = dtype == BDLRU_F32 ? 4 : 2;
  CHECK_VIEW("a", a, es, true); CHECK_VIEW("b", b, es, true); CHECK_VIEW("h", h, es, true);
  BDLRU_REQUIRE(h0_bstride == 0 || h0_bstride == C, "h0_bstride must be 0 or C");
  GScanParams p{};
  p.x = mk(b); p.r = mk(a); p.h = mk(h); p.h0 = h0; p.h0_bs = h0_bstride; p.B = B; p.T = T; p.C = C;
  set_wide(p);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dtype == BDLRU_F32 ? launch_fwd<float, false, false>(p, st) : launch_fwd<__nv_bfloat16, false, false>(p, st);
}

extern "C" BDLRU_API int bdlru_scan_cl_bwd(bdlru_view a, const float* h0, int64_t h0_bstride, bdlru_view h, bdlru_view grad,
                                 bdlru_view da, bdlru_view db, float* dh0, void* workspace, size_t workspace_bytes,
                                 int B, int T, int C, int dtype, void* stream) {
  int rc = check_common(B, T, C, dtype);
  if (rc) return rc;
  const int es = dtype == BDLRU_F32 ? 4 : 2;
  CHECK_VIEW("a", a, es, true); CHECK_VIEW("h", h, es, true); CHECK_VIEW("grad", grad, es, true);
  CHECK_VIEW("da", da, es, true); CHECK_VIEW("db", db, es, true);
  BDLRU_REQUIRE(h0_bstride == 0 || h0_bstride == C, "h0_bstride must be 0 or C");
  GScanParams p{};
  p.r = mk(a); p.h = mk(h); p.g = mk(grad); p.dx = mk(db); p.dr = mk(da);
  p.h0 = h0; p.h0_bs = h0_bstride; p.B = B; p.T = T; p.C = C;
  set_wide(p);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dtype == BDLRU_F32
             ? launch_bwd<float, false, false>(p, nullptr, dh0, workspace, workspace_bytes, st)
             : launch_bwd<__nv_bfloat16, false, false>(p, nullptr, dh0, workspace, workspace_bytes, st);
}
