// The whole first half of RecurrentLayer.forward for INFERENCE (RecBLR.py:140-142 + 170-206) as ONE tcgen05 kernel:
//     out = LayerNorm( W_out ( silu(z) * BD-LRU( silu(conv(x)), W_g . + b_g ) ) + X ),      (x | z) = W_in X
// It extends csrc/fused_core.cu by the two projections and the residual +
// LayerNorm epilogue (f1): the kernel reads the layer input X [B, T, D] once and writes the layer's post-LayerNorm
// output [B, T, D]; xz, x', r|i, h and y never touch HBM (2 D-wide units per layer instead of ~14 C-wide ones).
//
// Mapping (same transposed idea): TMEM lanes = channels, columns = time, 32 time steps per tile.
//   MMA1  xz^T [2C, 32] = W_in [2C, D] * X^T      A = W_in in shared memory (K-major, TMA-swizzled), B = the X tile as TMA
//                                                 delivers it ([32 rows, D] K-major)
//   MMA2  ri^T [2C, 32] = W_g [2C, C] * x'^T      A = W_g resident in TMEM, B = the x' tile the compute threads write
//   MMA3  o^T  [D, 32]  = W_out [D, C] * y^T      A = W_out in shared memory (rows D..127 zero-filled by TMA), B = the y tile
// Every compute thread owns one channel c (= TMEM lane): conv + SiLU over its 32 steps straight out of its TMEM row, gate
// math + recurrence + z-gate in registers (three phases, only h = a h + b' is a dependent chain), x' and y written to
// shared memory in the 128-byte-swizzled K-major layout the tensor core expects.  The LayerNorm runs on lanes 0..D-1
// (two warps): the sums over d for the 32 time steps are a warp transpose-reduce (31 shuffles) + one exchange between the
// two warps through shared memory.
// TMEM (256 columns => two CTAs per SM): [0,128) W_g; [128,192) one accumulator used first for xz then for r|i;
// [192,224) o^T.  The three MMA phases of a tile are sequential; the two resident CTAs fill each other's hand-off gaps.
//
// Scope: d_model 64, expand 2 (C = 128), conv width 4, bf16 activations, inference.
#include <cuda.h>

#include "common.cuh"
#include "tc05.cuh"
#include "tma_host.cuh"

namespace bdlru {

constexpr int kLC = 128;   // channels
constexpr int kLD = 64;    // d_model
constexpr int kLT = 32;    // time steps per tile
constexpr int kLStages = 3;
constexpr uint32_t kLSlab32 = kLT * 128;     // [32 rows x 64 ch] slab
constexpr uint32_t kLSlab128 = 128 * 128;    // [128 rows x 64 ch] slab (weights)

struct FusedLayerParams {
  const float* conv_w;   // [C, 4] or null
  const float* conv_b;   // [C]
  const void* gates_w;   // [2C, C] bf16
  const float* gates_b;  // [2C]
  const float* lambda;   // [C]
  const float* h0;       // [C] or null
  const float* ln_g;     // [D]
  const float* ln_b;     // [D]
  void* out;             // [B, T, D] bf16
  float eps;
  int B, T, n_tiles;
};

// (row t, channel c) inside an operand made of 64-channel slabs of `slab_bytes` each, 128-byte rows, TMA's 128-byte swizzle
__device__ __forceinline__ uint32_t swz(int t, int c, uint32_t slab_bytes) {
  return (uint32_t)(c >> 6) * slab_bytes + (uint32_t)t * 128u + ((((uint32_t)(c & 63) >> 3) ^ ((uint32_t)t & 7u)) << 4) +
         ((uint32_t)(c & 7) << 1);
}
__device__ __forceinline__ float ldb(const uint8_t* p) {
  return __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(p)) << 16);
}
__device__ __forceinline__ float rbf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// v[t] summed over the 32 lanes of the warp, result for time step t left in lane t (31 shuffles instead of 32 x 5)
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float keep = up ? v[i + off] : v[i];
      const float send = up ? v[i] : v[i + off];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <bool USE_CONV>
__global__ void __launch_bounds__(192, 2)
fused_layer_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWin,
                       const __grid_constant__ CUtensorMap tmWout, const FusedLayerParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sWin = smem;                               // 2 blocks x 16 KB
  uint8_t* sWout = sWin + 2 * kLSlab128;              // 2 K-slabs x 16 KB
  uint8_t* sX = sWout + 2 * kLSlab128;                // kLStages x 4 KB
  uint8_t* sXp = sX + kLStages * kLSlab32;            // 2 slabs x 4 KB
  uint8_t* sY = sXp + 2 * kLSlab32;                   // 2 slabs x 4 KB
  float* ln_part = reinterpret_cast<float*>(sY + 2 * kLSlab32);   // [2 warps][32][2]
  float* ln_stat = ln_part + 2 * 32 * 2;                           // [32][2] mean, rstd
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_stat + 32 * 2);
  uint64_t* w_full = bars;                 // weights landed
  uint64_t* x_full = w_full + 1;           // [kLStages]
  uint64_t* x_empty = x_full + kLStages;   // [kLStages] (4 warps)
  uint64_t* xz_full = x_empty + kLStages;  // MMA1 done
  uint64_t* xp_ready = xz_full + 1;        // x' tile written, xz accumulator read out (4 warps)
  uint64_t* ri_full = xp_ready + 1;        // MMA2 done
  uint64_t* acc_free = ri_full + 1;        // r|i accumulator read out (4 warps)
  uint64_t* y_ready = acc_free + 1;        // y tile written (4 warps)
  uint64_t* out_full = y_ready + 1;        // MMA3 done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tc::prefetch_tensormap(&tmX);
    tc::prefetch_tensormap(&tmWin);
    tc::prefetch_tensormap(&tmWout);
    tc::mbar_init(w_full, 1);
    for (int s = 0; s < kLStages; ++s) {
      tc::mbar_init(&x_full[s], 1);
      tc::mbar_init(&x_empty[s], 4);
    }
    tc::mbar_init(xz_full, 1);
    tc::mbar_init(xp_ready, 4);
    tc::mbar_init(ri_full, 1);
    tc::mbar_init(acc_free, 4);
    tc::mbar_init(y_ready, 4);
    tc::mbar_init(out_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 5) {
    tc::tmem_alloc(tmem_slot, 256);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kAccA = 128, kAccO = 192;
  static_assert(kAccO + kLT <= 256, "TMEM budget: the allocation is 256 columns");

  const long my_rows = (p.B - (long)blockIdx.x + gridDim.x - 1) / gridDim.x;
  const long my_tiles = my_rows * p.n_tiles;

  if (warp < 4) {
    // ============================================================ compute threads: thread == channel == TMEM lane
    const int c = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    {  // W_g rows of this channel -> TMEM
      const __nv_bfloat16* gw = reinterpret_cast<const __nv_bfloat16*>(p.gates_w);
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        const uint4* src = reinterpret_cast<const uint4*>(gw + (size_t)(blk * kLC + c) * kLC);
#pragma unroll
        for (int j = 0; j < kLC / 16; ++j) {
          const uint4 lo = src[2 * j], hi = src[2 * j + 1];
          const uint32_t wv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
          tc::tmem_st_32x32_x8(lane_addr + (uint32_t)(blk * 64 + j * 8), wv);
        }
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
    }
    float cw[4] = {0.f, 0.f, 0.f, 1.f}, cb = 0.f;
    if (USE_CONV) {
#pragma unroll
      for (int j = 0; j < 4; ++j) cw[j] = p.conv_w[c * 4 + j];
      cb = p.conv_b[c];
    }
    const float br = p.gates_b[c], bi = p.gates_b[kLC + c];
    const float csp = softplus_acc(p.lambda[c]);
    const float h0 = p.h0 ? p.h0[c] : 0.f;
    const bool ln_thread = c < kLD;                       // lanes 0..D-1 own one output feature d = c
    const float lg = ln_thread ? p.ln_g[c] : 0.f, lb = ln_thread ? p.ln_b[c] : 0.f;
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
    float x1 = 0.f, x2 = 0.f, x3 = 0.f, h = h0;

    for (long k = 0; k < my_tiles; ++k) {
      const int s = (int)(k % kLStages);
      const uint32_t ph = (uint32_t)k & 1u;
      const int tile = (int)(k % p.n_tiles);
      const long b = (long)blockIdx.x + (k / p.n_tiles) * gridDim.x;
      const int t0 = tile * kLT;
      const int tmax = min(kLT, p.T - t0);
      if (tile == 0) { x1 = x2 = x3 = 0.f; h = h0; }

      // ---- phase A: x, z out of the xz accumulator; conv + SiLU; x' tile -> shared memory
      float xcr[32], zg[32];
      {
        uint32_t xa[32], za[32];
        tc::mbar_wait(xz_full, ph);
        tc::fence_after_sync();
        tc::tmem_ld_32x32(lane_addr + kAccA, xa);
        tc::tmem_ld_32x32(lane_addr + kAccA + kLT, za);
        tc::tmem_ld_wait();
        tc::fence_before_sync();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x0 = rbf(__uint_as_float(xa[j]));    // the in-projection's output is a bf16 activation
          float xc = x0;
          if (USE_CONV) {
            const float pre = fmaf(cw[3], x0, fmaf(cw[2], x1, fmaf(cw[1], x2, fmaf(cw[0], x3, cb))));
            x3 = x2; x2 = x1; x1 = x0;
            xc = pre * sigmoid_t<true>(pre);
          }
          const __nv_bfloat16 xb = __float2bfloat16_rn(xc);
          *reinterpret_cast<__nv_bfloat16*>(sXp + swz(j, c, kLSlab32)) = xb;
          xcr[j] = __bfloat162float(xb);
          const float zv = rbf(__uint_as_float(za[j]));
          zg[j] = zv * sigmoid_t<true>(zv);
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(xp_ready);
      }

      // ---- phase B: r, i out of the accumulator; gate math, recurrence, z-gate; y tile -> shared memory
      {
        uint32_t rr[32], ii[32];
        tc::mbar_wait(ri_full, ph);
        tc::fence_after_sync();
        tc::tmem_ld_32x32(lane_addr + kAccA, rr);
        tc::tmem_ld_32x32(lane_addr + kAccA + kLT, ii);
        tc::tmem_ld_wait();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(acc_free);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const Gate g = gate_full<true>(csp, __uint_as_float(rr[j]) + br, __uint_as_float(ii[j]) + bi);
          rr[j] = __float_as_uint(g.a);
          ii[j] = __float_as_uint(g.q * g.si * xcr[j]);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < tmax) h = fmaf(__uint_as_float(rr[j]), h, __uint_as_float(ii[j]));
          *reinterpret_cast<__nv_bfloat16*>(sY + swz(j, c, kLSlab32)) = __float2bfloat16_rn(zg[j] * h);
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(y_ready);
      }

      // ---- phase C: o^T out of TMEM (lanes 0..D-1), + residual, LayerNorm over d, store
      tc::mbar_wait(out_full, ph);
      tc::fence_after_sync();
      if (warp < 2) {
        uint32_t oa[32];
        tc::tmem_ld_32x32(lane_addr + kAccO, oa);
        tc::tmem_ld_wait();
        tc::fence_before_sync();
        float v[32], sq[32], sm[32];
        const uint8_t* sx = sX + (size_t)s * kLSlab32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = rbf(__uint_as_float(oa[j])) + ldb(sx + swz(j, c, kLSlab32));   // out-projection output is a bf16 activation
          sm[j] = v[j];
          sq[j] = v[j] * v[j];
        }
        const float s1 = transpose_reduce32(sm, lane);   // lane t: sum over this warp's 32 features of step t
        const float s2 = transpose_reduce32(sq, lane);
        ln_part[(warp * 32 + lane) * 2 + 0] = s1;
        ln_part[(warp * 32 + lane) * 2 + 1] = s2;
        asm volatile("bar.sync 1, 64;" ::: "memory");
        if (warp == 0) {
          const float t1 = ln_part[lane * 2] + ln_part[(32 + lane) * 2];
          const float t2 = ln_part[lane * 2 + 1] + ln_part[(32 + lane) * 2 + 1];
          const float mean = t1 * (1.0f / kLD);
          const float var = fmaxf(t2 * (1.0f / kLD) - mean * mean, 0.f);
          ln_stat[lane * 2] = mean;
          ln_stat[lane * 2 + 1] = rsqrtf(var + p.eps);
        }
        asm volatile("bar.sync 1, 64;" ::: "memory");
        __nv_bfloat16* orow = outp + ((size_t)b * p.T + t0) * kLD + c;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < tmax) {
            const float2 st = *reinterpret_cast<const float2*>(ln_stat + j * 2);
            orow[(size_t)j * kLD] = __float2bfloat16_rn(fmaf((v[j] - st.x) * st.y, lg, lb));
          }
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&x_empty[s]);
    }
  } else if (warp == 4) {
    // ============================================================ TMA producer: weights once, then the X tiles
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(w_full, 4 * kLSlab128);
      tc::tma_load_2d(sWin, &tmWin, w_full, 0, 0);                      // W_in rows 0..127   (x)
      tc::tma_load_2d(sWin + kLSlab128, &tmWin, w_full, 0, 128);        // W_in rows 128..255 (z)
      tc::tma_load_2d(sWout, &tmWout, w_full, 0, 0);                    // W_out, K 0..63   (rows >= D: zero fill)
      tc::tma_load_2d(sWout + kLSlab128, &tmWout, w_full, 64, 0);       // W_out, K 64..127
    }
    __syncwarp();
    for (long k = 0; k < my_tiles; ++k) {
      const int s = (int)(k % kLStages);
      const uint32_t ph = (uint32_t)(k / kLStages) & 1u;
      tc::mbar_wait(&x_empty[s], ph ^ 1u);
      if (tc::elect_one()) {
        const long b = (long)blockIdx.x + (k / p.n_tiles) * gridDim.x;
        const int row = (int)(b * p.T + (k % p.n_tiles) * kLT);
        tc::mbar_arrive_expect_tx(&x_full[s], kLSlab32);
        tc::tma_load_2d(sX + (size_t)s * kLSlab32, &tmX, &x_full[s], 0, row);
      }
      __syncwarp();
    }
  } else {
    // ============================================================ MMA issuer
    constexpr uint32_t idesc = tc::idesc_bf16_f32(128, kLT, 0, 0);
    tc::mbar_wait(w_full, 0);
    for (long k = 0; k < my_tiles; ++k) {
      const int s = (int)(k % kLStages);
      const uint32_t xph = (uint32_t)(k / kLStages) & 1u;
      const uint32_t ph = (uint32_t)k & 1u;
      // MMA1: xz^T = W_in X^T  (needs the X tile and the accumulator free of the previous tile's r|i)
      tc::mbar_wait(&x_full[s], xph);
      tc::mbar_wait(acc_free, ph ^ 1u);
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint64_t bx = tc::smem_desc_sw128(tc::smem_u32(sX + (size_t)s * kLSlab32), 16, 1024);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const uint64_t aw = tc::smem_desc_sw128(tc::smem_u32(sWin + (size_t)blk * kLSlab128), 16, 1024);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            tc::umma_bf16(tmem_base + kAccA + (uint32_t)(blk * kLT), aw + (uint64_t)(k4 * 2), bx + (uint64_t)(k4 * 2), idesc,
                          (uint32_t)(k4 != 0));
        }
        tc::umma_commit(xz_full);
      }
      __syncwarp();
      // MMA2: ri^T = W_g x'^T
      tc::mbar_wait(xp_ready, ph);
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint64_t bp = tc::smem_desc_sw128(tc::smem_u32(sXp), 16, 1024);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk)
#pragma unroll
          for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              tc::umma_bf16_ts(tmem_base + kAccA + (uint32_t)(blk * kLT), tmem_base + (uint32_t)(blk * 64 + (sl * 4 + k4) * 8),
                               bp + (uint64_t)((uint32_t)sl * (kLSlab32 >> 4) + (uint32_t)k4 * 2), idesc,
                               (uint32_t)((sl | k4) != 0));
        tc::umma_commit(ri_full);
      }
      __syncwarp();
      // MMA3: o^T = W_out y^T
      tc::mbar_wait(y_ready, ph);
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint64_t by = tc::smem_desc_sw128(tc::smem_u32(sY), 16, 1024);
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const uint64_t ao = tc::smem_desc_sw128(tc::smem_u32(sWout + (size_t)sl * kLSlab128), 16, 1024);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            tc::umma_bf16(tmem_base + kAccO, ao + (uint64_t)(k4 * 2), by + (uint64_t)((uint32_t)sl * (kLSlab32 >> 4) + k4 * 2),
                          idesc, (uint32_t)((sl | k4) != 0));
        }
        tc::umma_commit(out_full);
      }
      __syncwarp();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 256);
  }
}

// [rows, cols] bf16 row-major -> boxes of box_rows x 64 columns (make_rows_map with explicit names)
static int rows_map(CUtensorMap* m, const void* base, long rows, int cols, int box_rows) {
  return make_rows_map(m, base, rows, cols, box_rows);
}

}  // namespace bdlru

using namespace bdlru;

extern "C" BDLRU_API int bdlru_layer_fwd_supported(int d_model, int C, int conv_width, int dtype) {
  return d_model == kLD && C == kLC && conv_width == 4 && dtype == BDLRU_BF16;
}

extern "C" BDLRU_API int bdlru_layer_fwd(const void* x, const void* in_w, const float* conv_w, const float* conv_b,
                                         const void* gates_w, const float* gates_b, const float* Lambda, const float* h0,
                                         const void* out_w, const float* ln_gamma, const float* ln_beta, float eps, void* out,
                                         int B, int T, int d_model, int C, void* stream) {
  BDLRU_REQUIRE(x && in_w && gates_w && gates_b && Lambda && out_w && ln_gamma && ln_beta && out, "layer_fwd: null pointer");
  BDLRU_REQUIRE(d_model == kLD && C == kLC, "layer_fwd: d_model=%d C=%d (built for %d / %d)", d_model, C, kLD, kLC);
  BDLRU_REQUIRE(B >= 1 && T >= 1 && (long)B * T < (1L << 31), "layer_fwd: bad sizes B=%d T=%d", B, T);
  BDLRU_REQUIRE((conv_w == nullptr) == (conv_b == nullptr), "layer_fwd: conv_w and conv_b go together");
  BDLRU_REQUIRE(aligned(x, 16) && aligned(in_w, 16) && aligned(gates_w, 16) && aligned(out_w, 16),
                "layer_fwd: x / in_w / gates_w / out_w must be 16-byte aligned");
  CUtensorMap tx, twi, two;
  int rc;
  if ((rc = rows_map(&tx, x, (long)B * T, kLD, kLT))) return rc;
  if ((rc = rows_map(&twi, in_w, 2 * kLC, kLD, 128))) return rc;
  if ((rc = rows_map(&two, out_w, kLD, kLC, 128))) return rc;
  FusedLayerParams p = {};
  p.conv_w = conv_w; p.conv_b = conv_b; p.gates_w = gates_w; p.gates_b = gates_b; p.lambda = Lambda; p.h0 = h0;
  p.ln_g = ln_gamma; p.ln_b = ln_beta; p.out = out; p.eps = eps;
  p.B = B; p.T = T; p.n_tiles = (T + kLT - 1) / kLT;
  const size_t smem = 1024 + 4 * (size_t)kLSlab128 + (size_t)(kLStages + 4) * kLSlab32 + 1024;
  int grid = 2 * sm_count();
  if (grid > B) grid = B;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (conv_w) {
    BDLRU_CUDA(cudaFuncSetAttribute(fused_layer_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fused_layer_fwd_kernel<true><<<grid, 192, smem, st>>>(tx, twi, two, p);
  } else {
    BDLRU_CUDA(cudaFuncSetAttribute(fused_layer_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fused_layer_fwd_kernel<false><<<grid, 192, smem, st>>>(tx, twi, two, p);
  }
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}
