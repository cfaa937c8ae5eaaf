// VERDICT r1 row N1 — the core of GatedRecurrentLayer.forward (RecBLR.py:182-206 without the two projections) as ONE
// tcgen05 kernel for inference:   y = silu(z) * BD-LRU( x' = silu(conv(x)),  (r | i) = W_g x' + b_g )
// x' and the [B, T, 2C] gate pre-activations never touch HBM: the kernel reads xz [B, T, 2C] once and writes y [B, T, C].
//
// TRANSPOSED mapping: TMEM lanes = channels, TMEM columns = time.
//   r|i^T [2C, Tc] = W_g [2C, C] * x'^T :  A = W_g rows, resident in TMEM for the whole kernel (packed bf16 pairs, written
//   once per CTA with tcgen05.st); B = the x' tile [Tc time rows, C channels] in shared memory — exactly the K-major,
//   128-byte-swizzled layout in which TMA delivers the rows of xz, so the tile is loaded by TMA and convolved IN PLACE.
// Every compute thread owns ONE channel: it walks the tile's rows for the depthwise causal conv (3-value register halo
// carried across tiles), later receives the 32 consecutive time steps of its channel's r and i from its TMEM lane with one
// tcgen05.ld each, and runs gate math + the recurrence h = a h + b' sequentially in registers — no shuffles, no chunk
// aggregates, no block barriers.  A CTA processes one batch row at a time (persistent over rows), 64 time steps per tile,
// a 3-deep TMA ring of xz tiles; the conv of tile k+1 runs while the tensor core works on tile k.
// TMEM: 128 columns of weights + ONE accumulator stage of (r | i) x 64 = 128 columns = 256 of 512 => two CTAs per SM (a
// second accumulator stage would need 384 -> 512 columns and halve the resident compute warps, which are the bottleneck).
//
// Scope: C = 128 (the reference's hidden_size 64 x expand 2), bf16 activations, inference only (the training path keeps
// the separate kernels, whose backward needs x' and r|i saved).  DESIGN.md §7.1 discusses D = 128 and the backward.
#include <cuda.h>

#include "common.cuh"
#include "tc05.cuh"
#include "tma_host.cuh"

namespace bdlru {

constexpr int kFC = 128;        // channels (TMEM lanes)
constexpr int kFTc = 64;        // time steps per tile (MMA N)
constexpr int kFStages = 3;     // xz tiles in flight
constexpr uint32_t kFSlab = kFTc * 128;          // bytes of one [Tc rows x 64 channels] swizzled slab
constexpr uint32_t kFStageB = 4 * kFSlab;        // x (2 slabs) + z (2 slabs)

struct FusedCoreParams {
  const float* conv_w;   // [C, 4] (null: no conv, x' = x)
  const float* conv_b;   // [C]
  const void* gates_w;   // [2C, C] bf16 row-major
  const float* gates_b;  // [2C]
  const float* lambda;   // [C]
  const float* h0;       // [C] or null
  void* y;               // [B, T, C] bf16
  int B, T, n_tiles;
};

// byte offset of element (row t, channel c) of a [Tc x 128-channel] operand made of two 64-channel slabs in the
// 128-byte swizzle TMA writes: 16-byte chunk index XOR (row mod 8)
__device__ __forceinline__ uint32_t sw_off(int t, int c) {
  return (uint32_t)(c >> 6) * kFSlab + (uint32_t)t * 128u + ((((uint32_t)(c & 63) >> 3) ^ ((uint32_t)t & 7u)) << 4) +
         ((uint32_t)(c & 7) << 1);
}
__device__ __forceinline__ float ld_bf16(const uint8_t* p) {
  return __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(p)) << 16);
}

template <bool USE_CONV>
__global__ void __launch_bounds__(192, 2) fused_core_fwd_kernel(const __grid_constant__ CUtensorMap tmXZ,
                                                                const FusedCoreParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kFStages * kFStageB);
  uint64_t* full = bars;                       // TMA landed                       (tx)
  uint64_t* conv_done = full + kFStages;       // x' written in place              (4 warps)
  uint64_t* slot_empty = conv_done + kFStages; // tile fully consumed              (4 warps)
  uint64_t* acc_full = slot_empty + kFStages;  // MMAs of the tile complete        (commit)
  uint64_t* acc_empty = acc_full + 1;          // accumulator read out             (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tc::prefetch_tensormap(&tmXZ);
    for (int s = 0; s < kFStages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&conv_done[s], 4);
      tc::mbar_init(&slot_empty[s], 4);
    }
    tc::mbar_init(acc_full, 1);
    tc::mbar_init(acc_empty, 4);
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
  // TMEM columns: [0, 64) W_g rows of the r gates, [64, 128) of the i gates (packed pairs), [128, 256) the (r | i) x Tc
  // accumulator
  constexpr uint32_t kAccCol = 128;
  static_assert(kAccCol + 2 * kFTc <= 256, "TMEM budget: the allocation is 256 columns");

  // tiles owned by this CTA: batch rows blockIdx.x, + gridDim.x, ...; n_tiles per row
  const long my_rows = (p.B - (long)blockIdx.x + gridDim.x - 1) / gridDim.x;
  const long my_tiles = my_rows * p.n_tiles;

  if (warp < 4) {
    // ============================================================ compute threads: thread == channel == TMEM lane
    const int c = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    {  // this channel's two rows of W_g -> TMEM (A operand of every MMA of this CTA)
      const __nv_bfloat16* gw = reinterpret_cast<const __nv_bfloat16*>(p.gates_w);
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        const uint4* src = reinterpret_cast<const uint4*>(gw + (size_t)(blk * kFC + c) * kFC);
#pragma unroll
        for (int j = 0; j < kFC / 16; ++j) {
          const uint4 lo = src[2 * j], hi = src[2 * j + 1];
          const uint32_t wv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
          tc::tmem_st_32x32_x8(lane_addr + (uint32_t)(blk * 64 + j * 8), wv);
        }
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
    }
    // (the MMA warp may not read the weights before every compute warp has stored them: named barrier over the 4
    // compute warps + the MMA warp would do; the first conv_done hand-off below already orders them, because every
    // compute warp arrives on it only after its tcgen05.st has completed and been fenced)
    float cw[4] = {0.f, 0.f, 0.f, 1.f}, cb = 0.f;
    if (USE_CONV) {
#pragma unroll
      for (int j = 0; j < 4; ++j) cw[j] = p.conv_w[c * 4 + j];
      cb = p.conv_b[c];
    }
    const float br = p.gates_b[c], bi = p.gates_b[kFC + c];
    const float csp = softplus_acc(p.lambda[c]);
    const float h0 = p.h0 ? p.h0[c] : 0.f;
    __nv_bfloat16* yout = reinterpret_cast<__nv_bfloat16*>(p.y);

    float x1 = 0.f, x2 = 0.f, x3 = 0.f;   // conv halo: x_{t-1}, x_{t-2}, x_{t-3}
    float h = h0;

    auto conv_tile = [&](long k) {   // k-th tile of this CTA: convolve x in place (USE_CONV), publish the B operand
      const int s = (int)(k % kFStages);
      const uint32_t ph = (uint32_t)(k / kFStages) & 1u;
      const int tile = (int)(k % p.n_tiles);
      if (tile == 0) { x1 = x2 = x3 = 0.f; }
      tc::mbar_wait(&full[s], ph);
      if (USE_CONV) {
        uint8_t* sx = smem + (size_t)s * kFStageB;
        const int rows = min(kFTc, p.T - tile * kFTc);   // rows past the end of the sequence are never read back
#pragma unroll 1
        for (int tb = 0; tb < rows; tb += 16) {   // 16 rows at a time: all loads first, so they overlap
          float xin[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) xin[j] = ld_bf16(sx + sw_off(tb + j, c));
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x0 = xin[j];
            const float pre = fmaf(cw[3], x0, fmaf(cw[2], x1, fmaf(cw[1], x2, fmaf(cw[0], x3, cb))));
            x3 = x2; x2 = x1; x1 = x0;
            const float xc = pre * sigmoid_t<true>(pre);
            *reinterpret_cast<__nv_bfloat16*>(sx + sw_off(tb + j, c)) = __float2bfloat16_rn(xc);
          }
        }
        tc::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async proxy
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&conv_done[s]);
    };

    auto scan_tile = [&](long k) {
      const int s = (int)(k % kFStages);
      const uint32_t aph = (uint32_t)k & 1u;
      const int tile = (int)(k % p.n_tiles);
      const long b = (long)blockIdx.x + (k / p.n_tiles) * gridDim.x;
      const int t0 = tile * kFTc;
      if (tile == 0) h = h0;
      const int tmax = min(kFTc, p.T - t0);
      const uint8_t* sx = smem + (size_t)s * kFStageB;
      const uint8_t* sz = sx + 2 * kFSlab;
      tc::mbar_wait(acc_full, aph);
      tc::fence_after_sync();
      __nv_bfloat16* yrow = yout + ((size_t)b * p.T + t0) * kFC + c;
      const int n_half = (tmax + 31) >> 5;   // a tail tile (T = 200: 8 of 64 steps) skips the halves it does not need
#pragma unroll 1
      for (int half = 0; half < n_half; ++half) {
        uint32_t rr[32], ii[32];
        tc::tmem_ld_32x32(lane_addr + kAccCol + (uint32_t)(half * 32), rr);
        tc::tmem_ld_32x32(lane_addr + kAccCol + (uint32_t)(kFTc + half * 32), ii);
        tc::tmem_ld_wait();
        if (half == n_half - 1) {   // the accumulator is in registers: the MMAs of the next tile may overwrite it
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(acc_empty);
        }
        // Three phases over the 32 time steps so that only the recurrence itself is a dependent chain: (1) gate math for
        // every step (independent: the MUFU / FMA latencies overlap), leaving a_t in rr and b'_t in ii; (2) the serial
        // h_t = a_t h_{t-1} + b'_t (one FMA per step), leaving h_t in rr; (3) z-gate and store (independent again).
        float zg[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int t = half * 32 + j;
          const float xc = ld_bf16(sx + sw_off(t, c));
          const float zv = ld_bf16(sz + sw_off(t, c));
          const Gate g = gate_full<true>(csp, __uint_as_float(rr[j]) + br, __uint_as_float(ii[j]) + bi);
          rr[j] = __float_as_uint(g.a);
          ii[j] = __float_as_uint(g.q * g.si * xc);
          zg[j] = zv * sigmoid_t<true>(zv);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (half * 32 + j < tmax) h = fmaf(__uint_as_float(rr[j]), h, __uint_as_float(ii[j]));
          rr[j] = __float_as_uint(h);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int t = half * 32 + j;
          if (t < tmax) yrow[(size_t)t * kFC] = __float2bfloat16_rn(zg[j] * __uint_as_float(rr[j]));
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&slot_empty[s]);
    };

    // software pipeline: conv(k + 1) is issued before scan(k), so the MMAs of tile k run under the conv of tile k + 1
    if (my_tiles > 0) conv_tile(0);
    for (long k = 0; k < my_tiles; ++k) {
      if (k + 1 < my_tiles) conv_tile(k + 1);
      scan_tile(k);
    }
  } else if (warp == 4) {
    // ============================================================ TMA producer
    for (long k = 0; k < my_tiles; ++k) {
      const int s = (int)(k % kFStages);
      const uint32_t ph = (uint32_t)(k / kFStages) & 1u;
      tc::mbar_wait(&slot_empty[s], ph ^ 1u);
      if (tc::elect_one()) {
        const long b = (long)blockIdx.x + (k / p.n_tiles) * gridDim.x;
        const int row = (int)(b * p.T + (k % p.n_tiles) * kFTc);
        uint8_t* dst = smem + (size_t)s * kFStageB;
        tc::mbar_arrive_expect_tx(&full[s], kFStageB);
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) tc::tma_load_2d(dst + (size_t)sl * kFSlab, &tmXZ, &full[s], sl * 64, row);
      }
      __syncwarp();
    }
  } else {
    // ============================================================ MMA issuer
    constexpr uint32_t idesc = tc::idesc_bf16_f32(kFC, kFTc, 0, 0);
    for (long k = 0; k < my_tiles; ++k) {
      const int s = (int)(k % kFStages);
      const uint32_t ph = (uint32_t)(k / kFStages) & 1u;
      const uint32_t aph = (uint32_t)k & 1u;
      tc::mbar_wait(&conv_done[s], ph);
      tc::mbar_wait(acc_empty, aph ^ 1u);
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint64_t bd0 = tc::smem_desc_sw128(tc::smem_u32(smem + (size_t)s * kFStageB), 16, 1024);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const uint32_t d_tmem = tmem_base + kAccCol + (uint32_t)(blk * kFTc);
#pragma unroll
          for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              tc::umma_bf16_ts(d_tmem, tmem_base + (uint32_t)(blk * 64 + (sl * 4 + k4) * 8),
                               bd0 + (uint64_t)((uint32_t)sl * (kFSlab >> 4) + (uint32_t)k4 * 2), idesc,
                               (uint32_t)((sl | k4) != 0));
        }
        tc::umma_commit(acc_full);
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

}  // namespace bdlru

using namespace bdlru;

extern "C" BDLRU_API int bdlru_core_fwd_supported(int C, int dtype) { return C == kFC && dtype == BDLRU_BF16; }

extern "C" BDLRU_API int bdlru_core_fwd(const void* xz, const float* conv_w, const float* conv_b, const void* gates_w,
                                        const float* gates_b, const float* Lambda, const float* h0, void* y, int B, int T,
                                        int C, void* stream) {
  BDLRU_REQUIRE(xz && gates_w && gates_b && Lambda && y, "core_fwd: null pointer");
  BDLRU_REQUIRE(C == kFC, "core_fwd: C=%d (only C = %d is built)", C, kFC);
  BDLRU_REQUIRE(B >= 1 && T >= 1 && (long)B * T < (1L << 31), "core_fwd: bad sizes B=%d T=%d", B, T);
  BDLRU_REQUIRE((conv_w == nullptr) == (conv_b == nullptr), "core_fwd: conv_w and conv_b go together");
  BDLRU_REQUIRE(aligned(xz, 16) && aligned(gates_w, 16) && aligned(y, 2), "core_fwd: xz / gates_w must be 16-byte aligned");
  CUtensorMap tm;
  int rc = make_rows_map(&tm, xz, (long)B * T, 2 * C, kFTc);
  if (rc) return rc;
  FusedCoreParams p = {};
  p.conv_w = conv_w; p.conv_b = conv_b; p.gates_w = gates_w; p.gates_b = gates_b; p.lambda = Lambda; p.h0 = h0; p.y = y;
  p.B = B; p.T = T; p.n_tiles = (T + kFTc - 1) / kFTc;
  const size_t smem = 1024 + (size_t)kFStages * kFStageB + 256;
  int grid = 2 * sm_count();
  if (grid > B) grid = B;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (conv_w) {
    BDLRU_CUDA(cudaFuncSetAttribute(fused_core_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fused_core_fwd_kernel<true><<<grid, 192, smem, st>>>(tm, p);
  } else {
    BDLRU_CUDA(cudaFuncSetAttribute(fused_core_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fused_core_fwd_kernel<false><<<grid, 192, smem, st>>>(tm, p);
  }
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}
