// Optimizer step of the row-sharded item table (sharded.ShardedItemTable): torch.optim.Adam's update of the fp32 master
// rows FUSED with the refresh of the bf16 compute copy — one pass over (p, g, m, v) that also writes the bf16 rows the next
// step's input gather and tensor-core CE read.  The stock path costs a fused multi-tensor Adam (7 x 4 bytes per element)
// plus a separate cast pass (4 + 2): 34 bytes per element; this kernel moves 30 and is one launch.
//
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// (torch.optim.Adam with amsgrad = False, maximize = False; weight_decay adds wd * p to g first — L2, not AdamW).
#include "common.cuh"

namespace bdlru {

__global__ void __launch_bounds__(256) table_adam_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                         float4* __restrict__ m, float4* __restrict__ v,
                                                         uint2* __restrict__ p_bf16, long n4, float b1, float b2,
                                                         float step_size, float inv_bc2_sqrt, float eps, float wd) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ga[4] = {gg.x, gg.y, gg.z, gg.w};
    float ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gr = fmaf(wd, pa[e], ga[e]);
      ma[e] = fmaf(b1, ma[e], (1.0f - b1) * gr);
      va[e] = fmaf(b2, va[e], (1.0f - b2) * gr * gr);
      const float denom = fmaf(sqrtf(va[e]), inv_bc2_sqrt, eps);
      pa[e] = fmaf(-step_size, ma[e] / denom, pa[e]);
    }
    p[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    m[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    v[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (p_bf16) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(pa[0], pa[1]), hi = __floats2bfloat162_rn(pa[2], pa[3]);
      uint2 u;
      u.x = *reinterpret_cast<const uint32_t*>(&lo);
      u.y = *reinterpret_cast<const uint32_t*>(&hi);
      p_bf16[i] = u;
    }
  }
}

}  // namespace bdlru

using namespace bdlru;

extern "C" BDLRU_API int bdlru_table_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                               void* param_bf16, int64_t n, float lr, float beta1, float beta2, float eps,
                                               float weight_decay, int64_t step, void* stream) {
  BDLRU_REQUIRE(param && grad && exp_avg && exp_avg_sq, "table_adam_step: null pointer");
  BDLRU_REQUIRE(n >= 4 && n % 4 == 0, "table_adam_step: n=%ld must be a positive multiple of 4", (long)n);
  BDLRU_REQUIRE(step >= 1, "table_adam_step: step=%ld must be >= 1", (long)step);
  BDLRU_REQUIRE(aligned(param, 16) && aligned(grad, 16) && aligned(exp_avg, 16) && aligned(exp_avg_sq, 16) &&
                    aligned(param_bf16, 8),
                "table_adam_step: pointers must be 16-byte aligned (bf16 copy: 8)");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const long n4 = n / 4;
  long blocks = (n4 + 255) / 256;
  const long cap = (long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  table_adam_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(exp_avg),
      reinterpret_cast<float4*>(exp_avg_sq), reinterpret_cast<uint2*>(param_bf16), n4, beta1, beta2,
      (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), eps, weight_decay);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}
