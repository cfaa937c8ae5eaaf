// Micro-benchmark (tuning aid, not product): issue rate of tcgen05.mma kind::f16 on one SM for several operand
// sources / shapes.  One CTA per SM, one elected lane issues REPS x 8 MMAs back to back, clock64 around them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I datamining_recblr_b200/csrc tools/mma_rate.cu -o tools/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc05.cuh"

template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = slot;
  if (warp == 1) {
    constexpr uint32_t idesc = tc::idesc_bf16_f32(128, N, 0, 0);
    long long t0 = 0, t1 = 0;
    if (tc::elect_one()) {
      const uint64_t ad = tc::smem_desc_sw128(tc::smem_u32(smem), 16, 1024);
      const uint64_t bd = tc::smem_desc_sw128(tc::smem_u32(smem + 32768), 16, 1024);
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t d = tb + 128 + (uint32_t)((k / 4) % NACC) * N;   // accumulators after 128 columns of "A"
          if (TS) tc::umma_bf16_ts(d, tb + (k & 3) * 8, bd + (uint64_t)((k & 3) * 2), idesc, 1);
          else tc::umma_bf16(d, ad + (uint64_t)((k & 3) * 2), bd + (uint64_t)((k & 3) * 2), idesc, 1);
        }
      }
      tc::umma_commit(&bar);
      tc::mbar_wait(&bar, 0);
      t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc::fence_after_sync(); tc::tmem_dealloc(tb, 512); }
}

template <int N, bool TS, int NACC>
void run(const char* name, int grid) {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  const int reps = 2048;
  cudaFuncSetAttribute(rate_kernel<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  rate_kernel<N, TS, NACC><<<grid, 128, 100 * 1024>>>(d, 64);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  rate_kernel<N, TS, NACC><<<grid, 128, 100 * 1024>>>(d, reps);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long h[148];
  cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
  double cyc = (double)h[0] / (reps * 8.0);
  double flops = 2.0 * 128 * N * 16 * reps * 8.0 * grid;
  printf("%-28s grid %3d: %7.1f cycles/MMA (ideal %5.1f)  %8.1f TFLOP/s  (%.3f ms) %s\n", name, grid, cyc,
         128.0 * N * 16 * 2 / 8192.0, flops / (ms * 1e-3) / 1e12, ms, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<64, false, 1>("SS M128 N64", grid);
    run<96, false, 1>("SS M128 N96", grid);
    run<128, false, 1>("SS M128 N128", grid);
    run<128, false, 2>("SS M128 N128 2acc", grid);
    run<256, false, 1>("SS M128 N256", grid);
    run<96, true, 1>("TS M128 N96", grid);
    run<96, true, 2>("TS M128 N96 2acc", grid);
    run<128, true, 1>("TS M128 N128", grid);
    run<256, true, 1>("TS M128 N256", grid);
  }
  return 0;
}
