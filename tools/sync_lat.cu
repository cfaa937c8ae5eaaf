// Micro-benchmark (tuning aid): latencies of the synchronisation hops used by the full-sort kernel, on one SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "tc05.cuh"

__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// mode 0: commit (no MMA) -> same thread waits.  mode 1: one MMA N=96 TS + commit -> wait.   mode 2: 8 MMAs + commit -> wait
// mode 3: ping-pong warp1 <-> warp2 with plain mbarrier arrive / try_wait.  mode 4: same with test_wait spin.
// mode 5: warp1 commit(after 8 MMAs) -> warp2 try_wait -> tcgen05.ld x3 + wait -> arrive -> warp1 try_wait  (the real loop)
__global__ void __launch_bounds__(128, 1) lat_kernel(long long* out, int mode, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(&bar[0], 1); tc::mbar_init(&bar[1], 1); tc::fence_barrier_init(); }
  if (warp == 0) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = slot;
  constexpr uint32_t idesc = tc::idesc_bf16_f32(128, 96, 0, 0);
  const uint64_t bd = tc::smem_desc_sw128(tc::smem_u32(smem), 16, 1024);
  if (warp == 1) {
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t ph = r & 1;
      if (mode <= 2 || mode == 5) {
        if (tc::elect_one()) {
          const int n = mode == 0 ? 0 : (mode == 1 ? 1 : 8);
          for (int k = 0; k < n; ++k) tc::umma_bf16_ts(tb + 128, tb + (k & 3) * 8, bd + (uint64_t)((k & 3) * 2), idesc, 1);
          tc::umma_commit(&bar[0]);
        }
        __syncwarp();
        if (mode == 5) tc::mbar_wait(&bar[1], ph); else tc::mbar_wait(&bar[0], ph);
      } else {
        if (tc::elect_one()) tc::mbar_arrive(&bar[0]);
        __syncwarp();
        if (mode == 3) tc::mbar_wait(&bar[1], ph); else while (!test_wait(&bar[1], ph)) {}
      }
    }
    long long t1 = clock64();
    if (threadIdx.x == 32) out[blockIdx.x] = (t1 - t0) / reps;
  } else if (warp == 2 && mode >= 3) {
    for (int r = 0; r < reps; ++r) {
      const uint32_t ph = r & 1;
      if (mode == 4) while (!test_wait(&bar[0], ph)) {} else tc::mbar_wait(&bar[0], ph);
      if (mode == 5) {
        tc::fence_after_sync();
        uint32_t v[32];
        for (int c = 0; c < 3; ++c) tc::tmem_ld_32x32(tb + (64u << 16) + 128 + c * 32, v);
        tc::tmem_ld_wait();
        tc::fence_before_sync();
      }
      __syncwarp();
      if (tc::elect_one()) tc::mbar_arrive(&bar[1]);
      __syncwarp();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc::fence_after_sync(); tc::tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char* names[] = {"commit(no MMA) -> wait", "1 MMA + commit -> wait", "8 MMA (384 cyc) + commit -> wait",
                         "ping-pong arrive/try_wait RTT", "ping-pong arrive/test_wait RTT", "8 MMA+commit -> ld x3 -> arrive RTT"};
  for (int mode = 0; mode < 6; ++mode) {
    lat_kernel<<<1, 128, 64 * 1024>>>(d, mode, 2000);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-40s %6lld cycles/iter  %s\n", names[mode], h, cudaGetErrorString(e));
  }
  return 0;
}
