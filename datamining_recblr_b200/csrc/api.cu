// Library-wide plumbing of the C ABI (include/bdlru.h): version, thread-local error text, launch counter.
#include <stdarg.h>

#include "common.cuh"

namespace bdlru {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();  // no device (build container): sizes are computed for a B200
    return 148;
  }
  cached = n;
  return n;
}

}  // namespace bdlru

extern "C" BDLRU_API int bdlru_version(void) { return 2; }
extern "C" BDLRU_API const char* bdlru_last_error(void) { return bdlru::g_err; }
extern "C" BDLRU_API uint64_t bdlru_launch_count(void) { return bdlru::g_launches.load(); }
