// Full-sort scoring / full-softmax CE entry points (include/bdlru.h).  Placeholder bodies until the tcgen05
// kernels land: every call reports BDLRU_ERR_UNSUPPORTED (never a silent fallback).
#include "common.cuh"

using namespace bdlru;

#define UNSUPPORTED(name)                                              \
  do {                                                                 \
    set_error(name ": tcgen05 kernel not built into this library yet"); \
    return BDLRU_ERR_UNSUPPORTED;                                      \
  } while (0)

extern "C" BDLRU_API int bdlru_fullsort_available(void) { return 0; }
extern "C" BDLRU_API size_t bdlru_fullsort_topk_workspace_bytes(int64_t, int64_t, int, int) { return 0; }
extern "C" BDLRU_API int bdlru_fullsort_topk(const void*, const void*, int64_t, int64_t, int, int, int64_t, int64_t,
                                             float*, int32_t*, void*, size_t, void*) {
  UNSUPPORTED("bdlru_fullsort_topk");
}
extern "C" BDLRU_API int bdlru_topk_merge(const float*, const int32_t*, int64_t, int, int, float*, int32_t*, void*) {
  UNSUPPORTED("bdlru_topk_merge");
}
extern "C" BDLRU_API size_t bdlru_fullsort_ce_workspace_bytes(int64_t, int64_t, int) { return 0; }
extern "C" BDLRU_API int bdlru_fullsort_ce_fwd(const void*, const void*, const int64_t*, int64_t, int64_t, int, int64_t,
                                               float*, float*, float*, void*, size_t, void*) {
  UNSUPPORTED("bdlru_fullsort_ce_fwd");
}
extern "C" BDLRU_API int bdlru_fullsort_ce_bwd(const void*, const void*, const int64_t*, const float*, float, int64_t,
                                               int64_t, int, int64_t, float*, float*, void*, size_t, void*) {
  UNSUPPORTED("bdlru_fullsort_ce_bwd");
}
