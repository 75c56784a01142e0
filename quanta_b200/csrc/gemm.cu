// Placeholder translation unit for the tcgen05 dequant-GEMM (rows G1-G3); filled in next.
#include "common.cuh"

namespace quanta {
size_t gemm_workspace_bytes(int64_t, int64_t) { return 256; }
size_t int8_outlier_workspace_bytes(int64_t, int64_t) { return 256; }
}  // namespace quanta

extern "C" int quanta_gemm_wna16(const void*, int, const uint8_t*, int, const float*, const float*, int64_t,
                                 const void*, void*, int64_t, int64_t, int64_t, void*, size_t, void*) {
    return QUANTA_EUNSUPPORTED;
}
extern "C" int quanta_int8_outlier_matmul(const void*, int, const int8_t*, const float*, float, const void*, void*,
                                          int64_t, int64_t, int64_t, void*, size_t, void*) {
    return QUANTA_EUNSUPPORTED;
}
