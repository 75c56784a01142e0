// Miscellaneous C-ABI entry points: version, error strings, workspace sizes.
#include "common.cuh"

namespace quanta {
size_t quantize_workspace_bytes(int64_t cols);
size_t gemm_workspace_bytes(int64_t M, int64_t N);
size_t int8_outlier_workspace_bytes(int64_t M, int64_t N);
size_t base_workspace_bytes(int64_t groups);
}  // namespace quanta

extern "C" int quanta_abi_version(void) { return QUANTA_B200_ABI_VERSION; }

extern "C" const char* quanta_error_string(int code) {
    switch (code) {
        case QUANTA_OK: return "ok";
        case QUANTA_EINVAL: return "invalid argument";
        case QUANTA_EUNSUPPORTED: return "unsupported combination of arguments";
        case QUANTA_EWORKSPACE: return "workspace too small (see quanta_workspace_bytes)";
        case QUANTA_EDRIVER: return "CUDA driver entry point cuTensorMapEncodeTiled unavailable or failed";
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown error";
}

extern "C" size_t quanta_workspace_bytes(int op, int64_t rows, int64_t cols) {
    switch (op) {
        case QUANTA_OP_QUANTIZE_AFFINE:
        case QUANTA_OP_BACKEND_QUANTIZE: return quanta::quantize_workspace_bytes(cols);
        case QUANTA_OP_BACKEND_DEQUANTIZE: return 256;
        case QUANTA_OP_GEMM: return quanta::gemm_workspace_bytes(rows, cols);
        case QUANTA_OP_INT8_OUTLIER: return quanta::int8_outlier_workspace_bytes(rows, cols);
        case QUANTA_OP_BASE_QUANTIZE: return quanta::base_workspace_bytes(cols);
    }
    return 256;
}
