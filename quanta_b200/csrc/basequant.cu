// Convention C — BaseQuantizer (Quanta/functional/base.py:5-72; row N4).
//
//   quantize:   mn, mx per tensor or per column (dim 0); if allclose(mn, mx) for ALL groups: scale = 1, zp = mn;
//               symmetric:  scale = rcp(max(|mn|,|mx|)) * Q, zp = 0,  q = clamp(round(x*scale), -Q, Q) + 2^(bits-1)
//               asymmetric: scale = rcp(mx - mn) * L,    zp = mn,     q = clamp(round((x - zp)*scale), 0, L)
//               (the degenerate case still computes codes with scale 1 / zp mn — unlike convention B)
//   dequantize: symmetric (int8(q) - 2^(bits-1)) / scale, asymmetric q / scale + zp — true divides.
//
// A legacy API next to the hot path: three small launches (reduce -> all-groups flag -> codes), every arithmetic
// step one IEEE float32 operation as in the oracle (oracle/oracle_np.py:base_quantize).
#include "common.cuh"

namespace quanta {

// order-preserving keys of float32 with -0.0 < +0.0; NaN maps to the winning extreme of each reduction
__device__ __forceinline__ uint32_t bq_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float bq_unkey(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
constexpr uint32_t kBqNanMin = 0u, kBqNanMax = 0xFFFFFFFFu;
__device__ __forceinline__ float bq_min_of(uint32_t k) { return k == kBqNanMin ? __uint_as_float(0x7FC00000u) : bq_unkey(k); }
__device__ __forceinline__ float bq_max_of(uint32_t k) { return k == kBqNanMax ? __uint_as_float(0x7FC00000u) : bq_unkey(k); }

// workspace: [0] all-close flag | [16 ..) min keys [groups] | max keys [groups]
__global__ void bq_init_kernel(uint32_t* ws, int64_t groups) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) ws[0] = 1u;
    if (i < groups) { ws[16 + i] = 0xFFFFFFFFu; ws[16 + groups + i] = 0u; }
}

template <typename T>
__global__ void bq_minmax_tensor_kernel(const T* __restrict__ x, int64_t n, uint32_t* ws) {
    uint32_t kmn = 0xFFFFFFFFu, kmx = 0u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float f = to_f32(x[i]);
        if (f != f) { kmn = kBqNanMin; kmx = kBqNanMax; }
        else { const uint32_t k = bq_key(f); kmn = min(kmn, k); kmx = max(kmx, k); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmn = min(kmn, __shfl_xor_sync(0xffffffffu, kmn, o));
        kmx = max(kmx, __shfl_xor_sync(0xffffffffu, kmx, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&ws[16], kmn); atomicMax(&ws[17], kmx); }
}

// thread = one column, rows chunked over blockIdx.y
template <typename T>
__global__ void bq_minmax_dim0_kernel(const T* __restrict__ x, int64_t rows, int64_t cols, uint32_t* ws) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = (int64_t)blockIdx.y * per, r1 = min(rows, r0 + per);
    uint32_t kmn = 0xFFFFFFFFu, kmx = 0u;
    for (int64_t r = r0; r < r1; ++r) {
        const float f = to_f32(x[r * cols + c]);
        if (f != f) { kmn = kBqNanMin; kmx = kBqNanMax; }
        else { const uint32_t k = bq_key(f); kmn = min(kmn, k); kmx = max(kmx, k); }
    }
    if (r0 < r1) { atomicMin(&ws[16 + c], kmn); atomicMax(&ws[16 + cols + c], kmx); }
}

__device__ __forceinline__ bool bq_isclose(float a, float b) {       // torch.isclose, rtol 1e-5, atol 1e-8
    if (a == b) return true;
    const float allowed = __fadd_rn(1e-8f, fabsf(__fmul_rn(1e-5f, b)));
    const float actual = fabsf(__fsub_rn(a, b));
    return (actual <= 3.402823466e38f) && actual <= allowed;
}

__global__ void bq_allclose_kernel(uint32_t* ws, int64_t groups) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups) return;
    if (!bq_isclose(bq_min_of(ws[16 + i]), bq_max_of(ws[16 + groups + i]))) atomicAnd(&ws[0], 0u);
}

__device__ __forceinline__ void bq_params(const uint32_t* ws, int64_t groups, int64_t g, bool symmetric, float Q, float L,
                                          float* scale, float* zp) {
    const float mn = bq_min_of(ws[16 + g]), mx = bq_max_of(ws[16 + groups + g]);
    if (ws[0] != 0u) { *scale = 1.0f; *zp = mn; return; }                 // base.py:26-27
    if (symmetric) {
        const float am = max_nan(fabsf(mn), fabsf(mx));
        *scale = __fmul_rn(__frcp_rn(am), Q);                             // int / Tensor == reciprocal * int
        *zp = 0.0f;
    } else {
        *scale = __fmul_rn(__frcp_rn(__fsub_rn(mx, mn)), L);
        *zp = mn;
    }
}

// element i belongs to group i % groups (groups = cols for per_channel, 1 otherwise)
template <typename T>
__global__ void bq_quantize_kernel(const T* __restrict__ x, int64_t n, int64_t groups, int symmetric, int bits,
                                   const uint32_t* __restrict__ ws, uint8_t* __restrict__ q, float* __restrict__ scale_out,
                                   float* __restrict__ zp_out) {
    const float Q = (float)((1 << (bits - 1)) - 1), L = (float)((1 << bits) - 1);
    const uint32_t off = 1u << (bits - 1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = groups == 1 ? 0 : i % groups;
        float s, z;
        bq_params(ws, groups, g, symmetric != 0, Q, L, &s, &z);
        if (i < groups) { scale_out[i] = s; zp_out[i] = z; }
        const float f = to_f32(x[i]);
        uint32_t code;
        if (symmetric) {
            const float t = __fmul_rn(f, s);
            float c = fminf(fmaxf(t, -Q), Q);
            c = (t != t) ? 0.0f : c;                                       // NaN -> int8 0
            code = (uint32_t)((int)rintf(c) + (int)off);
        } else {
            const float t = __fmul_rn(__fsub_rn(f, z), s);
            code = code_bits(t, L) - kMagicBits;                           // clamp(rint(t), 0, L), NaN -> 0
        }
        q[i] = (uint8_t)code;
    }
}

__global__ void bq_dequantize_kernel(const uint8_t* __restrict__ q, int64_t n, int64_t groups, int symmetric, int bits,
                                     const float* __restrict__ scale, const float* __restrict__ zp, float* __restrict__ out) {
    const int off = 1 << (bits - 1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = groups == 1 ? 0 : i % groups;
        const float s = scale[g];
        if (symmetric) {
            const int v = (int)(int8_t)((int)(int8_t)q[i] - off);          // int8 arithmetic, wraps
            out[i] = __fdiv_rn((float)v, s);
        } else {
            out[i] = __fadd_rn(__fdiv_rn((float)q[i], s), zp[g]);
        }
    }
}

size_t base_workspace_bytes(int64_t groups) { return (size_t)(16 + 2 * (groups > 0 ? groups : 1)) * 4 + 256; }

template <typename T>
static int bq_quantize_launch(const T* x, int64_t rows, int64_t cols, int per_channel, int symmetric, int bits, uint8_t* q,
                              float* scale, float* zp, uint32_t* ws, cudaStream_t st) {
    const int64_t n = rows * cols, groups = per_channel ? cols : 1;
    bq_init_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(ws, groups);
    if (per_channel) {
        int64_t chunks = rows / 64; if (chunks < 1) chunks = 1; if (chunks > 256) chunks = 256;
        bq_minmax_dim0_kernel<T><<<dim3((unsigned)((cols + 127) / 128), (unsigned)chunks), 128, 0, st>>>(x, rows, cols, ws);
    } else {
        int64_t blocks = (n + 1023) / 1024; if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
        bq_minmax_tensor_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(x, n, ws);
    }
    bq_allclose_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(ws, groups);
    int64_t blocks = (n + 1023) / 1024; if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    bq_quantize_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(x, n, groups, symmetric, bits, ws, q, scale, zp);
    return cuda_status(cudaGetLastError());
}

}  // namespace quanta

using namespace quanta;

extern "C" int quanta_base_quantize(const void* x, int x_dtype, int64_t rows, int64_t cols, int per_channel, int symmetric,
                                    int bits, uint8_t* q_out, float* scale_out, float* zp_out, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    if (!x || !q_out || !scale_out || !zp_out || rows <= 0 || cols <= 0) return QUANTA_EINVAL;
    if (bits != 4 && bits != 8) return QUANTA_EUNSUPPORTED;
    const int64_t groups = per_channel ? cols : 1;
    if (!workspace || workspace_bytes < base_workspace_bytes(groups)) return QUANTA_EWORKSPACE;
    uint32_t* ws = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (x_dtype) {
        case QUANTA_F32: return bq_quantize_launch((const float*)x, rows, cols, per_channel, symmetric, bits, q_out, scale_out, zp_out, ws, st);
        case QUANTA_F16: return bq_quantize_launch((const __half*)x, rows, cols, per_channel, symmetric, bits, q_out, scale_out, zp_out, ws, st);
        case QUANTA_BF16: return bq_quantize_launch((const __nv_bfloat16*)x, rows, cols, per_channel, symmetric, bits, q_out, scale_out, zp_out, ws, st);
    }
    return QUANTA_EINVAL;
}

extern "C" int quanta_base_dequantize(const uint8_t* q, int64_t rows, int64_t cols, int64_t nchan, int bits, int symmetric,
                                      const float* scale, const float* zp, float* out, void* stream) {
    if (!q || !scale || !zp || !out || rows <= 0 || cols <= 0) return QUANTA_EINVAL;
    if (bits != 4 && bits != 8) return QUANTA_EUNSUPPORTED;
    if (nchan != 1 && nchan != cols) return QUANTA_EINVAL;
    const int64_t n = rows * cols;
    int64_t blocks = (n + 1023) / 1024; if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    bq_dequantize_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(q, n, nchan, symmetric, bits, scale, zp, out);
    return cuda_status(cudaGetLastError());
}
