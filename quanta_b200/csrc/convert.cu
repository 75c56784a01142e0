// Fused convert_precision for per-tensor linear codes (row N3; Quanta/utils/utils.py:216-279).
//
// The reference dequantizes to fp32 and quantizes again: dequantize (1 B read + 4 B write) + min/max + quantize
// (4 B read twice + 1 B write) = 14 B per element through three full passes.  With ONE (scale, zero_point) for
// the whole source tensor the dequantized value is a function of the code alone, v(c) = c*s + z (multiply and
// add rounded separately, quantization.py:38), so
//   * the min / max of the dequantized tensor are the min / max of v over the codes that OCCUR,
//   * the new code is a function of the old code: a 256-entry table built with the reference's own arithmetic
//     (common.cuh: affine_params / affine_quotient / code_bits).
// Pass 1 marks the codes that occur (1 B read per element), pass 2 builds the table in every CTA and maps
// (1 B read + 1 B write): 3 B per element, bit-identical to dequantize + quantize for any scale (NaN / inf
// included: the table is computed from the same values the two-step path would see).
#include "common.cuh"

namespace quanta {

__global__ void __launch_bounds__(256) convert_presence_kernel(const uint8_t* __restrict__ q, int64_t n, uint32_t* __restrict__ flags) {
    __shared__ uint32_t bits[8];
    if (threadIdx.x < 8) bits[threadIdx.x] = 0u;
    __syncthreads();
    uint32_t mine[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    const int64_t nvec = ((reinterpret_cast<uintptr_t>(q) & 15) == 0) ? n / 16 : 0;
    const uint4* qv = reinterpret_cast<const uint4*>(q);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcs(qv + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int b = 0; b < 4; ++b) { const uint32_t c = (w[k] >> (8 * b)) & 0xFFu; mine[c >> 5] |= 1u << (c & 31); }
    }
    for (int64_t i = nvec * 16 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t c = q[i];
        mine[c >> 5] |= 1u << (c & 31);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t m = mine[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
        if ((threadIdx.x & 31) == 0 && m) atomicOr(&bits[k], m);
    }
    __syncthreads();
    if (threadIdx.x < 8 && bits[threadIdx.x]) atomicOr(&flags[threadIdx.x], bits[threadIdx.x]);
}

__global__ void __launch_bounds__(256) convert_map_kernel(const uint8_t* __restrict__ q, int64_t n, const float* __restrict__ src_scale,
                                                          const float* __restrict__ src_zp, int tgt_bits,
                                                          const uint32_t* __restrict__ flags, uint8_t* __restrict__ out,
                                                          float* __restrict__ scale_out, float* __restrict__ zp_out) {
    __shared__ float red_mn[8], red_mx[8];
    __shared__ uint32_t lut[256 * 32];                        // entry c replicated per lane: conflict-free lookups
    const int c = threadIdx.x, lane = c & 31, warp = c >> 5;
    const float s = src_scale[0], z = src_zp[0];
    const float v = __fadd_rn(__fmul_rn((float)c, s), z);     // dequantize_*bit, quantization.py:38 / :58
    const bool present = (flags[c >> 5] >> (c & 31)) & 1u;
    // min / max of the dequantized tensor = over the codes that occur (NaN propagates like torch.min / max)
    float mn = present ? v : __uint_as_float(0x7F800000u), mx = present ? v : __uint_as_float(0xFF800000u);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min_nan(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { red_mn[warp] = mn; red_mx[warp] = mx; }
    __syncthreads();
    mn = red_mn[0]; mx = red_mx[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) { mn = min_nan(mn, red_mn[k]); mx = max_nan(mx, red_mx[k]); }
    const float L = tgt_bits == 8 ? 255.0f : 15.0f;
    const AffineParams p = affine_params(mn, mx, L);          // quantize_*bit_linear, quantization.py:73-99 / :185-210
    const uint32_t code = code_bits(affine_quotient(v, p), L) - kMagicBits;
    for (int l = 0; l < 32; ++l) lut[c * 32 + l] = code;
    if (blockIdx.x == 0 && c == 0) { scale_out[0] = p.scale; zp_out[0] = p.mn; }
    __syncthreads();

    const bool vec_ok = ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int64_t nvec = vec_ok ? n / 16 : 0;
    const uint4* qv = reinterpret_cast<const uint4*>(q);
    uint4* ov = reinterpret_cast<uint4*>(out);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 in = __ldcs(qv + i);
        const uint32_t w[4] = {in.x, in.y, in.z, in.w};
        uint32_t r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            r[k] = lut[((w[k]) & 0xFFu) * 32 + lane] | (lut[((w[k] >> 8) & 0xFFu) * 32 + lane] << 8) |
                   (lut[((w[k] >> 16) & 0xFFu) * 32 + lane] << 16) | (lut[(w[k] >> 24) * 32 + lane] << 24);
        }
        __stcs(ov + i, make_uint4(r[0], r[1], r[2], r[3]));
    }
    for (int64_t i = nvec * 16 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (uint8_t)lut[(uint32_t)q[i] * 32 + lane];
}

}  // namespace quanta

using namespace quanta;

extern "C" int quanta_convert_linear(const uint8_t* q, int64_t n, const float* src_scale, const float* src_zp, int target_bits,
                                     uint8_t* q_out, float* scale_out, float* zp_out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    if (!q || !src_scale || !src_zp || !q_out || !scale_out || !zp_out || n <= 0) return QUANTA_EINVAL;
    if (target_bits != 4 && target_bits != 8) return QUANTA_EUNSUPPORTED;
    if (!workspace || workspace_bytes < 256) return QUANTA_EWORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint32_t* flags = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(workspace) + 31) & ~uintptr_t(31));
    cudaError_t e = cudaMemsetAsync(flags, 0, 32, st);
    if (e != cudaSuccess) return (int)e;
    int64_t blocks = (n / 16 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    convert_presence_kernel<<<(unsigned)blocks, 256, 0, st>>>(q, n, flags);
    if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;           // every CTA builds the 32 KB table once
    convert_map_kernel<<<(unsigned)blocks, 256, 0, st>>>(q, n, src_scale, src_zp, target_bits, flags, q_out, scale_out, zp_out);
    return cuda_status(cudaGetLastError());
}
