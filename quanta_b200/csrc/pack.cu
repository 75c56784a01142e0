// 4-bit nibble pack / unpack (rows P1, P2 of SURVEY §8):
// Quanta/utils/utils.py:23-48.  Pure byte shuffling, HBM-bound: 128-bit loads,
// 64/128-bit stores, four independent vectors in flight per thread.
#include "common.cuh"

namespace quanta {

// [b0 b1 b2 b3] -> byte0 = b0 | (b1 << 4), byte2 = b2 | (b3 << 4) in uint8
// arithmetic: the high nibble of b1/b3 drops, b0/b2 are NOT masked (reference
// behaviour for inputs > 15, utils.py:34).
// HI = false: utils.py order (even index -> low nibble).  HI = true: ModelQuantize._pack_tensor's order
// (Quanta/functional/model.py:73-82: byte = (t[2i] << 4) | t[2i+1], first element -> HIGH nibble).
template <bool HI>
__device__ __forceinline__ uint32_t pack_word(uint32_t w) {
    return HI ? (((w << 4) & 0x00F000F0u) | ((w >> 8) & 0x00FF00FFu))
              : ((w & 0x00FF00FFu) | ((w >> 4) & 0x00F000F0u));
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r;
}

// 16 codes (uint4) -> 8 bytes (uint2) per vector.
template <bool HI>
__global__ void __launch_bounds__(256) pack4_vec_kernel(const uint4* __restrict__ q, int64_t nvec, uint2* __restrict__ out) {
    pdl_wait(); 
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto pack = [](uint4 v) {
        uint32_t a = pack_word<HI>(v.x), b = pack_word<HI>(v.y), c = pack_word<HI>(v.z), d = pack_word<HI>(v.w);
        return make_uint2(prmt(a, b, 0x6420u), prmt(c, d, 0x6420u));
    };
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldcs(q + i + k * stride);
#pragma unroll
        for (int k = 0; k < 4; ++k) __stcs(out + i + k * stride, pack(v[k]));
    }
    for (; i < nvec; i += stride) __stcs(out + i, pack(__ldcs(q + i)));
}

template <bool HI>
__global__ void __launch_bounds__(256) pack4_tail_kernel(const uint8_t* __restrict__ q, int64_t start, int64_t n,
                                                         uint8_t* __restrict__ out) {
    const int64_t i = start + 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n) return;
    uint8_t lo = q[i];
    uint8_t hi = (i + 1 < n) ? q[i + 1] : 0;                 // one zero pad if n is odd (utils.py:31-32)
    out[i >> 1] = HI ? (uint8_t)((uint8_t)(lo << 4) | hi) : (uint8_t)(lo | (uint8_t)(hi << 4));
}

// 8 packed bytes (uint2) -> 16 codes (uint4) per vector.
template <bool HI>
__global__ void __launch_bounds__(256) unpack4_vec_kernel(const uint2* __restrict__ p, int64_t nvec, uint4* __restrict__ out) {
    pdl_wait(); 
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto unpack = [](uint2 v) {
        // lo* = the nibble that decodes to the even output index
        uint32_t lo0 = (HI ? v.x >> 4 : v.x) & 0x0F0F0F0Fu, hi0 = (HI ? v.x : v.x >> 4) & 0x0F0F0F0Fu;
        uint32_t lo1 = (HI ? v.y >> 4 : v.y) & 0x0F0F0F0Fu, hi1 = (HI ? v.y : v.y >> 4) & 0x0F0F0F0Fu;
        return make_uint4(prmt(lo0, hi0, 0x5140u), prmt(lo0, hi0, 0x7362u), prmt(lo1, hi1, 0x5140u), prmt(lo1, hi1, 0x7362u));
    };
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        uint2 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldcs(p + i + k * stride);
#pragma unroll
        for (int k = 0; k < 4; ++k) __stcs(out + i + k * stride, unpack(v[k]));
    }
    for (; i < nvec; i += stride) __stcs(out + i, unpack(__ldcs(p + i)));
}

template <bool HI>
__global__ void __launch_bounds__(256) unpack4_tail_kernel(const uint8_t* __restrict__ p, int64_t start, int64_t nbytes,
                                                           uint8_t* __restrict__ out) {
    const int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbytes) return;
    uint8_t b = p[i];
    out[2 * i] = HI ? (b >> 4) & 0x0F : b & 0x0F;
    out[2 * i + 1] = HI ? b & 0x0F : (b >> 4) & 0x0F;
}

static unsigned grid_for(int64_t nvec) {
    int64_t want = (nvec + 1023) / 1024;            // 4 vectors per thread
    int64_t cap = (int64_t)kNumSMs * 8;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace quanta

using namespace quanta;

template <bool HI>
static int pack4_launch(const uint8_t* q, int64_t n, uint8_t* packed, void* stream) {
    if (!q || !packed || n < 0) return QUANTA_EINVAL;
    if (n == 0) return QUANTA_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool ok = (reinterpret_cast<uintptr_t>(q) % 16 == 0) && (reinterpret_cast<uintptr_t>(packed) % 8 == 0);
    const int64_t nvec = ok ? n / 16 : 0;
    if (nvec > 0)
        launch_pdl(pack4_vec_kernel<HI>, dim3(grid_for(nvec)), dim3(256), 0, st, reinterpret_cast<const uint4*>(q), nvec,
                   reinterpret_cast<uint2*>(packed));
    const int64_t start = nvec * 16;
    if (start < n) {
        const int64_t pairs = (n - start + 1) / 2;
        pack4_tail_kernel<HI><<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(q, start, n, packed);
    }
    return cuda_status(cudaGetLastError());
}

template <bool HI>
static int unpack4_launch(const uint8_t* packed, int64_t nbytes, uint8_t* out, void* stream) {
    if (!packed || !out || nbytes < 0) return QUANTA_EINVAL;
    if (nbytes == 0) return QUANTA_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool ok = (reinterpret_cast<uintptr_t>(packed) % 8 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    const int64_t nvec = ok ? nbytes / 8 : 0;
    if (nvec > 0)
        launch_pdl(unpack4_vec_kernel<HI>, dim3(grid_for(nvec)), dim3(256), 0, st, reinterpret_cast<const uint2*>(packed), nvec,
                   reinterpret_cast<uint4*>(out));
    const int64_t start = nvec * 8;
    if (start < nbytes)
        unpack4_tail_kernel<HI><<<(unsigned)((nbytes - start + 255) / 256), 256, 0, st>>>(packed, start, nbytes, out);
    return cuda_status(cudaGetLastError());
}

extern "C" int quanta_pack4(const uint8_t* q, int64_t n, uint8_t* packed, void* stream) { return pack4_launch<false>(q, n, packed, stream); }
extern "C" int quanta_unpack4(const uint8_t* packed, int64_t nbytes, uint8_t* out, void* stream) { return unpack4_launch<false>(packed, nbytes, out, stream); }
extern "C" int quanta_pack4_hi(const uint8_t* q, int64_t n, uint8_t* packed, void* stream) { return pack4_launch<true>(q, n, packed, stream); }
extern "C" int quanta_unpack4_hi(const uint8_t* packed, int64_t nbytes, uint8_t* out, void* stream) { return unpack4_launch<true>(packed, nbytes, out, stream); }
