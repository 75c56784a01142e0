// W4A16 / W8A16 dequantize-then-matmul for decode-sized batches (M <= 4) on the CUDA cores
// (rows G1/G2 of SURVEY §8): y[M,N] = x[M,K] . dequant(Wq)[N,K]^T + bias.
//
// A tcgen05.mma costs ~70 cycles of dispatch however small it is, which caps the tensor-core
// kernel (gemm.cu) at ~29 weights/clk/SM — below what the HBM stream delivers (36.6).  For one
// to four activation rows the products are cheap enough to do in fp32 FFMAs instead:
//
//   * persistent CTAs; each converts x to fp32 once into shared memory, laid out so that lane l's
//     k-values are one conflict-free LDS.128 per 4 values, plus the per-lane sums of x;
//   * a warp streams 512 contiguous bytes of one weight row per step (LDG.128 per lane =
//     32 four-bit / 16 eight-bit codes, no allocation in L1), several steps in flight;
//   * codes are NOT dequantized one by one: PRMT drops a code into the mantissa of 128.0f (4-bit) /
//     32768.0f (8-bit), the FFMA accumulates raw = sum x_k (C + n_k), and scale / zero-point are
//     applied once per lane and step:  acc += s * raw + (z - C s) * sum x_k
//     -> 1 PRMT + M FFMA per weight;
//   * a CTA's warps split each row's K range S ways, partials meet in shared memory (fixed order,
//     deterministic), bias is added and y stored.
// Products use the exact fp32 weight q*s + z (the tensor-core path rounds it to the activation
// type first, like the reference's `.to(x.dtype)`); the difference is far inside the 1e-2 tolerance.
#include "common.cuh"

namespace quanta {

constexpr int kGemvWarps = 24;        // one CTA of 24 warps per SM: 24 x 4 steps x 512 B = 48 KB of codes in flight
constexpr int kGemvUnroll = 4;       // steps (LDG.128 per lane) in flight per warp

template <typename T> __device__ __forceinline__ float gv_to_float(T v);
template <> __device__ __forceinline__ float gv_to_float<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float gv_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T gv_from_float(float v);
template <> __device__ __forceinline__ __half gv_from_float<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 gv_from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t gv_prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r;
}
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

struct GemvParams {
    int M, N, K;
    int S;              // K splits of a row inside a CTA (warps per row)
    int rows_per_cta;   // warps / S
    int steps;          // ceil(K / (32 * G)): 512-byte steps per row
    int scale_stride;   // K / 64
};

// G = codes per lane and step: 32 (4-bit) or 16 (8-bit).
template <typename ACT, int BITS, int MB>
__global__ void __launch_bounds__(32 * kGemvWarps, 1)
gemv_wna16_kernel(const ACT* __restrict__ x, const uint8_t* __restrict__ wq, const float* __restrict__ scale,
                  const float* __restrict__ zp, const ACT* __restrict__ bias, ACT* __restrict__ y, const GemvParams p) {
    constexpr int G = BITS == 4 ? 32 : 16;
    constexpr int NJ = G / 4;                        // float4 groups of x per lane and step
    constexpr float C = BITS == 4 ? 128.0f : 32768.0f;
    extern __shared__ float4 gsm[];
    // xs[m][step][j][lane] (float4) | hs[m][step][lane] (float) | red[warp][MB]
    float4* xs = gsm;
    float* hs = reinterpret_cast<float*>(xs + (size_t)MB * p.steps * NJ * 32);
    float* red = hs + (size_t)MB * p.steps * 32;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;

    // ---- x -> fp32, permuted so that lane l's 4 values of group j sit at xs[..][j][l] ----
    for (int m = 0; m < MB; ++m) {
        for (int idx = tid; idx < p.steps * 32 * NJ; idx += nthreads) {
            const int st = idx / (32 * NJ), r = idx - st * 32 * NJ;
            const int j = r >> 5, l = r & 31;
            const int k = (st * 32 + l) * G + 4 * j;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < p.M && k < p.K) {                // K is a multiple of 4
                const ACT* px = x + (int64_t)m * p.K + k;
                v = make_float4(gv_to_float(px[0]), gv_to_float(px[1]), gv_to_float(px[2]), gv_to_float(px[3]));
            }
            xs[((size_t)m * p.steps + st) * NJ * 32 + r] = v;
        }
    }
    __syncthreads();
    for (int m = 0; m < MB; ++m) {
        for (int idx = tid; idx < p.steps * 32; idx += nthreads) {
            const int st = idx >> 5, l = idx & 31;
            float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float4 v = xs[((size_t)m * p.steps + st) * NJ * 32 + j * 32 + l];
                sum += (v.x + v.y) + (v.z + v.w);
            }
            hs[(size_t)m * p.steps * 32 + idx] = sum;
        }
    }
    __syncthreads();

    const int split = warp % p.S, rslot = warp / p.S;
    const int row_bytes = p.K * BITS / 8;
    const int n_groups = (p.N + p.rows_per_cta - 1) / p.rows_per_cta;
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int row = grp * p.rows_per_cta + rslot;
        float acc[MB];
#pragma unroll
        for (int m = 0; m < MB; ++m) acc[m] = 0.0f;
        if (row < p.N) {
            const uint8_t* wrow = wq + (int64_t)row * row_bytes + 16 * lane;
            const float* srow = scale + (int64_t)row * p.scale_stride;
            const float* zrow = zp + (int64_t)row * p.scale_stride;
            // steps split, split + S, ...; kGemvUnroll steps in flight
            for (int st0 = split; st0 < p.steps; st0 += kGemvUnroll * p.S) {
                uint4 raw[kGemvUnroll];
                float sc[kGemvUnroll], zc[kGemvUnroll];
                bool ok[kGemvUnroll];
#pragma unroll
                for (int u = 0; u < kGemvUnroll; ++u) {
                    const int st = st0 + u * p.S;
                    const int k = (st * 32 + lane) * G;
                    ok[u] = st < p.steps && k < p.K;
                    raw[u] = make_uint4(0, 0, 0, 0);
                    sc[u] = 0.0f; zc[u] = 0.0f;
                    if (ok[u]) {
                        raw[u] = ldg_stream(wrow + (int64_t)st * 512);
                        sc[u] = __ldg(srow + (k >> 6));
                        zc[u] = __ldg(zrow + (k >> 6));
                    }
                }
#pragma unroll
                for (int u = 0; u < kGemvUnroll; ++u) {
                    if (!ok[u]) continue;
                    const int st = st0 + u * p.S;
                    const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
                    float rsum[MB];
#pragma unroll
                    for (int m = 0; m < MB; ++m) rsum[m] = 0.0f;
                    const float4* xb = xs + (size_t)st * NJ * 32 + lane;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (BITS == 4) {
                            // word = nibbles n0..n7 = k 8i .. 8i+7; bytes of ev = (n0,n2,n4,n6), of od = (n1,n3,n5,n7)
                            const uint32_t ev = w[i] & 0x0F0F0F0Fu, od = (w[i] >> 4) & 0x0F0F0F0Fu;
                            float c[8];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                c[2 * e] = __uint_as_float(gv_prmt(ev, 0x43000000u, 0x7044u | (e << 8)));       // 128 + n
                                c[2 * e + 1] = __uint_as_float(gv_prmt(od, 0x43000000u, 0x7044u | (e << 8)));
                            }
#pragma unroll
                            for (int m = 0; m < MB; ++m) {
                                const float4 a = xb[((size_t)m * p.steps * NJ + 2 * i) * 32];
                                const float4 b = xb[((size_t)m * p.steps * NJ + 2 * i + 1) * 32];
                                // two independent chains per row keep the FFMA latency off the critical path
                                float r0 = fmaf(a.x, c[0], 0.0f), r1 = fmaf(a.y, c[1], 0.0f);
                                r0 = fmaf(a.z, c[2], r0); r1 = fmaf(a.w, c[3], r1);
                                r0 = fmaf(b.x, c[4], r0); r1 = fmaf(b.y, c[5], r1);
                                r0 = fmaf(b.z, c[6], r0); r1 = fmaf(b.w, c[7], r1);
                                rsum[m] += r0 + r1;
                            }
                        } else {
                            float c[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                c[e] = __uint_as_float(gv_prmt(w[i], 0x47000000u, 0x7504u | (e << 4)));         // 32768 + q
#pragma unroll
                            for (int m = 0; m < MB; ++m) {
                                const float4 a = xb[((size_t)m * p.steps * NJ + i) * 32];
                                float r0 = fmaf(a.x, c[0], 0.0f), r1 = fmaf(a.y, c[1], 0.0f);
                                r0 = fmaf(a.z, c[2], r0); r1 = fmaf(a.w, c[3], r1);
                                rsum[m] += r0 + r1;
                            }
                        }
                    }
                    const float zf = fmaf(-C, sc[u], zc[u]);
#pragma unroll
                    for (int m = 0; m < MB; ++m)
                        acc[m] += fmaf(sc[u], rsum[m], zf * hs[((size_t)m * p.steps + st) * 32 + lane]);
                }
            }
        }
        // lanes -> warp total
#pragma unroll
        for (int m = 0; m < MB; ++m) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int m = 0; m < MB; ++m) red[warp * MB + m] = acc[m];
        }
        __syncthreads();
        if (split == 0 && lane < p.M && row < p.N) {
            float v = 0.0f;
            for (int s = 0; s < p.S; ++s) v += red[(warp + s) * MB + lane];       // fixed order: deterministic
            if (bias != nullptr) v += gv_to_float(bias[row]);
            y[(int64_t)lane * p.N + row] = gv_from_float<ACT>(v);
        }
        __syncthreads();
    }
}

static size_t gemv_smem_bytes(int bits, int mb, int64_t K) {
    const int G = bits == 4 ? 32 : 16;
    const int64_t steps = (K + 32 * G - 1) / (32 * G);
    return (size_t)mb * steps * 32 * (G / 4) * 16 + (size_t)mb * steps * 32 * 4 + (size_t)kGemvWarps * mb * 4 + 64;
}

// The CUDA-core path serves M <= 4, block = 64, when x (as fp32) fits in shared memory.
bool gemv_eligible(int bits, int64_t M, int64_t K, int64_t block) {
    if (M > 4 || block != 64 || (K & 63) != 0) return false;
    const int mb = M <= 1 ? 1 : (M <= 2 ? 2 : 4);
    return gemv_smem_bytes(bits, mb, K) <= 200u * 1024u;
}

template <typename ACT, int BITS, int MB>
static int gemv_launch_mb(const ACT* x, const uint8_t* wq, const float* scale, const float* zp, const ACT* bias, ACT* y,
                          int64_t M, int64_t N, int64_t K, cudaStream_t st) {
    GemvParams p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    constexpr int G = BITS == 4 ? 32 : 16;
    p.steps = (int)((K + 32 * G - 1) / (32 * G));
    p.scale_stride = (int)(K / 64);
    const size_t smem = gemv_smem_bytes(BITS, MB, K);
    const int warps = kGemvWarps, grid_max = kNumSMs;
    // K splits per row (a divisor of the warp count): balance the row groups over the grid and
    // keep at least one full batch of steps per split
    int best_s = 1;
    double best_eff = -1.0;
    for (int s = 1; s <= warps; ++s) {
        if (warps % s) continue;
        if (s > 1 && p.steps / s < kGemvUnroll) break;
        const int64_t groups = (N + warps / s - 1) / (warps / s);
        const double per = (double)groups / grid_max;
        const double eff = per / (double)((groups + grid_max - 1) / grid_max);
        if (eff > best_eff + 0.02) { best_eff = eff; best_s = s; }
    }
    p.S = best_s;
    p.rows_per_cta = warps / p.S;
    const int64_t groups = (N + p.rows_per_cta - 1) / p.rows_per_cta;
    const int grid = (int)(groups < grid_max ? groups : grid_max);
    auto kern = gemv_wna16_kernel<ACT, BITS, MB>;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        smem_set = smem;
    }
    kern<<<grid, 32 * warps, smem, st>>>(x, wq, scale, zp, bias, y, p);
    return cuda_status(cudaGetLastError());
}

template <typename ACT, int BITS>
int gemv_launch(const ACT* x, const uint8_t* wq, const float* scale, const float* zp, const ACT* bias, ACT* y, int64_t M,
                int64_t N, int64_t K, cudaStream_t st) {
    if (M <= 1) return gemv_launch_mb<ACT, BITS, 1>(x, wq, scale, zp, bias, y, M, N, K, st);
    if (M <= 2) return gemv_launch_mb<ACT, BITS, 2>(x, wq, scale, zp, bias, y, M, N, K, st);
    return gemv_launch_mb<ACT, BITS, 4>(x, wq, scale, zp, bias, y, M, N, K, st);
}

template int gemv_launch<__nv_bfloat16, 4>(const __nv_bfloat16*, const uint8_t*, const float*, const float*, const __nv_bfloat16*, __nv_bfloat16*, int64_t, int64_t, int64_t, cudaStream_t);
template int gemv_launch<__nv_bfloat16, 8>(const __nv_bfloat16*, const uint8_t*, const float*, const float*, const __nv_bfloat16*, __nv_bfloat16*, int64_t, int64_t, int64_t, cudaStream_t);
template int gemv_launch<__half, 4>(const __half*, const uint8_t*, const float*, const float*, const __half*, __half*, int64_t, int64_t, int64_t, cudaStream_t);
template int gemv_launch<__half, 8>(const __half*, const uint8_t*, const float*, const float*, const __half*, __half*, int64_t, int64_t, int64_t, cudaStream_t);

}  // namespace quanta
