// W4A16 / W8A16 dequantize-then-matmul on the 5th-generation tensor cores
// (rows G1/G2 of SURVEY §8): y[M,N] = x[M,K] . dequant(Wq)[N,K]^T + bias.
//
// Mapping ("swap-AB"): the weight tile's 128 output features are the UMMA M
// dimension (one TMEM lane each), the batch rows are the UMMA N dimension
// (16..256), so small batches do not waste the 128-row MMA:
//     D[n, m] (fp32, TMEM) += A[n, k] (dequantized weights, smem) * B[m, k] (x, smem)
// Per 64-wide K block and pipeline stage:
//   warp 0      TMA: raw weight codes [128 x 64] (u8 or nibble-packed) and x [MB x 64]
//               (SWIZZLE_128B) -> shared memory, completes tma_full[s]
//   warps 2..5  dequantize: thread = weight row; codes -> act dtype with the block's
//               scale / zero-point, written as the K-major SWIZZLE_128B A tile;
//               fence.proxy.async; arrive a_full[s]
//   warp 1      one lane issues 4 x tcgen05.mma (K = 16 each); tcgen05.commit frees the stage
// After the last K block the four dequant warps read the accumulator with
// tcgen05.ld (lane = output feature) and store y (or a split-K partial).
#include "common.cuh"

#include <cstdlib>

namespace quanta {

constexpr int kDequantWarps = 8;      // warps that cooperate on one K block (each thread: half a weight row)
constexpr int kDequantGroups = 2;     // groups work on alternating K blocks -> 4 dequant warps per scheduler
constexpr int kFirstDequantWarp = 3;  // warp 0: raw-code TMA, warp 1: activation TMA, warp 2: MMA issuer + TMEM
constexpr int kGemmThreads = 32 * (kFirstDequantWarp + kDequantWarps * kDequantGroups);
constexpr int kMaxRawStages = 32;
constexpr int kAStages = 4;           // dequantized A tiles in flight (even: the groups alternate)
constexpr int kTileN = 128;          // output features per CTA (UMMA M)
constexpr int kBlockK = 64;          // K elements per stage = one 128-byte swizzle atom of 16-bit values
constexpr int kATileBytes = kTileN * kBlockK * 2;

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128g(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t prmt_(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r;
}
__device__ __forceinline__ void tma_load_2d_plain(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0,
                                                  int32_t c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}

// ---- tcgen05 wrappers -----------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// completes `bar` (one arrival) when all previously issued MMAs of this thread have finished
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row
// swizzle atoms 1024 bytes apart (SBO), LBO unused, descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// ---- activation-type traits -------------------------------------------------
template <typename ACT> struct ActTraits;
template <> struct ActTraits<__nv_bfloat16> {
    static constexpr uint32_t kFmt = 1;                  // UMMA F16F32Format::BF16
    static constexpr uint32_t kMagic = 0x43004300u;      // bf16x2 (128 + n)
    static constexpr CUtensorMapDataType kTma = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    using V2 = __nv_bfloat162;
    __device__ static __forceinline__ V2 bias2() { return __float2bfloat162_rn(128.0f); }
    __device__ static __forceinline__ V2 dup(float v) { return __float2bfloat162_rn(v); }
    __device__ static __forceinline__ uint32_t pack(float a, float b) {
        V2 t = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&t);
    }
    __device__ static __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ static __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
};
template <> struct ActTraits<__half> {
    static constexpr uint32_t kFmt = 0;                  // F16
    static constexpr uint32_t kMagic = 0x64006400u;      // fp16x2 (1024 + n)
    static constexpr CUtensorMapDataType kTma = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    using V2 = __half2;
    __device__ static __forceinline__ V2 bias2() { return __float2half2_rn(1024.0f); }
    __device__ static __forceinline__ V2 dup(float v) { return __float2half2_rn(v); }
    __device__ static __forceinline__ uint32_t pack(float a, float b) {
        V2 t = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&t);
    }
    __device__ static __forceinline__ float to_float(__half v) { return __half2float(v); }
    __device__ static __forceinline__ __half from_float(float v) { return __float2half_rn(v); }
};

// pack two fp32 values into the A operand's 16-bit format
template <uint32_t kAFmt>
__device__ __forceinline__ uint32_t pack_a(float a, float b) {
    if (kAFmt == 0) { __half2 t = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&t); }
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

// dequantize one packed pair register: ((magic | nibbles) - bias) * scale + zp in 16-bit x2 math
template <typename ACT>
__device__ __forceinline__ uint32_t dq_pair(uint32_t nib2, typename ActTraits<ACT>::V2 s2, typename ActTraits<ACT>::V2 z2) {
    using V2 = typename ActTraits<ACT>::V2;
    uint32_t m = nib2 | ActTraits<ACT>::kMagic;
    V2 t = *reinterpret_cast<V2*>(&m);
    t = __hsub2(t, ActTraits<ACT>::bias2());
    t = __hfma2(t, s2, z2);
    return *reinterpret_cast<uint32_t*>(&t);
}

struct GemmParams {
    int M, N, K;
    int mb;                 // UMMA N: batch rows per CTA tile (multiple of 16, <= 256)
    int split_k;            // number of K splits
    int kblocks_per_split;  // 64-wide K blocks per split
    int raw_stages;         // deep ring of raw weight codes: this is what keeps HBM busy
    int x_stages;           // ring of activation tiles
    int scale_stride;       // K / block
    int block_shift;        // log2(block / 64): K block kb uses scale column kb >> block_shift
    int tmem_cols;
    uint32_t x_bytes, raw_bytes;
    uint32_t x_ring_off, raw_ring_off;   // byte offsets of the rings behind the A ring
};

// kAFmt: UMMA format of the dequantized A operand (0 = fp16, 1 = bf16).  fp16 A with bf16
// activations is the mixed-format mode (selected at run time, see gemm_launch).
template <typename ACT, int BITS, uint32_t kAFmt>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_wna16_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                  const float* __restrict__ scale, const float* __restrict__ zp, const ACT* __restrict__ bias,
                  ACT* __restrict__ y, float* __restrict__ partial, const GemmParams p) {
    using AT = ActTraits<ACT>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t raw_full[kMaxRawStages], raw_empty[kMaxRawStages];
    __shared__ uint64_t x_full[8], x_empty[8], a_full[kAStages], a_empty[kAStages], tmem_full;
    __shared__ uint32_t tmem_base_slot;

    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tile = blockIdx.x, split = blockIdx.y, m_tile = blockIdx.z;
    const int n0 = n_tile * kTileN, m0 = m_tile * p.mb;
    const int kb0 = split * p.kblocks_per_split;
    const int nkb = p.kblocks_per_split;

    // shared memory: [A ring: kAStages x 16 KB][x ring: x_stages x mb*128 B][raw ring: raw_stages x 4|8 KB]
    auto a_addr = [&](int s) { return smem + s * kATileBytes; };
    auto x_addr = [&](int s) { return smem + p.x_ring_off + s * p.x_bytes; };
    auto r_addr = [&](int s) { return smem + p.raw_ring_off + s * p.raw_bytes; };

    if (tid == 0) {
        for (int s = 0; s < p.raw_stages; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], kDequantWarps); }
        for (int s = 0; s < p.x_stages; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
        for (int s = 0; s < kAStages; ++s) { mbar_init(&a_full[s], kDequantWarps); mbar_init(&a_empty[s], 1); }
        mbar_init(&tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(&tmem_base_slot), p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        // ===== raw weight codes: deep TMA ring (bytes in flight = raw_stages x raw_bytes) =====
        if (lane == 0) {
            prefetch_tensormap(&tmap_w);
            const uint64_t pol_w = policy_evict_first();     // weights are streamed once
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&raw_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&raw_full[s], p.raw_bytes);
                tma_load_2d_plain(r_addr(s), &tmap_w, smem_u32(&raw_full[s]), (kb0 + i) * (kBlockK * BITS / 8), n0, pol_w);
                if (++s == p.raw_stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== activation tiles =====
        if (lane == 0) {
            prefetch_tensormap(&tmap_x);
            const uint64_t pol_x = policy_evict_last();      // activations are re-read by every CTA
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&x_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&x_full[s], p.x_bytes);
                tma_load_2d_plain(x_addr(s), &tmap_x, smem_u32(&x_full[s]), (kb0 + i) * kBlockK, m0, pol_x);
                if (++s == p.x_stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (kAFmt << 7) | (AT::kFmt << 10) |
                                   ((uint32_t)(p.mb >> 3) << 17) | ((uint32_t)(kTileN >> 4) << 24);
            int sa = 0, sx = 0;
            uint32_t pa = 0, px = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&x_full[sx], px);              // activation tile landed
                mbar_wait(&a_full[sa], pa);              // A tile dequantized and fenced
                tc_fence_after();
                const uint64_t da = smem_desc_sw128(a_addr(sa));
                const uint64_t db = smem_desc_sw128(x_addr(sx));
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                    umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (i | k) != 0 ? 1u : 0u);   // +32 bytes per K=16 step
                umma_commit(&a_empty[sa]);               // both arrive when the MMAs above have read their operands
                umma_commit(&x_empty[sx]);
                if (++sa == kAStages) { sa = 0; pa ^= 1; }
                if (++sx == p.x_stages) { sx = 0; px ^= 1; }
            }
            umma_commit(&tmem_full);
        }
    } else {
        // ===== dequantize (16 warps in 2 groups), then epilogue =====
        // TMEM lanes are reachable per warp quarter (warp % 4); four warps share a quarter and
        // split the batch columns in the epilogue.
        const int quarter = warp & 3;
        const int epart = (warp - kFirstDequantWarp) >> 2;   // epilogue: which part of the batch columns (0..3)
        // main loop: group g handles K blocks g, g+2, ...; within the group warp dw owns weight rows
        // [16 dw, 16 dw + 16); lane -> (row, half of the 64 K values).  Adjacent lanes read adjacent
        // 16 bytes of raw codes and write disjoint swizzled chunks: no bank conflicts.
        const int group = (warp - kFirstDequantWarp) / kDequantWarps;
        const int dw = (warp - kFirstDequantWarp) % kDequantWarps;
        const int row = 16 * dw + (lane >> 1);
        const int half = lane & 1;
        const int gn_d = n0 + row;
        const float* srow = scale + (int64_t)(gn_d < p.N ? gn_d : 0) * p.scale_stride;
        const float* zrow = zp + (int64_t)(gn_d < p.N ? gn_d : 0) * p.scale_stride;
        // parameters of 4 consecutive scale columns at a time (one 16-byte load each)
        const bool vec4 = (p.scale_stride & 3) == 0 && p.block_shift == 0 && (kb0 & 3) == 0 &&
                          ((reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(zp)) & 15) == 0;
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), z4 = s4;
        if (vec4) {
            s4 = __ldg(reinterpret_cast<const float4*>(srow + kb0));
            z4 = __ldg(reinterpret_cast<const float4*>(zrow + kb0));
        }
        int s = group, r = group;                        // ring sizes are even: the group keeps its slot parity
        uint32_t ph = 0, pr = 0;
        for (int i = group; i < nkb; i += kDequantGroups) {
            float sc, z;
            if (vec4) {
                const int e = i & 3;
                sc = e == 0 ? s4.x : (e == 1 ? s4.y : (e == 2 ? s4.z : s4.w));
                z = e == 0 ? z4.x : (e == 1 ? z4.y : (e == 2 ? z4.z : z4.w));
                if (e >= 2 && i + kDequantGroups < nkb) {     // this group's next K block starts a new group of 4
                    s4 = __ldg(reinterpret_cast<const float4*>(srow + kb0 + (i & ~3) + 4));
                    z4 = __ldg(reinterpret_cast<const float4*>(zrow + kb0 + (i & ~3) + 4));
                }
            } else {
                sc = __ldg(srow + ((kb0 + i) >> p.block_shift));
                z = __ldg(zrow + ((kb0 + i) >> p.block_shift));
            }
            // raw codes -> registers, then hand the raw slot straight back to the TMA ring
            mbar_wait(&raw_full[r], pr);
            uint32_t w[BITS == 4 ? 4 : 8];
            {
                const uint32_t rrow = r_addr(r) + row * (kBlockK * BITS / 8) + (4 * BITS) * half;
                const uint4 rv = lds128g(rrow);
                w[0] = rv.x; w[1] = rv.y; w[2] = rv.z; w[3] = rv.w;
                if (BITS == 8) {
                    const uint4 rv2 = lds128g(rrow + 16);
                    w[4] = rv2.x; w[5] = rv2.y; w[6] = rv2.z; w[7] = rv2.w;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_empty[r]);
            r += kDequantGroups;
            if (r >= p.raw_stages) { r -= p.raw_stages; pr ^= 1; }
            mbar_wait(&a_empty[s], ph ^ 1);              // the MMA that read this A slot kAStages blocks ago is done
            const uint32_t arow = a_addr(s) + row * 128;
            const uint32_t sw = row & 7;
            if (BITS == 4) {
                if (kAFmt == 0) {
                    // fp16 A operand: packed fp16x2 math (11-bit significand keeps scale / zp accurate)
                    const __half2 s2 = __float2half2_rn(sc), z2 = __float2half2_rn(z);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // word = nibbles n0..n7 (8 consecutive K values); registers pair (n_i, n_{i+4})
                        const uint32_t r0 = dq_pair<__half>(w[j] & 0x000F000Fu, s2, z2);
                        const uint32_t r1 = dq_pair<__half>((w[j] >> 4) & 0x000F000Fu, s2, z2);
                        const uint32_t r2 = dq_pair<__half>((w[j] >> 8) & 0x000F000Fu, s2, z2);
                        const uint32_t r3 = dq_pair<__half>((w[j] >> 12) & 0x000F000Fu, s2, z2);
                        const uint32_t c = 4 * half + j;                   // 16-byte chunk = 8 K values
                        sts128(arow + ((c ^ sw) << 4), prmt_(r0, r1, 0x5410u), prmt_(r2, r3, 0x5410u),
                               prmt_(r0, r1, 0x7632u), prmt_(r2, r3, 0x7632u));
                    }
                } else {
                    // bf16 activations: bf16 has too few significand bits for scale / zero-point, so the
                    // multiply-add runs in fp32.  (0x4300 | n) is the bf16 (and, shifted, the fp32) value
                    // 128 + n; the offset is folded into the zero-point: w = (128 + n) * s + (z - 128 s).
                    const float zf = __fmaf_rn(-128.0f, sc, z);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // bytes [n0 n2 n4 n6] and [n1 n3 n5 n7]; one PRMT drops a nibble into bits 16..19 of
                        // 0x43000000, i.e. builds the fp32 value 128 + n
                        const uint32_t ev = w[j] & 0x0F0F0F0Fu, od = (w[j] >> 4) & 0x0F0F0F0Fu;
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            f[2 * e] = __fmaf_rn(__uint_as_float(prmt_(ev, 0x43000000u, 0x7044u | (e << 8))), sc, zf);
                            f[2 * e + 1] = __fmaf_rn(__uint_as_float(prmt_(od, 0x43000000u, 0x7044u | (e << 8))), sc, zf);
                        }
                        const uint32_t c = 4 * half + j;
                        sts128(arow + ((c ^ sw) << 4), pack_a<kAFmt>(f[0], f[1]), pack_a<kAFmt>(f[2], f[3]),
                               pack_a<kAFmt>(f[4], f[5]), pack_a<kAFmt>(f[6], f[7]));
                    }
                }
            } else {
                // this thread's 32 K values = 32 code bytes (w[0..7])
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        // q exact via the 2^23 magic; q*scale + zp with one fp32 rounding (the reference rounds
                        // twice: <= 1 fp32 ulp apart, far inside the 16-bit rounding that follows)
                        const uint32_t word = w[2 * j + (e >> 2)];
                        const float q = __fsub_rn(__uint_as_float(prmt_(word, 0x4B000000u, 0x7540u + (e & 3))), 8388608.0f);
                        f[e] = __fmaf_rn(q, sc, z);
                    }
                    const uint32_t c = 4 * half + j;
                    sts128(arow + ((c ^ sw) << 4), pack_a<kAFmt>(f[0], f[1]), pack_a<kAFmt>(f[2], f[3]),
                           pack_a<kAFmt>(f[4], f[5]), pack_a<kAFmt>(f[6], f[7]));
                }
            }
            fence_proxy_async_smem();                    // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[s]);
            s += kDequantGroups;
            if (s >= kAStages) { s -= kAStages; ph ^= 1; }
        }

        // ---- epilogue: TMEM lane = output feature; the quarter's two warps split the batch columns ----
        mbar_wait(&tmem_full, 0);
        tc_fence_after();
        const int gn = n0 + 32 * quarter + lane;
        const bool n_ok = gn < p.N;
        const uint32_t taddr = tmem_d + ((uint32_t)(32 * quarter) << 16);
        const float b = (bias != nullptr && n_ok && p.split_k == 1) ? AT::to_float(bias[gn]) : 0.0f;
        // mb is a multiple of 16: 4 column parts when mb % 32 == 0, else 2 parts (8-column TMEM loads)
        const int nparts = (p.mb & 31) == 0 ? 4 : 2;
        const int cpp = p.mb / nparts;
        const int c_begin = epart < nparts ? epart * cpp : 0, c_end = epart < nparts ? c_begin + cpp : 0;
        for (int c0 = c_begin; c0 < c_end; c0 += 8) {
            uint32_t r[8];
            tmem_ld8(taddr + c0, r);
            tmem_ld_wait();
            if (n_ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int m = m0 + c0 + j;
                    if (m < p.M) {
                        const float v = __uint_as_float(r[j]);
                        if (p.split_k == 1) y[(int64_t)m * p.N + gn] = AT::from_float(v + b);
                        else partial[((int64_t)split * p.M + m) * p.N + gn] = v;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_d, p.tmem_cols); }
}

// y[m,n] = sum_s partial[s][m][n] + bias[n]
template <typename ACT>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, const ACT* __restrict__ bias,
                                                            ACT* __restrict__ y, int64_t MN, int N, int split_k) {
    using AT = ActTraits<ACT>;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= MN) return;
    float acc = 0.0f;
    for (int s = 0; s < split_k; ++s) acc += partial[(int64_t)s * MN + i];
    if (bias) acc += AT::to_float(bias[i % N]);
    y[i] = AT::from_float(acc);
}

constexpr int kMaxSplitK = 8;

// Pick the K split that minimises (waves over the 148 SMs) x (K blocks per CTA + fixed
// prologue/epilogue cost, ~6 K-block equivalents); ties go to the smaller split.
static int choose_split_k(int n_tiles, int m_tiles, int total_kblocks) {
    int best = 1;
    int64_t best_cost = -1;
    for (int s = 1; s <= kMaxSplitK; ++s) {
        if (total_kblocks % s) continue;
        if (s > 1 && total_kblocks / s < 4) break;
        const int64_t ctas = (int64_t)n_tiles * m_tiles * s;
        const int64_t waves = (ctas + kNumSMs - 1) / kNumSMs;
        const int64_t cost = waves * (total_kblocks / s + 6);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
    }
    return best;
}

size_t gemm_workspace_bytes(int64_t M, int64_t N);
size_t gemm_workspace_bytes(int64_t M, int64_t N) {
    return (size_t)kMaxSplitK * (size_t)M * (size_t)N * sizeof(float) + 512;     // fp32 split-K partials
}
size_t int8_outlier_workspace_bytes(int64_t, int64_t) { return 256; }

template <typename ACT, int BITS, uint32_t kAFmt>
static cudaError_t launch_gemm_kernel(dim3 grid, int smem, cudaStream_t st, const CUtensorMap& tw, const CUtensorMap& tx,
                                      const float* scale, const float* zp, const ACT* bias, ACT* y, float* partial,
                                      const GemmParams& p) {
    auto kern = gemm_wna16_kernel<ACT, BITS, kAFmt>;
    static int smem_set = 0;
    if (smem > smem_set) {           // static __shared__ (barriers) also counts against the 227 KB opt-in limit
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        smem_set = smem;
    }
    kern<<<grid, kGemmThreads, smem, st>>>(tw, tx, scale, zp, bias, y, partial, p);
    return cudaGetLastError();
}

template <typename ACT, int BITS>
static int gemm_launch(const ACT* x, const uint8_t* wq, const float* scale, const float* zp, int64_t block,
                       const ACT* bias, ACT* y, int64_t M, int64_t N, int64_t K, void* workspace, size_t ws_bytes,
                       cudaStream_t st) {
    GemmParams p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    int mb = (int)((M + 15) / 16 * 16);
    if (mb > 256) mb = 256;
    p.mb = mb;
    const int m_tiles = (int)((M + mb - 1) / mb);
    const int n_tiles = (int)((N + kTileN - 1) / kTileN);
    const int total_kb = (int)(K / kBlockK);
    int split = choose_split_k(n_tiles, m_tiles, total_kb);
    if (const char* e = getenv("QUANTA_B200_SPLIT_K")) { int v = atoi(e); if (v >= 1 && total_kb % v == 0) split = v; }
    p.split_k = split;
    p.kblocks_per_split = total_kb / split;
    p.scale_stride = (int)(K / block);
    int bs = 0; while ((int64_t)(kBlockK << bs) < block) ++bs;
    p.block_shift = bs;
    p.x_bytes = (uint32_t)mb * 128u;
    p.raw_bytes = (uint32_t)(kTileN * kBlockK * BITS / 8);
    // Shared-memory budget: a shallow A ring (dequant -> MMA latency only), an activation ring, and
    // everything else for the raw-code ring: raw bytes in flight are what hides the HBM latency.
    const uint32_t budget = 218u * 1024u;
    const uint32_t a_ring = (uint32_t)kAStages * kATileBytes;
    int x_stages = mb <= 64 ? 8 : (mb <= 128 ? 4 : 3);
    int raw_stages = (int)((budget - a_ring - x_stages * p.x_bytes) / p.raw_bytes);
    if (raw_stages > kMaxRawStages) raw_stages = kMaxRawStages;
    raw_stages &= ~1;                                    // the two dequant groups alternate slots
    if (raw_stages < 2) return QUANTA_EUNSUPPORTED;
    p.raw_stages = raw_stages;
    p.x_stages = x_stages;
    p.x_ring_off = a_ring;
    p.raw_ring_off = a_ring + (uint32_t)x_stages * p.x_bytes;
    int cols = 32; while (cols < mb) cols <<= 1;
    p.tmem_cols = cols;
    float* partial = nullptr;
    if (split > 1) {
        const size_t need = (size_t)split * M * N * sizeof(float);
        if (!workspace || ws_bytes < need + 256) return QUANTA_EWORKSPACE;
        partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    }

    CUtensorMap tmap_w, tmap_x;
    const uint64_t wrow_bytes = (uint64_t)K * BITS / 8;
    int rc = make_tensor_map_2d(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, wq, wrow_bytes, (uint64_t)N, wrow_bytes,
                                (uint32_t)(kBlockK * BITS / 8), kTileN, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    rc = make_tensor_map_2d(&tmap_x, ActTraits<ACT>::kTma, 2, x, (uint64_t)K, (uint64_t)M, (uint64_t)K * 2, kBlockK,
                            (uint32_t)mb, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;

    // A-operand format = activation format.  (A mixed fp16-A x bf16-B MMA was tried so that bf16
    // activations could use the cheaper packed-fp16 dequant math: tcgen05.mma kind::f16 raises an
    // illegal-instruction fault on sm_100a when the A and B formats differ.)
    const int smem = (int)(p.raw_ring_off + (uint32_t)p.raw_stages * p.raw_bytes + 1024);
    dim3 grid(n_tiles, split, m_tiles);
    cudaError_t e;
    if (ActTraits<ACT>::kFmt == 0)
        e = launch_gemm_kernel<ACT, BITS, 0>(grid, smem, st, tmap_w, tmap_x, scale, zp, bias, y, partial, p);
    else
        e = launch_gemm_kernel<ACT, BITS, 1>(grid, smem, st, tmap_w, tmap_x, scale, zp, bias, y, partial, p);
    if (e != cudaSuccess) return (int)e;
    if (split > 1) {
        const int64_t MN = M * N;
        splitk_reduce_kernel<ACT><<<(unsigned)((MN + 255) / 256), 256, 0, st>>>(partial, bias, y, MN, (int)N, split);
        e = cudaGetLastError();
    }
    return cuda_status(e);
}

}  // namespace quanta

using namespace quanta;

extern "C" int quanta_gemm_wna16(const void* x, int act_dtype, const uint8_t* wq, int bits, const float* scale,
                                 const float* zp, int64_t block, const void* bias, void* y, int64_t M, int64_t N,
                                 int64_t K, void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !wq || !scale || !zp || !y || M <= 0 || N <= 0 || K <= 0) return QUANTA_EINVAL;
    if ((bits != 4 && bits != 8) || block <= 0 || block % kBlockK != 0 || K % block != 0) return QUANTA_EINVAL;
    if ((block / kBlockK) & (block / kBlockK - 1)) return QUANTA_EUNSUPPORTED;          // block = 64 * 2^j
    if ((K * bits / 8) % 16 != 0 || (K * 2) % 16 != 0) return QUANTA_EUNSUPPORTED;      // TMA row pitch
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wq)) & 15) return QUANTA_EUNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (act_dtype == QUANTA_BF16) {
        using T = __nv_bfloat16;
        return bits == 4 ? gemm_launch<T, 4>((const T*)x, wq, scale, zp, block, (const T*)bias, (T*)y, M, N, K, workspace, workspace_bytes, st)
                         : gemm_launch<T, 8>((const T*)x, wq, scale, zp, block, (const T*)bias, (T*)y, M, N, K, workspace, workspace_bytes, st);
    }
    if (act_dtype == QUANTA_F16) {
        using T = __half;
        return bits == 4 ? gemm_launch<T, 4>((const T*)x, wq, scale, zp, block, (const T*)bias, (T*)y, M, N, K, workspace, workspace_bytes, st)
                         : gemm_launch<T, 8>((const T*)x, wq, scale, zp, block, (const T*)bias, (T*)y, M, N, K, workspace, workspace_bytes, st);
    }
    return QUANTA_EINVAL;
}

extern "C" int quanta_int8_outlier_matmul(const void*, int, const int8_t*, const float*, float, const void*, void*,
                                          int64_t, int64_t, int64_t, void*, size_t, void*) {
    return QUANTA_EUNSUPPORTED;
}
