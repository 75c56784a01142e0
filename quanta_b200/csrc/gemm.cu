// W4A16 / W8A16 dequantize-then-matmul on the 5th-generation tensor cores
// (rows G1/G2 of SURVEY §8): y[M,N] = x[M,K] . dequant(Wq)[N,K]^T + bias.
//
// Mapping ("swap-AB"): the 128 output features of a weight tile are the UMMA M
// dimension (one TMEM lane each), the batch rows are the UMMA N dimension
// (16..256), so small batches do not waste the 128-row MMA:
//     D[n, m] (fp32, TMEM) += A[n, k] (dequantized weights, TMEM) * B[m, k] (x, smem)
// The dequantized A operand never touches shared memory: the dequant warps write
// it straight into tensor memory (tcgen05.st) and the MMA reads A from TMEM.
//
// Persistent, warp-specialised CTA (one per SM), 24 warps:
//   warp 0      TMA ring of raw weight codes: stages of [128 rows x 128 B] (SWIZZLE_128B),
//               i.e. 256 K values of 4-bit or 128 of 8-bit codes per box
//   warp 1      TMA ring of activation tiles [mb x 64] (SWIZZLE_128B, K-major B operand)
//   warp 2      TMEM allocation; one lane issues tcgen05.mma.kind::f16 (4 x K=16 per 64-K block)
//   warps 4-7   epilogue: tcgen05.ld the accumulator, add bias, store y — or store an fp32
//               partial and let the last-arriving CTA of the tile reduce all partials
//   warps 8-23  dequantize: 2 groups x 8 warps; a group owns every other stage and its own half of
//               the A slots and raw slots (so every mbarrier has one producer and one consumer
//               side and parity waits are unambiguous); a thread owns half a 64-K block of one
//               weight row (= its TMEM lane): 32 codes -> 16-bit A values -> one tcgen05.st
// Work is split stream-K style: the (tile, K-stage) units are cut into G equal
// contiguous ranges, one per CTA, so all SMs stream weights for the same time.
#include "common.cuh"

#include <cmath>
#include <cstdlib>
#include <cstdio>

namespace quanta {

constexpr int kTileN = 128;          // output features per tile (UMMA M, TMEM lanes)
constexpr int kBlockK = 64;          // K per A slot / MMA group: one 128-byte swizzle atom of 16-bit activations
constexpr int kKbPerStage = 4;       // 64-K blocks per raw-code stage
constexpr int kStageK = kBlockK * kKbPerStage;
constexpr int kASlots = 8;           // dequantized A tiles resident in TMEM
constexpr int kACols = kBlockK / 2;  // 32-bit TMEM columns per A slot (two 16-bit values per column)
constexpr int kDBase = kASlots * kACols;   // accumulators start at TMEM column 256
constexpr int kTmemCols = 512;
constexpr int kCtrlWarps = 4, kEpiWarps = 4, kDqGroups = 2, kDqGroupWarps = 8, kDqWarps = kDqGroupWarps * kDqGroups;
constexpr int kFirstEpiWarp = kCtrlWarps, kFirstDqWarp = kCtrlWarps + kEpiWarps;
constexpr int kGemmThreads = 32 * (kCtrlWarps + kEpiWarps + kDqWarps);
constexpr int kMaxRing = 12;
constexpr int kCounterBytes = 64 * 1024;     // per-tile arrival counters at the head of the workspace
constexpr int kMaxTiles = kCounterBytes / 4;

__device__ __forceinline__ uint4 lds128g(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t prmt_(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r;
}
// (a & b) | c in one LOP3
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
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
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// completes `bar` (one arrival) when all previously issued MMAs of this thread have finished
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

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
    static constexpr float kCentre = 136.0f;             // 128 + 8
    static constexpr CUtensorMapDataType kTma = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    using V2 = __nv_bfloat162;
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
    static constexpr float kCentre = 1032.0f;            // 1024 + 8
    static constexpr CUtensorMapDataType kTma = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    using V2 = __half2;
    __device__ static __forceinline__ V2 dup(float v) { return __float2half2_rn(v); }
    __device__ static __forceinline__ uint32_t pack(float a, float b) {
        V2 t = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&t);
    }
    __device__ static __forceinline__ float to_float(__half v) { return __half2float(v); }
    __device__ static __forceinline__ __half from_float(float v) { return __float2half_rn(v); }
};

// 4-bit: one 32-bit word = nibbles n0..n7 = 8 consecutive K values (pack_4bit_tensor layout).
// (magic | nibble) is the 16-bit float 128+n (bf16) / 1024+n (fp16), exactly; subtracting the
// centre leaves n-8 exactly, and one packed fma gives (n-8)*s + (z+8s) = n*s + z with the scale
// and the block midpoint z+8s rounded to the activation type (the midpoint is near zero for
// weight-like data, so its rounding error is far below the final 16-bit rounding of the weight).
template <typename ACT>
__device__ __forceinline__ void dequant_word4(uint32_t w, typename ActTraits<ACT>::V2 s2, typename ActTraits<ACT>::V2 z2,
                                              typename ActTraits<ACT>::V2 c2, uint32_t* out) {
    using V2 = typename ActTraits<ACT>::V2;
    constexpr uint32_t kM = ActTraits<ACT>::kMagic;
    uint32_t t[4];
    t[0] = and_or(w, 0x000F000Fu, kM);             // (n0, n4)
    t[1] = and_or(w >> 4, 0x000F000Fu, kM);        // (n1, n5)
    t[2] = and_or(w >> 8, 0x000F000Fu, kM);        // (n2, n6)
    t[3] = and_or(w >> 12, 0x000F000Fu, kM);       // (n3, n7)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        V2 v = *reinterpret_cast<V2*>(&t[i]);
        v = __hfma2(__hsub2(v, c2), s2, z2);
        t[i] = *reinterpret_cast<uint32_t*>(&v);
    }
    out[0] = prmt_(t[0], t[1], 0x5410u);           // (k0, k1)
    out[1] = prmt_(t[2], t[3], 0x5410u);           // (k2, k3)
    out[2] = prmt_(t[0], t[1], 0x7632u);           // (k4, k5)
    out[3] = prmt_(t[2], t[3], 0x7632u);           // (k6, k7)
}

// 8-bit: PRMT drops a code byte into bits 8..15 of 0x47000000, i.e. builds the fp32 value
// 32768 + q exactly; fma with zc = z - 32768*s gives q*s + z to within 2^-9 of a scale step,
// far inside the 16-bit rounding that follows (the reference rounds mul and add separately).
template <typename ACT>
__device__ __forceinline__ void dequant_word8(uint32_t w, float s, float zc, uint32_t* out) {
    float f[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
        f[e] = __fmaf_rn(__uint_as_float(prmt_(w, 0x47000000u, 0x7504u | (e << 4))), s, zc);
    out[0] = ActTraits<ACT>::pack(f[0], f[1]);
    out[1] = ActTraits<ACT>::pack(f[2], f[3]);
}

__constant__ float kGemmNf4Levels[16] = {
    -1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
    -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
    0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
    0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};

constexpr int kMaxOut = 8;
// NOUT = 1 for the plain GEMM (one map, compile-time loops), kMaxOut for the scatter entry
template <int NOUT>
struct GemmOutputs {
    CUtensorMap maps[NOUT];         // TMA-store maps of the output buffers ([M, col0 + N] windows, pitch ldy)
    void* y[NOUT];
};

struct GemmParams {
    int M, N, K;
    int mb;                 // UMMA N: batch rows per tile (multiple of 16, <= 256)
    int n_tiles, m_tiles;
    int total_kb;           // K / 64
    int S;                  // raw-code stages per tile: ceil(total_kb / 4)
    int G;                  // CTAs
    unsigned int U;         // work units: tiles * S  (U * G < 2^32, checked on the host)
    unsigned int uq, ur;    // U / G, U % G: CTA c owns units [c uq + min(c, ur), + uq + (c < ur))
    unsigned int s_magic;   // ceil(2^32 / S) (0: S == 1): u / S without a division — a 32-bit division is ~100
    unsigned int nt_magic;  // ceil(2^32 / n_tiles) (0: n_tiles == 1)   instructions on every role's critical path
    int my_acc;             // nacc / nmma
    int nbuf;               // accumulator buffers in TMEM (1 or 2)
    int nacc;               // independent accumulators per buffer (power of 2): consecutive MMAs rotate over them
    int nmma;               // MMA-issuing warps (2 when nacc >= 2: tcgen05.mma dispatch is per-warp bound)
    int raw_stages, x_stages;
    int xkb;                // 64-K blocks per activation slot: 4, 2 or 1 (slot <= 32 KB)
    uint32_t raw_bytes, x_kb_bytes, x_slot_bytes, x_ring_off;
    int scale_stride;       // K / block
    int block_shift;        // log2(block / 64)
    int vec4;               // scale / zero-point rows can be read as float4 per stage
    int vec_store;          // 8-byte stores of 4 outputs are aligned in every output buffer
    int y_tma;              // whole tiles leave through a TMA store (needs N % 8 == 0 and an aligned y)
    int n_out;              // output buffers (1, or one per tensor-parallel peer: the tile is written to all of them)
    int ldy;                // row pitch of y in elements (N, or the full out_features of a column-parallel layer)
    int col0;               // first output column of this call inside y
    int nf4;                // 4-bit codes index the NF4 table, `scale` holds abs_max per block, zero-point unused
    int dbg;                // experiment switches (QUANTA_B200_GEMM_DBG): 1 = no MMA, 2 = no dequant math/store
};

// Division-free index arithmetic (exactness of the multiplications is checked on the host, gemm_launch).
__device__ __forceinline__ unsigned int div_magic(unsigned int u, unsigned int magic) { return magic ? __umulhi(u, magic) : u; }
__device__ __forceinline__ unsigned int first_unit(unsigned int c, const GemmParams& p) { return c * p.uq + (c < p.ur ? c : p.ur); }
__device__ __forceinline__ int tile_of_unit(unsigned int u, const GemmParams& p) { return (int)div_magic(u, p.s_magic); }
// tile -> (n tile, m tile): tiles are numbered n-fastest
__device__ __forceinline__ void split_tile(int tile, const GemmParams& p, int& n_idx, int& m_idx) {
    m_idx = (int)div_magic((unsigned int)tile, p.nt_magic);
    n_idx = tile - m_idx * p.n_tiles;
}

// Contiguous range of work units of one CTA, walked tile by tile.
struct SegWalk {
    unsigned int u, u1, S, magic;
    __device__ __forceinline__ void init(const GemmParams& p, unsigned int cta) {
        u = first_unit(cta, p); u1 = first_unit(cta + 1u, p); S = (unsigned int)p.S; magic = p.s_magic;
    }
    // next segment: stages [s0, s1) of `tile`; false when the range is exhausted
    __device__ __forceinline__ bool next(int& tile, int& s0, int& s1) {
        if (u >= u1) return false;
        const unsigned int t = div_magic(u, magic);
        const unsigned int b = u - t * S, left = u1 - u;
        tile = (int)t; s0 = (int)b;
        s1 = (left < S - b) ? (int)(b + left) : (int)S;
        u += (unsigned int)(s1 - s0);
        return true;
    }
};

// CTA whose unit range contains unit `u`
__device__ __forceinline__ int cta_of_unit(unsigned int u, const GemmParams& p) {
    const unsigned int big = p.ur * (p.uq + 1u);
    return (int)(u < big ? u / (p.uq + 1u) : p.ur + (u - big) / p.uq);      // reducers only, once per tile
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(pred));
    return pred != 0;
}

// ---- CTA-pair helpers (CG = 2) ------------------------------------------------------------
// Two CTAs of a cluster work on ADJACENT 128-feature tiles over the same K range and the same batch rows.  Everything
// is private to a CTA (weight ring, dequant groups, TMEM, MMA issue, epilogue) except the activation tile, which both
// need: each CTA loads HALF of its rows and multicasts them into both CTAs' rings, so the pair pulls every activation
// byte out of L2 once instead of twice (at 256 batch rows the activation traffic — 128 KB per 16 KB of codes — is what
// bounds the kernel).  A slot may be refilled when BOTH CTAs' MMAs have read it: the commit that frees it is multicast
// too.  (The first form of the pair, one cta_group::2 MMA for both CTAs, coupled the two dequant pipelines through one
// barrier and measured 10-25 % slower than single CTAs; it is gone.)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// completes `bar` in this CTA, or at the same offset in both CTAs of the pair
template <int CG> __device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    if (CG == 2) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    } else {
        umma_commit(bar);
    }
}
// activation tile load; in a pair the rows land in both CTAs and signal both CTAs' barriers (same offsets)
template <int CG>
__device__ __forceinline__ void tma_load_x(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1, uint64_t policy) {
    if (CG == 2) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
            " [%0], [%1, {%4, %5}], [%2], %3, %6;"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "h"((uint16_t)3), "r"(c0), "r"(c1), "l"(policy)
            : "memory");
    } else {
        tma_load_2d_plain(dst, map, bar, c0, c1, policy);
    }
}

// Stream-K fix-up of one tile by `nthreads` threads (thread j): sum the fp32 partials of every
// contributing CTA in a fixed order (deterministic), add the bias, write y.  Thread -> 4 consecutive
// features, rows j/32, j/32 + nthreads/32, ...; kFixRows rows (float4 each) in flight per contributor.
constexpr int kFixRows = 6;
template <typename ACT, int CG, int NOUT>
__device__ __noinline__ void fixup_reduce(const GemmParams& p, const GemmOutputs<NOUT>& outs, const ACT* __restrict__ bias,
                                             const float* __restrict__ partial, int tile, int crank, int j, int nthreads) {
    using AT = ActTraits<ACT>;
    int n_idx, m_tile;
    split_tile(tile, p, n_idx, m_tile);
    const int n_tile = n_idx * CG + crank;
    const int m0 = m_tile * p.mb;
    const int m_valid = min(p.mb, p.M - m0);
    const int f4 = 4 * (j & 31), mq = j >> 5, mstep = nthreads >> 5;
    const int gn4 = n_tile * kTileN + f4;
    float b4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) b4[e] = (bias != nullptr && gn4 + e < p.N) ? AT::to_float(bias[gn4 + e]) : 0.0f;
    const bool vec_ok = p.vec_store && gn4 + 3 < p.N;
    const unsigned int tu0 = (unsigned int)tile * (unsigned int)p.S;
    const int c_first = cta_of_unit(tu0, p), c_last = cta_of_unit(tu0 + (unsigned int)p.S - 1u, p);
    const size_t slot_elems = (size_t)(kTileN * p.mb);
    for (int mb0 = mq; mb0 < m_valid; mb0 += kFixRows * mstep) {
        float4 acc[kFixRows];
#pragma unroll
        for (int i = 0; i < kFixRows; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = c_first; c <= c_last; ++c) {
            const int wc = (tile == tile_of_unit(first_unit((unsigned int)c, p), p)) ? 0 : 1;
            const float* src = partial + (((size_t)c * 2 + wc) * CG + crank) * slot_elems + f4;
            float4 v[kFixRows];
#pragma unroll
            for (int i = 0; i < kFixRows; ++i) {
                const int m = mb0 + mstep * i;
                v[i] = (m < m_valid) ? __ldcg(reinterpret_cast<const float4*>(src + m * kTileN))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < kFixRows; ++i) { acc[i].x += v[i].x; acc[i].y += v[i].y; acc[i].z += v[i].z; acc[i].w += v[i].w; }
        }
#pragma unroll
        for (int i = 0; i < kFixRows; ++i) {
            const int m = mb0 + mstep * i;
            if (m < m_valid) {
                const float o[4] = {acc[i].x + b4[0], acc[i].y + b4[1], acc[i].z + b4[2], acc[i].w + b4[3]};
                for (int ob = 0; ob < (NOUT == 1 ? 1 : p.n_out); ++ob) {
                    ACT* dst = static_cast<ACT*>(outs.y[ob]) + (int64_t)(m0 + m) * p.ldy + p.col0 + gn4;
                    if (vec_ok) {
                        uint2 pk;
                        pk.x = AT::pack(o[0], o[1]);
                        pk.y = AT::pack(o[2], o[3]);
                        *reinterpret_cast<uint2*>(dst) = pk;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) if (gn4 + e < p.N) dst[e] = AT::from_float(o[e]);
                    }
                }
            }
        }
    }
}

template <typename ACT, int BITS, int CG, int NOUT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_wna16_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                  const __grid_constant__ CUtensorMap tmap_s, const __grid_constant__ CUtensorMap tmap_z,
                  const __grid_constant__ GemmOutputs<NOUT> outs, const float* __restrict__ scale, const float* __restrict__ zp, const ACT* __restrict__ bias,
                  unsigned int* __restrict__ counters, float* __restrict__ partial,
                  const __grid_constant__ GemmParams p) {
    using AT = ActTraits<ACT>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t raw_full[kMaxRing], raw_empty[kMaxRing], x_full[kMaxRing], x_empty[kMaxRing];
    __shared__ uint64_t a_full[kDqGroups], a_empty[kDqGroups], d_full[2], d_empty[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ int fix_flag;
    __shared__ int fin_tile;            // tile whose stream-K fix-up is done by the whole CTA at the end (-1: none)
    __shared__ __align__(1024) float stage[16][kTileN];        // epilogue transpose buffer (8 KB)
    __shared__ uint32_t nf4_pairs[256];                        // NF4: byte -> (level[lo nibble], level[hi nibble]) in the act type
#ifdef QUANTA_GEMM_TRACE
    __shared__ long long trace[8][48];
    const bool tr = (p.dbg & 8) && blockIdx.x == 0;
    int tri = 0;
    const long long t_start = clock64();
#define TRACE2(role, cond) do { if (tr && (cond) && tri < 48) trace[role][tri++] = clock64() - t_start; } while (0)
#define TRACE(role) do { if (tr && tri < 48) trace[role][tri++] = clock64() - t_start; } while (0)
#else
#define TRACE2(role, cond) do { } while (0)
#define TRACE(role) do { } while (0)
#endif

    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp-uniform for the compiler
    // CG = 2: the two CTAs of a cluster form a pair: they walk the same unit range on adjacent 128-feature tiles
    // and share the activation loads (see the pair helpers above).  `cta` is the scheduling unit (CTA or pair).
    const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;
    const unsigned int cta = blockIdx.x / CG;

    // shared memory: [raw ring: raw_stages x 16|32 KB][x ring: x_stages x (xkb x mb x 128 B)]
    auto r_addr = [&](int s) { return smem + (uint32_t)s * p.raw_bytes; };
    auto x_addr = [&](int s) { return smem + p.x_ring_off + (uint32_t)s * p.x_slot_bytes; };

    SegWalk walk;
    walk.init(p, cta);
    int tile, s0, s1;
    if (tid == 32 * kFirstEpiWarp) fin_tile = -1;       // before the CTA-wide barrier below
    if (BITS == 4 && p.nf4 && tid >= 32 * kFirstDqWarp && tid < 32 * kFirstDqWarp + 256) {
        const int b = tid - 32 * kFirstDqWarp;      // visible to the dequant warps after the CTA-wide barrier below
        nf4_pairs[b] = AT::pack(kGemmNf4Levels[b & 15], kGemmNf4Levels[b >> 4]);
    }

    // Programmatic dependent launch: the next kernel on the stream may start its prologue while this grid is in
    // its tail; symmetrically this grid may have started during the previous kernel's tail, so everything that
    // depends on that kernel — the activation loads, bias, every global write — sits behind griddepcontrol.wait.
    // The weight stream (codes, scales, zero-points) does not and starts at once.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // Each producer initialises its own ring, checks in at the CTA-wide barrier without waiting
    // (barrier.arrive) and starts streaming at once; everybody else sees all barriers after
    // barrier.sync.  Producer loops are warp-uniform with one elected lane issuing, so the TMA
    // operands stay in uniform registers.
    if (warp == 0) {
        // ===== raw weight codes: deep TMA ring (bytes in flight hide the HBM latency) =====
        constexpr uint32_t kRawCodeBytes = BITS == 4 ? 16384u : 32768u;
        if (lane == 0) {
            for (int s = 0; s < p.raw_stages; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], kDqGroupWarps); }
            fence_barrier_init();
            prefetch_tensormap(&tmap_w);
            if (p.vec4) { prefetch_tensormap(&tmap_s); prefetch_tensormap(&tmap_z); }
        }
        __syncwarp();
        if (CG == 2) cluster_arrive();                       // checked in; the matching wait comes after the stream
        else asm volatile("barrier.arrive 2, %0;" ::"n"(kGemmThreads) : "memory");
        const uint64_t pol_w = policy_evict_first();         // weights are streamed once
        int slot = 0, issued = 0;
        uint32_t ph = 0;
        while (walk.next(tile, s0, s1)) {
            int n_idx, m_idx;
            split_tile(tile, p, n_idx, m_idx);
            const int n0 = (n_idx * CG + (int)crank) * kTileN;
            for (int s = s0; s < s1; ++s) {
                if (issued >= p.raw_stages) mbar_wait(&raw_empty[slot], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&raw_full[slot], p.raw_bytes);
                    const uint32_t bar = smem_u32(&raw_full[slot]);
                    if (BITS == 4) {
                        tma_load_2d_plain(r_addr(slot), &tmap_w, bar, s * 128, n0, pol_w);
                    } else {
                        tma_load_2d_plain(r_addr(slot), &tmap_w, bar, s * 256, n0, pol_w);
                        tma_load_2d_plain(r_addr(slot) + 16384u, &tmap_w, bar, s * 256 + 128, n0, pol_w);
                    }
                    if (p.vec4) {
                        // the stage's scales and zero-points ride on the same barrier: [128 rows x 4 floats] each
                        tma_load_2d_plain(r_addr(slot) + kRawCodeBytes, &tmap_s, bar, s * kKbPerStage, n0, pol_w);
                        tma_load_2d_plain(r_addr(slot) + kRawCodeBytes + 2048u, &tmap_z, bar, s * kKbPerStage, n0, pol_w);
                    }
                    TRACE(0);
                }
                __syncwarp();
                ++issued;
                if (++slot == p.raw_stages) { slot = 0; ph ^= 1; }
            }
        }
        if (CG == 2) cluster_wait();
    } else if (warp == 1) {
        // ===== activation tiles: xkb 64-K blocks per slot =====
        if (lane == 0) {
            for (int s = 0; s < p.x_stages; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], CG); }   // freed by both CTAs' MMAs
            fence_barrier_init();
            prefetch_tensormap(&tmap_x);
        }
        __syncwarp();
        if (CG == 2) { cluster_arrive(); cluster_wait(); }   // the peer's barriers must exist before the first load
        else asm volatile("barrier.arrive 2, %0;" ::"n"(kGemmThreads) : "memory");
        const uint64_t pol_x = policy_evict_last();          // activations are re-read by every CTA
        asm volatile("griddepcontrol.wait;" ::: "memory");   // x may be the previous kernel's output
        int slot = 0, issued = 0;
        uint32_t ph = 0;
        while (walk.next(tile, s0, s1)) {
            int n_idx, m_idx;
            split_tile(tile, p, n_idx, m_idx);
            const int m0 = m_idx * p.mb + (int)crank * (p.mb / CG);                  // this CTA loads (and multicasts) its half of the rows
            const uint32_t half_off = crank * (uint32_t)(p.mb / CG) * 128u;          // where they go inside a 64-K block of the slot
            for (int s = s0; s < s1; ++s) {
                for (int j = 0; j < kKbPerStage; j += p.xkb) {
                    if (issued >= p.x_stages) mbar_wait(&x_empty[slot], ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&x_full[slot], p.x_slot_bytes);               // all rows: own loads + the peer's
                        // blocks past the end of K are out of bounds and arrive zero-filled
                        for (int jj = 0; jj < p.xkb; ++jj)
                            tma_load_x<CG>(x_addr(slot) + (uint32_t)jj * p.x_kb_bytes + half_off, &tmap_x, smem_u32(&x_full[slot]),
                                           (s * kKbPerStage + j + jj) * kBlockK, m0, pol_x);
                    }
                    __syncwarp();
                    TRACE2(3, lane == 0);
                    ++issued;
                    if (++slot == p.x_stages) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else {
        if (warp == 2) {
            tmem_alloc(smem_u32(&tmem_base_slot), kTmemCols);
        } else if (warp == 3 && lane == 0) {
            for (int s = 0; s < kDqGroups; ++s) { mbar_init(&a_full[s], kDqGroupWarps); mbar_init(&a_empty[s], 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(&d_full[s], p.nmma); mbar_init(&d_empty[s], kEpiWarps); }
            fence_barrier_init();
        }
        __syncwarp();
        tc_fence_before();
        if (CG == 2) { cluster_arrive(); cluster_wait(); }
        else asm volatile("barrier.sync 2, %0;" ::"n"(kGemmThreads) : "memory");
    }
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;

    if (warp == 2 || (warp == 3 && p.nmma == 2)) {
        // ===== MMA issuers: the whole warp walks the schedule, one elected lane issues =====
        // A tcgen05.mma costs ~70 cycles of dispatch whatever its size (the tensor pipe itself needs only
        // N/2 cycles), so a stage is >= 1120 cycles of issue plus ~600 cycles of barrier waits and
        // commits.  With two issuers (small batches) warp 2 takes the even stages and warp 3 the odd
        // ones — one dequant group each — so one warp's waits hide behind the other's issue; each
        // warp rotates over its own half of the accumulators (the epilogue adds all of them up).
        const int mw = warp - 2;
        const uint32_t idesc = (1u << 4) | (AT::kFmt << 7) | (AT::kFmt << 10) |
                               ((uint32_t)(p.mb >> 3) << 17) | ((uint32_t)(kTileN >> 4) << 24);
        const uint32_t kb_desc = p.x_kb_bytes >> 4;          // descriptor step between 64-K blocks of one slot
        const uint32_t my_acc = (uint32_t)p.my_acc;                   // accumulators of this warp (power of 2)
        int sx = 0, seg = 0, sc = 0;
        uint32_t px = 0;
        while (walk.next(tile, s0, s1)) {
            const int buf = p.nbuf == 2 ? (seg & 1) : 0;
            const uint32_t use = (uint32_t)(p.nbuf == 2 ? (seg >> 1) : seg);      // how often `buf` was used before
            mbar_wait(&d_empty[buf], (use & 1) ^ 1);   // the epilogues have drained this accumulator
            const uint32_t tmem_d = tmem + kDBase + (uint32_t)((buf * p.nacc + mw * (int)my_acc) * p.mb);
            uint32_t touched = 0;                            // own accumulators already written in this segment
            for (int s = s0; s < s1; ++s, ++sc) {
                const int nkb = min(kKbPerStage, p.total_kb - s * kKbPerStage);
                const int g = sc & 1;                        // dequant group = A buffer
                if (p.nmma == 1 || g == mw) mbar_wait(&a_full[g], (uint32_t)(sc >> 1) & 1u);   // the stage's 4 A tiles are in TMEM
                const uint32_t ta = tmem + (uint32_t)(g * kKbPerStage * kACols);
                TRACE2(2, lane == 0);
#pragma unroll
                for (int j = 0; j < kKbPerStage; ++j) {
                    const int jj = j & (p.xkb - 1);
                    const bool mine = p.nmma == 1 || (sc & 1) == mw;
                    if (mine) {
                        if (jj == 0) mbar_wait(&x_full[sx], px);       // activation slot landed
                        tc_fence_after();
                        const bool live = j < nkb && !(p.dbg & 1);
                        uint32_t acc_idx[4], acc_flag[4];    // computed by every lane: stays warp-uniform
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            acc_idx[k] = (uint32_t)(4 * j + k) & (my_acc - 1u);
                            acc_flag[k] = (touched >> acc_idx[k]) & 1u;
                            if (live) touched |= 1u << acc_idx[k];
                        }
                        if (elect_one()) {
                            if (live) {
                                const uint64_t db = smem_desc_sw128(x_addr(sx)) + (uint64_t)(jj * kb_desc);
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k)    // K = 16 per MMA: 8 TMEM columns of A, 32 bytes of B
                                    umma_f16_ts(tmem_d + acc_idx[k] * (uint32_t)p.mb, ta + (uint32_t)(j * kACols + 8 * k),
                                                db + 2 * k, idesc, acc_flag[k]);
                            }
                            // both arrive when the MMAs above have read their operands
                            if (jj == p.xkb - 1) umma_commit_pair<CG>(&x_empty[sx]);   // in a pair: frees the slot in both CTAs
                            if (j == kKbPerStage - 1) umma_commit(&a_empty[g]);
                        }
                        __syncwarp();
                    }
                    if (jj == p.xkb - 1) { if (++sx == p.x_stages) { sx = 0; px ^= 1; } }
                }
                TRACE2(2, lane == 0);
            }
            if (elect_one()) umma_commit(&d_full[buf]);
            __syncwarp();
            ++seg;
        }
    } else if (warp >= kFirstEpiWarp && warp < kFirstDqWarp) {
        // ===== epilogue: TMEM lane = output feature =====
        asm volatile("griddepcontrol.wait;" ::: "memory");   // bias reads, y / partial / counter writes
        const int quarter = warp & 3;
        const int row = 32 * quarter + lane;
        const int etid = tid - 32 * kFirstEpiWarp;
        if (etid == 0 && p.y_tma) { for (int o = 0; o < (NOUT == 1 ? 1 : p.n_out); ++o) prefetch_tensormap(&outs.maps[o]); }
        int seg = 0, esc = 0;                 // esc: CTA-wide stage counter at the start of the segment
        while (walk.next(tile, s0, s1)) {
            const int buf = p.nbuf == 2 ? (seg & 1) : 0;
            const uint32_t use = (uint32_t)(p.nbuf == 2 ? (seg >> 1) : seg);
            int n_idx, m_tile;
            split_tile(tile, p, n_idx, m_tile);
            const int n_tile = n_idx * CG + (int)crank;
            const int gn = n_tile * kTileN + row, m0 = m_tile * p.mb;
            const bool n_ok = gn < p.N;
            const int m_valid = min(p.mb, p.M - m0);
            const bool whole = (s0 == 0 && s1 == p.S);
            const bool last_seg = walk.u >= walk.u1;                  // no further segment for this CTA
            const float b = (bias != nullptr && n_ok) ? AT::to_float(bias[gn]) : 0.0f;
            // this CTA's partial slot for the tile: 0 if the tile is the first one the CTA touches, else 1
            const int which = (tile == tile_of_unit(first_unit(cta, p), p)) ? 0 : 1;
            float* mine = partial + (((size_t)cta * 2 + which) * CG + crank) * (size_t)(kTileN * p.mb);

            TRACE2(4, etid == 0);
            mbar_wait(&d_full[buf], use & 1);
            tc_fence_after();
            TRACE2(4, etid == 0);
            const uint32_t taddr = tmem + ((uint32_t)(32 * quarter) << 16) + kDBase + (uint32_t)(buf * p.nacc * p.mb);
            // store mapping: thread -> 4 consecutive features (f4..f4+3), batch rows etid/32 + 4q
            const int f4 = 4 * (etid & 31);
            const int gn4 = n_tile * kTileN + f4;
            float b4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) b4[e] = (bias != nullptr && gn4 + e < p.N) ? AT::to_float(bias[gn4 + e]) : 0.0f;
            const bool vec_ok = p.vec_store && gn4 + 3 < p.N;
            // accumulators the MMA warps wrote in this segment: each warp issues 4 MMAs per 64-K block of
            // its half of the stage (all of it with one issuer), rotating over its own my_acc accumulators
            const int my_acc = p.my_acc;
            const int last_nkb = (s1 == p.S) ? p.total_kb - (p.S - 1) * kKbPerStage : kKbPerStage;
            int kb_w[2];
            if (p.nmma == 1) { kb_w[0] = (s1 - s0 - 1) * kKbPerStage + last_nkb; kb_w[1] = 0; }
            else {
                // warp w issued the stages whose CTA-wide stage counter has parity w
                kb_w[0] = kb_w[1] = 0;
                for (int s = s0; s < s1; ++s) kb_w[(esc + (s - s0)) & 1] += (s == p.S - 1) ? last_nkb : kKbPerStage;
            }
            for (int c0 = 0; c0 < p.mb; c0 += 16) {
                uint32_t r[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = 0u;
                for (int w = 0; w < p.nmma; ++w) {
                    const int nv = min(my_acc, 4 * kb_w[w]);
                    for (int a = 0; a < nv; ++a) {
                        uint32_t t[16];
                        tmem_ld16(taddr + (uint32_t)((w * my_acc + a) * p.mb) + c0, t);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(t[j]));
                    }
                }
                if (whole && p.y_tma) {
                    // Whole tile: bias, 16-bit conversion, [16 rows x 128 features] staged in shared memory
                    // (two 4 KB buffers), then ONE TMA store per chunk; rows past M / features past N are
                    // clipped by the tensor map.  (Plain global stores from this kernel cost ~10 cycles
                    // per warp-store per SM: 20K cycles for a 128 x 256 tile.)
                    ACT* sbuf = reinterpret_cast<ACT*>(&stage[0][0]) + ((c0 >> 4) & 1) * (16 * kTileN);
                    if (etid == 0 && c0 >= 32) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // buffer's previous store has read it
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) sbuf[j * kTileN + row] = AT::from_float(__uint_as_float(r[j]) + b);
                    fence_proxy_async_smem();
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                    if (etid == 0 && c0 < m_valid) {
                        // one store per output buffer: the local y, or every tensor-parallel peer's y over NVLink
                        for (int o = 0; o < (NOUT == 1 ? 1 : p.n_out); ++o)
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"(reinterpret_cast<uint64_t>(&outs.maps[o])), "r"(smem_u32(sbuf)),
                                           "r"(p.col0 + n_tile * kTileN), "r"(m0 + c0) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    continue;
                }
                // Otherwise transpose the chunk through shared memory so that the stores are full rows:
                // thread -> 4 consecutive features of 4 batch rows, 8 / 16 bytes per store.
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");      // previous chunk fully read
#pragma unroll
                for (int j = 0; j < 16; ++j) stage[j][row] = __uint_as_float(r[j]);
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                if (c0 < m_valid) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int ml = 4 * q + (etid >> 5);              // batch row within the chunk
                        const int m = c0 + ml;
                        if (m < m_valid) {
                            const float4 v = *reinterpret_cast<const float4*>(&stage[ml][f4]);
                            if (whole) {
                                const float o[4] = {v.x + b4[0], v.y + b4[1], v.z + b4[2], v.w + b4[3]};
                                for (int ob = 0; ob < (NOUT == 1 ? 1 : p.n_out); ++ob) {
                                    ACT* dst = static_cast<ACT*>(outs.y[ob]) + (int64_t)(m0 + m) * p.ldy + p.col0 + gn4;
                                    if (vec_ok) {
                                        uint2 pk;
                                        pk.x = AT::pack(o[0], o[1]);
                                        pk.y = AT::pack(o[2], o[3]);
                                        *reinterpret_cast<uint2*>(dst) = pk;
                                    } else {
#pragma unroll
                                        for (int e = 0; e < 4; ++e) if (gn4 + e < p.N) dst[e] = AT::from_float(o[e]);
                                    }
                                }
                            } else {
                                *reinterpret_cast<float4*>(mine + m * kTileN + f4) = v;
                            }
                        }
                    }
                }
            }
            if (whole && p.y_tma && etid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging free for the next segment
            TRACE2(4, etid == 0);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[buf]);   // accumulator may be overwritten
            if (!whole) {
                // stream-K fix-up: the CTA that arrives last at the tile's counter reduces every partial
                const unsigned int tu0 = (unsigned int)tile * (unsigned int)p.S;
                const int c_first = cta_of_unit(tu0, p), c_last = cta_of_unit(tu0 + (unsigned int)p.S - 1u, p);
                // release: CTA barrier, then one thread publishes with a gpu-scope fence + atomic (fences
                // are cumulative over the barrier); acquire on the way back mirrors it
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                if (etid == 0) {
                    __threadfence();
                    const unsigned int old = atomicAdd(&counters[tile * CG + (int)crank], 1u);
                    __threadfence();
                    const int last = (old == (unsigned int)(c_last - c_first)) ? 1 : 0;
                    if (last) counters[tile * CG + (int)crank] = 0u;      // leave the workspace clean for the next call
                    fix_flag = last;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                if (fix_flag) {
                    if (last_seg && p.mb >= 64) {
                        // the CTA has nothing left to do: hand the reduction to all 20 epilogue + dequant
                        // warps (5x the loads in flight) after the role loops
                        if (etid == 0) fin_tile = tile;
                    } else {
                        fixup_reduce<ACT, CG, NOUT>(p, outs, bias, partial, tile, (int)crank, etid, 32 * kEpiWarps);
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");      // fix_flag is reused
            }
            esc += s1 - s0;
            ++seg;
        }
    } else if (warp >= kFirstDqWarp) {
        // ===== dequantize: thread = (weight row = TMEM lane, half of each 64-K block) =====
        using V2 = typename AT::V2;
        const int quarter = warp & 3;                                  // TMEM lane quarter this warp may access
        const int half = ((warp - kFirstDqWarp) >> 2) & 1;            // which 32 of the 64 K values
        const int group = (warp - kFirstDqWarp) / kDqGroupWarps;      // stages group, group + 2, ...
        const int row = 32 * quarter + lane;
        // this group's A buffer: 4 tiles of 32 columns; this thread's half tile starts 16 columns in
        const uint32_t a_addr = tmem + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(group * kKbPerStage * kACols + 16 * half);
        const V2 c2 = AT::dup(AT::kCentre);
        // Scale / zero-point of a stage: delivered by TMA next to the raw codes (vec4), else loaded
        // with plain loads one stage of work ahead (odd shapes: K not a multiple of 256, block > 64).
        constexpr uint32_t kRawCodeBytes = BITS == 4 ? 16384u : 32768u;
        SegWalk ahead = walk;
        int a_tile = 0, a_s = 0, a_s1 = 0;
        bool a_ok = ahead.next(a_tile, a_s, a_s1);
        float sv[4], zv[4];
        auto fetch_params = [&](int t, int s) {
            int n_idx, m_idx;
            split_tile(t, p, n_idx, m_idx);
            const int gn = (n_idx * CG + (int)crank) * kTileN + row;
            const int64_t rbase = (int64_t)(gn < p.N ? gn : p.N - 1) * p.scale_stride;
            const int kb = s * kKbPerStage;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = min(kb + j, p.total_kb - 1) >> p.block_shift;
                sv[j] = __ldg(scale + rbase + col);
                zv[j] = __ldg(zp + rbase + col);
            }
        };
        // advance `ahead` by n stages (across segments)
        auto skip = [&](int n) {
            while (a_ok && n > 0) {
                const int room = a_s1 - a_s;
                if (n < room) { a_s += n; n = 0; }
                else { n -= room; a_ok = ahead.next(a_tile, a_s, a_s1); }
            }
        };
        if (!p.vec4) {
            skip(group);
            if (a_ok) fetch_params(a_tile, a_s);
        }

        int sc = 0;                 // global stage counter of this CTA
        int rslot = 0;
        uint32_t rph = 0;
        while (walk.next(tile, s0, s1)) {
            for (int s = s0; s < s1; ++s, ++sc) {
                if ((sc & (kDqGroups - 1)) == group) {
                    const int nkb = min(kKbPerStage, p.total_kb - s * kKbPerStage);
                    float cs[4], cz[4];
                    if (!p.vec4) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) { cs[j] = sv[j]; cz[j] = zv[j]; }
                        skip(kDqGroups);
                        if (a_ok) fetch_params(a_tile, a_s); // lands while this stage is processed
                    }
                    TRACE2(1, lane == 0 && quarter == 0 && half == 0 && group == 0 && sc >= 4 && sc <= 8);
                    mbar_wait(&raw_full[rslot], rph);
                    TRACE2(1, lane == 0 && quarter == 0 && half == 0 && group == 0 && sc >= 4 && sc <= 8);

                    const uint32_t rrow = r_addr(rslot) + (uint32_t)row * 128u;
                    const uint32_t sw = (uint32_t)(row & 7);
                    if (p.vec4) {
                        const uint4 a = lds128g(r_addr(rslot) + kRawCodeBytes + (uint32_t)row * 16u);
                        const uint4 c = lds128g(r_addr(rslot) + kRawCodeBytes + 2048u + (uint32_t)row * 16u);
                        cs[0] = __uint_as_float(a.x); cs[1] = __uint_as_float(a.y); cs[2] = __uint_as_float(a.z); cs[3] = __uint_as_float(a.w);
                        cz[0] = __uint_as_float(c.x); cz[1] = __uint_as_float(c.y); cz[2] = __uint_as_float(c.z); cz[3] = __uint_as_float(c.w);
                    }
                    // the MMAs that read this group's A buffer two stages ago are done
                    mbar_wait(&a_empty[group], ((uint32_t)(sc >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    TRACE2(1, lane == 0 && quarter == 0 && half == 0 && group == 0 && sc >= 4 && sc <= 8);
#pragma unroll
                    for (int j = 0; j < kKbPerStage; ++j) {
                        if (j < nkb && !(p.dbg & 2)) {
                            uint32_t out[16];
                            if (BITS == 4 && p.nf4) {
                                // NF4: one shared-memory lookup per code byte gives the K-adjacent pair
                                // (level[lo], level[hi]); one packed multiply applies the block's abs_max
                                const V2 s2 = AT::dup(cs[j]);
                                const uint4 rv = lds128g(rrow + (((uint32_t)(2 * j + half) ^ sw) << 4));
                                const uint32_t w4[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        uint32_t pr = nf4_pairs[(w4[i] >> (8 * e)) & 0xFFu];
                                        V2 v = __hmul2(*reinterpret_cast<V2*>(&pr), s2);
                                        out[4 * i + e] = *reinterpret_cast<uint32_t*>(&v);
                                    }
                                }
                            } else if (BITS == 4) {
                                const V2 s2 = AT::dup(cs[j]);
                                const V2 z2 = AT::dup(__fmaf_rn(8.0f, cs[j], cz[j]));
                                const uint4 rv = lds128g(rrow + (((uint32_t)(2 * j + half) ^ sw) << 4));
                                dequant_word4<ACT>(rv.x, s2, z2, c2, out);
                                dequant_word4<ACT>(rv.y, s2, z2, c2, out + 4);
                                dequant_word4<ACT>(rv.z, s2, z2, c2, out + 8);
                                dequant_word4<ACT>(rv.w, s2, z2, c2, out + 12);
                            } else {
                                const float zc = __fmaf_rn(-32768.0f, cs[j], cz[j]);
                                const uint32_t box = rrow + (uint32_t)(j >> 1) * 16384u;
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const uint4 rv = lds128g(box + (((uint32_t)(4 * (j & 1) + 2 * half + h) ^ sw) << 4));
                                    dequant_word8<ACT>(rv.x, cs[j], zc, out + 8 * h);
                                    dequant_word8<ACT>(rv.y, cs[j], zc, out + 8 * h + 2);
                                    dequant_word8<ACT>(rv.z, cs[j], zc, out + 8 * h + 4);
                                    dequant_word8<ACT>(rv.w, cs[j], zc, out + 8 * h + 6);
                                }
                            }
                            tmem_st16(a_addr + (uint32_t)(j * kACols), out);
                        }
                        TRACE2(1, lane == 0 && quarter == 0 && half == 0 && group == 0 && sc >= 4 && sc <= 8);
                    }
                    tmem_st_wait();
                    TRACE2(1, lane == 0 && quarter == 0 && half == 0 && group == 0 && sc >= 4 && sc <= 8);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&a_full[group]);
                        mbar_arrive(&raw_empty[rslot]);
                    }

                }
                if (++rslot == p.raw_stages) { rslot = 0; rph ^= 1; }
            }
        }
    }

    if (warp >= kFirstEpiWarp && p.mb >= 64) {
        // epilogue + dequant warps: a pending last-segment fix-up is reduced by all of them (small
        // batches have too few rows to share: the epilogue warps did it in their loop)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("bar.sync 3, %0;" ::"n"(32 * (kEpiWarps + kDqWarps)) : "memory");
        const int ft = *reinterpret_cast<volatile int*>(&fin_tile);
        if (ft >= 0)
            fixup_reduce<ACT, CG, NOUT>(p, outs, bias, partial, ft, (int)crank, tid - 32 * kFirstEpiWarp, 32 * (kEpiWarps + kDqWarps));
    }
    if (tid == 32 * kFirstEpiWarp) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // y stores have landed
    tc_fence_before();
    __syncthreads();
    if (CG == 2) { cluster_arrive(); cluster_wait(); }       // the peer may still be signalling into this CTA
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem, kTmemCols); }
#ifdef QUANTA_GEMM_TRACE
    if (tr && tid == 0) {
        printf("end %lld\n", clock64() - t_start);
        for (int r = 0; r < 8; ++r) { printf("role %d:", r); for (int i = 0; i < 48; ++i) printf(" %lld", trace[r][i]); printf("\n"); }
    }
#endif
}

// ---- host side --------------------------------------------------------------


// Number of scheduling units (CTAs, or CTA pairs when cg = 2): every SM with equal contiguous unit
// ranges (stream-K), or tiles x s units whose ranges coincide with tile boundaries (s = 1: no
// partials at all).  Costs in SM cycles, from the measurements in DESIGN.md: a stage (256 K) costs
// the MMA warp ~600 cycles of barrier traffic plus 16 MMAs of max(70, mb/2) cycles each, the
// dequant groups ~930 cycles, and one SM pulls its raw codes and its share of the activation
// tiles (re-read by every tile) out of L2 at ~29 B/cycle; a partial costs its fp32 store, ~4000
// cycles of release/acquire latency and the last arriver's (latency-bound) reads of every
// contributor's tile.
static int choose_units(int tiles, int S, int mb, int raw_bytes, int cg) {
    const int max_units = kNumSMs / cg;
    const long long U = (long long)tiles * S;
    const double mma = 600.0 + 16.0 * (mb / 2.0 > 70.0 ? mb / 2.0 : 70.0);
    const double pull = (512.0 * mb / cg + raw_bytes) / 29.0;
    const double stage = fmax(fmax(mma, 930.0), pull);
    const double tile_io = (double)kTileN * mb * 4.0 / 32.0;
    int best = 1;
    double best_cost = -1.0;
    auto consider = [&](int G) {
        if (G < 1 || G > max_units || (long long)G > U) return;
        double cost = (double)((U + G - 1) / G) * stage;
        if (G != tiles) {
            const double contributors = (double)G / tiles < 2.0 ? 2.0 : (double)G / tiles + 1.0;
            cost += 4000.0 + 2.0 * tile_io * (1.0 + contributors);
            // ranges that do not coincide with tile boundaries give most CTAs TWO partial tiles (two
            // accumulator drains and partial stores, uneven contributor counts): measured ~6 tile
            // transfers slower than the tile-aligned split on 4096 x 14336 at M = 64 / 128 / 256
            if (G % tiles != 0) cost += 6.0 * tile_io;
        }
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = G; }
    };
    consider(tiles <= max_units ? tiles : 0);
    for (int s = 2; tiles * s <= max_units && s <= S; ++s) consider(tiles * s);
    consider((int)(U < max_units ? U : max_units));
    return best;
}

size_t gemm_workspace_bytes(int64_t M, int64_t N);
size_t gemm_workspace_bytes(int64_t M, int64_t) {
    int64_t mb = (M + 31) / 32 * 32;
    if (mb > 256) mb = 256;
    return (size_t)kCounterBytes + (size_t)kNumSMs * 2 * kTileN * (size_t)mb * sizeof(float) + 512;
}

template <typename ACT, int BITS, int CG, int NOUT>
static int gemm_launch_cg(GemmParams& p, const ACT* x, const uint8_t* wq, const float* scale, const float* zp,
                          const ACT* bias, void* const* ys, int64_t M, int64_t N, int64_t K, void* workspace, size_t ws_bytes,
                          cudaStream_t st) {
    const int mb = p.mb;
    const int tiles = p.n_tiles * p.m_tiles;               // scheduling tiles: CG adjacent 128-feature tiles each
    if (tiles * CG > kMaxTiles) return QUANTA_EUNSUPPORTED;
    if ((unsigned long long)tiles * (unsigned long long)p.S * (unsigned long long)(kNumSMs + 1) >= (1ull << 32)) return QUANTA_EUNSUPPORTED;
    p.U = (unsigned int)tiles * (unsigned int)p.S;
    p.G = choose_units(tiles, p.S, mb, BITS == 4 ? 16384 : 32768, CG);
    { const int v = env_int("QUANTA_B200_GEMM_CTAS", 0); if (v >= 1 && v <= kNumSMs / CG && (unsigned int)v <= p.U) p.G = v; }
    p.uq = p.U / (unsigned int)p.G;
    p.ur = p.U % (unsigned int)p.G;
    {
        // x / d == __umulhi(x, ceil(2^32 / d)) for every x <= xmax when xmax * (magic * d - 2^32) < 2^32
        auto magic_for = [](unsigned int d, unsigned long long xmax, unsigned int* out) {
            if (d == 1) { *out = 0u; return true; }
            const unsigned long long m = ((1ull << 32) + d - 1ull) / d;
            *out = (unsigned int)m;
            return xmax * (m * d - (1ull << 32)) < (1ull << 32);
        };
        if (!magic_for((unsigned int)p.S, p.U, &p.s_magic) || !magic_for((unsigned int)p.n_tiles, (unsigned long long)tiles, &p.nt_magic))
            return QUANTA_EUNSUPPORTED;
    }
    p.my_acc = p.nacc / p.nmma;
    p.x_kb_bytes = (uint32_t)mb * 128u;                     // one 64-K block of the activation tile (all rows: a pair multicasts)
    p.xkb = p.x_kb_bytes <= 8192u ? 4 : (p.x_kb_bytes <= 16384u ? 2 : 1);    // activation slots of at most 32 KB
    p.x_slot_bytes = p.x_kb_bytes * (uint32_t)p.xkb;
    p.raw_bytes = (BITS == 4 ? 16384u : 32768u) + (p.vec4 ? 4096u : 0u);    // codes (+ scale / zero-point tiles)
    // Shared-memory budget: an activation ring sized for the MMA, everything else for the raw-code
    // ring — raw bytes in flight are what hides the HBM latency.
    const uint32_t budget = 214u * 1024u;
    int x_stages = p.x_slot_bytes <= 8192u ? 6 : (p.x_slot_bytes <= 16384u ? 4 : (p.x_kb_bytes <= 8192u ? 3 : 4));
    int raw_stages = (int)((budget - (uint32_t)x_stages * p.x_slot_bytes) / p.raw_bytes);
    if (raw_stages > kMaxRing) raw_stages = kMaxRing;
    raw_stages &= ~1;                  // even: a raw slot is always consumed by the same dequant group
    if (raw_stages < 2) return QUANTA_EUNSUPPORTED;
    p.raw_stages = raw_stages;
    p.x_stages = x_stages;
    p.x_ring_off = (uint32_t)raw_stages * p.raw_bytes;

    const size_t need = (size_t)kCounterBytes + (size_t)p.G * CG * 2 * kTileN * (size_t)mb * sizeof(float) + 256;
    if (!workspace || ws_bytes < need) return QUANTA_EWORKSPACE;
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
    float* partial = reinterpret_cast<float*>(ws + kCounterBytes);

    CUtensorMap tmap_w, tmap_x;
    const uint64_t wrow_bytes = (uint64_t)K * BITS / 8;
    int rc = make_tensor_map_2d(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, wq, wrow_bytes, (uint64_t)N, wrow_bytes,
                                128, kTileN, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tensor_map_2d(&tmap_x, ActTraits<ACT>::kTma, 2, x, (uint64_t)K, (uint64_t)M, (uint64_t)K * 2, kBlockK,
                            (uint32_t)(mb / CG), CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;

    CUtensorMap tmap_s = tmap_w, tmap_z = tmap_w;          // placeholders unless the parameters ride on TMA
    if (p.vec4) {
        rc = make_tensor_map_2d(&tmap_s, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, scale, (uint64_t)p.scale_stride, (uint64_t)N,
                                (uint64_t)p.scale_stride * 4, kKbPerStage, kTileN, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
        rc = make_tensor_map_2d(&tmap_z, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, zp, (uint64_t)p.scale_stride, (uint64_t)N,
                                (uint64_t)p.scale_stride * 4, kKbPerStage, kTileN, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
    }
    GemmOutputs<NOUT> outs;
    bool aligned16 = true, aligned8 = true;
    for (int o = 0; o < NOUT; ++o) {
        outs.maps[o] = tmap_w;                              // placeholder
        outs.y[o] = o < p.n_out ? ys[o] : nullptr;
        if (o < p.n_out) {
            aligned16 = aligned16 && (reinterpret_cast<uintptr_t>(ys[o]) & 15) == 0;
            aligned8 = aligned8 && (reinterpret_cast<uintptr_t>(ys[o]) & 7) == 0;
        }
    }
    p.vec_store = (aligned8 && (p.ldy & 3) == 0 && (p.col0 & 3) == 0) ? 1 : 0;
    p.y_tma = (aligned16 && (p.ldy & 7) == 0) ? 1 : 0;
    if (env_int("QUANTA_B200_GEMM_YTMA", 1) == 0) p.y_tma = 0;
    if (p.y_tma) {
        // window [M, col0 + N] of each buffer: features past this call's columns are clipped by the map
        for (int o = 0; o < p.n_out; ++o) {
            rc = make_tensor_map_2d(&outs.maps[o], ActTraits<ACT>::kTma, 2, ys[o], (uint64_t)(p.col0 + N), (uint64_t)M,
                                    (uint64_t)p.ldy * 2, kTileN, 16, CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc) return rc;
        }
    }
    auto kern = gemm_wna16_kernel<ACT, BITS, CG, NOUT>;
    const int smem = (int)(p.x_ring_off + (uint32_t)p.x_stages * p.x_slot_bytes + 1024);
    if (int e = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem)) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.G * CG));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (env_int("QUANTA_B200_GEMM_PDL", 1)) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (CG == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CG; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmap_w, tmap_x, tmap_s, tmap_z, outs, scale, zp, bias, counters, partial, p);
    return cuda_status(e != cudaSuccess ? e : cudaGetLastError());
}

// mma.sync weight-stream kernel for decode-sized batches (gemm_small.cu)
bool gemm_small_eligible(int64_t M, int64_t N, int64_t K, int64_t block, const void* scale, const void* zp);
template <typename ACT, int BITS>
int gemm_small_launch(const ACT* x, const uint8_t* wq, const float* scale, const float* zp, const ACT* bias, void* const* ys,
                      int n_out, int64_t ldy, int64_t col0, int64_t M, int64_t N, int64_t K, void* workspace, size_t ws_bytes,
                      cudaStream_t st, int nf4, const PeerSync* sync);

template <typename ACT, int BITS>
static int gemm_launch(const ACT* x, const uint8_t* wq, const float* scale, const float* zp, int64_t block,
                       const ACT* bias, ACT* y, int64_t M, int64_t N, int64_t K, void* workspace, size_t ws_bytes,
                       cudaStream_t st, int nf4 = 0, void* const* ys = nullptr, int n_out = 1, int64_t ldy = 0,
                       int64_t col0 = 0, const PeerSync* sync = nullptr) {
    void* one[1] = {y};
    if (!ys) { ys = one; n_out = 1; ldy = N; col0 = 0; }
    // M <= 16: the weight-stream kernel (HBM-bound regime; gemm_small.cu)
    if (gemm_small_eligible(M, N, K, block, scale, zp))
        return gemm_small_launch<ACT, BITS>(x, wq, scale, zp, bias, ys, n_out, ldy, col0, M, N, K, workspace, ws_bytes, st, nf4, sync);
    if (sync) return QUANTA_EUNSUPPORTED;                    // the in-kernel completion exists in the small-batch kernel only
    GemmParams p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    const int n_tiles = (int)((N + kTileN - 1) / kTileN);
    // CTA pairs (CG = 2: adjacent feature tiles, the activation tile loaded once and multicast to both CTAs) halve the
    // L2 reads of the activations — and measured no faster (M = 256: 42.8 vs 42.1 us on 4096 x 14336, 32.1 vs 31.8 on
    // 14336 x 4096; M = 64 / 128 1-4 % slower): every SM still has to take in the whole 128 KB tile per 256-K stage
    // (41 B/cycle at the measured stage time), so the bound is the SM's inbound path, not L2.  Off unless
    // QUANTA_B200_GEMM_PAIR=1 (kept working by tests/test_gpu_gemm.py::test_cta_pair_multicast_mode).
    int cg = 1;
    if (env_int("QUANTA_B200_GEMM_PAIR", 0) == 1 && n_tiles >= 2 && n_out == 1) cg = 2;
    const int mb_step = 16 * cg;                             // each CTA of a pair holds mb / 2 rows, a multiple of 16
    int mb = (int)((M + mb_step - 1) / mb_step * mb_step);
    if (mb > 256) mb = 256;
    { const int cap = env_int("QUANTA_B200_GEMM_MB", 0); if (cap >= 16 && cap % 16 == 0 && mb > cap) mb = cap; }
    p.mb = mb;
    p.m_tiles = (int)((M + mb - 1) / mb);
    p.n_tiles = (n_tiles + cg - 1) / cg;
    p.total_kb = (int)(K / kBlockK);
    p.S = (p.total_kb + kKbPerStage - 1) / kKbPerStage;
    // accumulators: nbuf x nacc x mb <= 256 TMEM columns
    p.nacc = mb <= 16 ? 8 : (mb <= 64 ? 4 : (mb <= 128 ? 2 : 1));
    p.nbuf = (2 * p.nacc * mb <= kTmemCols - kDBase) ? 2 : 1;
    // a second MMA-issuing warp (even / odd stages) is implemented but measured no faster: the body is
    // bound by the dequant groups, not by the issuer (QUANTA_B200_GEMM_NMMA=2 enables it)
    p.nmma = 1;
    if (env_int("QUANTA_B200_GEMM_NMMA", 1) == 2 && p.nacc >= 2) p.nmma = 2;
    p.dbg = 0;
    p.nf4 = nf4;
    p.n_out = n_out; p.ldy = (int)ldy; p.col0 = (int)col0;
    p.dbg = env_int("QUANTA_B200_GEMM_DBG", 0);
    p.scale_stride = (int)(K / block);
    int bs = 0; while ((int64_t)(kBlockK << bs) < block) ++bs;
    p.block_shift = bs;
    p.vec4 = (bs == 0 && (p.scale_stride & 3) == 0 &&
              ((reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(zp)) & 15) == 0) ? 1 : 0;
    if (n_out > 1) return gemm_launch_cg<ACT, BITS, 1, kMaxOut>(p, x, wq, scale, zp, bias, ys, M, N, K, workspace, ws_bytes, st);
    if (cg == 2) return gemm_launch_cg<ACT, BITS, 2, 1>(p, x, wq, scale, zp, bias, ys, M, N, K, workspace, ws_bytes, st);
    return gemm_launch_cg<ACT, BITS, 1, 1>(p, x, wq, scale, zp, bias, ys, M, N, K, workspace, ws_bytes, st);
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

extern "C" int quanta_gemm_nf4a16(const void* x, int act_dtype, const uint8_t* wq, const float* absmax, int64_t block,
                                  const void* bias, void* y, int64_t M, int64_t N, int64_t K, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    if (!x || !wq || !absmax || !y || M <= 0 || N <= 0 || K <= 0) return QUANTA_EINVAL;
    if (block <= 0 || block % kBlockK != 0 || K % block != 0) return QUANTA_EINVAL;
    if ((block / kBlockK) & (block / kBlockK - 1)) return QUANTA_EUNSUPPORTED;
    if ((K / 2) % 16 != 0 || (K * 2) % 16 != 0) return QUANTA_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wq)) & 15) return QUANTA_EUNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (act_dtype == QUANTA_BF16) {
        using T = __nv_bfloat16;
        return gemm_launch<T, 4>((const T*)x, wq, absmax, absmax, block, (const T*)bias, (T*)y, M, N, K, workspace, workspace_bytes, st, 1);
    }
    if (act_dtype == QUANTA_F16) {
        using T = __half;
        return gemm_launch<T, 4>((const T*)x, wq, absmax, absmax, block, (const T*)bias, (T*)y, M, N, K, workspace, workspace_bytes, st, 1);
    }
    return QUANTA_EINVAL;
}

static int gemm_scatter_entry(const void* x, int act_dtype, const uint8_t* wq, int bits, const float* scale,
                                         const float* zp, int64_t block, const void* bias, void* const* ys, int n_out,
                                         int64_t ldy, int64_t col0, int64_t M, int64_t N, int64_t K, void* workspace,
                                         size_t workspace_bytes, void* stream, const PeerSync* sync) {
    if (!x || !wq || !scale || !zp || !ys || n_out < 1 || n_out > kMaxOut || M <= 0 || N <= 0 || K <= 0) return QUANTA_EINVAL;
    if (ldy < col0 + N || col0 < 0) return QUANTA_EINVAL;
    for (int o = 0; o < n_out; ++o) if (!ys[o]) return QUANTA_EINVAL;
    if ((bits != 4 && bits != 8) || block <= 0 || block % kBlockK != 0 || K % block != 0) return QUANTA_EINVAL;
    if ((block / kBlockK) & (block / kBlockK - 1)) return QUANTA_EUNSUPPORTED;
    if ((K * bits / 8) % 16 != 0 || (K * 2) % 16 != 0) return QUANTA_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wq)) & 15) return QUANTA_EUNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (act_dtype == QUANTA_BF16) {
        using T = __nv_bfloat16;
        return bits == 4 ? gemm_launch<T, 4>((const T*)x, wq, scale, zp, block, (const T*)bias, nullptr, M, N, K, workspace, workspace_bytes, st, 0, ys, n_out, ldy, col0, sync)
                         : gemm_launch<T, 8>((const T*)x, wq, scale, zp, block, (const T*)bias, nullptr, M, N, K, workspace, workspace_bytes, st, 0, ys, n_out, ldy, col0, sync);
    }
    if (act_dtype == QUANTA_F16) {
        using T = __half;
        return bits == 4 ? gemm_launch<T, 4>((const T*)x, wq, scale, zp, block, (const T*)bias, nullptr, M, N, K, workspace, workspace_bytes, st, 0, ys, n_out, ldy, col0, sync)
                         : gemm_launch<T, 8>((const T*)x, wq, scale, zp, block, (const T*)bias, nullptr, M, N, K, workspace, workspace_bytes, st, 0, ys, n_out, ldy, col0, sync);
    }
    return QUANTA_EINVAL;
}

extern "C" int quanta_gemm_wna16_scatter(const void* x, int act_dtype, const uint8_t* wq, int bits, const float* scale,
                                         const float* zp, int64_t block, const void* bias, void* const* ys, int n_out,
                                         int64_t ldy, int64_t col0, int64_t M, int64_t N, int64_t K, void* workspace,
                                         size_t workspace_bytes, void* stream) {
    return gemm_scatter_entry(x, act_dtype, wq, bits, scale, zp, block, bias, ys, n_out, ldy, col0, M, N, K, workspace, workspace_bytes,
                              stream, nullptr);
}

// The same, with the ranks' synchronisation inside the kernel: when it has completed on this rank, every rank's columns
// have landed in this rank's buffer (see PeerSync in common.cuh) and no barrier kernel is needed behind it.
// QUANTA_EUNSUPPORTED when the shape is not served by the small-batch kernel (M > 16, block != 64, K % 256 != 0):
// the caller then uses quanta_gemm_wna16_scatter plus its own barrier.
extern "C" int quanta_gemm_wna16_scatter_sync(const void* x, int act_dtype, const uint8_t* wq, int bits, const float* scale,
                                              const float* zp, int64_t block, const void* bias, void* const* ys, int n_out,
                                              int64_t ldy, int64_t col0, int64_t M, int64_t N, int64_t K, void* workspace,
                                              size_t workspace_bytes, void* const* peer_flags, int rank, int world,
                                              unsigned int* epoch_counter, void* stream) {
    if (!peer_flags || !epoch_counter || world < 2 || world > 8 || rank < 0 || rank >= world) return QUANTA_EINVAL;
    PeerSync sync;
    for (int r = 0; r < 8; ++r) sync.flags[r] = r < world ? peer_flags[r] : nullptr;
    sync.rank = rank; sync.world = world; sync.epoch_counter = epoch_counter;
    return gemm_scatter_entry(x, act_dtype, wq, bits, scale, zp, block, bias, ys, n_out, ldy, col0, M, N, K, workspace, workspace_bytes,
                              stream, &sync);
}
