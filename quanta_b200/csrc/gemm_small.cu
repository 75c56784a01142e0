// W4A16 / W8A16 dequantize-then-matmul for decode-sized batches (M <= 16; rows G1/G2 of SURVEY §8):
//     y[M,N] = x[M,K] . dequant(Wq)[N,K]^T + bias,   Wq blockwise-64 along K (convention A).
//
// At these batch sizes the op is a pure weight stream (5.7 us of HBM time for a
// 4096 x 14336 W4 matrix), so the kernel is built around the stream and keeps everything else
// off its critical path:
//
//   * one persistent CTA per SM.  A producer thread keeps a ring of TMA stages full: SWIZZLE_128B
//     tiles of [128 rows x 128 B] codes, the step's [128 x 4] scale / zero-point tiles and the step's
//     activations [m_pad x 256] all on one mbarrier.  It starts before the rest of the CTA is set up
//     and requests only three stages until the first one has landed (the TMA unit works on all of
//     its outstanding copies at once: eight stages requested together by 148 SMs all complete
//     together, ~4.5 K cycles later).  The (row tile, 256-K step) units are cut into equal
//     contiguous ranges, one per CTA (stream-K), so every SM streams for the same time whatever
//     the shape;
//   * 16 consumer warps drain the ring (warp = 32 weight rows x one 64-K block of every step, so
//     a B fragment read from shared memory serves two row slabs).  Codes are never dequantized one
//     by one: LOP3 drops a nibble pair into the mantissas of the 16-bit constant 128.0 (bf16) /
//     1024.0 (fp16) — exact integers C + n — and the warp-level tensor-core MMA (mma.sync m16n8k16,
//     fp32 accumulate) forms raw = sum_k (C + n_k) x_k over one 64-K block; scale and zero-point are
//     applied once per block on the accumulators:
//         y += s * raw + (z - C s) * sum_k x_k
//     i.e. 7 integer instructions per 8 weights + 2 FMAs per accumulator, instead of 19
//     instructions per 8 weights for an element-wise dequantization.  The MMA's K order is free as
//     long as both operands agree, so x is staged in the order the LOP3 pairs come out
//     (k0,k4 | k1,k5 | k2,k6 | k3,k7) and no PRMT is needed.  8-bit codes go through
//     PRMT -> fp32 (32768 + q) -> q -> packed 16-bit, exact as well;
//   * four staging warps turn the raw activations of a stage into MMA fragment order plus the
//     per-block sums of x, in a 4-slot ring of their own (full / empty mbarriers), so the
//     consumers' instruction stream is the weight path only;
//   * a CTA that covers only part of a row tile's K range stores an fp32 partial and bumps the
//     tile's counter with ONE fire-and-forget release; the CTA that owns the tile's last K steps
//     is its reducer: once its own stream has ended it acquires the counter, sums the partials in
//     CTA order (deterministic), adds the bias and writes y — to every output buffer of a
//     tensor-parallel call (peer-mapped buffers over NVLink).  Nobody pays an atomic round trip
//     between two fences at the end of the kernel (measured 5-9 K cycles per CTA); when a range
//     holds whole tiles its partial segments are processed first, so the kernel ends on a plain
//     store of y.
//
// The products use the exact fp32 value q*s + z of the weight (the tcgen05 path in gemm.cu rounds it
// to the activation type first, like the reference's `.to(x.dtype)`); the difference is far inside
// the 1e-2 tolerance of rows G1/G2.
#include "common.cuh"

#include <cstdio>
#include <cstdlib>

namespace quanta {

constexpr int kSmRows = 128;               // weight rows per tile
constexpr int kSmConsWarps = 16;           // warp w: rows 32 (w & 3) .. +32, block (w >> 2) of every step
constexpr int kSmConsThreads = 32 * kSmConsWarps;
constexpr int kSmXWarps = 4;               // activation staging warps
constexpr int kSmThreads = kSmConsThreads + 64 + 32 * kSmXWarps;   // + producer warp + publisher warp + staging warps
constexpr int kSmStepK = 256;              // K per stage
constexpr int kSmMaxRing = 8;
constexpr int kSmXRing = 4;                // activation slots (one 256-K unit each)
constexpr int kSmStartWindow = 3;          // stages requested before the first one has landed
constexpr int kSmMaxOut = 8;
constexpr int kSmCounterBytes = 64 * 1024; // same workspace header as gemm.cu (zero before, zero after)

// The experiment switches cost ~25 instructions per unit in the consumers' loop: compiled in only on request.
#ifdef QUANTA_SMALL_DBG
#define SM_DBG(p, bit) (((p).dbg & (bit)) != 0)
#else
#define SM_DBG(p, bit) false
#endif

// Timeline of one launch (-DQUANTA_SMALL_TRACE build only): per CTA 40 clock64 stamps relative to kernel entry, + globaltimer at
// entry / exit in slots 38 / 39.  Read back with quanta_debug_small_trace().
#ifdef QUANTA_SMALL_TRACE
__device__ long long g_sm_trace[kNumSMs][40];
#define SM_TRACE(slot) do { if ((slot) < 38) g_sm_trace[blockIdx.x][(slot)] = clock64() - t_entry; } while (0)
#else
#define SM_TRACE(slot) do { } while (0)
#endif

struct SmallParams {
    int M, N, K;
    int S;                  // 256-K steps per row tile
    int n_tiles;
    int G;                  // CTAs
    unsigned int U;         // units = n_tiles * S
    int m_pad;              // 8 * NB
    int R;                  // weight ring stages
    int ldy, col0, n_out;
    int vec_y;              // 8-byte y stores are aligned in every output buffer
    int dbg;                // experiment switches (QUANTA_B200_SMALL_DBG; only in a -DQUANTA_SMALL_DBG build): 1 no compute,
                            // 2 no x staging, 4 no epilogue, 8 no TMA
    uint32_t mul4, mul12;   // 2^28, 2^20: `w >> 4` / `w >> 12` as IMAD.HI on the FMA pipe (see sm_shr)
    uint32_t stage_bytes;   // codes + scale tile + zero-point tile + raw activations [m_pad x 256]
    uint32_t code_bytes;
    uint32_t x_off;         // x ring
    uint32_t x_slot_bytes;  // m_pad * 512 (chunks) + NB * 128 (block sums)
    uint32_t red_off;       // [3][4 row groups][2 slabs][NB][4][32] floats: the K quarters of a tile meet here
    uint32_t bar_off;       // mbarriers + flags
    void* y[kSmMaxOut];
};

template <typename ACT> struct SmTraits;
template <> struct SmTraits<__nv_bfloat16> {
    static constexpr uint32_t kMagic = 0x43004300u;   // bf16x2 (128 + n)
    static constexpr float kOffset = 128.0f;
    __device__ static __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ static __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
    __device__ static __forceinline__ uint32_t pack(float lo, float hi) {
        __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&t);
    }
    __device__ static __forceinline__ void mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
        asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};
template <> struct SmTraits<__half> {
    static constexpr uint32_t kMagic = 0x64006400u;   // fp16x2 (1024 + n)
    static constexpr float kOffset = 1024.0f;
    __device__ static __forceinline__ float to_float(__half v) { return __half2float(v); }
    __device__ static __forceinline__ __half from_float(float v) { return __float2half_rn(v); }
    __device__ static __forceinline__ uint32_t pack(float lo, float hi) {
        __half2 t = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&t);
    }
    __device__ static __forceinline__ void mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
        asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};

__device__ __forceinline__ uint32_t sm_and_or(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t sm_prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r;
}
__device__ __forceinline__ uint2 sm_lds64(uint32_t addr) {
    uint2 v; asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v;
}
__device__ __forceinline__ uint4 sm_lds128(uint32_t addr) {
    uint4 v; asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sm_sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sm_sts32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float sm_lds32(uint32_t addr) {
    float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v;
}
__device__ __forceinline__ void sm_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void sm_bar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sm_bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SMW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SMD_%=;\n\t"
        "bra SMW_%=;\n\t"
        "SMD_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void sm_cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kSmConsThreads) : "memory"); }

// 4-bit: one 32-bit word = nibbles n0..n7 of 8 consecutive K values.  Four LOP3 (+ three shifts)
// give the exact 16-bit pairs (C+n0, C+n4), (C+n1, C+n5), (C+n2, C+n6), (C+n3, C+n7).
template <typename ACT>
__device__ __forceinline__ void sm_pairs4(uint32_t w, uint32_t* p) {
    constexpr uint32_t kM = SmTraits<ACT>::kMagic;
    p[0] = sm_and_or(w, 0x000F000Fu, kM);
    p[1] = sm_and_or(w >> 4, 0x000F000Fu, kM);
    p[2] = sm_and_or(w >> 8, 0x000F000Fu, kM);
    p[3] = sm_and_or(w >> 12, 0x000F000Fu, kM);
}
// 8-bit: one word = 4 codes; PRMT builds the fp32 value 32768 + q, one subtraction leaves q exactly,
// pairs (q0, q1), (q2, q3) in natural K order.
template <typename ACT>
__device__ __forceinline__ void sm_pairs8(uint32_t w, uint32_t* p) {
    float f[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
        f[e] = __fadd_rn(__uint_as_float(sm_prmt(w, 0x47000000u, 0x7504u | (e << 4))), -32768.0f);
    p[0] = SmTraits<ACT>::pack(f[0], f[1]);
    p[1] = SmTraits<ACT>::pack(f[2], f[3]);
}

// Row of the warp's 16-row slab that MMA row `gid` (0..7; +8 for the second half) stands for, chosen so
// that the shared-memory reads of the swizzled code tile are conflict-free: 4-bit reads 8 bytes per
// lane (a half-warp = MMA rows 0..3 must differ in bits 1-2 of the row), 8-bit 16 bytes per lane (a
// quarter-warp = MMA rows 2q, 2q+1 must differ in bit 2).
template <int BITS> __device__ __forceinline__ int sm_row_of(int gid) {
    return BITS == 4 ? (((gid & 3) << 1) | (gid >> 2)) : (((gid & 1) << 2) | (gid >> 1));
}

// CTA whose unit range [U c / G, U (c+1) / G) contains unit u
__device__ __forceinline__ int sm_cta_of_unit(unsigned int u, const SmallParams& p) {
    const unsigned int G = (unsigned int)p.G;
    unsigned int c = (unsigned int)(((unsigned long long)u * G) / p.U);
    while (c + 1 < G && p.U * (c + 1) / G <= u) ++c;
    while (c > 0 && p.U * c / G > u) --c;
    return (int)c;
}

// The segments of a CTA's unit range [u0, u1) — a segment = this CTA's steps [s0, s1) of one row tile — in
// PROCESSING order.  A segment that covers only part of a tile's K range ends with a hand-over through global
// memory, a whole tile ends with a plain store of y.  So when the range holds whole tiles, the partial segments
// (the tail of the first tile, the head of the last one) are processed FIRST and handed over while the stream
// runs on, and the kernel ends on a whole tile.  With no whole tile in the range the natural order is kept.
struct SmWalk {
    unsigned int rb[3], re[3], u, S;
    int nr, r;
    __device__ __forceinline__ void init(unsigned int u0, unsigned int u1, unsigned int S_) {
        S = S_; r = 0; nr = 0;
        const unsigned int t0 = u0 / S, tl = (u1 - 1u) / S;
        const bool f_part = (u0 - t0 * S) != 0u && t0 != tl;            // tail of the first tile
        const bool l_part = (u1 - tl * S) != S && t0 != tl;             // head of the last tile
        const unsigned int wb = f_part ? (t0 + 1u) * S : u0, we = l_part ? tl * S : u1;
        if (t0 != tl && we > wb) {
            if (l_part) { rb[nr] = tl * S; re[nr] = u1; ++nr; }
            if (f_part) { rb[nr] = u0; re[nr] = (t0 + 1u) * S; ++nr; }
            rb[nr] = wb; re[nr] = we; ++nr;
        } else {
            rb[0] = u0; re[0] = u1; nr = 1;
        }
        u = rb[0];
    }
    // next segment: steps [s0, s1) of `tile`; `final`: nothing follows it
    __device__ __forceinline__ bool next(int& tile, int& s0, int& s1, bool& final) {
        if (r < nr && u >= re[r]) { ++r; if (r < nr) u = rb[r]; }
        if (r >= nr) return false;
        const unsigned int t = u / S, b = u - t * S, left = re[r] - u;
        tile = (int)t; s0 = (int)b;
        s1 = (left < S - b) ? (int)(b + left) : (int)S;
        u += (unsigned int)(s1 - s0);
        final = (r == nr - 1) && u >= re[r];
        return true;
    }
};

// Stream-K hand-over of one tile.  The CTA that owns the tile's LAST K steps is its reducer; every other
// contributor stores its partial and bumps the tile's counter with ONE fire-and-forget release
// (red.release.gpu: the partial stores of the CTA — ordered before it by a CTA-level barrier — are visible to
// whoever acquires the incremented value).  The reducer's own share of the tile is the first thing it
// processes, so by the time its stream has ended the others have usually long arrived; it acquires the counter
// (bounded spin; contributors have lower CTA indices and were dispatched earlier), resets it for the next call
// and sums the partials.  A contributor never waits for anybody.
__device__ __forceinline__ void sm_contribute(unsigned int* counters, int tile) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counters + tile) : "memory");
}
__device__ __forceinline__ void sm_await_contributors(unsigned int* counters, int tile, int others) {
    unsigned int seen = 0, spins = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counters + tile) : "memory");
        if (seen >= (unsigned int)others || ++spins > (1u << 22)) break;
    }
    counters[tile] = 0u;                                     // leave the workspace clean for the next call
}

// Sum the partials of every contributor of `tile` in CTA order (deterministic), add the bias, write y.
// Thread j of `nthreads` (a multiple of 32): 4 consecutive features, batch rows j / 32 + (nthreads / 32) k;
// 8 contributors' loads in flight per thread.
template <typename ACT>
__device__ __forceinline__ void sm_fixup(const SmallParams& p, const ACT* __restrict__ bias, const float* __restrict__ partial,
                                         int tile, int c_first, int c_last, int j, int nthreads) {
    using T = SmTraits<ACT>;
    const unsigned int S = (unsigned int)p.S;
    const int n0 = tile * kSmRows;
    const int f4 = 4 * (j & 31), gn4 = n0 + f4;
    const size_t slot = (size_t)(kSmRows * p.m_pad);
    float b4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) b4[e] = (bias != nullptr && gn4 + e < p.N) ? T::to_float(bias[gn4 + e]) : 0.0f;
    for (int m = j >> 5; m < p.M; m += nthreads >> 5) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c0 = c_first; c0 <= c_last; c0 += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = c0 + u;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c <= c_last) {
                    const int wc = (tile == (int)((p.U * (unsigned int)c / (unsigned int)p.G) / S)) ? 0 : 1;
                    v[u] = __ldcg(reinterpret_cast<const float4*>(partial + ((size_t)c * 2 + wc) * slot + m * kSmRows + f4));
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        const float o4[4] = {acc.x + b4[0], acc.y + b4[1], acc.z + b4[2], acc.w + b4[3]};
        for (int o = 0; o < p.n_out; ++o) {
            ACT* dst = static_cast<ACT*>(p.y[o]) + (int64_t)m * p.ldy + p.col0 + gn4;
            if (p.vec_y && gn4 + 3 < p.N) {
                uint2 pk;
                pk.x = T::pack(o4[0], o4[1]);
                pk.y = T::pack(o4[2], o4[3]);
                *reinterpret_cast<uint2*>(dst) = pk;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (gn4 + e < p.N) dst[e] = T::from_float(o4[e]);
            }
        }
    }
}

template <typename ACT, int BITS, int NB>
__global__ void __launch_bounds__(kSmThreads, 1)
gemm_small_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_s,
                  const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_x, const ACT* __restrict__ bias,
                  unsigned int* __restrict__ counters, float* __restrict__ partial, const __grid_constant__ SmallParams p) {
    using T = SmTraits<ACT>;
    extern __shared__ __align__(1024) uint8_t sm_raw[];

    const uint32_t smem = smem_u32(sm_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t bars = smem + p.bar_off;                  // full[R] | empty[R] | xfull[XR] | xempty[XR]
    auto full_bar = [&](int s) { return bars + (uint32_t)s * 8u; };
    auto empty_bar = [&](int s) { return bars + (uint32_t)(kSmMaxRing + s) * 8u; };
    auto xfull_bar = [&](int s) { return bars + (uint32_t)(2 * kSmMaxRing + s) * 8u; };
    auto xempty_bar = [&](int s) { return bars + (uint32_t)(2 * kSmMaxRing + kSmXRing + s) * 8u; };

#ifdef QUANTA_SMALL_TRACE
    const long long t_entry = clock64();
    if (tid == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); g_sm_trace[blockIdx.x][38] = (long long)g; }
#endif
    const unsigned int cta = blockIdx.x;
    const unsigned int u0 = p.U * cta / (unsigned int)p.G, u1 = p.U * (cta + 1u) / (unsigned int)p.G;
    const int n_units = (int)(u1 - u0);
    const int S = p.S;
    const int tile0 = (int)(u0 / (unsigned int)S);       // the first tile of the range owns partial slot 0
    SmWalk walk;
    walk.init(u0, u1, (unsigned int)S);
    int tile, s0, s1;
    bool final_seg;

    // The producer initialises the weight ring itself, checks in at the CTA barrier WITHOUT waiting (bar.arrive)
    // and starts streaming at once; everybody else sees all barriers after bar.sync.
    if (tid == 32 * kSmConsWarps) {
        SM_TRACE(32);
        prefetch_tensormap(&tmap_w); prefetch_tensormap(&tmap_s); prefetch_tensormap(&tmap_z); prefetch_tensormap(&tmap_x);
        for (int s = 0; s < p.R; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_bar(s)), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty_bar(s)), "r"(kSmConsWarps + kSmXWarps));
        }
        fence_barrier_init();
        SM_TRACE(33);
    }
    if (tid == 0) {
        for (int s = 0; s < kSmXRing; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(xfull_bar(s)), "r"(kSmXWarps));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(xempty_bar(s)), "r"(kSmConsWarps));
        }
        fence_barrier_init();
    }
    if (warp == kSmConsWarps) {
        __syncwarp();
        asm volatile("barrier.arrive 0, %0;" ::"n"(kSmThreads) : "memory");
    } else {
        asm volatile("barrier.sync 0, %0;" ::"n"(kSmThreads) : "memory");
    }

    if (warp == kSmConsWarps) {
        // ===== producer: keeps the ring full across tile boundaries =====
        if (lane == 0) {
            SM_TRACE(34);
            const uint64_t pol = policy_evict_first();       // weights are streamed once
            const uint64_t pol_x = policy_evict_last();      // activations are re-read by every row tile
            int slot = 0, i = 0;
            uint32_t ph = 0;
            // Programmatic dependent launch: this grid may start while the previous kernel on the stream is still in
            // its tail.  The weights do not depend on it, so the first stages' codes / scales / zero-points are
            // requested at once; the activations (the previous kernel's output, in a chain of layers) and every
            // global write of this kernel wait for griddepcontrol.wait — a no-op in an ordinary launch.
            int pend_step[kSmStartWindow], npend = 0;
            bool waited = false;
            auto issue_weights = [&](uint32_t bar, uint32_t dst, int kb, int n0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(p.stage_bytes) : "memory");
                if (BITS == 4) {
                    sm_tma_2d(dst, &tmap_w, bar, kb / 2, n0, pol);
                } else {
                    sm_tma_2d(dst, &tmap_w, bar, kb, n0, pol);
                    sm_tma_2d(dst + 16384u, &tmap_w, bar, kb + 128, n0, pol);
                }
                sm_tma_2d(dst + p.code_bytes, &tmap_s, bar, kb / 64, n0, pol);
                sm_tma_2d(dst + p.code_bytes + 2048u, &tmap_z, bar, kb / 64, n0, pol);
            };
            auto dependency_wait = [&]() {
                asm volatile("griddepcontrol.wait;" ::: "memory");
                waited = true;
                // the activations [m_pad rows x 256 K] of the stages requested so far (rows past M arrive zero-filled)
                for (int k = 0; k < npend; ++k)
                    sm_tma_2d(smem + (uint32_t)k * p.stage_bytes + p.code_bytes + 4096u, &tmap_x, full_bar(k), pend_step[k] * kSmStepK, 0, pol_x);
            };
            while (walk.next(tile, s0, s1, final_seg)) {
                for (int step = s0; step < s1; ++step, ++i) {
                    if (SM_DBG(p, 8)) break;
                    if (i == 0) SM_TRACE(1);
                    if (i == kSmStartWindow) {
                        // start-up window (see the header): open the whole ring once the first stage has landed
                        dependency_wait();
                        SM_TRACE(2);
                        sm_bar_wait(full_bar(0), 0u);
                        SM_TRACE(3);
                    }
                    if (i >= p.R) sm_bar_wait(empty_bar(slot), ph ^ 1u);
                    const uint32_t bar = full_bar(slot);
                    const uint32_t dst = smem + (uint32_t)slot * p.stage_bytes;
                    issue_weights(bar, dst, step * kSmStepK, tile * kSmRows);
                    if (waited) sm_tma_2d(dst + p.code_bytes + 4096u, &tmap_x, bar, step * kSmStepK, 0, pol_x);
                    else pend_step[npend++] = step;
                    if (++slot == p.R) { slot = 0; ph ^= 1u; }
                }
            }
            if (!waited) dependency_wait();
            SM_TRACE(4);
        }
        return;
    }
    // Let the next kernel on the stream start its own prologue as soon as this grid's CTAs make room (its weights do
    // not depend on us); every other warp orders its global accesses behind the previous kernel.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (warp == kSmConsWarps + 1) {
        // ===== publisher: releases the partial of every contributor segment that is not the CTA's last one while
        //       the stream runs on (a segment that reaches the tile's last step makes this CTA the reducer: no release) =====
        while (walk.next(tile, s0, s1, final_seg)) {
            if (final_seg || s1 == S || SM_DBG(p, 4)) continue;
            asm volatile("bar.sync 2, 160;" ::: "memory");   // the 4 output warps have stored the partial
            if (lane == 0) sm_contribute(counters, tile);
        }
        return;
    }

    if (warp >= kSmConsWarps + 2) {
        // ===== activation staging warps =====
        // The unit's activations arrive with its weights (raw [m_pad x 256] rows behind the zero-point tile).  Chunk c
        // (c = 32 xw + lane + 128 it): batch row m = c / 32, block (c / 8) & 3, 8 consecutive K values 8 (c & 7) of the
        // block -> slot chunk (((nb 4 + blk) 2 + j) 32 + gid 4 + tig) with nb = m / 8, gid = m & 7, tig = (c & 7) / 2,
        // j = c & 1: a consumer warp's B-fragment read of one (nb, blk, j) is 512 contiguous bytes.
        constexpr int kChunks = 8 * NB * 32;                 // 16-byte x chunks per unit
        constexpr int kPer = kChunks / (32 * kSmXWarps);
        const int xw = warp - (kSmConsWarps + 2);
        uint32_t xsrc[kPer], xdst[kPer], xsum_dst[kPer];
#pragma unroll
        for (int it = 0; it < kPer; ++it) {
            const int c = 32 * xw + lane + 32 * kSmXWarps * it;
            const int m = c >> 5, blk = (c >> 3) & 3, c8 = c & 7;
            const int nb = m >> 3, g = m & 7, t = c8 >> 1, j = c8 & 1;
            xsrc[it] = p.code_bytes + 4096u + (uint32_t)(c * 16);
            xdst[it] = (uint32_t)((((nb * 4 + blk) * 2 + j) * 32 + g * 4 + t) * 16);
            xsum_dst[it] = (uint32_t)(NB * 4096 + ((nb * 4 + blk) * 8 + g) * 4);
        }
        const bool x_on = !SM_DBG(p, 2), w_on = !SM_DBG(p, 8);
        int xs = 0, ws_ = 0;
        uint32_t ph = 0, wph_ = 0;
        for (int i = 0; i < n_units; ++i) {
            if (w_on) sm_bar_wait(full_bar(ws_), wph_);
            if (!x_on) {                                     // experiment switch: release the stage, stage nothing
                __syncwarp();
                if (lane == 0) sm_bar_arrive(empty_bar(ws_));
                if (++ws_ == p.R) { ws_ = 0; wph_ ^= 1u; }
                continue;
            }
            if (i >= kSmXRing) sm_bar_wait(xempty_bar(xs), ph ^ 1u);
            const uint32_t stage = smem + (uint32_t)ws_ * p.stage_bytes;
            const uint32_t slot = smem + p.x_off + (uint32_t)xs * p.x_slot_bytes;
            uint4 raw[kPer];
#pragma unroll
            for (int it = 0; it < kPer; ++it) raw[it] = sm_lds128(stage + xsrc[it]);
#pragma unroll
            for (int it = 0; it < kPer; ++it) {
                uint4 v = raw[it];
                const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
                float sum = 0.0f;
                if (SmTraits<ACT>::kOffset == 128.0f) {
                    // bf16 -> fp32 is a shift / mask
#pragma unroll
                    for (int q = 0; q < 4; ++q) sum += __uint_as_float(w4[q] << 16) + __uint_as_float(w4[q] & 0xFFFF0000u);
                } else {
                    const ACT* e = reinterpret_cast<const ACT*>(&v);
#pragma unroll
                    for (int q = 0; q < 8; ++q) sum += T::to_float(e[q]);
                }
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                sum += __shfl_xor_sync(0xffffffffu, sum, 4);
                if (BITS == 4) {
                    uint4 o;
                    o.x = sm_prmt(v.x, v.z, 0x5410u);        // (x0, x4)
                    o.y = sm_prmt(v.x, v.z, 0x7632u);        // (x1, x5)
                    o.z = sm_prmt(v.y, v.w, 0x5410u);        // (x2, x6)
                    o.w = sm_prmt(v.y, v.w, 0x7632u);        // (x3, x7)
                    v = o;
                }
                sm_sts128(slot + xdst[it], v);
                if ((lane & 7) == 0) sm_sts32(slot + xsum_dst[it], sum);
            }
            __syncwarp();
            if (lane == 0) { sm_bar_arrive(xfull_bar(xs)); sm_bar_arrive(empty_bar(ws_)); }
            if (++xs == kSmXRing) { xs = 0; ph ^= 1u; }
            if (++ws_ == p.R) { ws_ = 0; wph_ ^= 1u; }
        }
        return;
    }

    // ===== consumers =====
    const int gid = lane >> 2, tig = lane & 3;
    const int rg = warp & 3, kq = warp >> 2;                 // rows 32 rg .. +32, block kq
    // rows[sl][h]: slab sl (16 rows), MMA row gid + 8 h
    int rows[2][2];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) { rows[sl][0] = (2 * rg + sl) * 16 + sm_row_of<BITS>(gid); rows[sl][1] = rows[sl][0] + 8; }

    const bool do_compute = !SM_DBG(p, 1), do_x = !SM_DBG(p, 2), do_epi = !SM_DBG(p, 4), do_wait = !SM_DBG(p, 8);
    // Per-thread shared-memory offsets (the block b = kq is fixed per warp; rows + 8 / + 16 keep the swizzle
    // phase, so every other address of a unit is one of these plus an immediate).
    const int r00 = rows[0][0];
    const uint32_t w_off = BITS == 4
        ? (uint32_t)r00 * 128u + ((((uint32_t)(2 * kq + (tig >> 1))) ^ (uint32_t)(r00 & 7)) << 4) + (uint32_t)(tig & 1) * 8u
        : (uint32_t)(kq >> 1) * 16384u + (uint32_t)r00 * 128u + ((((uint32_t)(4 * (kq & 1) + tig)) ^ (uint32_t)(r00 & 7)) << 4);
    const uint32_t s_off = p.code_bytes + (uint32_t)r00 * 16u + (uint32_t)kq * 4u;
    const uint32_t xb_off = p.x_off + (uint32_t)((kq * 2) * 512 + lane * 16);
    const uint32_t xs_off = p.x_off + (uint32_t)(NB * 4096 + (kq * 8 + 2 * tig) * 4);
    // Absolute shared-memory addresses of this thread's words in slot 0.  Made opaque so that they live in
    // registers: ptxas otherwise rematerialises the whole lane / warp arithmetic above inside the unit loop
    // (~30 of its ~160 instructions).
    uint32_t w_base = smem + w_off, s_base = smem + s_off, xb_base = smem + xb_off, xs_base0 = smem + xs_off;
    asm volatile("" : "+r"(w_base), "+r"(s_base), "+r"(xb_base), "+r"(xs_base0));
    const uint32_t stage_bytes = p.stage_bytes, x_slot_bytes = p.x_slot_bytes;

    float tot[2][NB][4];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) { tot[sl][nb][0] = tot[sl][nb][1] = tot[sl][nb][2] = tot[sl][nb][3] = 0.0f; }
    int wslot = 0, xs_cur = 0, red_tile = -1;
    uint32_t wph = 0, xph = 0;

#ifdef QUANTA_SMALL_TRACE
    int tr_unit = 0;
#endif
    if (tid == 0) SM_TRACE(5);
    while (walk.next(tile, s0, s1, final_seg)) {
        for (int step = s0; step < s1; ++step) {
            // The staging warps signal xfull only after they have seen the stage's `full` barrier complete (all of
            // its TMA bytes are in shared memory by then), so one wait covers weights and activations.
            if (do_x) sm_bar_wait(xfull_bar(xs_cur), xph);
            else if (do_wait) sm_bar_wait(full_bar(wslot), wph);

            if (do_compute) {
                const uint32_t w_cur = w_base + (uint32_t)wslot * stage_bytes, s_cur = s_base + (uint32_t)wslot * stage_bytes;
                const uint32_t xb_cur = xb_base + (uint32_t)xs_cur * x_slot_bytes, xs_addr = xs_base0 + (uint32_t)xs_cur * x_slot_bytes;
                // ---- every shared-memory read of the unit goes out first (they are ordered asm statements;
                //      the arithmetic below is free for the compiler to interleave) ----
                uint32_t wraw[2][2][BITS == 4 ? 2 : 4];      // [slab][row half][words]
#pragma unroll
                for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t wa = w_cur + (uint32_t)(sl * 2048 + h * 1024);
                        if (BITS == 4) {
                            const uint2 w = sm_lds64(wa);
                            wraw[sl][h][0] = w.x; wraw[sl][h][1] = w.y;
                        } else {
                            const uint4 w = sm_lds128(wa);
                            wraw[sl][h][0] = w.x; wraw[sl][h][1] = w.y; wraw[sl][h][BITS == 4 ? 0 : 2] = w.z; wraw[sl][h][BITS == 4 ? 1 : 3] = w.w;
                        }
                    }
                uint4 xb[NB][2];
                uint2 xs2[NB];
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    xb[nb][0] = sm_lds128(xb_cur + (uint32_t)(nb * 4096));
                    xb[nb][1] = sm_lds128(xb_cur + (uint32_t)(nb * 4096 + 512));
                    xs2[nb] = sm_lds64(xs_addr + (uint32_t)(nb * 128));
                }
                float sc[2][2], zc[2][2];
                const float off = BITS == 4 ? T::kOffset : 0.0f;
#pragma unroll
                for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        sc[sl][h] = sm_lds32(s_cur + (uint32_t)(sl * 256 + h * 128));
                        zc[sl][h] = sm_lds32(s_cur + (uint32_t)(sl * 256 + h * 128 + 2048));
                    }
                // ---- 4 MMAs (K = 16 each) per slab and batch block: k outermost, so that consecutive MMAs
                //      are independent (2 slabs x NB accumulators) ----
                float c[2][NB][4];
#pragma unroll
                for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) { c[sl][nb][0] = c[sl][nb][1] = c[sl][nb][2] = c[sl][nb][3] = 0.0f; }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t a[2][4];                        // {row lo, row+8 lo, row hi, row+8 hi}
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl) {
                        if (BITS == 4) {
                            // word k / 2: pairs (n0,n4) (n1,n5) for even k, (n2,n6) (n3,n7) for odd k
                            // the shifts by 4 and 12 are multiplications (IMAD.HI by 2^28 / 2^20 from the parameter bank,
                            // opaque to ptxas): they issue on the FMA pipe while the LOP3s keep the ALU pipe busy
                            const uint32_t r0 = wraw[sl][0][k >> 1], r1 = wraw[sl][1][k >> 1];
                            const uint32_t w0 = (k & 1) ? (r0 >> 8) : r0, w1 = (k & 1) ? (r1 >> 8) : r1;
                            const uint32_t m = (k & 1) ? p.mul12 : p.mul4;
                            a[sl][0] = sm_and_or(w0, 0x000F000Fu, T::kMagic);
                            a[sl][1] = sm_and_or(w1, 0x000F000Fu, T::kMagic);
                            a[sl][2] = sm_and_or(__umulhi(r0, m), 0x000F000Fu, T::kMagic);
                            a[sl][3] = sm_and_or(__umulhi(r1, m), 0x000F000Fu, T::kMagic);
                        } else {
                            uint32_t p0[2], p1[2];
                            sm_pairs8<ACT>(wraw[sl][0][k], p0); sm_pairs8<ACT>(wraw[sl][1][k], p1);
                            a[sl][0] = p0[0]; a[sl][1] = p1[0]; a[sl][2] = p0[1]; a[sl][3] = p1[1];
                        }
                    }
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
                        const uint4 xv4 = xb[nb][k >> 1];
                        const uint32_t b0 = (k & 1) ? xv4.z : xv4.x, b1 = (k & 1) ? xv4.w : xv4.y;
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl) T::mma(c[sl][nb], a[sl], b0, b1);
                    }
                }
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    const float z0 = __fmaf_rn(-off, sc[sl][0], zc[sl][0]), z1 = __fmaf_rn(-off, sc[sl][1], zc[sl][1]);
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
                        const float xs_a = __uint_as_float(xs2[nb].x), xs_b = __uint_as_float(xs2[nb].y);
                        tot[sl][nb][0] = __fmaf_rn(sc[sl][0], c[sl][nb][0], __fmaf_rn(z0, xs_a, tot[sl][nb][0]));
                        tot[sl][nb][1] = __fmaf_rn(sc[sl][0], c[sl][nb][1], __fmaf_rn(z0, xs_b, tot[sl][nb][1]));
                        tot[sl][nb][2] = __fmaf_rn(sc[sl][1], c[sl][nb][2], __fmaf_rn(z1, xs_a, tot[sl][nb][2]));
                        tot[sl][nb][3] = __fmaf_rn(sc[sl][1], c[sl][nb][3], __fmaf_rn(z1, xs_b, tot[sl][nb][3]));
                    }
                }
            }
            // every lane's reads of the stage have been consumed by the instructions above
            __syncwarp();
            if (lane == 0) { sm_bar_arrive(empty_bar(wslot)); if (do_x) sm_bar_arrive(xempty_bar(xs_cur)); }
            if (++wslot == p.R) { wslot = 0; wph ^= 1u; }
            if (++xs_cur == kSmXRing) { xs_cur = 0; xph ^= 1u; }
#ifdef QUANTA_SMALL_TRACE
            if (tid == 0 && tr_unit < 24) { SM_TRACE(8 + tr_unit); ++tr_unit; }
#endif
        }
        if (tid == 0) SM_TRACE(6);
        if (do_epi) {
            // ===== end of this CTA's segment [s0, s1) of `tile`: the 4 K quarters meet in shared memory =====
            const bool whole = s0 == 0 && s1 == S;
            const uint32_t red = smem + p.red_off + (uint32_t)(rg * (2 * NB * 4) * 128 + lane * 4);
            const uint32_t red_group = (uint32_t)(4 * 2 * NB * 4 * 128);
            if (kq > 0) {
#pragma unroll
                for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            sm_sts32(red + (uint32_t)(kq - 1) * red_group + (uint32_t)(((sl * NB + nb) * 4 + q) * 128), tot[sl][nb][q]);
            }
            sm_cons_sync();
            const int n0 = tile * kSmRows;
            if (kq == 0) {
                // tot[sl][nb][q]: feature n0 + rows[sl][q >> 1], batch row 8 nb + 2 tig + (q & 1)
#pragma unroll
                for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t ra = red + (uint32_t)(((sl * NB + nb) * 4 + q) * 128);
                            tot[sl][nb][q] = ((tot[sl][nb][q] + sm_lds32(ra)) + sm_lds32(ra + red_group)) + sm_lds32(ra + 2 * red_group);
                        }
                if (whole) {
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int gn = n0 + rows[sl][h];
                            const float bv = (bias != nullptr && gn < p.N) ? T::to_float(bias[gn]) : 0.0f;
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int m = 8 * nb + 2 * tig + e;
                                    if (m < p.M && gn < p.N) {
                                        const ACT v = T::from_float(tot[sl][nb][2 * h + e] + bv);
                                        for (int o = 0; o < p.n_out; ++o)
                                            static_cast<ACT*>(p.y[o])[(int64_t)m * p.ldy + p.col0 + gn] = v;
                                    }
                                }
                        }
                    }
                } else {
                    // this CTA's partial slot for the tile: 0 if it is the first tile the CTA touches, else 1;
                    // layout [m][128 features]
                    const int which = (tile == tile0) ? 0 : 1;
                    float* mine = partial + ((size_t)cta * 2 + which) * (size_t)(kSmRows * p.m_pad);
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                __stcg(mine + (8 * nb + 2 * tig + (q & 1)) * kSmRows + rows[sl][q >> 1], tot[sl][nb][q]);
                    if (!final_seg && s1 != S) {
                        // a contributor segment with more work behind it: the publisher warp releases it
                        __syncwarp();
                        asm volatile("bar.arrive 2, 160;" ::: "memory");
                    }
                }
            }
#pragma unroll
            for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) { tot[sl][nb][0] = tot[sl][nb][1] = tot[sl][nb][2] = tot[sl][nb][3] = 0.0f; }
            if (!whole && s1 == S) red_tile = tile;          // this CTA reduces the tile once its stream has ended
            sm_cons_sync();                                  // red is reused by the next segment; the partial stores are done
            if (!whole && final_seg && s1 != S && tid == 0) sm_contribute(counters, tile);
        }
    }
    if (tid == 0) SM_TRACE(7);
    if (red_tile >= 0 && do_epi) {
        // ===== reducer duty: the other contributors' partials (lower CTA indices) =====
        const unsigned int tu0 = (unsigned int)red_tile * (unsigned int)S;
        const int c_first = sm_cta_of_unit(tu0, p), c_last = (int)cta;
        if (tid == 0) sm_await_contributors(counters, red_tile, c_last - c_first);
        sm_cons_sync();
        if (tid == 0) SM_TRACE(36);
        sm_fixup<ACT>(p, bias, partial, red_tile, c_first, c_last, tid, kSmConsThreads);
    }
#ifdef QUANTA_SMALL_TRACE
    if (tid == 0) {
        SM_TRACE(37);
        unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); g_sm_trace[blockIdx.x][39] = (long long)g;
    }
#endif
}

#ifdef QUANTA_SMALL_TRACE
extern "C" __attribute__((visibility("default"))) int quanta_debug_small_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_sm_trace, sizeof(long long) * kNumSMs * 40);
}
#endif

// ---- host side --------------------------------------------------------------

struct SmallTuning { int ring; int max_m; int ctas; int dbg; int pdl; };
static SmallTuning small_tuning() {
    static const SmallTuning t = []() {
        SmallTuning v{0, 16, 0, 0, 1};
        if (const char* e = getenv("QUANTA_B200_SMALL_RING")) { int s = atoi(e); if (s >= 2 && s <= kSmMaxRing) v.ring = s; }
        if (const char* e = getenv("QUANTA_B200_SMALL_MAX_M")) { int m = atoi(e); if (m >= 0 && m <= 16) v.max_m = m; }
        if (const char* e = getenv("QUANTA_B200_SMALL_CTAS")) { int c = atoi(e); if (c >= 1 && c <= kNumSMs) v.ctas = c; }
        if (const char* e = getenv("QUANTA_B200_SMALL_DBG")) v.dbg = atoi(e);
        if (const char* e = getenv("QUANTA_B200_SMALL_PDL")) v.pdl = atoi(e) != 0;
        return v;
    }();
    return t;
}

// The small-batch kernel serves M <= 16 with blockwise-64 parameters on K % 256 == 0.
bool gemm_small_eligible(int64_t M, int64_t N, int64_t K, int64_t block, const void* scale, const void* zp) {
    if (M > small_tuning().max_m || block != 64 || (K % kSmStepK) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(zp)) & 15) return false;
    const int64_t n_tiles = (N + kSmRows - 1) / kSmRows;
    if (n_tiles > kSmCounterBytes / 4) return false;
    if ((unsigned long long)n_tiles * (unsigned long long)(K / kSmStepK) * (unsigned long long)(kNumSMs + 1) >= (1ull << 32)) return false;
    return true;
}

template <typename ACT, int BITS, int NB>
static int gemm_small_launch_nb(const ACT* x, const uint8_t* wq, const float* scale, const float* zp, const ACT* bias,
                                void* const* ys, int n_out, int64_t ldy, int64_t col0, int64_t M, int64_t N, int64_t K,
                                void* workspace, size_t ws_bytes, cudaStream_t st) {
    SmallParams p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.m_pad = 8 * NB;
    p.n_tiles = (int)((N + kSmRows - 1) / kSmRows);
    p.S = (int)(K / kSmStepK);
    p.U = (unsigned int)p.n_tiles * (unsigned int)p.S;
    int dev = 0;
    cudaGetDevice(&dev);
    int sms = kNumSMs;
    {
        static int cached[64] = {0};
        if (dev >= 0 && dev < 64) {
            if (cached[dev] == 0) {
                int n = 0;
                if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
                cached[dev] = n < kNumSMs ? n : kNumSMs;     // partial slots are sized for kNumSMs
            }
            sms = cached[dev];
        }
    }
    if (small_tuning().ctas) sms = small_tuning().ctas < sms ? small_tuning().ctas : sms;
    p.G = (int)(p.U < (unsigned int)sms ? p.U : (unsigned int)sms);
    p.code_bytes = BITS == 4 ? 16384u : 32768u;
    p.stage_bytes = p.code_bytes + 4096u + (uint32_t)p.m_pad * 512u;       // codes | scales | zero-points | activations
    p.x_slot_bytes = (uint32_t)(NB * 4096 + NB * 128);
    const uint32_t x_bytes = (uint32_t)kSmXRing * p.x_slot_bytes;
    const uint32_t red_bytes = (uint32_t)(3 * 4 * 2 * NB * 4 * 128);
    const uint32_t bar_bytes = 256;
    const uint32_t budget = 226u * 1024u;
    int ring = (int)((budget - x_bytes - red_bytes - bar_bytes) / p.stage_bytes);
    if (ring > kSmMaxRing) ring = kSmMaxRing;
    if (small_tuning().ring && small_tuning().ring < ring) ring = small_tuning().ring;
    if (ring < 2) return QUANTA_EUNSUPPORTED;
    p.R = ring;
    p.x_off = (uint32_t)ring * p.stage_bytes;
    p.red_off = p.x_off + x_bytes;
    p.bar_off = p.red_off + red_bytes;
    const int smem = (int)(p.bar_off + bar_bytes);
    p.ldy = (int)ldy; p.col0 = (int)col0; p.n_out = n_out;
    p.dbg = small_tuning().dbg;
    p.mul4 = 1u << 28;
    p.mul12 = 1u << 20;
    bool aligned8 = true;
    for (int o = 0; o < kSmMaxOut; ++o) {
        p.y[o] = o < n_out ? ys[o] : nullptr;
        if (o < n_out) aligned8 = aligned8 && (reinterpret_cast<uintptr_t>(ys[o]) & 7) == 0;
    }
    p.vec_y = (aligned8 && (ldy & 3) == 0 && (col0 & 3) == 0) ? 1 : 0;

    const size_t need = (size_t)kSmCounterBytes + (size_t)p.G * 2 * kSmRows * (size_t)p.m_pad * sizeof(float) + 256;
    if (!workspace || ws_bytes < need) return QUANTA_EWORKSPACE;
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
    float* partial = reinterpret_cast<float*>(ws + kSmCounterBytes);

    CUtensorMap tmap_w, tmap_s, tmap_z;
    const uint64_t wrow_bytes = (uint64_t)K * BITS / 8;
    int rc = make_tensor_map_2d(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, wq, wrow_bytes, (uint64_t)N, wrow_bytes, 128,
                                kSmRows, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const uint64_t sstride = (uint64_t)(K / 64);
    rc = make_tensor_map_2d(&tmap_s, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, scale, sstride, (uint64_t)N, sstride * 4, 4, kSmRows,
                            CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    rc = make_tensor_map_2d(&tmap_z, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, zp, sstride, (uint64_t)N, sstride * 4, 4, kSmRows,
                            CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;

    CUtensorMap tmap_x;
    rc = make_tensor_map_2d(&tmap_x, SmTraits<ACT>::kOffset == 128.0f ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                            2, x, (uint64_t)K, (uint64_t)M, (uint64_t)K * 2, kSmStepK, (uint32_t)p.m_pad, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    auto kern = gemm_small_kernel<ACT, BITS, NB>;
    if (int e = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem)) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.G);
    cfg.blockDim = dim3(kSmThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = small_tuning().pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmap_w, tmap_s, tmap_z, tmap_x, bias, counters, partial, p);
    return cuda_status(e != cudaSuccess ? e : cudaGetLastError());
}

template <typename ACT, int BITS>
int gemm_small_launch(const ACT* x, const uint8_t* wq, const float* scale, const float* zp, const ACT* bias, void* const* ys,
                      int n_out, int64_t ldy, int64_t col0, int64_t M, int64_t N, int64_t K, void* workspace, size_t ws_bytes,
                      cudaStream_t st) {
    if (M <= 8) return gemm_small_launch_nb<ACT, BITS, 1>(x, wq, scale, zp, bias, ys, n_out, ldy, col0, M, N, K, workspace, ws_bytes, st);
    return gemm_small_launch_nb<ACT, BITS, 2>(x, wq, scale, zp, bias, ys, n_out, ldy, col0, M, N, K, workspace, ws_bytes, st);
}

#define QUANTA_SMALL_INST(ACT, BITS)                                                                                          \
    template int gemm_small_launch<ACT, BITS>(const ACT*, const uint8_t*, const float*, const float*, const ACT*, void* const*, \
                                              int, int64_t, int64_t, int64_t, int64_t, int64_t, void*, size_t, cudaStream_t);
QUANTA_SMALL_INST(__nv_bfloat16, 4)
QUANTA_SMALL_INST(__nv_bfloat16, 8)
QUANTA_SMALL_INST(__half, 4)
QUANTA_SMALL_INST(__half, 8)

}  // namespace quanta
