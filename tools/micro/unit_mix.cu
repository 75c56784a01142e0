// Microbenchmark: the instruction mix of one gemm_small consumer unit (per warp: 4 x LDS.64 codes, 2 x LDS.128 + LDS.64
// activations, 8 scale loads, 32 LOP3 + 24 shifts, 8 HMMA.16816 in two dependent chains of 4, 20 FFMA) run by 16 warps
// per SM out of shared memory, with parts switched off, to see which pipe bounds it on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o unit_mix unit_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b, unsigned c) { unsigned r; asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ void mma(float* c, const unsigned* a, unsigned b0, unsigned b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint2 lds64(const void* p) { uint2 v; unsigned a = (unsigned)__cvta_generic_to_shared(p); asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds128(const void* p) { uint4 v; unsigned a = (unsigned)__cvta_generic_to_shared(p); asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ float lds32(const void* p) { float v; unsigned a = (unsigned)__cvta_generic_to_shared(p); asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
// MODE bits: 1 unpack, 2 mma, 4 ffma, 8 shifts via IMAD.HI, 16: 4 independent chains (k steps do not depend)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, unsigned mul4, unsigned mul12) {
    extern __shared__ unsigned sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 16384; i += 512) sm[i] = i * 2654435761u;
    __syncthreads();
    float tot[2][4] = {};
    const unsigned* wp = sm + warp * 256 + lane * 2;
    const unsigned* xp = sm + 8192 + lane * 4;
    const float* sp = reinterpret_cast<const float*>(sm) + 12288 + (lane >> 2) * 4;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const unsigned o = (it & 7) * 1024;
        uint2 w[2][2];
        for (int sl = 0; sl < 2; ++sl) for (int h = 0; h < 2; ++h) w[sl][h] = lds64(wp + ((o + sl * 128 + h * 64) & 8191));
        uint4 xb0 = lds128(xp + (o & 2047)), xb1 = lds128(xp + ((o + 128) & 2047));
        uint2 xsu = lds64(sp + 64 + (o & 255)); float2 xs = make_float2(__uint_as_float(xsu.x), __uint_as_float(xsu.y));
        float sc[2][2], zc[2][2];
        for (int sl = 0; sl < 2; ++sl) for (int h = 0; h < 2; ++h) { sc[sl][h] = lds32(sp + sl * 32 + h * 16 + (o & 255)); zc[sl][h] = lds32(sp + 512 + sl * 32 + h * 16 + (o & 255)); }
        float c[2][4] = {};
        float c2[2][4][4] = {};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            unsigned a[2][4];
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                const unsigned r0 = (kk >> 1) ? w[sl][0].y : w[sl][0].x, r1 = (kk >> 1) ? w[sl][1].y : w[sl][1].x;
                if (MODE & 1) {
                    const unsigned w0 = (kk & 1) ? (r0 >> 8) : r0, w1 = (kk & 1) ? (r1 >> 8) : r1;
                    unsigned h0, h1;
                    if (MODE & 8) { const unsigned m = (kk & 1) ? mul12 : mul4; h0 = __umulhi(r0, m); h1 = __umulhi(r1, m); }
                    else { h0 = r0 >> ((kk & 1) ? 12 : 4); h1 = r1 >> ((kk & 1) ? 12 : 4); }
                    a[sl][0] = lop(w0, 0x000F000Fu, 0x43004300u); a[sl][1] = lop(w1, 0x000F000Fu, 0x43004300u);
                    a[sl][2] = lop(h0, 0x000F000Fu, 0x43004300u); a[sl][3] = lop(h1, 0x000F000Fu, 0x43004300u);
                } else { a[sl][0] = r0; a[sl][1] = r1; a[sl][2] = r0 ^ kk; a[sl][3] = r1 ^ kk; }
            }
            const unsigned b0 = (kk & 1) ? ((kk >> 1) ? xb1.z : xb0.z) : ((kk >> 1) ? xb1.x : xb0.x);
            const unsigned b1 = (kk & 1) ? ((kk >> 1) ? xb1.w : xb0.w) : ((kk >> 1) ? xb1.y : xb0.y);
            if (MODE & 2) {
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) { if (MODE & 16) mma(c2[sl][kk], a[sl], b0, b1); else mma(c[sl], a[sl], b0, b1); }
            } else {
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) { c[sl][0] += __uint_as_float(a[sl][0] ^ b0); c[sl][1] += __uint_as_float(a[sl][1] ^ b1); c[sl][2] += __uint_as_float(a[sl][2]); c[sl][3] += __uint_as_float(a[sl][3]); }
            }
        }
        if (MODE & 16) for (int sl = 0; sl < 2; ++sl) for (int q = 0; q < 4; ++q) c[sl][q] = (c2[sl][0][q] + c2[sl][1][q]) + (c2[sl][2][q] + c2[sl][3][q]);
        if (MODE & 4) {
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                const float z0 = fmaf(-128.f, sc[sl][0], zc[sl][0]), z1 = fmaf(-128.f, sc[sl][1], zc[sl][1]);
                tot[sl][0] = fmaf(sc[sl][0], c[sl][0], fmaf(z0, xs.x, tot[sl][0])); tot[sl][1] = fmaf(sc[sl][0], c[sl][1], fmaf(z0, xs.y, tot[sl][1]));
                tot[sl][2] = fmaf(sc[sl][1], c[sl][2], fmaf(z1, xs.x, tot[sl][2])); tot[sl][3] = fmaf(sc[sl][1], c[sl][3], fmaf(z1, xs.y, tot[sl][3]));
            }
        } else { for (int sl = 0; sl < 2; ++sl) for (int q = 0; q < 4; ++q) tot[sl][q] += c[sl][q] + sc[sl][q & 1] + zc[sl][q >> 1] + xs.x; }
    }
    long long t1 = clock64();
    float s = 0; for (int sl = 0; sl < 2; ++sl) for (int q = 0; q < 4; ++q) s += tot[sl][q];
    if (s == 123.456f) out[1] = s;
    if (tid == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0) / iters;
}
template <int MODE> void run(const char* name) {
    float* out; cudaMalloc(&out, 8);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<MODE><<<148, 512, 65536>>>(out, 64, 1u << 28, 1u << 20);
    k<MODE><<<148, 512, 65536>>>(out, 2048, 1u << 28, 1u << 20);
    float h[2]; cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-48s %7.1f cycles per unit per warp (16 warps/SM: 4 per SMSP) -> %6.1f per SMSP-unit\n", name, h[0], h[0] / 4);
    cudaFree(out);
}
int main() {
    run<1 | 2 | 4>("full (shifts on ALU)");
    run<1 | 2 | 4 | 8>("full (2 of 3 shifts as IMAD.HI)");
    run<1 | 2 | 4 | 8 | 16>("full, IMAD.HI, independent k-step accumulators");
    run<2 | 4>("no unpack");
    run<1 | 4 | 8>("no mma");
    run<1 | 2 | 8>("no ffma");
    run<2>("mma only");
    run<2 | 16>("mma only, independent accumulators");
    run<1 | 8>("unpack only");
    return 0;
}
