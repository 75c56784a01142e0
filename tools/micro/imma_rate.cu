// Microbenchmark: legacy mma.sync.m16n8k32 s32.u8.s8 (IMMA.16832) issue rate per SM on sm_100a, and the
// rate of mul.hi.u32 used as a right shift next to LOP3 (ALU pipe vs FMA pipe).
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(int* out, int iters) {
    int c[CHAINS][4];
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0;
    unsigned a0 = threadIdx.x * 0x01010101u, a1 = a0 ^ 0x3f803f80u, a2 = a0 + 7, a3 = a1 + 9, b0 = 0x3f803f80u, b1 = 0x40004000u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0;
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123456) out[0] = s;
}
// MODE 0: 3 SHF + 4 LOP3 per word (ALU pipe only); MODE 1: 3 mul.hi + 4 LOP3
template <int MODE>
__global__ void alu(unsigned* out, int iters) {
    unsigned w[4] = {threadIdx.x * 2654435761u, threadIdx.x * 40503u + 1, threadIdx.x ^ 0x5bd1e995u, threadIdx.x + 77u}, acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            unsigned s4, s8, s12;
            if (MODE == 0) { s4 = w[q] >> 4; s8 = w[q] >> 8; s12 = w[q] >> 12; }
            else {
                asm volatile("mul.hi.u32 %0, %1, 0x10000000;" : "=r"(s4) : "r"(w[q]));
                asm volatile("mul.hi.u32 %0, %1, 0x01000000;" : "=r"(s8) : "r"(w[q]));
                asm volatile("mul.hi.u32 %0, %1, 0x00100000;" : "=r"(s12) : "r"(w[q]));
            }
            unsigned p0, p1, p2, p3;
            asm volatile("lop3.b32 %0, %1, 0x000F000F, 0x43004300, 0xEA;" : "=r"(p0) : "r"(w[q]));
            asm volatile("lop3.b32 %0, %1, 0x000F000F, 0x43004300, 0xEA;" : "=r"(p1) : "r"(s4));
            asm volatile("lop3.b32 %0, %1, 0x000F000F, 0x43004300, 0xEA;" : "=r"(p2) : "r"(s8));
            asm volatile("lop3.b32 %0, %1, 0x000F000F, 0x43004300, 0xEA;" : "=r"(p3) : "r"(s12));
            w[q] = p0 ^ p1 ^ p2 ^ p3 ^ (unsigned)it;     // 4 more LOP3-class ops
        }
    }
    for (int q = 0; q < 4; ++q) acc ^= w[q];
    if (acc == 0x12345) out[0] = acc;
}
template <int CHAINS> void run(int warps) {
    int* out; cudaMalloc(&out, 4);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CHAINS><<<148, warps * 32>>>(out, 16);
    cudaEventRecord(e0);
    k<CHAINS><<<148, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas_per_sm = (double)warps * iters * CHAINS;
    double cyc = ms * 1e-3 * 1.965e9;
    printf("IMMA.16832 warps/SM %2d chains %d: %.2f cycles per IMMA per SM (%.2f per SMSP)\n", warps, CHAINS, cyc / mmas_per_sm, 4 * cyc / mmas_per_sm);
    cudaFree(out);
}
template <int MODE> void run_alu(int warps) {
    unsigned* out; cudaMalloc(&out, 4);
    const int iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    alu<MODE><<<148, warps * 32>>>(out, 16);
    cudaEventRecord(e0);
    alu<MODE><<<148, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double words_per_smsp = (double)warps * iters * 4 / 4;
    printf("A-gen mode %d warps/SM %2d: %.2f cycles per warp-word per SMSP\n", MODE, warps, ms * 1e-3 * 1.965e9 / words_per_smsp);
    cudaFree(out);
}
int main() {
    for (int w : {8, 16}) { run<1>(w); run<2>(w); run<4>(w); }
    for (int w : {8, 16}) { run_alu<0>(w); run_alu<1>(w); }
    return 0;
}
