// Microbenchmark: legacy mma.sync.m16n8k16 (HMMA.16816.F32.BF16) issue rate per SM on sm_100a,
// as a function of resident warps and independent accumulator chains.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(float* out, int iters) {
    float c[CHAINS][4];
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    unsigned a0 = threadIdx.x * 0x01010101u, a1 = a0 ^ 0x3f803f80u, a2 = a0 + 7, a3 = a1 + 9, b0 = 0x3f803f80u, b1 = 0x40004000u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[0] = s;
}
template <int CHAINS> void run(int warps) {
    float* out; cudaMalloc(&out, 4);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CHAINS><<<148, warps * 32>>>(out, 16);
    cudaEventRecord(e0);
    k<CHAINS><<<148, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas_per_sm = (double)warps * iters * CHAINS;
    double cyc = ms * 1e-3 * 1.965e9;
    printf("warps/SM %2d chains %d: %.2f cycles per HMMA per SM (%.2f per SMSP), %.0f TFLOP/s dense equiv\n", warps, CHAINS,
           cyc / mmas_per_sm, 4 * cyc / mmas_per_sm, 148.0 * mmas_per_sm * 4096 / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); }
    return 0;
}
