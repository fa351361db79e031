// ALU-pipe micro-benchmark behind the score-pass analysis (not part of the product): issue rate per SM of the integer
// instructions the 2-bit -> int8 expansion can be built from.  8 independent dependent chains per thread, 32 warps/SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu alu.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k_op(unsigned *out, int iters, unsigned lut_a, unsigned lut_b, unsigned mask) {
    unsigned x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x * 2654435761u + k * 40503u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (OP == 0) asm volatile("prmt.b32 %0, %1, %2, %0;" : "+r"(x[k]) : "r"(lut_a), "r"(lut_b));          // PRMT, selector = data (incl. sign-replicate nibbles)
            if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x[k]) : "r"(lut_a), "r"(mask));     // LOP3
            if (OP == 2) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(lut_b | 0x80000001u));           // IMAD.HI
            if (OP == 3) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(lut_a), "r"(mask & 7)); // SHF
            if (OP == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(lut_b | 5u), "r"(mask));     // IMAD
            if (OP == 5) { asm volatile("prmt.b32 %0, %1, %2, %0;" : "+r"(x[k]) : "r"(lut_a), "r"(lut_b));
                           asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[(k + 4) & 7]) : "r"(lut_b | 0x80000001u)); } // PRMT + IMAD.HI co-issue
            if (OP == 6) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x[k]) : "r"(lut_a), "r"(mask));
                           asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[(k + 4) & 7]) : "r"(lut_b | 0x80000001u)); } // LOP3 + IMAD.HI co-issue
        }
    }
    unsigned s = 0;
    for (int k = 0; k < 8; k++) s ^= x[k];
    if (s == 0x12345678u) out[0] = s;
}
template <int OP> void run(const char *name, int sms, int per_iter) {
    unsigned *d; cudaMalloc(&d, 64);
    const int iters = 1 << 14, blocks = sms * 4, threads = 256;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_op<OP><<<blocks, threads>>>(d, 256, 0xFEFEFEFEu, 0u, 0x01010101u);
    cudaEventRecord(a);
    k_op<OP><<<blocks, threads>>>(d, iters, 0xFEFEFEFEu, 0u, 0x01010101u);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ops = (double)iters * per_iter * blocks * threads;
    printf("%-22s %.3f ms  %.1f lane-ops/clk/SM (at %d MHz), %s\n", name, ms, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}
int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("PRMT (reg selector)", p.multiProcessorCount, 8);
    run<1>("LOP3", p.multiProcessorCount, 8);
    run<2>("IMAD.HI", p.multiProcessorCount, 8);
    run<3>("SHF.R", p.multiProcessorCount, 8);
    run<4>("IMAD", p.multiProcessorCount, 8);
    run<5>("PRMT + IMAD.HI pairs", p.multiProcessorCount, 16);
    run<6>("LOP3 + IMAD.HI pairs", p.multiProcessorCount, 16);
    return 0;
}
