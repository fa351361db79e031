// Micro-benchmarks that decide the score-pass design on B200 (not part of the product):
//   legacy mma.sync m16n8k32 s8 rate, PRMT/LOP3 (ALU pipe) rate, DFMA rate.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_imma(int *out, int iters) {
    int c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0}, c3[4] = {0, 0, 0, 0};
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    for (int i = 0; i < iters; i++) {
#define MMA(c) asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        MMA(c0) MMA(c1) MMA(c2) MMA(c3)
    }
    if (c0[0] + c1[1] + c2[2] + c3[3] == 0x7fffffff) out[0] = 1;
}
__global__ void k_prmt(unsigned *out, int iters) {
    unsigned x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    const unsigned lut = 0x00010002;
    for (int i = 0; i < iters; i++) {
        x0 = __byte_perm(lut, 0, x0 & 0x3333); x1 = __byte_perm(lut, 0, x1 & 0x3333);
        x2 = __byte_perm(lut, 0, x2 & 0x3333); x3 = __byte_perm(lut, 0, x3 & 0x3333);
        x4 = __byte_perm(lut, 0, x4 & 0x3333); x5 = __byte_perm(lut, 0, x5 & 0x3333);
        x6 = __byte_perm(lut, 0, x6 & 0x3333); x7 = __byte_perm(lut, 0, x7 & 0x3333);
    }
    if ((x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7) == 0xdeadbeef) out[0] = 1;
}
__global__ void k_dfma(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9);
        a4 = fma(a4, 1.0000001, 1e-9); a5 = fma(a5, 1.0000001, 1e-9); a6 = fma(a6, 1.0000001, 1e-9); a7 = fma(a7, 1.0000001, 1e-9);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 1.2345) out[0] = a0;
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    void *d; cudaMalloc(&d, 64);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, clock attr %d kHz\n", p.name, sms, clk);
    const int iters = 1 << 15, blocks = sms * 4, threads = 256;
    float ms = timeit([&] { k_imma<<<blocks, threads>>>((int *)d, iters); });
    double mmas = (double)iters * 4 * blocks * (threads / 32);
    printf("IMMA m16n8k32.s8: %.3f ms, %.1f TOPS, %.3f mma/clk/SM @1.9GHz\n", ms, mmas * 16 * 8 * 32 * 2 / ms / 1e9,
           mmas / (ms * 1e-3) / sms / 1.9e9);
    ms = timeit([&] { k_prmt<<<blocks, threads>>>((unsigned *)d, iters); });
    double ops = (double)iters * 16 * blocks * threads;
    printf("PRMT+LOP3 pairs: %.3f ms, %.1f lane-ops/clk/SM @1.9GHz\n", ms, ops / (ms * 1e-3) / sms / 1.9e9);
    ms = timeit([&] { k_dfma<<<blocks, threads>>>((double *)d, iters); });
    printf("DFMA: %.3f ms, %.2f TFLOP/s\n", ms, (double)iters * 8 * 2 * blocks * threads / ms / 1e9);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
