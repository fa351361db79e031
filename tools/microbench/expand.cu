// Micro-benchmark behind the score-pass analysis (not part of the product): cycles per packed 32-bit word (16 samples)
// of the 2-bit -> int8 expansion body (2 masks, 3 right shifts, 4 PRMT) for different ways of doing the shifts, at the
// occupancy the score kernel has (3 or 4 expanding warps per SM sub-partition) and at full occupancy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o expand expand.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned lop(unsigned x) { unsigned r; asm volatile("and.b32 %0, %1, 0x77777777;" : "=r"(r) : "r"(x) : "memory"); return r; }
__device__ __forceinline__ unsigned prm(unsigned lut, unsigned s) { unsigned r; asm volatile("prmt.b32 %0, %1, %1, %2;" : "=r"(r) : "r"(lut), "r"(s) : "memory"); return r; }
__device__ __forceinline__ unsigned mulhi(unsigned x, unsigned k) { unsigned r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(k) : "memory"); return r; }
__device__ __forceinline__ unsigned shr(unsigned x, unsigned k) { unsigned r; asm volatile("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(k) : "memory"); return r; }
__device__ __forceinline__ unsigned widehi(unsigned x, unsigned k) {
    unsigned lo, hi;
    asm volatile("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(x), "r"(k) : "memory");
    return hi ^ (lo & 0);
}
__device__ __forceinline__ unsigned bytes_hi(unsigned x) { unsigned r; asm volatile("prmt.b32 %0, %1, %1, 0x4432;" : "=r"(r) : "r"(x) : "memory"); return r; }

// V: 0 = three IMAD.HI (the kernel), 1 = three SHF, 2 = >>2 IMAD.HI and >>16 SHF, 3 = >>2 SHF and >>16 IMAD.HI,
//    4 = three IMAD.WIDE, 5 = >>16 as a byte PRMT and >>2 IMAD.HI, 6 = no shifts at all (floor: 2 masks + 4 PRMT),
//    7 = >>2 IMAD.HI, one >>16 IMAD.HI, one >>16 SHF
template <int V, bool NOSTORE = false>
__global__ void __launch_bounds__(1024, 1) k_expand(unsigned *out, int iters, unsigned lut, unsigned k30, unsigned k16, unsigned two, unsigned sixteen) {
    // the expanded words leave the registers as in the kernel: tcgen05.st 32x32b.x16 into this warp's lane quarter
    __shared__ unsigned slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned taddr = slot + (((unsigned)(warp & 3) * 32u) << 16) + (unsigned)((warp >> 2) & 7) * 64u;
    unsigned x[16], d[16];            // 16 live words, each refreshed once per iteration (one IMAD per word: ptxas would
                                      // otherwise hoist the loop-invariant expansion out of the loop)
    for (int k = 0; k < 16; k++) x[k] = threadIdx.x * 2654435761u + k * 40503u + blockIdx.x;
    for (int k = 0; k < 16; k++) d[k] = 0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned xw = x[4 * c + k];
                x[4 * c + k] = xw * k16 + two;
                const unsigned xe = lop(xw);
                unsigned t;
                if (V == 0 || V == 2 || V == 5 || V == 7) t = mulhi(xw, k30);
                else if (V == 4) t = widehi(xw, k30);
                else if (V == 6) t = xw;
                else t = shr(xw, two);
                const unsigned xo = lop(t);
                unsigned he, ho;
                if (V == 0 || V == 3) { ho = mulhi(xo, k16); he = mulhi(xe, k16); }
                else if (V == 1 || V == 2) { ho = shr(xo, sixteen); he = shr(xe, sixteen); }
                else if (V == 4) { ho = widehi(xo, k16); he = widehi(xe, k16); }
                else if (V == 5) { ho = bytes_hi(xo); he = bytes_hi(xe); }
                else if (V == 7) { ho = mulhi(xo, k16); he = shr(xe, sixteen); }
                else { ho = xo; he = xe; }
                d[4 * k + 0] = prm(lut, xo);
                d[4 * k + 1] = prm(lut, xe);
                d[4 * k + 2] = prm(lut, ho);
                d[4 * k + 3] = prm(lut, he);
            }
            if (!NOSTORE)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                             ::"r"(taddr + (unsigned)c * 16u), "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]),
                               "r"(d[8]), "r"(d[9]), "r"(d[10]), "r"(d[11]), "r"(d[12]), "r"(d[13]), "r"(d[14]), "r"(d[15]) : "memory");
            else {
                unsigned v = 0;
#pragma unroll
                for (int k = 0; k < 16; k++) v |= d[k];
                if (v == 0x12345679u) out[1] = v;
            }
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    unsigned s = 0;
    for (int k = 0; k < 16; k++) s ^= d[k] ^ x[k];
    if (s == 0x12345678u) out[0] = s;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int V> void run(const char *name, int sms, int warps_per_sm) {
    unsigned *d; cudaMalloc(&d, 64);
    const int iters = 1 << 12, blocks = sms, threads = warps_per_sm * 32;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_expand<V><<<blocks, threads>>>(d, 64, 0x00018002u, 1u << 30, 1u << 16, 2, 16);
    cudaEventRecord(a);
    k_expand<V><<<blocks, threads>>>(d, iters, 0x00018002u, 1u << 30, 1u << 16, 2, 16);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double words_per_smsp = (double)iters * 16 * warps_per_sm / 4;      // warp-words per sub-partition
    printf("%-44s %2d warps/SM  %.3f ms  %.2f clk per warp-word per sub-partition (%s)\n", name, warps_per_sm, ms,
           ms * 1e-3 * clk * 1e3 / words_per_smsp, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}
int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int w : {12, 16, 32}) {
        run<0>("3 x IMAD.HI (kernel)", p.multiProcessorCount, w);
        run<1>("3 x SHF", p.multiProcessorCount, w);
        run<2>(">>2 IMAD.HI, >>16 SHF", p.multiProcessorCount, w);
        run<3>(">>2 SHF, >>16 IMAD.HI", p.multiProcessorCount, w);
        run<4>("3 x IMAD.WIDE", p.multiProcessorCount, w);
        run<5>(">>2 IMAD.HI, >>16 byte PRMT", p.multiProcessorCount, w);
        run<7>(">>2 IMAD.HI, >>16 one IMAD.HI one SHF", p.multiProcessorCount, w);
        run<6>("no shifts (floor: 2 LOP3 + 4 PRMT)", p.multiProcessorCount, w);
    }
    return 0;
}
