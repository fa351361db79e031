// score.cuh -- K1+K2: fused 2-bit unpack + score pass (replaces GRAB.ReadGeno + ref :557-570, :579-603).
//
// For every SNP row of packed PLINK codes the pass produces, without ever materialising the
// N-double genotype vector the reference builds (`g = Geno.mtx[,i]`, ref :510):
//     asum  = n_het + 2 n_homA1          (integer, exact)
//     nmiss = number of 01 codes         (integer, exact)
//     D0,D1 = sum over non-missing of dosage_i * R_i, dosage_i * V_i
//     T0,T1 = sum over missing of R_i, V_i        (mean imputation is applied algebraically later)
// hom-A2 samples (code 11) contribute nothing; flips (MAF > 0.5) are applied algebraically in K3.
//
// v1 (this file): fp64 CUDA-core kernel.  Each thread owns 16 samples (one 32-bit word column) with
// R/V for those samples held in registers, and walks down a tile of SNP rows, so the residual
// vectors are read once per (CTA, 128 rows) instead of once per SNP.
#pragma once
#include "types.cuh"
#include <cuda_runtime.h>

namespace spagxe {

__device__ __forceinline__ double score_warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int kScoreThreads = 256;
constexpr int kScoreRows = 128;

__global__ void __launch_bounds__(kScoreThreads)
k_score_v1(const uint32_t *__restrict__ bed, int64_t pitch_words, int64_t n, int m_chunk, int m_pad,
           const double *__restrict__ R, const double *__restrict__ V, double *__restrict__ partial,
           int *__restrict__ ipartial) {
    __shared__ double s_part[kScoreRows][kScoreThreads / 32][4];
    __shared__ int s_ipart[kScoreRows][kScoreThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t wcol = (int64_t)blockIdx.x * kScoreThreads + threadIdx.x;
    const int64_t nwords = (n + 15) >> 4;
    const int row0 = blockIdx.y * kScoreRows;
    const int nrows = min(kScoreRows, m_chunk - row0);
    double r[16], v[16];
    const int64_t i0 = wcol * 16;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const bool ok = (i0 + k) < n;
        r[k] = ok ? R[i0 + k] : 0.0;
        v[k] = ok ? V[i0 + k] : 0.0;
    }
    uint32_t tailmask = 0;  // bits to force to 11 (samples >= n)
    if (wcol >= nwords) tailmask = 0xFFFFFFFFu;
    else if (n - i0 < 16) tailmask = 0xFFFFFFFFu << (2 * (int)(n - i0));
    const uint32_t *p = bed + (int64_t)row0 * pitch_words + wcol;
    const bool inb = wcol < pitch_words;
    for (int rr = 0; rr < nrows; rr++) {
        uint32_t w = inb ? __ldg(p + (int64_t)rr * pitch_words) : 0xFFFFFFFFu;
        w |= tailmask;
        double d0 = 0, d1 = 0, t0 = 0, t1 = 0;
        int cnt = 0;
        if (w != 0xFFFFFFFFu) {
            const uint32_t lo = w & 0x55555555u, hi = (w >> 1) & 0x55555555u;
            const uint32_t m_hom = ~lo & ~hi & 0x55555555u, m_het = ~lo & hi, m_mis = lo & ~hi;
            cnt = 2 * __popc(m_hom) + __popc(m_het) + (__popc(m_mis) << 16);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t c = (w >> (2 * k)) & 3u;
                const double wd = (c == 0u) ? 2.0 : ((c == 2u) ? 1.0 : 0.0);
                d0 = fma(wd, r[k], d0);
                d1 = fma(wd, v[k], d1);
            }
            if (m_mis) {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if ((m_mis >> (2 * k)) & 1u) { t0 += r[k]; t1 += v[k]; }
                }
            }
        }
        d0 = score_warp_sum(d0); d1 = score_warp_sum(d1); t0 = score_warp_sum(t0); t1 = score_warp_sum(t1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) {
            s_part[rr][warp][0] = d0; s_part[rr][warp][1] = d1;
            s_part[rr][warp][2] = t0; s_part[rr][warp][3] = t1;
            s_ipart[rr][warp] = cnt;
        }
    }
    __syncthreads();
    // fixed-order cross-warp sum; layout partial[(tile_s*4 + q) * m_pad + row]
    for (int idx = threadIdx.x; idx < nrows * 5; idx += kScoreThreads) {
        const int rr = idx / 5, q = idx % 5;
        if (q < 4) {
            double s = 0;
#pragma unroll
            for (int ww = 0; ww < kScoreThreads / 32; ww++) s += s_part[rr][ww][q];
            partial[((int64_t)blockIdx.x * 4 + q) * m_pad + row0 + rr] = s;
        } else {
            int a = 0, mm = 0;
#pragma unroll
            for (int ww = 0; ww < kScoreThreads / 32; ww++) { int c = s_ipart[rr][ww]; a += c & 0xFFFF; mm += c >> 16; }
            ipartial[((int64_t)blockIdx.x * 2 + 0) * m_pad + row0 + rr] = a;
            ipartial[((int64_t)blockIdx.x * 2 + 1) * m_pad + row0 + rr] = mm;
        }
    }
}

// fixed-order reduction over sample tiles. One thread per SNP row.
__global__ void k_score_reduce(const double *__restrict__ partial, const int *__restrict__ ipartial, int tiles_s,
                               int m_chunk, int m_pad, SnpStats st) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m_chunk) return;
    double d0 = 0, d1 = 0, t0 = 0, t1 = 0;
    int a = 0, mm = 0;
    for (int ts = 0; ts < tiles_s; ts++) {
        d0 += partial[((int64_t)ts * 4 + 0) * m_pad + j];
        d1 += partial[((int64_t)ts * 4 + 1) * m_pad + j];
        t0 += partial[((int64_t)ts * 4 + 2) * m_pad + j];
        t1 += partial[((int64_t)ts * 4 + 3) * m_pad + j];
        a += ipartial[((int64_t)ts * 2 + 0) * m_pad + j];
        mm += ipartial[((int64_t)ts * 2 + 1) * m_pad + j];
    }
    st.D0[j] = d0; st.D1[j] = d1; st.T0[j] = t0; st.T1[j] = t1; st.asum[j] = a; st.nmiss[j] = mm;
}

}  // namespace spagxe
