// kernels.cuh -- Step-2 device kernels (sm_100a).  See DESIGN.md for the data layout.
//
//   K0  k_vec_prep1/2     SNP-invariant sums and R.new            (ref :583, :592-594, :2034-2037)
//   K1  k_pack_dense      R matrix {0,1,2,NA} -> PLINK 2-bit rows  (Geno.mtx input, ref :483, :510)
//       k_bed_counts      per-code counts (bit-exact decode check)
//       k_bed_decode      2-bit -> double with NA_real_            (GRAB.ReadGeno, ref :484-487)
//   K2  score kernels live in score.cuh
//   K3  k_finalize_*      per-SNP scalar epilogue + routing lists  (ref :557-603, :2000-2045)
//   K4  k_spa_homo        Newton on the binomial CGF + tail        (ref :1795-1879, :2048-2062)
//   K5  bb_farm.cu        G-adjusted residual, SPA, Wald, CCT      (ref :604-678, :2070-2123)
#pragma once
#include "dmath.cuh"
#include "types.cuh"
#include "snp_common.cuh"
#include "stat_common.cuh"
#include <cooperative_groups.h>

// The kernels are spread over several translation units so that the heavy ones (minutes of ptxas each) compile side by
// side: a .cu file defines SPAGXE_TU before including this header and gets the kernels of its group only (0: all of them).
#ifndef SPAGXE_TU
#define SPAGXE_TU 0
#endif
#define SPAGXE_IN_TU(x) (SPAGXE_TU == 0 || SPAGXE_TU == (x))

namespace spagxe {

// ---------------------------------------------------------------------------
// K0: SNP-invariant vector statistics.  One CTA (deterministic tree).
#if SPAGXE_IN_TU(1)
__global__ void __launch_bounds__(1024) k_vec_prep1(const double *__restrict__ R, const double *__restrict__ E,
                                                    int64_t n, VecConst *vc) {
    __shared__ double scratch[3 * 32];
    double v[3] = {0, 0, 0};
    int bad = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double r = R[i], e = E[i], r2 = r * r;
        if (!isfinite(r) || !isfinite(e)) bad++;
        v[0] += r; v[1] += r2; v[2] += e * r2;
    }
    block_sum<3>(v, scratch);
    bad = __syncthreads_count(bad);
    if (threadIdx.x == 0) { vc->sumR = v[0]; vc->sumR2 = v[1]; vc->sumER2 = v[2]; vc->lambda = v[2] / v[1]; vc->nonfinite = bad; }
}
#endif
// mode 0: V = R.new (CCT); mode 1: V = E*R (mix / plus).  Rn_in != NULL: use the caller's R.new (plus).
#if SPAGXE_IN_TU(1)
__global__ void __launch_bounds__(1024) k_vec_prep2(const double *__restrict__ R, const double *__restrict__ E,
                                                    const double *__restrict__ Rn_in, int64_t n, int mode,
                                                    double *__restrict__ Rn, double *__restrict__ V, VecConst *vc) {
    __shared__ double scratch[4 * 32];
    const double lambda = vc->lambda;
    double v[4] = {0, 0, 0, 0};
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double r = R[i], e = E[i];
        double rn = Rn_in ? Rn_in[i] : (e - lambda) * r;
        Rn[i] = rn;
        double vv = (mode == 0) ? rn : e * r;
        V[i] = vv;
        v[0] += rn; v[1] += rn * rn; v[2] += vv; v[3] += vv * vv;
    }
    block_sum<4>(v, scratch);
    if (threadIdx.x == 0) { vc->sumRn = v[0]; vc->sumRn2 = v[1]; vc->sumV = v[2]; vc->sumV2 = v[3]; }
}
#endif

// ---------------------------------------------------------------------------
// K1 helpers
// dense column-major matrix -> packed rows. One thread per output 32-bit word (16 samples).
template <typename T>
__device__ __forceinline__ int dense_code(T x, int *bad);
template <>
__device__ __forceinline__ int dense_code<double>(double x, int *bad) {
    if (isnan(x)) return 1;
    if (x == 0.0) return 3;
    if (x == 1.0) return 2;
    if (x == 2.0) return 0;
    *bad = 1; return 1;
}
template <>
__device__ __forceinline__ int dense_code<int>(int x, int *bad) {
    if (x == INT_MIN) return 1;
    if (x == 0) return 3;
    if (x == 1) return 2;
    if (x == 2) return 0;
    *bad = 1; return 1;
}
template <typename T>
__global__ void k_pack_dense(const T *__restrict__ G, int64_t ld, int64_t n, int m, uint32_t *__restrict__ bed,
                             int64_t pitch_words, Counters *ctr) {
    const int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (w >= pitch_words || j >= m) return;
    const T *col = G + (int64_t)j * ld;
    uint32_t word = 0xFFFFFFFFu;
    int bad = 0;
    const int64_t i0 = w * 16;
    if (i0 < n) {
        word = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            int c = 3;
            if (i0 + k < n) c = dense_code<T>(col[i0 + k], &bad);
            word |= (uint32_t)c << (2 * k);
        }
    }
    bed[(int64_t)j * pitch_words + w] = word;
    if (bad) atomicAdd(&ctr->n_nonint, 1ULL);
}

// GRAB.ReadGeno's sample subset / reorder and allele order as a device gather, 2 bits -> 2 bits (ref :484-487).
// One thread per output 32-bit word (16 samples); idx == NULL keeps the sample order; padding samples get code 11.
#if SPAGXE_IN_TU(1)
__global__ void k_bed_prepare(const uint8_t *__restrict__ src, int64_t spitch, uint32_t *__restrict__ dst, int64_t dpitch_words,
                              int64_t n, int m, const int64_t *__restrict__ idx, int count_a2) {
    const int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (w >= dpitch_words || j >= m) return;
    const uint8_t *row = src + (int64_t)j * spitch;
    uint32_t word = 0xFFFFFFFFu;
    const int64_t i0 = w * 16;
    if (i0 < n) {
        word = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            int c = 3;
            if (i0 + k < n) {
                const int64_t sidx = idx ? idx[i0 + k] : i0 + k;
                c = (row[sidx >> 2] >> (2 * (int)(sidx & 3))) & 3;
                if (count_a2) c = (c == 0) ? 3 : (c == 3) ? 0 : c;   // hom-A1 <-> hom-A2; het and missing stay
            }
            word |= (uint32_t)c << (2 * k);
        }
    }
    dst[(int64_t)j * dpitch_words + w] = word;
}
#endif

// per-code counts (exact integers). One warp per SNP row; pitch multiple of 4.
#if SPAGXE_IN_TU(1)
__global__ void k_bed_counts(const uint8_t *__restrict__ bed, int64_t pitch, int64_t n, int m, long long *counts) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= m) return;
    const uint32_t *row = reinterpret_cast<const uint32_t *>(bed + (int64_t)warp * pitch);
    const int64_t nw = (n + 15) >> 4;
    int c0 = 0, c1 = 0, c2 = 0;
    for (int64_t w = lane; w < nw; w += 32) {
        uint32_t x = row[w];
        int64_t rem = n - w * 16;
        if (rem < 16) x |= 0xFFFFFFFFu << (2 * rem);
        uint32_t lo = x & 0x55555555u, hi = (x >> 1) & 0x55555555u;
        c0 += __popc(~lo & ~hi & 0x55555555u);
        c1 += __popc(lo & ~hi);
        c2 += __popc(~lo & hi);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, o);
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
    if (lane == 0) {
        counts[4 * warp + 0] = c0; counts[4 * warp + 1] = c1; counts[4 * warp + 2] = c2;
        counts[4 * warp + 3] = n - c0 - c1 - c2;
    }
}
#endif

#if SPAGXE_IN_TU(1)
__global__ void k_bed_decode(const uint8_t *__restrict__ bed, int64_t pitch, int64_t n, int m, double *__restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= n || j >= m) return;
    int c = (bed[(int64_t)j * pitch + (i >> 2)] >> (2 * (i & 3))) & 3;
    out[(int64_t)j * n + i] = (c == 0) ? 2.0 : (c == 2) ? 1.0 : (c == 3) ? 0.0 : d_na_real();
}
#endif

// K3 for SPAGxE_CCT (ref :543-603).  One thread per SNP of the chunk.
#if SPAGXE_IN_TU(1)
__global__ void k_finalize_cct(SnpStats st, int m_chunk, int64_t snp0, int64_t m_total, int64_t n,
                               const VecConst *__restrict__ vcp, double epsilon, double min_maf,
                               double *__restrict__ out, SpaItem *spa_list, BItem *b_list, const uint8_t *bed,
                               int64_t pitch, Counters *ctr, int g_model) {
    const int jl = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = jl < m_chunk;
    const VecConst vc = *vcp;
    bool want_spa = false, want_b = false;
    SpaItem item;
    const int64_t j = snp0 + jl;
    if (active) {
        const double NA = d_na_real();
        const int asum = st.asum[jl], nmiss = st.nmiss[jl];
        Prologue pr = make_prologue(asum, nmiss, n);
        double *o = out + j;
        o[0] = pr.maf; o[1 * m_total] = pr.miss_rate;
        if (nmiss == n) {  // mean(g, na.rm=T) is NaN; the reference stops at `if(MAF > 0.5)`
            for (int c = 0; c < 10; c++) o[c * m_total] = NA;
            o[1 * m_total] = 1.0;
            atomicAdd(&ctr->n_allmissing, 1ULL);
        } else if (pr.maf < min_maf) {
            for (int c = 2; c < 10; c++) o[c * m_total] = NA;
            atomicAdd(&ctr->n_filtered, 1ULL);
        } else {
            double S1 = st.D0[jl] + ((nmiss > 0) ? pr.u * st.T0[jl] : 0.0);
            double SV = st.D1[jl] + ((nmiss > 0) ? pr.u * st.T1[jl] : 0.0);
            if (pr.flip) { S1 = 2 * vc.sumR - S1; SV = 2 * vc.sumV - SV; }
            double sum_g = pr.sum_g;
            if (g_model != 0) {   // Dom / Rec (ref :573-574): g takes one value per genotype class, MAF stays additive
                const GClass gc = make_gclass(pr, g_model);
                S1 = class_dot(gc, st.D0[jl], st.T0[jl], st.H0[jl], vc.sumR);
                SV = class_dot(gc, st.D1[jl], st.T1[jl], st.H1[jl], vc.sumV);
                sum_g = class_sum(gc, make_gcounts(asum, nmiss, st.nhom[jl], n), 1);
            }
            const double VarG = 2 * pr.maf * (1 - pr.maf);
            const double VarS1 = vc.sumR2 * VarG;
            const double Z1 = S1 / sqrt(VarS1);
            const double pG = d_pnorm(-fabs(Z1), true) * 2;
            o[6 * m_total] = pG; o[7 * m_total] = S1; o[8 * m_total] = VarS1; o[9 * m_total] = Z1;
            if (pG > epsilon) {
                atomicAdd(&ctr->n_branch_a, 1ULL);
                double mafi, S, Smean, z;
                // Cutoff is not forwarded by the reference (:598): always 2, inner min.maf 1e-5
                int r = homo_scalar(sum_g, n, SV, vc.sumV, vc.sumV2, 0.00001, 2.0, &mafi, &S, &Smean, &z);
                if (r == 2) {
                    o[2 * m_total] = NA; o[3 * m_total] = NA; o[4 * m_total] = NA; o[5 * m_total] = NA;
                } else if (r == 0) {
                    double pn = d_pnorm(fabs(z), false) * 2;
                    o[2 * m_total] = pn; o[3 * m_total] = pn; o[4 * m_total] = pn; o[5 * m_total] = pn;
                } else {
                    o[5 * m_total] = d_pnorm(fabs(z), false) + d_pnorm(-fabs(z), true);
                    want_spa = true;
                    item.snp = (int)j; item.flags = 3;  // write p to 3 columns starting at col 2
                    item.maf = mafi; item.S = S; item.Smean = Smean; item.aux = 0;
                }
            } else {
                atomicAdd(&ctr->n_branch_b, 1ULL);
                want_b = true;
            }
        }
    }
    int pos = list_reserve(&ctr->spa_count, want_spa);
    if (want_spa) spa_list[pos] = item;
    pos = list_reserve(&ctr->b_count, want_b);
    if (want_b) b_list[pos] = BItem{bed + (int64_t)jl * pitch, (long long)j, st.asum[jl], st.nmiss[jl], st.D0[jl], st.T0[jl],
                                    g_model ? st.H0[jl] : 0.0, g_model ? st.nhom[jl] : 0, 0};
}
#endif

// Reduction policies: every thread of every participating CTA gets the same fixed-order total.
struct BlockRed {
    double *scratch;  // >= 4*32 doubles of shared memory
    template <int NV>
    __device__ __forceinline__ void sum(double (&v)[NV]) const { block_sum<NV>(v, scratch); }
    __device__ __forceinline__ int rank() const { return 0; }
    __device__ __forceinline__ int size() const { return 1; }
};
// A thread-block cluster cooperating on one SNP: CTA totals are exchanged through distributed
// shared memory and added in rank order by everyone.
struct ClusterRed {
    double *scratch;  // block scratch (>= 4*32 doubles)
    double *slot;     // this CTA's exchange slots in shared memory: two alternating buffers of 4 doubles
    mutable int par = 0;
    // ONE cluster barrier per call: a buffer is rewritten two calls later, and every CTA has passed the barrier of the
    // call in between after it finished reading, so no reader can still be on it
    template <int NV>
    __device__ __forceinline__ void sum(double (&v)[NV]) const {
        static_assert(NV <= 4, "exchange slot holds 4 doubles");
        namespace cg = cooperative_groups;
        cg::cluster_group cl = cg::this_cluster();
        block_sum<NV>(v, scratch);
        double *mine = slot + 4 * par;
        par ^= 1;
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < NV; k++) mine[k] = v[k];
        }
        cl.sync();
        const unsigned cs = cl.num_blocks();
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] = 0;
        for (unsigned r = 0; r < cs; r++) {
            const double *rs = cl.map_shared_rank(mine, r);
#pragma unroll
            for (int k = 0; k < NV; k++) v[k] += rs[k];
        }
    }
    __device__ __forceinline__ int rank() const { return (int)cooperative_groups::this_cluster().block_rank(); }
    __device__ __forceinline__ int size() const { return (int)cooperative_groups::this_cluster().num_blocks(); }
};

// Collective: fastgetroot_H1_new + GetProb_SPA_G_new (ref :1795-1879). All threads return p.
// Src::get(i, r, maf, w): term i contributes w * phi(r); exact sources have w == 1.
// max_abs_t: quadrature sources stop (return -1) when |t| leaves the range their error bound covers.
template <class Src, class Red>
__device__ double spa_tail(const Src &src, const Red &red, int64_t i_lo, int64_t i_hi, double s, bool lower_tail,
                           double max_abs_t, int *iters) {
    const double tol = 0x1p-13;
    double t = 0, H1 = 0, diff = CUDART_INF, H2e = 0;
    int iter;
    for (iter = 1; iter <= 100; iter++) {
        const double old_t = t, old_diff = diff, old_H1 = H1;
        if (fabs(t) > max_abs_t) { *iters = 0; return -1.0; }
        double v[2] = {0, 0};
        // two samples in flight: the exp / reciprocal chains interleave, the per-thread order of the adds is unchanged
#pragma unroll 2
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            double r, maf, w, k1, k2;
            src.get(i, r, maf, w);
            cgf_k12(t * r, maf, k1, k2);
            v[0] += w * (r * k1);
            v[1] += w * ((r * r) * k2);
        }
        red.template sum<2>(v);
        H1 = v[0] - s; H2e = v[1];
        diff = -1 * H1 / H2e;
        if (isnan(diff)) diff = 5;
        if (isnan(H1) || isinf(H1)) { t = d_sign(s) * CUDART_INF; H2e = 0; break; }
        if (d_sign(H1) != d_sign(old_H1)) {
            int guard = 0;
            while (fabs(diff) > fabs(old_diff) - tol && guard++ < 1200) diff = diff / 2;
        }
        if (isnan(diff)) diff = 5;
        if (fabs(diff) < tol) break;
        t = old_t + diff;
    }
    if (iter > 100) iter = 100;
    *iters = iter;
    if (iter == 100) t = CUDART_NAN;
    if (!isfinite(t)) return 0.0;  // every non-finite root yields NA -> 0 in the reference (:1873)
    if (fabs(t) > max_abs_t) { *iters = 0; return -1.0; }
    double v[2] = {0, 0};
#pragma unroll 2
    for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
        double r, maf, w, k1, k2;
        src.get(i, r, maf, w);
        v[0] += w * cgf_k0(t * r, maf);
        cgf_k12(t * r, maf, k1, k2);
        v[1] += w * ((r * r) * k2);
    }
    red.template sum<2>(v);
    const double temp1 = t * s - v[0];
    const double w = d_sign(t) * sqrt(2 * temp1);
    const double vv = t * sqrt(v[1]);
    double pval = d_pnorm(w + 1 / w * log(vv / w), lower_tail);
    if (isnan(pval)) pval = 0;
    return pval;
}

// two tails (ref :2048-2062). All threads return pval.spa.G (or -1 when a quadrature source bailed out)
template <class Src, class Red>
__device__ double spa_two_tail(const Src &src, const Red &red, int64_t i_lo, int64_t i_hi, double s_hi, double s_lo,
                               double max_abs_t, int *iters) {
    int i1, i2;
    double pval1 = spa_tail(src, red, i_lo, i_hi, s_hi, false, max_abs_t, &i1);
    if (pval1 < 0) return -1.0;
    double pval2 = spa_tail(src, red, i_lo, i_hi, s_lo, true, max_abs_t, &i2);
    if (pval2 < 0) return -1.0;
    if (isnan(pval2)) pval2 = pval1;
    *iters = i1 + i2;
    return pval1 + pval2;
}

// The same two tails advanced in lock step: one pass over the samples serves both Newton iterations (and both final
// evaluations), so a SNP needs max(iterations) + 1 passes instead of their sum + 2.  Every tail sees exactly the
// sequence of iterates and the summation order of spa_tail, so the result is bit-identical to spa_two_tail.
template <class Src, class Red>
__device__ double spa_two_tail_fused(const Src &src, const Red &red, int64_t i_lo, int64_t i_hi, double s_hi, double s_lo,
                                     int *iters, double max_abs_t = 1.7976931348623157e308, double arg_scale = 1.0) {
    const double tol = 0x1p-13;
    const double sv[2] = {s_hi, s_lo};
    double t[2] = {0, 0}, H1[2] = {0, 0}, diff[2] = {CUDART_INF, CUDART_INF}, pv[2] = {0, 0};
    int it[2] = {0, 0}, phase[2] = {0, 0};   // 0 Newton, 1 final evaluation pending, 2 done
    for (;;) {
        if (phase[0] == 2 && phase[1] == 2) break;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (phase[h] == 0 && ++it[h] > 100) { it[h] = 100; pv[h] = 0.0; phase[h] = 2; }   // maxiter: t = NA -> p = 0 (ref :1846, :1873)
        }
        if (phase[0] == 2 && phase[1] == 2) break;
        // a source with a bounded domain (the quadrature): an iterate or root outside it hands the SNP back, as spa_tail does
        if ((phase[0] != 2 && fabs(t[0]) > max_abs_t) || (phase[1] != 2 && fabs(t[1]) > max_abs_t)) { *iters = 0; return -1.0; }
        double v[4] = {0, 0, 0, 0};
#pragma unroll 2
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            double r, maf, w, k1, k2;
            src.get(i, r, maf, w);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (phase[h] == 2) continue;
                cgf_k12(t[h] * r, maf, k1, k2);
                v[2 * h] += w * ((phase[h] == 0) ? (r * k1) : cgf_k0(t[h] * r, maf));
                v[2 * h + 1] += w * ((r * r) * k2);
            }
        }
        red.template sum<4>(v);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (phase[h] == 1) {
                const double temp1 = t[h] * sv[h] - v[2 * h];
                const double w = d_sign(t[h]) * sqrt(2 * temp1);
                const double vv = t[h] * sqrt(v[2 * h + 1]);
                // arg_scale = sqrt(Var.ratio) of GetProb_SPA_G_new_adjust (mixPlus :456), 1 everywhere else (exact)
                double pval = d_pnorm((w + 1 / w * log(vv / w)) * arg_scale, h == 1);
                if (isnan(pval)) pval = 0;
                pv[h] = pval; phase[h] = 2;
            } else if (phase[h] == 0) {
                const double old_t = t[h], old_diff = diff[h], old_H1 = H1[h];
                H1[h] = v[2 * h] - sv[h];
                double d = -1 * H1[h] / v[2 * h + 1];
                if (isnan(d)) d = 5;
                bool brk = false;
                if (isnan(H1[h]) || isinf(H1[h])) { t[h] = d_sign(sv[h]) * CUDART_INF; brk = true; }
                else {
                    if (d_sign(H1[h]) != d_sign(old_H1)) {
                        int guard = 0;
                        while (fabs(d) > fabs(old_diff) - tol && guard++ < 1200) d = d / 2;
                    }
                    if (isnan(d)) d = 5;
                    if (fabs(d) < tol) brk = true;
                    else t[h] = old_t + d;
                }
                diff[h] = d;
                if (brk) {
                    if (it[h] == 100) t[h] = CUDART_NAN;
                    if (!isfinite(t[h])) { pv[h] = 0.0; phase[h] = 2; }   // every non-finite root yields NA -> 0 (ref :1873)
                    else phase[h] = 1;
                }
            }
        }
    }
    double p2 = pv[1];
    if (isnan(p2)) p2 = pv[0];
    *iters = it[0] + it[1];
    return pv[0] + p2;
}

struct SrcHomo {  // SNP-invariant vector, scalar MAF (SPAGxE_CCT branch A, SPAGxE_Plus)
    const double *rv; double maf;
    __device__ __forceinline__ void get(int64_t i, double &r, double &m, double &w) const { r = rv[i]; m = maf; w = 1.0; }
};

// ---- K4c: the same sums through a SNP-invariant piecewise-Chebyshev quadrature ----------------
// sum_i phi(R_i) = sum_{b,j} W[b][j] * phi(x[b][j]) + err, W[b][j] = sum_{i in bin b} L_j(R_i) (Lagrange
// basis on kQNodes Chebyshev points of bin b).  For the CGF terms phi(r) = r K1(tr), r^2 K2(tr), K0(tr)
// the error per sample is ~ (|t| h / (2 pi))^kQNodes / 2^(kQNodes-1) (see types.cuh), i.e. ~1e-13 relative while
// |t| h <= kQTheta; beyond that the item is re-done by the exact kernel.
__device__ __forceinline__ double cheb_node(int j) {  // cos((2j+1) pi / (2 kQNodes))
    return cospi((2.0 * j + 1.0) / (2.0 * kQNodes));
}

#if SPAGXE_IN_TU(1)
__global__ void __launch_bounds__(1024) k_quad_range(const double *__restrict__ rv, int64_t n, Quad q) {
    __shared__ double smin[32], smax[32];
    double mn = CUDART_INF, mx = -CUDART_INF;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { double r = rv[i]; mn = fmin(mn, r); mx = fmax(mx, r); }
    for (int o = 16; o > 0; o >>= 1) { mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 32; k++) { mn = fmin(mn, smin[k]); mx = fmax(mx, smax[k]); }
        const double h = (mx - mn) / kQBins;
        const bool ok = isfinite(h) && h > 0;
        q.range[0] = mn; q.range[1] = mx; q.range[2] = h; q.range[3] = ok ? kQTheta / h : 0.0;
    }
}
#endif
#if SPAGXE_IN_TU(1)
__global__ void k_quad_accumulate(const double *__restrict__ rv, int64_t n, Quad q) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const double mn = q.range[0], h = q.range[2];
    if (i >= n || !(q.range[3] > 0)) return;
    const double r = rv[i];
    int b = (int)floor((r - mn) / h);
    b = max(0, min(kQBins - 1, b));
    const double u = (r - (mn + (b + 0.5) * h)) / (0.5 * h);
    double xi[kQNodes];
#pragma unroll
    for (int j = 0; j < kQNodes; j++) xi[j] = cheb_node(j);
#pragma unroll
    for (int j = 0; j < kQNodes; j++) {
        double L = 1.0;
#pragma unroll
        for (int k = 0; k < kQNodes; k++) if (k != j) L *= (u - xi[k]) / (xi[j] - xi[k]);
        atomicAdd((unsigned long long *)&q.wfix[b * kQNodes + j], (unsigned long long)__double2ll_rn(L * kQScale));
    }
}
#endif
// nodes + weights, compacted in a fixed order to the bins that hold samples (single CTA, deterministic)
#if SPAGXE_IN_TU(1)
__global__ void __launch_bounds__(1024) k_quad_finish(Quad q) {
    constexpr int kTotal = kQBins * kQNodes, kPer = (kTotal + 1023) / 1024;
    __shared__ int s_cnt[1024];
    const double mn = q.range[0], h = q.range[2];
    const int i0 = threadIdx.x * kPer;
    int c = 0;
    for (int k = 0; k < kPer; k++) { const int idx = i0 + k; if (idx < kTotal && q.wfix[idx] != 0) c++; }
    s_cnt[threadIdx.x] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < 1024; t++) { int v = s_cnt[t]; s_cnt[t] = run; run += v; }
        *q.count = run;
    }
    __syncthreads();
    int pos = s_cnt[threadIdx.x];
    for (int k = 0; k < kPer; k++) {
        const int idx = i0 + k;
        if (idx < kTotal && q.wfix[idx] != 0) {
            const int b = idx / kQNodes, j = idx % kQNodes;
            q.x[pos] = mn + (b + 0.5) * h + 0.5 * h * cheb_node(j);
            q.w[pos] = (double)q.wfix[idx] / kQScale;
            pos++;
        }
    }
}
#endif
struct SrcQuad {
    const double *x, *wq; double maf;
    __device__ __forceinline__ void get(int64_t i, double &r, double &m, double &w) const { r = x[i]; m = maf; w = wq[i]; }
};

__device__ __forceinline__ void spa_item_tails(const SpaItem &it, double &s_hi, double &s_lo) {
    if (it.flags & 16) { s_hi = it.S; s_lo = it.Smean; }  // caller supplied both tails (plus)
    else { s_hi = fmax(it.S, 2 * it.Smean - it.S); s_lo = fmin(it.S, 2 * it.Smean - it.S); }
}
__device__ __forceinline__ void spa_item_write(const SpaItem &it, double p, int iters, int64_t m_total, double *out,
                                               Counters *ctr) {
    const int ncol = it.flags & 15;
    for (int c = 0; c < ncol; c++) out[(int64_t)it.snp + (int64_t)(2 + c) * m_total] = p;
    atomicAdd(&ctr->n_spa, 1ULL);
    atomicAdd(&ctr->newton_iters, (unsigned long long)iters);
}

// quadrature SPA: one CTA per item; items whose |t| leaves the covered range go to the exact list
#if SPAGXE_IN_TU(1)
__global__ void __launch_bounds__(kSpaQThreads) k_spa_quad(const SpaItem *__restrict__ list, Counters *ctr, Quad q,
                                                            SpaItem *__restrict__ exact_list, int64_t m_total,
                                                            double *__restrict__ out) {
    __shared__ double scratch[4 * 32];
    const int count = ctr->spa_count;
    const double max_t = q.range[3];
    BlockRed red{scratch};
    // dynamic queue: the Newton iteration counts of the items differ by an order of magnitude
    __shared__ int s_idx;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_idx = atomicAdd(&ctr->work_cursor, 1);
        __syncthreads();
        const int idx = s_idx;
        if (idx >= count) break;
        const SpaItem it = list[idx];
        double s_hi, s_lo;
        spa_item_tails(it, s_hi, s_lo);
        int iters = 0;
        double p = -1.0;
        if (max_t > 0) {
            SrcQuad src{q.x, q.w, it.maf};
            // (the fused two-tail walk was measured here too: 0.75 instead of 0.60 ms per 65,536-SNP step -- with ~10 nodes
            // per thread the pass is latency bound and two tails per pass do not shorten it)
            p = spa_two_tail(src, red, 0, (int64_t)(*q.count), s_hi, s_lo, max_t, &iters);
        }
        if (threadIdx.x == 0) {
            if (p < 0) exact_list[atomicAdd(&ctr->spax_count, 1)] = it;
            else spa_item_write(it, p, iters, m_total, out, ctr);
        }
    }
}
#endif

// exact SPA: persistent CTAs popping work items, N exp() per CGF evaluation
#if SPAGXE_IN_TU(1)
__global__ void __launch_bounds__(kSpaThreads) k_spa_homo(const SpaItem *__restrict__ list, const int *count_ptr,
                                                           Counters *ctr, const double *__restrict__ rv, int64_t n,
                                                           int64_t m_total, double *__restrict__ out) {
    __shared__ double scratch[4 * 32];
    const int count = *count_ptr;
    BlockRed red{scratch};
    for (int idx = blockIdx.x; idx < count; idx += gridDim.x) {
        const SpaItem it = list[idx];
        SrcHomo src{rv, it.maf};
        double s_hi, s_lo;
        spa_item_tails(it, s_hi, s_lo);
        int iters;
        const double p = spa_two_tail(src, red, 0, n, s_hi, s_lo, CUDART_INF, &iters);
        if (threadIdx.x == 0) spa_item_write(it, p, iters, m_total, out, ctr);
    }
}
#endif

// ---------------------------------------------------------------------------
// K5: branch B (ref :604-678).  One thread-block CLUSTER per SNP: each CTA owns a contiguous slice of
// the samples, CTA totals meet through distributed shared memory (ClusterRed), every CTA then runs
// the same scalar logic, so no broadcast is needed.
struct SrcAdj {  // R.new0_i = E_i * (R_i - fit[class_i]) computed from the packed row (ref :609-612)
    const uint8_t *row; const double *R, *E; double fit[4]; double maf;
    __device__ __forceinline__ void get(int64_t i, double &r, double &m, double &w) const {
        int c = (row[i >> 2] >> (2 * (i & 3))) & 3;
        r = E[i] * (R[i] - fit[c]);
        m = maf; w = 1.0;
    }
};

// Solve the equilibrated SPD system (single thread). A: q x q (ld kMaxQ), b: rhs.
// Returns false when a Cholesky pivot collapses (rank deficient design).
static __device__ bool chol_solve(double (*A)[kMaxQ], double *b, int q, double *x, double *inv_last) {
    double sc[kMaxQ];
    for (int i = 0; i < q; i++) { if (!(A[i][i] > 0)) return false; sc[i] = 1.0 / sqrt(A[i][i]); }
    for (int i = 0; i < q; i++) for (int j = 0; j < q; j++) A[i][j] *= sc[i] * sc[j];
    for (int j = 0; j < q; j++) {
        double d = A[j][j];
        for (int k = 0; k < j; k++) d -= A[j][k] * A[j][k];
        if (!(d > 1e-13)) return false;
        d = sqrt(d); A[j][j] = d;
        for (int i = j + 1; i < q; i++) {
            double sacc = A[i][j];
            for (int k = 0; k < j; k++) sacc -= A[i][k] * A[j][k];
            A[i][j] = sacc / d;
        }
    }
    double y[kMaxQ];
    for (int i = 0; i < q; i++) {
        double sacc = b[i] * sc[i];
        for (int k = 0; k < i; k++) sacc -= A[i][k] * y[k];
        y[i] = sacc / A[i][i];
    }
    for (int i = q - 1; i >= 0; i--) {
        double sacc = y[i];
        for (int k = i + 1; k < q; k++) sacc -= A[k][i] * x[k];
        x[i] = sacc / A[i][i];
    }
    for (int i = 0; i < q; i++) x[i] *= sc[i];
    const double l = A[q - 1][q - 1];
    *inv_last = sc[q - 1] * sc[q - 1] / (l * l);  // ((X'WX)^-1)[q-1][q-1]
    return true;
}

struct BShared {  // static shared state of the branch-B kernel
    double NE[kMaxQ][kMaxQ];
    double rhs[kMaxQ], beta[kMaxQ];
    double ne_part[kBMaxEnt];
    double scratch[4 * 32];
    double slot[8];
    double se2; int ok;
};

// dynamic smem: xs[kBTile][q+1] (x row, z in column q), ws[kBTile]
// One weighted-least-squares accumulation pass over this CTA's sample slice, then the cluster-wide
// normal equations land in sh.NE / sh.rhs of EVERY CTA.  Returns the cluster-wide deviance (binary)
// or residual sum of squares (lm second pass).  pass0: eta from mustart (ref glm.fit), else x.beta.
template <int NS, class Design>
__device__ double wald_pass(const Design &ds, int64_t i_lo, int64_t i_hi, int q, bool pass0, double *tile, BShared &sh,
                            const ClusterRed &red, int *invalid) {
    namespace cg = cooperative_groups;
    const int qs = (q + 1) | 1;  // odd row stride: conflict-free 64-bit reads
    const int trait = ds.trait();
    double *xs = tile, *ws = tile + (size_t)kBTile * qs;
    const int nent = q * (q + 3) / 2;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    // entries grp, grp + kBGroups, ... belong to this warp; all lanes of a warp share (a,b)
    int ea[NS], eb[NS];
    double acc[NS];
    DevProd dp;
#pragma unroll
    for (int k = 0; k < NS; k++) {
        acc[k] = 0; ea[k] = -1; eb[k] = 0;
        int e = grp + k * kBGroups;
        if (e < nent) {
            int a = 0, rem = e;
            for (a = 0; a < q; a++) { int len = q + 1 - a; if (rem < len) break; rem -= len; }
            ea[k] = a; eb[k] = a + rem;
        }
    }
    double dev = 0;
    int bad = 0;
    for (int64_t base = i_lo; base < i_hi; base += kBTile) {
        const int cnt = (int)min((int64_t)kBTile, i_hi - base);
        __syncthreads();
        if (threadIdx.x < cnt) {
            const int64_t i = base + threadIdx.x;
            double *xr = xs + (size_t)threadIdx.x * qs;
            const double y = ds.y(i);
            ds.fill_rt(i, xr, q);
            double eta = 0;
            if (!pass0) for (int c = 0; c < q; c++) eta += xr[c] * sh.beta[c];
            double w2, wz;
            irls_sample(trait, pass0, y, eta, w2, wz, dev, bad, dp);
            xr[q] = wz;   // column q already carries the weight
            ws[threadIdx.x] = w2;
        }
        __syncthreads();
        for (int sidx = lane; sidx < cnt; sidx += 32) {
            const double w = ws[sidx];
            const double *xr = xs + (size_t)sidx * qs;
#pragma unroll
            for (int k = 0; k < NS; k++)
                if (ea[k] >= 0) acc[k] = fma((eb[k] == q ? 1.0 : w) * xr[ea[k]], xr[eb[k]], acc[k]);
        }
    }
    dev += 2 * dp.log_value();
    // warp totals -> CTA partial normal equations -> cluster totals (rank order)
#pragma unroll
    for (int k = 0; k < NS; k++) {
        double t = warp_sum(acc[k]);
        int e = grp + k * kBGroups;
        if (lane == 0 && e < nent) sh.ne_part[e] = t;
    }
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();
    if (threadIdx.x < nent) {
        double t = 0;
        for (unsigned r = 0; r < cl.num_blocks(); r++) t += cl.map_shared_rank(sh.ne_part, r)[threadIdx.x];
        int a = 0, rem = threadIdx.x;
        for (a = 0; a < q; a++) { int len = q + 1 - a; if (rem < len) break; rem -= len; }
        const int b = a + rem;
        if (b == q) sh.rhs[a] = t; else { sh.NE[a][b] = t; sh.NE[b][a] = t; }
    }
    cl.sync();
    double v[2] = {dev, (double)bad};
    red.template sum<2>(v);
    *invalid = v[1] > 0;
    return v[0];
}

// Register-accumulator variant for narrow designs (Q = p + 4 <= 8): every thread keeps the whole upper
// triangle of X'WX (+ X'Wz) for its own samples in registers; no shared-memory tile, no per-tile barrier.
template <int Q, bool PASS0, class Design>
__device__ double wald_pass_reg(const Design &ds, int64_t i_lo, int64_t i_hi, double *tile, BShared &sh,
                                const ClusterRed &red, int *invalid) {
    namespace cg = cooperative_groups;
    constexpr bool pass0 = PASS0;   // compile-time: the first-pass table must not occupy registers in the later passes
    constexpr int NENT = Q * (Q + 3) / 2;
    const int trait = ds.trait();
    double acc[NENT];
#pragma unroll
    for (int e = 0; e < NENT; e++) acc[e] = 0;
    double beta[Q];
#pragma unroll
    for (int c = 0; c < Q; c++) beta[c] = sh.beta[c];
    DevProd dp;
    double dev = 0;
    int bad = 0;
    auto consume = [&](const double (&x)[Q], double y) {
        double eta = 0;
        if (!pass0) {
#pragma unroll
            for (int c = 0; c < Q; c++) eta = fma(x[c], beta[c], eta);
        }
        double w2, wz;
        irls_sample(trait, pass0, y, eta, w2, wz, dev, bad, dp);
        int e = 0;
#pragma unroll
        for (int a = 0; a < Q; a++) {
            const double wa = w2 * x[a];
#pragma unroll
            for (int b = a; b < Q; b++) { acc[e] = fma(wa, x[b], acc[e]); e++; }
            acc[e] = fma(x[a], wz, acc[e]); e++;
        }
    };
    auto accumulate = [&](const double (&x)[Q], double w2, double wz) {
        int e = 0;
#pragma unroll
        for (int a = 0; a < Q; a++) {
            const double wa = w2 * x[a];
#pragma unroll
            for (int b = a; b < Q; b++) { acc[e] = fma(wa, x[b], acc[e]); e++; }
            acc[e] = fma(x[a], wz, acc[e]); e++;
        }
    };
    // kWS samples per trip.  IRLS passes of a 0/1 response take a branch-free path in which the kWS exp / reciprocal
    // chains are independent instruction streams (the fp64 pipe is latency-bound at 16 warps per SM); every sum is
    // still updated in sample order, so the result does not depend on kWS.
    constexpr int kWS = 2;
    // first pass of glm.fit on a 0/1 response: eta = logit((y + 0.5) / 2) = -+log 3, so the working quantities take
    // two values; they are computed once per thread with the very operations irls_sample applies per sample
    double p0_eta[2] = {0, 0}, p0_ope[2] = {1, 1}, p0_w2[2] = {0, 0}, p0_wz[2] = {0, 0};
    if (pass0 && trait == 1) {
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const double y = (double)b, mus = (y + 0.5) / 2, e0 = log(mus / (1 - mus)), ee = exp(e0);
            const double opexp = 1 + ee, r = 1.0 / opexp, mu = ee * r, mev = mu * r;
            p0_eta[b] = e0; p0_ope[b] = opexp; p0_w2[b] = mev; p0_wz[b] = fma(mev, e0, y - mu);
        }
    }
    int64_t i = i_lo + threadIdx.x;
    for (; i + (kWS - 1) * (int64_t)blockDim.x < i_hi; i += kWS * (int64_t)blockDim.x) {
        double xs[kWS][Q], ys[kWS], eta[kWS];
        bool fast = (trait == 1);
#pragma unroll
        for (int k = 0; k < kWS; k++) {
            ys[k] = ds.y(i + k * (int64_t)blockDim.x);
            ds.template fill<Q>(i + k * (int64_t)blockDim.x, xs[k]);
            eta[k] = 0;
            if (!pass0) {
#pragma unroll
                for (int c = 0; c < Q; c++) eta[k] = fma(xs[k][c], beta[c], eta[k]);
            }
            fast = fast && (ys[k] == 0.0 || ys[k] == 1.0) && !(eta[k] > 30. || eta[k] < -30.);
        }
        if (fast) {
            double w2[kWS], wz[kWS], ope[kWS];
#pragma unroll
            for (int k = 0; k < kWS; k++) {   // same operations as the common branch of irls_sample
                if (pass0) {                   // mustart is 0.75 / 0.25: two possible values of everything
                    const bool one = (ys[k] == 1.0);   // selects, not indexed loads: the table stays in registers
                    eta[k] = one ? p0_eta[1] : p0_eta[0]; ope[k] = one ? p0_ope[1] : p0_ope[0];
                    w2[k] = one ? p0_w2[1] : p0_w2[0];    wz[k] = one ? p0_wz[1] : p0_wz[0];
                    continue;
                }
                const double ee = exp(eta[k]);
                ope[k] = 1 + ee;
                const double r = 1.0 / ope[k];
                const double mu = ee * r;
                w2[k] = mu * r;
                wz[k] = fma(w2[k], eta[k], ys[k] - mu);
            }
#pragma unroll
            for (int k = 0; k < kWS; k++) {
                dp.mul(ope[k]);
                dev -= 2 * (ys[k] * eta[k]);
                accumulate(xs[k], w2[k], wz[k]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kWS; k++) {
                double w2, wz;
                irls_sample(trait, pass0, ys[k], eta[k], w2, wz, dev, bad, dp);
                accumulate(xs[k], w2, wz);
            }
        }
    }
    for (; i < i_hi; i += blockDim.x) {
        double xa[Q];
        const double ya = ds.y(i);
        ds.template fill<Q>(i, xa);
        consume(xa, ya);
    }
    dev += 2 * dp.log_value();
    // warp totals -> per-warp rows in the (otherwise unused) tile -> CTA partial -> cluster total
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < NENT; e++) {
        const double t = warp_sum(acc[e]);
        if (lane == 0) tile[warp * NENT + e] = t;
    }
    __syncthreads();
    if (threadIdx.x < NENT) {
        double t = 0;
        for (int w = 0; w < nw; w++) t += tile[w * NENT + threadIdx.x];
        sh.ne_part[threadIdx.x] = t;
    }
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();
    if (threadIdx.x < NENT) {
        double t = 0;
        for (unsigned r = 0; r < cl.num_blocks(); r++) t += cl.map_shared_rank(sh.ne_part, r)[threadIdx.x];
        int a = 0, rem = threadIdx.x;
        for (a = 0; a < Q; a++) { int len = Q + 1 - a; if (rem < len) break; rem -= len; }
        const int b = a + rem;
        if (b == Q) sh.rhs[a] = t; else { sh.NE[a][b] = t; sh.NE[b][a] = t; }
    }
    cl.sync();
    double v[2] = {dev, (double)bad};
    red.template sum<2>(v);
    *invalid = v[1] > 0;
    return v[0];
}

struct GFromRow {
    const uint8_t *row; double gval[4];
    __device__ __forceinline__ double operator()(int64_t i) const { return gval[(row[i >> 2] >> (2 * (i & 3))) & 3]; }
};

__device__ __forceinline__ void wald_solve(BShared &sh, int q) {
    __syncthreads();
    if (threadIdx.x == 0) {
        double x[kMaxQ], il = 0;
        bool ok = chol_solve(sh.NE, sh.rhs, q, x, &il);
        for (int c = 0; c < q; c++) sh.beta[c] = ok ? x[c] : 0.0;
        sh.se2 = il; sh.ok = ok ? 1 : 0;
    }
    __syncthreads();
}

__device__ __forceinline__ void wald_solve(BShared &sh, int q);

// Design of the Wald model Y ~ Cova + E + g + E:g (ref :638, :652)
struct DesignWald {
    WaldData wd; const double *E; GFromRow gf; int64_t n;
    __device__ __forceinline__ int q() const { return wd.p + 4; }
    __device__ __forceinline__ int trait() const { return wd.trait; }
    __device__ __forceinline__ double y(int64_t i) const { return wd.Y[i]; }
    template <int Q>
    __device__ __forceinline__ void fill(int64_t i, double *x) const {
        constexpr int P = Q - 4;
        const double g = gf(i), ev = E[i];
        x[0] = 1.0;
#pragma unroll
        for (int c = 0; c < P; c++) x[1 + c] = wd.cova[i + (int64_t)c * n];
        x[P + 1] = ev; x[P + 2] = g; x[P + 3] = ev * g;
    }
    __device__ __forceinline__ void fill_rt(int64_t i, double *x, int q) const {
        const int p = q - 4;
        const double g = gf(i), ev = E[i];
        x[0] = 1.0;
        for (int c = 0; c < p; c++) x[1 + c] = wd.cova[i + (int64_t)c * n];
        x[p + 1] = ev; x[p + 2] = g; x[p + 3] = ev * g;
    }
};
// Design of the mix fallback glm(1{round(g) != 0} ~ selected PCs, binomial) (ref :1081-1089)
struct GDense {   // processed genotype: NA -> 2 MAF, flip, Dom / Rec (ref :562-574)
    const double *col; double imp; int flip, g_model;
    __device__ __forceinline__ double operator()(int64_t i) const {
        double g = col[i];
        if (isnan(g)) g = imp;
        if (flip) g = 2 - g;
        if (g_model == 1) g = (g >= 1) ? 1.0 : 0.0;
        if (g_model == 2) g = (g <= 1) ? 0.0 : 1.0;
        return g;
    }
};
struct SrcDenseAdj {  // R.new0_i = E_i (R_i - a - b g_i) (ref :609-612)
    const double *R, *E; GDense gf; double a, b, maf;
    __device__ __forceinline__ void get(int64_t i, double &r, double &m, double &w) const {
        r = E[i] * (R[i] - (a + b * gf(i))); m = maf; w = 1.0;
    }
};
struct DesignWaldDense {
    WaldData wd; const double *E; GDense gf; int64_t n;
    __device__ __forceinline__ int q() const { return wd.p + 4; }
    __device__ __forceinline__ int trait() const { return wd.trait; }
    __device__ __forceinline__ double y(int64_t i) const { return wd.Y[i]; }
    __device__ __forceinline__ void fill_rt(int64_t i, double *x, int q) const {
        const int p = q - 4;
        const double g = gf(i), ev = E[i];
        x[0] = 1.0;
        for (int c = 0; c < p; c++) x[1 + c] = wd.cova[i + (int64_t)c * n];
        x[p + 1] = ev; x[p + 2] = g; x[p + 3] = ev * g;
    }
    template <int Q>
    __device__ __forceinline__ void fill(int64_t i, double *x) const { fill_rt(i, x, Q); }
};
template <class GF> struct WaldDesignOf;
template <> struct WaldDesignOf<GFromRow> { using type = DesignWald; };
template <> struct WaldDesignOf<GDense> { using type = DesignWaldDense; };

template <class GF>
struct DesignAfT {
    const double *pcs; int64_t n; int nsel; const int *sel; GF gf;
    __device__ __forceinline__ int q() const { return nsel + 1; }
    __device__ __forceinline__ int trait() const { return 1; }
    __device__ __forceinline__ double y(int64_t i) const { return (rint(gf(i)) != 0.0) ? 1.0 : 0.0; }
    template <int Q>
    __device__ __forceinline__ void fill(int64_t i, double *x) const {
        x[0] = 1.0;
#pragma unroll
        for (int c = 0; c < Q - 1; c++) x[1 + c] = pcs[i + (int64_t)sel[c] * n];
    }
    __device__ __forceinline__ void fill_rt(int64_t i, double *x, int q) const {
        x[0] = 1.0;
        for (int c = 0; c < q - 1; c++) x[1 + c] = pcs[i + (int64_t)sel[c] * n];
    }
};

template <class Design>
__device__ __forceinline__ double wald_pass_any(const Design &ds, int64_t i_lo, int64_t i_hi, int q, bool pass0,
                                                double *tile, BShared &sh, const ClusterRed &red, int *invalid) {
    switch (q) {
        case 2: return pass0 ? wald_pass_reg<2, true>(ds, i_lo, i_hi, tile, sh, red, invalid)
                             : wald_pass_reg<2, false>(ds, i_lo, i_hi, tile, sh, red, invalid);
        case 3: return pass0 ? wald_pass_reg<3, true>(ds, i_lo, i_hi, tile, sh, red, invalid)
                             : wald_pass_reg<3, false>(ds, i_lo, i_hi, tile, sh, red, invalid);
        case 4: return pass0 ? wald_pass_reg<4, true>(ds, i_lo, i_hi, tile, sh, red, invalid)
                             : wald_pass_reg<4, false>(ds, i_lo, i_hi, tile, sh, red, invalid);
        case 5: return pass0 ? wald_pass_reg<5, true>(ds, i_lo, i_hi, tile, sh, red, invalid)
                             : wald_pass_reg<5, false>(ds, i_lo, i_hi, tile, sh, red, invalid);
        case 6: return pass0 ? wald_pass_reg<6, true>(ds, i_lo, i_hi, tile, sh, red, invalid)
                             : wald_pass_reg<6, false>(ds, i_lo, i_hi, tile, sh, red, invalid);
        default: break;
    }
    const int ns = (q * (q + 3) / 2 + kBGroups - 1) / kBGroups;
    switch (ns) {
        case 3: return wald_pass<3>(ds, i_lo, i_hi, q, pass0, tile, sh, red, invalid);
        case 4: return wald_pass<4>(ds, i_lo, i_hi, q, pass0, tile, sh, red, invalid);
        case 5: return wald_pass<5>(ds, i_lo, i_hi, q, pass0, tile, sh, red, invalid);
        case 6: return wald_pass<6>(ds, i_lo, i_hi, q, pass0, tile, sh, red, invalid);
        case 7: return wald_pass<7>(ds, i_lo, i_hi, q, pass0, tile, sh, red, invalid);
        case 8: return wald_pass<8>(ds, i_lo, i_hi, q, pass0, tile, sh, red, invalid);
        case 9: return wald_pass<9>(ds, i_lo, i_hi, q, pass0, tile, sh, red, invalid);
        default: return wald_pass<kBSlots>(ds, i_lo, i_hi, q, pass0, tile, sh, red, invalid);
    }
}

// glm.fit(binomial) IRLS on a design (cluster-collective). On return sh.beta holds the final coefficients and
// *se2_last the last-WLS variance of the last coefficient. Returns 0 ok, 1 rank deficient, 2 invalid.
template <class Design>
__device__ int irls_fit(const Design &ds, int64_t i_lo, int64_t i_hi, double *tile, BShared &sh, const ClusterRed &red,
                        double *se2_last) {
    const int q = ds.q();
    int invalid;
    double devold = wald_pass_any(ds, i_lo, i_hi, q, true, tile, sh, red, &invalid);
    for (int it = 1; it <= 25; it++) {
        wald_solve(sh, q);
        if (!sh.ok) return 1;
        *se2_last = sh.se2;   // covariance of the WLS just solved = the one summary.glm reports on convergence
        double dev = wald_pass_any(ds, i_lo, i_hi, q, false, tile, sh, red, &invalid);
        if (!isfinite(dev) || invalid) return 2;
        if (fabs(dev - devold) / (fabs(dev) + 0.1) < 1e-8) break;
        devold = dev;
    }
    return 0;
}

// Cluster-collective Wald p-value of E:g (ref :636-662): glm.fit IRLS for binary, lm for quantitative.
template <class GFun>
__device__ double wald_eg(const GFun &gf, const WaldData &wd, const double *E, int64_t n, int64_t i_lo, int64_t i_hi,
                          double *tile, BShared &sh, const ClusterRed &red, int *fail) {
    typename WaldDesignOf<GFun>::type ds{wd, E, gf, n};
    const int q = ds.q();
    *fail = 0;
    if (wd.trait == 1) {
        double se2 = 0;
        int rc = irls_fit(ds, i_lo, i_hi, tile, sh, red, &se2);
        if (rc) { *fail = rc; return d_na_real(); }
        const double zst = sh.beta[q - 1] / sqrt(se2);
        return 2 * d_pnorm(-fabs(zst), true);
    }
    int invalid;
    wald_pass_any(ds, i_lo, i_hi, q, true, tile, sh, red, &invalid);
    wald_solve(sh, q);
    if (!sh.ok) { *fail = 1; return d_na_real(); }
    const double il = sh.se2, bl = sh.beta[q - 1];
    double rss = wald_pass_any(ds, i_lo, i_hi, q, false, tile, sh, red, &invalid);
    const double rdf = (double)(n - q);
    const double se = sqrt(il * (rss / rdf));
    return d_pt2(bl / se, rdf);
}

// (branch B of SPAGxE_CCT is the solver farm in bb_farm.cu)


// ===========================================================================
// SPAGxE_Plus (R/SPAGxEPlus_functions_MYZ.R:248-405)
// Step-1 scalars of SPAGxE_Plus_Nullmodel (ref R/SPAGxEPlus_functions_MYZ.R:490-537): the three sparse-GRM quadratic
// forms and R.new = (E - lambda) R with lambda = R_GRM_RE / R_GRM_R.  One CTA, fixed-order sums (runs once per analysis).
#if SPAGXE_IN_TU(1)
__global__ void __launch_bounds__(1024) k_plus_nullmodel(const double *__restrict__ R, const double *__restrict__ E,
                                                         int64_t n, GrmCoo grm, double *__restrict__ Rnew,
                                                         double *__restrict__ out3) {
    __shared__ double scratch[5 * 32];
    __shared__ double s_lambda;
    double v[5] = {0, 0, 0, 0, 0};
    for (long long e = threadIdx.x; e < grm.nnz; e += blockDim.x) {
        const int i = grm.gi[e], j = grm.gj[e];
        const double val = grm.gv[e], ri = R[i], rj = R[j], rei = ri * E[i], rej = rj * E[j];
        v[0] += (val * ri) * rj;          // Cov_R          (:507)
        v[1] += (val * ri) * rej;         // Cov_R_RE_part1 (:515)
        v[2] += (val * rj) * rei;         // Cov_R_RE_part2 (:516)
    }
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double r = R[i], re = r * E[i];
        v[3] += r * r; v[4] += re * r;
    }
    block_sum<5>(v, scratch);
    const double R_GRM_R = v[0] * 2 - v[3];                  // :510
    const double R_GRM_RE = v[1] + v[2] - v[4];              // :523-525
    if (threadIdx.x == 0) { out3[0] = R_GRM_R; out3[1] = R_GRM_RE; s_lambda = R_GRM_RE / R_GRM_R; }
    __syncthreads();
    const double lambda = s_lambda;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) Rnew[i] = (E[i] - lambda) * R[i];   // :530
    __syncthreads();   // the COO pass below gathers R.new written by other threads of this CTA
    double w[2] = {0, 0};
    for (long long e = threadIdx.x; e < grm.nnz; e += blockDim.x) w[0] += (grm.gv[e] * Rnew[grm.gi[e]]) * Rnew[grm.gj[e]];
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const double r = Rnew[i]; w[1] += r * r; }
    block_sum<2>(w, scratch);
    if (threadIdx.x == 0) out3[2] = w[0] * 2 - w[1];         // :537
}
#endif

// K3 for SPAGxE_Plus: V = E*R, so D1/T1 carry sum(g*E*R).  One thread per SNP.  out is M x 8.
#if SPAGXE_IN_TU(1)
__global__ void k_finalize_plus(SnpStats st, int m_chunk, int64_t snp0, int64_t m_total, int64_t n,
                                const VecConst *__restrict__ vcp, PlusConst pc, double epsilon, double cutoff,
                                double min_maf, double *__restrict__ out, SpaItem *spa_list, BItem *b_list,
                                const uint8_t *bed, int64_t pitch, Counters *ctr, int g_model) {
    const int jl = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = jl < m_chunk;
    const VecConst vc = *vcp;
    bool want_spa = false, want_b = false;
    SpaItem item;
    const int64_t j = snp0 + jl;
    if (active) {
        const double NA = d_na_real();
        const int asum = st.asum[jl], nmiss = st.nmiss[jl];
        Prologue pr = make_prologue(asum, nmiss, n);
        double *o = out + j;
        o[0] = pr.maf; o[1 * m_total] = pr.miss_rate;
        if (nmiss == n) {
            for (int c = 0; c < 8; c++) o[c * m_total] = NA;
            o[1 * m_total] = 1.0;
            atomicAdd(&ctr->n_allmissing, 1ULL);
        } else if (pr.maf < min_maf) {
            for (int c = 2; c < 8; c++) o[c * m_total] = NA;
            atomicAdd(&ctr->n_filtered, 1ULL);
        } else {
            double S1 = st.D0[jl] + ((nmiss > 0) ? pr.u * st.T0[jl] : 0.0);
            double S2 = st.D1[jl] + ((nmiss > 0) ? pr.u * st.T1[jl] : 0.0);
            if (pr.flip) { S1 = 2 * vc.sumR - S1; S2 = 2 * vc.sumV - S2; }
            if (g_model != 0) {   // Dom / Rec (ref R/SPAGxEPlus_functions_MYZ.R:286-287)
                const GClass gc = make_gclass(pr, g_model);
                S1 = class_dot(gc, st.D0[jl], st.T0[jl], st.H0[jl], vc.sumR);
                S2 = class_dot(gc, st.D1[jl], st.T1[jl], st.H1[jl], vc.sumV);
            }
            const double gvar = 2 * pr.maf * (1 - pr.maf);
            const double VarS1 = gvar * pc.R_GRM_R;             // ref :300
            const double Z1 = (S1 - 0) / sqrt(VarS1);
            const double pG = d_pnorm(-fabs(Z1), true) * 2;
            o[4 * m_total] = pG; o[5 * m_total] = S1; o[6 * m_total] = VarS1; o[7 * m_total] = Z1;
            if (pG > epsilon) {
                atomicAdd(&ctr->n_branch_a, 1ULL);
                const double lambda = pc.R_GRM_RE / pc.R_GRM_R;  // ref :310
                const double S_GxE = S2 - lambda * S1;
                const double Svar = gvar * pc.Rnew_GRM_Rnew;
                const double z = S_GxE / sqrt(Svar);
                if (fabs(z) < cutoff) {
                    const double pn = d_pnorm(fabs(z), false) * 2;
                    o[2 * m_total] = pn; o[3 * m_total] = pn;
                } else {
                    const double ratio = (vc.sumRn2 * gvar) / Svar;  // sum(R.new^2 * g.var.est) / var (ref :322-323)
                    const double Snew = S_GxE * sqrt(ratio);
                    o[3 * m_total] = d_pnorm(fabs(z), false) + d_pnorm(-fabs(z), true);
                    want_spa = true;
                    item.snp = (int)j; item.flags = 1 | 16;  // 1 column, explicit tails
                    item.maf = pr.maf; item.S = fabs(Snew); item.Smean = -fabs(Snew); item.aux = 0;
                }
            } else {
                atomicAdd(&ctr->n_branch_b, 1ULL);
                want_b = true;
            }
        }
    }
    int pos = list_reserve(&ctr->spa_count, want_spa);
    if (want_spa) spa_list[pos] = item;
    pos = list_reserve(&ctr->b_count, want_b);
    if (want_b) b_list[pos] = BItem{bed + (int64_t)jl * pitch, (long long)j, st.asum[jl], st.nmiss[jl], st.D0[jl], st.T0[jl],
                                    g_model ? st.H0[jl] : 0.0, g_model ? st.nhom[jl] : 0, 0};
}
#endif

// branch B of SPAGxE_Plus (ref :342-400): cluster per SNP; K7 = sparse-GRM quadratic form of R.new0
#if SPAGXE_IN_TU(1)
__global__ void __launch_bounds__(kBThreads) k_branch_b_plus(const BItem *__restrict__ b_list, const int *__restrict__ count_ptr, Counters *ctr,
                                                              int64_t m_total, int64_t n,
                                                              const double *__restrict__ R, const double *__restrict__ E,
                                                              const VecConst *__restrict__ vcp, GrmCoo grm, double cutoff,
                                                              double *__restrict__ out, int g_model) {
    namespace cg = cooperative_groups;
    __shared__ BShared sh;
    cg::cluster_group cl = cg::this_cluster();
    const int cs = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int cluster_id = blockIdx.x / cs, n_clusters = gridDim.x / cs;
    const VecConst vc = *vcp;
    const int count = *count_ptr;   // read on the device: the host never synchronises to size this launch
    ClusterRed red{sh.scratch, sh.slot};
    const int64_t per = ((n + cs - 1) / cs + 3) / 4 * 4;
    const int64_t i_lo = min(n, per * rank), i_hi = min(n, per * (rank + 1));
    const long long eper = (grm.nnz + cs - 1) / cs;
    const long long e_lo = min(grm.nnz, eper * rank), e_hi = min(grm.nnz, eper * (rank + 1));
    for (int idx = cluster_id; idx < count; idx += n_clusters) {
        const BItem bi = b_list[idx];
        const int64_t j = bi.snp;
        const uint8_t *row = bi.row;
        const int asum = bi.asum, nmiss = bi.nmiss;
        const Prologue pr = make_prologue(asum, nmiss, n);
        GFromRow gf;
        gf.row = row;
        const GClass gc = make_gclass(pr, g_model);
        for (int c = 0; c < 4; c++) gf.gval[c] = gc.val[c];
        double S1 = bi.D0 + ((nmiss > 0) ? pr.u * bi.T0 : 0.0);
        if (pr.flip) S1 = 2 * vc.sumR - S1;
        double sum_g = pr.sum_g;
        if (g_model != 0) {
            S1 = class_dot(gc, bi.D0, bi.T0, bi.H0, vc.sumR);
            sum_g = class_sum(gc, make_gcounts(asum, nmiss, bi.nhom, n), 1);
        }
        double v3[3] = {0, 0, 0};
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) { double g = gf(i); v3[0] += g * g; }
        red.sum<3>(v3);
        const double a = (double)n, b = sum_g, d = v3[0];
        const double det = a * d - b * b;
        const double NA = d_na_real();
        double *o = out + j;
        const bool writer = (rank == 0 && threadIdx.x == 0);
        if (det == 0 || !isfinite(det)) {
            if (writer) { o[2 * m_total] = NA; o[3 * m_total] = NA; atomicAdd(&ctr->n_singular, 1ULL); }
            continue;
        }
        const double i00 = d / det, i01 = -b / det, i11 = a / det;
        SrcAdj src;
        src.row = row; src.R = R; src.E = E; src.maf = pr.maf;
        for (int c = 0; c < 4; c++) {
            double wi0 = i00 + gf.gval[c] * i01, wi1 = i01 + gf.gval[c] * i11;
            src.fit[c] = wi0 * vc.sumR + wi1 * S1;
        }
        // S0 = sum(g*E*R0), sum(x), sum(x^2) with x = R.new0 = E*R0
        v3[0] = v3[1] = v3[2] = 0;
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            double r, m, w; src.get(i, r, m, w);
            v3[0] += gf(i) * r; v3[1] += r; v3[2] += r * r;
        }
        red.sum<3>(v3);
        // K7: sum over the COO entries of Value * x[ID1] * x[ID2] (ref :356-358)
        double vq[1] = {0};
        for (long long e = e_lo + threadIdx.x; e < e_hi; e += blockDim.x) {
            double r1, r2, m, w;
            src.get(grm.gi[e], r1, m, w); src.get(grm.gj[e], r2, m, w);
            vq[0] += (grm.gv[e] * r1) * r2;
        }
        red.sum<1>(vq);
        const double gvar = 2 * pr.maf * (1 - pr.maf);
        const double S0 = v3[0];
        const double Smean = 2 * v3[1] * pr.maf;                       // ref :360
        const double Svar = 2 * vq[0] * gvar - v3[2] * gvar;          // ref :361
        const double z = (S0 - Smean) / sqrt(Svar);
        double pspa, pnorm_;
        int iters = 0;
        const bool spa = !(fabs(z) < cutoff);
        if (!spa) { pnorm_ = d_pnorm(fabs(z), false) * 2; pspa = pnorm_; }
        else {
            const double ratio = (v3[2] * gvar) / Svar;
            const double Snew = S0 * sqrt(ratio), Smean_new = Smean * sqrt(ratio);
            const double dlt = fabs(Snew - Smean_new);
            pnorm_ = d_pnorm(fabs(z), false) + d_pnorm(-fabs(z), true);
            pspa = spa_two_tail(src, red, i_lo, i_hi, Smean + dlt, Smean - dlt, CUDART_INF, &iters);
        }
        if (writer) {
            o[2 * m_total] = pspa; o[3 * m_total] = pnorm_;
            if (spa) { atomicAdd(&ctr->n_spa, 1ULL); atomicAdd(&ctr->newton_iters, (unsigned long long)iters); }
        }
    }
    cl.sync();
}
#endif


// ===========================================================================
// SPAGxEmix_CCT (R/SPAGxECCT.R:1013-1269): individual-level allele frequencies from the PCs.
// allele-frequency model of one SNP, kept in shared memory and evaluated on the fly per sample
struct AfModel {
    int kind;               // 0 uniform (MAF), 1 clamp(x'beta / 2), 2 logistic 1 - sqrt(1 - mu)
    int k, nsel;
    int sel[kMaxPC];
    double maf;
    double beta[kMaxPC + 1];
};
__device__ __forceinline__ double af_eval(const AfModel &af, const double *__restrict__ pcs, int64_t n, int64_t i) {
    if (af.kind == 0) return af.maf;
    if (af.kind == 1) {
        double f = af.beta[0];
        for (int c = 0; c < af.k; c++) f += pcs[i + (int64_t)c * n] * af.beta[1 + c];
        double m = f / 2;                              // Fit.lm$fitted.values/2 (ref :1067)
        if (m <= 0) m = 0;                             // ref :1095-1096
        if (m >= 1) m = 1;
        return m;
    }
    double eta = af.beta[0];
    for (int c = 0; c < af.nsel; c++) eta += pcs[i + (int64_t)af.sel[c] * n] * af.beta[1 + c];
    const double tmp = (eta < -30.) ? DBL_EPSILON : ((eta > 30.) ? 1 / DBL_EPSILON : exp(eta));
    const double mu = tmp / (1 + tmp);
    return 1 - sqrt(1 - mu);                           // ref :1089
}
// genotype source of the per-SNP SPAGxEmix code: a packed 2-bit row (values and fitted values of W = [1, g] per class) ...
struct MixGenoRow {
    GFromRow gf; double fit[4];
    __device__ __forceinline__ double g(int64_t i) const { return gf(i); }
    __device__ __forceinline__ void set_fit(double i00, double i01, double i11, double sumR, double S1) {
        for (int c = 0; c < 4; c++) {
            const double wi0 = i00 + gf.gval[c] * i01, wi1 = i01 + gf.gval[c] * i11;
            fit[c] = wi0 * sumR + wi1 * S1;
        }
    }
    __device__ __forceinline__ double fit_at(int64_t i) const { return fit[(gf.row[i >> 2] >> (2 * (i & 3))) & 3]; }
};
// ... or a dense double column (imputed dosages): the same expressions per sample
struct MixGenoDense {
    GDense gf; double i00, i01, i11, sumR, S1;
    __device__ __forceinline__ double g(int64_t i) const { return gf(i); }
    __device__ __forceinline__ void set_fit(double a00, double a01, double a11, double sR, double s1) { i00 = a00; i01 = a01; i11 = a11; sumR = sR; S1 = s1; }
    __device__ __forceinline__ double fit_at(int64_t i) const {
        const double gg = gf(i), wi0 = i00 + gg * i01, wi1 = i01 + gg * i11;
        return wi0 * sumR + wi1 * S1;
    }
};
template <class MG>
struct SrcMixT {  // residual vector (E - lambda) R or E * (R - fitted(g_i)) with per-sample allele frequency
    const AfModel *af; const double *pcs; int64_t n;
    const double *R, *E; MG mg; double lambda; int adj;   // mg by value: the class table stays in registers
    const double *mcache;   // m_i of this SNP, written once by the statistics pass (same thread <-> same samples in every pass)
    __device__ __forceinline__ void get(int64_t i, double &r, double &m, double &w) const {
        if (adj) r = E[i] * (R[i] - mg.fit_at(i));
        else r = (E[i] - lambda) * R[i];
        m = mcache ? mcache[i] : af_eval(*af, pcs, n, i); w = 1.0;
    }
};
using SrcMix = SrcMixT<MixGenoRow>;

// K3 for mix: filtered rows are written here; every other SNP goes to the cluster kernel's list
#if SPAGXE_IN_TU(1)
__global__ void k_finalize_mix(SnpStats st, int m_chunk, int64_t snp0, int64_t m_total, int64_t n, double min_maf,
                               double *__restrict__ out, int *b_list, Counters *ctr) {
    const int jl = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = jl < m_chunk;
    bool want = false;
    if (active) {
        const double NA = d_na_real();
        const int asum = st.asum[jl], nmiss = st.nmiss[jl];
        Prologue pr = make_prologue(asum, nmiss, n);
        double *o = out + snp0 + jl;
        o[0] = pr.maf; o[1 * m_total] = pr.miss_rate;
        if (nmiss == n) {
            for (int c = 0; c < 12; c++) o[c * m_total] = NA;
            o[1 * m_total] = 1.0;
            atomicAdd(&ctr->n_allmissing, 1ULL);
        } else if (pr.maf < min_maf) {
            for (int c = 2; c < 12; c++) o[c * m_total] = NA;
            atomicAdd(&ctr->n_filtered, 1ULL);
        } else want = true;
    }
    int pos = list_reserve(&ctr->b_count, want);
    if (want) b_list[pos] = jl;
}
#endif

struct MixShared { AfModel af; double xg[kMaxPC + 1]; double pv[kMaxPC + 1]; };

// PLUS: SPAGxEmix_CCT_Plus_one_SNP (R/SPAmixGxECCTPlus_functions_MYZ_2024-09-17.R:144-427): the same allele-frequency
// model, every variance through the sparse GRM with the per-SNP weights sqrt(g.var.est_i) (:222-232, :260, :269-299,
// :318-349), the saddlepoint p-value scaled by sqrt(Var.ratio) (GetProb_SPA_G_new_adjust, :435-462).  The branch-B
// Wald test there is a mixed model (lme4::glmer / lmer, :384 / :404) that stays with the caller: columns 4 and 5 of such
// rows are NA.  m_i is exchanged between the CTAs of the cluster through the per-cluster cache in global memory.
__device__ __forceinline__ double mixplus_sq(const double *mc, int64_t i) { const double m = __ldcg(mc + i); return sqrt(2 * m * (1 - m)); }

// where a deferred Wald item of the packed-row kernels comes from (nullptr row: dense genotypes, nothing is deferred)
struct MixDeferSrc { const uint8_t *row; long long snp; int asum, nmiss, jl; };

// One SNP from the allele-frequency model on (ref R/SPAGxECCT.R:1055-1269; PLUS: R/SPAmixGxECCTPlus_functions_MYZ_2024-09-17.R:190-427).
// maf, S1 = sum(g R), S2 = sum(g E R), sum_g describe the processed genotype vector mg.g(.); hb: the batched path's OLS
// coefficients of this SNP or nullptr; o: the SNP's output row.  The kernel parameters are handed through one by one (not
// in a struct) so that they stay reads of the constant bank after inlining.
template <bool PLUS, class MG>
__device__ __forceinline__ void mix_snp_core(Counters *ctr, const int64_t n, const int64_t m_total, const int64_t i_lo, const int64_t i_hi,
                                             const int cs, const int rank, const double *__restrict__ R, const double *__restrict__ E,
                                             const VecConst &vc, const WaldData &wd, const MixData &md, const double epsilon,
                                             const double cutoff, const GrmCoo &grm, double *tile, BShared &sh, MixShared &ms,
                                             double *mcache_in, const SnpStats &st, const ClusterRed &red, const MG &mg, double maf,
                                             double S1, double S2, double sum_g, const double *hb, double *o, const MixDeferSrc &df) {
    const int k = md.k, k1 = md.k + 1;
    const double NA = d_na_real();
    const double MAC = maf * 2 * (double)n;                                   // ref :1058
    double negnum = 0;
    __syncthreads();
    if (threadIdx.x == 0) { ms.af.kind = 0; ms.af.maf = maf; ms.af.k = k; ms.af.nsel = 0; }
    __syncthreads();
    if (!(MAC <= 20)) {
        // ---- lm(g ~ topPCs): X'g, beta = (X'X)^-1 X'g (ref :1066); the batched path has them already ----
        if (hb) {
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int a = 0; a < k1; a++) ms.af.beta[a] = hb[a];
                ms.af.kind = 1;
            }
            __syncthreads();
        } else {
        for (int c0 = 0; c0 < k1; c0 += 4) {
            double v4[4] = {0, 0, 0, 0};
            #pragma unroll 4
            for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
                const double g = mg.g(i);
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (c0 + c < k1) v4[c] += g * ((c0 + c == 0) ? 1.0 : md.pcs[i + (int64_t)(c0 + c - 1) * n]);
            }
            red.sum<4>(v4);
            if (threadIdx.x == 0) for (int c = 0; c < 4; c++) if (c0 + c < k1) ms.xg[c0 + c] = v4[c];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int a = 0; a < k1; a++) {
                double b = 0;
                for (int c = 0; c < k1; c++) b += md.xtx_inv[a + c * k1] * ms.xg[c];
                ms.af.beta[a] = b;
            }
            ms.af.kind = 1;
        }
        __syncthreads();
        }
        // fitted values: negative count and residual sum of squares
        double v2[2] = {0, 0};
        #pragma unroll 4
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            double f = ms.af.beta[0];
            for (int c = 0; c < k; c++) f += md.pcs[i + (int64_t)c * n] * ms.af.beta[1 + c];
            if (f / 2 < 0) v2[0] += 1.0;
            const double e = mg.g(i) - f;
            v2[1] += e * e;
            if (mcache_in) mcache_in[i] = f / 2;   // Fit.lm$fitted.values / 2: the statistics pass clamps it instead of a second walk over the PCs
        }
        red.sum<2>(v2);
        negnum = v2[0];
        const double negratio = negnum / (double)n;
        if (negratio > md.neg_ratio_cut) {                                        // ref :1074
            const double rdf = (double)(n - k1), resvar = v2[1] / rdf;
            __syncthreads();
            if (threadIdx.x < k) {                                                // PC p-values (summary.lm)
                const int a = 1 + threadIdx.x;
                const double se = sqrt(md.xtx_inv[a + a * k1] * resvar);
                ms.pv[a] = d_pt2(ms.af.beta[a] / se, rdf);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int ns = 0;
                for (int a = 1; a <= k; a++) if (ms.pv[a] < md.pc_pcut) ms.af.sel[ns++] = a - 1;
                ms.af.nsel = ns;
                // sum(round(g)) == 0 | sum(2 - round(g)) == 0 from the class counts (ref :1084)
                ms.af.kind = 0;
            }
            __syncthreads();
            const int nsel = ms.af.nsel;
            if (nsel > 0) {
                double vs[2] = {0, 0};
                for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) { double gt = rint(mg.g(i)); vs[0] += gt; vs[1] += 2 - gt; }
                red.sum<2>(vs);
                if (!(vs[0] == 0 || vs[1] == 0)) {
                    DesignAfT<decltype(mg.gf)> ds{md.pcs, n, nsel, ms.af.sel, mg.gf};
                    double se2;
                    int frc = irls_fit(ds, i_lo, i_hi, tile, sh, red, &se2);
                    (void)frc;   // the reference uses Fit$fitted.values whatever glm() warned about
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        for (int c = 0; c <= nsel; c++) ms.af.beta[c] = sh.beta[c];
                        ms.af.kind = 2;
                        if (rank == 0) atomicAdd(&ctr->n_mix_logistic, 1ULL);
                    }
                    __syncthreads();
                }
            }
        }
    }
    __syncthreads();
    // ---- statistics with per-sample allele frequencies (ref :1102-1108) ----
    double v4[4] = {0, 0, 0, 0};
    // m_i is evaluated once per SNP (k loads + k multiply-adds, or an exp and a sqrt) and kept for the later passes
    double *mcache = mcache_in;
    const bool fitted_cached = (mcache != nullptr) && ms.af.kind == 1;   // kind 1 <=> the fitted-values pass above ran and stands
    #pragma unroll 4
    for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
        double m;
        if (fitted_cached) { m = mcache[i]; if (m <= 0) m = 0; if (m >= 1) m = 1; }   // same thread wrote it; af_eval's clamp (ref :1095-1096)
        else m = af_eval(ms.af, md.pcs, n, i);
        const double gv = 2 * m * (1 - m), r = R[i], r2 = r * r;
        if (mcache) mcache[i] = m;
        v4[0] += r * m; v4[1] += r2 * gv; v4[2] += (gv * E[i]) * r2;
    }
    red.sum<4>(v4);
    // edges of the GRM table this CTA sums (PLUS); the barrier inside red.sum has published every CTA's part of mcache
    const int64_t e_per = (grm.nnz + cs - 1) / cs;
    const int64_t e_lo = min((int64_t)grm.nnz, e_per * rank), e_hi = min((int64_t)grm.nnz, e_per * (rank + 1));
    double cov12 = 0;
    if (PLUS) {
        double q2[2] = {0, 0};
        #pragma unroll 4
        for (int64_t e = e_lo + threadIdx.x; e < e_hi; e += blockDim.x) {
            const int a = grm.gi[e], b = grm.gj[e];
            const double xa = R[a] * mixplus_sq(mcache, a), xb = R[b] * mixplus_sq(mcache, b), gvv = grm.gv[e];
            q2[0] += (gvv * xa) * xb;
            q2[1] += (gvv * xa) * (xb * E[b]) + (gvv * xb) * (xa * E[a]);
        }
        red.sum<2>(q2);
        cov12 = q2[1] - v4[2];                                                    // Cov_S1_S2 (ref mixPlus :279-281)
        v4[1] = q2[0] * 2 - v4[1];                                                // VarS1 (ref mixPlus :260)
    }
    const double meanS1 = 2 * v4[0], VarS1 = v4[1];
    const double Z1 = (S1 - meanS1) / sqrt(VarS1);
    const double pG = d_pnorm(-fabs(Z1), true) * 2;
    const bool writer = (rank == 0 && threadIdx.x == 0);
    SrcMixT<MG> src;
    src.af = &ms.af; src.pcs = md.pcs; src.n = n; src.R = R; src.E = E; src.mg = mg; src.mcache = mcache;
    double po[4];
    int iters = 0; bool spa = false;
    if (pG > epsilon) {
        const double lambda = PLUS ? cov12 / VarS1 : v4[2] / v4[1];               // ref :1115; mixPlus :285
        const double S_GxE = S2 - lambda * S1;
        src.lambda = lambda; src.adj = 0;
        double v2[2] = {0, 0};
        #pragma unroll 4
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            double r, m, w; src.get(i, r, m, w);
            v2[0] += r * m; v2[1] += (r * r) * (2 * m * (1 - m));
        }
        red.sum<2>(v2);
        double Svar = v2[1], vscale = 1.0;
        if (PLUS) {   // S_GxE.var through the GRM (ref mixPlus :296), Var.ratio for the saddlepoint (:308-309)
            double q1[1] = {0};
            #pragma unroll 4
            for (int64_t e = e_lo + threadIdx.x; e < e_hi; e += blockDim.x) {
                const int a = grm.gi[e], b = grm.gj[e];
                const double xa = ((E[a] - lambda) * R[a]) * mixplus_sq(mcache, a), xb = ((E[b] - lambda) * R[b]) * mixplus_sq(mcache, b);
                q1[0] += (grm.gv[e] * xa) * xb;
            }
            red.sum<1>(q1);
            Svar = q1[0] * 2 - v2[1];
            vscale = sqrt(v2[1] / Svar);
        }
        const double Smean = 2 * v2[0];
        const double z = (S_GxE - Smean) / sqrt(Svar);
        if (fabs(z) < cutoff) {
            const double pn = d_pnorm(fabs(z), false) * 2;
            po[0] = po[1] = po[2] = po[3] = pn;
        } else {
            spa = true;
            const double ps = spa_two_tail_fused(src, red, i_lo, i_hi, fmax(S_GxE, 2 * Smean - S_GxE), fmin(S_GxE, 2 * Smean - S_GxE), &iters,
                                                 1.7976931348623157e308, vscale);
            po[0] = po[1] = po[2] = ps;
            po[3] = d_pnorm(fabs(z), false) + d_pnorm(-fabs(z), true);
        }
        if (writer) atomicAdd(&ctr->n_branch_a, 1ULL);
    } else {
        if (writer) atomicAdd(&ctr->n_branch_b, 1ULL);
        double v3[3] = {0, 0, 0};
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) { double g = mg.g(i); v3[0] += g * g; }
        red.sum<3>(v3);
        const double a = (double)n, b = sum_g, d = v3[0];
        const double det = a * d - b * b;
        if (det == 0 || !isfinite(det)) {
            if (writer) {
                o[2 * m_total] = NA; o[3 * m_total] = NA; o[4 * m_total] = NA; o[5 * m_total] = NA;
                o[6 * m_total] = pG; o[7 * m_total] = S1; o[8 * m_total] = VarS1; o[9 * m_total] = Z1;
                o[10 * m_total] = negnum; o[11 * m_total] = MAC;
                atomicAdd(&ctr->n_singular, 1ULL);
            }
            return;
        }
        const double i00 = d / det, i01 = -b / det, i11 = a / det;
        src.mg.set_fit(i00, i01, i11, vc.sumR, S1);
        src.adj = 1; src.lambda = 0;
        v3[0] = v3[1] = v3[2] = 0;
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            double r, m, w; src.get(i, r, m, w);
            v3[0] += mg.g(i) * r; v3[1] += r * m; v3[2] += (r * r) * (2 * m * (1 - m));
        }
        red.sum<3>(v3);
        double Svar = v3[2], vscale = 1.0;
        if (PLUS) {   // S_GxE0.var through the GRM (ref mixPlus :343), Var.ratio (:355-356)
            double q1[1] = {0};
            for (int64_t e = e_lo + threadIdx.x; e < e_hi; e += blockDim.x) {
                const int a = grm.gi[e], b = grm.gj[e];
                double ra, rb, m_, w_;
                src.get(a, ra, m_, w_); src.get(b, rb, m_, w_);
                q1[0] += (grm.gv[e] * (ra * mixplus_sq(mcache, a))) * (rb * mixplus_sq(mcache, b));
            }
            red.sum<1>(q1);
            Svar = q1[0] * 2 - v3[2];
            vscale = sqrt(v3[2] / Svar);
        }
        const double S0 = v3[0], Smean = 2 * v3[1];
        const double z = (S0 - Smean) / sqrt(Svar);
        double pspa, pn;
        if (fabs(z) < cutoff) { pn = d_pnorm(fabs(z), false) * 2; pspa = pn; }
        else {
            spa = true;
            pspa = spa_two_tail_fused(src, red, i_lo, i_hi, fmax(S0, 2 * Smean - S0), fmin(S0, 2 * Smean - S0), &iters,
                                      1.7976931348623157e308, vscale);
            pn = d_pnorm(fabs(z), false) + d_pnorm(-fabs(z), true);
        }
        double pw = NA, pc = NA;
        if (!PLUS && (wd.trait == 3 || wd.trait == 4)) {
            // coxph / clm (ref :1187-1193, :1244-1250): the item goes to the Cox / cumulative-link kernel that follows on the
            // stream; it reads pval.spaGxE from column 3 of the row and writes the Wald and CCT columns
            if (writer && df.row) {
                const int pos = atomicAdd(&ctr->spax_count, 1);
                const int gm = md.g_model, jl = df.jl;
                md.wald_items[pos] = BItem{df.row, df.snp, df.asum, df.nmiss, st.D0[jl], st.T0[jl], gm ? st.H0[jl] : 0.0, gm ? st.nhom[jl] : 0, 0};
            }
        } else if (!PLUS) {
            int wfail;
            pw = wald_eg(mg.gf, wd, E, n, i_lo, i_hi, tile, sh, red, &wfail);
            if (isnan(pw)) pw = pspa;                                             // ref :1194, :1213, :1232
            int crc = d_cct2(pspa, pw, &pc);
            if (writer) {
                if (wfail) atomicAdd(&ctr->n_wald_fail, 1ULL);
                if (crc) atomicAdd(&ctr->n_cct_stop, 1ULL);
            }
        }
        po[0] = pspa; po[1] = pw; po[2] = pc; po[3] = pn;
    }
    if (writer) {
        o[2 * m_total] = po[0]; o[3 * m_total] = po[1]; o[4 * m_total] = po[2]; o[5 * m_total] = po[3];
        o[6 * m_total] = pG; o[7 * m_total] = S1; o[8 * m_total] = VarS1; o[9 * m_total] = Z1;
        o[10 * m_total] = negnum; o[11 * m_total] = MAC;
        if (spa) { atomicAdd(&ctr->n_spa, 1ULL); atomicAdd(&ctr->newton_iters, (unsigned long long)iters); }
    }
}

template <bool PLUS>
__device__ __forceinline__ void mix_snp_body(const int *__restrict__ list, Counters *ctr,
                                             const uint8_t *__restrict__ bed, int64_t pitch, SnpStats st,
                                             int64_t snp0, int64_t m_total, int64_t n,
                                             const double *__restrict__ R, const double *__restrict__ E,
                                             const VecConst *__restrict__ vcp, WaldData wd, MixData md,
                                             double epsilon, double cutoff, double *__restrict__ out,
                                             double *__restrict__ mcache_all, const double *__restrict__ hbeta,
                                             int hbeta_ld, const MixHand *__restrict__ hand, GrmCoo grm) {
    namespace cg = cooperative_groups;
    extern __shared__ double tile[];
    __shared__ BShared sh;
    __shared__ MixShared ms;
    cg::cluster_group cl = cg::this_cluster();
    const int cs = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int cluster_id = blockIdx.x / cs, n_clusters = gridDim.x / cs;
    const int count = ctr->b_count;
    const VecConst vc = *vcp;
    ClusterRed red{sh.scratch, sh.slot};
    const int64_t per = ((n + cs - 1) / cs + 3) / 4 * 4;
    const int64_t i_lo = min(n, per * rank), i_hi = min(n, per * (rank + 1));
    const int k = md.k, k1 = md.k + 1;
    const double NA = d_na_real();
    // Dynamic queue over the item list, heavy items first: the first sweep takes the SNPs that still need their allele-frequency
    // model or branch B (several times the cost of the others), the second the saddlepoint-only SNPs the batched path
    // handed over; a static round-robin left the clusters up to a third apart at the end of the kernel.
    __shared__ int s_next;
    (void)n_clusters;
    for (;;) {
        cl.sync();   // every CTA of the cluster is done with the previous item (and has read s_next)
        if (rank == 0 && threadIdx.x == 0) s_next = atomicAdd(&ctr->work_cursor, 1);
        cl.sync();
        const int v = *cl.map_shared_rank(&s_next, 0);
        if (v >= 2 * count) break;
        const int idx = (v < count) ? v : v - count;
        const int entry = list[idx];
        if ((v < count) == (hand != nullptr && (entry & kMixHandFlag) != 0)) continue;
        const int jl = entry & (kMixHandFlag - 1);
        const int64_t j = snp0 + jl;
        const uint8_t *row = bed + (int64_t)jl * pitch;
        const int asum = st.asum[jl], nmiss = st.nmiss[jl];
        const Prologue pr = make_prologue(asum, nmiss, n);
        if (hand && (entry & kMixHandFlag)) {
            // ---- the batched path finished everything but the saddlepoint tails (ref :1127-1140) ----
            const MixHand h = hand[jl];
            __syncthreads();
            if (threadIdx.x == 0) {
                ms.af.kind = 1; ms.af.maf = pr.maf; ms.af.k = k; ms.af.nsel = 0;
                for (int a = 0; a < k1; a++) ms.af.beta[a] = hbeta[(int64_t)jl * hbeta_ld + a];
            }
            __syncthreads();
            double *mc = mcache_all ? mcache_all + (int64_t)cluster_id * n : nullptr;
            if (mc) for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) mc[i] = af_eval(ms.af, md.pcs, n, i);
            SrcMix src;
            src.af = &ms.af; src.pcs = md.pcs; src.n = n; src.R = R; src.E = E; src.mcache = mc;   // adj = 0: mg is not read
            src.lambda = h.lambda; src.adj = 0;
            int iters = 0;
            const double ps = spa_two_tail_fused(src, red, i_lo, i_hi, fmax(h.S_GxE, 2 * h.Smean - h.S_GxE),
                                                 fmin(h.S_GxE, 2 * h.Smean - h.S_GxE), &iters);
            if (rank == 0 && threadIdx.x == 0) {
                double *o = out + j;
                o[2 * m_total] = ps; o[3 * m_total] = ps; o[4 * m_total] = ps;
                o[5 * m_total] = d_pnorm(fabs(h.z), false) + d_pnorm(-fabs(h.z), true);
                o[6 * m_total] = h.pG; o[7 * m_total] = h.S1; o[8 * m_total] = h.VarS1; o[9 * m_total] = h.Z1;
                o[10 * m_total] = h.negnum; o[11 * m_total] = h.MAC;
                atomicAdd(&ctr->n_branch_a, 1ULL); atomicAdd(&ctr->n_spa, 1ULL);
                atomicAdd(&ctr->newton_iters, (unsigned long long)iters);
            }
            continue;
        }
        GFromRow gf;
        gf.row = row;
        const int g_model = md.g_model;
        const GClass gc = make_gclass(pr, g_model);
        for (int c = 0; c < 4; c++) gf.gval[c] = gc.val[c];
        double S1 = st.D0[jl] + ((nmiss > 0) ? pr.u * st.T0[jl] : 0.0);
        double S2 = st.D1[jl] + ((nmiss > 0) ? pr.u * st.T1[jl] : 0.0);
        if (pr.flip) { S1 = 2 * vc.sumR - S1; S2 = 2 * vc.sumV - S2; }
        double sum_g = pr.sum_g;
        if (g_model != 0) {   // Dom / Rec (ref :1046-1047)
            S1 = class_dot(gc, st.D0[jl], st.T0[jl], st.H0[jl], vc.sumR);
            S2 = class_dot(gc, st.D1[jl], st.T1[jl], st.H1[jl], vc.sumV);
            sum_g = class_sum(gc, make_gcounts(asum, nmiss, st.nhom[jl], n), 1);
        }
        MixGenoRow mg; mg.gf = gf;
        mix_snp_core<PLUS>(ctr, n, m_total, i_lo, i_hi, cs, rank, R, E, vc, wd, md, epsilon, cutoff, grm, tile, sh, ms,
                           mcache_all ? mcache_all + (int64_t)cluster_id * n : nullptr, st, red, mg, pr.maf, S1, S2, sum_g,
                           hbeta ? hbeta + (int64_t)jl * hbeta_ld : nullptr, out + j, MixDeferSrc{row, (long long)j, asum, nmiss, jl});
    }
    cl.sync();
}

#if SPAGXE_IN_TU(1)
__global__ void __launch_bounds__(kBThreads) k_mix_snp(const int *__restrict__ list, Counters *ctr,
                                                        const uint8_t *__restrict__ bed, int64_t pitch, SnpStats st,
                                                        int64_t snp0, int64_t m_total, int64_t n,
                                                        const double *__restrict__ R, const double *__restrict__ E,
                                                        const VecConst *__restrict__ vcp, WaldData wd, MixData md,
                                                        double epsilon, double cutoff, double *__restrict__ out,
                                                        double *__restrict__ mcache_all, const double *__restrict__ hbeta,
                                                        int hbeta_ld, const MixHand *__restrict__ hand) {
    mix_snp_body<false>(list, ctr, bed, pitch, st, snp0, m_total, n, R, E, vcp, wd, md, epsilon, cutoff, out, mcache_all, hbeta,
                        hbeta_ld, hand, GrmCoo{nullptr, nullptr, nullptr, 0});
}
#endif
// SPAGxEmix_CCT_Plus: every SNP of the chunk's list through the per-SNP route (the GRM forms are not polynomial in the
// OLS coefficients, so there is no batched path); mcache_all must hold n doubles per cluster
#if SPAGXE_IN_TU(4)
__global__ void __launch_bounds__(kBThreads) k_mixplus_snp(const int *__restrict__ list, Counters *ctr,
                                                            const uint8_t *__restrict__ bed, int64_t pitch, SnpStats st,
                                                            int64_t snp0, int64_t m_total, int64_t n,
                                                            const double *__restrict__ R, const double *__restrict__ E,
                                                            const VecConst *__restrict__ vcp, MixData md, GrmCoo grm,
                                                            double epsilon, double cutoff, double *__restrict__ out,
                                                            double *__restrict__ mcache_all) {
    mix_snp_body<true>(list, ctr, bed, pitch, st, snp0, m_total, n, R, E, vcp, WaldData{nullptr, nullptr, 0, 0}, md, epsilon, cutoff,
                       out, mcache_all, nullptr, 0, nullptr, grm);
}
#endif

// SPAGxE_Plus on imputed (non-integer) dosages (ref R/SPAGxEPlus_functions_MYZ.R:248-405 on a real-valued g): one cluster per
// SNP, both branches; the statistics of k_finalize_plus / k_branch_b_plus with the genotype read as doubles.
#if SPAGXE_IN_TU(3)
__global__ void __launch_bounds__(kBThreads) k_plus_dosage_snp(PlusDosageArgs A) {
    namespace cg = cooperative_groups;
    __shared__ BShared sh;
    cg::cluster_group cl = cg::this_cluster();
    const int cs = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int cluster_id = blockIdx.x / cs, n_clusters = gridDim.x / cs;
    const VecConst vc = *A.vc;
    ClusterRed red{sh.scratch, sh.slot};
    const int64_t n = A.n, m_total = A.m_total;
    const int64_t per = ((n + cs - 1) / cs + 3) / 4 * 4;
    const int64_t i_lo = min(n, per * rank), i_hi = min(n, per * (rank + 1));
    const long long eper = (A.grm.nnz + cs - 1) / cs;
    const long long e_lo = min(A.grm.nnz, eper * rank), e_hi = min(A.grm.nnz, eper * (rank + 1));
    const double NA = d_na_real();
    const bool writer = (rank == 0 && threadIdx.x == 0);
    for (int jl = cluster_id; jl < A.m; jl += n_clusters) {
        const double *col = A.G + (int64_t)jl * A.ld;
        double *o = A.out + A.snp0 + jl;
        double a2[2] = {0, 0};
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            const double g = col[i];
            if (!isnan(g)) { a2[0] += g; a2[1] += 1.0; }
        }
        red.sum<2>(a2);
        const double n_obs = a2[1], n_miss = (double)n - n_obs;
        if (n_obs == 0) {
            if (writer) { for (int c = 0; c < 8; c++) o[c * m_total] = NA; o[m_total] = 1.0; atomicAdd(&A.ctr->n_allmissing, 1ULL); }
            continue;
        }
        double MAF = (a2[0] / n_obs) / 2;
        const double miss_rate = n_miss / (double)n;
        GDense gf; gf.col = col; gf.imp = 2 * MAF; gf.flip = 0; gf.g_model = A.g_model;
        if (MAF > 0.5) { MAF = 1 - MAF; gf.flip = 1; }
        if (writer) { o[0] = MAF; o[m_total] = miss_rate; }
        if (MAF < A.min_maf) {
            if (writer) { for (int c = 2; c < 8; c++) o[c * m_total] = NA; atomicAdd(&A.ctr->n_filtered, 1ULL); }
            continue;
        }
        double s4[4] = {0, 0, 0, 0};
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            const double g = gf(i);
            s4[0] += g * A.R[i]; s4[1] += (g * A.E[i]) * A.R[i]; s4[2] += g; s4[3] += g * g;
        }
        red.sum<4>(s4);
        const double S1 = s4[0], S2 = s4[1], sum_g = s4[2];
        const double gvar = 2 * MAF * (1 - MAF);
        const double VarS1 = gvar * A.pc.R_GRM_R;                                   // ref :300
        const double Z1 = (S1 - 0) / sqrt(VarS1);
        const double pG = d_pnorm(-fabs(Z1), true) * 2;
        if (writer) { o[4 * m_total] = pG; o[5 * m_total] = S1; o[6 * m_total] = VarS1; o[7 * m_total] = Z1; }
        int iters = 0;
        if (pG > A.epsilon) {
            if (writer) atomicAdd(&A.ctr->n_branch_a, 1ULL);
            const double lambda = A.pc.R_GRM_RE / A.pc.R_GRM_R;                     // ref :310
            const double S_GxE = S2 - lambda * S1;
            const double Svar = gvar * A.pc.Rnew_GRM_Rnew;
            const double z = S_GxE / sqrt(Svar);
            if (fabs(z) < A.cutoff) {
                const double pn = d_pnorm(fabs(z), false) * 2;
                if (writer) { o[2 * m_total] = pn; o[3 * m_total] = pn; }
            } else {
                const double ratio = (vc.sumRn2 * gvar) / Svar;                      // ref :322-323
                const double Snew = S_GxE * sqrt(ratio);
                SrcHomo src; src.rv = A.Rn; src.maf = MAF;
                const double ps = spa_two_tail_fused(src, red, i_lo, i_hi, fabs(Snew), -fabs(Snew), &iters);
                if (writer) {
                    o[2 * m_total] = ps; o[3 * m_total] = d_pnorm(fabs(z), false) + d_pnorm(-fabs(z), true);
                    atomicAdd(&A.ctr->n_spa, 1ULL); atomicAdd(&A.ctr->newton_iters, (unsigned long long)iters);
                }
            }
            continue;
        }
        if (writer) atomicAdd(&A.ctr->n_branch_b, 1ULL);
        const double a = (double)n, b = sum_g, d = s4[3];
        const double det = a * d - b * b;
        if (det == 0 || !isfinite(det)) {
            if (writer) { o[2 * m_total] = NA; o[3 * m_total] = NA; atomicAdd(&A.ctr->n_singular, 1ULL); }
            continue;
        }
        const double i00 = d / det, i01 = -b / det, i11 = a / det;
        SrcDenseAdj src;
        src.R = A.R; src.E = A.E; src.gf = gf; src.maf = MAF;
        src.a = i00 * vc.sumR + i01 * S1; src.b = i01 * vc.sumR + i11 * S1;
        double v3[3] = {0, 0, 0};
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            double r, mm, w; src.get(i, r, mm, w);
            v3[0] += gf(i) * r; v3[1] += r; v3[2] += r * r;
        }
        red.sum<3>(v3);
        double vq[1] = {0};                                                        // K7 (ref :356-358)
        for (long long e = e_lo + threadIdx.x; e < e_hi; e += blockDim.x) {
            double r1, r2, mm, w;
            src.get(A.grm.gi[e], r1, mm, w); src.get(A.grm.gj[e], r2, mm, w);
            vq[0] += (A.grm.gv[e] * r1) * r2;
        }
        red.sum<1>(vq);
        const double S0 = v3[0];
        const double Smean = 2 * v3[1] * MAF;                                      // ref :360
        const double Svar = 2 * vq[0] * gvar - v3[2] * gvar;                       // ref :361
        const double z = (S0 - Smean) / sqrt(Svar);
        double pspa, pnorm_;
        const bool spa = !(fabs(z) < A.cutoff);
        if (!spa) { pnorm_ = d_pnorm(fabs(z), false) * 2; pspa = pnorm_; }
        else {
            const double ratio = (v3[2] * gvar) / Svar;
            const double Snew = S0 * sqrt(ratio), Smean_new = Smean * sqrt(ratio);
            const double dlt = fabs(Snew - Smean_new);
            pnorm_ = d_pnorm(fabs(z), false) + d_pnorm(-fabs(z), true);
            pspa = spa_two_tail_fused(src, red, i_lo, i_hi, Smean + dlt, Smean - dlt, &iters);
        }
        if (writer) {
            o[2 * m_total] = pspa; o[3 * m_total] = pnorm_;
            if (spa) { atomicAdd(&A.ctr->n_spa, 1ULL); atomicAdd(&A.ctr->newton_iters, (unsigned long long)iters); }
        }
    }
    cl.sync();
}
#endif

// SPAGxEmix_CCT / SPAGxEmix_CCT_Plus on imputed (non-integer) dosages: a dense double column per SNP, the prologue of
// ref :1030-1051 computed here (k_cct_dosage_snp's layout), then the same per-SNP code as the packed-row kernels.
template <bool PLUS>
__device__ __forceinline__ void mix_dosage_body(const MixDosageArgs &A) {
    namespace cg = cooperative_groups;
    extern __shared__ double tile[];
    __shared__ BShared sh;
    __shared__ MixShared ms;
    cg::cluster_group cl = cg::this_cluster();
    const int cs = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int cluster_id = blockIdx.x / cs, n_clusters = gridDim.x / cs;
    const VecConst vc = *A.vc;
    ClusterRed red{sh.scratch, sh.slot};
    const int64_t n = A.n, m_total = A.m_total;
    const int64_t per = ((n + cs - 1) / cs + 3) / 4 * 4;
    const int64_t i_lo = min(n, per * rank), i_hi = min(n, per * (rank + 1));
    const double NA = d_na_real();
    const bool writer = (rank == 0 && threadIdx.x == 0);
    const SnpStats no_stats{};
    for (int jl = cluster_id; jl < A.m; jl += n_clusters) {
        cl.sync();   // every CTA is done with the previous SNP's shared state and allele-frequency cache
        const double *col = A.G + (int64_t)jl * A.ld;
        double *o = A.out + A.snp0 + jl;
        double a2[2] = {0, 0};
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            const double g = col[i];
            if (!isnan(g)) { a2[0] += g; a2[1] += 1.0; }
        }
        red.sum<2>(a2);
        const double n_obs = a2[1], n_miss = (double)n - n_obs;
        if (n_obs == 0) {
            if (writer) { for (int c = 0; c < 12; c++) o[c * m_total] = NA; o[m_total] = 1.0; atomicAdd(&A.ctr->n_allmissing, 1ULL); }
            continue;
        }
        double MAF = (a2[0] / n_obs) / 2;
        const double miss_rate = n_miss / (double)n;
        MixGenoDense mg;
        mg.gf.col = col; mg.gf.imp = 2 * MAF; mg.gf.flip = 0; mg.gf.g_model = A.md.g_model;
        if (MAF > 0.5) { MAF = 1 - MAF; mg.gf.flip = 1; }
        if (writer) { o[0] = MAF; o[m_total] = miss_rate; }
        if (MAF < A.min_maf) {
            if (writer) { for (int c = 2; c < 12; c++) o[c * m_total] = NA; atomicAdd(&A.ctr->n_filtered, 1ULL); }
            continue;
        }
        double s3[3] = {0, 0, 0};
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            const double g = mg.g(i);
            s3[0] += g * A.R[i]; s3[1] += (g * A.E[i]) * A.R[i]; s3[2] += g;
        }
        red.sum<3>(s3);
        mix_snp_core<PLUS>(A.ctr, n, m_total, i_lo, i_hi, cs, rank, A.R, A.E, vc, A.wd, A.md, A.epsilon, A.cutoff, A.grm, tile, sh, ms,
                           A.mcache_all + (int64_t)cluster_id * n, no_stats, red, mg, MAF, s3[0], s3[1], s3[2], nullptr, o,
                           MixDeferSrc{nullptr, 0, 0, 0, 0});
    }
    cl.sync();
}
#if SPAGXE_IN_TU(3)
__global__ void __launch_bounds__(kBThreads) k_mix_dosage_snp(MixDosageArgs A) { mix_dosage_body<false>(A); }
#endif
#if SPAGXE_IN_TU(3)
__global__ void __launch_bounds__(kBThreads) k_mixplus_dosage_snp(MixDosageArgs A) { mix_dosage_body<true>(A); }
#endif

// ===========================================================================
// SPAGxEmixCCT_localance (R/SPAGxECCT.R:1401-1647): ancestry-specific G x E test with local ancestry.  One cluster per
// SNP; g (ancestry-specific genotype), haplo_numVec and the Cova.haplo.mtx.list columns are per-SNP 2-bit rows.
// G_i ~ Binomial(h_i, MAF): the CGF of a sample is h_i / 2 times the CGF of the two-allele case, so the saddlepoint
// code above runs with the weight w = h_i / 2 (exact for h in {0, 1, 2}).
__device__ __forceinline__ int code_at(const uint8_t *row, int64_t i) { return (row[i >> 2] >> (2 * (i & 3))) & 3; }
__device__ __forceinline__ double hap_of_code(int c) { return (c == 0) ? 2.0 : ((c == 2) ? 1.0 : 0.0); }

struct SrcLocal {  // residual vector (E - lambda) R or E * (R - fit[class]); scalar allele frequency; weight h_i / 2
    const double *R, *E; const uint8_t *row_g, *row_h; double fit[4]; double lambda, maf; int adj;
    __device__ __forceinline__ void get(int64_t i, double &r, double &m, double &w) const {
        if (adj) r = E[i] * (R[i] - fit[code_at(row_g, i)]);
        else r = (E[i] - lambda) * R[i];
        m = maf; w = 0.5 * hap_of_code(code_at(row_h, i));
    }
};

// Wald design of the local-ancestry test (ref :1593, :1607): [1, Cova.haplo (L), Cova (p), haplo_numVec (quantitative only),
// E, g, E:g] restricted to the columns lm() / glm() keep (map[c] = column of the full design)
struct DesignLocal {
    WaldData wd; const double *E; GFromRow gf; int64_t n;
    const uint8_t *row_h; const uint8_t *ch[kMaxLocalCov]; int L, has_h, nq; const int *map;
    __device__ __forceinline__ int q() const { return nq; }
    __device__ __forceinline__ int q_full() const { return 1 + L + wd.p + has_h + 3; }
    __device__ __forceinline__ int trait() const { return wd.trait; }
    __device__ __forceinline__ double y(int64_t i) const { return wd.Y[i]; }
    __device__ __forceinline__ double col(int id, int64_t i) const {
        if (id == 0) return 1.0;
        id -= 1;
        if (id < L) return hap_of_code(code_at(ch[id], i));
        id -= L;
        if (id < wd.p) return wd.cova[i + (int64_t)id * n];
        id -= wd.p;
        if (has_h) { if (id == 0) return hap_of_code(code_at(row_h, i)); id -= 1; }
        if (id == 0) return E[i];
        if (id == 1) return gf(i);
        return E[i] * gf(i);
    }
    __device__ __forceinline__ void fill_rt(int64_t i, double *x, int q) const {
        for (int c = 0; c < q; c++) x[c] = col(map ? map[c] : c, i);
    }
    template <int Q>
    __device__ __forceinline__ void fill(int64_t i, double *x) const {
#pragma unroll
        for (int c = 0; c < Q; c++) x[c] = col(map ? map[c] : c, i);
    }
};

struct LocalShared { int map[kMaxQ]; int nq; int last_kept; };

// which columns of the design survive lm()'s / glm()'s rank check: Cholesky of the equilibrated Gram matrix, a pivot
// below 1e-13 (remaining norm < 3e-7 of the column norm; dqrdc2 uses 1e-7) marks the column as aliased
static __device__ void local_alias_map(const double (*NE)[kMaxQ], int q0, LocalShared &ls) {
    double A[kMaxQ][kMaxQ], sc[kMaxQ];
    int kept[kMaxQ], nk = 0;
    for (int i = 0; i < q0; i++) sc[i] = (NE[i][i] > 0) ? 1.0 / sqrt(NE[i][i]) : 0.0;
    for (int j = 0; j < q0; j++) {
        // row of L for candidate column j against the kept columns
        double lrow[kMaxQ];
        double d = (sc[j] > 0) ? 1.0 : 0.0;
        for (int a = 0; a < nk; a++) {
            const int ca = kept[a];
            double sacc = NE[j][ca] * sc[j] * sc[ca];
            for (int b = 0; b < a; b++) sacc -= lrow[b] * A[a][b];
            lrow[a] = sacc / A[a][a];
            d -= lrow[a] * lrow[a];
        }
        if (!(d > 1e-13)) continue;
        for (int a = 0; a < nk; a++) A[nk][a] = lrow[a];
        A[nk][nk] = sqrt(d);
        kept[nk++] = j;
    }
    for (int a = 0; a < nk; a++) ls.map[a] = kept[a];
    ls.nq = nk;
    ls.last_kept = (nk > 0 && kept[nk - 1] == q0 - 1) ? 1 : 0;
}

#if SPAGXE_IN_TU(2)
__global__ void __launch_bounds__(kBThreads) k_localance_snp(LocalArgs A) {
    namespace cg = cooperative_groups;
    extern __shared__ double tile[];
    __shared__ BShared sh;
    __shared__ LocalShared ls;
    cg::cluster_group cl = cg::this_cluster();
    const int cs = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int cluster_id = blockIdx.x / cs, n_clusters = gridDim.x / cs;
    const VecConst vc = *A.vc;
    ClusterRed red{sh.scratch, sh.slot};
    const int64_t n = A.n, m_total = A.m_total;
    const int64_t per = ((n + cs - 1) / cs + 3) / 4 * 4;
    const int64_t i_lo = min(n, per * rank), i_hi = min(n, per * (rank + 1));
    const double NA = d_na_real();
    const bool writer = (rank == 0 && threadIdx.x == 0);
    for (int jl = cluster_id; jl < A.m; jl += n_clusters) {
        const uint8_t *row_g = A.g_rows + (int64_t)jl * A.pitch, *row_h = A.h_rows + (int64_t)jl * A.pitch;
        double *o = A.out + A.snp0 + jl;
        // ---- class counts and the SNP-specific sums with h (ref :1514-1543) ----
        double c4[4] = {0, 0, 0, 0}, t4[4] = {0, 0, 0, 0}, u2[2] = {0, 0};
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            const int cg_ = code_at(row_g, i), ch_ = code_at(row_h, i);
            const double r = A.R[i], h = hap_of_code(ch_), r2 = r * r;
            c4[0] += (cg_ == 0); c4[1] += (cg_ == 2); c4[2] += (cg_ == 1) + (ch_ == 1); c4[3] += h;
            t4[0] += (cg_ == 0) ? r : 0.0; t4[1] += (cg_ == 2) ? r : 0.0; t4[2] += h * r; t4[3] += h * r2;
            u2[0] += (h * A.E[i]) * r2;
        }
        red.sum<4>(c4); red.sum<4>(t4); red.sum<2>(u2);
        const double n_hom = c4[0], n_het = c4[1], n_hz = (double)n - n_hom - n_het, sum_h = c4[3];
        if (c4[2] > 0) {   // an NA in g or haplo_numVec: `if (MAF > 0.5)` stops the reference
            if (writer) { for (int c = 0; c < 10; c++) o[c * m_total] = NA; atomicAdd(&A.ctr->n_allmissing, 1ULL); }
            continue;
        }
        double MAF = (n_het + 2 * n_hom) / sum_h;
        double gv[4] = {2.0, 0.0, 1.0, 0.0};                           // by PLINK code: 00 -> 2, 10 -> 1, 11 -> 0
        if (MAF > 0.5) { MAF = 1 - MAF; for (int c = 0; c < 4; c++) gv[c] = 2 - gv[c]; }
        if (A.g_model == 1) for (int c = 0; c < 4; c++) gv[c] = (gv[c] >= 1) ? 1.0 : 0.0;
        if (A.g_model == 2) for (int c = 0; c < 4; c++) gv[c] = (gv[c] <= 1) ? 0.0 : 1.0;
        if (!(MAF >= A.min_maf)) {
            if (writer) { o[0] = MAF; o[m_total] = 0.0; for (int c = 2; c < 10; c++) o[c * m_total] = NA; atomicAdd(&A.ctr->n_filtered, 1ULL); }
            continue;
        }
        const double T_hz = vc.sumR - t4[0] - t4[1];
        const double S1 = gv[0] * t4[0] + gv[2] * t4[1] + gv[3] * T_hz;
        const double S1mean = MAF * t4[2], VarS1 = (MAF * (1 - MAF)) * t4[3];
        const double Z1 = (S1 - S1mean) / sqrt(VarS1);
        const double pG = d_pnorm(-fabs(Z1), true) * 2;
        // the inner SPAmix_localance_one_SNP re-derives the frequency from the processed g and may flip again (ref :2334-2350)
        const double sum_g = gv[0] * n_hom + gv[2] * n_het + gv[3] * n_hz;
        double MAFin = sum_g / sum_h, gin[4];
        for (int c = 0; c < 4; c++) gin[c] = gv[c];
        if (MAFin > 0.5) { MAFin = 1 - MAFin; for (int c = 0; c < 4; c++) gin[c] = 2 - gin[c]; }
        const bool inner_filtered = !(MAFin >= A.min_maf);
        SrcLocal src;
        src.R = A.R; src.E = A.E; src.row_g = row_g; src.row_h = row_h; src.maf = MAFin;
        GFromRow gf; gf.row = row_g;
        for (int c = 0; c < 4; c++) gf.gval[c] = gv[c];
        double po[4];
        int iters = 0; bool spa = false;
        const bool branch_a = pG > A.epsilon;
        bool singular = false;
        if (branch_a) {
            src.lambda = u2[0] / t4[3]; src.adj = 0;                              // ref :1551
        } else {
            const double a = (double)n, b = sum_g, d = gv[0] * gv[0] * n_hom + gv[2] * gv[2] * n_het + gv[3] * gv[3] * n_hz;
            const double det = a * d - b * b;
            if (det == 0 || !isfinite(det)) singular = true;
            else {
                const double i00 = d / det, i01 = -b / det, i11 = a / det;
                for (int c = 0; c < 4; c++) {
                    const double wi0 = i00 + gv[c] * i01, wi1 = i01 + gv[c] * i11;
                    src.fit[c] = wi0 * vc.sumR + wi1 * S1;
                }
                src.adj = 1; src.lambda = 0;
            }
        }
        if (singular) {
            if (writer) {
                o[0] = MAF; o[m_total] = 0.0;
                o[2 * m_total] = NA; o[3 * m_total] = NA; o[4 * m_total] = NA; o[5 * m_total] = NA;
                o[6 * m_total] = pG; o[7 * m_total] = S1; o[8 * m_total] = VarS1; o[9 * m_total] = Z1;
                atomicAdd(&A.ctr->n_branch_b, 1ULL); atomicAdd(&A.ctr->n_singular, 1ULL);
            }
            continue;
        }
        double pspa = NA, pn = NA;
        if (!inner_filtered) {
            double v3[3] = {0, 0, 0};
            for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
                double r, mm, w; src.get(i, r, mm, w);
                const double h = 2 * w;
                v3[0] += gin[code_at(row_g, i)] * r; v3[1] += h * r; v3[2] += h * (r * r);
            }
            red.sum<3>(v3);
            const double S = v3[0], Smean = MAFin * v3[1], Svar = (MAFin * (1 - MAFin)) * v3[2];
            const double z = (S - Smean) / sqrt(Svar);
            if (fabs(z) < 2.0) { pn = d_pnorm(fabs(z), false) * 2; pspa = pn; }   // Cutoff is not forwarded (ref :1559, :1579)
            else {
                spa = true;
                pspa = spa_two_tail_fused(src, red, i_lo, i_hi, fmax(S, 2 * Smean - S), fmin(S, 2 * Smean - S), &iters);
                pn = d_pnorm(fabs(z), false) + d_pnorm(-fabs(z), true);
            }
        }
        if (branch_a) {
            po[0] = po[1] = po[2] = pspa; po[3] = pn;
            if (writer) atomicAdd(&A.ctr->n_branch_a, 1ULL);
        } else {
            if (writer) atomicAdd(&A.ctr->n_branch_b, 1ULL);
            // ---- Wald test: which columns does the fit keep, then glm.fit / lm on them ----
            DesignLocal ds;
            ds.wd = A.wd; ds.E = A.E; ds.gf = gf; ds.n = n; ds.row_h = row_h; ds.L = A.n_ch; ds.has_h = (A.wd.trait == 2) ? 1 : 0;
            for (int l = 0; l < kMaxLocalCov; l++) ds.ch[l] = (l < A.n_ch) ? A.ch_rows[l] + (int64_t)jl * A.pitch : nullptr;
            ds.map = nullptr; ds.nq = ds.q_full();
            const int q0 = ds.nq;
            int invalid;
            {   // X'X of the full design: the lm-type first pass (unit weights)
                DesignLocal dl = ds; dl.wd.trait = 2;
                wald_pass_any(dl, i_lo, i_hi, q0, true, tile, sh, red, &invalid);
            }
            __syncthreads();
            if (threadIdx.x == 0) local_alias_map(sh.NE, q0, ls);
            __syncthreads();
            double pw = NA; int wfail = 0;
            if (!ls.last_kept) wfail = 1;                                         // "E:g" itself aliased: the reference stops
            else {
                ds.map = ls.map; ds.nq = ls.nq;
                const int q = ds.nq;
                if (A.wd.trait == 1) {
                    double se2 = 0;
                    const int rc = irls_fit(ds, i_lo, i_hi, tile, sh, red, &se2);
                    if (rc) wfail = rc;
                    else pw = 2 * d_pnorm(-fabs(sh.beta[q - 1] / sqrt(se2)), true);
                } else {
                    wald_pass_any(ds, i_lo, i_hi, q, true, tile, sh, red, &invalid);
                    wald_solve(sh, q);
                    if (!sh.ok) wfail = 1;
                    else {
                        const double il = sh.se2, bl = sh.beta[q - 1];
                        const double rss = wald_pass_any(ds, i_lo, i_hi, q, false, tile, sh, red, &invalid);
                        const double rdf = (double)(n - q);
                        pw = d_pt2(bl / sqrt(il * (rss / rdf)), rdf);
                    }
                }
            }
            double pc = NA;
            const int crc = d_cct2(pspa, pw, &pc);
            if (writer) {
                if (wfail) atomicAdd(&A.ctr->n_wald_fail, 1ULL);
                if (crc) atomicAdd(&A.ctr->n_cct_stop, 1ULL);
            }
            po[0] = pspa; po[1] = pw; po[2] = crc ? NA : pc; po[3] = pn;
        }
        if (writer) {
            o[0] = MAF; o[m_total] = 0.0;
            o[2 * m_total] = po[0]; o[3 * m_total] = po[1]; o[4 * m_total] = po[2]; o[5 * m_total] = po[3];
            o[6 * m_total] = pG; o[7 * m_total] = S1; o[8 * m_total] = VarS1; o[9 * m_total] = Z1;
            if (spa) { atomicAdd(&A.ctr->n_spa, 1ULL); atomicAdd(&A.ctr->newton_iters, (unsigned long long)iters); }
        }
    }
    cl.sync();
}
#endif

// ===========================================================================
// SPAGxE_CCT_one_SNP (R/SPAGxECCT.R:543-683) on a real-valued genotype vector (imputed dosages): the 2-bit path above
// needs g in {0, 1, 2, NA}; a matrix with other values goes through this kernel -- one cluster per SNP, g read as
// doubles (8 bytes per sample instead of 0.25: a compatibility path, HBM / PCIe bound by its input).

#if SPAGXE_IN_TU(4)
__global__ void __launch_bounds__(kBThreads) k_cct_dosage_snp(DosageArgs A) {
    namespace cg = cooperative_groups;
    extern __shared__ double tile[];
    __shared__ BShared sh;
    cg::cluster_group cl = cg::this_cluster();
    const int cs = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int cluster_id = blockIdx.x / cs, n_clusters = gridDim.x / cs;
    const VecConst vc = *A.vc;
    ClusterRed red{sh.scratch, sh.slot};
    const int64_t n = A.n, m_total = A.m_total;
    const int64_t per = ((n + cs - 1) / cs + 3) / 4 * 4;
    const int64_t i_lo = min(n, per * rank), i_hi = min(n, per * (rank + 1));
    const double NA = d_na_real();
    const bool writer = (rank == 0 && threadIdx.x == 0);
    for (int jl = cluster_id; jl < A.m; jl += n_clusters) {
        const double *col = A.G + (int64_t)jl * A.ld;
        double *o = A.out + A.snp0 + jl;
        // ---- mean(g, na.rm = T), missing rate (ref :557-560) ----
        double a2[2] = {0, 0};
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            const double g = col[i];
            if (!isnan(g)) { a2[0] += g; a2[1] += 1.0; }
        }
        red.sum<2>(a2);
        const double n_obs = a2[1], n_miss = (double)n - n_obs;
        if (n_obs == 0) {
            if (writer) { for (int c = 0; c < 10; c++) o[c * m_total] = NA; o[m_total] = 1.0; atomicAdd(&A.ctr->n_allmissing, 1ULL); }
            continue;
        }
        double MAF = (a2[0] / n_obs) / 2;
        const double miss_rate = n_miss / (double)n;
        GDense gf; gf.col = col; gf.imp = 2 * MAF; gf.flip = 0; gf.g_model = A.g_model;
        if (MAF > 0.5) { MAF = 1 - MAF; gf.flip = 1; }
        if (MAF < A.min_maf) {
            if (writer) { o[0] = MAF; o[m_total] = miss_rate; for (int c = 2; c < 10; c++) o[c * m_total] = NA; atomicAdd(&A.ctr->n_filtered, 1ULL); }
            continue;
        }
        // ---- sums of the processed vector: S1 = sum(g R), sum(g R.new), sum(g), sum(g^2) ----
        double s4[4] = {0, 0, 0, 0};
        for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
            const double g = gf(i);
            s4[0] += g * A.R[i]; s4[1] += g * A.Rn[i]; s4[2] += g; s4[3] += g * g;
        }
        red.sum<4>(s4);
        const double S1 = s4[0], sum_g = s4[2];
        const double VarS1 = vc.sumR2 * (2 * MAF * (1 - MAF));
        const double Z1 = S1 / sqrt(VarS1);
        const double pG = d_pnorm(-fabs(Z1), true) * 2;
        double po[4];
        int iters = 0; bool spa = false;
        if (pG > A.epsilon) {
            if (writer) atomicAdd(&A.ctr->n_branch_a, 1ULL);
            double mafi, S, Smean, z;
            const int r = homo_scalar(sum_g, n, s4[1], vc.sumRn, vc.sumRn2, 0.00001, 2.0, &mafi, &S, &Smean, &z);   // ref :598
            if (r == 2) { po[0] = po[1] = po[2] = po[3] = NA; }
            else if (r == 0) { const double pn = d_pnorm(fabs(z), false) * 2; po[0] = po[1] = po[2] = po[3] = pn; }
            else {
                spa = true;
                SrcHomo src; src.rv = A.Rn; src.maf = mafi;
                const double ps = spa_two_tail_fused(src, red, i_lo, i_hi, fmax(S, 2 * Smean - S), fmin(S, 2 * Smean - S), &iters);
                po[0] = po[1] = po[2] = ps;
                po[3] = d_pnorm(fabs(z), false) + d_pnorm(-fabs(z), true);
            }
        } else {
            if (writer) atomicAdd(&A.ctr->n_branch_b, 1ULL);
            const double a = (double)n, b = sum_g, d = s4[3];
            const double det = a * d - b * b;
            if (det == 0 || !isfinite(det)) {
                if (writer) {
                    o[0] = MAF; o[m_total] = miss_rate;
                    o[2 * m_total] = NA; o[3 * m_total] = NA; o[4 * m_total] = NA; o[5 * m_total] = NA;
                    o[6 * m_total] = pG; o[7 * m_total] = S1; o[8 * m_total] = VarS1; o[9 * m_total] = Z1;
                    atomicAdd(&A.ctr->n_singular, 1ULL);
                }
                continue;
            }
            const double i00 = d / det, i01 = -b / det, i11 = a / det;
            SrcDenseAdj src;
            src.R = A.R; src.E = A.E; src.gf = gf;
            src.a = i00 * vc.sumR + i01 * S1; src.b = i01 * vc.sumR + i11 * S1;
            double v3[3] = {0, 0, 0};
            for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
                double r, mm, w; src.get(i, r, mm, w);
                v3[0] += gf(i) * r; v3[1] += r; v3[2] += r * r;
            }
            red.sum<3>(v3);
            double mafi, S, Smean, z, pspa, pn;
            const int r = homo_scalar(sum_g, n, v3[0], v3[1], v3[2], A.min_maf, 2.0, &mafi, &S, &Smean, &z);   // ref :616
            if (r == 2) { pspa = NA; pn = NA; }
            else if (r == 0) { pn = d_pnorm(fabs(z), false) * 2; pspa = pn; }
            else {
                spa = true; src.maf = mafi;
                pspa = spa_two_tail_fused(src, red, i_lo, i_hi, fmax(S, 2 * Smean - S), fmin(S, 2 * Smean - S), &iters);
                pn = d_pnorm(fabs(z), false) + d_pnorm(-fabs(z), true);
            }
            DesignWaldDense ds{A.wd, A.E, gf, n};
            const int q = ds.q();
            double pw = NA; int wfail = 0, invalid;
            if (A.wd.trait == 1) {
                double se2 = 0;
                const int rc = irls_fit(ds, i_lo, i_hi, tile, sh, red, &se2);
                if (rc) wfail = rc; else pw = 2 * d_pnorm(-fabs(sh.beta[q - 1] / sqrt(se2)), true);
            } else {
                wald_pass_any(ds, i_lo, i_hi, q, true, tile, sh, red, &invalid);
                wald_solve(sh, q);
                if (!sh.ok) wfail = 1;
                else {
                    const double il = sh.se2, bl = sh.beta[q - 1];
                    const double rss = wald_pass_any(ds, i_lo, i_hi, q, false, tile, sh, red, &invalid);
                    const double rdf = (double)(n - q);
                    pw = d_pt2(bl / sqrt(il * (rss / rdf)), rdf);
                }
            }
            double pc = NA;
            const int crc = d_cct2(pspa, pw, &pc);
            if (writer) {
                if (wfail) atomicAdd(&A.ctr->n_wald_fail, 1ULL);
                if (crc) atomicAdd(&A.ctr->n_cct_stop, 1ULL);
            }
            po[0] = pspa; po[1] = pw; po[2] = crc ? NA : pc; po[3] = pn;
        }
        if (writer) {
            o[0] = MAF; o[m_total] = miss_rate;
            o[2 * m_total] = po[0]; o[3 * m_total] = po[1]; o[4 * m_total] = po[2]; o[5 * m_total] = po[3];
            o[6 * m_total] = pG; o[7 * m_total] = S1; o[8 * m_total] = VarS1; o[9 * m_total] = Z1;
            if (spa) { atomicAdd(&A.ctr->n_spa, 1ULL); atomicAdd(&A.ctr->newton_iters, (unsigned long long)iters); }
        }
    }
    cl.sync();
}
#endif

}  // namespace spagxe
