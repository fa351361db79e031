// mix_batch.cu -- batched common path of SPAGxEmix_CCT (ref R/SPAGxECCT.R:1055-1147).
//
// k_mix_snp (kernels.cuh) walks one SNP at a time through a thread-block cluster and re-reads the N x k PC matrix
// from L2 four times per SNP, which makes the loop L2-bandwidth bound (21 k SNPs/s at N = 200,000, k = 10).
// Most SNPs, however, take the same short route through the reference:
//     MAC > 20, lm(g ~ topPCs), less than MAF.est.negative.ratio.cutoff negative fitted frequencies -> clamp,
//     marginal p_G > epsilon (branch A), |z_GxE| < Cutoff (normal approximation)
// and for that route everything is a sum over samples of products  f(x_i' beta_j) * (per-sample constant),
// i.e. a skinny contraction shared by all SNPs of a tile.  Here a thread owns ONE SNP and the sample rows
// [1, PCs, R, R^2, E R^2, E R, E^2 R^2] are staged through shared memory once per 128 SNPs:
//     F1  k_mix_tile<K, 0>   X'g partial sums per (sample split, SNP)            (ref :1066, lm's X'y)
//     F2  k_mix_beta         beta = (X'X)^-1 X'g in fixed order
//     F3  k_mix_tile<K, 1>   negative count and the five weighted sums of m_i = clamp(x_i' beta / 2)   (ref :1067-1120)
//     F4  k_mix_finish       scalar epilogue; writes the output row of a common-path SNP, appends every other SNP
//                            (MAC <= 20, too many negative fits, branch B, SPA) to the list k_mix_snp then serves.
// Sums are accumulated per (split, SNP) and added in split order: results do not depend on the launch shape.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <string>

#include "mix_batch.h"
#include "dmath.cuh"
#include "snp_common.cuh"

namespace spagxe {

constexpr int kMbThreads = 128;   // SNPs per CTA (one thread each)
constexpr int kMbTile = 128;      // samples per staged tile
constexpr int kMbStats = 6;       // neg, sum r m, sum r^2 gv, sum gv E r^2, sum E r m, sum E^2 r^2 gv

__global__ void k_mix_mark(const int *__restrict__ list, const int *__restrict__ count, uint8_t *__restrict__ want) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < *count) want[list[t]] = 1;
}

// One CTA = 128 SNPs x one sample split.  MODE 0: acc[c] = sum_i g_ij x_ic.  MODE 1: statistics of m_ij.
template <int KPAD, int MODE>
__global__ void __launch_bounds__(kMbThreads)
k_mix_tile(const uint8_t *__restrict__ bed, int64_t pitch, int mc, int64_t n, int k, const double *__restrict__ pcs,
           const double *__restrict__ R, const double *__restrict__ E, const double *__restrict__ beta,
           const int *__restrict__ asum, const int *__restrict__ nmiss, double *__restrict__ partial, int64_t per_split, int g_model,
           const VecConst *__restrict__ vcp) {
    // E enters the lambda-expanded statistics centred by the SNP-invariant c = sum(E R^2) / sum(R^2): with an uncentred E
    // (a birth year, a large offset) the expansion in lambda would lose log10(E_rms^2 / Var E) digits to cancellation
    const double e_centre = (MODE == 1) ? vcp->lambda : 0.0;
    constexpr int ROW = (MODE == 0) ? KPAD : KPAD + 6;    // doubles per staged sample row (even)
    constexpr int NP = (MODE == 0) ? KPAD : kMbStats;
    __shared__ __align__(16) double xs[kMbTile * ROW];
    __shared__ uint32_t gs[(MODE == 0) ? kMbThreads * 9 : 1];   // 128 rows x 8 words (+1 pad word: conflict-free)
    const int jb = blockIdx.x * kMbThreads, j = jb + threadIdx.x;
    const bool valid = j < mc;
    const int64_t i0 = (int64_t)blockIdx.y * per_split, i1 = min(n, i0 + per_split);
    double acc[NP];
#pragma unroll
    for (int c = 0; c < NP; c++) acc[c] = 0;
    double gval[4] = {0, 0, 0, 0}, b[KPAD];
#pragma unroll
    for (int c = 0; c < KPAD; c++) b[c] = 0;
    if (MODE == 0 && valid) {
        const Prologue pr = make_prologue(asum[j], nmiss[j], n);
        const GClass gc = make_gclass(pr, g_model);   // ref :1030-1047
#pragma unroll
        for (int c = 0; c < 4; c++) gval[c] = gc.val[c];
    }
    if (MODE == 1 && valid) {
#pragma unroll
        for (int c = 0; c < KPAD; c++) b[c] = beta[(int64_t)j * KPAD + c];
    }
    for (int64_t it = i0; it < i1; it += kMbTile) {
        __syncthreads();
        {   // stage the sample rows: thread t owns sample it + t (coalesced over the column-major PC matrix)
            const int64_t i = it + threadIdx.x;
            double *xr = xs + threadIdx.x * ROW;
            const bool in = i < i1;
            xr[0] = in ? 1.0 : 0.0;
#pragma unroll
            for (int c = 1; c < KPAD; c++) xr[c] = (in && c <= k) ? pcs[i + (int64_t)(c - 1) * n] : 0.0;
            if (MODE == 1) {
                const double r = in ? R[i] : 0.0, e = in ? E[i] - e_centre : 0.0, r2 = r * r;
                xr[KPAD + 0] = r; xr[KPAD + 1] = r2; xr[KPAD + 2] = e * r2; xr[KPAD + 3] = e * r;
                xr[KPAD + 4] = (e * e) * r2; xr[KPAD + 5] = 0.0;
            }
        }
        if (MODE == 0) {   // stage 32 packed bytes (128 samples) of each of the 128 rows, 16 bytes per thread and trip
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int q = threadIdx.x + kMbThreads * e, row = q >> 1, half = q & 1;
                uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);   // code 11 = dosage 0
                const int64_t off = (it >> 2) + half * 16;
                if (jb + row < mc && off + 16 <= pitch) v = *(const uint4 *)(bed + (int64_t)(jb + row) * pitch + off);
                uint32_t *g = gs + row * 9 + half * 4;
                g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
            }
        }
        __syncthreads();
        if (MODE == 0) {
#pragma unroll 1
            for (int w = 0; w < 8; w++) {
                const uint32_t word = gs[threadIdx.x * 9 + w];
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    const uint32_t code = (word >> (2 * s)) & 3u;
                    const double g = (code & 1u) ? ((code & 2u) ? gval[3] : gval[1]) : ((code & 2u) ? gval[2] : gval[0]);
                    const double2 *xr = (const double2 *)(xs + (w * 16 + s) * ROW);
#pragma unroll
                    for (int c = 0; c < KPAD / 2; c++) {
                        const double2 x2 = xr[c];
                        acc[2 * c] += g * x2.x; acc[2 * c + 1] += g * x2.y;
                    }
                }
            }
        } else {
#pragma unroll 2
            for (int s = 0; s < kMbTile; s++) {
                const double2 *xr = (const double2 *)(xs + s * ROW);
                double xv[KPAD + 6];
#pragma unroll
                for (int c = 0; c < (KPAD + 6) / 2; c++) { const double2 x2 = xr[c]; xv[2 * c] = x2.x; xv[2 * c + 1] = x2.y; }
                // fitted value in the order of the cluster kernel: f = beta0; f += pc_c * beta_c
                double f = b[0] * xv[0];
#pragma unroll
                for (int c = 1; c < KPAD; c++) f += xv[c] * b[c];
                double m = f / 2;                                  // Fit.lm$fitted.values / 2 (ref :1067)
                if (m < 0) acc[0] += 1.0;                          // ref :1071 (padding rows give f = 0)
                if (m <= 0) m = 0;                                 // ref :1095-1096
                if (m >= 1) m = 1;
                const double gv = 2 * m * (1 - m);
                acc[1] += xv[KPAD + 0] * m;                        // sum R m          (ref :1105)
                acc[2] += xv[KPAD + 1] * gv;                       // sum R^2 gvar     (ref :1106)
                acc[3] += gv * xv[KPAD + 2];                       // sum gvar E R^2   (ref :1115)
                acc[4] += xv[KPAD + 3] * m;                        // sum E R m
                acc[5] += xv[KPAD + 4] * gv;                       // sum E^2 R^2 gvar
            }
        }
    }
    if (valid) {
        double *dst = partial + ((int64_t)blockIdx.y * mc + j) * NP;
#pragma unroll
        for (int c = 0; c < NP; c++) dst[c] = acc[c];
    }
}

template <int KPAD>
__global__ void k_mix_beta(const double *__restrict__ partial, int n_splits, int mc, int k, const double *__restrict__ xtx_inv,
                           double *__restrict__ beta) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= mc) return;
    double xg[KPAD];
#pragma unroll
    for (int c = 0; c < KPAD; c++) xg[c] = 0;
    for (int s = 0; s < n_splits; s++) {
        const double *p = partial + ((int64_t)s * mc + j) * KPAD;
#pragma unroll
        for (int c = 0; c < KPAD; c++) xg[c] += p[c];
    }
    const int k1 = k + 1;
#pragma unroll
    for (int a = 0; a < KPAD; a++) {
        double bsum = 0;
        if (a < k1) {
#pragma unroll
            for (int c = 0; c < KPAD; c++) if (c < k1) bsum += xtx_inv[a + c * k1] * xg[c];   // same order as k_mix_snp
        }
        beta[(int64_t)j * KPAD + a] = bsum;
    }
}

__global__ void k_mix_finish(const double *__restrict__ partial, int n_splits, int mc, int64_t snp0, int64_t m_total, int64_t n,
                             SnpStats st, const VecConst *__restrict__ vcp, const uint8_t *__restrict__ want, MixBatchParams prm,
                             double *__restrict__ out, int *__restrict__ slow_list, Counters *ctr, MixHand *__restrict__ hand) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    bool slow = false, spa_only = false;
    if (j < mc && want[j]) {
        const VecConst vc = *vcp;
        double a[kMbStats] = {0, 0, 0, 0, 0, 0};
        for (int s = 0; s < n_splits; s++) {
            const double *p = partial + ((int64_t)s * mc + j) * kMbStats;
#pragma unroll
            for (int c = 0; c < kMbStats; c++) a[c] += p[c];
        }
        const int asum = st.asum[j], nmiss = st.nmiss[j];
        const Prologue pr = make_prologue(asum, nmiss, n);
        const double MAC = pr.maf * 2 * (double)n;                                    // ref :1058
        const double negnum = a[0], negratio = negnum / (double)n;
        slow = (MAC <= 20) || (negratio > prm.neg_ratio_cut);                          // ref :1060, :1074
        if (!slow) {
            double S1 = st.D0[j] + ((nmiss > 0) ? pr.u * st.T0[j] : 0.0);
            double S2 = st.D1[j] + ((nmiss > 0) ? pr.u * st.T1[j] : 0.0);
            if (pr.flip) { S1 = 2 * vc.sumR - S1; S2 = 2 * vc.sumV - S2; }
            if (prm.g_model != 0) {
                const GClass gc = make_gclass(pr, prm.g_model);
                S1 = class_dot(gc, st.D0[j], st.T0[j], st.H0[j], vc.sumR);
                S2 = class_dot(gc, st.D1[j], st.T1[j], st.H1[j], vc.sumV);
            }
            const double meanS1 = 2 * a[1], VarS1 = a[2];                              // ref :1105-1106
            const double Z1 = (S1 - meanS1) / sqrt(VarS1);
            const double pG = d_pnorm(-fabs(Z1), true) * 2;
            slow = !(pG > prm.epsilon);                                               // branch B (ref :1150)
            if (!slow) {
                // a[3..5] were accumulated with E' = E - c (c = vc.lambda, SNP-invariant): lambda' = lambda - c
                const double lam_c = a[3] / a[2];
                const double lambda = lam_c + vc.lambda;                              // ref :1115
                const double S_GxE = S2 - lambda * S1;
                // sum R.new m and sum R.new^2 gvar with R.new = (E - lambda) R = (E' - lambda') R, expanded in lambda'
                const double Smean = 2 * (a[4] - lam_c * a[1]);
                const double Svar = a[5] - 2 * lam_c * a[3] + (lam_c * lam_c) * a[2];
                const double z = (S_GxE - Smean) / sqrt(Svar);
                slow = !(fabs(z) < prm.cutoff);                                       // SPA (ref :1127)
                if (slow) {   // everything but the two tail probabilities is final: hand the statistics over
                    spa_only = true;
                    hand[j] = MixHand{S1, VarS1, Z1, pG, lambda, S_GxE, Smean, Svar, z, negnum, MAC};
                }
                if (!slow) {
                    const double pn = d_pnorm(fabs(z), false) * 2;
                    double *o = out + snp0 + j;
                    o[2 * m_total] = pn; o[3 * m_total] = pn; o[4 * m_total] = pn; o[5 * m_total] = pn;
                    o[6 * m_total] = pG; o[7 * m_total] = S1; o[8 * m_total] = VarS1; o[9 * m_total] = Z1;
                    o[10 * m_total] = negnum; o[11 * m_total] = MAC;
                    atomicAdd(&ctr->n_branch_a, 1ULL);
                }
            }
        }
    }
    const int pos = list_reserve(&ctr->b_count, slow);
    if (slow) slow_list[pos] = spa_only ? (j | kMixHandFlag) : j;
}

// ---- host side ----------------------------------------------------------------------------------
struct MixScratch {
    uint8_t *want = nullptr; double *part = nullptr, *beta = nullptr; MixHand *hand = nullptr;
    int64_t cap_want = 0, cap_part = 0, cap_beta = 0, cap_hand = 0; int beta_ld = 0;
};
const double *mix_scratch_beta(const MixScratch *ms, int *ld) { *ld = ms->beta_ld; return ms->beta; }
const MixHand *mix_scratch_hand(const MixScratch *ms) { return ms->hand; }
MixScratch *mix_scratch_create() { return new MixScratch(); }
void mix_scratch_destroy(MixScratch *ms) {
    if (!ms) return;
    cudaFree(ms->want); cudaFree(ms->part); cudaFree(ms->beta); cudaFree(ms->hand);
    delete ms;
}

template <typename T>
static cudaError_t grow(T **p, int64_t *cap, int64_t need) {
    if (need <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc((void **)p, (size_t)need * sizeof(T));
    if (e == cudaSuccess) *cap = need;
    return e;
}

template <int KPAD>
static cudaError_t run_k(cudaStream_t s, int num_sms, MixScratch &ms, const uint8_t *d_rows, int64_t pitch, SnpStats st, int64_t j0,
                         int mc, int64_t M, int64_t n, const double *R, const double *E, const VecConst *vc, const double *pcs,
                         const double *xtx_inv, int k, const MixBatchParams &prm, double *d_out, int *b_list, Counters *ctr,
                         int64_t *launches) {
    const int blocks = (mc + kMbThreads - 1) / kMbThreads;
    int64_t tiles = (n + kMbTile - 1) / kMbTile;
    int n_splits = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, (4 * num_sms + blocks - 1) / blocks));
    const int64_t per_split = ((tiles + n_splits - 1) / n_splits) * kMbTile;
    n_splits = (int)((n + per_split - 1) / per_split);
    cudaError_t e;
    const int64_t np = std::max<int64_t>(KPAD, kMbStats);
    if ((e = grow(&ms.part, &ms.cap_part, (int64_t)n_splits * mc * np)) != cudaSuccess) return e;
    if ((e = grow(&ms.beta, &ms.cap_beta, (int64_t)mc * KPAD)) != cudaSuccess) return e;
    if ((e = grow(&ms.hand, &ms.cap_hand, (int64_t)mc)) != cudaSuccess) return e;
    ms.beta_ld = KPAD;
    dim3 grid((unsigned)blocks, (unsigned)n_splits);
    k_mix_tile<KPAD, 0><<<grid, kMbThreads, 0, s>>>(d_rows, pitch, mc, n, k, pcs, R, E, nullptr, st.asum, st.nmiss, ms.part, per_split, prm.g_model, vc);
    k_mix_beta<KPAD><<<(mc + 127) / 128, 128, 0, s>>>(ms.part, n_splits, mc, k, xtx_inv, ms.beta);
    k_mix_tile<KPAD, 1><<<grid, kMbThreads, 0, s>>>(d_rows, pitch, mc, n, k, pcs, R, E, ms.beta, st.asum, st.nmiss, ms.part, per_split, prm.g_model, vc);
    k_mix_finish<<<(mc + 127) / 128, 128, 0, s>>>(ms.part, n_splits, mc, j0, M, n, st, vc, ms.want, prm, d_out, b_list, ctr, ms.hand);
    *launches += 4;
    return cudaGetLastError();
}

cudaError_t mix_batch_run(cudaStream_t s, MixScratch *scratch, int num_sms, const uint8_t *d_rows, int64_t pitch, SnpStats st, int64_t j0, int mc,
                          int64_t M, int64_t n, const double *R, const double *E, const VecConst *vc, const double *pcs,
                          const double *xtx_inv, int k, const MixBatchParams &prm, double *d_out, int *b_list, Counters *ctr,
                          int64_t *launches) {
    if (!scratch) return cudaErrorInvalidValue;
    MixScratch &ms = *scratch;
    cudaError_t e;
    if ((e = grow(&ms.want, &ms.cap_want, (int64_t)mc)) != cudaSuccess) return e;
    // the candidate list written by k_finalize_mix becomes a per-SNP flag; the list is then refilled with the slow SNPs
    if ((e = cudaMemsetAsync(ms.want, 0, (size_t)mc, s)) != cudaSuccess) return e;
    k_mix_mark<<<(mc + 255) / 256, 256, 0, s>>>(b_list, &ctr->b_count, ms.want);
    if ((e = cudaMemsetAsync(&ctr->b_count, 0, sizeof(int), s)) != cudaSuccess) return e;
    *launches += 1;
    const int k1 = k + 1;
    if (k1 <= 4) return run_k<4>(s, num_sms, ms, d_rows, pitch, st, j0, mc, M, n, R, E, vc, pcs, xtx_inv, k, prm, d_out, b_list, ctr, launches);
    if (k1 <= 8) return run_k<8>(s, num_sms, ms, d_rows, pitch, st, j0, mc, M, n, R, E, vc, pcs, xtx_inv, k, prm, d_out, b_list, ctr, launches);
    if (k1 <= 12) return run_k<12>(s, num_sms, ms, d_rows, pitch, st, j0, mc, M, n, R, E, vc, pcs, xtx_inv, k, prm, d_out, b_list, ctr, launches);
    return run_k<16>(s, num_sms, ms, d_rows, pitch, st, j0, mc, M, n, R, E, vc, pcs, xtx_inv, k, prm, d_out, b_list, ctr, launches);
}

}  // namespace spagxe
