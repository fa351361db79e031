// cox_wald.cu -- Wald test of the E:g coefficient for traits = "survival" in branch B of SPAGxE_CCT.
//
// Reference (R/SPAGxECCT.R:622-634):
//     data0 = coxph(Surv(Phen.mtx$surv.time, Phen.mtx$event) ~ as.matrix(Cova.mtx) + E + g + g*E, iter.max = 1000)
//     pval.wald = summary(data0)$coefficients["E:g", "Pr(>|z|)"]
// `survival` is a third-party package that is not part of the reference tree; what is restated here is the published
// algorithm of coxph.fit / coxfit6.c: Efron ties, start at beta = 0, Newton-Raphson on the partial likelihood with
// `fabs(1 - loglik_old / loglik_new) <= 1e-9` as the stop, step halving when the likelihood drops, variance = inverse
// information at the LAST iterate, z = beta / se, p = pchisq(z^2, 1, lower = FALSE) = 2 pnorm(-|z|).
//
// Data layout (SNP-invariant, built once per analysis by cox_setup): samples in DESCENDING time order, so that the risk
// set of a death time is a prefix; the covariate columns and E gathered into that order and centred (Newton iterates
// are invariant under the affine reparametrisation, coxfit6 centres too); per sorted sample an int: bit 0 = death,
// bits 1.. = number of deaths of the tie group that ENDS at this sample (0 elsewhere).  Samples behind the last death
// are in no risk set and are dropped.
//
// Kernel: one thread-block cluster (8 CTAs x 8 warps) per work item, persistent over the device-side item list.
// A warp owns a contiguous run of sorted samples.  Per likelihood evaluation: pass A = per-warp totals of
// v = (w, w x_a, w x_a x_b), w = exp(x'beta), and of the deaths behind the warp's last group end; the totals meet
// through distributed shared memory and every warp forms its carry-in in a fixed order; pass B = the same sweep with
// warp scans in the chunks that hold a death or a group end, the Efron terms of every death group evaluated by the lane
// that owns its last sample.  All sums have a fixed order, so results are reproducible; every CTA redoes the scalar
// Newton logic on identical totals.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <numeric>
#include <vector>

#include "cox_wald.h"
#include "dmath.cuh"
#include "snp_common.cuh"

namespace spagxe {
namespace cg = cooperative_groups;

constexpr int kCxWarps = 8;
constexpr int kCxThreads = kCxWarps * 32;
constexpr int kCxCluster = 8;
constexpr int kCxQ = 5;                                  // design columns: E, g, E:g, cova[0], cova[1]
constexpr int kCxNC = kCxQ * (kCxQ + 1) / 2;             // 15
constexpr int kCxNV = 1 + kCxQ + kCxNC;                  // 21 values per sample
constexpr int kCxMaxIter = 1000;                         // iter.max of the reference call
constexpr double kCxEps = 1e-9;                          // coxph.control(eps)
constexpr double kCxTolerChol = 1.818989403545856e-12;   // .Machine$double.eps^0.75

struct CoxArgs {
    const BItem *items; const int *count;
    int64_t m_total, n, n_eff;
    const int *perm;        // sorted position -> sample index
    const double *x;        // (p + 1) columns x n_eff: E, cova[0..p) gathered and centred
    const int *info;        // bit 0 death, bits 1..: deaths of the tie group ending here
    int p, g_model, unit_len;
    double *out; Counters *ctr;
};

__device__ __forceinline__ constexpr int cx_pair(int a, int b) { return 1 + kCxQ + a * (a + 1) / 2 + b; }   // b <= a

struct CxShared {
    double totP[kCxWarps][kCxNV], tailD[kCxWarps][kCxNV];
    double carryP[kCxWarps][kCxNV], carryG[kCxWarps][kCxNV];
    double contrib[kCxWarps][kCxNV];          // loglik, u[5], imat lower triangle[15]
    double ctaP[kCxNV], ctaTailD[kCxNV], slotC[kCxNV];   // read by the other CTAs of the cluster
    double red[kCxNV];
    double beta[kCxQ], newbeta[kCxQ];
    double loglik_old, pw;
    int hasb[kCxWarps], ctaHasb;
    int iter, halving, status;                // status: 0 running, 1 done, 2 failed (singular information)
};

// LDL' of the leading q x q block (cholesky2 of the survival package): returns the rank, zeroes deficient pivots
__device__ int cx_cholesky(double (*m)[kCxQ], int q, double toler) {
    double eps = 0;
    for (int i = 0; i < q; i++) { if (m[i][i] > eps) eps = m[i][i]; for (int j = i + 1; j < q; j++) m[j][i] = m[i][j]; }
    eps = (eps == 0) ? toler : eps * toler;
    int rank = 0;
    for (int i = 0; i < q; i++) {
        const double pivot = m[i][i];
        if (!isfinite(pivot) || pivot < eps) { m[i][i] = 0; continue; }
        rank++;
        for (int j = i + 1; j < q; j++) {
            const double t = m[j][i] / pivot;
            m[j][i] = t;
            m[j][j] -= t * t * pivot;
            for (int k = j + 1; k < q; k++) m[k][j] -= t * m[k][i];
        }
    }
    return rank;
}
__device__ void cx_solve(double (*m)[kCxQ], int q, double *y) {   // chsolve2
    for (int i = 0; i < q; i++) { double t = y[i]; for (int j = 0; j < i; j++) t -= y[j] * m[i][j]; y[i] = t; }
    for (int i = q - 1; i >= 0; i--) {
        if (m[i][i] == 0) { y[i] = 0; continue; }
        double t = y[i] / m[i][i];
        for (int j = i + 1; j < q; j++) t -= y[j] * m[j][i];
        y[i] = t;
    }
}
// diagonal element `k` of the inverse from the LDL' factors: e_k solved through the factorisation
__device__ double cx_inv_diag(double (*m)[kCxQ], int q, int k) {
    double e[kCxQ];
    for (int i = 0; i < q; i++) e[i] = (i == k) ? 1.0 : 0.0;
    cx_solve(m, q, e);
    return e[k];
}

// one sample of the sorted order: design row, linear predictor, the 21 values
struct CxSample { double x[kCxQ]; double zb; double v[kCxNV]; };
__device__ __forceinline__ void cx_eval(const CoxArgs &A, const uint8_t *__restrict__ row, const double (&gval)[4],
                                        const double (&b)[kCxQ], int64_t s, bool valid, CxSample &o) {
    if (!valid) {
#pragma unroll
        for (int k = 0; k < kCxNV; k++) o.v[k] = 0;
#pragma unroll
        for (int k = 0; k < kCxQ; k++) o.x[k] = 0;
        o.zb = 0;
        return;
    }
    const int i = A.perm[s];
    const int code = (row[i >> 2] >> (2 * (i & 3))) & 3;
    const double g = (code & 1) ? ((code & 2) ? gval[3] : gval[1]) : ((code & 2) ? gval[2] : gval[0]);
    const double e = A.x[s];
    o.x[0] = e; o.x[1] = g; o.x[2] = e * g;
    o.x[3] = (A.p > 0) ? A.x[A.n_eff + s] : 0.0;
    o.x[4] = (A.p > 1) ? A.x[2 * A.n_eff + s] : 0.0;
    double eta = 0;
#pragma unroll
    for (int k = 0; k < kCxQ; k++) eta += b[k] * o.x[k];
    eta = fmin(fmax(eta, -200.0), 22.0);                    // coxsafe()
    o.zb = eta;
    const double w = exp(eta);
    o.v[0] = w;
#pragma unroll
    for (int a = 0; a < kCxQ; a++) {
        const double wa = w * o.x[a];
        o.v[1 + a] = wa;
#pragma unroll
        for (int c = 0; c <= a; c++) o.v[cx_pair(a, c)] = wa * o.x[c];
    }
}

__global__ void __launch_bounds__(kCxThreads, 1) k_cox_wald(CoxArgs A) {
    __shared__ CxShared sh;
    cg::cluster_group cl = cg::this_cluster();
    const int cs = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int cluster_id = blockIdx.x / cs, n_clusters = gridDim.x / cs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = rank * kCxWarps + warp;
    const int64_t s_lo = min(A.n_eff, (int64_t)unit * A.unit_len), s_hi = min(A.n_eff, s_lo + A.unit_len);
    const int q = 3 + A.p;
    const int count = *A.count;
    const double NA = d_na_real();
    for (int idx = cluster_id; idx < count; idx += n_clusters) {
        const BItem bi = A.items[idx];
        const uint8_t *row = bi.row;
        const Prologue pr = make_prologue(bi.asum, bi.nmiss, A.n);
        const GClass gc = make_gclass(pr, A.g_model);
        double gval[4];
#pragma unroll
        for (int c = 0; c < 4; c++) gval[c] = gc.val[c];
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 0; k < kCxQ; k++) { sh.beta[k] = 0; sh.newbeta[k] = 0; }
            sh.loglik_old = 0; sh.iter = 0; sh.halving = 0; sh.status = 0; sh.pw = NA;
        }
        __syncthreads();
        for (;;) {
            double b[kCxQ];
#pragma unroll
            for (int k = 0; k < kCxQ; k++) b[k] = sh.newbeta[k];
            // ---------------- pass A: totals of the run, deaths behind its last group end ----------------
            {
                double accP[kCxNV], accD[kCxNV];
#pragma unroll
                for (int k = 0; k < kCxNV; k++) { accP[k] = 0; accD[k] = 0; }
                int hasb = 0;
                for (int64_t s0 = s_lo; s0 < s_hi; s0 += 32) {
                    const int64_t s = s0 + lane;
                    const bool valid = s < s_hi;
                    const int info = valid ? A.info[s] : 0;
                    CxSample sm;
                    cx_eval(A, row, gval, b, s, valid, sm);
                    const bool death = info & 1;
                    const unsigned bm = __ballot_sync(0xffffffffu, (info >> 1) > 0);
#pragma unroll
                    for (int k = 0; k < kCxNV; k++) { accP[k] += sm.v[k]; if (death) accD[k] += sm.v[k]; }
                    if (bm) {   // every death up to the last group end of this chunk is closed
                        const int lb = 31 - __clz(bm);
#pragma unroll
                        for (int k = 0; k < kCxNV; k++) accD[k] = (lane > lb && death) ? sm.v[k] : 0.0;
                        hasb = 1;
                    }
                }
#pragma unroll
                for (int k = 0; k < kCxNV; k++) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        accP[k] += __shfl_xor_sync(0xffffffffu, accP[k], o);
                        accD[k] += __shfl_xor_sync(0xffffffffu, accD[k], o);
                    }
                }
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < kCxNV; k++) { sh.totP[warp][k] = accP[k]; sh.tailD[warp][k] = accD[k]; sh.contrib[warp][k] = 0; }
                    sh.hasb[warp] = hasb;
                }
            }
            __syncthreads();
            if (warp == 0 && lane < kCxNV) {
                double t = 0;
                for (int w = 0; w < kCxWarps; w++) t += sh.totP[w][lane];
                sh.ctaP[lane] = t;
                double d = 0; int hb = 0;
                for (int w = kCxWarps - 1; w >= 0; w--) { d += sh.tailD[w][lane]; if (sh.hasb[w]) { hb = 1; break; } }
                sh.ctaTailD[lane] = d;
                if (lane == 0) sh.ctaHasb = hb;
            }
            cl.sync();
            if (lane < kCxNV) {   // carry-in of this warp: prefix of the totals, deaths since the last group end before it
                double cp = 0;
                for (int r = 0; r < rank; r++) cp += cl.map_shared_rank(sh.ctaP, r)[lane];
                for (int w = 0; w < warp; w++) cp += sh.totP[w][lane];
                sh.carryP[warp][lane] = cp;
                double cgv = 0; bool open = true;
                for (int w = warp - 1; w >= 0 && open; w--) { cgv += sh.tailD[w][lane]; if (sh.hasb[w]) open = false; }
                for (int r = rank - 1; r >= 0 && open; r--) {
                    cgv += cl.map_shared_rank(sh.ctaTailD, r)[lane];
                    if (*cl.map_shared_rank(&sh.ctaHasb, r)) open = false;
                }
                sh.carryG[warp][lane] = cgv;
            }
            __syncwarp();
            // ---------------- pass B: Efron terms at the group ends ----------------
            {
                double runP[kCxNV];
#pragma unroll
                for (int k = 0; k < kCxNV; k++) runP[k] = 0;
                for (int64_t s0 = s_lo; s0 < s_hi; s0 += 32) {
                    const int64_t s = s0 + lane;
                    const bool valid = s < s_hi;
                    const int info = valid ? A.info[s] : 0;
                    CxSample sm;
                    cx_eval(A, row, gval, b, s, valid, sm);
                    const bool death = info & 1;
                    const int nd = info >> 1;
                    const unsigned bm = __ballot_sync(0xffffffffu, nd > 0);
                    const unsigned dm = __ballot_sync(0xffffffffu, death);
                    if (!(bm | dm)) {
#pragma unroll
                        for (int k = 0; k < kCxNV; k++) runP[k] += sm.v[k];
                        continue;
                    }
                    // fold the lanes' private sums of the quiet chunks into the carry
#pragma unroll
                    for (int k = 0; k < kCxNV; k++) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) runP[k] += __shfl_xor_sync(0xffffffffu, runP[k], o);
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < kCxNV; k++) sh.carryP[warp][k] += runP[k];
                    }
#pragma unroll
                    for (int k = 0; k < kCxNV; k++) runP[k] = 0;
                    // inclusive scans over the lanes: all samples (sv, in place) and deaths only (sd)
                    double sd[kCxNV];
#pragma unroll
                    for (int k = 0; k < kCxNV; k++) {
                        double a = sm.v[k], d = death ? sm.v[k] : 0.0;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const double ta = __shfl_up_sync(0xffffffffu, a, o), td = __shfl_up_sync(0xffffffffu, d, o);
                            if (lane >= o) { a += ta; d += td; }
                        }
                        sm.v[k] = a; sd[k] = d;
                    }
                    __syncwarp();
                    // deaths since the previous group end: inside the chunk a difference of scans, else carry + scan
                    const unsigned below = bm & ((1u << lane) - 1u);
                    const int pbl = below ? 31 - __clz(below) : -1;
                    const int lb = bm ? 31 - __clz(bm) : -1;
                    double gsum[kCxNV];
#pragma unroll
                    for (int k = 0; k < kCxNV; k++) {
                        const double t = __shfl_sync(0xffffffffu, sd[k], pbl < 0 ? 0 : pbl);
                        gsum[k] = (pbl >= 0) ? sd[k] - t : sh.carryG[warp][k] + sd[k];
                    }
                    // direct terms of the deaths: loglik += zbeta, u += x (lane order)
                    for (unsigned m = dm; m; m &= m - 1) {
                        if (lane == __ffs(m) - 1) {
                            sh.contrib[warp][0] += sm.zb;
#pragma unroll
                            for (int a = 0; a < kCxQ; a++) sh.contrib[warp][1 + a] += sm.x[a];
                        }
                        __syncwarp();
                    }
                    // Efron terms of the groups that end in this chunk (lane order)
                    for (unsigned m = bm; m; m &= m - 1) {
                        if (lane == __ffs(m) - 1) {
                            double P[kCxNV];
#pragma unroll
                            for (int k = 0; k < kCxNV; k++) P[k] = sh.carryP[warp][k] + sm.v[k];
                            for (int jd = 1; jd <= nd; jd++) {
                                const double f = (double)(nd - jd) / (double)nd;   // share of the death sums left out
                                const double denom = P[0] - f * gsum[0];
                                sh.contrib[warp][0] -= log(denom);
                                double av[kCxQ];
#pragma unroll
                                for (int a = 0; a < kCxQ; a++) av[a] = P[1 + a] - f * gsum[1 + a];
#pragma unroll
                                for (int a = 0; a < kCxQ; a++) {
                                    const double mean = av[a] / denom;
                                    sh.contrib[warp][1 + a] -= mean;
#pragma unroll
                                    for (int c = 0; c <= a; c++) {
                                        const double cm = P[cx_pair(a, c)] - f * gsum[cx_pair(a, c)];
                                        sh.contrib[warp][cx_pair(a, c)] += (cm - mean * av[c]) / denom;
                                    }
                                }
                            }
                        }
                        __syncwarp();
                    }
                    // carries behind this chunk
#pragma unroll
                    for (int k = 0; k < kCxNV; k++) {
                        const double t31 = __shfl_sync(0xffffffffu, sm.v[k], 31);
                        const double e31 = __shfl_sync(0xffffffffu, sd[k], 31);
                        const double elb = __shfl_sync(0xffffffffu, sd[k], lb < 0 ? 0 : lb);
                        if (lane == 0) {
                            sh.carryP[warp][k] += t31;
                            sh.carryG[warp][k] = (lb >= 0) ? e31 - elb : sh.carryG[warp][k] + e31;
                        }
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
            if (warp == 0 && lane < kCxNV) {
                double t = 0;
                for (int w = 0; w < kCxWarps; w++) t += sh.contrib[w][lane];
                sh.slotC[lane] = t;
            }
            cl.sync();
            if (warp == 0 && lane < kCxNV) {
                double t = 0;
                for (int r = 0; r < cs; r++) t += cl.map_shared_rank(sh.slotC, r)[lane];
                sh.red[lane] = t;
            }
            cl.sync();   // the exchange slots may be rewritten from here on
            // ---------------- Newton-Raphson bookkeeping of coxfit6 (every CTA, identical totals) ----------------
            if (threadIdx.x == 0) {
                double im[kCxQ][kCxQ], u[kCxQ];
                for (int a = 0; a < kCxQ; a++) {
                    u[a] = sh.red[1 + a];
                    for (int c = 0; c <= a; c++) { im[a][c] = sh.red[cx_pair(a, c)]; im[c][a] = im[a][c]; }
                }
                const double lk = sh.red[0];
                const int rk = cx_cholesky(im, q, kCxTolerChol);
                if (rk < q) sh.status = 2;
                else if (sh.iter == 0) {
                    sh.loglik_old = lk;
                    cx_solve(im, q, u);
                    for (int a = 0; a < q; a++) sh.newbeta[a] = sh.beta[a] + u[a];
                    sh.iter = 1;
                } else {
                    const bool conv = fabs(1 - sh.loglik_old / lk) <= kCxEps && !sh.halving;
                    if (conv || sh.iter == kCxMaxIter) {
                        const double var = cx_inv_diag(im, q, 2);           // "E:g" is column 2 here
                        const double z = sh.newbeta[2] / sqrt(var);
                        sh.pw = 2 * d_pnorm(-fabs(z), true);
                        sh.status = 1;
                    } else if (!(lk >= sh.loglik_old)) {                    // not converging (or not finite): step halving
                        sh.halving = 1;
                        for (int a = 0; a < q; a++) sh.newbeta[a] = (sh.newbeta[a] + sh.beta[a]) / 2;
                        sh.iter++;
                    } else {
                        sh.halving = 0;
                        sh.loglik_old = lk;
                        cx_solve(im, q, u);
                        for (int a = 0; a < q; a++) { sh.beta[a] = sh.newbeta[a]; sh.newbeta[a] += u[a]; }
                        sh.iter++;
                    }
                }
            }
            __syncthreads();
            if (sh.status != 0) break;
        }
        if (rank == 0 && threadIdx.x == 0) {
            double *o = A.out + bi.snp;
            const int64_t M = A.m_total;
            double pw = sh.pw;
            if (sh.status == 2 || isnan(pw)) { pw = NA; atomicAdd(&A.ctr->n_wald_fail, 1ULL); }
            double pc = NA;
            const int crc = d_cct2(o[2 * M], pw, &pc);   // p.spa of this item was written by the solver farm
            if (crc) atomicAdd(&A.ctr->n_cct_stop, 1ULL);
            o[3 * M] = pw; o[4 * M] = pc;
        }
    }
    cl.sync();
}

// ---- setup kernels: column means (one CTA per column, fixed order) and the gather into time order ----
__global__ void __launch_bounds__(1024) k_cox_means(const double *__restrict__ E, const double *__restrict__ cova, int64_t n, double *means) {
    __shared__ double red[1024];
    const double *col = (blockIdx.x == 0) ? E : cova + (int64_t)(blockIdx.x - 1) * n;
    double a = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) a += col[i];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) means[blockIdx.x] = red[0] / (double)n;
}
__global__ void k_cox_gather(const double *__restrict__ E, const double *__restrict__ cova, int64_t n, int64_t n_eff,
                             const int *__restrict__ perm, const double *__restrict__ means, double *__restrict__ x) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n_eff) return;
    const int c = blockIdx.y;
    const double *col = (c == 0) ? E : cova + (int64_t)(c - 1) * n;
    x[(int64_t)c * n_eff + s] = col[perm[s]] - means[c];
}

// ---- host side -----------------------------------------------------------------------------------------------------
struct CoxData {
    int64_t n = 0, n_eff = 0; int p = 0; bool ready = false;
    int *d_perm = nullptr, *d_info = nullptr; double *d_x = nullptr, *d_means = nullptr;
};
CoxData *cox_create() { return new CoxData(); }
void cox_destroy(CoxData *c) {
    if (!c) return;
    cudaFree(c->d_perm); cudaFree(c->d_info); cudaFree(c->d_x); cudaFree(c->d_means);
    delete c;
}
void cox_invalidate(CoxData *c) { if (c) c->ready = false; }
bool cox_ready(const CoxData *c) { return c && c->ready; }

cudaError_t cox_setup(CoxData *c, cudaStream_t s, int64_t n, const double *h_time, const double *h_event, const double *d_E,
                      const double *d_cova, int p, const char **what) {
    *what = "cox_setup";
    c->ready = false;
    if (p > kCoxMaxCov) { *what = "survival Wald: at most 2 covariate columns"; return cudaErrorInvalidValue; }
    std::vector<int> idx((size_t)n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return h_time[a] < h_time[b]; });   // coxph.fit: order(time)
    std::vector<int> perm((size_t)n), info((size_t)n, 0);
    for (int64_t k = 0; k < n; k++) perm[(size_t)k] = idx[(size_t)(n - 1 - k)];                       // coxfit6 walks from the largest time
    int64_t n_eff = 0;
    for (int64_t a = 0; a < n;) {
        int64_t e = a; int nd = 0;
        const double t = h_time[perm[(size_t)a]];
        while (e < n && h_time[perm[(size_t)e]] == t) {
            if (h_event[perm[(size_t)e]] != 0.0) { nd++; info[(size_t)e] |= 1; }
            e++;
        }
        if (nd > 0) { info[(size_t)(e - 1)] |= nd << 1; n_eff = e; }
        a = e;
    }
    if (n_eff == 0) { *what = "survival Wald: no events"; return cudaErrorInvalidValue; }
    cudaError_t err;
    cudaFree(c->d_perm); cudaFree(c->d_info); cudaFree(c->d_x); cudaFree(c->d_means);
    c->d_perm = c->d_info = nullptr; c->d_x = c->d_means = nullptr;
    if ((err = cudaMalloc((void **)&c->d_perm, sizeof(int) * (size_t)n_eff)) != cudaSuccess) return err;
    if ((err = cudaMalloc((void **)&c->d_info, sizeof(int) * (size_t)n_eff)) != cudaSuccess) return err;
    if ((err = cudaMalloc((void **)&c->d_x, sizeof(double) * (size_t)n_eff * (p + 1))) != cudaSuccess) return err;
    if ((err = cudaMalloc((void **)&c->d_means, sizeof(double) * (p + 1))) != cudaSuccess) return err;
    if ((err = cudaMemcpyAsync(c->d_perm, perm.data(), sizeof(int) * (size_t)n_eff, cudaMemcpyHostToDevice, s)) != cudaSuccess) return err;
    if ((err = cudaMemcpyAsync(c->d_info, info.data(), sizeof(int) * (size_t)n_eff, cudaMemcpyHostToDevice, s)) != cudaSuccess) return err;
    k_cox_means<<<p + 1, 1024, 0, s>>>(d_E, d_cova, n, c->d_means);
    k_cox_gather<<<dim3((unsigned)((n_eff + 255) / 256), (unsigned)(p + 1)), 256, 0, s>>>(d_E, d_cova, n, n_eff, c->d_perm, c->d_means, c->d_x);
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    if ((err = cudaStreamSynchronize(s)) != cudaSuccess) return err;   // perm / info are host vectors of this frame
    c->n = n; c->n_eff = n_eff; c->p = p; c->ready = true;
    return cudaSuccess;
}

cudaError_t cox_launch(CoxData *c, cudaStream_t s, int num_sms, const BItem *items, const int *count, int64_t m_total, int g_model,
                       double *out, Counters *ctr, int64_t *launches) {
    if (!c || !c->ready) return cudaErrorInvalidValue;
    CoxArgs a;
    a.items = items; a.count = count; a.m_total = m_total; a.n = c->n; a.n_eff = c->n_eff;
    a.perm = c->d_perm; a.x = c->d_x; a.info = c->d_info; a.p = c->p; a.g_model = g_model;
    const int units = kCxCluster * kCxWarps;
    a.unit_len = (int)(((c->n_eff + units - 1) / units + 31) / 32 * 32);
    a.out = out; a.ctr = ctr;
    cudaLaunchConfig_t cfg = {};
    const int n_clusters = std::max(1, num_sms / kCxCluster);
    cfg.gridDim = dim3((unsigned)(n_clusters * kCxCluster));
    cfg.blockDim = dim3(kCxThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCxCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_cox_wald, a);
    if (e == cudaSuccess) (*launches)++;
    return e;
}

}  // namespace spagxe
