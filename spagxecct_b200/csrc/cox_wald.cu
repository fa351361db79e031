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
// Every LANE owns a contiguous run of sorted samples (the arrays are stored lane-interleaved so that the loads of a warp
// stay coalesced).  Per likelihood evaluation: pass A = per-lane totals of v = (w, w x_a, w x_a x_b), w = exp(x'beta),
// and of the deaths behind the lane's last group end; one warp scan, then the warp / CTA totals meet through
// distributed shared memory and every lane forms its carry-in in a fixed order; pass B = the same walk with a running
// prefix in registers, the Efron terms of a death group evaluated by the lane that reaches its last sample.  No
// shuffles inside the walks, all sums have a fixed order (reproducible results); every CTA redoes the scalar Newton
// logic on identical totals.
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
    int p, g_model, unit_len, na_to_spa;
    double *out; Counters *ctr;
    uint8_t *codes;         // per cluster: the item's 2-bit codes gathered into time order (n_eff bytes each)
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

// one sample of the sorted order: raw operands (loaded one chunk ahead: every sweep is a dependent chain of loads
// otherwise), then the design row, the linear predictor and the 21 values
struct CxIn { int info, code; double e, c0, c1; };
__device__ __forceinline__ CxIn cx_load(const CoxArgs &A, const uint8_t *__restrict__ codes, int64_t s, int64_t s_hi) {
    CxIn in; in.info = 0; in.code = 3; in.e = 0; in.c0 = 0; in.c1 = 0;
    if (s < s_hi) {
        in.info = A.info[s];                      // bit 30: a valid sample (clear in the padding of the storage order)
        in.code = __ldcg(codes + s);              // written by other SMs for this item: not through L1
        in.e = A.x[s];
        if (A.p > 0) in.c0 = A.x[A.n_eff + s];
        if (A.p > 1) in.c1 = A.x[2 * A.n_eff + s];
    }
    return in;
}
struct CxSample { double x[kCxQ]; double zb; double v[kCxNV]; };
__device__ __forceinline__ void cx_eval(const CxIn &in, const double (&gval)[4], const double (&b)[kCxQ], CxSample &o) {
    if (!(in.info >> 30)) {
#pragma unroll
        for (int k = 0; k < kCxNV; k++) o.v[k] = 0;
#pragma unroll
        for (int k = 0; k < kCxQ; k++) o.x[k] = 0;
        o.zb = 0;
        return;
    }
    const int code = in.code;
    const double g = (code & 1) ? ((code & 2) ? gval[3] : gval[1]) : ((code & 2) ? gval[2] : gval[0]);
    const double e = in.e;
    o.x[0] = e; o.x[1] = g; o.x[2] = e * g;
    o.x[3] = in.c0;
    o.x[4] = in.c1;
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
    // storage order: unit u, step t, lane l at index u * unit_len + 32 t + l holds sorted position u * unit_len + l * (unit_len / 32) + t,
    // i.e. a lane walks a contiguous run of the time order while the warp's loads stay coalesced
    const int64_t j_lo = (int64_t)unit * A.unit_len, j_hi = j_lo + A.unit_len;
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
        // the item's genotype codes in time order, gathered once (every sweep then streams them)
        uint8_t *codes = A.codes + (size_t)cluster_id * (size_t)A.n_eff;
        cl.sync();   // the previous item's sweeps are over in every CTA of the cluster
        for (int64_t s = (int64_t)rank * kCxThreads + threadIdx.x; s < A.n_eff; s += (int64_t)cs * kCxThreads) {
            const int i = A.perm[s];   // -1: padding of the storage order
            codes[s] = (i < 0) ? (uint8_t)3 : (uint8_t)((row[i >> 2] >> (2 * (i & 3))) & 3);
        }
        __threadfence();
        cl.sync();
        if (threadIdx.x == 0) {
            for (int k = 0; k < kCxQ; k++) { sh.beta[k] = 0; sh.newbeta[k] = 0; }
            sh.loglik_old = 0; sh.iter = 0; sh.halving = 0; sh.status = 0; sh.pw = NA;
        }
        __syncthreads();
        for (;;) {
            double b[kCxQ];
#pragma unroll
            for (int k = 0; k < kCxQ; k++) b[k] = sh.newbeta[k];
            // ---------------- pass A: totals of every lane's sub-run, deaths behind its last group end ----------------
            double P[kCxNV], G[kCxNV];      // become the running prefix / open-death sums of pass B
            bool open_lane;                 // no group end in this warp before this lane
            {
#pragma unroll
                for (int k = 0; k < kCxNV; k++) { P[k] = 0; G[k] = 0; }
                int hasb = 0;
                CxIn nxt = cx_load(A, codes, j_lo + lane, j_hi);
                for (int64_t j = j_lo; j < j_hi; j += 32) {
                    const CxIn cur = nxt;
                    nxt = cx_load(A, codes, j + 32 + lane, j_hi);
                    const int info = cur.info & 0x3fffffff;
                    CxSample sm;
                    cx_eval(cur, gval, b, sm);
                    const bool death = info & 1;
#pragma unroll
                    for (int k = 0; k < kCxNV; k++) { P[k] += sm.v[k]; if (death) G[k] += sm.v[k]; }
                    if (info >> 1) {   // a death group ends here: every death so far is closed
#pragma unroll
                        for (int k = 0; k < kCxNV; k++) G[k] = 0;
                        hasb = 1;
                    }
                }
                // lanes in order: exclusive prefix of the totals, segmented (reset at a group end) sum of the open deaths
                int f = hasb;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int tf = __shfl_up_sync(0xffffffffu, f, o);
#pragma unroll
                    for (int k = 0; k < kCxNV; k++) {
                        const double tp = __shfl_up_sync(0xffffffffu, P[k], o), tg = __shfl_up_sync(0xffffffffu, G[k], o);
                        if (lane >= o) { P[k] += tp; if (!f) G[k] += tg; }
                    }
                    if (lane >= o) f |= tf;
                }
                if (lane == 31) {
#pragma unroll
                    for (int k = 0; k < kCxNV; k++) { sh.totP[warp][k] = P[k]; sh.tailD[warp][k] = G[k]; }
                    sh.hasb[warp] = f;
                }
                // shift to exclusive: what the lanes before this one hold
                const int fprev = __shfl_up_sync(0xffffffffu, f, 1);
#pragma unroll
                for (int k = 0; k < kCxNV; k++) {
                    const double tp = __shfl_up_sync(0xffffffffu, P[k], 1), tg = __shfl_up_sync(0xffffffffu, G[k], 1);
                    P[k] = lane ? tp : 0.0; G[k] = lane ? tg : 0.0;
                }
                open_lane = lane ? !fprev : true;
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
#pragma unroll
            for (int k = 0; k < kCxNV; k++) { P[k] += sh.carryP[warp][k]; if (open_lane) G[k] += sh.carryG[warp][k]; }
            // ---------------- pass B: running prefix per lane, Efron terms where a death group ends ----------------
            {
                double ct[kCxNV];           // loglik, u[5], information (lower triangle): this lane's share
#pragma unroll
                for (int k = 0; k < kCxNV; k++) ct[k] = 0;
                CxIn nxt = cx_load(A, codes, j_lo + lane, j_hi);
                for (int64_t j = j_lo; j < j_hi; j += 32) {
                    const CxIn cur = nxt;
                    nxt = cx_load(A, codes, j + 32 + lane, j_hi);
                    const int info = cur.info & 0x3fffffff;
                    CxSample sm;
                    cx_eval(cur, gval, b, sm);
                    const int nd = info >> 1;
#pragma unroll
                    for (int k = 0; k < kCxNV; k++) P[k] += sm.v[k];
                    if (info & 1) {   // a death: direct terms, and it joins the open group
                        ct[0] += sm.zb;
#pragma unroll
                        for (int a = 0; a < kCxQ; a++) ct[1 + a] += sm.x[a];
#pragma unroll
                        for (int k = 0; k < kCxNV; k++) G[k] += sm.v[k];
                    }
                    if (nd > 0) {     // the group ends here: nd Efron terms, then the group is closed
                        for (int jd = 1; jd <= nd; jd++) {
                            const double f = (double)(nd - jd) / (double)nd;   // share of the death sums left out
                            const double denom = P[0] - f * G[0];
                            const double rden = 1.0 / denom;
                            ct[0] -= log(denom);
                            double mean[kCxQ];
#pragma unroll
                            for (int a = 0; a < kCxQ; a++) { mean[a] = (P[1 + a] - f * G[1 + a]) * rden; ct[1 + a] -= mean[a]; }
#pragma unroll
                            for (int a = 0; a < kCxQ; a++) {
#pragma unroll
                                for (int c = 0; c <= a; c++)
                                    ct[cx_pair(a, c)] += (P[cx_pair(a, c)] - f * G[cx_pair(a, c)]) * rden - mean[a] * mean[c];
                            }
                        }
#pragma unroll
                        for (int k = 0; k < kCxNV; k++) G[k] = 0;
                    }
                }
#pragma unroll
                for (int k = 0; k < kCxNV; k++) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) ct[k] += __shfl_xor_sync(0xffffffffu, ct[k], o);
                }
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < kCxNV; k++) sh.contrib[warp][k] = ct[k];
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
            atomicAdd(&A.ctr->cox_evals, (unsigned long long)(sh.iter + 1));   // likelihood evaluations of this item
            if (sh.status == 2 || isnan(pw)) { pw = NA; atomicAdd(&A.ctr->n_wald_fail, 1ULL); }
            if (A.na_to_spa && isnan(pw)) pw = o[2 * M];                 // SPAGxEmix_CCT (ref :1190, :1249)
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
    const int i = perm[s];
    x[(int64_t)c * n_eff + s] = (i < 0) ? 0.0 : col[i] - means[c];
}

// ---- host side -----------------------------------------------------------------------------------------------------
struct CoxData {
    int64_t n = 0, n_eff = 0; int p = 0, unit_len = 0; bool ready = false;   // n_eff: length of the (padded) storage order
    int *d_perm = nullptr, *d_info = nullptr; double *d_x = nullptr, *d_means = nullptr;
    uint8_t *d_codes = nullptr; size_t cap_codes = 0;
};
CoxData *cox_create() { return new CoxData(); }
void cox_destroy(CoxData *c) {
    if (!c) return;
    cudaFree(c->d_perm); cudaFree(c->d_info); cudaFree(c->d_x); cudaFree(c->d_means); cudaFree(c->d_codes);
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
    // storage order (see k_cox_wald): unit u, step t, lane l  <-  sorted position u * unit_len + l * (unit_len / 32) + t
    {
        const int units = kCxCluster * kCxWarps;
        const int64_t unit_len = ((n_eff + units - 1) / units + 31) / 32 * 32, run = unit_len / 32, n_pad = unit_len * units;
        std::vector<int> perm_st((size_t)n_pad, -1), info_st((size_t)n_pad, 0);
        for (int64_t j = 0; j < n_pad; j++) {
            const int64_t u = j / unit_len, r = j % unit_len, t = r / 32, l = r % 32;
            const int64_t sp = u * unit_len + l * run + t;
            if (sp < n_eff) { perm_st[(size_t)j] = perm[(size_t)sp]; info_st[(size_t)j] = info[(size_t)sp] | (1 << 30); }
        }
        perm.swap(perm_st); info.swap(info_st);
        n_eff = n_pad; c->unit_len = (int)unit_len;
    }
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
                       double *out, Counters *ctr, int64_t *launches, bool na_to_spa) {
    if (!c || !c->ready) return cudaErrorInvalidValue;
    CoxArgs a;
    a.items = items; a.count = count; a.m_total = m_total; a.n = c->n; a.n_eff = c->n_eff;
    a.perm = c->d_perm; a.x = c->d_x; a.info = c->d_info; a.p = c->p; a.g_model = g_model;
    a.unit_len = c->unit_len;
    a.out = out; a.ctr = ctr; a.na_to_spa = na_to_spa ? 1 : 0;
    cudaLaunchConfig_t cfg = {};
    const int n_clusters = std::max(1, num_sms / kCxCluster);
    const size_t need = (size_t)n_clusters * (size_t)c->n_eff;
    if (need > c->cap_codes) {
        cudaError_t e0 = cudaStreamSynchronize(s);
        if (e0 != cudaSuccess) return e0;
        cudaFree(c->d_codes); c->d_codes = nullptr; c->cap_codes = 0;
        if ((e0 = cudaMalloc((void **)&c->d_codes, need)) != cudaSuccess) return e0;
        c->cap_codes = need;
    }
    a.codes = c->d_codes;
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
