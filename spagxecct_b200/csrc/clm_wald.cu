// clm_wald.cu -- Wald test of the E:g coefficient for traits = "categorical" in branch B of SPAGxE_CCT.
//
// Reference (R/SPAGxECCT.R:664-677):
//     data0 = clm(formula = as.factor(Phen.mtx$Y) ~ as.matrix(Cova.mtx) + E + g + g*E)
//     pval.wald = summary(data0)$coefficients["E:g", "Pr(>|z|)"]
// `ordinal` is a third-party package that the reference neither ships nor imports; what is restated is
// the model of that call -- logit link, flexible thresholds: P(Y <= j) = F(alpha_j - x'beta), j = 1 .. J-1 -- and its
// summary: z = estimate / sqrt(diag(solve(-Hessian))) at the maximum-likelihood estimate, p = 2 pnorm(-|z|).  The
// likelihood is strictly concave, so clm's Newton iteration (stop: max |gradient| < 1e-6) and the one below (stop: Newton
// step < 1e-10 (1 + max |theta|)) meet at the same estimate to ~1e-11.  Iteration: alpha_j = qlogis(j / J), beta = 0, step halving while the
// likelihood drops (the CPU test checker in tests/ walks the same path).
//
// Data layout (SNP-invariant, built once per analysis by clm_setup): samples ordered by category, every category's
// run padded to a multiple of 32, so that a warp-wide chunk of 32 samples has ONE category: which thresholds a
// sample touches is then warp-uniform, and its gradient / Hessian contributions go to a fixed set of accumulators
// (6 scalars, 3 q-vectors per category; the q x q block of the slopes is common to all categories).  E and the
// covariates are gathered into that order once; the item's genotype codes once per item.
//
// Kernel: one thread-block cluster (8 CTAs x 8 warps) per work item, persistent over the device-side item list.
// Chunks are dealt round-robin to the 64 warps of the cluster; a warp flushes its per-category accumulators when the
// category of its chunk changes (at most J times per evaluation).  Warp -> CTA -> cluster totals are added in a fixed
// order (reproducible results); every CTA redoes the scalar Newton logic on identical totals.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "clm_wald.h"
#include "dmath.cuh"
#include "snp_common.cuh"

namespace spagxe {
namespace cg = cooperative_groups;

constexpr int kClWarps = 8;
constexpr int kClThreads = kClWarps * 32;
constexpr int kClCluster = 8;
constexpr int kClQ = 5;                                  // design columns: E, g, E:g, cova[0], cova[1]
constexpr int kClNB = kClQ * (kClQ + 1) / 2;             // 15: slope block of the Hessian
constexpr int kClNC = 6 + 3 * kClQ;                      // 21 per-category sums
constexpr int kClMaxP = kClmMaxJ - 1 + kClQ;             // 12 parameters
constexpr int kClNTot = kClmMaxJ * kClNC + kClNB;        // 183 totals per evaluation
constexpr int kClMaxIter = 100;
constexpr double kClTolerChol = 1.818989403545856e-12;   // .Machine$double.eps^0.75

struct ClmArgs {
    const BItem *items; const int *count;
    int64_t m_total, n, n_pad;
    const int *perm;            // storage position -> sample index (-1: padding)
    const double *x;            // (p + 1) columns x n_pad: E, cova[0..p) gathered
    const uint8_t *chunk_cat;   // category of every 32-sample chunk
    int p, J, g_model, na_to_spa;
    double *out; Counters *ctr;
    uint8_t *codes;             // per cluster: the item's 2-bit codes in storage order, bit 2 = a real sample
};

struct ClShared {
    double cat[kClWarps][kClmMaxJ][kClNC];
    double bb[kClWarps][kClNB];
    double slot[kClNTot];                     // this CTA's totals, read by the other CTAs of the cluster
    double red[kClNTot];
    double th[kClMaxP], prev[kClMaxP];
    double ll_old, pw;
    int iter, halv, status, evals;            // status: 0 running, 1 done, 2 failed (singular Hessian / no convergence)
};

// LDL' in place (as cholesky2 of the survival package, cox_wald.cu): returns the rank, zeroes deficient pivots
__device__ int cl_cholesky(double (*m)[kClMaxP], int n, double toler) {
    double eps = 0;
    for (int i = 0; i < n; i++) { if (m[i][i] > eps) eps = m[i][i]; for (int j = i + 1; j < n; j++) m[j][i] = m[i][j]; }
    eps = (eps == 0) ? toler : eps * toler;
    int rank = 0;
    for (int i = 0; i < n; i++) {
        const double pivot = m[i][i];
        if (!isfinite(pivot) || pivot < eps) { m[i][i] = 0; continue; }
        rank++;
        for (int j = i + 1; j < n; j++) {
            const double t = m[j][i] / pivot;
            m[j][i] = t;
            m[j][j] -= t * t * pivot;
            for (int k = j + 1; k < n; k++) m[k][j] -= t * m[k][i];
        }
    }
    return rank;
}
__device__ void cl_solve(double (*m)[kClMaxP], int n, double *y) {
    for (int i = 0; i < n; i++) { double t = y[i]; for (int j = 0; j < i; j++) t -= y[j] * m[i][j]; y[i] = t; }
    for (int i = n - 1; i >= 0; i--) {
        if (m[i][i] == 0) { y[i] = 0; continue; }
        double t = y[i] / m[i][i];
        for (int j = i + 1; j < n; j++) t -= y[j] * m[j][i];
        y[i] = t;
    }
}

struct ClIn { int code; double e, c0, c1; };
__device__ __forceinline__ ClIn cl_load(const ClmArgs &A, const uint8_t *__restrict__ codes, int64_t s, int64_t s_hi) {
    ClIn in; in.code = 3; in.e = 0; in.c0 = 0; in.c1 = 0;
    if (s < s_hi) {
        in.code = __ldcg(codes + s);              // written by other SMs for this item: not through L1
        in.e = A.x[s];
        if (A.p > 0) in.c0 = A.x[A.n_pad + s];
        if (A.p > 1) in.c1 = A.x[2 * A.n_pad + s];
    }
    return in;
}

__global__ void __launch_bounds__(kClThreads, 1) k_clm_wald(ClmArgs A) {
    __shared__ ClShared sh;
    cg::cluster_group cl = cg::this_cluster();
    const int cs = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int cluster_id = blockIdx.x / cs, n_clusters = gridDim.x / cs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int J = A.J, nth = J - 1, q = 3 + A.p, np = nth + q;
    const int64_t n_chunks = A.n_pad / 32;
    const int gw = rank * kClWarps + warp, W = cs * kClWarps;
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
        uint8_t *codes = A.codes + (size_t)cluster_id * (size_t)A.n_pad;
        cl.sync();   // the previous item's sweeps are over in every CTA of the cluster
        for (int64_t s = (int64_t)rank * kClThreads + threadIdx.x; s < A.n_pad; s += (int64_t)cs * kClThreads) {
            const int i = A.perm[s];
            codes[s] = (i < 0) ? (uint8_t)3 : (uint8_t)(((row[i >> 2] >> (2 * (i & 3))) & 3) | 4);
        }
        __threadfence();
        if (threadIdx.x == 0) {
            for (int j = 0; j < nth; j++) { const double prb = (double)(j + 1) / (double)J; sh.th[j] = log(prb / (1 - prb)); }
            for (int a = 0; a < q; a++) sh.th[nth + a] = 0;
            for (int a = 0; a < kClMaxP; a++) sh.prev[a] = (a < np) ? sh.th[a] : 0;
            sh.ll_old = -CUDART_INF; sh.iter = 0; sh.halv = 0; sh.status = 0; sh.evals = 0; sh.pw = NA;
        }
        cl.sync();
        for (;;) {
            double b[kClQ];
#pragma unroll
            for (int k = 0; k < kClQ; k++) b[k] = (k < q) ? sh.th[nth + k] : 0.0;
            for (int k = lane; k < kClmMaxJ * kClNC; k += 32) (&sh.cat[warp][0][0])[k] = 0;
            __syncwarp();
            double acc[kClNC], bb[kClNB];
#pragma unroll
            for (int k = 0; k < kClNC; k++) acc[k] = 0;
#pragma unroll
            for (int k = 0; k < kClNB; k++) bb[k] = 0;
            int cur = -1;
            bool has_u = false, has_l = false;
            double thu = 0, thl = 0;
            auto flush = [&]() {
#pragma unroll
                for (int k = 0; k < kClNC; k++) {
                    double v = acc[k];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) sh.cat[warp][cur][k] += v;
                    acc[k] = 0;
                }
            };
            ClIn nxt = cl_load(A, codes, (int64_t)gw * 32 + lane, A.n_pad);
            for (int64_t ch = gw; ch < n_chunks; ch += W) {
                const ClIn in = nxt;
                nxt = cl_load(A, codes, (ch + W) * 32 + lane, A.n_pad);
                const int c = A.chunk_cat[ch];
                if (c != cur) {
                    if (cur >= 0) flush();
                    cur = c;
                    has_u = c < nth; has_l = c > 0;
                    thu = has_u ? sh.th[c] : 0.0;
                    thl = has_l ? sh.th[c - 1] : 0.0;
                }
                const bool valid = (in.code & 4) != 0;
                const int code = in.code & 3;
                const double g = (code & 1) ? ((code & 2) ? gval[3] : gval[1]) : ((code & 2) ? gval[2] : gval[0]);
                double x[kClQ];
                x[0] = in.e; x[1] = g; x[2] = in.e * g; x[3] = in.c0; x[4] = in.c1;
                double eta = 0;
#pragma unroll
                for (int k = 0; k < kClQ; k++) eta += b[k] * x[k];
                double Fu = 1, fu = 0, su = 0, Fl = 0, fl = 0, sl = 0;   // F, F' = F (1 - F), F'' = F' (1 - 2 F)
                if (has_u) { Fu = 1 / (1 + exp(-(thu - eta))); fu = Fu * (1 - Fu); su = fu * (1 - 2 * Fu); }
                if (has_l) { Fl = 1 / (1 + exp(-(thl - eta))); fl = Fl * (1 - Fl); sl = fl * (1 - 2 * Fl); }
                const double pi = Fu - Fl, rp = 1 / pi, d = (fu - fl) * rp;
                const double cu = fu * rp, cl_ = -fl * rp;
                // padding lanes (no sample behind them) contribute exact zeros
                acc[0] += valid ? log(pi) : 0.0;
                acc[1] += valid ? cu : 0.0;
                acc[2] += valid ? cl_ : 0.0;
                acc[3] += valid ? su * rp - cu * cu : 0.0;
                acc[4] += valid ? -sl * rp - cl_ * cl_ : 0.0;
                acc[5] += valid ? fu * fl * rp * rp : 0.0;
                const double dm = valid ? d : 0.0, hub = valid ? -su * rp + cu * d : 0.0, hlb = valid ? sl * rp + cl_ * d : 0.0,
                             hbb = valid ? (su - sl) * rp - d * d : 0.0;
#pragma unroll
                for (int a = 0; a < kClQ; a++) {
                    acc[6 + a] += dm * x[a];
                    acc[6 + kClQ + a] += hub * x[a];
                    acc[6 + 2 * kClQ + a] += hlb * x[a];
                    const double ha = hbb * x[a];
#pragma unroll
                    for (int c2 = 0; c2 <= a; c2++) bb[a * (a + 1) / 2 + c2] += ha * x[c2];
                }
            }
            if (cur >= 0) flush();
#pragma unroll
            for (int k = 0; k < kClNB; k++) {
                double v = bb[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) sh.bb[warp][k] = v;
            }
            __syncthreads();
            for (int k = threadIdx.x; k < kClNTot; k += kClThreads) {
                double t = 0;
                if (k < kClmMaxJ * kClNC) { for (int w = 0; w < kClWarps; w++) t += (&sh.cat[w][0][0])[k]; }
                else { for (int w = 0; w < kClWarps; w++) t += sh.bb[w][k - kClmMaxJ * kClNC]; }
                sh.slot[k] = t;
            }
            cl.sync();
            for (int k = threadIdx.x; k < kClNTot; k += kClThreads) {
                double t = 0;
                for (int r = 0; r < cs; r++) t += cl.map_shared_rank(sh.slot, r)[k];
                sh.red[k] = t;
            }
            cl.sync();   // the exchange slots may be rewritten from here on
            // ---------------- Newton-Raphson bookkeeping (every CTA, identical totals) ----------------
            if (threadIdx.x == 0) {
                double H[kClMaxP][kClMaxP], gr[kClMaxP];
                for (int a = 0; a < np; a++) { gr[a] = 0; for (int c2 = 0; c2 < np; c2++) H[a][c2] = 0; }
                double ll = 0;
                for (int c = 0; c < J; c++) {
                    const double *t = &sh.red[c * kClNC];
                    const int u = (c < nth) ? c : -1, l = (c > 0) ? c - 1 : -1;
                    ll += t[0];
                    if (u >= 0) { gr[u] += t[1]; H[u][u] += t[3]; }
                    if (l >= 0) { gr[l] += t[2]; H[l][l] += t[4]; }
                    if (u >= 0 && l >= 0) { H[u][l] += t[5]; H[l][u] += t[5]; }
                    for (int a = 0; a < q; a++) {
                        gr[nth + a] -= t[6 + a];
                        if (u >= 0) { H[u][nth + a] += t[6 + kClQ + a]; H[nth + a][u] += t[6 + kClQ + a]; }
                        if (l >= 0) { H[l][nth + a] += t[6 + 2 * kClQ + a]; H[nth + a][l] += t[6 + 2 * kClQ + a]; }
                    }
                }
                for (int a = 0; a < q; a++)
                    for (int c2 = 0; c2 <= a; c2++) {
                        const double v = sh.red[kClmMaxJ * kClNC + a * (a + 1) / 2 + c2];
                        H[nth + a][nth + c2] = v; H[nth + c2][nth + a] = v;
                    }
                sh.evals++;
                if (sh.iter > 0 && !(ll >= sh.ll_old) && sh.halv < 30) {   // the step overshot: halve it
                    for (int a = 0; a < np; a++) sh.th[a] = (sh.th[a] + sh.prev[a]) / 2;
                    sh.halv++;
                } else {
                    sh.halv = 0;
                    for (int a = 0; a < np; a++) for (int c2 = 0; c2 < np; c2++) H[a][c2] = -H[a][c2];
                    if (cl_cholesky(H, np, kClTolerChol) < np) sh.status = 2;
                    else {
                        cl_solve(H, np, gr);      // gr is now the Newton step
                        double ms = 0, mt = 0;
                        for (int a = 0; a < np; a++) { ms = fmax(ms, fabs(gr[a])); mt = fmax(mt, fabs(sh.th[a])); }
                        if (ms < 1e-10 * (1 + mt)) {
                            double e[kClMaxP];
                            const int k = nth + 2;                           // "E:g" is design column 2 here
                            for (int a = 0; a < np; a++) e[a] = (a == k) ? 1.0 : 0.0;
                            cl_solve(H, np, e);
                            const double z = sh.th[k] / sqrt(e[k]);
                            sh.pw = 2 * d_pnorm(-fabs(z), true);
                            sh.status = 1;
                        } else if (sh.iter + 1 >= kClMaxIter) sh.status = 2;
                        else {
                            for (int a = 0; a < np; a++) { sh.prev[a] = sh.th[a]; sh.th[a] += gr[a]; }
                            sh.ll_old = ll;
                            sh.iter++;
                        }
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
            atomicAdd(&A.ctr->clm_evals, (unsigned long long)sh.evals);
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

__global__ void k_clm_gather(const double *__restrict__ E, const double *__restrict__ cova, int64_t n, int64_t n_pad,
                             const int *__restrict__ perm, double *__restrict__ x) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n_pad) return;
    const int c = blockIdx.y;
    const double *col = (c == 0) ? E : cova + (int64_t)(c - 1) * n;
    const int i = perm[s];
    x[(int64_t)c * n_pad + s] = (i < 0) ? 0.0 : col[i];
}

// ---- host side -----------------------------------------------------------------------------------------------------
struct ClmData {
    int64_t n = 0, n_pad = 0; int p = 0, J = 0; bool ready = false;
    int *d_perm = nullptr; double *d_x = nullptr; uint8_t *d_chunk_cat = nullptr;
    uint8_t *d_codes = nullptr; size_t cap_codes = 0;
};
ClmData *clm_create() { return new ClmData(); }
void clm_destroy(ClmData *c) {
    if (!c) return;
    cudaFree(c->d_perm); cudaFree(c->d_x); cudaFree(c->d_chunk_cat); cudaFree(c->d_codes);
    delete c;
}
void clm_invalidate(ClmData *c) { if (c) c->ready = false; }
bool clm_ready(const ClmData *c) { return c && c->ready; }

cudaError_t clm_setup(ClmData *c, cudaStream_t s, int64_t n, const double *h_ycat, const double *d_E, const double *d_cova, int p,
                      const char **what) {
    *what = "clm_setup";
    c->ready = false;
    if (p > kClmMaxCov) { *what = "categorical Wald: at most 2 covariate columns"; return cudaErrorInvalidValue; }
    int J = 0;
    std::vector<int64_t> cnt(kClmMaxJ, 0);
    for (int64_t i = 0; i < n; i++) {
        const double y = h_ycat[i];
        if (!(y >= 0 && y < kClmMaxJ) || y != std::floor(y)) { *what = "categorical Wald: Y must hold level indices 0 .. J-1, J <= 8"; return cudaErrorInvalidValue; }
        cnt[(size_t)y]++;
        J = std::max(J, (int)y + 1);
    }
    if (J < 2) { *what = "categorical Wald: fewer than two levels"; return cudaErrorInvalidValue; }
    for (int k = 0; k < J; k++)
        if (cnt[(size_t)k] == 0) { *what = "categorical Wald: a level index without samples (as.factor has no empty levels)"; return cudaErrorInvalidValue; }
    std::vector<int64_t> start(J + 1, 0);
    for (int k = 0; k < J; k++) start[(size_t)k + 1] = start[(size_t)k] + (cnt[(size_t)k] + 31) / 32 * 32;
    const int64_t n_pad = start[(size_t)J];
    std::vector<int> perm((size_t)n_pad, -1);
    std::vector<uint8_t> chunk_cat((size_t)(n_pad / 32));
    {
        std::vector<int64_t> cur(start.begin(), start.end() - 1);
        for (int64_t i = 0; i < n; i++) perm[(size_t)cur[(size_t)h_ycat[i]]++] = (int)i;
        for (int k = 0; k < J; k++)
            for (int64_t ch = start[(size_t)k] / 32; ch < start[(size_t)k + 1] / 32; ch++) chunk_cat[(size_t)ch] = (uint8_t)k;
    }
    cudaError_t err;
    cudaFree(c->d_perm); cudaFree(c->d_x); cudaFree(c->d_chunk_cat);
    c->d_perm = nullptr; c->d_x = nullptr; c->d_chunk_cat = nullptr;
    if ((err = cudaMalloc((void **)&c->d_perm, sizeof(int) * (size_t)n_pad)) != cudaSuccess) return err;
    if ((err = cudaMalloc((void **)&c->d_x, sizeof(double) * (size_t)n_pad * (p + 1))) != cudaSuccess) return err;
    if ((err = cudaMalloc((void **)&c->d_chunk_cat, chunk_cat.size())) != cudaSuccess) return err;
    if ((err = cudaMemcpyAsync(c->d_perm, perm.data(), sizeof(int) * (size_t)n_pad, cudaMemcpyHostToDevice, s)) != cudaSuccess) return err;
    if ((err = cudaMemcpyAsync(c->d_chunk_cat, chunk_cat.data(), chunk_cat.size(), cudaMemcpyHostToDevice, s)) != cudaSuccess) return err;
    k_clm_gather<<<dim3((unsigned)((n_pad + 255) / 256), (unsigned)(p + 1)), 256, 0, s>>>(d_E, d_cova, n, n_pad, c->d_perm, c->d_x);
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    if ((err = cudaStreamSynchronize(s)) != cudaSuccess) return err;   // perm / chunk_cat are host vectors of this frame
    c->n = n; c->n_pad = n_pad; c->p = p; c->J = J; c->ready = true;
    return cudaSuccess;
}

cudaError_t clm_launch(ClmData *c, cudaStream_t s, int num_sms, const BItem *items, const int *count, int64_t m_total, int g_model,
                       double *out, Counters *ctr, int64_t *launches, bool na_to_spa) {
    if (!c || !c->ready) return cudaErrorInvalidValue;
    ClmArgs a;
    a.items = items; a.count = count; a.m_total = m_total; a.n = c->n; a.n_pad = c->n_pad;
    a.perm = c->d_perm; a.x = c->d_x; a.chunk_cat = c->d_chunk_cat; a.p = c->p; a.J = c->J; a.g_model = g_model;
    a.out = out; a.ctr = ctr; a.na_to_spa = na_to_spa ? 1 : 0;
    const int n_clusters = std::max(1, num_sms / kClCluster);
    const size_t need = (size_t)n_clusters * (size_t)c->n_pad;
    if (need > c->cap_codes) {
        cudaError_t e0 = cudaStreamSynchronize(s);
        if (e0 != cudaSuccess) return e0;
        cudaFree(c->d_codes); c->d_codes = nullptr; c->cap_codes = 0;
        if ((e0 = cudaMalloc((void **)&c->d_codes, need)) != cudaSuccess) return e0;
        c->cap_codes = need;
    }
    a.codes = c->d_codes;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_clusters * kClCluster));
    cfg.blockDim = dim3(kClThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kClCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_clm_wald, a);
    if (e == cudaSuccess) (*launches)++;
    return e;
}

}  // namespace spagxe
