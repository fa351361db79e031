// bb_farm.cu -- K5: the SPAGxE_CCT branch-B solver farm (see bb_farm.h for the design).
//
// Reference path per work item (R/SPAGxECCT.R):
//   :606-612  W = [1, g]; R0 = R - W (W'W)^-1 W'R; R.new0 = E * R0
//   :616      SPA_G_one_SNP_homo_new(g, R.new0, min.maf)      -> :1991-2065, :1795-1879 (Newton + tail)
//   :636-662  glm(Y ~ Cova + E + g + E:g, binomial) / lm(...)   -> p of "E:g"
//   :629-671  CCT(c(p.spa, p.wald))                             -> :2070-2123
//
// Execution model: cooperative persistent kernel, grid = one CTA per SM, 16 warps per CTA.
//   * CTA c owns samples [c * len, (c+1) * len); R, E, Y and the covariate columns of that slice live in shared memory
//     for the whole kernel (135 KB at N = 500,000, p = 2), so the per-sample operands of every pass come from shared
//     memory; only the 2-bit genotype bytes of the item are read from global memory (L1-resident after the first pass).
//   * A "solver" is a per-item iteration-synchronous state machine: the Wald fit (IRLS or lm) and the two SPA tails.
//     Per farm iteration, a unit = (active solver, half of the CTA's slice) is swept by one warp; lanes keep the whole
//     upper triangle of X'WX in registers (designs of <= 6 columns) or share it entry-wise through a per-warp tile.
//   * Unit partial sums go to global memory; after the grid barrier every CTA adds them in the same fixed order and
//     redoes the scalar update (Cholesky solve, Newton step, convergence tests), so all CTAs hold identical state and
//     agree on termination without another exchange.  Results are bit-reproducible for a given (N, SM count).
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>

#include "bb_farm.h"
#include "dmath.cuh"
#include "snp_common.cuh"
#include "stat_common.cuh"

namespace spagxe {

constexpr int kFWarps = 16;
constexpr int kFThreads = kFWarps * 32;
constexpr int kFJ = 2;             // units per (solver, CTA): halves of the CTA's slice
constexpr int kFBatch = 128;       // items advanced in lock step (a 65,536-SNP call at epsilon = 0.001 holds 66 +- 8 branch-B SNPs: one batch)
constexpr int kFRegQ = 6;          // widest design with register accumulators
constexpr int kFMaxIt = 25;        // glm.fit maxit

struct FItem {   // per-item state in shared memory (identical in every CTA)
    const uint8_t *row; long long snp;
    double gval[4], fit[4];
    double beta[kMaxQ];
    double devold, se2, bl, pw;
    double mafi, S, Smean, z, pnorm_, pspa;
    double tail_s[2], tail_t[2], tail_H1[2], tail_diff[2], tail_p[2];
    double D0, T0, sum_g, u;             // u: imputed dosage 2 * MAF_raw before the flip
    double H0;                           // sum of R over hom-A1 samples (G.model Dom / Rec only)
    int tail_iter[2], tail_phase[2];     // phase: 0 Newton, 1 final evaluation, 2 done
    int w_it, w_status;                  // Wald: 0 active, 1 done, 2 rank deficient, 3 invalid
    int spa_kind;                        // homo_scalar: 0 normal, 1 SPA, 2 inner min.maf NA, 3 singular W'W
    int asum, nmiss, flip;
    int prep, pad_;                      // preparation stage: 0 class counts pending, 1 adjusted-residual sums pending, 2 done
};

static_assert(sizeof(FItem) % 8 == 0, "item states are published as 8-byte words");

struct FGeom {
    int64_t n, len;          // samples, samples per CTA slice (multiple of 4)
    int q, p, trait;
    int ncols;               // columns: R, E, Y, cova[0..p)
    int n_staged;            // the first n_staged of them live in shared memory (all of them in the BASELINE configs)
    int ld;                  // 8 or 16: leading dimension of the Cholesky matrix of wide designs
    int kw;                  // doubles per partial row
    int rows;                // partial rows per solver = gridDim.x * kFJ
    int tile_stride;         // generic path: doubles per tile row
};

// ---- grid barrier (cooperative launch guarantees co-residency) -----------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned &phase) {
    __syncthreads();
    if (threadIdx.x == 0) {
        phase++;
        const unsigned target = phase * gridDim.x;
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned v;
        unsigned spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (++spins > (1u << 24)) __trap();   // a protocol bug must fault (after a few seconds), not hang the GPU
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// ---- column access ---------------------------------------------------------------------------------------------------
template <bool ALL>   // ALL: every column is staged (the BASELINE configs); the accessor is then a plain shared-memory load
struct ColsT {   // column c of the CTA's slice: 0 R, 1 E, 2 Y, 3.. covariates; i = index inside the slice
    const double *s; int stride; int n_staged;              // shared-memory copy of the first n_staged columns
    const double *R, *E, *Y, *cova; int64_t n, g0;          // global fall-back for what did not fit
    __device__ __forceinline__ double get(int c, int i) const {
        if (ALL || c < n_staged) return s[c * stride + i];
        const double *src = (c == 0) ? R : (c == 1) ? E : (c == 2) ? Y : cova + (int64_t)(c - 3) * n;
        return __ldg(src + g0 + i);
    }
};

__device__ __forceinline__ int code_at(const uint8_t *__restrict__ row, int64_t i) {
    return (__ldg(row + (i >> 2)) >> (2 * (int)(i & 3))) & 3;
}

// warp butterfly: every lane ends with the total (fixed tree => deterministic)
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- sweeps (one warp, samples [a, b) of the CTA's slice, lanes stride 32) ---------------------------------------------
// class counts for sum(g^2): n00 and n10 (exact integers)
__device__ void sweep_counts(const FItem &it, int64_t g0, int a, int b, double *prow, int lane) {
    int c00 = 0, c10 = 0;
    for (int i = a + lane; i < b; i += 32) {
        const int c = code_at(it.row, g0 + i);
        c00 += (c == 0); c10 += (c == 2);
    }
    c00 = __reduce_add_sync(0xffffffffu, c00); c10 = __reduce_add_sync(0xffffffffu, c10);
    if (lane == 0) { prow[0] = (double)c00; prow[1] = (double)c10; }
}

// S = sum(g * rn), sum(rn), sum(rn^2) with rn = E (R - fit[class])   (ref :609-612, :2026-2037)
template <class CT>
__device__ void sweep_adj(const CT &cols, const FItem &it, int64_t g0, int a, int b, double *prow, int lane) {
    double v0 = 0, v1 = 0, v2 = 0;
    for (int i = a + lane; i < b; i += 32) {
        const int c = code_at(it.row, g0 + i);
        const double rn = cols.get(1, i) * (cols.get(0, i) - it.fit[c]);
        v0 += it.gval[c] * rn; v1 += rn; v2 += rn * rn;
    }
    v0 = wsum(v0); v1 = wsum(v1); v2 = wsum(v2);
    if (lane == 0) { prow[0] = v0; prow[1] = v1; prow[2] = v2; }
}

// one CGF evaluation of a SPA tail: Newton passes need sum r K1(t r), sum r^2 K2(t r); the final pass sum K0, sum r^2 K2
template <bool FINAL, class CT>
__device__ void sweep_tail(const CT &cols, const FItem &it, int h, int64_t g0, int a, int b, double *prow, int lane) {
    const double t = it.tail_t[h], maf = it.mafi;
    double v0 = 0, v1 = 0;
#pragma unroll 2
    for (int i = a + lane; i < b; i += 32) {
        const int c = code_at(it.row, g0 + i);
        const double r = cols.get(1, i) * (cols.get(0, i) - it.fit[c]);
        double k1, k2;
        cgf_k12(t * r, maf, k1, k2);
        if (FINAL) v0 += cgf_k0(t * r, maf); else v0 += r * k1;
        v1 += (r * r) * k2;
    }
    v0 = wsum(v0); v1 = wsum(v1);
    if (lane == 0) { prow[0] = v0; prow[1] = v1; }
}

// Wald pass with register accumulators: X = [1, cova(P), E, g, E g], Q = P + 4 columns.
// MODE 0: glm.fit first pass (eta from mustart), 1: IRLS pass at beta, 2: lm normal equations, 3: lm residual sum of squares.
template <int Q, int MODE, class CT>
__device__ void sweep_wald_reg(const CT &cols, const FItem &it, int64_t g0, int a, int b, double *prow, int lane) {
    constexpr int P = Q - 4;
    constexpr int NENT = Q * (Q + 3) / 2;
    constexpr int U = 4;
    double acc[NENT];
#pragma unroll
    for (int e = 0; e < NENT; e++) acc[e] = 0;
    double beta[Q];
#pragma unroll
    for (int c = 0; c < Q; c++) beta[c] = it.beta[c];
    DevProd dp;
    double dev = 0;
    int bad = 0;
    // glm.fit's first pass on a 0/1 response: eta = logit((y + 0.5) / 2) takes two values
    double p0_eta[2] = {0, 0}, p0_ope[2] = {1, 1}, p0_w2[2] = {0, 0}, p0_wz[2] = {0, 0};
    if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double y = (double)k, mus = (y + 0.5) / 2, e0 = log(mus / (1 - mus)), ee = exp(e0);
            const double opexp = 1 + ee, r = 1.0 / opexp, mu = ee * r, mev = mu * r;
            p0_eta[k] = e0; p0_ope[k] = opexp; p0_w2[k] = mev; p0_wz[k] = fma(mev, e0, y - mu);
        }
    }
    auto accumulate = [&](const double (&x)[Q], double w2, double wz) {
        int e = 0;
#pragma unroll
        for (int r = 0; r < Q; r++) {
            const double wa = (r == 0) ? w2 : w2 * x[r];
#pragma unroll
            for (int c = r; c < Q; c++) { acc[e] = fma(wa, x[c], acc[e]); e++; }
            acc[e] = fma(x[r], wz, acc[e]); e++;
        }
    };
    auto load = [&](int i, double (&x)[Q], double &y) {
        const int c = code_at(it.row, g0 + i);
        const double g = it.gval[c], ev = cols.get(1, i);
        y = cols.get(2, i);
        x[0] = 1.0;
#pragma unroll
        for (int k = 0; k < P; k++) x[1 + k] = cols.get(3 + k, i);
        x[P + 1] = ev; x[P + 2] = g; x[P + 3] = ev * g;
    };
    auto one = [&](const double (&x)[Q], double y) {   // any sample through the general code
        double eta = 0;
        if (MODE == 1 || MODE == 3) {
#pragma unroll
            for (int c = 0; c < Q; c++) eta = fma(x[c], beta[c], eta);
        }
        if (MODE == 3) { dev += (y - eta) * (y - eta); return; }
        double w2, wz;
        irls_sample(MODE <= 1 ? 1 : 2, MODE == 0 || MODE == 2, y, eta, w2, wz, dev, bad, dp);
        accumulate(x, w2, wz);
    };
    int i = a + lane;
    if (MODE <= 1) {
        // U samples per trip.  Phase 1 builds eta (the design row is consumed as it is loaded), phase 2 runs the U
        // exp / reciprocal chains side by side, phase 3 re-reads each row from shared memory and accumulates: the rows
        // are not kept live across the chains, which is what lets four samples fit into 128 registers.
        for (; i + 32 * (U - 1) < b; i += 32 * U) {
            double ys[U], gs[U], eta[U];
            bool fast = true;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int ii = i + 32 * u;
                const int c = code_at(it.row, g0 + ii);
                gs[u] = it.gval[c];
                ys[u] = cols.get(2, ii);
                eta[u] = 0;
                if (MODE == 1) {
                    const double ev = cols.get(1, ii);
                    double e = beta[0];
#pragma unroll
                    for (int k = 0; k < P; k++) e = fma(cols.get(3 + k, ii), beta[1 + k], e);
                    e = fma(ev, beta[P + 1], e);
                    e = fma(gs[u], beta[P + 2], e);
                    e = fma(ev * gs[u], beta[P + 3], e);
                    eta[u] = e;
                }
                fast = fast && (ys[u] == 0.0 || ys[u] == 1.0) && !(eta[u] > 30. || eta[u] < -30.);
            }
            if (fast) {
                double w2[U], wz[U], ope[U];
#pragma unroll
                for (int u = 0; u < U; u++) {   // same operations as the common branch of irls_sample
                    if (MODE == 0) {
                        const bool is1 = (ys[u] == 1.0);
                        eta[u] = is1 ? p0_eta[1] : p0_eta[0]; ope[u] = is1 ? p0_ope[1] : p0_ope[0];
                        w2[u] = is1 ? p0_w2[1] : p0_w2[0];    wz[u] = is1 ? p0_wz[1] : p0_wz[0];
                    } else {
                        const double ee = d_exp_fast(eta[u]);
                        ope[u] = 1 + ee;
                        const double r = d_rcp_fast(ope[u]);
                        const double mu = ee * r;
                        w2[u] = mu * r;
                        wz[u] = fma(w2[u], eta[u], ys[u] - mu);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int ii = i + 32 * u;
                    dp.mul(ope[u]);
                    dev -= 2 * (ys[u] * eta[u]);
                    double x[Q];
                    const double ev = cols.get(1, ii);
                    x[0] = 1.0;
#pragma unroll
                    for (int k = 0; k < P; k++) x[1 + k] = cols.get(3 + k, ii);
                    x[P + 1] = ev; x[P + 2] = gs[u]; x[P + 3] = ev * gs[u];
                    accumulate(x, w2[u], wz[u]);
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; u++) {
                    double x[Q], y;
                    load(i + 32 * u, x, y);
                    one(x, y);
                }
            }
        }
    }
    for (; i < b; i += 32) {
        double x[Q], y;
        load(i, x, y);
        one(x, y);
    }
    if (MODE <= 1) dev += 2 * dp.log_value();
    // warp totals by a transposing butterfly: at every step a lane hands over the half of the vector its partner keeps,
    // so 31 exchanges instead of 5 per entry; lane e ends with the total of entry e (fixed tree => deterministic)
    static_assert(NENT + 2 <= 32, "one entry per lane");
    bad = __reduce_add_sync(0xffffffffu, bad);
    double v[32];
#pragma unroll
    for (int e = 0; e < 32; e++) v[e] = (e < NENT) ? acc[e] : 0.0;
    v[NENT] = dev;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int k = 0; k < half; k++) {
            const double send = up ? v[k] : v[k + half];
            const double keep = up ? v[k + half] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    if (lane <= NENT) prow[lane] = v[0];
    if (lane == NENT + 1) prow[lane] = (double)bad;
}

// Wald pass for wide designs (7..16 columns): the warp stages 32 samples' rows [x, w x, wz] in its tile, then lane L
// accumulates the normal-equation entries L, L + 32, ... over those 32 samples.
template <int MODE, class CT>
__device__ void sweep_wald_tile(const CT &cols, const FItem &it, const FGeom &gm, int64_t g0, int a, int b, double *prow,
                                double *tile, int lane) {
    const int q = gm.q, p = gm.p, stride = gm.tile_stride;
    const int nent = q * (q + 3) / 2;
    constexpr int NS = (kMaxQ * (kMaxQ + 3) / 2 + 31) / 32;   // 5 entries per lane at most
    int ea[NS], eb[NS];
    double acc[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) {
        acc[k] = 0; ea[k] = -1; eb[k] = 0;
        const int e = lane + 32 * k;
        if (e < nent) {
            int r = 0, rem = e;
            for (r = 0; r < q; r++) { const int ln = q + 1 - r; if (rem < ln) break; rem -= ln; }
            ea[k] = r; eb[k] = r + rem;        // eb == q: the rhs column (wz)
        }
    }
    DevProd dp;
    double dev = 0;
    int bad = 0;
    for (int base = a; base < b; base += 32) {
        const int i = base + lane;
        const int cnt = min(32, b - base);
        __syncwarp();
        if (i < b) {
            double *xr = tile + (size_t)lane * stride;      // [0, q): x   [q]: wz   [q+1, 2q+1): w x
            const int c = code_at(it.row, g0 + i);
            const double g = it.gval[c], ev = cols.get(1, i), y = cols.get(2, i);
            xr[0] = 1.0;
            for (int k = 0; k < p; k++) xr[1 + k] = cols.get(3 + k, i);
            xr[p + 1] = ev; xr[p + 2] = g; xr[p + 3] = ev * g;
            double eta = 0;
            if (MODE == 1 || MODE == 3) for (int k = 0; k < q; k++) eta = fma(xr[k], it.beta[k], eta);
            double w2 = 0, wz = 0;
            if (MODE == 3) dev += (y - eta) * (y - eta);
            else irls_sample(MODE <= 1 ? 1 : 2, MODE == 0 || MODE == 2, y, eta, w2, wz, dev, bad, dp);
            xr[q] = wz;
            for (int k = 0; k < q; k++) xr[q + 1 + k] = w2 * xr[k];
        }
        __syncwarp();
        if (MODE != 3) {
            for (int s = 0; s < cnt; s++) {
                const double *xr = tile + (size_t)s * stride;
#pragma unroll
                for (int k = 0; k < NS; k++)
                    if (ea[k] >= 0) acc[k] = fma(eb[k] == q ? xr[ea[k]] : xr[q + 1 + ea[k]], xr[eb[k]], acc[k]);
            }
        }
    }
    if (MODE <= 1) dev += 2 * dp.log_value();
#pragma unroll
    for (int k = 0; k < NS; k++) { const int e = lane + 32 * k; if (e < nent) prow[e] = acc[k]; }
    dev = wsum(dev);
    bad = __reduce_add_sync(0xffffffffu, bad);
    if (lane == 0) { prow[nent] = dev; prow[nent + 1] = (double)bad; }
}

// ---- scalar updates (one warp per solver; lane 0 does the serial part) ---------------------------------------------------
// Solve the equilibrated SPD system in place.  A: q x q row-major (ld q), rhs b; returns false when a pivot collapses.
template <int LD>
__device__ bool chol_solve_small(double (*A)[LD], double *b, int q, double *x, double *inv_last) {
    double sc[LD], y[LD];
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
    *inv_last = sc[q - 1] * sc[q - 1] / (l * l);
    return true;
}

// Add the partial rows of one solver: lane l adds rows l, l + 32, ... (each row is a contiguous, 16-byte aligned run of
// kw >= NV doubles), then a butterfly over the lanes; every lane ends with all NV totals.  The order is fixed.  All NV
// entries of a row are read without guards (the loads of a row are in flight together); entries the caller does not
// use hold whatever the row holds there.
template <int NV>
__device__ __forceinline__ void reduce_rows(const double *__restrict__ part, int rows, int kw, int lane, double (&tot)[NV]) {
    static_assert(NV % 2 == 0, "rows are read as double2");
#pragma unroll
    for (int k = 0; k < NV; k++) tot[k] = 0;
    for (int r = lane; r < rows; r += 32) {
        const double2 *pr = reinterpret_cast<const double2 *>(part + (size_t)r * kw);
        double2 v[NV / 2];
#pragma unroll
        for (int k = 0; k < NV / 2; k++) v[k] = __ldcg(pr + k);
#pragma unroll
        for (int k = 0; k < NV / 2; k++) { tot[2 * k] += v[k].x; tot[2 * k + 1] += v[k].y; }
    }
#pragma unroll
    for (int k = 0; k < NV; k++) tot[k] = wsum(tot[k]);
}

// The same solve with every index known at compile time (Q <= 6): the matrix lives in registers.
// ent: upper triangle row by row, each row followed by its right-hand side (the order the sweeps accumulate in).
template <int Q>
__device__ bool chol_solve_reg(const double *ent, double *x_out, double *inv_last) {
    double A[Q][Q], b[Q], sc[Q], y[Q], x[Q];
    {
        int e = 0;
#pragma unroll
        for (int r = 0; r < Q; r++) {
#pragma unroll
            for (int c = r; c < Q; c++) { A[r][c] = ent[e]; A[c][r] = ent[e]; e++; }
            b[r] = ent[e]; e++;
        }
    }
    bool ok = true;
#pragma unroll
    for (int i = 0; i < Q; i++) { if (!(A[i][i] > 0)) ok = false; sc[i] = 1.0 / sqrt(A[i][i]); }
    if (!ok) return false;
#pragma unroll
    for (int i = 0; i < Q; i++)
#pragma unroll
        for (int j = 0; j < Q; j++) A[i][j] *= sc[i] * sc[j];
#pragma unroll
    for (int j = 0; j < Q; j++) {
        double d = A[j][j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= A[j][k] * A[j][k];
        if (!(d > 1e-13)) ok = false;
        d = sqrt(d); A[j][j] = d;
#pragma unroll
        for (int i = j + 1; i < Q; i++) {
            double sacc = A[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) sacc -= A[i][k] * A[j][k];
            A[i][j] = sacc / d;
        }
    }
    if (!ok) return false;
#pragma unroll
    for (int i = 0; i < Q; i++) {
        double sacc = b[i] * sc[i];
#pragma unroll
        for (int k = 0; k < i; k++) sacc -= A[i][k] * y[k];
        y[i] = sacc / A[i][i];
    }
#pragma unroll
    for (int i = Q - 1; i >= 0; i--) {
        double sacc = y[i];
#pragma unroll
        for (int k = i + 1; k < Q; k++) sacc -= A[k][i] * x[k];
        x[i] = sacc / A[i][i];
    }
#pragma unroll
    for (int i = 0; i < Q; i++) x_out[i] = x[i] * sc[i];
    const double l = A[Q - 1][Q - 1];
    *inv_last = sc[Q - 1] * sc[Q - 1] / (l * l);
    return true;
}

// CTA-cooperative sum of one solver's partial rows: thread (g, e) = (tid / 32, tid % 32) adds entry e of the rows of
// group g (rows split into 16 contiguous groups), the 16 group sums meet in shared memory and are added in group order.
// Every thread of the CTA must call it; out[0 .. nval) is valid after it returns.  Fixed order => same bits everywhere.
__device__ void cta_reduce_rows(const double *__restrict__ part, int rows, int kw, int nval, double *s_red, double *out) {
    const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int per = (rows + kFWarps - 1) / kFWarps;
    const int r0 = min(rows, g * per), r1 = min(rows, r0 + per);
    for (int e0 = 0; e0 < nval; e0 += 32) {
        double acc = 0;
        const double *p = part + e0 + e;
#pragma unroll 8
        for (int r = r0; r < r1; r++) acc += __ldcg(p + (size_t)r * kw);
        s_red[g * 32 + e] = acc;
        __syncthreads();
        if (g == 0 && e0 + e < nval) {
            double t = 0;
#pragma unroll
            for (int k = 0; k < kFWarps; k++) t += s_red[k * 32 + e];
            out[e0 + e] = t;
        }
        __syncthreads();
    }
}

// Wald state machine step after sweep number it.w_it (ref: glm.fit / summary.glm, lm / summary.lm)
__device__ void update_wald(FItem &it, const FGeom &gm, double *wscr, int lane) {   // wscr[0 .. nent + 2): the totals
    const int q = gm.q, nent = q * (q + 3) / 2;
    if (lane == 0) {
        const double dev = wscr[nent];
        const bool bad = wscr[nent + 1] > 0;
        const int sweep = it.w_it;
        bool need_solve = true;
        if (gm.trait == 1) {
            if (sweep == 0) it.devold = dev;
            else {
                if (!isfinite(dev) || bad) { it.w_status = 3; need_solve = false; }
                else if (fabs(dev - it.devold) / (fabs(dev) + 0.1) < 1e-8 || sweep == kFMaxIt) { it.w_status = 1; need_solve = false; }
                else it.devold = dev;
            }
        } else if (sweep == 1) {   // lm: residual sum of squares is in; finish
            const double rdf = (double)(gm.n - q);
            const double se = sqrt(it.se2 * (dev / rdf));
            it.pw = d_pt2(it.bl / se, rdf);
            it.w_status = 1; need_solve = false;
        }
        if (need_solve) {
            double rhs[kMaxQ], x[kMaxQ], il = 0;
            bool ok;
            if (q == 6) ok = chol_solve_reg<6>(wscr, x, &il);
            else if (q == 5) ok = chol_solve_reg<5>(wscr, x, &il);
            else if (q == 4) ok = chol_solve_reg<4>(wscr, x, &il);
            else if (gm.ld == 8) {
                double (*A)[8] = reinterpret_cast<double (*)[8]>(wscr + 160);
                int e = 0;
                for (int r = 0; r < q; r++) {
                    for (int c = r; c < q; c++) { A[r][c] = wscr[e]; A[c][r] = wscr[e]; e++; }
                    rhs[r] = wscr[e]; e++;
                }
                ok = chol_solve_small<8>(A, rhs, q, x, &il);
            } else {
                double (*A)[kMaxQ] = reinterpret_cast<double (*)[kMaxQ]>(wscr + 160);
                int e = 0;
                for (int r = 0; r < q; r++) {
                    for (int c = r; c < q; c++) { A[r][c] = wscr[e]; A[c][r] = wscr[e]; e++; }
                    rhs[r] = wscr[e]; e++;
                }
                ok = chol_solve_small<kMaxQ>(A, rhs, q, x, &il);
            }
            if (!ok) it.w_status = 2;
            else {
                for (int c = 0; c < q; c++) it.beta[c] = x[c];
                it.se2 = il; it.bl = x[q - 1];
            }
        }
        if (it.w_status == 1 && gm.trait == 1) {
            const double zst = it.beta[q - 1] / sqrt(it.se2);   // summary.glm: covariance of the last WLS
            it.pw = 2 * d_pnorm(-fabs(zst), true);
        }
        it.w_it = sweep + 1;
    }
    __syncwarp();
}

// one Newton step of fastgetroot_H1_new (ref :1795-1850) or the Barndorff-Nielsen tail (ref :1863-1875)
__device__ void update_tail(FItem &it, int h, const FGeom &gm, const double *part, int lane) {
    double tot[2];
    reduce_rows<2>(part, gm.rows, 2, lane, tot);
    const double v0 = tot[0], v1 = tot[1];
    if (lane == 0) {
        const double tol = 0x1p-13;
        const double s = it.tail_s[h];
        if (it.tail_phase[h] == 0) {
            const int iter = it.tail_iter[h] + 1;
            const double old_t = it.tail_t[h], old_diff = it.tail_diff[h], old_H1 = it.tail_H1[h];
            const double H1 = v0 - s, H2e = v1;
            double diff = -1 * H1 / H2e;
            double t = old_t;
            bool stop = false, finite_root = true;
            if (isnan(diff)) diff = 5;
            if (isnan(H1) || isinf(H1)) { t = d_sign(s) * CUDART_INF; stop = true; finite_root = false; }
            else {
                if (d_sign(H1) != d_sign(old_H1)) {
                    int guard = 0;
                    while (fabs(diff) > fabs(old_diff) - tol && guard++ < 1200) diff = diff / 2;
                }
                if (isnan(diff)) diff = 5;
                if (fabs(diff) < tol) stop = true;
                else t = old_t + diff;
            }
            it.tail_iter[h] = iter;
            it.tail_t[h] = t; it.tail_H1[h] = H1; it.tail_diff[h] = diff;
            if (!stop && iter == 100) { stop = true; }                     // loop exhausted
            if (stop) {
                if (iter == 100) finite_root = false;                        // `if(iter == maxiter) t = NA` semantics of the loop
                if (!finite_root || !isfinite(t)) { it.tail_p[h] = 0.0; it.tail_phase[h] = 2; }   // NA -> 0 (ref :1873)
                else it.tail_phase[h] = 1;
            }
        } else if (it.tail_phase[h] == 1) {
            const double t = it.tail_t[h];
            const double temp1 = t * s - v0;
            const double w = d_sign(t) * sqrt(2 * temp1);
            const double vv = t * sqrt(v1);
            double pval = d_pnorm(w + 1 / w * log(vv / w), h == 1);   // tail 0: upper (lower.tail = F), tail 1: lower
            if (isnan(pval)) pval = 0;
            it.tail_p[h] = pval; it.tail_phase[h] = 2;
        }
    }
    __syncwarp();
}

template <class CT>
__device__ __forceinline__ void run_wald_unit(const CT &cols, const FItem &it, const FGeom &gm, int mode, int64_t g0, int a, int b,
                                              double *pr, double *tile, int lane) {
    if (gm.q <= kFRegQ) {
        switch (gm.q * 4 + mode) {
            case 16: sweep_wald_reg<4, 0>(cols, it, g0, a, b, pr, lane); break;
            case 17: sweep_wald_reg<4, 1>(cols, it, g0, a, b, pr, lane); break;
            case 18: sweep_wald_reg<4, 2>(cols, it, g0, a, b, pr, lane); break;
            case 19: sweep_wald_reg<4, 3>(cols, it, g0, a, b, pr, lane); break;
            case 20: sweep_wald_reg<5, 0>(cols, it, g0, a, b, pr, lane); break;
            case 21: sweep_wald_reg<5, 1>(cols, it, g0, a, b, pr, lane); break;
            case 22: sweep_wald_reg<5, 2>(cols, it, g0, a, b, pr, lane); break;
            case 23: sweep_wald_reg<5, 3>(cols, it, g0, a, b, pr, lane); break;
            case 24: sweep_wald_reg<6, 0>(cols, it, g0, a, b, pr, lane); break;
            case 25: sweep_wald_reg<6, 1>(cols, it, g0, a, b, pr, lane); break;
            case 26: sweep_wald_reg<6, 2>(cols, it, g0, a, b, pr, lane); break;
            default: sweep_wald_reg<6, 3>(cols, it, g0, a, b, pr, lane); break;
        }
    } else {
        switch (mode) {
            case 0: sweep_wald_tile<0>(cols, it, gm, g0, a, b, pr, tile, lane); break;
            case 1: sweep_wald_tile<1>(cols, it, gm, g0, a, b, pr, tile, lane); break;
            case 2: sweep_wald_tile<2>(cols, it, gm, g0, a, b, pr, tile, lane); break;
            default: sweep_wald_tile<3>(cols, it, gm, g0, a, b, pr, tile, lane); break;
        }
    }
}

struct FarmWorkCounters { unsigned long long irls, tail, prep, iters, items; };

// ---- the kernel -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFThreads, 1)
k_bb_farm(BbFarmArgs A, FGeom gm, double *__restrict__ part /*2 x batch x rows x kw*/,
          double *__restrict__ tpart /*2 x batch x 2 x rows x 2*/, double *__restrict__ qpart /*2 x batch x rows x 4*/,
          FItem *__restrict__ gitems /*batch*/, unsigned *barrier,
          FarmWorkCounters *work) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int count = *A.count;
    if (count <= 0) return;          // uniform: every CTA leaves before the first barrier
    // shared-memory carve-up
    double *s_cols = reinterpret_cast<double *>(smem_raw);
    const size_t cols_doubles = (size_t)gm.n_staged * gm.len;
    FItem *s_items = reinterpret_cast<FItem *>(s_cols + cols_doubles);
    int *s_act = reinterpret_cast<int *>(s_items + kFBatch);                   // per item: who takes part in this iteration
    int *s_units = s_act + kFBatch;                                             // unit list of the iteration (kFBatch x 4 kinds x kFJ)
    int *s_ctl = s_units + kFBatch * 4 * kFJ;                                   // [0] units, [1] next unit, [2] active items (+ pad)
    double *s_red = reinterpret_cast<double *>(s_ctl + 4);                      // 16 x 32 group sums of the cooperative reduction
    double *s_tot = s_red + kFWarps * 32;                                       // 160 totals of the solver being updated,
                                                                                // then kMaxQ x kMaxQ for the wide-design Cholesky
    double *s_tiles = s_tot + 160 + kMaxQ * kMaxQ;                              // per warp: 32 x tile_stride (wide designs)
    double *tile = s_tiles + (size_t)warp * 32 * gm.tile_stride;

    const int64_t g0 = (int64_t)blockIdx.x * gm.len;                            // first sample of this CTA's slice
    const int mylen = (int)max((int64_t)0, min(gm.n, g0 + gm.len) - g0);
    const VecConst vc = *A.vc;
    // stage the SNP-invariant columns of the slice once
    ColsT<true> colsA;
    ColsT<false> colsB;
    for (int c = 0; c < gm.n_staged; c++) {
        const double *src = (c == 0) ? A.R : (c == 1) ? A.E : (c == 2) ? A.Y : A.cova + (int64_t)(c - 3) * gm.n;
        for (int i = threadIdx.x; i < mylen; i += blockDim.x) s_cols[(size_t)c * gm.len + i] = src[g0 + i];
    }
    colsB.s = s_cols; colsB.stride = (int)gm.len; colsB.n_staged = gm.n_staged;
    colsB.R = A.R; colsB.E = A.E; colsB.Y = A.Y; colsB.cova = A.cova; colsB.n = gm.n; colsB.g0 = g0;
    colsA.s = colsB.s; colsA.stride = colsB.stride; colsA.n_staged = colsB.n_staged;
    colsA.R = A.R; colsA.E = A.E; colsA.Y = A.Y; colsA.cova = A.cova; colsA.n = gm.n; colsA.g0 = g0;
    const bool all_staged = (gm.n_staged == gm.ncols);
    __syncthreads();
    unsigned bphase = 0;
    int pbuf = 0;
    const int half = ((mylen + kFJ - 1) / kFJ + 3) / 4 * 4;                     // unit boundaries stay byte aligned
    auto unit_range = [&](int j, int &a, int &b) { a = min(mylen, j * half); b = min(mylen, (j + 1) * half); };
    auto prow_of = [&](int buf, int item, int j) {
        return part + (((size_t)buf * kFBatch + item) * gm.rows + (size_t)blockIdx.x * kFJ + j) * gm.kw;
    };
    auto qrow_of = [&](int buf, int item, int j) {
        return qpart + (((size_t)buf * kFBatch + item) * gm.rows + (size_t)blockIdx.x * kFJ + j) * 4;
    };
    auto trow_of = [&](int buf, int item, int h, int j) {
        return tpart + ((((size_t)buf * kFBatch + item) * 2 + h) * gm.rows + (size_t)blockIdx.x * kFJ + j) * 2;
    };
    unsigned long long w_irls = 0, w_tail = 0, w_prep = 0, w_iters = 0;

    for (int b0 = 0; b0 < count; b0 += kFBatch) {
        const int nb = min(kFBatch, count - b0);
        __syncthreads();
        // ---- item setup (ref :557-570 prologue from the integer counts) ----
        if (threadIdx.x < nb) {
            FItem &it = s_items[threadIdx.x];
            const BItem bi = A.items[b0 + threadIdx.x];
            const Prologue pr = make_prologue(bi.asum, bi.nmiss, gm.n);
            it.row = bi.row; it.snp = bi.snp; it.asum = bi.asum; it.nmiss = bi.nmiss; it.D0 = bi.D0; it.T0 = bi.T0;
            const GClass gc = make_gclass(pr, A.g_model);
            for (int c = 0; c < 4; c++) it.gval[c] = gc.val[c];
            it.H0 = bi.H0;
            for (int c = 0; c < 4; c++) it.fit[c] = 0;
            it.flip = pr.flip; it.sum_g = pr.sum_g; it.u = pr.u;
            for (int c = 0; c < kMaxQ; c++) it.beta[c] = 0;
            it.devold = 0; it.se2 = 0; it.bl = 0; it.pw = d_na_real();
            it.w_it = 0; it.w_status = (gm.trait >= 3) ? 1 : 0; it.spa_kind = 0; it.prep = 0; it.pad_ = 0;   // survival / categorical: the Wald test is cox_wald.cu's / clm_wald.cu's
            it.mafi = 0; it.S = 0; it.Smean = 0; it.z = 0;
            it.pspa = d_na_real(); it.pnorm_ = d_na_real();
            for (int h = 0; h < 2; h++) { it.tail_phase[h] = 2; it.tail_iter[h] = 0; it.tail_p[h] = 0; it.tail_t[h] = 0; it.tail_H1[h] = 0; it.tail_diff[h] = CUDART_INF; it.tail_s[h] = 0; }
        }
        // ---- the iteration loop: every active solver of every item advances one step per pair of grid barriers ----
        // solvers of an item: preparation (class counts -> fitted values -> adjusted-residual sums -> SPA set-up), the Wald
        // fit (runs from the first iteration on: it only needs the genotypes) and, once the preparation is done, the tails
        for (;;) {
            __syncthreads();
            // snapshot of who takes part in this iteration: unit numbering and ownership must not move while item
            // states are being updated
            if (threadIdx.x < nb) {
                const FItem &it = s_items[threadIdx.x];
                s_act[threadIdx.x] = ((it.w_status == 0 && it.spa_kind != 3) ? 1 : 0) | (it.tail_phase[0] < 2 ? 2 : 0) |
                                     (it.tail_phase[1] < 2 ? 4 : 0) | (it.prep == 0 ? 8 : 0) | (it.prep == 1 ? 16 : 0);
            }
            __syncthreads();
            // unit list of this iteration, most expensive kinds first (Wald, tails, preparation); one warp builds it,
            // then the warps of the CTA take units from it dynamically (the partial row of a unit does not depend on
            // which warp computes it, so the results stay reproducible)
            if (warp == 0) {
                int n_units = 0, n_act = 0;
                for (int pass = 0; pass < 3; pass++) {
                    for (int i0 = 0; i0 < nb; i0 += 32) {
                        const int item = i0 + lane;
                        const int act = (item < nb) ? s_act[item] : 0;
                        if (pass == 0) n_act += __popc(__ballot_sync(0xffffffffu, act != 0));
                        const int kinds = (pass == 0) ? (act & 1) : (pass == 1) ? ((act >> 1) & 3) : ((act >> 3) & 3);
                        // per lane: number of solver kinds of this pass that are active (0..2), kFJ units each
                        const int mine = __popc(kinds) * kFJ;
                        int pre = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += t; }
                        int pos = n_units + pre - mine;
                        for (int bit = 0; bit < 2; bit++) {
                            if (!(kinds & (1 << bit))) continue;
                            const int kind = (pass == 0) ? 0 : (pass == 1) ? 1 + bit : 3 + bit;   // 0 Wald, 1/2 tails, 3 counts, 4 sums
                            for (int j = 0; j < kFJ; j++) s_units[pos++] = (item << 8) | (kind << 4) | j;
                        }
                        n_units += __shfl_sync(0xffffffffu, pre, 31);
                    }
                }
                if (lane == 0) { s_ctl[0] = n_units; s_ctl[1] = 0; s_ctl[2] = n_act; }
            }
            __syncthreads();
            const int n_units = s_ctl[0], n_active = s_ctl[2];
            for (;;) {
                int k = 0;
                if (lane == 0) k = atomicAdd(&s_ctl[1], 1);
                k = __shfl_sync(0xffffffffu, k, 0);
                if (k >= n_units) break;
                const int desc = s_units[k];
                const int item = desc >> 8, kind = (desc >> 4) & 15, j = desc & 15;
                const FItem &it = s_items[item];
                int a, b; unit_range(j, a, b);
                if (kind == 0) {
                    double *pr = prow_of(pbuf, item, j);
                    const int mode = (gm.trait == 1) ? (it.w_it == 0 ? 0 : 1) : (it.w_it == 0 ? 2 : 3);
                    if (all_staged) run_wald_unit(colsA, it, gm, mode, g0, a, b, pr, tile, lane);
                    else run_wald_unit(colsB, it, gm, mode, g0, a, b, pr, tile, lane);
                    w_irls += (unsigned long long)(b - a);
                } else if (kind <= 2) {
                    const int h = kind - 1;
                    double *tr = trow_of(pbuf, item, h, j);
                    if (all_staged) {
                        if (it.tail_phase[h] == 0) sweep_tail<false>(colsA, it, h, g0, a, b, tr, lane);
                        else sweep_tail<true>(colsA, it, h, g0, a, b, tr, lane);
                    } else {
                        if (it.tail_phase[h] == 0) sweep_tail<false>(colsB, it, h, g0, a, b, tr, lane);
                        else sweep_tail<true>(colsB, it, h, g0, a, b, tr, lane);
                    }
                    w_tail += (unsigned long long)(b - a);
                } else {
                    double *qr = qrow_of(pbuf, item, j);
                    if (kind == 3) sweep_counts(it, g0, a, b, qr, lane);
                    else sweep_adj(colsB, it, g0, a, b, qr, lane);
                    w_prep += (unsigned long long)(b - a);
                }
            }
            if (n_active == 0) break;                     // identical item tables in every CTA: a uniform decision
            w_iters++;
            grid_barrier(barrier, bphase);
            // ---- scalar updates: item k of the batch belongs to CTA k mod grid.  The owner adds the partial rows once
            // (fixed order), advances the item's solvers and publishes the new state; after the second barrier every
            // CTA copies the states of the items that took part, so all CTAs keep identical item tables.
            for (int item = blockIdx.x; item < nb; item += gridDim.x) {   // the items this CTA owns
                const int act = s_act[item];
                if (!act) continue;                                           // uniform in the CTA
                FItem &it = s_items[item];
                if (act & 1)
                    cta_reduce_rows(part + ((size_t)pbuf * kFBatch + item) * gm.rows * gm.kw, gm.rows, gm.kw,
                                    gm.q * (gm.q + 3) / 2 + 2, s_red, s_tot);
                if (warp == 0) {
                    if (act & 1) update_wald(it, gm, s_tot, lane);
                } else if (warp == 1 || warp == 2) {
                    const int h = warp - 1;
                    if (act & (2 << h)) update_tail(it, h, gm, tpart + (((size_t)pbuf * kFBatch + item) * 2 + h) * gm.rows * 2, lane);
                } else if (warp == 3 && (act & 24)) {
                    double tot[4];
                    reduce_rows<4>(qpart + ((size_t)pbuf * kFBatch + item) * gm.rows * 4, gm.rows, 4, lane, tot);
                    if (lane == 0) {
                        if (act & 8) {
                            // W'W, W'R -> fitted value per genotype class (ref :606-611); sum(g^2) from the exact class counts
                            const double n00 = tot[0], n10 = tot[1];
                            const double n01 = (double)it.nmiss, n11 = (double)gm.n - n00 - n10 - n01;
                            const double sg2 = ((n00 * (it.gval[0] * it.gval[0]) + n01 * (it.gval[1] * it.gval[1])) +
                                                n10 * (it.gval[2] * it.gval[2])) + n11 * (it.gval[3] * it.gval[3]);
                            double S1 = it.D0 + ((it.nmiss > 0) ? it.u * it.T0 : 0.0);   // sum(g R) from the score pass (ref :582)
                            if (it.flip) S1 = 2 * vc.sumR - S1;
                            if (A.g_model != 0) {   // recoded genotypes: class values times class sums / counts
                                GClass gc; for (int c = 0; c < 4; c++) gc.val[c] = it.gval[c];
                                S1 = class_dot(gc, it.D0, it.T0, it.H0, vc.sumR);
                                it.sum_g = ((n00 * it.gval[0] + n01 * it.gval[1]) + n10 * it.gval[2]) + n11 * it.gval[3];
                            }
                            const double a_ = (double)gm.n, b_ = it.sum_g, d_ = sg2;
                            const double det = a_ * d_ - b_ * b_;
                            if (det == 0 || !isfinite(det)) { it.spa_kind = 3; it.prep = 2; }   // solve() would stop (ref :609)
                            else {
                                const double i00 = d_ / det, i01 = -b_ / det, i11 = a_ / det;
                                for (int c = 0; c < 4; c++) {
                                    const double wi0 = i00 + it.gval[c] * i01, wi1 = i01 + it.gval[c] * i11;
                                    it.fit[c] = wi0 * vc.sumR + wi1 * S1;
                                }
                                it.prep = 1;
                            }
                        } else {
                            double mafi, S, Smean = 0, z = 0;
                            // branch B forwards min.maf to the inner call (ref :616); Cutoff stays 2
                            const int r = homo_scalar(it.sum_g, gm.n, tot[0], tot[1], tot[2], A.min_maf, 2.0, &mafi, &S, &Smean, &z);
                            it.spa_kind = r; it.mafi = mafi; it.S = S; it.Smean = Smean; it.z = z;
                            if (r == 2) { it.pspa = d_na_real(); it.pnorm_ = d_na_real(); }
                            else if (r == 0) { it.pspa = d_pnorm(fabs(z), false) * 2; it.pnorm_ = it.pspa; }
                            else {
                                it.pnorm_ = d_pnorm(fabs(z), false) + d_pnorm(-fabs(z), true);
                                it.tail_s[0] = fmax(S, 2 * Smean - S); it.tail_s[1] = fmin(S, 2 * Smean - S);
                                it.tail_phase[0] = 0; it.tail_phase[1] = 0;
                            }
                            it.prep = 2;
                        }
                    }
                }
                __syncthreads();
                // publish (one warp copies the whole state)
                if (warp == 4) {
                    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&s_items[item]);
                    unsigned long long *dst = reinterpret_cast<unsigned long long *>(gitems + item);
                    for (int k = lane; k < (int)(sizeof(FItem) / 8); k += 32) dst[k] = src[k];
                }
            }
            grid_barrier(barrier, bphase);
            for (int item = warp; item < nb; item += kFWarps) {
                if (!s_act[item] || (item % (int)gridDim.x) == (int)blockIdx.x) continue;
                const unsigned long long *src = reinterpret_cast<const unsigned long long *>(gitems + item);
                unsigned long long *dst = reinterpret_cast<unsigned long long *>(&s_items[item]);
                for (int k = lane; k < (int)(sizeof(FItem) / 8); k += 32) dst[k] = __ldcg(src + k);
            }
            pbuf ^= 1;
        }
        // ---- results (CTA 0 writes; every CTA holds the same values) ----
        if (blockIdx.x == 0 && threadIdx.x < nb) {
            FItem &it = s_items[threadIdx.x];
            const double NA = d_na_real();
            double *o = A.out + it.snp;
            const int64_t M = A.m_total;
            if (it.spa_kind == 3) {
                o[2 * M] = NA; o[3 * M] = NA; o[4 * M] = NA; o[5 * M] = NA;
                atomicAdd(&A.ctr->n_singular, 1ULL);
            } else {
                double pspa = it.pspa;
                if (it.spa_kind == 1) {
                    double p1 = it.tail_p[0], p2 = it.tail_p[1];
                    if (isnan(p2)) p2 = p1;
                    pspa = p1 + p2;
                    atomicAdd(&A.ctr->n_spa, 1ULL);
                    atomicAdd(&A.ctr->newton_iters, (unsigned long long)(it.tail_iter[0] + it.tail_iter[1]));
                }
                if (gm.trait >= 3) {   // survival / categorical: k_cox_wald / k_clm_wald fill the Wald and CCT columns from p.spa
                    o[2 * M] = pspa; o[3 * M] = NA; o[4 * M] = NA; o[5 * M] = it.pnorm_;
                } else {
                double pw = it.pw;
                if (it.w_status >= 2) { pw = NA; atomicAdd(&A.ctr->n_wald_fail, 1ULL); }
                double pc = NA;
                const int crc = d_cct2(pspa, pw, &pc);
                if (crc) atomicAdd(&A.ctr->n_cct_stop, 1ULL);
                o[2 * M] = pspa; o[3 * M] = pw; o[4 * M] = pc; o[5 * M] = it.pnorm_;
                }
            }
        }
    }
    // work accounting for the fp64 roofline (sample-passes actually swept)
    if (lane == 0) {   // per-warp values are lane-uniform
        atomicAdd(&work->irls, w_irls); atomicAdd(&work->tail, w_tail); atomicAdd(&work->prep, w_prep);
        if (blockIdx.x == 0 && warp == 0) { atomicAdd(&work->iters, w_iters); atomicAdd(&work->items, (unsigned long long)count); }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------------
struct BbFarm {
    int device = 0, num_sms = 0, grid = 0;
    size_t smem_max = 0;
    double *d_part = nullptr; size_t cap_part = 0;
    double *d_tpart = nullptr; size_t cap_tpart = 0;   // tails (2 doubles per row) followed by the preparation rows (4)
    unsigned *d_barrier = nullptr;
    FarmWorkCounters *d_work = nullptr;
    FItem *d_items = nullptr;
};

cudaError_t bb_farm_create(BbFarm **out, int device, int num_sms) {
    BbFarm *f = new BbFarm();
    f->device = device; f->num_sms = num_sms;
    cudaError_t e;
    int smem_optin = 0;
    if ((e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)) != cudaSuccess) { delete f; return e; }
    f->smem_max = (size_t)smem_optin;
    if ((e = cudaFuncSetAttribute(k_bb_farm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin)) != cudaSuccess) { delete f; return e; }
    int per_sm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bb_farm, kFThreads, (size_t)smem_optin)) != cudaSuccess) { delete f; return e; }
    if (per_sm < 1) { delete f; return cudaErrorLaunchOutOfResources; }
    f->grid = num_sms;   // one CTA per SM: the slices are sized for exactly this
    if ((e = cudaMalloc((void **)&f->d_barrier, sizeof(unsigned))) != cudaSuccess) { delete f; return e; }
    if ((e = cudaMalloc((void **)&f->d_work, sizeof(FarmWorkCounters))) != cudaSuccess) { cudaFree(f->d_barrier); delete f; return e; }
    cudaMemset(f->d_work, 0, sizeof(FarmWorkCounters));
    if ((e = cudaMalloc((void **)&f->d_items, sizeof(FItem) * kFBatch)) != cudaSuccess) { cudaFree(f->d_barrier); cudaFree(f->d_work); delete f; return e; }
    *out = f;
    return cudaSuccess;
}

void bb_farm_destroy(BbFarm *f) {
    if (!f) return;
    cudaFree(f->d_part); cudaFree(f->d_tpart); cudaFree(f->d_barrier); cudaFree(f->d_work); cudaFree(f->d_items);
    delete f;
}

cudaError_t bb_farm_launch(BbFarm *f, cudaStream_t s, const BbFarmArgs &args, int64_t *launches, const char **what) {
    cudaError_t e;
    FGeom gm;
    gm.n = args.n; gm.p = args.p; gm.q = args.p + 4; gm.trait = args.trait;
    gm.ncols = 3 + args.p;
    // fewer CTAs than SMs when N is small: at least 2,048 samples per slice
    int grid = (int)std::min<int64_t>(f->grid, std::max<int64_t>(1, (args.n + 2047) / 2048));
    gm.len = ((args.n + grid - 1) / grid + 3) / 4 * 4;
    grid = (int)((args.n + gm.len - 1) / gm.len);
    gm.rows = grid * kFJ;
    const int nent = gm.q * (gm.q + 3) / 2;
    gm.kw = (nent + 2 + 31) / 32 * 32;
    gm.tile_stride = (gm.q <= kFRegQ) ? 0 : ((2 * gm.q + 1) | 1);
    gm.ld = (gm.q <= 8) ? 8 : kMaxQ;
    const size_t fixed = sizeof(FItem) * kFBatch + sizeof(int) * (kFBatch + kFBatch * 4 * kFJ + 4) +
                         sizeof(double) * (kFWarps * 32 + 160 + kMaxQ * kMaxQ) +
                         sizeof(double) * kFWarps * 32 * (size_t)gm.tile_stride + 64;
    if (fixed > f->smem_max) { *what = "bb_farm: shared-memory layout"; return cudaErrorInvalidConfiguration; }
    // as many SNP-invariant columns of the slice as fit (R, E, Y first); the rest is read through L1/L2
    gm.n_staged = (int)std::min<size_t>((size_t)gm.ncols, (f->smem_max - fixed) / (sizeof(double) * (size_t)gm.len));
    const size_t smem = fixed + sizeof(double) * (size_t)gm.n_staged * gm.len;
    const size_t need_part = (size_t)2 * kFBatch * gm.rows * gm.kw;
    if (need_part > f->cap_part) {
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) { *what = "sync"; return e; }
        cudaFree(f->d_part); f->d_part = nullptr; f->cap_part = 0;
        if ((e = cudaMalloc((void **)&f->d_part, need_part * sizeof(double))) != cudaSuccess) { *what = "cudaMalloc(part)"; return e; }
        f->cap_part = need_part;
    }
    const size_t need_t = (size_t)2 * kFBatch * 2 * gm.rows * 2 + (size_t)2 * kFBatch * gm.rows * 4;
    if (need_t > f->cap_tpart) {
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) { *what = "sync"; return e; }
        cudaFree(f->d_tpart); f->d_tpart = nullptr; f->cap_tpart = 0;
        if ((e = cudaMalloc((void **)&f->d_tpart, need_t * sizeof(double))) != cudaSuccess) { *what = "cudaMalloc(tpart)"; return e; }
        f->cap_tpart = need_t;
    }
    if ((e = cudaMemsetAsync(f->d_barrier, 0, sizeof(unsigned), s)) != cudaSuccess) { *what = "memset(barrier)"; return e; }
    BbFarmArgs a = args;
    double *part = f->d_part, *tpart = f->d_tpart;
    unsigned *bar = f->d_barrier;
    FarmWorkCounters *work = f->d_work;
    FItem *gitems = f->d_items;
    double *qpart = f->d_tpart + (size_t)2 * kFBatch * 2 * gm.rows * 2;
    void *kargs[] = {&a, &gm, &part, &tpart, &qpart, &gitems, &bar, &work};
    e = cudaLaunchCooperativeKernel((const void *)k_bb_farm, dim3((unsigned)grid), dim3(kFThreads), kargs, smem, s);
    if (e != cudaSuccess) { *what = "cudaLaunchCooperativeKernel(k_bb_farm)"; return e; }
    (*launches)++;
    return cudaSuccess;
}

cudaError_t bb_farm_read_work(BbFarm *f, cudaStream_t s, BbFarmWork *w, bool reset) {
    FarmWorkCounters h;
    cudaError_t e = cudaMemcpyAsync(&h, f->d_work, sizeof h, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return e;
    w->irls_sample_passes = h.irls; w->tail_sample_passes = h.tail; w->prep_sample_passes = h.prep;
    w->iterations = h.iters; w->items = h.items;
    if (reset) e = cudaMemsetAsync(f->d_work, 0, sizeof h, s);
    return e;
}

}  // namespace spagxe
