// stat_common.cuh -- per-sample / per-SNP device helpers shared by the translation units of the per-SNP tests
// (kernels.cuh: SPA, SPAGxEmix, SPAGxE_Plus; bb_farm.cu: the SPAGxE_CCT branch-B solver farm).
#pragma once
#include "dmath.cuh"
#include "types.cuh"

namespace spagxe {

// SPA_G_one_SNP_homo_new's scalar part (ref :2000-2045) given S = sum(g * Rv) and the vector's
// SNP-invariant sums.  Returns: 0 -> p-values final (normal), 1 -> SPA needed, 2 -> inner min.maf NA.
__device__ __forceinline__ int homo_scalar(double sum_g, int64_t n, double S_in, double sumRv, double sumRv2,
                                           double min_maf_inner, double cutoff, double *maf_i, double *S_out,
                                           double *Smean, double *z) {
    double mafi = (sum_g / (double)n) / 2;  // mean(g)/2 on the imputed, flipped vector (ref :2000)
    double S = S_in;
    if (mafi > 0.5) { mafi = 1 - mafi; S = 2 * sumRv - S; }  // ref :2010-2013 (ties only)
    *maf_i = mafi; *S_out = S;
    if (mafi < min_maf_inner) return 2;
    const double gvar = 2 * mafi * (1 - mafi);
    const double Svar = sumRv2 * gvar;   // sum(R^2 * g.var.est) (ref :2034)
    *Smean = 2 * (sumRv * mafi);         // 2 * sum(R * MAF.est)  (ref :2037)
    *z = (S - *Smean) / sqrt(Svar);
    return (fabs(*z) < cutoff) ? 0 : 1;
}

// ---------------------------------------------------------------------------
// K4: saddlepoint approximation.  CGF pieces evaluated exactly as the reference writes them
// (ref :1652-1681) so overflow/NaN appear in the same places.
__device__ __forceinline__ void cgf_k12(double tr, double maf, double &k1, double &k2) {
    const double e = (fabs(tr) < 700.0) ? d_exp_fast(tr) : exp(tr);   // library exp only where Inf / 0 / NaN can arise
    const double x = maf * e;
    const double u = (1 - maf) + x;
    const double m0 = u * u;
    const double m1 = (2 * x) * u;
    const double m2 = 2 * (x * x) + (2 * x) * u;
    const double num = m0 * m2 - m1 * m1;
    if (m0 < 1e150 && m0 > 1e-150) {   // common range: one reciprocal instead of two divisions (same values to rounding)
        const double r = d_rcp_fast(m0);
        k1 = m1 * r;
        k2 = (num * r) * r;
    } else {                     // overflow neighbourhood: keep the reference's exact Inf/NaN behaviour
        k1 = m1 / m0;
        k2 = num / (m0 * m0);
    }
}
__device__ __forceinline__ double cgf_k0(double tr, double maf) {
    const double u = (1 - maf) + maf * exp(tr);
    return log(u * u);
}

struct DevProd {   // prod (1+e^eta) = m * 2^k; log(prod) costs one log per thread instead of one per sample
    double m = 1.0; int k = 0; int cnt = 0;
    __device__ __forceinline__ void mul(double x) {
        m *= x;                                   // x in [1, 1+e^30]: 8 factors stay far below 2^1023
        if (++cnt == 8) {                         // m is a normal number >= 1: renormalise through the exponent field
            cnt = 0;
            const int hi = __double2hiint(m);
            const int e = ((hi >> 20) & 0x7ff) - 1023;
            k += e;
            m = __hiloint2double(hi - (e << 20), __double2loint(m));
        }
    }
    __device__ __forceinline__ double log_value() const { return log(m) + (double)k * 0.693147180559945309417232121458; }
};

// Per-sample IRLS working quantities (glm.fit binomial/logit restated).
// Returns the working weight w and wz = w * z (z = eta + (y - mu)/mu.eta), which is all X'WX and X'Wz need.
__device__ __forceinline__ void irls_sample(int trait, bool pass0, double y, double eta_in, double &w2, double &wz,
                                            double &dev, int &bad, DevProd &dp) {
    const double THRESH = 30., MTHRESH = -30., INVEPS = 1 / DBL_EPSILON;
    if (trait == 1) {
        double eta = eta_in;
        if (pass0) { double mus = (y + 0.5) / 2; eta = log(mus / (1 - mus)); }
        const bool clampd = (eta > THRESH || eta < MTHRESH);
        const double ee = exp(eta);
        if (!clampd && (y == 0.0 || y == 1.0)) {
            // mu = e/(1+e), mu.eta = e/(1+e)^2 = Var(mu) => w = mu.eta, w z = w eta + (y - mu);
            // dev_i = -2 log(mu^y (1-mu)^(1-y)) = 2 (log(1+e) - y eta): one reciprocal, one log, no division
            const double opexp = 1 + ee;
            const double r = 1.0 / opexp;
            const double mu = ee * r;
            const double mev = mu * r;
            dp.mul(opexp);                 // 2 log(1+e) is added once per thread from the running product
            dev -= 2 * (y * eta);
            w2 = mev;
            wz = fma(mev, eta, y - mu);
            return;
        }
        const double tmp = (eta < MTHRESH) ? DBL_EPSILON : ((eta > THRESH) ? INVEPS : ee);
        const double opexp = 1 + ee;
        const double mu = tmp / (1 + tmp);
        if (!(mu > 0 && mu < 1)) bad = 1;
        double d;
        if (y == 1.0) d = -log(mu);                // y*log(y/mu); y_log_y(0, .) = 0
        else if (y == 0.0) d = -log(1 - mu);
        else d = y * log(y / mu) + (1 - y) * log((1 - y) / (1 - mu));
        dev += 2 * d;
        const double varmu = mu * (1 - mu);
        const double mev = clampd ? DBL_EPSILON : ee / (opexp * opexp);
        w2 = (mev * mev) / varmu;
        wz = w2 * (eta + (y - mu) / mev);
    } else {
        w2 = 1.0; wz = y;
        if (!pass0) dev += (y - eta_in) * (y - eta_in);  // lm second pass: residual sum of squares
    }
}


}  // namespace spagxe
