// dmath.cuh -- fp64 device math shared by the Step-2 kernels (sm_100a).
//
// Restates base-R numerics the reference calls on its hot path:
//   stats::pnorm   (R/SPAGxECCT.R:585, :1871, :2043, :2056)  -> d_pnorm   (Cody 1969)
//   stats::pcauchy (R/SPAGxECCT.R:2120)                      -> d_pcauchy
//   CCT()          (R/SPAGxECCT.R:2070-2123)                 -> d_cct2
//   stats::pt via summary.lm (R/SPAGxECCT.R:654, :1069)      -> d_pt2 (incomplete beta CF)
#pragma once
#include <cstdint>
#include <cfloat>
#include <math_constants.h>

namespace spagxe {

__device__ __forceinline__ double d_na_real() { return __longlong_as_double(0x7FF00000000007A2LL); }
__device__ __forceinline__ double d_sign(double x) { return (x > 0.0) - (x < 0.0); }  // NaN handled by callers

// ---- branch-free exp / reciprocal for the per-sample hot loops ---------------------------------------------------
// exp(x) for |x| < 700 (result normal): k = round(x / ln 2), r = x - k ln 2 in two parts, e^r by a degree-11 polynomial
// (Chebyshev interpolant on |r| <= 0.3466, relative error 1.7e-17 before rounding), 2^k through the exponent field.
// The CUDA library exp() computes the same thing but wraps it in range branches (BSSY/BSYNC pairs) that keep ptxas
// from interleaving the chains of neighbouring samples; the callers guard the range instead.
__device__ __forceinline__ double d_exp_fast(double x) {
    const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);   // 1.5 * 2^52: the low word is round(x / ln 2)
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 2.5110049204818658e-08;
    p = fma(p, r, 2.763265472252779e-07);
    p = fma(p, r, 2.755724088722987e-06);
    p = fma(p, r, 2.4801485441561313e-05);
    p = fma(p, r, 0.00019841269890076403);
    p = fma(p, r, 0.0013888888952352863);
    p = fma(p, r, 0.008333333333319589);
    p = fma(p, r, 0.04166666666648795);
    p = fma(p, r, 0.1666666666666668);
    p = fma(p, r, 0.5000000000000019);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
// 1 / x for normal x with a normal reciprocal: hardware seed (2^-23) + two Newton steps, no special-case branch
__device__ __forceinline__ double d_rcp_fast(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// ---- pnorm (both tails, full relative accuracy down to ~1e-308) -------------
static __device__ __noinline__ double d_pnorm(double x, bool lower_tail) {
    const double a[5] = {2.2352520354606839287, 161.02823106855587881, 1067.6894854603709582,
                         18154.981253343561249, 0.065682337918207449113};
    const double b[4] = {47.20258190468824187, 976.09855173777669322, 10260.932208618978205,
                         45507.789335026729956};
    const double c[9] = {0.39894151208813466764, 8.8831497943883759412, 93.506656132177855979,
                         597.27027639480026226, 2494.5375852903726711, 6848.1904505362823326,
                         11602.651437647350124, 9842.7148383839780218, 1.0765576773720192317e-8};
    const double d[8] = {22.266688044328115691, 235.38790178262499861, 1519.377599407554805,
                         6485.558298266760755, 18615.571640885098091, 34900.952721145977266,
                         38912.003286093271411, 19685.429676859990727};
    const double p[6] = {0.21589853405795699, 0.1274011611602473639, 0.022235277870649807,
                         0.001421619193227893466, 2.9112874951168792e-5, 0.02307344176494017303};
    const double q[5] = {1.28426009614491121, 0.468238212480865118, 0.0659881378689285515,
                         0.00378239633202758244, 7.29751555083966205e-5};
    if (isnan(x)) return x;
    if (isinf(x)) return (x < 0) ? (lower_tail ? 0. : 1.) : (lower_tail ? 1. : 0.);
    double cum, ccum, xnum, xden, temp, xsq, del;
    const double y = fabs(x);
    if (y <= 0.67448975) {
        if (y > DBL_EPSILON * 0.5) {
            xsq = x * x;
            xnum = a[4] * xsq; xden = xsq;
#pragma unroll
            for (int i = 0; i < 3; ++i) { xnum = (xnum + a[i]) * xsq; xden = (xden + b[i]) * xsq; }
        } else { xnum = xden = 0.0; }
        temp = x * (xnum + a[3]) / (xden + b[3]);
        cum = 0.5 + temp; ccum = 0.5 - temp;
    } else if (y <= 5.656854249492380195206754896838) {
        xnum = c[8] * y; xden = y;
#pragma unroll
        for (int i = 0; i < 7; ++i) { xnum = (xnum + c[i]) * y; xden = (xden + d[i]) * y; }
        temp = (xnum + c[7]) / (xden + d[7]);
        xsq = trunc(y * 16) / 16;
        del = (y - xsq) * (y + xsq);
        cum = exp(-xsq * xsq * 0.5) * exp(-del * 0.5) * temp;
        ccum = 1.0 - cum;
        if (x > 0.) { temp = cum; cum = ccum; ccum = temp; }
    } else if ((lower_tail && -37.5193 < x && x < 8.2924) || (!lower_tail && -8.2924 < x && x < 37.5193)) {
        xsq = 1.0 / (x * x);
        xnum = p[5] * xsq; xden = xsq;
#pragma unroll
        for (int i = 0; i < 4; ++i) { xnum = (xnum + p[i]) * xsq; xden = (xden + q[i]) * xsq; }
        temp = xsq * (xnum + p[4]) / (xden + q[4]);
        temp = (0.398942280401432677939946059934 - temp) / y;
        xsq = trunc(x * 16) / 16;
        del = (x - xsq) * (x + xsq);
        cum = exp(-xsq * xsq * 0.5) * exp(-del * 0.5) * temp;
        ccum = 1.0 - cum;
        if (x > 0.) { temp = cum; cum = ccum; ccum = temp; }
    } else {
        if (x > 0) { cum = 1.; ccum = 0.; } else { cum = 0.; ccum = 1.; }
    }
    return lower_tail ? cum : ccum;
}

__device__ __forceinline__ double d_pcauchy(double x) {
    if (isnan(x)) return x;
    if (isinf(x)) return (x < 0) ? 0. : 1.;
    if (fabs(x) > 1) {
        double y = atan(1. / x) / CUDART_PI;
        return (x > 0) ? (0.5 - y + 0.5) : (-y);
    }
    return 0.5 + atan(x) / CUDART_PI;
}

// CCT(c(p1, p2)) with equal weights; returns 0 or the stop() code (1 NA, 2 range, 3 both 0 and 1).
static __device__ __noinline__ int d_cct2(double p1, double p2, double *out) {
    *out = d_na_real();
    if (isnan(p1) || isnan(p2)) return 1;
    if ((p1 < 0) || (p2 < 0) || (p1 > 1) || (p2 > 1)) return 2;
    const bool is_zero = (p1 == 0) || (p2 == 0);
    const bool is_one = (p1 == 1) || (p2 == 1);
    if (is_zero && is_one) return 3;
    if (is_zero) { *out = 0; return 0; }
    if (p1 == 1) p1 = 0.999;
    if (p2 == 1) p2 = 0.999;
    const double w = 0.5;
    const bool s1 = p1 < 1e-16, s2 = p2 < 1e-16;
    double stat;
    if (!s1 && !s2) {
        stat = w * tan((0.5 - p1) * CUDART_PI) + w * tan((0.5 - p2) * CUDART_PI);
    } else {
        double a = 0, b = 0;
        if (s1) a += (w / p1) / CUDART_PI; else b += w * tan((0.5 - p1) * CUDART_PI);
        if (s2) a += (w / p2) / CUDART_PI; else b += w * tan((0.5 - p2) * CUDART_PI);
        stat = a + b;
    }
    if (stat > 1e15) *out = (1 / stat) / CUDART_PI;
    else *out = 1 - d_pcauchy(stat);
    return 0;
}

// ---- regularized incomplete beta by modified Lentz; 2*pt(|t|, df, lower=FALSE) ----
static __device__ __noinline__ double d_betacf(double a, double b, double x) {
    const double FPMIN = 1e-300, EPS = 1e-16;
    double qab = a + b, qap = a + 1, qam = a - 1, c = 1, d = 1 - qab * x / qap;
    if (fabs(d) < FPMIN) d = FPMIN;
    d = 1 / d;
    double h = d;
    for (int m = 1; m <= 200000; m++) {
        double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
        c = 1 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
        d = 1 / d; h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
        c = 1 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
        d = 1 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1) < EPS) break;
    }
    return h;
}
// log B(a, 1/2) without the lgamma cancellation for large a
__device__ __forceinline__ double d_lbeta_half(double a) {
    if (a < 50.0) return lgamma(a) + 0.57236494292470008707 /*lgamma(.5)*/ - lgamma(a + 0.5);
    // lgamma(a) - lgamma(a+1/2) = -0.5*log(a) + 1/(8a) - 1/(192 a^3) + 1/(640 a^5) - ...
    double ia = 1.0 / a, ia2 = ia * ia;
    double dlt = -0.5 * log(a) + ia * (0.125 + ia2 * (-1.0 / 192 + ia2 * (1.0 / 640 - ia2 * 17.0 / 14336)));
    return dlt + 0.57236494292470008707;
}
// I_x(a, 1/2) with y = 1 - x supplied separately
static __device__ __noinline__ double d_betainc_half(double a, double x, double y) {
    if (x <= 0) return 0;
    if (y <= 0) return 1;
    const double A = a, B = 0.5;
    double lbt = -d_lbeta_half(a) + A * ((y < 0.5) ? log1p(-y) : log(x)) + B * log(y);
    if (x < (A + 1) / (A + B + 2)) return exp(lbt) * d_betacf(A, B, x) / A;
    return 1 - exp(lbt) * d_betacf(B, A, y) / B;
}
// 2*pt(|t|, df, lower=FALSE) = I_{df/(df+t^2)}(df/2, 1/2)
static __device__ __noinline__ double d_pt2(double t, double df) {
    if (isnan(t) || isnan(df)) return t + df;
    const double x2 = t * t, den = df + x2;
    return d_betainc_half(df / 2, df / den, x2 / den);
}

// ---- block reductions (fixed order => deterministic) -----------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sum NV doubles over the block; result valid in ALL threads. scratch: >= NV*32 doubles of smem.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = warp_sum(v[k]);
    __syncthreads();  // scratch may still be read from a previous call
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) scratch[k * 32 + warp] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double x = (lane < nw) ? scratch[k * 32 + lane] : 0.0;
        v[k] = warp_sum(x);
    }
}

}  // namespace spagxe
