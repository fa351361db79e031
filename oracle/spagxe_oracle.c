/*
 * spagxe_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the Step-2 per-SNP testing arithmetic of the
 * SPAGxECCT R package, written from the published method and the behaviour
 * of the reference's R functions.  Every function cites the reference
 * file:line it follows (paths relative to /root/reference).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the CUDA product path never
 * links or calls it.
 *
 * PARITY STATUS: genotype decode is pinned by the reference's own
 * inst/GenoMat_*.raw fixtures (tests/test_oracle_golden.py).  No statistic
 * or p-value is pinned by any file in the reference and R is not installed
 * in the build container, so for every numeric output this oracle is
 * "parity unpinned": it is cross-checked only against independent
 * numpy / scipy / mpmath restatements written inside
 * tests/test_oracle_golden.py, and -- whenever an Rscript is on PATH --
 * against the reference itself (tools/run_reference.R driven by
 * tests/test_vs_rscript.py).
 *
 * R numeric semantics mirrored here:
 *   - sum(x): sequential long-double accumulation of doubles that were
 *     rounded to double element-wise first;
 *   - mean(x): long-double mean, plus the second refinement pass for
 *     REALSXP input (selectable, see orc_set_mean_mode);
 *   - pnorm: Cody (1969) rational approximations as in R's nmath;
 *   - pcauchy / CCT cancellation, .Machine$double.eps^0.25 == 2^-13.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef long double ld;

#define ORC_NA (orc_na_real())

/* R's NA_real_: quiet NaN with payload 1954 */
static double orc_na_real(void) {
    union { double d; uint64_t u; } x;
    x.u = 0x7FF00000000007A2ULL;
    return x.d;
}
double orc_na(void) { return orc_na_real(); }

/* mean mode: 0 = shared contract RN53(A/n) (integer-valued data);
 *            1 = R-faithful long-double mean (with 2nd pass for doubles)   */
static int g_mean_mode = 0;
void orc_set_mean_mode(int m) { g_mean_mode = m; }
/* G.model of the OUTER *_one_SNP functions: 0 "Add", 1 "Dom", 2 "Rec" (R/SPAGxECCT.R:572-574, :1045-1047,
 * R/SPAGxEPlus_functions_MYZ.R:285-287).  The inner SPA_G_one_SNP_homo_new call of SPAGxE_CCT never receives
 * G.model (:598, :616), so orc_spa_homo always runs its prologue with "Add". */
static int g_gmodel = 0;
void orc_set_gmodel(int m) { g_gmodel = m; }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* R's sum() on a double vector: LDOUBLE accumulation.                 */
static double r_sum(const double *x, long n) {
    ld s = 0.0L;
    for (long i = 0; i < n; i++) s += x[i];
    return (double)s;
}

/* mean(g, na.rm=TRUE): R/SPAGxECCT.R:557, :2000.
 * g_is_int: the R vector was INTSXP (single long-double pass).         */
static double r_mean_narm(const double *g, long n, int g_is_int, long *n_obs_out) {
    ld s = 0.0L; long m = 0;
    for (long i = 0; i < n; i++) if (!isnan(g[i])) { s += g[i]; m++; }
    if (n_obs_out) *n_obs_out = m;
    if (m == 0) return NAN;
    if (g_mean_mode == 0) {
        /* contract: one correctly rounded double division */
        return (double)s / (double)m;
    }
    s /= (ld)m;
    if (!g_is_int) {
        ld t = 0.0L;
        for (long i = 0; i < n; i++) if (!isnan(g[i])) t += ((ld)g[i] - s);
        s += t / (ld)m;
    }
    return (double)s;
}

/* ------------------------------------------------------------------ */
/* PLINK .bed 2-bit decode, A1-count convention of GRAB::GRAB.ReadGeno
 * (call sites R/SPAGxECCT.R:484-487); pinned by inst/GenoMat_*.raw:
 * 00 -> 2, 01 -> NA, 10 -> 1, 11 -> 0; sample 0 in the low bits.       */
void orc_bed_decode_row(const uint8_t *row, long n, double *g) {
    static const double lut_val[4] = {2.0, 0.0 /*NA*/, 1.0, 0.0};
    for (long i = 0; i < n; i++) {
        int c = (row[i >> 2] >> (2 * (i & 3))) & 3;
        g[i] = (c == 1) ? orc_na_real() : lut_val[c];
    }
}

/* per-code counts of one packed row: cnt[c] for c = 0..3 */
void orc_bed_counts(const uint8_t *row, long n, long cnt[4]) {
    cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
    for (long i = 0; i < n; i++) cnt[(row[i >> 2] >> (2 * (i & 3))) & 3]++;
}

/* ------------------------------------------------------------------ */
/* stats::pnorm (R nmath pnorm_both, Cody 1969); used at
 * R/SPAGxECCT.R:585, :1871, :2043, :2056-2057.                         */
#define ORC_M_SQRT_32 5.656854249492380195206754896838
#define ORC_M_1_SQRT_2PI 0.398942280401432677939946059934

static void pnorm_both(double x, double *cum, double *ccum, int i_tail) {
    static const double a[5] = {2.2352520354606839287, 161.02823106855587881,
        1067.6894854603709582, 18154.981253343561249, 0.065682337918207449113};
    static const double b[4] = {47.20258190468824187, 976.09855173777669322,
        10260.932208618978205, 45507.789335026729956};
    static const double c[9] = {0.39894151208813466764, 8.8831497943883759412,
        93.506656132177855979, 597.27027639480026226, 2494.5375852903726711,
        6848.1904505362823326, 11602.651437647350124, 9842.7148383839780218,
        1.0765576773720192317e-8};
    static const double d[8] = {22.266688044328115691, 235.38790178262499861,
        1519.377599407554805, 6485.558298266760755, 18615.571640885098091,
        34900.952721145977266, 38912.003286093271411, 19685.429676859990727};
    static const double p[6] = {0.21589853405795699, 0.1274011611602473639,
        0.022235277870649807, 0.001421619193227893466, 2.9112874951168792e-5,
        0.02307344176494017303};
    static const double q[5] = {1.28426009614491121, 0.468238212480865118,
        0.0659881378689285515, 0.00378239633202758244, 7.29751555083966205e-5};
    double xden, xnum, temp, del, eps, xsq, y;
    int i, lower, upper;
    if (isnan(x)) { *cum = *ccum = NAN; return; }
    eps = DBL_EPSILON * 0.5;
    lower = i_tail != 1;
    upper = i_tail != 0;
    y = fabs(x);
    if (y <= 0.67448975) {
        if (y > eps) {
            xsq = x * x;
            xnum = a[4] * xsq; xden = xsq;
            for (i = 0; i < 3; ++i) { xnum = (xnum + a[i]) * xsq; xden = (xden + b[i]) * xsq; }
        } else xnum = xden = 0.0;
        temp = x * (xnum + a[3]) / (xden + b[3]);
        if (lower) *cum = 0.5 + temp;
        if (upper) *ccum = 0.5 - temp;
    } else if (y <= ORC_M_SQRT_32) {
        xnum = c[8] * y; xden = y;
        for (i = 0; i < 7; ++i) { xnum = (xnum + c[i]) * y; xden = (xden + d[i]) * y; }
        temp = (xnum + c[7]) / (xden + d[7]);
        xsq = trunc(y * 16) / 16;
        del = (y - xsq) * (y + xsq);
        *cum = exp(-xsq * xsq * 0.5) * exp(-del * 0.5) * temp;
        *ccum = 1.0 - *cum;
        if (x > 0.) { temp = *cum; if (lower) *cum = *ccum; *ccum = temp; }
    } else if ((lower && -37.5193 < x && x < 8.2924) || (upper && -8.2924 < x && x < 37.5193)) {
        xsq = 1.0 / (x * x);
        xnum = p[5] * xsq; xden = xsq;
        for (i = 0; i < 4; ++i) { xnum = (xnum + p[i]) * xsq; xden = (xden + q[i]) * xsq; }
        temp = xsq * (xnum + p[4]) / (xden + q[4]);
        temp = (ORC_M_1_SQRT_2PI - temp) / y;
        xsq = trunc(x * 16) / 16;
        del = (x - xsq) * (x + xsq);
        *cum = exp(-xsq * xsq * 0.5) * exp(-del * 0.5) * temp;
        *ccum = 1.0 - *cum;
        if (x > 0.) { temp = *cum; if (lower) *cum = *ccum; *ccum = temp; }
    } else {
        if (x > 0) { *cum = 1.; *ccum = 0.; } else { *cum = 0.; *ccum = 1.; }
    }
}

double orc_pnorm(double x, int lower_tail) {
    double p, cp;
    if (isnan(x)) return NAN;
    if (isinf(x)) return (x < 0) ? (lower_tail ? 0. : 1.) : (lower_tail ? 1. : 0.);
    pnorm_both(x, &p, &cp, lower_tail ? 0 : 1);
    return lower_tail ? p : cp;
}

/* stats::pcauchy(x) (location 0, scale 1, lower tail); R/SPAGxECCT.R:2120 */
double orc_pcauchy(double x) {
    if (isnan(x)) return NAN;
    if (isinf(x)) return (x < 0) ? 0. : 1.;
    if (fabs(x) > 1) {
        double y = atan(1. / x) / M_PI;
        return (x > 0) ? (0.5 - y + 0.5) : (-y);
    }
    return 0.5 + atan(x) / M_PI;
}

/* CCT of two p-values with equal weights: R/SPAGxECCT.R:2070-2123.
 * return 0 ok; 1 = "Cannot have NAs"; 2 = "must be between 0 and 1";
 * 3 = "Cannot have both 0 and 1".  *warn = 1 if a p-value was exactly 1. */
int orc_cct2(double p1, double p2, double *out, int *warn) {
    double pv[2] = {p1, p2};
    if (warn) *warn = 0;
    *out = NAN;
    if (isnan(p1) || isnan(p2)) return 1;
    if ((p1 < 0) + (p2 < 0) + (p1 > 1) + (p2 > 1) > 0) return 2;
    int is_zero = (p1 == 0) || (p2 == 0);
    int is_one = (p1 == 1) || (p2 == 1);
    if (is_zero && is_one) return 3;
    if (is_zero) { *out = 0; return 0; }
    if (is_one) { if (warn) *warn = 1; for (int i = 0; i < 2; i++) if (pv[i] == 1) pv[i] = 0.999; }
    double w = 1.0 / 2.0;
    int small0 = pv[0] < 1e-16, small1 = pv[1] < 1e-16;
    double stat;
    if (!small0 && !small1) {
        ld s = 0.0L;
        for (int i = 0; i < 2; i++) s += w * tan((0.5 - pv[i]) * M_PI);
        stat = (double)s;
    } else {
        ld s = 0.0L;
        for (int i = 0; i < 2; i++) if (pv[i] < 1e-16) s += (w / pv[i]) / M_PI;
        stat = (double)s;
        ld s2 = 0.0L;
        for (int i = 0; i < 2; i++) if (!(pv[i] < 1e-16)) s2 += w * tan((0.5 - pv[i]) * M_PI);
        stat = stat + (double)s2;
    }
    if (stat > 1e15) *out = (1 / stat) / M_PI;
    else *out = 1 - orc_pcauchy(stat);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Binomial(2,p) MGF/CGF pieces: R/SPAGxECCT.R:1652-1681 (evaluated the
 * naive way so that overflow -> NaN happens where the reference's does) */
static inline double M_G0(double t, double maf) { double e = exp(t); double u = (1 - maf + maf * e); return u * u; }
static inline double M_G1(double t, double maf) { double e = exp(t); return 2 * (maf * e) * (1 - maf + maf * e); }
static inline double M_G2(double t, double maf) { double e = exp(t); double x = maf * e; return 2 * (x * x) + 2 * (maf * e) * (1 - maf + maf * e); }
static inline double K_G0(double t, double maf) { return log(M_G0(t, maf)); }
static inline double K_G1(double t, double maf) { return M_G1(t, maf) / M_G0(t, maf); }
static inline double K_G2(double t, double maf) {
    double m0 = M_G0(t, maf), m1 = M_G1(t, maf), m2 = M_G2(t, maf);
    return (m0 * m2 - m1 * m1) / (m0 * m0);
}
#define MAF_AT(i) (mafvec ? mafvec[i] : maf)

/* H_org :1683, H1_adj :1694, H2 :1706 */
static double H_org(double t, const double *R, const double *mafvec, double maf, long n) {
    ld s = 0.0L;
    for (long i = 0; i < n; i++) s += K_G0(t * R[i], MAF_AT(i));
    return (double)s;
}
static double H1_adj(double t, const double *R, double sstat, const double *mafvec, double maf, long n) {
    ld s = 0.0L;
    for (long i = 0; i < n; i++) s += R[i] * K_G1(t * R[i], MAF_AT(i));
    return (double)s - sstat;
}
static double H2(double t, const double *R, const double *mafvec, double maf, long n) {
    ld s = 0.0L;
    for (long i = 0; i < n; i++) s += (R[i] * R[i]) * K_G2(t * R[i], MAF_AT(i));
    return (double)s;
}
static inline double r_sign(double x) { return isnan(x) ? NAN : (x > 0) - (x < 0); }

/* fastgetroot_H1_new: R/SPAGxECCT.R:1795-1850.
 * root may be +-Inf or NaN (the reference's NA).                        */
void orc_spa_root(const double *R, const double *mafvec, double maf, long n, double sstat,
                  double init_t, double *root, int *iter_out, int *converge, double *h2_out) {
    const double tol = 0x1p-13; /* .Machine$double.eps^0.25 */
    const int maxiter = 100;
    double t = init_t, H1 = 0, diff_t = INFINITY, H2e = 0;
    int conv = 1, iter;
    for (iter = 1; iter <= maxiter; iter++) {
        double old_t = t, old_diff = diff_t, old_H1 = H1;
        H1 = H1_adj(t, R, sstat, mafvec, maf, n);
        H2e = H2(t, R, mafvec, maf, n);
        diff_t = -1 * H1 / H2e;
        if (isnan(diff_t)) diff_t = 5;
        if (isnan(H1) || isinf(H1)) { t = r_sign(sstat) * INFINITY; H2e = 0; break; }
        if (r_sign(H1) != r_sign(old_H1)) {
            int guard = 0;
            while (fabs(diff_t) > fabs(old_diff) - tol && guard++ < 1200) diff_t = diff_t / 2;
        }
        if (isnan(diff_t)) diff_t = 5;
        if (fabs(diff_t) < tol) break;
        t = old_t + diff_t;
    }
    if (iter > maxiter) iter = maxiter; /* R's loop variable stays at maxiter */
    if (iter == maxiter) { conv = 0; t = NAN; }
    *root = t; if (iter_out) *iter_out = iter; if (converge) *converge = conv; if (h2_out) *h2_out = H2e;
}

/* GetProb_SPA_G_new: R/SPAGxECCT.R:1855-1879 */
double orc_getprob_spa(const double *R, const double *mafvec, double maf, long n, double sstat,
                       int lower_tail, int *iter_out) {
    double zeta, h2; int it, cv;
    orc_spa_root(R, mafvec, maf, n, sstat, 0.0, &zeta, &it, &cv, &h2);
    if (iter_out) *iter_out = it;
    double k1 = H_org(zeta, R, mafvec, maf, n);
    double k2 = H2(zeta, R, mafvec, maf, n);
    double temp1 = zeta * sstat - k1;
    double w = r_sign(zeta) * sqrt(2 * temp1);
    double v = zeta * sqrt(k2);
    double pval = orc_pnorm(w + 1 / w * log(v / w), lower_tail);
    if (isnan(pval)) pval = 0;
    return pval;
}

/* ------------------------------------------------------------------ */
/* genotype prologue: R/SPAGxECCT.R:557-577 (dup :1030-1051, :2000-2020,
 * R/SPAGxEPlus_functions_MYZ.R:270-290).  g is modified in place.
 * returns 1 if MAF < min.maf (early NA row).  gm: G.model 0 Add, 1 Dom, 2 Rec. */
static int prologue(double *g, long n, int g_is_int, double min_maf, double *maf_out, double *miss_out, int *touched, int gm) {
    int tch = 0;
    long n_obs;
    double MAF = r_mean_narm(g, n, g_is_int, &n_obs) / 2;
    double missing_rate = (double)(n - n_obs) / (double)n;
    if (missing_rate != 0) {
        double imp = 2 * MAF;
        for (long i = 0; i < n; i++) if (isnan(g[i])) g[i] = imp;
        tch = 1;
    }
    if (MAF > 0.5) {
        MAF = 1 - MAF;
        for (long i = 0; i < n; i++) g[i] = 2 - g[i];
        tch = 1;
    }
    if (gm == 1) { for (long i = 0; i < n; i++) g[i] = (g[i] >= 1) ? 1.0 : 0.0; tch = 1; }   /* ifelse(g>=1,1,0) */
    if (gm == 2) { for (long i = 0; i < n; i++) g[i] = (g[i] <= 1) ? 0.0 : 1.0; tch = 1; }   /* ifelse(g<=1,0,1) */
    *maf_out = MAF; *miss_out = missing_rate;
    if (touched) *touched = tch;
    return (MAF < min_maf);
}

/* SPA_G_one_SNP_homo_new: R/SPAGxECCT.R:1991-2065.
 * out[7] = MAF, missing.rate, p.spa, p.norm, S, S.var, z
 * info bit0: filtered by min.maf; bit1: SPA taken; *iters = Newton its  */
void orc_spa_homo(const double *g_in, int g_is_int, const double *R, long n, double cutoff,
                  double min_maf, double out[7], int *info, int iters[2]) {
    double *g = (double *)malloc(sizeof(double) * n);
    memcpy(g, g_in, sizeof(double) * n);
    double MAF, miss; int touched = 0;
    if (info) *info = 0;
    if (iters) iters[0] = iters[1] = 0;
    if (prologue(g, n, g_is_int, min_maf, &MAF, &miss, &touched, 0)) {
        out[0] = MAF; out[1] = miss;
        for (int k = 2; k < 7; k++) out[k] = ORC_NA;
        if (info) *info |= 1;
        free(g); return;
    }
    ld acc = 0.0L;
    for (long i = 0; i < n; i++) acc += g[i] * R[i];
    double S = (double)acc;
    double gvar = 2 * MAF * (1 - MAF);
    acc = 0.0L;
    for (long i = 0; i < n; i++) acc += (R[i] * R[i]) * gvar;
    double Svar = (double)acc;
    acc = 0.0L;
    for (long i = 0; i < n; i++) acc += R[i] * MAF;
    double Smean = 2 * (double)acc;
    double z = (S - Smean) / sqrt(Svar);
    if (fabs(z) < cutoff) {
        double pn = orc_pnorm(fabs(z), 0) * 2;
        out[0] = MAF; out[1] = miss; out[2] = pn; out[3] = pn; out[4] = S; out[5] = Svar; out[6] = z;
        free(g); return;
    }
    if (info) *info |= 2;
    double s_hi = fmax(S, 2 * Smean - S), s_lo = fmin(S, 2 * Smean - S);
    int it1 = 0, it2 = 0;
    double pval1 = orc_getprob_spa(R, NULL, MAF, n, s_hi, 0, &it1);
    double pval2 = orc_getprob_spa(R, NULL, MAF, n, s_lo, 1, &it2);
    if (iters) { iters[0] = it1; iters[1] = it2; }
    if (isnan(pval2)) pval2 = pval1;
    double pval3 = orc_pnorm(fabs(z), 0), pval4 = orc_pnorm(-fabs(z), 1);
    out[0] = MAF; out[1] = miss; out[2] = pval1 + pval2; out[3] = pval3 + pval4;
    out[4] = S; out[5] = Svar; out[6] = z;
    free(g);
}

/* ------------------------------------------------------------------ */
/* Least squares by Householder QR (no pivoting; rank check in the spirit
 * of LINPACK dqrdc2's limited pivoting: a column whose remaining norm
 * falls below tol * original norm is declared aliased).  Stands in for
 * stats:::Cdqrls used by lm()/glm.fit() (call sites R/SPAGxECCT.R:638,
 * :652, :1066, :1088).  A is n x q column-major, overwritten.
 * coef[q]; covu[q*q] = (R'R)^-1; returns rank (q if full).             */
static int qr_ls(double *A, double *yv, long n, int q, double tol, double *coef, double *covu, double *rss_out) {
    ld *cn0 = (ld *)malloc(sizeof(ld) * q);
    double *Rm = (double *)calloc((size_t)q * q, sizeof(double));
    int rank = q;
    for (int j = 0; j < q; j++) { ld s = 0; for (long i = 0; i < n; i++) s += (ld)A[i + (long)j * n] * A[i + (long)j * n]; cn0[j] = sqrtl(s); }
    for (int j = 0; j < q; j++) {
        double *aj = A + (long)j * n;
        ld s = 0; for (long i = j; i < n; i++) s += (ld)aj[i] * aj[i];
        ld nrm = sqrtl(s);
        if (nrm <= tol * cn0[j] || cn0[j] == 0) { rank = j; break; }
        ld alpha = (aj[j] > 0) ? -nrm : nrm;
        ld v0 = aj[j] - alpha;
        /* v = (v0, aj[j+1..]) ; H = I - 2 v v'/(v'v) */
        ld vtv = v0 * v0 + (s - (ld)aj[j] * aj[j]);
        aj[j] = (double)v0;
        for (int k = j + 1; k < q; k++) {
            double *ak = A + (long)k * n;
            ld d = 0; for (long i = j; i < n; i++) d += (ld)aj[i] * ak[i];
            ld f = 2 * d / vtv;
            for (long i = j; i < n; i++) ak[i] = (double)(ak[i] - f * aj[i]);
        }
        {
            ld d = 0; for (long i = j; i < n; i++) d += (ld)aj[i] * yv[i];
            ld f = 2 * d / vtv;
            for (long i = j; i < n; i++) yv[i] = (double)(yv[i] - f * aj[i]);
        }
        Rm[j + j * q] = (double)alpha;
        for (int k = j + 1; k < q; k++) Rm[j + k * q] = A[j + (long)k * n];
    }
    free(cn0);
    if (rank < q) { free(Rm); return rank; }
    /* back substitution R coef = (Q'y)[0..q) */
    for (int j = q - 1; j >= 0; j--) {
        ld s = yv[j];
        for (int k = j + 1; k < q; k++) s -= (ld)Rm[j + k * q] * coef[k];
        coef[j] = (double)(s / Rm[j + j * q]);
    }
    if (rss_out) { ld s = 0; for (long i = q; i < n; i++) s += (ld)yv[i] * yv[i]; *rss_out = (double)s; }
    if (covu) {
        /* Rinv then Rinv Rinv' */
        double *Ri = (double *)calloc((size_t)q * q, sizeof(double));
        for (int c = 0; c < q; c++) {
            for (int j = q - 1; j >= 0; j--) {
                ld s = (j == c) ? 1.0L : 0.0L;
                for (int k = j + 1; k < q; k++) s -= (ld)Rm[j + k * q] * Ri[k + c * q];
                Ri[j + c * q] = (double)(s / Rm[j + j * q]);
            }
        }
        for (int a = 0; a < q; a++) for (int b2 = 0; b2 < q; b2++) {
            ld s = 0; for (int k = 0; k < q; k++) s += (ld)Ri[a + k * q] * Ri[b2 + k * q];
            covu[a + b2 * q] = (double)s;
        }
        free(Ri);
    }
    free(Rm);
    return rank;
}

/* lgamma-based regularized incomplete beta I_x(a,b) (continued fraction,
 * modified Lentz).  Stands in for stats::pt -> pbeta (TOMS 708), used by
 * summary.lm at R/SPAGxECCT.R:654 and :1069.                            */
static ld betacf(ld a, ld b, ld x) {
    const ld FPMIN = 1e-4000L, EPS = 1e-19L;
    ld qab = a + b, qap = a + 1, qam = a - 1, c = 1, d = 1 - qab * x / qap;
    if (fabsl(d) < FPMIN) d = FPMIN;
    d = 1 / d; ld h = d;
    for (int m = 1; m <= 100000; m++) {
        int m2 = 2 * m;
        ld aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1 + aa * d; if (fabsl(d) < FPMIN) d = FPMIN;
        c = 1 + aa / c; if (fabsl(c) < FPMIN) c = FPMIN;
        d = 1 / d; h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1 + aa * d; if (fabsl(d) < FPMIN) d = FPMIN;
        c = 1 + aa / c; if (fabsl(c) < FPMIN) c = FPMIN;
        d = 1 / d; ld del = d * c; h *= del;
        if (fabsl(del - 1) < EPS) break;
    }
    return h;
}
/* I_x(a,b) with y = 1 - x supplied separately (no cancellation when x is close to 1) */
static ld betainc_reg(ld a, ld b, ld x, ld y) {
    if (x <= 0) return 0; if (y <= 0) return 1;
    ld lbt = lgammal(a + b) - lgammal(a) - lgammal(b) + a * ((y < 0.5L) ? log1pl(-y) : logl(x)) + b * logl(y);
    if (x < (a + 1) / (a + b + 2)) return expl(lbt) * betacf(a, b, x) / a;
    return 1 - expl(lbt) * betacf(b, a, y) / b;
}
/* 2*pt(|t|, df, lower.tail=FALSE) = I_{df/(df+t^2)}(df/2, 1/2)  (R: pt -> pbeta, both tails kept
 * accurate; here the continued fraction is evaluated on the small side directly) */
double orc_pt2(double t, double df) {
    if (isnan(t) || isnan(df)) return NAN;
    ld x2 = (ld)t * t, den = (ld)df + x2;
    return (double)betainc_reg((ld)df / 2, 0.5L, (ld)df / den, x2 / den);
}

/* glm.fit(family=binomial) IRLS restatement: mustart
 * (y+0.5)/2, |dev-devold|/(|dev|+0.1) < 1e-8, maxit 25; covariance from
 * the last WLS (summary.glm).  X n x q col-major (intercept included).
 * returns 0 ok, 1 rank-deficient, 2 non-finite deviance / invalid mu.
 * mu_out (optional) = fitted values.                                    */
int orc_glm_binomial(const double *X, const double *y, long n, int q, double *coef, double *se,
                     int *iters, int *converged, double *mu_out) {
    const double THRESH = 30., MTHRESH = -30., INVEPS = 1 / DBL_EPSILON;
    double *eta = (double *)malloc(sizeof(double) * n), *mu = (double *)malloc(sizeof(double) * n);
    double *A = (double *)malloc(sizeof(double) * n * q), *zw = (double *)malloc(sizeof(double) * n);
    double *covu = (double *)malloc(sizeof(double) * q * q);
    int rc = 0, conv = 0, it;
    ld dsum = 0;
    for (long i = 0; i < n; i++) {
        double mus = (y[i] + 0.5) / 2;
        eta[i] = log(mus / (1 - mus));
        double tmp = (eta[i] < MTHRESH) ? DBL_EPSILON : ((eta[i] > THRESH) ? INVEPS : exp(eta[i]));
        mu[i] = tmp / (1 + tmp);
        double yl = (y[i] != 0.) ? y[i] * log(y[i] / mu[i]) : 0;
        double y1 = 1 - y[i];
        double yl1 = (y1 != 0.) ? y1 * log(y1 / (1 - mu[i])) : 0;
        dsum += 2 * (yl + yl1);
    }
    double devold = (double)dsum, dev = devold;
    for (it = 1; it <= 25; it++) {
        for (long i = 0; i < n; i++) {
            double varmu = mu[i] * (1 - mu[i]);
            double opexp = 1 + exp(eta[i]);
            double mev = (eta[i] > THRESH || eta[i] < MTHRESH) ? DBL_EPSILON : exp(eta[i]) / (opexp * opexp);
            double z = eta[i] + (y[i] - mu[i]) / mev;
            double w = sqrt((mev * mev) / varmu);
            for (int j = 0; j < q; j++) A[i + (long)j * n] = X[i + (long)j * n] * w;
            zw[i] = z * w;
        }
        int rank = qr_ls(A, zw, n, q, 1e-11, coef, covu, NULL);
        if (rank < q) { rc = 1; break; }
        dsum = 0; int valid = 1;
        for (long i = 0; i < n; i++) {
            double e = 0;
            { ld s = 0; for (int j = 0; j < q; j++) s += (ld)X[i + (long)j * n] * coef[j]; e = (double)s; }
            eta[i] = e;
            double tmp = (e < MTHRESH) ? DBL_EPSILON : ((e > THRESH) ? INVEPS : exp(e));
            mu[i] = tmp / (1 + tmp);
            if (!(mu[i] > 0 && mu[i] < 1) || !isfinite(mu[i])) valid = 0;
            double yl = (y[i] != 0.) ? y[i] * log(y[i] / mu[i]) : 0;
            double y1 = 1 - y[i];
            double yl1 = (y1 != 0.) ? y1 * log(y1 / (1 - mu[i])) : 0;
            dsum += 2 * (yl + yl1);
        }
        dev = (double)dsum;
        if (!isfinite(dev) || !valid) { rc = 2; break; }
        if (fabs(dev - devold) / (fabs(dev) + 0.1) < 1e-8) { conv = 1; break; }
        devold = dev;
    }
    if (it > 25) it = 25;
    if (rc == 0) for (int j = 0; j < q; j++) se[j] = sqrt(covu[j + j * q]);
    if (mu_out) memcpy(mu_out, mu, sizeof(double) * n);
    if (iters) *iters = it; if (converged) *converged = conv;
    free(eta); free(mu); free(A); free(zw); free(covu);
    return rc;
}

/* lm() + summary.lm: coef, se, two-sided t p-values. returns 0 / 1(rank) */
int orc_lm(const double *X, const double *y, long n, int q, double *coef, double *se, double *pval, double *fitted) {
    double *A = (double *)malloc(sizeof(double) * n * q), *yy = (double *)malloc(sizeof(double) * n);
    double *covu = (double *)malloc(sizeof(double) * q * q);
    memcpy(A, X, sizeof(double) * n * q); memcpy(yy, y, sizeof(double) * n);
    double rss;
    int rank = qr_ls(A, yy, n, q, 1e-7, coef, covu, &rss);
    if (rank < q) { free(A); free(yy); free(covu); return 1; }
    double rdf = (double)(n - q), resvar = rss / rdf;
    for (int j = 0; j < q; j++) {
        se[j] = sqrt(covu[j + j * q] * resvar);
        if (pval) pval[j] = orc_pt2(coef[j] / se[j], rdf);
    }
    if (fitted) for (long i = 0; i < n; i++) { ld s = 0; for (int j = 0; j < q; j++) s += (ld)X[i + (long)j * n] * coef[j]; fitted[i] = (double)s; }
    free(A); free(yy); free(covu);
    return 0;
}

/* ------------------------------------------------------------------ */
/* survival::coxph(Surv(time, status) ~ X, ties = "efron", iter.max)  (call site R/SPAGxECCT.R:624).
 * `survival` is a third-party package absent from the reference tree (DESCRIPTION Imports, no version pin); this is a
 * restatement of its published fitter coxph.fit -> coxfit6.c: subjects ordered by time, covariates centred and scaled
 * (columns whose values all lie in {-1, 0, 1} are left alone: coxph.control(nocenter)), accumulation from the largest
 * time down, Efron approximation for tied deaths, beta0 = 0, Newton-Raphson with
 * |1 - loglik_old / loglik_new| <= eps (1e-9) as the stop and step halving when the likelihood decreases, LDL'
 * (cholesky2, toler.chol = eps^0.75), variance = inverse information at the last iterate.
 * X is n x q column-major.  coef[q], se[q]; returns 0, or 1 when the information matrix is rank deficient.  */
static double coxsafe(double x) { return x < -200 ? -200 : (x > 22 ? 22 : x); }
static int cholesky2(double *m, int n, double toler) {   /* m[i*n + j], lower triangle holds L, diagonal D */
    double eps = 0;
    for (int i = 0; i < n; i++) { if (m[i * n + i] > eps) eps = m[i * n + i]; for (int j = i + 1; j < n; j++) m[j * n + i] = m[i * n + j]; }
    eps = (eps == 0) ? toler : eps * toler;
    int rank = 0;
    for (int i = 0; i < n; i++) {
        double pivot = m[i * n + i];
        if (!isfinite(pivot) || pivot < eps) { m[i * n + i] = 0; continue; }
        rank++;
        for (int j = i + 1; j < n; j++) {
            double temp = m[j * n + i] / pivot;
            m[j * n + i] = temp;
            m[j * n + j] -= temp * temp * pivot;
            for (int k = j + 1; k < n; k++) m[k * n + j] -= temp * m[k * n + i];
        }
    }
    return rank;
}
static void chsolve2(const double *m, int n, double *y) {
    for (int i = 0; i < n; i++) { double t = y[i]; for (int j = 0; j < i; j++) t -= y[j] * m[i * n + j]; y[i] = t; }
    for (int i = n - 1; i >= 0; i--) {
        if (m[i * n + i] == 0) { y[i] = 0; continue; }
        double t = y[i] / m[i * n + i];
        for (int j = i + 1; j < n; j++) t -= y[j] * m[j * n + i];
        y[i] = t;
    }
}
typedef struct { double t; long i; } cox_key;
static int cox_key_cmp(const void *a, const void *b) {
    const cox_key *x = (const cox_key *)a, *y = (const cox_key *)b;
    if (x->t < y->t) return -1;
    if (x->t > y->t) return 1;
    return (x->i > y->i) - (x->i < y->i);   /* stable, like R's order() */
}
/* one pass of coxfit6's accumulation loop at `beta`: loglik, u[q], imat[q*q] (upper triangle filled like imat[j][i]) */
static double cox_pass(const double *cov, const double *stime, const int *status, long n, int q, const double *beta,
                       double *u, double *imat) {
    double a[16], a2[16], cmat[256], cmat2[256];
    double loglik = 0, denom = 0;
    for (int i = 0; i < q; i++) { u[i] = 0; a[i] = 0; a2[i] = 0; for (int j = 0; j < q; j++) { imat[i * q + j] = 0; cmat[i * q + j] = 0; cmat2[i * q + j] = 0; } }
    for (long person = n - 1; person >= 0;) {
        double dtime = stime[person];
        int ndead = 0; double denom2 = 0;
        while (person >= 0 && stime[person] == dtime) {
            double zbeta = 0;
            for (int i = 0; i < q; i++) zbeta += beta[i] * cov[(long)i * n + person];
            zbeta = coxsafe(zbeta);
            double risk = exp(zbeta);
            if (status[person] == 0) {
                denom += risk;
                for (int i = 0; i < q; i++) {
                    a[i] += risk * cov[(long)i * n + person];
                    for (int j = 0; j <= i; j++) cmat[i * q + j] += risk * cov[(long)i * n + person] * cov[(long)j * n + person];
                }
            } else {
                ndead++;
                denom2 += risk;
                loglik += zbeta;
                for (int i = 0; i < q; i++) {
                    u[i] += cov[(long)i * n + person];
                    a2[i] += risk * cov[(long)i * n + person];
                    for (int j = 0; j <= i; j++) cmat2[i * q + j] += risk * cov[(long)i * n + person] * cov[(long)j * n + person];
                }
            }
            person--;
        }
        if (ndead > 0) {   /* Efron: the deaths enter the sums in ndead equal steps */
            for (int k = 0; k < ndead; k++) {
                denom += denom2 / ndead;
                loglik -= log(denom);
                for (int i = 0; i < q; i++) {
                    a[i] += a2[i] / ndead;
                    double temp2 = a[i] / denom;
                    u[i] -= temp2;
                    for (int j = 0; j <= i; j++) {
                        cmat[i * q + j] += cmat2[i * q + j] / ndead;
                        imat[j * q + i] += (cmat[i * q + j] - temp2 * a[j]) / denom;
                    }
                }
            }
            for (int i = 0; i < q; i++) { a2[i] = 0; for (int j = 0; j < q; j++) cmat2[i * q + j] = 0; }
        }
    }
    return loglik;
}
int orc_coxph(const double *X, const double *time, const double *event, long n, int q, int maxiter,
              double *coef, double *se, int *iters_out) {
    if (q > 16) return 2;
    cox_key *key = (cox_key *)malloc(sizeof(cox_key) * n);
    for (long i = 0; i < n; i++) { key[i].t = time[i]; key[i].i = i; }
    qsort(key, n, sizeof(cox_key), cox_key_cmp);
    double *cov = (double *)malloc(sizeof(double) * n * q), *stime = (double *)malloc(sizeof(double) * n);
    int *status = (int *)malloc(sizeof(int) * n);
    for (long k = 0; k < n; k++) { stime[k] = key[k].t; status[k] = event[key[k].i] != 0; }
    double scale[16];
    for (int c = 0; c < q; c++) {
        int zero_one = 1;
        for (long k = 0; k < n; k++) { double v = X[(long)c * n + key[k].i]; cov[(long)c * n + k] = v; if (v != -1 && v != 0 && v != 1) zero_one = 0; }
        scale[c] = 1.0;
        if (!zero_one) {
            double temp = 0;
            for (long k = 0; k < n; k++) temp += cov[(long)c * n + k];
            temp /= (double)n;
            for (long k = 0; k < n; k++) cov[(long)c * n + k] -= temp;
            temp = 0;
            for (long k = 0; k < n; k++) temp += fabs(cov[(long)c * n + k]);
            temp = (temp > 0) ? (double)n / temp : 1.0;
            scale[c] = temp;
            for (long k = 0; k < n; k++) cov[(long)c * n + k] *= temp;
        }
    }
    const double eps = 1e-9, toler = pow(DBL_EPSILON, 0.75);
    double beta[16], newbeta[16], u[16], a[16], imat[256];
    for (int i = 0; i < q; i++) beta[i] = 0;
    double loglik = cox_pass(cov, stime, status, n, q, beta, u, imat);
    for (int i = 0; i < q; i++) a[i] = u[i];
    int rank = cholesky2(imat, q, toler);
    chsolve2(imat, q, a);
    for (int i = 0; i < q; i++) newbeta[i] = beta[i] + a[i];
    int halving = 0, iter, rc = (rank < q);
    for (iter = 1; iter <= maxiter && !rc; iter++) {
        double newlk = cox_pass(cov, stime, status, n, q, newbeta, u, imat);
        rank = cholesky2(imat, q, toler);
        if (rank < q) { rc = 1; break; }
        if ((fabs(1 - (loglik / newlk)) <= eps && halving == 0) || iter == maxiter) break;
        if (!(newlk >= loglik)) {
            halving = 1;
            for (int i = 0; i < q; i++) newbeta[i] = (newbeta[i] + beta[i]) / 2;
        } else {
            halving = 0;
            loglik = newlk;
            chsolve2(imat, q, u);
            for (int i = 0; i < q; i++) { beta[i] = newbeta[i]; newbeta[i] = newbeta[i] + u[i]; }
        }
    }
    if (iters_out) *iters_out = iter;
    if (!rc) {   /* chinv2: diagonal of the inverse, column by column through the factorisation; undo the scaling */
        for (int k = 0; k < q; k++) {
            double e[16];
            for (int i = 0; i < q; i++) e[i] = (i == k);
            chsolve2(imat, q, e);
            coef[k] = newbeta[k] * scale[k];
            se[k] = sqrt(e[k] * scale[k] * scale[k]);
        }
    }
    free(key); free(cov); free(stime); free(status);
    return rc;
}

/* ------------------------------------------------------------------ */
/* ordinal::clm(as.factor(Y) ~ X) with the defaults of the call site R/SPAGxECCT.R:666 (logit link, flexible
 * thresholds): the proportional-odds model P(Y <= j) = F(alpha_j - x'beta), j = 1 .. J-1.  `ordinal` is a third-party
 * package that the reference does not even import; what is restated is the model, not clm's iteration path: clm runs
 * Newton-Raphson to max|gradient| < 1e-6 on a strictly concave likelihood, i.e. to the maximum-likelihood estimate
 * within ~1e-11, and reports z = estimate / sqrt(diag(solve(-Hessian))), p = 2 pnorm(-|z|).  Here: Newton-Raphson from
 * alpha_j = qlogis(j / J), beta = 0, step halving while the likelihood drops, stop when the Newton step is below
 * 1e-10 (1 + max|theta|); the Hessian of that last evaluation gives the standard errors.
 * X is n x q column-major (no intercept), ycat[i] in 0 .. J-1.  coef / se: J-1 thresholds, then q slopes.
 * returns 0, 1 singular Hessian, 2 no convergence in 100 iterations.                                              */
int orc_clm(const double *X, const int *ycat, long n, int q, int J, double *coef, double *se, int *iters_out) {
    const int nth = J - 1, np = nth + q;
    if (J < 2 || np > 24) return 3;
    double th[24], prev[24], g[24], H[24 * 24], step[24];
    for (int j = 0; j < nth; j++) { double pr = (double)(j + 1) / (double)J; th[j] = log(pr / (1 - pr)); }
    for (int a = 0; a < q; a++) th[nth + a] = 0;
    double ll_old = -INFINITY; int halv = 0, iter, rc = 2;
    memcpy(prev, th, sizeof th);
    for (iter = 0; iter < 100; iter++) {
        ld ll = 0;
        for (int a = 0; a < np; a++) { g[a] = 0; for (int b = 0; b < np; b++) H[a * np + b] = 0; }
        for (long i = 0; i < n; i++) {
            double eta = 0;
            for (int a = 0; a < q; a++) eta += th[nth + a] * X[(long)a * n + i];
            const int y = ycat[i], u = (y < nth) ? y : -1, l = (y > 0) ? y - 1 : -1;
            double Fu = 1, fu = 0, su = 0, Fl = 0, fl = 0, sl = 0;   /* F, F' = F(1-F), F'' = F'(1-2F) */
            if (u >= 0) { Fu = 1 / (1 + exp(-(th[u] - eta))); fu = Fu * (1 - Fu); su = fu * (1 - 2 * Fu); }
            if (l >= 0) { Fl = 1 / (1 + exp(-(th[l] - eta))); fl = Fl * (1 - Fl); sl = fl * (1 - 2 * Fl); }
            const double pi = Fu - Fl, rp = 1 / pi, d = (fu - fl) * rp;
            ll += log(pi);
            const double cu = fu * rp, cl = -fl * rp;
            if (u >= 0) { g[u] += cu; H[u * np + u] += su * rp - cu * cu; }
            if (l >= 0) { g[l] += cl; H[l * np + l] += -sl * rp - cl * cl; }
            if (u >= 0 && l >= 0) { H[u * np + l] += fu * fl * rp * rp; H[l * np + u] += fu * fl * rp * rp; }
            const double hub = -su * rp + cu * d, hlb = sl * rp + cl * d, hbb = (su - sl) * rp - d * d;
            for (int a = 0; a < q; a++) {
                const double xa = X[(long)a * n + i];
                g[nth + a] -= d * xa;
                if (u >= 0) { H[u * np + nth + a] += hub * xa; H[(nth + a) * np + u] += hub * xa; }
                if (l >= 0) { H[l * np + nth + a] += hlb * xa; H[(nth + a) * np + l] += hlb * xa; }
                for (int b = 0; b <= a; b++) {
                    const double v = hbb * xa * X[(long)b * n + i];
                    H[(nth + a) * np + nth + b] += v;
                    if (b != a) H[(nth + b) * np + nth + a] += v;
                }
            }
        }
        const double llv = (double)ll;
        if (iter > 0 && !(llv >= ll_old) && halv < 30) {   /* the step overshot: halve it */
            for (int a = 0; a < np; a++) th[a] = (th[a] + prev[a]) / 2;
            halv++; iter--;
            continue;
        }
        halv = 0;
        /* Newton step: solve (-H) step = g */
        double A[24 * 24];
        for (int a = 0; a < np * np; a++) A[a] = -H[a];
        for (int a = 0; a < np; a++) step[a] = g[a];
        if (cholesky2(A, np, pow(DBL_EPSILON, 0.75)) < np) { rc = 1; break; }
        chsolve2(A, np, step);
        double ms = 0, mt = 0;
        for (int a = 0; a < np; a++) { if (fabs(step[a]) > ms) ms = fabs(step[a]); if (fabs(th[a]) > mt) mt = fabs(th[a]); }
        if (ms < 1e-10 * (1 + mt)) {
            for (int k = 0; k < np; k++) {
                double e[24];
                for (int a = 0; a < np; a++) e[a] = (a == k);
                chsolve2(A, np, e);
                coef[k] = th[k]; se[k] = sqrt(e[k]);
            }
            rc = 0; break;
        }
        memcpy(prev, th, sizeof th);
        ll_old = llv;
        for (int a = 0; a < np; a++) th[a] += step[a];
    }
    if (iters_out) *iters_out = iter;
    return rc;
}

/* Wald p-value of the E:g term for Y ~ Cova + E + g + E:g
 * (R/SPAGxECCT.R:622-677).  trait: 1 binary (glm binomial), 2 quantitative (lm), 3 survival (coxph: no intercept
 * column; Y then holds 2n doubles, surv.time followed by event), 4 categorical (clm; Y = category index 0 .. J-1).
 * returns 0 ok, else error code; *pw = NA on failure.                    */
static int wald_eg(int trait, const double *g, const double *E, const double *Y, const double *Cova, int p, long n, double *pw) {
    if (trait == 3) {   /* coxph(Surv(surv.time, event) ~ as.matrix(Cova.mtx) + E + g + g*E, iter.max = 1000): ref :624 */
        int qc = p + 3;
        double *Xc = (double *)malloc(sizeof(double) * n * qc), cf[16], sc[16];
        for (int j = 0; j < p; j++) memcpy(Xc + (long)j * n, Cova + (long)j * n, sizeof(double) * n);
        for (long i = 0; i < n; i++) { Xc[i + (long)p * n] = E[i]; Xc[i + (long)(p + 1) * n] = g[i]; Xc[i + (long)(p + 2) * n] = E[i] * g[i]; }
        int itc, rcx = orc_coxph(Xc, Y, Y + n, n, qc, 1000, cf, sc, &itc);
        *pw = ORC_NA;
        if (rcx == 0) *pw = 2 * orc_pnorm(-fabs(cf[qc - 1] / sc[qc - 1]), 1);   /* pchisq(z^2, 1, lower.tail = FALSE) */
        free(Xc);
        return rcx;
    }
    if (trait == 4) {   /* clm(as.factor(Phen.mtx$Y) ~ as.matrix(Cova.mtx) + E + g + g*E): ref :666; Y holds the category indices 0 .. J-1 */
        int qc = p + 3, J = 0;
        int *yc = (int *)malloc(sizeof(int) * n);
        for (long i = 0; i < n; i++) { yc[i] = (int)Y[i]; if (yc[i] + 1 > J) J = yc[i] + 1; }
        double *Xc = (double *)malloc(sizeof(double) * n * qc), cf[24], sc[24];
        for (int j = 0; j < p; j++) memcpy(Xc + (long)j * n, Cova + (long)j * n, sizeof(double) * n);
        for (long i = 0; i < n; i++) { Xc[i + (long)p * n] = E[i]; Xc[i + (long)(p + 1) * n] = g[i]; Xc[i + (long)(p + 2) * n] = E[i] * g[i]; }
        int itc, rcx = orc_clm(Xc, yc, n, qc, J, cf, sc, &itc);
        *pw = ORC_NA;
        if (rcx == 0) *pw = 2 * orc_pnorm(-fabs(cf[J - 1 + qc - 1] / sc[J - 1 + qc - 1]), 1);
        free(Xc); free(yc);
        return rcx;
    }
    int q = p + 4;
    double *X = (double *)malloc(sizeof(double) * n * q);
    for (long i = 0; i < n; i++) X[i] = 1.0;
    for (int j = 0; j < p; j++) memcpy(X + (long)(1 + j) * n, Cova + (long)j * n, sizeof(double) * n);
    for (long i = 0; i < n; i++) { X[i + (long)(p + 1) * n] = E[i]; X[i + (long)(p + 2) * n] = g[i]; X[i + (long)(p + 3) * n] = E[i] * g[i]; }
    double *coef = (double *)malloc(sizeof(double) * q), *se = (double *)malloc(sizeof(double) * q);
    int rc;
    *pw = ORC_NA;
    if (trait == 1) {
        int it, cv;
        rc = orc_glm_binomial(X, Y, n, q, coef, se, &it, &cv, NULL);
        if (rc == 0) *pw = 2 * orc_pnorm(-fabs(coef[q - 1] / se[q - 1]), 1);
    } else {
        double *pv = (double *)malloc(sizeof(double) * q);
        rc = orc_lm(X, Y, n, q, coef, se, pv, NULL);
        if (rc == 0) *pw = pv[q - 1];
        free(pv);
    }
    free(X); free(coef); free(se);
    return rc;
}

/* R0 = R - W (W'W)^-1 W'R with W = [1, g]: R/SPAGxECCT.R:606-609.
 * returns 1 if W'W is singular (the reference's solve() would stop()).   */
static int adjust_for_g(const double *g, const double *R, long n, double *R0) {
    ld s1 = 0, sg = 0, sgg = 0, r0 = 0, r1 = 0;
    for (long i = 0; i < n; i++) { s1 += 1; sg += g[i]; sgg += (ld)g[i] * g[i]; r0 += R[i]; r1 += (ld)g[i] * R[i]; }
    double a = (double)s1, b = (double)sg, d = (double)sgg;
    double det = a * d - b * b;
    if (det == 0 || !isfinite(det)) return 1;
    double i00 = d / det, i01 = -b / det, i11 = a / det;
    double w0 = (double)r0, w1 = (double)r1;
    for (long i = 0; i < n; i++) {
        double wi0 = i00 + g[i] * i01, wi1 = i01 + g[i] * i11;
        R0[i] = R[i] - (wi0 * w0 + wi1 * w1);
    }
    return 0;
}

/* SPAGxE_CCT_one_SNP: R/SPAGxECCT.R:543-683.  trait 1=binary 2=quantitative.
 * g may contain NaN (missing).  Cutoff is NOT forwarded by the reference
 * (:598,:616) so the inner cutoff is always 2; branch A inner min.maf is
 * 1e-5, branch B forwards min_maf.
 * route bits: 1 filtered, 2 branch B, 4 SPA taken, 8 wald failed, 16 CCT stop, 32 solve singular
 * returns 0, or the CCT stop code (>0) mirroring stop().                */
int orc_cct_one_snp(int trait, const double *g_in, int g_is_int, const double *R, const double *E,
                    const double *Y, const double *Cova, int p, long n, double epsilon, double min_maf,
                    double out[10], int *route, int iters[2]) {
    double *g = (double *)malloc(sizeof(double) * n);
    memcpy(g, g_in, sizeof(double) * n);
    double MAF, miss; int rt = 0, rc = 0, touched = 0;
    if (iters) iters[0] = iters[1] = 0;
    if (prologue(g, n, g_is_int, min_maf, &MAF, &miss, &touched, g_gmodel)) {
        out[0] = MAF; out[1] = miss; for (int k = 2; k < 10; k++) out[k] = ORC_NA;
        if (route) *route = 1; free(g); return 0;
    }
    double VarG = 2 * MAF * (1 - MAF);
    ld acc = 0; for (long i = 0; i < n; i++) acc += g[i] * R[i];
    double S1 = (double)acc;
    acc = 0; for (long i = 0; i < n; i++) acc += R[i] * R[i];
    double sumR2 = (double)acc;
    double VarS1 = sumR2 * VarG;
    double Z1 = S1 / sqrt(VarS1);
    double pnorm1 = orc_pnorm(-fabs(Z1), 1) * 2;
    double po[5];
    /* after imputation / flip the vector is REALSXP unless untouched */
    int inner_is_int = g_is_int && !touched;
    if (pnorm1 > epsilon) {
        acc = 0; for (long i = 0; i < n; i++) acc += E[i] * (R[i] * R[i]);
        double lambda = (double)acc / sumR2;
        double *Rn = (double *)malloc(sizeof(double) * n);
        for (long i = 0; i < n; i++) Rn[i] = (E[i] - lambda) * R[i];
        double res[7]; int info;
        orc_spa_homo(g, inner_is_int, Rn, n, 2.0, 0.00001, res, &info, iters);
        if (info & 2) rt |= 4;
        po[0] = res[2]; po[1] = res[2]; po[2] = res[2]; po[3] = res[3]; po[4] = pnorm1;
        free(Rn);
    } else {
        rt |= 2;
        double *R0 = (double *)malloc(sizeof(double) * n);
        if (adjust_for_g(g, R, n, R0)) {
            rt |= 32; for (int k = 0; k < 4; k++) po[k] = ORC_NA; po[4] = pnorm1;
        } else {
            for (long i = 0; i < n; i++) R0[i] = E[i] * R0[i]; /* R.new0 */
            double res[7]; int info;
            orc_spa_homo(g, inner_is_int, R0, n, 2.0, min_maf, res, &info, iters);
            if (info & 2) rt |= 4;
            double pw; int wrc = wald_eg(trait, g, E, Y, Cova, p, n, &pw);
            if (wrc) rt |= 8;
            double pc = ORC_NA; int warn;
            int crc = wrc ? 1 : orc_cct2(res[2], pw, &pc, &warn);
            if (crc) { rt |= 16; rc = crc; pc = ORC_NA; }
            po[0] = res[2]; po[1] = pw; po[2] = pc; po[3] = res[3]; po[4] = pnorm1;
        }
        free(R0);
    }
    out[0] = MAF; out[1] = miss; for (int k = 0; k < 5; k++) out[2 + k] = po[k];
    out[7] = S1; out[8] = VarS1; out[9] = Z1;
    if (route) *route = rt;
    free(g);
    return rc;
}

/* SPAGxE_CCT loop body over a column-major N x M double matrix
 * (R/SPAGxECCT.R:509-525); out is M x 10 column-major.                  */
int orc_cct_matrix(int trait, const double *G, int g_is_int, long n, long m, const double *R, const double *E,
                   const double *Y, const double *Cova, int p, double epsilon, double min_maf,
                   double *out, int *route, int *iters) {
    int rc_all = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (long j = 0; j < m; j++) {
        double o[10]; int rt, it[2];
        int rc = orc_cct_one_snp(trait, G + j * n, g_is_int, R, E, Y, Cova, p, n, epsilon, min_maf, o, &rt, it);
        for (int k = 0; k < 10; k++) out[j + k * m] = o[k];
        if (route) route[j] = rt;
        if (iters) { iters[2 * j] = it[0]; iters[2 * j + 1] = it[1]; }
        if (rc) {
#pragma omp critical
            rc_all = rc;
        }
    }
    return rc_all;
}

/* same, genotypes given as packed .bed rows (pitch bytes per SNP)        */
int orc_cct_bed(int trait, const uint8_t *bed, long pitch, long n, long m, const double *R, const double *E,
                const double *Y, const double *Cova, int p, double epsilon, double min_maf,
                double *out, int *route, int *iters) {
    int rc_all = 0;
#pragma omp parallel
    {
        double *g = (double *)malloc(sizeof(double) * n);
#pragma omp for schedule(dynamic, 1)
        for (long j = 0; j < m; j++) {
            double o[10]; int rt, it[2];
            orc_bed_decode_row(bed + j * pitch, n, g);
            int rc = orc_cct_one_snp(trait, g, 0, R, E, Y, Cova, p, n, epsilon, min_maf, o, &rt, it);
            for (int k = 0; k < 10; k++) out[j + k * m] = o[k];
            if (route) route[j] = rt;
            if (iters) { iters[2 * j] = it[0]; iters[2 * j + 1] = it[1]; }
            if (rc) {
#pragma omp critical
                rc_all = rc;
            }
        }
        free(g);
    }
    return rc_all;
}

/* ------------------------------------------------------------------ */
/* SPAGxEmix_CCT_one_SNP: R/SPAGxECCT.R:1013-1269.
 * topPCs n x k col-major.  out[12].
 * route bits as above plus 64: MAC<=20 uniform AF, 128: negative-ratio
 * fallback entered, 256: logistic AF used.                              */
static double two_tail_spa(const double *Rv, const double *mafvec, double maf, long n, double S, double Smean,
                           double z, double *pnorm_out, int iters[2]) {
    double s_hi = fmax(S, 2 * Smean - S), s_lo = fmin(S, 2 * Smean - S);
    int i1 = 0, i2 = 0;
    double pval1 = orc_getprob_spa(Rv, mafvec, maf, n, s_hi, 0, &i1);
    double pval2 = orc_getprob_spa(Rv, mafvec, maf, n, s_lo, 1, &i2);
    if (iters) { iters[0] = i1; iters[1] = i2; }
    if (isnan(pval2)) pval2 = pval1;
    *pnorm_out = orc_pnorm(fabs(z), 0) + orc_pnorm(-fabs(z), 1);
    return pval1 + pval2;
}

/* individual-level allele frequencies of SPAGxEmix_CCT_one_SNP (R/SPAGxECCT.R:1055-1100; the same block opens
 * SPAGxEmix_CCT_Plus_one_SNP, R/SPAmixGxECCTPlus_functions_MYZ_2024-09-17.R:190-232): g is the imputed / flipped
 * genotype vector, m receives MAF.est, *negnum_out MAF.est.negative.num; returns the route bits 64 / 128 / 256.    */
static int mix_af_estimate(const double *g, long n, double MAF, double MAC, const double *PCs, int k, double pc_pcut,
                           double neg_ratio_cut, double *m, double *negnum_out) {
    int rt = 0;
    double negnum = 0;
    if (MAC <= 20) {
        for (long i = 0; i < n; i++) m[i] = MAF;
        rt |= 64;
    } else {
        int q = k + 1;
        double *X = (double *)malloc(sizeof(double) * n * q);
        for (long i = 0; i < n; i++) X[i] = 1.0;
        memcpy(X + n, PCs, sizeof(double) * n * k);
        double *coef = (double *)malloc(sizeof(double) * q), *se = (double *)malloc(sizeof(double) * q), *pv = (double *)malloc(sizeof(double) * q);
        int lrc = orc_lm(X, g, n, q, coef, se, pv, m);
        if (lrc) { for (int j = 0; j < q; j++) pv[j] = NAN; }
        long nn = 0;
        for (long i = 0; i < n; i++) { m[i] = m[i] / 2; if (m[i] < 0) nn++; }
        negnum = (double)nn;
        double negratio = negnum / (double)n;
        if (negratio > neg_ratio_cut) {
            rt |= 128;
            int nsel = 0; int *sel = (int *)malloc(sizeof(int) * k);
            for (int j = 0; j < k; j++) if (pv[1 + j] < pc_pcut) sel[nsel++] = j;
            if (nsel == 0) {
                for (long i = 0; i < n; i++) m[i] = MAF;
            } else {
                ld sg = 0, s2g = 0;
                double *gt = (double *)malloc(sizeof(double) * n);
                for (long i = 0; i < n; i++) { gt[i] = nearbyint(g[i]); sg += gt[i]; s2g += 2 - gt[i]; }
                if ((double)sg == 0 || (double)s2g == 0) {
                    for (long i = 0; i < n; i++) m[i] = MAF;
                } else {
                    for (long i = 0; i < n; i++) if (gt[i] != 0) gt[i] = 1;
                    int q2 = nsel + 1;
                    double *X2 = (double *)malloc(sizeof(double) * n * q2);
                    for (long i = 0; i < n; i++) X2[i] = 1.0;
                    for (int j = 0; j < nsel; j++) memcpy(X2 + (long)(1 + j) * n, PCs + (long)sel[j] * n, sizeof(double) * n);
                    double *c2 = (double *)malloc(sizeof(double) * q2), *s2 = (double *)malloc(sizeof(double) * q2);
                    double *mu = (double *)malloc(sizeof(double) * n);
                    int it, cv;
                    int grc = orc_glm_binomial(X2, gt, n, q2, c2, s2, &it, &cv, mu);
                    (void)grc;
                    for (long i = 0; i < n; i++) m[i] = 1 - sqrt(1 - mu[i]);
                    rt |= 256;
                    free(X2); free(c2); free(s2); free(mu);
                }
                free(gt);
            }
            free(sel);
        } else {
            for (long i = 0; i < n; i++) { if (m[i] <= 0) m[i] = 0; if (m[i] >= 1) m[i] = 1; }
        }
        free(X); free(coef); free(se); free(pv);
    }
    *negnum_out = negnum;
    return rt;
}

int orc_mix_one_snp(int trait, const double *g_in, int g_is_int, const double *R, const double *E,
                    const double *Y, const double *Cova, int p, const double *PCs, int k, long n,
                    double epsilon, double cutoff, double min_maf, double pc_pcut, double neg_ratio_cut,
                    double out[12], int *route, int iters[2], double *mafest_out) {
    double *g = (double *)malloc(sizeof(double) * n);
    memcpy(g, g_in, sizeof(double) * n);
    double MAF, miss; int rt = 0, rc = 0, touched = 0;
    if (iters) iters[0] = iters[1] = 0;
    if (prologue(g, n, g_is_int, min_maf, &MAF, &miss, &touched, g_gmodel)) {
        out[0] = MAF; out[1] = miss; for (int j = 2; j < 12; j++) out[j] = ORC_NA;
        if (route) *route = 1; free(g); return 0;
    }
    double MAC = MAF * 2 * (double)n;
    double *m = (double *)malloc(sizeof(double) * n);
    double negnum = 0;
    rt |= mix_af_estimate(g, n, MAF, MAC, PCs, k, pc_pcut, neg_ratio_cut, m, &negnum);
    if (mafest_out) memcpy(mafest_out, m, sizeof(double) * n);
    double *gv = (double *)malloc(sizeof(double) * n);
    for (long i = 0; i < n; i++) gv[i] = 2 * m[i] * (1 - m[i]);
    ld acc = 0; for (long i = 0; i < n; i++) acc += g[i] * R[i];
    double S1 = (double)acc;
    acc = 0; for (long i = 0; i < n; i++) acc += R[i] * m[i];
    double meanS1 = 2 * (double)acc;
    acc = 0; for (long i = 0; i < n; i++) acc += (R[i] * R[i]) * gv[i];
    double VarS1 = (double)acc;
    double Z1 = (S1 - meanS1) / sqrt(VarS1);
    double pnorm1 = orc_pnorm(-fabs(Z1), 1) * 2;
    double po[5];
    if (pnorm1 > epsilon) {
        acc = 0; for (long i = 0; i < n; i++) acc += (g[i] * E[i]) * R[i];
        double S2 = (double)acc;
        ld a1 = 0, a2 = 0;
        for (long i = 0; i < n; i++) { a1 += (gv[i] * E[i]) * (R[i] * R[i]); a2 += gv[i] * (R[i] * R[i]); }
        double lambda = (double)a1 / (double)a2;
        double S_GxE = S2 - lambda * S1;
        double *Rn = (double *)malloc(sizeof(double) * n);
        for (long i = 0; i < n; i++) Rn[i] = (E[i] - lambda) * R[i];
        acc = 0; for (long i = 0; i < n; i++) acc += Rn[i] * m[i];
        double Smean = 2 * (double)acc;
        acc = 0; for (long i = 0; i < n; i++) acc += (Rn[i] * Rn[i]) * gv[i];
        double Svar = (double)acc;
        double z = (S_GxE - Smean) / sqrt(Svar);
        if (fabs(z) < cutoff) {
            double pn = orc_pnorm(fabs(z), 0) * 2;
            po[0] = po[1] = po[2] = po[3] = pn; po[4] = pnorm1;
        } else {
            rt |= 4;
            double pn; double ps = two_tail_spa(Rn, m, 0, n, S_GxE, Smean, z, &pn, iters);
            po[0] = po[1] = po[2] = ps; po[3] = pn; po[4] = pnorm1;
        }
        free(Rn);
    } else {
        rt |= 2;
        double *R0 = (double *)malloc(sizeof(double) * n);
        if (adjust_for_g(g, R, n, R0)) {
            rt |= 32; for (int j = 0; j < 4; j++) po[j] = ORC_NA; po[4] = pnorm1;
        } else {
            acc = 0; for (long i = 0; i < n; i++) acc += (g[i] * E[i]) * R0[i];
            double S0 = (double)acc;
            for (long i = 0; i < n; i++) R0[i] = E[i] * R0[i];
            acc = 0; for (long i = 0; i < n; i++) acc += R0[i] * m[i];
            double Smean = 2 * (double)acc;
            acc = 0; for (long i = 0; i < n; i++) acc += (R0[i] * R0[i]) * gv[i];
            double Svar = (double)acc;
            double z = (S0 - Smean) / sqrt(Svar);
            double pspa, pn;
            if (fabs(z) < cutoff) { pn = orc_pnorm(fabs(z), 0) * 2; pspa = pn; }
            else { rt |= 4; pspa = two_tail_spa(R0, m, 0, n, S0, Smean, z, &pn, iters); }
            double pw; int wrc = wald_eg(trait, g, E, Y, Cova, p, n, &pw);
            if (wrc) rt |= 8;
            if (isnan(pw)) pw = pspa; /* :1194,:1213,:1232 */
            double pc = ORC_NA; int warn;
            int crc = orc_cct2(pspa, pw, &pc, &warn);
            if (crc) { rt |= 16; rc = crc; pc = ORC_NA; }
            po[0] = pspa; po[1] = pw; po[2] = pc; po[3] = pn; po[4] = pnorm1;
        }
        free(R0);
    }
    out[0] = MAF; out[1] = miss; for (int j = 0; j < 5; j++) out[2 + j] = po[j];
    out[7] = S1; out[8] = VarS1; out[9] = Z1; out[10] = negnum; out[11] = MAC;
    if (route) *route = rt;
    free(g); free(m); free(gv);
    return rc;
}

int orc_mix_bed(int trait, const uint8_t *bed, long pitch, long n, long mm, const double *R, const double *E,
                const double *Y, const double *Cova, int p, const double *PCs, int k,
                double epsilon, double cutoff, double min_maf, double pc_pcut, double neg_ratio_cut,
                double *out, int *route, int *iters) {
    int rc_all = 0;
#pragma omp parallel
    {
        double *g = (double *)malloc(sizeof(double) * n);
#pragma omp for schedule(dynamic, 1)
        for (long j = 0; j < mm; j++) {
            double o[12]; int rt, it[2];
            orc_bed_decode_row(bed + j * pitch, n, g);
            int rc = orc_mix_one_snp(trait, g, 0, R, E, Y, Cova, p, PCs, k, n, epsilon, cutoff, min_maf,
                                     pc_pcut, neg_ratio_cut, o, &rt, it, NULL);
            for (int c = 0; c < 12; c++) out[j + c * mm] = o[c];
            if (route) route[j] = rt;
            if (iters) { iters[2 * j] = it[0]; iters[2 * j + 1] = it[1]; }
            if (rc) {
#pragma omp critical
                rc_all = rc;
            }
        }
        free(g);
    }
    return rc_all;
}

/* ------------------------------------------------------------------ */
/* SPAGxEmixCCT_localance (R/SPAGxECCT.R:1401-1647): ancestry-specific G x E test with local ancestry.
 * g = ancestry-specific genotype (0/1/2, never NA: the reference stops on `if (MAF > 0.5)` with an NA), h = number of
 * haplotypes of that ancestry (0/1/2); G ~ Binomial(h_i, MAF).                                                      */
static inline double r_pow_small(double x, double y) {   /* R_pow for the exponents that occur here (-2 .. 2) */
    if (y == 2.0) return x * x;
    if (y == 1.0) return x;
    if (y == 0.0) return 1.0;
    return pow(x, y);
}
/* M_G0_local :2132, M_G1_local :2138, M_G2_local :2144, K_G*_local :2152-2168 */
static inline double MGl0(double t, double maf, double h) { return r_pow_small(1 - maf + maf * exp(t), h); }
static inline double MGl1(double t, double maf, double h) { return h * (maf * exp(t)) * r_pow_small(1 - maf + maf * exp(t), h - 1); }
static inline double MGl2(double t, double maf, double h) {
    double me = maf * exp(t), u = 1 - maf + maf * exp(t);
    return h * (h - 1) * (me * me) * r_pow_small(u, h - 2) + h * me * r_pow_small(u, h - 1);
}
static double Hl_org(double t, const double *R, double maf, const double *h, long n) {
    ld s = 0; for (long i = 0; i < n; i++) s += log(MGl0(t * R[i], maf, h[i])); return (double)s;
}
static double Hl1_adj(double t, const double *R, double sstat, double maf, const double *h, long n) {
    ld s = 0; for (long i = 0; i < n; i++) s += R[i] * (MGl1(t * R[i], maf, h[i]) / MGl0(t * R[i], maf, h[i])); return (double)s - sstat;
}
static double Hl2(double t, const double *R, double maf, const double *h, long n) {
    ld s = 0;
    for (long i = 0; i < n; i++) {
        double m0 = MGl0(t * R[i], maf, h[i]), m1 = MGl1(t * R[i], maf, h[i]), m2 = MGl2(t * R[i], maf, h[i]);
        s += (R[i] * R[i]) * ((m0 * m2 - m1 * m1) / (m0 * m0));
    }
    return (double)s;
}
/* GetProb_SPA_G_new_local :2287-2311 with fastgetroot_H1_new_local :2223-2281 */
static double getprob_spa_local(const double *R, double maf, const double *h, long n, double sstat, int lower_tail, int *iter_out) {
    const double tol = 0x1p-13; const int maxiter = 100;
    double t = 0, H1 = 0, diff_t = INFINITY;
    int iter;
    for (iter = 1; iter <= maxiter; iter++) {
        double old_t = t, old_diff = diff_t, old_H1 = H1;
        H1 = Hl1_adj(t, R, sstat, maf, h, n);
        double H2e = Hl2(t, R, maf, h, n);
        diff_t = -1 * H1 / H2e;
        if (isnan(diff_t)) diff_t = 5;
        if (isnan(H1) || isinf(H1)) { t = r_sign(sstat) * INFINITY; break; }
        if (r_sign(H1) != r_sign(old_H1)) {
            int guard = 0;
            while (fabs(diff_t) > fabs(old_diff) - tol && guard++ < 1200) diff_t = diff_t / 2;
        }
        if (isnan(diff_t)) diff_t = 5;
        if (fabs(diff_t) < tol) break;
        t = old_t + diff_t;
    }
    if (iter > maxiter) iter = maxiter;
    if (iter == maxiter) t = NAN;
    if (iter_out) *iter_out = iter;
    double k1 = Hl_org(t, R, maf, h, n), k2 = Hl2(t, R, maf, h, n);
    double temp1 = t * sstat - k1;
    double w = r_sign(t) * sqrt(2 * temp1), v = t * sqrt(k2);
    double pval = orc_pnorm(w + 1 / w * log(v / w), lower_tail);
    if (isnan(pval)) pval = 0;
    return pval;
}
/* SPAmix_localance_one_SNP :2324-2395 (Cutoff = 2, G.model = "Add": the caller does not forward them).
 * g is modified in place.  res[0] = p.spa (res[3] of the reference), res[1] = p.norm; info bit0 filtered, bit1 SPA  */
static void spa_localance(double *g, const double *R, const double *h, long n, double min_maf, double res[2], int *info, int iters[2]) {
    ld sg = 0, sh = 0;
    for (long i = 0; i < n; i++) { sg += g[i]; sh += h[i]; }
    double MAF = (double)sg / (double)sh;
    *info = 0; if (iters) iters[0] = iters[1] = 0;
    if (MAF > 0.5) { MAF = 1 - MAF; for (long i = 0; i < n; i++) g[i] = 2 - g[i]; }
    if (MAF < min_maf) { res[0] = res[1] = ORC_NA; *info = 1; return; }
    ld a = 0, b = 0, c = 0;
    for (long i = 0; i < n; i++) {
        a += g[i] * R[i];
        b += (MAF * h[i]) * R[i];
        c += (R[i] * R[i]) * ((h[i] * MAF) * (1 - MAF));
    }
    double S = (double)a, Smean = (double)b, Svar = (double)c;
    double z = (S - Smean) / sqrt(Svar);
    if (fabs(z) < 2.0) { res[0] = res[1] = orc_pnorm(fabs(z), 0) * 2; return; }
    *info = 2;
    int i1 = 0, i2 = 0;
    double p1 = getprob_spa_local(R, MAF, h, n, fmax(S, 2 * Smean - S), 0, &i1);
    double p2 = getprob_spa_local(R, MAF, h, n, fmin(S, 2 * Smean - S), 1, &i2);
    if (iters) { iters[0] = i1; iters[1] = i2; }
    res[0] = p1 + p2;
    res[1] = orc_pnorm(fabs(z), 0) + orc_pnorm(-fabs(z), 1);
}
/* columns of X (n x q column-major) that lm() / glm() would report as aliased (dqrdc2's limited pivoting moves a column
 * whose remaining norm falls below tol x its norm to the end; the fit then ignores it): keep[j] = 0 for those.
 * returns the rank                                                                                                  */
static int drop_aliased(const double *X, long n, int q, double tol, int *keep) {
    double *Q = (double *)malloc(sizeof(double) * n * q);
    int rank = 0;
    for (int j = 0; j < q; j++) {
        double *v = Q + (long)rank * n;
        ld n0 = 0;
        for (long i = 0; i < n; i++) { v[i] = X[i + (long)j * n]; n0 += (ld)v[i] * v[i]; }
        for (int pass = 0; pass < 2; pass++)
            for (int k = 0; k < rank; k++) {
                const double *u = Q + (long)k * n;
                ld d = 0; for (long i = 0; i < n; i++) d += (ld)u[i] * v[i];
                for (long i = 0; i < n; i++) v[i] = (double)(v[i] - d * u[i]);
            }
        ld n1 = 0; for (long i = 0; i < n; i++) n1 += (ld)v[i] * v[i];
        if (n0 == 0 || sqrtl(n1) <= tol * sqrtl(n0)) { keep[j] = 0; continue; }
        ld inv = 1 / sqrtl(n1);
        for (long i = 0; i < n; i++) v[i] = (double)(v[i] * inv);
        keep[j] = 1; rank++;
    }
    free(Q);
    return rank;
}
/* Wald model of the local-ancestry test (ref :1593, :1607, :1621): binary glm(Y ~ Cova.haplo + Cova + E + g + g*E),
 * quantitative lm(Y ~ Cova.haplo + Cova + haplo_numVec + E + g + g*E); aliased columns (two-way admixture with both
 * ancestries' counts in Cova.haplo.mtx.list: h1 + h2 = 2) are dropped as lm / glm do.  returns 0, 1 ("E:g" itself
 * aliased or rank failure), 2 (IRLS invalid)                                                                        */
static int wald_localance(int trait, const double *g, const double *h, const double *E, const double *Y, const double *Cova, int p,
                          const double *CH, int L, long n, double *pw) {
    const int q0 = 1 + L + p + (trait == 2 ? 1 : 0) + 3;
    double *X = (double *)malloc(sizeof(double) * n * q0);
    int c = 0;
    for (long i = 0; i < n; i++) X[i] = 1.0;
    c = 1;
    for (int j = 0; j < L; j++, c++) memcpy(X + (long)c * n, CH + (long)j * n, sizeof(double) * n);
    for (int j = 0; j < p; j++, c++) memcpy(X + (long)c * n, Cova + (long)j * n, sizeof(double) * n);
    if (trait == 2) { memcpy(X + (long)c * n, h, sizeof(double) * n); c++; }
    for (long i = 0; i < n; i++) { X[i + (long)c * n] = E[i]; X[i + (long)(c + 1) * n] = g[i]; X[i + (long)(c + 2) * n] = E[i] * g[i]; }
    int keep[64];
    *pw = ORC_NA;
    drop_aliased(X, n, q0, 1e-7, keep);
    if (!keep[q0 - 1]) { free(X); return 1; }
    int q = 0;
    for (int j = 0; j < q0; j++) if (keep[j]) { if (q != j) memmove(X + (long)q * n, X + (long)j * n, sizeof(double) * n); q++; }
    double *coef = (double *)malloc(sizeof(double) * q), *se = (double *)malloc(sizeof(double) * q);
    int rc;
    if (trait == 1) {
        int it, cv;
        rc = orc_glm_binomial(X, Y, n, q, coef, se, &it, &cv, NULL);
        if (rc == 0) *pw = 2 * orc_pnorm(-fabs(coef[q - 1] / se[q - 1]), 1);
    } else {
        double *pv = (double *)malloc(sizeof(double) * q);
        rc = orc_lm(X, Y, n, q, coef, se, pv, NULL);
        if (rc == 0) *pw = pv[q - 1];
        free(pv);
    }
    free(X); free(coef); free(se);
    return rc;
}
/* SPAGxEmixCCT_localance_one_SNP :1496-1647.  CH: the L columns of Cova.haplo.mtx.list for this SNP (n x L).
 * trait 1 binary, 2 quantitative.  route bits as orc_cct_one_snp; rc 40: g holds NA (the reference stops).          */
int orc_localance_one_snp(int trait, const double *g_in, const double *h, const double *R, const double *E, const double *Y,
                          const double *Cova, int p, const double *CH, int L, long n, double epsilon, double min_maf,
                          double out[10], int *route, int iters[2]) {
    double *g = (double *)malloc(sizeof(double) * n);
    memcpy(g, g_in, sizeof(double) * n);
    int rt = 0, rc = 0;
    if (iters) iters[0] = iters[1] = 0;
    ld sg = 0, sh = 0;
    for (long i = 0; i < n; i++) {
        if (isnan(g[i])) { for (int k = 0; k < 10; k++) out[k] = ORC_NA; if (route) *route = 1; free(g); return 40; }
        sg += g[i]; sh += h[i];
    }
    double MAF = (double)sg / (double)sh, miss = 0.0;
    if (MAF > 0.5) { MAF = 1 - MAF; for (long i = 0; i < n; i++) g[i] = 2 - g[i]; }
    if (g_gmodel == 1) for (long i = 0; i < n; i++) g[i] = (g[i] >= 1) ? 1.0 : 0.0;
    if (g_gmodel == 2) for (long i = 0; i < n; i++) g[i] = (g[i] <= 1) ? 0.0 : 1.0;
    if (MAF < min_maf || isnan(MAF)) {
        out[0] = MAF; out[1] = miss; for (int k = 2; k < 10; k++) out[k] = ORC_NA;
        if (route) *route = 1; free(g); return 0;
    }
    ld a = 0, b = 0, c = 0;
    for (long i = 0; i < n; i++) {
        a += g[i] * R[i];
        b += (MAF * h[i]) * R[i];
        c += (R[i] * R[i]) * ((h[i] * MAF) * (1 - MAF));
    }
    double S1 = (double)a, S1mean = (double)b, VarS1 = (double)c;
    double Z1 = (S1 - S1mean) / sqrt(VarS1);
    double pnorm1 = orc_pnorm(-fabs(Z1), 1) * 2;
    double po[5];
    double *Rn = (double *)malloc(sizeof(double) * n);
    if (pnorm1 > epsilon) {
        ld u = 0, v = 0;
        for (long i = 0; i < n; i++) { u += (h[i] * E[i]) * (R[i] * R[i]); v += h[i] * (R[i] * R[i]); }
        double lambda = (double)u / (double)v;
        for (long i = 0; i < n; i++) Rn[i] = (E[i] - lambda) * R[i];
        double res[2]; int info;
        spa_localance(g, Rn, h, n, min_maf, res, &info, iters);
        if (info & 2) rt |= 4;
        po[0] = res[0]; po[1] = res[0]; po[2] = res[0]; po[3] = res[1]; po[4] = pnorm1;
    } else {
        rt |= 2;
        if (adjust_for_g(g, R, n, Rn)) {
            rt |= 32; for (int k = 0; k < 4; k++) po[k] = ORC_NA; po[4] = pnorm1;
        } else {
            for (long i = 0; i < n; i++) Rn[i] = E[i] * Rn[i];
            double *g2 = (double *)malloc(sizeof(double) * n);   /* the inner function may flip its own copy of g */
            memcpy(g2, g, sizeof(double) * n);
            double res[2]; int info;
            spa_localance(g2, Rn, h, n, min_maf, res, &info, iters);
            free(g2);
            if (info & 2) rt |= 4;
            double pw; int wrc = wald_localance(trait, g, h, E, Y, Cova, p, CH, L, n, &pw);
            if (wrc) rt |= 8;
            double pc = ORC_NA; int warn;
            int crc = wrc ? 1 : orc_cct2(res[0], pw, &pc, &warn);
            if (crc) { rt |= 16; rc = crc; pc = ORC_NA; }
            po[0] = res[0]; po[1] = pw; po[2] = pc; po[3] = res[1]; po[4] = pnorm1;
        }
    }
    free(Rn);
    out[0] = MAF; out[1] = miss; for (int k = 0; k < 5; k++) out[2 + k] = po[k];
    out[7] = S1; out[8] = VarS1; out[9] = Z1;
    if (route) *route = rt;
    free(g);
    return rc;
}
/* loop of SPAGxEmixCCT_localance :1443-1476 over N x M column-major matrices; CHs: L matrices, each N x M */
int orc_localance_matrix(int trait, const double *G, const double *H, long n, long m, const double *R, const double *E,
                         const double *Y, const double *Cova, int p, const double *const *CHs, int L, double epsilon, double min_maf,
                         double *out, int *route, int *iters) {
    int rc_all = 0;
#pragma omp parallel
    {
        double *ch = (double *)malloc(sizeof(double) * n * (L > 0 ? L : 1));
#pragma omp for schedule(dynamic, 1)
        for (long j = 0; j < m; j++) {
            double o[10]; int rt, it[2];
            for (int l = 0; l < L; l++) memcpy(ch + (long)l * n, CHs[l] + j * n, sizeof(double) * n);
            int rc = orc_localance_one_snp(trait, G + j * n, H + j * n, R, E, Y, Cova, p, ch, L, n, epsilon, min_maf, o, &rt, it);
            for (int c = 0; c < 10; c++) out[j + c * m] = o[c];
            if (route) route[j] = rt;
            if (iters) { iters[2 * j] = it[0]; iters[2 * j + 1] = it[1]; }
            if (rc) {
#pragma omp critical
                rc_all = rc;
            }
        }
        free(ch);
    }
    return rc_all;
}

/* ------------------------------------------------------------------ */
/* sparse-GRM quadratic form 2*sum_{nnz} v x_i x_j - sum x^2
 * (R/SPAGxEPlus_functions_MYZ.R:352-361, :509); i,j 0-based indices.    */
double orc_grm_quad(long nnz, const int32_t *gi, const int32_t *gj, const double *gv, const double *x, long n) {
    ld s = 0; for (long e = 0; e < nnz; e++) s += (gv[e] * x[gi[e]]) * x[gj[e]];
    ld s2 = 0; for (long i = 0; i < n; i++) s2 += x[i] * x[i];
    return (double)s * 2 - (double)s2;
}

/* the three Step-1 GRM scalars + R.new: R/SPAGxEPlus_functions_MYZ.R:490-537 */
void orc_plus_nullmodel_scalars(long nnz, const int32_t *gi, const int32_t *gj, const double *gv,
                                const double *R, const double *E, long n,
                                double *R_GRM_R, double *R_GRM_RE, double *Rnew, double *Rnew_GRM_Rnew) {
    double *RE = (double *)malloc(sizeof(double) * n);
    for (long i = 0; i < n; i++) RE[i] = R[i] * E[i];
    *R_GRM_R = orc_grm_quad(nnz, gi, gj, gv, R, n);
    ld p1 = 0, p2 = 0, s3 = 0;
    for (long e = 0; e < nnz; e++) { p1 += (gv[e] * R[gi[e]]) * RE[gj[e]]; p2 += (gv[e] * R[gj[e]]) * RE[gi[e]]; }
    for (long i = 0; i < n; i++) s3 += RE[i] * R[i];
    *R_GRM_RE = (double)p1 + (double)p2 - (double)s3;
    double lambda = *R_GRM_RE / *R_GRM_R;
    for (long i = 0; i < n; i++) Rnew[i] = (E[i] - lambda) * R[i];
    *Rnew_GRM_Rnew = orc_grm_quad(nnz, gi, gj, gv, Rnew, n);
    free(RE);
}

/* SPAGxE_Plus_one_SNP: R/SPAGxEPlus_functions_MYZ.R:248-405.  out[8].     */
int orc_plus_one_snp(const double *g_in, int g_is_int, const double *R, const double *E, const double *Rnew,
                     double R_GRM_R, double R_GRM_RE, double Rnew_GRM_Rnew,
                     long nnz, const int32_t *gi, const int32_t *gj, const double *gv, long n,
                     double epsilon, double cutoff, double min_maf, double out[8], int *route, int iters[2]) {
    double *g = (double *)malloc(sizeof(double) * n);
    memcpy(g, g_in, sizeof(double) * n);
    double MAF, miss; int rt = 0, touched = 0;
    if (iters) iters[0] = iters[1] = 0;
    if (prologue(g, n, g_is_int, min_maf, &MAF, &miss, &touched, g_gmodel)) {
        out[0] = MAF; out[1] = miss; for (int j = 2; j < 8; j++) out[j] = ORC_NA;
        if (route) *route = 1; free(g); return 0;
    }
    double gvar = 2 * MAF * (1 - MAF);
    ld acc = 0; for (long i = 0; i < n; i++) acc += g[i] * R[i];
    double S1 = (double)acc;
    double VarS1 = gvar * R_GRM_R;
    double Z1 = (S1 - 0) / sqrt(VarS1);
    double pnorm1 = orc_pnorm(-fabs(Z1), 1) * 2;
    double po[3];
    if (pnorm1 > epsilon) {
        acc = 0; for (long i = 0; i < n; i++) acc += (g[i] * E[i]) * R[i];
        double S2 = (double)acc;
        double lambda = R_GRM_RE / R_GRM_R;
        double S_GxE = S2 - lambda * S1;
        double Svar = gvar * Rnew_GRM_Rnew;
        double z = S_GxE / sqrt(Svar);
        if (fabs(z) < cutoff) {
            double pn = orc_pnorm(fabs(z), 0) * 2;
            po[0] = pn; po[1] = pn; po[2] = pnorm1;
        } else {
            rt |= 4;
            acc = 0; for (long i = 0; i < n; i++) acc += (Rnew[i] * Rnew[i]) * gvar;
            double ratio = (double)acc / Svar;
            double Snew = S_GxE * sqrt(ratio);
            int i1, i2;
            double pval1 = orc_getprob_spa(Rnew, NULL, MAF, n, fabs(Snew), 0, &i1);
            double pval2 = orc_getprob_spa(Rnew, NULL, MAF, n, -fabs(Snew), 1, &i2);
            if (iters) { iters[0] = i1; iters[1] = i2; }
            if (isnan(pval2)) pval2 = pval1;
            po[0] = pval1 + pval2; po[1] = orc_pnorm(fabs(z), 0) + orc_pnorm(-fabs(z), 1); po[2] = pnorm1;
        }
    } else {
        rt |= 2;
        double *R0 = (double *)malloc(sizeof(double) * n);
        if (adjust_for_g(g, R, n, R0)) {
            rt |= 32; po[0] = po[1] = ORC_NA; po[2] = pnorm1;
        } else {
            acc = 0; for (long i = 0; i < n; i++) acc += (g[i] * E[i]) * R0[i];
            double S0 = (double)acc;
            for (long i = 0; i < n; i++) R0[i] = E[i] * R0[i];
            ld cs = 0; for (long e = 0; e < nnz; e++) cs += (gv[e] * R0[gi[e]]) * R0[gj[e]];
            ld sx = 0, sxx = 0; for (long i = 0; i < n; i++) { sx += R0[i]; sxx += R0[i] * R0[i]; }
            double Smean = 2 * (double)sx * MAF;
            double Svar = 2 * (double)cs * gvar - (double)sxx * gvar;
            double z = (S0 - Smean) / sqrt(Svar);
            if (fabs(z) < cutoff) {
                double pn = orc_pnorm(fabs(z), 0) * 2;
                po[0] = pn; po[1] = pn;
            } else {
                rt |= 4;
                double var_spa = (double)sxx * gvar;
                double ratio = var_spa / Svar;
                double Snew = S0 * sqrt(ratio), Smean_new = Smean * sqrt(ratio);
                int i1, i2;
                double pval1 = orc_getprob_spa(R0, NULL, MAF, n, Smean + fabs(Snew - Smean_new), 0, &i1);
                double pval2 = orc_getprob_spa(R0, NULL, MAF, n, Smean - fabs(Snew - Smean_new), 1, &i2);
                if (iters) { iters[0] = i1; iters[1] = i2; }
                if (isnan(pval2)) pval2 = pval1;
                po[0] = pval1 + pval2; po[1] = orc_pnorm(fabs(z), 0) + orc_pnorm(-fabs(z), 1);
            }
            po[2] = pnorm1;
        }
        free(R0);
    }
    out[0] = MAF; out[1] = miss; out[2] = po[0]; out[3] = po[1]; out[4] = po[2];
    out[5] = S1; out[6] = VarS1; out[7] = Z1;
    if (route) *route = rt;
    free(g);
    return 0;
}

int orc_plus_bed(const uint8_t *bed, long pitch, long n, long mm, const double *R, const double *E, const double *Rnew,
                 double R_GRM_R, double R_GRM_RE, double Rnew_GRM_Rnew,
                 long nnz, const int32_t *gi, const int32_t *gj, const double *gv,
                 double epsilon, double cutoff, double min_maf, double *out, int *route, int *iters) {
#pragma omp parallel
    {
        double *g = (double *)malloc(sizeof(double) * n);
#pragma omp for schedule(dynamic, 1)
        for (long j = 0; j < mm; j++) {
            double o[8]; int rt, it[2];
            orc_bed_decode_row(bed + j * pitch, n, g);
            orc_plus_one_snp(g, 0, R, E, Rnew, R_GRM_R, R_GRM_RE, Rnew_GRM_Rnew, nnz, gi, gj, gv, n,
                             epsilon, cutoff, min_maf, o, &rt, it);
            for (int c = 0; c < 8; c++) out[j + c * mm] = o[c];
            if (route) route[j] = rt;
            if (iters) { iters[2 * j] = it[0]; iters[2 * j + 1] = it[1]; }
        }
        free(g);
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* SPAGxEmix_CCT_Plus_one_SNP: R/SPAmixGxECCTPlus_functions_MYZ_2024-09-17.R:144-427 (SPAGxEmix_CCT with the sparse GRM
 * in every variance).  out[12] as orc_mix_one_snp.  The reference reads the residual vector `R` as a free variable
 * (:235 ff.; its scripts define R = ResidMat$Resid); gi / gj are the 0-based positions of sparseGRM$ID1 / $ID2 in
 * ResidMat$SubjID (the match() calls of :228-229).  Branch B's Wald test is a mixed model fitted by lme4::glmer / lmer
 * (:384, :404), which is not restated: out[3] (Wald) and out[4] (CCT) are NA for branch-B SNPs and the caller combines
 * them (the R wrapper calls glmer / lmer itself for those few SNPs).  route bits as orc_mix_one_snp.               */
/* GetProb_SPA_G_new_adjust: :435-462 */
static double getprob_spa_adjust(const double *R, const double *mafvec, long n, double sstat, double var_ratio,
                                 int lower_tail, int *iter_out) {
    double zeta, h2; int it, cv;
    orc_spa_root(R, mafvec, 0, n, sstat, 0.0, &zeta, &it, &cv, &h2);
    if (iter_out) *iter_out = it;
    double k1 = H_org(zeta, R, mafvec, 0, n);
    double k2 = H2(zeta, R, mafvec, 0, n);
    double temp1 = zeta * sstat - k1;
    double w = r_sign(zeta) * sqrt(2 * temp1);
    double v = zeta * sqrt(k2);
    double a = w + 1 / w * log(v / w);
    double pval = orc_pnorm(a * sqrt(var_ratio), lower_tail);
    if (isnan(pval)) pval = 0;
    return pval;
}
/* sum(Value * x[ID1] * y[ID2]) over the GRM table (:230-232, :247-255) */
static double grm_cross(long nnz, const int32_t *gi, const int32_t *gj, const double *gv, const double *x, const double *y) {
    ld s = 0; for (long e = 0; e < nnz; e++) s += (gv[e] * x[gi[e]]) * y[gj[e]];
    return (double)s;
}
/* the statistic block shared by both branches (:268-299, :318-349): z, then normal or adjusted saddlepoint p-values */
static void mixplus_tails(const double *Rv, const double *m, const double *gvar, const double *sq, long n, long nnz,
                          const int32_t *gi, const int32_t *gj, const double *gv, double S, double cutoff,
                          double *pspa, double *pn, int *rt, int iters[2]) {
    double *x = (double *)malloc(sizeof(double) * n);
    for (long i = 0; i < n; i++) x[i] = Rv[i] * sq[i];
    ld acc = 0; for (long i = 0; i < n; i++) acc += Rv[i] * m[i];
    double Smean = 2 * (double)acc;
    acc = 0; for (long i = 0; i < n; i++) acc += (Rv[i] * Rv[i]) * gvar[i];
    double Svar_mix = (double)acc;
    double Svar = grm_cross(nnz, gi, gj, gv, x, x) * 2 - Svar_mix;
    double z = (S - Smean) / sqrt(Svar);
    if (fabs(z) < cutoff) {
        *pn = orc_pnorm(fabs(z), 0) * 2; *pspa = *pn;
    } else {
        *rt |= 4;
        double ratio = Svar_mix / Svar;
        int i1 = 0, i2 = 0;
        double pval1 = getprob_spa_adjust(Rv, m, n, fmax(S, 2 * Smean - S), ratio, 0, &i1);
        double pval2 = getprob_spa_adjust(Rv, m, n, fmin(S, 2 * Smean - S), ratio, 1, &i2);
        if (iters) { iters[0] = i1; iters[1] = i2; }
        if (isnan(pval2)) pval2 = pval1;
        *pspa = pval1 + pval2;
        *pn = orc_pnorm(fabs(z), 0) + orc_pnorm(-fabs(z), 1);
    }
    free(x);
}

int orc_mixplus_one_snp(const double *g_in, int g_is_int, const double *R, const double *E, const double *PCs, int k,
                        long n, long nnz, const int32_t *gi, const int32_t *gj, const double *gv,
                        double epsilon, double cutoff, double min_maf, double pc_pcut, double neg_ratio_cut,
                        double out[12], int *route, int iters[2]) {
    double *g = (double *)malloc(sizeof(double) * n);
    memcpy(g, g_in, sizeof(double) * n);
    double MAF, miss; int rt = 0, touched = 0;
    if (iters) iters[0] = iters[1] = 0;
    if (prologue(g, n, g_is_int, min_maf, &MAF, &miss, &touched, g_gmodel)) {
        out[0] = MAF; out[1] = miss; for (int j = 2; j < 12; j++) out[j] = ORC_NA;
        if (route) *route = 1; free(g); return 0;
    }
    double MAC = MAF * 2 * (double)n;
    double *m = (double *)malloc(sizeof(double) * n), *gvar = (double *)malloc(sizeof(double) * n);
    double *sq = (double *)malloc(sizeof(double) * n), *x = (double *)malloc(sizeof(double) * n);
    double negnum = 0;
    rt |= mix_af_estimate(g, n, MAF, MAC, PCs, k, pc_pcut, neg_ratio_cut, m, &negnum);
    for (long i = 0; i < n; i++) { gvar[i] = 2 * m[i] * (1 - m[i]); sq[i] = sqrt(gvar[i]); x[i] = R[i] * sq[i]; }   /* :222-224 */
    ld acc = 0; for (long i = 0; i < n; i++) acc += g[i] * R[i];
    double S1 = (double)acc;
    acc = 0; for (long i = 0; i < n; i++) acc += R[i] * m[i];
    double meanS1 = 2 * (double)acc;
    acc = 0; for (long i = 0; i < n; i++) acc += (R[i] * R[i]) * gvar[i];
    double VarS1 = grm_cross(nnz, gi, gj, gv, x, x) * 2 - (double)acc;                                             /* :260 */
    double Z1 = (S1 - meanS1) / sqrt(VarS1);
    double pnorm1 = orc_pnorm(-fabs(Z1), 1) * 2;
    double po[5];
    if (pnorm1 > epsilon) {
        acc = 0; for (long i = 0; i < n; i++) acc += (g[i] * E[i]) * R[i];
        double S2 = (double)acc;
        double *y = (double *)malloc(sizeof(double) * n);
        for (long i = 0; i < n; i++) y[i] = (R[i] * E[i]) * sq[i];                                                 /* :269 */
        acc = 0; for (long i = 0; i < n; i++) acc += (gvar[i] * E[i]) * (R[i] * R[i]);
        double Cov = grm_cross(nnz, gi, gj, gv, x, y) + grm_cross(nnz, gi, gj, gv, y, x) - (double)acc;            /* :279-281 */
        double lambda = Cov / VarS1;
        double S_GxE = S2 - lambda * S1;
        for (long i = 0; i < n; i++) y[i] = (E[i] - lambda) * R[i];
        double pspa, pn;
        mixplus_tails(y, m, gvar, sq, n, nnz, gi, gj, gv, S_GxE, cutoff, &pspa, &pn, &rt, iters);
        po[0] = po[1] = po[2] = pspa; po[3] = pn; po[4] = pnorm1;
        free(y);
    } else {
        rt |= 2;
        double *R0 = (double *)malloc(sizeof(double) * n);
        if (adjust_for_g(g, R, n, R0)) {
            rt |= 32; for (int j = 0; j < 4; j++) po[j] = ORC_NA; po[4] = pnorm1;
        } else {
            acc = 0; for (long i = 0; i < n; i++) acc += (g[i] * E[i]) * R0[i];
            double S0 = (double)acc;
            for (long i = 0; i < n; i++) R0[i] = E[i] * R0[i];
            double pspa, pn;
            mixplus_tails(R0, m, gvar, sq, n, nnz, gi, gj, gv, S0, cutoff, &pspa, &pn, &rt, iters);
            po[0] = pspa; po[1] = ORC_NA; po[2] = ORC_NA; po[3] = pn; po[4] = pnorm1;   /* Wald (glmer / lmer) and CCT: caller */
        }
        free(R0);
    }
    out[0] = MAF; out[1] = miss; for (int j = 0; j < 5; j++) out[2 + j] = po[j];
    out[7] = S1; out[8] = VarS1; out[9] = Z1; out[10] = negnum; out[11] = MAC;
    if (route) *route = rt;
    free(g); free(m); free(gvar); free(sq); free(x);
    return 0;
}

int orc_mixplus_bed(const uint8_t *bed, long pitch, long n, long mm, const double *R, const double *E, const double *PCs,
                    int k, long nnz, const int32_t *gi, const int32_t *gj, const double *gv, double epsilon, double cutoff,
                    double min_maf, double pc_pcut, double neg_ratio_cut, double *out, int *route, int *iters) {
    double *g = (double *)malloc(sizeof(double) * n);
    for (long j = 0; j < mm; j++) {
        double o[12]; int rt, it[2];
        orc_bed_decode_row(bed + j * pitch, n, g);
        orc_mixplus_one_snp(g, 0, R, E, PCs, k, n, nnz, gi, gj, gv, epsilon, cutoff, min_maf, pc_pcut, neg_ratio_cut, o, &rt, it);
        for (int c = 0; c < 12; c++) out[j + c * mm] = o[c];
        if (route) route[j] = rt;
        if (iters) { iters[2 * j] = it[0]; iters[2 * j + 1] = it[1]; }
    }
    free(g);
    return 0;
}
