// snp_common.cuh -- small per-SNP device helpers shared by the translation units (kernels.cuh, mix_batch.cu):
// the genotype prologue of every *_one_SNP function (ref R/SPAGxECCT.R:557-577) and warp-aggregated list appends.
#pragma once
#include <cstdint>

namespace spagxe {

// ---------------------------------------------------------------------------
// warp-aggregated list append
__device__ __forceinline__ int list_reserve(int *count, bool want) {
    unsigned mask = __ballot_sync(0xffffffffu, want);
    if (!mask) return -1;
    int leader = __ffs(mask) - 1, lane = threadIdx.x & 31;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return want ? base + __popc(mask & ((1u << lane) - 1)) : -1;
}

// genotype prologue from the integer counts (ref :557-577). Returns false if the row is all-missing.
struct Prologue {
    double maf_raw, maf, miss_rate, u, sum_g;
    bool flip;
};
__device__ __forceinline__ Prologue make_prologue(int asum, int nmiss, int64_t n) {
    Prologue p;
    const double nobs = (double)(n - nmiss);
    p.maf_raw = ((double)asum / nobs) / 2;  // contract: one correctly rounded division
    p.miss_rate = (double)nmiss / (double)n;
    p.u = 2 * p.maf_raw;
    p.sum_g = (double)asum + p.u * (double)nmiss;
    p.flip = p.maf_raw > 0.5;
    p.maf = p.flip ? 1 - p.maf_raw : p.maf_raw;
    if (p.flip) p.sum_g = 2.0 * (double)n - p.sum_g;
    return p;
}

// G.model (ref :572-574, dup :1045-1047, R/SPAGxEPlus_functions_MYZ.R:285-287): value of g per PLINK code
// (00 hom A1, 01 missing, 10 het, 11 hom A2) after mean imputation, the MAF > 0.5 flip and the Dom / Rec recode
// `ifelse(g>=1,1,0)` / `ifelse(g<=1,0,1)`.  MAF itself stays the additive one.  g_model: 0 Add, 1 Dom, 2 Rec.
struct GClass { double val[4]; };
__device__ __forceinline__ GClass make_gclass(const Prologue &pr, int g_model) {
    GClass gc;
    gc.val[0] = 2.0; gc.val[1] = pr.u; gc.val[2] = 1.0; gc.val[3] = 0.0;
    if (pr.flip) for (int c = 0; c < 4; c++) gc.val[c] = 2 - gc.val[c];
    if (g_model == 1) for (int c = 0; c < 4; c++) gc.val[c] = (gc.val[c] >= 1) ? 1.0 : 0.0;
    if (g_model == 2) for (int c = 0; c < 4; c++) gc.val[c] = (gc.val[c] <= 1) ? 0.0 : 1.0;
    return gc;
}
// class counts from the score pass: asum = n10 + 2 n00, nhom = n00
struct GCounts { double n[4]; };
__device__ __forceinline__ GCounts make_gcounts(int asum, int nmiss, int nhom, int64_t n) {
    GCounts c;
    c.n[0] = (double)nhom; c.n[1] = (double)nmiss; c.n[2] = (double)(asum - 2 * nhom);
    c.n[3] = (double)(n - nhom - nmiss - (asum - 2 * nhom));
    return c;
}
// sum(g' x) from the class sums of x: D = sum over non-missing of dosage * x = 2 T00 + T10, T = T01, H = T00
__device__ __forceinline__ double class_dot(const GClass &gc, double D, double T, double H, double sum_x) {
    const double t10 = D - 2 * H, t11 = ((sum_x - H) - T) - t10;
    return ((gc.val[0] * H + gc.val[1] * T) + gc.val[2] * t10) + gc.val[3] * t11;
}
__device__ __forceinline__ double class_sum(const GClass &gc, const GCounts &c, int power) {
    double s = 0;
    for (int k = 0; k < 4; k++) s += c.n[k] * (power == 2 ? gc.val[k] * gc.val[k] : gc.val[k]);
    return s;
}

}  // namespace spagxe
