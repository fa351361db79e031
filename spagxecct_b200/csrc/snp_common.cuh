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

}  // namespace spagxe
