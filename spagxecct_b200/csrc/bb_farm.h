// bb_farm.h -- host interface of the SPAGxE_CCT branch-B solver farm (bb_farm.cu): K5 of DESIGN.md.
//
// Branch B of SPAGxE_CCT_one_SNP (ref R/SPAGxECCT.R:604-678) runs for the SNPs whose marginal genetic effect is
// significant: genotype-adjusted residual R.new0 = E (R - W (W'W)^-1 W'R), the homogeneous SPA test on it, a per-SNP
// Wald test (glm.fit IRLS for a binary trait, lm for a quantitative one) and the Cauchy combination of the two.
// All of it is iteration-synchronous work over the N samples (Newton steps on the CGF, IRLS passes), so the farm runs
// every work item of a flush in lock step on the whole GPU: one persistent cooperative kernel, one CTA per SM, each
// CTA owning a fixed contiguous slice of the samples whose SNP-invariant columns (R, E, Y, covariates) are staged in
// shared memory once; per iteration every CTA sweeps its slice for every active solver, partial sums meet in global
// memory at a grid barrier and every CTA redoes the (tiny) scalar update in the same fixed order.  The number of
// items is read on the device, so the host never synchronises to size the launch.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "types.cuh"

namespace spagxe {

struct BbFarmArgs {
    const BItem *items;        // work list written by k_finalize_cct
    const int *count;          // its length (device memory)
    int64_t m_total, n;
    const double *R, *E;
    const VecConst *vc;
    const double *Y, *cova;    // Wald data: Phen.mtx$Y and Cova.mtx (n x p column-major)
    int p, trait;              // trait: 1 binary (glm.fit), 2 quantitative (lm)
    double min_maf;            // forwarded to the inner SPA call (ref :616)
    int g_model;               // 0 Add, 1 Dom, 2 Rec (ref :572-574): the item's H0 separates het and hom sums
    double *out;               // M x 10 column-major result matrix (device)
    Counters *ctr;
};

struct BbFarm;   // device scratch + launch geometry, owned by the context

// variant: 0 = product (16 warps, 2 samples per lane in flight); others are tuning variants with identical results
// up to summation order (see bb_farm.cu)
cudaError_t bb_farm_create(BbFarm **out, int device, int num_sms);
void bb_farm_destroy(BbFarm *f);
// enqueue the farm for the items currently in args.items[0 .. *args.count); resets nothing (the caller zeroes *count)
cudaError_t bb_farm_launch(BbFarm *f, cudaStream_t s, const BbFarmArgs &args, int64_t *launches, const char **what);
// flop model of the last launch geometry (for the bench's fp64 roofline): sample-passes are counted on the device
struct BbFarmWork { unsigned long long irls_sample_passes, tail_sample_passes, prep_sample_passes, iterations, items; };
cudaError_t bb_farm_read_work(BbFarm *f, cudaStream_t s, BbFarmWork *w, bool reset);

}  // namespace spagxe
