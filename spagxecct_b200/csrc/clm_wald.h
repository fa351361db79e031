// clm_wald.h -- host interface of the categorical-trait Wald test (clm_wald.cu): ordinal::clm(as.factor(Y) ~ Cova + E + g + g*E)
// for the branch-B items of SPAGxE_CCT (ref R/SPAGxECCT.R:664-677).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "types.cuh"

namespace spagxe {

constexpr int kClmMaxCov = 2;   // covariate columns of the register-resident design (E, g, E:g + 2)
constexpr int kClmMaxJ = 8;     // response categories

struct ClmData;   // SNP-invariant device data: category order, gathered design columns
ClmData *clm_create();
void clm_destroy(ClmData *c);
void clm_invalidate(ClmData *c);           // E, Y or the covariates changed
bool clm_ready(const ClmData *c);
// h_ycat: host vector of length n, the index of every sample's level in sort(unique(Y)) (0 .. J-1, every level present);
// d_E, d_cova: device (n, n x p column-major)
cudaError_t clm_setup(ClmData *c, cudaStream_t s, int64_t n, const double *h_ycat, const double *d_E, const double *d_cova, int p,
                      const char **what);
// Wald p-value and CCT(p.spa, p.wald) for items[0 .. *count): reads p.spa from column 2 of `out` (written by the solver
// farm earlier on the same stream), writes columns 3 and 4
cudaError_t clm_launch(ClmData *c, cudaStream_t s, int num_sms, const BItem *items, const int *count, int64_t m_total, int g_model,
                       double *out, Counters *ctr, int64_t *launches, bool na_to_spa = false);
// na_to_spa: SPAGxEmix_CCT's rule `if (is.na(pval.wald)) pval.wald = pval.spaGxE` (ref R/SPAGxECCT.R:1190, :1249); SPAGxE_CCT has none

}  // namespace spagxe
