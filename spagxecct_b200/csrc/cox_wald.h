// cox_wald.h -- host interface of the survival-trait Wald test (cox_wald.cu): coxph(Surv(time, event) ~ Cova + E + g + g*E)
// for the branch-B items of SPAGxE_CCT (ref R/SPAGxECCT.R:622-634).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "types.cuh"

namespace spagxe {

constexpr int kCoxMaxCov = 2;   // covariate columns of the register-resident design (E, g, E:g + 2)

struct CoxData;   // SNP-invariant device data: time order, gathered and centred columns, tie-group markers
CoxData *cox_create();
void cox_destroy(CoxData *c);
void cox_invalidate(CoxData *c);           // E or the covariates changed
bool cox_ready(const CoxData *c);
// h_time / h_event: host vectors of length n (Phen.mtx$surv.time, Phen.mtx$event); d_E, d_cova: device (n, n x p column-major)
cudaError_t cox_setup(CoxData *c, cudaStream_t s, int64_t n, const double *h_time, const double *h_event, const double *d_E,
                      const double *d_cova, int p, const char **what);
// Wald p-value and CCT(p.spa, p.wald) for items[0 .. *count): reads p.spa from column 2 of `out` (written by the solver
// farm earlier on the same stream), writes columns 3 and 4
cudaError_t cox_launch(CoxData *c, cudaStream_t s, int num_sms, const BItem *items, const int *count, int64_t m_total, int g_model,
                       double *out, Counters *ctr, int64_t *launches, bool na_to_spa = false);
// na_to_spa: SPAGxEmix_CCT's rule `if (is.na(pval.wald)) pval.wald = pval.spaGxE` (ref R/SPAGxECCT.R:1190, :1249); SPAGxE_CCT has none

}  // namespace spagxe
