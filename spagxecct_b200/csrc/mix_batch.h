// mix_batch.h -- host interface of the batched SPAGxEmix_CCT common path (mix_batch.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "types.cuh"

namespace spagxe {

struct MixBatchParams { double epsilon, cutoff, neg_ratio_cut; int g_model; };

// device scratch of the batched path (per-SNP flags, partial sums, coefficients); owned by the calling context
struct MixScratch;
MixScratch *mix_scratch_create();
void mix_scratch_destroy(MixScratch *ms);
// what the last mix_batch_run left for the per-SNP kernel: OLS coefficients (row j, leading dimension *ld) and the
// statistics of the SNPs whose work-list entry carries kMixHandFlag
const double *mix_scratch_beta(const MixScratch *ms, int *ld);
const MixHand *mix_scratch_hand(const MixScratch *ms);

// Consumes the candidate list k_finalize_mix left in (b_list, ctr->b_count): writes the output rows of the SNPs that
// take the common route and leaves the others in (b_list, ctr->b_count) for k_mix_snp.  All work is queued on `s`.
cudaError_t mix_batch_run(cudaStream_t s, MixScratch *scratch, int num_sms, const uint8_t *d_rows, int64_t pitch, SnpStats st, int64_t j0, int mc,
                          int64_t M, int64_t n, const double *R, const double *E, const VecConst *vc, const double *pcs,
                          const double *xtx_inv, int k, const MixBatchParams &prm, double *d_out, int *b_list, Counters *ctr,
                          int64_t *launches);

}  // namespace spagxe
