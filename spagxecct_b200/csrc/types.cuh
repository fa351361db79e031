// types.cuh -- plain structs shared by the score-pass translation unit (score_launch.cu) and the per-SNP test
// kernels (kernels.cuh / spagxe_cuda.cu).
#pragma once
#include <cstdint>

namespace spagxe {

// PLINK code -> A1 dosage: 00 -> 2, 01 -> missing, 10 -> 1, 11 -> 0 (pinned by inst/GenoMat_*.raw)
constexpr int kMaxQ = 16;   // Wald design columns: 1 + p + 3, p <= 12
constexpr int kMaxPC = 15;

struct VecConst {
    double sumR, sumR2, sumER2, lambda;  // lambda = sum(E R^2)/sum(R^2) (ref :592)
    double sumV, sumV2;                  // second score vector V (R.new for CCT, E*R for mix/plus)
    double sumRn, sumRn2;                // R.new = (E - lambda) R (ref :594)
    int nonfinite, pad_;                 // number of non-finite entries of R and E (set_vectors refuses them)
};

struct SnpStats {  // per-SNP outputs of the score pass (chunk-local index)
    double *D0, *D1;  // sum over non-missing of dosage * R, dosage * V
    double *T0, *T1;  // sum over missing samples of R, V
    int *asum;        // allele sum over non-missing = n_het + 2 n_homA1
    int *nmiss;       // number of missing calls
};

struct SpaItem { int snp; int flags; double maf; double S; double Smean; double aux; };
// deferred branch-B work item: everything the cluster kernel needs, valid until the next flush
struct BItem { const uint8_t *row; long long snp; int asum, nmiss; double D0, T0; };

struct Counters {  // device-side diagnostics / work-list cursors
    unsigned long long n_filtered, n_branch_a, n_branch_b, n_spa, n_wald_fail, n_cct_stop, n_singular,
        n_mix_logistic, newton_iters, n_allmissing, n_nonint;
    int spa_count, spax_count, b_count, pad_;
};

struct BmatInfo { int sR, sV; double maxR, maxV; };   // fixed-point scales of the tensor-core B matrix

}  // namespace spagxe
