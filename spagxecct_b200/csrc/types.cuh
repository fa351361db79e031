// types.cuh -- plain structs shared by the score-pass translation unit (score_launch.cu) and the per-SNP test
// kernels (kernels.cuh / spagxe_cuda.cu).
#pragma once
#include <cstdint>
#include <cstddef>

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
    // G.model Dom / Rec only (second score pass with a 0/1 LUT): sums over hom-A1 samples of R, V and their number
    double *H0, *H1;
    int *nhom;
};

struct SpaItem { int snp; int flags; double maf; double S; double Smean; double aux; };
// deferred branch-B work item: everything the cluster kernel needs, valid until the next flush
struct BItem { const uint8_t *row; long long snp; int asum, nmiss; double D0, T0; double H0; int nhom, pad_; };   // H0, nhom: G.model Dom / Rec

struct Counters {  // device-side diagnostics / work-list cursors
    unsigned long long n_filtered, n_branch_a, n_branch_b, n_spa, n_wald_fail, n_cct_stop, n_singular,
        n_mix_logistic, newton_iters, n_allmissing, n_nonint, cox_evals, clm_evals;
    int spa_count, spax_count, b_count, work_cursor;   // work_cursor: dynamic item queue of k_mix_snp (reset with the lists)
};

struct BmatInfo { int sR, sV; double maxR, maxV; };   // fixed-point scales of the tensor-core B matrix


// ---- launch geometry and kernel argument structs shared by the host pipeline (spagxe_cuda.cu) and the per-SNP test
// kernels (kernels.cuh, compiled in snp_kernels.cu) ---------------------------------------------------------------
struct WaldData {
    const double *Y, *cova;  // cova: n x p column-major
    int p, trait;
};

// 256 bins x 12 Chebyshev nodes.  Interpolation error of phi(t r) on a bin of width h: (|t| h / 2)^n |phi^(n)| / (2^(n-1) n!);
// the CGF pieces are analytic within |Im x| < pi, so |phi^(n)| ~ n! / pi^n and the error is ~ (|t| h / (2 pi))^12 / 2^11
// = 1.3e-13 at |t| h = 1 (round 1 used 2048 bins x 6 nodes: four times the nodes, 5e-13 at its limit |t| h = 0.1).
constexpr int kQNodes = 12;
constexpr int kQBins = 256;
constexpr double kQTheta = 1.0;
constexpr double kQScale = 274877906944.0;  // 2^38 fixed-point weights (exact integer atomics => deterministic)

struct Quad {
    double *x;          // [kQBins * kQNodes] nodes
    double *w;          // [kQBins * kQNodes] weights (double, after conversion)
    long long *wfix;    // fixed-point accumulation buffer
    double *range;      // [0]=rmin, [1]=rmax, [2]=h, [3]=max |t| covered (kQTheta / h; 0 = disabled)
    int *count;         // number of nodes with non-zero weight (x, w are compacted to that length)
};

constexpr int kSpaQThreads = 256;   // k_spa_quad
constexpr int kSpaThreads = 512;    // k_spa_homo
constexpr int kBThreads = 512;
constexpr int kBTile = 512;                 // samples staged per tile (one per thread)
constexpr int kBCluster = 8;                // CTAs cooperating on one SNP
constexpr int kBGroups = kBThreads / 32;    // entry groups (one warp each) in the accumulation phase
constexpr int kBMaxEnt = (kMaxQ * (kMaxQ + 3)) / 2;          // 152 normal-equation entries incl. rhs
constexpr int kBSlots = (kBMaxEnt + kBGroups - 1) / kBGroups; // <= 10 accumulators per thread
inline __host__ __device__ size_t branch_b_smem(int q) { return sizeof(double) * kBTile * (size_t)(((q + 1) | 1) + 1); }

struct PlusConst { double R_GRM_R, R_GRM_RE, Rnew_GRM_Rnew; };
struct GrmCoo { const int *gi, *gj; const double *gv; long long nnz; };

// hand-over of the batched SPAGxEmix path (mix_batch.cu) to the per-SNP kernel for a SNP whose only open item is the
// saddlepoint approximation: its statistics are final, k_mix_snp starts from the batch's OLS coefficients
struct MixHand { double S1, VarS1, Z1, pG, lambda, S_GxE, Smean, Svar, z, negnum, MAC; };
constexpr int kMixHandFlag = 1 << 28;   // set in a work-list entry whose MixHand record is valid

struct MixData {
    const double *pcs;      // n x k column-major (topPCs)
    const double *xtx_inv;  // (k+1) x (k+1) inverse of [1,PCs]'[1,PCs] (SNP-invariant)
    int k;
    double pc_pcut, neg_ratio_cut;
    int g_model;            // 0 Add, 1 Dom, 2 Rec
    BItem *wald_items;      // traits survival / categorical: branch-B SNPs are appended here (count: Counters::spax_count) and
                            // their Wald test + CCT run afterwards in k_cox_wald / k_clm_wald (ref R/SPAGxECCT.R:1187-1262)
};

// SPAGxEmixCCT_localance (ref R/SPAGxECCT.R:1401-1647): packed 2-bit rows (code -> value: 00 2, 10 1, 11 0; 01 = NA is refused)
// of the ancestry-specific genotypes, of the local ancestry counts and of the Cova.haplo.mtx.list columns of every SNP
constexpr int kMaxLocalCov = 4;
struct LocalArgs {
    const uint8_t *g_rows, *h_rows; const uint8_t *ch_rows[kMaxLocalCov]; int n_ch; int64_t pitch;
    int m; int64_t snp0, m_total, n;
    const double *R, *E; const VecConst *vc; WaldData wd;
    double epsilon, min_maf; int g_model;
    double *out; Counters *ctr;
};

// SPAGxE_CCT on a dense real-valued genotype matrix (imputed dosages; ref R/SPAGxECCT.R:543-683 with non-integer g)
struct MixDosageArgs {   // SPAGxEmix_CCT / SPAGxEmix_CCT_Plus on a dense column block (imputed dosages)
    const double *G; int64_t ld;       // device, column-major n x m block (NaN = NA)
    int m; int64_t snp0, m_total, n;
    const double *R, *E; const VecConst *vc; WaldData wd; MixData md; GrmCoo grm;
    double epsilon, cutoff, min_maf;
    double *out; Counters *ctr; double *mcache_all;
};
struct PlusDosageArgs {  // SPAGxE_Plus on a dense column block (imputed dosages)
    const double *G; int64_t ld; int m; int64_t snp0, m_total, n;
    const double *R, *E, *Rn; const VecConst *vc; PlusConst pc; GrmCoo grm;
    double epsilon, cutoff, min_maf; int g_model;
    double *out; Counters *ctr;
};
struct DosageArgs {
    const double *G; int64_t ld;       // device, column-major n x m block (NaN = NA)
    int m; int64_t snp0, m_total, n;
    const double *R, *E, *Rn; const VecConst *vc; WaldData wd;
    double epsilon, min_maf; int g_model;
    double *out; Counters *ctr;
};

}  // namespace spagxe
