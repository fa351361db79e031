/*
 * spagxe.h -- C ABI of libspagxe_cuda.so
 *
 * The drop-in boundary for SPAGxECCT's Step-2 per-SNP testing loop on NVIDIA
 * B200 (sm_100a).  The reference (YuzhuoMa97/SPAGxECCT) is a pure-R package with
 * no native interface, so each entry point below states which R closure (and which
 * lines of it) it replaces; the R-side `.Call` shim that binds them is shown in
 * INTEGRATION.md and spagxecct_b200/R/.
 *
 * Conventions
 *   - plain C types only; every function returns an int status (SPAGXE_OK == 0);
 *     the message of the last failure is available from spagxe_last_error().
 *   - no C++ exception and no CUDA error crosses this boundary.
 *   - matrices are column-major (R layout); outputs are M x ncol column-major.
 *   - filtered SNPs (MAF < min.maf) carry R's NA_real_ bit pattern
 *     (0x7FF00000000007A2) in the statistic columns, like `matrix(NA, M, k)`.
 *   - there is no CPU fallback: every entry point fails with
 *     SPAGXE_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SPAGXE_H
#define SPAGXE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spagxe_ctx spagxe_ctx;

enum {
    SPAGXE_OK = 0,
    SPAGXE_ERR_CUDA = 1,        /* CUDA runtime failure (message has the detail)          */
    SPAGXE_ERR_ARG = 2,         /* bad argument (NULL, size mismatch, unknown enum)       */
    SPAGXE_ERR_STATE = 3,       /* a required spagxe_set_* call is missing                */
    SPAGXE_ERR_UNSUPPORTED = 4, /* e.g. non-integer dosages with a survival / categorical trait    */
    SPAGXE_ERR_CCT_STOP = 5,    /* CCT() would stop(): NA / out-of-range p (ref :2072-2087)*/
    SPAGXE_ERR_SINGULAR = 6     /* solve(t(W)%*%W) singular (ref :609)                    */
};

/* traits: the reference's `traits=` argument (R/SPAGxECCT.R:455, :622-677) */
enum { SPAGXE_TRAIT_BINARY = 1, SPAGXE_TRAIT_QUANTITATIVE = 2, SPAGXE_TRAIT_SURVIVAL = 3, SPAGXE_TRAIT_CATEGORICAL = 4 };

/* genotype sources (R/SPAGxECCT.R:483-488: Geno.mtx or GenoFile) */
enum {
    SPAGXE_GENO_BED = 0, /* PLINK .bed rows, variant-major, 2 bits/sample, WITHOUT the 3-byte magic */
    SPAGXE_GENO_F64 = 1, /* R numeric matrix  N x M, NA/NaN = missing, values in {0,1,2}             */
    SPAGXE_GENO_I32 = 2  /* R integer matrix  N x M, INT_MIN (NA_integer_) = missing                  */
};
enum { SPAGXE_LOC_HOST = 0, SPAGXE_LOC_DEVICE = 1 };

typedef struct {
    int32_t kind;      /* SPAGXE_GENO_*                                                  */
    int32_t location;  /* SPAGXE_LOC_*: where `data` lives                               */
    const void *data;  /* first SNP's first byte/element                                 */
    int64_t n_samples; /* N                                                              */
    int64_t n_snps;    /* M                                                              */
    int64_t pitch;     /* BED: bytes between consecutive SNP rows (>= ceil(N/4));
                          F64/I32: elements between consecutive columns (>= N)           */
    /* what GRAB::GRAB.ReadGeno does between the file and the matrix (R/SPAGxECCT.R:484-487), BED only, optional: */
    const int64_t *sample_index; /* HOST array of n_samples entries or NULL: sample_index[i] = position inside a
                                    file row of analysis sample i (the SampleIDs subset / reorder); the gather runs
                                    on the device, 2 bits -> 2 bits                                             */
    int64_t n_file_samples;      /* samples per packed row of the source when sample_index is given (pitch >=
                                    ceil(n_file_samples/4)); 0 = n_samples                                      */
    int32_t flags;               /* SPAGXE_GENO_COUNT_A2: count the second allele (control$AlleleOrder="ref-first") */
    int32_t reserved;
} spagxe_geno;
enum { SPAGXE_GENO_COUNT_A2 = 1 };

/* the tuning arguments shared by SPAGxE_CCT / SPAGxEmix_CCT / SPAGxE_Plus
 * (R/SPAGxECCT.R:464-469, :926-933; R/SPAGxEPlus_functions_MYZ.R:168-173)               */
typedef struct {
    double epsilon;          /* 0.001 */
    double cutoff;           /* 2; NOT used by run_cct, exactly like the reference (:598,:616) */
    double min_maf;          /* 0.001 */
    double missing_cutoff;   /* 0.15; accepted and ignored like the reference (never applied)  */
    double pc_pvalue_cutoff; /* topPCs.pvalue.cutoff = 0.05 (mix only)                  */
    double neg_ratio_cutoff; /* MAF.est.negative.ratio.cutoff = 0.1 (mix only)          */
    int32_t g_model;         /* 0 = "Add", 1 = "Dom", 2 = "Rec" (R/SPAGxECCT.R:572-574)  */
    int32_t impute_method;   /* 0 = "fixed"                                             */
    int32_t out_location;    /* SPAGXE_LOC_HOST (default) or SPAGXE_LOC_DEVICE for `out` */
    int32_t reserved;
    int64_t out_ld;          /* rows of the caller's (host) result matrix when it is larger than this call's M: the
                                call writes a row block of it (0 = M; used by the multi-GPU driver)              */
} spagxe_params;

/* diagnostics filled by every run (all counters are totals over the call) */
typedef struct {
    int64_t n_snps;
    int64_t n_filtered;     /* MAF < min.maf                                             */
    int64_t n_branch_a;     /* p.value.betaG >  epsilon                                  */
    int64_t n_branch_b;     /* p.value.betaG <= epsilon (genotype-adjusted + Wald + CCT) */
    int64_t n_spa;          /* SNPs whose |z| >= cutoff went through the saddlepoint     */
    int64_t n_wald_fail;    /* rank-deficient / non-finite IRLS -> NA                    */
    int64_t n_cct_stop;     /* CCT() stop() conditions met                               */
    int64_t n_singular;     /* singular W'W                                              */
    int64_t n_mix_logistic; /* mix: SNPs that used the logistic AF fallback              */
    int64_t newton_iters;   /* total Newton iterations over all tails                    */
    int64_t kernel_launches;/* kernels launched by this call                             */
    int64_t h2d_bytes, d2h_bytes;
    double ms_total;        /* CUDA-event time of the whole call on the ctx stream       */
    double ms_score;        /* sum over chunks: unpack + score kernel(s)                 */
    double ms_tests;        /* sum over chunks: finalize + SPA + branch-B kernels        */
    int64_t score_bytes;    /* algorithmic bytes of the score pass: M * ceil(N/4)        */
    /* stage split of ms_tests and the work counted on the device (fp64 rooflines of the SPA and branch-B stages) */
    double ms_spa;          /* sum over chunks: per-SNP epilogue + SPA kernels           */
    double ms_branch_b;     /* deferred branch-B launches (solver farm / cluster kernel) */
    int64_t spa_nodes;      /* quadrature nodes per CGF evaluation (0: per-sample sums)  */
    int64_t bb_irls_sample_passes; /* samples x IRLS / lm passes swept by the farm       */
    int64_t bb_tail_sample_passes; /* samples x CGF evaluations of branch-B SPA tails    */
    int64_t bb_prep_sample_passes; /* samples x preparation passes (counts, R.new0 sums) */
    int64_t bb_iterations;  /* lock-step iterations (grid barriers) of the farm          */
    int64_t cox_evals;      /* traits = survival: partial-likelihood evaluations, all items */
    int64_t clm_evals;      /* traits = categorical: likelihood evaluations, all items   */
} spagxe_diag;

/* ---- lifetime ------------------------------------------------------------- */
/* device: CUDA ordinal. Replaces nothing in the reference (which has no state). */
int spagxe_ctx_create(int device, spagxe_ctx **out);
void spagxe_ctx_destroy(spagxe_ctx *ctx);
const char *spagxe_last_error(const spagxe_ctx *ctx); /* ctx may be NULL: last create error */
/* run all work on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = own stream */
int spagxe_ctx_set_stream(spagxe_ctx *ctx, void *cuda_stream);
const char *spagxe_version(void);
/* Test / tuning hooks -- NOT part of the reference-facing surface and never needed by a caller of the R functions.
 * Each option switches between implementations that must give the same results (the tests compare them); the
 * defaults are the product configuration.  The library reads no environment variable.                       */
enum {
    SPAGXE_OPT_CHUNK_SNPS = 1,  /* SNPs per pipeline chunk (0 = automatic, ~1 GiB of packed rows)             */
    SPAGXE_OPT_EXACT_SPA = 2,   /* 1: per-sample CGF sums instead of the SNP-invariant quadrature (K4 vs K4c) */
    SPAGXE_OPT_MIX_PER_SNP = 3, /* 1: every SPAGxEmix SNP through the per-SNP kernel (K6b) instead of K6a      */
    SPAGXE_OPT_MIX_CLUSTER = 5, /* CTAs per SNP in the per-SNP SPAGxEmix kernel (1..8; tuning, results agree to rounding) */
    SPAGXE_OPT_SCORE_IMPL = 4,  /* 0: tcgen05 score kernel (default), 1: fp64 CUDA-core cross-check kernel    */
    SPAGXE_OPT_PROBE_CFG = 100  /* builds with -DSPAGXE_PROBES only (tools/): score-kernel timing probes       */
};
int spagxe_ctx_set_option(spagxe_ctx *ctx, int option, int64_t value);

/* ---- per-analysis inputs --------------------------------------------------- */
/* R (Step-1 residuals from SPA_G_Get_Resid, R/SPAGxECCT.R:48-96) and E; host pointers,
 * length N.  Replaces the R/E arguments of SPAGxE_CCT (:460-461) / SPAGxEmix_CCT (:921-922). */
int spagxe_set_vectors(spagxe_ctx *ctx, int64_t n, const double *R, const double *E);
/* same with device pointers (multi-GPU: rank 0's copy arrives by one NCCL broadcast) */
int spagxe_set_vectors_device(spagxe_ctx *ctx, int64_t n, const double *dR, const double *dE);

/* data of the Wald model Y ~ Cova + E + g + E:g (R/SPAGxECCT.R:622-677): Phen.mtx$Y and
 * as.matrix(Cova.mtx) (N x p column-major, p <= 12).  trait = SPAGXE_TRAIT_*.
 * SPAGXE_TRAIT_CATEGORICAL (R/SPAGxECCT.R:664-677, :1244-1262: ordinal::clm(as.factor(Phen.mtx$Y) ~ ..., logit link,
 * flexible thresholds): Y holds the index of every sample's level in sort(unique(Y)), 0 .. J-1 with J <= 8 and every level
 * present; p <= 2; SPAGxE_CCT and SPAGxEmix_CCT.  Survival data go through spagxe_set_wald_survival.   */
int spagxe_set_wald(spagxe_ctx *ctx, int trait, const double *Y, const double *cova, int p);

/* same with device pointers (the multi-GPU driver hands over what arrived by the NCCL broadcast) */
int spagxe_set_wald_device(spagxe_ctx *ctx, int trait, const double *dY, const double *dcova, int p);

/* traits = "survival" (R/SPAGxECCT.R:622-634): the Wald model is coxph(Surv(Phen.mtx$surv.time, Phen.mtx$event) ~
 * as.matrix(Cova.mtx) + E + g + g*E, iter.max = 1000) with Efron ties; host pointers, event in {0, 1}, p <= 2.
 * SPAGxE_CCT and SPAGxEmix_CCT (:1187-1193; spagxe_run_cct / spagxe_run_mix); sets the trait to SPAGXE_TRAIT_SURVIVAL. */
int spagxe_set_wald_survival(spagxe_ctx *ctx, const double *surv_time, const double *event, const double *cova, int p);

/* topPCs of SPAGxEmix_CCT (R/SPAGxECCT.R:925, :1066): N x k column-major, k <= 15         */
int spagxe_set_pcs(spagxe_ctx *ctx, const double *pcs, int k);

/* obj.SPAGxE_Plus_Nullmodel (R/SPAGxEPlus_functions_MYZ.R:539-545): R via spagxe_set_vectors,
 * R.new (length N), the three GRM scalars and the sparse GRM as 0-based COO triplets
 * (lower triangle incl. diagonal, as in data/sparseGRM.SPAGxEPlus.rda).                   */
int spagxe_set_grm(spagxe_ctx *ctx, int64_t nnz, const int32_t *gi, const int32_t *gj, const double *gv,
                   double R_GRM_R, double R_GRM_RE, double Rnew_GRM_Rnew, const double *Rnew);

/* The three sparse-GRM quadratic forms and R.new of SPAGxE_Plus_Nullmodel (R/SPAGxEPlus_functions_MYZ.R:490-537:
 * the dplyr match()/mutate() block), computed on the device with the same COO kernel family as the Step-2 branch B.
 * All pointers are HOST arrays; gi/gj are 0-based sample indices (the R shim does the match() on SubjID once).
 * scalars3 = {R_GRM_R, R_GRM_RE, R.new_GRM_R.new}; Rnew has length n.  Does not change the context's analysis state. */
int spagxe_plus_nullmodel(spagxe_ctx *ctx, int64_t n, const double *R, const double *E, int64_t nnz, const int32_t *gi,
                          const int32_t *gj, const double *gv, double *scalars3, double *Rnew);

/* ---- the Step-2 loops ------------------------------------------------------ */
/* SPAGxE_CCT loop body, R/SPAGxECCT.R:509-525 -> SPAGxE_CCT_one_SNP :543-683.
 * out: M x 10 column-major: MAF, missing.rate, p.value.spaGxE, p.value.spaGxE.Wald,
 * p.value.spaGxE.CCT.Wald, p.value.normGxE, p.value.betaG, Stat.betaG, Var.betaG, z.betaG */
int spagxe_run_cct(spagxe_ctx *ctx, const spagxe_geno *geno, const spagxe_params *prm, double *out,
                   spagxe_diag *diag);
/* SPAGxEmix_CCT loop body, R/SPAGxECCT.R:975-997 -> :1013-1269.  out: M x 12 (adds
 * MAF.est.negative.num, MAC).                                                            */
int spagxe_run_mix(spagxe_ctx *ctx, const spagxe_geno *geno, const spagxe_params *prm, double *out,
                   spagxe_diag *diag);
/* SPAGxE_Plus loop body, R/SPAGxEPlus_functions_MYZ.R:214-233 -> :248-405.  out: M x 8:
 * MAF, missing.rate, p.value.spaGxE.plus, p.value.normGxE.plus, p.value.betaG, Stat.betaG,
 * Var.betaG, z.betaG                                                                     */
int spagxe_run_plus(spagxe_ctx *ctx, const spagxe_geno *geno, const spagxe_params *prm, double *out,
                    spagxe_diag *diag);

/* SPAGxEmix_CCT_Plus (R/SPAmixGxECCTPlus_functions_MYZ_2024-09-17.R:38-133 -> SPAGxEmix_CCT_Plus_one_SNP :144-427): SPAGxEmix_CCT
 * with the sparse GRM in every variance (t(R sqrt(g.var.est)) GRM (R sqrt(g.var.est)), :222-232, :260, :269-299, :318-349)
 * and GetProb_SPA_G_new_adjust (:435-462).  spagxe_set_vectors (R = ResidMat$Resid: the reference's one-SNP function
 * reads `R` as a free variable), spagxe_set_pcs and spagxe_set_sparse_grm first; no Wald data.
 * spagxe_set_sparse_grm: the sparseGRM table as 0-based COO triplets, gi / gj = positions of ID1 / ID2 in ResidMat$SubjID
 * after the reference's filter (:220; the match() calls of :228-229 are done once by the caller).
 * out: M x 12 like spagxe_run_mix (p.value.spaGxEmix.plus, ...Wald, ...CCT.Wald, p.value.normGxEmix.plus, ...).  For the
 * SNPs of branch B (p.value.betaG <= epsilon) the reference's Wald test is a mixed model fitted by lme4::glmer / lmer
 * (:384, :404); that fit is NOT done here: columns 4 and 5 of those rows hold NA_real_, column 3 holds pval.spaGxE, and
 * the caller (spagxecct_b200/R/SPAGxECCT_cuda.R; api.SPAGxEmix_CCT_Plus) fits the model for those few SNPs and applies
 * CCT.  diag->n_branch_b counts them.                                                                             */
int spagxe_set_sparse_grm(spagxe_ctx *ctx, int64_t nnz, const int32_t *gi, const int32_t *gj, const double *gv);
int spagxe_run_mix_plus(spagxe_ctx *ctx, const spagxe_geno *geno, const spagxe_params *prm, double *out,
                        spagxe_diag *diag);

/* SPAGxEmixCCT_localance loop body, R/SPAGxECCT.R:1443-1476 -> SPAGxEmixCCT_localance_one_SNP :1496-1647 (with
 * SPAmix_localance_one_SNP :2324-2395 and the binomial(h_i, MAF) saddlepoint functions :2132-2311): ancestry-specific
 * G x E test with local ancestry.  The reference takes R matrices only: G (Geno.mtx, ancestry-specific genotypes), H
 * (haplo.mtx, local ancestry counts) and the n_cov_haplo matrices of Cova.haplo.mtx.list, all HOST, double,
 * column-major n x m with leading dimension ld, values in {0, 1, 2} (an NA stops the reference at `if (MAF > 0.5)`).
 * spagxe_set_vectors and spagxe_set_wald (binary or quantitative) first.  Wald designs: binary glm(Y ~ Cova.haplo +
 * Cova + E + g + g*E), quantitative lm(Y ~ Cova.haplo + Cova + haplo_numVec + E + g + g*E); columns that lm / glm would
 * report as aliased (both ancestries' counts of a two-way admixture: h1 + h2 = 2) are dropped as there.  out: HOST
 * m x 10, column-major: MAF, missing.rate, p.value.spaGxE.index.ance, p.value.spaGxE.Wald.index.ance,
 * p.value.spaGxE.CCT.Wald.index.ance, p.value.normGxE.index.ance, p.value.betaG.index.ance, Stat / Var / z.betaG.      */
int spagxe_run_localance(spagxe_ctx *ctx, int64_t n, int64_t m, const double *G, const double *H, int n_cov_haplo,
                         const double *const *cov_haplo, int64_t ld, const spagxe_params *prm, double *out,
                         spagxe_diag *diag);

/* ---- the same loop on several GPUs of one node ----------------------------------
 * One process, one host thread + stream per GPU.  SNPs are independent, so the M SNPs of a call are cut into contiguous
 * ranges (a .bed file is variant-major: a range of SNPs is a range of bytes); the per-analysis vectors are uploaded to
 * the first GPU and reach the others through ONE ncclBroadcast of [R | E | Y | Cova]; every GPU copies its result rows
 * into its own row block of the caller's M x ncol matrix.  No other collective, no reduction across GPUs.  NCCL is
 * loaded with dlopen("libnccl.so.2") when the first multi-GPU context is created (single-GPU use never needs it).
 * Replaces the serial for-loop of SPAGxE_CCT (R/SPAGxECCT.R:509-525) exactly like spagxe_run_cct.               */
typedef struct spagxe_mgpu spagxe_mgpu;
int spagxe_mgpu_create(int n_gpus, const int *devices /* NULL: 0 .. n_gpus-1 */, spagxe_mgpu **out);
void spagxe_mgpu_destroy(spagxe_mgpu *mg);
const char *spagxe_mgpu_last_error(const spagxe_mgpu *mg);   /* mg may be NULL: last create error */
int spagxe_mgpu_n_gpus(const spagxe_mgpu *mg);
/* R, E (length n) and the Wald data (trait, Y, Cova n x p) of SPAGxE_CCT; host pointers.  One broadcast. */
int spagxe_mgpu_set_inputs(spagxe_mgpu *mg, int64_t n, const double *R, const double *E, int trait, const double *Y,
                           const double *cova, int p);
/* geno: n_shards == 1: one HOST genotype source, cut into contiguous SNP ranges, one per GPU;
 *       n_shards == n_gpus: shard k is tested on GPU k and may live on the host or in that GPU's memory.
 * out: HOST matrix (sum of the shards' SNPs) x 10, column-major, rows in shard order.  diag: totals (times: maximum). */
int spagxe_mgpu_run_cct(spagxe_mgpu *mg, const spagxe_geno *geno, int n_shards, const spagxe_params *prm, double *out,
                        spagxe_diag *diag);
/* the other two loops and the survival trait on several GPUs (SPAGxEmix_CCT R/SPAGxECCT.R:975-997, SPAGxE_Plus
 * R/SPAGxEPlus_functions_MYZ.R:214-233): spagxe_mgpu_set_inputs first (trait 0 = no Wald data: SPAGxE_Plus, or survival
 * whose data follow here), then the per-analysis extras, which every GPU receives from the host arrays; out is
 * M x 12 (mix) / M x 8 (plus).  Same sharding, still no collective on the data path.                                */
int spagxe_mgpu_set_wald_survival(spagxe_mgpu *mg, const double *surv_time, const double *event, const double *cova, int p);
int spagxe_mgpu_set_pcs(spagxe_mgpu *mg, const double *pcs, int k);
int spagxe_mgpu_set_grm(spagxe_mgpu *mg, int64_t nnz, const int32_t *gi, const int32_t *gj, const double *gv,
                        double R_GRM_R, double R_GRM_RE, double Rnew_GRM_Rnew, const double *Rnew);
int spagxe_mgpu_run_mix(spagxe_mgpu *mg, const spagxe_geno *geno, int n_shards, const spagxe_params *prm, double *out,
                        spagxe_diag *diag);
int spagxe_mgpu_run_plus(spagxe_mgpu *mg, const spagxe_geno *geno, int n_shards, const spagxe_params *prm, double *out,
                         spagxe_diag *diag);
/* SPAGxEmix_CCT_Plus on several GPUs: spagxe_mgpu_set_inputs (trait 0), spagxe_mgpu_set_pcs, spagxe_mgpu_set_sparse_grm;
 * out is M x 12 with NA in the Wald / CCT columns of branch-B rows (see spagxe_run_mix_plus).                    */
int spagxe_mgpu_set_sparse_grm(spagxe_mgpu *mg, int64_t nnz, const int32_t *gi, const int32_t *gj, const double *gv);
int spagxe_mgpu_run_mix_plus(spagxe_mgpu *mg, const spagxe_geno *geno, int n_shards, const spagxe_params *prm, double *out,
                             spagxe_diag *diag);

/* ---- genotype unpack (K1) exposed for the bit-exact decode checks ----------- */
/* per-SNP code counts of packed rows: counts[4*j + c], c = 00,01,10,11 (GRAB decode: 2,NA,1,0);
 * replaces the counting part of R/SPAGxECCT.R:557-560.  counts is a HOST array of 4*M int64. */
int spagxe_bed_counts(spagxe_ctx *ctx, const spagxe_geno *geno, int64_t *counts);
/* decode packed rows into an N x M double matrix (host), NA_real_ for missing: what
 * GRAB::GRAB.ReadGeno(...)$GenoMat returns (R/SPAGxECCT.R:484-487).                       */
int spagxe_bed_decode(spagxe_ctx *ctx, const spagxe_geno *geno, double *out_host);

/* raw outputs of the fused unpack+score pass (K1+K2) for M SNPs, all HOST arrays of length M:
 *   D0,D1 = sum over non-missing of dosage*R, dosage*V;  T0,T1 = sum over missing of R, V;
 *   asum = n_het + 2 n_homA1;  nmiss = #missing.  impl: 0 = tensor-core kernel (default), 1 = fp64
 *   CUDA-core kernel.  Replaces the vector passes at R/SPAGxECCT.R:557-560, :582, :591; exposed so
 *   that the two kernels can be checked against each other and against the oracle.            */
int spagxe_score_stats(spagxe_ctx *ctx, const spagxe_geno *geno, int impl, double *D0, double *D1, double *T0,
                       double *T1, int32_t *asum, int32_t *nmiss);

/* measured device peaks used by bench.py for the SPA stage's fp64 fraction (DFMA loop)   */
int spagxe_measure_fp64_peak(spagxe_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* SPAGXE_H */
