// spagxe_cuda.cu -- C ABI of libspagxe_cuda.so (include/spagxe.h) and the host-side pipeline.
//
// One context drives one GPU.  A run walks the SNP axis in chunks; each chunk is
//   H2D (copy stream, pinned staging, double-buffered)  ->  score pass  ->  per-SNP epilogue + routing
//   ->  SPA work list  ->  branch-B work list
// with no host synchronisation inside the loop: work lists are consumed by persistent CTAs that read
// their length from device memory.
#include <cuda_runtime.h>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/spagxe.h"
#include "snp_kernels.h"
#include "score_launch.h"
#include "mix_batch.h"
#include "bb_farm.h"
#include "cox_wald.h"
#include "clm_wald.h"

using namespace spagxe;

static std::string g_create_error;

struct spagxe_ctx {
    int device = 0;
    std::string err;
    cudaStream_t s_main = nullptr, s_copy = nullptr, s_user = nullptr;
    bool user_stream = false;
    int num_sms = 148;
    // analysis inputs
    int64_t n = 0;
    double *d_R = nullptr, *d_E = nullptr, *d_Rn = nullptr, *d_V = nullptr;
    VecConst *d_vc = nullptr;
    Counters *d_ctr = nullptr;
    int vec_mode = -1;  // which V is currently prepared (-1 none)
    // wald
    double *d_Y = nullptr, *d_cova = nullptr;
    int wald_p = 0, wald_trait = 0;
    // pcs / grm (mix, plus)
    double *d_pcs = nullptr; int n_pcs = 0;
    double *d_xtx_inv = nullptr;
    int64_t grm_nnz = 0; int *d_gi = nullptr, *d_gj = nullptr; double *d_gv = nullptr;
    double R_GRM_R = 0, R_GRM_RE = 0, Rnew_GRM_Rnew = 0; double *d_Rn_user = nullptr; bool have_grm = false, have_coo = false;
    // chunk scratch
    int64_t cap_chunk_bytes = 0; uint8_t *d_bed[2] = {nullptr, nullptr};
    int cap_m = 0; double *d_stats = nullptr; int *d_istats = nullptr;
    int64_t cap_spa = 0;                                   // capacity of the saddlepoint work lists (SNPs)
    int64_t cap_partial = 0; double *d_partial = nullptr; int *d_ipartial = nullptr;
    SpaItem *d_spa_list = nullptr, *d_spax_list = nullptr; int *d_b_list = nullptr;
    BItem *d_bitems = nullptr; int64_t cap_bitems = 0;
    Quad quad = {nullptr, nullptr, nullptr, nullptr, nullptr}; bool quad_on = false;
    // tensor-core score pass
    int8_t *d_bmat = nullptr; int64_t cap_bmat = 0; BmatInfo *d_binfo = nullptr; long long *d_acc = nullptr; int cap_acc = 0;
    bool use_tc = true; int tc_cfg = 0;
    // test / tuning hooks (spagxe_ctx_set_option); no environment variable is ever read by this library
    int64_t opt_chunk_snps = 0; bool opt_exact_spa = false, opt_mix_per_snp = false; int opt_mix_cluster = 4;   // 4..6 CTAs per SNP measure alike at N = 200,000 (8: 6 % slower, 2: 12 % slower)
    int64_t cap_out = 0; double *d_out = nullptr;
    int64_t cap_dense = 0; void *d_dense = nullptr;
    int64_t cap_local = 0; uint8_t *d_local = nullptr;     // SPAGxEmixCCT_localance: packed rows of g, haplo and the haplotype covariates
    BbFarm *farm = nullptr;                                // SPAGxE_CCT branch-B solver farm (K5)
    CoxData *cox = nullptr;                                // traits = "survival": time order, tie groups, gathered columns
    std::vector<double> h_time, h_event;                   // Phen.mtx$surv.time / $event (host copies for the sort)
    ClmData *clm = nullptr;                                // traits = "categorical": category order, gathered columns
    std::vector<double> h_ycat;                            // level index of Phen.mtx$Y (host copy for the ordering)
    MixScratch *mix_scratch = nullptr;                    // SPAGxEmix: scratch of the batched common path
    int64_t cap_mcache = 0; double *d_mcache = nullptr;   // SPAGxEmix: cached allele frequencies, one vector per cluster
    // pinned staging
    uint8_t *h_stage[2] = {nullptr, nullptr}; size_t stage_bytes = 0; cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;   // timing events, created once and reused by every call
    int64_t cap_src = 0; uint8_t *d_src = nullptr;           // file-layout rows of a chunk before the sample gather
    int64_t cap_sidx = 0; int64_t *d_sidx = nullptr;         // device copy of spagxe_geno.sample_index
    int64_t cap_util = 0; void *d_util = nullptr;            // scratch of spagxe_bed_counts / spagxe_bed_decode
    int *h_flag = nullptr;                                   // pinned word for small device -> host flags
    int64_t launches = 0;
    cudaStream_t stream() const { return user_stream ? s_user : s_main; }
};

static int check_grm_indices(spagxe_ctx *ctx, const char *who, int64_t nnz, const int32_t *gi, const int32_t *gj, int64_t n);
static int fail(spagxe_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, SPAGXE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                       \
    } while (0)
#define LAUNCH_CHECK()                                                                             \
    do {                                                                                           \
        ctx->launches++;                                                                           \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, SPAGXE_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

template <typename T>
static cudaError_t dev_realloc(T **p, size_t count) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) return cudaSuccess;
    return cudaMalloc((void **)p, count * sizeof(T));
}

extern "C" const char *spagxe_version(void) { return "spagxe-b200 0.1 (sm_100a)"; }

extern "C" const char *spagxe_last_error(const spagxe_ctx *ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int spagxe_ctx_create(int device, spagxe_ctx **out) {
    spagxe_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, SPAGXE_ERR_ARG, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SPAGXE_ERR_CUDA, "no CUDA device: %s (libspagxe_cuda has no CPU fallback)",
                    cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, SPAGXE_ERR_ARG, "device %d out of range (%d)", device, ndev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, SPAGXE_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, SPAGXE_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    ctx = new spagxe_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    auto bail = [&](const char *what, cudaError_t ee) {
        int rc = fail(nullptr, SPAGXE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(ee));
        delete ctx;
        return rc;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->s_main, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
    for (int b = 0; b < 2; b++) {
        if ((e = cudaEventCreateWithFlags(&ctx->ev_stage[b], cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
        if ((e = cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
        if ((e = cudaEventCreateWithFlags(&ctx->ev_done[b], cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
    }
    if ((e = cudaMalloc((void **)&ctx->d_vc, sizeof(VecConst))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&ctx->d_ctr, sizeof(Counters))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = bb_farm_create(&ctx->farm, device, ctx->num_sms)) != cudaSuccess) return bail("bb_farm_create", e);
    // opt in to the large dynamic shared memory of the per-SNP SPAGxEmix kernel
    e = cudaFuncSetAttribute(k_mix_snp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)branch_b_smem(kMaxQ));
    if (e != cudaSuccess) return bail("cudaFuncSetAttribute(k_mix_snp)", e);
    e = cudaFuncSetAttribute(k_mixplus_snp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)branch_b_smem(kMaxQ));
    if (e != cudaSuccess) return bail("cudaFuncSetAttribute(k_mixplus_snp)", e);
    e = cudaFuncSetAttribute(k_mix_dosage_snp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)branch_b_smem(kMaxQ));
    if (e != cudaSuccess) return bail("cudaFuncSetAttribute(k_mix_dosage_snp)", e);
    e = cudaFuncSetAttribute(k_mixplus_dosage_snp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)branch_b_smem(kMaxQ));
    if (e != cudaSuccess) return bail("cudaFuncSetAttribute(k_mixplus_dosage_snp)", e);
    e = cudaFuncSetAttribute(k_cct_dosage_snp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)branch_b_smem(kMaxQ));
    if (e != cudaSuccess) return bail("cudaFuncSetAttribute(k_cct_dosage_snp)", e);
    e = cudaFuncSetAttribute(k_localance_snp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)branch_b_smem(kMaxQ));
    if (e != cudaSuccess) return bail("cudaFuncSetAttribute(k_localance_snp)", e);
    if ((e = cudaMalloc((void **)&ctx->quad.x, sizeof(double) * kQBins * kQNodes)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&ctx->quad.w, sizeof(double) * kQBins * kQNodes)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&ctx->quad.wfix, sizeof(long long) * kQBins * kQNodes)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&ctx->quad.range, sizeof(double) * 4)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&ctx->quad.count, sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&ctx->d_binfo, sizeof(BmatInfo))) != cudaSuccess) return bail("cudaMalloc", e);
    {
        std::string serr;
        if ((e = score_init(&serr)) != cudaSuccess) return bail(serr.c_str(), e);
    }
    if ((e = cudaMallocHost((void **)&ctx->h_flag, 4 * sizeof(int))) != cudaSuccess) return bail("cudaMallocHost", e);
    *out = ctx;
    return SPAGXE_OK;
}

extern "C" void spagxe_ctx_destroy(spagxe_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(ctx->d_R); cudaFree(ctx->d_E); cudaFree(ctx->d_Rn); cudaFree(ctx->d_V); cudaFree(ctx->d_vc);
    cudaFree(ctx->d_ctr); cudaFree(ctx->d_Y); cudaFree(ctx->d_cova); cudaFree(ctx->d_pcs); cudaFree(ctx->d_xtx_inv);
    cudaFree(ctx->d_gi); cudaFree(ctx->d_gj); cudaFree(ctx->d_gv); cudaFree(ctx->d_Rn_user);
    cudaFree(ctx->d_bed[0]); cudaFree(ctx->d_bed[1]); cudaFree(ctx->d_stats); cudaFree(ctx->d_istats);
    cudaFree(ctx->d_partial); cudaFree(ctx->d_ipartial); cudaFree(ctx->d_spa_list); cudaFree(ctx->d_b_list);
    cudaFree(ctx->d_out); cudaFree(ctx->d_dense); cudaFree(ctx->d_local); cudaFree(ctx->d_spax_list); cudaFree(ctx->d_mcache);
    mix_scratch_destroy(ctx->mix_scratch);
    bb_farm_destroy(ctx->farm);
    cox_destroy(ctx->cox);
    clm_destroy(ctx->clm);
    cudaFree(ctx->d_bmat); cudaFree(ctx->d_binfo); cudaFree(ctx->d_acc); cudaFree(ctx->d_bitems);
    if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
    cudaFree(ctx->d_util); cudaFree(ctx->d_src); cudaFree(ctx->d_sidx);
    for (auto ev : ctx->ev_pool) cudaEventDestroy(ev);
    cudaFree(ctx->quad.x); cudaFree(ctx->quad.w); cudaFree(ctx->quad.wfix); cudaFree(ctx->quad.range); cudaFree(ctx->quad.count);
    for (int b = 0; b < 2; b++) {
        if (ctx->h_stage[b]) cudaFreeHost(ctx->h_stage[b]);
        if (ctx->ev_stage[b]) cudaEventDestroy(ctx->ev_stage[b]);
        if (ctx->ev_copied[b]) cudaEventDestroy(ctx->ev_copied[b]);
        if (ctx->ev_done[b]) cudaEventDestroy(ctx->ev_done[b]);
    }
    if (ctx->s_main) cudaStreamDestroy(ctx->s_main);
    if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
    delete ctx;
}

extern "C" int spagxe_ctx_set_stream(spagxe_ctx *ctx, void *cuda_stream) {
    if (!ctx) return SPAGXE_ERR_ARG;
    ctx->s_user = (cudaStream_t)cuda_stream;
    ctx->user_stream = true;  // NULL here means the legacy default stream, which torch uses by default
    return SPAGXE_OK;
}

// Test / tuning hooks.  They select between bit-compatible implementations (chunking, exact vs quadrature SPA sums,
// batched vs per-SNP SPAGxEmix route, fp64 vs tensor-core score kernel) so that the tests can compare them; defaults are
// the product configuration.  Nothing here is read from the environment.
extern "C" int spagxe_ctx_set_option(spagxe_ctx *ctx, int option, int64_t value) {
    if (!ctx) return SPAGXE_ERR_ARG;
    switch (option) {
        case SPAGXE_OPT_CHUNK_SNPS: if (value < 0) break; ctx->opt_chunk_snps = value; return SPAGXE_OK;
        case SPAGXE_OPT_EXACT_SPA: ctx->opt_exact_spa = value != 0; ctx->vec_mode = -1; return SPAGXE_OK;
        case SPAGXE_OPT_MIX_PER_SNP: ctx->opt_mix_per_snp = value != 0; return SPAGXE_OK;
        case SPAGXE_OPT_MIX_CLUSTER: if (value < 1 || value > 8) break; ctx->opt_mix_cluster = (int)value; return SPAGXE_OK;
        case SPAGXE_OPT_SCORE_IMPL: if (value != 0 && value != 1) break; ctx->use_tc = (value == 0); return SPAGXE_OK;
#ifdef SPAGXE_PROBES
        case SPAGXE_OPT_PROBE_CFG: ctx->tc_cfg = (int)value; return SPAGXE_OK;
#endif
        default: break;
    }
    return fail(ctx, SPAGXE_ERR_ARG, "set_option: unknown option %d or bad value %lld", option, (long long)value);
}

static int set_vectors_common(spagxe_ctx *ctx, int64_t n, const double *R, const double *E, cudaMemcpyKind kind) {
    if (!ctx) return SPAGXE_ERR_ARG;
    if (n <= 0 || !R || !E) return fail(ctx, SPAGXE_ERR_ARG, "set_vectors: n=%lld, R=%p, E=%p", (long long)n, R, E);
    CU(cudaSetDevice(ctx->device));
    if (n != ctx->n) {
        CU(dev_realloc(&ctx->d_R, n)); CU(dev_realloc(&ctx->d_E, n));
        CU(dev_realloc(&ctx->d_Rn, n)); CU(dev_realloc(&ctx->d_V, n));
        // inputs that depend on N are invalidated
        cudaFree(ctx->d_Y); ctx->d_Y = nullptr; cudaFree(ctx->d_cova); ctx->d_cova = nullptr; ctx->wald_trait = 0;
        cudaFree(ctx->d_pcs); ctx->d_pcs = nullptr; ctx->n_pcs = 0;
        ctx->have_grm = false; ctx->have_coo = false;
        ctx->n = n;
    }
    cudaStream_t s = ctx->stream();
    CU(cudaMemcpyAsync(ctx->d_R, R, n * sizeof(double), kind, s));
    CU(cudaMemcpyAsync(ctx->d_E, E, n * sizeof(double), kind, s));
    k_vec_prep1<<<1, 1024, 0, s>>>(ctx->d_R, ctx->d_E, n, ctx->d_vc);
    LAUNCH_CHECK();
    ctx->vec_mode = -1;
    cox_invalidate(ctx->cox);   // E is one of the Cox design columns
    clm_invalidate(ctx->clm);   // and of the cumulative-link design
    // R and E must be finite: the fixed-point digits of the score pass and the SPA sums are undefined otherwise
    // (R would carry the NaN through sum() and stop in the first `if`)
    CU(cudaMemcpyAsync(ctx->h_flag, &ctx->d_vc->nonfinite, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (ctx->h_flag[0]) {
        ctx->n = 0;
        return fail(ctx, SPAGXE_ERR_ARG, "set_vectors: R or E holds %d non-finite value(s)", ctx->h_flag[0]);
    }
    return SPAGXE_OK;
}
extern "C" int spagxe_set_vectors(spagxe_ctx *ctx, int64_t n, const double *R, const double *E) {
    return set_vectors_common(ctx, n, R, E, cudaMemcpyHostToDevice);
}
extern "C" int spagxe_set_vectors_device(spagxe_ctx *ctx, int64_t n, const double *dR, const double *dE) {
    return set_vectors_common(ctx, n, dR, dE, cudaMemcpyDeviceToDevice);
}

static int set_wald_common(spagxe_ctx *ctx, int trait, const double *Y, const double *cova, int p, cudaMemcpyKind kind);
extern "C" int spagxe_set_wald(spagxe_ctx *ctx, int trait, const double *Y, const double *cova, int p) {
    return set_wald_common(ctx, trait, Y, cova, p, cudaMemcpyHostToDevice);
}
extern "C" int spagxe_set_wald_device(spagxe_ctx *ctx, int trait, const double *dY, const double *dcova, int p) {
    return set_wald_common(ctx, trait, dY, dcova, p, cudaMemcpyDeviceToDevice);
}
static int set_wald_common(spagxe_ctx *ctx, int trait, const double *Y, const double *cova, int p, cudaMemcpyKind kind) {
    if (!ctx) return SPAGXE_ERR_ARG;
    if (ctx->n <= 0) return fail(ctx, SPAGXE_ERR_STATE, "set_wald: call spagxe_set_vectors first");
    if (trait == SPAGXE_TRAIT_SURVIVAL)
        return fail(ctx, SPAGXE_ERR_ARG, "traits = survival takes two vectors: use spagxe_set_wald_survival(time, event, cova, p)");
    if (trait != SPAGXE_TRAIT_BINARY && trait != SPAGXE_TRAIT_QUANTITATIVE && trait != SPAGXE_TRAIT_CATEGORICAL)
        return fail(ctx, SPAGXE_ERR_ARG, "traits must be binary(1), quantitative(2), survival(3) or categorical(4)");
    if (!Y || p < 0 || (p > 0 && !cova)) return fail(ctx, SPAGXE_ERR_ARG, "set_wald: bad arguments");
    if (p + 4 > kMaxQ) return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "set_wald: at most %d covariates", kMaxQ - 4);
    if (trait == SPAGXE_TRAIT_CATEGORICAL) {
        // ref R/SPAGxECCT.R:666: clm(as.factor(Phen.mtx$Y) ~ ...); Y holds the index of each sample's level, 0 .. J-1
        if (p > kClmMaxCov) return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "set_wald: the categorical Wald test takes at most %d covariate columns", kClmMaxCov);
        CU(cudaSetDevice(ctx->device));
        ctx->h_ycat.resize((size_t)ctx->n);
        if (kind == cudaMemcpyHostToDevice) std::copy(Y, Y + ctx->n, ctx->h_ycat.begin());
        else CU(cudaMemcpy(ctx->h_ycat.data(), Y, ctx->n * sizeof(double), cudaMemcpyDeviceToHost));
        bool seen[kClmMaxJ] = {false}; int J = 0;
        for (int64_t i = 0; i < ctx->n; i++) {
            const double y = ctx->h_ycat[(size_t)i];
            if (!(y >= 0 && y < kClmMaxJ) || y != std::floor(y))
                return fail(ctx, SPAGXE_ERR_ARG, "set_wald: categorical Y[%lld] = %g is not a level index 0 .. %d", (long long)i, y, kClmMaxJ - 1);
            seen[(int)y] = true; J = std::max(J, (int)y + 1);
        }
        if (J < 2) return fail(ctx, SPAGXE_ERR_ARG, "set_wald: a categorical trait needs at least two levels");
        for (int k = 0; k < J; k++)
            if (!seen[k]) return fail(ctx, SPAGXE_ERR_ARG, "set_wald: level index %d has no samples (as.factor() has no empty levels)", k);
        if (!ctx->clm) ctx->clm = clm_create();
        clm_invalidate(ctx->clm);
    }
    CU(cudaSetDevice(ctx->device));
    CU(dev_realloc(&ctx->d_Y, ctx->n));
    CU(dev_realloc(&ctx->d_cova, (size_t)ctx->n * std::max(p, 1)));
    CU(cudaMemcpyAsync(ctx->d_Y, Y, ctx->n * sizeof(double), kind, ctx->stream()));
    if (p > 0) CU(cudaMemcpyAsync(ctx->d_cova, cova, (size_t)ctx->n * p * sizeof(double), kind, ctx->stream()));
    CU(cudaStreamSynchronize(ctx->stream()));
    ctx->wald_p = p; ctx->wald_trait = trait;
    return SPAGXE_OK;
}

// traits = "survival" (ref R/SPAGxECCT.R:622-634): Phen.mtx$surv.time, Phen.mtx$event and as.matrix(Cova.mtx); host pointers.
// The time order and the gathered design columns are built at the next run (they also depend on E).
extern "C" int spagxe_set_wald_survival(spagxe_ctx *ctx, const double *surv_time, const double *event, const double *cova, int p) {
    if (!ctx) return SPAGXE_ERR_ARG;
    if (ctx->n <= 0) return fail(ctx, SPAGXE_ERR_STATE, "set_wald_survival: call spagxe_set_vectors first");
    if (!surv_time || !event || p < 0 || (p > 0 && !cova)) return fail(ctx, SPAGXE_ERR_ARG, "set_wald_survival: bad arguments");
    if (p > kCoxMaxCov) return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "set_wald_survival: at most %d covariate columns", kCoxMaxCov);
    int64_t n_events = 0;
    for (int64_t i = 0; i < ctx->n; i++) {
        if (!std::isfinite(surv_time[i])) return fail(ctx, SPAGXE_ERR_ARG, "set_wald_survival: surv.time[%lld] is not finite", (long long)i);
        if (event[i] != 0.0 && event[i] != 1.0) return fail(ctx, SPAGXE_ERR_ARG, "set_wald_survival: event must be 0 / 1");
        n_events += event[i] != 0.0;
    }
    if (n_events == 0) return fail(ctx, SPAGXE_ERR_ARG, "set_wald_survival: no events");
    CU(cudaSetDevice(ctx->device));
    ctx->h_time.assign(surv_time, surv_time + ctx->n);
    ctx->h_event.assign(event, event + ctx->n);
    CU(dev_realloc(&ctx->d_Y, ctx->n));     // the solver farm stages a Y column; it is not used for this trait
    CU(dev_realloc(&ctx->d_cova, (size_t)ctx->n * std::max(p, 1)));
    CU(cudaMemcpyAsync(ctx->d_Y, event, ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream()));
    if (p > 0) CU(cudaMemcpyAsync(ctx->d_cova, cova, (size_t)ctx->n * p * sizeof(double), cudaMemcpyHostToDevice, ctx->stream()));
    CU(cudaStreamSynchronize(ctx->stream()));
    ctx->wald_p = p; ctx->wald_trait = SPAGXE_TRAIT_SURVIVAL;
    if (!ctx->cox) ctx->cox = cox_create();
    cox_invalidate(ctx->cox);
    return SPAGXE_OK;
}

// ---------------------------------------------------------------------------
static int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static int ensure_vec_mode(spagxe_ctx *ctx, int mode, const double *d_Rn_in) {
    if (ctx->vec_mode == mode) return SPAGXE_OK;
    k_vec_prep2<<<1, 1024, 0, ctx->stream()>>>(ctx->d_R, ctx->d_E, d_Rn_in, ctx->n, mode == 0 ? 0 : 1, ctx->d_Rn, ctx->d_V, ctx->d_vc);
    LAUNCH_CHECK();
    ctx->vec_mode = mode;
    {   // fixed-point int8 digit matrix B of (R, V) for the tensor-core score pass
        cudaStream_t s = ctx->stream();
        const int64_t need = score_bmat_bytes(ctx->n);
        if (need > ctx->cap_bmat) { CU(cudaStreamSynchronize(s)); CU(dev_realloc(&ctx->d_bmat, need)); ctx->cap_bmat = need; }
        CU(score_build_b(s, ctx->d_R, ctx->d_V, ctx->n, ctx->d_binfo, ctx->d_bmat, &ctx->launches));
    }
    // SNP-invariant quadrature of the SPA vector (R.new): only worth it when N >> number of nodes
    const double *rv = (mode == 0) ? ctx->d_V : ctx->d_Rn;
    ctx->quad_on = ctx->n >= 8ll * kQBins * kQNodes && !ctx->opt_exact_spa;
    if (ctx->quad_on) {
        cudaStream_t s = ctx->stream();
        CU(cudaMemsetAsync(ctx->quad.wfix, 0, sizeof(long long) * kQBins * kQNodes, s));
        k_quad_range<<<1, 1024, 0, s>>>(rv, ctx->n, ctx->quad);
        LAUNCH_CHECK();
        k_quad_accumulate<<<(unsigned)((ctx->n + 255) / 256), 256, 0, s>>>(rv, ctx->n, ctx->quad);
        LAUNCH_CHECK();
        k_quad_finish<<<1, 1024, 0, s>>>(ctx->quad);
        LAUNCH_CHECK();
    }
    return SPAGXE_OK;
}

// SPA work list of the current chunk: quadrature first, exact kernel for what it hands back
static int launch_spa(spagxe_ctx *ctx, const double *rv, int64_t N, int64_t M, double *d_out) {
    cudaStream_t s = ctx->stream();
    if (ctx->quad_on) {
        CU(cudaMemsetAsync(&ctx->d_ctr->work_cursor, 0, sizeof(int), s));   // item queue of k_spa_quad
        k_spa_quad<<<ctx->num_sms * 3, kSpaQThreads, 0, s>>>(ctx->d_spa_list, ctx->d_ctr, ctx->quad, ctx->d_spax_list, M, d_out);   // 3 resident CTAs per SM (80 registers)
        LAUNCH_CHECK();
        k_spa_homo<<<ctx->num_sms * 2, kSpaThreads, 0, s>>>(ctx->d_spax_list, &ctx->d_ctr->spax_count, ctx->d_ctr, rv, N, M, d_out);
        LAUNCH_CHECK();
    } else {
        k_spa_homo<<<ctx->num_sms * 2, kSpaThreads, 0, s>>>(ctx->d_spa_list, &ctx->d_ctr->spa_count, ctx->d_ctr, rv, N, M, d_out);
        LAUNCH_CHECK();
    }
    return SPAGXE_OK;
}

// Deferred branch B: work items accumulate on the device; a flush launches the kernel that drains them.  The item count
// is read ON THE DEVICE (persistent grids), so the host never synchronises here and the next chunk's copies keep flowing.
static int flush_branch_b(spagxe_ctx *ctx, int mode, int64_t M, int64_t N, WaldData wd, double min_maf, double cutoff,
                          double *d_out, int g_model) {
    cudaStream_t s = ctx->stream();
    if (mode == 0) {   // SPAGxE_CCT: the solver farm (cooperative persistent kernel, one CTA per SM)
        BbFarmArgs a;
        a.items = ctx->d_bitems; a.count = &ctx->d_ctr->b_count; a.m_total = M; a.n = N;
        a.R = ctx->d_R; a.E = ctx->d_E; a.vc = ctx->d_vc; a.Y = wd.Y; a.cova = wd.cova; a.p = wd.p; a.trait = wd.trait;
        a.min_maf = min_maf; a.g_model = g_model; a.out = d_out; a.ctr = ctx->d_ctr;
        const char *what = "bb_farm_launch";
        cudaError_t e = bb_farm_launch(ctx->farm, s, a, &ctx->launches, &what);
        if (e != cudaSuccess) return fail(ctx, SPAGXE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
        if (wd.trait == SPAGXE_TRAIT_SURVIVAL) {   // the farm left the Wald / CCT columns to the Cox kernel
            e = cox_launch(ctx->cox, s, ctx->num_sms, ctx->d_bitems, &ctx->d_ctr->b_count, M, g_model, d_out, ctx->d_ctr, &ctx->launches);
            if (e != cudaSuccess) return fail(ctx, SPAGXE_ERR_CUDA, "cox_launch: %s", cudaGetErrorString(e));
        }
        if (wd.trait == SPAGXE_TRAIT_CATEGORICAL) {   // ... or to the cumulative-link kernel
            e = clm_launch(ctx->clm, s, ctx->num_sms, ctx->d_bitems, &ctx->d_ctr->b_count, M, g_model, d_out, ctx->d_ctr, &ctx->launches);
            if (e != cudaSuccess) return fail(ctx, SPAGXE_ERR_CUDA, "clm_launch: %s", cudaGetErrorString(e));
        }
    } else {           // SPAGxE_Plus: one 4-CTA cluster per item, persistent over the device-side list
        constexpr int cs = 4;
        cudaLaunchConfig_t cfg = {};
        const int n_clusters = std::max(1, ctx->num_sms / cs);
        cfg.gridDim = dim3((unsigned)(n_clusters * cs));
        cfg.blockDim = dim3(kBThreads);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const BItem *items = ctx->d_bitems;
        const int *count = &ctx->d_ctr->b_count;
        Counters *ctr = ctx->d_ctr;
        const double *R = ctx->d_R, *E = ctx->d_E;
        const VecConst *vc = ctx->d_vc;
        GrmCoo grm{ctx->d_gi, ctx->d_gj, ctx->d_gv, (long long)ctx->grm_nnz};
        CU(cudaLaunchKernelEx(&cfg, k_branch_b_plus, items, count, ctr, M, N, R, E, vc, grm, cutoff, d_out, g_model));
        ctx->launches++;
    }
    CU(cudaMemsetAsync(&ctx->d_ctr->b_count, 0, sizeof(int), s));
    return SPAGXE_OK;
}

struct ChunkPlan {
    int64_t n, m, row_bytes, dpitch;  // dpitch: device pitch of staged rows (multiple of 128)
    int chunk_m;
};

static int ensure_chunk_scratch(spagxe_ctx *ctx, const ChunkPlan &pl, bool need_bed_buffers, int64_t out_elems) {
    const int64_t bytes = pl.dpitch * pl.chunk_m;
    if (need_bed_buffers && bytes > ctx->cap_chunk_bytes) {
        CU(dev_realloc(&ctx->d_bed[0], bytes)); CU(dev_realloc(&ctx->d_bed[1], bytes));
        ctx->cap_chunk_bytes = bytes;
    }
    if (pl.chunk_m > ctx->cap_m) {
        CU(dev_realloc(&ctx->d_stats, (size_t)6 * pl.chunk_m)); CU(dev_realloc(&ctx->d_istats, (size_t)3 * pl.chunk_m));
        CU(dev_realloc(&ctx->d_b_list, pl.chunk_m));
        ctx->cap_m = pl.chunk_m;
    }
    {   // saddlepoint work lists: the items carry no genotype rows, so they are collected over the chunks of a call and
        // solved by one launch (a list holds up to 2^20 SNPs' worth; longer calls flush in between)
        const int64_t want = std::max<int64_t>(pl.chunk_m, std::min<int64_t>(pl.m, 1ll << 20));
        if (want > ctx->cap_spa) {
            CU(dev_realloc(&ctx->d_spa_list, want)); CU(dev_realloc(&ctx->d_spax_list, want));
            ctx->cap_spa = want;
        }
    }
    if (out_elems > ctx->cap_out) { CU(dev_realloc(&ctx->d_out, out_elems)); ctx->cap_out = out_elems; }
    return SPAGXE_OK;
}

// copy `rows` packed rows from host memory into a pitched device buffer on the copy stream
static int stage_rows(spagxe_ctx *ctx, const uint8_t *src, int64_t spitch, int64_t row_bytes, int64_t rows,
                      uint8_t *dst, int64_t dpitch, bool pinned) {
    if (pinned) {
        CU(cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, rows, cudaMemcpyHostToDevice, ctx->s_copy));
        return SPAGXE_OK;
    }
    if (!ctx->h_stage[0]) {
        ctx->stage_bytes = 32u << 20;
        for (int b = 0; b < 2; b++) CU(cudaMallocHost((void **)&ctx->h_stage[b], ctx->stage_bytes));
    }
    const int64_t rows_per = std::max<int64_t>(1, (int64_t)ctx->stage_bytes / row_bytes);
    if (row_bytes > (int64_t)ctx->stage_bytes) return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "row larger than staging buffer");
    int b = 0;
    for (int64_t r0 = 0; r0 < rows; r0 += rows_per, b ^= 1) {
        const int64_t nr = std::min(rows_per, rows - r0);
        CU(cudaEventSynchronize(ctx->ev_stage[b]));
        if (spitch == row_bytes) memcpy(ctx->h_stage[b], src + r0 * spitch, nr * row_bytes);
        else for (int64_t r = 0; r < nr; r++) memcpy(ctx->h_stage[b] + r * row_bytes, src + (r0 + r) * spitch, row_bytes);
        CU(cudaMemcpy2DAsync(dst + r0 * dpitch, dpitch, ctx->h_stage[b], row_bytes, row_bytes, nr,
                             cudaMemcpyHostToDevice, ctx->s_copy));
        CU(cudaEventRecord(ctx->ev_stage[b], ctx->s_copy));
    }
    return SPAGXE_OK;
}

static bool host_ptr_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int validate_geno(spagxe_ctx *ctx, const spagxe_geno *g) {
    if (!g || !g->data) return fail(ctx, SPAGXE_ERR_ARG, "geno is NULL");
    if (g->n_samples <= 0 || g->n_snps < 0) return fail(ctx, SPAGXE_ERR_ARG, "geno: bad shape");
    if (g->n_snps > INT_MAX / 2) return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "geno: too many SNPs in one call");
    const bool prep = (g->sample_index != nullptr) || (g->flags & SPAGXE_GENO_COUNT_A2);
    if (prep && g->kind != SPAGXE_GENO_BED) return fail(ctx, SPAGXE_ERR_ARG, "geno: sample_index / flags apply to .bed input only");
    if (g->kind == SPAGXE_GENO_BED) {
        const int64_t nfile = (g->sample_index && g->n_file_samples > 0) ? g->n_file_samples : g->n_samples;
        if (g->pitch < (nfile + 3) / 4) return fail(ctx, SPAGXE_ERR_ARG, "geno: pitch < ceil(samples per file row / 4)");
        if (g->sample_index)
            for (int64_t i = 0; i < g->n_samples; i++)
                if (g->sample_index[i] < 0 || g->sample_index[i] >= nfile)
                    return fail(ctx, SPAGXE_ERR_ARG, "geno: sample_index[%lld] = %lld is outside the file row (%lld samples)",
                                (long long)i, (long long)g->sample_index[i], (long long)nfile);
        if (!prep && g->location == SPAGXE_LOC_DEVICE && ((g->pitch & 15) || ((uintptr_t)g->data & 15)))
            return fail(ctx, SPAGXE_ERR_ARG, "geno: device-resident .bed rows need a 16-byte aligned base and pitch");
    } else if (g->kind == SPAGXE_GENO_F64 || g->kind == SPAGXE_GENO_I32) {
        if (g->pitch < g->n_samples) return fail(ctx, SPAGXE_ERR_ARG, "geno: leading dimension < N");
    } else return fail(ctx, SPAGXE_ERR_ARG, "geno: unknown kind %d", g->kind);
    if (g->location != SPAGXE_LOC_HOST && g->location != SPAGXE_LOC_DEVICE) return fail(ctx, SPAGXE_ERR_ARG, "geno: bad location");
    return SPAGXE_OK;
}

static bool geno_needs_prepare(const spagxe_geno *g) {
    return g->kind == SPAGXE_GENO_BED && (g->sample_index != nullptr || (g->flags & SPAGXE_GENO_COUNT_A2));
}

static ChunkPlan make_plan(const spagxe_geno *g, int64_t chunk_override) {
    ChunkPlan pl;
    pl.n = g->n_samples; pl.m = g->n_snps;
    pl.row_bytes = (g->n_samples + 3) / 4;
    if (g->kind == SPAGXE_GENO_BED && g->location == SPAGXE_LOC_DEVICE && !geno_needs_prepare(g)) pl.dpitch = g->pitch;
    else pl.dpitch = round_up(pl.row_bytes, 128);
    // ~1 GiB of packed rows per chunk (>> L2), dense inputs are limited by their 8 B/sample staging
    int64_t budget = (g->kind == SPAGXE_GENO_BED) ? (1ll << 30) : (1ll << 26);
    int64_t cm = std::max<int64_t>(kScoreRowsPad, budget / pl.dpitch / kScoreRowsPad * kScoreRowsPad);
    cm = std::min<int64_t>(cm, 65536);
    if (chunk_override > 0) cm = chunk_override;   // test hook: force many small chunks
    pl.chunk_m = (int)std::min<int64_t>(cm, std::max<int64_t>(pl.m, 1));
    return pl;
}

// Bring chunk [j0, j0+mc) into packed device rows. Returns the device pointer/pitch to use.
static int stage_chunk(spagxe_ctx *ctx, const spagxe_geno *g, const ChunkPlan &pl, int64_t j0, int mc, int buf,
                       const uint8_t **d_rows, int64_t *h2d_bytes) {
    cudaStream_t s = ctx->stream();
    if (geno_needs_prepare(g)) {
        // file-layout rows -> (device) -> gather of the analysis samples / allele order -> packed rows of N samples
        const int64_t nfile = (g->sample_index && g->n_file_samples > 0) ? g->n_file_samples : g->n_samples;
        const int64_t src_row = (nfile + 3) / 4;
        const uint8_t *d_src_rows;
        int64_t spitch;
        if (g->location == SPAGXE_LOC_DEVICE) { d_src_rows = (const uint8_t *)g->data + j0 * g->pitch; spitch = g->pitch; }
        else {
            spitch = round_up(src_row, 16);
            const int64_t need = spitch * pl.chunk_m;
            if (need > ctx->cap_src) { CU(cudaStreamSynchronize(s)); CU(dev_realloc(&ctx->d_src, need)); ctx->cap_src = need; }
            const uint8_t *src = (const uint8_t *)g->data + j0 * g->pitch;
            CU(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_done[buf], 0));
            CU(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_done[buf ^ 1], 0));   // d_src is a single buffer
            int rc = stage_rows(ctx, src, g->pitch, src_row, mc, ctx->d_src, spitch, host_ptr_is_pinned(src));
            if (rc) return rc;
            CU(cudaEventRecord(ctx->ev_copied[buf], ctx->s_copy));
            CU(cudaStreamWaitEvent(s, ctx->ev_copied[buf], 0));
            *h2d_bytes += src_row * mc;
            d_src_rows = ctx->d_src;
        }
        const int64_t pw = pl.dpitch / 4;
        dim3 grid((unsigned)((pw + 255) / 256), (unsigned)mc);
        k_bed_prepare<<<grid, 256, 0, s>>>(d_src_rows, spitch, (uint32_t *)ctx->d_bed[buf], pw, pl.n, mc,
                                           g->sample_index ? ctx->d_sidx : nullptr, (g->flags & SPAGXE_GENO_COUNT_A2) ? 1 : 0);
        LAUNCH_CHECK();
        *d_rows = ctx->d_bed[buf];
        return SPAGXE_OK;
    }
    if (g->kind == SPAGXE_GENO_BED) {
        if (g->location == SPAGXE_LOC_DEVICE) {
            *d_rows = (const uint8_t *)g->data + j0 * g->pitch;
            return SPAGXE_OK;
        }
        const uint8_t *src = (const uint8_t *)g->data + j0 * g->pitch;
        CU(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_done[buf], 0));
        int rc = stage_rows(ctx, src, g->pitch, pl.row_bytes, mc, ctx->d_bed[buf], pl.dpitch, host_ptr_is_pinned(src));
        if (rc) return rc;
        CU(cudaEventRecord(ctx->ev_copied[buf], ctx->s_copy));
        CU(cudaStreamWaitEvent(s, ctx->ev_copied[buf], 0));
        *h2d_bytes += pl.row_bytes * mc;
        *d_rows = ctx->d_bed[buf];
        return SPAGXE_OK;
    }
    // dense R matrix: stage (if on host) then pack into 2-bit rows on the device
    const size_t esz = (g->kind == SPAGXE_GENO_F64) ? 8 : 4;
    const void *dsrc;
    if (g->location == SPAGXE_LOC_DEVICE) dsrc = (const uint8_t *)g->data + (size_t)j0 * g->pitch * esz;
    else {
        const int64_t need = (int64_t)pl.n * pl.chunk_m * (int64_t)esz;
        if (need > ctx->cap_dense) {
            CU(cudaStreamSynchronize(s));
            if (ctx->d_dense) cudaFree(ctx->d_dense);
            ctx->d_dense = nullptr;
            CU(cudaMalloc(&ctx->d_dense, need));
            ctx->cap_dense = need;
        }
        CU(cudaMemcpy2DAsync(ctx->d_dense, pl.n * esz, (const uint8_t *)g->data + (size_t)j0 * g->pitch * esz,
                             g->pitch * esz, pl.n * esz, mc, cudaMemcpyHostToDevice, s));
        *h2d_bytes += (int64_t)pl.n * esz * mc;
        dsrc = ctx->d_dense;
    }
    const int64_t ld = (g->location == SPAGXE_LOC_DEVICE) ? g->pitch : pl.n;
    const int64_t pw = pl.dpitch / 4;
    CU(launch_pack_dense(s, g->kind == SPAGXE_GENO_F64, dsrc, ld, pl.n, mc, (uint32_t *)ctx->d_bed[buf], pw, ctx->d_ctr));
    ctx->launches++;
    *d_rows = ctx->d_bed[buf];
    return SPAGXE_OK;
}

// device copy of the sample gather index of this call (if any)
static int upload_sample_index(spagxe_ctx *ctx, const spagxe_geno *g) {
    if (!g->sample_index) return SPAGXE_OK;
    if (g->n_samples > ctx->cap_sidx) {
        CU(cudaStreamSynchronize(ctx->stream()));
        CU(dev_realloc(&ctx->d_sidx, g->n_samples)); ctx->cap_sidx = g->n_samples;
    }
    CU(cudaMemcpyAsync(ctx->d_sidx, g->sample_index, g->n_samples * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream()));
    return SPAGXE_OK;
}

static SnpStats stats_view(spagxe_ctx *ctx, int cap) {
    SnpStats st;
    st.D0 = ctx->d_stats; st.D1 = ctx->d_stats + cap; st.T0 = ctx->d_stats + 2 * (size_t)cap; st.T1 = ctx->d_stats + 3 * (size_t)cap;
    st.asum = ctx->d_istats; st.nmiss = ctx->d_istats + cap;
    st.H0 = ctx->d_stats + 4 * (size_t)cap; st.H1 = ctx->d_stats + 5 * (size_t)cap; st.nhom = ctx->d_istats + 2 * (size_t)cap;
    return st;
}

static int launch_score(spagxe_ctx *ctx, const uint8_t *d_rows, const ChunkPlan &pl, int mc, SnpStats st, int g_model = 0) {
    cudaStream_t s = ctx->stream();
    const ScoreGeom sg{pl.n, pl.row_bytes, pl.dpitch, mc};
    if (!ctx->use_tc) {   // fp64 cross-check kernel: its partial-sum scratch is only allocated when it is used
        const int64_t need = score_v1_partial_elems(sg, mc);
        if (need > ctx->cap_partial) {
            CU(cudaStreamSynchronize(s));
            CU(dev_realloc(&ctx->d_partial, need)); CU(dev_realloc(&ctx->d_ipartial, need / 2));
            ctx->cap_partial = need;
        }
        if (g_model != 0) return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "G.model Dom / Rec needs the tcgen05 score kernel (SPAGXE_OPT_SCORE_IMPL = 0)");
        CU(score_launch_v1(s, d_rows, sg, ctx->d_R, ctx->d_V, ctx->d_partial, ctx->d_ipartial, st, &ctx->launches));
        return SPAGXE_OK;
    }
    const int m_pad = score_acc_rows(mc);
    if (m_pad > ctx->cap_acc) {
        CU(cudaStreamSynchronize(s));
        CU(dev_realloc(&ctx->d_acc, (size_t)m_pad * 32));
        ctx->cap_acc = m_pad;
    }
    std::string serr;
    cudaError_t e = score_launch_tc(s, ctx->num_sms, ctx->tc_cfg, d_rows, sg, ctx->d_bmat, ctx->d_binfo, ctx->d_acc, st, &ctx->launches, &serr);
    if (e != cudaSuccess) return fail(ctx, SPAGXE_ERR_CUDA, "%s: %s", serr.c_str(), cudaGetErrorString(e));
    if (g_model != 0) {   // hom-A1 class sums: the same kernel with a 0/1 LUT
        e = score_launch_tc(s, ctx->num_sms, 0, d_rows, sg, ctx->d_bmat, ctx->d_binfo, ctx->d_acc, st, &ctx->launches, &serr, true);
        if (e != cudaSuccess) return fail(ctx, SPAGXE_ERR_CUDA, "%s (hom pass): %s", serr.c_str(), cudaGetErrorString(e));
    }
    return SPAGXE_OK;
}

struct EventPool {   // view of the context's reusable timing events for one call
    spagxe_ctx *ctx;
    explicit EventPool(spagxe_ctx *c) : ctx(c) { c->ev_used = 0; }
    cudaEvent_t rec(cudaStream_t s) {
        if (ctx->ev_used == ctx->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); ctx->ev_pool.push_back(e); }
        cudaEvent_t e = ctx->ev_pool[ctx->ev_used++];
        cudaEventRecord(e, s);
        return e;
    }
};

static void fill_diag(spagxe_diag *diag, const Counters &c, int64_t m) {
    diag->n_snps = m; diag->n_filtered = c.n_filtered; diag->n_branch_a = c.n_branch_a; diag->n_branch_b = c.n_branch_b;
    diag->n_spa = c.n_spa; diag->n_wald_fail = c.n_wald_fail; diag->n_cct_stop = c.n_cct_stop;
    diag->n_singular = c.n_singular; diag->n_mix_logistic = c.n_mix_logistic; diag->newton_iters = c.newton_iters;
}

static const spagxe_params kDefaultParams = {0.001, 2.0, 0.001, 0.15, 0.05, 0.1, 0, 0, 0, 0, 0};

enum RunMode { RUN_CCT = 0, RUN_PLUS = 1, RUN_MIX = 2, RUN_MIXPLUS = 3 };

static int launch_mix(spagxe_ctx *ctx, const uint8_t *d_rows, int64_t dpitch, SnpStats st, int64_t j0, int mc, int64_t M,
                      int64_t N, WaldData wd, const spagxe_params *prm, double *d_out, bool plus) {
    cudaStream_t s = ctx->stream();
    k_finalize_mix<<<(mc + 127) / 128, 128, 0, s>>>(st, mc, j0, M, N, prm->min_maf, d_out, ctx->d_b_list, ctx->d_ctr);
    LAUNCH_CHECK();
    if (!ctx->opt_mix_per_snp && !plus) {   // test hook: every SNP through the cluster kernel (SPAGxEmix_CCT_Plus: always)
        const MixBatchParams bp{prm->epsilon, prm->cutoff, prm->neg_ratio_cutoff, prm->g_model};
        if (!ctx->mix_scratch) ctx->mix_scratch = mix_scratch_create();
        CU(mix_batch_run(s, ctx->mix_scratch, ctx->num_sms, d_rows, dpitch, st, j0, mc, M, N, ctx->d_R, ctx->d_E, ctx->d_vc, ctx->d_pcs,
                         ctx->d_xtx_inv, ctx->n_pcs, bp, d_out, ctx->d_b_list, ctx->d_ctr, &ctx->launches));
    }
    cudaLaunchConfig_t cfg = {};
    const int mcs = ctx->opt_mix_cluster;      // CTAs cooperating on one SNP of the per-SNP route
    const int n_clusters = std::max(1, ctx->num_sms / mcs);
    cfg.gridDim = dim3((unsigned)(n_clusters * mcs));
    cfg.blockDim = dim3(kBThreads);
    cfg.dynamicSmemBytes = branch_b_smem(kMaxQ);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)mcs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const int *list = ctx->d_b_list;
    Counters *ctr = ctx->d_ctr;
    const double *R = ctx->d_R, *E = ctx->d_E;
    const VecConst *vc = ctx->d_vc;
    MixData md{ctx->d_pcs, ctx->d_xtx_inv, ctx->n_pcs, prm->pc_pvalue_cutoff, prm->neg_ratio_cutoff, prm->g_model, ctx->d_bitems};
    {   // per-cluster cache of the individual allele frequencies m_i (N doubles each)
        const int64_t need = (int64_t)n_clusters * N;
        if (need > ctx->cap_mcache) { CU(cudaStreamSynchronize(s)); CU(dev_realloc(&ctx->d_mcache, need)); ctx->cap_mcache = need; }
    }
    double *mcache = ctx->d_mcache;
    const double *hbeta = nullptr; int hbeta_ld = 0; const MixHand *hand = nullptr;
    if (plus) {
        GrmCoo grm{ctx->d_gi, ctx->d_gj, ctx->d_gv, (long long)ctx->grm_nnz};
        CU(cudaLaunchKernelEx(&cfg, k_mixplus_snp, list, ctr, d_rows, dpitch, st, j0, M, N, R, E, vc, md, grm, prm->epsilon, prm->cutoff, d_out, mcache));
        ctx->launches++;
        return SPAGXE_OK;
    }
    if (!ctx->opt_mix_per_snp) { hbeta = mix_scratch_beta(ctx->mix_scratch, &hbeta_ld); hand = mix_scratch_hand(ctx->mix_scratch); }
    CU(cudaLaunchKernelEx(&cfg, k_mix_snp, list, ctr, d_rows, dpitch, st, j0, M, N, R, E, vc, wd, md, prm->epsilon, prm->cutoff, d_out, mcache,
                          hbeta, hbeta_ld, hand));
    ctx->launches++;
    // traits survival / categorical: the Wald test + CCT of the chunk's branch-B SNPs (k_mix_snp appended them to d_bitems)
    if (wd.trait == SPAGXE_TRAIT_SURVIVAL) {
        cudaError_t e = cox_launch(ctx->cox, s, ctx->num_sms, ctx->d_bitems, &ctx->d_ctr->spax_count, M, prm->g_model, d_out, ctx->d_ctr, &ctx->launches, true);
        if (e != cudaSuccess) return fail(ctx, SPAGXE_ERR_CUDA, "cox_launch: %s", cudaGetErrorString(e));
    } else if (wd.trait == SPAGXE_TRAIT_CATEGORICAL) {
        cudaError_t e = clm_launch(ctx->clm, s, ctx->num_sms, ctx->d_bitems, &ctx->d_ctr->spax_count, M, prm->g_model, d_out, ctx->d_ctr, &ctx->launches, true);
        if (e != cudaSuccess) return fail(ctx, SPAGXE_ERR_CUDA, "clm_launch: %s", cudaGetErrorString(e));
    }
    return SPAGXE_OK;
}

// SPAGxE_CCT on a real-valued matrix (imputed dosages): the 2-bit pipeline refuses values outside {0, 1, 2, NA}, so a
// double matrix that holds such values is tested by k_cct_dosage_snp instead (one cluster per SNP, g read as doubles).
// Binary and quantitative traits.
// kind 0: SPAGxE_CCT; 1: SPAGxEmix_CCT; 2: SPAGxEmix_CCT_Plus (k_mix_dosage_snp / k_mixplus_dosage_snp: the per-SNP SPAGxEmix code on a
// dense column, M x 12)
static int run_cct_dosage(spagxe_ctx *ctx, const spagxe_geno *g, const spagxe_params *prm, double *out, spagxe_diag *diag, int kind = 0) {
    if (g->kind != SPAGXE_GENO_F64) return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "dosage input must be a double matrix");
    if (kind < 2 && ctx->wald_trait != SPAGXE_TRAIT_BINARY && ctx->wald_trait != SPAGXE_TRAIT_QUANTITATIVE)
        return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "Geno.mtx holds non-integer dosages: the survival / categorical Wald tests need hard calls");
    const int ncol = kind == 0 ? 10 : (kind == 3 ? 8 : 12);   // kind 3: SPAGxE_Plus (k_plus_dosage_snp)
    if (prm->out_location == SPAGXE_LOC_DEVICE) return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "dosage input: the result matrix lives on the host");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream();
    int rc = ensure_vec_mode(ctx, kind == 0 ? 0 : (kind == 3 ? 1 : 2), kind == 3 ? ctx->d_Rn_user : nullptr);
    if (rc) return rc;
    const int64_t n = g->n_samples, m = g->n_snps, ld = g->pitch;
    spagxe_diag dg; memset(&dg, 0, sizeof dg);
    ctx->launches = 0;
    CU(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(Counters), s));
    EventPool evp(ctx);
    cudaEvent_t ev0 = evp.rec(s);
    const bool on_dev = (g->location == SPAGXE_LOC_DEVICE);
    const int64_t mc_max = std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(m, 1), (256ll << 20) / (n * 8)));
    const int64_t need_dense = mc_max * n * 8, need_out = std::max<int64_t>(m, 1) * ncol;
    if (!on_dev && need_dense > ctx->cap_dense) { CU(cudaStreamSynchronize(s)); cudaFree(ctx->d_dense); ctx->d_dense = nullptr; CU(cudaMalloc(&ctx->d_dense, need_dense)); ctx->cap_dense = need_dense; }
    if (need_out > ctx->cap_out) { CU(cudaStreamSynchronize(s)); CU(dev_realloc(&ctx->d_out, need_out)); ctx->cap_out = need_out; }
    cudaLaunchConfig_t cfg = {};
    const int mcs = 8;
    const int n_clusters = std::max(1, ctx->num_sms / mcs);
    cfg.gridDim = dim3((unsigned)(n_clusters * mcs));
    cfg.blockDim = dim3(kBThreads);
    cfg.dynamicSmemBytes = branch_b_smem(kMaxQ);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)mcs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const double *G = (const double *)g->data;
    if (kind == 1 || kind == 2) {   // per-cluster cache of the individual allele frequencies m_i
        const int64_t need = (int64_t)n_clusters * n;
        if (need > ctx->cap_mcache) { CU(cudaStreamSynchronize(s)); CU(dev_realloc(&ctx->d_mcache, need)); ctx->cap_mcache = need; }
    }
    for (int64_t j0 = 0; j0 < m; j0 += mc_max) {
        const int mc = (int)std::min<int64_t>(mc_max, m - j0);
        if (kind == 3) {
            PlusDosageArgs a;
            if (on_dev) { a.G = G + j0 * ld; a.ld = ld; }
            else {
                CU(cudaMemcpy2DAsync(ctx->d_dense, n * 8, G + j0 * ld, ld * 8, n * 8, mc, cudaMemcpyHostToDevice, s));
                dg.h2d_bytes += n * 8 * mc;
                a.G = (const double *)ctx->d_dense; a.ld = n;
            }
            a.m = mc; a.snp0 = j0; a.m_total = m; a.n = n;
            a.R = ctx->d_R; a.E = ctx->d_E; a.Rn = ctx->d_Rn; a.vc = ctx->d_vc;
            a.pc = PlusConst{ctx->R_GRM_R, ctx->R_GRM_RE, ctx->Rnew_GRM_Rnew};
            a.grm = GrmCoo{ctx->d_gi, ctx->d_gj, ctx->d_gv, (long long)ctx->grm_nnz};
            a.epsilon = prm->epsilon; a.cutoff = prm->cutoff; a.min_maf = prm->min_maf; a.g_model = prm->g_model;
            a.out = ctx->d_out; a.ctr = ctx->d_ctr;
            cudaLaunchConfig_t cfg0 = cfg; cfg0.dynamicSmemBytes = 0;   // no Wald tile in this kernel
            CU(cudaLaunchKernelEx(&cfg0, k_plus_dosage_snp, a));
            ctx->launches++;
            continue;
        }
        if (kind != 0) {
            MixDosageArgs a;
            if (on_dev) { a.G = G + j0 * ld; a.ld = ld; }
            else {
                CU(cudaMemcpy2DAsync(ctx->d_dense, n * 8, G + j0 * ld, ld * 8, n * 8, mc, cudaMemcpyHostToDevice, s));
                dg.h2d_bytes += n * 8 * mc;
                a.G = (const double *)ctx->d_dense; a.ld = n;
            }
            a.m = mc; a.snp0 = j0; a.m_total = m; a.n = n;
            a.R = ctx->d_R; a.E = ctx->d_E; a.vc = ctx->d_vc;
            a.wd = WaldData{ctx->d_Y, ctx->d_cova, ctx->wald_p, kind == 2 ? 0 : ctx->wald_trait};
            a.md = MixData{ctx->d_pcs, ctx->d_xtx_inv, ctx->n_pcs, prm->pc_pvalue_cutoff, prm->neg_ratio_cutoff, prm->g_model, nullptr};
            a.grm = GrmCoo{ctx->d_gi, ctx->d_gj, ctx->d_gv, kind == 2 ? (long long)ctx->grm_nnz : 0};
            a.epsilon = prm->epsilon; a.cutoff = prm->cutoff; a.min_maf = prm->min_maf;
            a.out = ctx->d_out; a.ctr = ctx->d_ctr; a.mcache_all = ctx->d_mcache;
            if (kind == 2) CU(cudaLaunchKernelEx(&cfg, k_mixplus_dosage_snp, a));
            else CU(cudaLaunchKernelEx(&cfg, k_mix_dosage_snp, a));
            ctx->launches++;
            continue;
        }
        DosageArgs a;
        if (on_dev) { a.G = G + j0 * ld; a.ld = ld; }
        else {
            CU(cudaMemcpy2DAsync(ctx->d_dense, n * 8, G + j0 * ld, ld * 8, n * 8, mc, cudaMemcpyHostToDevice, s));
            dg.h2d_bytes += n * 8 * mc;
            a.G = (const double *)ctx->d_dense; a.ld = n;
        }
        a.m = mc; a.snp0 = j0; a.m_total = m; a.n = n;
        a.R = ctx->d_R; a.E = ctx->d_E; a.Rn = ctx->d_Rn; a.vc = ctx->d_vc;
        a.wd = WaldData{ctx->d_Y, ctx->d_cova, ctx->wald_p, ctx->wald_trait};
        a.epsilon = prm->epsilon; a.min_maf = prm->min_maf; a.g_model = prm->g_model;
        a.out = ctx->d_out; a.ctr = ctx->d_ctr;
        CU(cudaLaunchKernelEx(&cfg, k_cct_dosage_snp, a));
        ctx->launches++;
    }
    cudaEvent_t ev1 = evp.rec(s);
    Counters hc;
    CU(cudaMemcpyAsync(&hc, ctx->d_ctr, sizeof hc, cudaMemcpyDeviceToHost, s));
    if (m > 0) {
        const int64_t old = prm->out_ld > 0 ? prm->out_ld : m;
        CU(cudaMemcpy2DAsync(out, old * 8, ctx->d_out, m * 8, m * 8, ncol, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    float ms = 0; cudaEventElapsedTime(&ms, ev0, ev1);
    dg.ms_total = ms; dg.ms_tests = ms; dg.d2h_bytes = m * ncol * 8;
    fill_diag(&dg, hc, m);
    dg.kernel_launches = ctx->launches;
    if (diag) *diag = dg;
    if (hc.n_allmissing) return fail(ctx, SPAGXE_ERR_ARG, "%llu SNP(s) have every genotype missing (the reference stops here)", hc.n_allmissing);
    if (hc.n_singular) return fail(ctx, SPAGXE_ERR_SINGULAR, "solve(t(W)%%*%%W) is singular for %llu SNP(s)", hc.n_singular);
    if (hc.n_cct_stop) return fail(ctx, SPAGXE_ERR_CCT_STOP, "CCT() would stop() for %llu SNP(s) (NA or out-of-range p-value)", hc.n_cct_stop);
    return SPAGXE_OK;
}

static int run_generic(spagxe_ctx *ctx, const spagxe_geno *g, const spagxe_params *prm, double *out, spagxe_diag *diag,
                       RunMode mode) {
    if (!ctx) return SPAGXE_ERR_ARG;
    if (!prm) prm = &kDefaultParams;
    int rc = validate_geno(ctx, g);
    if (rc) return rc;
    const bool mixplus = (mode == RUN_MIXPLUS);
    if (mixplus) mode = RUN_MIX;   // the same pipeline; the per-SNP kernel differs (launch_mix)
    const char *name = mode == RUN_CCT ? "run_cct" : mode == RUN_PLUS ? "run_plus" : mixplus ? "run_mix_plus" : "run_mix";
    const int ncol = mode == RUN_CCT ? 10 : mode == RUN_PLUS ? 8 : 12;
    if (!out) return fail(ctx, SPAGXE_ERR_ARG, "out is NULL");
    if (ctx->n <= 0) return fail(ctx, SPAGXE_ERR_STATE, "%s: call spagxe_set_vectors first", name);
    if (g->n_samples != ctx->n) return fail(ctx, SPAGXE_ERR_ARG, "geno has %lld samples but R has %lld",
                                            (long long)g->n_samples, (long long)ctx->n);
    if (mode != RUN_PLUS && !mixplus && ctx->wald_trait == 0)
        return fail(ctx, SPAGXE_ERR_STATE, "%s: call spagxe_set_wald first (Wald test data)", name);
    if (mode == RUN_PLUS && !ctx->have_grm) return fail(ctx, SPAGXE_ERR_STATE, "run_plus: call spagxe_set_grm first");
    if (mode == RUN_MIX && ctx->n_pcs <= 0) return fail(ctx, SPAGXE_ERR_STATE, "%s: call spagxe_set_pcs first", name);
    if (mixplus && !ctx->have_coo) return fail(ctx, SPAGXE_ERR_STATE, "run_mix_plus: call spagxe_set_sparse_grm first");
    const bool wald_dev = (mode == RUN_CCT) || (mode == RUN_MIX && !mixplus);   // loops whose branch B runs a Wald kernel
    if (wald_dev && ctx->wald_trait == SPAGXE_TRAIT_CATEGORICAL && !clm_ready(ctx->clm)) {
        CU(cudaSetDevice(ctx->device));
        const char *what = "clm_setup";
        cudaError_t ce = clm_setup(ctx->clm, ctx->stream(), ctx->n, ctx->h_ycat.data(), ctx->d_E, ctx->d_cova, ctx->wald_p, &what);
        if (ce != cudaSuccess) return fail(ctx, SPAGXE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(ce));
    }
    if (wald_dev && ctx->wald_trait == SPAGXE_TRAIT_SURVIVAL && !cox_ready(ctx->cox)) {
        CU(cudaSetDevice(ctx->device));
        const char *what = "cox_setup";
        cudaError_t ce = cox_setup(ctx->cox, ctx->stream(), ctx->n, ctx->h_time.data(), ctx->h_event.data(), ctx->d_E, ctx->d_cova,
                                   ctx->wald_p, &what);
        if (ce != cudaSuccess) return fail(ctx, SPAGXE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(ce));
    }
    if (prm->g_model < 0 || prm->g_model > 2) return fail(ctx, SPAGXE_ERR_ARG, "g_model must be 0 (Add), 1 (Dom) or 2 (Rec)");
    if (prm->impute_method != 0) return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "only impute.method=\"fixed\" exists");
    if (prm->out_ld != 0 && (prm->out_ld < g->n_snps || prm->out_location == SPAGXE_LOC_DEVICE))
        return fail(ctx, SPAGXE_ERR_ARG, "out_ld must be 0 or >= M, and only applies to a host result matrix");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream();
    const int64_t M = g->n_snps, N = g->n_samples;
    spagxe_diag dg; memset(&dg, 0, sizeof dg);
    ctx->launches = 0;
    if (M == 0) { if (diag) *diag = dg; return SPAGXE_OK; }
    ChunkPlan pl = make_plan(g, ctx->opt_chunk_snps);
    const bool own_bed = !(g->kind == SPAGXE_GENO_BED && g->location == SPAGXE_LOC_DEVICE) || geno_needs_prepare(g);
    double *d_out = nullptr;
    if (prm->out_location == SPAGXE_LOC_DEVICE) { d_out = out; rc = ensure_chunk_scratch(ctx, pl, own_bed, 0); }
    else { rc = ensure_chunk_scratch(ctx, pl, own_bed, M * ncol); d_out = ctx->d_out; }
    if (rc) return rc;
    if ((rc = upload_sample_index(ctx, g))) return rc;
    // CCT contracts with (R, R.new); plus and mix with (R, E*R); plus takes R.new from the null-model object
    const int vmode = (mode == RUN_CCT) ? 0 : (mode == RUN_PLUS ? 1 : 2);
    {
        const int64_t need = own_bed ? 2ll * pl.chunk_m : M;
        if (need > ctx->cap_bitems) { CU(dev_realloc(&ctx->d_bitems, need)); ctx->cap_bitems = need; }
    }
    if ((rc = ensure_vec_mode(ctx, vmode, mode == RUN_PLUS ? ctx->d_Rn_user : nullptr))) return rc;
    CU(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(Counters), s));
    EventPool evp(ctx);
    cudaEvent_t e_begin = evp.rec(s);
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> score_ev, test_ev, spa_ev, bb_ev;
    SnpStats st = stats_view(ctx, ctx->cap_m);
    WaldData wd{ctx->d_Y, ctx->d_cova, ctx->wald_p, ctx->wald_trait};
    PlusConst pc{ctx->R_GRM_R, ctx->R_GRM_RE, ctx->Rnew_GRM_Rnew};
    int buf = 0;
    int64_t spa_pending = 0;   // SNPs whose saddlepoint items are waiting in the work list
    for (int64_t j0 = 0; j0 < M; j0 += pl.chunk_m, buf ^= 1) {
        const int mc = (int)std::min<int64_t>(pl.chunk_m, M - j0);
        const uint8_t *d_rows = nullptr;
        if ((rc = stage_chunk(ctx, g, pl, j0, mc, buf, &d_rows, &dg.h2d_bytes))) return rc;
        // SPAGxEmix: per-chunk work list and queue cursor.  The other loops collect their saddlepoint items over the chunks
        // (flush_spa) and their branch-B items until a flush.
        if (mode == RUN_MIX) CU(cudaMemsetAsync(&ctx->d_ctr->spa_count, 0, 4 * sizeof(int), s));
        cudaEvent_t a = evp.rec(s);
        if ((rc = launch_score(ctx, d_rows, pl, mc, st, prm->g_model))) return rc;
        cudaEvent_t b = evp.rec(s);
        score_ev.push_back({a, b});
        const bool last = (j0 + pl.chunk_m >= M);
        // one saddlepoint launch for the items of many chunks: the slowest SNP of a launch (a dozen Newton iterations where
        // the mean is two) sets its duration, so eight launches per 65,536 SNPs cost 0.6 ms and one costs a quarter of that
        auto flush_spa = [&](const double *rv) -> int {
            if (spa_pending == 0) return SPAGXE_OK;
            int r2 = launch_spa(ctx, rv, N, M, d_out);
            if (r2) return r2;
            CU(cudaMemsetAsync(&ctx->d_ctr->spa_count, 0, 2 * sizeof(int), s));
            spa_pending = 0;
            return SPAGXE_OK;
        };
        if (mode == RUN_CCT) {
            if (spa_pending + mc > ctx->cap_spa && (rc = flush_spa(ctx->d_V))) return rc;
            k_finalize_cct<<<(mc + 127) / 128, 128, 0, s>>>(st, mc, j0, M, N, ctx->d_vc, prm->epsilon, prm->min_maf, d_out,
                                                            ctx->d_spa_list, ctx->d_bitems, d_rows, pl.dpitch, ctx->d_ctr, prm->g_model);
            LAUNCH_CHECK();
            spa_pending += mc;
            if (last && (rc = flush_spa(ctx->d_V))) return rc;
        } else if (mode == RUN_PLUS) {
            if (spa_pending + mc > ctx->cap_spa && (rc = flush_spa(ctx->d_Rn))) return rc;
            k_finalize_plus<<<(mc + 127) / 128, 128, 0, s>>>(st, mc, j0, M, N, ctx->d_vc, pc, prm->epsilon, prm->cutoff,
                                                             prm->min_maf, d_out, ctx->d_spa_list, ctx->d_bitems, d_rows, pl.dpitch,
                                                             ctx->d_ctr, prm->g_model);
            LAUNCH_CHECK();
            spa_pending += mc;
            if (last && (rc = flush_spa(ctx->d_Rn))) return rc;
        } else {
            if ((rc = launch_mix(ctx, d_rows, pl.dpitch, st, j0, mc, M, N, wd, prm, d_out, mixplus))) return rc;
        }
        // staged rows live in two alternating buffers: pending branch-B items must run before their buffer is reused;
        // caller-owned device rows stay valid for the whole call, so everything is deferred to one launch at the end
        bool flushed = false;
        cudaEvent_t c = evp.rec(s);
        spa_ev.push_back({b, c});
        if (mode != RUN_MIX && (last || (own_bed && buf == 1))) {
            if ((rc = flush_branch_b(ctx, mode == RUN_CCT ? 0 : 1, M, N, wd, prm->min_maf, prm->cutoff, d_out, prm->g_model))) return rc;
            flushed = true;
        }
        cudaEvent_t d = evp.rec(s);
        test_ev.push_back({b, d});
        if (flushed) bb_ev.push_back({c, d});
        if (own_bed) {
            CU(cudaEventRecord(ctx->ev_done[buf], s));
            // the flush also read the OTHER staging buffer (deferred items of the previous chunk): it may only be
            // refilled after the cluster kernel, so move its release point behind the flush as well
            if (flushed) CU(cudaEventRecord(ctx->ev_done[buf ^ 1], s));
        }
    }
    Counters hc;
    CU(cudaMemcpyAsync(&hc, ctx->d_ctr, sizeof hc, cudaMemcpyDeviceToHost, s));
    int h_nodes = 0;
    if (ctx->quad_on && mode != RUN_MIX) CU(cudaMemcpyAsync(&h_nodes, ctx->quad.count, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (prm->out_location != SPAGXE_LOC_DEVICE) {
        const int64_t ld = prm->out_ld ? prm->out_ld : M;
        CU(cudaMemcpy2DAsync(out, ld * sizeof(double), d_out, M * sizeof(double), M * sizeof(double), ncol,
                             cudaMemcpyDeviceToHost, s));
        dg.d2h_bytes += M * ncol * sizeof(double);
    }
    cudaEvent_t e_end = evp.rec(s);
    CU(cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, e_begin, e_end); dg.ms_total = ms;
    for (auto &p : score_ev) { cudaEventElapsedTime(&ms, p.first, p.second); dg.ms_score += ms; }
    for (auto &p : test_ev) { cudaEventElapsedTime(&ms, p.first, p.second); dg.ms_tests += ms; }
    for (auto &p : spa_ev) { cudaEventElapsedTime(&ms, p.first, p.second); dg.ms_spa += ms; }
    for (auto &p : bb_ev) { cudaEventElapsedTime(&ms, p.first, p.second); dg.ms_branch_b += ms; }
    dg.spa_nodes = h_nodes;
    if (mode == RUN_CCT) {
        BbFarmWork w;
        CU(bb_farm_read_work(ctx->farm, s, &w, true));
        dg.bb_irls_sample_passes = (int64_t)w.irls_sample_passes; dg.bb_tail_sample_passes = (int64_t)w.tail_sample_passes;
        dg.bb_prep_sample_passes = (int64_t)w.prep_sample_passes; dg.bb_iterations = (int64_t)w.iterations;
    }
    fill_diag(&dg, hc, M);
    dg.cox_evals = (int64_t)hc.cox_evals;
    dg.clm_evals = (int64_t)hc.clm_evals;
    dg.kernel_launches = ctx->launches;
    dg.score_bytes = M * pl.row_bytes;
    if (diag) *diag = dg;
    if (hc.n_nonint) {
        // imputed dosages: SPAGxE_CCT has a dense per-SNP path for them; the other loops rest on the 2-bit coding
        if (mode == RUN_CCT && g->kind == SPAGXE_GENO_F64) return run_cct_dosage(ctx, g, prm, out, diag);
        if (mode == RUN_MIX && g->kind == SPAGXE_GENO_F64) return run_cct_dosage(ctx, g, prm, out, diag, mixplus ? 2 : 1);
        if (mode == RUN_PLUS && g->kind == SPAGXE_GENO_F64) return run_cct_dosage(ctx, g, prm, out, diag, 3);
        return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "Geno.mtx holds values outside {0,1,2,NA}: imputed dosages need a double matrix");
    }
    if (hc.n_allmissing) return fail(ctx, SPAGXE_ERR_ARG, "%llu SNP(s) have every genotype missing (the reference stops here)",
                                     hc.n_allmissing);
    if (hc.n_singular) return fail(ctx, SPAGXE_ERR_SINGULAR, "solve(t(W)%%*%%W) is singular for %llu SNP(s)", hc.n_singular);
    if (hc.n_cct_stop) return fail(ctx, SPAGXE_ERR_CCT_STOP, "CCT() would stop() for %llu SNP(s) (NA or out-of-range p-value)",
                                   hc.n_cct_stop);
    return SPAGXE_OK;
}

extern "C" int spagxe_run_cct(spagxe_ctx *ctx, const spagxe_geno *g, const spagxe_params *prm, double *out, spagxe_diag *diag) {
    return run_generic(ctx, g, prm, out, diag, RUN_CCT);
}
extern "C" int spagxe_run_plus(spagxe_ctx *ctx, const spagxe_geno *g, const spagxe_params *prm, double *out, spagxe_diag *diag) {
    return run_generic(ctx, g, prm, out, diag, RUN_PLUS);
}
extern "C" int spagxe_run_mix(spagxe_ctx *ctx, const spagxe_geno *g, const spagxe_params *prm, double *out, spagxe_diag *diag) {
    return run_generic(ctx, g, prm, out, diag, RUN_MIX);
}
extern "C" int spagxe_run_mix_plus(spagxe_ctx *ctx, const spagxe_geno *g, const spagxe_params *prm, double *out, spagxe_diag *diag) {
    return run_generic(ctx, g, prm, out, diag, RUN_MIXPLUS);
}

// ---------------------------------------------------------------------------
// SPAGxEmixCCT_localance (ref R/SPAGxECCT.R:1401-1647).  The reference takes R matrices only (Geno.mtx, haplo.mtx and
// the matrices of Cova.haplo.mtx.list, all N x M): they are copied in column blocks, packed to 2 bits on the device
// (values outside {0, 1, 2} are refused) and tested by one cluster per SNP (k_localance_snp).
extern "C" int spagxe_run_localance(spagxe_ctx *ctx, int64_t n, int64_t m, const double *G, const double *H, int n_cov_haplo,
                                    const double *const *cov_haplo, int64_t ld, const spagxe_params *prm, double *out,
                                    spagxe_diag *diag) {
    if (!ctx) return SPAGXE_ERR_ARG;
    if (!prm) prm = &kDefaultParams;
    if (!G || !H || !out || m < 0 || n <= 0) return fail(ctx, SPAGXE_ERR_ARG, "run_localance: bad arguments");
    if (ctx->n <= 0) return fail(ctx, SPAGXE_ERR_STATE, "run_localance: call spagxe_set_vectors first");
    if (n != ctx->n) return fail(ctx, SPAGXE_ERR_ARG, "run_localance: matrices have %lld rows but R has %lld", (long long)n, (long long)ctx->n);
    if (ld < n) return fail(ctx, SPAGXE_ERR_ARG, "run_localance: leading dimension < N");
    if (n_cov_haplo < 0 || n_cov_haplo > kMaxLocalCov || (n_cov_haplo > 0 && !cov_haplo))
        return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "run_localance: at most %d matrices in Cova.haplo.mtx.list", kMaxLocalCov);
    if (ctx->wald_trait != SPAGXE_TRAIT_BINARY && ctx->wald_trait != SPAGXE_TRAIT_QUANTITATIVE)
        return fail(ctx, ctx->wald_trait == 0 ? SPAGXE_ERR_STATE : SPAGXE_ERR_UNSUPPORTED,
                    "run_localance: call spagxe_set_wald first (binary or quantitative trait; the survival / categorical Wald tests "
                    "are implemented for SPAGxE_CCT only)");
    const int q_full = 1 + n_cov_haplo + ctx->wald_p + (ctx->wald_trait == SPAGXE_TRAIT_QUANTITATIVE ? 1 : 0) + 3;
    if (q_full > kMaxQ) return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "run_localance: the Wald design has %d columns (at most %d)", q_full, kMaxQ);
    if (prm->g_model < 0 || prm->g_model > 2) return fail(ctx, SPAGXE_ERR_ARG, "g_model must be 0 (Add), 1 (Dom) or 2 (Rec)");
    if (prm->impute_method != 0) return fail(ctx, SPAGXE_ERR_UNSUPPORTED, "only impute.method=\"fixed\" exists");
    if (prm->out_location == SPAGXE_LOC_DEVICE) return fail(ctx, SPAGXE_ERR_ARG, "run_localance: the result matrix lives on the host");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream();
    int rc = ensure_vec_mode(ctx, 1, nullptr);   // R, E and the constants of R on the device (V = E R is not used here)
    if (rc) return rc;
    spagxe_diag dg; memset(&dg, 0, sizeof dg);
    CU(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(Counters), s));
    EventPool evp(ctx);
    cudaEvent_t ev0 = evp.rec(s);
    const int nmat = 2 + n_cov_haplo;
    const int64_t row_bytes = (n + 3) / 4, pitch = round_up(row_bytes, 16);
    // column blocks: the dense staging buffer holds `mc` columns of one matrix (<= 256 MB), the packed rows of all matrices follow
    const int64_t mc_max = std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(m, 1), (256ll << 20) / (n * 8)));
    const int64_t need_dense = mc_max * n * 8, need_rows = (int64_t)nmat * mc_max * pitch, need_out = std::max<int64_t>(m, 1) * 10;
    if (need_dense > ctx->cap_dense) { CU(cudaStreamSynchronize(s)); cudaFree(ctx->d_dense); ctx->d_dense = nullptr; CU(cudaMalloc(&ctx->d_dense, need_dense)); ctx->cap_dense = need_dense; }
    if (need_rows > ctx->cap_local) { CU(cudaStreamSynchronize(s)); CU(dev_realloc(&ctx->d_local, need_rows)); ctx->cap_local = need_rows; }
    if (need_out > ctx->cap_out) { CU(cudaStreamSynchronize(s)); CU(dev_realloc(&ctx->d_out, need_out)); ctx->cap_out = need_out; }
    const double *mats[2 + kMaxLocalCov] = {G, H};
    for (int l = 0; l < n_cov_haplo; l++) {
        if (!cov_haplo[l]) return fail(ctx, SPAGXE_ERR_ARG, "run_localance: Cova.haplo.mtx.list[%d] is NULL", l);
        mats[2 + l] = cov_haplo[l];
    }
    cudaLaunchConfig_t cfg = {};
    const int mcs = 8;
    const int n_clusters = std::max(1, ctx->num_sms / mcs);
    cfg.gridDim = dim3((unsigned)(n_clusters * mcs));
    cfg.blockDim = dim3(kBThreads);
    cfg.dynamicSmemBytes = branch_b_smem(kMaxQ);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)mcs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int64_t j0 = 0; j0 < m; j0 += mc_max) {
        const int mc = (int)std::min<int64_t>(mc_max, m - j0);
        for (int k = 0; k < nmat; k++) {
            CU(cudaMemcpy2DAsync(ctx->d_dense, n * 8, mats[k] + j0 * ld, ld * 8, n * 8, mc, cudaMemcpyHostToDevice, s));
            dg.h2d_bytes += n * 8 * mc;
            CU(launch_pack_dense(s, 1, ctx->d_dense, n, n, mc, (uint32_t *)(ctx->d_local + (int64_t)k * mc_max * pitch), pitch / 4, ctx->d_ctr));
            ctx->launches++;
        }
        LocalArgs a;
        a.g_rows = ctx->d_local; a.h_rows = ctx->d_local + mc_max * pitch; a.n_ch = n_cov_haplo; a.pitch = pitch;
        for (int l = 0; l < kMaxLocalCov; l++) a.ch_rows[l] = (l < n_cov_haplo) ? ctx->d_local + (int64_t)(2 + l) * mc_max * pitch : nullptr;
        a.m = mc; a.snp0 = j0; a.m_total = m; a.n = n;
        a.R = ctx->d_R; a.E = ctx->d_E; a.vc = ctx->d_vc; a.wd = WaldData{ctx->d_Y, ctx->d_cova, ctx->wald_p, ctx->wald_trait};
        a.epsilon = prm->epsilon; a.min_maf = prm->min_maf; a.g_model = prm->g_model;
        a.out = ctx->d_out; a.ctr = ctx->d_ctr;
        CU(cudaLaunchKernelEx(&cfg, k_localance_snp, a));
        ctx->launches++;
    }
    cudaEvent_t ev1 = evp.rec(s);
    Counters hc;
    CU(cudaMemcpyAsync(&hc, ctx->d_ctr, sizeof hc, cudaMemcpyDeviceToHost, s));
    if (m > 0) CU(cudaMemcpyAsync(out, ctx->d_out, sizeof(double) * m * 10, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    float ms = 0; cudaEventElapsedTime(&ms, ev0, ev1);
    dg.ms_total = ms; dg.ms_tests = ms; dg.d2h_bytes = m * 10 * 8;
    fill_diag(&dg, hc, m);
    dg.kernel_launches = ctx->launches;
    if (diag) *diag = dg;
    if (hc.n_nonint) return fail(ctx, SPAGXE_ERR_UNSUPPORTED,
                                 "run_localance: Geno.mtx / haplo.mtx / Cova.haplo.mtx.list hold values outside {0, 1, 2}");
    if (hc.n_allmissing) return fail(ctx, SPAGXE_ERR_ARG, "run_localance: %llu SNP(s) hold NA in g or haplo_numVec (the reference stops at `if (MAF > 0.5)`)",
                                     hc.n_allmissing);
    if (hc.n_singular) return fail(ctx, SPAGXE_ERR_SINGULAR, "solve(t(W)%%*%%W) is singular for %llu SNP(s)", hc.n_singular);
    if (hc.n_cct_stop) return fail(ctx, SPAGXE_ERR_CCT_STOP, "CCT() would stop() for %llu SNP(s) (NA or out-of-range p-value)",
                                   hc.n_cct_stop);
    return SPAGXE_OK;
}

extern "C" int spagxe_set_grm(spagxe_ctx *ctx, int64_t nnz, const int32_t *gi, const int32_t *gj, const double *gv,
                              double R_GRM_R, double R_GRM_RE, double Rnew_GRM_Rnew, const double *Rnew) {
    if (!ctx) return SPAGXE_ERR_ARG;
    if (ctx->n <= 0) return fail(ctx, SPAGXE_ERR_STATE, "set_grm: call spagxe_set_vectors first");
    if (nnz < 0 || (nnz > 0 && (!gi || !gj || !gv)) || !Rnew) return fail(ctx, SPAGXE_ERR_ARG, "set_grm: bad arguments");
    { int rc = check_grm_indices(ctx, "set_grm", nnz, gi, gj, ctx->n); if (rc) return rc; }
    CU(cudaSetDevice(ctx->device));
    CU(dev_realloc(&ctx->d_gi, std::max<int64_t>(nnz, 1))); CU(dev_realloc(&ctx->d_gj, std::max<int64_t>(nnz, 1)));
    CU(dev_realloc(&ctx->d_gv, std::max<int64_t>(nnz, 1))); CU(dev_realloc(&ctx->d_Rn_user, ctx->n));
    if (nnz > 0) {
        CU(cudaMemcpy(ctx->d_gi, gi, nnz * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(ctx->d_gj, gj, nnz * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(ctx->d_gv, gv, nnz * sizeof(double), cudaMemcpyHostToDevice));
    }
    CU(cudaMemcpy(ctx->d_Rn_user, Rnew, ctx->n * sizeof(double), cudaMemcpyHostToDevice));
    ctx->grm_nnz = nnz; ctx->R_GRM_R = R_GRM_R; ctx->R_GRM_RE = R_GRM_RE; ctx->Rnew_GRM_Rnew = Rnew_GRM_Rnew;
    ctx->have_grm = true; ctx->have_coo = true;
    ctx->vec_mode = -1;
    return SPAGXE_OK;
}

// sparseGRM of SPAGxEmix_CCT_Plus (ref R/SPAmixGxECCTPlus_functions_MYZ_2024-09-17.R:216-232): the table alone, no Step-1 scalars
extern "C" int spagxe_set_sparse_grm(spagxe_ctx *ctx, int64_t nnz, const int32_t *gi, const int32_t *gj, const double *gv) {
    if (!ctx) return SPAGXE_ERR_ARG;
    if (ctx->n <= 0) return fail(ctx, SPAGXE_ERR_STATE, "set_sparse_grm: call spagxe_set_vectors first");
    if (nnz < 0 || (nnz > 0 && (!gi || !gj || !gv))) return fail(ctx, SPAGXE_ERR_ARG, "set_sparse_grm: bad arguments");
    { int rc = check_grm_indices(ctx, "set_sparse_grm", nnz, gi, gj, ctx->n); if (rc) return rc; }
    CU(cudaSetDevice(ctx->device));
    CU(dev_realloc(&ctx->d_gi, std::max<int64_t>(nnz, 1))); CU(dev_realloc(&ctx->d_gj, std::max<int64_t>(nnz, 1)));
    CU(dev_realloc(&ctx->d_gv, std::max<int64_t>(nnz, 1)));
    if (nnz > 0) {
        CU(cudaMemcpy(ctx->d_gi, gi, nnz * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(ctx->d_gj, gj, nnz * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(ctx->d_gv, gv, nnz * sizeof(double), cudaMemcpyHostToDevice));
    }
    ctx->grm_nnz = nnz;
    ctx->have_coo = true; ctx->have_grm = false;   // SPAGxE_Plus needs its scalars and R.new again (spagxe_set_grm)
    return SPAGXE_OK;
}

static int check_grm_indices(spagxe_ctx *ctx, const char *who, int64_t nnz, const int32_t *gi, const int32_t *gj, int64_t n) {
    for (int64_t e = 0; e < nnz; e++)
        if (gi[e] < 0 || gi[e] >= n || gj[e] < 0 || gj[e] >= n)
            return fail(ctx, SPAGXE_ERR_ARG, "%s: index out of range at entry %lld", who, (long long)e);
    return SPAGXE_OK;
}

// SPAGxE_Plus_Nullmodel's quadratic forms on the device (ref R/SPAGxEPlus_functions_MYZ.R:490-537)
extern "C" int spagxe_plus_nullmodel(spagxe_ctx *ctx, int64_t n, const double *R, const double *E, int64_t nnz,
                                     const int32_t *gi, const int32_t *gj, const double *gv, double *scalars3, double *Rnew) {
    if (!ctx) return SPAGXE_ERR_ARG;
    if (n <= 0 || !R || !E || nnz < 0 || (nnz > 0 && (!gi || !gj || !gv)) || !scalars3 || !Rnew)
        return fail(ctx, SPAGXE_ERR_ARG, "plus_nullmodel: bad arguments");
    int rc = check_grm_indices(ctx, "plus_nullmodel", nnz, gi, gj, n);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream();
    // one scratch block: R | E | Rnew | out3 (8-byte items), then gv, then gi | gj
    const int64_t nz = std::max<int64_t>(nnz, 1);
    const int64_t need = (3 * n + 4 + nz) * (int64_t)sizeof(double) + 2 * nz * (int64_t)sizeof(int);
    if (need > ctx->cap_util) {
        CU(cudaStreamSynchronize(s));
        cudaFree(ctx->d_util); ctx->d_util = nullptr; ctx->cap_util = 0;
        CU(cudaMalloc(&ctx->d_util, need)); ctx->cap_util = need;
    }
    double *dR = (double *)ctx->d_util, *dE = dR + n, *dRn = dE + n, *dout = dRn + n, *dgv = dout + 4;
    int *dgi = (int *)(dgv + nz), *dgj = dgi + nz;
    CU(cudaMemcpyAsync(dR, R, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dE, E, n * sizeof(double), cudaMemcpyHostToDevice, s));
    if (nnz > 0) {
        CU(cudaMemcpyAsync(dgv, gv, nnz * sizeof(double), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(dgi, gi, nnz * sizeof(int), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(dgj, gj, nnz * sizeof(int), cudaMemcpyHostToDevice, s));
    }
    GrmCoo grm{dgi, dgj, dgv, (long long)nnz};
    k_plus_nullmodel<<<1, 1024, 0, s>>>(dR, dE, n, grm, dRn, dout);
    LAUNCH_CHECK();
    CU(cudaMemcpyAsync(scalars3, dout, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(Rnew, dRn, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return SPAGXE_OK;
}

// raw per-SNP outputs of the score pass (K1+K2) for cross-checking the two kernels
extern "C" int spagxe_score_stats(spagxe_ctx *ctx, const spagxe_geno *g, int impl, double *D0, double *D1, double *T0,
                                  double *T1, int32_t *asum, int32_t *nmiss) {
    if (!ctx) return SPAGXE_ERR_ARG;
    int rc = validate_geno(ctx, g);
    if (rc) return rc;
    if (ctx->n <= 0 || g->n_samples != ctx->n) return fail(ctx, SPAGXE_ERR_STATE, "score_stats: set_vectors first / N mismatch");
    if (!D0 || !D1 || !T0 || !T1 || !asum || !nmiss) return fail(ctx, SPAGXE_ERR_ARG, "score_stats: NULL output");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream();
    ChunkPlan pl = make_plan(g, ctx->opt_chunk_snps);
    const bool own_bed = !(g->kind == SPAGXE_GENO_BED && g->location == SPAGXE_LOC_DEVICE) || geno_needs_prepare(g);
    if ((rc = ensure_chunk_scratch(ctx, pl, own_bed, 0))) return rc;
    if ((rc = upload_sample_index(ctx, g))) return rc;
    if (ctx->vec_mode < 0 && (rc = ensure_vec_mode(ctx, 0, nullptr))) return rc;
    CU(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(Counters), s));
    SnpStats st = stats_view(ctx, ctx->cap_m);
    const bool saved = ctx->use_tc;
    ctx->use_tc = (impl != 1);
    int buf = 0; int64_t h2d = 0;
    for (int64_t j0 = 0; j0 < g->n_snps && rc == 0; j0 += pl.chunk_m, buf ^= 1) {
        const int mc = (int)std::min<int64_t>(pl.chunk_m, g->n_snps - j0);
        const uint8_t *d_rows = nullptr;
        if ((rc = stage_chunk(ctx, g, pl, j0, mc, buf, &d_rows, &h2d))) break;
        if ((rc = launch_score(ctx, d_rows, pl, mc, st))) break;
        cudaError_t e = cudaSuccess;
        double *dsts[4] = {D0, D1, T0, T1};
        const double *srcs[4] = {st.D0, st.D1, st.T0, st.T1};
        for (int k = 0; k < 4 && e == cudaSuccess; k++)
            e = cudaMemcpyAsync(dsts[k] + j0, srcs[k], sizeof(double) * mc, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(asum + j0, st.asum, sizeof(int) * mc, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nmiss + j0, st.nmiss, sizeof(int) * mc, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = fail(ctx, SPAGXE_ERR_CUDA, "score_stats: %s", cudaGetErrorString(e));
        if (own_bed) cudaEventRecord(ctx->ev_done[buf], s);
    }
    ctx->use_tc = saved;
    return rc;
}

// ---------------------------------------------------------------------------
extern "C" int spagxe_bed_counts(spagxe_ctx *ctx, const spagxe_geno *g, int64_t *counts) {
    if (!ctx) return SPAGXE_ERR_ARG;
    int rc = validate_geno(ctx, g);
    if (rc) return rc;
    if (!counts) return fail(ctx, SPAGXE_ERR_ARG, "counts is NULL");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream();
    ChunkPlan pl = make_plan(g, ctx->opt_chunk_snps);
    ctx->launches = 0;
    const bool own_bed = !(g->kind == SPAGXE_GENO_BED && g->location == SPAGXE_LOC_DEVICE) || geno_needs_prepare(g);
    if ((rc = ensure_chunk_scratch(ctx, pl, own_bed, 0))) return rc;
    if ((rc = upload_sample_index(ctx, g))) return rc;
    CU(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(Counters), s));
    long long *d_counts = nullptr;
    {
        const int64_t need = (int64_t)sizeof(long long) * 4 * pl.chunk_m;
        if (need > ctx->cap_util) { cudaFree(ctx->d_util); ctx->d_util = nullptr; ctx->cap_util = 0; CU(cudaMalloc(&ctx->d_util, need)); ctx->cap_util = need; }
        d_counts = (long long *)ctx->d_util;
    }
    int buf = 0; int64_t h2d = 0;
    for (int64_t j0 = 0; j0 < g->n_snps; j0 += pl.chunk_m, buf ^= 1) {
        const int mc = (int)std::min<int64_t>(pl.chunk_m, g->n_snps - j0);
        const uint8_t *d_rows = nullptr;
        if ((rc = stage_chunk(ctx, g, pl, j0, mc, buf, &d_rows, &h2d))) return rc;
        k_bed_counts<<<(mc * 32 + 255) / 256, 256, 0, s>>>(d_rows, pl.dpitch, pl.n, mc, d_counts);
        ctx->launches++;
        cudaError_t e = cudaMemcpyAsync(counts + 4 * j0, d_counts, sizeof(long long) * 4 * mc, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return fail(ctx, SPAGXE_ERR_CUDA, "bed_counts: %s", cudaGetErrorString(e));
        if (own_bed) cudaEventRecord(ctx->ev_done[buf], s);
    }
    return SPAGXE_OK;
}

extern "C" int spagxe_bed_decode(spagxe_ctx *ctx, const spagxe_geno *g, double *out_host) {
    if (!ctx) return SPAGXE_ERR_ARG;
    int rc = validate_geno(ctx, g);
    if (rc) return rc;
    if (!out_host) return fail(ctx, SPAGXE_ERR_ARG, "out is NULL");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream();
    ChunkPlan pl = make_plan(g, ctx->opt_chunk_snps);
    // decode expands 32x: keep chunks small
    pl.chunk_m = (int)std::min<int64_t>(pl.chunk_m, std::max<int64_t>(1, (1ll << 28) / (pl.n * 8)));
    const bool own_bed = !(g->kind == SPAGXE_GENO_BED && g->location == SPAGXE_LOC_DEVICE) || geno_needs_prepare(g);
    if ((rc = ensure_chunk_scratch(ctx, pl, own_bed, 0))) return rc;
    if ((rc = upload_sample_index(ctx, g))) return rc;
    double *d_dec = nullptr;
    {
        const int64_t need = (int64_t)sizeof(double) * pl.n * pl.chunk_m;
        if (need > ctx->cap_util) { cudaFree(ctx->d_util); ctx->d_util = nullptr; ctx->cap_util = 0; CU(cudaMalloc(&ctx->d_util, need)); ctx->cap_util = need; }
        d_dec = (double *)ctx->d_util;
    }
    int buf = 0; int64_t h2d = 0;
    for (int64_t j0 = 0; j0 < g->n_snps; j0 += pl.chunk_m, buf ^= 1) {
        const int mc = (int)std::min<int64_t>(pl.chunk_m, g->n_snps - j0);
        const uint8_t *d_rows = nullptr;
        if ((rc = stage_chunk(ctx, g, pl, j0, mc, buf, &d_rows, &h2d))) return rc;
        dim3 grid((unsigned)((pl.n + 255) / 256), (unsigned)mc);
        k_bed_decode<<<grid, 256, 0, s>>>(d_rows, pl.dpitch, pl.n, mc, d_dec);
        ctx->launches++;
        cudaError_t e = cudaMemcpyAsync(out_host + j0 * pl.n, d_dec, sizeof(double) * pl.n * mc, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return fail(ctx, SPAGXE_ERR_CUDA, "bed_decode: %s", cudaGetErrorString(e));
        if (own_bed) cudaEventRecord(ctx->ev_done[buf], s);
    }
    return SPAGXE_OK;
}

// ---------------------------------------------------------------------------
__global__ void k_dfma_peak(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}
extern "C" int spagxe_measure_fp64_peak(spagxe_ctx *ctx, double *tflops) {
    if (!ctx || !tflops) return SPAGXE_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream();
    double *d = nullptr;
    CU(cudaMalloc((void **)&d, 8));
    const int iters = 1 << 16, blocks = ctx->num_sms * 8, threads = 256;
    k_dfma_peak<<<blocks, threads, 0, s>>>(d, 1024);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a, s);
        k_dfma_peak<<<blocks, threads, 0, s>>>(d, iters);
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        double fl = 2.0 * 8 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        best = std::max(best, fl);
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SPAGXE_ERR_CUDA, "fp64 peak: %s", cudaGetErrorString(e));
    *tflops = best;
    return SPAGXE_OK;
}

extern "C" int spagxe_set_pcs(spagxe_ctx *ctx, const double *pcs, int k) {
    if (!ctx) return SPAGXE_ERR_ARG;
    if (ctx->n <= 0) return fail(ctx, SPAGXE_ERR_STATE, "set_pcs: call spagxe_set_vectors first");
    if (!pcs || k < 1 || k > kMaxPC) return fail(ctx, SPAGXE_ERR_ARG, "set_pcs: need 1 <= k <= %d principal components", kMaxPC);
    const int64_t n = ctx->n;
    const int k1 = k + 1;
    // (X'X)^-1 for X = [1, PCs] is SNP-invariant: long-double normal matrix + Cholesky on the host, once
    std::vector<long double> A((size_t)k1 * k1, 0.0L);
    for (int a = 0; a < k1; a++)
        for (int b = a; b < k1; b++) {
            long double sacc = 0;
            const double *ca = a ? pcs + (size_t)(a - 1) * n : nullptr, *cb = b ? pcs + (size_t)(b - 1) * n : nullptr;
            for (int64_t i = 0; i < n; i++) sacc += (long double)(ca ? ca[i] : 1.0) * (cb ? cb[i] : 1.0);
            A[a + (size_t)b * k1] = A[b + (size_t)a * k1] = sacc;
        }
    std::vector<long double> L((size_t)k1 * k1, 0.0L), Li((size_t)k1 * k1, 0.0L);
    for (int j = 0; j < k1; j++) {
        long double d = A[j + (size_t)j * k1];
        for (int c = 0; c < j; c++) d -= L[j + (size_t)c * k1] * L[j + (size_t)c * k1];
        if (!(d > 0)) return fail(ctx, SPAGXE_ERR_SINGULAR, "set_pcs: [1, topPCs] is rank deficient");
        d = sqrtl(d); L[j + (size_t)j * k1] = d;
        for (int i = j + 1; i < k1; i++) {
            long double sacc = A[i + (size_t)j * k1];
            for (int c = 0; c < j; c++) sacc -= L[i + (size_t)c * k1] * L[j + (size_t)c * k1];
            L[i + (size_t)j * k1] = sacc / d;
        }
    }
    for (int c = 0; c < k1; c++)   // Li = L^-1 (lower triangular)
        for (int i = 0; i < k1; i++) {
            long double sacc = (i == c) ? 1.0L : 0.0L;
            for (int t = 0; t < i; t++) sacc -= L[i + (size_t)t * k1] * Li[t + (size_t)c * k1];
            Li[i + (size_t)c * k1] = sacc / L[i + (size_t)i * k1];
        }
    std::vector<double> inv((size_t)k1 * k1);
    for (int a = 0; a < k1; a++)
        for (int b = 0; b < k1; b++) {
            long double sacc = 0;
            for (int t = 0; t < k1; t++) sacc += Li[t + (size_t)a * k1] * Li[t + (size_t)b * k1];
            inv[a + (size_t)b * k1] = (double)sacc;
        }
    CU(cudaSetDevice(ctx->device));
    CU(dev_realloc(&ctx->d_pcs, (size_t)n * k));
    CU(dev_realloc(&ctx->d_xtx_inv, (size_t)k1 * k1));
    CU(cudaMemcpy(ctx->d_pcs, pcs, (size_t)n * k * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_xtx_inv, inv.data(), (size_t)k1 * k1 * sizeof(double), cudaMemcpyHostToDevice));
    ctx->n_pcs = k;
    return SPAGXE_OK;
}
