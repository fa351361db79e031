// mgpu.cu -- the Step-2 loop on several GPUs of one node behind the C ABI (include/spagxe.h, spagxe_mgpu_*).
//
// SNPs are independent and a .bed file is variant-major, so the M SNPs of a call are cut into contiguous
// ranges, one per GPU.  One process; every GPU has its own single-GPU context (own stream, own scratch) and is driven by
// its own host thread for the duration of a call.  The per-analysis vectors are uploaded to the first GPU and reach the
// others through ONE ncclBroadcast of [R | E | Y | Cova]; each GPU writes its rows into its own row block of the caller's
// column-major result matrix (params.out_ld).  No other collective and no reduction across GPUs.
//
// NCCL is bound at run time (dlopen) so that libspagxe_cuda.so has no link-time dependency on it: single-GPU callers
// never load it, and inside a torch process the already loaded libnccl.so.2 is the one that is used.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/spagxe.h"

namespace {

// the six NCCL entry points this file needs (types as in nccl.h)
typedef void *nccl_comm_t;
typedef int nccl_result_t;                 // ncclSuccess == 0
constexpr int kNcclFloat64 = 8;            // ncclDataType_t: ncclFloat64 / ncclDouble
struct NcclApi {
    void *handle = nullptr;
    nccl_result_t (*CommInitAll)(nccl_comm_t *, int, const int *) = nullptr;
    nccl_result_t (*CommDestroy)(nccl_comm_t) = nullptr;
    nccl_result_t (*Broadcast)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    nccl_result_t (*GroupStart)() = nullptr;
    nccl_result_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(nccl_result_t) = nullptr;
    bool load(std::string *err) {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) { *err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
        CommInitAll = (decltype(CommInitAll))dlsym(handle, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        Broadcast = (decltype(Broadcast))dlsym(handle, "ncclBroadcast");
        GroupStart = (decltype(GroupStart))dlsym(handle, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(handle, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !Broadcast || !GroupStart || !GroupEnd || !GetErrorString) {
            *err = "libnccl.so.2 lacks one of ncclCommInitAll/ncclBroadcast/ncclGroupStart/ncclGroupEnd/ncclCommDestroy";
            return false;
        }
        return true;
    }
};
NcclApi g_nccl;
std::string g_mgpu_create_error;

}  // namespace

struct spagxe_mgpu {
    int n = 0;
    std::vector<int> devices;
    std::vector<spagxe_ctx *> ctx;
    std::vector<nccl_comm_t> comms;
    std::vector<cudaStream_t> streams;      // one per GPU for the broadcast
    std::vector<double *> d_in;             // per GPU: [R | E | Y | Cova] as received
    int64_t cap_in = 0;
    int64_t n_samples = 0;
    std::string err;
};

static int mfail(spagxe_mgpu *mg, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (mg) mg->err = buf; else g_mgpu_create_error = buf;
    return code;
}

extern "C" const char *spagxe_mgpu_last_error(const spagxe_mgpu *mg) {
    return mg ? mg->err.c_str() : g_mgpu_create_error.c_str();
}
extern "C" int spagxe_mgpu_n_gpus(const spagxe_mgpu *mg) { return mg ? mg->n : 0; }

extern "C" void spagxe_mgpu_destroy(spagxe_mgpu *mg) {
    if (!mg) return;
    for (int k = 0; k < (int)mg->ctx.size(); k++) {
        cudaSetDevice(mg->devices[k]);
        if (k < (int)mg->d_in.size() && mg->d_in[k]) cudaFree(mg->d_in[k]);
        if (k < (int)mg->streams.size() && mg->streams[k]) cudaStreamDestroy(mg->streams[k]);
        if (k < (int)mg->comms.size() && mg->comms[k] && g_nccl.CommDestroy) g_nccl.CommDestroy(mg->comms[k]);
        if (mg->ctx[k]) spagxe_ctx_destroy(mg->ctx[k]);
    }
    delete mg;
}

extern "C" int spagxe_mgpu_create(int n_gpus, const int *devices, spagxe_mgpu **out) {
    if (!out) return mfail(nullptr, SPAGXE_ERR_ARG, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return mfail(nullptr, SPAGXE_ERR_CUDA, "no CUDA device: %s (libspagxe_cuda has no CPU fallback)", cudaGetErrorString(e));
    if (n_gpus < 1 || n_gpus > ndev) return mfail(nullptr, SPAGXE_ERR_ARG, "n_gpus = %d but %d device(s) are visible", n_gpus, ndev);
    spagxe_mgpu *mg = new spagxe_mgpu();
    mg->n = n_gpus;
    for (int k = 0; k < n_gpus; k++) mg->devices.push_back(devices ? devices[k] : k);
    mg->ctx.assign(n_gpus, nullptr); mg->comms.assign(n_gpus, nullptr); mg->streams.assign(n_gpus, nullptr);
    mg->d_in.assign(n_gpus, nullptr);
    for (int k = 0; k < n_gpus; k++) {
        int rc = spagxe_ctx_create(mg->devices[k], &mg->ctx[k]);
        if (rc) { int r = mfail(nullptr, rc, "GPU %d: %s", mg->devices[k], spagxe_last_error(nullptr)); spagxe_mgpu_destroy(mg); return r; }
        cudaSetDevice(mg->devices[k]);
        if ((e = cudaStreamCreateWithFlags(&mg->streams[k], cudaStreamNonBlocking)) != cudaSuccess) {
            int r = mfail(nullptr, SPAGXE_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); spagxe_mgpu_destroy(mg); return r;
        }
    }
    if (n_gpus > 1) {
        std::string lerr;
        if (!g_nccl.load(&lerr)) { int r = mfail(nullptr, SPAGXE_ERR_CUDA, "%s", lerr.c_str()); spagxe_mgpu_destroy(mg); return r; }
        nccl_result_t nr = g_nccl.CommInitAll(mg->comms.data(), n_gpus, mg->devices.data());
        if (nr != 0) {
            int r = mfail(nullptr, SPAGXE_ERR_CUDA, "ncclCommInitAll: %s", g_nccl.GetErrorString(nr));
            for (auto &c : mg->comms) c = nullptr;
            spagxe_mgpu_destroy(mg); return r;
        }
    }
    *out = mg;
    return SPAGXE_OK;
}

#define MCU(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) return mfail(mg, SPAGXE_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

extern "C" int spagxe_mgpu_set_inputs(spagxe_mgpu *mg, int64_t n, const double *R, const double *E, int trait, const double *Y,
                                      const double *cova, int p) {
    if (!mg) return SPAGXE_ERR_ARG;
    // trait == 0: no Wald data (SPAGxE_Plus, or a survival trait whose data follow through spagxe_mgpu_set_wald_survival)
    if (trait == 0) { Y = nullptr; cova = nullptr; p = 0; }
    if (n <= 0 || !R || !E || (trait != 0 && !Y) || p < 0 || (p > 0 && !cova)) return mfail(mg, SPAGXE_ERR_ARG, "set_inputs: bad arguments");
    const int64_t total = n * (3 + (int64_t)p);
    if (total > mg->cap_in) {
        for (int k = 0; k < mg->n; k++) {
            MCU(cudaSetDevice(mg->devices[k]));
            if (mg->d_in[k]) cudaFree(mg->d_in[k]);
            mg->d_in[k] = nullptr;
            MCU(cudaMalloc((void **)&mg->d_in[k], total * sizeof(double)));
        }
        mg->cap_in = total;
    }
    // upload to the first GPU ...
    MCU(cudaSetDevice(mg->devices[0]));
    cudaStream_t s0 = mg->streams[0];
    double *b0 = mg->d_in[0];
    MCU(cudaMemcpyAsync(b0, R, n * sizeof(double), cudaMemcpyHostToDevice, s0));
    MCU(cudaMemcpyAsync(b0 + n, E, n * sizeof(double), cudaMemcpyHostToDevice, s0));
    if (Y) MCU(cudaMemcpyAsync(b0 + 2 * n, Y, n * sizeof(double), cudaMemcpyHostToDevice, s0));
    else MCU(cudaMemsetAsync(b0 + 2 * n, 0, n * sizeof(double), s0));
    if (p > 0) MCU(cudaMemcpyAsync(b0 + 3 * n, cova, n * (int64_t)p * sizeof(double), cudaMemcpyHostToDevice, s0));
    // ... and ONE broadcast to the others (the only collective of the whole analysis)
    if (mg->n > 1) {
        nccl_result_t nr = g_nccl.GroupStart();
        for (int k = 0; k < mg->n && nr == 0; k++)
            nr = g_nccl.Broadcast(mg->d_in[0], mg->d_in[k], (size_t)total, kNcclFloat64, 0, mg->comms[k], mg->streams[k]);
        if (nr == 0) nr = g_nccl.GroupEnd(); else g_nccl.GroupEnd();
        if (nr != 0) return mfail(mg, SPAGXE_ERR_CUDA, "ncclBroadcast: %s", g_nccl.GetErrorString(nr));
    }
    for (int k = 0; k < mg->n; k++) {
        MCU(cudaSetDevice(mg->devices[k]));
        MCU(cudaStreamSynchronize(mg->streams[k]));
    }
    for (int k = 0; k < mg->n; k++) {
        const double *b = mg->d_in[k];
        int rc = spagxe_set_vectors_device(mg->ctx[k], n, b, b + n);
        if (!rc && trait != 0) rc = spagxe_set_wald_device(mg->ctx[k], trait, b + 2 * n, b + 3 * n, p);
        if (rc) return mfail(mg, rc, "GPU %d: %s", mg->devices[k], spagxe_last_error(mg->ctx[k]));
    }
    mg->n_samples = n;
    return SPAGXE_OK;
}

// per-analysis data beyond [R | E | Y | Cova]: every GPU receives its copy from the host arrays (tens of MB at most)
#define MG_EACH(expr)                                                                                  \
    do {                                                                                               \
        if (!mg) return SPAGXE_ERR_ARG;                                                                \
        if (mg->n_samples <= 0) return mfail(mg, SPAGXE_ERR_STATE, "call spagxe_mgpu_set_inputs first"); \
        for (int k = 0; k < mg->n; k++) {                                                              \
            cudaSetDevice(mg->devices[k]);                                                             \
            spagxe_ctx *c = mg->ctx[k];                                                                \
            const int rc = (expr);                                                                     \
            if (rc) return mfail(mg, rc, "GPU %d: %s", mg->devices[k], spagxe_last_error(c));          \
        }                                                                                              \
        return SPAGXE_OK;                                                                              \
    } while (0)
extern "C" int spagxe_mgpu_set_wald_survival(spagxe_mgpu *mg, const double *surv_time, const double *event, const double *cova, int p) {
    MG_EACH(spagxe_set_wald_survival(c, surv_time, event, cova, p));
}
extern "C" int spagxe_mgpu_set_pcs(spagxe_mgpu *mg, const double *pcs, int k_pcs) { MG_EACH(spagxe_set_pcs(c, pcs, k_pcs)); }
extern "C" int spagxe_mgpu_set_grm(spagxe_mgpu *mg, int64_t nnz, const int32_t *gi, const int32_t *gj, const double *gv,
                                   double R_GRM_R, double R_GRM_RE, double Rnew_GRM_Rnew, const double *Rnew) {
    MG_EACH(spagxe_set_grm(c, nnz, gi, gj, gv, R_GRM_R, R_GRM_RE, Rnew_GRM_Rnew, Rnew));
}
extern "C" int spagxe_mgpu_set_sparse_grm(spagxe_mgpu *mg, int64_t nnz, const int32_t *gi, const int32_t *gj, const double *gv) {
    MG_EACH(spagxe_set_sparse_grm(c, nnz, gi, gj, gv));
}

static int mgpu_run(spagxe_mgpu *mg, const spagxe_geno *geno, int n_shards, const spagxe_params *prm, double *out,
                    spagxe_diag *diag, int which);
extern "C" int spagxe_mgpu_run_cct(spagxe_mgpu *mg, const spagxe_geno *geno, int n_shards, const spagxe_params *prm, double *out,
                                   spagxe_diag *diag) { return mgpu_run(mg, geno, n_shards, prm, out, diag, 0); }
extern "C" int spagxe_mgpu_run_mix(spagxe_mgpu *mg, const spagxe_geno *geno, int n_shards, const spagxe_params *prm, double *out,
                                   spagxe_diag *diag) { return mgpu_run(mg, geno, n_shards, prm, out, diag, 1); }
extern "C" int spagxe_mgpu_run_plus(spagxe_mgpu *mg, const spagxe_geno *geno, int n_shards, const spagxe_params *prm, double *out,
                                    spagxe_diag *diag) { return mgpu_run(mg, geno, n_shards, prm, out, diag, 2); }
extern "C" int spagxe_mgpu_run_mix_plus(spagxe_mgpu *mg, const spagxe_geno *geno, int n_shards, const spagxe_params *prm, double *out,
                                        spagxe_diag *diag) { return mgpu_run(mg, geno, n_shards, prm, out, diag, 3); }

static int mgpu_run(spagxe_mgpu *mg, const spagxe_geno *geno, int n_shards, const spagxe_params *prm, double *out,
                    spagxe_diag *diag, int which) {
    if (!mg) return SPAGXE_ERR_ARG;
    if (!geno || !out) return mfail(mg, SPAGXE_ERR_ARG, "run_cct: geno / out is NULL");
    if (n_shards != 1 && n_shards != mg->n) return mfail(mg, SPAGXE_ERR_ARG, "run_cct: n_shards must be 1 or n_gpus (%d)", mg->n);
    if (mg->n_samples <= 0) return mfail(mg, SPAGXE_ERR_STATE, "run_cct: call spagxe_mgpu_set_inputs first");
    std::vector<spagxe_geno> sh(mg->n);
    std::vector<int64_t> row0(mg->n, 0);
    int64_t M = 0;
    if (n_shards == 1) {
        if (geno->location != SPAGXE_LOC_HOST && mg->n > 1)
            return mfail(mg, SPAGXE_ERR_ARG, "run_cct: a single genotype source must be host memory (it is cut across the GPUs)");
        M = geno->n_snps;
        const size_t esz = geno->kind == SPAGXE_GENO_BED ? 1 : geno->kind == SPAGXE_GENO_F64 ? 8 : 4;
        const int64_t base = M / mg->n, rem = M % mg->n;
        int64_t lo = 0;
        for (int k = 0; k < mg->n; k++) {
            const int64_t cnt = base + (k < rem ? 1 : 0);
            sh[k] = *geno;
            sh[k].data = (const char *)geno->data + (size_t)lo * (size_t)geno->pitch * esz;
            sh[k].n_snps = cnt;
            row0[k] = lo;
            lo += cnt;
        }
    } else {
        for (int k = 0; k < mg->n; k++) { sh[k] = geno[k]; row0[k] = M; M += geno[k].n_snps; }
    }
    spagxe_params p0;
    if (prm) p0 = *prm; else { memset(&p0, 0, sizeof p0); p0.epsilon = 0.001; p0.cutoff = 2.0; p0.min_maf = 0.001; p0.missing_cutoff = 0.15; p0.pc_pvalue_cutoff = 0.05; p0.neg_ratio_cutoff = 0.1; }
    if (p0.out_location != SPAGXE_LOC_HOST) return mfail(mg, SPAGXE_ERR_ARG, "run_cct: the result matrix must be host memory");
    p0.out_ld = M;
    std::vector<int> rcs(mg->n, 0);
    std::vector<spagxe_diag> dgs(mg->n);
    std::vector<std::thread> th;
    for (int k = 0; k < mg->n; k++) {
        th.emplace_back([&, k]() {
            memset(&dgs[k], 0, sizeof(spagxe_diag));
            if (sh[k].n_snps == 0) return;
            cudaSetDevice(mg->devices[k]);
            rcs[k] = which == 0 ? spagxe_run_cct(mg->ctx[k], &sh[k], &p0, out + row0[k], &dgs[k])
                   : which == 1 ? spagxe_run_mix(mg->ctx[k], &sh[k], &p0, out + row0[k], &dgs[k])
                   : which == 3 ? spagxe_run_mix_plus(mg->ctx[k], &sh[k], &p0, out + row0[k], &dgs[k])
                                : spagxe_run_plus(mg->ctx[k], &sh[k], &p0, out + row0[k], &dgs[k]);
        });
    }
    for (auto &t : th) t.join();
    spagxe_diag tot; memset(&tot, 0, sizeof tot);
    for (int k = 0; k < mg->n; k++) {
        const spagxe_diag &d = dgs[k];
        tot.n_snps += d.n_snps; tot.n_filtered += d.n_filtered; tot.n_branch_a += d.n_branch_a; tot.n_branch_b += d.n_branch_b;
        tot.n_spa += d.n_spa; tot.n_wald_fail += d.n_wald_fail; tot.n_cct_stop += d.n_cct_stop; tot.n_singular += d.n_singular;
        tot.n_mix_logistic += d.n_mix_logistic; tot.newton_iters += d.newton_iters; tot.kernel_launches += d.kernel_launches;
        tot.h2d_bytes += d.h2d_bytes; tot.d2h_bytes += d.d2h_bytes; tot.score_bytes += d.score_bytes;
        tot.bb_irls_sample_passes += d.bb_irls_sample_passes; tot.bb_tail_sample_passes += d.bb_tail_sample_passes;
        tot.bb_prep_sample_passes += d.bb_prep_sample_passes; tot.bb_iterations += d.bb_iterations; tot.cox_evals += d.cox_evals; tot.clm_evals += d.clm_evals;
        if (d.ms_total > tot.ms_total) tot.ms_total = d.ms_total;
        if (d.ms_score > tot.ms_score) tot.ms_score = d.ms_score;
        if (d.ms_tests > tot.ms_tests) tot.ms_tests = d.ms_tests;
        if (d.ms_spa > tot.ms_spa) tot.ms_spa = d.ms_spa;
        if (d.ms_branch_b > tot.ms_branch_b) tot.ms_branch_b = d.ms_branch_b;
        if (d.spa_nodes > tot.spa_nodes) tot.spa_nodes = d.spa_nodes;
    }
    if (diag) *diag = tot;
    for (int k = 0; k < mg->n; k++)
        if (rcs[k]) return mfail(mg, rcs[k], "GPU %d: %s", mg->devices[k], spagxe_last_error(mg->ctx[k]));
    return SPAGXE_OK;
}
