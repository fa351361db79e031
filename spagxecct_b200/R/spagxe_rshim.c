/* spagxe_rshim.c -- the thin .Call layer between R and libspagxe_cuda.so (include/spagxe.h).
 *
 * Compiled only where R exists (R CMD SHLIB spagxe_rshim.c -L.. -lspagxe_cuda); the build image has no R, so
 * this file is not part of the default build.  It only unwraps SEXPs, allocates the result matrix and turns
 * status codes into R errors; no C++ object is alive when Rf_error() long-jumps.  R owns every input (read-only,
 * REAL()/INTEGER()/RAW() taken once on the main thread); contexts are external pointers that the wrappers destroy with
 * on.exit() (the finalizer is only the safety net).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/spagxe.h"

static void ctx_finalizer(SEXP ptr) {
    spagxe_ctx *ctx = (spagxe_ctx *)R_ExternalPtrAddr(ptr);
    if (ctx) { spagxe_ctx_destroy(ctx); R_ClearExternalPtr(ptr); }
}
static void mgpu_finalizer(SEXP ptr) {
    spagxe_mgpu *mg = (spagxe_mgpu *)R_ExternalPtrAddr(ptr);
    if (mg) { spagxe_mgpu_destroy(mg); R_ClearExternalPtr(ptr); }
}

SEXP spagxe_r_create(SEXP device) {
    spagxe_ctx *ctx = NULL;
    if (spagxe_ctx_create(asInteger(device), &ctx) != SPAGXE_OK) Rf_error("%s", spagxe_last_error(NULL));
    SEXP ptr = PROTECT(R_MakeExternalPtr(ctx, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}
SEXP spagxe_r_destroy(SEXP ptr) { ctx_finalizer(ptr); return R_NilValue; }

static spagxe_ctx *get_ctx(SEXP ptr) {
    spagxe_ctx *ctx = (spagxe_ctx *)R_ExternalPtrAddr(ptr);
    if (!ctx) Rf_error("spagxe: context already released");
    return ctx;
}
#define CHECK(ctx, call) do { if ((call) != SPAGXE_OK) Rf_error("%s", spagxe_last_error(ctx)); } while (0)

SEXP spagxe_r_set_vectors(SEXP ptr, SEXP R, SEXP E) {
    spagxe_ctx *ctx = get_ctx(ptr);
    if (XLENGTH(R) != XLENGTH(E)) Rf_error("R and E should be of the same length");
    CHECK(ctx, spagxe_set_vectors(ctx, XLENGTH(R), REAL(R), REAL(E)));
    return R_NilValue;
}
SEXP spagxe_r_set_wald(SEXP ptr, SEXP trait, SEXP Y, SEXP cova) {
    spagxe_ctx *ctx = get_ctx(ptr);
    CHECK(ctx, spagxe_set_wald(ctx, asInteger(trait), REAL(Y), REAL(cova), ncols(cova)));
    return R_NilValue;
}
/* traits = "survival": Phen.mtx$surv.time, Phen.mtx$event (R/SPAGxECCT.R:624) */
SEXP spagxe_r_set_wald_survival(SEXP ptr, SEXP time, SEXP event, SEXP cova) {
    spagxe_ctx *ctx = get_ctx(ptr);
    CHECK(ctx, spagxe_set_wald_survival(ctx, REAL(time), REAL(event), REAL(cova), ncols(cova)));
    return R_NilValue;
}
SEXP spagxe_r_set_pcs(SEXP ptr, SEXP pcs) {
    spagxe_ctx *ctx = get_ctx(ptr);
    CHECK(ctx, spagxe_set_pcs(ctx, REAL(pcs), ncols(pcs)));
    return R_NilValue;
}
SEXP spagxe_r_set_grm(SEXP ptr, SEXP gi, SEXP gj, SEXP gv, SEXP scalars, SEXP Rnew) {
    spagxe_ctx *ctx = get_ctx(ptr);
    CHECK(ctx, spagxe_set_grm(ctx, XLENGTH(gv), INTEGER(gi), INTEGER(gj), REAL(gv), REAL(scalars)[0], REAL(scalars)[1],
                              REAL(scalars)[2], REAL(Rnew)));
    return R_NilValue;
}
/* sparseGRM of SPAGxEmix_CCT_Plus: 0-based positions of ID1 / ID2 in ResidMat$SubjID and Value */
SEXP spagxe_r_set_sparse_grm(SEXP ptr, SEXP gi, SEXP gj, SEXP gv) {
    spagxe_ctx *ctx = get_ctx(ptr);
    CHECK(ctx, spagxe_set_sparse_grm(ctx, XLENGTH(gv), INTEGER(gi), INTEGER(gj), REAL(gv)));
    return R_NilValue;
}
/* list(scalars = c(R_GRM_R, R_GRM_RE, R.new_GRM_R.new), R.new = numeric(n)) */
SEXP spagxe_r_plus_nullmodel(SEXP ptr, SEXP R, SEXP E, SEXP gi, SEXP gj, SEXP gv) {
    spagxe_ctx *ctx = get_ctx(ptr);
    const R_xlen_t n = XLENGTH(R);
    SEXP sc = PROTECT(allocVector(REALSXP, 3)), rn = PROTECT(allocVector(REALSXP, n));
    int rc = spagxe_plus_nullmodel(ctx, n, REAL(R), REAL(E), XLENGTH(gv), INTEGER(gi), INTEGER(gj), REAL(gv), REAL(sc), REAL(rn));
    if (rc != SPAGXE_OK) { UNPROTECT(2); Rf_error("%s", spagxe_last_error(ctx)); }
    SEXP out = PROTECT(allocVector(VECSXP, 2)), nm = PROTECT(allocVector(STRSXP, 2));
    SET_VECTOR_ELT(out, 0, sc); SET_VECTOR_ELT(out, 1, rn);
    SET_STRING_ELT(nm, 0, mkChar("scalars")); SET_STRING_ELT(nm, 1, mkChar("R.new"));
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(4);
    return out;
}

/* geno: numeric/integer matrix (N x M) or a raw vector of packed .bed rows; sidx: NULL or 0-based file positions
 * (doubles: R has no 64-bit integer) of the analysis samples; idx_out: malloc'ed int64 copy the caller frees */
static void geno_from_sexp(SEXP g, SEXP n_, SEXP m_, SEXP sidx, SEXP nfile, SEXP a2, spagxe_geno *d, int64_t **idx_out) {
    memset(d, 0, sizeof *d);
    *idx_out = NULL;
    d->location = SPAGXE_LOC_HOST;
    if (TYPEOF(g) == RAWSXP) {
        d->kind = SPAGXE_GENO_BED; d->data = RAW(g);
        d->n_samples = (int64_t)asReal(n_); d->n_snps = (int64_t)asReal(m_);
        const int64_t nf = (int64_t)asReal(nfile);
        d->pitch = ((sidx != R_NilValue && nf > 0 ? nf : d->n_samples) + 3) / 4;
        if (sidx != R_NilValue) {
            int64_t *ix = (int64_t *)malloc(sizeof(int64_t) * (size_t)d->n_samples);
            if (!ix) Rf_error("spagxe: out of memory");
            for (int64_t i = 0; i < d->n_samples; i++) ix[i] = (int64_t)REAL(sidx)[i];
            d->sample_index = ix; d->n_file_samples = nf; *idx_out = ix;
        }
        d->flags = asInteger(a2) ? SPAGXE_GENO_COUNT_A2 : 0;
    } else if (TYPEOF(g) == INTSXP) {
        d->kind = SPAGXE_GENO_I32; d->data = INTEGER(g); d->n_samples = nrows(g); d->n_snps = ncols(g); d->pitch = nrows(g);
    } else {
        d->kind = SPAGXE_GENO_F64; d->data = REAL(g); d->n_samples = nrows(g); d->n_snps = ncols(g); d->pitch = nrows(g);
    }
}

static void params_from_sexp(SEXP prm, spagxe_params *sp) {
    const double *p = REAL(prm);
    memset(sp, 0, sizeof *sp);
    sp->epsilon = p[0]; sp->cutoff = p[1]; sp->min_maf = p[2]; sp->missing_cutoff = p[3];
    sp->pc_pvalue_cutoff = p[4]; sp->neg_ratio_cutoff = p[5]; sp->out_location = SPAGXE_LOC_HOST;
    if (XLENGTH(prm) > 6) sp->g_model = (int)p[6];   /* 0 "Add", 1 "Dom", 2 "Rec" (R/SPAGxECCT.R:572-574) */
}

/* attr(out, "diag") <- named numeric vector of the routing counts and device times */
static void attach_diag(SEXP out, const spagxe_diag *d) {
    static const char *names[] = {"n_snps", "n_filtered", "n_branch_a", "n_branch_b", "n_spa", "n_wald_fail", "n_mix_logistic",
                                  "newton_iters", "kernel_launches", "ms_total", "ms_score", "ms_tests"};
    const double vals[] = {(double)d->n_snps, (double)d->n_filtered, (double)d->n_branch_a, (double)d->n_branch_b, (double)d->n_spa,
                           (double)d->n_wald_fail, (double)d->n_mix_logistic, (double)d->newton_iters, (double)d->kernel_launches,
                           d->ms_total, d->ms_score, d->ms_tests};
    const int k = (int)(sizeof vals / sizeof vals[0]);
    SEXP v = PROTECT(allocVector(REALSXP, k)), nm = PROTECT(allocVector(STRSXP, k));
    for (int i = 0; i < k; i++) { REAL(v)[i] = vals[i]; SET_STRING_ELT(nm, i, mkChar(names[i])); }
    setAttrib(v, R_NamesSymbol, nm);
    setAttrib(out, install("diag"), v);
    UNPROTECT(2);
}

/* which: 0 cct (M x 10), 1 mix (M x 12), 2 plus (M x 8), 3 mix_plus (M x 12); prm = c(epsilon, Cutoff, min.maf, missing.cutoff, pc.cut, neg.cut) */
SEXP spagxe_r_run(SEXP ptr, SEXP which, SEXP geno, SEXP n_, SEXP m_, SEXP prm, SEXP sidx, SEXP nfile, SEXP a2) {
    spagxe_ctx *ctx = get_ctx(ptr);
    spagxe_geno gd; int64_t *ix = NULL;
    geno_from_sexp(geno, n_, m_, sidx, nfile, a2, &gd, &ix);
    spagxe_params sp; params_from_sexp(prm, &sp);
    const int w = asInteger(which), ncol = (w == 0) ? 10 : ((w == 1 || w == 3) ? 12 : 8);
    SEXP out = PROTECT(allocMatrix(REALSXP, (int)gd.n_snps, ncol));
    spagxe_diag diag; memset(&diag, 0, sizeof diag);
    int rc = (w == 0) ? spagxe_run_cct(ctx, &gd, &sp, REAL(out), &diag)
           : (w == 1) ? spagxe_run_mix(ctx, &gd, &sp, REAL(out), &diag)
           : (w == 3) ? spagxe_run_mix_plus(ctx, &gd, &sp, REAL(out), &diag)
                      : spagxe_run_plus(ctx, &gd, &sp, REAL(out), &diag);
    free(ix);
    if (rc != SPAGXE_OK) { UNPROTECT(1); Rf_error("%s", spagxe_last_error(ctx)); }
    attach_diag(out, &diag);
    UNPROTECT(1);
    return out;
}

/* SPAGxEmixCCT_localance (R/SPAGxECCT.R:1443-1476): Geno.mtx, haplo.mtx (REALSXP n x m) and the list of the
 * Cova.haplo.mtx.list matrices (each REALSXP n x m) */
SEXP spagxe_r_run_localance(SEXP ptr, SEXP G, SEXP H, SEXP chl, SEXP prm) {
    spagxe_ctx *ctx = get_ctx(ptr);
    const int n = nrows(G), m = ncols(G), L = (int)XLENGTH(chl);
    if (nrows(H) != n || ncols(H) != m) Rf_error("haplo.mtx must have the shape of Geno.mtx");
    if (L > 4) Rf_error("at most 4 matrices in Cova.haplo.mtx.list");
    const double *ch[4] = {NULL, NULL, NULL, NULL};
    for (int l = 0; l < L; l++) {
        SEXP c = VECTOR_ELT(chl, l);
        if (!isReal(c) || nrows(c) != n || ncols(c) != m) Rf_error("Cova.haplo.mtx.list[[%d]] must be a numeric matrix of the shape of Geno.mtx", l + 1);
        ch[l] = REAL(c);
    }
    spagxe_params sp; params_from_sexp(prm, &sp);
    SEXP out = PROTECT(allocMatrix(REALSXP, m, 10));
    spagxe_diag diag; memset(&diag, 0, sizeof diag);
    int rc = spagxe_run_localance(ctx, n, m, REAL(G), REAL(H), L, ch, n, &sp, REAL(out), &diag);
    if (rc != SPAGXE_OK) { UNPROTECT(1); Rf_error("%s", spagxe_last_error(ctx)); }
    attach_diag(out, &diag);
    UNPROTECT(1);
    return out;
}

/* ---- several GPUs of the node (SPAGxE_CCT's n.gpus argument) ---- */
SEXP spagxe_r_mgpu_create(SEXP n_gpus) {
    spagxe_mgpu *mg = NULL;
    if (spagxe_mgpu_create(asInteger(n_gpus), NULL, &mg) != SPAGXE_OK) Rf_error("%s", spagxe_mgpu_last_error(NULL));
    SEXP ptr = PROTECT(R_MakeExternalPtr(mg, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, mgpu_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}
SEXP spagxe_r_mgpu_destroy(SEXP ptr) { mgpu_finalizer(ptr); return R_NilValue; }
static spagxe_mgpu *get_mgpu(SEXP ptr) {
    spagxe_mgpu *mg = (spagxe_mgpu *)R_ExternalPtrAddr(ptr);
    if (!mg) Rf_error("spagxe: multi-GPU context already released");
    return mg;
}
SEXP spagxe_r_mgpu_set_inputs(SEXP ptr, SEXP R, SEXP E, SEXP trait, SEXP Y, SEXP cova) {
    spagxe_mgpu *mg = get_mgpu(ptr);
    if (spagxe_mgpu_set_inputs(mg, XLENGTH(R), REAL(R), REAL(E), asInteger(trait), REAL(Y), REAL(cova), ncols(cova)) != SPAGXE_OK)
        Rf_error("%s", spagxe_mgpu_last_error(mg));
    return R_NilValue;
}
SEXP spagxe_r_mgpu_run_cct(SEXP ptr, SEXP geno, SEXP n_, SEXP m_, SEXP prm, SEXP sidx, SEXP nfile, SEXP a2) {
    spagxe_mgpu *mg = get_mgpu(ptr);
    spagxe_geno gd; int64_t *ix = NULL;
    geno_from_sexp(geno, n_, m_, sidx, nfile, a2, &gd, &ix);
    spagxe_params sp; params_from_sexp(prm, &sp);
    SEXP out = PROTECT(allocMatrix(REALSXP, (int)gd.n_snps, 10));
    spagxe_diag diag; memset(&diag, 0, sizeof diag);
    int rc = spagxe_mgpu_run_cct(mg, &gd, 1, &sp, REAL(out), &diag);
    free(ix);
    if (rc != SPAGXE_OK) { UNPROTECT(1); Rf_error("%s", spagxe_mgpu_last_error(mg)); }
    attach_diag(out, &diag);
    UNPROTECT(1);
    return out;
}

/* the other loops on several GPUs: inputs without Wald data (trait 0), the per-analysis extras, one generic run.
 * which: 1 mix (M x 12), 2 plus (M x 8), 3 mix_plus (M x 12) */
SEXP spagxe_r_mgpu_set_vectors(SEXP ptr, SEXP R, SEXP E) {
    spagxe_mgpu *mg = get_mgpu(ptr);
    if (spagxe_mgpu_set_inputs(mg, XLENGTH(R), REAL(R), REAL(E), 0, NULL, NULL, 0) != SPAGXE_OK) Rf_error("%s", spagxe_mgpu_last_error(mg));
    return R_NilValue;
}
SEXP spagxe_r_mgpu_set_wald_survival(SEXP ptr, SEXP time, SEXP event, SEXP cova) {
    spagxe_mgpu *mg = get_mgpu(ptr);
    if (spagxe_mgpu_set_wald_survival(mg, REAL(time), REAL(event), REAL(cova), ncols(cova)) != SPAGXE_OK) Rf_error("%s", spagxe_mgpu_last_error(mg));
    return R_NilValue;
}
SEXP spagxe_r_mgpu_set_pcs(SEXP ptr, SEXP pcs) {
    spagxe_mgpu *mg = get_mgpu(ptr);
    if (spagxe_mgpu_set_pcs(mg, REAL(pcs), ncols(pcs)) != SPAGXE_OK) Rf_error("%s", spagxe_mgpu_last_error(mg));
    return R_NilValue;
}
SEXP spagxe_r_mgpu_set_grm(SEXP ptr, SEXP gi, SEXP gj, SEXP gv, SEXP scalars, SEXP Rnew) {
    spagxe_mgpu *mg = get_mgpu(ptr);
    if (spagxe_mgpu_set_grm(mg, XLENGTH(gv), INTEGER(gi), INTEGER(gj), REAL(gv), REAL(scalars)[0], REAL(scalars)[1], REAL(scalars)[2],
                            REAL(Rnew)) != SPAGXE_OK) Rf_error("%s", spagxe_mgpu_last_error(mg));
    return R_NilValue;
}
SEXP spagxe_r_mgpu_set_sparse_grm(SEXP ptr, SEXP gi, SEXP gj, SEXP gv) {
    spagxe_mgpu *mg = get_mgpu(ptr);
    if (spagxe_mgpu_set_sparse_grm(mg, XLENGTH(gv), INTEGER(gi), INTEGER(gj), REAL(gv)) != SPAGXE_OK) Rf_error("%s", spagxe_mgpu_last_error(mg));
    return R_NilValue;
}
SEXP spagxe_r_mgpu_run(SEXP ptr, SEXP which, SEXP geno, SEXP n_, SEXP m_, SEXP prm, SEXP sidx, SEXP nfile, SEXP a2) {
    spagxe_mgpu *mg = get_mgpu(ptr);
    spagxe_geno gd; int64_t *ix = NULL;
    geno_from_sexp(geno, n_, m_, sidx, nfile, a2, &gd, &ix);
    spagxe_params sp; params_from_sexp(prm, &sp);
    const int w = asInteger(which), ncol = (w == 2) ? 8 : 12;
    SEXP out = PROTECT(allocMatrix(REALSXP, (int)gd.n_snps, ncol));
    spagxe_diag diag; memset(&diag, 0, sizeof diag);
    int rc = (w == 1) ? spagxe_mgpu_run_mix(mg, &gd, 1, &sp, REAL(out), &diag)
           : (w == 2) ? spagxe_mgpu_run_plus(mg, &gd, 1, &sp, REAL(out), &diag)
                      : spagxe_mgpu_run_mix_plus(mg, &gd, 1, &sp, REAL(out), &diag);
    free(ix);
    if (rc != SPAGXE_OK) { UNPROTECT(1); Rf_error("%s", spagxe_mgpu_last_error(mg)); }
    attach_diag(out, &diag);
    UNPROTECT(1);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"spagxe_r_create", (DL_FUNC)&spagxe_r_create, 1},
    {"spagxe_r_destroy", (DL_FUNC)&spagxe_r_destroy, 1},
    {"spagxe_r_set_vectors", (DL_FUNC)&spagxe_r_set_vectors, 3},
    {"spagxe_r_set_wald", (DL_FUNC)&spagxe_r_set_wald, 4},
    {"spagxe_r_set_wald_survival", (DL_FUNC)&spagxe_r_set_wald_survival, 4},
    {"spagxe_r_set_pcs", (DL_FUNC)&spagxe_r_set_pcs, 2},
    {"spagxe_r_set_grm", (DL_FUNC)&spagxe_r_set_grm, 6},
    {"spagxe_r_plus_nullmodel", (DL_FUNC)&spagxe_r_plus_nullmodel, 6},
    {"spagxe_r_set_sparse_grm", (DL_FUNC)&spagxe_r_set_sparse_grm, 4},
    {"spagxe_r_run", (DL_FUNC)&spagxe_r_run, 9},
    {"spagxe_r_run_localance", (DL_FUNC)&spagxe_r_run_localance, 5},
    {"spagxe_r_mgpu_create", (DL_FUNC)&spagxe_r_mgpu_create, 1},
    {"spagxe_r_mgpu_destroy", (DL_FUNC)&spagxe_r_mgpu_destroy, 1},
    {"spagxe_r_mgpu_set_inputs", (DL_FUNC)&spagxe_r_mgpu_set_inputs, 6},
    {"spagxe_r_mgpu_run_cct", (DL_FUNC)&spagxe_r_mgpu_run_cct, 8},
    {"spagxe_r_mgpu_set_vectors", (DL_FUNC)&spagxe_r_mgpu_set_vectors, 3},
    {"spagxe_r_mgpu_set_wald_survival", (DL_FUNC)&spagxe_r_mgpu_set_wald_survival, 4},
    {"spagxe_r_mgpu_set_pcs", (DL_FUNC)&spagxe_r_mgpu_set_pcs, 2},
    {"spagxe_r_mgpu_set_grm", (DL_FUNC)&spagxe_r_mgpu_set_grm, 6},
    {"spagxe_r_mgpu_set_sparse_grm", (DL_FUNC)&spagxe_r_mgpu_set_sparse_grm, 4},
    {"spagxe_r_mgpu_run", (DL_FUNC)&spagxe_r_mgpu_run, 9},
    {NULL, NULL, 0}};

void R_init_spagxe_rshim(DllInfo *dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
