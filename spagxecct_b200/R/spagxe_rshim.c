/* spagxe_rshim.c -- the thin .Call layer between R and libspagxe_cuda.so (include/spagxe.h).
 *
 * Compiled only where R exists (R CMD SHLIB spagxe_rshim.c -L.. -lspagxe_cuda); the build image has no R, so
 * this file is not part of the default build.  It only unwraps SEXPs, allocates the result matrix and turns
 * status codes into R errors; no C++ object is alive when Rf_error() long-jumps.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include "../../include/spagxe.h"

static void ctx_finalizer(SEXP ptr) {
    spagxe_ctx *ctx = (spagxe_ctx *)R_ExternalPtrAddr(ptr);
    if (ctx) { spagxe_ctx_destroy(ctx); R_ClearExternalPtr(ptr); }
}

SEXP spagxe_r_create(SEXP device) {
    spagxe_ctx *ctx = NULL;
    if (spagxe_ctx_create(asInteger(device), &ctx) != SPAGXE_OK) Rf_error("%s", spagxe_last_error(NULL));
    SEXP ptr = PROTECT(R_MakeExternalPtr(ctx, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

static spagxe_ctx *get_ctx(SEXP ptr) {
    spagxe_ctx *ctx = (spagxe_ctx *)R_ExternalPtrAddr(ptr);
    if (!ctx) Rf_error("spagxe: context already released");
    return ctx;
}
#define CHECK(ctx, call) do { if ((call) != SPAGXE_OK) Rf_error("%s", spagxe_last_error(ctx)); } while (0)

SEXP spagxe_r_set_vectors(SEXP ptr, SEXP R, SEXP E) {
    spagxe_ctx *ctx = get_ctx(ptr);
    CHECK(ctx, spagxe_set_vectors(ctx, XLENGTH(R), REAL(R), REAL(E)));
    return R_NilValue;
}
SEXP spagxe_r_set_wald(SEXP ptr, SEXP trait, SEXP Y, SEXP cova) {
    spagxe_ctx *ctx = get_ctx(ptr);
    CHECK(ctx, spagxe_set_wald(ctx, asInteger(trait), REAL(Y), REAL(cova), ncols(cova)));
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

/* geno: numeric/integer matrix (N x M) or a raw vector of packed .bed rows with attributes n, m */
static void geno_from_sexp(SEXP g, SEXP n_, SEXP m_, spagxe_geno *d) {
    d->location = SPAGXE_LOC_HOST;
    if (TYPEOF(g) == RAWSXP) {
        d->kind = SPAGXE_GENO_BED; d->data = RAW(g);
        d->n_samples = (int64_t)asReal(n_); d->n_snps = (int64_t)asReal(m_); d->pitch = (d->n_samples + 3) / 4;
    } else if (TYPEOF(g) == INTSXP) {
        d->kind = SPAGXE_GENO_I32; d->data = INTEGER(g); d->n_samples = nrows(g); d->n_snps = ncols(g); d->pitch = nrows(g);
    } else {
        d->kind = SPAGXE_GENO_F64; d->data = REAL(g); d->n_samples = nrows(g); d->n_snps = ncols(g); d->pitch = nrows(g);
    }
}

/* which: 0 cct (M x 10), 1 mix (M x 12), 2 plus (M x 8); prm = c(epsilon, Cutoff, min.maf, missing.cutoff, pc.cut, neg.cut) */
SEXP spagxe_r_run(SEXP ptr, SEXP which, SEXP geno, SEXP n_, SEXP m_, SEXP prm) {
    spagxe_ctx *ctx = get_ctx(ptr);
    spagxe_geno gd; geno_from_sexp(geno, n_, m_, &gd);
    const double *p = REAL(prm);
    spagxe_params sp = {p[0], p[1], p[2], p[3], p[4], p[5], 0, 0, SPAGXE_LOC_HOST, 0};
    const int w = asInteger(which), ncol = (w == 0) ? 10 : (w == 1 ? 12 : 8);
    SEXP out = PROTECT(allocMatrix(REALSXP, (int)gd.n_snps, ncol));
    spagxe_diag diag;
    int rc = (w == 0) ? spagxe_run_cct(ctx, &gd, &sp, REAL(out), &diag)
           : (w == 1) ? spagxe_run_mix(ctx, &gd, &sp, REAL(out), &diag)
                      : spagxe_run_plus(ctx, &gd, &sp, REAL(out), &diag);
    UNPROTECT(1);
    if (rc != SPAGXE_OK) Rf_error("%s", spagxe_last_error(ctx));
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"spagxe_r_create", (DL_FUNC)&spagxe_r_create, 1},
    {"spagxe_r_set_vectors", (DL_FUNC)&spagxe_r_set_vectors, 3},
    {"spagxe_r_set_wald", (DL_FUNC)&spagxe_r_set_wald, 4},
    {"spagxe_r_set_pcs", (DL_FUNC)&spagxe_r_set_pcs, 2},
    {"spagxe_r_set_grm", (DL_FUNC)&spagxe_r_set_grm, 6},
    {"spagxe_r_run", (DL_FUNC)&spagxe_r_run, 6},
    {NULL, NULL, 0}};

void R_init_spagxe_rshim(DllInfo *dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
