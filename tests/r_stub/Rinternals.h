/* Minimal declarations of the R C API used by spagxecct_b200/R/spagxe_rshim.c, for a -fsyntax-only check in an image without R. */
#include "R.h"
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
extern SEXP R_NilValue, R_NamesSymbol;
enum { INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, RAWSXP = 24 };
void *R_ExternalPtrAddr(SEXP); void R_ClearExternalPtr(SEXP); SEXP R_MakeExternalPtr(void *, SEXP, SEXP);
void R_RegisterCFinalizerEx(SEXP, void (*)(SEXP), Rboolean);
void Rf_error(const char *, ...); int asInteger(SEXP); double asReal(SEXP); R_xlen_t XLENGTH(SEXP);
double *REAL(SEXP); int *INTEGER(SEXP); unsigned char *RAW(SEXP); int TYPEOF(SEXP); int nrows(SEXP); int ncols(SEXP);
SEXP allocVector(int, R_xlen_t); SEXP allocMatrix(int, int, int); SEXP PROTECT(SEXP); void UNPROTECT(int);
void SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP); void SET_STRING_ELT(SEXP, R_xlen_t, SEXP); SEXP mkChar(const char *);
SEXP VECTOR_ELT(SEXP, R_xlen_t); int isReal(SEXP);
SEXP setAttrib(SEXP, SEXP, SEXP); SEXP install(const char *);
