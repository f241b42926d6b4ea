/* Minimal DECLARATIONS of the R C API used by r/frk_b200_shim.c -- test infrastructure so that the shim can be
 * syntax- and type-checked with gcc in a container without R.  Not a substitute for R's headers. */
#ifndef R_STUB_RINTERNALS_H
#define R_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
extern SEXP R_NamesSymbol;
extern SEXP R_NilValue;
#define NILSXP 0
#define LGLSXP 10
#define RAWSXP 24
int TYPEOF(SEXP);
Rboolean Rf_isMatrix(SEXP);
R_xlen_t XLENGTH(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
double *REAL(SEXP);
int *INTEGER(SEXP);
unsigned char *RAW(SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_mkChar(const char *);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
void Rf_error(const char *, ...) __attribute__((noreturn));
#endif
