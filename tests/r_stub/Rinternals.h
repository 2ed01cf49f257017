/* TEST INFRASTRUCTURE ONLY -- see R.h in this directory. */
#ifndef AQ_RINTERNALS_STUB_H
#define AQ_RINTERNALS_STUB_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
typedef unsigned char Rbyte;
#define TRUE 1
#define FALSE 0
#define VECSXP 19
#define STRSXP 16
#define REALSXP 14
#define INTSXP 13
#define RAWSXP 24
extern SEXP R_NilValue, R_NamesSymbol;
extern double R_NaReal;
#define NA_REAL R_NaReal
void Rf_error(const char*, ...) __attribute__((noreturn));
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
int Rf_isNull(SEXP), Rf_isReal(SEXP), Rf_isInteger(SEXP), Rf_isMatrix(SEXP);
int Rf_nrows(SEXP), Rf_ncols(SEXP), Rf_length(SEXP), Rf_asInteger(SEXP), Rf_asLogical(SEXP), TYPEOF(SEXP);
R_xlen_t XLENGTH(SEXP);
double Rf_asReal(SEXP);
double* REAL(SEXP);
int* INTEGER(SEXP);
Rbyte* RAW(SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t), Rf_protect(SEXP), Rf_mkChar(const char*), Rf_ScalarReal(double), Rf_install(const char*);
SEXP VECTOR_ELT(SEXP, R_xlen_t), SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP), Rf_setAttrib(SEXP, SEXP, SEXP);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP), Rf_unprotect(int);
char* R_alloc(size_t, int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#endif
