/* Minimal stand-in for R's C API, ONLY for syntax-checking r_shim.c where R is not installed (this image has no R).
 * Declarations follow the public R API names; nothing here is linked or run. */
#ifndef FLX_STUB_RINTERNALS_H
#define FLX_STUB_RINTERNALS_H
#include <stddef.h>
#include <stdio.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef void* (*DL_FUNC)(void);
typedef struct _DllInfo DllInfo;
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
enum { INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19 };
extern SEXP R_NamesSymbol;
extern double R_NaReal;
#define NA_REAL R_NaReal
#define FALSE 0
#define TRUE 1
#define ISNAN(x) ((x) != (x))
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP Rf_protect(SEXP); void Rf_unprotect(int);
R_xlen_t XLENGTH(SEXP); double* REAL(SEXP); int* INTEGER(SEXP); const char* CHAR(SEXP); SEXP STRING_ELT(SEXP, R_xlen_t);
SEXP Rf_allocVector(unsigned, R_xlen_t); SEXP Rf_allocMatrix(unsigned, int, int);
void SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP); void SET_STRING_ELT(SEXP, R_xlen_t, SEXP); SEXP Rf_mkChar(const char*);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP); SEXP Rf_ScalarReal(double);
int Rf_asInteger(SEXP); double Rf_asReal(SEXP); int Rf_asLogical(SEXP);
void Rf_error(const char*, ...) __attribute__((noreturn));
int R_registerRoutines(DllInfo*, const void*, const R_CallMethodDef*, const void*, const void*);
int R_useDynamicSymbols(DllInfo*, int);
#endif
