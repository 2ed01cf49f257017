/* TEST INFRASTRUCTURE ONLY: a miniature runtime behind the stand-in R API declarations of this directory, so that the
 * SHIPPED .Call shim (bindings/R/atlasqtl_b200_shim.c) can be compiled, loaded and EXECUTED in an image without R:
 * SEXPs are small tagged records, Rf_error() longjmps back to the caller of rstub_call(), PROTECT / UNPROTECT are
 * counted so that a stack imbalance of a wrapper is detected, R_registerRoutines() keeps the table R_init_atlasqtl
 * hands over.  Semantics follow "Writing R Extensions" for the few entry points the shim uses; nothing here comes from
 * R's sources.  Driven from Python through ctypes (tests/r_shim_real.py). */
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R.h"
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

#define NILSXP 0
#define SYMSXP 1
#define CHARSXP 9
#define LGLSXP 10
#define EXTPTRSXP 22
#define NA_INT (-2147483647 - 1)

struct SEXPREC {
  int type;
  R_xlen_t len;
  void* data;        /* element storage (double / int / Rbyte / SEXP) or the string of a CHARSXP / SYMSXP */
  int owns;          /* data was malloc'ed here (0: borrowed from the caller, e.g. a NumPy buffer) */
  int nrow, ncol;    /* dim attribute; nrow < 0: none */
  SEXP names;
  SEXP attr_sym[4], attr_val[4];
  int nattr;
  void* ext;         /* EXTPTRSXP address */
  R_CFinalizer_t fin;
};

static struct SEXPREC nil_rec = {NILSXP, 0, NULL, 0, -1, -1, NULL, {0}, {0}, 0, NULL, NULL};
static struct SEXPREC names_rec = {SYMSXP, 0, (void*)"names", 0, -1, -1, NULL, {0}, {0}, 0, NULL, NULL};
SEXP R_NilValue = &nil_rec;
SEXP R_NamesSymbol = &names_rec;
double R_NaReal;

static jmp_buf* cur_jmp = NULL;
static char err_msg[1024];
static int protect_depth = 0;
static void* ralloc_list[64];
static int ralloc_n = 0;
static const R_CallMethodDef* routines = NULL;
static SEXP symbols[32];
static int nsymbols = 0;

__attribute__((constructor)) static void rstub_init(void) {
  union { uint64_t u; double d; } na;
  na.u = 0x7FF00000000007A2ULL;   /* NA_real_: a NaN whose low word is 1954 */
  R_NaReal = na.d;
}

static SEXP new_sexp(int type, R_xlen_t len) {
  SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
  s->type = type;
  s->len = len;
  s->nrow = s->ncol = -1;
  s->names = R_NilValue;
  return s;
}

void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_msg, sizeof err_msg, fmt, ap);
  va_end(ap);
  if (!cur_jmp) { fprintf(stderr, "Rf_error outside rstub_call: %s\n", err_msg); abort(); }
  longjmp(*cur_jmp, 1);
}

SEXP Rf_allocVector(unsigned int type, R_xlen_t n) {
  size_t w = type == REALSXP ? sizeof(double) : type == INTSXP || type == LGLSXP ? sizeof(int)
             : type == RAWSXP ? 1 : type == VECSXP || type == STRSXP ? sizeof(SEXP) : 0;
  if (!w) Rf_error("allocVector: type %u is not supported by the stub", type);
  SEXP s = new_sexp((int)type, n);
  s->data = calloc((size_t)(n > 0 ? n : 1), w);
  s->owns = 1;
  if (type == VECSXP || type == STRSXP)
    for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
  return s;
}
SEXP Rf_protect(SEXP s) { ++protect_depth; return s; }
void Rf_unprotect(int n) { protect_depth -= n; }
SEXP Rf_mkChar(const char* c) {
  SEXP s = new_sexp(CHARSXP, (R_xlen_t)strlen(c));
  s->data = strdup(c);
  s->owns = 1;
  return s;
}
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); ((double*)s->data)[0] = v; return s; }
SEXP Rf_install(const char* name) {
  if (!strcmp(name, "names")) return R_NamesSymbol;
  for (int i = 0; i < nsymbols; ++i)
    if (!strcmp((const char*)symbols[i]->data, name)) return symbols[i];
  SEXP s = new_sexp(SYMSXP, 0);
  s->data = strdup(name);
  s->owns = 1;
  if (nsymbols < 32) symbols[nsymbols++] = s;
  return s;
}
char* R_alloc(size_t n, int size) {
  void* p = calloc(n ? n : 1, (size_t)size);
  if (ralloc_n < 64) ralloc_list[ralloc_n++] = p;
  return (char*)p;
}

int TYPEOF(SEXP s) { return s->type; }
R_xlen_t XLENGTH(SEXP s) { return s->len; }
int Rf_length(SEXP s) { return (int)s->len; }
int Rf_isNull(SEXP s) { return s->type == NILSXP; }
int Rf_isReal(SEXP s) { return s->type == REALSXP; }
int Rf_isInteger(SEXP s) { return s->type == INTSXP; }
int Rf_isMatrix(SEXP s) { return s->nrow >= 0; }
int Rf_nrows(SEXP s) { return s->nrow >= 0 ? s->nrow : (int)s->len; }
int Rf_ncols(SEXP s) { return s->nrow >= 0 ? s->ncol : 1; }
double* REAL(SEXP s) { if (s->type != REALSXP) Rf_error("REAL() can only be applied to a 'numeric', not a type %d", s->type); return (double*)s->data; }
int* INTEGER(SEXP s) { if (s->type != INTSXP && s->type != LGLSXP) Rf_error("INTEGER() can only be applied to a 'integer', not a type %d", s->type); return (int*)s->data; }
Rbyte* RAW(SEXP s) { if (s->type != RAWSXP) Rf_error("RAW() can only be applied to a 'raw', not a type %d", s->type); return (Rbyte*)s->data; }
double Rf_asReal(SEXP s) {
  if (s->len < 1) return R_NaReal;
  if (s->type == REALSXP) return ((double*)s->data)[0];
  if (s->type == INTSXP || s->type == LGLSXP) { int v = ((int*)s->data)[0]; return v == NA_INT ? R_NaReal : (double)v; }
  return R_NaReal;
}
int Rf_asInteger(SEXP s) {
  if (s->len < 1) return NA_INT;
  if (s->type == INTSXP || s->type == LGLSXP) return ((int*)s->data)[0];
  if (s->type == REALSXP) { double v = ((double*)s->data)[0]; return v != v || v > 2147483647.0 || v <= -2147483648.0 ? NA_INT : (int)v; }
  return NA_INT;
}
int Rf_asLogical(SEXP s) {
  if (s->len < 1) return NA_INT;
  if (s->type == LGLSXP) return ((int*)s->data)[0];
  if (s->type == INTSXP) { int v = ((int*)s->data)[0]; return v == NA_INT ? NA_INT : v != 0; }
  if (s->type == REALSXP) { double v = ((double*)s->data)[0]; return v != v ? NA_INT : v != 0; }
  return NA_INT;
}
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) {
  if (s->type != VECSXP || i < 0 || i >= s->len) Rf_error("VECTOR_ELT: not a list or index out of range");
  return ((SEXP*)s->data)[i];
}
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) {
  if (s->type != VECSXP || i < 0 || i >= s->len) Rf_error("SET_VECTOR_ELT: not a list or index out of range");
  ((SEXP*)s->data)[i] = v;
  return v;
}
void SET_STRING_ELT(SEXP s, R_xlen_t i, SEXP v) {
  if (s->type != STRSXP || i < 0 || i >= s->len || v->type != CHARSXP) Rf_error("SET_STRING_ELT: bad arguments");
  ((SEXP*)s->data)[i] = v;
}
SEXP Rf_setAttrib(SEXP s, SEXP sym, SEXP v) {
  if (sym == R_NamesSymbol) { s->names = v; return v; }
  for (int i = 0; i < s->nattr; ++i)
    if (s->attr_sym[i] == sym) { s->attr_val[i] = v; return v; }
  if (s->nattr >= 4) Rf_error("setAttrib: the stub keeps four attributes per object");
  s->attr_sym[s->nattr] = sym;
  s->attr_val[s->nattr++] = v;
  return v;
}

SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) { (void)tag; (void)prot; SEXP s = new_sexp(EXTPTRSXP, 1); s->ext = p; return s; }
void* R_ExternalPtrAddr(SEXP s) { return s->type == EXTPTRSXP ? s->ext : NULL; }
void R_ClearExternalPtr(SEXP s) { s->ext = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t f, Rboolean onexit) { (void)onexit; s->fin = f; }

int R_registerRoutines(DllInfo* dll, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
  (void)dll; (void)c; (void)f; (void)e;
  routines = call;
  return 1;
}
static int dynamic_symbols = 1;
int R_useDynamicSymbols(DllInfo* dll, int v) { (void)dll; int old = dynamic_symbols; dynamic_symbols = v; return old; }

/* ------------------------------------------------------------------------------------------------ driver side */
int rstub_routine(int i, const char** name, DL_FUNC* fun, int* nargs) {
  if (!routines) return 0;
  for (int k = 0; k <= i; ++k) if (!routines[k].name) return 0;
  *name = routines[i].name; *fun = routines[i].fun; *nargs = routines[i].numArgs;
  return 1;
}
int rstub_dynamic_symbols(void) { return dynamic_symbols; }

/* wrap caller-owned storage (no copy): what R does when it hands the payload of a vector to .Call */
SEXP rstub_wrap(int type, void* data, R_xlen_t len, int nrow, int ncol) {
  SEXP s = new_sexp(type, len);
  s->data = data;
  s->nrow = nrow;
  s->ncol = ncol;
  return s;
}
SEXP rstub_nil(void) { return R_NilValue; }
int rstub_dim(SEXP s, int which) { return which == 0 ? s->nrow : s->ncol; }
void* rstub_data(SEXP s) { return s->data; }
SEXP rstub_names(SEXP s) { return s->names; }
const char* rstub_string(SEXP strsxp, R_xlen_t i) { return (const char*)((SEXP*)strsxp->data)[i]->data; }
SEXP rstub_attr(SEXP s, const char* name) {
  for (int i = 0; i < s->nattr; ++i)
    if (!strcmp((const char*)s->attr_sym[i]->data, name)) return s->attr_val[i];
  return R_NilValue;
}
void rstub_finalize(SEXP s) { if (s->type == EXTPTRSXP && s->fin) s->fin(s); }

typedef SEXP (*F0)(void);
/* .Call(fn, args...): returns 0 and *out, or 1 with the Rf_error() message in msg.  *imbalance = PROTECT depth left
 * behind by a call that returned normally (R reports "stack imbalance in .Call" for a non-zero value). */
int rstub_call(DL_FUNC fn, int nargs, SEXP* a, SEXP* out, char* msg, int msglen, int* imbalance) {
  jmp_buf env, *saved = cur_jmp;
  const int depth0 = protect_depth;
  volatile int rc = 0;
  cur_jmp = &env;
  if (setjmp(env)) {
    snprintf(msg, (size_t)msglen, "%s", err_msg);
    protect_depth = depth0;      /* R unwinds the protect stack on error */
    rc = 1;
  } else {
    SEXP r;
    switch (nargs) {
      case 1: r = ((SEXP(*)(SEXP))fn)(a[0]); break;
      case 2: r = ((SEXP(*)(SEXP, SEXP))fn)(a[0], a[1]); break;
      case 3: r = ((SEXP(*)(SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2]); break;
      case 4: r = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2], a[3]); break;
      case 5: r = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2], a[3], a[4]); break;
      case 6: r = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
      case 15: r = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(
                   a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14]); break;
      case 16: r = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(
                   a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15]); break;
      default: snprintf(msg, (size_t)msglen, "rstub_call: %d arguments are not wired", nargs); cur_jmp = saved; return 1;
    }
    *out = r;
    *imbalance = protect_depth - depth0;
    protect_depth = depth0;
  }
  cur_jmp = saved;
  for (int i = 0; i < ralloc_n; ++i) free(ralloc_list[i]);   /* R_alloc memory lives until the end of .Call */
  ralloc_n = 0;
  return rc;
}
