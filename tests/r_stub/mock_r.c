/* A small mock of the part of R's C API that r/src/frk_b200_shim.c uses -- TEST INFRASTRUCTURE.
 * It lets the shim be linked and EXECUTED without R: vectors / matrices / lists as plain heap objects, PROTECT as a
 * no-op (nothing is ever collected; mock_free_all releases everything), Rf_error as a longjmp back into mock_dotcall,
 * R_registerRoutines recording the table so that calls go through the registered names and arities the way .Call does.
 * Declarations: tests/r_stub/Rinternals.h, tests/r_stub/R_ext/Rdynload.h. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "R_ext/Rdynload.h"

struct SEXPREC {
    int type;
    R_xlen_t len;
    int nrow, ncol;     /* -1: no dim attribute */
    void *data;         /* double[] / int[] / SEXP[] / char[] */
    SEXP names;
    struct SEXPREC *next;
};

static struct SEXPREC nil_obj = {NILSXP, 0, -1, -1, NULL, NULL, NULL};
static struct SEXPREC names_sym = {1, 0, -1, -1, NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj;
SEXP R_NamesSymbol = &names_sym;

static SEXP all_objects = NULL;
static jmp_buf err_jmp;
static int err_armed = 0;
static char err_msg[1024];
static const R_CallMethodDef *call_table = NULL;

static SEXP new_obj(int type, R_xlen_t len, size_t elt) {
    SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
    s->type = type; s->len = len; s->nrow = s->ncol = -1;
    s->data = calloc(len > 0 ? (size_t)len : 1, elt);
    s->names = R_NilValue;
    s->next = all_objects;
    all_objects = s;
    return s;
}
static size_t elt_size(unsigned type) {
    return type == REALSXP ? sizeof(double) : (type == INTSXP || type == LGLSXP) ? sizeof(int) : type == RAWSXP ? 1 : sizeof(SEXP);
}

int TYPEOF(SEXP x) { return x->type; }
Rboolean Rf_isMatrix(SEXP x) { return x->nrow >= 0 ? TRUE : FALSE; }
R_xlen_t XLENGTH(SEXP x) { return x->len; }
int Rf_nrows(SEXP x) { return x->nrow >= 0 ? x->nrow : (int)x->len; }
int Rf_ncols(SEXP x) { return x->ncol >= 0 ? x->ncol : 1; }
double *REAL(SEXP x) {
    if (x->type != REALSXP) Rf_error("REAL() can only be applied to a 'numeric', not a type-%d object", x->type);
    return (double *)x->data;
}
int *INTEGER(SEXP x) {
    if (x->type != INTSXP && x->type != LGLSXP) Rf_error("INTEGER() can only be applied to a 'integer', not a type-%d object", x->type);
    return (int *)x->data;
}
unsigned char *RAW(SEXP x) {
    if (x->type != RAWSXP) Rf_error("RAW() can only be applied to a 'raw', not a type-%d object", x->type);
    return (unsigned char *)x->data;
}
SEXP Rf_allocVector(unsigned type, R_xlen_t n) { return new_obj((int)type, n, elt_size(type)); }
SEXP Rf_allocMatrix(unsigned type, int nr, int nc) {
    SEXP s = new_obj((int)type, (R_xlen_t)nr * nc, elt_size(type));
    s->nrow = nr; s->ncol = nc;
    return s;
}
SEXP Rf_mkChar(const char *c) {
    SEXP s = new_obj(9 /* CHARSXP */, (R_xlen_t)strlen(c) + 1, 1);
    memcpy(s->data, c, strlen(c) + 1);
    return s;
}
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP *)x->data)[i] = v; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP *)x->data)[i] = v; return v; }
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP v) { if (name == R_NamesSymbol) x->names = v; return v; }
SEXP Rf_protect(SEXP s) { return s; }
void Rf_unprotect(int n) { (void)n; }
int Rf_asInteger(SEXP x) { return x->type == REALSXP ? (int)((double *)x->data)[0] : ((int *)x->data)[0]; }
int Rf_asLogical(SEXP x) { return Rf_asInteger(x) != 0; }
double Rf_asReal(SEXP x) { return x->type == REALSXP ? ((double *)x->data)[0] : (double)((int *)x->data)[0]; }
void Rf_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_msg, sizeof err_msg, fmt, ap);
    va_end(ap);
    if (err_armed) longjmp(err_jmp, 1);
    fprintf(stderr, "mock R: error outside a call: %s\n", err_msg);
    abort();
}
int R_registerRoutines(DllInfo *dll, const void *c, const R_CallMethodDef *call, const void *f, const void *e) {
    (void)dll; (void)c; (void)f; (void)e;
    call_table = call;
    return 1;
}
Rboolean R_useDynamicSymbols(DllInfo *dll, Rboolean v) { (void)dll; return v; }

/* ---- what the Python harness drives ------------------------------------------------------------------------- */
void R_init_autoFRK(DllInfo *dll);
void mock_load(void) { R_init_autoFRK(NULL); }
int mock_nroutines(void) {
    int n = 0;
    while (call_table && call_table[n].name) ++n;
    return n;
}
const char *mock_routine_name(int i) { return call_table[i].name; }
int mock_routine_nargs(int i) { return call_table[i].numArgs; }
const char *mock_last_error(void) { return err_msg; }

SEXP mock_real_matrix(const double *v, int nr, int nc) {
    SEXP s = Rf_allocMatrix(REALSXP, nr, nc);
    memcpy(s->data, v, sizeof(double) * (size_t)nr * nc);
    return s;
}
SEXP mock_real_vector(const double *v, R_xlen_t n) {
    SEXP s = Rf_allocVector(REALSXP, n);
    memcpy(s->data, v, sizeof(double) * (size_t)n);
    return s;
}
SEXP mock_int_vector(const int *v, R_xlen_t n, int logical) {
    SEXP s = Rf_allocVector(logical ? LGLSXP : INTSXP, n);
    memcpy(s->data, v, sizeof(int) * (size_t)n);
    return s;
}
SEXP mock_int_matrix(const int *v, int nr, int nc) {
    SEXP s = Rf_allocMatrix(INTSXP, nr, nc);
    memcpy(s->data, v, sizeof(int) * (size_t)nr * nc);
    return s;
}
SEXP mock_nil(void) { return R_NilValue; }
int mock_type(SEXP s) { return s->type; }
long long mock_length(SEXP s) { return (long long)s->len; }
int mock_nrow(SEXP s) { return s->nrow; }
int mock_ncol(SEXP s) { return s->ncol; }
const double *mock_real_ptr(SEXP s) { return (const double *)s->data; }
SEXP mock_list_elt(SEXP s, int i) { return ((SEXP *)s->data)[i]; }
const char *mock_list_name(SEXP s, int i) {
    if (s->names == R_NilValue) return "";
    return (const char *)((SEXP *)s->names->data)[i]->data;
}

/* .Call(name, args...): looks the routine up in the registered table and enforces its arity like R does */
typedef SEXP (*F1)(SEXP); typedef SEXP (*F2)(SEXP, SEXP); typedef SEXP (*F3)(SEXP, SEXP, SEXP);
typedef SEXP (*F4)(SEXP, SEXP, SEXP, SEXP); typedef SEXP (*F5)(SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP); typedef SEXP (*F7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
int mock_dotcall(const char *name, int nargs, SEXP *a, SEXP *result) {
    const R_CallMethodDef *r = call_table;
    while (r && r->name && strcmp(r->name, name)) ++r;
    if (!r || !r->name) { snprintf(err_msg, sizeof err_msg, "\"%s\" not available for .Call() for package \"autoFRK\"", name); return 1; }
    if (r->numArgs != nargs) {
        snprintf(err_msg, sizeof err_msg, "Incorrect number of arguments (%d), expecting %d for '%s'", nargs, r->numArgs, name);
        return 1;
    }
    err_armed = 1;
    if (setjmp(err_jmp)) { err_armed = 0; return 1; }
    void *f = (void *)r->fun;
    SEXP out = R_NilValue;
    switch (nargs) {
        case 1: out = ((F1)f)(a[0]); break;
        case 2: out = ((F2)f)(a[0], a[1]); break;
        case 3: out = ((F3)f)(a[0], a[1], a[2]); break;
        case 4: out = ((F4)f)(a[0], a[1], a[2], a[3]); break;
        case 5: out = ((F5)f)(a[0], a[1], a[2], a[3], a[4]); break;
        case 6: out = ((F6)f)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
        case 7: out = ((F7)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
        case 8: out = ((F8)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
        default: err_armed = 0; snprintf(err_msg, sizeof err_msg, "mock R: %d arguments unsupported", nargs); return 1;
    }
    err_armed = 0;
    *result = out;
    return 0;
}
void mock_free_all(void) {
    while (all_objects) {
        SEXP n = all_objects->next;
        free(all_objects->data);
        free(all_objects);
        all_objects = n;
    }
}
