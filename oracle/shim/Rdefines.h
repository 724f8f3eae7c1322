/*
 * minir -- a tiny stand-in for the R C API (TEST INFRASTRUCTURE ONLY).
 *
 * Purpose: let the reference's hot-path C files (match_pdict*.c & friends under
 * /root/reference/src) compile *unmodified, in place* into oracle/_ref/ so the
 * real reference implementation can be executed as the parity oracle and the
 * CPU baseline.  Nothing in the product (biostrings_b200/, include/) includes
 * or links this.  It implements only what those files use.
 */
#ifndef MINIR_RDEFINES_H
#define MINIR_RDEFINES_H

#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include <limits.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
	NILSXP = 0, SYMSXP = 1, CHARSXP = 9, LGLSXP = 10, INTSXP = 13,
	REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22, RAWSXP = 24,
	S4SXP = 25
};

typedef struct minir_sexp *SEXP;
typedef int Rboolean;
typedef ptrdiff_t R_xlen_t;

struct minir_slot {
	char *name;
	SEXP value;
	struct minir_slot *next;
};

struct minir_sexp {
	int type;
	int gen;                 /* 1 = persistent, 2 = per-call scratch */
	long length;
	void *data;              /* int* / double* / unsigned char* / char* / SEXP* */
	SEXP names;
	char *classname;         /* S4 objects */
	struct minir_slot *slots;/* S4 objects */
	void *xptr;              /* external pointers / shim payload */
	SEXP tag, prot;          /* external pointers */
	void (*fin)(struct minir_sexp *);  /* C finalizer of an external pointer (run when it is freed) */
	struct minir_sexp *prev, *next;
};

extern struct minir_sexp minir_nil;
#define R_NilValue (&minir_nil)

#define NA_INTEGER INT_MIN
#define NA_LOGICAL INT_MIN
SEXP minir_NA_STRING(void);
#define NA_STRING minir_NA_STRING()
#define TRUE 1
#define FALSE 0

SEXP minir_alloc(int type, long length);
SEXP Rf_allocVector(int type, long length);
SEXP Rf_allocMatrix(int type, int nrow, int ncol);
#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix

#define NEW_INTEGER(n)   minir_alloc(INTSXP, (n))
#define NEW_LOGICAL(n)   minir_alloc(LGLSXP, (n))
#define NEW_NUMERIC(n)   minir_alloc(REALSXP, (n))
#define NEW_RAW(n)       minir_alloc(RAWSXP, (n))
#define NEW_LIST(n)      minir_alloc(VECSXP, (n))
#define NEW_CHARACTER(n) minir_alloc(STRSXP, (n))

#define INTEGER(x) ((int *) (x)->data)
#define LOGICAL(x) ((int *) (x)->data)
#define REAL(x)    ((double *) (x)->data)
#define RAW(x)     ((unsigned char *) (x)->data)
#define LENGTH(x)  ((int) (x)->length)
#define XLENGTH(x) ((x)->length)
#define TYPEOF(x)  ((x)->type)
#define IS_INTEGER(x) ((x)->type == INTSXP)
#define IS_NUMERIC(x) ((x)->type == REALSXP)
#define IS_LOGICAL(x) ((x)->type == LGLSXP)
#define IS_CHARACTER(x) ((x)->type == STRSXP)
#define isNull(x) ((x) == R_NilValue)

SEXP VECTOR_ELT(SEXP x, long i);
SEXP SET_VECTOR_ELT(SEXP x, long i, SEXP v);
#define SET_ELEMENT(x, i, v) SET_VECTOR_ELT(x, i, v)
SEXP STRING_ELT(SEXP x, long i);
void SET_STRING_ELT(SEXP x, long i, SEXP v);
const char *minir_CHAR(SEXP x);
#define CHAR(x) minir_CHAR(x)
SEXP Rf_mkChar(const char *s);
#define mkChar Rf_mkChar
SEXP Rf_install(const char *s);
#define install Rf_install
SEXP Rf_mkString(const char *s);
#define mkString Rf_mkString

SEXP GET_SLOT(SEXP x, SEXP name);
SEXP SET_SLOT(SEXP x, SEXP name, SEXP value);
void SET_NAMES(SEXP x, SEXP names);
SEXP GET_NAMES(SEXP x);
/* dim / dimnames attributes (src/letter_frequency.c); kept beside the slots */
void SET_DIM(SEXP x, SEXP dim);
SEXP GET_DIM(SEXP x);
void SET_DIMNAMES(SEXP x, SEXP dimnames);
SEXP GET_DIMNAMES(SEXP x);
SEXP Rf_alloc3DArray(int type, int nrow, int ncol, int nface);
#define alloc3DArray Rf_alloc3DArray
int Rf_asLogical(SEXP x);
#define asLogical Rf_asLogical
SEXP Rf_list2(SEXP a, SEXP b);
SEXP Rf_list3(SEXP a, SEXP b, SEXP c);
#define list2 Rf_list2
#define list3 Rf_list3
SEXP MAKE_CLASS(const char *name);
SEXP NEW_OBJECT(SEXP classdef);

SEXP Rf_ScalarInteger(int v);
SEXP Rf_ScalarLogical(int v);
#define ScalarInteger Rf_ScalarInteger
#define ScalarLogical Rf_ScalarLogical
SEXP Rf_duplicate(SEXP x);
#define duplicate Rf_duplicate

#define PROTECT(x) (x)
#define UNPROTECT(n) ((void) 0)

SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
SEXP R_ExternalPtrTag(SEXP x);
SEXP R_ExternalPtrProtected(SEXP x);
void *R_ExternalPtrAddr(SEXP x);
void R_SetExternalPtrAddr(SEXP x, void *p);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP x, R_CFinalizer_t fun, Rboolean onexit);
SEXP R_lsInternal(SEXP env, Rboolean all);

char *S_alloc(long n, int size);
char *R_alloc(size_t n, int size);

void Rf_error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
void Rf_warning(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void Rprintf(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
#define error Rf_error
#define warning Rf_warning

void R_CheckUserInterrupt(void);

#ifdef __cplusplus
}
#endif
#endif
