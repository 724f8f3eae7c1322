/* minir.h -- harness-side entry points of the R stand-in (TEST INFRASTRUCTURE ONLY). */
#ifndef MINIR_H
#define MINIR_H
#include <setjmp.h>
#include "Rdefines.h"

extern jmp_buf minir_jmpbuf;
extern int minir_jmp_armed;
extern char minir_errmsg[4096];
extern char minir_warnmsg[4096];

void minir_set_generation(int gen);     /* 1 = persistent, 2 = per-call scratch */
void minir_persist(SEXP x);            /* promote an object graph to persistent */
void minir_release_scratch(void);       /* what the end of a .Call does in R */
void minir_free_object(SEXP x);         /* demote a persistent object to scratch */
void minir_free_AEbufs(void);

struct minir_memfile { const char *data; long nbytes; long pos; };
SEXP minir_new_filexp(const char *data, long nbytes);
SEXP minir_new_XString(const char *classname, const char *ptr, int length);
SEXP minir_new_XStringSet(const char *classname, int n, const char *const *ptrs,
			  const int *widths);
#endif
