/*
 * minir.c -- implementation of the tiny R / S4Vectors / IRanges / XVector
 * stand-in declared in the headers of this directory.
 *
 * TEST INFRASTRUCTURE ONLY.  It exists so that the reference's own hot-path C
 * (compiled unmodified from /root/reference/src into oracle/_ref/) can run
 * without R.  The product never links it.
 *
 * Memory model: every SEXP carries a generation (1 = persistent, 2 = per-call
 * scratch).  Storing a scratch object into a persistent container promotes it
 * (this is what keeps ACtree2 node blocks, which the reference allocates lazily
 * *during* a search, alive across calls).  minir_release_scratch() frees all
 * scratch SEXPs, S_alloc() memory and auto-extending buffers, like the end of
 * a .Call does in R.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <setjmp.h>

#include "Rdefines.h"
#include "S4Vectors_interface.h"
#include "IRanges_interface.h"
#include "XVector_interface.h"
#include "minir.h"

struct minir_sexp minir_nil = { NILSXP, 1, 0, NULL, &minir_nil, NULL, NULL, NULL,
				&minir_nil, &minir_nil, NULL, NULL };
static struct minir_sexp minir_na_string = { CHARSXP, 1, 2, (void *) "NA", &minir_nil,
				NULL, NULL, NULL, &minir_nil, &minir_nil, NULL, NULL };
SEXP minir_NA_STRING(void) { return &minir_na_string; }

static SEXP all_head = NULL;	/* doubly linked list of live SEXPs */
static int cur_gen = 1;

/* ---- error handling -------------------------------------------------- */

jmp_buf minir_jmpbuf;
int minir_jmp_armed = 0;
char minir_errmsg[4096];
char minir_warnmsg[4096];

void Rf_error(const char *fmt, ...)
{
	va_list ap;

	va_start(ap, fmt);
	vsnprintf(minir_errmsg, sizeof(minir_errmsg), fmt, ap);
	va_end(ap);
	if (minir_jmp_armed)
		longjmp(minir_jmpbuf, 1);
	fprintf(stderr, "minir: uncaught error: %s\n", minir_errmsg);
	abort();
}

void Rf_warning(const char *fmt, ...)
{
	va_list ap;

	va_start(ap, fmt);
	vsnprintf(minir_warnmsg, sizeof(minir_warnmsg), fmt, ap);
	va_end(ap);
}

void Rprintf(const char *fmt, ...)
{
	va_list ap;

	va_start(ap, fmt);
	vfprintf(stdout, fmt, ap);
	va_end(ap);
}

void R_CheckUserInterrupt(void) {}

static void *xmalloc(size_t n)
{
	void *p = malloc(n ? n : 1);

	if (p == NULL)
		Rf_error("minir: out of memory (%zu bytes)", n);
	return p;
}

/* ---- scratch (non-SEXP) memory --------------------------------------- */

struct scratch_blk { struct scratch_blk *next; };
static struct scratch_blk *scratch_head = NULL;

static void *scratch_alloc(size_t n, int zero)
{
	struct scratch_blk *b = xmalloc(sizeof(*b) + 16 + n);

	b->next = scratch_head;
	scratch_head = b;
	if (zero)
		memset((char *) b + sizeof(*b) + 16 - sizeof(*b) % 16, 0, n);
	return (char *) b + sizeof(*b) + 16 - sizeof(*b) % 16;
}

char *S_alloc(long n, int size) { return scratch_alloc((size_t) n * size, 1); }
char *R_alloc(size_t n, int size) { return scratch_alloc(n * size, 0); }

/* ---- SEXP allocation -------------------------------------------------- */

static size_t elt_size(int type)
{
	switch (type) {
	case LGLSXP: case INTSXP: return sizeof(int);
	case REALSXP: return sizeof(double);
	case RAWSXP: case CHARSXP: case SYMSXP: return 1;
	case STRSXP: case VECSXP: return sizeof(SEXP);
	}
	return 0;
}

SEXP minir_alloc(int type, long length)
{
	SEXP x = xmalloc(sizeof(*x));
	long i;

	memset(x, 0, sizeof(*x));
	x->type = type;
	x->gen = cur_gen;
	x->length = length;
	x->names = x->tag = x->prot = R_NilValue;
	if (elt_size(type) != 0) {
		x->data = xmalloc(elt_size(type) * (size_t) (length + 1));
		if (type == STRSXP || type == VECSXP)
			for (i = 0; i < length; i++)
				((SEXP *) x->data)[i] = R_NilValue;
	}
	x->next = all_head;
	if (all_head != NULL)
		all_head->prev = x;
	all_head = x;
	return x;
}

SEXP Rf_allocVector(int type, long length) { return minir_alloc(type, length); }

static SEXP new_dim(int n, int d0, int d1, int d2)
{
	SEXP dim = minir_alloc(INTSXP, n);

	((int *) dim->data)[0] = d0;
	((int *) dim->data)[1] = d1;
	if (n > 2)
		((int *) dim->data)[2] = d2;
	return dim;
}

SEXP Rf_allocMatrix(int type, int nrow, int ncol)
{
	SEXP x = minir_alloc(type, (long) nrow * ncol);

	SET_DIM(x, new_dim(2, nrow, ncol, 0));
	return x;
}

SEXP Rf_alloc3DArray(int type, int nrow, int ncol, int nface)
{
	SEXP x = minir_alloc(type, (long) nrow * ncol * nface);

	SET_DIM(x, new_dim(3, nrow, ncol, nface));
	return x;
}

static void free_sexp(SEXP x)
{
	struct minir_slot *s, *s2;

	if (x->type == EXTPTRSXP && x->fin != NULL) {
		void (*fin)(SEXP) = x->fin;

		x->fin = NULL;
		fin(x);
	}

	if (x->prev != NULL)
		x->prev->next = x->next;
	else
		all_head = x->next;
	if (x->next != NULL)
		x->next->prev = x->prev;
	for (s = x->slots; s != NULL; s = s2) {
		s2 = s->next;
		free(s->name);
		free(s);
	}
	free(x->classname);
	free(x->data);
	free(x);
}

static void promote(SEXP v, int gen)
{
	long i;
	struct minir_slot *s;

	if (v == NULL || v == R_NilValue || v == &minir_na_string || v->gen <= gen)
		return;
	v->gen = gen;
	if (v->type == VECSXP || v->type == STRSXP)
		for (i = 0; i < v->length; i++)
			promote(((SEXP *) v->data)[i], gen);
	promote(v->names, gen);
	promote(v->tag, gen);
	promote(v->prot, gen);
	for (s = v->slots; s != NULL; s = s->next)
		promote(s->value, gen);
}

void minir_set_generation(int gen) { cur_gen = gen; }

/* Make 'x' and everything reachable from it persistent. */
void minir_persist(SEXP x) { promote(x, 1); }

void minir_release_scratch(void)
{
	SEXP x, x2;
	struct scratch_blk *b, *b2;

	for (x = all_head; x != NULL; x = x2) {
		x2 = x->next;
		if (x->gen >= 2)
			free_sexp(x);
	}
	for (b = scratch_head; b != NULL; b = b2) {
		b2 = b->next;
		free(b);
	}
	scratch_head = NULL;
	minir_free_AEbufs();
}

void minir_free_object(SEXP x)
{
	/* Demote everything reachable from 'x' to scratch; the next
	   minir_release_scratch() frees it. */
	long i;
	struct minir_slot *s;

	if (x == NULL || x == R_NilValue || x == &minir_na_string || x->gen >= 2)
		return;
	x->gen = 2;
	if (x->type == VECSXP || x->type == STRSXP)
		for (i = 0; i < x->length; i++)
			minir_free_object(((SEXP *) x->data)[i]);
	minir_free_object(x->names);
	minir_free_object(x->tag);
	minir_free_object(x->prot);
	for (s = x->slots; s != NULL; s = s->next)
		minir_free_object(s->value);
}

/* ---- vector / string accessors ---------------------------------------- */

SEXP VECTOR_ELT(SEXP x, long i)
{
	if (x->type != VECSXP || i < 0 || i >= x->length)
		Rf_error("minir: bad VECTOR_ELT(type=%d, i=%ld, length=%ld)",
			 x->type, i, x->length);
	return ((SEXP *) x->data)[i];
}

SEXP SET_VECTOR_ELT(SEXP x, long i, SEXP v)
{
	if (x->type != VECSXP || i < 0 || i >= x->length)
		Rf_error("minir: bad SET_VECTOR_ELT(type=%d, i=%ld, length=%ld)",
			 x->type, i, x->length);
	promote(v, x->gen);
	((SEXP *) x->data)[i] = v;
	return v;
}

SEXP STRING_ELT(SEXP x, long i)
{
	if (x->type != STRSXP || i < 0 || i >= x->length)
		Rf_error("minir: bad STRING_ELT");
	return ((SEXP *) x->data)[i];
}

void SET_STRING_ELT(SEXP x, long i, SEXP v)
{
	if (x->type != STRSXP || i < 0 || i >= x->length)
		Rf_error("minir: bad SET_STRING_ELT");
	promote(v, x->gen);
	((SEXP *) x->data)[i] = v;
}

const char *minir_CHAR(SEXP x)
{
	if (x->type != CHARSXP && x->type != SYMSXP)
		Rf_error("minir: CHAR() on a non-CHARSXP");
	return (const char *) x->data;
}

static SEXP mk_chars(int type, const char *s)
{
	size_t n = strlen(s);
	SEXP x = minir_alloc(type, (long) n);

	memcpy(x->data, s, n + 1);
	return x;
}

SEXP Rf_mkChar(const char *s) { return mk_chars(CHARSXP, s); }
SEXP Rf_install(const char *s)
{
	/* symbols are interned for the life of the process in R */
	int g = cur_gen;
	SEXP x;

	cur_gen = 1;
	x = mk_chars(SYMSXP, s);
	cur_gen = g;
	return x;
}

SEXP Rf_mkString(const char *s)
{
	SEXP x = minir_alloc(STRSXP, 1);

	SET_STRING_ELT(x, 0, Rf_mkChar(s));
	return x;
}

SEXP Rf_ScalarInteger(int v)
{
	SEXP x = minir_alloc(INTSXP, 1);

	INTEGER(x)[0] = v;
	return x;
}

SEXP Rf_ScalarLogical(int v)
{
	SEXP x = minir_alloc(LGLSXP, 1);

	LOGICAL(x)[0] = v;
	return x;
}

SEXP Rf_duplicate(SEXP x)
{
	SEXP y;
	long i;
	struct minir_slot *s;

	if (x == R_NilValue || x == &minir_na_string || x->type == SYMSXP)
		return x;
	y = minir_alloc(x->type, x->length);
	if (x->type == VECSXP || x->type == STRSXP) {
		for (i = 0; i < x->length; i++)
			((SEXP *) y->data)[i] = Rf_duplicate(((SEXP *) x->data)[i]);
	} else if (elt_size(x->type) != 0) {
		memcpy(y->data, x->data, elt_size(x->type) * (size_t) (x->length + 1));
	}
	y->names = Rf_duplicate(x->names);
	if (x->classname != NULL)
		y->classname = strdup(x->classname);
	for (s = x->slots; s != NULL; s = s->next)
		SET_SLOT(y, Rf_mkChar(s->name), Rf_duplicate(s->value));
	y->xptr = x->xptr;
	y->tag = x->tag;
	y->prot = x->prot;
	return y;
}

/* ---- S4 slots, names, classes ------------------------------------------ */

SEXP GET_SLOT(SEXP x, SEXP name)
{
	const char *nm = minir_CHAR(name);
	struct minir_slot *s;

	for (s = x->slots; s != NULL; s = s->next)
		if (strcmp(s->name, nm) == 0)
			return s->value;
	Rf_error("minir: no slot of name \"%s\" for this object of class \"%s\"",
		 nm, x->classname ? x->classname : "?");
}

SEXP SET_SLOT(SEXP x, SEXP name, SEXP value)
{
	const char *nm = minir_CHAR(name);
	struct minir_slot *s;

	promote(value, x->gen);
	for (s = x->slots; s != NULL; s = s->next)
		if (strcmp(s->name, nm) == 0) {
			s->value = value;
			return x;
		}
	s = xmalloc(sizeof(*s));
	s->name = strdup(nm);
	s->value = value;
	s->next = x->slots;
	x->slots = s;
	return x;
}

void SET_NAMES(SEXP x, SEXP names)
{
	promote(names, x->gen);
	x->names = names;
}

SEXP GET_NAMES(SEXP x) { return x->names; }

/* dim / dimnames live in the slot list under names no S4 class uses */
static SEXP get_attr(SEXP x, const char *nm)
{
	struct minir_slot *s;

	for (s = x->slots; s != NULL; s = s->next)
		if (strcmp(s->name, nm) == 0)
			return s->value;
	return R_NilValue;
}

void SET_DIM(SEXP x, SEXP dim) { SET_SLOT(x, Rf_install(".dim"), dim); }
SEXP GET_DIM(SEXP x) { return get_attr(x, ".dim"); }
void SET_DIMNAMES(SEXP x, SEXP dimnames) { SET_SLOT(x, Rf_install(".dimnames"), dimnames); }
SEXP GET_DIMNAMES(SEXP x) { return get_attr(x, ".dimnames"); }

int Rf_asLogical(SEXP x)
{
	if (x->length < 1)
		return NA_LOGICAL;
	if (x->type == LGLSXP || x->type == INTSXP)
		return ((int *) x->data)[0];
	if (x->type == REALSXP)
		return ((double *) x->data)[0] != 0.0;
	return NA_LOGICAL;
}

/* pairlists are only used as dimnames containers here: a plain list serves */
SEXP Rf_list2(SEXP a, SEXP b)
{
	SEXP x = minir_alloc(VECSXP, 2);

	SET_VECTOR_ELT(x, 0, a);
	SET_VECTOR_ELT(x, 1, b);
	return x;
}

SEXP Rf_list3(SEXP a, SEXP b, SEXP c)
{
	SEXP x = minir_alloc(VECSXP, 3);

	SET_VECTOR_ELT(x, 0, a);
	SET_VECTOR_ELT(x, 1, b);
	SET_VECTOR_ELT(x, 2, c);
	return x;
}

SEXP MAKE_CLASS(const char *name) { return Rf_mkChar(name); }

SEXP NEW_OBJECT(SEXP classdef)
{
	SEXP x = minir_alloc(S4SXP, 0);

	x->classname = strdup(minir_CHAR(classdef));
	return x;
}

const char *get_classname(SEXP x)
{
	if (x->classname == NULL)
		Rf_error("minir: get_classname() on an object without a class");
	return x->classname;
}

const char *get_List_elementType(SEXP x)
{
	return minir_CHAR(STRING_ELT(GET_SLOT(x, Rf_install("elementType")), 0));
}

void set_List_elementType(SEXP x, const char *type)
{
	SET_SLOT(x, Rf_install("elementType"), Rf_mkString(type));
}

/* ---- external pointers -------------------------------------------------- */

SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot)
{
	SEXP x = minir_alloc(EXTPTRSXP, 0);

	x->xptr = p;
	x->tag = tag;
	x->prot = prot;
	promote(tag, x->gen);
	promote(prot, x->gen);
	return x;
}

SEXP R_ExternalPtrTag(SEXP x) { return x->tag; }
SEXP R_ExternalPtrProtected(SEXP x) { return x->prot; }
void *R_ExternalPtrAddr(SEXP x) { return x->xptr; }
void R_SetExternalPtrAddr(SEXP x, void *p) { x->xptr = p; }

/* the finalizer runs when the external pointer is freed (minir_release_scratch after
   minir_free_object: the stand-in for R's garbage collector) */
void R_RegisterCFinalizerEx(SEXP x, R_CFinalizer_t fun, Rboolean onexit)
{
	(void) onexit;
	x->fin = fun;
}

SEXP R_lsInternal(SEXP env, Rboolean all)
{
	Rf_error("minir: environments are not supported");
}

/* ---- S4Vectors: auto-extending buffers ---------------------------------- */

/* All AE buffers are released together by minir_free_AEbufs(). */
static void **ae_pool = NULL;
static size_t ae_pool_n = 0, ae_pool_cap = 0;

static void *ae_track(void *p)
{
	if (ae_pool_n == ae_pool_cap) {
		ae_pool_cap = ae_pool_cap ? 2 * ae_pool_cap : 1024;
		ae_pool = realloc(ae_pool, ae_pool_cap * sizeof(void *));
		if (ae_pool == NULL)
			Rf_error("minir: out of memory");
	}
	ae_pool[ae_pool_n++] = p;
	return p;
}

void minir_free_AEbufs(void)
{
	size_t i;

	/* entries are IntAE* (which own elts) or IntAEAE* (tagged via low bit) */
	for (i = 0; i < ae_pool_n; i++) {
		size_t v = (size_t) ae_pool[i];

		if (v & 1) {
			IntAEAE *aeae = (IntAEAE *) (v & ~(size_t) 1);

			free(aeae->elts);
			free(aeae);
		} else {
			IntAE *ae = (IntAE *) v;

			free(ae->elts);
			free(ae);
		}
	}
	ae_pool_n = 0;
}

static void IntAE_extend(IntAE *ae, size_t new_buflength)
{
	ae->elts = realloc(ae->elts, (new_buflength ? new_buflength : 1) * sizeof(int));
	if (ae->elts == NULL)
		Rf_error("minir: out of memory");
	ae->_buflength = new_buflength;
}

IntAE *new_IntAE(size_t buflength, size_t nelt, int val)
{
	IntAE *ae = xmalloc(sizeof(*ae));
	size_t i;

	ae->_buflength = 0;
	ae->_nelt = 0;
	ae->elts = NULL;
	ae_track(ae);
	if (buflength != 0)
		IntAE_extend(ae, buflength);
	if (nelt > buflength)
		Rf_error("minir: new_IntAE(): nelt > buflength");
	ae->_nelt = nelt;
	for (i = 0; i < nelt; i++)
		ae->elts[i] = val;
	return ae;
}

size_t IntAE_get_nelt(const IntAE *ae) { return ae->_nelt; }

size_t IntAE_set_nelt(IntAE *ae, size_t nelt)
{
	if (nelt > ae->_buflength)
		Rf_error("minir: IntAE_set_nelt(): nelt > buflength");
	return ae->_nelt = nelt;
}

void IntAE_set_val(const IntAE *ae, int val)
{
	size_t i;

	for (i = 0; i < ae->_nelt; i++)
		ae->elts[i] = val;
}

void IntAE_insert_at(IntAE *ae, size_t at, int val)
{
	size_t nelt = ae->_nelt, i;

	if (at > nelt)
		Rf_error("minir: IntAE_insert_at(): at > nelt");
	if (nelt >= ae->_buflength)
		IntAE_extend(ae, ae->_buflength < 8 ? 8 : 2 * ae->_buflength);
	for (i = nelt; i > at; i--)
		ae->elts[i] = ae->elts[i - 1];
	ae->elts[at] = val;
	ae->_nelt = nelt + 1;
}

void IntAE_append(IntAE *ae, const int *newvals, size_t nnewval)
{
	size_t new_nelt = ae->_nelt + nnewval, cap;

	if (new_nelt > ae->_buflength) {
		cap = ae->_buflength < 8 ? 8 : ae->_buflength;
		while (cap < new_nelt)
			cap *= 2;
		IntAE_extend(ae, cap);
	}
	memcpy(ae->elts + ae->_nelt, newvals, nnewval * sizeof(int));
	ae->_nelt = new_nelt;
}

void IntAE_shift(const IntAE *ae, size_t offset, int shift)
{
	size_t i;

	for (i = offset; i < ae->_nelt; i++)
		ae->elts[i] += shift;
}

static int cmp_int_asc(const void *a, const void *b)
{
	int x = *(const int *) a, y = *(const int *) b;

	return (x > y) - (x < y);
}

static int cmp_int_desc(const void *a, const void *b) { return cmp_int_asc(b, a); }

void sort_int_array(int *x, size_t nelt, int desc)
{
	qsort(x, nelt, sizeof(int), desc ? cmp_int_desc : cmp_int_asc);
}

void get_order_of_int_array(const int *x, int nelt, int desc, int *out, int out_shift)
{
	Rf_error("minir: get_order_of_int_array() is not needed on the PDict path");
}

void IntAE_qsort(const IntAE *ae, size_t offset, int desc)
{
	sort_int_array(ae->elts + offset, ae->_nelt - offset, desc);
}

void IntAE_uniq(IntAE *ae, size_t offset)
{
	size_t nelt = ae->_nelt, i1, i2;

	if (nelt <= offset + 1)
		return;
	for (i1 = offset, i2 = offset + 1; i2 < nelt; i2++)
		if (ae->elts[i2] != ae->elts[i1])
			ae->elts[++i1] = ae->elts[i2];
	ae->_nelt = i1 + 1;
}

LLongAE *new_LLongAE(size_t buflength, size_t nelt, long long val)
{
	LLongAE *ae = ae_track(xmalloc(sizeof(*ae)));
	size_t i;

	ae->_buflength = buflength > nelt ? buflength : nelt;
	ae->_nelt = nelt;
	ae->elts = ae->_buflength ? xmalloc(ae->_buflength * sizeof(long long)) : NULL;
	for (i = 0; i < nelt; i++)
		ae->elts[i] = val;
	return ae;
}

size_t LLongAE_get_nelt(const LLongAE *ae) { return ae->_nelt; }

void LLongAE_insert_at(LLongAE *ae, size_t at, long long val)
{
	if (ae->_nelt >= ae->_buflength) {
		ae->_buflength = ae->_buflength ? 2 * ae->_buflength : 16;
		ae->elts = realloc(ae->elts, ae->_buflength * sizeof(long long));
	}
	memmove(ae->elts + at + 1, ae->elts + at, (ae->_nelt - at) * sizeof(long long));
	ae->elts[at] = val;
	ae->_nelt++;
}

CharAEAE *new_CharAEAE(size_t buflength, size_t nelt)
{
	CharAEAE *aeae = ae_track(xmalloc(sizeof(*aeae)));

	(void) nelt;
	aeae->_buflength = buflength;
	aeae->_nelt = 0;
	aeae->elts = buflength ? xmalloc(buflength * sizeof(char *)) : NULL;
	return aeae;
}

size_t CharAEAE_get_nelt(const CharAEAE *aeae) { return aeae->_nelt; }

void CharAEAE_append_string(CharAEAE *aeae, const char *string)
{
	if (aeae->_nelt >= aeae->_buflength) {
		aeae->_buflength = aeae->_buflength ? 2 * aeae->_buflength : 16;
		aeae->elts = realloc(aeae->elts, aeae->_buflength * sizeof(char *));
	}
	aeae->elts[aeae->_nelt++] = strdup(string);
}

SEXP new_CHARACTER_from_CharAEAE(const CharAEAE *aeae)
{
	SEXP x = NEW_CHARACTER((long) aeae->_nelt);
	size_t i;

	for (i = 0; i < aeae->_nelt; i++)
		SET_STRING_ELT(x, (long) i, Rf_mkChar(aeae->elts[i]));
	return x;
}

/* IN-PLACE coercion in S4Vectors; here the list just remembers its row count */
void list_as_data_frame(SEXP x, int nrow)
{
	SET_SLOT(x, Rf_install(".nrow"), Rf_ScalarInteger(nrow));
}

SEXP new_INTEGER_from_IntAE(const IntAE *ae)
{
	SEXP ans = NEW_INTEGER((long) ae->_nelt);

	memcpy(INTEGER(ans), ae->elts, ae->_nelt * sizeof(int));
	return ans;
}

IntAE *new_IntAE_from_CHARACTER(SEXP x, int keyshift)
{
	Rf_error("minir: new_IntAE_from_CHARACTER() is not needed on the PDict path");
}

IntAEAE *new_IntAEAE(size_t buflength, size_t nelt)
{
	IntAEAE *aeae = xmalloc(sizeof(*aeae));
	size_t i;

	if (nelt > buflength)
		Rf_error("minir: new_IntAEAE(): nelt > buflength");
	aeae->_buflength = buflength;
	aeae->_nelt = nelt;
	aeae->elts = xmalloc((buflength ? buflength : 1) * sizeof(IntAE *));
	ae_track((void *) ((size_t) aeae | 1));
	for (i = 0; i < nelt; i++)
		aeae->elts[i] = new_IntAE(0, 0, 0);
	return aeae;
}

size_t IntAEAE_get_nelt(const IntAEAE *aeae) { return aeae->_nelt; }

void IntAEAE_sum_and_shift(const IntAEAE *aeae1, const IntAEAE *aeae2, int shift)
{
	size_t i, j;

	if (aeae1->_nelt != aeae2->_nelt)
		Rf_error("minir: IntAEAE_sum_and_shift(): incompatible lengths");
	for (i = 0; i < aeae1->_nelt; i++) {
		const IntAE *a = aeae1->elts[i], *b = aeae2->elts[i];

		if (a->_nelt != b->_nelt)
			Rf_error("minir: IntAEAE_sum_and_shift(): incompatible lengths");
		for (j = 0; j < a->_nelt; j++)
			a->elts[j] += b->elts[j] + shift;
	}
}

/* mode 0: integer(0) for empty elements; mode 1: NULL (the PDict path uses 1) */
SEXP new_LIST_from_IntAEAE(const IntAEAE *aeae, int mode)
{
	SEXP ans = NEW_LIST((long) aeae->_nelt);
	size_t i;

	for (i = 0; i < aeae->_nelt; i++) {
		const IntAE *ae = aeae->elts[i];

		if (ae->_nelt != 0 || mode == 0)
			SET_VECTOR_ELT(ans, (long) i, new_INTEGER_from_IntAE(ae));
		else if (mode != 1)
			SET_VECTOR_ELT(ans, (long) i, NEW_LOGICAL(1));
	}
	return ans;
}

SEXP IntAEAE_toEnvir(const IntAEAE *aeae, SEXP envir, int keyshift)
{
	Rf_error("minir: environments are not supported");
}

/* ---- IRanges ------------------------------------------------------------- */

SEXP get_H2LGrouping_high2low(SEXP x) { return GET_SLOT(x, Rf_install("high2low")); }
SEXP get_H2LGrouping_low2high(SEXP x) { return GET_SLOT(x, Rf_install("low2high")); }

SEXP new_IRanges(const char *classname, SEXP start, SEXP width, SEXP names)
{
	SEXP ans = NEW_OBJECT(MAKE_CLASS(classname));

	SET_SLOT(ans, Rf_install("start"), start);
	SET_SLOT(ans, Rf_install("width"), width);
	SET_SLOT(ans, Rf_install("NAMES"), names);
	return ans;
}

/* ---- XVector --------------------------------------------------------------- */

/* XString: S4 object whose xptr points at the bytes and length is the width. */
SEXP minir_new_XString(const char *classname, const char *ptr, int length)
{
	SEXP x = NEW_OBJECT(MAKE_CLASS(classname));

	x->xptr = (void *) ptr;
	x->length = length;
	return x;
}

Chars_holder hold_XRaw(SEXP x)
{
	Chars_holder h;

	h.ptr = (const char *) x->xptr;
	h.length = (int) x->length;
	return h;
}

/* XStringSet: S4 object; xptr -> array of element pointers; slot "width". */
SEXP minir_new_XStringSet(const char *classname, int n, const char *const *ptrs,
			  const int *widths)
{
	SEXP x = NEW_OBJECT(MAKE_CLASS(classname)), w = NEW_INTEGER(n), p;

	/* keep the pointer table inside a RAW vector so it is owned by 'x' */
	p = NEW_RAW((long) (n ? n : 1) * (long) sizeof(char *));
	if (n)
		memcpy(RAW(p), ptrs, (size_t) n * sizeof(char *));
	if (n)
		memcpy(INTEGER(w), widths, (size_t) n * sizeof(int));
	SET_SLOT(x, Rf_install("width"), w);
	SET_SLOT(x, Rf_install(".ptrs"), p);
	x->xptr = RAW(p);
	x->length = n;
	return x;
}

XVectorList_holder hold_XVectorList(SEXP x)
{
	XVectorList_holder h;

	h.classname = get_classname(x);
	h.length = (int) x->length;
	h.ptrs = (const char *const *) x->xptr;
	h.widths = INTEGER(GET_SLOT(x, Rf_install("width")));
	return h;
}

int get_length_from_XVectorList_holder(const XVectorList_holder *x_holder)
{
	return x_holder->length;
}

Chars_holder get_elt_from_XRawList_holder(const XVectorList_holder *x_holder, int i)
{
	Chars_holder h;

	h.ptr = x_holder->ptrs[i];
	h.length = x_holder->widths[i];
	return h;
}

XVectorList_holder get_linear_subset_from_XVectorList_holder(
		const XVectorList_holder *x_holder, int offset, int length)
{
	XVectorList_holder h = *x_holder;

	h.length = length;
	h.ptrs += offset;
	h.widths += offset;
	return h;
}

int get_XVectorList_length(SEXP x) { return (int) x->length; }
SEXP get_XVectorList_width(SEXP x) { return GET_SLOT(x, Rf_install("width")); }
void set_XVectorList_names(SEXP x, SEXP names) { SET_SLOT(x, Rf_install("NAMES"), names); }

SEXP get_XVector_tag(SEXP x) { return GET_SLOT(x, Rf_install("tag")); }

SEXP new_XInteger_from_tag(const char *classname, SEXP tag)
{
	SEXP x = NEW_OBJECT(MAKE_CLASS(classname));

	SET_SLOT(x, Rf_install("tag"), tag);
	x->length = tag->length;
	{
		/* XVector: x@shared is a SharedVector whose @xp external pointer carries the data
		   vector as its tag (XVector/src/SharedVector_class.c) */
		SEXP shared = NEW_OBJECT(MAKE_CLASS("SharedInteger"));

		SET_SLOT(shared, Rf_install("xp"), R_MakeExternalPtr(NULL, tag, R_NilValue));
		SET_SLOT(x, Rf_install("shared"), shared);
	}
	return x;
}

SEXP new_XRaw_from_tag(const char *classname, SEXP tag)
{
	SEXP x = new_XInteger_from_tag(classname, tag);

	x->xptr = RAW(tag);
	return x;
}

SEXP alloc_XRaw(const char *classname, int length)
{
	return new_XRaw_from_tag(classname, NEW_RAW(length));
}

/* One RAW pool holding the elements back to back (what XVector does for small sets). */
SEXP alloc_XRawList(const char *classname, const char *element_type, SEXP width)
{
	int n = LENGTH(width), i;
	size_t total = 0, off = 0;
	const char **ptrs = xmalloc((size_t) (n ? n : 1) * sizeof(char *));
	SEXP pool, x;

	(void) element_type;
	for (i = 0; i < n; i++)
		total += (size_t) INTEGER(width)[i];
	pool = NEW_RAW((long) (total ? total : 1));
	for (i = 0; i < n; i++) {
		ptrs[i] = (const char *) RAW(pool) + off;
		off += (size_t) INTEGER(width)[i];
	}
	x = minir_new_XStringSet(classname, n, ptrs, INTEGER(width));
	SET_SLOT(x, Rf_install(".pool"), pool);
	if (GET_NAMES(width) != R_NilValue)	/* XVector propagates names(width) */
		set_XVectorList_names(x, GET_NAMES(width));
	free(ptrs);
	return x;
}

SEXP get_XVectorList_names(SEXP x)
{
	struct minir_slot *s;

	for (s = x->slots; s != NULL; s = s->next)
		if (strcmp(s->name, "NAMES") == 0)
			return s->value;
	return R_NilValue;
}

void Ocopy_bytes_from_i1i2_with_lkup(int i1, int i2, char *dest, int dest_nbytes,
		const char *src, int src_nbytes, const int *lkup, int lkup_length)
{
	Rf_error("minir: Ocopy_bytes_from_i1i2_with_lkup() is only used when writing FASTA");
}

/* ---- XVector io_utils: a "file" is an in-memory buffer behind an external pointer ---- */

SEXP minir_new_filexp(const char *data, long nbytes)
{
	SEXP holder = NEW_RAW((long) sizeof(struct minir_memfile));	/* owned by the pointer */
	struct minir_memfile *f = (struct minir_memfile *) RAW(holder);

	f->data = data;
	f->nbytes = nbytes;
	f->pos = 0;
	return R_MakeExternalPtr(f, R_NilValue, holder);
}

/* fgets() semantics plus XVector's end-of-line detection: a sentinel in the last byte of the
 * buffer tells whether the buffer was filled; not filled, or filled up to a newline, means the
 * end of the line (or of the file) is in the buffer.  Returns 0 at end of file. */
int filexp_gets(SEXP filexp, char *buf, int buf_size, int *EOL_in_buf)
{
	struct minir_memfile *f = R_ExternalPtrAddr(filexp);
	int n = 0;

	if (f->pos >= f->nbytes)
		return 0;
	buf[buf_size - 1] = 'N';
	while (n < buf_size - 1 && f->pos < f->nbytes) {
		char c = f->data[f->pos++];

		buf[n++] = c;
		if (c == '\n')
			break;
	}
	buf[n] = 0;
	*EOL_in_buf = buf[buf_size - 1] == 'N' || buf[buf_size - 2] == '\n';
	return *EOL_in_buf ? 2 : 1;
}

void filexp_seek(SEXP filexp, long long int offset, int whence)
{
	struct minir_memfile *f = R_ExternalPtrAddr(filexp);

	(void) whence;			/* only SEEK_SET is used */
	f->pos = (long) offset;
}

long long int filexp_tell(SEXP filexp)
{
	return ((struct minir_memfile *) R_ExternalPtrAddr(filexp))->pos;
}

int filexp_puts(SEXP filexp, const char *s)
{
	Rf_error("minir: filexp_puts() is only used when writing FASTA");
}

int delete_trailing_LF_or_CRLF(const char *buf, int buf_len)
{
	if (buf_len == -1)
		buf_len = (int) strlen(buf);
	if (buf_len == 0)
		return 0;
	if (buf[--buf_len] != '\n')
		return ++buf_len;
	if (buf_len == 0)
		return 0;
	if (buf[--buf_len] != '\r')
		return ++buf_len;
	return buf_len;
}

void Ocopy_bytes_to_i1i2_with_lkup(int i1, int i2, char *dest, int dest_nbytes,
		const char *src, int src_nbytes, const int *lkup, int lkup_length)
{
	Rf_error("minir: Ocopy_bytes_to_i1i2_with_lkup() is not needed on the PDict path");
}
